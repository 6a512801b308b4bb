#!/usr/bin/env python3
"""Dump the reference's builtin pore-model TABLES (data, not source) into
nanocall_b200/data/builtin_models.{bin,json}.

Source of the numbers: Builtin_Model::init_lists / names / strands
(/root/reference/src/nanocall/Builtin_Model.cpp:1-19, src/builtin_models/*.inl), read through
oracle/_ref/libncref.so so the float32 bit patterns are exactly those compiled into the
reference binary.  Layout of the .bin: n_models x 4096 x 4 float32 little-endian
(level_mean, level_stdv, sd_mean, sd_stdv), state index = 2-bit packed k-mer (A=0,C=1,G=2,T=3,
first base in the high bits; Kmer.hpp:13-50).  Runs only where the reference tree exists."""
import ctypes, json, os, sys
import numpy as np

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(root, "oracle", "_ref", "libncref.so"))
n = lib.ncref_n_builtin()
tables = np.zeros((n, 4096, 4), dtype=np.float32)
meta = []
for i in range(n):
    strand = ctypes.c_int()
    name = ctypes.create_string_buffer(128)
    rc = lib.ncref_builtin(i, tables[i].ctypes.data_as(ctypes.c_void_p), ctypes.byref(strand), name, 128)
    assert rc == 0, rc
    meta.append({"name": name.value.decode(), "strand": strand.value})
out = os.path.join(root, "nanocall_b200", "data")
os.makedirs(out, exist_ok=True)
tables.astype("<f4").tofile(os.path.join(out, "builtin_models.bin"))
with open(os.path.join(out, "builtin_models.json"), "w") as f:
    json.dump({"n_states": 4096, "fields": ["level_mean", "level_stdv", "sd_mean", "sd_stdv"], "models": meta}, f, indent=1)
print(meta)
