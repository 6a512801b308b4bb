#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libncref.so = the reference's
own headers compiled in place).  Runs only in the build container (needs /root/reference); the
fixtures are committed so the GPU box and the CPU test tier never need the reference tree.

The reference ships no golden vectors of its own; these are outputs of its
code on seeded synthetic inputs, with the inputs stored next to them.
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib  # noqa: E402
from nanocall_b200 import models, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
R = oracle_lib.ref()
M = {m["name"]: m for m in models.load_builtin_models()}
T, C1, C2 = (M[k]["table"] for k in ("r73.t.006.ont.model", "r73.c.p1.006.ont.model", "r73.c.p2.006.ont.model"))


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def mask_of(i, j):
    m = 1 if i == j else 0
    for l in range(1, 6):
        if (i & ((1 << (2 * (6 - l))) - 1)) == (j >> (2 * l)):
            m |= 1 << l
    return m


# ---- logsum table, transitions, scaled models, mean/stdv
tbl = R.flogsum_table()
g = dict(flogsum_table_crc=crc(tbl), flogsum_table_head=tbl[:64].copy(), flogsum_table_tail=tbl[-64:].copy())
pairs = np.array([[-1.0, -2.5], [-3.25, -3.25], [-100.0, -84.5], [-np.inf, -7.0], [-5.0, -np.inf], [0.0, -15.9989], [0.0, -16.0],
                  [-1234.5, -1234.5005], [-2.0, -9.4321]], np.float32)
g["flogsum_pairs"] = pairs
g["flogsum_vals"] = np.array([R.flogsum(a, b) for a, b in pairs], np.float32)
st_cases = np.array([[0.1, 0.3], [0.09, 0.28], [0.05, 0.4], [0.4, 0.05], [0.1476, 0.2379]], np.float32)
g["st_cases"] = st_cases
for k, (ps, pk) in enumerate(st_cases):
    tr = R.transitions(float(ps), float(pk))
    g[f"tr{k}_from_cnt"] = tr["from_cnt"].astype(np.uint8)
    for name in ("from_idx", "from_lp", "to_idx", "to_lp", "to_cnt"):
        g[f"tr{k}_{name}_crc"] = crc(tr[name])
    lut = np.full(64, np.nan, np.float32)
    for j in range(4096):
        for q in range(tr["from_cnt"][j]):
            i = int(tr["from_idx"][j, q])
            m = mask_of(i, j)
            v = tr["from_lp"][j, q]
            assert np.isnan(lut[m]) or lut[m].view(np.uint32) == v.view(np.uint32), "weight is not a function of the mask"
            lut[m] = v
    g[f"tr{k}_lut"] = lut
pm_cases = np.array([[1, 0, 0, 1, 1, 1], [1.05, -2.0, 0.003, 1.1, 0.9, 1.2], [0.93, 4.5, -0.004, 1.27, 1.18, 0.83]], np.float32)
g["pm_cases"] = pm_cases
for k, pm in enumerate(pm_cases):
    for name, tab in (("t", T), ("c1", C1)):
        sm = R.scaled_model(tab, pm)
        for f in ("level_mean", "level_stdv", "log_level_stdv", "sd_mean", "sd_lambda", "log_sd_lambda"):
            g[f"sm{k}_{name}_{f}_crc"] = crc(sm[f])
            g[f"sm{k}_{name}_{f}_head"] = sm[f][:8].copy()
        g[f"sm{k}_{name}_stats"] = np.array([sm["mean"], sm["stdv"]], np.float32)
rng = np.random.default_rng(99)
x = rng.normal(58, 6, 777).astype(np.float32)
g["ms_x"] = x
g["ms_out"] = np.array(R.mean_stdv(x), np.float32)
g["st_train_kmers_crc"] = crc(R.st_train_kmers())
g["st_train_kmers_n"] = np.uint32(R.st_train_kmers().size)
np.savez_compressed(os.path.join(OUT, "tables.npz"), **g)

# ---- Viterbi: several reads incl. scaled / custom transitions / ties / zero stdv
v = {}
cases = []
rng = np.random.default_rng(7)
for k, (tab, name, n, pm, st) in enumerate([
        (T, "t", 200, pm_cases[0], (0.1, 0.3)),
        (T, "t", 1, pm_cases[0], (0.1, 0.3)),
        (T, "t", 2, pm_cases[1], (0.1, 0.3)),
        (C1, "c1", 333, pm_cases[1], (0.1476, 0.2379)),
        (C2, "c2", 150, pm_cases[2], (0.05, 0.4)),
        (T, "t", 1500, pm_cases[0], (0.1, 0.3))]):
    rd = synth.make_read(rng, tab, n, tuple(pm))
    if k == 4:
        rd["stdv"][[3, 77]] = 0.0
    r = R.viterbi(tab, pm, st[0], st[1], rd["mean"], rd["stdv"], rd["start"], dump=(n <= 2))
    v[f"c{k}_model"] = np.array(name)
    v[f"c{k}_pm"] = pm
    v[f"c{k}_st"] = np.array(st, np.float32)
    for f in ("mean", "stdv", "start"):
        v[f"c{k}_{f}"] = rd[f]
    v[f"c{k}_path_prob"] = np.float32(r["path_prob"])
    v[f"c{k}_states"] = r["states"].astype(np.uint16)
    v[f"c{k}_moves"] = r["moves"].astype(np.uint8)
    v[f"c{k}_bases"] = np.array(r["bases"])
    if n <= 2:
        v[f"c{k}_alpha"] = r["alpha"]
# plateau case: constant events
n = 120
mean = np.full(n, 58.0, np.float32); stdv = np.full(n, 0.9, np.float32); start = (np.arange(n) * 0.02).astype(np.float32)
r = R.viterbi(T, pm_cases[0], 0.1, 0.3, mean, stdv, start)
k = 6
v[f"c{k}_model"] = np.array("t"); v[f"c{k}_pm"] = pm_cases[0]; v[f"c{k}_st"] = np.array((0.1, 0.3), np.float32)
v[f"c{k}_mean"], v[f"c{k}_stdv"], v[f"c{k}_start"] = mean, stdv, start
v[f"c{k}_path_prob"] = np.float32(r["path_prob"]); v[f"c{k}_states"] = r["states"].astype(np.uint16)
v[f"c{k}_moves"] = r["moves"].astype(np.uint8); v[f"c{k}_bases"] = np.array(r["bases"])
v["n_cases"] = np.uint32(7)
np.savez_compressed(os.path.join(OUT, "viterbi.npz"), **v)

# ---- Forward/Backward: full alpha/beta for a short sequence, logZ for longer ones
f = {}
rd = synth.make_read(rng, T, 100, tuple(pm_cases[1]))
r = R.fwbw(T, pm_cases[1], 0.1, 0.3, rd["mean"][:6], rd["stdv"][:6], rd["start"][:6])
f["a_pm"], f["a_st"] = pm_cases[1], np.array((0.1, 0.3), np.float32)
for q in ("mean", "stdv", "start"):
    f[f"a_{q}"] = rd[q][:6]
f["a_alpha"], f["a_beta"], f["a_logz"] = r["alpha"], r["beta"], np.float32(r["log_pr_data"])
r = R.fwbw(T, pm_cases[1], 0.1476, 0.2379, rd["mean"], rd["stdv"], rd["start"])
f["b_pm"], f["b_st"] = pm_cases[1], np.array((0.1476, 0.2379), np.float32)
for q in ("mean", "stdv", "start"):
    f[f"b_{q}"] = rd[q]
f["b_logz"] = np.float32(r["log_pr_data"])
f["b_alpha_crc"], f["b_beta_crc"] = crc(r["alpha"]), crc(r["beta"])
f["b_alpha_last"], f["b_beta_first"] = r["alpha"][-1].copy(), r["beta"][0].copy()
np.savez_compressed(os.path.join(OUT, "fwbw.npz"), **f)

# ---- train_one_round: double-strand (3 chained rounds) and single-strand
t = {}
true = np.array([1.05, -2.0, 0.003, 1.1, 0.9, 1.2], np.float32)
r0 = synth.make_read(rng, T, 400, tuple(true))
r1 = synth.make_read(rng, C1, 400, tuple(true))


def seqs(r, sd, n=50):
    return [(sd, r["mean"][:n], r["stdv"][:n], r["start"][:n]), (sd, r["mean"][-n:], r["stdv"][-n:], r["start"][-n:])]


S = seqs(r0, 0) + seqs(r1, 1)
t["n_seqs"] = np.uint32(4)
for k, (sd, m, s, st_) in enumerate(S):
    t[f"s{k}_strand"] = np.uint32(sd); t[f"s{k}_mean"] = m; t[f"s{k}_stdv"] = s; t[f"s{k}_start"] = st_
pm = np.array([1.0, 0.5, 0, 1, 1, 1], np.float32)
st = np.array([.1, .3, .1, .3], np.float32)
for rnd in range(3):
    o = R.train_one_round(S, T, C1, pm, st)
    t[f"d{rnd}_in_pm"], t[f"d{rnd}_in_st"] = pm, st
    t[f"d{rnd}_pm"], t[f"d{rnd}_st"], t[f"d{rnd}_fit"], t[f"d{rnd}_done"] = o["pm"], o["st"], o["fit"], np.uint8(o["done"])
    pm, st = o["pm"], o["st"]
o = R.train_one_round(seqs(r1, 1), C1, C1, pm, st)
t["s_in_pm"], t["s_in_st"] = pm, st
t["s_pm"], t["s_st"], t["s_fit"], t["s_done"] = o["pm"], o["st"], o["fit"], np.uint8(o["done"])
o = R.train_one_round(S, T, C1, pm, st, train_scaling=False)
t["ns_pm"], t["ns_st"], t["ns_fit"] = o["pm"], o["st"], o["fit"]
o = R.train_one_round(S, T, C1, pm, st, train_transitions=False)
t["nt_pm"], t["nt_st"], t["nt_fit"] = o["pm"], o["st"], o["fit"]
np.savez_compressed(os.path.join(OUT, "train.npz"), **t)
print({fn: os.path.getsize(os.path.join(OUT, fn)) for fn in os.listdir(OUT)})
