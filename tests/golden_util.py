import os
import zlib

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODEL_KEYS = {"t": "r73.t.006.ont.model", "c1": "r73.c.p1.006.ont.model", "c2": "r73.c.p2.006.ont.model"}


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def bits(x):
    return np.asarray(x, np.float32).view(np.uint32)


def same_bits(a, b):
    return np.array_equal(bits(a), bits(b))


def train_seqs(t):
    return [(int(t[f"s{k}_strand"]), t[f"s{k}_mean"], t[f"s{k}_stdv"], t[f"s{k}_start"]) for k in range(int(t["n_seqs"]))]
