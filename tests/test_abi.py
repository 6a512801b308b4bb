"""CPU tier: the C-ABI library loads, exports every symbol include/nanocall_b200.h declares, fails
loudly without a GPU, and its host-side helpers reproduce the reference arithmetic (golden vectors)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from golden_util import MODEL_KEYS, load, same_bits
from nanocall_b200 import _lib, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nanocall_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)
    assert b"sm_100a" in lib.nc_version()


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.NanocallError) as ei:
        api.Context(0)
    assert ei.value.code == _lib.NC_ERR_CUDA and "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "nanocall_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle_lib" not in src and "nc_oracle" not in src and "libncref" not in src, fn


def test_transition_lut_matches_reference_weights():
    g = load("tables")
    for k, (ps, pk) in enumerate(g["st_cases"]):
        lut = api.transition_lut(ps, pk)
        want = g[f"tr{k}_lut"]
        seen = ~np.isnan(want)
        assert seen.sum() == 18  # exactly 18 masks occur
        assert same_bits(lut[seen], want[seen]), k


def test_mean_stdv_base_seq_min_skip():
    g = load("tables")
    assert same_bits(api.mean_stdv(g["ms_x"]), g["ms_out"])
    v = load("viterbi")
    lib = _lib.load()
    for k in range(int(v["n_cases"])):
        st, mv = v[f"c{k}_states"], v[f"c{k}_moves"]
        assert api.base_seq(st, mv) == str(v[f"c{k}_bases"])
        for i in range(1, min(st.size, 400)):
            assert lib.nc_min_skip(int(st[i - 1]), int(st[i])) == int(mv[i])
    assert api.base_seq(np.array([0x1B], np.uint16), np.array([0], np.uint8)) == "AAACGT"


def test_dispatch_order_planner():
    """nc_plan_dispatch_order (host helper behind nc_viterbi_packed): a permutation; identity when the first wave fits
    the pool; otherwise the memory of the jobs running at any time of the simulated schedule stays within the pool,
    and the longest job still starts first."""
    import ctypes as C
    from nanocall_b200 import _lib, synth
    lib = _lib.load()
    lens = np.sort(synth.mixture_lengths(2026, 4000).astype(np.uint32))[::-1].copy()
    perm = np.zeros(lens.size, np.uint32)
    pool, workers = 8_000_000, 145
    changed = lib.nc_plan_dispatch_order(lens.size, lens.ctypes.data, C.c_uint64(pool), workers, perm.ctypes.data)
    assert changed == 1
    assert np.array_equal(np.sort(perm), np.arange(lens.size, dtype=np.uint32))
    assert perm[0] == 0                                   # the longest read starts first
    # replay an order with the planner's timing and memory model (a job starts when a forward CTA is idle and its
    # columns are free): the planned order finishes no later than plain longest-first
    import heapq

    def replay(order):
        running, free, idle, now, end = [], int(pool * 0.9), workers, 0.0, 0.0
        for k in order:
            n = int(lens[k])
            while idle == 0 or n > free:
                t, m = heapq.heappop(running)
                now, free, idle = t, free + m, idle + 1
            free, idle = free - n, idle - 1
            heapq.heappush(running, (now + 1.1 * n, n))
            end = max(end, now + 1.1 * n)
        return end

    planned, lpt = replay(perm), replay(np.arange(lens.size))
    assert planned <= lpt, (planned, lpt)
    # a batch whose first wave fits keeps longest-first
    small = np.full(1000, 10000, np.uint32)
    p2 = np.zeros(small.size, np.uint32)
    assert lib.nc_plan_dispatch_order(small.size, small.ctypes.data, C.c_uint64(pool), workers, p2.ctypes.data) == 0
    assert np.array_equal(p2, np.arange(small.size, dtype=np.uint32))


def test_dispatch_order_planner_accepts_unsorted_lengths():
    """The first-wave check takes the n_workers LONGEST jobs wherever they sit in lens[]: short jobs in front of
    long ones must not make a batch look as if its first wave fitted the pool."""
    import ctypes as C
    from nanocall_b200 import _lib
    lib = _lib.load()
    workers, pool = 4, 1000
    lens = np.array([10] * 8 + [400] * 8, np.uint32)          # ascending: the four longest need 1600 > 900 columns
    perm = np.zeros(lens.size, np.uint32)
    assert lib.nc_plan_dispatch_order(lens.size, lens.ctypes.data, C.c_uint64(pool), workers, perm.ctypes.data) == 1
    assert np.array_equal(np.sort(perm), np.arange(lens.size, dtype=np.uint32))
    assert lens[perm[0]] == 400                                # the longest job that fits starts first
