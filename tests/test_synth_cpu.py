"""CPU tier: the synthetic workload generators the benchmarks and GPU tests rely on, and the committed pipeline golden."""
import gzip
import json
import os

import numpy as np

from nanocall_b200 import models, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mixture_lengths_match_the_named_shape():
    L = synth.mixture_lengths(5, 20000)
    assert L.min() >= 500 and L.max() < 150000
    frac_long = np.mean(L >= 100000)
    frac_mid = np.mean((L >= 20000) & (L < 100000))
    assert 0.005 < frac_long < 0.015 and 0.07 < frac_mid < 0.11
    assert 4000 < np.median(L[L < 20000]) < 6000
    assert np.array_equal(L, synth.mixture_lengths(5, 20000))     # deterministic


def test_make_batch_with_lengths_restarts_the_clock_per_read():
    table = models.builtin_model("r73.t")["table"]
    lengths = [7, 1, 300, 33]
    b = synth.make_batch_uniform(9, table, 0, 0, lengths=lengths)
    off = b["ev_off"].astype(np.int64)
    assert list(np.diff(off)) == lengths and b["mean"].size == sum(lengths)
    for a, e in zip(off[:-1], off[1:]):
        st = b["start"][a:e]
        assert st[0] == 0.0 and np.all(np.diff(st) > 0)
    assert np.all(b["stdv"] > 0) and np.all(b["stdv"] <= 4.0)
    assert b["truth"].max() < 4096


def test_pipeline_golden_is_complete():
    path = os.path.join(ROOT, "tests", "golden", "pipeline_r73.json.gz")
    with gzip.open(path, "rt") as f:
        g = json.load(f)
    assert g["n_reads"] == len(g["reads"]) == 200
    for r in g["reads"]:
        assert len(r["bases"]) == 2 and all(set(b) <= set("ACGT") and len(b) > 500 for b in r["bases"])
        assert len(r["pm"]) == 2 and all(len(v) == 6 for v in r["pm"].values())
        assert set(r["rounds"]) == set(r["pm"]) and all(1 <= k <= 20 for k in r["rounds"].values())
