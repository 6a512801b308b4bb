"""Deterministic synthetic 2D reads for the pipeline parity runs (test infrastructure; shared by
tests/test_pipeline_gpu.py and tools/make_pipeline_golden.py so both sides decode the same events)."""
import numpy as np

from nanocall_b200 import synth

R73 = ["r73.t.006.ont.model", "r73.c.p1.006.ont.model", "r73.c.p2.006.ont.model"]


def make_reads(models, seed, n_reads, nt=600, nc=500, comp="r73.c.p1.006.ont.model"):
    rng = np.random.default_rng(seed)
    reads = []
    for k in range(n_reads):
        pm = tuple(synth.random_params(rng, 1)[0])
        t = synth.make_read(rng, models[R73[0]]["table"], nt + 17 * (k % 40), pm)
        c = synth.make_read(rng, models[comp if k % 2 == 0 else R73[2]]["table"], nc + 11 * (k % 40), pm)
        # complement starts where the template ended (start is relative to the template start when scaled together)
        c["start"] = (c["start"] + t["start"][-1] + np.float32(0.5)).astype(np.float32)
        reads.append((f"read{k}", [t, c]))
    return reads
