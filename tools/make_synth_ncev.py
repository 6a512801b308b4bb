#!/usr/bin/env python3
"""Write a synthetic 2D read set (template + complement, random per-read scaling) as one .ncev container.
usage: make_synth_ncev.py out.ncev n_reads n_template_events n_complement_events [seed]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanocall_b200 import evio, models, synth  # noqa: E402

out, n_reads, nt, nc = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
seed = int(sys.argv[5]) if len(sys.argv) > 5 else 1
rng = np.random.default_rng(seed)
T = models.builtin_model("r73.t")["table"]
C = [models.builtin_model("r73.c.p1")["table"], models.builtin_model("r73.c.p2")["table"]]
reads = []
for k in range(n_reads):
    pm = tuple(synth.random_params(rng, 1)[0])
    t = synth.make_read(rng, T, nt, pm)
    c = synth.make_read(rng, C[k % 2], nc, pm) if nc > 0 else None
    if c is not None:
        c["start"] = (c["start"] + t["start"][-1] + np.float32(0.5)).astype(np.float32)
    reads.append((f"read{k:06d}", [t, c]))
evio.write_ncev(out, reads)
print(out, n_reads, "reads")
