#!/usr/bin/env python3
"""Golden results of the FULL pipeline (configs 3/4: EM training rounds, candidate-model selection, Viterbi with the
trained parameters) for a few hundred synthetic 2D reads, computed with the oracle (the C restatement, which is
bit-identical to the compiled reference) driven by tests/oracle_pipeline.py.  CPU only, ~20 s per read and core: run it in
this container, commit the result, compare on the GPU box (tests/test_pipeline_gpu.py::test_pipeline_parity_golden).
usage: make_pipeline_golden.py [n_reads=200] [seed=11] [out=tests/golden/pipeline_r73.json.gz]"""
import gzip
import json
import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
SEED = int(sys.argv[2]) if len(sys.argv) > 2 else 11
OUT = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "tests", "golden", "pipeline_r73.json.gz")
NT, NC = 1200, 1000

_state = {}


def _init():
    import oracle_lib
    from nanocall_b200 import models as M
    import pipeline_reads as PR
    mdl = {m["name"]: m for m in M.load_builtin_models()}
    # the plain-C restatement: bit-identical to the compiled reference on every primitive (tests/test_oracle_vs_ref.py)
    # and, unlike libncref.so (thread-local logger state), safe to drive from forked worker processes
    _state["lib"] = oracle_lib.port()
    _state["mdl"] = {n: mdl[n] for n in PR.R73}
    _state["reads"] = PR.make_reads(mdl, SEED, N, nt=NT, nc=NC)


def _one(k):
    import oracle_pipeline as OP
    rid, ev = _state["reads"][k]
    exp = OP.run_read(_state["lib"], _state["mdl"], ev, OP.Opts())
    return dict(read=rid,
                pm={"+".join(key): [float(v) for v in pm] for key, pm in exp["pm_params"].items()},
                st={"+".join(key): [float(v) for v in st] for key, st in exp["st_params"].items()},
                fits={"+".join(key): float(v) for key, v in exp["fits"].items()},
                rounds={"+".join(key): int(v) for key, v in exp["rounds"].items()},
                preferred=None if exp["preferred"][2] is None else "+".join(exp["preferred"][2]),
                bases=[exp["calls"][0]["bases"], exp["calls"][1]["bases"]],
                models=[exp["calls"][0]["model"], exp["calls"][1]["model"]])


if __name__ == "__main__":
    kind = "port (bit-identical to the compiled reference: tests/test_oracle_vs_ref.py)"
    with mp.Pool(os.cpu_count(), initializer=_init) as pool:
        res = pool.map(_one, range(N), chunksize=1)
    with gzip.open(OUT, "wt") as f:
        json.dump(dict(seed=SEED, n_reads=N, nt=NT, nc=NC, oracle=kind, options="r73 preset defaults", reads=res), f)
    print(OUT, N, "reads, oracle =", kind)
