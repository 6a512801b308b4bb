"""Datasets of the parity runs against the REAL reference program (oracle/_ref/nanocall_ref = unmodified nanocall.cpp,
compiled in place over a stand-in fast5::File).  Shared by tools/make_ref_golden.py, which runs the reference on them
where /root/reference exists and commits its outputs under tests/golden/, and by tests/test_ref_binary_gpu.py, which
regenerates the same inputs and runs nanocall-b200 on them.  Test infrastructure only."""
import os

import numpy as np

from nanocall_b200 import evio, models, synth

# name -> (seed, [(n_reads, nt, nc)], pore of the SYNTHETIC data, CLI options common to both programs)
DATASETS = {
    # config 3/4 in the small: 2D reads, r73 preset (double-strand scaling, <= 20 EM rounds, two candidate complement models)
    "r73_2d_small": (101, [(600, 1200, 1000)], "r73", ["--pore", "r73"]),
    # config 3 at its own shape: 5000 + 5000 events
    "r73_2d_5k": (102, [(100, 5000, 5000)], "r73", ["--pore", "r73"]),
    # per-strand scaling and selection (nanocall.cpp:463-571, 787-855), some 1D reads
    "r73_single": (103, [(80, 900, 800)], "r73", ["--pore", "r73", "--single-strand-scaling"]),
    # --no-train: initial scaling only, every model ranked by its Viterbi path (strands separately, nanocall.cpp:1012-1026)
    "r73_notrain": (104, [(40, 800, 700)], "r73", ["--pore", "r73", "--no-train"]),
    # the r9 preset (default pore: r9 models, no drift training, abasic offset 0) and --1d
    "r9_2d": (105, [(60, 900, 800)], "r9", ["--pore", "r9"]),
    "r9_1d": (106, [(40, 1500, 0)], "r9", ["--1d"]),
    # trims, --max-ed-events truncation, narrower FASTA lines, fewer training events
    "r73_opts": (107, [(60, 1500, 1300)], "r73", ["--pore", "r73", "--max-ed-events", "2500", "--trim-ed-sq-start", "30",
                                                   "--trim-ed-hp-end", "70", "--fasta-line-width", "60",
                                                   "--scaling-num-events", "120", "--scaling-max-rounds", "4"]),
    # --trans: a custom initial transition table in the reference's file format, produced by the reference's own
    # compute-state-transitions (full form with a probability cutoff: neighbour sets differ from the fast table's);
    # in force while a strand's transition parameters are the defaults, i.e. in the first training round, and for
    # the Viterbi pass as well when transitions are not trained
    "r73_trans": (108, [(50, 900, 800)], "r73", ["--pore", "r73", "--trans", "@TRANS:-k,0.3,-t,0.1,-p,0.0005@"]),
    "r73_trans_fixed": (109, [(40, 800, 700)], "r73", ["--pore", "r73", "--no-train-transitions", "--trans", "@TRANS:-k,0.25,-t,0.12,-p,0.002@"]),
}

CST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "compute_state_transitions")


def materialize_options(opts, out_dir):
    """Replace "@TRANS:<args>@" by the path of a transition table generated with the reference's tool (None if the tool
    is not built)."""
    out = []
    for o in opts:
        if o.startswith("@TRANS:"):
            if not os.path.exists(CST):
                return None
            import subprocess
            path = os.path.join(out_dir, "trans.tsv")
            subprocess.run([CST] + o[7:-1].split(",") + ["-o", path], check=True)
            out.append(path)
        else:
            out.append(o)
    return out


def write_inputs(name, out_dir):
    """-> list of file paths (one raw event table per read, named like fast5 files), in processing order."""
    seed, shapes, pore, _ = DATASETS[name]
    os.makedirs(out_dir, exist_ok=True)
    T = models.builtin_model(pore + ".t")["table"]
    C = [models.builtin_model(pore + ".c.p1")["table"], models.builtin_model(pore + ".c.p2")["table"]]
    rng = np.random.default_rng(seed)
    files = []
    k = 0
    for n_reads, nt, nc in shapes:
        for _ in range(n_reads):
            pm = tuple(synth.random_params(rng, 1)[0])
            a, b = nt + 13 * (k % 17), (nc + 7 * (k % 11)) if nc else 0
            if nc and k % 16 == 5:
                b = 0                      # a 1D read among the 2D ones
            level = 115.0 if pore == "r73" else 200.0
            ev = synth.make_raw_2d_read(rng, (T, C), a, b, pm, comp=k % 2, hairpin_level=level)
            if k % 16 == 9:
                ev["stdv"][120] = 0.0      # Event::update_logs: 0 -> 0.01
                ev["stdv"][130] = 4.5      # dropped by the event filter
            fn = os.path.join(out_dir, f"{name}_{k:04d}.fast5")
            evio.write_ncrw(fn, [(f"{name}.{k}" if k % 3 else "", 5000.0, ev)])
            files.append(fn)
            k += 1
    return files
