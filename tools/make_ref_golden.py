#!/usr/bin/env python3
"""Run the REAL reference program (oracle/_ref/nanocall_ref: unmodified nanocall.cpp) on the datasets of
tests/ref_reads.py and commit what it wrote: FASTA, --stats TSV and the scaling_result / selected_model / best_model
log lines -> tests/golden/ref_<dataset>.json.gz.  Needs /root/reference (to build nanocall_ref); CPU only, slow
(the reference needs ~20 s per 2D read and core).
usage: make_ref_golden.py [dataset ...] [-t threads]"""
import gzip
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_reads  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "nanocall_ref")
KEEP = re.compile(r"(scaling_result|selected_model|best_model|means_apart|scaling_regression) .*")


def main():
    args = sys.argv[1:]
    threads = 6
    if "-t" in args:
        i = args.index("-t")
        threads = int(args[i + 1])
        del args[i:i + 2]
    names = args or list(ref_reads.DATASETS)
    subprocess.run(["make", "-s", "-f", os.path.join(ROOT, "oracle", "Makefile"), "ref"], check=True)
    for name in names:
        with tempfile.TemporaryDirectory() as d:
            files = ref_reads.write_inputs(name, d)
            fofn = os.path.join(d, "fofn.txt")
            open(fofn, "w").write("\n".join(files) + "\n")
            opts = ref_reads.DATASETS[name][3]
            run_opts = ref_reads.materialize_options(opts, d)
            t0 = time.time()
            p = subprocess.run([REF] + run_opts + ["-t", str(threads), "-o", os.path.join(d, "out.fa"), "--stats",
                                               os.path.join(d, "stats.tsv"), fofn], capture_output=True, text=True)
            assert p.returncode == 0, p.stderr[-2000:]
            lines = [m.group(0) for m in (KEEP.search(l) for l in p.stderr.split("\n")) if m]
            out = {"dataset": name, "options": opts, "n_reads": len(files), "fasta": open(os.path.join(d, "out.fa")).read(),
                   "stats": open(os.path.join(d, "stats.tsv")).read(), "log": sorted(lines),
                   "reference_seconds": time.time() - t0, "threads": threads}
            path = os.path.join(ROOT, "tests", "golden", f"ref_{name}.json.gz")
            with gzip.open(path, "wt") as f:
                json.dump(out, f)
            print(name, len(files), "reads", f"{time.time() - t0:.0f} s", os.path.getsize(path), "bytes", flush=True)


if __name__ == "__main__":
    main()
