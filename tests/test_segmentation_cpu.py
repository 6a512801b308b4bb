"""Fast5_Summary on raw event tables (host/reads.cpp: abasic level, hairpin detection, trims, --max-ed-events, --1d) against
the unmodified reference program, WITHOUT a GPU: `nanocall-b200 --summarize-only --stats` runs the loaders and the
segmentation only; the first eight --stats columns (file, read, number of events, abasic level, the four strand bounds) of
every read must be what the reference program printed (tests/golden/ref_*.json.gz, tools/make_ref_golden.py).  The trained
parameter columns are the GPU tests' business (tests/test_ref_binary_gpu.py)."""
import gzip
import json
import os
import subprocess

import pytest

import ref_reads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "nanocall_b200", "bin", "nanocall-b200")


@pytest.mark.parametrize("name", sorted(ref_reads.DATASETS))
def test_segmentation_matches_the_reference_program(name, tmp_path):
    path = os.path.join(ROOT, "tests", "golden", f"ref_{name}.json.gz")
    if not os.path.exists(path) or not os.path.exists(CLI):
        pytest.skip("golden or CLI not built")
    with gzip.open(path, "rt") as f:
        gold = json.load(f)
    d = str(tmp_path)
    files = ref_reads.write_inputs(name, d)
    fofn = os.path.join(d, "fofn.txt")
    open(fofn, "w").write("\n".join(files) + "\n")
    opts = ref_reads.materialize_options(gold["options"], d)
    if opts is None:
        # (a --trans table is not needed to segment reads: drop the option and its file)
        opts, skip = [], False
        for o in gold["options"]:
            if skip:
                skip = False
                continue
            if o in ("-s", "--trans"):
                skip = True
                continue
            opts.append(o)
    stats = os.path.join(d, "stats.tsv")
    p = subprocess.run([CLI] + opts + ["--summarize-only", "--stats", stats, fofn], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    got = [l.split("\t")[:8] for l in open(stats).read().strip().split("\n")]
    exp = [l.split("\t")[:8] for l in gold["stats"].strip().split("\n")]
    assert len(got) == len(exp) == gold["n_reads"] + 1
    bad = [(g, e) for g, e in zip(got, exp) if g != e]
    assert not bad, (len(bad), bad[:3])
