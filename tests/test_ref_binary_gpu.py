"""nanocall-b200 against the REAL reference program: tests/golden/ref_*.json.gz hold what the unmodified
nanocall.cpp (compiled in place over a stand-in fast5::File, oracle/Makefile target nanocall_ref) wrote for the
datasets of tests/ref_reads.py -- FASTA, --stats TSV and its scaling_result / selected_model / best_model log lines.
The same raw event tables are regenerated here and run through the CLI on the GPU.

north_star's bars: basecalls identical on >= 99.9 % of reads (edit distance reported for the rest), trained scaling
parameters within 1e-4, path log-likelihoods within 1e-5 relative."""
import gzip
import json
import os
import re
import subprocess

import numpy as np
import pytest

import ref_reads

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "nanocall_b200", "bin", "nanocall-b200")
KEEP = re.compile(r"(scaling_result|selected_model|best_model|means_apart|scaling_regression) .*")
NUM = re.compile(r"-?(?:\d+\.?\d*(?:e[-+]?\d+)?|nan|inf)", re.I)


def edit_distance(a, b):
    """Myers' bit-parallel Levenshtein distance (Python integers as bit vectors)."""
    if not a:
        return len(b)
    peq = {}
    for i, ch in enumerate(a):
        peq[ch] = peq.get(ch, 0) | (1 << i)
    m = len(a)
    mask, top = (1 << m) - 1, 1 << (m - 1)
    pv, mv, score = mask, 0, m
    for ch in b:
        eq = peq.get(ch, 0)
        xv = eq | mv
        xh = (((eq & pv) + pv) ^ pv) | eq
        ph = mv | ~(xh | pv)
        mh = pv & xh
        if ph & top:
            score += 1
        elif mh & top:
            score -= 1
        ph = ((ph << 1) | 1) & mask
        mh = (mh << 1) & mask
        pv = (mh | ~(xv | ph)) & mask
        mv = ph & xv
    return score


def _fasta(text):
    rec, name = {}, None
    for line in text.split("\n"):
        if line.startswith(">"):
            name = line[1:]
            rec[name] = []
        elif line:
            rec[name].append(line)
    return rec


def _close(a, b, rel):
    a, b = float(a), float(b)
    if np.isnan(a) or np.isnan(b):
        return np.isnan(a) and np.isnan(b)
    return abs(a - b) <= rel * max(abs(a), abs(b), 1e-2)


def _same_line(x, y, rel):
    """Same text up to the numbers, numbers within rel."""
    if NUM.sub("#", x) != NUM.sub("#", y):
        return False
    return all(_close(p, q, rel) for p, q in zip(NUM.findall(x), NUM.findall(y)))


def _golden(name):
    path = os.path.join(ROOT, "tests", "golden", f"ref_{name}.json.gz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated (tools/make_ref_golden.py needs /root/reference and CPU hours)")
    with gzip.open(path, "rt") as f:
        return json.load(f)


@pytest.mark.parametrize("name", sorted(ref_reads.DATASETS))
def test_cli_matches_the_reference_program(name, tmp_path):
    gold = _golden(name)
    d = str(tmp_path)
    files = ref_reads.write_inputs(name, d)
    assert len(files) == gold["n_reads"]
    fofn = os.path.join(d, "fofn.txt")
    open(fofn, "w").write("\n".join(files) + "\n")
    out, stats = os.path.join(d, "out.fa"), os.path.join(d, "stats.tsv")
    opts = ref_reads.materialize_options(gold["options"], d)
    if opts is None:
        pytest.skip("oracle/_ref/compute_state_transitions (the reference's --trans table generator) is not built")
    p = subprocess.run([CLI] + opts + ["-o", out, "--stats", stats, fofn], capture_output=True, text=True, timeout=1800)
    assert p.returncode == 0, p.stderr[-3000:]

    # ---- FASTA: same records in the same order, same line wrapping; sequences identical on >= 99.9 % of the records
    got, exp = _fasta(open(out).read()), _fasta(gold["fasta"])
    assert list(got) == list(exp)
    diff = [(k, edit_distance("".join(got[k]), "".join(exp[k]))) for k in exp if got[k] != exp[k]]
    report = {"dataset": name, "records": len(exp), "identical": len(exp) - len(diff), "edit_distances": diff[:20]}
    with open(os.path.join(ROOT, "gpurun_out", f"ref_parity_{name}.json") if os.path.isdir(os.path.join(ROOT, "gpurun_out"))
              else os.path.join(d, "report.json"), "w") as f:
        json.dump(report, f)
    assert len(diff) <= 1e-3 * len(exp), report

    # ---- log lines: rounds, model selection and ranking identical; parameters within 1e-4, path probabilities 1e-5
    mine = sorted(m.group(0) for m in (KEEP.search(l) for l in p.stderr.split("\n")) if m)
    theirs = gold["log"]
    kinds = lambda ls, k: [l for l in ls if l.startswith(k)]
    for kind, rel in (("selected_model", 0.0), ("scaling_result", 1e-4), ("best_model", 1e-4), ("means_apart", 1e-3)):
        a, b = kinds(mine, kind), kinds(theirs, kind)
        assert len(a) == len(b), (kind, len(a), len(b))
        bad = [(x, y) for x, y in zip(a, b) if not _same_line(x, y, rel)]
        assert not bad, (kind, len(bad), bad[:2])
    for x, y in zip(kinds(mine, "best_model"), kinds(theirs, "best_model")):
        px, py = float(x.rsplit("[", 1)[1].rstrip("]")), float(y.rsplit("[", 1)[1].rstrip("]"))
        assert _close(px, py, 1e-5), (x, y)

    # ---- --stats: same rows; text identical up to the trained parameters (five decimals printed), those within 1e-4
    a, b = open(stats).read().split("\n"), gold["stats"].split("\n")
    assert len(a) == len(b)
    assert a[0] == b[0]
    for x, y in zip(a[1:], b[1:]):
        fx, fy = x.split("\t"), y.split("\t")
        assert len(fx) == len(fy)
        for c, (u, v) in enumerate(zip(fx, fy)):
            if u == v:
                continue
            assert c >= 9 and abs(float(u) - float(v)) <= 1e-4 * max(abs(float(v)), 1.0) + 1.5e-5, (x, y)
