"""Parity of the CUDA Forward/Backward and training-round path with the oracle.
alpha, beta and log Pr[data] are compared BIT FOR BIT (the table-driven p7_FLogsum and the
reference's accumulation order are reproduced on the device).  Trained parameters are compared at
1e-5 relative (north_star allows 1e-4): they go through device expf/logf and tree-shaped float sums."""
import numpy as np
import pytest

from golden_util import MODEL_KEYS, load, same_bits, train_seqs
from nanocall_b200 import synth

pytestmark = pytest.mark.gpu

T, C1, C2 = "r73.t.006.ont.model", "r73.c.p1.006.ont.model", "r73.c.p2.006.ont.model"
RTOL = 2e-7  # a couple of float ulps: the round is reproduced essentially bit for bit


def _close(a, b, rtol=RTOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.all((np.isnan(a) & np.isnan(b)) | (np.abs(a - b) <= rtol * np.maximum(np.abs(b), 1e-3)))


@pytest.mark.parametrize("model,n,pm,st", [
    (T, 60, (1, 0, 0, 1, 1, 1), (0.1, 0.3)),
    (C1, 100, (1.05, -2.0, 0.003, 1.1, 0.9, 1.2), (0.1476, 0.2379)),
    (C2, 33, (0.93, 4.5, -0.004, 1.27, 1.18, 0.83), (0.05, 0.4)),
    (T, 1, (1, 0, 0, 1, 1, 1), (0.1, 0.3)),
    (T, 2, (1, 0, 0, 1, 1, 1), (0.4, 0.05)),
])
def test_forward_backward_bit_exact(ctx, port, models, model, n, pm, st):
    table = models[model]["table"]
    mid = ctx.register_model(table, 2)
    rd = synth.make_read(np.random.default_rng(n), table, n, pm)
    got = ctx.forward_backward(mid, pm, st, rd["mean"], rd["stdv"], rd["start"])
    exp = port.fwbw(table, np.array(pm, np.float32), st[0], st[1], rd["mean"], rd["stdv"], rd["start"])
    bad = np.argwhere(got["alpha"].view(np.uint32) != exp["alpha"].view(np.uint32))
    assert bad.size == 0, ("alpha", bad[:5], got["alpha"][tuple(bad[0])], exp["alpha"][tuple(bad[0])])
    bad = np.argwhere(got["beta"].view(np.uint32) != exp["beta"].view(np.uint32))
    assert bad.size == 0, ("beta", bad[:5], got["beta"][tuple(bad[0])], exp["beta"][tuple(bad[0])])
    assert same_bits(got["log_pr_data"], exp["log_pr_data"])


def test_forward_backward_golden(ctx, models):
    f = load("fwbw")
    mid = ctx.register_model(models[MODEL_KEYS["t"]]["table"], 0)
    r = ctx.forward_backward(mid, f["a_pm"], f["a_st"], f["a_mean"], f["a_stdv"], f["a_start"])
    assert same_bits(r["alpha"], f["a_alpha"]) and same_bits(r["beta"], f["a_beta"]) and same_bits(r["log_pr_data"], f["a_logz"])
    r = ctx.forward_backward(mid, f["b_pm"], f["b_st"], f["b_mean"], f["b_stdv"], f["b_start"])
    assert same_bits(r["log_pr_data"], f["b_logz"])
    assert same_bits(r["alpha"][-1], f["b_alpha_last"]) and same_bits(r["beta"][0], f["b_beta_first"])


def test_forward_backward_zero_stdv(ctx, port, models):
    table = models[T]["table"]
    mid = ctx.register_model(table, 0)
    rd = synth.make_read(np.random.default_rng(4), table, 20)
    rd["stdv"][[0, 7, 19]] = 0.0
    got = ctx.forward_backward(mid, None, None, rd["mean"], rd["stdv"], rd["start"])
    exp = port.fwbw(table, np.array([1, 0, 0, 1, 1, 1], np.float32), 0.1, 0.3, rd["mean"], rd["stdv"], rd["start"])
    assert same_bits(got["alpha"], exp["alpha"]) and same_bits(got["beta"], exp["beta"])


def _check_round(got, exp, strands=(0, 1)):
    assert same_bits(got["fit"], exp["fit"]), (got["fit"], exp["fit"])
    assert got["done"] == exp["done"]
    assert _close(got["pm"], exp["pm"]), (got["pm"], exp["pm"])
    for sd in strands:
        assert _close(got["st"][2 * sd:2 * sd + 2], exp["st"][2 * sd:2 * sd + 2]), (got["st"], exp["st"])


def test_train_round_golden_chain(ctx, models):
    """Three chained double-strand rounds from the golden file: same inputs as the reference saw."""
    t = load("train")
    Tt, Ct = models[MODEL_KEYS["t"]]["table"], models[MODEL_KEYS["c1"]]["table"]
    m0, m1 = ctx.register_model(Tt, 0), ctx.register_model(Ct, 1)
    S = train_seqs(t)
    for rnd in range(3):
        got = ctx.train_one_round(S, (m0, m1), t[f"d{rnd}_in_pm"], t[f"d{rnd}_in_st"])
        _check_round(got, dict(pm=t[f"d{rnd}_pm"], st=t[f"d{rnd}_st"], fit=t[f"d{rnd}_fit"], done=bool(t[f"d{rnd}_done"])))
    got = ctx.train_one_round(S[2:], (m1, m1), t["s_in_pm"], t["s_in_st"])
    _check_round(got, dict(pm=t["s_pm"], st=t["s_st"], fit=t["s_fit"], done=bool(t["s_done"])), strands=(1,))
    assert np.isnan(got["st"][:2]).all()  # strand without sequences: NaN, as the reference leaves it
    got = ctx.train_one_round(S, (m0, m1), t["s_in_pm"], t["s_in_st"], train_scaling=False)
    _check_round(got, dict(pm=t["ns_pm"], st=t["ns_st"], fit=t["ns_fit"], done=False))
    assert same_bits(got["pm"], t["s_in_pm"])
    got = ctx.train_one_round(S, (m0, m1), t["s_in_pm"], t["s_in_st"], train_transitions=False)
    _check_round(got, dict(pm=t["nt_pm"], st=t["nt_st"], fit=t["nt_fit"], done=False))
    assert same_bits(got["st"], t["s_in_st"])


@pytest.mark.parametrize("train_drift", [True, False])
def test_train_round_batch_vs_oracle(ctx, port, models, train_drift):
    """A batch of groups (double- and single-strand, different candidates) in one call."""
    rng = np.random.default_rng(23)
    tabs = {k: models[k]["table"] for k in (T, C1, C2)}
    ids = {k: ctx.register_model(tabs[k], 0 if k == T else 1) for k in tabs}
    groups, expect = [], []
    for g in range(6):
        true = synth.random_params(rng, 1)[0]
        cm = C1 if g % 2 == 0 else C2
        r0 = synth.make_read(rng, tabs[T], 300, tuple(true))
        r1 = synth.make_read(rng, tabs[cm], 260, tuple(true))
        n = int(rng.integers(20, 60))
        s0 = [(0, r0["mean"][:n], r0["stdv"][:n], r0["start"][:n]), (0, r0["mean"][-n:], r0["stdv"][-n:], r0["start"][-n:])]
        s1 = [(1, r1["mean"][:n], r1["stdv"][:n], r1["start"][:n]), (1, r1["mean"][-n:], r1["stdv"][-n:], r1["start"][-n:])]
        pm = np.array([1.0, float(rng.uniform(-1, 1)), 0, 1, 1, 1], np.float32)
        st = np.array([0.1, 0.3, 0.1, 0.3], np.float32) if g < 3 else \
            np.array([rng.uniform(.06, .3), rng.uniform(.1, .35), rng.uniform(.06, .3), rng.uniform(.1, .35)], np.float32)
        if g == 4:
            seqs, mids, tb = s0, (ids[T], ids[T]), (tabs[T], tabs[T])      # single-strand template
        elif g == 5:
            seqs, mids, tb = s1, (ids[cm], ids[cm]), (tabs[cm], tabs[cm])  # single-strand complement
        else:
            seqs, mids, tb = s0 + s1, (ids[T], ids[cm]), (tabs[T], tabs[cm])
        groups.append(dict(seqs=seqs, model_id=mids, pm=pm, st=st))
        expect.append(port.train_one_round(seqs, tb[0], tb[1], pm, st, train_drift=train_drift))
    got = ctx.train_round_batch(groups, train_drift=train_drift)
    for g, (a, b) in enumerate(zip(got, expect)):
        strands = (0, 1) if g < 4 else ((0,) if g == 4 else (1,))
        _check_round(a, b, strands)


def test_train_round_waves(ctx, port, models):
    """Force several waves (tiny scratch limit is not exposed; use many groups instead) and check order."""
    rng = np.random.default_rng(5)
    table = models[T]["table"]
    mid = ctx.register_model(table, 0)
    groups = []
    for g in range(40):
        rd = synth.make_read(rng, table, 30)
        groups.append(dict(seqs=[(0, rd["mean"], rd["stdv"], rd["start"])], model_id=(mid, mid),
                           pm=np.array([1, 0, 0, 1, 1, 1], np.float32), st=np.array([.1, .3, .1, .3], np.float32)))
    got = ctx.train_round_batch(groups)
    for g in (0, 17, 39):
        exp = port.train_one_round(groups[g]["seqs"], table, table, groups[g]["pm"], groups[g]["st"])
        _check_round(got[g], exp, strands=(0,))


def test_custom_transition_table_kernels_match_the_oracle(ctx, port, models):
    """--trans: a table given as stored edges takes the list-walking kernels.  Fed with the edges of the parametric table
    itself (the oracle's to_v lists, in order), Viterbi and Forward/Backward must reproduce the oracle bit for bit; fed
    with the same edges in REVERSED file order, Viterbi's tie rule and the fold order follow the file."""
    import numpy as np
    from nanocall_b200 import api, synth
    table = models["r73.t.006.ont.model"]["table"]
    mid = ctx.register_model(table, 0)
    tr = port.transitions(0.1, 0.3)
    frm, to, lp = [], [], []
    for i in range(4096):
        for k in range(int(tr["to_cnt"][i])):
            frm.append(i); to.append(int(tr["to_idx"][i, k])); lp.append(tr["to_lp"][i, k])
    rng = np.random.default_rng(3)
    pm = synth.random_params(rng, 1)[0]
    rd = synth.make_read(rng, table, 300, tuple(pm))
    try:
        ctx.set_default_transitions(0.1, 0.3, frm, to, lp)
        out = ctx.viterbi(np.array([0, 300], np.uint64), rd["mean"], rd["stdv"], rd["start"], mid, pm=pm)
        exp = port.viterbi(table, pm, 0.1, 0.3, rd["mean"], rd["stdv"], rd["start"])
        assert out["path_logprob"].view(np.uint32)[0] == np.float32(exp["path_prob"]).view(np.uint32)
        assert np.array_equal(out["states"].astype(np.uint32), exp["states"])
        assert np.array_equal(out["moves"].astype(np.int32), exp["moves"])
        fb = ctx.forward_backward(mid, pm, (0.1, 0.3), rd["mean"][:60], rd["stdv"][:60], rd["start"][:60])
        efb = port.fwbw(table, pm, 0.1, 0.3, rd["mean"][:60], rd["stdv"][:60], rd["start"][:60])
        assert np.array_equal(fb["alpha"].view(np.uint32), efb["alpha"].view(np.uint32))
        assert np.array_equal(fb["beta"].view(np.uint32), efb["beta"].view(np.uint32))
        assert fb["log_pr_data"].view(np.uint32) == efb["log_pr_data"].view(np.uint32)
        # other parameters than the defaults: the parametric kernels, untouched by the table
        out2 = ctx.viterbi(np.array([0, 300], np.uint64), rd["mean"], rd["stdv"], rd["start"], mid, pm=pm, st=(0.12, 0.25))
        exp2 = port.viterbi(table, pm, 0.12, 0.25, rd["mean"], rd["stdv"], rd["start"])
        assert out2["path_logprob"].view(np.uint32)[0] == np.float32(exp2["path_prob"]).view(np.uint32)
        # reversed file order: same Viterbi score (max is order-free), forward sums folded in another order
        ctx.set_default_transitions(0.1, 0.3, frm[::-1], to[::-1], lp[::-1])
        out3 = ctx.viterbi(np.array([0, 300], np.uint64), rd["mean"], rd["stdv"], rd["start"], mid, pm=pm)
        assert out3["path_logprob"].view(np.uint32)[0] == np.float32(exp["path_prob"]).view(np.uint32)
    finally:
        ctx.set_default_transitions(0.1, 0.3)
