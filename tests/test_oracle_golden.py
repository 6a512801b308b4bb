"""The C restatement (oracle/nc_oracle.c) against the golden vectors produced by the reference's own
compiled headers (tools/make_golden.py).  Bit-for-bit: floats are compared by bit pattern."""
import numpy as np

from golden_util import MODEL_KEYS, crc, load, same_bits, train_seqs


def test_flogsum_table_and_values(port):
    g = load("tables")
    tbl = port.flogsum_table()
    assert crc(tbl) == g["flogsum_table_crc"]
    assert same_bits(tbl[:64], g["flogsum_table_head"]) and same_bits(tbl[-64:], g["flogsum_table_tail"])
    got = np.array([port.flogsum(a, b) for a, b in g["flogsum_pairs"]], np.float32)
    assert same_bits(got, g["flogsum_vals"])


def test_transitions(port):
    g = load("tables")
    for k, (ps, pk) in enumerate(g["st_cases"]):
        tr = port.transitions(float(ps), float(pk))
        assert np.array_equal(tr["from_cnt"].astype(np.uint8), g[f"tr{k}_from_cnt"])
        assert int(tr["from_cnt"].sum()) == 85936 and int(tr["to_cnt"].sum()) == 85936
        for name in ("from_idx", "from_lp", "to_idx", "to_lp", "to_cnt"):
            assert crc(tr[name]) == g[f"tr{k}_{name}_crc"], (k, name)


def test_scaled_models_and_stats(port, models):
    g = load("tables")
    for k, pm in enumerate(g["pm_cases"]):
        for name in ("t", "c1"):
            sm = port.scaled_model(models[MODEL_KEYS[name]]["table"], pm)
            for f in ("level_mean", "level_stdv", "log_level_stdv", "sd_mean", "sd_lambda", "log_sd_lambda"):
                assert crc(sm[f]) == g[f"sm{k}_{name}_{f}_crc"], (k, name, f)
            assert same_bits([sm["mean"], sm["stdv"]], g[f"sm{k}_{name}_stats"])
    assert same_bits(port.mean_stdv(g["ms_x"]), g["ms_out"])
    km = port.st_train_kmers()
    assert km.size == int(g["st_train_kmers_n"]) == 2160 and crc(km) == g["st_train_kmers_crc"]


def test_viterbi(port, models):
    v = load("viterbi")
    for k in range(int(v["n_cases"])):
        table = models[MODEL_KEYS[str(v[f"c{k}_model"])]]["table"]
        st = v[f"c{k}_st"]
        dump = f"c{k}_alpha" in v
        r = port.viterbi(table, v[f"c{k}_pm"], float(st[0]), float(st[1]), v[f"c{k}_mean"], v[f"c{k}_stdv"],
                         v[f"c{k}_start"], dump=dump)
        assert same_bits(r["path_prob"], v[f"c{k}_path_prob"]), k
        assert np.array_equal(r["states"].astype(np.uint16), v[f"c{k}_states"]), k
        assert np.array_equal(r["moves"].astype(np.uint8), v[f"c{k}_moves"]), k
        assert r["bases"] == str(v[f"c{k}_bases"]), k
        if dump:
            assert same_bits(r["alpha"], v[f"c{k}_alpha"])


def test_forward_backward(port, models):
    f = load("fwbw")
    table = models[MODEL_KEYS["t"]]["table"]
    r = port.fwbw(table, f["a_pm"], float(f["a_st"][0]), float(f["a_st"][1]), f["a_mean"], f["a_stdv"], f["a_start"])
    assert same_bits(r["alpha"], f["a_alpha"]) and same_bits(r["beta"], f["a_beta"])
    assert same_bits(r["log_pr_data"], f["a_logz"])
    r = port.fwbw(table, f["b_pm"], float(f["b_st"][0]), float(f["b_st"][1]), f["b_mean"], f["b_stdv"], f["b_start"])
    assert same_bits(r["log_pr_data"], f["b_logz"])
    assert crc(r["alpha"]) == f["b_alpha_crc"] and crc(r["beta"]) == f["b_beta_crc"]


def test_train_one_round(port, models):
    t = load("train")
    T, C1 = models[MODEL_KEYS["t"]]["table"], models[MODEL_KEYS["c1"]]["table"]
    S = train_seqs(t)
    for rnd in range(3):
        o = port.train_one_round(S, T, C1, t[f"d{rnd}_in_pm"], t[f"d{rnd}_in_st"])
        assert same_bits(o["pm"], t[f"d{rnd}_pm"]) and same_bits(o["st"], t[f"d{rnd}_st"])
        assert same_bits(o["fit"], t[f"d{rnd}_fit"]) and o["done"] == bool(t[f"d{rnd}_done"])
    o = port.train_one_round(S[2:], C1, C1, t["s_in_pm"], t["s_in_st"])
    assert same_bits(o["pm"], t["s_pm"]) and same_bits(o["fit"], t["s_fit"])
    assert np.isnan(o["st"][:2]).all() and same_bits(o["st"][2:], t["s_st"][2:])
    o = port.train_one_round(S, T, C1, t["s_in_pm"], t["s_in_st"], train_scaling=False)
    assert same_bits(o["pm"], t["ns_pm"]) and same_bits(o["st"], t["ns_st"]) and same_bits(o["fit"], t["ns_fit"])
    o = port.train_one_round(S, T, C1, t["s_in_pm"], t["s_in_st"], train_transitions=False)
    assert same_bits(o["pm"], t["nt_pm"]) and same_bits(o["st"], t["nt_st"]) and same_bits(o["fit"], t["nt_fit"])
