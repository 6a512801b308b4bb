"""Pins oracle/nc_oracle.c against the reference's own code (oracle/_ref/libncref.so) on fresh seeded
inputs, bit for bit.  Skipped where neither the prebuilt .so nor /root/reference exists."""
import numpy as np
import pytest

from golden_util import same_bits
from nanocall_b200 import synth

T, C1, C2 = "r73.t.006.ont.model", "r73.c.p1.006.ont.model", "r73.c.p2.006.ont.model"


def test_tables(port, ref):
    assert same_bits(port.flogsum_table(), ref.flogsum_table())
    assert np.array_equal(port.st_train_kmers(), ref.st_train_kmers())
    rng = np.random.default_rng(3)
    for _ in range(3):
        ps, pk = float(rng.uniform(0.05, 0.4)), float(rng.uniform(0.05, 0.4))
        a, b = ref.transitions(ps, pk), port.transitions(ps, pk)
        for k in a:
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_viterbi_random(port, ref, models, seed):
    rng = np.random.default_rng(seed)
    table = models[[T, C1, C2][seed % 3]]["table"]
    pm = synth.random_params(rng, 1)[0]
    ps, pk = float(rng.uniform(0.05, 0.4)), float(rng.uniform(0.05, 0.4))
    rd = synth.make_read(rng, table, int(rng.integers(100, 600)), tuple(pm))
    a = ref.viterbi(table, pm, ps, pk, rd["mean"], rd["stdv"], rd["start"])
    b = port.viterbi(table, pm, ps, pk, rd["mean"], rd["stdv"], rd["start"])
    assert same_bits(a["path_prob"], b["path_prob"])
    assert np.array_equal(a["states"], b["states"]) and np.array_equal(a["moves"], b["moves"]) and a["bases"] == b["bases"]


def test_viterbi_batch_threaded(port, ref, models):
    table = models[T]["table"]
    batch = synth.make_batch(9, table, [120, 333, 57, 200, 90])
    pm = np.tile(np.array([1, 0, 0, 1, 1, 1], np.float32), (5, 1))
    st = np.tile(np.array([0.1, 0.3], np.float32), (5, 1))
    a = ref.viterbi_batch(table, batch["ev_off"], batch["mean"], batch["stdv"], batch["start"], pm, st, n_threads=3)
    b = port.viterbi_batch(table, batch["ev_off"], batch["mean"], batch["stdv"], batch["start"], pm, st, n_threads=3)
    assert same_bits(a["path_prob"], b["path_prob"]) and np.array_equal(a["states"], b["states"])


def test_fwbw_and_training_random(port, ref, models):
    rng = np.random.default_rng(17)
    pm = synth.random_params(rng, 1)[0]
    r0 = synth.make_read(rng, models[T]["table"], 300, tuple(pm))
    r1 = synth.make_read(rng, models[C2]["table"], 300, tuple(pm))
    a = ref.fwbw(models[T]["table"], pm, 0.13, 0.21, r0["mean"][:40], r0["stdv"][:40], r0["start"][:40])
    b = port.fwbw(models[T]["table"], pm, 0.13, 0.21, r0["mean"][:40], r0["stdv"][:40], r0["start"][:40])
    assert same_bits(a["alpha"], b["alpha"]) and same_bits(a["beta"], b["beta"]) and same_bits(a["log_pr_data"], b["log_pr_data"])
    S = [(0, r0["mean"][:40], r0["stdv"][:40], r0["start"][:40]), (0, r0["mean"][-40:], r0["stdv"][-40:], r0["start"][-40:]),
         (1, r1["mean"][:40], r1["stdv"][:40], r1["start"][:40]), (1, r1["mean"][-40:], r1["stdv"][-40:], r1["start"][-40:])]
    guess = np.array([1.0, 0.3, 0, 1, 1, 1], np.float32)
    st = np.array([.1, .3, .1, .3], np.float32)
    a = ref.train_one_round(S, models[T]["table"], models[C2]["table"], guess, st)
    b = port.train_one_round(S, models[T]["table"], models[C2]["table"], guess, st)
    assert same_bits(a["pm"], b["pm"]) and same_bits(a["st"], b["st"]) and same_bits(a["fit"], b["fit"]) and a["done"] == b["done"]


@pytest.mark.gpu
def test_port_equals_reference_in_the_gpu_session(port, ref, models):
    """The GPU parity tests compare the CUDA path with the C port; the port is pinned against the compiled reference by the
    tests above, which run in the CPU session.  The same comparison once more inside the `-m gpu` session closes the chain
    CUDA == port == reference on ONE box; it needs no GPU itself."""
    test_tables(port, ref)
    test_viterbi_random(port, ref, models, 2)
    test_fwbw_and_training_random(port, ref, models)
