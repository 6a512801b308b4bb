"""Parity of the CUDA Viterbi path (through the C ABI) with the oracle: bit-exact path
log-probability, identical states / moves / base sequence (integer work: exact; the float
path log-probability is compared by bit pattern, tighter than north_star's 1e-5 relative)."""
import numpy as np
import pytest

from nanocall_b200 import api, synth

pytestmark = pytest.mark.gpu

R73T = "r73.t.006.ont.model"
R73C1 = "r73.c.p1.006.ont.model"
R73C2 = "r73.c.p2.006.ont.model"


@pytest.fixture(autouse=True, params=["alpha", "backpointer"])
def vit_mode(request, ctx):
    """Every test runs against both Viterbi kernels: the alpha-column kernel (fast path, arg max evaluated in the
    traceback) and the backpointer kernel (long-read path); both must give the reference's bits."""
    from nanocall_b200 import _lib as L
    ctx.set_viterbi_mode(L.NC_VIT_BACKPOINTER if request.param == "backpointer" else L.NC_VIT_AUTO)
    yield request.param
    ctx.set_viterbi_mode(L.NC_VIT_AUTO)


def _bits(x):
    return np.asarray(x, np.float32).view(np.uint32)


def _check_batch(ctx, port, table, mid, batch, pm, st):
    out = ctx.viterbi(batch["ev_off"], batch["mean"], batch["stdv"], batch["start"], mid, pm, st)
    n = batch["ev_off"].size - 1
    pm_a, st_a = api._pm_array(pm, n), api._st_array(st, n)
    exp = port.viterbi_batch(table, batch["ev_off"], batch["mean"], batch["stdv"], batch["start"], pm_a, st_a,
                             n_threads=8)
    assert np.array_equal(_bits(out["path_logprob"]), _bits(exp["path_prob"])), \
        (out["path_logprob"], exp["path_prob"])
    assert np.array_equal(out["states"].astype(np.uint32), exp["states"])
    assert np.array_equal(out["moves"].astype(np.int32), exp["moves"])
    return out


def test_identity_scaling_various_lengths(ctx, port, models):
    table = models[R73T]["table"]
    mid = ctx.register_model(table, 0)
    lengths = [1, 2, 3, 17, 64, 65, 127, 128, 129, 130, 257, 500, 1000, 2049]
    batch = synth.make_batch(11, table, lengths)
    _check_batch(ctx, port, table, mid, batch, None, None)


def test_scaled_and_custom_transitions(ctx, port, models):
    table = models[R73C1]["table"]
    mid = ctx.register_model(table, 1)
    rng = np.random.default_rng(5)
    n = 24
    lengths = rng.integers(50, 900, n)
    pm = synth.random_params(rng, n)
    st = np.stack([rng.uniform(0.05, 0.4, n), rng.uniform(0.05, 0.4, n)], 1).astype(np.float32)
    st[0] = (0.1, 0.3)
    st[1] = (0.05, 0.4)
    st[2] = (0.4, 0.05)
    batch = synth.make_batch(12, table, lengths, pm)
    _check_batch(ctx, port, table, mid, batch, pm, st)


def test_more_jobs_than_sms_and_job_order(ctx, port, models):
    """400 short jobs > 148 CTAs: exercises the persistent job loop; outputs must land at the job's
    own offsets whatever order the CTAs pick them up in."""
    table = models[R73T]["table"]
    mid = ctx.register_model(table, 0)
    rng = np.random.default_rng(6)
    lengths = rng.integers(20, 260, 400)
    batch = synth.make_batch(13, table, lengths)
    _check_batch(ctx, port, table, mid, batch, None, None)


def test_ties_and_plateaus(ctx, port, models):
    """Constant events make whole columns of near-ties; the lowest predecessor / lowest final state
    must win exactly as the reference's strict '>' scans do (Viterbi.hpp:84,127)."""
    table = models[R73T]["table"]
    mid = ctx.register_model(table, 0)
    n = 300
    off = np.array([0, n, 2 * n], np.uint64)
    mean = np.concatenate([np.full(n, 58.0, np.float32), np.tile(np.array([50., 66.], np.float32), n // 2)])
    stdv = np.concatenate([np.full(n, 0.9, np.float32), np.full(n, 1.1, np.float32)])
    start = np.tile(np.arange(n, dtype=np.float32) * 0.02, 2)
    batch = dict(ev_off=off, mean=mean, stdv=stdv, start=start)
    _check_batch(ctx, port, table, mid, batch, None, None)


def test_flat_model_all_states_tie(ctx, port):
    """A model whose 4096 states are identical: every comparison is an exact tie."""
    table = np.tile(np.array([60.0, 1.0, 0.8, 0.25], np.float32), (4096, 1))
    mid = ctx.register_model(table, 2)
    rng = np.random.default_rng(2)
    n = 150
    batch = dict(ev_off=np.array([0, n], np.uint64), mean=rng.normal(60, 1, n).astype(np.float32),
                 stdv=np.full(n, 0.8, np.float32), start=(np.arange(n) * 0.02).astype(np.float32))
    out = _check_batch(ctx, port, table, mid, batch, None, None)
    assert out["states"][-1] == 0  # lowest final state on a full tie


def test_zero_stdv_event_is_fixed_up(ctx, port, models):
    """Event::update_logs turns stdv == 0 into 0.01 (Event.hpp:39-42)."""
    table = models[R73T]["table"]
    mid = ctx.register_model(table, 0)
    batch = synth.make_batch(21, table, [120])
    batch["stdv"][[0, 17, 119]] = 0.0
    _check_batch(ctx, port, table, mid, batch, None, None)


def test_long_read_traceback_blocks(ctx, port, models):
    """12k events: the blocked speculative traceback uses > 100 blocks."""
    table = models[R73T]["table"]
    mid = ctx.register_model(table, 0)
    batch = synth.make_batch(31, table, [12000, 700])
    _check_batch(ctx, port, table, mid, batch, None, None)


_LONG_ORACLE = {}


def test_read_of_100k_events(ctx, port, models, vit_mode):
    """BASELINE.json configs[4] names reads of 100k+ events: one 102,400-event read next to a short one, bit-identical
    to the oracle through either kernel (the reference's matrix for this read is 3.3 GB; here 1.6 GB of alpha columns
    or 400 MB of backpointers)."""
    table = models[R73T]["table"]
    mid = ctx.register_model(table, 0)
    batch = synth.make_batch_uniform(77, table, 0, 0, lengths=[102400, 900])
    out = ctx.viterbi(batch["ev_off"], batch["mean"], batch["stdv"], batch["start"], mid)
    if "exp" not in _LONG_ORACLE:   # ~30 s of CPU: computed once for both kernel modes
        _LONG_ORACLE["exp"] = port.viterbi_batch(table, batch["ev_off"], batch["mean"], batch["stdv"], batch["start"],
                                                 api._pm_array(None, 2), api._st_array(None, 2), n_threads=2)
    exp = _LONG_ORACLE["exp"]
    assert np.array_equal(_bits(out["path_logprob"]), _bits(exp["path_prob"]))
    assert np.array_equal(out["states"].astype(np.uint32), exp["states"])
    assert np.array_equal(out["moves"].astype(np.int32), exp["moves"])


def test_per_job_pointer_api_and_base_seq(ctx, port, models):
    table = models[R73C2]["table"]
    mid = ctx.register_model(table, 1)
    rng = np.random.default_rng(8)
    jobs = []
    for k in range(5):
        pm = tuple(synth.random_params(rng, 1)[0])
        rd = synth.make_read(rng, table, int(rng.integers(30, 400)), pm)
        jobs.append(dict(mean=rd["mean"], stdv=rd["stdv"], start=rd["start"], model_id=mid, pm=pm, st=(0.12, 0.25)))
    outs = ctx.viterbi_jobs(jobs)
    for jb, o in zip(jobs, outs):
        e = port.viterbi(table, np.array(jb["pm"], np.float32), 0.12, 0.25, jb["mean"], jb["stdv"], jb["start"])
        assert _bits(o["path_logprob"]) == _bits(e["path_prob"])
        assert np.array_equal(o["states"], e["states"])
        assert o["bases"] == e["bases"]
        assert api.base_seq(o["states"], o["moves"]) == e["bases"]


def test_empty_job_is_an_error(ctx, models):
    mid = ctx.register_model(models[R73T]["table"], 0)
    with pytest.raises(api.NanocallError) as ei:
        ctx.viterbi(np.array([0, 0], np.uint64), np.zeros(1, np.float32), np.ones(1, np.float32),
                    np.zeros(1, np.float32), mid)
    assert ei.value.code == -1


def test_device_resident_path_matches_host_path(ctx, models):
    """Device-resident events, no log_stdv supplied: the kernel derives it with the glibc-compatible logf."""
    import torch
    table = models[R73T]["table"]
    mid = ctx.register_model(table, 0)
    batch = synth.make_batch(41, table, [300, 900, 150])
    batch["stdv"][5] = 0.0
    host = ctx.viterbi(batch["ev_off"], batch["mean"], batch["stdv"], batch["start"], mid)
    dev = torch.device("cuda:0")
    d = {k: torch.from_numpy(batch[k]).to(dev) for k in ("mean", "stdv", "start")}
    total = int(batch["ev_off"][-1])
    d_states = torch.zeros(total, dtype=torch.int16, device=dev)
    d_moves = torch.zeros(total, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    path = ctx.viterbi_device(batch["ev_off"], d["mean"].data_ptr(), d["stdv"].data_ptr(), d["start"].data_ptr(),
                              None, mid, d_states=d_states.data_ptr(), d_moves=d_moves.data_ptr())
    assert np.array_equal(_bits(path), _bits(host["path_logprob"]))
    assert np.array_equal(d_states.cpu().numpy().view(np.uint16), host["states"])
    assert np.array_equal(d_moves.cpu().numpy(), host["moves"])


def test_host_supplied_log_stdv_equals_device_logf(ctx, port, models):
    """log_stdv computed by libm on the host and passed in gives the same bits as the device's own logf."""
    table = models[R73T]["table"]
    mid = ctx.register_model(table, 0)
    batch = synth.make_batch(43, table, [700])
    a = ctx.viterbi(batch["ev_off"], batch["mean"], batch["stdv"], batch["start"], mid)
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.logf.restype = ctypes.c_float
    libm.logf.argtypes = [ctypes.c_float]
    lstd = np.array([libm.logf(float(v)) for v in batch["stdv"]], np.float32)
    b = ctx.viterbi(batch["ev_off"], batch["mean"], batch["stdv"], batch["start"], mid, log_stdv=lstd)
    assert np.array_equal(_bits(a["path_logprob"]), _bits(b["path_logprob"]))
    assert np.array_equal(a["states"], b["states"])


def test_mixed_dispatch_small_pool(port, models, vit_mode):
    """A pool too small for the long job's alpha columns: that job takes the backpointer kernel, the short ones the
    alpha kernel, in the same call (two launches); outputs identical to the oracle either way."""
    if vit_mode != "alpha":
        pytest.skip("dispatch test runs once")
    table = models[R73T]["table"]
    lengths = [3000, 40, 900, 2500, 200, 100]
    # a 32 MiB pool holds 2048 alpha columns: the 3000- and 2500-event jobs take the backpointer kernel (2 slabs of
    # 12.3 MB), which leaves 548 columns, so the 900-event job joins them; 200, 100 and 40 take the alpha kernel
    c = api.Context(0, bp_pool_bytes=32 << 20)
    try:
        mid = c.register_model(table, 0)
        batch = synth.make_batch(51, table, lengths)
        _check_batch(c, port, table, mid, batch, None, None)
        assert c.last_launches() == 2
    finally:
        c.close()


def test_column_allocator_under_pressure(port, models, vit_mode):
    """Alpha kernel with a pool that holds only ~6 of the 40 jobs at a time: forward CTAs wait for columns, the
    traceback service releases and coalesces extents of different sizes; same bits as the oracle."""
    if vit_mode != "alpha":
        pytest.skip("allocator test runs once")
    table = models[R73T]["table"]
    lengths = [1500, 90, 700, 1100, 33, 400, 1300, 250, 999, 64] * 4
    c = api.Context(0, bp_pool_bytes=100 << 20)   # 6400 columns
    try:
        mid = c.register_model(table, 0)
        batch = synth.make_batch(57, table, lengths)
        _check_batch(c, port, table, mid, batch, None, None)
        assert c.last_launches() == 1
    finally:
        c.close()


def test_many_short_jobs_with_a_full_pool(models, vit_mode, monkeypatch):
    """2000 jobs of 300-1400 events and a pool of 4 GiB: every forward CTA may hold five jobs, more than the pool has
    columns for, so most forward CTAs wait for columns while the 48 service warps release them.  (With a
    compare-and-swap lock on the extent list the waiting warps starved the service warps of the lock for minutes on
    some GPUs; the list now has a FIFO lock and waiters that cannot fit do not take it.)  Must finish in well under the
    10 s wait limit, three times in a row, with the same results as the backpointer kernel."""
    if vit_mode != "alpha":
        pytest.skip("runs once")
    monkeypatch.setenv("NC_WAIT_LIMIT_S", "10")
    from nanocall_b200 import _lib as L
    import time
    table = models[R73T]["table"]
    rng = np.random.default_rng(77)
    lengths = [int(v) for v in rng.integers(300, 1429, size=2000)]
    batch = synth.make_batch(78, table, lengths)
    for pool in (4 << 30, 1 << 30):
        c = api.Context(0, bp_pool_bytes=pool)
        try:
            mid = c.register_model(table, 0)
            c.set_viterbi_mode(L.NC_VIT_BACKPOINTER)
            ref = c.viterbi(batch["ev_off"], batch["mean"], batch["stdv"], batch["start"], mid)
            c.set_viterbi_mode(L.NC_VIT_AUTO)
            for _ in range(3):
                t0 = time.time()
                out = c.viterbi(batch["ev_off"], batch["mean"], batch["stdv"], batch["start"], mid)
                assert time.time() - t0 < 5.0
                assert c.last_launches() == 1
                assert np.array_equal(_bits(out["path_logprob"]), _bits(ref["path_logprob"]))
                assert np.array_equal(out["states"], ref["states"]) and np.array_equal(out["moves"], ref["moves"])
        finally:
            c.close()


def test_few_long_jobs_take_the_cluster_kernel(port, models, vit_mode, monkeypatch):
    """NC_VIT_CLUSTER=1, at most one job per two SMs, each of at least 2000 events: every job is decoded by a cluster of
    two CTAs (half of the states each, class candidates exchanged through distributed shared memory).  Same bits as
    the oracle and as the one-CTA-per-job kernel, odd and even lengths, one job and several.  (The kernel is off by
    default: it is slower than one CTA per job, see nc_viterbi_alpha.cu; the latency of both is printed.)"""
    if vit_mode != "alpha":
        pytest.skip("runs once")
    import time
    table = models[R73T]["table"]
    for lengths in ([2500, 7001, 3000, 2048, 12345], [4097], [2000] * 74):
        batch = synth.make_batch(91 + len(lengths), table, lengths)
        res = {}
        for flag in ("1", "0"):
            monkeypatch.setenv("NC_VIT_CLUSTER", flag)
            c = api.Context(0, bp_pool_bytes=4 << 30)
            try:
                mid = c.register_model(table, 0)
                if flag == "1" and len(lengths) < 10:
                    _check_batch(c, port, table, mid, batch, None, None)
                res[flag] = c.viterbi(batch["ev_off"], batch["mean"], batch["stdv"], batch["start"], mid)
                assert c.last_launches() == 1
            finally:
                c.close()
        assert np.array_equal(_bits(res["1"]["path_logprob"]), _bits(res["0"]["path_logprob"]))
        assert np.array_equal(res["1"]["states"], res["0"]["states"]) and np.array_equal(res["1"]["moves"], res["0"]["moves"])
    # one read of 60k events: latency with and without the split (printed; the assertion is only on the results)
    batch = synth.make_batch(97, table, [60000])
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("NC_VIT_CLUSTER", flag)
        c = api.Context(0, bp_pool_bytes=4 << 30)
        try:
            mid = c.register_model(table, 0)
            c.viterbi(batch["ev_off"], batch["mean"], batch["stdv"], batch["start"], mid)
            out[flag] = c.viterbi(batch["ev_off"], batch["mean"], batch["stdv"], batch["start"], mid)
            print(f"60k-event read, NC_VIT_CLUSTER={flag}: kernel {c.last_kernel_ms():.2f} ms")
        finally:
            c.close()
    assert np.array_equal(out["1"]["states"], out["0"]["states"])


def test_streamed_event_upload(port, models, vit_mode, monkeypatch):
    """Host-memory calls above a size threshold start the kernels first and stream the events behind them in chunks
    (jobs wait for their events to land).  Forced here on a small batch with 4096-event chunks: same bits."""
    monkeypatch.setenv("NC_STREAM_IN_MIN_EVENTS", "0")
    monkeypatch.setenv("NC_STREAM_IN_CHUNK", "4096")
    table = models[R73T]["table"]
    c = api.Context(0, bp_pool_bytes=2 << 30)
    try:
        from nanocall_b200 import _lib as L
        c.set_viterbi_mode(L.NC_VIT_BACKPOINTER if vit_mode == "backpointer" else L.NC_VIT_AUTO)
        mid = c.register_model(table, 0)
        batch = synth.make_batch(53, table, [5000, 33, 1200, 7000, 250, 4097, 640] * 3)
        # the streamed path needs PINNED host memory (pageable calls take the plain copy path)
        import torch
        keep = {k: torch.from_numpy(batch[k]).pin_memory() for k in ("mean", "stdv", "start")}
        for k, v in keep.items():
            batch[k] = v.numpy()
        _check_batch(c, port, table, mid, batch, None, None)
        # and the same call from pageable memory (plain copies): same bits
        for k in keep:
            batch[k] = batch[k].copy()
        _check_batch(c, port, table, mid, batch, None, None)
    finally:
        c.close()


def test_path_probability_only_needs_no_scratch(port, models, vit_mode):
    """states=NULL, moves=NULL (candidate ranking): no columns are stored, so even a tiny pool serves long jobs."""
    if vit_mode != "alpha":
        pytest.skip("runs once")
    table = models[R73T]["table"]
    c = api.Context(0, bp_pool_bytes=1 << 20)
    try:
        mid = c.register_model(table, 0)
        batch = synth.make_batch(52, table, [5000, 300])
        out = c.viterbi(batch["ev_off"], batch["mean"], batch["stdv"], batch["start"], mid,
                        want_states=False, want_moves=False)
        exp = port.viterbi_batch(table, batch["ev_off"], batch["mean"], batch["stdv"], batch["start"],
                                 api._pm_array(None, 2), api._st_array(None, 2), n_threads=8)
        assert np.array_equal(_bits(out["path_logprob"]), _bits(exp["path_prob"]))
    finally:
        c.close()


def test_two_contexts_one_thread_and_concurrent_training(port, models, vit_mode):
    """Two contexts driven from one host thread (every entry point selects its own device), and the alpha-column
    kernel's persistent grid running while another context's Forward/Backward kernels occupy SMs: the grid is
    launched cooperatively, so it waits for the SMs it needs instead of deadlocking on half-resident CTAs."""
    if vit_mode != "alpha":
        pytest.skip("runs once")
    import threading
    table = models[R73T]["table"]
    a, b = api.Context(0, bp_pool_bytes=4 << 30), api.Context(0, bp_pool_bytes=4 << 30)
    try:
        ma, mb = a.register_model(table, 0), b.register_model(table, 0)
        batch = synth.make_batch(61, table, [3000, 500, 1800, 2500] * 40)
        rng = np.random.default_rng(5)
        groups = []
        for k in range(64):
            rd = synth.make_read(rng, table, 200)
            seqs = [(0, rd["mean"][:100], rd["stdv"][:100], rd["start"][:100]), (0, rd["mean"][100:], rd["stdv"][100:], rd["start"][100:])]
            groups.append(dict(seqs=seqs, model_id=(mb, mb), pm=(1, 0, 0, 1, 1, 1), st=(0.1, 0.3, 0.1, 0.3)))
        # interleaved calls from ONE thread
        first = b.train_round_batch(groups[:4])
        _check_batch(a, port, table, ma, batch, None, None)
        again = b.train_round_batch(groups[:4])
        assert all(np.array_equal(x["pm"].view(np.uint32), y["pm"].view(np.uint32)) for x, y in zip(first, again))
        # training on context b in a second thread while context a decodes
        stop = threading.Event()
        errs = []

        def train_loop():
            try:
                while not stop.is_set():
                    b.train_round_batch(groups)
            except Exception as e:   # noqa: BLE001
                errs.append(e)
        th = threading.Thread(target=train_loop)
        th.start()
        try:
            for _ in range(3):
                _check_batch(a, port, table, ma, batch, None, None)
        finally:
            stop.set()
            th.join()
        assert not errs, errs
    finally:
        a.close()
        b.close()


def test_stalled_grid_is_an_error_not_a_hang(models, vit_mode, monkeypatch):
    """A pool too small for the first wave of jobs plus a wait limit of a few milliseconds: forward CTAs cannot get
    columns in time, the grid drains and the call fails with NC_ERR_STATE (or succeeds if the hardware was quick) --
    it never hangs."""
    if vit_mode != "alpha":
        pytest.skip("runs once")
    monkeypatch.setenv("NC_WAIT_LIMIT_S", "0.002")
    from nanocall_b200 import _lib as L
    table = models[R73T]["table"]
    c = api.Context(0, bp_pool_bytes=200 << 20)   # 12800 columns for ~300 jobs of 6000 events
    try:
        mid = c.register_model(table, 0)
        batch = synth.make_batch(71, table, [6000] * 300)
        try:
            c.viterbi(batch["ev_off"], batch["mean"], batch["stdv"], batch["start"], mid)
        except api.NanocallError as e:
            assert e.code == L.NC_ERR_STATE and "alpha-column kernel stopped" in str(e)
    finally:
        c.close()
