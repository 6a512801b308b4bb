"""Thin numpy-facing wrapper over the C ABI, used by tests/ and bench.py.

Names follow the reference: a `Context` owns the registered `Pore_Model`s; `viterbi()` is the
batched Viterbi::fill (Viterbi.hpp:44-150 via basecall_strand, nanocall.cpp:645-690);
`forward_backward()` is Forward_Backward::fill (Forward_Backward.hpp:46-135);
`train_one_round()` is Parameter_Trainer::train_one_round (Parameter_Trainer.hpp:541-579).
Every call goes through libnanocall_b200.so; nothing here computes.
"""
import ctypes as C
import numpy as np

from . import _lib as L

N_STATES = 4096
IDENTITY_PM = (1.0, 0.0, 0.0, 1.0, 1.0, 1.0)
DEFAULT_ST = (0.1, 0.3)  # CLI defaults --pr-stay / --pr-skip (nanocall.cpp:84-85)


class NanocallError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"nanocall_b200 error {code}: {msg}")
        self.code = code


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a)  # raw device pointer


def _as(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _pm_array(pm, n):
    a = np.empty((n, 6), np.float32)
    a[:] = np.asarray(IDENTITY_PM if pm is None else pm, np.float32).reshape(-1, 6)
    return a


def _st_array(st, n):
    a = np.empty((n, 2), np.float32)
    a[:] = np.asarray(DEFAULT_ST if st is None else st, np.float32).reshape(-1, 2)
    return a


class Context:
    def __init__(self, device=0, bp_pool_bytes=0):
        self.lib = L.load()
        h = C.c_void_p()
        rc = self.lib.nc_ctx_create(int(device), int(bp_pool_bytes), C.byref(h))
        if rc != L.NC_OK:
            raise NanocallError(rc, self.lib.nc_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.nc_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != L.NC_OK:
            raise NanocallError(rc, self.lib.nc_last_error(self.h).decode())

    @property
    def stream(self):
        return self.lib.nc_ctx_stream(self.h)

    def sync(self):
        self._check(self.lib.nc_ctx_sync(self.h))

    def last_kernel_ms(self):
        return float(self.lib.nc_ctx_last_kernel_ms(self.h))

    def last_launches(self):
        return int(self.lib.nc_ctx_last_launches(self.h))

    def viterbi_stats(self, reset=True):
        """Counters of the alpha-column kernel (see nc_ctx_viterbi_stats)."""
        out = np.zeros(8, np.uint64)
        self._check(self.lib.nc_ctx_viterbi_stats(self.h, out.ctypes.data, int(reset)))
        keys = ("fwd_cycles", "fwd_wait_slab_cycles", "tb_busy_cycles", "tb_wait_cycles", "tb_passes", "tb_lane_steps", "tb_jobs")
        return {k: int(v) for k, v in zip(keys, out)}

    def train_stats(self, reset=True):
        """Device time of the training kernels since the last reset (see nc_ctx_train_stats)."""
        out = np.zeros(8, np.float64)
        self._check(self.lib.nc_ctx_train_stats(self.h, out.ctypes.data, int(reset)))
        keys = ("emission_ms", "fwbw_ms", "pm_stats_ms", "st_stats_ms", "events", "launches", "waves")
        return {k: float(v) for k, v in zip(keys, out)}

    def set_default_transitions(self, p_stay, p_skip, edges_from=None, edges_to=None, logp=None):
        """Custom initial transition table (--trans): edges in file order; None removes it."""
        if edges_from is None:
            self._check(self.lib.nc_ctx_set_default_transitions(self.h, float(p_stay), float(p_skip), 0, None, None, None))
            return
        f, t, lp = _as(edges_from, np.uint16), _as(edges_to, np.uint16), _as(logp, np.float32)
        self._check(self.lib.nc_ctx_set_default_transitions(self.h, float(p_stay), float(p_skip), f.size, _ptr(f), _ptr(t), _ptr(lp)))

    def set_viterbi_mode(self, mode):
        """L.NC_VIT_AUTO (alpha-column kernel where it fits) or L.NC_VIT_BACKPOINTER."""
        self._check(self.lib.nc_ctx_set_viterbi_mode(self.h, int(mode)))

    def device_info(self):
        n, mem, name = C.c_int(), C.c_size_t(), C.create_string_buffer(256)
        self._check(self.lib.nc_ctx_device_info(self.h, C.byref(n), C.byref(mem), name, 256))
        return dict(n_sms=n.value, total_mem=mem.value, name=name.value.decode())

    # ---- Pore_Model::load_from_vector
    def register_model(self, table, strand=2):
        t = _as(table, np.float32)
        assert t.shape == (N_STATES, 4)
        mid = C.c_int()
        self._check(self.lib.nc_model_register(self.h, t.ctypes.data, int(strand), C.byref(mid)))
        return mid.value

    def model_stats(self, model_id):
        m, s = C.c_float(), C.c_float()
        self._check(self.lib.nc_model_stats(self.h, model_id, C.byref(m), C.byref(s)))
        return np.float32(m.value), np.float32(s.value)

    # ---- Viterbi
    def viterbi(self, ev_off, mean, stdv, start, model_id, pm=None, st=None, log_stdv=None,
                want_states=True, want_moves=True):
        """Host-memory packed batch. Returns dict(path_logprob, states, moves)."""
        ev_off = _as(ev_off, np.uint64)
        n = ev_off.size - 1
        mean, stdv, start = _as(mean, np.float32), _as(stdv, np.float32), _as(start, np.float32)
        log_stdv = None if log_stdv is None else _as(log_stdv, np.float32)
        mid = np.empty(n, np.int32)
        mid[:] = model_id
        pm_a, st_a = _pm_array(pm, n), _st_array(st, n)
        path = np.zeros(n, np.float32)
        total = int(ev_off[-1])
        states = np.zeros(total, np.uint16) if (want_states or want_moves) else None
        moves = np.zeros(total, np.uint8) if want_moves else None
        self._check(self.lib.nc_viterbi_packed(
            self.h, n, _ptr(ev_off), _ptr(mean), _ptr(stdv), _ptr(start), _ptr(log_stdv),
            _ptr(mid), _ptr(pm_a), _ptr(st_a), L.NC_MEM_HOST, _ptr(path), _ptr(states), _ptr(moves)))
        return dict(path_logprob=path, states=states, moves=moves)

    def viterbi_device(self, ev_off, d_mean, d_stdv, d_start, d_log_stdv, model_id, pm=None, st=None,
                       d_states=None, d_moves=None):
        """Device-resident events/outputs (raw pointers, e.g. torch.Tensor.data_ptr())."""
        ev_off = _as(ev_off, np.uint64)
        n = ev_off.size - 1
        mid = np.empty(n, np.int32)
        mid[:] = model_id
        pm_a, st_a = _pm_array(pm, n), _st_array(st, n)
        path = np.zeros(n, np.float32)
        self._check(self.lib.nc_viterbi_packed(
            self.h, n, _ptr(ev_off), _ptr(d_mean), _ptr(d_stdv), _ptr(d_start), _ptr(d_log_stdv),
            _ptr(mid), _ptr(pm_a), _ptr(st_a), L.NC_MEM_DEVICE, _ptr(path), _ptr(d_states), _ptr(d_moves)))
        return path

    def viterbi_jobs(self, jobs, want_bases=True):
        """Per-job-pointer form (nc_viterbi_batch). jobs: list of dict(mean, stdv, start, model_id, pm, st)."""
        n = len(jobs)
        J = (L.VitJob * n)()
        O = (L.VitOut * n)()
        keep = []
        for k, jb in enumerate(jobs):
            m, s, t = _as(jb["mean"], np.float32), _as(jb["stdv"], np.float32), _as(jb["start"], np.float32)
            ne = m.size
            st_arr, mv_arr = np.zeros(ne, np.uint16), np.zeros(ne, np.uint8)
            bases = C.create_string_buffer(6 * ne + 8) if want_bases else None
            keep.append((m, s, t, st_arr, mv_arr, bases))
            J[k].mean, J[k].stdv, J[k].start = m.ctypes.data, s.ctypes.data, t.ctypes.data
            J[k].n_events, J[k].model_id = ne, int(jb["model_id"])
            J[k].pm = L.PmParams(*[float(v) for v in jb.get("pm", IDENTITY_PM)])
            J[k].st = L.StParams(*[float(v) for v in jb.get("st", DEFAULT_ST)])
            O[k].states, O[k].moves = st_arr.ctypes.data, mv_arr.ctypes.data
            if want_bases:
                O[k].bases, O[k].bases_cap = C.addressof(bases), 6 * ne + 8
        self._check(self.lib.nc_viterbi_batch(self.h, n, J, O))
        out = []
        for k in range(n):
            _, _, _, st_arr, mv_arr, bases = keep[k]
            out.append(dict(path_logprob=np.float32(O[k].path_logprob), states=st_arr, moves=mv_arr,
                            bases=bases.raw[:O[k].n_bases].decode() if want_bases else None))
        return out

    # ---- Forward_Backward
    def forward_backward(self, model_id, pm, st, mean, stdv, start, want_matrices=True):
        mean, stdv, start = _as(mean, np.float32), _as(stdv, np.float32), _as(start, np.float32)
        n = mean.size
        pmv = L.PmParams(*[float(v) for v in (IDENTITY_PM if pm is None else pm)])
        stv = L.StParams(*[float(v) for v in (DEFAULT_ST if st is None else st)])
        alpha = np.zeros((n, N_STATES), np.float32) if want_matrices else None
        beta = np.zeros((n, N_STATES), np.float32) if want_matrices else None
        lz = C.c_float()
        self._check(self.lib.nc_fwbw(self.h, int(model_id), C.addressof(pmv), C.addressof(stv), n,
                                     _ptr(mean), _ptr(stdv), _ptr(start), _ptr(alpha), _ptr(beta), C.byref(lz)))
        return dict(alpha=alpha, beta=beta, log_pr_data=np.float32(lz.value))

    # ---- Parameter_Trainer
    def train_round_batch(self, groups, train_scaling=True, train_transitions=True, train_drift=True):
        """groups: list of dict(seqs=[(strand, mean, stdv, start)], model_id=(m0, m1), pm=6 floats,
        st=(p_stay0, p_skip0, p_stay1, p_skip1)).  Returns list of dict(pm, st, fit, done)."""
        ng = len(groups)
        seq_off = np.zeros(ng + 1, np.uint32)
        strands, lens, means, stdvs, starts = [], [], [], [], []
        tin = (L.TrainIn * ng)()
        for g, grp in enumerate(groups):
            seq_off[g + 1] = seq_off[g] + len(grp["seqs"])
            for (sd, m, s, t) in grp["seqs"]:
                strands.append(sd)
                lens.append(len(m))
                means.append(_as(m, np.float32)); stdvs.append(_as(s, np.float32)); starts.append(_as(t, np.float32))
            tin[g].model_id[0], tin[g].model_id[1] = int(grp["model_id"][0]), int(grp["model_id"][1])
            tin[g].pm = L.PmParams(*[float(v) for v in grp["pm"]])
            stp = [float(v) for v in grp["st"]]
            tin[g].st[0] = L.StParams(stp[0], stp[1])
            tin[g].st[1] = L.StParams(stp[2], stp[3])
        ev_off = np.zeros(len(lens) + 1, np.uint64)
        ev_off[1:] = np.cumsum(lens)
        mean = np.concatenate(means) if means else np.zeros(0, np.float32)
        stdv = np.concatenate(stdvs) if stdvs else np.zeros(0, np.float32)
        start = np.concatenate(starts) if starts else np.zeros(0, np.float32)
        strand_a = np.asarray(strands, np.uint8)
        opts = L.TrainOpts(int(train_scaling), int(train_transitions), int(train_drift))
        tout = (L.TrainOut * ng)()
        self._check(self.lib.nc_train_round_batch(self.h, ng, _ptr(seq_off), _ptr(ev_off), _ptr(strand_a),
                                                  _ptr(mean), _ptr(stdv), _ptr(start),
                                                  C.addressof(tin), C.addressof(opts), C.addressof(tout)))
        res = []
        for g in range(ng):
            o = tout[g]
            res.append(dict(
                pm=np.array([o.pm.scale, o.pm.shift, o.pm.drift, o.pm.var, o.pm.scale_sd, o.pm.var_sd], np.float32),
                st=np.array([o.st[0].p_stay, o.st[0].p_skip, o.st[1].p_stay, o.st[1].p_skip], np.float32),
                fit=np.float32(o.fit), done=bool(o.done)))
        return res

    def train_one_round(self, seqs, model_id, pm, st, **kw):
        return self.train_round_batch([dict(seqs=seqs, model_id=model_id, pm=pm, st=st)], **kw)[0]


def transition_lut(p_stay, p_skip):
    out = np.zeros(64, np.float32)
    L.load().nc_transition_lut(float(p_stay), float(p_skip), out.ctypes.data)
    return out


def mean_stdv(x):
    x = _as(x, np.float32)
    m, s = C.c_float(), C.c_float()
    L.load().nc_mean_stdv(x.size, x.ctypes.data, C.byref(m), C.byref(s))
    return np.float32(m.value), np.float32(s.value)


def base_seq(states, moves):
    states, moves = _as(states, np.uint16), _as(moves, np.uint8)
    cap = 6 * states.size + 8
    buf = C.create_string_buffer(cap)
    n = L.load().nc_base_seq(states.size, states.ctypes.data, moves.ctypes.data, C.addressof(buf), cap)
    return buf.raw[:n].decode()
