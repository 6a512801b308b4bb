"""ctypes bindings for the TEST-ONLY oracle libraries.

  port()  -> oracle/libnc_oracle.so   plain-C restatement (always buildable; travels)
  ref()   -> oracle/_ref/libncref.so  the reference's own headers compiled in place (built only
                                      where /root/reference exists; the prebuilt .so travels)
Nothing under nanocall_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")
S = 4096

_f = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def build():
    subprocess.run(["make", "-s", "-f", os.path.join(ORACLE, "Makefile"), "oracle", "ref"], check=True)


class _Common:
    """Shared call shapes: both libraries export the same signatures under different prefixes."""
    prefix = ""

    def __init__(self, lib):
        self.lib = lib

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def viterbi(self, table, pm, p_stay, p_skip, mean, stdv, start, dump=False):
        n = mean.size
        table = np.ascontiguousarray(table, np.float32)
        pm = np.ascontiguousarray(pm, np.float32)
        pp = C.c_float()
        states = np.zeros(n, np.uint32)
        moves = np.zeros(n, np.int32)
        cap = 6 * n + 8
        bases = C.create_string_buffer(cap)
        nb = C.c_uint32()
        alpha = np.zeros((n, S), np.float32) if dump else None
        beta = np.zeros((n, S), np.uint32) if dump else None
        rc = self.fn("viterbi")(_p(table), _p(pm), C.c_float(p_stay), C.c_float(p_skip), C.c_uint32(n),
                                _p(mean), _p(stdv), _p(start), C.byref(pp), _p(states), _p(moves),
                                bases, C.c_uint32(cap), C.byref(nb), _p(alpha), _p(beta))
        assert rc == 0, rc
        out = dict(path_prob=np.float32(pp.value), states=states, moves=moves,
                   bases=bases.raw[:nb.value].decode())
        if dump:
            out["alpha"], out["beta"] = alpha, beta
        return out

    def viterbi_batch(self, table, ev_off, mean, stdv, start, pm, st, n_threads=1, want_paths=True):
        n_jobs = ev_off.size - 1
        table = np.ascontiguousarray(table, np.float32)
        pm = np.ascontiguousarray(pm, np.float32).reshape(n_jobs, 6)
        st = np.ascontiguousarray(st, np.float32).reshape(n_jobs, 2)
        ev_off = np.ascontiguousarray(ev_off, np.uint64)
        pp = np.zeros(n_jobs, np.float32)
        states = np.zeros(mean.size, np.uint32) if want_paths else None
        moves = np.zeros(mean.size, np.int32) if want_paths else None
        rc = self.fn("viterbi_batch")(_p(table), C.c_uint32(n_jobs), _p(ev_off), _p(mean), _p(stdv), _p(start),
                                      _p(pm), _p(st), C.c_uint32(n_threads), _p(pp), _p(states), _p(moves))
        assert rc == 0, rc
        return dict(path_prob=pp, states=states, moves=moves)

    def fwbw(self, table, pm, p_stay, p_skip, mean, stdv, start):
        n = mean.size
        table = np.ascontiguousarray(table, np.float32)
        pm = np.ascontiguousarray(pm, np.float32)
        alpha = np.zeros((n, S), np.float32)
        beta = np.zeros((n, S), np.float32)
        lz = C.c_float()
        rc = self.fn("fwbw")(_p(table), _p(pm), C.c_float(p_stay), C.c_float(p_skip), C.c_uint32(n),
                             _p(mean), _p(stdv), _p(start), _p(alpha), _p(beta), C.byref(lz))
        assert rc == 0, rc
        return dict(alpha=alpha, beta=beta, log_pr_data=np.float32(lz.value))

    def mean_stdv(self, x):
        x = np.ascontiguousarray(x, np.float32)
        m, s = C.c_float(), C.c_float()
        self.fn("mean_stdv")(C.c_uint32(x.size), _p(x), C.byref(m), C.byref(s))
        return np.float32(m.value), np.float32(s.value)

    def flogsum(self, a, b):
        f = self.fn("flogsum")
        f.restype = C.c_float
        return np.float32(f(C.c_float(a), C.c_float(b)))


class Port(_Common):
    prefix = "nco_"

    def __init__(self):
        path = os.path.join(ORACLE, "libnc_oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-f", os.path.join(ORACLE, "Makefile"), "oracle"], check=True)
        super().__init__(C.CDLL(path))
        self.lib.nco_flogsum_table.restype = C.POINTER(C.c_float)
        self.lib.nco_trans_prob.restype = C.c_float

    def flogsum_table(self):
        return np.ctypeslib.as_array(self.lib.nco_flogsum_table(), shape=(16000,)).copy()

    def st_train_kmers(self):
        out = np.zeros(S, np.uint32)
        n = self.lib.nco_st_train_kmers(_p(out))
        return out[:n]

    def trans_mask(self, i, j):
        return self.lib.nco_trans_mask(C.c_uint(i), C.c_uint(j))

    def transitions(self, p_stay, p_skip):
        class T(C.Structure):
            _fields_ = [("from_cnt", C.c_uint32 * S), ("to_cnt", C.c_uint32 * S),
                        ("from_idx", C.c_uint32 * (S * 21)), ("to_idx", C.c_uint32 * (S * 21)),
                        ("from_lp", C.c_float * (S * 21)), ("to_lp", C.c_float * (S * 21))]
        t = T()
        self.lib.nco_transitions(C.c_float(p_stay), C.c_float(p_skip), C.byref(t))
        g = lambda a, dt, shape: np.frombuffer(a, dtype=dt).reshape(shape).copy()
        return dict(from_cnt=g(t.from_cnt, np.uint32, (S,)), to_cnt=g(t.to_cnt, np.uint32, (S,)),
                    from_idx=g(t.from_idx, np.uint32, (S, 21)), to_idx=g(t.to_idx, np.uint32, (S, 21)),
                    from_lp=g(t.from_lp, np.float32, (S, 21)), to_lp=g(t.to_lp, np.float32, (S, 21)))

    def scaled_model(self, table, pm=None):
        class M(C.Structure):
            _fields_ = [(k, C.c_float * S) for k in
                        ("level_mean", "level_stdv", "sd_mean", "sd_stdv", "sd_lambda",
                         "log_level_stdv", "log_sd_lambda")] + [("mean", C.c_float), ("stdv", C.c_float)]
        table = np.ascontiguousarray(table, np.float32)
        m0, m1 = M(), M()
        self.lib.nco_model_prepare(_p(table), C.byref(m0))
        m = m0
        if pm is not None:
            pm = np.ascontiguousarray(pm, np.float32)
            self.lib.nco_model_scale(C.byref(m0), _p(pm), C.byref(m1))
            m = m1
        out = {k: np.frombuffer(getattr(m, k), np.float32).copy() for k, _ in M._fields_[:7]}
        out["mean"], out["stdv"] = np.float32(m.mean), np.float32(m.stdv)
        return out

    def train_one_round(self, seqs, table0, table1, pm, st, train_scaling=True, train_transitions=True,
                        train_drift=True):
        return _train(self.lib.nco_train_one_round, seqs, table0, table1, pm, st, train_scaling,
                      train_transitions, train_drift)


class Ref(_Common):
    prefix = "ncref_"

    def __init__(self, default_p_stay=0.1, default_p_skip=0.3, train_drift=1):
        path = os.path.join(ORACLE, "_ref", "libncref.so")
        if not os.path.exists(path):
            if os.path.isdir("/root/reference/src/nanocall"):
                subprocess.run(["make", "-s", "-f", os.path.join(ORACLE, "Makefile"), "ref"], check=True)
            else:
                raise FileNotFoundError(path)
        super().__init__(C.CDLL(path))
        self.train_drift = train_drift
        self.n_train_kmers = self.lib.ncref_init(C.c_float(default_p_stay), C.c_float(default_p_skip),
                                                 C.c_int(train_drift))

    def flogsum_table(self):
        out = np.zeros(16000, np.float32)
        self.lib.ncref_flogsum_table(_p(out))
        return out

    def st_train_kmers(self):
        out = np.zeros(S, np.uint32)
        n = self.lib.ncref_st_train_kmers(_p(out))
        return out[:n]

    def transitions(self, p_stay, p_skip):
        d = dict(from_cnt=np.zeros(S, np.uint32), from_idx=np.zeros((S, 21), np.uint32),
                 from_lp=np.zeros((S, 21), np.float32), to_cnt=np.zeros(S, np.uint32),
                 to_idx=np.zeros((S, 21), np.uint32), to_lp=np.zeros((S, 21), np.float32))
        rc = self.lib.ncref_transitions(C.c_float(p_skip), C.c_float(p_stay), _p(d["from_cnt"]), _p(d["from_idx"]),
                                        _p(d["from_lp"]), _p(d["to_cnt"]), _p(d["to_idx"]), _p(d["to_lp"]))
        assert rc == 0
        return d

    def scaled_model(self, table, pm=None):
        table = np.ascontiguousarray(table, np.float32)
        out = np.zeros((S, 8), np.float32)
        stats = np.zeros(2, np.float32)
        pm = None if pm is None else np.ascontiguousarray(pm, np.float32)
        self.lib.ncref_scaled_model(_p(table), _p(pm), _p(out), _p(stats))
        keys = ("level_mean", "level_stdv", "log_level_stdv", "sd_mean", "sd_lambda", "log_sd_lambda", "sd_stdv")
        d = {k: out[:, i].copy() for i, k in enumerate(keys)}
        d["mean"], d["stdv"] = stats[0], stats[1]
        return d

    def emissions(self, table, pm, mean, stdv, start):
        table = np.ascontiguousarray(table, np.float32)
        pm = np.ascontiguousarray(pm, np.float32)
        out = np.zeros((mean.size, S), np.float32)
        self.lib.ncref_emissions(_p(table), _p(pm), C.c_uint32(mean.size), _p(mean), _p(stdv), _p(start), _p(out))
        return out

    def train_one_round(self, seqs, table0, table1, pm, st, train_scaling=True, train_transitions=True,
                        train_drift=None):
        assert train_drift is None or int(train_drift) == int(self.train_drift), "set train_drift in Ref()"
        return _train(self.lib.ncref_train_one_round, seqs, table0, table1, pm, st, train_scaling,
                      train_transitions, None)


def _train(fn, seqs, table0, table1, pm, st, train_scaling, train_transitions, train_drift):
    """seqs: list of (strand, mean, stdv, start)."""
    lens = np.array([s[1].size for s in seqs], np.uint32)
    strands = np.array([s[0] for s in seqs], np.uint32)
    mean = np.ascontiguousarray(np.concatenate([s[1] for s in seqs]), np.float32)
    stdv = np.ascontiguousarray(np.concatenate([s[2] for s in seqs]), np.float32)
    start = np.ascontiguousarray(np.concatenate([s[3] for s in seqs]), np.float32)
    table0 = np.ascontiguousarray(table0, np.float32)
    table1 = np.ascontiguousarray(table1, np.float32)
    pm = np.ascontiguousarray(pm, np.float32)
    st = np.ascontiguousarray(st, np.float32)
    new_pm = np.zeros(6, np.float32)
    new_st = np.zeros(4, np.float32)
    fit = C.c_float()
    done = C.c_int()
    args = [C.c_uint32(len(seqs)), _p(lens), _p(strands), _p(mean), _p(stdv), _p(start), _p(table0), _p(table1),
            _p(pm), _p(st), C.c_int(int(train_scaling)), C.c_int(int(train_transitions))]
    if train_drift is not None:
        args.append(C.c_int(int(train_drift)))
    args += [_p(new_pm), _p(new_st), C.byref(fit), C.byref(done)]
    rc = fn(*args)
    assert rc == 0, rc
    return dict(pm=new_pm, st=new_st, fit=np.float32(fit.value), done=bool(done.value))


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        _port = Port()
    return _port


def have_ref():
    return os.path.exists(os.path.join(ORACLE, "_ref", "libncref.so")) or os.path.isdir("/root/reference/src/nanocall")


def ref():
    global _ref
    if _ref is None:
        _ref = Ref()
    return _ref
