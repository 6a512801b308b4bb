"""Synthetic R7.3-style event tables, used by tests, bench.py and the
golden-vector generator.  Pure numpy, deterministic for a given seed.

Per read: a uniform random base stream; event i sits on the 6-mer starting at base pos_i, where
pos advances by 0 / 1 / 2 bases with probability p_stay / 1-p_stay-p_skip / p_skip -- the
stay/step/skip walk behind nanocall's transition model (State_Transitions.hpp:125-144,
CLI defaults --pr-stay .1 --pr-skip .3, nanocall.cpp:84-85).  Per event:
mean ~ N(scale*mu_s + shift + drift*t, (var*sigma_s)^2), stdv ~ InvGauss(scale_sd*eta_s,
var_sd*lambda_s) with lambda = eta^3 / sd_stdv^2 (Pore_Model.hpp:112), clamped to (0, 4]
(the reader's filter, Fast5_Summary.hpp:734-745), length ~ max(0.002, Exp(0.02 s)),
start = running sum of lengths in seconds from strand start.
"""
import numpy as np

N_STATES = 4096
IDENTITY_PARAMS = (1.0, 0.0, 0.0, 1.0, 1.0, 1.0)  # scale, shift, drift, var, scale_sd, var_sd


def kmer_walk(rng, n_events, p_stay=0.1, p_skip=0.3):
    """State index per event for one read (A=0,C=1,G=2,T=3, first base in the high bits)."""
    u = rng.random(n_events)
    move = np.where(u < p_stay, 0, np.where(u < 1.0 - p_skip, 1, 2)).astype(np.int64)
    move[0] = 0
    pos = np.cumsum(move)
    bases = rng.integers(0, 4, size=int(pos[-1]) + 6, dtype=np.int64)
    kmer = np.zeros(bases.size - 5, dtype=np.int64)
    for k in range(6):
        kmer = (kmer << 2) | bases[k:k + kmer.size]
    return kmer[pos].astype(np.uint16)


def make_read(rng, table, n_events, params=IDENTITY_PARAMS, p_stay=0.1, p_skip=0.3):
    """-> dict(mean, stdv, start: float32[n_events], states: uint16[n_events] ground truth)."""
    scale, shift, drift, var, scale_sd, var_sd = params
    states = kmer_walk(rng, n_events, p_stay, p_skip)
    t = table[states].astype(np.float64)
    mu, sigma, eta, sd_stdv = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
    lam = eta ** 3 / sd_stdv ** 2
    length = np.maximum(0.002, rng.exponential(0.02, n_events))
    start = np.concatenate([[0.0], np.cumsum(length)[:-1]])
    mean = rng.normal(scale * mu + shift + drift * start, var * sigma)
    stdv = rng.wald(scale_sd * eta, var_sd * lam)
    stdv = np.clip(stdv, 1e-3, 4.0)
    return {
        "mean": mean.astype(np.float32),
        "stdv": stdv.astype(np.float32),
        "start": start.astype(np.float32),
        "states": states,
    }


def make_batch(seed, table, lengths, params=None, p_stay=0.1, p_skip=0.3):
    """Packed batch: concatenated float32 mean/stdv/start + uint64 offsets[n_reads+1].
    params: None (identity for every read) or array[n_reads, 6]."""
    rng = np.random.default_rng(seed)
    lengths = np.asarray(lengths, dtype=np.int64)
    off = np.zeros(lengths.size + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lengths)
    total = int(off[-1])
    mean = np.empty(total, np.float32)
    stdv = np.empty(total, np.float32)
    start = np.empty(total, np.float32)
    truth = np.empty(total, np.uint16)
    for r, n in enumerate(lengths):
        p = IDENTITY_PARAMS if params is None else tuple(float(v) for v in params[r])
        rd = make_read(rng, table, int(n), p, p_stay, p_skip)
        a, b = int(off[r]), int(off[r + 1])
        mean[a:b], stdv[a:b], start[a:b], truth[a:b] = rd["mean"], rd["stdv"], rd["start"], rd["states"]
    return {"ev_off": off, "mean": mean, "stdv": stdv, "start": start, "truth": truth}


def random_params(rng, n):
    """Ground-truth scaling for the training configs."""
    p = np.empty((n, 6), np.float32)
    p[:, 0] = rng.uniform(0.9, 1.1, n)
    p[:, 1] = rng.uniform(-5, 5, n)
    p[:, 2] = rng.uniform(-0.005, 0.005, n)
    p[:, 3] = rng.uniform(0.9, 1.3, n)
    p[:, 4] = rng.uniform(0.8, 1.2, n)
    p[:, 5] = rng.uniform(0.8, 1.5, n)
    return p


def mixture_lengths(seed, n_reads):
    """Read lengths of BASELINE.json configs[4]: 90 % LogNormal(median 5000, sigma 0.5)
    clipped to [500, 20000), 9 % uniform 20k-50k, 1 % uniform 100k-150k events."""
    rng = np.random.default_rng(seed)
    u = rng.random(n_reads)
    short = np.clip(rng.lognormal(np.log(5000.0), 0.5, n_reads), 500, 19999)
    mid = rng.uniform(20000, 50000, n_reads)
    lng = rng.uniform(100000, 150000, n_reads)
    return np.where(u < 0.90, short, np.where(u < 0.99, mid, lng)).astype(np.int64)


def make_batch_uniform(seed, table, n_reads, n_events, p_stay=0.1, p_skip=0.3, lengths=None):
    """Vectorised generator for n_reads reads of exactly n_events events (or of the given lengths), identity
    scaling (bench workloads: 10k x 10k, the length mixture).  One long stay/step/skip walk over a random base
    stream is cut into reads; `start` restarts at 0 in every read."""
    rng = np.random.default_rng(seed)
    if lengths is not None:
        lengths = np.asarray(lengths, np.int64)
        n_reads = lengths.size
        total = int(lengths.sum())
    else:
        total = n_reads * n_events
    u = rng.random(total, dtype=np.float32)
    move = np.where(u < p_stay, 0, np.where(u < 1.0 - p_skip, 1, 2)).astype(np.int64)
    del u
    move[0] = 0
    pos = np.cumsum(move)
    del move
    bases = rng.integers(0, 4, size=int(pos[-1]) + 6, dtype=np.uint16)
    kmer = np.zeros(bases.size - 5, dtype=np.uint16)
    for k in range(6):
        kmer = (kmer << 2) | bases[k:k + kmer.size]
    states = kmer[pos]
    del pos, bases, kmer
    t = table[states]
    mu, sigma = t[:, 0], t[:, 1]
    eta = t[:, 2].astype(np.float64)
    lam = eta ** 3 / t[:, 3].astype(np.float64) ** 2
    mean = (mu + sigma * rng.standard_normal(total, dtype=np.float32)).astype(np.float32)
    stdv = np.clip(rng.wald(eta, lam), 1e-3, 4.0).astype(np.float32)
    del eta, lam, t
    if lengths is not None:
        length = np.maximum(0.002, rng.exponential(0.02, total))
        off = np.zeros(n_reads + 1, np.uint64)
        off[1:] = np.cumsum(lengths)
        run = np.cumsum(length) - length                      # start of each event on one global clock
        first = np.repeat(run[off[:-1].astype(np.int64)], lengths)
        return {"ev_off": off, "mean": mean, "stdv": stdv, "start": (run - first).astype(np.float32), "truth": states}
    length = np.maximum(0.002, rng.exponential(0.02, total)).reshape(n_reads, n_events)
    start = np.cumsum(length, axis=1) - length
    off = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(n_events))
    return {"ev_off": off, "mean": mean, "stdv": stdv, "start": start.astype(np.float32).reshape(-1),
            "truth": states}


def make_raw_2d_read(rng, tables, nt, nc, params, sampling_rate=5000.0, lead=60, tail=60, hairpin=8, comp=0,
                     hairpin_level=115.0):
    """One raw 2D read as a fast5 file holds it (structured array: mean, stdv as float64, start, length in samples):
    lead-in, template (nt events), a hairpin island of abasic-level events, complement (nc events), tail.
    tables = (template table, [complement tables]); params = the read's scaling, applied with the time base the
    reference uses when both strands are scaled together (seconds from the template's first event).
    nc == 0: a 1D read without hairpin."""
    from .evio import RAW_DTYPE
    parts = []
    clock = 1000

    def strand(table, n, pm, t0_ref):
        nonlocal clock
        rd = make_read(rng, table, n, IDENTITY_PARAMS)
        length = np.maximum(10, np.rint(np.maximum(0.002, rng.exponential(0.02, n)) * sampling_rate)).astype(np.int64)
        start = clock + np.concatenate([[0], np.cumsum(length)[:-1]])
        t0 = start[0] if t0_ref is None else t0_ref
        t = (start - t0) / sampling_rate
        tb = table[rd["states"]].astype(np.float64)
        scale, shift, drift, var, scale_sd, var_sd = [float(v) for v in pm]
        mean = rng.normal(scale * tb[:, 0] + shift + drift * t, var * tb[:, 1]).astype(np.float32)
        lam = tb[:, 2] ** 3 / tb[:, 3] ** 2
        stdv = np.clip(rng.wald(scale_sd * tb[:, 2], var_sd * lam), 1e-3, 4.0).astype(np.float32)
        ev = np.zeros(n, RAW_DTYPE)
        ev["mean"], ev["stdv"], ev["start"], ev["length"] = mean, stdv, start, length
        clock = int(start[-1] + length[-1])
        return ev, int(t0)

    ttab, ctabs = tables
    ev, _ = strand(ttab, lead, IDENTITY_PARAMS, None)
    parts.append(ev)
    ev, t0 = strand(ttab, nt, params, None)
    parts.append(ev)
    if nc:
        h = np.zeros(hairpin, RAW_DTYPE)
        h["mean"] = (hairpin_level + 2.0 * rng.standard_normal(hairpin)).astype(np.float32)
        h["stdv"] = (1.0 + 0.2 * rng.random(hairpin)).astype(np.float32)
        h["length"] = 100
        h["start"] = clock + 100 * np.arange(hairpin)
        clock += 100 * hairpin
        parts.append(h)
        ev, _ = strand(ctabs[comp], nc, params, t0)
        parts.append(ev)
    ev, _ = strand(ctabs[comp] if nc else ttab, tail, IDENTITY_PARAMS, None)
    parts.append(ev)
    return np.concatenate(parts)
