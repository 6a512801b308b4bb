"""TEST INFRASTRUCTURE: the reference's per-read driver logic restated over the oracle's primitives.

The arithmetic (train_one_round, Viterbi, mean_stdv) is the oracle's (= the reference's, bit for bit);
what is restated here is the orchestration that lives in lambdas of nanocall.cpp and cannot be linked:
  initial scaling        Fast5_Summary.hpp:223-278
  candidate lists        nanocall.cpp:300-323
  training slices        nanocall.cpp:327-338
  EM loops + stop rules  nanocall.cpp:356-436 (double strand), :463-551 (single strand)
  model selection        nanocall.cpp:437-459, :552-570
  basecall + ranking     nanocall.cpp:645-690, :692-855
"""
import numpy as np

F = np.float32


class Opts:
    min_ed_events = 10
    scaling_num_events = 200
    scaling_max_rounds = 10
    scaling_min_progress = F(1.0)
    scaling_select_threshold = F(20.0)
    double_strand_scaling = True
    train_scaling = True
    train_transitions = True
    train_drift = True
    pr_stay = F(0.1)
    pr_skip = F(0.3)


def run_read(lib, models, events, opts, train=True):
    """models: dict name -> dict(table, strand); events: [strand0 dict or None, strand1 dict or None].
    Returns dict(pm_params, st_params, fits, preferred, calls={st: dict(model, path_prob, bases, states)})."""
    names = sorted(models)  # std::map order
    stats = {n: lib.scaled_model(models[n]["table"]) for n in names}
    nev = [0 if e is None else e["mean"].size for e in events]
    together = opts.double_strand_scaling and nev[0] >= opts.min_ed_events and nev[1] >= opts.min_ed_events
    dst = np.array([opts.pr_stay, opts.pr_skip], F)
    pm_m, st_m = {}, {}
    ident = np.array([1, 0, 0, 1, 1, 1], F)
    if together:
        r0, r1 = lib.mean_stdv(events[0]["mean"]), lib.mean_stdv(events[1]["mean"])
        for n0 in names:
            if models[n0]["strand"] in (0, 2):
                for n1 in names:
                    if models[n1]["strand"] in (1, 2):
                        pm = ident.copy()
                        pm[0] = F(F(F(r0[1] / stats[n0]["stdv"]) + F(r1[1] / stats[n1]["stdv"])) / F(2))
                        pm[1] = F(F(F(F(r0[0] - F(pm[0] * stats[n0]["mean"])) + r1[0]) - F(pm[0] * stats[n1]["mean"])) / F(2))
                        pm_m[(n0, n1)] = pm
                        st_m[(n0, n1)] = np.concatenate([dst, dst])
    else:
        for st in range(2):
            if nev[st] < opts.min_ed_events:
                continue
            r = lib.mean_stdv(events[st]["mean"])
            for n in names:
                if models[n]["strand"] in (st, 2):
                    key = (n, "") if st == 0 else ("", n)
                    pm = ident.copy()
                    pm[0] = F(r[1] / stats[n]["stdv"])
                    pm[1] = F(r[0] - F(pm[0] * stats[n]["mean"]))
                    pm_m[key] = pm
                    st_m[key] = np.concatenate([dst, dst])
    preferred = {0: "", 1: "", 2: None}
    fits, rounds = {}, {}

    def seqs_of(st):
        e = events[st]
        n = min(opts.scaling_num_events, nev[st])
        h = n // 2
        return [(st, e["mean"][:h], e["stdv"][:h], e["start"][:h]),
                (st, e["mean"][nev[st] - h:], e["stdv"][nev[st] - h:], e["start"][nev[st] - h:])]

    def em(key, seqs, t0, t1, max_rounds):
        crt_pm, crt_st, crt_fit, rnd = pm_m[key], st_m[key], F(-np.inf), 0
        while True:
            old_pm, old_st, old_fit = crt_pm, crt_st, crt_fit
            o = lib.train_one_round(seqs, t0, t1, old_pm, old_st, opts.train_scaling, opts.train_transitions,
                                    train_drift=opts.train_drift)
            crt_pm, crt_st, crt_fit = o["pm"], o["st"], o["fit"]
            if o["done"]:
                break
            if crt_fit < old_fit:
                crt_pm, crt_st, crt_fit = old_pm, old_st, old_fit
                break
            rnd += 1
            if rnd >= max_rounds or (rnd > 1 and crt_fit < F(old_fit + opts.scaling_min_progress)):
                break
        pm_m[key], st_m[key], fits[key], rounds[key] = crt_pm, crt_st, crt_fit, rnd

    def select(keys):
        best = keys[0]
        for k in keys:
            if fits[best] < fits[k]:
                best = k
        if all(k == best or F(fits[k] + opts.scaling_select_threshold) < fits[best] for k in keys):
            return best
        return None

    if train:
        lists = {st: [n for n in names if models[n]["strand"] in (st, 2)] if nev[st] >= opts.min_ed_events else []
                 for st in range(2)}
        if together:
            keys = [(a, b) for a in lists[0] for b in lists[1]]
            for k in keys:
                em(k, seqs_of(0) + seqs_of(1), models[k[0]]["table"], models[k[1]]["table"], 2 * opts.scaling_max_rounds)
            if np.isfinite(opts.scaling_select_threshold):
                preferred[2] = select(sorted(keys))
        else:
            for st in range(2):
                if nev[st] < opts.min_ed_events:
                    continue
                keys = [((n, "") if st == 0 else ("", n)) for n in lists[st]]
                for k in keys:
                    em(k, seqs_of(st), models[k[st]]["table"], models[k[st]]["table"], opts.scaling_max_rounds)
                if np.isfinite(opts.scaling_select_threshold):
                    b = select(sorted(keys))
                    if b is not None:
                        preferred[st] = b[st]

    def basecall(st, key):
        pm, sp = pm_m[key], st_m[key][2 * st:2 * st + 2]
        e = events[st]
        return lib.viterbi(models[key[st]]["table"], pm, float(sp[0]), float(sp[1]), e["mean"], e["stdv"], e["start"])

    calls = {}
    if together:
        sub = [preferred[2]] if preferred[2] is not None else sorted(k for k in pm_m if k[0] and k[1])
        res = []
        for k in sub:
            a, b = basecall(0, k), basecall(1, k)
            res.append((F(a["path_prob"] + b["path_prob"]), k, a, b))
        best = res[0]
        for r in res[1:]:
            if not (r[0] < best[0]):
                best = r
        for st in range(2):
            calls[st] = dict(model=best[1][st], path_prob=best[2 + st]["path_prob"], bases=best[2 + st]["bases"],
                             states=best[2 + st]["states"], key=best[1])
    else:
        for st in range(2):
            if nev[st] < opts.min_ed_events:
                continue
            if preferred[st]:
                sub = [(preferred[st], "") if st == 0 else ("", preferred[st])]
            else:
                sub = sorted(k for k in pm_m if k[st] and not k[1 - st])
            res = [(basecall(st, k), k) for k in sub]
            best = res[0]
            for r in res[1:]:
                if not (r[0]["path_prob"] < best[0]["path_prob"]):
                    best = r
            calls[st] = dict(model=best[1][st], path_prob=best[0]["path_prob"], bases=best[0]["bases"],
                             states=best[0]["states"], key=best[1])
    return dict(pm_params=pm_m, st_params=st_m, fits=fits, rounds=rounds, preferred=preferred, calls=calls,
                together=together)
