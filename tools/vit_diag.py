"""Diagnostics for the alpha-column Viterbi kernel on one GPU: throughput of (a) the full call (forward + stores +
traceback), (b) path probability only (forward pass, no stores, no traceback), with the kernel's device counters.
usage: python tools/vit_diag.py [--reads 4000] [--events 10000]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=4000)
    ap.add_argument("--events", type=int, default=10000)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--seed", type=int, default=11)
    ap.add_argument("--mix", action="store_true", help="read-length mixture of configs[4] instead of --events per read")
    args = ap.parse_args()
    import torch
    from nanocall_b200 import api, models, synth

    table = models.builtin_model("r73.t")["table"]
    if args.mix:
        lengths = synth.mixture_lengths(args.seed, args.reads)
        batch = synth.make_batch_uniform(args.seed, table, 0, 0, lengths=lengths)
        total = int(lengths.sum())
    else:
        batch = synth.make_batch_uniform(11, table, args.reads, args.events)
        total = args.reads * args.events
    dev = torch.device("cuda", 0)
    ctx = api.Context(0)
    mid = ctx.register_model(table, 0)
    d = {k: torch.from_numpy(batch[k]).to(dev) for k in ("mean", "stdv", "start")}
    d_states = torch.empty(total, dtype=torch.int16, device=dev)
    d_moves = torch.empty(total, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    sm_hz = 1.965e9

    def run(tag, states, moves):
        out = {}
        for _ in range(args.reps + 1):
            ctx.viterbi_stats(reset=True)
            ctx.viterbi_device(batch["ev_off"], d["mean"].data_ptr(), d["stdv"].data_ptr(), d["start"].data_ptr(), None, mid,
                               d_states=states, d_moves=moves)
            ms = ctx.last_kernel_ms()
            st = ctx.viterbi_stats(reset=True)
            out = {"mode": tag, "ms": ms, "events_per_s": total / ms * 1e3, **st}
            n_fwd = 145 if states else 148
            out["fwd_busy_frac"] = st["fwd_cycles"] / (n_fwd * ms * 1e-3 * sm_hz)
            out["cycles_per_column"] = st["fwd_cycles"] / total
            if states:
                out["slab_wait_frac"] = st["fwd_wait_slab_cycles"] / (n_fwd * ms * 1e-3 * sm_hz)
                out["tb_busy_frac_of_64_warps"] = st["tb_busy_cycles"] / (64 * ms * 1e-3 * sm_hz)
                out["tb_steps_per_event"] = st["tb_lane_steps"] / total
        print(json.dumps(out), flush=True)

    # pure-write HBM bandwidth for comparison (the alpha kernel is a 16 KiB/event write stream)
    x = torch.empty(1 << 30, dtype=torch.float32, device=dev)
    for _ in range(2):
        x.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        x.zero_()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"write_only_GBps": 5 * x.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9}), flush=True)
    del x
    run("full", d_states.data_ptr(), d_moves.data_ptr())
    run("path_only", None, None)
    ctx.close()


if __name__ == "__main__":
    main()
