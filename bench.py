#!/usr/bin/env python3
"""bench.py -- Viterbi events/sec of the nanocall decoding hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--reads R] [--events E] [--impl ours|reference]

A "step" is one pass of the hot path (scale + transitions + drift + Viterbi forward + traceback +
moves) over one batch of synthetic reads.  Default workload = BASELINE.json configs[1]:
10k reads x 10k events of R7.3 template, fixed identity scaling, default transitions, 1 B200.
Under torchrun (N > 1) every rank decodes its own R reads on its own GPU (reads shard naturally,
no data-path collective): scaling = "weak", value = all ranks' events / max-over-ranks time.

JSON line keys beyond the base contract:
  roofline      HBM view of the dominant kernel (viterbi_alpha_kernel): algorithmic bytes = 4112 B/event
                (4096 B backpointers + 12 B event + 1 B traceback read + 3 B state/move) over the
                kernel's own launch duration (CUDA events on the context's stream around each launch);
                `traffic` = DRAM bytes per launch from the committed ncu capture (the kernel streams the alpha
                columns, 16 KiB/event, by design: DESIGN.md K1a)
  roofline_issue executed work: the kernel's warp instructions per event (ncu smsp__inst_executed on this very launch,
                profiles/r2_ncu_viterbi_alpha_10kx10k_final_summary.md) x events/s against the issue peak of 148 SMs x 4
                sub-partitions x SM clock under load; the FMA-heavy pipe, the busiest one, was 67.0 % active in that capture.
                (The algorithmic count of SURVEY 8d, 245,600 FP32 op/event, is NOT used as a roof: sharing the class maxima
                removes two thirds of the per-edge operations, so a fraction built on it exceeds 1.)
  roofline_rf   what binds (profiles/r1_viterbi_alpha_experiments.md): register-file operand reads,
                404 per thread and column, against the measured 2 reads per cycle and lane
  cpu_baseline  the reference's own Viterbi (oracle/_ref, kind "reference"; the C port otherwise)
                on a bounded sample of the same reads on this box's host cores
  e2e           same metric through nc_viterbi_packed with pinned HOST buffers: H2D of the events
                and D2H of states/moves/scores inside the timed region
  parity        the reads of the cpu_baseline sample decoded by the GPU in the timed step, compared bit for bit
                (path log-probability, every state, every move) with the reference's own Viterbi::fill
  mixture       BASELINE.json configs[4] shape (90 % ~5k / 9 % 20-50k / 1 % 100-150k events) on the same context:
                events/s device-resident, with the longest read
  pipeline      BASELINE.json configs[2]/[3]: 2D reads (5000 + 5000 events) through the host pipeline
                (nanocall-b200: segmentation, Forward/Backward training rounds, model selection, Viterbi with the
                trained parameters, FASTA) from host memory; Forward/Backward events/s with the per-kernel device
                times, roofline against FP32 issue (1.78 M op per F/B event), and the reference's
                train_one_round timed on a sample of the same reads
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HBM_BYTES_PER_EVENT = 4112
FP32_OPS_PER_EVENT = 245600
RF_READS_PER_EVENT = 404 * 512  # operand reads per thread and column x threads (nc_viterbi_alpha.cu header)
# dram__bytes_read.sum + dram__bytes_write.sum per event of viterbi_alpha_kernel, from the committed capture
# measured on the bench's own launch (10000 reads x 10000 events, one kernel): 1.6387 TB written + 47.4 GB read
NCU_DRAM_BYTES_PER_EVENT = {"viterbi_alpha_kernel": 16861.0, "source": "profiles/r2_ncu_viterbi_alpha_10kx10k_final_summary.md"}
# executed work of the same launch: smsp__inst_executed.sum = 217.47e9 warp instructions for 1e8 events (248.68e9 before the
# exchange-buffer addresses were kept in registers and the column loop ran eight columns per trip)
NCU_WARP_INSTR_PER_EVENT = {"viterbi_alpha_kernel": 2174.7, "fmaheavy_pipe_pct": 67.0, "issue_active_pct": 53.6,
                            "source": "profiles/r2_ncu_viterbi_alpha_10kx10k_final_summary.md"}
MODEL = "r73.t.006.ont.model"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--reads", type=int, default=10000, help="reads per GPU")
    ap.add_argument("--events", type=int, default=10000, help="events per read")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-reads", type=int, default=0, help="0 = 2 per host thread")
    ap.add_argument("--mix", action="store_true",
                    help="read-length mixture of BASELINE.json configs[4] (90 %% ~5k, 9 %% 20-50k, 1 %% 100-150k events) "
                         "instead of --events per read; not the default line")
    ap.add_argument("--pipeline-reads", type=int, default=2000, help="2D reads of the pipeline section (0 = skip)")
    ap.add_argument("--mix-reads", type=int, default=4000, help="reads of the mixture section (0 = skip)")
    ap.add_argument("--vit-mode", default="auto", choices=["auto", "backpointer"],
                    help="auto = alpha-column kernel where the columns fit the pool; backpointer = long-read kernel only")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows)}


def host_threads_for_cpu_baseline(n_events):
    """All host threads, bounded so the reference's 32 KiB/event matrices fit in RAM."""
    cores = os.cpu_count() or 1
    try:
        import psutil
        avail = psutil.virtual_memory().available
        per_thread = n_events * 32768 + (64 << 20)
        cores = max(1, min(cores, int(avail * 0.5 // per_thread)))
    except Exception:
        pass
    return cores


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline(table, batch, n_events, sample_reads=0, want_paths=False):
    """Time the reference's Viterbi on a bounded sample (first reads of the batch). test-infra import."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    threads = host_threads_for_cpu_baseline(n_events)
    if oracle_lib.have_ref():
        lib, kind = oracle_lib.ref(), "reference"
    else:
        lib, kind = oracle_lib.port(), "port"
    n_reads = sample_reads or min(batch["ev_off"].size - 1, max(64, 4 * threads))   # >= 64 reads
    off = batch["ev_off"][:n_reads + 1]
    tot = int(off[-1])
    pm = np.tile(np.array([1, 0, 0, 1, 1, 1], np.float32), (n_reads, 1))
    st = np.tile(np.array([0.1, 0.3], np.float32), (n_reads, 1))
    t0 = time.perf_counter()
    res = lib.viterbi_batch(table, off, batch["mean"][:tot], batch["stdv"][:tot], batch["start"][:tot], pm, st,
                            n_threads=threads, want_paths=want_paths)
    dt = time.perf_counter() - t0
    out = {"value": tot / dt, "unit": "events/s", "cores": threads, "cpu": cpu_model(), "kind": kind,
           "sample": f"{n_reads} reads x {n_events} events ({tot} events) in {dt:.1f} s, "
                     f"{threads} threads, one read per worker (pfor chunk 1)"}
    return out, dt, res


def cpu_train_baseline(n_reads, threads):
    """The reference's Parameter_Trainer::train_one_round (oracle/_ref) on the first training round of n_reads synthetic 2D
    reads x 2 candidate model pairs, one group per worker thread: Forward/Backward events/s on the host cores."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from concurrent.futures import ThreadPoolExecutor
    from nanocall_b200 import models, synth
    if not oracle_lib.have_ref():
        return None
    ref = oracle_lib.ref()
    T = models.builtin_model("r73.t")["table"]
    C = [models.builtin_model("r73.c.p1")["table"], models.builtin_model("r73.c.p2")["table"]]
    rng = np.random.default_rng(7)
    groups = []
    for k in range(n_reads):
        pm = tuple(synth.random_params(rng, 1)[0])
        t = synth.make_read(rng, T, 200, pm)
        c = synth.make_read(rng, C[k % 2], 200, pm)
        seqs = [(0, t["mean"][:100], t["stdv"][:100], t["start"][:100]), (0, t["mean"][100:], t["stdv"][100:], t["start"][100:]),
                (1, c["mean"][:100], c["stdv"][:100], c["start"][:100]), (1, c["mean"][100:], c["stdv"][100:], c["start"][100:])]
        for cm in C:
            groups.append((seqs, cm))
    pm0 = np.array([1, 0, 0, 1, 1, 1], np.float32)
    st0 = np.array([0.1, 0.3, 0.1, 0.3], np.float32)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:   # ctypes releases the GIL inside the call
        list(ex.map(lambda g: ref.train_one_round(g[0], T, g[1], pm0, st0), groups))
    dt = time.perf_counter() - t0
    ev = 400 * len(groups)
    return {"value": ev / dt, "unit": "Forward/Backward events/s", "cores": threads, "cpu": cpu_model(), "kind": "reference",
            "sample": f"{len(groups)} groups (= {n_reads} reads x 2 candidate pairs) x 4 sequences x 100 events, one "
                      f"train_one_round each ({ev} events) in {dt:.1f} s, {threads} threads"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU Viterbi on the box's host cores."""
    if rank != 0:
        return
    from nanocall_b200 import synth, models
    table = models.builtin_model(MODEL)["table"]
    threads = host_threads_for_cpu_baseline(args.events)
    n_reads = args.cpu_sample_reads or min(args.reads, max(64, 4 * threads))   # >= 64 reads
    batch = synth.make_batch_uniform(args.seed, table, n_reads, args.events)
    times = []
    res = None
    for it in range(args.warmup + args.steps):
        res, dt, _ = cpu_baseline(table, batch, args.events, n_reads)
        if it >= args.warmup:
            times.append(dt)
        if sum(times) > 240:  # keep the whole run within a few minutes
            break
    tot = n_reads * args.events
    v = tot * len(times) / sum(times)
    res["value"] = v
    line = {"impl": "reference", "metric": "viterbi_events_per_sec", "value": v, "unit": "events/s",
            "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.reads} reads x {args.events} events R7.3 template, Viterbi only "
                                   f"(configs[1]); each step = bounded sample of {n_reads} reads",
                       "model": MODEL},
            "cpu_baseline": res,
            "e2e": {"value": v, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_OUT = None


def emit(line):
    """The one JSON line of the contract, on the process's original stdout."""
    print(json.dumps(line), file=_OUT or sys.stdout, flush=True)


def main():
    # stdout carries exactly one JSON line: anything a library prints there (NCCL prints its version at the first
    # collective) goes to stderr instead; the line itself is written to the saved descriptor
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from nanocall_b200 import api, synth, models
    from nanocall_b200 import dist as ncd

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: nanocall_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    table = models.builtin_model(MODEL)["table"]
    if args.mix:
        lengths = synth.mixture_lengths(args.seed + rank, args.reads)
        batch = synth.make_batch_uniform(args.seed + rank, table, args.reads, 0, lengths=lengths)
        total = int(lengths.sum())
    else:
        batch = synth.make_batch_uniform(args.seed + rank, table, args.reads, args.events)
        total = args.reads * args.events

    ctx = api.Context(local_rank)
    mid = ctx.register_model(table, 0)
    info = ctx.device_info()
    if args.vit_mode == "backpointer":
        from nanocall_b200 import _lib as L
        ctx.set_viterbi_mode(L.NC_VIT_BACKPOINTER)

    # pinned host copies (e2e) and device-resident copies (value)
    host = {k: torch.from_numpy(v).pin_memory() for k, v in
            (("mean", batch["mean"]), ("stdv", batch["stdv"]), ("start", batch["start"]))}
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    d_states = torch.empty(total, dtype=torch.int16, device=dev)
    d_moves = torch.empty(total, dtype=torch.uint8, device=dev)
    h_states = torch.empty(total, dtype=torch.int16).pin_memory()
    h_moves = torch.empty(total, dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def step_device():
        # log_stdv is not supplied: the kernel derives it on the device (glibc-compatible logf)
        return ctx.viterbi_device(batch["ev_off"], d["mean"].data_ptr(), d["stdv"].data_ptr(), d["start"].data_ptr(),
                                  None, mid, d_states=d_states.data_ptr(), d_moves=d_moves.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    path = None
    launches = 0
    kernel_ms = []
    for _ in range(args.steps):
        path = step_device()
        launches += ctx.last_launches()
        kernel_ms.append(ctx.last_kernel_ms())   # CUDA events on the context's stream around the kernel launch
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    sampler.stop_flag.set()
    sampler.join()
    clocks = sampler.summary()

    # e2e: pinned host buffers in, host results out, wall clock around the public call
    e2e = None
    if not args.no_e2e:
        hp = {k: v.numpy() for k, v in host.items()}
        hs, hm = h_states.numpy().view(np.uint16), h_moves.numpy()
        lib = ctx.lib
        n = args.reads
        midv = np.full(n, mid, np.int32)
        pm_a, st_a = api._pm_array(None, n), api._st_array(None, n)
        pathv = np.zeros(n, np.float32)

        def step_host():
            ctx._check(lib.nc_viterbi_packed(ctx.h, n, batch["ev_off"].ctypes.data, hp["mean"].ctypes.data,
                                             hp["stdv"].ctypes.data, hp["start"].ctypes.data, None,
                                             midv.ctypes.data, pm_a.ctypes.data, st_a.ctypes.data, 0,
                                             pathv.ctypes.data, hs.ctypes.data, hm.ctypes.data))
        step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        barrier()
        e2e_s = time.perf_counter() - t0
        if not np.array_equal(pathv.view(np.uint32), path.view(np.uint32)):
            raise SystemExit("e2e and device-resident paths disagree")
        e2e = (e2e_s, total * 12 + n * (288 + 4), total * 3 + n * 4)

    # ---- parity of the timed workload itself: the first reads of the batch as the GPU decoded them in the timed steps,
    # bit for bit against the reference's Viterbi::fill (the same call is the cpu_baseline measurement)
    cpu_res = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.mix:
        cpu_res, _, ref_out = cpu_baseline(table, batch, args.events, args.cpu_sample_reads, want_paths=True)
        n_chk = ref_out["path_prob"].size
        tot_chk = int(batch["ev_off"][n_chk])
        g_states = d_states[:tot_chk].cpu().numpy().view(np.uint16)
        g_moves = d_moves[:tot_chk].cpu().numpy()
        same_path = path[:n_chk].view(np.uint32) == ref_out["path_prob"].view(np.uint32)
        off = batch["ev_off"].astype(np.int64)
        same_read = [bool(same_path[k]) and np.array_equal(g_states[off[k]:off[k + 1]], ref_out["states"][off[k]:off[k + 1]])
                     and np.array_equal(g_moves[off[k]:off[k + 1]], ref_out["moves"][off[k]:off[k + 1]]) for k in range(n_chk)]
        parity = {"reads_checked": int(n_chk), "events_checked": tot_chk, "identical": int(sum(same_read)),
                  "path_logprob_bit_identical": int(same_path.sum()), "against": cpu_res["kind"],
                  "what": "path log-probability (bits), every state and every move of the timed device-resident step"}

    # ---- mixture (configs[4] shape) on the same context
    mixture = None
    if args.mix_reads > 0 and not args.mix:
        del d, d_states, d_moves, h_states, h_moves, host
        torch.cuda.empty_cache()
        lengths = synth.mixture_lengths(args.seed + 100 + rank, args.mix_reads)
        mb = synth.make_batch_uniform(args.seed + 100 + rank, table, args.mix_reads, 0, lengths=lengths)
        mtotal = int(lengths.sum())
        md = {k: torch.from_numpy(mb[k]).to(dev) for k in ("mean", "stdv", "start")}
        ms_states = torch.empty(mtotal, dtype=torch.int16, device=dev)
        ms_moves = torch.empty(mtotal, dtype=torch.uint8, device=dev)

        def step_mix():
            return ctx.viterbi_device(mb["ev_off"], md["mean"].data_ptr(), md["stdv"].data_ptr(), md["start"].data_ptr(),
                                      None, mid, d_states=ms_states.data_ptr(), d_moves=ms_moves.data_ptr())
        for _ in range(2):
            step_mix()
        barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record(stream)
        n_mix = max(1, args.steps)
        for _ in range(n_mix):
            step_mix()
        m1.record(stream)
        barrier()
        mix_ms = ncd.max_over_ranks([m0.elapsed_time(m1) / n_mix], dev)[0]
        mix_events = ncd.sum_over_ranks([mtotal], dev)[0]
        mixture = {"value": mix_events / (mix_ms * 1e-3), "unit": "events/s", "ms_per_step": mix_ms, "reads_per_gpu": args.mix_reads,
                   "events": mix_events, "longest_read": int(lengths.max()),
                   "workload": "90 % ~5k / 9 % 20-50k / 1 % 100-150k events, R7.3 template, Viterbi + traceback, device-resident"}
        del md, ms_states, ms_moves
    info_name, n_sms = info["name"], info["n_sms"]
    ctx.close()
    torch.cuda.empty_cache()

    # ---- full pipeline (configs[2]/[3]) through the host program, one process per GPU, host buffers
    pipeline = None
    if args.pipeline_reads > 0 and not args.mix:
        cli = os.path.join(ROOT, "nanocall_b200", "bin", "nanocall-b200")
        js = f"/tmp/nc_bench_pipeline_{os.getpid()}.json"
        cmd = [cli, "--pore", "r73", "--synth", f"{args.pipeline_reads}:{args.seed + rank}:{min(args.pipeline_reads, 2048)}:2d:5000:5000",
               "-o", "/dev/null", "--log", "warning", "--device", str(local_rank), "--batch-reads", str(args.pipeline_reads),
               "--batch-mevents", "64", "--summary-json", js]
        barrier()
        pr = subprocess.run(cmd, capture_output=True, text=True)
        if pr.returncode != 0:
            raise SystemExit("pipeline section failed: " + pr.stderr[-2000:])
        with open(js) as f:
            pj = json.load(f)
        os.remove(js)
        dv = pj["devices"][0]
        agg = ncd.sum_over_ranks([pj["read_events"], dv["fwbw_events"], dv["viterbi_events"]], dev)
        slow = ncd.max_over_ranks([pj["steady_wall_s"], dv["train_kernel_ms"], dv["viterbi_kernel_ms"], dv["fwbw_ms"],
                                   dv["pm_stats_ms"], dv["st_stats_ms"], dv["emission_ms"]], dev)
        pipeline = {"reads_per_gpu": args.pipeline_reads, "read_events": agg[0],
                    "read_events_per_s": agg[0] / slow[0], "steady_wall_s": slow[0],
                    "fwbw_events": agg[1], "fwbw_events_per_s": agg[1] / (slow[1] * 1e-3),
                    "train_kernel_ms": slow[1], "viterbi_events": agg[2],
                    "viterbi_events_per_s": agg[2] / (slow[2] * 1e-3), "viterbi_kernel_ms": slow[2],
                    "kernel_ms": {"fwbw_kernel": slow[3], "pm_stats_kernel": slow[4], "st_stats_kernel": slow[5], "emission_kernel": slow[6]},
                    "train_rounds": dv["train_rounds"], "init_s": dv["init_s"],
                    "workload": f"{args.pipeline_reads} synthetic 2D reads per GPU x (5000 template + 5000 complement events, hairpin), "
                                "r73 preset: segmentation, <= 20 EM rounds on 4 x 100 events for 2 candidate model pairs, "
                                "selection, Viterbi with the trained parameters, FASTA; timed from the first batch handed to the "
                                "GPU to the last record written (steady_wall_s); kernel times are CUDA events"}
    else:
        barrier()

    # whole-job numbers: every rank decoded `total` events; the job is as slow as its slowest rank
    ms, e2e_max = ncd.max_over_ranks([ms, e2e[0] if e2e else 0.0], dev)
    if e2e:
        e2e = (e2e_max,) + e2e[1:]

    if rank == 0:
        ms_per_step = ms / args.steps
        value = world * total / (ms_per_step * 1e-3)
        hbm_peak, peak_kind = peaks()
        per_gpu_evs = total / (ms_per_step * 1e-3)
        k_ms = sum(kernel_ms) / len(kernel_ms)
        ach_gbs = total * HBM_BYTES_PER_EVENT / (k_ms * 1e-3) / 1e9
        kname = "viterbi_kernel" if args.vit_mode == "backpointer" else "viterbi_alpha_kernel"
        traffic = NCU_DRAM_BYTES_PER_EVENT.get(kname)
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        fp32_peak = n_sms * 128 * sm_mhz * 1e6 / 1e12  # T FP32 instr/s (non-FMA issue)
        line = {
            "metric": "viterbi_events_per_sec", "value": value, "unit": "events/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": (f"{args.reads} reads per GPU, length mixture 90 % ~5k / 9 % 20-50k / 1 % 100-150k events "
                                    f"({total} events, longest {int(np.diff(batch['ev_off'].astype(np.int64)).max())}), R7.3 template, "
                                    "fixed identity scaling, default transitions, Viterbi + traceback (configs[4] shape)")
                       if args.mix else
                                   f"{args.reads} reads x {args.events} events per GPU, R7.3 template, "
                                   "fixed identity scaling, default transitions, Viterbi + traceback (configs[1])",
                       "model": MODEL, "reads_per_gpu": args.reads, "events_per_read": args.events,
                       "l2": "inputs larger than L2: 12 B/event of events (1.2 GB per 1e8 events) and 16 KiB/event of alpha columns (164 MB per 10k-event read) stream through HBM, nothing is reused across steps",
                       "device": info_name, "n_sms": n_sms},
            "clocks": clocks,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": ach_gbs / hbm_peak, "traffic": traffic * total if traffic else None,
                         "traffic_source": NCU_DRAM_BYTES_PER_EVENT["source"] if traffic else None,
                         "peak_kind": peak_kind, "kernel": kname, "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": HBM_BYTES_PER_EVENT * total,
                         "algorithmic_bytes_per_event": HBM_BYTES_PER_EVENT},
        }
        if kname in NCU_WARP_INSTR_PER_EVENT:
            wi = NCU_WARP_INSTR_PER_EVENT[kname]
            issue_peak = n_sms * 4 * sm_mhz * 1e6 / 1e12   # T warp-instructions/s
            line["roofline_issue"] = {"bound": "warp_issue", "achieved": per_gpu_evs * wi / 1e12, "peak": issue_peak,
                                      "unit": "T warp-instr/s", "frac": per_gpu_evs * wi / 1e12 / issue_peak,
                                      "warp_instructions_per_event": wi, "fmaheavy_pipe_pct_ncu": NCU_WARP_INSTR_PER_EVENT["fmaheavy_pipe_pct"],
                                      "issue_active_pct_ncu": NCU_WARP_INSTR_PER_EVENT["issue_active_pct"],
                                      "source": NCU_WARP_INSTR_PER_EVENT["source"],
                                      "algorithmic_fp32_ops_per_event": FP32_OPS_PER_EVENT,
                                      "peak_kind": f"{n_sms} SMs x 4 sub-partitions x {sm_mhz:.0f} MHz under load"}
        if args.vit_mode != "backpointer":
            rf_peak = n_sms * 128 * 2 * sm_mhz * 1e6 / 1e12
            line["roofline_rf"] = {"bound": "register_operand_reads", "achieved": per_gpu_evs * RF_READS_PER_EVENT / 1e12,
                                   "peak": rf_peak, "unit": "Treads/s", "frac": per_gpu_evs * RF_READS_PER_EVENT / 1e12 / rf_peak,
                                   "reads_per_event": RF_READS_PER_EVENT,
                                   "peak_kind": f"{n_sms} SMs x 128 lanes x 2 reads/clk (tools/ubench) x {sm_mhz:.0f} MHz"}
        if e2e:
            line["e2e"] = {"value": world * total * args.steps / e2e[0], "unit": "events/s",
                           "h2d_bytes_per_step": e2e[1], "d2h_bytes_per_step": e2e[2],
                           "timing": "wall clock around nc_viterbi_packed(NC_MEM_HOST), pinned buffers"}
        if cpu_res:   # (not with --mix: a 150k-event read needs 4.9 GB per CPU thread)
            line["cpu_baseline"] = cpu_res
            line["parity"] = parity
        if mixture:
            line["mixture"] = mixture
        if pipeline:
            fb_ops = 1.78e6   # FP32 operations per Forward/Backward event
            ach = pipeline["fwbw_events_per_s"] / world * fb_ops / 1e12
            pipeline["roofline"] = {"bound": "fp32_issue", "achieved": ach, "peak": fp32_peak, "unit": "Tinstr/s", "frac": ach / fp32_peak,
                                    "algorithmic_ops_per_event": fb_ops, "kernels": "emission + fwbw + pm_stats + st_stats (all training kernels)",
                                    "peak_kind": f"{n_sms} SMs x 128 lanes x {sm_mhz:.0f} MHz"}
            if world == 1 and not args.no_cpu_baseline:
                pipeline["cpu_baseline"] = cpu_train_baseline(16, host_threads_for_cpu_baseline(400))
            line["pipeline"] = pipeline
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
