#!/usr/bin/env python
"""bench.py -- image-pairs/sec of the DeMoN two-view inference path at 256x192 (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the full pipeline (bootstrap + 3 x iterative + refinement, examples/example.py:87-99)
over one batch of 64 synthetic pairs per GPU: BASELINE.json configs[2] at N=1, configs[3] (512 pairs over 8
GPUs) at N=8 -- weak scaling, pairs are independent, the only collective is one NCCL all-gather of the final
depth / motion tensors, inside the timed region.

One JSON line on stdout (rank 0):
  value        pairs/s, inputs resident in HBM, whole job; `--inflight` (default 2) batches per GPU are in flight on their
               own streams and pipelines (every step is still one full batch through the whole path)
  e2e          pairs/s through the C-ABI host-buffer entry (pinned host -> device -> pipeline -> host), same scheme
  roofline     the dominant kernel (the tcgen05 conv kernel; the fp32 SIMT conv kernel when the net runs in
               fp32 mode) timed with CUDA events around its launches in a second, single-stream timed region
  cpu_baseline the CPU oracle (torch-CPU fp32 + C geometry ops) on a bounded sample, host cores stated
`--impl reference` times that CPU path alone (TensorFlow 1.4 cannot be installed here).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

PER_GPU_BATCH = 64
ITERATIONS = 3
METRIC = "image_pairs_per_sec_256x192"
N_INPUT_SETS = 4     # rotating input batches: 4 x 75.5 MB > 126 MB L2 (plus a ~2.6 GB activation workspace per step)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            p = json.load(open(path))
            burst = float(p.get("bf16_tflops", 1590.0))
            return {"hbm_gbs": float(p.get("hbm_gbs", 6650.0)), "bf16_tflops": burst,
                    "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", burst)), "source": "measured"}
        except (OSError, ValueError, TypeError):
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """CPU threads this process may really use: scheduler affinity, capped by the cgroup CPU quota
    (os.cpu_count() reports the machine's cores, not the container's share)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, int(q / per + 0.5)))
            break
        except Exception:
            continue
    return max(1, n)


def best_thread_count(net, ip, i22):
    """torch's CPU convolutions do not scale to every core count (oversubscription, tiny layers): try the
    full count and a few smaller ones on ONE pair each and keep the fastest, so the baseline is the best the
    host can do."""
    full = host_threads()
    # more than ~32 threads only oversubscribes these small convolutions (measured: 128 threads are 100x slower than 8)
    cands = sorted({min(full, 32), min(full, 16), min(full, 8)}, reverse=True)
    best, best_t = full, None
    for c in cands:
        torch.set_num_threads(c)
        net.bootstrap(ip[:1], i22[:1])          # warm
        t0 = time.perf_counter()
        net.bootstrap(ip[:1], i22[:1])
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def synthetic_inputs(batch, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, 6, 192, 256, generator=g) - 0.5


def cpu_oracle_rate(pairs, repeats=1):
    """pairs/s of the CPU restatement of the reference path on `pairs` synthetic pairs (bounded sample).
    Returns (rate, seconds, threads used)."""
    from demon_b200 import weights as W
    from oracle import ops as oops
    from oracle.network import OracleNets
    net = OracleNets(W.synthetic_weights(0))
    ip = synthetic_inputs(pairs, 1234).numpy()
    i22 = oops.median3x3_downsample(oops.median3x3_downsample(np.ascontiguousarray(ip[:, 3:6])))
    threads = best_thread_count(net, ip, i22)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        net.pipeline(ip, i22, iterations=ITERATIONS)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return pairs / best, best, threads


def run_reference(args, rank):
    """Reference arm: the reference's own CPU implementation of the path.  TensorFlow 1.4 / Eigen cannot be
    installed offline, so this is the oracle port (kind "port"); each step is a bounded sample of 2 pairs."""
    if rank != 0:
        return
    from demon_b200 import weights as W
    from oracle import ops as oops
    from oracle.network import OracleNets
    net = OracleNets(W.synthetic_weights(0))
    sample = 2
    ip = synthetic_inputs(sample, 1234).numpy()
    i22 = oops.median3x3_downsample(oops.median3x3_downsample(np.ascontiguousarray(ip[:, 3:6])))
    threads = best_thread_count(net, ip, i22)
    steps, warmup = max(1, min(args.steps, 10)), max(1, min(args.warmup, 2))
    for _ in range(warmup):
        net.pipeline(ip, i22, iterations=ITERATIONS)
    t0 = time.perf_counter()
    for _ in range(steps):
        net.pipeline(ip, i22, iterations=ITERATIONS)
    dt = time.perf_counter() - t0
    value = sample * steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "full pipeline (bootstrap + 3x iterative + refinement) at 256x192, bounded sample of %d pairs per step "
                                   "of the batch-64 workload, CPU" % sample},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                             "sample": "%d steps x %d pairs, torch-CPU fp32 convolutions + C geometry ops (oracle/), all host threads" % (steps, sample)},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=None, choices=["fp32", "3xtf32", "tf32"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="pairs per GPU (default: BASELINE.json configs[2])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=2, help="batches in flight per GPU (pipelines on their own streams)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    graph_warmup = 2 * N_INPUT_SETS + 1     # every rotating input set must be seen twice before its CUDA graph replays

    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    from demon_b200 import _lib, parallel, weights as W
    from demon_b200.networks_original import Session, DemonPipeline, DEFAULT_PRECISION
    rank, local, world = parallel.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    B = args.batch
    precision = args.precision or DEFAULT_PRECISION

    sess = Session(precision=precision)
    sess.load_weights(W.synthetic_weights(0))
    pipe = DemonPipeline(sess, batch_size=B, iterations=ITERATIONS)
    net = pipe.net
    inputs = [synthetic_inputs(B, 1234 + rank + 1000 * i).to(dev) for i in range(N_INPUT_SETS)]
    outs = {"predict_depth0": torch.empty(B, 1, 192, 256, device=dev), "predict_rotation": torch.empty(B, 3, device=dev),
            "predict_translation": torch.empty(B, 3, device=dev)}
    gather = parallel.OutputGather(B, world, device=dev)

    def step(i):
        pipe.forward(inputs[i % N_INPUT_SETS], None, outs)
        gather(outs["predict_depth0"], outs["predict_rotation"], outs["predict_translation"])

    # Two batches in flight per GPU: a second pipeline (own workspace, own output buffers) on a second stream, steps
    # alternate between the two.  The kernels of one batch fill the SMs the other leaves idle (tails of the persistent
    # kernels, the low-resolution layers, the dense layers); every step is still one full batch through the whole path.
    NF = max(1, args.inflight)
    pipes = [pipe] + [DemonPipeline(sess, batch_size=B, iterations=ITERATIONS, private_net=True) for _ in range(NF - 1)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(NF)]
    outs2 = [outs] + [{k: torch.empty_like(v) for k, v in outs.items()} for _ in range(NF - 1)]
    gathers = [gather] + [parallel.OutputGather(B, world, device=dev) for _ in range(NF - 1)]

    def step2(i):
        k = i % NF
        with torch.cuda.stream(streams[k]):
            pipes[k].forward(inputs[(i // NF) % N_INPUT_SETS], None, outs2[k])
            gathers[k](outs2[k]["predict_depth0"], outs2[k]["predict_rotation"], outs2[k]["predict_translation"])

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------------------------------
    # Two timed regions: the first gives `value` (no instrumentation, two batches in flight); the second runs the same
    # steps on one stream with CUDA events around every layer launch on the launching stream (demon_net_profile_*) and
    # feeds the roofline figures.
    for i in range(NF * max(args.warmup, graph_warmup)):
        step2(i)
    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.demon_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(streams[0])
    for k in range(1, NF):
        streams[k].wait_event(ev0)
    for i in range(args.steps):
        step2(i)
    for k in range(1, NF):
        ev_b = torch.cuda.Event()
        ev_b.record(streams[k])
        streams[0].wait_event(ev_b)
    ev1.record(streams[0])
    barrier()
    ms_value = ev0.elapsed_time(ev1)
    launches = int(lib.demon_launch_count() - launches0)
    _lib.check(lib.demon_net_profile_begin(net.ptr))
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev2.record()
    for i in range(args.steps):
        step(i)
    ev3.record()
    barrier()
    ms_local = ev2.elapsed_time(ev3)
    _lib.check(lib.demon_net_profile_end(net.ptr))
    clocks = sampler.stop() if rank == 0 else None
    ms = parallel.max_over_ranks(ms_value, dev)
    value = world * B * args.steps / (ms / 1e3)

    # ---- per-layer device time -> roofline of the dominant kernel ------------------------------------
    lm = W.layer_macs()
    FAMILIES = {0: "conv_simt_kernel (fp32 CUDA-core implicit GEMM)",
                1: "conv_tc_kernel (tcgen05 kind::tf32, operands in shared memory, %s)" % precision,
                2: "conv_tc_halo_kernel<false> (tcgen05 kind::tf32, TMA halo tile, A operand in TMEM, %s)" % precision,
                3: "conv_tc_halo_kernel<true> (tcgen05 kind::tf32, per-tap tiles, A operand in TMEM, %s)" % precision}
    fam = {k: [0.0, 0.0, 0] for k in FAMILIES}   # ms, MACs, launches
    top = []
    for i in range(lib.demon_net_num_layers(net.ptr)):
        name = lib.demon_net_layer_name(net.ptr, i).decode()
        t, calls, lpc, tc = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int(), ctypes.c_int()
        lib.demon_net_layer_profile(net.ptr, i, ctypes.byref(t), ctypes.byref(calls), ctypes.byref(lpc), ctypes.byref(tc))
        if calls.value == 0:
            continue
        macs = lm[name] * B * calls.value
        f = fam[tc.value]
        f[0] += t.value; f[1] += macs; f[2] += calls.value * lpc.value
        top.append((t.value, name, macs, tc.value))
    top.sort(reverse=True)
    pk = peaks()
    dom = max(fam, key=lambda k: fam[k][0])
    d_ms, d_macs, d_launches = fam[dom]
    achieved = 2.0 * d_macs / (d_ms / 1e3) / 1e12 if d_ms > 0 else 0.0
    if dom != 0:
        peak = pk["bf16_tflops_sustained"] / 2.0
        peak_note = ("%s bf16 sustained %.1f TF/s / 2 (kind::tf32 issues at half the bf16 rate); %s spends %d tensor MACs per "
                     "algorithmic MAC, so frac <= %.2f by construction" % (pk["source"], pk["bf16_tflops_sustained"], precision,
                                                                          3 if precision == "3xtf32" else 1, 1 / 3 if precision == "3xtf32" else 1))
    else:
        peak = 2 * 128 * 148 * 1.965e9 / 1e12     # fp32 FFMA peak at clocks.max.sm
        peak_note = "fp32 FFMA peak 128 FMA/clk/SM x 148 SMs x 1965 MHz (no measured fp32 figure in MEASURED_PEAKS.json)"
    # DRAM traffic of the heaviest launch of the dominant kernel, from the committed `ncu --set full` capture (profiles/)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.isfile(tpath):
        tj = json.load(open(tpath))
        key = {2: "refine0_upconv", 1: "refine2_upconv", 3: "conv3x"}.get(dom)
        if key in tj:
            traffic = {"bytes": tj[key], "launch": key, "algorithmic_bytes": {"refine0_upconv": 2 * 64 * 96 * 128 * 128 * 4 + 4 * 4 * 128 * 32 * 4,
                                                                              "refine2_upconv": 64 * 24 * 32 * 256 * 4 + 64 * 48 * 64 * 64 * 4,
                                                                              "conv3x": 64 * 24 * 64 * 128 * 4 + 64 * 24 * 32 * 128 * 4}.get(key),
                       "source": "profiles/r01_ncu_full.md (one launch at batch 64, dram__bytes_read.sum + dram__bytes_write.sum)"}
    roofline = {"bound": "tensor", "kernel": FAMILIES[dom], "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak if peak else None, "traffic": traffic,
                "algorithmic_flops_per_launch": 2.0 * d_macs / max(1, d_launches), "avg_launch_ms": d_ms / max(1, d_launches),
                "launches_timed": d_launches, "kernel_share_of_step": d_ms / ms_local if ms_local else None, "peak_note": peak_note,
                "instrumented_ms_per_step": ms_local / args.steps,
                "kernels": [{"kernel": FAMILIES[k], "ms_per_step": fam[k][0] / args.steps, "share_of_step": fam[k][0] / ms_local,
                             "tflops": 2.0 * fam[k][1] / (fam[k][0] / 1e3) / 1e12 if fam[k][0] > 0 else 0, "launches_per_step": fam[k][2] // args.steps}
                            for k in sorted(fam, key=lambda k: -fam[k][0]) if fam[k][2]],
                "top_layers_ms_per_step": [{"layer": n, "ms": t / args.steps, "tflops": 2.0 * m / (t / 1e3) / 1e12 if t > 0 else 0, "kernel": k}
                                           for t, n, m, k in top[:8]]}

    # ---- end to end through the C-ABI host entry (pinned host buffers, copies inside the timed region) ----
    # Two pipelines (own workspaces) on two streams, async host entry: the H2D / D2H copies of one batch overlap the
    # compute of the other; every step still moves its full input from pinned host memory and its result back, and the
    # timed region ends when the last result is on the host.
    e2e = None
    if True:
        NE = max(2, NF) if NF > 1 else 1
        h_in = [synthetic_inputs(B, 4321 + rank + 1000 * i).pin_memory() for i in range(NE)]
        h_depth = [torch.empty(B, 1, 192, 256).pin_memory() for _ in range(NE)]
        h_rot = [torch.empty(B, 3).pin_memory() for _ in range(NE)]
        h_tr = [torch.empty(B, 3).pin_memory() for _ in range(NE)]
        e2e_steps = max(4, min(args.steps, 20))

        def e2e_step(i):
            k = i % NE
            streams[k].synchronize()            # the previous result of this slot is on the host (and may be consumed)
            pipes[k].forward_host_async(h_in[k], None, h_depth[k], h_rot[k], h_tr[k], streams[k])

        for i in range(3 * NE):
            e2e_step(i)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_step(i)
        torch.cuda.synchronize()
        dt_local = time.perf_counter() - t0
        barrier()
        dt = parallel.max_over_ranks(dt_local, dev)
        e2e = {"value": world * B * e2e_steps / dt, "unit": "pairs/s", "h2d_bytes_per_step": int(h_in[0].numel() * 4),
               "d2h_bytes_per_step": int((h_depth[0].numel() + h_rot[0].numel() + h_tr[0].numel()) * 4), "steps": e2e_steps,
               "api": "demon_pipeline_forward_host_async (C ABI) via DemonPipeline.forward_host_async: pinned host buffers, two "
                      "pipelines on two streams so that one batch's copies overlap the other's compute"}

    # ---- CPU baseline on the host cores (rank 0, N=1 only) --------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, secs, threads = cpu_oracle_rate(8)
        cpu = {"value": rate, "unit": "pairs/s", "cores": threads, "kind": "port",
               "sample": "8 pairs of the same synthetic workload, one pass (%.1f s), torch-CPU fp32 convolutions + C geometry ops; "
                         "thread count picked as the fastest of a short calibration (host offers %d)" % (secs, host_threads())}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if precision == "fp32" else ("tf32x3" if precision == "3xtf32" else "tf32"), "data": "synthetic",
                "config": {"workload": "BASELINE.json configs[%d]: batch=%d synthetic 256x192 pairs per GPU, full pipeline "
                                       "(bootstrap + 3x iterative + refinement), %d GPU(s)" % (2 if world == 1 else 3, B, world),
                           "global_batch": world * B, "precision": precision, "iterations": ITERATIONS,
                           "l2": "%d rotating input batches of %.1f MB (> 126 MB L2) and a %.2f GB activation workspace rewritten every step"
                                 % (N_INPUT_SETS, B * 6 * 192 * 256 * 4 / 1e6, lib.demon_net_workspace_bytes(net.ptr) / 1e9),
                           "parallelism": "dp%d, one NCCL all-gather of depth0+motion per step" % world if world > 1 else "single GPU",
                           "batches_in_flight": "%d per GPU (pipelines with own workspaces on their own CUDA streams, steps alternate); the " % NF +
                                                "instrumented region behind `roofline` runs the same steps on one stream",
                           "flops_per_pair": 2.0 * W.macs_per_pair()["pipeline"]},
                "gpu_launches": launches, "clocks": clocks, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
                "algorithmic_tflops": 2.0 * W.macs_per_pair()["pipeline"] * value / 1e12}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
