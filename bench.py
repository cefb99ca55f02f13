#!/usr/bin/env python
"""bench.py -- images/sec of the full ColorHandPose3DNetwork.inference pipeline on synthetic 320x320 batches.

  python bench.py --gpus N --steps K --warmup W            # our sm_100a path (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...   # the CPU restatement of the TF1 reference (oracle)

One JSON line on stdout (rank 0).  A "step" is one pass of the full pipeline (HandSegNet -> mask/crop ->
PoseNet2D -> PosePrior/Viewpoint lifting -> x8 up-sampling -> key-point arg-max) over one batch of
`--batch` synthetic images PER GPU (weak scaling; BASELINE config 4 shards 32 images per GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GFLOP_PER_IMAGE = 142.258408192          # conv + FC FLOPs of the full pipeline (SURVEY.md 8a.1 / arch.conv_flops_per_image)
METRIC = "images/sec full pipeline 320x320"


def measured_traffic(precision="bf16x3"):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/*_summary.json), or None."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    if os.path.isdir(pdir):
        for fn in sorted(os.listdir(pdir)):
            if fn.endswith("_summary.json") and (("f8c" in fn) == (precision == "fp16_f8c")):
                try:
                    d = json.load(open(os.path.join(pdir, fn)))
                    if "tc_conv" in d:
                        best = (fn, d["tc_conv"])
                except Exception:
                    pass
    return best


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {"tflops_burst": float(d["bf16_tflops"]), "tflops_sustained": float(d["bf16_tflops_sustained"]),
                    "hbm_gbs": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
        except Exception:
            pass
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons through NVML while the timed region runs."""
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            names = {getattr(N, k): k for k in dir(N) if k.startswith("nvmlClocksEventReason") or k.startswith("nvmlClocksThrottleReason")}
            while not self.stop_flag:
                self.sm.append(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM))
                try:
                    mask = N.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if isinstance(bit, int) and bit and (mask & bit) == bit and bin(bit).count("1") == 1:
                        self.reasons.add(nm.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
                time.sleep(0.002)
        except Exception as e:  # NVML unavailable: report that instead of clocks
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        r = sorted(x for x in self.reasons if x not in ("GpuIdle", "None", "ApplicationsClocksSetting"))
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": r, "samples": len(self.sm)}


def cpu_reference_throughput(n_images, H, W, seconds_cap=40.0):
    """Times the oracle (CPU restatement of the TF1 graph; the reference itself needs TensorFlow 1.3, which cannot be
    installed here) on the host cores.  The sample is one batched oracle.inference() call per thread-count candidate;
    the best throughput is reported.  The mask grower runs in its fast boolean form (bit-identical to the literal
    32 x dilation2d sequence, tests/test_oracle_kat.py), which is the generous choice for the CPU side.
    Returns (images/s, threads used, sample description)."""
    import torch
    from hand3d_b200 import weights as Wt
    from oracle import hand3d_oracle as O
    cores = os.cpu_count() or 1
    wd = Wt.synthetic_weights(0)
    img = Wt.synthetic_images(n_images, H, W, seed=100)
    hs = Wt.synthetic_hand_side(n_images, seed=2)
    best, best_thr, t_start = 0.0, cores, time.perf_counter()
    cands = sorted({cores, max(1, cores // 2), max(1, cores // 4), min(cores, 32), min(cores, 16)}, reverse=True)
    for thr in cands:
        torch.set_num_threads(thr)
        O.inference(img[:1], hs[:1], wd, literal_mask=False)                  # warm-up for this thread count
        t0 = time.perf_counter()
        O.inference(img, hs, wd, literal_mask=False)
        v = n_images / (time.perf_counter() - t0)
        if v > best:
            best, best_thr = v, thr
        if time.perf_counter() - t_start > seconds_cap:
            break
    torch.set_num_threads(cores)
    return best, best_thr, "one batched oracle inference() of [%d,%d,%d,3] per thread-count candidate %s, best kept (%d threads of %d cores)" % (
        n_images, H, W, cands, best_thr, cores)


def run_reference(args):
    """Reference arm: the CPU restatement of the TF1 graph (oracle/) on the host cores -- the unmodified reference cannot run
    (TensorFlow 1.3 is not installable offline, and the repo ships no weights).  Warm-up = thread-count sweep; each timed
    step = one batched oracle inference() sized so that the whole run stays within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from hand3d_b200 import weights as Wt
    from oracle import hand3d_oracle as O
    v0, thr, sample = cpu_reference_throughput(min(args.ref_images, 4), args.height, args.width, seconds_cap=60.0)
    torch.set_num_threads(thr)
    n = int(max(1, min(args.ref_images, round(150.0 * v0 / max(1, args.steps)))))      # ~150 s of timed CPU work in total
    wd = Wt.synthetic_weights(0)
    img = Wt.synthetic_images(n, args.height, args.width, seed=100)
    hs = Wt.synthetic_hand_side(n, seed=2)
    for _ in range(max(0, args.warmup - 1)):
        O.inference(img, hs, wd, literal_mask=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.inference(img, hs, wd, literal_mask=False)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "full ColorHandPose3DNetwork.inference %dx%d, CPU restatement of the TF1 reference (oracle/); "
                               "TensorFlow 1.3 is not installable" % (args.height, args.width), "images_per_step": n,
                   "threads": thr},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": thr, "kind": "port",
                         "sample": "%d steps x one batched oracle inference() of %d images, %d torch threads (best of the warm-up sweep: %s)" % (
                             args.steps, n, thr, sample)},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from hand3d_b200 import runtime, weights as Wt
    from hand3d_b200.distributed import P2PGather, gather_records, pack_records

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device -- hand3d_b200 has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, H, W = args.batch, args.height, args.width

    ctx = runtime.Context(local_rank, precision=args.precision)
    ctx.load_weights(Wt.synthetic_weights(0))
    # synthetic inputs: NBUF different batches per rank (rotated every step so that inputs > L2 never repeat back to back)
    NBUF = 4
    host_imgs = [torch.from_numpy(Wt.synthetic_images(B, H, W, seed=1000 + 17 * rank + i)).pin_memory() for i in range(NBUF)]
    host_hs = [torch.from_numpy(Wt.synthetic_hand_side(B, seed=2000 + 17 * rank + i)).pin_memory() for i in range(NBUF)]
    dev_imgs = [t.to(dev) for t in host_imgs]
    dev_hs = [t.to(dev) for t in host_hs]

    # multi-GPU result exchange: fused pack + peer-memory all-gather kernel (NCCL all_gather only with --gather nccl)
    p2p = None
    if world > 1 and args.gather == "p2p":
        p2p = P2PGather(ctx, max_batch=B)

    def exchange(r):
        if p2p is not None:
            return p2p.gather(r["keypoint_coord3d"], r["keypoints_uv"], r["center"], r["scale_crop"])
        rec = pack_records(r["keypoint_coord3d"], r["keypoints_uv"], r["center"], r["scale_crop"])
        return gather_records(rec) if world > 1 else rec

    # one CUDA graph per input buffer (the forward pass is sync-free with fixed pointers: ~90 launches replay as one)
    graphs = None
    if args.cuda_graph:
        graphs = []
        for k in range(NBUF):
            c0 = ctx.launch_count
            replay, res = ctx.capture_pipeline(dev_imgs[k], dev_hs[k], True, outputs="keypoints")
            graphs.append((replay, res, (ctx.launch_count - c0) // 2))   # warm-up + capture each issue the step once
    graph_launches = [0]

    def step_device(i):
        if graphs is not None:
            replay, r, nl = graphs[i % NBUF]
            replay()
            graph_launches[0] += nl
        else:
            r = ctx.pipeline(dev_imgs[i % NBUF], dev_hs[i % NBUF], True, outputs="keypoints")
        return exchange(r)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for i in range(args.warmup):
        step_device(i)
    barrier()

    # ---- timed region (device-resident inputs)
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count + graph_launches[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step_device(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count + graph_launches[0] - l0
    sampler.stop_flag = True
    sampler.join(timeout=2.0)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms / 1000.0)

    # ---- end-to-end: pinned host -> device copy of the step's inputs and device -> host read of the gathered key-point records
    # Double-buffered: the pinned-host -> device copy of step i+1 runs on a copy stream while step i computes.
    copy_stream = torch.cuda.Stream(device=dev)
    stage_img = [torch.empty((B, H, W, 3), dtype=torch.float32, device=dev) for _ in range(2)]
    stage_hs = [torch.empty((B, 2), dtype=torch.float32, device=dev) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        k = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[k])            # the compute stream is done reading this staging buffer
            stage_img[k].copy_(host_imgs[i % NBUF], non_blocking=True)
            stage_hs[k].copy_(host_hs[i % NBUF], non_blocking=True)
            ready[k].record(copy_stream)

    def run_e2e(n):
        cur = torch.cuda.current_stream()
        for k in range(2):
            consumed[k].record(cur)
        prefetch(0)
        for i in range(n):
            k = i & 1
            if i + 1 < n:
                prefetch(i + 1)
            cur.wait_event(ready[k])
            r = ctx.pipeline(stage_img[k], stage_hs[k], True, outputs="keypoints")
            consumed[k].record(cur)
            out_host.copy_(exchange(r), non_blocking=True)  # device -> host read of the step's (gathered) result

    out_host = torch.empty((world * B, 108), dtype=torch.float32).pin_memory()
    run_e2e(max(2, args.warmup // 2))
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    f0.record()
    run_e2e(args.steps)
    f1.record()
    barrier()
    t = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(t.item()) / 1000.0)
    h2d = B * H * W * 3 * 4 + B * 2 * 4
    d2h = world * B * 108 * 4

    # ---- per-kernel-class timing (CUDA events around every launch; separate pass so it does not perturb `value`)
    prof_steps = min(3, args.steps)
    ctx.profile_begin()
    for i in range(prof_steps):
        ctx.pipeline(dev_imgs[i % NBUF], dev_hs[i % NBUF], True, outputs="keypoints")
    prof = ctx.profile_end()
    peaks = measured_peaks()
    roof = None
    dominant = "tc_conv" if prof["tc_conv"]["launches"] else "direct_conv"
    d = prof[dominant]
    if d["ms"] > 0:
        achieved = d["flops"] / (d["ms"] * 1e-3) / 1e12
        peak = peaks["tflops_sustained"] if dominant == "tc_conv" else 75.0
        tr = measured_traffic(args.precision) if (dominant == "tc_conv" and args.precision in ("bf16x3", "fp16_f8c")) else None
        passes = 3 if args.precision in ("bf16x3", "fp16x3") else (2 if args.precision == "fp16_f8c" else 1)
        roof = {"bound": "tensor", "kernel": "conv_tc_kernel / conv_tc2_kernel / conv_c64_kernel (tcgen05 implicit GEMM: all conv layers with Cin >= 21 and the FC stacks)" if dominant == "tc_conv" else "conv_direct_kernel (fp32 FFMA)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": tr[1]["dram_bytes_per_launch"] if tr else None,
                "traffic_source": ("ncu dram__bytes_read+write per launch, B=32, profiles/%s" % tr[0]) if tr else None,
                "achieved_per_launch": {"gflop": d["flops"] / max(1, d["launches"]) / 1e9, "us": 1e3 * d["ms"] / max(1, d["launches"])},
                "peak_source": peaks["source"] + (", bf16 sustained" if dominant == "tc_conv" else ", nominal fp32 FFMA"),
                "launches_per_step": d["launches"] // prof_steps, "ms_per_step": d["ms"] / prof_steps,
                "mma_passes": passes,
                # fp32 parity costs `passes` tensor-core passes per algorithmic FLOP: the executed rate is what the tensor pipe sees
                "executed": {"value": achieved * passes, "unit": "TFLOP/s", "frac": achieved * passes / peak} if dominant == "tc_conv" else None,
                "share_of_step": (d["ms"] / prof_steps) / (ms / args.steps),
                "by_class_ms_per_step": {k: v["ms"] / prof_steps for k, v in prof.items()}}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample = cpu_reference_throughput(args.cpu_images, H, W, seconds_cap=25.0)
            cpu = {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16x3": "bf16 hi/lo split x3 MMA passes, fp32 accumulate (fp32 parity, 1e-3)",
                      "fp16x3": "fp16 hi/lo split x3 MMA passes, fp32 accumulate (fp32 parity, 1e-3)",
                      "fp16_f8c": "fp16 main pass + two fp8 (e4m3) correction passes, fp32 accumulate (fp32 parity, 1e-3)",
                      "fp16": "fp16 (1e-2 path)", "bf16": "bf16", "fp32_ffma": "f32"}[args.precision],
            "data": "synthetic",
            "config": {"workload": "full ColorHandPose3DNetwork.inference (HandSegNet+PoseNet2D+PosePrior/Viewpoint, %dx%d input, 256x256 crop), "
                                   "%d images per GPU per step (BASELINE config 4 shard)" % (H, W, B),
                       "global_batch": world * B, "precision": args.precision, "parallelism": "dp%d" % world, "cuda_graph": bool(args.cuda_graph),
                       "l2": "inputs rotate over %d distinct batches per rank (%.0f MB > L2); activations per step %.1f GB" % (
                           NBUF, NBUF * B * H * W * 12 / 1e6, B * 0.312),
                       "collective": ("none (single GPU)" if world == 1 else
                                      "fused pack + all-gather of 432 B/image records over NVLink peer memory (h3d_gather_records_p2p, %s)" % (
                                          "multimem store" if (p2p is not None and p2p.mc) else "peer stores")
                                      if p2p is not None else "NCCL all_gather of 432 B/image records")},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": roof,
            "cpu_baseline": cpu,
            "tflops_algorithmic": value * GFLOP_PER_IMAGE / 1e3,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--height", type=int, default=320)
    ap.add_argument("--width", type=int, default=320)
    ap.add_argument("--precision", default=os.environ.get("H3D_PRECISION", "bf16x3"), choices=["bf16x3", "fp16x3", "fp16", "bf16", "fp32_ffma", "fp16_f8c"])
    ap.add_argument("--cpu-images", type=int, default=8, help="bounded CPU-baseline sample (images per oracle call)")
    ap.add_argument("--ref-images", type=int, default=8, help="--impl reference: images per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", default=os.environ.get("H3D_GATHER", "p2p"), choices=["p2p", "nccl"], help="multi-GPU result exchange")
    ap.add_argument("--cuda-graph", type=int, default=int(os.environ.get("H3D_CUDA_GRAPH", "0")), help="replay the step from a CUDA graph")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
