#!/usr/bin/env python
"""Throughput harness for the DPT-Hybrid-384 hot path (BASELINE.json metric: 384x384 images/sec).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference CPU path

A "step" is one forward pass of the depth model over one batch of synthetic 384x384 RGB images.
N=1 workload = BASELINE.json configs[1] (depth head, bf16, batch 32, one B200).  For N>1 (launched
by torch.distributed.run, one rank per GPU) every rank runs the same per-GPU batch on its own
images — independent units, no data-path collective ("weak" scaling); weights are NCCL-broadcast
from rank 0 once, timings are reduced with MAX over ranks.

One JSON line on stdout (rank 0):  value = whole-job images/s with inputs resident in HBM (CUDA
events, device time, max over ranks);  e2e = the same metric through the public API
(`DPTDepthModel.forward`) with pinned-host inputs, H2D copy and D2H read-back inside the timed
region;  roofline = the ViT-block GEMM launches of the tcgen05 kernel, timed per launch with CUDA
events on the launching stream in an instrumented pass;  cpu_baseline = the oracle (the
reference's CPU PyTorch arithmetic, fp32) timed on the host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

IMG = 384
GFLOP_PER_IMAGE = 255.23          # algorithmic, reference formulation
VIT_GEMM_GFLOP_PER_IMAGE = 2 * 49.007  # 12 x (qkv + proj + fc1 + fc2) at 577 tokens
METRIC = "384x384 images/sec (DPT-Hybrid-384 depth forward)"


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"],
                "tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def pick_threads() -> int:
    """All host threads the CPU path can USE: intra-op oversubscription on a 100+-core box makes the
    convolutions slower, so time one image at a few thread counts and keep the fastest."""
    import torch
    from oracle import dpt_oracle, make_golden, weights
    cores = os.cpu_count() or 1
    cands = sorted({min(cores, c) for c in (16, 32, 64, cores)})
    sd = weights.make_state_dict(0, 1)
    x = make_golden.golden_input(1, seed=0)
    best, best_t = cands[0], float("inf")
    with torch.no_grad():
        for c in cands:
            torch.set_num_threads(c)
            dpt_oracle.forward_fp32(sd, x)
            t0 = time.perf_counter()
            dpt_oracle.forward_fp32(sd, x)
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = c, dt
    return best


def cpu_forward_timer(batch: int, reps: int, threads: int):
    """Times the oracle (reference CPU PyTorch arithmetic, fp32) on `batch` images; returns img/s."""
    import torch
    from oracle import dpt_oracle, make_golden, weights
    torch.set_num_threads(threads)
    sd = weights.make_state_dict(0, 1)
    x = make_golden.golden_input(batch, seed=0)
    with torch.no_grad():
        dpt_oracle.forward_fp32(sd, x)                      # warm-up
        best = float("inf")
        for _ in range(reps):
            t0 = time.perf_counter()
            dpt_oracle.forward_fp32(sd, x)
            best = min(best, time.perf_counter() - t0)
    return batch / best, best


def cpu_model_name():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference_arm(args, rank: int, world: int):
    """--impl reference: the reference's CPU implementation of the path on the host cores.
    /root/reference is absent on the GPU box and the reference is pure Python over the un-vendored
    timm, so the oracle port (bit-identical to the reference module in the build container,
    tests/test_oracle_cpu.py) is what runs; kind = "port"."""
    if rank != 0:
        return
    import torch
    from oracle import dpt_oracle, make_golden, weights
    cores = pick_threads()
    torch.set_num_threads(cores)
    batch = args.cpu_batch
    sd = weights.make_state_dict(0, 1)
    x = make_golden.golden_input(batch, seed=0)
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 2))):
            dpt_oracle.forward_fp32(sd, x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            dpt_oracle.forward_fp32(sd, x)
        dt = time.perf_counter() - t0
    value = batch * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"DPT-Hybrid-384 depth forward, reference CPU PyTorch arithmetic, {batch} images/step "
                               f"(bounded sample of configs[1])", "global_batch": batch},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "host_cores": os.cpu_count(),
                         "kind": "port", "cpu": cpu_model_name(), "sample": f"{args.steps} steps x {batch} images, fp32, "
                                                             f"torch.set_num_threads({cores})"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def classify_gemm(info: dict, batch: int, ntok: int) -> str:
    m, n, k = info["m"], info["n"], info["k"]
    if m == batch * ntok and (k, n) in ((768, 2304), (768, 768), (768, 3072), (3072, 768)):
        return "vit_gemm"
    return "other_gemm"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step (configs[1]: 32)")
    ap.add_argument("--cpu-batch", type=int, default=2, help="images per CPU step (reference arm / cpu_baseline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    # keep stdout to the one JSON line: NCCL prints its version banner there when NCCL_DEBUG=VERSION
    os.environ["NCCL_DEBUG"] = os.environ.get("ODB_NCCL_DEBUG", "WARN")
    import torch
    from omnidata_b200 import _capi, ops, parallel
    from omnidata_b200.model import DPTDepthModel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the native arm has no CPU fallback")
    rank, world, local = parallel.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.batch
    peaks = load_peaks()

    # ---- model: rank 0 owns the seeded weights, the others receive them over NCCL
    from omnidata_b200 import synthetic
    model = DPTDepthModel(backbone="vitb_rn50_384")
    if rank == 0:
        model.load_state_dict(synthetic.make_state_dict(0, 1), strict=True)
    model = model.to(dev).eval()
    t0 = time.perf_counter()
    bcast_bytes = parallel.broadcast_state_dict(model, src=0)
    torch.cuda.synchronize()
    bcast_ms = 1e3 * (time.perf_counter() - t0)
    model._invalidate()
    model.use_cuda_graph = not args.no_graph

    # ---- synthetic inputs: 4 distinct batches rotate so that no step re-reads a hot input; the
    # activations streamed per step (several GB) exceed the 126 MB L2 many times over anyway.
    gen = torch.Generator(device="cpu").manual_seed(1000 + rank)
    host_inputs = [(torch.rand(B, 3, IMG, IMG, generator=gen) * 2 - 1).pin_memory() for _ in range(4)]
    dev_inputs = [h.to(dev) for h in host_inputs]

    with torch.no_grad():
        # launches per forward (eager, counted by the library itself)
        model.use_cuda_graph = False
        n0 = _capi.launch_count()
        model(dev_inputs[0])
        torch.cuda.synchronize()
        launches_per_fwd = _capi.launch_count() - n0
        model.use_cuda_graph = not args.no_graph

        for i in range(args.warmup):
            model(dev_inputs[i % 4])
        torch.cuda.synchronize()

        # ---------------- timed region 1: device-resident inputs
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        parallel.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            y = model(dev_inputs[i % 4])
        e1.record()
        torch.cuda.synchronize()
        parallel.barrier()
        ms = parallel.reduce_max(e0.elapsed_time(e1), dev)
        clocks = sampler.stop() if rank == 0 else None

        # ---------------- timed region 2: end to end through the public API, host buffers
        # StreamingPredictor.run = DPTDepthModel.forward per batch, with the pinned-host -> device copy
        # of the next batch and the device -> host read of the previous depth maps on copy streams.
        from omnidata_b200.pipeline import StreamingPredictor
        predictor = StreamingPredictor(model, dev)
        host_outs = [torch.empty((B, IMG, IMG), dtype=torch.float32).pin_memory() for _ in range(2)]
        predictor.run((host_inputs[i % 4] for i in range(3)), host_outs)
        torch.cuda.synchronize()
        parallel.barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        predictor.run((host_inputs[i % 4] for i in range(args.steps)), host_outs)
        e3.record()
        torch.cuda.synchronize()
        parallel.barrier()
        ms_e2e = parallel.reduce_max(e2.elapsed_time(e3), dev)

        # ---------------- instrumented pass: per-launch CUDA events (roofline)
        roof = None
        roof_hbm = None
        detail = {}
        if rank == 0:
            model.use_cuda_graph = False
            with ops.LaunchTimer() as lt:
                for i in range(3):
                    model(dev_inputs[i % 4])
            recs = lt.results()
            per_fwd = len(recs) // 3
            recs = recs[per_fwd:]                                        # drop the first pass
            ntok = (IMG // 16) ** 2 + 1
            agg = {}
            for name, info, t_ms in recs:
                key = name
                if name == "odb_conv_gemm":
                    key = classify_gemm(info, B, ntok)
                a = agg.setdefault(key, {"ms": 0.0, "launches": 0, "flops": 0.0, "bytes": 0.0})
                a["ms"] += t_ms / 2
                a["launches"] += 0.5
                if name == "odb_conv_gemm":
                    a["flops"] += 2.0 * info["m"] * info["n"] * info["k"] / 2
                a["flops"] += info.get("flops", 0.0) / 2
                a["bytes"] += info.get("bytes", 0.0) / 2
            total_ms = sum(a["ms"] for a in agg.values())
            for k, a in agg.items():
                d = {"ms_per_step": round(a["ms"], 4), "launches": int(a["launches"]),
                     "share": round(a["ms"] / total_ms, 4)}
                if a["flops"]:
                    d["tflops"] = round(a["flops"] / (a["ms"] * 1e-3) / 1e12, 2)
                if a["bytes"]:
                    d["gbs"] = round(a["bytes"] / (a["ms"] * 1e-3) / 1e9, 1)
                detail[k] = d
            v = agg.get("vit_gemm")
            if v:
                achieved = VIT_GEMM_GFLOP_PER_IMAGE * 1e9 * B / (v["ms"] * 1e-3) / 1e12
                peak = peaks["tflops_sustained"]
                roof = {"kernel": "conv_gemm_kernel (tcgen05) — the 48 ViT-block GEMM launches",
                        "bound": "tensor", "achieved": round(achieved, 2), "peak": peak, "unit": "TFLOP/s",
                        "frac": round(achieved / peak, 4),
                        # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the four ViT GEMM
                        # shapes, from the committed `ncu --set full` capture (profiles/r01d_vit_gemm_*.csv);
                        # algorithmic bytes per launch (operands + output + residual, bf16): 131 MB
                        "traffic": 92.8e6, "traffic_unit": "bytes/launch (ncu, profiles/r01d)",
                        "peak_source": f"{peaks['source']} sustained bf16 (MEASURED_PEAKS.json)",
                        "avg_launch_ms": round(v["ms"] / v["launches"], 4),
                        "how": "CUDA events around every launch on the launching stream, eager instrumented pass "
                               "after the timed region (2 forwards averaged)"}
            u = agg.get("odb_upsample2x_add")
            if u and u["bytes"]:
                gbs = u["bytes"] / (u["ms"] * 1e-3) / 1e9
                roof_hbm = {"kernel": "upsample2x_add_kernel - bilinear x2 (+skip add, +relu copy), the HBM-bound "
                                      "kernel of the FeatureFusionBlock decoder and head (5 launches)",
                            "bound": "hbm", "achieved": round(gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": round(gbs / peaks["hbm_gbs"], 4),
                            "traffic": 1.45e9, "traffic_unit": "bytes, largest launch (ncu, profiles/r01c): equals its "
                                                               "algorithmic 0.30 GB read + 1.21 GB write",
                            "peak_source": f"{peaks['source']} copy bandwidth (MEASURED_PEAKS.json)"}
            model.use_cuda_graph = not args.no_graph

    images = B * world * args.steps
    value = images / (ms * 1e-3)
    e2e_value = images / (ms_e2e * 1e-3)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "configs[1]: DPT-Hybrid-384 depth head, bf16, batch 32 per GPU, synthetic 384x384 RGB",
                       "global_batch": B * world, "per_gpu_batch": B, "parallelism": f"dp{world} (independent images)",
                       "cuda_graph": not args.no_graph,
                       "l2": "4 rotating input batches; per-step activation traffic >> 126 MB L2",
                       "weights": "seeded synthetic (no checkpoint offline), NCCL-broadcast from rank 0"},
            "e2e": {"value": round(e2e_value, 2), "unit": "images/s", "ms_per_step": round(ms_e2e / args.steps, 4),
                    "h2d_bytes_per_step": B * 3 * IMG * IMG * 4, "d2h_bytes_per_step": B * IMG * IMG * 4},
            "gpu_launches": int(launches_per_fwd * args.steps),
            "launches_per_step": int(launches_per_fwd),
            "clocks": clocks,
            "model_tflops": round(GFLOP_PER_IMAGE * 1e9 * value / 1e12, 2),
            "model_frac_of_sustained_peak": round(GFLOP_PER_IMAGE * 1e9 * value / world / 1e12 / peaks["tflops_sustained"], 4),
            "weight_broadcast": {"bytes": bcast_bytes, "ms": round(bcast_ms, 2)},
            "roofline": roof,
            "roofline_hbm": roof_hbm,
            "roofline_detail": detail,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = pick_threads()
            v, secs = cpu_forward_timer(args.cpu_batch, reps=3, threads=cores)
            line["cpu_baseline"] = {"value": round(v, 3), "unit": "images/s", "cores": cores,
                                    "host_cores": os.cpu_count(), "kind": "port",
                                    "cpu": cpu_model_name(),
                                    "sample": f"oracle fp32 forward (bit-identical to the reference module in the build "
                                              f"container), {args.cpu_batch} images, best of 3 after warm-up, "
                                              f"{secs:.2f} s per pass"}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
