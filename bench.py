#!/usr/bin/env python
"""Throughput harness for the DPT-Hybrid-384 hot path (BASELINE.json metric: 384x384 images/sec).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference CPU path

A "step" is one forward pass over one batch of synthetic 384x384 RGB images (`--config` picks the BASELINE.json
configuration: 1 = depth b32 (default, the metric's own), 2 = dual depth+normal b64, 3 = 512 images strong scaling,
4 = the train_depth.py step).
N=1 workload = BASELINE.json configs[1] (depth head, bf16, batch 32, one B200).  For N>1 (launched
by torch.distributed.run, one rank per GPU) every rank runs the same per-GPU batch on its own
images — independent units, no data-path collective ("weak" scaling); weights are NCCL-broadcast
from rank 0 once, timings are reduced with MAX over ranks.

One JSON line on stdout (rank 0):  value = whole-job images/s with inputs resident in HBM (CUDA
events, device time, max over ranks);  e2e = the same metric through the public API
(`DPTDepthModel.forward`) with pinned-host inputs, H2D copy and D2H read-back inside the timed
region;  roofline = ALL launches of the dominant kernel (`conv_gemm_kernel`, the tcgen05 implicit GEMM), timed per
launch with CUDA events on the launching stream in an instrumented eager pass (`roofline_vit_blocks`: its ViT-block
subset);  cpu_baseline = the oracle (the reference's CPU PyTorch arithmetic, fp32) timed on the host cores on a
bounded sample;  gpu_eager_baseline = the same arithmetic in eager torch on this GPU (fp32 and autocast bf16).
`--config 4` (omnidata_b200/train_bench.py) times the whole train step, replayed as one CUDA graph on one GPU
(`--no-graph`: the eager launch sequence).
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
    tests/test_oracle_cpu.py) is what runs; kind = "port".  Fixed policy: batch 8 per step,
    torch.set_num_threads(min(host cores, 32)) — on the 128-thread hosts of this pool more intra-op threads are
    slower — K timed steps after W warm-up steps; the best single step is reported beside the mean."""
    if rank != 0:
        return
    import torch
    from oracle import dpt_oracle, make_golden, weights
    cores = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(cores)
    batch = args.cpu_batch
    dual = args.config == 2
    sds = [weights.make_state_dict(0, 1)] + ([weights.make_state_dict(0, 3)] if dual else [])
    x = make_golden.golden_input(batch, seed=0)
    xs = [x] + ([(x + 1) / 2] if dual else [])
    times = []
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 2))):
            for sd, xi in zip(sds, xs):
                dpt_oracle.forward_fp32(sd, xi)
        for _ in range(args.steps):
            t0 = time.perf_counter()
            for sd, xi in zip(sds, xs):
                dpt_oracle.forward_fp32(sd, xi)
            times.append(time.perf_counter() - t0)
    dt = sum(times)
    value = batch * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{WORKLOADS[args.config]} — reference CPU PyTorch arithmetic, {batch} images/step "
                               f"(bounded sample)", "global_batch": batch},
        "cpu_baseline": {"value": value, "best_step_value": batch / min(times), "unit": "images/s", "cores": cores,
                         "host_cores": os.cpu_count(), "kind": "port", "cpu": cpu_model_name(),
                         "sample": f"{args.steps} steps x {batch} images, fp32, torch.set_num_threads({cores})"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


WORKLOADS = {
    1: "configs[1]: DPT-Hybrid-384 depth head, bf16, batch 32 per GPU, synthetic 384x384 RGB",
    2: "configs[2]: DPT-Hybrid-384 dual depth + normal networks on the same images, bf16, batch 64 per GPU",
    3: "configs[3]: DPT-Hybrid-384 depth, global batch 512 sharded 512/N per GPU (strong scaling)",
    4: "configs[4]: train_depth.py step — DPT forward + MiDaS SSI + virtual-normal loss + backward + gradient "
       "all-reduce + clip + Adam, batch 16 per GPU (128 at 8 GPUs)",
}


def classify_gemm(info: dict, batch: int, ntok: int) -> str:
    m, n, k = info["m"], info["n"], info["k"]
    if m == batch * ntok and (k, n) in ((768, 2304), (768, 768), (768, 3072), (3072, 768)):
        return "vit_gemm"
    return "other_gemm"


def gpu_eager_baseline(batch: int, dev):
    """The GPU yardstick BASELINE.md 3.5 asks for: the reference network (oracle restatement = the reference module's
    arithmetic, torch library kernels: cuDNN / cuBLAS) in eager PyTorch on the SAME B200, fp32 (TF32 off) and under
    torch.autocast(bf16).  Baseline leg only — never on the product path."""
    import torch
    from oracle import dpt_oracle, make_golden, weights
    sd = {k: v.to(dev) for k, v in weights.make_state_dict(0, 1).items()}
    x = make_golden.golden_input(batch, seed=3).to(dev)
    out = {}
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        for name in ("fp32", "autocast_bf16"):
            def run():
                if name == "fp32":
                    return dpt_oracle.forward_fp32(sd, x)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return dpt_oracle.forward_fp32(sd, x)
            with torch.no_grad():
                run(); run()
                torch.cuda.synchronize()
                best = float("inf")
                for _ in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); run(); e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
            out[name] = {"images_per_s": round(batch / (best * 1e-3), 1), "ms_per_step": round(best, 3)}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    out["what"] = (f"reference network in eager torch {torch.__version__} (cuDNN/cuBLAS library kernels) on this GPU, "
                   f"batch {batch}, device-resident input, best of 3 after 2 warm-up passes, CUDA events")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4],
                    help="BASELINE.json configs index: 1 depth b32 (default, the metric's config), 2 dual depth+normal "
                         "b64, 3 global batch 512 strong scaling, 4 train step")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (0 = the config's own)")
    ap.add_argument("--cpu-batch", type=int, default=8, help="images per CPU step (reference arm / cpu_baseline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    # stdout carries exactly one JSON line: whatever NCCL_DEBUG level the caller asked for goes to stderr
    if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
    if args.config == 4:
        from omnidata_b200 import train_bench
        train_bench.main(args)
        return
    import torch
    from omnidata_b200 import _capi, ops, parallel
    from omnidata_b200.model import DPTDepthModel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the native arm has no CPU fallback")
    rank, world, local = parallel.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dual = args.config == 2
    strong = args.config == 3
    if args.batch:
        B = args.batch
    elif strong:
        if 512 % world:
            raise SystemExit("config 3: 512 images must divide over the ranks")
        B = 512 // world                      # images per GPU per step (the whole global batch is one step)
    else:
        B = 64 if dual else 32
    chunk = min(B, 128) if strong else B      # config 3: the per-GPU share runs as forwards of <= 128 images
    if B % chunk:
        raise SystemExit("per-GPU batch must be a multiple of the chunk size")
    n_chunks = B // chunk
    peaks = load_peaks()

    # ---- models: rank 0 owns the seeded weights and packs them; the PACKED (bf16) form is NCCL-broadcast
    from omnidata_b200 import synthetic
    chans = (1, 3) if dual else (1,)
    models = []
    bcast_bytes, bcast_ms = 0, 0.0
    for c in chans:
        m = DPTDepthModel(backbone="vitb_rn50_384", num_channels=c)
        if rank == 0:
            m.load_state_dict(synthetic.make_state_dict(0, c), strict=True)
        m = m.to(dev).eval()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bcast_bytes += parallel.broadcast_packed_weights(m, dev, src=0)
        torch.cuda.synchronize()
        bcast_ms += 1e3 * (time.perf_counter() - t0)
        weights_identical = parallel.packed_weights_identical(m, dev) and (weights_identical if models else True)
        m.use_cuda_graph = not args.no_graph
        models.append(m)

    def forward_all(x):
        """one chunk of images through every network of the config (config 2: depth on [-1,1], normal on [0,1])."""
        outs = [models[0](x)]
        if dual:
            outs.append(models[1](x * 0.5 + 0.5))      # demo.py:74-76 / 92-95: the normal net takes RGB in [0,1]
        return outs

    # ---- synthetic inputs: 4 distinct batches rotate so that no step re-reads a hot input; the
    # activations streamed per step (several GB) exceed the 126 MB L2 many times over anyway.
    gen = torch.Generator(device="cpu").manual_seed(1000 + rank)
    n_rot = 4 if chunk <= 64 else 2
    host_inputs = [(torch.rand(chunk, 3, IMG, IMG, generator=gen) * 2 - 1).pin_memory() for _ in range(n_rot)]
    dev_inputs = [h.to(dev) for h in host_inputs]

    with torch.no_grad():
        # launches per forward (eager, counted by the library itself)
        for m in models:
            m.use_cuda_graph = False
        n0 = _capi.launch_count()
        forward_all(dev_inputs[0])
        torch.cuda.synchronize()
        launches_per_chunk = _capi.launch_count() - n0
        for m in models:
            m.use_cuda_graph = not args.no_graph

        for i in range(args.warmup):
            forward_all(dev_inputs[i % n_rot])
        torch.cuda.synchronize()

        # ---------------- timed region 1: device-resident inputs
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        parallel.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps * n_chunks):
            y = forward_all(dev_inputs[i % n_rot])
        e1.record()
        torch.cuda.synchronize()
        parallel.barrier()
        ms = parallel.reduce_max(e0.elapsed_time(e1), dev)
        clocks = sampler.stop() if rank == 0 else None

        # ---------------- timed region 2: end to end through the public API, host buffers
        # StreamingPredictor.run = DPTDepthModel.forward per batch, with the pinned-host -> device copy
        # of the next batch and the device -> host read of the previous outputs on copy streams.
        from omnidata_b200.pipeline import StreamingPredictor
        predictor = StreamingPredictor(forward_all, dev)
        out_shapes = [(chunk, IMG, IMG)] + ([(chunk, 3, IMG, IMG)] if dual else [])
        host_outs = [[torch.empty(s, dtype=torch.float32).pin_memory() for s in out_shapes] for _ in range(2)]
        predictor.run((host_inputs[i % n_rot] for i in range(3)), host_outs)
        torch.cuda.synchronize()
        parallel.barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        predictor.run((host_inputs[i % n_rot] for i in range(args.steps * n_chunks)), host_outs)
        e3.record()
        torch.cuda.synchronize()
        parallel.barrier()
        ms_e2e = parallel.reduce_max(e2.elapsed_time(e3), dev)

        # ---------------- instrumented pass: per-launch CUDA events (roofline)
        roof = roof_vit = roof_hbm = None
        detail = {}
        if rank == 0:
            for m in models:
                m.use_cuda_graph = False
            with ops.LaunchTimer() as lt:
                for i in range(3):
                    forward_all(dev_inputs[i % n_rot])
            recs = lt.results()
            per_fwd = len(recs) // 3
            recs = recs[per_fwd:]                                        # drop the first pass
            ntok = (IMG // 16) ** 2 + 1
            agg = {}
            for name, info, t_ms in recs:
                keys = [name]
                if name == "odb_conv_gemm":
                    keys = ["conv_gemm_all", classify_gemm(info, chunk, ntok)]
                for key in keys:
                    a = agg.setdefault(key, {"ms": 0.0, "launches": 0, "flops": 0.0, "bytes": 0.0})
                    a["ms"] += t_ms / 2
                    a["launches"] += 0.5
                    if name == "odb_conv_gemm":
                        a["flops"] += 2.0 * info["m"] * info["n"] * info["k"] / 2
                    a["flops"] += info.get("flops", 0.0) / 2
                    a["bytes"] += info.get("bytes", 0.0) / 2
            total_ms = sum(a["ms"] for k, a in agg.items() if k != "conv_gemm_all")
            for k, a in agg.items():
                d = {"ms_per_step": round(a["ms"], 4), "launches": int(a["launches"]),
                     "share": round(a["ms"] / total_ms, 4)}
                if a["flops"]:
                    d["tflops"] = round(a["flops"] / (a["ms"] * 1e-3) / 1e12, 2)
                if a["bytes"]:
                    d["gbs"] = round(a["bytes"] / (a["ms"] * 1e-3) / 1e9, 1)
                detail[k] = d
            peak = peaks["tflops_sustained"]
            how = ("CUDA events around every launch on the launching stream, eager instrumented pass after the timed "
                   "region (2 forwards averaged)")
            g = agg.get("conv_gemm_all")
            if g:
                achieved = g["flops"] / (g["ms"] * 1e-3) / 1e12
                roof = {"kernel": f"conv_gemm_kernel (tcgen05 implicit GEMM) — ALL {int(g['launches'])} launches of a "
                                  f"forward: every conv / linear layer of the network",
                        "bound": "tensor", "achieved": round(achieved, 2), "peak": peak, "unit": "TFLOP/s",
                        "frac": round(achieved / peak, 4),
                        "algorithmic_gflop_per_step": round(g["flops"] / 1e9, 1),
                        "avg_launch_ms": round(g["ms"] / g["launches"], 4), "share_of_step": detail["conv_gemm_all"]["share"],
                        "traffic": NCU_TRAFFIC.get("conv_gemm_all"), "traffic_unit": NCU_TRAFFIC.get("unit"),
                        "peak_source": f"{peaks['source']} sustained bf16 (MEASURED_PEAKS.json)", "how": how}
            v = agg.get("vit_gemm")
            if v:
                achieved = VIT_GEMM_GFLOP_PER_IMAGE * 1e9 * chunk * len(models) / (v["ms"] * 1e-3) / 1e12
                roof_vit = {"kernel": "conv_gemm_kernel — the 48 ViT-block GEMM launches only (qkv, proj, fc1, fc2)",
                            "bound": "tensor", "achieved": round(achieved, 2), "peak": peak, "unit": "TFLOP/s",
                            "frac": round(achieved / peak, 4), "avg_launch_ms": round(v["ms"] / v["launches"], 4)}
            u = agg.get("odb_upsample2x_add")
            if u and u["bytes"]:
                gbs = u["bytes"] / (u["ms"] * 1e-3) / 1e9
                roof_hbm = {"kernel": "upsample2x_add_kernel - bilinear x2 (+skip add, +relu copy), the HBM-bound "
                                      "kernel of the FeatureFusionBlock decoder and head (5 launches)",
                            "bound": "hbm", "achieved": round(gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": round(gbs / peaks["hbm_gbs"], 4),
                            "peak_source": f"{peaks['source']} copy bandwidth (MEASURED_PEAKS.json)"}
            for m in models:
                m.use_cuda_graph = not args.no_graph

    images = B * world * args.steps
    value = images / (ms * 1e-3)
    e2e_value = images / (ms_e2e * 1e-3)

    if rank == 0:
        nets = len(models)
        line = {
            "metric": METRIC if not dual else "384x384 images/sec (DPT-Hybrid-384 depth + normal forward)",
            "value": round(value, 2), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.config], "index": args.config,
                       "global_batch": B * world, "per_gpu_batch": B,
                       "forward_chunk": chunk, "chunks_per_step": n_chunks, "networks": nets,
                       "parallelism": f"dp{world} (independent images)",
                       "cuda_graph": not args.no_graph,
                       "l2": f"{n_rot} rotating input batches; per-step activation traffic >> 126 MB L2",
                       "residual_stream": "fp32 (ViT tokens), other activations bf16",
                       "weights": "seeded synthetic (no checkpoint offline); packed bf16 form NCCL-broadcast from rank 0"},
            "e2e": {"value": round(e2e_value, 2), "unit": "images/s", "ms_per_step": round(ms_e2e / args.steps, 4),
                    "h2d_bytes_per_step": B * 3 * IMG * IMG * 4,
                    "d2h_bytes_per_step": B * IMG * IMG * 4 * (4 if dual else 1)},
            "gpu_launches": int(launches_per_chunk * n_chunks * args.steps),
            "launches_per_step": int(launches_per_chunk * n_chunks),
            "clocks": clocks,
            "model_tflops": round(GFLOP_PER_IMAGE * nets * 1e9 * value / 1e12, 2),
            "model_frac_of_sustained_peak": round(GFLOP_PER_IMAGE * nets * 1e9 * value / world / 1e12 / peaks["tflops_sustained"], 4),
            "weight_broadcast": {"bytes": bcast_bytes, "ms": round(bcast_ms, 2), "what": "packed bf16 kernel operands",
                                 "identical_on_all_ranks": weights_identical},
            "roofline": roof,
            "roofline_vit_blocks": roof_vit,
            "roofline_hbm": roof_hbm,
            "roofline_detail": detail,
        }
        if world == 1 and not args.no_gpu_baseline and args.config == 1:
            torch.cuda.empty_cache()
            line["gpu_eager_baseline"] = gpu_eager_baseline(B, dev)
        if world == 1 and not args.no_cpu_baseline:
            cores = min(os.cpu_count() or 1, 32)
            v, secs = cpu_forward_timer(args.cpu_batch, reps=3, threads=cores)
            line["cpu_baseline"] = {"value": round(v / nets, 3), "unit": "images/s", "cores": cores,
                                    "host_cores": os.cpu_count(), "kind": "port",
                                    "cpu": cpu_model_name(),
                                    "sample": f"oracle fp32 forward (bit-identical to the reference module in the build "
                                              f"container), {args.cpu_batch} images, best of 3 after warm-up, "
                                              f"{secs:.2f} s per pass" + (" (one network timed, halved for two)" if dual else "")}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture of this round (profiles/r02_*):
# filled by hand from the capture (batch 32, configs[1])
NCU_TRAFFIC = {"unit": "bytes per launch = (dram__bytes_read.sum + dram__bytes_write.sum) summed over the 130 conv_gemm launches of "
                       "one batch-32 forward / 130 (profiles/r02_conv_gemm_ncu.csv: 9.96 GB read + 3.29 GB written)",
               "conv_gemm_all": 101.9e6}


if __name__ == "__main__":
    main()
