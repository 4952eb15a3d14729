"""DPT-Large (backbone 'vitl16_384') on one GPU: images/s (CUDA graph, CUDA events) and the per-family launch
times of one eager forward.   python profiles/large.py [batch] > gpurun_out/large.txt"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from omnidata_b200 import ops, synthetic  # noqa: E402
from omnidata_b200.model import DPTDepthModel, state_dict_spec  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    for backbone, gflop in (("vitl16_384", None), ("vitb16_384", None)):
        model = DPTDepthModel(backbone=backbone)
        model.load_state_dict(synthetic.make_state_dict(0, 1, spec=state_dict_spec(1, backbone=backbone)))
        model = model.cuda().eval()
        x = torch.rand(batch, 3, 384, 384, device="cuda") * 2 - 1
        with torch.no_grad():
            model.use_cuda_graph = True
            for _ in range(3):
                model(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                model(x)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            model.use_cuda_graph = False
            model(x)
            with ops.LaunchTimer() as lt:
                model(x)
            recs = lt.results()
        agg, flops = {}, 0.0
        for name, info, t in recs:
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += t
            if name == "odb_conv_gemm":
                flops += 2.0 * info["m"] * info["n"] * info["k"]
            flops += info.get("flops", 0.0)
        print(f"{backbone}: batch {batch}: {ms:.3f} ms / step = {batch / ms * 1e3:.1f} images/s; {len(recs)} launches; "
              f"{flops / batch / 1e9:.1f} GFLOP/image executed = {flops / ms / 1e9:.0f} TFLOP/s")
        tot = sum(v[1] for v in agg.values())
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"   {k:28s} {n:4d} launches {t:8.3f} ms  {100 * t / tot:5.1f} %")
        del model
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
