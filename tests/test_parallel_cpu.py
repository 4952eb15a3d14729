"""N>1 host logic on CPU: world_size-2 gloo — sharding, weight broadcast, max/sum reductions."""
import os
import socket
import sys
from pathlib import Path

import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, str(ROOT))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from omnidata_b200 import parallel
    r, w, _ = parallel.init_from_env("gloo")
    torch.manual_seed(100 + rank)                      # ranks start with DIFFERENT weights
    m = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.Conv2d(3, 5, 3))
    sent = parallel.broadcast_state_dict(m, src=0, bucket_bytes=1024)
    sig = float(sum(p.double().sum() for p in m.parameters()))
    lo, hi = parallel.shard_range(11, r, w)
    mx = parallel.reduce_max(10.0 + r, torch.device("cpu"))
    sm = parallel.reduce_sum(float(hi - lo), torch.device("cpu"))
    parallel.barrier()
    out.put((r, sig, (lo, hi), mx, sm, sent))
    torch.distributed.destroy_process_group()


def test_world2_gloo_broadcast_shard_reduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, sig0, rng0, mx0, sm0, sent0), (_, sig1, rng1, mx1, sm1, _) = res
    assert sig0 == sig1                                # weights identical after broadcast
    assert rng0 == (0, 6) and rng1 == (6, 11)          # contiguous, covers all 11 items
    assert mx0 == mx1 == 11.0 and sm0 == sm1 == 11.0
    assert sent0 > 0


def test_shard_range_properties():
    from omnidata_b200.parallel import shard_range
    for n in (0, 1, 7, 512):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_grad_bucket_plan_covers_the_flat_gradient_in_completion_order():
    """train step (configs[4]): the flat gradient is all-reduced in four contiguous ranges as the backward completes them."""
    from omnidata_b200.model import state_dict_spec
    from omnidata_b200.train import plan_grad_buckets
    import math
    spec = state_dict_spec(1)
    names = [k for k, _ in spec]
    sizes = [(math.prod(s) + 3) // 4 * 4 for _, s in spec]
    buckets = plan_grad_buckets(names, sizes)
    assert [t for _, _, t in buckets] == ["decoder", "vit_hi", "vit_lo", "resnet"]
    spans = sorted((s, e) for s, e, _ in buckets)
    assert spans[0][0] == 0 and spans[-1][1] == sum(sizes)
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    offs, off = {}, 0
    for n, s in zip(names, sizes):
        offs[n] = off
        off += s
    rng = {t: (s, e) for s, e, t in buckets}
    inside = lambda n, t: rng[t][0] <= offs[n] < rng[t][1]
    assert inside("scratch.output_conv.0.weight", "decoder") and inside("pretrained.act_postprocess3.3.weight", "decoder")
    assert inside("pretrained.model.blocks.11.mlp.fc2.weight", "vit_hi") and inside("pretrained.model.blocks.6.norm1.weight", "vit_hi")
    assert inside("pretrained.model.blocks.5.mlp.fc2.bias", "vit_lo") and inside("pretrained.model.patch_embed.proj.weight", "vit_lo")
    assert inside("pretrained.model.pos_embed", "resnet") and inside("pretrained.model.patch_embed.backbone.stem.conv.weight", "resnet")


def _bucket_worker(rank, world, port, out):
    sys.path.insert(0, str(ROOT))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    from omnidata_b200 import parallel
    from omnidata_b200.train import plan_grad_buckets
    parallel.init_from_env("gloo")
    names = ["pretrained.model.cls_token", "pretrained.model.patch_embed.backbone.stem.conv.weight",
             "pretrained.model.patch_embed.proj.weight", "pretrained.model.blocks.0.norm1.weight",
             "pretrained.model.blocks.6.norm1.weight", "pretrained.model.norm.weight", "scratch.output_conv.0.weight"]
    sizes = [8, 12, 16, 8, 8, 4, 20]
    flat = torch.arange(sum(sizes), dtype=torch.float32) * (rank + 1)
    for s, e, _ in plan_grad_buckets(names, sizes):          # data-parallel mean, bucket by bucket (gloo has no AVG)
        dist.all_reduce(flat[s:e], op=dist.ReduceOp.SUM)
        flat[s:e] /= world
    out.put((rank, flat.tolist()))            # plain lists: tensors in a Queue need the sender alive at receive time
    dist.destroy_process_group()


def test_world2_gloo_bucketed_gradient_mean():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = (torch.arange(76, dtype=torch.float32) * 1.5).tolist()
    assert res[0] == want and res[1] == want


def _packed_worker(rank, world, port, out):
    sys.path.insert(0, str(ROOT))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from omnidata_b200 import parallel
    from omnidata_b200.model import DPTDepthModel
    parallel.init_from_env("gloo")
    model = DPTDepthModel()
    if rank == 1:                                           # a different checkpoint on the receiving rank
        with torch.no_grad():
            for p in model.parameters():
                p.mul_(1.5).add_(0.01)
    sent = parallel.broadcast_packed_weights(model, torch.device("cpu"), src=0)
    sig = 0.0
    n_bf16 = 0
    for _, t in parallel._packed_tensors(model._packed):
        sig += float(t.double().abs().sum())
        n_bf16 += t.numel() if t.dtype == torch.bfloat16 else 0
    fresh = model._packed_sig == model._weights_signature()      # the staleness check will not re-pack the received weights
    out.put((rank, sig, sent, n_bf16, fresh))
    torch.distributed.destroy_process_group()


def test_world2_gloo_packed_bf16_weight_broadcast():
    """the PACKED kernel operands (bf16 GEMM weights + fp32 vectors, ~248 MB) travel, not the fp32 state_dict."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_packed_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, sig0, sent0, nb0, fresh0), (_, sig1, sent1, nb1, fresh1) = res
    assert sig0 == sig1 and sig0 > 0                       # rank 1 now holds rank 0's packed weights
    assert fresh0 and fresh1
    assert 240e6 < sent0 < 256e6 and sent0 == sent1        # about half of the 493 MB fp32 state_dict
    assert nb0 > 100e6
