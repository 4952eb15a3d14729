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
