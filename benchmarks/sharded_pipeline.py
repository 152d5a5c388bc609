"""Where the host-buffer time of the row-sharded search goes: device-resident steps, blocking search_host,
pipelined search_host_submit/_wait at depth 2 / 3, and the bare copies.  Run under torchrun (any world size) or
plainly (world 1).  Prints one JSON line on rank 0."""
import json
import os
import sys
import time
from collections import deque

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from cmw_rag_b200.engine import pinned_empty  # noqa: E402
from cmw_rag_b200.sharded import PeerGather, ShardedSearcher, shard_bounds  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
rows, dim, B, k, steps = int(os.environ.get("ROWS", "1000000")), 1536, 4096, 100, int(os.environ.get("STEPS", "20"))
lo, hi = shard_bounds(rows, world)[rank]
st, first = bench.build_store(torch, dim, dev, lo, hi)
q, _ = bench.make_queries(torch, dist, first, lo, B, dim, dev, 7, rank, world)
q_host = pinned_empty((B, dim), np.float32)
q_host[:] = q.cpu().numpy()
pg = PeerGather(device=local, max_batch=B, max_k=k) if world > 1 and os.environ.get("EXCHANGE", "gather") == "gather" else None
s = ShardedSearcher(st, gather=pg)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def outs(n):
    return [(pinned_empty((B, k), np.float32), pinned_empty((B, k), np.int64), pinned_empty((B,), np.int32))
            for _ in range(n)]


res = {}
for _ in range(5):
    s.search(q, k)
o = outs(3)
for i in range(3):
    s.search_host(q_host, k, out=o[i])
barrier()
t0 = time.perf_counter()
for _ in range(steps):
    s.search(q, k)
barrier()
res["device_ms"] = (time.perf_counter() - t0) * 1e3 / steps
t0 = time.perf_counter()
for _ in range(steps):
    s.search_host(q_host, k, out=o[0])
barrier()
res["blocking_ms"] = (time.perf_counter() - t0) * 1e3 / steps
for depth in (2, 3):
    for t in [s.search_host_submit(q_host, k, out=o[i]) for i in range(depth)]:
        s.search_host_wait(t)
    pend = deque()
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        if len(pend) == depth:
            s.search_host_wait(pend.popleft())
        pend.append(s.search_host_submit(q_host, k, out=o[i % depth]))
    while pend:
        s.search_host_wait(pend.popleft())
    barrier()
    res[f"pipelined_depth{depth}_ms"] = (time.perf_counter() - t0) * 1e3 / steps
# the bare copies
qd = torch.empty((B, dim), device=dev)
barrier()
t0 = time.perf_counter()
for _ in range(steps):
    qd.copy_(torch.from_numpy(q_host), non_blocking=True)
torch.cuda.synchronize(dev)
res["h2d_ms"] = (time.perf_counter() - t0) * 1e3 / steps
res["h2d_pinned"] = bool(torch.from_numpy(q_host).is_pinned())
ms, mi, fl = s.search(q, k)
barrier()
t0 = time.perf_counter()
for _ in range(steps):
    torch.from_numpy(o[0][0]).copy_(ms, non_blocking=True)
    torch.from_numpy(o[0][1]).copy_(mi, non_blocking=True)
torch.cuda.synchronize(dev)
res["d2h_ms"] = (time.perf_counter() - t0) * 1e3 / steps
# host time of one submit (enqueue only)
barrier()
t0 = time.perf_counter()
t = s.search_host_submit(q_host, k, out=o[0])
res["submit_host_ms"] = (time.perf_counter() - t0) * 1e3
s.search_host_wait(t)
if world > 1:
    v = torch.tensor([res[n] for n in sorted(res) if n != "h2d_pinned"], dtype=torch.float64, device=dev)
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    for n, x in zip([n for n in sorted(res) if n != "h2d_pinned"], v.tolist()):
        res[n] = x
if rank == 0:
    res["world"] = world
    print(json.dumps(res), flush=True)
if pg is not None:
    pg.close()
st.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
