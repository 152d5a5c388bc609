"""torchrun worker: two ranks exchange top-k candidates through the fused NVLink peer-memory kernels (K5x), and run the
two-phase search with both of its exchanges over peer memory (cmw_peer_gather), each checked against the oracle.  Used by __graft_entry__.smoke() on boxes with two or more GPUs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import synth  # noqa: E402
from cmw_rag_b200 import DenseStore  # noqa: E402
from cmw_rag_b200.sharded import PeerExchange, PeerGather, ShardedSearcher, shard_bounds  # noqa: E402
from oracle.cport import exact_topk_c  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n, d, k = 8192, 256, 10
c = synth.make_corpus(n, d, seed=12)
q, _ = synth.make_queries(c, 16, seed=13)
lo, hi = shard_bounds(n, world)[rank]
st = DenseStore(d, hi - lo, device=local, id_offset=lo)
st.append(c[lo:hi])
ex = PeerExchange(device=local, max_batch=16, max_k=16)
ms, mi, fl = ShardedSearcher(st, exchange=ex).search(torch.from_numpy(q).to(dev), k)
torch.cuda.synchronize()
ref_ids, _, _ = exact_topk_c(c, q, k)
assert (mi.cpu().numpy() == ref_ids).all() and int(fl.sum()) == 0
ex.close()
pg = PeerGather(device=local, max_batch=16, max_k=16)
ms2, mi2, fl2 = ShardedSearcher(st, gather=pg).search(torch.from_numpy(q).to(dev), k)
torch.cuda.synchronize()
assert (mi2.cpu().numpy() == ref_ids).all() and int(fl2.sum()) == 0 and torch.equal(ms2, ms)
pg.close()
dist.barrier()
dist.destroy_process_group()
print("peer exchange ok", rank)
