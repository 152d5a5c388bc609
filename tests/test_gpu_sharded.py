"""Row-sharded search on real GPUs: ShardedSearcher over NCCL (all visible GPUs, or the world-size-1
degenerate case on a single-GPU box) must return the global oracle top-k on every rank."""
from __future__ import annotations

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
import synth
from oracle.cport import exact_topk_c
from cmw_rag_b200 import DenseStore
from cmw_rag_b200.sharded import PeerExchange, ShardedSearcher, shard_bounds
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
dev = torch.device(f"cuda:{{local}}"); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n, d, k = 30011, 256, 50
c = synth.make_corpus(n, d, seed=4)
q, _ = synth.make_queries(c, 37, seed=5)
lo, hi = shard_bounds(n, world)[rank]
st = DenseStore(d, max(1, hi - lo), device=local, id_offset=lo)
st.append(c[lo:hi])
s = ShardedSearcher(st)
qd = torch.from_numpy(q).to(dev)
ref_ids, ref_sc, _ = exact_topk_c(c, q, k)
for mode in ("f32", "bf16"):
    ms, mi, fl = s.search(qd, k, mode=mode)
    torch.cuda.synchronize()
    if mode == "f32":
        assert (mi.cpu().numpy() == ref_ids).all(), rank
        assert np.abs(ms.cpu().numpy() - ref_sc).max() <= 1e-5
        assert int(fl.sum()) == 0
    else:
        rec = np.mean([len(set(mi[b].tolist()) & set(ref_ids[b])) / k for b in range(37)])
        assert rec >= 0.95, rec
# the fused NVLink peer-memory exchange must give the same answer as all-gather + merge, call after call
ex = PeerExchange(device=local, max_batch=64, max_k=64)
fused = ShardedSearcher(st, exchange=ex)
for it in range(5):
    qi = torch.from_numpy(np.roll(q, it, axis=0).copy()).to(dev)
    a_s, a_i, _ = s.search(qi, k, mode="f32")
    b_s, b_i, _ = fused.search(qi, k, mode="f32")
    torch.cuda.synchronize()
    assert torch.equal(a_i, b_i) and torch.equal(a_s, b_s), (rank, it)
assert (b_i.cpu().numpy() == np.roll(ref_ids, 4, axis=0)).all()
ex.close()
dist.barrier(); dist.destroy_process_group()
print("sharded ok", rank, world)
"""


def test_sharded_searcher_nccl(tmp_path):
    import torch

    ngpu = torch.cuda.device_count()
    world = 2 if ngpu >= 2 else 1
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", "29547", str(script)],
        capture_output=True, text=True, timeout=600, env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert res.stdout.count("sharded ok") == world


def test_two_devices_in_one_process():
    """One process driving two GPUs (the C ABI selects the store's device per call; shared-memory
    attributes and TMA descriptors are per device)."""
    import numpy as np
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth
    from cmw_rag_b200 import DenseStore
    from oracle.cport import exact_topk_c

    c = synth.make_corpus(12000, 256, seed=2)
    q, _ = synth.make_queries(c, 40, seed=3)
    ref, ref_sc, _ = exact_topk_c(c, q, 10)
    stores = [DenseStore(256, 12000, device=dv) for dv in (0, 1)]
    for st in stores:
        st.append(c)
    for rep in range(2):
        for dv, st in enumerate(stores):
            sc, ids, fl = st.search_host(q, 10)
            assert (ids == ref).all() and (fl == 0).all(), dv
            qd = torch.from_numpy(q[:3]).to(f"cuda:{dv}")
            with torch.cuda.device(dv):
                sc2, ids2, _ = st.search(qd, 10, algo="scan")
                torch.cuda.synchronize(dv)
            assert (ids2.cpu().numpy() == ref[:3]).all()
    for st in stores:
        st.close()
