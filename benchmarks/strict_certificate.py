import sys, json, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/benchmarks')
import torch, bench
from cmw_rag_b200 import _native as N
dev=torch.device("cuda:0")
class A: pass
a=A(); a.rows,a.dim,a.shard,a.no_f32=1_000_000,1536,"queries",False
st,first=bench.build_store(torch,a,dev,0,1)
q,needle=bench.make_queries(torch,first,4096,1536,dev,7)
for eps,kp in ((5e-4,0),(3.9e-3,512),(3.9e-3,768)):
    N.set_option("bf16_eps",eps); N.set_option("kprime",kp)
    for _ in range(3): sc,ids,fl=st.search(q,100)
    torch.cuda.synchronize()
    N.profile_enable(True)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(6): st.search(q,100)
    e1.record(); torch.cuda.synchronize()
    prof=N.profile_read(); N.profile_enable(False)
    print(json.dumps({"bf16_eps":eps,"kprime":kp,"ms":e0.elapsed_time(e1)/6,"uncertified":int(fl.sum()),"phases":{k:round(v[0]/6,3) for k,v in prof.items()}}))
