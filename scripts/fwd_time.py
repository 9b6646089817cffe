import sys, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sde4mbrl_b200.engine import Engine, PackedParams
w = bench.build_workload()
eng = Engine(w["model"]); pk = PackedParams(w["model"], w["nn"], eng.desc)
dev = torch.device("cuda")
ts = np.full(100, 0.01, np.float32)
y0 = torch.as_tensor(w["y0"], device=dev)[None]; opt = torch.as_tensor(w["opt"], device=dev)
key = torch.as_tensor(w["key"].view(np.int32), device=dev)
out = torch.empty((101, 13, 4096), device=dev)
for lanes in (1, 4, 8):
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.rollout(pk, y0, opt, ts, key=key, N=4096, lanes=lanes, out=out); e1.record(); torch.cuda.synchronize()
    ts_ = []
    for rep in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.rollout(pk, y0, opt, ts, key=key, N=4096, lanes=lanes, out=out); e1.record(); torch.cuda.synchronize()
        ts_.append(e0.elapsed_time(e1))
    print("lanes", lanes, "fwd ms", np.mean(ts_), np.min(ts_))
