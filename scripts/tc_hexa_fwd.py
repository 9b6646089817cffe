"""Hexacopter forward rollout, 4096 problems x 32 particles x 20 steps (the candidate evaluations of C4 / N = 32 without the
stage cost): tensor-core kernel vs the FP32 one-thread-per-sample kernel vs the lane choice of the library."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import bench
from sde4mbrl_b200 import _lib as L, configs, sde_models
from sde4mbrl_b200.engine import Engine
w = bench.build_workload(); dev = torch.device("cuda")
for P, N in ((4096, 32), (1024, 32), (4096, 8)):
    H = 20
    pm = configs.hexa_model(horizon=H, num_particles=N, stepsize=0.05)
    eng = Engine(sde_models.SDERotorModel(pm))
    rng = np.random.default_rng(0)
    y0 = np.tile(w["y0"], (P, 1)).astype(np.float32)
    u = torch.full((P, H, 6), 0.33, device=dev)
    keys = torch.stack([torch.zeros(P, dtype=torch.int32, device=dev), torch.arange(P, dtype=torch.int32, device=dev)], dim=1).contiguous()
    ts = np.full(H, 0.05, np.float32)
    out = None
    for name, lanes in (("tc", L.LANES_TC), ("fp32", 1), ("auto", 0)):
        tms = []
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); xb = eng.rollout(w["nn"], y0, u, ts, key=keys, N=N, lanes=lanes); e1.record(); torch.cuda.synchronize()
            if rep: tms.append(e0.elapsed_time(e1))
        if out is None: out = xb.clone()
        else: assert float((xb - out).abs().max()) < 1e-4
        print("P %d N %d %-5s %.3f ms" % (P, N, name, np.mean(tms)), flush=True)
