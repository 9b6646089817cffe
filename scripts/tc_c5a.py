"""BASELINE configs[4] (2^20 cartpole start states x 30 steps) through the tensor-core rollout (lanes = TC) and the FP32
one-thread-per-sample kernel (lanes = 1): CUDA-event times.  `--one` runs a single tensor-core launch (ncu target)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from sde4mbrl_b200 import _lib as L, configs, sde_models
from sde4mbrl_b200.engine import Engine, PackedParams

dev = torch.device("cuda:0")
PS, HS = 1 << 20, 30
pmc = configs.cartpole_model(horizon=HS, num_particles=1)
mc = sde_models.CartPoleSDE(pmc)
engc = Engine(mc)
pkc = PackedParams(mc, configs.random_params(mc, seed=0), engc.desc)
g = torch.Generator(device=dev).manual_seed(0)
obs = (torch.rand((PS, 4), device=dev, generator=g) * 2 - 1) * torch.tensor([2.0, 3.0, 3.14159, 6.0], device=dev)
act = torch.rand((PS, HS, 1), device=dev, generator=g) * 2 - 1
keys = torch.stack([torch.zeros(PS, dtype=torch.int32, device=dev), torch.arange(PS, dtype=torch.int32, device=dev)], dim=1).contiguous()
out = torch.empty((HS + 1, 4, PS), device=dev)
ts = np.full(HS, 0.02, np.float32)
if "--one" in sys.argv:
    engc.rollout(pkc, obs, act, ts, key=keys, N=1, out=out, lanes=L.LANES_TC)
    torch.cuda.synchronize()
    sys.exit(0)
for name, lanes in (("tensor-core", L.LANES_TC), ("fp32", 1)):
    tms = []
    for rep in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engc.rollout(pkc, obs, act, ts, key=keys, N=1, out=out, lanes=lanes)
        e1.record()
        torch.cuda.synchronize()
        if rep:
            tms.append(e0.elapsed_time(e1))
    tm = float(np.mean(tms))
    print("%-12s %.3f ms  %.2f G particle-steps/s  %.1f TFLOP/s algorithmic" % (name, tm, PS * HS / tm / 1e6, 3135.0 * PS * HS / tm / 1e9))
