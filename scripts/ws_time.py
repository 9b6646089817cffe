"""C3 value+gradient launch: warp-specialised kernel vs the 8-lane one-role kernel (CUDA events, L2 flushed)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sde4mbrl_b200 import configs, costs, sde_models, _lib as L
from sde4mbrl_b200.engine import Engine
from sde4mbrl_b200.nsde import compute_timesteps

H, N = 100, 4096
pm = configs.hexa_model(horizon=H, num_particles=N)
pm["stepsize"] = 0.01
model = sde_models.SDERotorModel(pm)
nn = configs.random_params(model, seed=0, scale="fan_in")
eng = Engine(model)
rng = np.random.default_rng(0)
y0 = configs.hover_state()[None]
u = rng.uniform(0.30, 0.36, (H, 6)).astype(np.float32)
cp = configs.hexa_mpc()["cost_params"]
stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
ts = compute_timesteps(pm)
key = np.array([0, 2050], np.uint32)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
y0d, ud, xd = torch.tensor(y0, device="cuda"), torch.tensor(u, device="cuda"), torch.tensor(xref, device="cuda")
res = {}
for name, lanes in (("lanes8", 8), ("ws", L.LANES_WS), ("auto", 0)):
    for _ in range(5):
        c, g, _ = eng.cost_grad(nn, y0d, ud, ts, xd, stage.desc, N=N, key=key, lanes=lanes)
    times = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        c, g, _ = eng.cost_grad(nn, y0d, ud, ts, xd, stage.desc, N=N, key=key, lanes=lanes)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    res[name] = (float(np.median(times)), float(np.min(times)), float(c[0]), g.cpu().numpy())
    print("%-8s median %.4f ms  min %.4f ms  cost %.6f" % (name, res[name][0], res[name][1], res[name][2]))
print("grad rel diff ws vs lanes8: %.2e" % (np.abs(res["ws"][3] - res["lanes8"][3]).max() / np.abs(res["lanes8"][3]).max()))
