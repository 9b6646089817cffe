"""A few C3 value+gradient launches with the chosen kernel (for ncu).  usage: ws_probe.py [ws|8|0] [launches]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sde4mbrl_b200 import configs, costs, sde_models, _lib as L
from sde4mbrl_b200.engine import Engine
from sde4mbrl_b200.nsde import compute_timesteps

mode = sys.argv[1] if len(sys.argv) > 1 else "ws"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
lanes = L.LANES_WS if mode == "ws" else int(mode)
H, N = 100, 4096
pm = configs.hexa_model(horizon=H, num_particles=N)
pm["stepsize"] = 0.01
model = sde_models.SDERotorModel(pm)
nn = configs.random_params(model, seed=0, scale="fan_in")
eng = Engine(model)
rng = np.random.default_rng(0)
y0d = torch.tensor(configs.hover_state()[None], device="cuda")
ud = torch.tensor(rng.uniform(0.30, 0.36, (H, 6)).astype(np.float32), device="cuda")
cp = configs.hexa_mpc()["cost_params"]
stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
xd = torch.tensor(np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32), device="cuda")
ts = compute_timesteps(pm)
key = np.array([0, 2050], np.uint32)
for _ in range(reps):
    c, g, _ = eng.cost_grad(nn, y0d, ud, ts, xd, stage.desc, N=N, key=key, lanes=lanes)
torch.cuda.synchronize()
print("cost", float(c[0]))
