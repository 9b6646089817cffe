"""A few value+gradient launches at a batched MPC shape (P problems x N particles, H = 20) with the chosen kernel, for ncu.
usage: ws_probe_batch.py P N [ws|1|4|8] [launches]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sde4mbrl_b200 import configs, costs, sde_models, _lib as L
from sde4mbrl_b200.engine import Engine
from sde4mbrl_b200.nsde import compute_timesteps
from oracle import threefry as tf  # (key chain of the probe inputs only)

P, N = int(sys.argv[1]), int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "ws"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
lanes = L.LANES_WS if mode == "ws" else int(mode)
H = 20
pm = configs.hexa_model(horizon=H, num_particles=1)
model = sde_models.SDERotorModel(pm)
nn = configs.random_params(model, seed=0, scale="fan_in")
eng = Engine(model)
rng = np.random.default_rng(0)
cp = configs.hexa_mpc()["cost_params"]
stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
ts = compute_timesteps(pm)
xd = torch.tensor(np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32), device="cuda")
y0 = torch.tensor(np.tile(configs.hover_state(), (P, 1)), device="cuda")
u = torch.tensor(rng.uniform(0.30, 0.36, (P, H, 6)).astype(np.float32), device="cuda")
keys = tf.split(tf.PRNGKey(1), P)
for _ in range(reps):
    c, g, _ = eng.cost_grad(nn, y0, u, ts, xd, stage.desc, N=N, key=keys, lanes=lanes)
torch.cuda.synchronize()
print("mean cost", float(c.mean()))
