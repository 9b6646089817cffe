"""Value+gradient launch time, warp-specialised kernel vs the lane-split kernels, over launch sizes (H = 20 MPC shape)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sde4mbrl_b200 import configs, costs, sde_models, _lib as L
from sde4mbrl_b200.engine import Engine
from sde4mbrl_b200.nsde import compute_timesteps
from oracle import threefry as tf
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
for P, N in ((256, 16), (512, 32), (1024, 32), (2048, 32), (4096, 32)):
    y0 = torch.tensor(np.tile(configs.hover_state(), (P, 1)), device="cuda")
    u = torch.tensor(rng.uniform(0.30, 0.36, (P, H, 6)).astype(np.float32), device="cuda")
    keys = tf.split(tf.PRNGKey(1), P)
    row = []
    for lanes in ((L.LANES_WS,) if os.environ.get("WS_ONLY") else (1, 4, 8, L.LANES_WS)):
        try:
            for _ in range(3):
                eng.cost_grad(nn, y0, u, ts, xd, stage.desc, N=N, key=keys, lanes=lanes)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                eng.cost_grad(nn, y0, u, ts, xd, stage.desc, N=N, key=keys, lanes=lanes)
            e1.record(); torch.cuda.synchronize()
            row.append("%7.1f" % (e0.elapsed_time(e1) * 100))
        except Exception as ex:
            row.append("   n/a ")
    print("P=%5d N=%4d G=%6d  us per launch (lanes 1 / 4 / 8 / ws, or ws only):" % (P, N, P * N), " ".join(row))
