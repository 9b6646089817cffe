"""Where the end-to-end (host arrays -> public API -> host result) time of one value+grad step goes."""
import contextlib, io, os, sys, time, cProfile, pstats
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sde4mbrl_b200 import nsde, configs, sde_models
from sde4mbrl_b200.engine import PackedParams
w = bench.build_workload()
mpc = configs.hexa_mpc(horizon=100, dt=0.01, num_particles=4096)
pm2 = configs.hexa_model(horizon=100, num_particles=4096, stepsize=0.01)
with contextlib.redirect_stdout(io.StringIO()):
    _, multi_cost, _, _, _ = nsde.create_online_cost_sampling_fn(pm2, mpc, sde_constr=sde_models.SDERotorModel)
eng = multi_cost.engine
pk = PackedParams(eng.model, w["nn"], eng.desc)
opt_leaf = torch.as_tensor(w["opt"]).clone().requires_grad_(True)
key_np = np.ascontiguousarray(w["key"])
out_h = torch.empty(601)
def step():
    c, _ = multi_cost(pk, w["y0"], opt_leaf, key_np, w["stage"], extra_cost_args=w["xref"])
    (g,) = torch.autograd.grad(c, opt_leaf)
    out_h[0] = c; out_h[1:] = g.reshape(-1)
def raw():
    return eng.cost_grad_host(pk, w["y0"], w["opt"], multi_cost.time_steps, w["xref"], w["stage"].desc, key=key_np, N=4096)
for fn, name in ((step, "public API"), (raw, "Engine.cost_grad_host")):
    for _ in range(5): fn()
    t0 = time.perf_counter()
    for _ in range(100): fn()
    print(name, "ms/step", (time.perf_counter() - t0) / 100 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(100): step()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[-4200:])
