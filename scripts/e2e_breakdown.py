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
    (c, _), g = multi_cost.value_and_grad(pk, w["y0"], w["opt"], key_np, w["stage"], extra_cost_args=w["xref"])
def step_autograd():
    c, _ = multi_cost(pk, w["y0"], opt_leaf, key_np, w["stage"], extra_cost_args=w["xref"])
    (g,) = torch.autograd.grad(c, opt_leaf)
def raw():
    return eng.cost_grad_host(pk, w["y0"], w["opt"], multi_cost.time_steps, w["xref"], w["stage"].desc, key=key_np, N=4096)
def raw_xtraj():
    return eng.cost_grad_host(pk, w["y0"], w["opt"], multi_cost.time_steps, w["xref"], w["stage"].desc, key=key_np, N=4096, want_xtraj=True)
def dev_only_xtraj():
    eng.cost_grad(pk, y0_d, opt_d, multi_cost.time_steps, xref_d, w["stage"].desc, key=key_d, N=4096, want_xtraj=True); torch.cuda.synchronize()
def dev_only():
    eng.cost_grad(pk, y0_d, opt_d, multi_cost.time_steps, xref_d, w["stage"].desc, key=key_d, N=4096); torch.cuda.synchronize()
y0_d = torch.as_tensor(w["y0"], device="cuda")[None]; opt_d = torch.as_tensor(w["opt"], device="cuda")
xref_d = torch.as_tensor(w["xref"], device="cuda"); key_d = torch.as_tensor(key_np.view(np.int32), device="cuda")
for fn, name in ((step, "value_and_grad (host arrays)"), (step_autograd, "multi_cost + autograd.grad (host arrays)"), (raw, "Engine.cost_grad_host"), (raw_xtraj, "Engine.cost_grad_host + xtraj"), (dev_only, "Engine.cost_grad on device tensors + sync"), (dev_only_xtraj, "Engine.cost_grad on device tensors + xtraj + sync")):
    for _ in range(5): fn()
    t0 = time.perf_counter()
    for _ in range(300): fn()
    print(name, "ms/step", (time.perf_counter() - t0) / 300 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(100): step()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[-4200:])
