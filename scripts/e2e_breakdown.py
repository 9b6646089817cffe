"""Where the end-to-end (host buffers -> public API -> host result) time of one value+grad step goes."""
import contextlib, io, os, sys, time, cProfile, pstats
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sde4mbrl_b200 import nsde, configs, sde_models
from sde4mbrl_b200.engine import PackedParams
w = bench.build_workload()
dev = torch.device("cuda")
mpc = configs.hexa_mpc(horizon=100, dt=0.01, num_particles=4096)
pm2 = configs.hexa_model(horizon=100, num_particles=4096, stepsize=0.01)
with contextlib.redirect_stdout(io.StringIO()):
    _, multi_cost, _, _, _ = nsde.create_online_cost_sampling_fn(pm2, mpc, sde_constr=sde_models.SDERotorModel)
eng = multi_cost.engine
pk = PackedParams(eng.model, w["nn"], eng.desc)
y0_h = torch.as_tensor(w["y0"]).pin_memory(); opt_h = torch.as_tensor(w["opt"]).pin_memory()
xref_h = torch.as_tensor(w["xref"]).pin_memory(); key_h = torch.as_tensor(w["key"].view(np.int32)).pin_memory()
out_h = torch.empty(601).pin_memory()
def step():
    y = y0_h.to(dev, non_blocking=True); o = opt_h.to(dev, non_blocking=True).requires_grad_(True)
    xr = xref_h.to(dev, non_blocking=True); k = key_h.to(dev, non_blocking=True)
    c, _ = multi_cost(pk, y, o, k, w["stage"], extra_cost_args=xr)
    (g,) = torch.autograd.grad(c, o)
    out_h[0:1].copy_(c.reshape(1), non_blocking=True); out_h[1:].copy_(g.reshape(-1), non_blocking=True)
    torch.cuda.synchronize()
for _ in range(5): step()
t0 = time.perf_counter()
for _ in range(50): step()
print("e2e ms/step", (time.perf_counter() - t0) / 50 * 1e3)
# host-only cost: same without waiting for the GPU
def step_nosync():
    y = y0_h.to(dev, non_blocking=True); o = opt_h.to(dev, non_blocking=True).requires_grad_(True)
    xr = xref_h.to(dev, non_blocking=True); k = key_h.to(dev, non_blocking=True)
    c, _ = multi_cost(pk, y, o, k, w["stage"], extra_cost_args=xr)
    (g,) = torch.autograd.grad(c, o)
    out_h[0:1].copy_(c.reshape(1), non_blocking=True); out_h[1:].copy_(g.reshape(-1), non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): step_nosync()
t_host = (time.perf_counter() - t0) / 50 * 1e3
torch.cuda.synchronize()
print("host enqueue ms/step", t_host)
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step_nosync()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[-3500:])
