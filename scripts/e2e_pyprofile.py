"""cProfile of the host-array value_and_grad call (C3 shape): where the Python time of the e2e path goes."""
import sys, os, cProfile, pstats, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sde4mbrl_b200 import nsde, sde_models, configs

w = bench.build_workload()
N, H = 4096, 100
pm = configs.hexa_model(horizon=H, num_particles=N, stepsize=0.01)
mpc = configs.hexa_mpc(horizon=H, dt=0.01, num_particles=N)
with contextlib.redirect_stdout(io.StringIO()):
    _, mc, _, _, _ = nsde.create_online_cost_sampling_fn(pm, mpc, sde_constr=sde_models.SDERotorModel)
from sde4mbrl_b200.engine import PackedParams
nn, stage = PackedParams(mc.engine.model, w["nn"], mc.engine.desc), w["stage"]
y0 = w["y0"]; opt = np.ascontiguousarray(w["opt"]); xref = w["xref"]; key = np.ascontiguousarray(w["key"])
f = lambda: mc.value_and_grad(nn, y0, opt, key, stage, extra_cost_args=xref)
for _ in range(20): f()
import time
t0 = time.perf_counter()
for _ in range(500): f()
print("wall per call: %.1f us" % ((time.perf_counter() - t0) / 500 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(500): f()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:5000])
