"""C4 (4096 hexacopter MPC solves, H=20, 20 APG iterations): batch time for different lane-splits of the
three cost launches of an iteration."""
import sys, os, itertools, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sde4mbrl_b200 import configs, sde_models
from sde4mbrl_b200.engine import Engine
from sde4mbrl_b200.apg_batched import BatchedAPG
w = bench.build_workload(); dev = torch.device("cuda")
P, HM = 4096, 20
rng = np.random.default_rng(123)
y0m = np.tile(w["y0"], (P, 1)); y0m[:, 3:6] += rng.normal(0, 0.05, (P, 3)).astype(np.float32)
keys = np.stack([np.zeros(P, np.uint32), np.arange(P, dtype=np.uint32)], axis=1)
xref = np.tile(w["xref"][:1], (HM + 1, 1))
n_part = int(sys.argv[1]) if len(sys.argv) > 1 else 1
combos = [dict(cand=0, xv=0, y=0)] if "--default-only" in sys.argv else [dict(cand=c, xv=x, y=y) for c, x, y in [(0, 0, 0), (1, 8, 8), (4, 8, 8), (8, 8, 8), (4, 4, 8), (4, 4, 4), (1, 4, 8), (1, 1, 1), (1, 1, 4), (1, 1, 8)]]
for lanes in combos:
    pmm = configs.hexa_model(horizon=HM, num_particles=n_part, stepsize=0.05)
    mpcm = configs.hexa_mpc(horizon=HM, dt=0.05, num_particles=n_part, max_iter=20)
    eng = Engine(sde_models.SDERotorModel(pmm))
    try:
        solver = BatchedAPG(eng, w["nn"], w["stage"], mpcm["apg_mpc"], np.full(HM, 0.05, np.float32), n_part, lb=1e-4, ub=1.0, lanes=lanes)
        x0 = torch.full((HM, 6), 0.33 + 2e-4, device=dev)
        y0d = torch.as_tensor(y0m, device=dev); kd = torch.as_tensor(keys.view(np.int32), device=dev)
        tms = []
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); solver.init(x0, y0d, xref, kd).run(20); e1.record(); torch.cuda.synchronize()
            if rep: tms.append(e0.elapsed_time(e1))
        print(n_part, lanes, "ms/batch %.3f" % np.mean(tms), "opt", float(solver.t["opt_cost"].mean()))
    except Exception as ex:
        print(n_part, lanes, "failed", repr(ex)[:100])
