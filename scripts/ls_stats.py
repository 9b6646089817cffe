import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from sde4mbrl_b200 import configs, sde_models
from sde4mbrl_b200.engine import Engine
from sde4mbrl_b200.apg_batched import BatchedAPG
w = bench.build_workload(); dev = torch.device("cuda")
P, HM = 4096, 20
rng = np.random.default_rng(123)
y0m = np.tile(w["y0"], (P, 1)); y0m[:, 3:6] += rng.normal(0, 0.05, (P, 3)).astype(np.float32); y0m[:, 10:13] += rng.normal(0, 0.05, (P, 3)).astype(np.float32)
keys = np.stack([np.zeros(P, np.uint32), np.arange(P, dtype=np.uint32)], axis=1)
xref = np.tile(w["xref"][:1], (HM + 1, 1))
for n_part in (1, 32):
    pmm = configs.hexa_model(horizon=HM, num_particles=n_part, stepsize=0.05)
    mpcm = configs.hexa_mpc(horizon=HM, dt=0.05, num_particles=n_part, max_iter=20)
    print(mpcm["apg_mpc"]["linesearch"])
    eng = Engine(sde_models.SDERotorModel(pmm))
    s = BatchedAPG(eng, w["nn"], w["stage"], mpcm["apg_mpc"], np.full(HM, 0.05, np.float32), n_part, lb=1e-4, ub=1.0)
    s.init(torch.full((HM, 6), 0.33 + 2e-4, device=dev), torch.as_tensor(y0m, device=dev), xref, torch.as_tensor(keys.view(np.int32), device=dev)).run(20)
    al = s.t["avg_linesearch"].cpu().numpy()
    print(n_part, "avg linesearch steps: mean %.3f max %.3f; fraction of problems with avg<=1.5: %.2f" % (al.mean(), al.max(), (al <= 1.5).mean()), "not_done", int(s.t["not_done"].sum()))
    # fraction of problems whose x (plain step) beats v (momentum point), per iteration
    s.init(torch.full((HM, 6), 0.33 + 2e-4, device=dev), torch.as_tensor(y0m, device=dev), xref, torch.as_tensor(keys.view(np.int32), device=dev))
    fr = []
    for it in range(20):
        s.step(); fr.append(float(s.pick.float().mean()))
    print(n_part, "fraction x <= v per iteration:", " ".join("%.2f" % f for f in fr))
