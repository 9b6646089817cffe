"""Batched APG (hexacopter, H=20, 20 iterations): batch time vs (P, N) for the two x/v evaluation orders and the two
line-search evaluation orders -- data for the automatic choices in apg_batched.BatchedAPG._alloc."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sde4mbrl_b200 import configs, sde_models
from sde4mbrl_b200.engine import Engine
from sde4mbrl_b200.apg_batched import BatchedAPG
w = bench.build_workload(); dev = torch.device("cuda"); HM = 20
rng = np.random.default_rng(123)
for P, N in ((1, 1024), (64, 32), (256, 1), (512, 32), (1024, 32), (2048, 32), (4096, 1), (4096, 8), (4096, 32)):
    y0 = np.tile(w["y0"], (P, 1)); y0[:, 3:6] += rng.normal(0, 0.05, (P, 3)).astype(np.float32)
    keys = np.stack([np.zeros(P, np.uint32), np.arange(P, dtype=np.uint32)], axis=1)
    xref = np.tile(w["xref"][:1], (HM + 1, 1))
    pm = configs.hexa_model(horizon=HM, num_particles=N, stepsize=0.05)
    mpc = configs.hexa_mpc(horizon=HM, dt=0.05, num_particles=N, max_iter=20)
    eng = Engine(sde_models.SDERotorModel(pm))
    out = []
    for merge, prog in ((False, False), (True, False), (False, True)):
        if prog and N % 32: out.append(float("nan")); continue
        s = BatchedAPG(eng, w["nn"], w["stage"], mpc["apg_mpc"], np.full(HM, 0.05, np.float32), N, lb=1e-4, ub=1.0,
                       merge_xv_grad=merge, progressive_ls=prog)
        x0 = torch.full((HM, 6), 0.33 + 2e-4, device=dev)
        y0d = torch.as_tensor(y0, device=dev); kd = torch.as_tensor(keys.view(np.int32), device=dev)
        tms = []
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.init(x0, y0d, xref, kd).run(20); e1.record(); torch.cuda.synchronize()
            if rep: tms.append(e0.elapsed_time(e1))
        out.append(np.mean(tms))
    print("P %5d N %5d  samples 2PN %7d | separate %.2f ms | merged x/v %.2f ms | progressive LS %.2f ms" % (P, N, 2 * P * N, *out))
