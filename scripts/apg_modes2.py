"""Round 2 refit of the automatic mode choices of the batched APG (hexacopter, H = 20, 20 iterations) now that the
value+gradient launches of 2 k - 98 k samples run on the warp-specialised kernel: batch time vs (P, N) for
auto / speculative / merged x/v / merged + progressive / separate / progressive line search."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sde4mbrl_b200 import configs, sde_models
from sde4mbrl_b200.engine import Engine
from sde4mbrl_b200.apg_batched import BatchedAPG
w = bench.build_workload(); dev = torch.device("cuda"); HM = 20
rng = np.random.default_rng(123)
sizes = ((512, 1), (1024, 1), (2048, 1), (4096, 1), (512, 32), (1024, 32), (2048, 32), (4096, 32))
if len(sys.argv) > 1:
    sizes = tuple(tuple(int(v) for v in a.split("x")) for a in sys.argv[1:])
for P, N in sizes:
    y0 = np.tile(w["y0"], (P, 1)); y0[:, 3:6] += rng.normal(0, 0.05, (P, 3)).astype(np.float32)
    keys = np.stack([np.zeros(P, np.uint32), np.arange(P, dtype=np.uint32)], axis=1)
    xref = np.tile(w["xref"][:1], (HM + 1, 1))
    pm = configs.hexa_model(horizon=HM, num_particles=N, stepsize=0.05)
    mpc = configs.hexa_mpc(horizon=HM, dt=0.05, num_particles=N, max_iter=20)
    eng = Engine(sde_models.SDERotorModel(pm))
    out = []
    for name, kw in (("auto", {}), ("spec", {"speculative": True}), ("merged", {"merge_xv_grad": True, "progressive_ls": False}),
                     ("merged+prog", {"merge_xv_grad": True, "progressive_ls": True}),
                     ("separate", {"merge_xv_grad": False, "progressive_ls": False}), ("prog", {"merge_xv_grad": False, "progressive_ls": True})):
        if kw.get("progressive_ls") and N % 32:
            out.append("%s nan" % name); continue
        try:
            s = BatchedAPG(eng, w["nn"], w["stage"], mpc["apg_mpc"], np.full(HM, 0.05, np.float32), N, lb=1e-4, ub=1.0, **kw)
            x0 = torch.full((HM, 6), 0.33 + 2e-4, device=dev)
            y0d = torch.as_tensor(y0, device=dev); kd = torch.as_tensor(keys.view(np.int32), device=dev)
            tms = []
            for rep in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); s.init(x0, y0d, xref, kd).run(20); e1.record(); torch.cuda.synchronize()
                if rep: tms.append(e0.elapsed_time(e1))
            out.append("%s %.2f" % (name, np.mean(tms)))
            del s
        except Exception as ex:
            out.append("%s ERR(%s)" % (name, str(ex)[:40]))
    print("P %5d N %3d | " % (P, N) + " | ".join(out), flush=True)
