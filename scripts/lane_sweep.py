"""Time of one cost launch (hexacopter, H=20, N=1) vs number of problems G and lanes per sample: data for the
launcher's pick_lanes rule."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sde4mbrl_b200 import configs, sde_models
from sde4mbrl_b200.engine import Engine
w = bench.build_workload(); dev = torch.device("cuda")
H = 20
pm = configs.hexa_model(horizon=H, num_particles=1, stepsize=0.05)
eng = Engine(sde_models.SDERotorModel(pm)); pk = eng.packed(w["nn"])
ts = np.full(H, 0.05, np.float32)
xref = torch.as_tensor(np.tile(w["xref"][:1], (H + 1, 1)), device=dev)
for grad in (False, True):
    for G in (2048, 4096, 8192, 16384, 28672, 49152, 65536, 98304, 131072):
        y0 = torch.as_tensor(np.tile(w["y0"], (G, 1)), device=dev)
        opt = torch.full((G, H, 6), 0.33, device=dev) + 0.01 * torch.rand((G, H, 6), device=dev)
        keys = torch.stack([torch.zeros(G, dtype=torch.int32, device=dev), torch.arange(G, dtype=torch.int32, device=dev)], dim=1).contiguous()
        cost = torch.empty(G, device=dev); gout = torch.empty_like(opt) if grad else None
        res = []
        for lanes in (1, 4, 8):
            ts_ = []
            for rep in range(8):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.cost_grad(pk, y0, opt, ts, xref, w["stage"].desc, key=keys, N=1, want_grad=grad, lanes=lanes, cost_out=cost, grad_out=gout)
                e1.record(); torch.cuda.synchronize()
                if rep > 2: ts_.append(e0.elapsed_time(e1))
            res.append(np.mean(ts_) * 1e3)
        print("grad" if grad else "value", "G", G, "us: L1 %.0f  L4 %.0f  L8 %.0f" % tuple(res), " best L", (1, 4, 8)[int(np.argmin(res))])
