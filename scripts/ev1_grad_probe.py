"""One-thread-per-sample gradient kernel at the C4 / N=32 shape (4096 problems x 32 particles, H=20): timing, and the
launch to point `ncu --set full -k regex:k_cost` at."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sde4mbrl_b200 import configs, sde_models
from sde4mbrl_b200.engine import Engine
w = bench.build_workload(); dev = torch.device("cuda")
H, P, N = 20, 4096, 32
threads = int(sys.argv[1]) if len(sys.argv) > 1 else 0
pm = configs.hexa_model(horizon=H, num_particles=N, stepsize=0.05)
eng = Engine(sde_models.SDERotorModel(pm)); pk = eng.packed(w["nn"])
ts = np.full(H, 0.05, np.float32)
xref = torch.as_tensor(np.tile(w["xref"][:1], (H + 1, 1)), device=dev)
y0 = torch.as_tensor(np.tile(w["y0"], (P, 1)), device=dev)
opt = torch.full((P, H, 6), 0.33, device=dev) + 0.01 * torch.rand((P, H, 6), device=dev)
keys = torch.stack([torch.zeros(P, dtype=torch.int32, device=dev), torch.arange(P, dtype=torch.int32, device=dev)], dim=1).contiguous()
cost = torch.empty(P, device=dev); gout = torch.empty_like(opt)
for grad in (True, False):
    tms = []
    for rep in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.cost_grad(pk, y0, opt, ts, xref, w["stage"].desc, key=keys, N=N, want_grad=grad, lanes=1, cost_out=cost,
                      grad_out=gout if grad else None, block_threads=threads)
        e1.record(); torch.cuda.synchronize()
        if rep > 1: tms.append(e0.elapsed_time(e1))
    ps = P * N * H / (np.mean(tms) * 1e-3)
    print("grad" if grad else "value", "threads", threads, "ms %.3f  particle-steps/s %.3g  algorithmic TFLOP/s %.1f" %
          (np.mean(tms), ps, ps * 3880 * (2 if grad else 1) / 1e12))
