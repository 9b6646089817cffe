"""Barrier timeline of one CTA of the warp-specialised kernel (library built with -DSDE4_WS_TRACE).
Prints, per step, when each role arrives at the two barriers of the step relative to the barrier's release."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sde4mbrl_b200 import configs, costs, sde_models, _lib as L
from sde4mbrl_b200.engine import Engine
from sde4mbrl_b200.nsde import compute_timesteps
H, N = 100, 4096
pm = configs.hexa_model(horizon=H, num_particles=N); pm["stepsize"] = 0.01
model = sde_models.SDERotorModel(pm)
nn = configs.random_params(model, seed=0, scale="fan_in")
eng = Engine(model)
rng = np.random.default_rng(0)
y0d = torch.tensor(configs.hover_state()[None], device="cuda")
ud = torch.tensor(rng.uniform(0.30, 0.36, (H, 6)).astype(np.float32), device="cuda")
cp = configs.hexa_mpc()["cost_params"]
stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
xd = torch.tensor(np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32), device="cuda")
ts = compute_timesteps(pm); key = np.array([0, 2050], np.uint32)
lib = L.load()
out = np.zeros((4, 1024), np.int64); cnt = np.zeros(4, np.int32)
for _ in range(3):
    eng.cost_grad(nn, y0d, ud, ts, xd, stage.desc, N=N, key=key, lanes=L.LANES_WS)
    lib.sde4_debug_ws_trace(out.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p))
print("events per role", cnt)
names = ["PHY", "DENS", "RES", "AUX"]
ev = out.reshape(4, 512, 2)[:, 1:]  # (event 0 of every role is B0, the start barrier for the first noise slot)
# PHY and RES meet at every barrier (B1, B2 per forward step; B3, B4 per backward step); DENS joins B1, B2 and, in the
# backward sweep, B4 only (it works one step ahead through two output slots); AUX joins B2 / B4
pr = ev[[0, 2], :4 * H]
arr, dep = pr[:, :, 0], pr[:, :, 1]
dens_f = ev[1, :2 * H]
dens_b = ev[1, 2 * H:3 * H]           # its k-th backward event is B4 of backward step k
rel = dep.max(axis=0)
relf = np.maximum(rel[:2 * H], dens_f[:, 1])
gaps = np.diff(relf)
print("forward: mean cycles B1->B2 %.0f, B2->B1 %.0f ; total fwd %.0f" % (gaps[0::2].mean(), gaps[1::2].mean(), relf[-1] - relf[0]))
relb = rel[2 * H:].copy()
relb[1::2] = np.maximum(relb[1::2], dens_b[:, 1])
gb = np.diff(relb)
print("backward: mean cycles B3->B4 %.0f, B4->B3 %.0f ; total bwd %.0f" % (gb[0::2].mean(), gb[1::2].mean(), relb[-1] - relb[0]))
for name, sl in (("fwd B1", slice(0, 2 * H, 2)), ("fwd B2", slice(1, 2 * H, 2))):
    a3 = np.stack([arr[0, sl], dens_f[sl, 0], arr[1, sl]])
    last = a3.argmax(axis=0)
    wait = (relf[sl][None, :] - a3).mean(axis=1)
    print(name, "last arriver", {n: int((last == r).sum()) for r, n in enumerate(["PHY", "DENS", "RES"])}, "mean wait", {n: int(wait[r]) for r, n in enumerate(["PHY", "DENS", "RES"])})
b3 = arr[:, 2 * H::2]
print("bwd B3 (PHY, RES)", "last arriver", {n: int((b3.argmax(axis=0) == r).sum()) for r, n in enumerate(["PHY", "RES"])},
      "mean wait", {n: int((relb[0::2][None, :] - b3).mean(axis=1)[r]) for r, n in enumerate(["PHY", "RES"])})
a4 = np.stack([arr[0, 2 * H + 1::2], dens_b[:, 0], arr[1, 2 * H + 1::2]])
print("bwd B4", "last arriver", {n: int((a4.argmax(axis=0) == r).sum()) for r, n in enumerate(["PHY", "DENS", "RES"])},
      "mean wait", {n: int((relb[1::2][None, :] - a4).mean(axis=1)[r]) for r, n in enumerate(["PHY", "DENS", "RES"])})
aux_f = ev[3, :H]                      # AUX: B2 of every forward step
print("fwd B2 AUX mean wait %d (its noise generation for the next step takes the rest of the %d-cycle step)"
      % (int((relf[1::2] - aux_f[:, 0]).mean()), int(np.diff(relf[1::2]).mean())))
