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
nb = int(cnt[:3].min())
ev = out.reshape(4, 512, 2)[:3, :nb]
arr, dep = ev[:, :, 0], ev[:, :, 1]
rel = dep.max(axis=0)  # release time of each barrier (B1 / B3: three warps; B2 / B4: all four)
gaps = np.diff(rel)
fw = gaps[:2 * H - 1]; bw = gaps[2 * H:2 * H + 2 * H - 1]
print("forward: mean cycles B1->B2 %.0f, B2->B1 %.0f ; total fwd %.0f" % (fw[0::2].mean(), fw[1::2].mean(), rel[2 * H - 1] - rel[0]))
print("backward: mean cycles B3->B4 %.0f, B4->B3 %.0f ; total bwd %.0f" % (bw[0::2].mean(), bw[1::2].mean(), rel[4 * H - 1] - rel[2 * H]))
last = arr.argmax(axis=0)
for name, sl in (("fwd B1", slice(0, 2 * H, 2)), ("fwd B2", slice(1, 2 * H, 2)), ("bwd B3", slice(2 * H, 4 * H, 2)), ("bwd B4", slice(2 * H + 1, 4 * H, 2))):
    l = last[sl]
    wait = (rel[sl][None, :] - arr[:, sl]).mean(axis=1)
    print(name, "last arriver", {names[r]: int((l == r).sum()) for r in range(3)}, "mean wait", {names[r]: int(wait[r]) for r in range(3)})
# AUX: barriers B2 (H of them) then B4 (H) + final
aux = out.reshape(4, 512, 2)[3, :cnt[3]]
print("AUX mean wait at its barriers: fwd %.0f bwd %.0f" % ((aux[:H, 1] - aux[:H, 0]).mean(), (aux[H:2 * H, 1] - aux[H:2 * H, 0]).mean()))
