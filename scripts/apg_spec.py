"""Speculative vs two-launch APG iteration: ms per 20-iteration solve (scripts/apg_spec.py, run on the GPU box)."""
import contextlib, io, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sde4mbrl_b200 import configs, costs, sde_models
from sde4mbrl_b200.engine import Engine
from sde4mbrl_b200.apg_batched import BatchedAPG

dev = torch.device("cuda:0")


def run(solver, x0, y0, xref, key, reps=6):
    tms = []
    for rep in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        solver.init(x0, y0, xref, key).run(20)
        e1.record()
        torch.cuda.synchronize()
        if rep > 1:
            tms.append(e0.elapsed_time(e1))
    return float(np.mean(tms)), float(solver.t["opt_cost"].mean())


H = 20
for P, N in ((1, 1), (1, 32), (1, 128), (1, 256), (1, 1024), (16, 1), (64, 1), (256, 1), (1024, 1)):
    pm = configs.hexa_model(horizon=H, num_particles=N, stepsize=0.05)
    mpc = configs.hexa_mpc(horizon=H, dt=0.05, num_particles=N, max_iter=20)
    mpc["apg_mpc"].update(atol=0.0, rtol=0.0, max_no_improvement_iter=20)
    model = sde_models.SDERotorModel(pm)
    nn = configs.random_params(model, seed=0)
    cp = mpc["cost_params"]
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    rng = np.random.default_rng(5)
    y0 = np.tile(configs.hover_state(), (P, 1)).astype(np.float32)
    y0[:, 3:6] += rng.normal(0, 0.05, (P, 3)).astype(np.float32)
    xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
    xref[:, :3] = [0.0, 0.5, 1.4]
    keys = np.stack([np.zeros(P, np.uint32), np.arange(P, dtype=np.uint32)], axis=1)
    x0 = torch.full((H, 6), 0.33 + 2e-4, device=dev)
    eng = Engine(model)
    out = []
    for spec in (False, True):
        s = BatchedAPG(eng, nn, stage, mpc["apg_mpc"], np.full(H, 0.05, np.float32), N, lb=1e-4, ub=1.0, speculative=spec)
        out.append(run(s, x0, y0, xref, keys))
    print("hexa P=%d N=%d: two-launch %.3f ms (cost %.6f)   speculative %.3f ms (cost %.6f)" %
          (P, N, out[0][0], out[0][1], out[1][0], out[1][1]), flush=True)

# cartpole C2
HC, NC = 50, 1024
for NC in (64, 256, 1024):
    pmc = configs.cartpole_model(horizon=HC, num_particles=NC)
    mpcc = configs.cartpole_mpc(horizon=HC, dt=0.02, num_particles=NC, max_iter=20)
    mpcc["apg_mpc"].update(atol=0.0, rtol=0.0, max_no_improvement_iter=20)
    modc = sde_models.CartPoleSDE(pmc)
    nnc = configs.random_params(modc, seed=0)
    for k_, v_ in nnc.items():
        if k_.endswith("linear_2") and "init_residual_networks" in k_:
            v_["w"] = v_["w"] * np.float32(25.0)
            v_["b"] = v_["b"] * np.float32(25.0)
    cpc = mpcc["cost_params"]
    stagec = costs.QuadraticCost(cpc["xerr"], cpc["uerr"], cpc["uref"], feat="cartpole", res_mult=cpc["res_mult"])
    y0c = (np.array([0.0, 0.0, np.pi, 0.0], np.float32) + np.random.default_rng(3).normal(0, 0.01, 4).astype(np.float32))[None]
    xrefc = np.zeros((HC + 1, 4), np.float32)
    keyc = np.array([0, 7], np.uint32)
    x0c = torch.full((HC, 1), 1e-4, device=dev)
    engc = Engine(modc)
    tsc = np.full(HC, 0.02, np.float32)
    out = []
    for spec in (False, True):
        s = BatchedAPG(engc, nnc, stagec, mpcc["apg_mpc"], tsc, NC, lb=-1.0, ub=1.0, speculative=spec)
        out.append(run(s, x0c, y0c, xrefc, keyc))
    print("cartpole N=%d: two-launch %.3f ms (cost %.6f)   speculative %.3f ms (cost %.6f)" %
          (NC, out[0][0], out[0][1], out[1][0], out[1][1]), flush=True)
