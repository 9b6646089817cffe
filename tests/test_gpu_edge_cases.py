"""C-ABI edge cases: argument validation (every failure is an error code + message, never a silent fallback or a
crash) and ragged sizes -- horizons of one step, particle counts that fill neither a lane group nor a warp, problem
counts that leave a partly filled last block."""
import ctypes as C

import numpy as np
import pytest
import torch

import helpers as hp
from oracle import nsde_oracle as orc
from sde4mbrl_b200 import _lib as L
from sde4mbrl_b200 import configs, costs

pytestmark = pytest.mark.gpu

ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE = -1, -2, -3, -4


def _rollout_args(eng, pk, P, N, H, dev):
    """A VALID sde4_rollout_args for the hexacopter (given noise), plus the tensors that keep its buffers alive."""
    t = {"y0": torch.as_tensor(np.tile(configs.hover_state(), (P, 1)), device=dev),
         "u": torch.full((P, H, eng.n_u), 0.33, device=dev),
         "dt": torch.full((H,), 0.01, device=dev),
         "noise": torch.zeros((H, eng.n_x, P * N), device=dev),
         "out": torch.empty((H + 1, eng.n_x, P * N), device=dev)}
    a = L.RolloutArgs()
    a.struct_size = C.sizeof(L.RolloutArgs)
    a.P, a.N, a.H = P, N, H
    a.params, a.y0, a.u, a.dt = pk.dev.data_ptr(), t["y0"].data_ptr(), t["u"].data_ptr(), t["dt"].data_ptr()
    a.u_stride_p, a.u_stride_t, a.u_stride_j = t["u"].stride()
    a.noise_mode, a.noise = L.NOISE_GIVEN, t["noise"].data_ptr()
    a.xevol, a.x_rows = t["out"].data_ptr(), eng.n_x
    return a, t


def _cost_args(eng, pk, stage, P, N, H, dev, grad=True):
    lib = L.load()
    ws_bytes = lib.sde4_cost_workspace_bytes(C.byref(eng.desc), P, N, H, 0, int(grad), 0)
    t = {"y0": torch.as_tensor(np.tile(configs.hover_state(), (P, 1)), device=dev),
         "opt": torch.full((P, H, eng.n_u), 0.33, device=dev),
         "dt": torch.full((H,), 0.01, device=dev),
         "xref": torch.as_tensor(np.tile(configs.hover_state(), (H + 1, 1)), device=dev),
         "noise": torch.zeros((H, eng.n_x, P * N), device=dev),
         "cost": torch.full((P,), -7.0, device=dev), "grad": torch.full((P, H, eng.n_u), -7.0, device=dev),
         "ws": torch.empty(ws_bytes + 512, dtype=torch.uint8, device=dev)}
    a = L.CostArgs()
    a.struct_size = C.sizeof(L.CostArgs)
    a.P, a.N, a.H = P, N, H
    a.params, a.y0, a.opt, a.dt = pk.dev.data_ptr(), t["y0"].data_ptr(), t["opt"].data_ptr(), t["dt"].data_ptr()
    a.opt_stride_p, a.opt_stride_t, a.opt_stride_j = t["opt"].stride()
    a.xref, a.xref_stride_p = t["xref"].data_ptr(), 0
    a.noise_mode, a.noise = L.NOISE_GIVEN, t["noise"].data_ptr()
    a.stage = C.pointer(stage.desc)
    a.cost, a.grad = t["cost"].data_ptr(), (t["grad"].data_ptr() if grad else None)
    base = t["ws"].data_ptr()
    a.workspace, a.workspace_bytes = (base + 255) & ~255, ws_bytes
    return a, t, ws_bytes


def _expect(rc, code, text):
    assert rc == code, (rc, L.last_error())
    assert text in L.last_error(), L.last_error()


def test_rollout_argument_validation(cuda):
    from sde4mbrl_b200.engine import Engine
    pm, model, nn = hp.build("hexa", 4, 3)
    eng = Engine(model)
    pk = eng.packed(nn)
    lib, dev = L.load(), pk.dev.device
    call = lambda a: lib.sde4_rollout(C.byref(eng.desc), C.byref(a), None)
    a, keep = _rollout_args(eng, pk, 2, 3, 4, dev)
    assert call(a) == 0  # the template itself is valid
    torch.cuda.synchronize()
    for field in ("P", "N", "H"):  # empty problem / particle / time axes are rejected, not silently skipped
        b, _k = _rollout_args(eng, pk, 2, 3, 4, dev)
        setattr(b, field, 0)
        _expect(call(b), ERR_INVALID, "must be positive")
        setattr(b, field, -5)
        _expect(call(b), ERR_INVALID, "must be positive")
    b, _k = _rollout_args(eng, pk, 2, 3, 4, dev)
    b.struct_size -= 4
    _expect(call(b), ERR_INVALID, "size mismatch")
    for field in ("params", "y0", "u", "dt", "xevol"):
        b, _k = _rollout_args(eng, pk, 2, 3, 4, dev)
        setattr(b, field, None)
        _expect(call(b), ERR_INVALID, "null buffer")
    b, _k = _rollout_args(eng, pk, 2, 3, 4, dev)
    b.params += 4
    _expect(call(b), ERR_INVALID, "aligned")
    b, _k = _rollout_args(eng, pk, 2, 3, 4, dev)
    b.noise = None
    _expect(call(b), ERR_INVALID, "given noise needs a buffer")
    b, _k = _rollout_args(eng, pk, 2, 3, 4, dev)
    b.noise_mode = L.NOISE_THREEFRY
    _expect(call(b), ERR_INVALID, "needs a key")
    b, _k = _rollout_args(eng, pk, 2, 3, 4, dev)
    b.noise_mode = 9
    _expect(call(b), ERR_INVALID, "bad noise/rng mode")
    b, _k = _rollout_args(eng, pk, 2, 3, 4, dev)
    b.x_rows = eng.n_x - 1
    _expect(call(b), ERR_INVALID, "x_rows")
    b, _k = _rollout_args(eng, pk, 2, 3, 4, dev)
    b.solver = 7
    _expect(call(b), ERR_INVALID, "unknown solver")
    b, _k = _rollout_args(eng, pk, 2, 3, 4, dev)
    b.particle_offset, b.particle_total = 2, 4  # 2 + 3 particles > 4
    _expect(call(b), ERR_INVALID, "particle shard")
    b, _k = _rollout_args(eng, pk, 2, 3, 4, dev)
    b.lanes = 16  # not compiled for this architecture
    _expect(call(b), ERR_UNSUPPORTED, "lane count")
    _expect(lib.sde4_rollout(None, C.byref(a), None), ERR_INVALID, "null")
    # an architecture nobody compiled: error names what IS compiled
    odd = type(eng.desc).from_buffer_copy(eng.desc)
    odd.density.h1 = 48
    _expect(lib.sde4_rollout(C.byref(odd), C.byref(a), None), ERR_UNSUPPORTED, "no compiled specialisation")
    assert "hexa_d32_swish" in L.last_error()
    # after all the rejected calls the valid one still runs and nothing was written out of bounds
    keep["out"].fill_(float("nan"))
    assert call(a) == 0
    torch.cuda.synchronize()
    assert bool(torch.isfinite(keep["out"]).all())


def test_cost_grad_argument_validation(cuda):
    from sde4mbrl_b200.engine import Engine
    pm, model, nn = hp.build("hexa", 4, 3)
    eng = Engine(model)
    pk = eng.packed(nn)
    cp = configs.hexa_mpc()["cost_params"]
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    lib, dev = L.load(), pk.dev.device
    call = lambda a: lib.sde4_cost_grad(C.byref(eng.desc), C.byref(a), None)
    a, keep, ws_bytes = _cost_args(eng, pk, stage, 2, 3, 4, dev)
    assert call(a) == 0
    torch.cuda.synchronize()
    good_cost, good_grad = keep["cost"].clone(), keep["grad"].clone()
    assert bool(torch.isfinite(good_cost).all()) and float(good_grad.abs().max()) > 0
    mk = lambda: _cost_args(eng, pk, stage, 2, 3, 4, dev)
    for field in ("P", "N", "H"):
        b, _k, _ = mk()
        setattr(b, field, 0)
        _expect(call(b), ERR_INVALID, "must be positive")
    b, _k, _ = mk()
    b.struct_size += 8
    _expect(call(b), ERR_INVALID, "size mismatch")
    for field in ("params", "y0", "opt", "dt", "xref", "cost", "workspace"):
        b, _k, _ = mk()
        setattr(b, field, None)
        assert call(b) == ERR_INVALID, (field, L.last_error())
    b, _k, _ = mk()
    b.stage = None
    assert call(b) == ERR_INVALID, L.last_error()
    b, _k, _ = mk()
    b.workspace_bytes = 64
    _expect(call(b), ERR_WORKSPACE, "workspace too small")
    b, _k, _ = mk()
    b.workspace += 16
    _expect(call(b), ERR_INVALID, "256-byte aligned")
    b, _k, _ = mk()
    b.lanes = 2
    _expect(call(b), ERR_UNSUPPORTED, "lane count")
    # a rejected call leaves the outputs untouched
    b, k2, _ = mk()
    b.H = 0
    assert call(b) == ERR_INVALID
    torch.cuda.synchronize()
    assert float(k2["cost"].min()) == -7.0 and float(k2["grad"].max()) == -7.0
    # and the valid call is reproducible afterwards, bit for bit
    assert call(a) == 0
    torch.cuda.synchronize()
    assert torch.equal(keep["cost"], good_cost) and torch.equal(keep["grad"], good_grad)
    # training entry point on an architecture without a training kernel
    _pm, prior_model, prior_nn = hp.build("hexa", 4, 3)
    assert lib.sde4_loss_workspace_bytes(C.byref(eng.desc), 2, 3, 4) > 0
    _expect(lib.sde4_rng_choice_mask(None, 4, 8, 3, 0, None, None), ERR_INVALID, "bad argument")
    buf = torch.zeros(64, device=dev)
    keys = torch.zeros((4, 2), dtype=torch.int32, device=dev)
    _expect(lib.sde4_rng_choice_mask(keys.data_ptr(), 4, 8, 9, 0, buf.data_ptr(), None), ERR_INVALID, "choice")
    _expect(lib.sde4_rng_choice_mask(keys.data_ptr(), 4, 2000, 9, 0, buf.data_ptr(), None), ERR_UNSUPPORTED, "1625")


@pytest.mark.parametrize("kind", ["msd", "cartpole", "hexa"])
@pytest.mark.parametrize("P,N,H,lanes", [(1, 1, 1, 0), (3, 5, 1, 0), (1, 33, 3, 0), (7, 1, 2, 0), (5, 3, 4, 1), (2, 37, 2, 8),
                                         (129, 1, 2, 0)])
def test_ragged_sizes_against_the_oracle(cuda, kind, P, N, H, lanes):
    """Horizon of one step, particle counts that fill neither a lane group nor a warp, problem counts that leave a
    partly filled last block: rollout, cost and gradient per problem vs the oracle on the same noise."""
    from sde4mbrl_b200.engine import Engine
    if lanes == 8 and kind == "msd":
        lanes = 4  # the mass-spring-damper kernels are compiled for 4 lanes
    pm, model, nn = hp.build(kind, H, N, seed=2)
    eng = Engine(model)
    rng = np.random.default_rng(P * 100 + N * 10 + H)
    y0 = np.stack([hp.start_state(kind, rng) for _ in range(P)])
    opt = np.stack([hp.controls(kind, H, rng) for _ in range(P)])
    dw = rng.standard_normal((P, N, H, model.n_x)).astype(np.float32)
    if kind == "hexa":
        cp = configs.hexa_mpc()["cost_params"]
        stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
        xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
        xref[:, :3] = [0.0, 0.5, 1.4]
    elif kind == "cartpole":
        cp = configs.cartpole_mpc()["cost_params"]
        stage = costs.QuadraticCost(cp["xerr"], cp["uerr"], cp["uref"], feat="cartpole", res_mult=cp["res_mult"])
        xref = np.zeros((H + 1, 4), np.float32)
    else:
        cp = {"feat": "identity", "xerr": [2.0, 0.5], "uerr": [0.1], "uref": [0.0], "res_mult": 1.0}
        stage = costs.QuadraticCost(cp["xerr"], cp["uerr"], cp["uref"])
        xref = np.tile(np.array([0.05, 0.0], np.float32), (H + 1, 1))
    ts = orc.compute_timesteps(pm)
    # native noise layout [H, n_x, P*N], sample g = p*N + n
    noise = np.ascontiguousarray(np.transpose(dw.reshape(P * N, H, model.n_x), (1, 2, 0)))
    x = eng.rollout(nn, y0, opt, ts, N=N, noise=noise, lanes=lanes)
    cost, grad, _ = eng.cost_grad(nn, y0, opt, ts, xref, stage.desc, N=N, noise=noise, want_grad=True, lanes=lanes)
    x = x.cpu().numpy().reshape(H + 1, model.n_x, P, N)
    om64 = hp.oracle_model(kind, pm, nn, torch.float64)
    fn = hp.stage_cost_fn(kind, cp)
    for p in sorted(set([0, P // 2, P - 1])):
        o = torch.tensor(opt[p], dtype=torch.float64, requires_grad=True)
        c, xt = orc.compute_cost(om64, torch.tensor(y0[p], dtype=torch.float64), o, torch.tensor(dw[p], dtype=torch.float64), fn,
                                 torch.tensor(xref, dtype=torch.float64))
        if kind == "hexa":
            c = c + orc.u_slew_cost(o, ts, cp, om64.n_u)
        (g,) = torch.autograd.grad(c, o)
        assert abs(float(cost[p]) - float(c)) <= 2e-5 * abs(float(c)) + 1e-7
        assert hp.rel_err(grad[p].cpu().numpy(), g.numpy()) < 1e-3
        want_x = xt.detach().numpy()[..., :model.n_x]  # [N, H+1, n_x]
        assert hp.rel_err(np.transpose(x[:, :, p, :], (2, 0, 1)), want_x) < 2e-5
