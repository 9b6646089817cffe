"""Thin launcher over the C-ABI: owns the device copies of the packed parameters, the time
steps and a grow-only workspace, and turns torch CUDA tensors into the pointer/stride structs of
``include/sde4b200.h``.  torch is plumbing only (device memory + streams); all arithmetic of the
hot path happens inside ``libsde4b200.so``.
"""
import ctypes as C
import zlib

import numpy as np
import torch

from . import _lib as L
from ._lib import Sde4Error


_HAVE_CUDA = None


def _dev():
    global _HAVE_CUDA
    if _HAVE_CUDA is None:
        _HAVE_CUDA = torch.cuda.is_available()
    if not _HAVE_CUDA:
        raise Sde4Error("sde4mbrl_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def as_f32(x, device=None):
    """numpy / list / torch -> float32 CUDA tensor (no copy if already there)."""
    if isinstance(x, torch.Tensor):
        if x.is_cuda and x.dtype == torch.float32 and device is None:
            return x
        return x.to(device=_dev() if device is None else device, dtype=torch.float32)
    arr = np.asarray(x, dtype=np.float32)
    if not arr.flags.writeable:  # torch refuses to alias read-only arrays silently
        arr = arr.copy()
    return torch.as_tensor(arr, device=_dev() if device is None else device)


def as_key(k, device=None):
    """jax-layout key(s) uint32[...,2] -> int32 CUDA tensor holding the same bits."""
    if isinstance(k, torch.Tensor) and k.is_cuda and k.dtype == torch.int32 and device is None and k.is_contiguous():
        return k
    device = _dev() if device is None else device
    if isinstance(k, torch.Tensor):
        if k.dtype == torch.uint32:
            k = k.view(torch.int32)
        elif k.dtype == torch.int64:
            k = (k & 0xFFFFFFFF).to(torch.int32) if False else k.to(torch.int32)
        if k.dtype != torch.int32:
            raise Sde4Error("keys must be uint32/int32 tensors")
        return k.to(device).contiguous()
    arr = np.asarray(k)
    if arr.dtype != np.uint32:
        arr = arr.astype(np.uint32)
    return torch.as_tensor(arr.view(np.int32), device=device).contiguous()


def host_f32(x):
    """numpy / list / CPU tensor -> C-contiguous float32 numpy array (no copy when it already is one)."""
    if isinstance(x, torch.Tensor):
        x = x.detach().numpy()
    return np.ascontiguousarray(x, dtype=np.float32)


def host_key(k):
    """jax-layout key(s) -> C-contiguous uint32 numpy array."""
    if isinstance(k, torch.Tensor):
        k = k.detach().cpu().numpy()
    k = np.asarray(k)
    if k.dtype == np.int32:
        k = k.view(np.uint32)
    return np.ascontiguousarray(k, dtype=np.uint32)


def is_host(x):
    return not (isinstance(x, torch.Tensor) and x.is_cuda)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """Raw handle of torch's current CUDA stream (the launches of this library are ordered with torch's own work)."""
    if _raw_stream is not None:  # ~0.3 us instead of ~4 us for the Stream object round trip
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def params_fingerprint(nn_params):
    """CRC-32 over the leaves of a haiku-style parameter tree, in tree order (~8 us for the hexacopter's 32 leaves).
    The engine caches the packed device block of the last tree it saw; a training loop that updates the leaves IN
    PLACE keeps the same dict object, so identity alone would keep launching with the first values."""
    c = 0
    for mod in nn_params.values():
        for leaf in (mod.values() if isinstance(mod, dict) else (mod,)):
            try:
                c = zlib.crc32(leaf, c)
            except (TypeError, ValueError, BufferError):  # python scalars, CPU tensors, non-contiguous arrays
                if isinstance(leaf, torch.Tensor):
                    leaf = leaf.detach().cpu().numpy()
                c = zlib.crc32(np.ascontiguousarray(np.asarray(leaf, dtype=np.float32)), c)
    return c


class PackedParams:
    """Device-resident packed parameter block (see ``sde4_param_layout``)."""

    def __init__(self, model, nn_params, desc, fingerprint=None):
        self.host = model.pack(nn_params, desc)
        self.dev = torch.as_tensor(self.host, device=_dev())
        self.source = nn_params
        self.fingerprint = params_fingerprint(nn_params) if fingerprint is None else fingerprint

    def private_copy(self):
        """A block with its own device buffer (for owners that bake the address into a CUDA graph)."""
        c = PackedParams.__new__(PackedParams)
        c.host, c.dev, c.source, c.fingerprint = self.host, self.dev.clone(), self.source, self.fingerprint
        return c

    def assign(self, other):
        """Take the values of ``other`` while keeping this block's device buffer (and address)."""
        self.dev.copy_(other.dev)
        self.host, self.source, self.fingerprint = other.host, other.source, other.fingerprint


class Engine:
    """One model (descriptor) bound to the current CUDA device."""

    def __init__(self, model, rng_mode=L.RNG_ORIGINAL):
        self.model = model
        self.desc = model.describe()
        L.load()
        self.offsets, self.param_total = model.layout(self.desc)
        self.n_x, self.n_u = model.n_x, model.n_u
        self.rng_mode = rng_mode
        self.noisy = model.is_noisy()
        self._packed = None
        self._ws = None
        self._dt_cache = {}
        self._dt_last = None
        self._ws_bytes = {}
        self._host_plans = {}
        self._stage = None

    # ---- parameters -----------------------------------------------------
    def packed(self, nn_params):
        """Device block of ``nn_params``.  Cached for the last tree seen, keyed on the tree's identity AND a CRC of its
        leaves, so leaves updated in place are re-packed (a ``PackedParams`` passes through untouched)."""
        if isinstance(nn_params, PackedParams):
            return nn_params
        fp = params_fingerprint(nn_params)
        pk = self._packed
        if pk is None or pk.source is not nn_params or pk.fingerprint != fp:
            self._packed = pk = PackedParams(self.model, nn_params, self.desc, fp)
        return pk

    def invalidate_params(self):
        """Forget the cached parameter block (the next call packs and uploads again)."""
        self._packed = None

    def dt_tensor(self, dt):
        hit = self._dt_last
        if hit is not None and hit[0] is dt:  # same READ-ONLY time-step array as last call
            return hit[1]
        arr = np.ascontiguousarray(np.asarray(dt, np.float32))
        key = arr.tobytes()
        t = self._dt_cache.get(key)
        if t is None:
            t = torch.as_tensor(np.array(arr), device=_dev())  # private writable copy (dt may be a read-only array)
            self._dt_cache[key] = t
        if isinstance(dt, np.ndarray) and not dt.flags.writeable:
            self._dt_last = (dt, t)
        return t

    def workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=_dev())
        return self._ws

    def _em_only(self):
        if self.model.params.get("sde_solver", "euler_maruyama") != "euler_maruyama":
            raise Sde4Error("the cost / gradient / training kernels integrate with euler_maruyama only "
                            "(the Stratonovich schemes are a rollout-only extension)")

    # ---- K1 --------------------------------------------------------------
    def rollout(self, nn_params, y0, u, dt, key=None, N=1, noise=None, noise_mode=None, x_rows=None,
                block_threads=0, out=None, particle_offset=0, particle_total=0, lanes=0, solver=None, sim=None):
        """y0[P,n_x], u[P,H,n_u] (or [H,n_u], shared), key [2] or [P,2].  Returns the native
        buffer ``xevol[H+1, x_rows, P*N]`` (sample-minor)."""
        pk = self.packed(nn_params)
        y0 = as_f32(y0).reshape(-1, self.n_x).contiguous()
        P = y0.shape[0]
        dt_t = self.dt_tensor(dt)
        H = dt_t.numel()
        u = as_f32(u)
        if u.dim() == 2:
            sp, (st, sj) = 0, u.stride()
        else:
            sp, st, sj = u.stride()
            if u.shape[0] == 1:
                sp = 0
            elif u.shape[0] != P:
                raise Sde4Error("u batch size does not match y0")
        if u.shape[-2] != H or u.shape[-1] != self.n_u:
            raise AssertionError("Dimension mismatch on the input of the sde solver")
        G = P * N
        x_rows = self.n_x if x_rows is None else x_rows
        if out is None and sim is None:
            out = torch.empty((H + 1, x_rows, G), dtype=torch.float32, device=y0.device)
        if noise_mode is None:
            noise_mode = L.NOISE_GIVEN if noise is not None else (L.NOISE_THREEFRY if self.noisy else L.NOISE_NONE)
        a = L.RolloutArgs()
        a.struct_size = C.sizeof(L.RolloutArgs)
        a.P, a.N, a.H = P, N, H
        a.params, a.y0, a.u = _ptr(pk.dev), _ptr(y0), _ptr(u)
        a.u_stride_p, a.u_stride_t, a.u_stride_j = sp, st, sj
        a.dt = _ptr(dt_t)
        keyt = None
        if noise_mode == L.NOISE_THREEFRY:
            keyt = as_key(key).reshape(-1, 2)
            a.key = _ptr(keyt)
            a.key_stride_p = 0 if keyt.shape[0] == 1 else 2
            if keyt.shape[0] not in (1, P):
                raise Sde4Error("need one key or one key per problem")
        if noise_mode == L.NOISE_GIVEN:
            noise = as_f32(noise).contiguous()
            if tuple(noise.shape) != (H, self.n_x, G):
                raise Sde4Error("noise must be [H, n_x, P*N] (sample-minor), got %s" % (tuple(noise.shape),))
            a.noise = _ptr(noise)
        a.rng_mode, a.noise_mode = self.rng_mode, noise_mode
        a.xevol, a.x_rows, a.block_threads = _ptr(out), x_rows, block_threads
        a.particle_offset, a.particle_total, a.lanes = particle_offset, particle_total, lanes
        name = self.model.params.get("sde_solver", "euler_maruyama") if solver is None else solver
        if name not in L.SOLVERS:
            raise Sde4Error("unknown sde_solver '%s' (known: %s)" % (name, ", ".join(L.SOLVERS)))
        a.solver = L.SOLVERS[name]
        if sim is not None:  # _lib.SimOut: simulator summary (last state, particle mean / std, rewards); `out` may be None
            L.check(L.load().sde4_sim_rollout(C.byref(self.desc), C.byref(a), C.byref(sim), _stream()))
            return out
        L.check(L.load().sde4_rollout(C.byref(self.desc), C.byref(a), _stream()))
        return out

    def sim_rollout(self, nn_params, y0, u, dt, key, N=1, reward=None, want_std=False, want_step=False, want_traj=False,
                    lanes=0):
        """The nSDE as a simulator: P start states x N particles x H steps, WITHOUT writing the trajectories
        (``sde4_sim_rollout``).  Returns a dict: ``mean[P, n_x]`` (particle mean of the last state), ``std`` (population
        std, with ``want_std``), ``last[n_x, P*N]``; with ``reward='cartpole_swingup'`` also ``reward_mean[P]`` (per-particle
        reward summed over the steps up to termination, averaged over particles) and, with ``want_step``,
        ``reward_of_mean[P]`` / ``done_of_mean[P]`` of the mean next state (mbrl ModelEnv.step)."""
        y0 = as_f32(y0).reshape(-1, self.n_x).contiguous()
        P, dev = y0.shape[0], y0.device
        G = P * N
        kind = {None: L.REWARD_NONE, "cartpole_swingup": L.REWARD_CARTPOLE_SWINGUP}[reward]
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        res = {"last": f(self.n_x, G), "mean": f(P, self.n_x)}
        so = L.SimOut()
        so.reward_kind = kind
        so.last, so.mean = res["last"].data_ptr(), res["mean"].data_ptr()
        if want_std:
            res["std"] = f(P, self.n_x)
            so.std = res["std"].data_ptr()
        if kind:
            res["racc"], res["reward_mean"] = f(G), f(P)
            so.racc, so.reward_mean = res["racc"].data_ptr(), res["reward_mean"].data_ptr()
            if want_step:
                res["reward_of_mean"] = f(P)
                res["done_of_mean"] = torch.empty(P, dtype=torch.int32, device=dev)
                so.reward_of_mean, so.done_of_mean = res["reward_of_mean"].data_ptr(), res["done_of_mean"].data_ptr()
        out = torch.empty((np.asarray(dt).shape[0] + 1, self.n_x, G), dtype=torch.float32, device=dev) if want_traj else None
        res["xevol"] = self.rollout(nn_params, y0, u, dt, key=key, N=N, out=out, lanes=lanes, sim=so)
        return res

    # ---- K2 + K3 ----------------------------------------------------------
    def cost_grad(self, nn_params, y0, opt, dt, xref, stage, terminal=None, key=None, N=1, noise=None,
                  noise_mode=None, want_grad=True, want_xtraj=False, block_threads=0, grad_out=None,
                  particle_offset=0, particle_total=0, lanes=0, data_mod=0, cost_out=None, problem_mask=None,
                  workspace=None, xgpu=None):
        """y0[P,n_x], opt[P,H,n_opt] (any strides) or [H,n_opt], xref[P,H+1,n_x] or [H+1,n_x].
        Returns (cost[P], grad (same shape/strides as opt) or None, xtraj[H+1,n_x+1,G] or None)."""
        self._em_only()
        pk = self.packed(nn_params)
        y0 = as_f32(y0).reshape(-1, self.n_x).contiguous()
        P = y0.shape[0]
        dt_t = self.dt_tensor(dt)
        H = dt_t.numel()
        opt = as_f32(opt)
        if data_mod:
            assert P == data_mod, "data_mod must equal the number of problem data rows"
            P = opt.shape[0]  # stacked candidate evaluations: opt holds k*data_mod rows
        n_opt = self.n_u + stage.n_slack
        squeeze = opt.dim() == 2
        if squeeze:
            opt = opt.unsqueeze(0)
        if opt.shape[0] != P or opt.shape[1] != H or opt.shape[2] != n_opt:
            raise AssertionError("Shape of the opt params do not match: %s vs (%d,%d,%d)" % (tuple(opt.shape), P, H, n_opt))
        xref = as_f32(xref)
        if xref.dim() == 2:
            xref = xref.contiguous()
            xsp = 0
        else:
            xref = xref.contiguous()
            xsp = xref.stride(0) if xref.shape[0] > 1 else 0
        if xref.shape[-2] != H + 1 or xref.shape[-1] != self.n_x:
            raise Sde4Error("xref must be [H+1, n_x]")
        G = P * N
        lib = L.load()
        wkey = (P, N, H, stage.n_slack, bool(want_grad), bool(want_xtraj))
        ws_bytes = self._ws_bytes.get(wkey)
        if ws_bytes is None:
            ws_bytes = lib.sde4_cost_workspace_bytes(C.byref(self.desc), P, N, H, stage.n_slack, int(want_grad),
                                                     int(want_xtraj))
            self._ws_bytes[wkey] = ws_bytes
        if workspace is None:
            ws = self.workspace(ws_bytes)
        else:  # caller-owned scratch (a CUDA graph bakes its address in: it must not be the engine's grow-only one)
            ws = workspace
            if ws.numel() * ws.element_size() < ws_bytes:
                raise Sde4Error("caller-supplied workspace too small: %d < %d bytes" % (ws.numel() * ws.element_size(), ws_bytes))
        cost = torch.empty(P, dtype=torch.float32, device=y0.device) if cost_out is None else cost_out
        grad = None
        if want_grad:
            if grad_out is None:
                grad = torch.empty_strided(opt.shape, opt.stride(), dtype=torch.float32, device=y0.device)
            else:
                # K3 writes the gradient with opt's element strides (misc_kernels.cuh): the output must be laid out alike
                grad = grad_out.unsqueeze(0) if (squeeze and grad_out.dim() == 2) else grad_out
                if tuple(grad.shape) != tuple(opt.shape) or tuple(grad.stride()) != tuple(opt.stride()):
                    opt = opt.contiguous()
                    if tuple(grad.shape) != tuple(opt.shape) or tuple(grad.stride()) != tuple(opt.stride()):
                        raise Sde4Error("grad_out must have the shape and strides of opt (got %s/%s vs %s/%s)" % (
                            tuple(grad.shape), tuple(grad.stride()), tuple(opt.shape), tuple(opt.stride())))
        xtraj = torch.empty((H + 1, self.n_x + 1, G), dtype=torch.float32, device=y0.device) if want_xtraj else None
        if noise_mode is None:
            noise_mode = L.NOISE_GIVEN if noise is not None else (L.NOISE_THREEFRY if self.noisy else L.NOISE_NONE)
        a = L.CostArgs()
        a.struct_size = C.sizeof(L.CostArgs)
        a.P, a.N, a.H = P, N, H
        a.params, a.y0, a.opt = _ptr(pk.dev), _ptr(y0), _ptr(opt)
        a.opt_stride_p, a.opt_stride_t, a.opt_stride_j = opt.stride()
        a.dt, a.xref, a.xref_stride_p = _ptr(dt_t), _ptr(xref), xsp
        keyt = None
        if noise_mode == L.NOISE_THREEFRY:
            keyt = as_key(key).reshape(-1, 2)
            if keyt.shape[0] not in (1, P, data_mod):
                raise Sde4Error("need one key or one key per problem")
            a.key = _ptr(keyt)
            a.key_stride_p = 0 if keyt.shape[0] == 1 else 2
        if noise_mode == L.NOISE_GIVEN:
            noise = as_f32(noise).contiguous()
            if tuple(noise.shape) != (H, self.n_x, G):
                raise Sde4Error("noise must be [H, n_x, P*N] (sample-minor), got %s" % (tuple(noise.shape),))
            a.noise = _ptr(noise)
        a.rng_mode, a.noise_mode = self.rng_mode, noise_mode
        a.stage = C.pointer(stage)
        a.terminal = C.pointer(terminal) if terminal is not None else None
        a.cost, a.grad, a.xtraj = _ptr(cost), _ptr(grad), _ptr(xtraj)
        a.workspace, a.workspace_bytes = _ptr(ws), ws.numel() * ws.element_size()
        a.block_threads = block_threads
        a.particle_offset, a.particle_total, a.lanes = particle_offset, particle_total, lanes
        a.data_mod = data_mod
        if xgpu is not None:  # _lib.XgpuReduce: the cross-GPU sum of [cost | grad] happens inside K3 (dist.CostGradReducer)
            a.xgpu = C.addressof(xgpu)
        if problem_mask is not None:  # int32[P] device tensor: problems with 0 are skipped, their outputs untouched
            assert problem_mask.dtype == torch.int32 and problem_mask.is_cuda and problem_mask.numel() == P
            a.problem_mask = problem_mask.data_ptr()
        L.check(lib.sde4_cost_grad(C.byref(self.desc), C.byref(a), _stream()))
        if squeeze and grad is not None:
            grad = grad.squeeze(0)
        return cost, grad, xtraj

    # ---- K2 + K3 with HOST buffers (one packed copy each way) ----------------
    def cost_grad_host(self, nn_params, y0, opt, dt, xref, stage, terminal=None, key=None, N=1, want_grad=True,
                       want_xtraj=False, block_threads=0, lanes=0, particle_offset=0, particle_total=0, xgpu=None):
        """Same evaluation as ``cost_grad`` for arguments that live in HOST memory (numpy arrays / CPU
        tensors): ``sde4_cost_grad_host`` packs them into a pinned staging buffer, does one H2D copy,
        the launches, one D2H copy of [cost | grad] and a stream synchronise.  Returns numpy
        ``cost[P]``, ``grad`` (shape of ``opt``) or None, and the device ``xtraj`` or None.
        ``particle_offset`` / ``particle_total`` / ``xgpu`` (``dist.CostGradReducer.launch_kwargs``): this rank's
        particles of a problem sharded over GPUs; K3 sums the world's partials and the TOTAL comes back."""
        self._em_only()
        pk = self.packed(nn_params)
        y0 = host_f32(y0).reshape(-1, self.n_x)
        P = y0.shape[0]
        dt_t = self.dt_tensor(dt)
        H = dt_t.numel()
        opt = host_f32(opt)
        n_opt = self.n_u + stage.n_slack
        squeeze = opt.ndim == 2
        opt3 = opt[None] if squeeze else opt
        if opt3.shape != (P, H, n_opt):
            raise AssertionError("Shape of the opt params do not match: %s vs (%d,%d,%d)" % (opt3.shape, P, H, n_opt))
        xref = host_f32(xref)
        if xref.shape[-2:] != (H + 1, self.n_x) or (xref.ndim == 3 and xref.shape[0] not in (1, P)):
            raise Sde4Error("xref must be [H+1, n_x] or [P, H+1, n_x]")
        xsp = (H + 1) * self.n_x if (xref.ndim == 3 and xref.shape[0] == P and P > 1) else 0
        lib = L.load()
        dev = pk.dev.device
        keyed = self.noisy
        keyh = None
        if keyed:
            keyh = host_key(key).reshape(-1, 2)
            if keyh.shape[0] not in (1, P):
                raise Sde4Error("need one key or one key per problem")
        # everything that only depends on the call shape is prepared once per shape
        pkey = (P, N, H, bytes(stage), None if terminal is None else bytes(terminal), bool(want_grad), bool(want_xtraj),
                lanes, block_threads, xsp,
                0 if keyh is None else keyh.shape[0])
        plan = self._host_plans.get(pkey)
        if plan is None:
            ws_bytes = lib.sde4_cost_workspace_bytes(C.byref(self.desc), P, N, H, stage.n_slack, int(want_grad),
                                                     int(want_xtraj))
            sbytes = lib.sde4_cost_grad_host_staging_bytes(C.byref(self.desc), P, H, stage.n_slack)
            a = L.CostArgs()
            a.struct_size = C.sizeof(L.CostArgs)
            a.P, a.N, a.H = P, N, H
            a.opt_stride_p, a.opt_stride_t, a.opt_stride_j = H * n_opt, n_opt, 1
            a.dt, a.xref_stride_p = _ptr(dt_t), xsp
            a.key_stride_p = 0 if (keyh is None or keyh.shape[0] == 1) else 2
            a.noise_mode = L.NOISE_THREEFRY if keyed else L.NOISE_NONE
            a.stage = C.pointer(stage)
            a.terminal = C.pointer(terminal) if terminal is not None else None
            a.block_threads, a.lanes = block_threads, lanes
            if len(self._host_plans) > 64:
                self._host_plans.clear()
            plan = (a, ws_bytes, sbytes, stage, terminal, dt_t)  # keeps the pointed-to descriptors alive
            self._host_plans[pkey] = plan
        a, ws_bytes, sbytes = plan[0], plan[1], plan[2]
        ws = self.workspace(ws_bytes)
        if self._stage is None or self._stage[0].numel() < sbytes:
            self._stage = (torch.empty(sbytes, dtype=torch.uint8).pin_memory(),
                           torch.empty(sbytes, dtype=torch.uint8, device=dev))
        cost = np.empty(P, np.float32)
        grad = np.empty_like(opt3) if want_grad else None
        xtraj = torch.empty((H + 1, self.n_x + 1, P * N), dtype=torch.float32, device=dev) if want_xtraj else None
        a.params, a.y0, a.opt, a.xref = pk.dev.data_ptr(), y0.ctypes.data, opt3.ctypes.data, xref.ctypes.data
        a.dt = dt_t.data_ptr()  # the plan is keyed on the call shape only: a different time-step array with the same H
        a.key = keyh.ctypes.data if keyed else None
        a.rng_mode = self.rng_mode
        a.cost, a.grad = cost.ctypes.data, (grad.ctypes.data if want_grad else None)
        a.xtraj = xtraj.data_ptr() if want_xtraj else None
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        a.particle_offset, a.particle_total = particle_offset, particle_total
        a.xgpu = C.addressof(xgpu) if xgpu is not None else None
        L.check(lib.sde4_cost_grad_host(C.byref(self.desc), C.byref(a), _ptr(self._stage[0]), _ptr(self._stage[1]),
                                        sbytes, _stream()))
        if squeeze and grad is not None:
            grad = grad[0]
        return cost, grad, xtraj

    # ---- K2 (training variant) + K3 + K5 -----------------------------------
    def loss_grad(self, nn_params, ymeas, u, dt, data_cost, key, N=1, noise=None, noise_mode=None, block_threads=0, lanes=0,
                  step_weight=None, particle_offset=0, particle_total=0):
        """Data term of the training loss and its gradient w.r.t. the PACKED parameter block.
        ymeas[B, H+1, n_x] (rows = solver steps), u[B, H, n_u], key[B, 2] = split(rng, B).
        Returns (loss[B], grad_packed[param_total]); d(mean_b loss)/d params = grad_packed.
        ``particle_offset`` / ``particle_total``: this launch evaluates particles [offset, offset + N) of
        ``particle_total`` per element (their keys, the 1 / particle_total of both means); the partial results of the
        launches that cover all particles add up."""
        self._em_only()
        pk = self.packed(nn_params)
        ymeas = as_f32(ymeas).contiguous()
        B = ymeas.shape[0]
        dt_t = self.dt_tensor(dt)
        H = dt_t.numel()
        u = as_f32(u).contiguous()
        assert tuple(ymeas.shape) == (B, H + 1, self.n_x) and tuple(u.shape) == (B, H, self.n_u)
        y0 = ymeas[:, 0, :].contiguous()
        lib = L.load()
        ws_bytes = lib.sde4_loss_workspace_bytes(C.byref(self.desc), B, N, H)
        if ws_bytes == 0:
            raise Sde4Error("no training-loss kernel compiled for this architecture")
        ws = self.workspace(ws_bytes)
        loss = torch.empty(B, dtype=torch.float32, device=y0.device)
        gpar = torch.empty(self.param_total, dtype=torch.float32, device=y0.device)
        if noise_mode is None:
            noise_mode = L.NOISE_GIVEN if noise is not None else (L.NOISE_THREEFRY if self.noisy else L.NOISE_NONE)
        a = L.CostArgs()
        a.struct_size = C.sizeof(L.CostArgs)
        a.P, a.N, a.H = B, N, H
        a.params, a.y0, a.opt = _ptr(pk.dev), _ptr(y0), _ptr(u)
        a.opt_stride_p, a.opt_stride_t, a.opt_stride_j = u.stride()
        a.dt, a.xref, a.xref_stride_p = _ptr(dt_t), _ptr(ymeas), ymeas.stride(0)
        if noise_mode == L.NOISE_THREEFRY:
            keyt = as_key(key).reshape(-1, 2)
            if keyt.shape[0] != B:
                raise Sde4Error("need one key per minibatch element (split(rng, B))")
            a.key, a.key_stride_p = _ptr(keyt), 2
        if noise_mode == L.NOISE_GIVEN:
            noise = as_f32(noise).contiguous()
            if tuple(noise.shape) != (H, self.n_x, B * N):
                raise Sde4Error("noise must be [H, n_x, B*N] (sample-minor)")
            a.noise = _ptr(noise)
        a.rng_mode, a.noise_mode = self.rng_mode, noise_mode
        a.stage = C.pointer(data_cost)
        a.cost = _ptr(loss)
        a.workspace, a.workspace_bytes = _ptr(ws), ws.numel()
        a.block_threads = block_threads
        a.particle_offset, a.particle_total = particle_offset, particle_total
        a.lanes = lanes  # 0: lane-split training kernel while the minibatch is latency bound; 1: one thread per sample
        if step_weight is not None:  # float32[H, B*N]: weight of the data term of sample g at time t+1 (num_sample2consider)
            assert step_weight.dtype == torch.float32 and step_weight.is_cuda and tuple(step_weight.shape) == (H, B * N)
            step_weight = step_weight.contiguous()
            a.step_weight = step_weight.data_ptr()
        L.check(lib.sde4_loss_grad(C.byref(self.desc), C.byref(a), _ptr(gpar), _stream()))
        return loss, gpar
