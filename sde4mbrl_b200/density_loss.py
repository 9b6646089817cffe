"""Distance-aware regulariser of the training loss: gradient-norm, local strong-convexity hinge and
mu-coefficient terms of ``ControlledSDE.density_loss`` / ``mu_coeff_nn`` / ``sample_for_loss_computation``
and their weighting in ``create_model_loss_fn.loss_fn``
(``sde4mbrl/nsde.py:597-680, 771-804, 1205-1254``).

NOT the sequential hot path: the terms are evaluated at the DATA points (``ymeas[:-1]``, ``uVal``), i.i.d. over
(minibatch, particle, time), with no rollout.  That is a batch of K = B*N*H_d*(1+ball_nsamples) independent
evaluations of two small MLPs plus a second-order derivative (the parameter gradient of ``|grad_z density|^2``),
i.e. plain library GEMMs + reverse-over-reverse differentiation -- the work the reference leaves to XLA.  Here it is
torch (cuBLAS) on the device with ``torch.autograd``; what is hand-written is the bit-exact key chain and the ball
noise (``sde4_rng_split_many`` / ``sde4_rng_normal_many``), so that the terms are comparable with the reference for
the same ``rng``.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import random as _random
from ._lib import Sde4Error

_ACT = {"tanh": torch.tanh, "swish": torch.nn.functional.silu, "sigmoid": torch.sigmoid, "relu": torch.relu}


def _mlp(tree, prefix, z, act_name):
    """``hk.nets.MLP``: ``act(z @ w + b)`` for all but the last layer."""
    act = _ACT[act_name]
    n = 0
    while "%s/linear_%d" % (prefix, n) in tree:
        n += 1
    if n == 0:
        raise Sde4Error("no MLP parameters under " + prefix)
    for k in range(n):
        leaf = tree["%s/linear_%d" % (prefix, k)]
        z = z @ leaf["w"] + leaf["b"]
        if k + 1 < n:
            z = act(z)
    return z


def reduced_inputs(model, x, u):
    """``noise_aug_state_ctrl(x, u)`` (nsde.py:353-363) on batches: reduced state of the model, restricted to
    ``indx_noise_in``."""
    p = model.params
    name = model.module_name
    if name == "cart_pole_sde" and "state_scaling" in p:  # cartpole_sde.py:41
        red = torch.stack((x[..., 1], torch.sin(x[..., 2]), torch.cos(x[..., 2]), x[..., 3]), dim=-1) / float(p["state_scaling"][-1])
    elif name == "sde_rotor_model" and "state_scaling" in p:  # sde_rotor_model.py:139-144
        red = x[..., :13] / float(max(p["state_scaling"]))
    else:
        red = x
    nin = list(p["diffusion_density_nn"].get("indx_noise_in", range(model.n_x)))
    # (columns picked by slicing, not by an index tensor: no host -> device copy, so the chain can be graph-captured)
    z = torch.stack([red[..., i] for i in nin], dim=-1) if len(nin) < red.shape[-1] else red
    if p.get("control_dependent_noise", False):  # nsde.py:362
        z = torch.cat((z, u.expand(*z.shape[:-1], u.shape[-1])), dim=-1)
    return z


def init_mu_params(model, dl, rng):
    """Parameters ``mu_coeff_nn`` adds to the tree (nsde.py:657-680)."""
    lm = dl["learn_mucoeff"]
    m = model.module_name
    if lm["type"] == "constant":
        return {}
    if lm["type"] == "global":
        return {m: {"mu_coeff": rng.uniform(-0.001, 0.001, ()).astype(np.float32)}}
    iv = lm.get("init_value", 0.01)
    nin = len(list(model.params["diffusion_density_nn"].get("indx_noise_in", range(model.n_x))))
    if model.params.get("control_dependent_noise", False):  # mu_coeff_nn reads the same augmented vector (nsde.py:640)
        nin += model.n_u
    dims = [nin] + list(lm["hidden_layers"]) + [1]
    return {"%s/~mu_coeff_nn/mu_coeff/~/linear_%d" % (m, k): {"w": rng.uniform(-iv, iv, (dims[k], dims[k + 1])).astype(np.float32),
                                                              "b": np.zeros(dims[k + 1], np.float32)}
            for k in range(len(dims) - 1)}


def density_keys(rng, B, N, Hd):
    """rng_ball for every (b, n, t): ``split(rng, B)[b]`` (nsde.py:1196) -> ``split(., N)[n]`` (:1187) ->
    ``split(., 3)[2]`` (:709) -> ``split(., Hd)[t]`` (:781) -> ``split(.)[1]`` (:615)."""
    kb = _random.split(rng, B)
    kbn = _random.split_many(kb, N)                      # [B, N, 2]
    kd = _random.split_many(kbn, 3)[:, :, 2].contiguous()  # [B, N, 2]
    kt = _random.split_many(kd, Hd)                      # [B, N, Hd, 2]
    return _random.split_many(kt, 2)[..., 1, :].contiguous()


def ball_radius(dl, dim, device):
    """``params['density_loss']['ball_radius']`` broadcast to the augmented vector (nsde.py:625-627) as a device tensor."""
    return torch.as_tensor(np.broadcast_to(np.asarray(dl["ball_radius"], np.float32), (dim,)).copy(), device=device)


def density_terms(model, tree, y, u, rng, N, dl, partial=None, radius=None):
    """Per (minibatch element, particle) terms of ``sample_for_loss_computation`` (nsde.py:771-804).
    ``tree``: torch parameter tree (requires_grad leaves), ``y[B, Hd+1, n_y]``, ``u[B, Hd, n_u]`` device tensors.
    ``partial = (y_values[B, H+1, n_y], u_values[B, H, n_u], step_weight[H, B*N])``: ``density_on_partial`` with
    ``num_sample2consider`` (nsde.py:794-797) -- the terms are taken at the solver-step data points the sample keeps
    (0/1 weights), with the keys ``split(rng_density, Hd)[idx]``.
    Returns ``grad_norm[B, N]`` (mean over time), ``sconvex[B, N]`` (sum over time), ``mu[B, N]`` (mean),
    ``density[B, N]`` (mean)."""
    if dl["learn_mucoeff"].get("quad_approx", False):
        raise Sde4Error("density_loss with quad_approx (density_loss_theory) is not built")
    B, Hd = y.shape[0], y.shape[1] - 1
    Hk = Hd  # rng_density is always split over the DATA horizon (nsde.py:790)
    wgt = None
    if partial is not None:
        y, u, sw = partial
        Hd = y.shape[1] - 1
        wgt = sw.t().reshape(B, N, Hd)
    m = model.module_name
    dn = model.params["diffusion_density_nn"]["density_nn"]
    pre = "%s/~construct_diffusion_density_nn/density/~" % m
    act = dn.get("activation_fn", "tanh")

    def density(zz):  # density_function, nsde.py:395
        return torch.sigmoid(_mlp(tree, pre, zz, act)[..., 0] - 7.0)

    z = reduced_inputs(model, y[:, :-1], u)  # [B, Hd, dim]: obs2state is the identity for the shipped models
    z = z.detach().requires_grad_(True)
    dim = z.shape[-1]
    den = density(z)  # [B, Hd]
    (gz,) = torch.autograd.grad(den.sum(), z, create_graph=True)  # rows are independent: d den_k / d z_k
    grad_norm = (gz ** 2).sum(-1)  # [B, Hd]
    ns = int(dl["ball_nsamples"])
    if radius is None:  # (a host -> device copy: the graphed path passes the device tensor in)
        radius = ball_radius(dl, dim, y.device)
    keys = density_keys(rng, B, N, Hk)[:, :, :Hd].contiguous()  # [B, N, Hd, 2]
    ball = _random.normal_many(keys, (ns, dim)) * radius  # [B, N, Hd, ns, dim]
    zb = z[:, None, :, None, :] + ball
    den_ball = density(zb)  # [B, N, Hd, ns]
    lm = dl["learn_mucoeff"]
    mu0 = float(dl["mu_coeff"])
    if lm["type"] == "constant":
        mu = torch.full_like(den, mu0)
    elif lm["type"] == "global":
        mu = (mu0 * torch.exp(tree[m]["mu_coeff"])).expand_as(den)
    else:
        mu = torch.exp(_mlp(tree, "%s/~mu_coeff_nn/mu_coeff/~" % m, z, lm["activation_fn"])[..., 0]) * mu0
    cond = den_ball - den[:, None, :, None] - (gz[:, None, :, None, :] * ball).sum(-1) \
        - 0.5 * mu[:, None, :, None] * (ball ** 2).sum(-1)
    sconvex = (torch.clamp(cond, max=0.0) ** 2).sum(-1)  # [B, N, Hd]
    expand = lambda a: a[:, None, :].expand(B, N, Hd)
    if wgt is not None:
        cnt = wgt.sum(-1)
        wmean = lambda a: (expand(a) * wgt).sum(-1) / cnt
        return wmean(grad_norm), (sconvex * wgt).sum(-1), wmean(mu), wmean(den)
    return expand(grad_norm).mean(-1), sconvex.sum(-1), expand(mu).mean(-1), expand(den).mean(-1)


def to_torch_tree(nn_params, device, names=("density", "mu_coeff")):
    """float32 device copies (requires_grad) of the leaves the regulariser depends on."""
    out = {}
    for mod, leaves in nn_params.items():
        if not (any(n in mod for n in names) or any(n in leaves for n in ("mu_coeff",))):
            continue
        out[mod] = {k: torch.as_tensor(np.asarray(v, np.float32), device=device).requires_grad_(True) for k, v in leaves.items()}
    return out


def weighted_terms(loss_params, dl, grad_norm, sconvex, mu):
    """The regulariser's share of ``total_sum`` (nsde.py:1205-1250) and the logged scalars."""
    loss_grad_density = grad_norm.mean(dim=1).mean()
    loss_density_scvex = sconvex.mean(dim=1).mean()
    mu_mean = mu.mean(dim=1).mean()
    mult = loss_params.get("pen_scvex_mult", 1.0)
    total = torch.zeros((), device=grad_norm.device)
    if loss_params.get("pen_grad_density", 0) > 0:
        p_mu = dl["mu_coeff"]
        eff = loss_params["pen_grad_density"] * p_mu * dl.get("grad_scaler", 1.0 / p_mu)
        total = total + eff * loss_grad_density * mult
    if loss_params.get("pen_density_scvex", 0) > 0:
        total = total + loss_params["pen_density_scvex"] * loss_density_scvex * mult
    loss_mu = torch.zeros((), device=grad_norm.device)
    if loss_params.get("pen_mu_coeff", 0) > 0:
        kind = loss_params.get("pen_mu_type", "quad_inv")
        if kind == "quad_inv":
            loss_mu = 1.0 / mu_mean ** 2
        elif kind == "lin_inv":
            loss_mu = 1.0 / mu_mean
        elif kind == "exp_inv":
            loss_mu = torch.exp(-mu_mean * loss_params["pen_mu_temp"])
        else:
            raise ValueError("Unknown pen_mu_type: {}. Choose from quad_inv, lin_inv, exp_inv".format(kind))
        total = total + loss_mu * loss_params["pen_mu_coeff"] * mult
    return total, {"gradDensity": loss_grad_density, "densitySCVEX": loss_density_scvex, "muCoeff": mu_mean,
                   "lossMuCoeff": loss_mu}


_INFO_KEYS = ("gradDensity", "densitySCVEX", "muCoeff", "lossMuCoeff")


class GraphedRegulariser:
    """The regulariser's value, logged scalars and parameter gradient as ONE CUDA-graph replay per call.

    ``density_terms`` + ``weighted_terms`` + the reverse-over-reverse backward are ~250 small launches that cost ~4 ms of
    host dispatch per call against well under 1 ms of GPU work at the cartpole_sde.yaml shape (B = 512, H = 30,
    20 ball samples; scripts/loss_api_breakdown.py).  The whole chain -- key derivation (``sde4_rng_split_many``), ball
    noise (``sde4_rng_normal_many``), the cuBLAS evaluations, ``torch.autograd.grad`` -- is captured once per input
    shape on static buffers; a call then copies [parameters | y | u | key] into them, replays, and reads back one
    packed result [total | 4 scalars | gradient block].  ``density_on_partial`` (per-sample step weights) keeps the
    eager path."""

    def __init__(self, model, loss_params, dl, num_particles):
        self.model, self.lp, self.dl, self.N = model, loss_params, dl, num_particles
        self._cache = {}

    def _leaves(self, nn_params):
        out = []
        for mod, leaves in nn_params.items():
            if not (any(n in mod for n in ("density", "mu_coeff")) or "mu_coeff" in leaves):
                continue
            for k, v in leaves.items():
                out.append((mod, k, tuple(np.shape(v))))
        return out

    @torch.enable_grad()
    def _body(self, st):
        p = st["flat"].clone().requires_grad_(True)
        tree, o = {}, 0
        for mod, k, shp in st["leaves"]:
            n = int(np.prod(shp)) if len(shp) else 1
            tree.setdefault(mod, {})[k] = p[o:o + n].view(shp)
            o += n
        gn, sc, mu, _ = density_terms(self.model, tree, st["y"], st["u"], st["key"], self.N, self.dl, None, st["radius"])
        dtot, dinfo = weighted_terms(self.lp, self.dl, gn, sc, mu)
        g = torch.autograd.grad(dtot, p, allow_unused=True)[0] if dtot.requires_grad else None
        scal = torch.stack([dtot.detach().float()] + [torch.as_tensor(dinfo[k], dtype=torch.float32, device=p.device).detach().reshape(())
                                                      for k in _INFO_KEYS])
        return torch.cat((scal, g if g is not None else torch.zeros_like(p)))

    def _capture(self, leaves, y, u, key):
        dev = y.device
        ntot = sum(int(np.prod(s)) if len(s) else 1 for _, _, s in leaves)
        st = {"leaves": leaves, "flat": torch.zeros(ntot, dtype=torch.float32, device=dev),
              "host": torch.zeros(ntot, dtype=torch.float32).pin_memory(),
              "y": y.clone(), "u": u.clone(), "key": key.clone()}
        with torch.no_grad():
            st["radius"] = ball_radius(self.dl, reduced_inputs(self.model, y[:, :-1], u).shape[-1], dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up off the capture (cuBLAS handles / workspaces, autograd threads)
            for _ in range(2):
                self._body(st)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        st["graph"] = torch.cuda.CUDAGraph()
        with torch.cuda.graph(st["graph"]):
            st["out"] = self._body(st)
        st["out_host"] = torch.zeros(st["out"].numel(), dtype=torch.float32).pin_memory()
        return st

    def __call__(self, nn_params, y, u, key):
        """-> (total, {logged scalars}, {mod: {leaf: numpy gradient}}) for device tensors y[B, Hd+1, n_y], u[B, Hd, n_u]
        and a device key int32[2]."""
        leaves = self._leaves(nn_params)
        sig = (tuple(y.shape), tuple(u.shape), tuple(leaves), str(y.device), _random.rng_mode())
        st = self._cache.get(sig)
        if st is None:
            if len(self._cache) >= 8:  # a training loop has one or two batch shapes; do not hoard graphs for more
                self._cache.pop(next(iter(self._cache)))
            st = self._cache[sig] = self._capture(leaves, y, u, key)
        o = 0
        hv = st["host"].numpy()
        for mod, k, shp in leaves:
            n = int(np.prod(shp)) if len(shp) else 1
            hv[o:o + n] = np.asarray(nn_params[mod][k], np.float32).reshape(-1)
            o += n
        st["flat"].copy_(st["host"], non_blocking=True)
        st["y"].copy_(y)
        st["u"].copy_(u)
        st["key"].copy_(key)
        st["graph"].replay()
        st["out_host"].copy_(st["out"], non_blocking=True)
        torch.cuda.current_stream(y.device).synchronize()
        res = st["out_host"].numpy()
        info = {k: float(res[1 + i]) for i, k in enumerate(_INFO_KEYS)}
        grads, o = {}, 1 + len(_INFO_KEYS)
        for mod, k, shp in leaves:
            n = int(np.prod(shp)) if len(shp) else 1
            grads.setdefault(mod, {})[k] = res[o:o + n].reshape(shp).copy()
            o += n
        return float(res[0]), info, grads


def regulariser_weights(loss_params, dl):
    """(w_grad, w_scvex, w_mu): the weights of gradDensity, densitySCVEX and lossMuCoeff in ``total_sum``
    (nsde.py:1223-1250), ``pen_scvex_mult`` folded in; a term whose penalty is not positive is left out."""
    mult = float(loss_params.get("pen_scvex_mult", 1.0))
    w_g = w_s = w_m = 0.0
    if loss_params.get("pen_grad_density", 0) > 0:
        p_mu = dl["mu_coeff"]
        w_g = loss_params["pen_grad_density"] * p_mu * dl.get("grad_scaler", 1.0 / p_mu) * mult
    if loss_params.get("pen_density_scvex", 0) > 0:
        w_s = loss_params["pen_density_scvex"] * mult
    if loss_params.get("pen_mu_coeff", 0) > 0:
        w_m = loss_params["pen_mu_coeff"] * mult
    return float(w_g), float(w_s), float(w_m)


class NativeRegulariser:
    """The regulariser's value, logged scalars and parameter gradient from the library's own kernels
    (``sde4_density_reg``, csrc/reg_kernels.cuh): centres -> ball samples -> scalars -> second-order columns -> the
    weight-gradient reduction; no library GEMM, no autograd graph.  One pinned H2D copy of the mu parameters (the
    density network travels in the engine's packed block), one D2H copy of [5 scalars | gradient blocks].

    Covers what the reference's configurations use: ``learn_mucoeff`` of type constant / global / dnn with
    hidden_layers [8, 8] and tanh; no ``quad_approx``; the terms at every data point (``density_on_partial`` keeps the
    torch path)."""

    def __init__(self, model, engine, loss_params, dl, num_particles):
        self.model, self.engine, self.lp, self.dl, self.N = model, engine, loss_params, dl, num_particles
        lm = dl["learn_mucoeff"]
        self.mu_type = {"constant": L.MU_CONSTANT, "global": L.MU_GLOBAL}.get(lm["type"], L.MU_DNN)
        self.dim = engine.desc.density.n_in
        self.mu_desc = L.NetDesc()
        self.mu_size = 0
        if self.mu_type == L.MU_DNN:
            hl = list(lm["hidden_layers"])
            act = {"tanh": L.ACT_TANH, "swish": L.ACT_SWISH}.get(lm["activation_fn"], -1)
            if len(hl) != 2:
                raise Sde4Error("mu_coeff network: two hidden layers expected")
            self.mu_desc.n_in, self.mu_desc.h1, self.mu_desc.h2, self.mu_desc.n_out, self.mu_desc.act = self.dim, hl[0], hl[1], 1, act
            up4 = lambda n: (n + 3) & ~3
            self.mu_size = self.dim * up4(hl[0]) + up4(hl[0]) + hl[0] * up4(hl[1]) + up4(hl[1]) + up4(hl[1]) + 4
        elif self.mu_type == L.MU_GLOBAL:
            self.mu_size = 1
        self._st = {}

    def supported(self, B=1, Hd=1):
        """Whether the library has regulariser kernels for this model and mu network."""
        if self.dl["learn_mucoeff"].get("quad_approx", False) or self.dim <= 0 or self.dim > L.REG_MAXD:
            return False
        if self.mu_type == L.MU_DNN and (self.mu_desc.h1, self.mu_desc.h2, self.mu_desc.act) != (8, 8, L.ACT_TANH):
            return False
        if int(self.dl["ball_nsamples"]) > 160:
            return False
        return L.load().sde4_density_reg_workspace_bytes(C.byref(self.engine.desc), B, self.N, Hd, int(self.dl["ball_nsamples"])) > 0

    def _state(self, B, Hd, dev):
        sig = (B, Hd, str(dev))
        st = self._st.get(sig)
        if st is None:
            ns = int(self.dl["ball_nsamples"])
            nb = L.load().sde4_density_reg_workspace_bytes(C.byref(self.engine.desc), B, self.N, Hd, ns)
            if nb == 0:
                raise Sde4Error("no regulariser kernels compiled for this architecture")
            self._st.clear()  # one batch shape at a time: the workspace holds the streams (MBs)
            nout = 5 + self.engine.param_total + self.mu_size
            st = self._st[sig] = {
                "ws": torch.empty(nb, dtype=torch.uint8, device=dev),
                "out": torch.zeros(nout, dtype=torch.float32, device=dev),
                "out_host": torch.zeros(nout, dtype=torch.float32).pin_memory(),
                "mu": torch.zeros(max(self.mu_size, 1), dtype=torch.float32, device=dev),
                "mu_host": torch.zeros(max(self.mu_size, 1), dtype=torch.float32).pin_memory()}
        return st

    def _pack_mu(self, nn_params, out):
        m = self.model.module_name
        if self.mu_type == L.MU_GLOBAL:
            out[0] = float(np.asarray(nn_params[m]["mu_coeff"]))
        elif self.mu_type == L.MU_DNN:
            self.model._pack_net(out, 0, nn_params, "%s/~mu_coeff_nn/mu_coeff/~" % m, self.mu_desc)

    def _unpack_mu(self, g):
        m = self.model.module_name
        if self.mu_type == L.MU_GLOBAL:
            return {m: {"mu_coeff": np.array(g[0], np.float32)}}
        if self.mu_type != L.MU_DNN:
            return {}
        up4 = lambda n: (n + 3) & ~3
        d = [self.dim, self.mu_desc.h1, self.mu_desc.h2, 1]
        pre = "%s/~mu_coeff_nn/mu_coeff/~" % m
        tree, o = {}, 0
        for k in range(3):
            if k < 2:
                outp = up4(d[k + 1])
                w = g[o:o + d[k] * outp].reshape(d[k], outp)[:, :d[k + 1]].copy()
                o += d[k] * outp
            else:
                inp = up4(d[k])
                w = g[o:o + d[k + 1] * inp].reshape(d[k + 1], inp)[:, :d[k]].T.copy()
                o += d[k + 1] * inp
            b = g[o:o + d[k + 1]].copy()
            o += up4(d[k + 1])
            tree["%s/linear_%d" % (pre, k)] = {"w": w, "b": b}
        return tree

    def __call__(self, nn_params, y, u, key):
        """-> (total, {logged scalars}, {mod: {leaf: numpy gradient}}) for device tensors y[B, Hd+1, n_y], u[B, Hd, n_u]
        and a device key int32[2]."""
        st = self.launch(nn_params, y, u, key)
        torch.cuda.current_stream(y.device).synchronize()
        return self.finish(st)

    def launch(self, nn_params, y, u, key, pk=None):
        """Queue the kernels and the read-back on the current stream; ``finish`` parses the result after the caller has
        synchronised the stream."""
        eng = self.engine
        y, u = y.contiguous(), u.contiguous()
        B, Hd = y.shape[0], y.shape[1] - 1
        assert tuple(y.shape) == (B, Hd + 1, eng.n_x) and tuple(u.shape) == (B, Hd, eng.n_u)
        st = self._state(B, Hd, y.device)
        pk = eng.packed(nn_params) if pk is None else pk
        if self.mu_size:
            self._pack_mu(nn_params, st["mu_host"].numpy())
            st["mu"].copy_(st["mu_host"], non_blocking=True)
        a = L.DensityRegArgs()
        a.struct_size = C.sizeof(L.DensityRegArgs)
        a.B, a.N, a.Hd = B, self.N, Hd
        a.params, a.y, a.u, a.key = pk.dev.data_ptr(), y.data_ptr(), u.data_ptr(), key.data_ptr()
        a.rng_mode = _random.rng_mode()
        a.ball_nsamples = int(self.dl["ball_nsamples"])
        rad = np.broadcast_to(np.asarray(self.dl["ball_radius"], np.float32), (self.dim,))
        for i in range(self.dim):
            a.ball_radius[i] = float(rad[i])
        a.mu_type, a.mu_net = self.mu_type, self.mu_desc
        a.mu_params = st["mu"].data_ptr() if self.mu_size else None
        a.mu_coeff = float(self.dl["mu_coeff"])
        a.w_grad, a.w_scvex, a.w_mu = regulariser_weights(self.lp, self.dl)
        kind = self.lp.get("pen_mu_type", "quad_inv")
        if kind not in L.PEN_MU:
            raise ValueError("Unknown pen_mu_type: {}. Choose from quad_inv, lin_inv, exp_inv".format(kind))
        a.pen_mu_type = L.PEN_MU[kind]
        a.pen_mu_temp = float(self.lp.get("pen_mu_temp", 0.0))
        out = st["out"]
        a.out = out.data_ptr()
        a.grad_params = out[5:].data_ptr()
        a.grad_mu = out[5 + eng.param_total:].data_ptr() if self.mu_size else None
        a.workspace, a.workspace_bytes = st["ws"].data_ptr(), st["ws"].numel()
        stream = torch.cuda.current_stream(y.device)
        L.check(L.load().sde4_density_reg(C.byref(eng.desc), C.byref(a), C.c_void_p(stream.cuda_stream)))
        st["out_host"].copy_(out, non_blocking=True)
        st["keep"] = (y, u, key, pk)  # inputs stay alive until the stream has run
        return st

    def packed_grad(self, st):
        """The regulariser's gradient w.r.t. the packed model block (valid after the stream has been synchronised)."""
        return st["out_host"].numpy()[5:5 + self.engine.param_total]

    def finish(self, st, mu_only=False):
        """-> (total, logged scalars, gradient tree); ``mu_only``: the tree holds the mu parameters only (the caller adds
        ``packed_grad`` to its own packed block before unpacking)."""
        eng = self.engine
        st.pop("keep", None)
        res = st["out_host"].numpy()
        info = {k: float(res[1 + i]) for i, k in enumerate(_INFO_KEYS)}
        grads = {}
        if not mu_only:
            gtree = self.model.unpack_grads(res[5:5 + eng.param_total], eng.desc)
            pre = "%s/~construct_diffusion_density_nn/density/~" % self.model.module_name
            grads = {mod: leaves for mod, leaves in gtree.items() if mod.startswith(pre)}
        grads.update(self._unpack_mu(res[5 + eng.param_total:]))
        return float(res[0]), info, grads
