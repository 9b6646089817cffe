"""Host-side mirrors of the reference's ``ControlledSDE`` subclasses.

In the reference these are haiku modules whose methods are traced by JAX
(``sde4mbrl/nsde.py:134-864`` and the three model files under
``sde4mbrlExamples``).  Here a class only *describes* its model to the CUDA
library: it turns the reference's ``params_model`` dict into the POD
``sde4_model_desc`` (static architecture + numeric constants) and packs a
haiku-style parameter tree ``{module_path: {name: array}}`` -- the exact layout of
the shipped ``*_sde.pkl`` files -- into the flat float32 block the kernels read.
Anything the compiled kernels cannot express raises ``Sde4Error`` (no CPU
fallback).
"""
import numpy as np

from . import _lib as L
from ._lib import ModelDesc, NetDesc, Sde4Error

_ACT = {"tanh": L.ACT_TANH, "swish": L.ACT_SWISH, "silu": L.ACT_SWISH}


def _mask(indices):
    m = 0
    prev = -1
    for i in indices:
        i = int(i)
        if i <= prev:
            raise Sde4Error("index lists must be strictly increasing for the compiled kernels: %s" % (list(indices),))
        m |= 1 << i
        prev = i
    return m


def _net_desc(n_in, hidden, n_out, act_name):
    hidden = list(hidden)
    if len(hidden) != 2:
        raise Sde4Error("compiled kernels support MLPs with exactly two hidden layers, got %s" % (hidden,))
    if act_name not in _ACT:
        raise Sde4Error("unsupported activation '%s' (tanh, swish)" % act_name)
    return NetDesc(int(n_in), int(hidden[0]), int(hidden[1]), int(n_out), _ACT[act_name])


def _up4(n):
    return (n + 3) & ~3


class ControlledSDE:
    """Descriptor base class (reference: ``sde4mbrl/nsde.py:134``)."""
    module_name = "controlled_sde"
    model_id = 0
    n_reduced = None  # size of reduced_state(x); default n_x

    def __init__(self, params, name=None):
        self.params = params
        for k in ("n_x", "n_u", "n_y", "horizon", "stepsize"):
            if k not in params:
                raise AssertionError("'%s' must be specified in the params" % k)
        if params.get("sde_solver", "euler_maruyama") not in L.SOLVERS:
            # sde_solver.py:321-325: euler_maruyama is the reference's only registered solver; the two Stratonovich
            # schemes of its commented-out code are offered as an extension of the ROLLOUT kernel (sde4b200.h)
            raise Sde4Error("unknown sde_solver '%s' (known: %s)" % (params["sde_solver"], ", ".join(L.SOLVERS)))
        if params["n_x"] != params["n_y"]:
            raise Sde4Error("only identity observation maps are supported (n_x == n_y)")
        self.n_x, self.n_u, self.n_y = params["n_x"], params["n_u"], params["n_y"]

    # ---- to be specialised ----------------------------------------------
    def _fill(self, d):
        raise NotImplementedError

    def nets(self):
        """[(slot, haiku prefix, n_in, hidden, n_out, act, init_value)] for density, net_a, net_b."""
        raise NotImplementedError

    def scalar_names(self):
        return []

    # ---- shared ---------------------------------------------------------
    def _density_setup(self, d, n_reduced):
        p = self.params
        dd = p.get("diffusion_density_nn", None)
        d.aggr = 1.0
        if dd is None or "density_nn" not in dd:
            return None
        nin = list(dd.get("indx_noise_in", range(self.n_x)))
        nout = list(dd.get("indx_noise_out", range(self.n_x)))
        if max(nin) >= n_reduced:
            raise Sde4Error("indx_noise_in exceeds the reduced state size")
        d.noise_in_mask = _mask(nin)
        d.noise_out_mask = _mask(nout)
        dn = dd["density_nn"]
        # control_dependent_noise (nsde.py:362): the density net reads concat(noise inputs, u); the kernels recognise it
        # by a density input wider than the noise-input mask (csrc/models.cuh, Diffusion::CDN)
        n_in = len(nin) + (self.n_u if p.get("control_dependent_noise", False) else 0)
        d.density = _net_desc(n_in, dn["hidden_layers"], 1, dn.get("activation_fn", "tanh"))
        d.aggr = float(dn.get("aggressiveness", 1.0))
        sc = dd.get("scaler_nn", None)
        if sc is not None and sc.get("type", "none") != "none":
            if sc.get("type", "scaler") != "scaler":
                raise Sde4Error("scaler_nn type 'nn' has no compiled kernel (only 'scaler')")
            d.n_scaler = 2 * len(nout)
        return dd

    def describe(self):
        d = ModelDesc()
        d.model_id = self.model_id
        d.n_x, d.n_u = self.n_x, self.n_u
        prior = list(self.params["noise_prior_params"])
        if len(prior) != self.n_x:
            raise Sde4Error("noise_prior_params must have n_x entries")
        ss = list(self.params.get("state_scaling", [1.0] * self.n_x))
        for i in range(self.n_x):
            d.noise_prior[i] = float(prior[i])
            d.inv_state_scaling[i] = float(np.float32(1.0) / np.float32(ss[i]))
        d.inv_red_scale = 1.0
        self._fill(d)
        return d

    @staticmethod
    def layout(desc):
        off = (L.C.c_int32 * 6)()
        total = L.load().sde4_param_layout(L.C.byref(desc), off)
        if total < 0:
            raise Sde4Error(L.last_error())
        return list(off), total

    def is_noisy(self):
        return any(float(v) != 0.0 for v in self.params["noise_prior_params"])

    # ---- parameter trees ------------------------------------------------
    def _pack_net(self, out, base, tree, prefix, nd):
        """w0[in][h1p] b0[h1p] w1[h1][h2p] b1[h2p] w2T[out][h2p] b2[outp]  (csrc/nets.cuh)."""
        o = base
        dims = [nd.n_in, nd.h1, nd.h2, nd.n_out]
        for k in range(3):
            key = "%s/linear_%d" % (prefix, k)
            if key not in tree:
                raise Sde4Error("missing parameters '%s'" % key)
            w = np.asarray(tree[key]["w"], np.float32)
            b = np.asarray(tree[key]["b"], np.float32)
            if w.shape != (dims[k], dims[k + 1]) or b.shape != (dims[k + 1],):
                raise Sde4Error("parameter shape mismatch at '%s': %s" % (key, w.shape))
            if k < 2:
                outp = _up4(dims[k + 1])
                wp = np.zeros((dims[k], outp), np.float32)
                wp[:, :dims[k + 1]] = w
            else:  # last layer stored transposed: [out][up4(in)]
                wp = np.zeros((dims[k + 1], _up4(dims[k])), np.float32)
                wp[:, :dims[k]] = w.T
            out[o:o + wp.size] = wp.ravel()
            o += wp.size
            out[o:o + dims[k + 1]] = b
            o += _up4(dims[k + 1])
        return o

    def pack(self, nn_params, desc=None):
        """haiku-style tree -> flat float32 block (layout: include/sde4b200.h, sde4_param_layout)."""
        desc = self.describe() if desc is None else desc
        off, total = self.layout(desc)
        out = np.zeros(total, np.float32)
        top = nn_params.get(self.module_name, {})
        self._pack_scalars(out, off[0], top)
        if desc.n_scaler:
            sc = np.asarray(top["scaler"], np.float32)
            if sc.shape != (desc.n_scaler,):
                raise Sde4Error("scaler must have shape (%d,)" % desc.n_scaler)
            out[off[1]:off[1] + desc.n_scaler] = sc
        for slot, prefix, nd in self._net_slots(desc):
            if nd.n_in > 0:
                self._pack_net(out, off[slot], nn_params, prefix, nd)
        return out

    def _pack_scalars(self, out, base, top):
        pass

    def unpack_grads(self, gpacked, desc=None):
        """Inverse of ``pack`` for a gradient block: flat float32 -> haiku-style tree of numpy arrays
        (network weights / biases and the scaler; model scalars are not differentiated)."""
        desc = self.describe() if desc is None else desc
        off, total = self.layout(desc)
        g = np.asarray(gpacked, np.float32)
        tree = {}
        if desc.n_scaler:
            tree.setdefault(self.module_name, {})["scaler"] = g[off[1]:off[1] + desc.n_scaler].copy()
        for slot, prefix, nd in self._net_slots(desc):
            if nd.n_in == 0:
                continue
            o = off[slot]
            dims = [nd.n_in, nd.h1, nd.h2, nd.n_out]
            for k in range(3):
                if k < 2:
                    outp = _up4(dims[k + 1])
                    w = g[o:o + dims[k] * outp].reshape(dims[k], outp)[:, :dims[k + 1]].copy()
                    o += dims[k] * outp
                else:
                    inp = _up4(dims[k])
                    w = g[o:o + dims[k + 1] * inp].reshape(dims[k + 1], inp)[:, :dims[k]].T.copy()
                    o += dims[k + 1] * inp
                b = g[o:o + dims[k + 1]].copy()
                o += _up4(dims[k + 1])
                tree["%s/linear_%d" % (prefix, k)] = {"w": w, "b": b}
        return tree

    def _net_slots(self, desc):
        m = self.module_name
        return [(2, "%s/~construct_diffusion_density_nn/density/~" % m, desc.density),
                (3, "%s/~init_residual_networks/%s/~" % (m, self.net_a_name), desc.net_a),
                (4, "%s/~init_residual_networks/%s/~" % (m, self.net_b_name), desc.net_b)]

    net_a_name = "res_forces"
    net_b_name = "res_moments"

    def init_params(self, seed=0):
        """Random initial parameter tree, distributed like the reference's haiku initialisers
        (``w ~ U(-init_value, init_value)``, ``b = 0``; nsde.py:372-390 and the model files).
        Haiku's own key-splitting order is not reproduced (not part of the hot path)."""
        rng = np.random.default_rng(seed)
        desc = self.describe()
        tree = {}
        top = self._init_scalars(rng)
        if desc.n_scaler:
            iv = self.params["diffusion_density_nn"]["scaler_nn"].get("init_value", 0.001)
            top["scaler"] = rng.uniform(-iv, iv, desc.n_scaler).astype(np.float32)
        if top:
            tree[self.module_name] = top
        inits = self._init_values()
        for (slot, prefix, nd), iv in zip(self._net_slots(desc), inits):
            if nd.n_in == 0:
                continue
            dims = [nd.n_in, nd.h1, nd.h2, nd.n_out]
            for k in range(3):
                tree["%s/linear_%d" % (prefix, k)] = {
                    "w": rng.uniform(-iv, iv, (dims[k], dims[k + 1])).astype(np.float32),
                    "b": np.zeros(dims[k + 1], np.float32)}
        return tree

    def _init_scalars(self, rng):
        return {}

    def _init_values(self):
        p = self.params
        dv = p.get("diffusion_density_nn", {}).get("density_nn", {}).get("init_value", 0.01)
        return [dv, p.get("residual_forces", {}).get("init_value", 1e-3), p.get("residual_moments", {}).get("init_value", 1e-3)]


class MassSpringDamper(ControlledSDE):
    """``sde4mbrlExamples/mass_spring_damper/mass_spring_model.py:21-123``."""
    module_name = "mass_spring_damper"
    model_id = L.MODEL_MSD
    net_a_name = "Fres"

    def _fill(self, d):
        p = self.params
        ip = p["init_params"]
        if p.get("encoder_type", "identity") not in ("identity", "ground_truth"):
            raise ValueError("Unknown encoder type")
        d.cfloat[0], d.cfloat[1], d.cfloat[2] = float(ip["mass"]), float(ip["kq_coeff"]), float(ip["bqdot_coeff"])
        flags = 0
        gt = bool(p.get("ground_truth", False))
        si = bool(p.get("include_side_info", False))
        if gt:
            flags |= L.MSD_GROUND_TRUTH
        elif si:
            flags |= L.MSD_SIDE_INFO
        if p["control"] and (gt or si):
            flags |= L.MSD_CONTROL  # the black-box residual ignores u (mass_spring_model.py:84-86)
        d.flags = flags
        self._density_setup(d, 2)
        if not gt:
            if p.get("prior", False):
                raise Sde4Error("'prior' (zero residual) MSD variant has no compiled kernel")
            rf = p["residual_forces"]
            if rf["type"] != "dnn":
                raise NotImplementedError
            d.net_a = _net_desc(1 if si else 2, rf["hidden_layers"], 1, rf["activation_fn"])


class CartPoleSDE(ControlledSDE):
    """``sde4mbrlExamples/cartpole/cartpole_sde.py:20-106``."""
    module_name = "cart_pole_sde"
    model_id = L.MODEL_CARTPOLE
    net_b_name = "control_nn"

    def _fill(self, d):
        p = self.params
        si = bool(p.get("side_info", False))
        d.flags = L.CART_SIDE_INFO if si else 0
        sc = p["nn_out_scale"]
        d.cfloat[0], d.cfloat[1] = float(sc[0]), float(sc[1])
        ss = p.get("state_scaling", [1.0] * self.n_x)
        # reduced_state is only redefined when state_scaling is given (cartpole_sde.py:39-41)
        if "diffusion_density_nn" in p and "density_nn" in p["diffusion_density_nn"] and "state_scaling" not in p:
            raise Sde4Error("cartpole density net without state_scaling has no compiled kernel")
        d.inv_red_scale = float(np.float32(1.0) / np.float32(ss[-1]))
        self._density_setup(d, 4)
        rf = p["residual_forces"]
        d.net_a = _net_desc(3 if si else 4, rf["hidden_layers"], 2, rf["activation_fn"])
        if si:
            cn = p["control_nn"]
            d.net_b = _net_desc(2, cn["hidden_layers"], 2, cn["activation_fn"])

    def _init_values(self):
        p = self.params
        dv = p.get("diffusion_density_nn", {}).get("density_nn", {}).get("init_value", 0.01)
        return [dv, p["residual_forces"].get("init_value", 1e-3), p.get("control_nn", {}).get("init_value", 1e-3)]


_ROTOR_SCALARS = ["mass", "Ixx", "Iyy", "Izz", "Lmx", "Lmy", "CoeffDrag"]


class SDERotorModel(ControlledSDE):
    """``sde4mbrlExamples/rotor_uav/sde_rotor_model.py:121-458``."""
    module_name = "sde_rotor_model"
    model_id = L.MODEL_ROTOR

    def _fill(self, d):
        p = self.params
        if p.get("control_augmented_state", False):
            raise Sde4Error("control_augmented_state has no compiled kernel (unused by the shipped configs)")
        if self.n_x != 13 or self.n_u not in (4, 6):
            raise Sde4Error("rotor model needs n_x=13 and n_u in (4, 6)")
        ip = p["init_params"]
        assert "fm_model" in ip, "The motor models must be specified"
        if ip["fm_model"].get("learn_mixer", False):
            raise Sde4Error("fm_model.learn_mixer has no compiled kernel (unused by the shipped configs)")
        d.cfloat[0] = float(ip.get("gravity", 9.81))
        d.flags = L.ROTOR_AERO_DRAG if p.get("aero_drag_effect", False) else 0
        if "state_scaling" in p:
            d.inv_red_scale = float(np.float32(1.0) / np.float32(max(p["state_scaling"])))
        self._density_setup(d, 13)
        if "residual_forces" in p:
            rf = p["residual_forces"]
            ai = rf.get("active_indx", None)
            ai = list(range(6)) if ai is None else list(ai)
            d.net_a = _net_desc(len(ai), rf["hidden_layers"], 3, rf["activation_fn"])
            d.net_a_in_mask = _mask(ai)
        if "residual_moments" in p:
            rm = p["residual_moments"]
            ai = rm.get("active_indx", None)
            ai = list(range(6)) if ai is None else list(ai)
            d.net_b = _net_desc(len(ai), rm["hidden_layers"], 3, rm["activation_fn"])
            d.net_b_in_mask = _mask(ai)

    def get_param(self, top, name):
        """``get_param`` (sde_rotor_model.py:146-158): YAML constant when ``fixed_init_params``."""
        ip = self.params["init_params"]
        if name in ip and self.params.get("fixed_init_params", False):
            return float(ip[name])
        if name not in top:
            raise Sde4Error("missing learnable scalar '%s'" % name)
        return float(top[name])

    def _pack_scalars(self, out, base, top):
        fm = self.params["init_params"]["fm_model"]
        order = int(fm.get("order", 1))
        const = bool(fm.get("use_constant_term", False))
        assume_sym = bool(fm.get("assume_sym", True))
        if order not in (1, 2, 3):
            raise ValueError("Invalid order")
        vals = {n: self.get_param(top, n) for n in ("mass", "Ixx", "Iyy", "Izz", "Lmx", "CoeffDrag")}
        vals["Lmy"] = vals["Lmx"] if assume_sym else self.get_param(top, "Lmy")
        for k, n in enumerate(_ROTOR_SCALARS):
            out[base + k] = vals[n]
        # motor polynomial: monomials [u^order, ..., u, (1)] x coeffMot{i} (motor_models.py:47-57)
        # stored highest power first as c3,c2,c1,c0
        n_mono = order + (1 if const else 0)
        coeffs = [self.get_param(top, "coeffMot%d" % i) for i in range(n_mono)]
        poly = [0.0, 0.0, 0.0, 0.0]
        for i in range(order):
            power = order - i
            poly[3 - power] = coeffs[i]
        if const:
            poly[3] = coeffs[order]
        out[base + 7:base + 11] = poly
        if self.params.get("aero_drag_effect", False):
            for k, n in enumerate(("aero_kdx", "aero_kdy", "aero_kdz", "aero_kh")):
                out[base + 11 + k] = float(top[n])

    def unpack_grads(self, gpacked, desc=None):
        """Network / scaler gradients as in the base class, plus the gradients of the LEARNABLE physical scalars
        (chain rule through ``_pack_scalars``: ``Lmy = Lmx`` when ``assume_sym``, motor polynomial coefficients)."""
        desc = self.describe() if desc is None else desc
        tree = super().unpack_grads(gpacked, desc)
        off, _ = self.layout(desc)
        g = np.asarray(gpacked, np.float32)[off[0]:off[0] + 16]
        ip = self.params["init_params"]
        fixed = self.params.get("fixed_init_params", False)
        fm = ip["fm_model"]
        order, const = int(fm.get("order", 1)), bool(fm.get("use_constant_term", False))
        assume_sym = bool(fm.get("assume_sym", True))
        out = {}
        idx = {n: k for k, n in enumerate(_ROTOR_SCALARS)}

        def learnable(n):
            return not (n in ip and fixed)
        for n in ("mass", "Ixx", "Iyy", "Izz", "Lmx", "CoeffDrag"):
            if learnable(n):
                out[n] = np.float32(g[idx[n]])
        if assume_sym:
            if "Lmx" in out:
                out["Lmx"] = np.float32(out["Lmx"] + g[idx["Lmy"]])
        elif learnable("Lmy"):
            out["Lmy"] = np.float32(g[idx["Lmy"]])
        for i in range(order + (1 if const else 0)):
            n = "coeffMot%d" % i
            if learnable(n):
                out[n] = np.float32(g[7 + (3 - (order - i) if i < order else 3)])
        if self.params.get("aero_drag_effect", False):
            for k, n in enumerate(("aero_kdx", "aero_kdy", "aero_kdz", "aero_kh")):
                out[n] = np.float32(g[11 + k])
        if out:
            tree.setdefault(self.module_name, {}).update(out)
        return tree

    def _init_scalars(self, rng):
        ip = self.params["init_params"]
        top = {}
        fm = ip["fm_model"]
        n_mono = int(fm.get("order", 1)) + (1 if fm.get("use_constant_term", False) else 0)
        names = ["mass", "Ixx", "Iyy", "Izz", "Lmx", "CoeffDrag"] + ["coeffMot%d" % i for i in range(n_mono)]
        if not fm.get("assume_sym", True):
            names.append("Lmy")
        fixed = self.params.get("fixed_init_params", False)
        for n in names:
            if n in ip and fixed:
                continue
            top[n] = np.float32(ip.get(n, 1e-3))
        if self.params.get("aero_drag_effect", False):
            for n, hi in (("aero_kdx", 1e-3), ("aero_kdy", 1e-3), ("aero_kdz", 1e-3), ("aero_kh", 1e-4)):
                top[n] = np.float32(rng.uniform(0.0, hi))
        return top

    def _init_values(self):
        dv = self.params.get("diffusion_density_nn", {}).get("density_nn", {}).get("init_value", 0.01)
        return [dv, 1e-3, 1e-3]  # sde_rotor_model.py:402,411
