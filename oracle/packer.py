"""Oracle-side descriptor + parameter packer for the multirotor SDE -- TEST INFRASTRUCTURE (and the CPU arm of
bench.py), never imported by the product.

The C restatement (oracle/c) consumes the POD descriptors and the packed parameter block that
``include/sde4b200.h`` documents.  This module builds both straight from the reference's inputs -- the
``params_model`` dict and the haiku parameter tree -- WITHOUT going through ``sde4mbrl_b200.sde_models`` /
``libsde4b200.so``, so that a test feeding the C oracle does not share the product's packer (a packer error is
no longer common mode), and so that ``bench.py --impl reference`` loads nothing of the product.

Layout restated from the header (``sde4_param_layout``): 16 rotor scalars | scaler (padded to 4) | density | net_a |
net_b; a network is ``w0[in][h1_4] b0[h1_4] w1[h1][h2_4] b1[h2_4] w2T[out][h2_4] b2[out_4]``.
Reference semantics: sde4mbrlExamples/rotor_uav/sde_rotor_model.py:146-158 (get_param), :298-351 (network inputs),
motor_models.py:47-60 (motor polynomial), sde4mbrl/nsde.py:353-373 (noise masks, scaler).
"""
import ctypes as C

import numpy as np

MAX_NX, MAX_NU = 16, 8


class NetDesc(C.Structure):  # sde4_net_desc
    _fields_ = [("n_in", C.c_int32), ("h1", C.c_int32), ("h2", C.c_int32), ("n_out", C.c_int32), ("act", C.c_int32)]


class ModelDesc(C.Structure):  # sde4_model_desc
    _fields_ = [("model_id", C.c_int32), ("n_x", C.c_int32), ("n_u", C.c_int32), ("flags", C.c_int32),
                ("density", NetDesc), ("net_a", NetDesc), ("net_b", NetDesc),
                ("noise_in_mask", C.c_uint32), ("noise_out_mask", C.c_uint32),
                ("net_a_in_mask", C.c_uint32), ("net_b_in_mask", C.c_uint32),
                ("n_scaler", C.c_int32), ("aggr", C.c_float),
                ("noise_prior", C.c_float * MAX_NX), ("inv_state_scaling", C.c_float * MAX_NX),
                ("inv_red_scale", C.c_float), ("cfloat", C.c_float * 8)]


class CostDesc(C.Structure):  # sde4_cost_desc
    _fields_ = [("kind", C.c_int32), ("feat", C.c_int32), ("sig_mask", C.c_uint32), ("use_res_sig", C.c_int32),
                ("res_mult", C.c_float), ("res_sig_lo", C.c_float), ("res_sig_mult", C.c_float),
                ("w_x", C.c_float * MAX_NX), ("w_x_sig", C.c_float * MAX_NX),
                ("w_u", C.c_float * MAX_NU), ("w_u_sig", C.c_float * MAX_NU), ("uref", C.c_float * MAX_NU),
                ("constr_mode", C.c_int32), ("n_slack", C.c_int32), ("slack_col", C.c_int32 * MAX_NX),
                ("pen", C.c_float * MAX_NX), ("lb", C.c_float * MAX_NX), ("ub", C.c_float * MAX_NX),
                ("slack_scale", C.c_float * MAX_NX), ("constr_weight", C.c_float),
                ("has_slew", C.c_int32), ("u_slew_coeff", C.c_float * MAX_NU)]


ACT = {"tanh": 1, "swish": 2}
MODULE = "sde_rotor_model"


def _up4(n):
    return (n + 3) & ~3


def _mask(idx):
    m = 0
    for i in idx:
        m |= 1 << int(i)
    return m


def _net(n_in, hidden, n_out, act):
    assert len(hidden) == 2
    return NetDesc(n_in, int(hidden[0]), int(hidden[1]), n_out, ACT[act])


def rotor_desc(pm):
    """``sde4_model_desc`` of an SDERotorModel ``params_model`` dict."""
    d = ModelDesc()
    d.model_id, d.n_x, d.n_u = 3, int(pm["n_x"]), int(pm["n_u"])
    d.flags = 1 if pm.get("aero_drag_effect", False) else 0
    ss = list(pm.get("state_scaling", [1.0] * 13))
    for i in range(13):
        d.noise_prior[i] = float(pm["noise_prior_params"][i])
        d.inv_state_scaling[i] = float(np.float32(1.0) / np.float32(ss[i]))
    d.inv_red_scale = float(np.float32(1.0) / np.float32(max(ss))) if "state_scaling" in pm else 1.0
    d.cfloat[0] = float(pm["init_params"].get("gravity", 9.81))
    d.aggr = 1.0
    dd = pm.get("diffusion_density_nn", None)
    if dd is not None and "density_nn" in dd:
        nin = list(dd.get("indx_noise_in", range(13)))
        nout = list(dd.get("indx_noise_out", range(13)))
        d.noise_in_mask, d.noise_out_mask = _mask(nin), _mask(nout)
        n_in = len(nin) + (d.n_u if pm.get("control_dependent_noise", False) else 0)
        dn = dd["density_nn"]
        d.density = _net(n_in, dn["hidden_layers"], 1, dn["activation_fn"])
        d.aggr = float(dn.get("aggressiveness", 1.0))
        if dd.get("scaler_nn", {}).get("type", None) == "scaler":
            d.n_scaler = 2 * len(nout)
    for key, fld, mfld in (("residual_forces", "net_a", "net_a_in_mask"), ("residual_moments", "net_b", "net_b_in_mask")):
        if key in pm:
            r = pm[key]
            ai = r.get("active_indx", None)
            ai = list(range(6)) if ai is None else list(ai)
            setattr(d, fld, _net(len(ai), r["hidden_layers"], 3, r["activation_fn"]))
            setattr(d, mfld, _mask(ai))
    return d


def _net_floats(nd):
    if nd.n_in == 0:
        return 0
    h1p, h2p = _up4(nd.h1), _up4(nd.h2)
    return nd.n_in * h1p + h1p + nd.h1 * h2p + h2p + nd.n_out * h2p + _up4(nd.n_out)


def rotor_layout(d):
    """Offsets (floats) [scalars, scaler, density, net_a, net_b, total] of the packed block."""
    off = [0, 16]
    off.append(off[1] + _up4(d.n_scaler))
    off.append(off[2] + _net_floats(d.density))
    off.append(off[3] + _net_floats(d.net_a))
    off.append(_up4(off[4] + _net_floats(d.net_b)))
    return off


def _put_net(out, base, tree, prefix, nd):
    o = base
    dims = [nd.n_in, nd.h1, nd.h2, nd.n_out]
    for k in range(3):
        w = np.asarray(tree["%s/linear_%d" % (prefix, k)]["w"], np.float32)
        b = np.asarray(tree["%s/linear_%d" % (prefix, k)]["b"], np.float32)
        assert w.shape == (dims[k], dims[k + 1]) and b.shape == (dims[k + 1],)
        if k < 2:  # row-major [in][out], out padded to a multiple of 4
            ld = _up4(dims[k + 1])
            for i in range(dims[k]):
                out[o + i * ld:o + i * ld + dims[k + 1]] = w[i]
            o += dims[k] * ld
        else:  # last layer transposed: one contiguous row per output
            ld = _up4(dims[k])
            for j in range(dims[k + 1]):
                out[o + j * ld:o + j * ld + dims[k]] = w[:, j]
            o += dims[k + 1] * ld
        out[o:o + dims[k + 1]] = b
        o += _up4(dims[k + 1])
    return o


def rotor_pack(pm, nn, d=None):
    """Packed float32 parameter block of an SDERotorModel haiku tree ``nn``."""
    d = rotor_desc(pm) if d is None else d
    off = rotor_layout(d)
    out = np.zeros(off[5], np.float32)
    top = nn.get(MODULE, {})
    ip = pm["init_params"]
    fixed = bool(pm.get("fixed_init_params", False))

    def get_param(name):  # sde_rotor_model.py:146-158
        if name in ip and fixed:
            return float(ip[name])
        return float(top[name])
    fm = ip["fm_model"]
    order, const = int(fm.get("order", 1)), bool(fm.get("use_constant_term", False))
    sym = bool(fm.get("assume_sym", True))
    sc = [get_param(n) for n in ("mass", "Ixx", "Iyy", "Izz", "Lmx")]
    sc.append(sc[4] if sym else get_param("Lmy"))
    sc.append(get_param("CoeffDrag"))
    out[0:7] = sc
    # thrust = sum_i coeffMot_i * u^(order - i) (+ coeffMot_order): stored as c3, c2, c1, c0 (motor_models.py:47-60)
    for i in range(order):
        out[7 + 3 - (order - i)] = get_param("coeffMot%d" % i)
    if const:
        out[10] = get_param("coeffMot%d" % order)
    if pm.get("aero_drag_effect", False):
        for k, n in enumerate(("aero_kdx", "aero_kdy", "aero_kdz", "aero_kh")):
            out[11 + k] = float(top[n])
    if d.n_scaler:
        out[off[1]:off[1] + d.n_scaler] = np.asarray(top["scaler"], np.float32)
    for slot, prefix, nd in ((2, "%s/~construct_diffusion_density_nn/density/~" % MODULE, d.density),
                             (3, "%s/~init_residual_networks/res_forces/~" % MODULE, d.net_a),
                             (4, "%s/~init_residual_networks/res_moments/~" % MODULE, d.net_b)):
        if nd.n_in:
            _put_net(out, off[slot], nn, prefix, nd)
    return out


def rotor_track_cost_desc(cp, n_u, with_slew=True):
    """``sde4_cost_desc`` of ``one_step_cost_f`` (sde_mpc_design.py:73-132) + the control-slew term (:141-148)."""
    c = CostDesc()
    for i in range(MAX_NX):
        c.slack_col[i] = -1
    c.kind, c.res_mult, c.constr_weight = 1, float(cp.get("res_mult", 1.0)), 1.0
    for key, bit, idx in (("perr", 0, (0, 1, 2)), ("verr", 1, (3, 4, 5)), ("qerr", 2, (7, 8, 9)), ("werr", 3, (10, 11, 12))):
        if key in cp:
            w = np.broadcast_to(np.asarray(cp[key], np.float64), (3,))
            for k, i in enumerate(idx):
                c.w_x[i] = float(w[k])
            if key + "_sig" in cp:
                c.sig_mask |= 1 << bit
                ws = np.broadcast_to(np.asarray(cp[key + "_sig"], np.float64), (3,))
                for k, i in enumerate(idx):
                    c.w_x_sig[i] = float(ws[k])
    if "uerr" in cp:
        w = np.broadcast_to(np.asarray(cp["uerr"], np.float64), (n_u,))
        ur = np.broadcast_to(np.asarray(cp["uref"], np.float64), (n_u,))
        for j in range(n_u):
            c.w_u[j], c.uref[j] = float(w[j]), float(ur[j])
        if "uerr_sig" in cp:
            c.sig_mask |= 1 << 4
            ws = np.broadcast_to(np.asarray(cp["uerr_sig"], np.float64), (n_u,))
            for j in range(n_u):
                c.w_u_sig[j] = float(ws[j])
    if "res_sig" in cp:
        c.use_res_sig, c.res_sig_lo, c.res_sig_mult = 1, float(cp["res_sig"][0]), float(cp["res_sig"][1])
    if with_slew and "u_slew_coeff" in cp:
        c.has_slew = 1
        for j, w in enumerate(np.broadcast_to(np.asarray(cp["u_slew_coeff"], np.float64), (n_u,))):
            c.u_slew_coeff[j] = float(w)
    return c
