"""Generates the committed golden fixtures under tests/golden/.

Run in the build container (where /root/reference is mounted):
    python tests/golden/make_golden.py

1. ref_weights_*.npz  -- the reference's SHIPPED learned parameters (data files, not source):
   sde4mbrlExamples/rotor_uav/hexa_ahg/my_models/hexa_ahg_v4_sde.pkl and
   sde4mbrlExamples/cartpole/my_models/cartpole_bb_learned_si_sde.pkl, flattened to
   {haiku_path::name: array} plus the 'nominal' model dict as JSON, so real-weight parity tests
   run on the GPU box (which has no /root/reference).
2. golden_<config>.npz -- oracle outputs (float32 trajectories with jax-layout threefry noise,
   float64 cost / gradient) for reduced-size instances of the BASELINE configs, produced by the
   CPU oracle (oracle/nsde_oracle.py).  The reference itself cannot be run here (no jax), so these
   pin the ORACLE (regression) and the CUDA path against it; see oracle/__init__.py.
"""
import json
import os
import pickle
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers as hp  # noqa: E402
from oracle import nsde_oracle as orc  # noqa: E402
from oracle import threefry as tf  # noqa: E402
from sde4mbrl_b200 import configs  # noqa: E402

REF = "/root/reference/sde4mbrlExamples"


def _jsonable(o):
    if isinstance(o, dict):
        return {k: _jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_jsonable(v) for v in o]
    if isinstance(o, (np.floating, np.integer)):
        return o.item()
    return o


def export_weights(pkl, out):
    d = pickle.load(open(os.path.join(REF, pkl), "rb"))
    flat = {}
    for mod, leaves in d["sde"].items():
        for name, arr in leaves.items():
            flat["%s::%s" % (mod, name)] = np.asarray(arr)
    flat["__nominal_json__"] = np.frombuffer(json.dumps(_jsonable(d["nominal"])).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, out), **flat)
    print("wrote", out, len(flat) - 1, "tensors")


def golden(kind, H, N, seed, mode):
    rng = np.random.default_rng(seed)
    pm, model, nn = hp.build(kind, H, N, seed=seed)
    y0 = hp.start_state(kind, rng)
    u = hp.controls(kind, H, rng)
    key = tf.PRNGKey(2050)
    om = hp.oracle_model(kind, pm, nn)
    x, dw = orc.sample_rollouts(om, y0, u, key, N, mode=mode)
    out = dict(y0=y0, u=u, key=key, xevol=x.numpy(), rng_mode=np.int32(mode), H=np.int32(H), N=np.int32(N),
               seed=np.int32(seed), dw=dw.numpy())
    # cost + gradient (float64) on the same noise
    if kind.startswith("hexa"):
        cp = configs.hexa_mpc()["cost_params"]
        xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
        xref[:, :3] = [0.0, 0.5, 1.4]
    elif kind.startswith("cartpole"):
        cp = configs.cartpole_mpc()["cost_params"]
        xref = np.zeros((H + 1, 4), np.float32)
    else:
        cp = {"feat": "identity", "xerr": [2.0, 0.5], "uerr": [0.1], "uref": [0.0], "res_mult": 1.0}
        xref = np.tile(np.array([0.05, 0.0], np.float32), (H + 1, 1))
    om64 = hp.oracle_model(kind, pm, nn, torch.float64)
    o = torch.tensor(u, dtype=torch.float64, requires_grad=True)
    c, _ = orc.compute_cost(om64, torch.tensor(y0, dtype=torch.float64), o, dw.double(), hp.stage_cost_fn(kind, cp),
                            torch.tensor(xref, dtype=torch.float64))
    if kind.startswith("hexa"):
        c = c + orc.u_slew_cost(o, orc.compute_timesteps(pm), cp, 6)
    (g,) = torch.autograd.grad(c, o)
    out.update(cost=np.float64(c.item()), grad=g.numpy(), xref=xref)
    name = "golden_%s_H%d_N%d_m%d.npz" % (kind, H, N, mode)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote", name, "cost", float(c))


if __name__ == "__main__":
    if os.path.isdir(REF):
        export_weights("rotor_uav/hexa_ahg/my_models/hexa_ahg_v4_sde.pkl", "ref_weights_hexa_ahg_v4.npz")
        export_weights("cartpole/my_models/cartpole_bb_learned_si_sde.pkl", "ref_weights_cartpole_bb_learned_si.npz")
        export_weights("rotor_uav/iris_sitl/my_models/iris_sitl_sde.pkl", "ref_weights_iris_sitl.npz")
    else:
        print("no /root/reference: skipping the weight export")
    golden("msd", 100, 24, 1, 0)      # C1 reduced (N 200 -> 24)
    golden("cartpole", 50, 16, 2, 0)  # C2 reduced (N 1024 -> 16)
    golden("hexa", 100, 8, 3, 0)      # C3 reduced (N 4096 -> 8)
    golden("hexa", 20, 8, 4, 1)       # C4 shape (H=20), partitionable RNG layout
