"""ctypes wrapper of the C restatement (oracle/c) -- test infrastructure / CPU baseline only.

The C library consumes the same POD descriptors and packed parameter block as the CUDA library
(include/sde4b200.h), so it also cross-checks the product's packer."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "c", "libsde4oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError("oracle/c/libsde4oracle.so missing: run `make -C oracle/c`")
        _lib = C.CDLL(LIB)
        _lib.sde4o_rotor_cost_grad.restype = C.c_int
        _lib.sde4o_simd_width.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else C.c_void_p(0)


def rotor_cost_grad(desc, cost_desc, packed, y0, opt, dt, xref, N, key=None, rng_mode=0, noise=None,
                    noisy=True, want_grad=True, nthreads=0, want_xtraj=False):
    """Mean-over-particles MPC cost (and gradient wrt opt[H, n_u]) of ONE multirotor problem.
    noise: None (in-line threefry from `key`) or float32 [N, H, 13]."""
    lib = load()
    opt = np.ascontiguousarray(opt, np.float32)
    H, nu = opt.shape
    packed = np.ascontiguousarray(packed, np.float32)
    y0 = np.ascontiguousarray(y0, np.float32)
    dt = np.ascontiguousarray(dt, np.float32)
    xref = np.ascontiguousarray(xref, np.float32)
    keya = np.ascontiguousarray(key if key is not None else [0, 0], np.uint32)
    noise_a = np.ascontiguousarray(noise, np.float32) if noise is not None else None
    cost = np.zeros(1, np.float32)
    grad = np.zeros((H, nu), np.float32)
    xtraj = np.zeros((N, H + 1, 13), np.float32) if want_xtraj else None
    rc = lib.sde4o_rotor_cost_grad(C.byref(desc), C.byref(cost_desc), _p(packed), _p(y0), _p(opt), _p(dt), _p(xref),
                                   C.c_int(N), C.c_int(H), _p(keya), C.c_int(rng_mode), _p(noise_a), C.c_int(int(noisy)),
                                   C.c_int(int(want_grad)), C.c_int(nthreads), _p(cost), _p(grad), _p(xtraj))
    assert rc == 0
    return float(cost[0]), (grad if want_grad else None), xtraj
