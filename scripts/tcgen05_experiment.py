"""The hidden->hidden product on tcgen05 (3xTF32, accumulator in tensor memory) against the FP32 pipe: time per layer of a
chain of 32 x 32 swish layers over 128-sample tiles, and the error of both against float64 (profiles/r2_tcgen05_experiment.md).
Usage: python scripts/tcgen05_experiment.py [tiles_per_sm ...]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from sde4mbrl_b200 import _lib

lib = _lib.load()
fn = lib.sde4_debug_layer_chain
fn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_void_p]
fn.restype = C.c_int
dev = torch.device("cuda:0")
sms = torch.cuda.get_device_properties(0).multi_processor_count
rng = np.random.default_rng(0)
w = rng.uniform(-1, 1, (32, 32)).astype(np.float32) / np.sqrt(32.0).astype(np.float32) * 1.7
b = rng.uniform(-0.2, 0.2, 32).astype(np.float32)
wd, bd = torch.as_tensor(w, device=dev), torch.as_tensor(b, device=dev)


def run(variant, tiles, layers, reps):
    x = torch.as_tensor(rng.normal(0, 1, (tiles * 128, 32)).astype(np.float32), device=dev)
    y = torch.empty_like(x)
    ms = C.c_float(0)
    rc = fn(variant, tiles, layers, reps, wd.data_ptr(), bd.data_ptr(), x.data_ptr(), y.data_ptr(), C.byref(ms), None)
    if rc != 0:
        raise RuntimeError("sde4_debug_layer_chain(variant %d) -> %d" % (variant, rc))
    return x.cpu().numpy(), y.cpu().numpy(), ms.value


def ref64(x, layers):
    v = x.astype(np.float64)
    W, B = w.astype(np.float64), b.astype(np.float64)
    for _ in range(layers):
        a = v @ W + B
        v = a / (1.0 + np.exp(-a))
    return v


# accuracy: one and eight layers deep, a few tiles
for layers in (1, 8):
    for variant, name in ((0, "fp32"), (1, "tcgen05 3xTF32")):
        x, y, _ = run(variant, 4, layers, 1)
        want = ref64(x, layers)
        err = np.max(np.abs(y - want)) / np.max(np.abs(want))
        print("accuracy  layers=%d  %-15s max abs err / max |ref| = %.3e" % (layers, name, err))

# throughput: the whole GPU, `tps` tiles per SM, 64 layers per launch
for tps in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16]:
    tiles, layers = sms * tps, 64
    out = {}
    for variant, name in ((0, "fp32"), (1, "tcgen05")):
        _, _, ms = run(variant, tiles, layers, 20)
        out[name] = ms
    flop = 2.0 * 32 * 32 * 128 * tiles * layers
    print("tiles/SM=%2d  fp32 %.4f ms (%.1f TFLOP/s, %.1f ns per layer-tile-wave)  tcgen05 %.4f ms (%.1f TFLOP/s algorithmic)  ratio fp32/tc %.2f"
          % (tps, out["fp32"], flop / out["fp32"] / 1e9, out["fp32"] * 1e6 / layers, out["tcgen05"], flop / out["tcgen05"] / 1e9,
             out["fp32"] / out["tcgen05"]))
