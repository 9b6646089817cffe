"""The XLA-FFI shim (ffi/sde4_xla_ffi.cc, ffi/nsde_b200_jax.py) as far as it can be exercised without jaxlib:
both sides of the `__has_include` guard compile -- the real handlers against tests/mock_xla (a mock of the header's
binding API that static-asserts every handler is callable with the parameter list its binding declares) -- the built
library exports the guard's probe, and the tree <-> packed-block gather the jax half relies on equals model.pack."""
import ctypes
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

import helpers as hp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "ffi", "sde4_xla_ffi.cc")
CUDA_INC = "/usr/local/cuda/include"

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


def test_stub_variant_builds_and_reports_unavailable(tmp_path):
    so = str(tmp_path / "stub.so")
    subprocess.run(["g++", "-std=c++17", "-fPIC", "-shared", "-o", so, SRC], check=True)
    lib = ctypes.CDLL(so)
    assert lib.sde4_xla_ffi_available() == 0
    assert not hasattr(lib, "Sde4CostGrad")


@pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")), reason="needs the CUDA headers")
def test_handlers_compile_against_the_mock_header(tmp_path):
    obj = str(tmp_path / "shim.o")
    subprocess.run(["g++", "-std=c++17", "-fPIC", "-c", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "mock_xla"),
                    "-I", CUDA_INC, "-o", obj, SRC], check=True)
    syms = subprocess.run(["nm", obj], check=True, capture_output=True, text=True).stdout
    for name in ("Sde4Rollout", "Sde4CostGrad", "Sde4LossGrad", "sde4_xla_ffi_available"):
        assert " T " + name in syms, name
    for name in ("sde4_rollout", "sde4_cost_grad", "sde4_loss_grad", "sde4_last_error"):  # resolved by libsde4b200.so
        assert " U " + name in syms, name


def test_a_handler_with_the_wrong_signature_is_rejected(tmp_path):
    """The mock is not a rubber stamp: swapping two parameters of a handler must fail to compile."""
    bad = tmp_path / "bad.cc"
    src = open(SRC).read().replace("ffi::Error RolloutImpl(cudaStream_t stream, ffi::Span<const uint8_t> model_desc, int32_t num_particles,",
                                   "ffi::Error RolloutImpl(cudaStream_t stream, int32_t num_particles, ffi::Span<const uint8_t> model_desc,")
    assert src != open(SRC).read()
    src = src.replace('#include "../include/sde4b200.h"', '#include "%s"' % os.path.join(ROOT, "include", "sde4b200.h"))
    bad.write_text(src)
    r = subprocess.run(["g++", "-std=c++17", "-fPIC", "-c", "-I", os.path.join(ROOT, "tests", "mock_xla"), "-I", CUDA_INC,
                        "-o", str(tmp_path / "bad.o"), str(bad)], capture_output=True, text=True)
    assert r.returncode != 0 and "not callable with the parameter list" in r.stderr


def test_built_shim_exports_the_probe():
    so = os.path.join(ROOT, "sde4mbrl_b200", "libsde4_xla_ffi.so")
    if not os.path.exists(so):
        pytest.skip("libsde4_xla_ffi.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    ctypes.CDLL(os.path.join(ROOT, "sde4mbrl_b200", "libsde4b200.so"), mode=ctypes.RTLD_GLOBAL)
    lib = ctypes.CDLL(so)
    assert lib.sde4_xla_ffi_available() in (0, 1)


@pytest.mark.parametrize("kind", ["hexa", "hexa16", "cartpole", "cartpole_bb", "msd"])
def test_tree_packer_is_the_product_packer(kind):
    sys.path.insert(0, os.path.join(ROOT, "ffi"))
    import nsde_b200_jax as J
    pm, model, nn = hp.build(kind, 10, 2, seed=3)
    tp = J.TreePacker(model, nn)
    assert np.array_equal(tp.pack(nn), model.pack(nn))
    # every leaf element lands somewhere in the block, and the gather is a pure copy
    flat = tp.flatten(nn)
    used = np.zeros(flat.shape[0], bool)
    used[tp.src[tp.src >= 0]] = True
    unused = [tp.keys[np.searchsorted(tp.offsets, i, side="right") - 1] for i in np.nonzero(~used)[0]]
    assert all("mu_coeff" in m for m, _ in unused), unused  # (the mu-coefficient net is not part of the rollout block)
    rt = tp.unflatten(flat)
    assert all(np.array_equal(rt[m][k], nn[m][k]) for m, k in tp.keys)
