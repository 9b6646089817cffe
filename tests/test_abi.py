"""The C-ABI library loads on a CPU-only box and exports every symbol include/sde4b200.h declares
(no compute calls here)."""
import ctypes as C
import os
import re

from sde4mbrl_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sde4b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sde4_[a-z_0-9]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = L.load()
    syms = _declared_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    assert set(L.EXPORTED_SYMBOLS) == set(syms)
    assert lib.sde4_version() == 1


def test_struct_sizes_match():
    lib = L.load()
    sizes = (C.c_int32 * 4)()
    lib.sde4_struct_sizes(sizes)
    assert list(sizes) == [C.sizeof(L.ModelDesc), C.sizeof(L.CostDesc), C.sizeof(L.RolloutArgs), C.sizeof(L.CostArgs)]
    aux = (C.c_int32 * 3)()
    lib.sde4_aux_struct_sizes(aux)
    assert list(aux) == [C.sizeof(L.SimOut), C.sizeof(L.XgpuReduce), C.sizeof(L.DensityRegArgs)]


def test_supported_archs_and_layout_query():
    lib = L.load()
    archs = lib.sde4_supported_archs().decode().split()
    for a in ("msd_bb_d32", "cartpole_si_d32", "hexa_d32_swish", "hexa_d16_tanh", "iris_d32_swish"):
        assert a in archs
    from sde4mbrl_b200 import configs, sde_models
    m = sde_models.SDERotorModel(configs.hexa_model(horizon=5, num_particles=2))
    d = m.describe()
    off, total = m.layout(d)
    assert off[0] == 0 and off[1] == 16 and total % 4 == 0 and off[5] == total
    # 6->32->32->1 density: 6*32+32 + 32*32+32 + 1*32+4
    assert off[3] - off[2] == 6 * 32 + 32 + 32 * 32 + 32 + 32 + 4


def test_unsupported_architecture_is_an_error_not_a_fallback():
    import pytest
    from sde4mbrl_b200 import configs, sde_models
    pm = configs.hexa_model(horizon=5, num_particles=2, density=(20, 20))
    m = sde_models.SDERotorModel(pm)
    with pytest.raises(L.Sde4Error, match="no compiled specialisation"):
        m.layout(m.describe())


def test_argument_validation_without_gpu():
    lib = L.load()
    from sde4mbrl_b200 import configs, sde_models
    m = sde_models.SDERotorModel(configs.hexa_model(horizon=5, num_particles=2))
    d = m.describe()
    a = L.RolloutArgs()
    a.struct_size = 3  # wrong ABI size
    assert lib.sde4_rollout(C.byref(d), C.byref(a), None) == -1
    assert b"size mismatch" in lib.sde4_last_error()
    a.struct_size = C.sizeof(L.RolloutArgs)
    a.P, a.N, a.H = 1, 1, 0
    assert lib.sde4_rollout(C.byref(d), C.byref(a), None) == -1
