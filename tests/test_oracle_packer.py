"""oracle/packer.py (descriptor + parameter block built straight from the reference's params dict and haiku tree) against
the product's describe() / pack() -- two independent implementations of the layout include/sde4b200.h documents -- and
the C oracle fed by it against the torch oracle fed by the raw tree."""
import os

import numpy as np
import pytest
import torch

import helpers as hp
from oracle import c_oracle, packer
from oracle import nsde_oracle as orc
from sde4mbrl_b200 import configs, costs, io, sde_models

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("kind,scale", [("hexa", "fan_in"), ("hexa", "yaml"), ("hexa16", "fan_in")])
def test_packer_matches_product_on_random_trees(kind, scale):
    pm, model, nn = hp.build(kind, 20, 4, seed=2, scale=scale)
    d = model.describe()
    od = packer.rotor_desc(pm)
    assert bytes(d) == bytes(od)
    assert packer.rotor_layout(od) == model.layout(d)[0]
    assert np.array_equal(packer.rotor_pack(pm, nn, od), model.pack(nn, d))


@pytest.mark.parametrize("fixture", ["ref_weights_hexa_ahg_v4.npz", "ref_weights_iris_sitl.npz"])
def test_packer_matches_product_on_shipped_weights(fixture):
    nn, nominal = io.load_npz_fixture(os.path.join(GOLD, fixture))
    model = sde_models.SDERotorModel(nominal)
    d = model.describe()
    od = packer.rotor_desc(nominal)
    assert bytes(d) == bytes(od)
    assert np.array_equal(packer.rotor_pack(nominal, nn, od), model.pack(nn, d))


def test_cost_descriptor_matches_product():
    cp = dict(configs.hexa_mpc()["cost_params"])
    cp["verr_sig"] = [1.5, 1.5, 1.5]
    cp["res_sig"] = [0.5, 2.0]
    want = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6).desc
    assert bytes(want) == bytes(packer.rotor_track_cost_desc(cp, 6))


def test_c_oracle_fed_by_oracle_packer_matches_torch_oracle():
    H, N = 20, 16
    rng = np.random.default_rng(4)
    pm, _, nn = hp.build("hexa", H, N, seed=7)
    y0, opt = hp.start_state("hexa", rng), hp.controls("hexa", H, rng)
    dw = rng.standard_normal((N, H, 13)).astype(np.float32)
    cp = configs.hexa_mpc()["cost_params"]
    xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
    ts = orc.compute_timesteps(pm)
    od = packer.rotor_desc(pm)
    c, g, _ = c_oracle.rotor_cost_grad(od, packer.rotor_track_cost_desc(cp, 6), packer.rotor_pack(pm, nn, od), y0, opt, ts,
                                       xref, N, noise=dw)
    om = hp.oracle_model("hexa", pm, nn, torch.float64)
    o = torch.tensor(opt, dtype=torch.float64, requires_grad=True)
    c64, _ = orc.compute_cost(om, torch.tensor(y0, dtype=torch.float64), o, torch.tensor(dw, dtype=torch.float64),
                              hp.stage_cost_fn("hexa", cp), torch.tensor(xref, dtype=torch.float64))
    c64 = c64 + orc.u_slew_cost(o, ts, cp, 6)
    (g64,) = torch.autograd.grad(c64, o)
    assert abs(c - c64.item()) <= 2e-5 * abs(c64.item())
    assert hp.rel_err(g, g64.numpy()) < 1e-3
