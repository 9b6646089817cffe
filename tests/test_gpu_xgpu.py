"""Cross-GPU sum of [cost | grad] folded into the reduction kernel (sde4_xgpu_reduce): two devices driven from ONE
process with plain peer access (the kernel only needs pointers that are valid on the launching device).  Also covers
the per-device kernel attributes: device 1 launches kernels that need more than 48 KB of dynamic shared memory."""
import numpy as np
import pytest
import torch

import helpers as hp
from oracle import nsde_oracle as orc
from oracle import threefry as tf
from sde4mbrl_b200 import _lib as L
from sde4mbrl_b200 import configs, costs

pytestmark = pytest.mark.gpu


def _two_gpus():
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2 and torch.cuda.can_device_access_peer(0, 1)


@pytest.mark.skipif(not _two_gpus(), reason="needs two GPUs with peer access")
@pytest.mark.parametrize("lanes,N", [(0, 2048), (8, 64), (1, 32)])
def test_fused_cross_gpu_reduce(lanes, N):
    from sde4mbrl_b200.engine import Engine
    H, world = 20, 2
    rng = np.random.default_rng(3)
    pm, model, nn = hp.build("hexa", H, N, seed=2)
    y0, opt = hp.start_state("hexa", rng), hp.controls("hexa", H, rng)
    cp = configs.hexa_mpc()["cost_params"]
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    xref = np.tile(configs.hover_state(), (H + 1, 1)).astype(np.float32)
    ts = orc.compute_timesteps(pm)
    key = tf.PRNGKey(5)
    n = 1 + H * 6
    nf = (n + 3) & ~3
    torch.zeros(1, device="cuda:0").to("cuda:1")  # torch enables peer access on the first cross-device copy
    torch.zeros(1, device="cuda:1").to("cuda:0")
    eng, sym, out, cnt, y0d, optd, xrefd = [], [], [], [], [], [], []
    for r in range(world):
        with torch.cuda.device(r):
            eng.append(Engine(model))
            sym.append(torch.zeros(2 * nf + 8, device="cuda:%d" % r))
            out.append(torch.zeros(n, device="cuda:%d" % r))
            cnt.append(torch.zeros(1, dtype=torch.int32, device="cuda:%d" % r))
            y0d.append(torch.tensor(y0[None], device="cuda:%d" % r))
            optd.append(torch.tensor(opt, device="cuda:%d" % r))
            xrefd.append(torch.tensor(xref, device="cuda:%d" % r))
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    with torch.cuda.device(0):  # the answer: all 2 N particles on one device
        cw, gw, _ = eng[0].cost_grad(nn, y0d[0], optd[0], ts, xrefd[0], stage.desc, key=key, N=world * N, lanes=lanes)
        cw, gw = float(cw[0]), gw.cpu().numpy()
    for epoch in (1, 2, 3):  # both parities of the partial buffers, flags carried over
        xs = []
        for r in range(world):
            with torch.cuda.device(r):
                xg = L.XgpuReduce()
                xg.world, xg.rank, xg.n, xg.epoch = world, r, n, epoch
                for q in range(world):
                    xg.buf[q] = sym[q].data_ptr() + 4 * nf * (epoch & 1)
                    xg.flag[q] = sym[q].data_ptr() + 4 * 2 * nf
                xg.out, xg.counter = out[r].data_ptr(), cnt[r].data_ptr()
                xs.append(xg)
                eng[r].cost_grad(nn, y0d[r], optd[r], ts, xrefd[r], stage.desc, key=key, N=N, lanes=lanes,
                                 particle_offset=r * N, particle_total=world * N, xgpu=xg)
        torch.cuda.synchronize(0)
        torch.cuda.synchronize(1)
        o0, o1 = out[0].cpu().numpy(), out[1].cpu().numpy()
        assert np.array_equal(o0, o1)  # rank-ordered sum: the same bits on both devices
        assert abs(o0[0] - cw) <= 2e-6 * abs(cw)
        assert hp.rel_err(o0[1:].reshape(H, 6), gw) < 1e-5
        out[0].zero_()
        out[1].zero_()
