"""N>1 host logic on CPU (gloo, world size 2): particle / problem sharding and the single
[cost, grad] SUM all-reduce reproduce the single-process mean over all particles.  The per-shard
partials come from the oracle (the CUDA kernels implement the same particle_offset / particle_total
contract; that part is covered on the GPU by test_gpu_costgrad.py::test_particle_sharding)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world_size, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as hp
    from oracle import nsde_oracle as orc
    from oracle import threefry as tf
    from sde4mbrl_b200 import dist as D

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    torch.set_num_threads(1)
    kind, H, N_total = "cartpole", 10, 7
    pm, model, nn = hp.build(kind, H, N_total, seed=2)
    om = hp.oracle_model(kind, pm, nn, torch.float64)
    rng = np.random.default_rng(0)
    y0 = torch.tensor(hp.start_state(kind, rng), dtype=torch.float64)
    opt = torch.tensor(hp.controls(kind, H, rng), dtype=torch.float64)
    xref = torch.zeros(H + 1, 4, dtype=torch.float64)
    from sde4mbrl_b200 import configs
    cp = configs.cartpole_mpc()["cost_params"]
    dw_all = torch.tensor(tf.rollout_noise(tf.PRNGKey(5), N_total, H, 4), dtype=torch.float64)

    def partial(lo, n):
        o = opt.clone().requires_grad_(True)
        c, _ = orc.compute_cost(om, y0, o, dw_all[lo:lo + n], hp.stage_cost_fn(kind, cp), xref)
        c = c * n / N_total  # shard mean -> contribution to the global mean (the kernels scale by 1/N_total)
        (g,) = torch.autograd.grad(c, o)
        return c.detach().float().reshape(1), g.float()

    off, cnt = D.particle_shard(N_total)
    c, g = partial(off, cnt)
    D.allreduce_cost_grad(c, g)
    cf, gf = partial(0, N_total)
    ok = abs(float(c) - float(cf)) < 1e-5 * abs(float(cf)) and float((g - gf).abs().max()) < 1e-5 * float(gf.abs().max())
    # the reducer object (CPU / gloo: plain buffer + all_reduce; symmetric memory is a CUDA-only fast path)
    c2, g2 = partial(off, cnt)
    red = D.CostGradReducer(1 + g2.numel(), "cpu", grad_shape=tuple(g2.shape))
    cv, gv = red.views(1)
    cv.copy_(c2)
    gv.copy_(g2)
    tot = red.reduce()
    ok = ok and red.transport == "gloo" and not red.symm and abs(float(tot[0]) - float(cf)) < 1e-5 * abs(float(cf)) \
        and float((tot[1:].view_as(gf) - gf).abs().max()) < 1e-5 * float(gf.abs().max())
    shards = [None] * world_size
    dist.all_gather_object(shards, (off, cnt))
    if rank == 0:
        out.put((ok, shards, D.problem_shard(10)))
    dist.destroy_process_group()


def test_particle_sharding_allreduce_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, shards, pshard = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
    assert shards == [(0, 4), (4, 3)]  # 7 particles over 2 ranks: remainder to the first rank
    assert pshard == (0, 5)


def test_shard_arithmetic():
    from sde4mbrl_b200 import dist as D
    for n, w in ((4096, 8), (7, 2), (5, 8), (1, 4)):
        parts = [D.particle_shard(n, r, w) for r in range(w)]
        assert sum(c for _, c in parts) == n
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
