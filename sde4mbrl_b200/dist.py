"""Multi-GPU plumbing of the path (one process per GPU, ``torch.distributed``).

The path shards with no data-path collective when the units are independent problems / start
states (``problem_shard``).  When ONE problem's particles span GPUs (``particle_shard``) every rank
evaluates its particles with keys ``split(rng, N_total)[offset + n]`` and costs / gradients scaled
by ``1/N_total`` (``sde4_cost_args.particle_offset / particle_total``); a single SUM all-reduce of
``[cost, grad]`` (``H*n_opt + 1`` floats per problem) then yields the mean over all particles
(``jnp.mean``, ``sde4mbrl/nsde.py:1586``).  That is the only collective of the path; over NVLink /
NVSwitch it is latency bound (2.4 KB for the hexacopter, H = 100)."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def particle_shard(n_total, rank=None, world_size=None):
    """Contiguous particle range [offset, offset+count) of this rank (remainder to the first ranks)."""
    if rank is None:
        rank, world_size = world()
    base, rem = divmod(int(n_total), int(world_size))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def problem_shard(num_problems, rank=None, world_size=None):
    """Contiguous problem range of this rank (no collective needed afterwards)."""
    return particle_shard(num_problems, rank, world_size)


def allreduce_cost_grad(cost, grad, group=None, buf=None):
    """In-place SUM all-reduce of the per-problem partial means.  ``cost[P]``, ``grad[P,H,n_opt]`` (or
    ``[H,n_opt]``); packs both into ONE message so a single latency is paid."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return cost, grad
    n = cost.numel() + grad.numel()
    if buf is None or buf.numel() < n or buf.device != cost.device:
        buf = torch.empty(n, dtype=torch.float32, device=cost.device)
    flat = buf[:n]
    flat[:cost.numel()].copy_(cost.reshape(-1))
    flat[cost.numel():].copy_(grad.reshape(-1))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    cost.copy_(flat[:cost.numel()].view_as(cost))
    grad.copy_(flat[cost.numel():].view_as(grad))
    return cost, grad
