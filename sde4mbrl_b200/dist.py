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


class CostGradReducer:
    """SUM all-reduce of ``[cost | grad]`` (``n = P*(H*n_opt + 1)`` floats) for particle-sharded evaluations.

    The message is a few KB, so the collective is pure latency.  ``transport`` is EXPLICIT -- nothing is probed and
    silently replaced:

    ``"fused"``  (default on CUDA) the sum is folded into the reduction kernel K3 (``sde4_xgpu_reduce``): K3 writes this
                 rank's partial into a buffer in torch's symmetric memory, its last CTA publishes a flag in every peer's
                 memory, waits for the peers' flags and sums the world's buffers with peer loads over NVLink / NVSwitch.
                 Two launches per sharded evaluation, no library collective.  Raises if symmetric memory is unavailable.
    ``"symm"``   K3 writes into a symmetric-memory buffer, ``torch.ops.symm_mem.one_shot_all_reduce`` reduces it.
    ``"nccl"``   plain device tensor + ``dist.all_reduce`` (NCCL).
    ``"gloo"``   (default on CPU) plain tensor + ``dist.all_reduce`` on a gloo group: the host-logic tests.

    Use: ``kw = red.launch_kwargs(num_cost)`` -> ``Engine.cost_grad(..., **kw)`` -> ``tot = red.reduce()``."""

    TRANSPORTS = ("fused", "symm", "nccl", "gloo")

    def __init__(self, n, device, group=None, grad_shape=None, transport=None):
        self.group = group if group is not None else (dist.group.WORLD if dist.is_initialized() else None)
        self.n = int(n)
        self.world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        self.device = torch.device(device)
        if transport is None:
            transport = "fused" if self.device.type == "cuda" else "gloo"
        if transport not in self.TRANSPORTS:
            raise ValueError("transport must be one of %s" % (self.TRANSPORTS,))
        if transport in ("fused", "symm") and self.device.type != "cuda":
            raise ValueError("transport '%s' needs a CUDA device" % transport)
        self.transport = transport if self.world > 1 else "none (one rank)"
        self.grad_shape = grad_shape
        self.symm = transport in ("fused", "symm") and self.world > 1
        self._epoch = 0
        self._xg = None
        if self.world > 1 and transport == "fused":
            import torch.distributed._symmetric_memory as sm
            from . import _lib as L
            if self.world > 8:
                raise ValueError("the fused reduction covers one box (<= 8 ranks)")
            # symmetric allocation: [ partial buffer, parity 0 | parity 1 | flags uint32[world] (+ padding) ]
            self._nf = (self.n + 3) & ~3
            tot = 2 * self._nf + 8
            self._sym = sm.empty(tot, dtype=torch.float32, device=self.device)
            self._sym.zero_()
            hdl = sm.rendezvous(self._sym, self.group.group_name)
            self._hdl = hdl
            self._ptrs = [int(p) for p in hdl.buffer_ptrs]
            self.out = torch.zeros(self.n, dtype=torch.float32, device=self.device)
            self._counter = torch.zeros(1, dtype=torch.int32, device=self.device)
            torch.cuda.synchronize(self.device)
            dist.barrier(self.group)  # every rank's flags are zero before anybody publishes
            xg = L.XgpuReduce()
            xg.world, xg.rank, xg.n = self.world, self.rank, self.n
            for r in range(self.world):
                xg.flag[r] = self._ptrs[r] + 4 * 2 * self._nf
            xg.out, xg.counter = self.out.data_ptr(), self._counter.data_ptr()
            self._xg = xg
            self.buf = None
        elif self.world > 1 and transport == "symm":
            import torch.distributed._symmetric_memory as sm
            self.buf = sm.empty(self.n, dtype=torch.float32, device=self.device)
            sm.rendezvous(self.buf, self.group.group_name)
            self.buf.zero_()
        else:
            self.buf = torch.zeros(self.n, dtype=torch.float32, device=self.device)

    def views(self, num_cost):
        """(cost[num_cost], grad[...]) views of the buffer the producer writes (not used by the fused transport)."""
        buf = self.buf if self.buf is not None else self.out
        g = buf[num_cost:]
        return buf[:num_cost], (g.view(self.grad_shape) if self.grad_shape is not None else g)

    def launch_kwargs(self, num_cost):
        """Keyword arguments of ``Engine.cost_grad`` that route its [cost | grad] into this reducer."""
        if self._xg is not None:
            self._epoch += 1
            xg = self._xg
            xg.epoch = self._epoch
            off = 4 * self._nf * (self._epoch & 1)
            for r in range(self.world):
                xg.buf[r] = self._ptrs[r] + off
            c, g = self.views(num_cost)  # placeholders (shape checks only): the kernel writes the symmetric buffer
            return dict(cost_out=c, grad_out=g, xgpu=xg)
        c, g = self.views(num_cost)
        return dict(cost_out=c, grad_out=g)

    def reduce(self):
        """Returns the summed ``[cost | grad]`` vector (valid until the next evaluation)."""
        if self.world == 1:
            return self.buf
        if self._xg is not None:
            return self.out  # K3 already summed the world's partials
        if self.transport == "symm":
            return torch.ops.symm_mem.one_shot_all_reduce(self.buf, "sum", self.group.group_name)
        dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.group)
        return self.buf
