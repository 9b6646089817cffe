"""Small-message all-reduce latency: NCCL vs torch symmetric-memory one-shot all-reduce (P2P over NVLink).
run: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/symm_allreduce_test.py"""
import os, sys, time, torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = 601
x = torch.full((n,), float(rank + 1), device=dev)
def timeit(fn, reps=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
t_nccl = timeit(lambda: dist.all_reduce(x.clone()))
if rank == 0: print("nccl all_reduce us", t_nccl, flush=True)
try:
    import torch.distributed._symmetric_memory as sm
    g = dist.group.WORLD
    buf = sm.empty(n, dtype=torch.float32, device=dev)
    hdl = sm.rendezvous(buf, g.group_name)
    buf.fill_(float(rank + 1)); torch.cuda.synchronize(); dist.barrier()
    out = torch.ops.symm_mem.one_shot_all_reduce(buf, "sum", g.group_name)
    torch.cuda.synchronize()
    want = world * (world + 1) / 2
    ok = bool((out == want).all())
    t_sm = timeit(lambda: torch.ops.symm_mem.one_shot_all_reduce(buf, "sum", g.group_name))
    if rank == 0: print("symm_mem one_shot_all_reduce us", t_sm, "correct", ok, flush=True)
except Exception as ex:
    if rank == 0: print("symm_mem unavailable:", repr(ex)[:300], flush=True)
dist.destroy_process_group()
