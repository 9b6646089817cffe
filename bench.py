#!/usr/bin/env python
"""bench.py -- headline benchmark of the sde4mbrl hot path on B200.

Workload (BASELINE.json configs[2], the one the >=100x target is quoted on): hexacopter nSDE,
ONE problem, 4096 particles x 100 Euler-Maruyama steps with the distance-aware diffusion
(in-kernel threefry noise) AND the reverse-mode gradient pass  ->  one "step" = one fused
value+gradient evaluation of the MPC cost (sde4_cost_grad: K2 rollout+cost+adjoint, K3 reduce).
Metric: particle-steps/s = particles x horizon / time per step (whole job, all GPUs).
With N GPUs (weak scaling) every rank holds 4096 particles of the SAME problem (global
N = 4096 x GPUs, keys = split(rng, N_global)) and the only collective is the NCCL all-reduce of
the per-problem [cost, grad] sums (601 floats) -- the north star's single exchange step.

Extra fields: forward-only rollout throughput, batched-MPC numbers, the measured FP32 FMA peak
that the roofline fraction is taken against, the CPU baseline (oracle port on the host cores).

`--impl reference` times the reference's CPU algorithm (the oracle port -- jax is not installed in
this image) on the same workload / metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic work, SURVEY.md 8(a)/(d): hexacopter F_fwd = 3790 flop per particle-step
# (+90 stage cost); value+grad = 2 x F_fwd
F_FWD_HEXA = 3790.0
F_VALGRAD_HEXA = 2.0 * 3880.0
N_PART, HORIZON = 4096, 100


def bench_config(gpus):
    """`config` of the JSON line -- the SAME dict from both arms (the driver compares them)."""
    return {"workload": "hexacopter nSDE (13 states, 6 rotors), 1 problem, 4096 particles per GPU x 100 steps, "
                        "distance-aware diffusion with in-kernel threefry noise + gradient pass (BASELINE configs[2])",
            "particles_per_gpu": N_PART, "horizon": HORIZON, "gpus": gpus,
            "parallelism": "particles of the one problem sharded over the GPUs, one all-reduce of [cost, grad] (601 floats)",
            "l2": "GPU arm: flushed between timed steps (256 MB write); CPU arm: 44 MB working set per step",
            "timing": "GPU arm: CUDA events per step, max over ranks; CPU arm: perf_counter per step"}


def build_workload(seed=0):
    from sde4mbrl_b200 import configs, costs, sde_models
    pm = configs.hexa_model(horizon=HORIZON, num_particles=N_PART, stepsize=0.01)
    model = sde_models.SDERotorModel(pm)
    nn = configs.random_params(model, seed=seed, scale="fan_in")
    rng = np.random.default_rng(0)
    y0 = configs.hover_state()
    y0[3:6] += rng.normal(0, 0.01, 3).astype(np.float32)
    y0[10:13] += rng.normal(0, 0.01, 3).astype(np.float32)
    opt = rng.uniform(0.30, 0.36, (HORIZON, 6)).astype(np.float32)
    xref = np.tile(configs.hover_state(), (HORIZON + 1, 1)).astype(np.float32)
    xref[:, :3] = [0.0, 0.5, 1.4]
    cp = configs.hexa_mpc()["cost_params"]
    stage = costs.with_slew(costs.one_step_cost_f(cp, 6), cp, 6)
    return dict(pm=pm, model=model, nn=nn, y0=y0, opt=opt, xref=xref, cp=cp, stage=stage,
                key=np.array([0, 2050], np.uint32))


class ClockSampler:
    """Samples SM clocks / throttle reasons DURING the timed region: NVML (what nvidia-smi reads) polled every
    2 ms from a thread -- the timed region is only ~10-20 ms long; `nvidia-smi` itself (one process per sample) is the
    fall-back when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        nv = self._nvml
        sm = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        flag = lambda bit: "Active" if (r & bit) else "Not Active"
        # bit masks of nvmlClocksEventReasons: SwPowerCap 0x4, HwSlowdown 0x8, SwThermalSlowdown 0x20, HwThermalSlowdown 0x40
        return [str(sm), str(mx), "", flag(0x8), flag(0x40), flag(0x20), flag(0x4)]

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self.samples.append(self._sample_nvml())
                    self._stop.wait(0.002)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples), "source": "nvml" if self._nvml is not None else "nvidia-smi"}


def cpu_reference_step(w, n_particles, threads):
    """One value+gradient evaluation with the CPU restatement of the reference's algorithm:
    oracle/c (plain C, AVX2/AVX-512 lanes over particles, OpenMP over particle chunks, in-line
    threefry noise) -- the kind of code XLA:CPU would emit for the reference's vmapped scan."""
    from oracle import c_oracle, packer
    if "c_packed" not in w:  # descriptors and parameter block from the oracle-side packer: nothing of libsde4b200.so
        w["c_desc"] = packer.rotor_desc(w["pm"])
        w["c_cost"] = packer.rotor_track_cost_desc(w["cp"], 6)
        w["c_packed"] = packer.rotor_pack(w["pm"], w["nn"], w["c_desc"])
    ts = np.full(HORIZON, 0.01, np.float32)
    t0 = time.perf_counter()
    c, g, _ = c_oracle.rotor_cost_grad(w["c_desc"], w["c_cost"], w["c_packed"], w["y0"], w["opt"], ts, w["xref"],
                                       n_particles, key=w["key"], nthreads=threads)
    return time.perf_counter() - t0, c, g


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; jax absent)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    w = build_workload()
    n_sample = N_PART  # both CPU arms run the full 4096-particle workload per step
    warm = max(args.warmup, 3)
    # first choice: the reference itself, jit-compiled by jax on the host cores (the arm the >= 100x target is defined
    # on); only where jax + dm-haiku and a copy of the reference exist (baseline/ref_jax_cpu.py) -- not in the build image
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_jax_cpu
    ok, why = ref_jax_cpu.available()
    kind, sample = "port", None
    if ok:
        try:
            from sde4mbrl_b200 import configs
            mpc = configs.hexa_mpc(horizon=HORIZON, dt=0.01, num_particles=N_PART)
            times = ref_jax_cpu.time_rotor_value_and_grad(w["pm"], mpc, w["cp"], w["nn"], w["y0"], w["opt"], w["key"],
                                                          w["xref"], args.steps, warm)
            kind = "jax"
            sample = ("%d particles x %d steps value+grad per step; the unmodified reference, jax.jit(jax.value_and_grad("
                      "cost_xt)) on the CPU backend" % (n_sample, HORIZON))
        except Exception as ex:  # a broken jax install must not lose the arm
            why = "jax arm failed: %r" % (ex,)
            ok = False
    if not ok:
        from oracle import c_oracle
        simd = c_oracle.load().sde4o_simd_width()
        for _ in range(warm):
            cpu_reference_step(w, n_sample, threads)
        times = [cpu_reference_step(w, n_sample, threads)[0] for _ in range(args.steps)]
        sample = ("%d particles x %d steps value+grad per step; oracle/c (C restatement of the reference, AVX-512/AVX2 + "
                  "OpenMP, SIMD width %d) because %s" % (n_sample, HORIZON, simd, why))
    t = float(np.mean(times))
    val = n_sample * HORIZON / t
    line = {"impl": "reference", "metric": "particle-steps/s (value+grad, hexacopter nSDE)", "value": val,
            "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": bench_config(args.gpus),
            "cpu_baseline": {"value": val, "unit": "particle-steps/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def mpc_batch_run(w, n_part, world, rank, dev, barrier, reps=4, e2e=False):
    """BASELINE configs[3]: 4096 independent hexacopter MPC solves (H = 20, dt = 0.05, 20 APG iterations,
    mpc_test_cfg.yaml), PARTITIONED over the ranks with no collective (strong scaling).  Device time per batch (CUDA
    events, max over ranks); with `e2e` also the wall time from pinned host start states / keys to host controls."""
    import torch
    import torch.distributed as dist
    from sde4mbrl_b200 import configs, sde_models
    from sde4mbrl_b200.apg_batched import BatchedAPG
    from sde4mbrl_b200.engine import Engine
    P_TOTAL, HM = 4096, 20
    p_lo, p_cnt = rank * (P_TOTAL // world), P_TOTAL // world
    rngm = np.random.default_rng(123)
    y0m = np.tile(w["y0"], (P_TOTAL, 1))
    y0m[:, 3:6] += rngm.normal(0, 0.05, (P_TOTAL, 3)).astype(np.float32)
    y0m[:, 10:13] += rngm.normal(0, 0.05, (P_TOTAL, 3)).astype(np.float32)
    keysm = np.stack([np.zeros(P_TOTAL, np.uint32), np.arange(P_TOTAL, dtype=np.uint32)], axis=1)
    xrefm = np.tile(w["xref"][:1], (HM + 1, 1))
    pmm = configs.hexa_model(horizon=HM, num_particles=n_part, stepsize=0.05)
    mpcm = configs.hexa_mpc(horizon=HM, dt=0.05, num_particles=n_part, max_iter=20)
    engm = Engine(sde_models.SDERotorModel(pmm))
    solver = BatchedAPG(engm, w["nn"], w["stage"], mpcm["apg_mpc"], np.full(HM, 0.05, np.float32), n_part, lb=1e-4, ub=1.0)
    x0m = torch.full((HM, 6), 0.33 + 2e-4, device=dev)
    y0d = torch.as_tensor(y0m[p_lo:p_lo + p_cnt], device=dev)
    kd = torch.as_tensor(keysm[p_lo:p_lo + p_cnt].view(np.int32), device=dev)
    tms = []
    for rep in range(reps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        solver.init(x0m, y0d, xrefm, kd).run(20)
        e1.record()
        torch.cuda.synchronize()
        if rep:
            tms.append(e0.elapsed_time(e1))
    tm = float(np.mean(tms))
    if world > 1:
        t = torch.tensor([tm], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tm = float(t.item())
    out = {"solves_per_s": P_TOTAL / (tm * 1e-3), "ms_per_batch": tm, "problems": P_TOTAL, "problems_per_gpu": p_cnt,
           "particles": n_part, "horizon": HM, "apg_iterations": 20, "mean_init_cost": float(solver.t["init_cost"].mean()),
           "mean_opt_cost": float(solver.t["opt_cost"].mean()), "rollout_equiv_per_solve": 2 + 20 * 9,
           "mode": "speculative" if solver.speculative else ("merged x/v" if solver.merge_xv_grad else
                                                             ("progressive line search" if solver.progressive_ls else "stacked line search"))}
    if e2e:
        y0p = torch.as_tensor(y0m[p_lo:p_lo + p_cnt]).pin_memory()
        kp = torch.as_tensor(keysm[p_lo:p_lo + p_cnt].view(np.int32)).pin_memory()
        uo = torch.empty((p_cnt, HM, 6)).pin_memory()
        tws = []
        for rep in range(reps):
            barrier()
            t0 = time.perf_counter()
            solver.init(x0m, y0p.to(dev, non_blocking=True), xrefm, kp.to(dev, non_blocking=True)).run(20)
            uo.copy_(solver.x_opt(), non_blocking=True)
            torch.cuda.synchronize()
            if rep:
                tws.append(time.perf_counter() - t0)
        te = float(np.mean(tws))
        if world > 1:
            t = torch.tensor([te], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        out["e2e_s"] = te
        out["h2d_bytes"] = int(y0p.numel() * 4 + kp.numel() * 4)
        out["d2h_bytes"] = int(uo.numel() * 4)
    return out


def run_mpc_batch(args):
    """--workload mpc_batch: ONE JSON line whose value is whole-job APG-MPC solves/s (strong scaling over the GPUs)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    w = build_workload()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    main_r = mpc_batch_run(w, args.particles, world, rank, dev, barrier, reps=max(args.steps, 3) + 1, e2e=True)
    other = mpc_batch_run(w, 1 if args.particles != 1 else 32, world, rank, dev, barrier, reps=4)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        line = {"metric": "APG-MPC solves/s (hexacopter nSDE, 4096 independent problems)", "value": main_r["solves_per_s"],
                "unit": "solves/s", "n_gpus": world, "steps": max(args.steps, 3), "warmup": 1,
                "ms_per_step": main_r["ms_per_batch"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "4096 independent hexacopter MPC solves (H = 20, dt = 0.05, 20 APG iterations, box "
                                       "prox, line search maxls 4), %d particle(s) per problem, problems partitioned over the "
                                       "GPUs, no collective (BASELINE configs[3])" % args.particles,
                           "problems": 4096, "particles": args.particles, "horizon": 20, "apg_iterations": 20, "gpus": world,
                           "l2": "working set per batch exceeds L2 at 32 particles (283 MB of checkpoints); not flushed",
                           "timing": "CUDA events around one batch (init + 20 iterations), max over ranks"},
                "clocks": clocks,
                "e2e": {"value": 4096 / main_r["e2e_s"], "unit": "solves/s", "h2d_bytes_per_step": main_r["h2d_bytes"],
                        "d2h_bytes_per_step": main_r["d2h_bytes"], "ms_per_step": main_r["e2e_s"] * 1e3,
                        "api": "apg_batched.BatchedAPG.init(pinned host start states / keys).run(20) + x_opt to pinned host"},
                "gpu_launches": "one CUDA-graph replay per APG iteration (6 bookkeeping + 2-4 cost + reduce launches)",
                "mpc_batch": {"N%d" % args.particles: main_r, "N%d" % other["particles"]: other}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--flush-mode", default="write", choices=["write", "write+read", "none"],
                    help="diagnostic: how L2 is flushed between timed steps (the bench line is quoted with 'write')")
    ap.add_argument("--block-threads", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=0, help="lanes per sample (0 = library heuristic)")
    ap.add_argument("--workload", default="particles", choices=["particles", "mpc_batch"],
                    help="particles: BASELINE configs[2] (the default line); mpc_batch: BASELINE configs[3], 4096 independent "
                         "hexacopter MPC solves partitioned over the GPUs, value = solves/s, strong scaling")
    ap.add_argument("--particles", type=int, default=32, help="mpc_batch: particles per problem (the reference deploys 1)")
    ap.add_argument("--transport", default="fused", choices=["fused", "symm", "nccl"],
                    help="cross-GPU sum of [cost, grad] at N > 1: folded into K3 (peer loads over NVLink), torch one-shot, NCCL")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "mpc_batch":
        return run_mpc_batch(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from sde4mbrl_b200 import _lib as L
    from sde4mbrl_b200.engine import Engine, PackedParams

    w = build_workload()
    eng = Engine(w["model"])
    pk = PackedParams(w["model"], w["nn"], eng.desc)
    ts = np.full(HORIZON, 0.01, np.float32)
    y0_d = torch.as_tensor(w["y0"], device=dev)[None]
    opt_d = torch.as_tensor(w["opt"], device=dev)
    xref_d = torch.as_tensor(w["xref"], device=dev)
    key_d = torch.as_tensor(w["key"].view(np.int32), device=dev)
    n_total = N_PART * world
    red = torch.zeros(HORIZON * 6 + 1, device=dev)
    # N > 1: K3 writes its [cost | grad] straight into the reducer's (symmetric-memory) buffer, ONE small all-reduce
    from sde4mbrl_b200.dist import CostGradReducer
    reducer = CostGradReducer(HORIZON * 6 + 1, dev, grad_shape=(HORIZON, 6), transport=args.transport) if world > 1 else None

    def step_device():
        if world > 1:
            eng.cost_grad(pk, y0_d, opt_d, ts, xref_d, w["stage"].desc, key=key_d, N=N_PART, want_grad=True,
                          block_threads=args.block_threads, lanes=args.lanes, particle_offset=rank * N_PART,
                          particle_total=n_total, **reducer.launch_kwargs(1))
            tot = reducer.reduce()
            return tot[0:1], tot[1:].view(HORIZON, 6)
        cost, grad, _ = eng.cost_grad(pk, y0_d, opt_d, ts, xref_d, w["stage"].desc, key=key_d, N=N_PART,
                                      want_grad=True, block_threads=args.block_threads, lanes=args.lanes,
                                      particle_offset=rank * N_PART, particle_total=n_total)
        return cost, grad

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- device-resident timing: K steps, each bracketed by CUDA events, L2 flushed in between
    evs = []
    barrier()
    t_wall0 = time.perf_counter()
    flush_rd = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if args.flush_mode == "write+read" else None
    for _ in range(args.steps):
        if args.flush_mode != "none":
            flush.fill_(1)
        if flush_rd is not None:
            flush_rd.view(torch.int32).sum()  # evicts the dirty flush lines: nothing of the flush is written back in the timed region
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_device()
        e1.record()
        evs.append((e0, e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    ms_per_step = float(np.mean(step_ms))
    if world > 1:
        t = torch.tensor([ms_per_step], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step = float(t.item())

    # ---- end-to-end through the public API with HOST buffers (pinned): H2D + kernels + D2H
    from sde4mbrl_b200 import nsde, random as R
    from sde4mbrl_b200 import configs, sde_models
    mpc = configs.hexa_mpc(horizon=HORIZON, dt=0.01, num_particles=N_PART)
    pm2 = configs.hexa_model(horizon=HORIZON, num_particles=N_PART, stepsize=0.01)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        _, multi_cost, _, _, _ = nsde.create_online_cost_sampling_fn(pm2, mpc, sde_constr=sde_models.SDERotorModel)
    e2e_eng = multi_cost.engine
    e2e_pk = PackedParams(e2e_eng.model, w["nn"], e2e_eng.desc)
    y0_h = torch.as_tensor(w["y0"]).pin_memory()
    opt_h = torch.as_tensor(w["opt"]).pin_memory()
    xref_h = torch.as_tensor(w["xref"]).pin_memory()
    key_h = torch.as_tensor(w["key"].view(np.int32)).pin_memory()
    out_h = torch.empty(HORIZON * 6 + 1).pin_memory()
    h2d = (y0_h.numel() + opt_h.numel() + xref_h.numel()) * 4 + 8
    d2h = out_h.numel() * 4

    opt_host = np.ascontiguousarray(w["opt"])
    key_np = np.ascontiguousarray(w["key"])

    def step_e2e_host():
        # N = 1: host arrays straight into the reference-facing call; libsde4b200 packs them into its pinned
        # staging buffer, does one H2D copy, K2 + K3, one D2H copy of [cost | grad] and a stream synchronise
        (c, _), g = multi_cost.value_and_grad(e2e_pk, w["y0"], opt_host, key_np, w["stage"], extra_cost_args=w["xref"])
        return c, g  # CPU tensors: the device->host read of [cost | grad] happened inside the call

    if world > 1:
        # N > 1: the same host-array call with the particles of the problem sharded over the ranks -- one packed pinned
        # H2D copy, K2 + K3 into the symmetric-memory buffer, one-shot all-reduce, one D2H copy (nsde.py, particle_reducer)
        multi_cost.particle_reducer = CostGradReducer(HORIZON * 6 + 1, dev, grad_shape=(HORIZON, 6), transport=args.transport)

    def step_e2e():
        return step_e2e_host()

    def step_e2e_autograd():  # (the generic autograd route, kept for scripts/e2e_breakdown.py comparisons)
        y = y0_h.to(dev, non_blocking=True)
        o = opt_h.to(dev, non_blocking=True).requires_grad_(True)
        xr = xref_h.to(dev, non_blocking=True)
        k = key_h.to(dev, non_blocking=True)
        c, _ = multi_cost(e2e_pk, y, o, k, w["stage"], extra_cost_args=xr)
        (g,) = torch.autograd.grad(c, o)
        if world > 1:
            red[0:1].copy_(c.reshape(1))
            red[1:].copy_(g.reshape(-1))
            dist.all_reduce(red)
            out_h.copy_(red, non_blocking=True)
        else:
            out_h[0:1].copy_(c.reshape(1), non_blocking=True)
            out_h[1:].copy_(g.reshape(-1), non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(3):
        step_e2e()
    if world > 1:  # the host-array route and the device-resident step agree on the all-reduced [cost | grad]
        c_chk, g_chk = step_e2e()
        c_dev, g_dev = step_device()
        torch.cuda.synchronize()
        assert abs(float(c_chk) - float(c_dev[0])) <= 1e-5 * abs(float(c_dev[0])), (float(c_chk), float(c_dev[0]))
        assert float((g_chk - g_dev.cpu()).abs().max()) <= 1e-5 * float(g_dev.abs().max()) + 1e-9
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    # ---- BASELINE configs[4] (second half): cartpole training loss forward/backward, B=512, N=1, H=30
    train_extra = {}
    if rank == 0:
        try:
            from sde4mbrl_b200 import costs as _c
            BT, HT = 512, 30
            pmt = configs.cartpole_model(horizon=HT, num_particles=1)
            mt = sde_models.CartPoleSDE(pmt)
            engt = Engine(mt)
            pkt = PackedParams(mt, configs.random_params(mt, seed=0), engt.desc)
            g = torch.Generator(device=dev).manual_seed(5)
            ym = torch.randn((BT, HT + 1, 4), device=dev, generator=g) * 0.5
            ut = torch.rand((BT, HT, 1), device=dev, generator=g) * 2 - 1
            kt = torch.stack([torch.zeros(BT, dtype=torch.int32, device=dev), torch.arange(BT, dtype=torch.int32, device=dev)], dim=1).contiguous()
            owt = np.array([9, 5, 1, 1.0, 10])
            dc = _c.QuadraticCost(list(1.0 / owt ** 2), [0.0], [0.0], feat="cartpole")
            tst = np.full(HT, 0.02, np.float32)
            tms = []
            for rep in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                engt.loss_grad(pkt, ym, ut, tst, dc.desc, kt, N=1)
                e1.record()
                torch.cuda.synchronize()
                if rep > 1:
                    tms.append(e0.elapsed_time(e1))
            tm = float(np.mean(tms))
            train_extra = {"batch": BT, "particles": 1, "horizon": HT, "ms_fwd_bwd": tm,
                           "particle_steps_per_s": BT * HT / (tm * 1e-3),
                           "fp32_tflops_algorithmic": 3.0 * 3135.0 * BT * HT / (tm * 1e-3) / 1e12}
            # the complete loss_fn of cartpole_sde.yaml:95-137 through the public API: data term (above) +
            # distance-aware regulariser at the data points (20 ball samples) + weight penalty, value and gradient
            import contextlib as _cl, io as _io2
            pml = configs.cartpole_model(horizon=1, num_particles=1)
            lpl = {"seed": 0, "stepsize": 0.02, "data_stepsize": 0.02, "num_particles": 1, "num_particles_test": 1, "horizon": HT,
                   "obs_weights": [9, 5, 1, 1.0, 10], "pen_data": 1.0, "pen_weights": 1e-3, "default_weights": 1.0,
                   "special_parameters_pen": {"scaler": 0, "density": 0},
                   "density_loss": {"learn_mucoeff": {"type": "dnn", "init_value": 0.01, "activation_fn": "tanh", "hidden_layers": [8, 8]},
                                    "mu_coeff": 10.0, "ball_radius": [0.1, 0.1, 0.4], "ball_nsamples": 20},
                   "pen_grad_density": 1.0, "pen_density_scvex": 1.0, "pen_mu_type": "lin_inv", "pen_mu_coeff": 1000.0,
                   "pen_scvex_mult": 0.01}
            with _cl.redirect_stdout(_io2.StringIO()):
                nnl, loss_fn_l, _, _ = nsde.create_model_loss_fn(pml, lpl, sde_constr=sde_models.CartPoleSDE, verbose=False)
            tws = []
            for rep in range(5):
                t0 = time.perf_counter()
                tot_l, info_l, grads_l = loss_fn_l.value_and_grad(nnl, ym, ut, R.PRNGKey(rep))
                torch.cuda.synchronize()
                if rep > 1:
                    tws.append((time.perf_counter() - t0) * 1e3)
            train_extra["ms_full_loss_value_and_grad_api"] = float(np.mean(tws))
            train_extra["full_loss_terms"] = {k: float(v) for k, v in info_l.items()}
        except Exception as ex:
            train_extra = {"error": repr(ex)}
    # ---- BASELINE configs[4] (first half): cartpole nSDE as an MBRL simulator, 2^20 start states x
    # horizon 30, one particle each; start states sharded over the ranks (no collective)
    sim_extra = {}
    try:
        PS, HS = (1 << 20) // world, 30
        pmc = configs.cartpole_model(horizon=HS, num_particles=1)
        mc = sde_models.CartPoleSDE(pmc)
        engc = Engine(mc)
        nnc = configs.random_params(mc, seed=0)
        pkc = PackedParams(mc, nnc, engc.desc)
        g = torch.Generator(device=dev).manual_seed(rank)
        obs = (torch.rand((PS, 4), device=dev, generator=g) * 2 - 1) * torch.tensor([2.0, 3.0, 3.14159, 6.0], device=dev)
        actc = torch.rand((PS, HS, 1), device=dev, generator=g) * 2 - 1
        keysc = torch.stack([torch.full((PS,), rank, dtype=torch.int32, device=dev),
                             torch.arange(PS, dtype=torch.int32, device=dev)], dim=1).contiguous()
        outc = torch.empty((HS + 1, 4, PS), device=dev)
        tsc = np.full(HS, 0.02, np.float32)
        tms = []
        for rep in range(4):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engc.rollout(pkc, obs, actc, tsc, key=keysc, N=1, out=outc)
            e1.record()
            torch.cuda.synchronize()
            if rep:
                tms.append(e0.elapsed_time(e1))
        tm = float(np.mean(tms))
        if world > 1:
            t = torch.tensor([tm], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tm = float(t.item())
        tot = PS * world * HS
        sim_extra = {"start_states": PS * world, "horizon": HS, "ms": tm, "particle_steps_per_s": tot / (tm * 1e-3),
                     "fp32_tflops": 3135.0 * tot / (tm * 1e-3) / 1e12,
                     "trajectory_writeout_gbs": 4.0 * PS * world * (HS + 1) * 4 / (tm * 1e-3) / 1e9}
        # the same rollouts as a simulator step (sde4_sim_rollout): only the last states leave the kernel
        # (16 MB instead of 520 MB), cartpole swing-up reward accumulated in the kernel, particle mean on the device
        tms = []
        for rep in range(4):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engc.sim_rollout(pkc, obs, actc, tsc, keysc, N=1, reward="cartpole_swingup")
            e1.record()
            torch.cuda.synchronize()
            if rep:
                tms.append(e0.elapsed_time(e1))
        tmp = float(np.mean(tms))
        if world > 1:
            t = torch.tensor([tmp], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tmp = float(t.item())
        sim_extra["predict_ms"] = tmp
        sim_extra["predict_particle_steps_per_s"] = tot / (tmp * 1e-3)
        sim_extra["predict_bytes_written"] = 4 * PS * (4 + 1 + 4 + 1)
        del outc, obs, actc
    except Exception as ex:
        sim_extra = {"error": repr(ex)}
    # ---- BASELINE configs[3]: batch of 4096 independent hexacopter MPC solves (H=20, dt=0.05, 20 APG
    # iterations, mpc_test_cfg.yaml), PARTITIONED over the ranks with no collective (strong scaling)
    mpc_extra = {}
    try:
        for n_part in (1, 32):
            mpc_extra["N%d" % n_part] = mpc_batch_run(w, n_part, world, rank, dev, barrier, reps=4)
        mpc_extra["scaling"] = "strong: 4096 problems partitioned over the GPUs, no collective (see --workload mpc_batch)"
    except Exception as ex:  # never lose the headline line because of the extra
        mpc_extra = {"error": repr(ex)}

    # ---- BASELINE configs[1]: cartpole APG-MPC, ONE problem, 1024 particles, horizon 50, 20 APG iterations
    # (latency of a single solve): (i) device-side APG, one CUDA-graph replay per iteration; (ii) the
    # reference-facing apg.py entry point driven with host arrays (every cost call = sde4_cost_grad_host)
    c2_extra = {}
    if rank == 0:
        try:
            import contextlib, io
            from sde4mbrl_b200 import apg as _apg, costs as _costs
            from sde4mbrl_b200.apg_batched import BatchedAPG
            HC, NC = 50, 1024
            pmc = configs.cartpole_model(horizon=HC, num_particles=NC)
            mpcc = configs.cartpole_mpc(horizon=HC, dt=0.02, num_particles=NC, max_iter=20)
            mpcc["apg_mpc"].update(atol=0.0, rtol=0.0, max_no_improvement_iter=20)  # exactly 20 iterations
            with contextlib.redirect_stdout(io.StringIO()):
                _, mcost, vprox, _, _ = nsde.create_online_cost_sampling_fn(pmc, mpcc, sde_constr=sde_models.CartPoleSDE)
            modc = mcost.engine.model
            nnc = configs.random_params(modc, seed=0)
            for k_, v_ in nnc.items():  # random_params damps the residual nets (x0.02) for long rollouts; an MPC problem
                if k_.endswith("linear_2") and "init_residual_networks" in k_:  # needs dynamics that react to the control
                    v_["w"] = v_["w"] * np.float32(25.0)
                    v_["b"] = v_["b"] * np.float32(25.0)
            cpc = mpcc["cost_params"]
            stagec = _costs.QuadraticCost(cpc["xerr"], cpc["uerr"], cpc["uref"], feat="cartpole", res_mult=cpc["res_mult"])
            y0c = np.array([0.0, 0.0, np.pi, 0.0], np.float32) + np.random.default_rng(3).normal(0, 0.01, 4).astype(np.float32)
            xrefc = np.zeros((HC + 1, 4), np.float32)
            keyc = np.array([0, 7], np.uint32)
            x0c = np.full((HC, 1), 1e-4, np.float32)
            solver = BatchedAPG(mcost.engine, nnc, stagec, mpcc["apg_mpc"], mcost.time_steps, NC, lb=-1.0, ub=1.0)
            tms = []
            for rep in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                solver.init(torch.as_tensor(x0c, device=dev), y0c[None], xrefc, keyc).run(20)
                e1.record()
                torch.cuda.synchronize()
                if rep:
                    tms.append(e0.elapsed_time(e1))
            dev_ms, dev_cost = float(np.mean(tms)), float(solver.t["opt_cost"][0])
            cost_h = lambda o: mcost(nnc, y0c, o, keyc, stagec, extra_cost_args=xrefc)[0]
            prox_h = lambda x, stepsize=None: vprox(x)
            tms = []
            for rep in range(4):
                t0 = time.perf_counter()
                st_h = _apg.apg(_apg.init_apg(torch.as_tensor(x0c), cost_h, mpcc["apg_mpc"]), cost_h, mpcc["apg_mpc"],
                                proximal_fn=prox_h)
                if rep:
                    tms.append((time.perf_counter() - t0) * 1e3)
            # the same reference-facing calls with the cost function made by multi_sampling.bind(...) instead of a lambda:
            # apg.apg hands the solve to the device-side solver
            cost_b = mcost.bind(nnc, y0c, keyc, stagec, extra_cost_args=xrefc)
            prox_b = lambda x, stepsize=None: vprox(x)
            prox_b.box = vprox
            tmb = []
            for rep in range(4):
                t0 = time.perf_counter()
                st_b = _apg.apg(_apg.init_apg(torch.as_tensor(x0c), cost_b, mpcc["apg_mpc"]), cost_b, mpcc["apg_mpc"],
                                proximal_fn=prox_b)
                if rep:
                    tmb.append((time.perf_counter() - t0) * 1e3)
            c2_extra = {"problem": "cartpole swing-up, 1 problem, %d particles, horizon %d, 20 APG iterations" % (NC, HC),
                        "apg_api_bound_cost_ms_per_solve": float(np.mean(tmb)), "apg_api_bound_cost_opt_cost": float(st_b.opt_cost),
                        "device_apg_ms_per_solve": dev_ms, "device_apg_opt_cost": dev_cost,
                        "host_apg_api_ms_per_solve": float(np.mean(tms)), "host_apg_opt_cost": float(st_h.opt_cost),
                        "host_apg_iterations": int(st_h.num_steps) - 1}
        except Exception as ex:
            c2_extra = {"error": repr(ex)}

    # ---- BASELINE configs[0]: mass-spring-damper nSDE, 200 particles x 100-step Euler-Maruyama rollout (the case the
    # reference runs on CPU via JAX), device time of one rollout call
    msd_extra = {}
    if rank == 0:
        try:
            pmm0 = configs.msd_model(horizon=100, num_particles=200)
            mm0 = sde_models.MassSpringDamper(pmm0)
            engm0 = Engine(mm0)
            pkm0 = PackedParams(mm0, configs.random_params(mm0, seed=0), engm0.desc)
            y0m0 = torch.tensor([[0.05, -0.02]], device=dev)
            um0 = torch.zeros((100, 1), device=dev)
            tsm0 = np.full(100, float(pmm0["stepsize"]), np.float32)
            km0 = torch.as_tensor(np.array([0, 1], np.uint32).view(np.int32), device=dev)
            outm0 = torch.empty((101, 2, 200), device=dev)
            tms = []
            for rep in range(13):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                engm0.rollout(pkm0, y0m0, um0, tsm0, key=km0, N=200, out=outm0)
                e1.record()
                torch.cuda.synchronize()
                if rep > 2:
                    tms.append(e0.elapsed_time(e1))
            msd_extra = {"particles": 200, "horizon": 100, "ms": float(np.mean(tms)),
                         "particle_steps_per_s": 200 * 100 / (float(np.mean(tms)) * 1e-3)}
        except Exception as ex:
            msd_extra = {"error": repr(ex)}

    # ---- latency of ONE hexacopter MPC solve as the reference deploys it (mpc_test_cfg.yaml: H=20, dt=0.05, 1 particle;
    # also 32 and 1024 particles), 20 APG iterations, device-side APG (one CUDA-graph replay per iteration)
    single_extra = {}
    if rank == 0:
        try:
            from sde4mbrl_b200.apg_batched import BatchedAPG
            HM1 = 20
            for n_part in (1, 32, 1024):
                pm1 = configs.hexa_model(horizon=HM1, num_particles=n_part, stepsize=0.05)
                mpc1 = configs.hexa_mpc(horizon=HM1, dt=0.05, num_particles=n_part, max_iter=20)
                eng1 = Engine(sde_models.SDERotorModel(pm1))
                sol1 = BatchedAPG(eng1, w["nn"], w["stage"], mpc1["apg_mpc"], np.full(HM1, 0.05, np.float32), n_part,
                                  lb=1e-4, ub=1.0)
                x01 = torch.full((HM1, 6), 0.33 + 2e-4, device=dev)
                xr1 = np.tile(w["xref"][:1], (HM1 + 1, 1))
                tms = []
                for rep in range(6):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    sol1.init(x01, w["y0"][None], xr1, w["key"]).run(20)
                    e1.record()
                    torch.cuda.synchronize()
                    if rep > 1:
                        tms.append(e0.elapsed_time(e1))
                single_extra["N%d" % n_part] = {"ms_per_solve": float(np.mean(tms)), "particles": n_part, "horizon": HM1,
                                                "apg_iterations": 20, "opt_cost": float(sol1.t["opt_cost"][0]),
                                                "init_cost": float(sol1.t["init_cost"][0])}
        except Exception as ex:
            single_extra = {"error": repr(ex)}

    # ---- extras (rank 0, N=1 only): forward rollout, FP32 peak, CPU baseline
    extras = {}
    roof = None
    cpu_b = None
    if rank == 0:
        # measured FP32 FMA peak (roofline denominator; MEASURED_PEAKS.json has no FP32 number)
        sink = torch.zeros(4, device=dev)
        peaks = {}
        for variant in (0, 1):
            best = 0.0
            for _ in range(3):
                fl = L.C.c_double(0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                L.check(L.load().sde4_fma_peak(variant, torch.cuda.get_device_properties(dev).multi_processor_count * 8, 256, 4000, L.C.c_void_p(sink.data_ptr()),
                                               L.C.byref(fl), L.C.c_void_p(torch.cuda.current_stream().cuda_stream)))
                e1.record()
                torch.cuda.synchronize()
                best = max(best, fl.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
            peaks["ffma" if variant == 0 else "ffma2_f32x2"] = best
        peak = max(peaks.values())
        # the kernel alone (K2; K3 is ~2 us): same launch, events on the launch stream, no collective
        kt = []
        for _ in range(min(args.steps, 20)):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.cost_grad(pk, y0_d, opt_d, ts, xref_d, w["stage"].desc, key=key_d, N=N_PART, want_grad=True,
                          block_threads=args.block_threads, particle_offset=rank * N_PART, particle_total=n_total,
                          lanes=args.lanes)
            e1.record()
            torch.cuda.synchronize()
            kt.append(e0.elapsed_time(e1))
        k_ms = float(np.mean(kt))
        flop = F_VALGRAD_HEXA * N_PART * HORIZON
        achieved = flop / (k_ms * 1e-3) / 1e12
        # the same launch against the HBM number of MEASURED_PEAKS.json (driver-written): algorithmic bytes = the
        # checkpoint / noise write-out; shows why "hbm" is not the bound of this path
        hbm_view = None
        try:
            mp = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))
            wbytes = 4.0 * N_PART * (HORIZON * (13 + 13 + 1) + 13)
            gbs = wbytes / (k_ms * 1e-3) / 1e9
            hbm_view = {"bound": "hbm", "achieved": gbs, "peak": mp["hbm_gbs"], "unit": "GB/s", "frac": gbs / mp["hbm_gbs"],
                        "algorithmic_bytes_per_launch": wbytes, "peak_source": "MEASURED_PEAKS.json hbm_gbs"}
        except Exception:
            pass
        roof = {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "hbm_view": hbm_view,
                "traffic": 3342080, "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one k_cost_ws launch, profiles/r2_h_ws_final_ncu.txt (checkpoints stay in L2)", "peak_source": "measured in this run by sde4_fma_peak (FFMA / fma.rn.f32x2 loop)",
                "peaks_tflops": peaks, "kernel_ms": k_ms,
                "algorithmic_flop_per_launch": flop,
                "note": "path is FP32-FMA bound (SURVEY 8d), not HBM/tensor; HBM write-out of this launch = %.1f MB"
                        % (4.0 * N_PART * (HORIZON + 1) * 13 * 2 / 1e6)}
        # forward-only rollout (K1) for context
        xout = torch.empty((HORIZON + 1, 13, N_PART), device=dev)
        ft = []
        for rep in range(13):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.rollout(pk, y0_d, opt_d, ts, key=key_d, N=N_PART, lanes=args.lanes, out=xout)
            e1.record()
            torch.cuda.synchronize()
            if rep >= 3:
                ft.append(e0.elapsed_time(e1))
        extras["forward_rollout_ms"] = float(np.mean(ft))
        extras["forward_particle_steps_per_s"] = N_PART * HORIZON / (float(np.mean(ft)) * 1e-3)
        extras["forward_fp32_tflops"] = F_FWD_HEXA * N_PART * HORIZON / (float(np.mean(ft)) * 1e-3) / 1e12
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n_sample = N_PART
            for _ in range(3):
                cpu_reference_step(w, n_sample, threads)
            tcpu = [cpu_reference_step(w, n_sample, threads)[0] for _ in range(20)]
            v = n_sample * HORIZON / float(np.mean(tcpu))
            cpu_b = {"value": v, "unit": "particle-steps/s", "cores": threads, "kind": "port",
                     "sample": "all %d particles x %d steps, value+grad, 20 reps; oracle/c (C restatement, AVX-512/AVX2 lanes + OpenMP)" % (n_sample, HORIZON)}

    if rank == 0:
        total_ps = n_total * HORIZON
        line = {
            "metric": "particle-steps/s (value+grad, hexacopter nSDE)", "value": total_ps / (ms_per_step * 1e-3),
            "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": bench_config(world),
            "collective": "none (1 GPU)" if world == 1 else reducer.transport,
            "l2_flush": {"write": "256 MB write between timed steps", "write+read": "DIAGNOSTIC: 256 MB write then 256 MB read between timed steps", "none": "DIAGNOSTIC: not flushed"}[args.flush_mode],
            "clocks": clocks,
            "e2e": {"value": total_ps / e2e_s, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3,
                    "api": "nsde.create_online_cost_sampling_fn(...).multi_sampling.value_and_grad on HOST numpy arrays "
                           "(sde4_cost_grad_host: pinned staging, one H2D + one D2H copy, stream sync inside)"
                           if world == 1 else
                           "nsde.create_online_cost_sampling_fn(...).multi_sampling.value_and_grad on HOST numpy arrays with "
                           "multi_sampling.particle_reducer = dist.CostGradReducer, transport '%s' (fused: sde4_cost_grad_host with "
                           "sde4_xgpu_reduce -- one packed pinned H2D copy, K2, K3 whose last CTA sums the peers' partials over "
                           "NVLink and writes the total into the mapped staging buffer, stream sync; other transports: K3 + "
                           "library all-reduce + D2H copy)" % args.transport},
            "gpu_launches": 2 * args.steps,
            "wall_s_timed_region": t_wall,
        }
        if roof:
            line["roofline"] = roof
        if cpu_b:
            line["cpu_baseline"] = cpu_b
        line.update(extras)
        line["msd_rollout"] = msd_extra
        line["cartpole_mpc_single"] = c2_extra
        line["hexa_mpc_single"] = single_extra
        line["mpc_batch"] = mpc_extra
        line["simulator_rollouts"] = sim_extra
        line["training_loss"] = train_extra
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
