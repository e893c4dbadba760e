#!/usr/bin/env python
"""bench.py -- CDPR instance-steps/s of the batched hot path on N B200s (weak scaling), one JSON line.

  python bench.py --gpus N --steps K --warmup W            own arm (CUDA path through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  reference arm: the reference's own force-law
                                                           code (oracle/_ref) in the reduced model on the host cores

A bench "step" is one pass of the hot path over one batch: `instances` independent robots per GPU, each
advanced `sim_steps` physics steps by ONE persistent kernel launch (BASELINE.json configs[2]: 2^20 instances x
1000 steps under per-instance sine velocity commands, fp64; NC = 8 is the north_star's synthetic 8-cable
extension of the 4-cable reference robot).  For N > 1 every rank owns its own 2^20 instances (configs[3]) and the
decimated trajectory (a snapshot every 100 steps) is gathered to every rank inside the timed region; after the timed
regions the gathered trajectory is CHECKED (bitwise against an NCCL all-gather of the same pass and against a fresh
single-GPU recompute of a sample of every rank's columns) and configs[4] -- the sampled-rollout batch with its
cross-GPU cost all-reduce -- is run, timed and checked.  Both verdicts are part of the JSON line.
"""
from __future__ import annotations

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

from cdpr_simulation_b200.flops import frozen_flops_per_instance_step, frozen_ik_flops_per_pose, ik_bytes_per_pose  # noqa: E402

METRIC = "CDPR instance-steps/sec"
UNIT = "instance-steps/s"


def flops_per_instance_step(nc: int) -> int:
    """The roofline numerator: FROZEN algorithmic count (cdpr_simulation_b200/flops.py; 1028 @ NC=8, 636 @ NC=4)."""
    return frozen_flops_per_instance_step(nc, "general")


def executed_flops_per_instance_step(nc: int):
    """What the built kernel's hot loop executes per step (SASS count; FMA = 2).  None when cuobjdump is unavailable."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import hot_loop_flops
        return int(hot_loop_flops.count()[nc]["executed_flops"])
    except Exception:
        try:
            return int(json.load(open(os.path.join(ROOT, "profiles", "hot_loop_flops.json")))[str(nc)]["executed_flops"])
        except Exception:
            return None


def state_bytes_per_instance(nc: int) -> int:
    # what the persistent kernel reads + writes per instance and LAUNCH: in: platform 13, per cable
    # i_err, target, window 11, moments 3, ctl (4 B), sine 3; out: platform 13, per cable i_err, last_time, window 11,
    # moments 3, 10 telemetry columns, vel_target, ctl (read-modify-write: 8 B)
    return 8 * (13 + 16 * nc + 3) + 4 * nc + 8 * (13 + 27 * nc) + 8 * nc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def make_oracle_config(nc: int):
    from oracle import binding as ob
    return ob.default_config(nc)


def cpu_arm(kind: str, nc: int, sim_steps: int, target_seconds: float, passes: int, cores: int):
    """Times the CPU oracle (kind 'port') or the reference's own force law inside the reduced model (kind
    'reference') on a bounded sample of the same workload.  Returns (instance-steps/s, sample description)."""
    from oracle import binding as ob
    from cdpr_simulation_b200 import workloads as wl
    cfg = make_oracle_config(nc)

    def run(n):
        amp, freq, phase, pose7, twist6 = wl.c3_instances(n, seed=1)
        b = ob.Batch(cfg, n, pose7, twist6, amp, freq, phase)
        t0 = time.perf_counter()
        if kind == "reference":
            b.step_reference_forcelaw(sim_steps, cores)
        else:
            b.step(sim_steps, cores)
        return time.perf_counter() - t0

    n0 = 8 * cores
    t = run(n0)
    rate = n0 * sim_steps / t
    n = int(max(cores, min(1 << 20, rate * target_seconds / sim_steps)))
    n = (n + cores - 1) // cores * cores
    times = [run(n) for _ in range(max(1, passes))]
    best = n * sim_steps / float(np.mean(times))
    return best, f"{n} instances x {sim_steps} steps (same generator and seed as the GPU batch, first {n} instances), " \
                 f"{len(times)} pass(es), {cores} OpenMP threads", float(np.mean(times)) * 1e3


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as ob
    ob.build()
    kind = "reference" if ob.ref_available() else "port"
    cores = host_cores()
    # K timed passes of a bounded sample; the sample shrinks with K so that the whole run stays around 1.5 minutes
    per_pass = max(0.5, min(args.ref_seconds, 90.0 / max(1, args.steps)))
    value, sample, ms = cpu_arm(kind, args.nc, args.sim_steps, per_pass, args.steps, cores)
    law = ("the reference's own Pid.cpp/JointForceCalculator.cpp, compiled unmodified (oracle/_ref), as the force law" if kind == "reference"
           else "FALLBACK: oracle/_ref is not built on this box, so the force law is the repository's C restatement (oracle port), not the reference's code")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C3 sample: sine velocity commands, NC={args.nc}, {args.sim_steps} physics steps per pass, "
                               f"reduced model with {law}, host cores only"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "reference_code_ran": kind == "reference",
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# multi-GPU correctness, inside the bench (driver-visible): configs[3] gather and configs[4] rollouts + all-reduce
# ------------------------------------------------------------------------------------------------------------------
def gather_check(cb, wl, D, torch, dist, batch, gather, load_inputs, args, rank, world, local_rank, stream):
    """After the timed region: one more pass through the FUSED gather, then the same pass again with local snapshots and
    an NCCL all_gather; the two trajectories must be bitwise equal on every rank.  Rank 0 also recomputes, on a fresh
    single-GPU handle, a sample of the instances of EVERY rank and compares those columns bitwise (a wrong column
    offset in the fused stores cannot hide behind two equally wrong gathers)."""
    n, k_sim, every = args.instances, args.sim_steps, args.snapshot_every
    n_snap = k_sim // every
    with torch.cuda.stream(stream):
        batch.reset(); load_inputs()
        gather.before_pass(); batch.step(k_sim); gather.after_pass(); gather.finish()
        stream.synchronize()
        if hasattr(gather, "latest"):
            fused = gather.latest().clone()                                   # [n_snap][13][world * n]
        else:
            fused = D.global_trajectory_to_instance_major(gather.recv).clone()
        local = torch.zeros((n_snap, 13, n), dtype=torch.float64, device=f"cuda:{local_rank}")
        torch.cuda.synchronize()
        batch.reset(); load_inputs()
        batch.set_snapshots(every, local.data_ptr(), n_snap)
        batch.step(k_sim); batch.synchronize()
        batch.set_snapshots(0, None, 0)
    gathered = D.gather_trajectory(local)
    traj = D.global_trajectory_to_instance_major(gathered)
    equal = bool(torch.equal(fused, traj))
    # rank 0: fresh single-GPU recompute of the first `m` instances of every rank
    m = min(n, 2048)
    sample_ok = True
    if rank == 0:
        cfg = cb.default_config(args.nc)
        for r in range(world):
            amp, freq, phase, pose7, twist6 = wl.c3_instances(n, seed=1 + 1000 * r)
            with cb.CdprBatch(cfg, m, device=local_rank) as g:
                snaps = torch.zeros((n_snap, 13, m), dtype=torch.float64, device=f"cuda:{local_rank}")
                torch.cuda.synchronize()
                g.set_platform_state(pose7[:m], twist6[:m]); g.set_sine_cmd(amp[:m], freq[:m], phase[:m])
                g.set_snapshots(every, snaps.data_ptr(), n_snap)
                g.step(k_sim); g.synchronize()
            sample_ok = sample_ok and bool(torch.equal(snaps, fused[:, :, r * n: r * n + m]))
    flags = torch.tensor([1.0 if equal else 0.0, 1.0 if sample_ok else 0.0], dtype=torch.float64, device=f"cuda:{local_rank}")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    del fused, traj, gathered, local
    return {"ranks": world, "bitwise_equal": bool(flags[0] > 0.5), "sample_recompute_equal": bool(flags[1] > 0.5),
            "snapshots": n_snap, "instances_total": world * n, "sample_instances_per_rank": m,
            "what": "fused gather buffer of every rank == NCCL all_gather of the same pass (bitwise); rank 0: columns of every "
                    "rank == fresh single-GPU recompute of that rank's first instances (bitwise)"}


def rollouts_c5(cb, wl, D, torch, dist, args, rank, world, local_rank):
    """configs[4]: 4096 command sequences x 256 steps per robot, robots sharded across the ranks (64 per GPU), in-kernel
    tracking cost, per-sequence sums over this rank's robots, cost vector all-reduced across the GPUs."""
    n_seq, n_cmd, spc, n_rob = 4096, 32, 8, args.rollout_robots          # 32 commands x 8 steps = 256 steps
    nc = args.nc
    cmds = wl.c5_rollouts(n_seq, n_cmd, nc)
    _, _, _, rp_all, rt_all = wl.c3_instances(n_rob * world, seed=5)
    lo, hi = rank * n_rob, (rank + 1) * n_rob                                 # contiguous robot range of this rank
    dev = f"cuda:{local_rank}"
    out = {}
    with cb.CdprBatch(cb.default_config(nc), n_rob * n_seq, device=local_rank) as g:
        cost = torch.zeros(n_seq, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        ms = []
        for _ in range(4):
            g.rollout(n_rob, n_seq, cmds, spc, [0.0, 0.0, 0.32], 0.05, rp_all[lo:hi], rt_all[lo:hi], dev_cost_seq=cost.data_ptr(), want_host_cost=False)
            g.synchronize()
            ms.append(g.last_kernel_ms)
        kernel_ms = float(np.mean(ms[1:]))
        partial = cost.clone()
        allreduce_us, check = None, None
        if world > 1:
            torch.cuda.synchronize(); dist.barrier()
            work = partial.clone()
            for _ in range(5):                                                # warm the communicator for this size
                D.allreduce_cost(work)
            reps = 50
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); dist.barrier()
            e0.record()
            for _ in range(reps):
                D.allreduce_cost(work)
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            allreduce_us = float(t[0])
            total = D.allreduce_cost(partial.clone())
            # check: gather every rank's partial vector and add them up in rank order
            parts = [torch.empty_like(partial) for _ in range(world)]
            dist.all_gather(parts, partial)
            expect = torch.zeros_like(partial)
            for p_ in parts:
                expect += p_
            rel = float(((total - expect).abs() / expect.abs().clamp_min(1e-300)).max())
            # and rank 0 re-runs the robots of the LAST rank on its own GPU: that partial vector must match bitwise
            same = True
            if rank == 0:
                l2, h2 = (world - 1) * n_rob, world * n_rob
                c2 = torch.zeros(n_seq, dtype=torch.float64, device=dev)
                torch.cuda.synchronize()
                g.rollout(n_rob, n_seq, cmds, spc, [0.0, 0.0, 0.32], 0.05, rp_all[l2:h2], rt_all[l2:h2], dev_cost_seq=c2.data_ptr(), want_host_cost=False)
                g.synchronize()
                same = bool(torch.equal(c2, parts[world - 1]))
            f = torch.tensor([1.0 if rel < 1e-13 else 0.0, 1.0 if same else 0.0], dtype=torch.float64, device=dev)
            dist.all_reduce(f, op=dist.ReduceOp.MIN)
            check = {"allreduce_vs_rank_ordered_sum_max_rel": rel, "allreduce_ok": bool(f[0] > 0.5),
                     "last_ranks_partial_recomputed_on_rank0_bitwise": bool(f[1] > 0.5)}
        t = torch.tensor([kernel_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kernel_ms = float(t[0])
        steps = n_cmd * spc
        out = {"value": world * n_rob * n_seq * steps / ((kernel_ms + (allreduce_us or 0.0) * 1e-3) * 1e-3), "unit": UNIT,
               "kernel_ms": kernel_ms, "allreduce_us": allreduce_us, "ranks": world,
               "robots_per_gpu": n_rob, "sequences": n_seq, "steps": steps, "n_cables": nc, "check": check,
               "what": f"{world} x {n_rob} robots x {n_seq} sequences x {steps} steps; cost vector of {n_seq} float64 all-reduced (NCCL) across the ranks; "
                       "value = rollout steps / (kernel + all-reduce time, max over ranks)"}
    return out


def bind_to_gpu_numa_node(gpu_index: int):
    """Pin this rank to the CPU cores next to its GPU (sysfs local_cpulist of the GPU's PCI device) BEFORE any pinned host
    memory is allocated, so the pinned buffers land on that NUMA node: with several ranks copying gigabytes per pass the
    host side of the D2H copies is the bound, and cross-socket traffic halves it.  (On the pool's B200 boxes the guest sees ONE
    NUMA node, so this changes nothing there -- measured at 4 GPUs: 85.4 vs 86.7 ms per pass; it matters on bare metal.)
    Returns a description for the JSON line."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return "unbound (no PCI bus id)"
        dev = "/sys/bus/pci/devices/" + bus[-12:]                      # 00000000:1b:00.0 -> 0000:1b:00.0
        cpus = set()
        for part in open(dev + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return "unbound (no local cores in this process's affinity mask)"
        os.sched_setaffinity(0, allowed)
        node = open(dev + "/numa_node").read().strip()
        return f"NUMA node {node}, {len(allowed)} cores"
    except Exception as e:
        return f"unbound ({type(e).__name__})"


def own_arm(args):
    import torch
    import torch.distributed as dist
    import cdpr_simulation_b200 as cb
    from cdpr_simulation_b200 import workloads as wl
    from cdpr_simulation_b200 import distributed as D
    from cdpr_simulation_b200.distributed import make_trajectory_gather

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if (world > 1 and not args.no_numa_bind) else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    nc, n, k_sim = args.nc, args.instances, args.sim_steps
    cfg = cb.default_config(nc)
    # rank-specific slice of the C3/C4 generator (global instance id = rank * n + local id)
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n, seed=1 + 1000 * rank)

    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True)
        t.numpy()[...] = a
        return t
    pin_in = [pinned(a) for a in (amp, freq, phase, pose7, twist6)]
    out_shapes = ((n, 7), (n, 6), (n, nc), (n, nc), (n, nc))
    pin_out = [torch.empty(s, dtype=torch.float64, pin_memory=True) for s in out_shapes]

    stream = torch.cuda.Stream()
    batch = cb.CdprBatch(cfg, n, device=local_rank)
    batch.set_stream(stream.cuda_stream)
    assert batch.kernel_variant == "fast"
    gather, gather_kind = (make_trajectory_gather(batch, args.snapshot_every, k_sim, stream, prefer_fused=(args.gather == "fused"))
                           if world > 1 else (None, None))

    def load_inputs():
        batch.set_platform_state(pin_in[3].numpy(), pin_in[4].numpy())
        batch.set_sine_cmd(pin_in[0].numpy(), pin_in[1].numpy(), pin_in[2].numpy())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = cb.measure_fp64_tflops(local_rank, 8192)

    # ---------------- value: inputs resident in HBM, device-timed ----------------
    with torch.cuda.stream(stream):
        load_inputs()
        for _ in range(args.warmup):
            if gather: gather.before_pass()
            batch.step(k_sim)
            if gather: gather.after_pass()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = batch.launch_count
        kernel_ms = []
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            if gather: gather.before_pass()
            batch.step(k_sim)
            if gather: gather.after_pass()
            kernel_ms.append(batch.last_kernel_ms)   # CUDA events recorded by the library around its launch
        if gather: gather.finish()
        ev1.record(stream)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms_total = ev0.elapsed_time(ev1)
        launches = batch.launch_count - launches0

        # ---------------- e2e: host buffers in, host buffers out, through the public API ----------------
        # Two handles (A/B) on two streams, each with its own pinned buffers: while pass k computes on one, the D2H of
        # pass k-1 and the H2D of pass k+1 run on the other (calls only enqueue: cdpr_set_async). Every pass still does
        # its own reset + H2D of all inputs + step + D2H of all outputs.  N > 1: every pass also runs the trajectory
        # gather, and each rank delivers its share of the GATHERED trajectory to the host -- all snapshots of the columns of
        # shard (rank + 1) mod N, read from its own copy of the gather buffer -- so the host receives the full decimated
        # trajectory of all N x 2^20 instances once per pass, spread evenly over the N PCIe links.
        n_snap = k_sim // args.snapshot_every
        src_shard = (rank + 1) % world          # the columns this rank delivers to the host come from its NEIGHBOUR's shard,
        col0 = src_shard * n                    # i.e. they exist on this GPU only because the gather put them there
        lanes = []
        for lane in range(2):
            st = stream if lane == 0 else torch.cuda.Stream()
            bt = batch if lane == 0 else cb.CdprBatch(cfg, n, device=local_rank)
            bt.set_stream(st.cuda_stream)
            bt.set_async(True)
            ins = pin_in if lane == 0 else [pinned(a) for a in (amp, freq, phase, pose7, twist6)]
            outs = pin_out if lane == 0 else [torch.empty(s, dtype=torch.float64, pin_memory=True) for s in out_shapes]
            gl = None
            if world > 1:
                gl = gather if lane == 0 else make_trajectory_gather(bt, args.snapshot_every, k_sim, st, prefer_fused=(args.gather == "fused"))[0]
            traj_host = torch.empty((n_snap, 13, n), dtype=torch.float64, pin_memory=True) if world > 1 else None
            traj_stage = None
            lanes.append((bt, ins, outs, gl, st, traj_host, traj_stage))
        if gather: gather.finish()

        def e2e_pass(k):
            bt, ins, outs, gl, st, traj_host, traj_stage = lanes[k % 2]
            bt.synchronize()                                   # the pinned buffers of this lane are free again
            if gl is not None:
                st.synchronize()
            bt.reset()
            bt.set_platform_state(ins[3].numpy(), ins[4].numpy())            # H2D
            bt.set_sine_cmd(ins[0].numpy(), ins[1].numpy(), ins[2].numpy())
            if gl is not None: gl.before_pass()
            bt.step(k_sim)
            if gl is not None:
                gl.after_pass(); gl.finish()
                src = gl.latest() if hasattr(gl, "latest") else None
                with torch.cuda.stream(st):                                 # D2H of this rank's share of the gathered trajectory
                    # the shard's columns are strided in the gather buffer, but every (snapshot, field) row of them is one
                    # contiguous 8 MB run: 130 plain D2H copies on the copy engine.  (A packing kernel, or torch's strided
                    # copy_, needs SMs and would queue behind the OTHER lane's step kernel, which holds every SM for the
                    # whole pass -- that serialises the two lanes.)
                    if src is not None:
                        for s_ in range(n_snap):
                            for f_ in range(13):
                                traj_host[s_, f_].copy_(src[s_, f_, col0:col0 + n], non_blocking=True)
                    else:
                        traj_host.copy_(gl.recv[src_shard], non_blocking=True)
            bt.platform_state((outs[0].numpy(), outs[1].numpy()))            # D2H
            bt.joint_states(tuple(t.numpy() for t in outs[2:]))

        def lanes_sync():
            for bt, _, _, _, st, _, _ in lanes:
                bt.synchronize(); st.synchronize()

        for k in range(2):
            e2e_pass(k)                                        # warm both lanes
        lanes_sync()
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            e2e_pass(k)
        lanes_sync()
        barrier()
        e2e_s = time.perf_counter() - t0
        for bt, _, _, _, _, _, _ in lanes:
            bt.set_async(False)
        traj_bytes = sum(t.numel() * t.element_size() for t in [lanes[0][5]] if t is not None)
        if lanes[1][0] is not batch:
            lanes[1][0].close()

    # ---------------- multi-GPU correctness + configs[4], after the timed regions ----------------
    gcheck = None
    if world > 1:
        gcheck = gather_check(cb, wl, D, torch, dist, batch, gather, load_inputs, args, rank, world, local_rank, stream)
    c5 = None
    if not args.no_rollouts:
        c5 = rollouts_c5(cb, wl, D, torch, dist, args, rank, world, local_rank)

    t = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    total_units = float(world) * n * k_sim * args.steps
    value = total_units / (ms_total * 1e-3)
    e2e_value = total_units / (e2e_ms * 1e-3)
    h2d = sum(x.numel() * x.element_size() for x in pin_in)
    d2h = sum(x.numel() * x.element_size() for x in pin_out) + traj_bytes

    if rank == 0:
        kms = float(np.mean(kernel_ms))
        frozen = flops_per_instance_step(nc)
        executed = executed_flops_per_instance_step(nc)
        achieved = frozen * float(n) * k_sim / (kms * 1e-3) / 1e12
        achieved_exec = executed * float(n) * k_sim / (kms * 1e-3) / 1e12 if executed else None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_bytes = state_bytes_per_instance(nc) * float(n)
        traffic, traffic_src = None, None
        for prof_name in ("r2_ncu_summary.json", "r1_ncu_summary.json"):
            try:   # DRAM bytes of this kernel from the committed `ncu --set full` capture of the same launch shape
                prof = json.load(open(os.path.join(ROOT, "profiles", prof_name)))[f"step_fast_nc{nc}"]
                if n == (1 << 20) and k_sim == 1000:
                    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                    traffic = sum(float(prof[m]["value"]) * scale[prof[m]["unit"]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                    traffic_src = f"profiles/{prof_name} (ncu --set full, one launch of 2^20 instances x 1000 steps)"
                    break
            except Exception:
                pass
        fp64_src = ("MEASURED_PEAKS.json" if "fp64_tflops" in peaks else
                    "DFMA issue rate measured in this run by cdpr_measure_fp64_tflops (MEASURED_PEAKS.json has no FP64 entry; nominal B200 FP64 = "
                    "148 SM x 64 lanes x 2 x 1.965 GHz = 37.2 TFLOP/s; tools/ubench_dfma.cu output with clocks: profiles/r2_fp64_peak.txt)")
        fp64_den = peaks.get("fp64_tflops", fp64_peak)
        roofline = {
            "bound": "fp64", "achieved": achieved, "peak": fp64_den, "unit": "TFLOP/s", "frac": achieved / fp64_den if fp64_den > 0 else None,
            "traffic": traffic, "traffic_unit": "bytes per launch", "traffic_source": traffic_src, "algorithmic_bytes_per_launch": hbm_bytes,
            "kernel": f"k_step_fast<{nc},11,VELOCITY,moments,spec>", "kernel_ms": kms,
            "flops_per_instance_step": frozen,
            "flops_note": "FROZEN algorithmic count (cdpr_simulation_b200/flops.py = BASELINE.md section 4 / SURVEY App. D: 244 + 98 NC, general inertia, "
                          "FIR form of the D-term); it does not follow the kernel",
            "executed_flops": executed,
            "executed_flops_note": "FP64 work the built kernel's hot loop really issues per instance-step (SASS count, FMA = 2; tools/hot_loop_flops.py); "
                                   "smaller than the frozen count because the kernel is specialised on this robot (isotropic inertia, anchors in the "
                                   "platform plane) and carries the D-term as a 3-value recursion",
            "achieved_executed": achieved_exec, "frac_executed": (achieved_exec / fp64_den) if (achieved_exec and fp64_den > 0) else None,
            "peak_source": fp64_src, "peak_measured_in_run": fp64_peak,
            "hbm": {"achieved": hbm_bytes / (kms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": hbm_bytes / (kms * 1e-3) / 1e9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                    "note": "state is read and written once per launch of 1000 steps: the kernel is FP64-issue bound, not HBM bound"},
        }
        extra, ik_c2 = None, None
        if world == 1 and not args.no_extras:
            extra, ik_c2 = side_measurements(cb, wl, torch, local_rank, hbm_peak, fp64_den)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, sample, _ = cpu_arm("port", nc, k_sim, args.cpu_seconds, 1, host_cores())
            cpu = {"value": v, "unit": UNIT, "cores": host_cores(), "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C3: {n} instances/GPU x {k_sim} physics steps per pass, per-instance sine velocity commands, "
                                   f"NC={nc} ({'synthetic 8-cable extension' if nc == 8 else 'reference 4-cable robot'})",
                       "instances_per_gpu": n, "sim_steps_per_pass": k_sim, "n_cables": nc,
                       "l2": "resident state per GPU (%.1f GB) is larger than L2; no flush needed" % (batch_state_gb(batch)),
                       "multi_gpu": None if world == 1 else f"C4: snapshot every {args.snapshot_every} steps gathered to every rank -- {gather_kind}",
                       "host_binding": numa},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps,
                    "what": "per pass and rank: reset + H2D(pose, twist, sine params) + step + D2H(platform pose/twist, joint states) through the C ABI, "
                            "pinned host buffers, two handles double-buffered so copies overlap the other handle's kernel"
                            + ("" if world == 1 else f"; plus the trajectory gather and the D2H of this rank's share of the gathered trajectory "
                                                     f"({n_snap} snapshots x 13 x the {n} columns of shard (rank + 1) mod {world}, read from this rank's gather buffer)")},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "gather_check": gcheck,
            "rollouts_c5": c5,
            "ik_c2": ik_c2,
            "extra": extra,
        }
        print(json.dumps(line), flush=True)
    batch.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ik_c2_measurement(cb, wl, torch, device, hbm_peak, nc, npose):
    """configs[1]: the kinematics sweep at its stated size, steady state through a CUDA graph (no launch gaps), with a
    cudaMemcpyAsync device-to-device copy moving the SAME number of DRAM bytes timed the same way as the ceiling."""
    p7, t6 = wl.c2_poses(npose, seed=0)
    st = np.ascontiguousarray(np.concatenate([p7[:, :3], p7[:, 6:7], p7[:, 3:6], t6], axis=1).T)
    nbytes = ik_bytes_per_pose(nc) * npose
    # enough rotating buffer sets that consecutive launches cannot be served from the 126 MB L2
    sets = max(2, int(np.ceil(3 * 126e6 / nbytes)))
    dev = f"cuda:{device}"
    d_in = [torch.from_numpy(st).to(dev) for _ in range(sets)]
    d_out = [torch.empty((nc, 8, npose), dtype=torch.float64, device=dev) for _ in range(sets)]
    # copy ceiling: a D2D copy of nbytes/2 reads nbytes/2 and writes nbytes/2 = the sweep's DRAM traffic
    half = (nbytes // 2 + 255) // 256 * 256
    c_src = [torch.empty(half, dtype=torch.uint8, device=dev) for _ in range(sets)]
    c_dst = [torch.empty(half, dtype=torch.uint8, device=dev) for _ in range(sets)]
    torch.cuda.synchronize()
    res = {}
    with cb.CdprBatch(cb.default_config(nc), 1, device=device) as g:
        single = []
        for k in range(12):
            g.ik_device(npose, d_in[k % sets].data_ptr(), d_out[k % sets].data_ptr()); single.append(g.last_kernel_ms)
        s = torch.cuda.Stream(device=device)
        g.set_stream(s.cuda_stream)
        g.set_option(cb.api.OPT_KERNEL_TIMING, 0)          # no event records: the launches go into a CUDA graph
        reps = 10
        def timed(fn_capture):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(s):
                fn_capture()                                 # warm
                s.synchronize()
                with torch.cuda.graph(graph, stream=s):
                    fn_capture()
                graph.replay(); s.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                for _ in range(reps):
                    graph.replay()
                e1.record(s)
                s.synchronize()
            return e0.elapsed_time(e1) / (reps * sets)
        def sweep():
            for k in range(sets):
                g.ik_device(npose, d_in[k].data_ptr(), d_out[k].data_ptr())
        def copies():
            for k in range(sets):
                c_dst[k].copy_(c_src[k], non_blocking=True)
        try:
            steady = timed(sweep)
            how = f"CUDA graph of {sets} launches over {sets} rotating buffer sets (> L2 in total), {reps} replays"
        except Exception as e:   # graph capture unavailable: plain back-to-back launches
            how = f"{reps * sets} back-to-back launches (graph capture failed: {type(e).__name__})"
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(s):
                sweep(); e0.record(s)
                for _ in range(reps):
                    sweep()
                e1.record(s); s.synchronize()
            steady = e0.elapsed_time(e1) / (reps * sets)
        try:
            copy_ms = timed(copies)
        except Exception:
            copy_ms = None
        g.set_option(cb.api.OPT_KERNEL_TIMING, 1)
    gbs = nbytes / (steady * 1e-3) / 1e9
    copy_gbs = (2 * half) / (copy_ms * 1e-3) / 1e9 if copy_ms else None
    res = {"poses": npose, "n_cables": nc, "poses_per_s": npose / (steady * 1e-3), "kernel_us_steady": steady * 1e3,
           "kernel_us_single_shot": float(np.median(single[2:])) * 1e3, "how": how,
           "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": None,
                        "algorithmic_bytes_per_launch": nbytes, "flops_per_pose": frozen_ik_flops_per_pose(nc),
                        "kernel": "k_ik_pair (one thread per cable and pose pair)" if (npose <= (1 << 19) and npose % 2 == 0) else "k_ik (one thread per pose)"},
           "copy_ceiling": {"us": copy_ms * 1e3 if copy_ms else None, "GBps": copy_gbs, "bytes_read_plus_written": 2 * half,
                            "what": "cudaMemcpyAsync D2D moving the same DRAM bytes (read + write), same graph, same rotation"},
           "frac_of_copy_ceiling": (gbs / copy_gbs) if copy_gbs else None}
    return res


def side_measurements(cb, wl, torch, device, hbm_peak, fp64_peak):
    """Not the headline: the reference's own 4-cable robot on the same workload, the other kernel shapes, plugin-style
    stepping, the hold / filter variant, and the config-2 kinematics sweep (returned separately as `ik_c2`)."""
    out = {}
    n, k = 1 << 20, 1000
    amp, freq, phase, pose7, twist6 = wl.c3_instances(n, seed=1)
    with cb.CdprBatch(cb.default_config(4), n, device=device) as g:
        g.set_platform_state(pose7, twist6); g.set_sine_cmd(amp, freq, phase)
        ms = []
        for _ in range(4):
            g.step(k); ms.append(g.last_kernel_ms)
        t = float(np.mean(ms[1:]))
        out["nc4_reference_robot"] = {"value": n * k / (t * 1e-3), "unit": UNIT, "kernel_ms": t,
                                      "fp64_frac": flops_per_instance_step(4) * n * k / (t * 1e-3) / 1e12 / fp64_peak,
                                      "fp64_frac_executed": (executed_flops_per_instance_step(4) or 0) * n * k / (t * 1e-3) / 1e12 / fp64_peak}
    # the other shapes of the same kernel: per-cable position targets (Position mode), and a run whose command clamp
    # fires on every step (the inline-clamping steady body instead of the optimistic one)
    with cb.CdprBatch(cb.default_config(8), n, device=device) as g:
        g.set_platform_state(pose7, twist6)
        g.set_position_cmd(np.random.default_rng(3).uniform(-0.02, 0.02, (n, 8)).astype(np.float32))
        ms = []
        for _ in range(3):
            g.step(k); ms.append(g.last_kernel_ms)
        t = float(np.mean(ms[1:]))
        out["position_mode_nc8"] = {"value": n * k / (t * 1e-3), "unit": UNIT, "kernel_ms": t}
    sat_cfg = cb.default_config(8)
    sat_cfg.vel_pid.cmd_limit = 3.5   # below the ~4 N equilibrium tension: saturated throughout
    with cb.CdprBatch(sat_cfg, n, device=device) as g:
        g.set_platform_state(pose7, twist6); g.set_sine_cmd(amp, freq, phase)
        ms = []
        for _ in range(3):
            g.step(k); ms.append(g.last_kernel_ms)
        t = float(np.mean(ms[1:]))
        out["always_saturated_nc8"] = {"value": n * k / (t * 1e-3), "unit": UNIT, "kernel_ms": t}
    # config 1 (the reference's own operating point): ONE 4-cable robot stepped like the plugin does it -- one update() per
    # physics step: command in, one step, joint states + platform state out (the Gazebo path is capped at ~1e3 steps/s)
    with cb.CdprBatch(cb.default_config(4), 1, device=device) as g:
        axes = np.full((1, 4), 0.01, dtype=np.float32)
        for _ in range(50):
            g.step(1)
        t0 = time.perf_counter()
        reps = 2000
        for kk in range(reps):
            if kk % 10 == 0:
                g.set_velocity_cmd(axes)
            g.step(1)
            g.joint_states(); g.platform_state()
        dt1 = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for kk in range(reps):
            g.step(1)
        g.synchronize()
        dt2 = (time.perf_counter() - t0) / reps
        fused = plugin_update_rate(g, axes, reps)
        out["plugin_style_single_robot"] = {"updates_per_s_with_readback": fused, "updates_per_s_separate_calls": 1.0 / dt1, "steps_per_s_no_readback": 1.0 / dt2,
                                            "what": "N=1, NC=4, one plugin update per physics step through the C ABI, a new velocity command every 10 steps, host buffers, "
                                                    "synchronous. with_readback = cdpr_update: command + one step + joint states + platform state in ONE call (the step "
                                                    "kernel publishes into mapped host memory: one launch + one synchronisation); separate_calls = the round-1 path "
                                                    "(set command + cdpr_step(1) + cdpr_get_joint_states + cdpr_get_platform_state)"}
    # the full-semantics kernel (step_flexr.cuh / step_flex.cuh) at the headline size: independent robots on the launch values,
    # velocity hold below 2 cm/s, and hold + one biquad stage on the P input and on the D output with the reference's filter
    # constants (launch:27-32; that loop lives on its clamps: two steps out of three saturate)
    def flex_case(edit, independent=False, square=False):
        cfg = cb.default_config(8)
        if edit:
            edit(cfg)
        with cb.CdprBatch(cfg, n, device=device) as g:
            if independent:
                g.set_independent(True)
            g.set_platform_state(pose7, twist6)
            if square:   # the reference's squarevelocitytest constants (0.06 m/s, 0.05 Hz, 10 Hz publisher), one phase per robot
                g.set_square_velocity_cmd(np.full(n, 0.06), np.full(n, 0.05), phase)
            else:
                g.set_sine_cmd(amp, freq, phase)
            ms = []
            for _ in range(3):
                g.step(k); ms.append(g.last_kernel_ms)
            t = float(np.mean(ms[1:]))
            return {"value": n * k / (t * 1e-3), "unit": UNIT, "kernel_ms": t, "variant": g.kernel_variant, "kernel": g.kernel_detail}
    def hold(cfg): cfg.velocity_epsilon = 0.02
    def hold_1p1d(cfg): cfg.velocity_epsilon = 0.02; cfg.vel_pid.p_cascade = 1; cfg.vel_pid.d_cascade = 1
    def hold_10hz(cfg): cfg.velocity_epsilon = 0.02; cfg.sine_publish_hz = 10.0
    out["full_semantics_nc8"] = {"independent_launch_values": flex_case(None, independent=True), "hold_2cm_s": flex_case(hold),
                                 "hold_1p_1d": flex_case(hold_1p1d),
                                 "squarevelocitytest_hold_2cm_s": flex_case(hold_10hz, square=True),
                                 "what": "2^20 instances x 1000 steps per launch, per-instance sine commands crossing the hold band (C3 inputs)"}
    out["general_variant_nc8"] = out["full_semantics_nc8"]["hold_1p_1d"]   # the name round 1 and 2 reported it under
    ik_c2 = None
    for nc in (4, 8):
        for npose in (65536, 1 << 22):
            r = ik_c2_measurement(cb, wl, torch, device, hbm_peak, nc, npose)
            if nc == 8 and npose == 65536:
                ik_c2 = r                                  # BASELINE.json configs[1] at its stated size
            out[f"ik_sweep_nc{nc}_{npose}"] = r
    return out, ik_c2


def plugin_update_rate(g, axes, reps):
    """One fused update per physics step (command in, step, joint + platform state out in ONE call, CUDA graph inside)."""
    for kk in range(50):
        g.update(axes if kk % 10 == 0 else None)
    t0 = time.perf_counter()
    for kk in range(reps):
        g.update(axes if kk % 10 == 0 else None)
    return reps / (time.perf_counter() - t0)


def batch_state_gb(batch) -> float:
    return batch._L.cdpr_state_bytes(batch._h) / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--nc", type=int, default=8, choices=[4, 8])
    ap.add_argument("--instances", type=int, default=1 << 20, help="instances per GPU")
    ap.add_argument("--sim-steps", type=int, default=1000, help="physics steps per pass (one kernel launch)")
    ap.add_argument("--snapshot-every", type=int, default=100)
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"], help="multi-GPU trajectory gather implementation")
    ap.add_argument("--rollout-robots", type=int, default=64, help="robots per GPU of the config-5 rollout batch (x 4096 sequences x 256 steps)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-seconds", type=float, default=6.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-rollouts", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="N > 1: do not pin the rank to the cores next to its GPU")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "own":
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args)
    else:
        own_arm(args)


if __name__ == "__main__":
    main()
