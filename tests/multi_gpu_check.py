#!/usr/bin/env python
"""Multi-GPU correctness check, run under torchrun on a real multi-GPU box (tests/test_multi_gpu.py spawns it when at
least two GPUs are visible):
   python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tests/multi_gpu_check.py
1. config 4: instances sharded by contiguous range over G ranks, snapshots all-gathered over NCCL ->
   rank 0 compares the gathered trajectory bitwise with a single-GPU run over all instances.
2. config 5: rollout cost vector, robots sharded over ranks, all-reduced -> compared with a single-GPU run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import cdpr_simulation_b200 as cb
from cdpr_simulation_b200 import workloads as wl, distributed as D

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nc, n_total, k, every = 8, 4096 * world + 3 * world, 300, 100   # not a multiple of the block size
cfg = cb.default_config(nc)
amp, freq, phase, pose7, twist6 = wl.c3_instances(n_total, seed=9)
lo, hi = D.shard_range(n_total, rank, world)

def run(lo, hi, device):
    n = hi - lo
    g = cb.CdprBatch(cfg, n, device=device)
    snaps = torch.zeros((k // every, 13, n), dtype=torch.float64, device=f"cuda:{device}")
    torch.cuda.synchronize()
    g.set_platform_state(pose7[lo:hi], twist6[lo:hi]); g.set_sine_cmd(amp[lo:hi], freq[lo:hi], phase[lo:hi])
    g.set_snapshots(every, snaps.data_ptr(), snaps.shape[0])
    g.step(k); g.synchronize()
    g.close()
    return snaps

local_snaps = run(lo, hi, local)
gathered = D.gather_trajectory(local_snaps)
traj = D.global_trajectory_to_instance_major(gathered)
ok1 = True
if rank == 0:
    ref = run(0, n_total, local)
    ok1 = bool(torch.equal(traj, ref))
    print(f"[config 4] {world} ranks x {hi - lo} instances, {k // every} snapshots gathered over NCCL: bitwise equal to 1-GPU run: {ok1}")

# config 4, fused: the step kernel writes its snapshots into every rank's symmetric-memory gather buffer over NVLink
ok3 = True
try:
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        g = cb.CdprBatch(cfg, hi - lo, device=local)
        g.set_stream(stream.cuda_stream)
        fg = D.FusedTrajectoryGather(g, every, k, stream, multicast=(os.environ.get("CDPR_NO_MULTICAST") is None))
        g.set_platform_state(pose7[lo:hi], twist6[lo:hi]); g.set_sine_cmd(amp[lo:hi], freq[lo:hi], phase[lo:hi])
        fg.before_pass(); g.step(k); fg.after_pass()
        stream.synchronize()
        fused = fg.latest().clone()
        g.close()
    ok3 = bool(torch.equal(fused, traj))
    print(f"[config 4 fused] rank {rank}: trajectory written by the step kernels over peer memory (multicast={fg.multicast}) == NCCL all-gather result: {ok3}")
except Exception as e:
    print(f"[config 4 fused] rank {rank}: symmetric memory unavailable here ({type(e).__name__}: {str(e)[:200]})")

# config 5: each rank owns n_robots_local robots, all evaluate the same command sequences
n_robots, n_seq, n_cmd, spc = 4 * world, 256, 8, 10
cmds = wl.c5_rollouts(n_seq, n_cmd, nc)
_, _, _, rp, rt = wl.c3_instances(n_robots, seed=5)
rlo, rhi = D.shard_range(n_robots, rank, world)
target, lam = [0.0, 0.0, 0.32], 0.05
g = cb.CdprBatch(cfg, (rhi - rlo) * n_seq, device=local)
cost_seq = torch.zeros(n_seq, dtype=torch.float64, device=f"cuda:{local}")
torch.cuda.synchronize()   # the handle runs on its own stream: buffers handed to it must be ready
g.rollout(rhi - rlo, n_seq, cmds, spc, target, lam, rp[rlo:rhi], rt[rlo:rhi], dev_cost_seq=cost_seq.data_ptr(), want_host_cost=False)
g.synchronize(); g.close()
D.allreduce_cost(cost_seq)
ok2 = True
if rank == 0:
    g = cb.CdprBatch(cfg, n_robots * n_seq, device=local)
    ref = torch.zeros(n_seq, dtype=torch.float64, device=f"cuda:{local}")
    torch.cuda.synchronize()
    g.rollout(n_robots, n_seq, cmds, spc, target, lam, rp, rt, dev_cost_seq=ref.data_ptr(), want_host_cost=False)
    g.synchronize(); g.close()
    rel = float(((cost_seq - ref).abs() / ref).max())
    ok2 = rel < 1e-13
    print(f"[config 5] {n_robots} robots x {n_seq} sequences x {n_cmd * spc} steps, cost vector all-reduced over {world} ranks: max rel diff vs 1-GPU {rel:.2e}")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (ok1 and ok2 and ok3) else 1)
