import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import cdpr_simulation_b200 as cb
from cdpr_simulation_b200 import workloads as wl
n = 1 << 20
cfg = cb.default_config(8)
amp, freq, phase, pose7, twist6 = wl.c3_instances(n, 1)
pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
pi = [pin(x) for x in (pose7, twist6, amp, freq, phase)]
outs = [torch.empty(sh, dtype=torch.float64).pin_memory() for sh in ((n, 7), (n, 6), (n, 8), (n, 8), (n, 8))]
g = cb.CdprBatch(cfg, n)
def t(f, reps=5):
    g.synchronize(); ts=[]
    for _ in range(reps):
        t0=time.perf_counter(); f(); g.synchronize(); ts.append((time.perf_counter()-t0)*1e3)
    return min(ts)
print("reset ms", t(g.reset))
print("set_platform_state ms", t(lambda: g.set_platform_state(pi[0].numpy(), pi[1].numpy())))
print("set_sine ms", t(lambda: g.set_sine_cmd(pi[2].numpy(), pi[3].numpy(), pi[4].numpy())))
print("step(1000) ms", t(lambda: g.step(1000), 2))
print("platform_state ms", t(lambda: g.platform_state((outs[0].numpy(), outs[1].numpy()))))
print("joint_states ms", t(lambda: g.joint_states(tuple(o.numpy() for o in outs[2:]))))

# pipelined two-lane loop (as in bench.py), with parts switched off to see what does not overlap
g.close()
def lanes_run(do_in=True, do_out=True, do_reset=True, passes=8):
    lanes = []
    for lane in range(2):
        st = torch.cuda.Stream(); bt = cb.CdprBatch(cfg, n); bt.set_stream(st.cuda_stream); bt.set_async(True)
        ins = [pin(x) for x in (pose7, twist6, amp, freq, phase)]
        os_ = [torch.empty(sh, dtype=torch.float64).pin_memory() for sh in ((n, 7), (n, 6), (n, 8), (n, 8), (n, 8))]
        bt.set_platform_state(ins[0].numpy(), ins[1].numpy()); bt.set_sine_cmd(ins[2].numpy(), ins[3].numpy(), ins[4].numpy())
        lanes.append((bt, ins, os_))
    def one(k):
        bt, ins, os_ = lanes[k % 2]
        bt.synchronize()
        if do_reset: bt.reset()
        if do_in:
            bt.set_platform_state(ins[0].numpy(), ins[1].numpy()); bt.set_sine_cmd(ins[2].numpy(), ins[3].numpy(), ins[4].numpy())
        bt.step(1000)
        if do_out:
            bt.platform_state((os_[0].numpy(), os_[1].numpy())); bt.joint_states(tuple(o.numpy() for o in os_[2:]))
    for k in range(2): one(k)
    for bt, _, _ in lanes: bt.synchronize()
    t0 = time.perf_counter()
    for k in range(passes): one(k)
    for bt, _, _ in lanes: bt.synchronize()
    dt = (time.perf_counter() - t0) / passes * 1e3
    for bt, _, _ in lanes: bt.close()
    return dt
print("pipelined ms/pass full           ", lanes_run())
print("pipelined ms/pass no D2H         ", lanes_run(do_out=False))
print("pipelined ms/pass no H2D         ", lanes_run(do_in=False))
print("pipelined ms/pass no reset       ", lanes_run(do_reset=False))
print("pipelined ms/pass step only      ", lanes_run(False, False, False))
