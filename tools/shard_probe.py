"""probe: S sub-batches on S streams, K steps per graph replay, branches phase-shifted so that one branch's latency-bound
k_step overlaps another branch's bandwidth-bound k_obs (throw-away measurement)"""
import sys, time, torch
sys.path.insert(0, ".")
from dcmrta_b200 import BatchedTaskEnv
B, A, T, K = 65536, 20, 50, 16
for S in (1, 2, 3, 4):
    envs, streams = [], []
    for s in range(S):
        e = BatchedTaskEnv(B // S, A, T, auto_reset=True, seed=1234, first_gid=s * (B // S))
        e.generate(); e.reset(); envs.append(e); streams.append(torch.cuda.Stream())
    for _ in range(600):
        for e in envs: e.step(policy="random")
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream()
    with torch.cuda.stream(cap):
        g.capture_begin()
        cur = torch.cuda.current_stream()
        for si, (e, st) in enumerate(zip(envs, streams)):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                for _ in range(si): e.build_obs(e.leader)          # phase shift
                for _ in range(K): e.step(policy="random")
        for st in streams: cur.wait_stream(st)
        g.capture_end()
    torch.cuda.synchronize()
    for _ in range(10): g.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100): g.replay()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"S={S}: {dt/100/K*1e6:.1f} us per full step, {B*100*K/dt:.3e} env-steps/s", flush=True)
    for e in envs: e.close()
