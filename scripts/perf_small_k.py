"""Scored-batch latency with K rollouts per candidate at small batch sizes: trajectories evaluated as
K*B rows of one rollout vs K sequential rollouts (HeadsBundle.TRAJECTORY_ROWS_MAX).  Developer tool.
  python scripts/perf_small_k.py [batches=1,256] [K=10]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
obs_all = bench.build_inputs(1).to(dev)
for b in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,256").split(",")]:
    for cap in (32768, 0):
        m = bench.build_scorer(dev)                      # fresh module: fresh graph cache
        m.heads.TRAJECTORY_ROWS_MAX = cap
        o = obs_all[:b].contiguous()
        f = lambda: m(o, horizon=bench.HORIZON, num_trajectories=K)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            f()
        e1.record()
        torch.cuda.synchronize()
        print(f"B={b} K={K} trajectories {'as rows' if cap else 'sequential'}: {e0.elapsed_time(e1) / 10:.2f} ms per scored batch", flush=True)
