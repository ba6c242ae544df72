"""Scored-batch latency at small batch sizes (CUDA-graph replay of sampler + EFE), developer tool.
  [AID_PAIRS=0] [AID_DEBUG=4] python scripts/perf_small.py [batches=1,256,4096]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
model = bench.build_scorer(dev)
obs_all = bench.build_inputs(1).to(dev)
for b in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,256,4096").split(",")]:
    o = obs_all[:b].contiguous()
    f = lambda: model(o, horizon=bench.HORIZON, num_trajectories=bench.K_TRAJ)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f()
    e1.record()
    torch.cuda.synchronize()
    print(f"B={b}: {e0.elapsed_time(e1) / 10:.2f} ms per scored batch  (AID_PAIRS={os.environ.get('AID_PAIRS', '1')} AID_DEBUG={os.environ.get('AID_DEBUG', '0')})", flush=True)
