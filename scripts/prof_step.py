"""One bench step (BASELINE configs[1]: 65,536 candidates x 50 denoise steps + EFE rollout) under the
CUDA profiler API, for `ncu --profile-from-start off ...` (developer tool; the numbers ncu prints are
never bench values).
  python scripts/prof_step.py [candidates=65536] [operand=bf16]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from active_inference_diffusion_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.CANDIDATES
operand = sys.argv[2] if len(sys.argv) > 2 else "bf16"
dev = torch.device("cuda", 0)
model = bench.build_scorer(dev)
obs = bench.build_inputs(1)[:B].to(dev)
with _lib.operand(operand):
    for _ in range(2):
        model(obs, horizon=bench.HORIZON, num_trajectories=bench.K_TRAJ)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model(obs, horizon=bench.HORIZON, num_trajectories=bench.K_TRAJ)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
