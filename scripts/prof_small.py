"""One scored batch at a small batch size (persistent sampler kernel + EFE rollout, launched directly, no
graph) under the CUDA profiler API, for `ncu --profile-from-start off ...` (developer tool; numbers printed
under ncu are never bench values).
  python scripts/prof_small.py [batch=1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda", 0)
model = bench.build_scorer(dev)
model.use_graph = False
model.latent_diffusion.use_graph = False
obs = bench.build_inputs(1)[:B].to(dev)
for _ in range(2):
    model(obs, horizon=bench.HORIZON, num_trajectories=bench.K_TRAJ)
torch.cuda.synchronize()
torch.cuda.profiler.start()
model(obs, horizon=bench.HORIZON, num_trajectories=bench.K_TRAJ)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
