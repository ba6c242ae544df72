"""Per-kernel device times inside one sampler run, from CUDA-event-free launch list (developer tool):
prints total time per kernel name using torch profiler-less cudaEvent bracketing of whole runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from active_inference_diffusion_b200 import DiffusionConfig, LatentDiffusionProcess
from tests.util import make_score_net
B, T, L, O, H, NB = 65536, 6, 128, 17, 512, 6
net, _ = make_score_net(L, O, H, NB, device="cuda")
diff = LatentDiffusionProcess(DiffusionConfig(num_diffusion_steps=T), L).cuda()
obs = torch.randn(B, O, device="cuda"); zT = torch.randn(B, L, device="cuda"); noise = torch.randn(T - 1, B, L, device="cuda")
run = lambda: diff.generate_latent_trajectory(net, B, obs, z_init=zT, noise=noise, return_trajectory=False)
for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): run()
e1.record(); torch.cuda.synchronize()
print(f"AID_DEBUG={os.environ.get('AID_DEBUG','0')} AID_CHAIN={os.environ.get('AID_CHAIN','1')}: {e0.elapsed_time(e1)/3/T:.3f} ms/step")
