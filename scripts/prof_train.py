"""Per-kernel device-time breakdown of one training step (torch.profiler, CUDA activities only).
Developer tool: python scripts/prof_train.py [batch=32768] [precision=bf16]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from active_inference_diffusion_b200 import ActiveInferenceConfig, DiffusionActiveInference, DiffusionConfig
from active_inference_diffusion_b200 import autograd_path as AP

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
AP.set_precision(sys.argv[2] if len(sys.argv) > 2 else "bf16")
L, A, H = 128, 6, 512
torch.manual_seed(0)
cfg = ActiveInferenceConfig(latent_dim=L, hidden_dim=H, device="cpu", diffusion=DiffusionConfig(num_diffusion_steps=50))
ai = DiffusionActiveInference(L, A, L, cfg).cuda()
ai.use_epistemic = False
obs, rew, lat = torch.randn(B, L).cuda(), torch.randn(B).cuda(), torch.randn(B, L).cuda()
params = list(ai.latent_score_network.parameters()) + list(ai.latent_diffusion.parameters())


def step():
    for p in ai.parameters():
        p.grad = None
    loss, _ = ai.elbo_device(obs, rew, lat)
    loss.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
total = sum(e.device_time_total for e in rows)
print(f"B={B} precision={AP.PRECISION}: {total / 1e3:.2f} ms of device time in {sum(e.count for e in rows)} launches")
for e in rows[:40]:
    print(f"{e.device_time_total / 1e3:9.3f} ms {100 * e.device_time_total / total:5.1f}%  n={e.count:5d}  {e.key[:110]}")
