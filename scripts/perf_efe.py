"""Device-timed EFE rollout at the bench shape (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from active_inference_diffusion_b200 import ActiveInferenceConfig, CandidateScorer, DiffusionConfig
B, L, O, A, H, h, K = 65536, 128, 17, 6, 512, 5, int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.manual_seed(0)
cfg = ActiveInferenceConfig(latent_dim=L, hidden_dim=H, efe_horizon=h, device="cpu", diffusion=DiffusionConfig(num_diffusion_steps=50))
m = CandidateScorer(O, A, cfg).eval().cuda()
lat = torch.randn(B, L, device="cuda")
pn = torch.randn(K * h, B, A, device="cuda"); rn = torch.randn(K * h, B, L, device="cuda")
run = lambda: m.heads.efe_rollout(lat, h, K, m.efe_config(), m.preference_temperature, pn, rn, None)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"EFE rollout B={B} K={K} h={h}: {ms:.2f} ms, {B*K*h*5781504/ms/1e9:.1f} TFLOP/s")
