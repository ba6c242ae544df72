"""EFE scoring with the epistemic (MINE) estimator on / off and K = 1 / 10 trajectories through the
reference-facing call `compute_expected_free_energy_diffusion` (developer tool, SURVEY §8d secondary numbers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from active_inference_diffusion_b200 import ActiveInferenceConfig, DiffusionActiveInference, DiffusionConfig

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
L, A, H, h = 128, 6, 512, 5
torch.manual_seed(0)
cfg = ActiveInferenceConfig(latent_dim=L, hidden_dim=H, efe_horizon=h, device="cuda", diffusion=DiffusionConfig(num_diffusion_steps=50))
m = DiffusionActiveInference(L, A, L, cfg).cuda().eval()
lat = torch.randn(B, L, device="cuda")
for epi in (False, True):
    for K in (1, 10):
        m.use_epistemic = epi
        with torch.no_grad():
            for _ in range(2):
                m.compute_expected_free_energy_diffusion(lat, horizon=h, num_trajectories=K)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                m.compute_expected_free_energy_diffusion(lat, horizon=h, num_trajectories=K)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        flop = B * K * h * (5781504 + (203.6e6 if epi else 0))
        print(f"EFE B={B} K={K} h={h} epistemic={'on ' if epi else 'off'}: {ms:8.2f} ms  {B / ms * 1e3:10.0f} candidates/s  {flop / ms / 1e9:7.1f} TFLOP/s")

if len(sys.argv) > 2 and sys.argv[2] == "profile":
    # per-kernel device time of one epistemic-on K=1 call (developer aid)
    from torch.profiler import profile, ProfilerActivity
    m.use_epistemic = True
    with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        m.compute_expected_free_energy_diffusion(lat, horizon=h, num_trajectories=1)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
