"""Quick device-timed run of the reverse-diffusion sampler (developer tool, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from active_inference_diffusion_b200 import DiffusionConfig, LatentDiffusionProcess, _lib
from tests.util import make_score_net

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 50
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
L, O, H, NB = 128, 17, 512, 6
net, _ = make_score_net(L, O, H, NB, device="cuda")
diff = LatentDiffusionProcess(DiffusionConfig(num_diffusion_steps=T), L).cuda()
obs = torch.randn(B, O, device="cuda")
zT = torch.randn(B, L, device="cuda")
noise = torch.randn(T - 1, B, L, device="cuda")
for _ in range(2):
    diff.generate_latent_trajectory(net, B, obs, z_init=zT, noise=noise, return_trajectory=False)
torch.cuda.synchronize()
_lib.reset_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    out = diff.generate_latent_trajectory(net, B, obs, z_init=zT, noise=noise, return_trajectory=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flops = B * (T * 42401792 + 1.066e6)
print(f"B={B} T={T}: {ms:.2f} ms/run, {B / ms * 1e3:.0f} samples/s, {flops / ms / 1e9:.1f} TFLOP/s "
      f"({flops / ms / 1e9 / 1623.1 * 100:.1f}% of 1623 TF), launches/run={_lib.launch_count() // reps}, finite={bool(torch.isfinite(out[-1]).all())}")
