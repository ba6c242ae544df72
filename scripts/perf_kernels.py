"""Device time of each GEMM class inside the sampler (CUDA events around every launch of the
class, no profiler).  Developer tool: AID_DEBUG knobs isolate parts of the epilogues."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from active_inference_diffusion_b200 import DiffusionConfig, LatentDiffusionProcess, _lib
from tests.util import make_score_net

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = 3
L, O, H, NB = 128, 17, 512, 6
net, _ = make_score_net(L, O, H, NB, device="cuda")
diff = LatentDiffusionProcess(DiffusionConfig(num_diffusion_steps=T), L).cuda()
obs = torch.randn(B, O, device="cuda"); zT = torch.randn(B, L, device="cuda"); noise = torch.randn(T - 1, B, L, device="cuda")
run = lambda: diff.generate_latent_trajectory(net, B, obs, z_init=zT, noise=noise, return_trajectory=False)
for _ in range(2): run()
torch.cuda.synchronize()
classes = [("modln  K=512  N=1024", 2, H, 2 * H, 2.0 * B * H * 2 * H), ("fc1    K=512  N=2048", 0, H, 4 * H, 2.0 * B * H * 4 * H),
           ("fc2    K=2048 N=512 ", 1, 4 * H, H, 2.0 * B * 4 * H * H), ("attn   K=512  N=512 ", 1, H, H, 2.0 * B * H * H)]
out = []
for name, epi, k, n, fl in classes:
    _lib.profile_select(epi, k, n)
    run(); run()
    ms, cnt = _lib.profile_collect()
    out.append(f"{name}: {ms / cnt * 1e3:7.1f} us  {fl / (ms / cnt) / 1e9:7.1f} TF/s (n={cnt})")
_lib.profile_select(-1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
print(f"AID_DEBUG={os.environ.get('AID_DEBUG', '0'):>3s} total {e0.elapsed_time(e1) / 2:.2f} ms/run | " + " | ".join(out))
