"""Sampler time per row tile for a sweep of batch sizes + row-independence check (developer tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from active_inference_diffusion_b200 import DiffusionConfig, LatentDiffusionProcess
from tests.util import make_score_net

T = 4
L, O, H, NB = 128, 17, 512, 6
net, _ = make_score_net(L, O, H, NB, device="cuda")
diff = LatentDiffusionProcess(DiffusionConfig(num_diffusion_steps=T), L).cuda()
g = torch.Generator(device="cuda").manual_seed(0)
BMAX = 65536
obs = torch.randn(BMAX, O, device="cuda", generator=g)
zT = torch.randn(BMAX, L, device="cuda", generator=g)
noise = torch.randn(T - 1, BMAX, L, device="cuda", generator=g)
ref = diff.generate_latent_trajectory(net, 256, obs[:256], z_init=zT[:256], noise=noise[:, :256].contiguous(), return_trajectory=False)[-1]
for tiles in [int(a) for a in sys.argv[1:]] or [144, 148, 150, 160, 176, 190, 192, 194, 208, 222, 224, 296, 298, 384]:
    B = tiles * 128
    n = noise[:, :B].contiguous()
    for _ in range(2):
        out = diff.generate_latent_trajectory(net, B, obs[:B], z_init=zT[:B], noise=n, return_trajectory=False)[-1]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = diff.generate_latent_trajectory(net, B, obs[:B], z_init=zT[:B], noise=n, return_trajectory=False)[-1]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"tiles={tiles:4d} B={B:6d}: {ms:8.2f} ms  {ms / tiles * 1e3:7.1f} us/tile  same_rows={bool(torch.equal(out[:256], ref))}")
