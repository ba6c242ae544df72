"""Time the DrQ-v2 encoder forward at the BASELINE cfg#5 image shape.
usage: python scripts/perf_encoder.py [batch] [iters] [precision]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from active_inference_diffusion_b200 import DrQV2Encoder, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
torch.manual_seed(0)
enc = DrQV2Encoder((3, 84, 84), feature_dim=128, frame_stack=3).cuda().eval()
enc.precision = prec
x = torch.randint(0, 256, (B, 9, 84, 84), dtype=torch.uint8, device="cuda")
for _ in range(2):
    y = enc(x)
torch.cuda.synchronize()
_lib.reset_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    y = enc(x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
# live FLOPs per image: 4 convs on the real 42x42 grid + the D -> 2F projection
hw = 42 * 42
flop = 2 * hw * (81 * 32 + 288 * 64 + 576 * 128 + 1152 * 256) + 2 * 451584 * 256 + 2 * 256 * 128
print(f"B={B} {prec}: {ms:.2f} ms/forward, {B / ms * 1e3:.0f} images/s, {B * flop / ms / 1e9:.1f} TFLOP/s "
      f"({flop / 1e9:.3f} GFLOP/image), launches/forward={_lib.launch_count() // iters}, finite={bool(torch.isfinite(y).all())}")
