"""Training step of the DrQ-v2 encoder (forward + backward, 9x84x84 frame stacks) with the native
convolution backend (conv_ops.conv3x3 on the tcgen05 GEMMs) next to the cuDNN-backed graph.  Developer tool.
  python scripts/perf_encoder_train.py [batch=256] [precision=bf16x3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from active_inference_diffusion_b200 import DrQV2Encoder

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
precision = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
torch.manual_seed(0)
enc = DrQV2Encoder((3, 84, 84), feature_dim=256, frame_stack=3).cuda().train()
enc.precision = precision
x = torch.rand(B, 9, 84, 84, device="cuda")


def step():
    enc.zero_grad(set_to_none=True)
    enc(x).square().mean().backward()


for backend in ("native", "torch"):
    enc.conv_backend = backend
    for allow in ((False, True) if backend == "torch" else (False,)):
        torch.backends.cudnn.allow_tf32 = allow
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        tag = backend + (" (cuDNN, TF32 %s)" % ("on" if allow else "off") if backend == "torch" else f" ({precision})")
        print(f"encoder training step B={B}: {tag:32s} {ms:8.2f} ms  {B / ms * 1e3:9.0f} images/s  peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
if len(sys.argv) > 3 and sys.argv[3] == "profile":
    from torch.profiler import profile, ProfilerActivity
    enc.conv_backend = "native"
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
