"""Device-timed runs of the BASELINE configs that are not the bench line (developer tool):
  cfg#4  Humanoid-v4 state shape (obs 376, act 17), horizon 15, 32,768 candidates per GPU
         (262,144 over 8), 50 cosine denoise steps
  cfg#5  pixel HalfCheetah 84x84x9 uint8 frame stacks -> DrQ-v2 encoder -> latent diffusion + EFE,
         batch 4,096
One JSON line per config; inputs resident in HBM, CUDA events, >= 3 warm-up passes.
  python scripts/perf_configs.py [reps=5]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from active_inference_diffusion_b200 import (ActiveInferenceConfig, CandidateScorer, DiffusionConfig, DrQV2Encoder, _lib)

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda", 0)
L, H, NB, T = 128, 512, 6, 50
F_STEP = 2 * ((13 * NB + 2.5) * H * H + 1.5 * L * H)


def f_efe(A):
    return 2 * ((L * H + 5 * H * H + H * A) + ((L + A) * H + 2 * H * H + H * L)
                + (L * H + H * H / 2 + H) + ((L + 128) * H + 2 * H * H + H))


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, _lib.launch_count() // reps


def scorer(O, A, h):
    torch.manual_seed(0)
    cfg = ActiveInferenceConfig(latent_dim=L, hidden_dim=H, efe_horizon=h, device="cpu",
                                diffusion=DiffusionConfig(num_diffusion_steps=T, beta_schedule="cosine"))
    return CandidateScorer(O, A, cfg).eval().to(dev)


# ---- cfg#4
O, A, h, B = 376, 17, 15, 32768
m = scorer(O, A, h)
obs = torch.randn(B, O, generator=torch.Generator().manual_seed(1)).clamp_(-1, 1).to(dev)
ms, n = timed(lambda: m(obs, horizon=h, num_trajectories=1))
flops = B * (T * F_STEP + 2 * (O * H + 2 * H * H) + h * f_efe(A))
print(json.dumps({"config": "cfg#4 Humanoid-v4 state shape (obs 376, act 17), horizon 15, 50 cosine steps",
                  "candidates_per_gpu": B, "ms_per_pass": ms, "candidates_per_s": B / ms * 1e3,
                  "tflops": flops / ms / 1e9, "launches_per_pass": n}))
del m, obs

# ---- cfg#5
O, A, h, B = 128, 6, 5, 4096
m = scorer(O, A, h)
torch.manual_seed(0)
enc = DrQV2Encoder((3, 84, 84), feature_dim=128, frame_stack=3).to(dev).eval()
px = torch.randint(0, 256, (B, 9, 84, 84), dtype=torch.uint8, generator=torch.Generator().manual_seed(2)).to(dev)
ms_enc, n_enc = timed(lambda: enc(px))
ms, n = timed(lambda: m.forward_pixels(enc, px, horizon=h, num_trajectories=1))
hw = 42 * 42
f_img = 2 * hw * (81 * 32 + 288 * 64 + 576 * 128 + 1152 * 256) + 2 * 451584 * 256 + 2 * 256 * 128
flops = B * (f_img + T * F_STEP + 2 * (O * H + 2 * H * H) + h * f_efe(A))
print(json.dumps({"config": "cfg#5 pixel HalfCheetah 9x84x84 uint8 -> DrQ-v2 encoder -> 50-step latent diffusion + horizon-5 EFE",
                  "batch": B, "ms_per_pass": ms, "images_per_s": B / ms * 1e3, "encoder_ms": ms_enc,
                  "tflops": flops / ms / 1e9, "launches_per_pass": n, "encoder_launches": n_enc}))
