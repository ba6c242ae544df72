"""Does scoring 65,536 candidates in row chunks (one wave of CTA pairs per chunk, L2-sized working
set) beat one 65,536-row call?  Developer experiment: same process, interleaved repetitions.
  python scripts/perf_chunks.py [reps=3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from active_inference_diffusion_b200 import ActiveInferenceConfig, CandidateScorer, DiffusionConfig
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
L, O, A, H, T, h, B = 128, 17, 6, 512, 50, 5, 65536
torch.manual_seed(0)
cfg = ActiveInferenceConfig(latent_dim=L, hidden_dim=H, efe_horizon=h, device="cpu",
                            diffusion=DiffusionConfig(num_diffusion_steps=T, beta_schedule="cosine"))
m = CandidateScorer(O, A, cfg).eval().cuda()
obs = torch.randn(B, O).clamp_(-1, 1).cuda()
plans = {"1x65536": [65536], "4x16384": [16384] * 4, "3x18944+8704": [18944] * 3 + [8704],
         "2x18944+27648": [18944, 18944, 27648], "2x32768": [32768] * 2, "18944+46592": [18944, 46592],
         "3x21888-ish": [21888, 21888, 21760]}


def run(plan):
    lo = 0
    outs = []
    for n in plan:
        outs.append(m(obs[lo:lo + n], horizon=h, num_trajectories=1)[0])
        lo += n
    return outs


for p in plans.values():
    run(p)
torch.cuda.synchronize()
res = {k: [] for k in plans}
for _ in range(reps):
    for k, p in plans.items():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(p)
        e1.record()
        torch.cuda.synchronize()
        res[k].append(e0.elapsed_time(e1))
for k, v in res.items():
    v = sorted(v)
    print(f"{k:16s} median {v[len(v) // 2]:8.2f} ms  min {v[0]:8.2f}  max {v[-1]:8.2f}  -> {B / v[len(v) // 2] * 1e3:9.0f} candidates/s")
