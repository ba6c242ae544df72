"""Round-2 device timings (developer tool): operand type x batch size x graph replay for the fused
scorer call (reverse diffusion + EFE), one JSON line per case.
  python scripts/perf_r2.py [reps=5]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from active_inference_diffusion_b200 import ActiveInferenceConfig, CandidateScorer, DiffusionConfig, _lib

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = torch.device("cuda", 0)
L, O, A, H, NB, T, h = 128, 17, 6, 512, 6, 50, 5


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


torch.manual_seed(0)
cfg = ActiveInferenceConfig(latent_dim=L, hidden_dim=H, efe_horizon=h, device="cpu",
                            diffusion=DiffusionConfig(num_diffusion_steps=T, beta_schedule="cosine"))
m = CandidateScorer(O, A, cfg).eval().to(dev)
from tests.util import perturb_state_dict
m.latent_score_network.load_state_dict(perturb_state_dict(m.latent_score_network.state_dict()))
for B in (65536, 4096, 256, 1):
    obs = torch.randn(B, O, generator=torch.Generator().manual_seed(1)).clamp_(-1, 1).to(dev)
    for operand in ("bf16", "f16"):
        for graph in ((False,) if B > 16384 else (False, True)):
            m.use_graph = graph
            m.latent_diffusion.use_graph = False
            for source in (("philox", "torch") if B == 65536 else ("philox",)):
                m.latent_diffusion.noise_source = source
                with _lib.operand(operand):
                    ms = timed(lambda: m(obs, horizon=h, num_trajectories=1), reps if B > 4096 else 4 * reps)
                print(json.dumps({"B": B, "operand": operand, "graph": graph, "noise": source, "ms": round(ms, 3),
                                  "candidates_per_s": round(B / ms * 1e3, 1)}), flush=True)
