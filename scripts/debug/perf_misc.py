import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench
from active_inference_diffusion_b200 import _lib
dev = torch.device("cuda", 0)
model = bench.build_scorer(dev)
obs_all = bench.build_inputs(1)
def wall(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for nenv in (8, 64, 256):
    o = obs_all[:nenv].pin_memory()
    print(f"collect_actions {nenv} envs, 20 diffusion steps: {wall(lambda: model.collect_actions(o, 20)):.2f} ms wall", flush=True)
with _lib.operand("f16"):
    ld = model.latent_diffusion
    for b in (1, 64):
        o = obs_all[:b].to(dev)
        f = lambda: ld.generate_latent_trajectory(model.latent_score_network, b, o, deterministic=False, return_trajectory=False)
        print(f"f16 operands sampler B={b}: {wall(f):.2f} ms wall", flush=True)
