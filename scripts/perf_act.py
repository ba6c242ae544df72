"""Wall-clock latency of DiffusionActiveInference.act() (one observation, as the reference's agents call it
per environment step: belief update by 50-step reverse diffusion, EFE over K = 10 rollouts x horizon 5, policy
head, one device->host read) and of its parts.  Developer tool.
  python scripts/perf_act.py [epistemic=1]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_efe import make_ai

L, A, H, T = 128, 6, 512, 50
ai, _, cfg = make_ai(L, A, H, T)
ai.use_epistemic = bool(int(sys.argv[1])) if len(sys.argv) > 1 else True
obs = torch.randn(L)


def wall(f, n=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


with torch.no_grad():
    print(f"act() one observation, K={10} h={cfg.efe_horizon} epistemic={'on' if ai.use_epistemic else 'off'}: {wall(lambda: ai.act(obs)):.2f} ms wall")
    o = obs.cuda().unsqueeze(0)
    print(f"  update_belief_via_diffusion: {wall(lambda: ai.update_belief_via_diffusion(o)):.2f} ms")
    lat = ai.update_belief_via_diffusion(o)["latent"]
    print(f"  compute_expected_free_energy_diffusion: {wall(lambda: ai.compute_expected_free_energy_diffusion(lat, horizon=cfg.efe_horizon)):.2f} ms")
    print(f"  policy head: {wall(lambda: ai.policy_network(lat)):.2f} ms")
    for nenv in (8, 64):
        ob = torch.randn(nenv, L).cuda()
        print(f"  update_belief_via_diffusion, {nenv} rows: {wall(lambda: ai.update_belief_via_diffusion(ob)):.2f} ms")
