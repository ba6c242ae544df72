"""One eager native training step under the CUDA profiler API (for `ncu --profile-from-start off`).
  python scripts/prof_train_native.py [batch=32768] [operand=f16]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from active_inference_diffusion_b200 import ActiveInferenceConfig, DiffusionActiveInference, DiffusionConfig

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
operand = sys.argv[2] if len(sys.argv) > 2 else "f16"
dev = torch.device("cuda", 0)
L, A, H = 128, 6, 512
torch.manual_seed(0)
cfg = ActiveInferenceConfig(latent_dim=L, hidden_dim=H, device="cpu", diffusion=DiffusionConfig(num_diffusion_steps=50))
ai = DiffusionActiveInference(L, A, L, cfg)
ai.latent_score_network.randomize_zero_init(123)
ai = ai.to(dev)
ai.training_path, ai.training_operand, ai.elbo_score_only = "native", operand, True
g = torch.Generator().manual_seed(1)
obs, rew, lat = torch.randn(B, L, generator=g).to(dev), torch.randn(B, generator=g).to(dev), torch.randn(B, L, generator=g).to(dev)
params = list(ai.latent_score_network.parameters()) + list(ai.latent_diffusion.parameters())


def step():
    for p in params:
        p.grad = None
    loss, _ = ai.elbo_device(obs, rew, lat)
    loss.backward()
    ai._join_time_importance()


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
