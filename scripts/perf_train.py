"""Device-timed training step of the hot path (BASELINE configs[2]): compute_diffusion_elbo forward,
backward (incl. the gradient penalty's double backward), gradient all-reduce over ranks, in the
bf16 and bf16x3 GEMM modes, next to the same graph on torch/cuBLAS fp32 (TF32 off and on).
Developer tool; under torchrun it shards the global batch over ranks.
  python scripts/perf_train.py [global_batch=32768] [reps=3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from active_inference_diffusion_b200 import ActiveInferenceConfig, DiffusionActiveInference, DiffusionConfig
from active_inference_diffusion_b200 import autograd_path as AP, distributed as D

GB = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
L, A, H = 128, 6, 512
torch.manual_seed(0)
cfg = ActiveInferenceConfig(latent_dim=L, hidden_dim=H, device="cpu", diffusion=DiffusionConfig(num_diffusion_steps=50))
ai = DiffusionActiveInference(L, A, L, cfg).to(dev)
ai.use_epistemic = False
B = GB // world
g = torch.Generator().manual_seed(1 + rank)
obs, rew, lat = torch.randn(B, L, generator=g).to(dev), torch.randn(B, generator=g).to(dev), torch.randn(B, L, generator=g).to(dev)
params = list(ai.latent_score_network.parameters()) + list(ai.latent_diffusion.parameters())


def step():
    for p in params:
        p.grad = None
    loss, _ = ai.compute_diffusion_elbo(obs, rew, lat)
    loss.backward()
    D.allreduce_grads(params)
    return loss


def timed(label):
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        fl = 9 * 48.1e6 * GB          # ~9 x F_fwd per sample
        print(f"{label:28s} global batch {GB} on {world} GPU(s): {float(ms):8.2f} ms/step  {GB / float(ms) * 1e3:9.0f} samples/s  "
              f"{fl / float(ms) / 1e9:7.1f} algorithmic TFLOP/s")


for prec in ("bf16", "bf16x3"):
    AP.set_precision(prec)
    timed(f"tcgen05 aid_gemm_nt [{prec}]")
# the same step captured as one CUDA graph (train_graph.GraphedElboStep): no launch gaps, no host syncs
from active_inference_diffusion_b200.train_graph import GraphedElboStep
for prec in ("bf16", "bf16x3"):
    AP.set_precision(prec)
    gs = GraphedElboStep(ai, B, params=params)
    eager_step, step = step, (lambda gs=gs: gs(obs, rew, lat)[0])
    timed(f"CUDA-graph step     [{prec}]")
    step = eager_step
    del gs
AP.set_precision("bf16x3")
# the same graph with torch's own matmul (cuBLAS) for comparison on the same box
AP.MatmulNT_apply_saved = AP.MatmulNT.apply
AP.linear = lambda x, w, b=None: torch.nn.functional.linear(x, w, b)
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    timed(f"torch F.linear (cuBLAS {'tf32' if tf32 else 'fp32'})")
if world > 1:
    dist.destroy_process_group()
