"""Device-timed training step (BASELINE configs[2]) on the native trunk kernels (csrc/train.inc) next to
the round-1 autograd path and torch/cuBLAS TF32, eager and as one CUDA graph.  Developer tool; under
torchrun it shards the global batch over ranks.
  python scripts/perf_train_native.py [global_batch=32768] [reps=5] [modes=native-f16,native-bf16,...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from active_inference_diffusion_b200 import ActiveInferenceConfig, DiffusionActiveInference, DiffusionConfig
from active_inference_diffusion_b200 import autograd_path as AP, distributed as D
from active_inference_diffusion_b200.train_graph import GraphedElboStep

GB = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
modes = (sys.argv[3] if len(sys.argv) > 3 else "native-f16,native-bf16,autograd-bf16,torch-tf32").split(",")
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
from tests.util import perturb_state_dict
ai.latent_score_network.load_state_dict({k: v.to(dev) for k, v in perturb_state_dict(
    {k: v.cpu() for k, v in ai.latent_score_network.state_dict().items()}).items()})
B = GB // world
g = torch.Generator().manual_seed(1 + rank)
obs, rew, lat = torch.randn(B, L, generator=g).to(dev), torch.randn(B, generator=g).to(dev), torch.randn(B, L, generator=g).to(dev)
params = list(ai.latent_score_network.parameters()) + list(ai.latent_diffusion.parameters())


def eager():
    for p in params:
        p.grad = None
    ai.elbo_score_only = True
    loss, _ = ai.elbo_device(obs, rew, lat)
    loss.backward()
    ai._join_time_importance()
    D.allreduce_grads(params)
    return loss


def timed(label, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{label:34s} global batch {GB} on {world} GPU(s): {float(ms):8.2f} ms/step  {GB / float(ms) * 1e3:9.0f} samples/s  "
              f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)


for mode in modes:
    path, prec = mode.split("-")
    if path == "torch":
        ai.training_path = "autograd"
        saved = AP.linear
        AP.linear = lambda x, w, b=None: torch.nn.functional.linear(x, w, b)
        torch.backends.cuda.matmul.allow_tf32 = prec == "tf32"
        timed(f"torch F.linear (cuBLAS {prec}) eager", eager)
        AP.linear = saved
        continue
    ai.training_path = path
    if path == "native":
        ai.training_operand = prec
    else:
        AP.set_precision(prec)
    timed(f"{mode} eager", eager)
    gs = GraphedElboStep(ai, B, params=params)
    timed(f"{mode} CUDA graph", lambda: gs(obs, rew, lat)[0])
    del gs
AP.set_precision("bf16x3")
if world > 1:
    dist.destroy_process_group()
