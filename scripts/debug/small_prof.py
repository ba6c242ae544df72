"""AID_PROF_LIB=<libaid built with -DAID_SMALL_PROFILE> python scripts/debug/small_prof.py [batches]:
in-kernel cycle breakdown of the persistent sampler (CTA 0 and the last CTA print their counters)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, bench
from active_inference_diffusion_b200 import _lib
if os.environ.get('AID_PROF_LIB'):
    _lib.LIB_PATHS['bf16'] = os.environ['AID_PROF_LIB']
dev = torch.device("cuda", 0)
model = bench.build_scorer(dev)
obs = bench.build_inputs(1).to(dev)
ld = model.latent_diffusion
ld.use_graph = False
for b in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,256").split(",")]:
    o = obs[:b].contiguous()
    for _ in range(2):
        ld.generate_latent_trajectory(model.latent_score_network, b, o, deterministic=False, return_trajectory=False)
    torch.cuda.synchronize()
    print("B", b, flush=True)
