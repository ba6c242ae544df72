import sys, torch
sys.path.insert(0, '/root/repo')
from active_inference_diffusion_b200 import _lib
for n in (4096, 32768, 262144):
    for dist_name, t in (("uniform", torch.rand(n, device="cuda")), ("concentrated", (torch.rand(n, device="cuda") * 0.05))):
        loss = torch.rand(n, device="cuda")
        w = torch.ones(100, device="cuda")
        for _ in range(2): _lib.time_importance_update(t, loss, w)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): _lib.time_importance_update(t, loss, w)
        e1.record(); torch.cuda.synchronize()
        print(f"k_time_importance n={n} t {dist_name}: {e0.elapsed_time(e1)/5*1e3:.0f} us")
