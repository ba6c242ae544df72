"""SM clock / power while the sampler runs in a given AID_DEBUG mode (developer tool)."""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from active_inference_diffusion_b200 import DiffusionConfig, LatentDiffusionProcess
from tests.util import make_score_net
B, T, L, O, H, NB = 65536, 6, 128, 17, 512, 6
net, _ = make_score_net(L, O, H, NB, device="cuda")
diff = LatentDiffusionProcess(DiffusionConfig(num_diffusion_steps=T), L).cuda()
obs = torch.randn(B, O, device="cuda"); zT = torch.randn(B, L, device="cuda"); noise = torch.randn(T - 1, B, L, device="cuda")
run = lambda: diff.generate_latent_trajectory(net, B, obs, z_init=zT, noise=noise, return_trajectory=False)
run(); torch.cuda.synchronize()
rows = []
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "50"],
                     stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append(l.strip()) for l in p.stdout], daemon=True).start()
time.sleep(0.3)
n0 = len(rows)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(40): run()
e1.record(); torch.cuda.synchronize()
n1 = len(rows); p.terminate()
clk = sorted(float(r.split(",")[0]) for r in rows[n0 + 2:n1] if r)
pw = sorted(float(r.split(",")[1]) for r in rows[n0 + 2:n1] if r)
print(f"AID_DEBUG={os.environ.get('AID_DEBUG','0'):>5s}: {e0.elapsed_time(e1)/40:.2f} ms/run, sm clock median {clk[len(clk)//2] if clk else None} MHz "
      f"(min {clk[0] if clk else None}), power median {pw[len(pw)//2] if pw else None} W, samples {len(clk)}")
