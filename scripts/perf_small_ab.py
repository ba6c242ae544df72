"""Small-batch sampler: persistent kernel (csrc/small.inc) vs the tcgen05 launch chain, developer tool.
  python scripts/perf_small_ab.py [batches=1,7,32,33,64,256]
Runs itself twice (AID_SMALL_MAX=0 and default) with the same Philox seed, compares the latents and
prints the time per sampler call (50 cosine steps, BASELINE dims) and per scored batch (sampler + EFE)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(batches, out_path):
    import torch
    import bench
    dev = torch.device("cuda", 0)
    model = bench.build_scorer(dev)
    obs_all = bench.build_inputs(1).to(dev)
    ld = model.latent_diffusion
    res = {}
    for b in batches:
        o = obs_all[:b].contiguous()
        ld.use_graph = False
        ld.seed_philox(1234, dev)
        z = ld.generate_latent_trajectory(model.latent_score_network, b, o, deterministic=False,
                                          return_trajectory=False)[-1].clone()
        res[b] = z.cpu()
        for graph in (False, True):
            ld.use_graph = graph
            f = lambda: ld.generate_latent_trajectory(model.latent_score_network, b, o, deterministic=False,
                                                      return_trajectory=False)
            for _ in range(3):
                f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                f()
            e1.record()
            torch.cuda.synchronize()
            print(f"  B={b:4d} sampler {'graph' if graph else 'eager'}: {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
        ld.use_graph = "auto"
        f = lambda: model(o, horizon=bench.HORIZON, num_trajectories=bench.K_TRAJ)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            f()
        e1.record()
        torch.cuda.synchronize()
        print(f"  B={b:4d} scored batch (graph): {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
    torch.save(res, out_path)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child([int(x) for x in sys.argv[3].split(",")], sys.argv[2])
        sys.exit(0)
    batches = sys.argv[1] if len(sys.argv) > 1 else "1,7,32,33,64,256"
    outs = []
    variants = [("chain", {"AID_SMALL_MAX": "0"}), ("persistent", {})]
    if os.environ.get("AB_ALL"):
        variants = [("chain", {"AID_SMALL_MAX": "0"}), ("grid", {"AID_SMALL_CLUSTER": "0"}),
                    ("cluster8", {"AID_SMALL_CLUSTER": "8"}), ("cluster16", {"AID_SMALL_CLUSTER": "16"})]
    for tag, env in variants:
        path = f"/tmp/small_ab_{tag}.pt"
        print(f"[{tag}]", flush=True)
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, __file__, "--child", path, batches], env=e, timeout=900)
        if r.returncode != 0:
            print(f"[{tag}] FAILED rc={r.returncode}")
            continue
        outs.append(path)
    import torch
    a = torch.load(outs[0])
    for path in outs[1:]:
        b = torch.load(path)
        for k in a:
            d = (a[k] - b[k]).norm() / a[k].norm()
            print(f"{os.path.basename(path)} B={k}: rel-L2 vs chain = {float(d):.3e}  max|d| = {float((a[k] - b[k]).abs().max()):.3e}  "
                  f"finite={bool(torch.isfinite(b[k]).all())}")
