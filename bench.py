#!/usr/bin/env python
"""bench.py — EFE candidate action samples/sec on B200 (BASELINE.json metric).

A "step" = one pass of the hot path over one batch of synthetic observations:
  per candidate row: own noise -> T=50 cosine-schedule reverse diffusion of the latent score network
  conditioned on its observation -> latent -> horizon-5 expected-free-energy rollout (K=1 trajectory,
  epistemic term off: it is one batch-constant scalar) -> efe + first action.
Workload = BASELINE.json configs[1]: HalfCheetah-v4 state shape (obs 17, act 6), latent 128,
hidden 512, 6 DiT blocks, 65,536 candidates per GPU (weak scaling: every rank scores its own 65,536).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's sm_100a path
  python bench.py --impl reference [...]                         # the reference algorithm on host cores
Under torchrun (N>1) one process per GPU; rows are independent so there is no data-path collective,
only the timing barrier / max-over-ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "efe_candidate_action_samples_per_sec"
UNIT = "candidates/s"
L, O, A, H, NB, T, HORIZON, K_TRAJ = 128, 17, 6, 512, 6, 50, 5, 1
CANDIDATES = 65536
CPU_SAMPLE = 256          # BASELINE configs[0]: batch 256 on CPU

# algorithmic FLOPs: live math only
F_STEP = 2 * ((13 * NB + 2.5) * H * H + 1.5 * L * H)          # per candidate per denoise step
F_OBS = 2 * (O * H + 2 * H * H)                               # per candidate once
F_EFE = 2 * ((L * H + 5 * H * H + H * A) + ((L + A) * H + 2 * H * H + H * L)
             + (L * H + H * H / 2 + H) + ((L + 128) * H + 2 * H * H + H))   # per (candidate, k, t)
F_CANDIDATE = T * F_STEP + F_OBS + K_TRAJ * HORIZON * F_EFE


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"burst": d.get("bf16_tflops"), "sustained": d.get("bf16_tflops_sustained"),
                "hbm": d.get("hbm_gbs"), "source": "measured"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_inputs(seed: int):
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randn(CANDIDATES, O, generator=g).clamp_(-1, 1)


# ---------------------------------------------------------------------------------------------
def cpu_port_step(params, heads, sched, obs, seed):
    """One pass of the oracle port (plain torch fp32, CPU) over `obs` — the reference's algorithm."""
    import torch
    from oracle import restatement as R
    g = torch.Generator().manual_seed(seed)
    B = obs.shape[0]
    zT = torch.randn(B, L, generator=g)
    noise = [torch.randn(B, L, generator=g) for _ in range(T - 1)]
    with torch.no_grad():
        latent = R.generate_latent_trajectory(params, sched, zT, obs, noise)[-1]
        nz = [dict(policy=torch.randn(B, A, generator=g), reparam=torch.randn(B, L, generator=g))
              for _ in range(K_TRAJ * HORIZON)]
        cfg = dict(epistemic_weight=0.1, pragmatic_weight=1.0, consistency_weight=0.1, discount_factor=0.99,
                   preference_temperature=1.0)
        efe, _, first = R.expected_free_energy(heads, cfg, latent, HORIZON, K_TRAJ, nz)
    return efe, first


def cpu_models():
    import torch
    from oracle import restatement as R
    from active_inference_diffusion_b200 import ActiveInferenceConfig, CandidateScorer, DiffusionConfig
    torch.manual_seed(0)
    cfg = ActiveInferenceConfig(latent_dim=L, hidden_dim=H, efe_horizon=HORIZON, device="cpu",
                                diffusion=DiffusionConfig(num_diffusion_steps=T, beta_schedule="cosine"))
    m = CandidateScorer(O, A, cfg).eval()
    sd = lambda mod: {k: v.detach().clone() for k, v in mod.state_dict().items()}
    heads = dict(policy=sd(m.policy_network), dynamics=sd(m.latent_dynamics), value=sd(m.value_network),
                 reward=sd(m.reward_predictor))
    return sd(m.latent_score_network), heads, R.make_schedule(T, "cosine")


def time_cpu(steps: int, warmup: int):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    params, heads, sched = cpu_models()
    obs = build_inputs(1)[:CPU_SAMPLE]
    for i in range(warmup):
        cpu_port_step(params, heads, sched, obs, 100 + i)
    t0 = time.perf_counter()
    for i in range(steps):
        cpu_port_step(params, heads, sched, obs, 200 + i)
    dt = (time.perf_counter() - t0) / steps
    return CPU_SAMPLE / dt, dt, torch.get_num_threads()


def run_reference_same_box(args):
    """--impl reference --ref-device cuda: the same oracle port (= the reference's eager PyTorch
    fp32 algorithm, TF32 off as in the reference's defaults) on this box's GPU instead of its host
    cores.  Not the driver's reference arm (that one is the CPU run below): it is the 'same box'
    bar of SURVEY.md §8(d), quoted in DESIGN.md."""
    import torch
    from oracle import restatement as R
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", 0)
    params, heads, sched = cpu_models()
    to = lambda d: {k: v.to(dev) for k, v in d.items()}
    params, heads, sched = to(params), {k: to(v) for k, v in heads.items()}, to(sched)
    B = args.ref_candidates
    obs = build_inputs(1)[:B].to(dev)
    cfg = dict(epistemic_weight=0.1, pragmatic_weight=1.0, consistency_weight=0.1, discount_factor=0.99,
               preference_temperature=1.0)

    def step():
        with torch.no_grad():
            zT = torch.randn(B, L, device=dev)
            noise = [torch.randn(B, L, device=dev) for _ in range(T - 1)]
            latent = R.generate_latent_trajectory(params, sched, zT, obs, noise)[-1]
            nz = [dict(policy=torch.randn(B, A, device=dev), reparam=torch.randn(B, L, device=dev))
                  for _ in range(K_TRAJ * HORIZON)]
            return R.expected_free_energy(heads, cfg, latent, HORIZON, K_TRAJ, nz)[0]

    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"impl": "reference", "device": "cuda (eager PyTorch fp32, TF32 off)", "metric": METRIC,
                      "value": B / ms * 1e3, "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": warmup,
                      "ms_per_step": ms, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": f"{B} candidates per step, T={T} cosine reverse diffusion + horizon-{HORIZON} EFE",
                                 "latent_dim": L, "hidden_dim": H, "num_blocks": NB}}))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.ref_device == "cuda":
        return run_reference_same_box(args)
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    value, dt, cores = time_cpu(steps, warmup)
    cb = {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
          "sample": f"{CPU_SAMPLE} candidates per step (BASELINE configs[0]); same T={T}, horizon={HORIZON}, K={K_TRAJ}; "
                    f"{steps} timed steps after {warmup} warm-up; torch fp32 oracle port of the reference algorithm "
                    "(/root/reference is not present on the GPU box)"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "HalfCheetah-v4 state shape (obs 17, act 6): T=50 cosine reverse diffusion + horizon-5 EFE, "
                               f"{CPU_SAMPLE}-candidate samples on host cores", "latent_dim": L, "hidden_dim": H,
                   "num_blocks": NB, "diffusion_steps": T, "horizon": HORIZON, "num_trajectories": K_TRAJ},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from active_inference_diffusion_b200 import ActiveInferenceConfig, CandidateScorer, DiffusionConfig, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)
    cfg = ActiveInferenceConfig(latent_dim=L, hidden_dim=H, efe_horizon=HORIZON, device="cpu",
                                diffusion=DiffusionConfig(num_diffusion_steps=T, beta_schedule="cosine"))
    model = CandidateScorer(O, A, cfg).eval().to(dev)
    obs_host = build_inputs(1 + rank).pin_memory()
    obs_dev = obs_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident():
        return model(obs_dev, horizon=HORIZON, num_trajectories=K_TRAJ)

    efe_host = torch.empty(CANDIDATES, dtype=torch.float32).pin_memory()
    act_host = torch.empty(CANDIDATES, A, dtype=torch.float32).pin_memory()

    def step_e2e():
        o = obs_host.to(dev, non_blocking=True)
        efe, first, _ = model(o, horizon=HORIZON, num_trajectories=K_TRAJ)
        efe_host.copy_(efe, non_blocking=True)
        act_host.copy_(first, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()

    # --- timed region (device-resident inputs); dominant-kernel events + clocks sampled inside it
    _lib.profile_select(0, H, 4 * H)          # mlp.0: K=H, N=4H, bias + GELU -> packed bf16 epilogue
    _lib.reset_launch_count()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_step = timed(step_resident, args.steps)
    launches = _lib.launch_count()
    clk = clocks.stop() if rank == 0 else None
    k_ms, k_n = _lib.profile_collect()
    _lib.profile_select(-1)

    # --- end to end: pinned host observations in, efe + first action back to the host, every step
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    total = CANDIDATES * world
    value = total / (ms_step * 1e-3)
    e2e_value = total / (ms_e2e * 1e-3)

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, cores = time_cpu(3, 1)
        cb = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
              "sample": f"{CPU_SAMPLE} candidates per step (BASELINE configs[0]), 3 timed steps after 1 warm-up, "
                        "torch fp32 oracle port of the reference algorithm on the box's host cores"}

    if rank == 0:
        peaks = measured_peaks()
        per_launch_ms = k_ms / max(k_n, 1)
        fc1_flops = 2.0 * CANDIDATES * H * (4 * H)
        achieved = fc1_flops / (per_launch_ms * 1e-3) / 1e12 if k_n else None
        peak = peaks["sustained"] or peaks["burst"]
        step_tflops = F_CANDIDATE * CANDIDATES / (ms_step * 1e-3) / 1e12
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": "HalfCheetah-v4 state shape (obs 17, act 6): 65,536 EFE candidates per GPU x horizon 5 x 50 "
                            "cosine-schedule denoise steps (BASELINE configs[1])",
                "candidates_per_gpu": CANDIDATES, "latent_dim": L, "hidden_dim": H, "num_blocks": NB,
                "diffusion_steps": T, "horizon": HORIZON, "num_trajectories": K_TRAJ, "epistemic": "off",
                "parallelism": f"rows sharded over {world} rank(s), no data-path collective",
                "l2": "working set per step (>1.5 GB of noise + >0.5 GB of activations) exceeds the 126 MB L2; no flush needed",
                "precision": "bf16 tensor-core operands, fp32 accumulate / LayerNorm / residual / reverse step"},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": CANDIDATES * O * 4 * world, "d2h_bytes_per_step": CANDIDATES * (1 + A) * 4 * world},
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "tensor", "kernel": "gemm2_kernel<EPI_PACK,resident A,GELU> (CTA-pair tcgen05 kernel; mlp.0: [65536x512]x[512x2048] + bias + GELU -> packed bf16; largest share of the step's FLOPs)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']}); burst figure {peaks['burst']}",
                "launches_timed": k_n, "avg_launch_ms": per_launch_ms, "flops_per_launch": fc1_flops,
                "traffic": 295.8e6, "traffic_note": "dram bytes read+written per launch (86.0 + 209.7 MB) from profiles/r1_ncu_full_gemm_kernels_v5_pairs.csv (ncu --set full); algorithmic: 67 MB packed A in + 268 MB packed activations out",
                "whole_step_tflops": step_tflops, "whole_step_frac_of_peak": step_tflops / peak},
            "cpu_baseline": cb,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="with --impl reference: cuda = the oracle port as eager PyTorch fp32 on this box's GPU")
    ap.add_argument("--ref-candidates", type=int, default=CANDIDATES)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
