"""TEST INFRASTRUCTURE — the parity oracle.  Never imported by the product path.

A plain-PyTorch (fp32, CPU) restatement of the reference's hot path, written as
pure functions over `state_dict`-style parameter dictionaries so that it can run
on the GPU box where /root/reference does not exist.  Every function cites the
reference lines it restates.  The restatement is pinned to the real reference by
`tests/test_oracle_vs_reference.py` (runs only where /root/reference exists) and
by the golden vectors under `tests/golden/` that `oracle/gen_golden.py` produced
by executing the unmodified reference in this container.

The reference has no tests, golden vectors or known-answer fixtures of its own
(SURVEY.md §4), so the pin is "outputs of the reference itself run here".

All randomness is injected: every function that draws noise in the reference
takes the noise tensors explicitly, in the reference's draw order.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


def sub(params: Params, prefix: str) -> Params:
    """View of `params` restricted to keys under `prefix.` (prefix stripped)."""
    p = prefix + "."
    return {k[len(p):]: v for k, v in params.items() if k.startswith(p)}


def linear(params: Params, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, params[name + ".weight"], params.get(name + ".bias"))


def layer_norm(params: Params, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), params[name + ".weight"], params[name + ".bias"], 1e-5)


# --------------------------------------------------------------------------
# Score network — models/score_networks.py
# --------------------------------------------------------------------------

def sinusoidal_embedding(time: torch.Tensor, dim: int, freq_scale: torch.Tensor) -> torch.Tensor:
    """models/score_networks.py:282-291 — [sin | cos], freq_i = exp(-i ln(1e4)/(half-1)) * freq_scale."""
    half = dim // 2
    k = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half, device=time.device) * -k) * freq_scale
    arg = time[:, None] * freqs[None, :]
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


def time_embed(p: Params, time: torch.Tensor) -> torch.Tensor:
    """models/score_networks.py:41-46."""
    dim = p["time_embed.1.weight"].shape[1]
    e = sinusoidal_embedding(time, dim, p["time_embed.0.freq_scale"])
    return linear(p, "time_embed.3", F.silu(linear(p, "time_embed.1", e)))


def continuous_time_embed(p: Params, t_norm: torch.Tensor) -> torch.Tensor:
    """models/score_networks.py:60-66."""
    h = F.silu(linear(p, "continuous_time_embed.0", t_norm))
    h = F.silu(linear(p, "continuous_time_embed.2", h))
    return linear(p, "continuous_time_embed.4", h)


def obs_encoder(p: Params, obs: torch.Tensor) -> torch.Tensor:
    """models/score_networks.py:49-59 in eval() mode (Dropout(0.1) is the identity)."""
    h = F.silu(layer_norm(p, "obs_encoder.1", linear(p, "obs_encoder.0", obs)))
    h = F.silu(layer_norm(p, "obs_encoder.5", linear(p, "obs_encoder.4", h)))
    return layer_norm(p, "obs_encoder.8", linear(p, "obs_encoder.7", h))


def ada_layer_norm(p: Params, name: str, x: torch.Tensor, cond: torch.Tensor) -> torch.Tensor:
    """models/score_networks.py:257-270 — LN(x)(1+scale)+shift, (scale,shift)=chunk(Linear(SiLU(c)))."""
    ss = linear(p, name + ".adaLN_modulation.1", F.silu(cond))
    scale, shift = ss.chunk(2, dim=-1)
    return F.layer_norm(x, (x.shape[-1],), None, None, 1e-5) * (1 + scale) + shift


def seq1_attention(p: Params, name: str, x: torch.Tensor) -> torch.Tensor:
    """models/score_networks.py:189-194,224-227 — nn.MultiheadAttention over ONE token.

    softmax over a single key is exactly 1, so the output is out_proj(W_v x + b_v);
    W_v = in_proj_weight[2H:3H] (verified == reference to 0.0 abs diff
    by tests/test_oracle_vs_reference.py).
    """
    H = x.shape[-1]
    wv = p[name + ".in_proj_weight"][2 * H:3 * H]
    bv = p[name + ".in_proj_bias"][2 * H:3 * H]
    v = F.linear(x, wv, bv)
    return F.linear(v, p[name + ".out_proj.weight"], p[name + ".out_proj.bias"])


def num_blocks(p: Params) -> int:
    n = 0
    while f"transformer_blocks.{n}.mlp.0.weight" in p:
        n += 1
    return n


def score_is_continuous(time: torch.Tensor) -> bool:
    """models/score_networks.py:121 — batch-GLOBAL branch (max/min over the whole tensor)."""
    return bool(time.max() <= 1.0 and time.min() >= 0.0)


def score_conditioning(p: Params, time: torch.Tensor, obs: Optional[torch.Tensor], batch: int
                       ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """models/score_networks.py:119-153 — returns (conditioning, time_weight or None)."""
    if score_is_continuous(time):
        t_sin = time_embed(p, time * 999.0)
        t_cont = continuous_time_embed(p, 2.0 * time.view(-1, 1) - 1.0)
        t_emb = t_sin + p["time_scale"] * t_cont
        time_weight = torch.sqrt(1.0 / (1e-5 + time.view(-1, 1)))
    else:
        t_emb = time_embed(p, time)
        time_weight = None
    if obs is not None:
        o_emb = obs_encoder(p, obs)
    else:
        o_emb = torch.zeros(batch, p["obs_encoder.8.weight"].shape[0], device=p["obs_encoder.8.weight"].device)
    return t_emb + o_emb, time_weight


def score_forward(p: Params, z_t: torch.Tensor, time: torch.Tensor,
                  obs: Optional[torch.Tensor] = None) -> torch.Tensor:
    """models/score_networks.py:101-171 (eval mode)."""
    cond, time_weight = score_conditioning(p, time, obs, z_t.shape[0])
    h = linear(p, "latent_proj", z_t)
    for i in range(num_blocks(p)):
        b = f"transformer_blocks.{i}"
        h = h + seq1_attention(p, b + ".attention", ada_layer_norm(p, b + ".norm1", h, cond))
        m = ada_layer_norm(p, b + ".norm2", h, cond)
        m = linear(p, b + ".mlp.2", F.gelu(linear(p, b + ".mlp.0", m)))
        h = h + m
    h = ada_layer_norm(p, "norm_final", h, cond)
    s = F.linear(F.silu(linear(p, "output_proj.0", h)), p["output_proj.2.weight"])
    s = torch.clamp(s, -10, 10) * p["output_multiplier"]
    if time_weight is not None:
        s = s * time_weight
    return s


# --------------------------------------------------------------------------
# Diffusion process — core/diffusion.py
# --------------------------------------------------------------------------

def make_schedule(num_steps: int, beta_schedule: str = "cosine",
                  beta_start: float = 1e-4, beta_end: float = 0.02) -> Dict[str, torch.Tensor]:
    """core/diffusion.py:106-144 — computed in torch fp32 exactly as the reference does."""
    steps = num_steps
    if beta_schedule == "cosine":
        s = 0.008
        x = torch.linspace(0, steps, steps + 1)
        ac = torch.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
        ac = ac / ac[0]
        betas = torch.clamp(1 - (ac[1:] / ac[:-1]), min=1e-4, max=0.999)
    elif beta_schedule == "linear":
        betas = torch.linspace(beta_start, beta_end, steps)
    else:
        raise ValueError(f"Unknown schedule: {beta_schedule}")
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, dim=0)
    ac_prev = F.pad(ac[:-1], (1, 0), value=1.0)
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return {
        "betas": betas, "alphas": alphas, "alphas_cumprod": ac, "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": torch.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": torch.log(torch.clamp(post_var, min=1e-20)),
    }


def reverse_step_coefficients(sched: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Per-step scalars consumed by p_sample (core/diffusion.py:208-255), computed with the
    same fp32 torch expressions the reference evaluates on every call."""
    betas, alphas = sched["betas"], sched["alphas"]
    ac, ac_prev = sched["alphas_cumprod"], sched["alphas_cumprod_prev"]
    return {
        "sqrt_one_minus_ac": sched["sqrt_one_minus_alphas_cumprod"],
        "sqrt_recip_alpha": 1.0 / torch.sqrt(alphas),
        "coef1": betas * torch.sqrt(ac_prev) / (1.0 - ac),
        "coef2": (1.0 - ac_prev) * torch.sqrt(alphas) / (1.0 - ac),
        "sigma": torch.sqrt(sched["posterior_variance"]),
    }


def p_sample(sched: Dict[str, torch.Tensor], z_t: torch.Tensor, t: int, score: torch.Tensor,
             noise: Optional[torch.Tensor], deterministic: bool = False) -> torch.Tensor:
    """core/diffusion.py:208-255 for a batch-constant step index t.

    pred_z0 = (z + sqrt(1-abar_t) s) / sqrt(alpha_t)   [sic: alpha_t, not abar_t  — :224-227]
    """
    c = reverse_step_coefficients(sched)
    pred = (z_t + c["sqrt_one_minus_ac"][t] * score) * c["sqrt_recip_alpha"][t]
    mean = c["coef1"][t] * pred + c["coef2"][t] * z_t
    if deterministic or t == 0:
        return mean
    return mean + c["sigma"][t] * noise


def generate_latent_trajectory(p: Params, sched: Dict[str, torch.Tensor], z_T: torch.Tensor,
                               obs: Optional[torch.Tensor], step_noise: Sequence[torch.Tensor],
                               deterministic: bool = False) -> List[torch.Tensor]:
    """core/diffusion.py:176-206.  `z_T` is the reference's initial `torch.randn(B, L)`;
    `step_noise[i]` is the i-th `randn_like` drawn (one per step with t>0, none at t=0).

    The score net is called with `t.float()` = T-1 … 0, so steps t=1 and t=0 take the
    continuous-time branch of models/score_networks.py:121."""
    T = sched["betas"].shape[0]
    z = z_T
    traj = [z]
    draw = 0
    for t in reversed(range(T)):
        tb = torch.full((z.shape[0],), float(t), device=z.device)
        s = score_forward(p, z, tb, obs)
        eps = None
        if not deterministic and t != 0:
            eps = step_noise[draw]
            draw += 1
        z = p_sample(sched, z, t, s, eps, deterministic)
        traj.append(z)
    return traj


def collector_sample(p: Params, sched: Dict[str, torch.Tensor], z_init: torch.Tensor,
                     obs: torch.Tensor, step_noise: Sequence[torch.Tensor],
                     max_diffusion_steps: int) -> torch.Tensor:
    """utils/async_collector.py:530-595 — second call site of the sampler: `num_steps =
    min(max_diffusion_steps, T)`; score called with t = step/(T-1) (always continuous
    branch); p_sample with the integer step.  (NaN re-init branch :591-593 not taken.)"""
    T = sched["betas"].shape[0]
    num_steps = min(max_diffusion_steps, T)
    max_index = T - 1
    z = z_init
    draw = 0
    for step in reversed(range(num_steps)):
        tc = torch.full((z.shape[0],), step / max_index if max_index > 0 else 0.0, device=z.device)
        s = score_forward(p, z, tc, obs)
        eps = None
        if step != 0:
            eps = step_noise[draw]
            draw += 1
        z = p_sample(sched, z, step, s, eps, False)
    return z


def q_sample(sched: Dict[str, torch.Tensor], z0: torch.Tensor, t: torch.Tensor,
             noise: torch.Tensor) -> torch.Tensor:
    """core/diffusion.py:154-174."""
    a = sched["sqrt_alphas_cumprod"].gather(-1, t).view(-1, 1)
    b = sched["sqrt_one_minus_alphas_cumprod"].gather(-1, t).view(-1, 1)
    return a * z0 + b * noise


def log_snr(dp: Params, t: torch.Tensor) -> torch.Tensor:
    """core/diffusion.py:56-60."""
    return dp["log_snr_min"] + (dp["log_snr_max"] - dp["log_snr_min"]) * (1 - t)


def continuous_q_sample(dp: Params, z0: torch.Tensor, t: torch.Tensor, noise: torch.Tensor):
    """core/diffusion.py:62-91 — the 'sigmoid' schedule: alpha=sigmoid(lambda), sigma=sigmoid(-lambda)."""
    lam = log_snr(dp, t)
    alpha = torch.sigmoid(lam).view(-1, 1)
    sigma = torch.sigmoid(-lam).view(-1, 1)
    return torch.sqrt(alpha) * z0 + torch.sqrt(sigma) * noise, lam, alpha, sigma


def loss_weight(dp: Params, t: torch.Tensor) -> torch.Tensor:
    """core/diffusion.py:93-104."""
    lam = log_snr(dp, t)
    return torch.exp(-0.5 * (lam ** 2) / 4.0) * (torch.sin(t * np.pi) + 0.1)


def sample_latent_prior(dp: Params, eps: torch.Tensor) -> torch.Tensor:
    """core/diffusion.py:146-152."""
    return dp["latent_prior_mean"].unsqueeze(0) + torch.exp(dp["latent_prior_log_std"]).unsqueeze(0) * eps


# --------------------------------------------------------------------------
# EFE heads — models/{policy,value,dynamics}_*.py, core/active_inference.py
# --------------------------------------------------------------------------

def policy_forward(p: Params, z: torch.Tensor, eps: Optional[torch.Tensor]):
    """models/policy_networks.py:95-146 (state-dependent std, no squashing).
    `eps` is the standard-normal draw of Normal.rsample (`normal_`); None = deterministic.
    Returns (action, log_prob, mean, std)."""
    h = linear(p, "latent_encoder.3", F.relu(layer_norm(p, "latent_encoder.1", linear(p, "latent_encoder.0", z))))
    t = h
    i = 0
    while f"trunk.{i}.weight" in p:
        t = F.relu(layer_norm(p, f"trunk.{i + 1}", linear(p, f"trunk.{i}", t)))
        i += 3
    h = h + t
    mean = linear(p, "mean_head.2", F.relu(linear(p, "mean_head.0", h)))
    log_std = linear(p, "log_std_head.2", F.relu(linear(p, "log_std_head.0", h)))
    log_std = torch.clamp(log_std, -20, 2)
    std = torch.exp(log_std)
    action = mean if eps is None else mean + std * eps
    var = std ** 2
    logp = (-((action - mean) ** 2) / (2 * var) - log_std - math.log(math.sqrt(2 * math.pi))).sum(-1)
    return action, logp, mean, std


def normal_entropy(std: torch.Tensor) -> torch.Tensor:
    """torch.distributions.Normal.entropy: 0.5 + 0.5 ln(2 pi) + ln(std)."""
    return 0.5 + 0.5 * math.log(2 * math.pi) + torch.log(std)


def mlp_ln_relu(p: Params, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """Linear→LayerNorm→ReLU repeated, then a final Linear (value_networks.py:30-45,
    dynamics_models.py:27-40)."""
    i = 0
    h = x
    while f"{prefix}.{i + 1}.weight" in p and p[f"{prefix}.{i + 1}.weight"].dim() == 1:
        h = F.relu(layer_norm(p, f"{prefix}.{i + 1}", linear(p, f"{prefix}.{i}", h)))
        i += 3
    return linear(p, f"{prefix}.{i}", h)


def dynamics_forward(p: Params, state: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
    """models/dynamics_models.py:47-67 (residual=True): state + net([state, action])."""
    return state + mlp_ln_relu(p, "network", torch.cat([state, action], dim=-1))


def value_forward(p: Params, state: torch.Tensor, time: torch.Tensor) -> torch.Tensor:
    """models/value_networks.py:47-60 → [B, 1]."""
    dim = p["time_embed.1.weight"].shape[1]
    e = sinusoidal_embedding(time, dim, p["time_embed.0.freq_scale"])
    t_emb = F.relu(linear(p, "time_embed.1", e))
    return mlp_ln_relu(p, "network", torch.cat([state, t_emb], dim=-1))


def reward_head(p: Params, z: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """core/active_inference.py:160-167,245-253."""
    h = F.relu(layer_norm(p, "1", linear(p, "0", z)))
    h = F.relu(linear(p, "3", h))
    out = linear(p, "5", h)
    return out[:, 0], torch.exp(torch.clamp(out[:, 1], -5, 2))


def decode_observation_state(p: Params, z: torch.Tensor) -> torch.Tensor:
    """core/active_inference.py:237-242 — state-mode decoder with skip, eval mode."""
    def blk(i, x):
        return F.silu(layer_norm(p, f"{i}.1", linear(p, f"{i}.0", x)))
    h1 = blk(0, z)
    h2 = blk(1, h1) + h1
    h3 = blk(2, h2)
    return linear(p, "3", h3)


def predict_next_latent(dyn: Params, z: torch.Tensor, a: torch.Tensor):
    """core/active_inference.py:447-464 — next_mean = z + dynamics(z,a) = 2z + f(z,a)."""
    mean = z + dynamics_forward(dyn, z, a)
    return mean, torch.full_like(mean, float(np.log(0.1)))


def epistemic_value(ep: Params, dec: Params, mean: torch.Tensor, logvar: torch.Tensor,
                    z_eps: Sequence[torch.Tensor], dir_eps: Sequence[torch.Tensor],
                    perms: Sequence[torch.Tensor], running_mean: float):
    """core/active_inference.py:940-1063 (state mode, eval) with the decoder shim of SURVEY §8c(1):
    `self.decoder(z)` on a ModuleList crashes in the reference (:953), the oracle substitutes
    `decode_observation_state`.  Returns (epistemic[B], mi, joint, marginal_term, new_running_mean).

    z_eps: S draws [B,L]; dir_eps: 4 draws [S*B,L]; perms: S permutations of B (int64).
    MINE marginal term follows ema_loss (:828-836): forward value log(mean(exp(T_marg)))."""
    B = mean.shape[0]
    S = len(z_eps)
    std = torch.exp(0.5 * logvar)
    z_all = torch.cat([mean + e * std for e in z_eps], dim=0)
    eps_scale = ep["perturbation_scale"] if "perturbation_scale" in ep else torch.tensor(0.1)
    f0 = decode_observation_state(dec, z_all)
    feats = []
    for d in dir_eps:
        delta = F.normalize(d, dim=-1) * eps_scale
        diff = (decode_observation_state(dec, z_all + delta) - f0) / eps_scale
        h = F.relu(linear(ep, "feature_extractor.0", diff))
        h = F.relu(linear(ep, "feature_extractor.2", h))
        feats.append(linear(ep, "feature_extractor.4", h))
    j = torch.cat(feats, dim=1)
    j = linear(ep, "jacobian_projector.4",
               F.relu(layer_norm(ep, "jacobian_projector.1", linear(ep, "jacobian_projector.0", j))))
    lat = linear(ep, "latent_processor.2", F.relu(linear(ep, "latent_processor.0", z_all)))

    def mine(x):
        h = F.relu(linear(ep, "mine_network.0", x))
        h = F.relu(linear(ep, "mine_network.3", h))
        return linear(ep, "mine_network.6", h)

    t_joint = mine(torch.cat([j, lat], dim=1))
    j_marg = torch.cat([j[i * B:(i + 1) * B][perms[i]] for i in range(S)], dim=0)
    t_marg = mine(torch.cat([j_marg, lat], dim=1))
    t_exp = torch.exp(torch.logsumexp(t_marg, 0) - math.log(t_marg.shape[0]))
    new_rm = float(t_exp) if running_mean == 0 else 0.01 * float(t_exp) + 0.99 * running_mean
    marg_term = t_marg.exp().mean().log()
    mi = t_joint.mean() - marg_term
    return torch.clamp(mi.expand(B), min=0.0), mi, t_joint.mean(), marg_term, new_rm


def expected_free_energy(nets: Dict[str, Params], cfg: Dict[str, float], latent: torch.Tensor,
                         horizon: int, num_trajectories: int,
                         noise: Sequence[Dict[str, object]],
                         epistemic: str = "zero", running_mean: float = 0.0):
    """core/active_inference.py:314-396.

    nets: {'policy','dynamics','value','reward', optionally 'epistemic','decoder'}.
    cfg : epistemic_weight, pragmatic_weight, consistency_weight, discount_factor,
          preference_temperature.
    noise[k*horizon + t] = {'policy': [B,A], 'reparam': [B,L], and for epistemic='mine':
          'z': S×[B,L], 'dir': 4×[S*B,L], 'perm': S×[B]}  — reference draw order.
    epistemic: 'zero' → term omitted (it is one scalar for the whole batch and cannot change any
          per-candidate ranking); 'mine' → faithful estimator with decoder shim.

    Note `pragmatic_weight` is applied twice (:353 and :371) as written.
    Returns (efe[B], info dict of last-step means, first_action[B,A] of trajectory 0).
    """
    B = latent.shape[0]
    total = torch.zeros(B, device=latent.device)
    epi_l, prag_l, cons_l = [], [], []
    first_action = None
    for k in range(num_trajectories):
        cur = latent.clone()
        traj = torch.zeros(B, device=latent.device)
        for t in range(horizon):
            nz = noise[k * horizon + t]
            action, _, _, std = policy_forward(nets["policy"], cur, nz["policy"])
            if k == 0 and t == 0:
                first_action = action
            mean, logvar = predict_next_latent(nets["dynamics"], cur, action)
            nxt = mean + nz["reparam"] * torch.exp(0.5 * logvar)
            r_mean, _ = reward_head(nets["reward"], nxt)
            prag = cfg["pragmatic_weight"] * (r_mean / cfg["preference_temperature"])
            prag = prag + value_forward(nets["value"], nxt, torch.full((B,), float(t), device=latent.device)).squeeze(-1)
            cons = -normal_entropy(std).sum(-1)
            if epistemic == "mine":
                epi, _, _, _, running_mean = epistemic_value(
                    nets["epistemic"], nets["decoder"], mean, logvar,
                    nz["z"], nz["dir"], nz["perm"], running_mean)
            else:
                epi = torch.zeros(B, device=latent.device)
            step = (cfg["epistemic_weight"] * epi + cfg["pragmatic_weight"] * prag
                    + cfg["consistency_weight"] * cons)
            traj = traj + (cfg["discount_factor"] ** t) * step
            cur = nxt
        total = total + traj / num_trajectories
        epi_l.append(epi)
        prag_l.append(prag)
        cons_l.append(cons)
    info = {
        "epistemic_mean": torch.stack(epi_l).mean(),
        "pragmatic_mean": torch.stack(prag_l).mean(),
        "consistency_mean": torch.stack(cons_l).mean(),
    }
    return total, info, first_action


# --------------------------------------------------------------------------
# Training loss — core/active_inference.py:533-636, 709-771
# --------------------------------------------------------------------------

def gradient_penalty(p: Params, noisy: torch.Tensor, t: torch.Tensor, obs: torch.Tensor) -> torch.Tensor:
    """core/active_inference.py:709-729 — double-backward through the score net."""
    x = noisy.detach().requires_grad_(True)
    s = score_forward(p, x, t, obs)
    g = torch.autograd.grad(s.sum(), x, create_graph=True, retain_graph=True)[0]
    return torch.mean((g.norm(2, dim=1) - 1.0) ** 2)


def diffusion_elbo(score: Params, diff: Params, decoder: Params, reward: Params,
                   cfg: Dict[str, float], observations: torch.Tensor, rewards: torch.Tensor,
                   latents: torch.Tensor, t: torch.Tensor, noise: torch.Tensor,
                   prior_eps: torch.Tensor):
    """core/active_inference.py:533-636 (state mode, eval).  Returns (loss, info, per_sample_losses).

    loss = -elbo,  elbo = -recon + kl_w*kl*exp(-5 mean t) + diff_w*sm + 0.1*gp - reward_w*rl
    (signs as written)."""
    recon = F.mse_loss(decode_observation_state(decoder, latents), observations)
    noisy, lam, alpha, sigma = continuous_q_sample(diff, latents, t, noise)
    pred = score_forward(score, noisy, t, observations)
    true_score = -noise / (sigma + 1e-8)
    w = loss_weight(diff, t)
    per_sample = w.view(-1) * torch.sum((pred - true_score) ** 2, dim=1)
    sm = per_sample.mean()
    gp = gradient_penalty(score, noisy, t, observations)
    prior = sample_latent_prior(diff, prior_eps)
    kl = (0.5 * torch.sum((latents - prior) ** 2, dim=-1)).mean()
    klw = torch.exp(-5.0 * t.mean())
    r_mean, r_std = reward_head(reward, latents)
    rl = -torch.distributions.Normal(r_mean, r_std).log_prob(rewards).mean()
    elbo = (-recon + cfg["kl_weight"] * kl * klw + cfg["diffusion_weight"] * sm
            + 0.1 * gp - cfg["reward_weight"] * rl)
    info = {"reconstruction_loss": recon, "kl_loss": kl, "score_matching_loss": sm, "elbo": elbo,
            "reward_loss": rl, "grad_penalty": gp, "mean_time": t.mean(),
            "loss_weight_mean": w.mean()}
    return -elbo, info, per_sample


def importance_sample_time(weights: torch.Tensor, indices: torch.Tensor, jitter: torch.Tensor) -> torch.Tensor:
    """core/active_inference.py:731-748 with the multinomial draw (`indices`) and `rand` injected."""
    del weights  # the probabilities only drive the multinomial, which is injected
    return (indices.float() + jitter) / 100.0


def update_time_importance(weights: torch.Tensor, t: torch.Tensor, loss: torch.Tensor) -> torch.Tensor:
    """core/active_inference.py:750-771 — sequential per-sample EMA over 100 bins (fp32 storage,
    python-double arithmetic exactly as `.item()` round trips do it)."""
    w = weights.clone()
    idx = (t * 99).long().clamp(0, 99)
    for i in range(idx.shape[0]):
        b = int(idx[i])
        w[b] = 0.99 * float(w[b]) + 0.01 * float(loss[i])
    return w


def time_importance_bins(t: torch.Tensor) -> torch.Tensor:
    return (t * 99).long().clamp(0, 99)


# --------------------------------------------------------------------------
# Free energy — core/free_energy.py:30-91
# --------------------------------------------------------------------------

def free_energy_loss(score: Params, log_precision: torch.Tensor, states: torch.Tensor,
                     observations: torch.Tensor, current_time: float = 0.0,
                     prior_mean: Optional[torch.Tensor] = None, prior_std: float = 1.0):
    B = states.shape[0]
    if prior_mean is None:
        prior_mean = torch.zeros_like(states)
    complexity = 0.5 * torch.sum((states - prior_mean) ** 2 / (prior_std ** 2), dim=-1).mean()
    obs_err = torch.sum((observations - states) ** 2, dim=-1)
    accuracy = -0.5 * torch.exp(log_precision) * obs_err.mean()
    s = score_forward(score, states, torch.full((B,), current_time, device=states.device), observations)
    reg = 0.01 * torch.sum(s ** 2, dim=-1).mean()
    return complexity - accuracy + reg, {"complexity": complexity, "accuracy": -accuracy,
                                         "observation_error": obs_err.mean(),
                                         "score_regularization": reg}


# --------------------------------------------------------------------------
# Fokker-Planck belief update — core/belief_dynamics.py:97-172 (fp64)
# --------------------------------------------------------------------------

def belief_update_diag(mean: np.ndarray, variance: np.ndarray, observation: np.ndarray,
                       score: np.ndarray, eps: np.ndarray, *, dt: float, D: float, lr: float,
                       noise_scale: float, min_variance: float, max_variance: float):
    """Restatement of BeliefDynamics.update for the default Gaussian observation model and
    diagonal covariance.  The reference method cannot run as shipped (calls the undefined
    `_record_state_enhanced` :170 and differentiates a detached gradient :210,:234,:261 —
    SURVEY §8c(2)); this follows :110-167 with the closed-form gradient of
    log p = -0.5|z-o|^2/ns^2 - 0.5|z|^2 + z·s  (equal to the autodiff one, checked in
    tests/test_oracle_vs_reference.py) and the closed-form Hessian diagonal -(1/ns^2 + 1).

    As written the mean moves along -lr * ∇log p (i.e. `mean_drift = -lr * F_gradient` with
    F_gradient = ∇ total_log_prob, :126-137).  The Hessian is evaluated at the UPDATED mean
    (:155-157) — irrelevant here because it is constant.
    All arrays float64.  Returns (mean', variance', precision')."""
    mean = np.asarray(mean, np.float64)
    g = -(mean - observation) / (noise_scale ** 2) - mean + score
    drift = -lr * g
    noise = math.sqrt(2 * D * dt) * eps * noise_scale
    adaptive_dt = dt / (1 + 0.1 * np.linalg.norm(g))
    new_mean = mean + drift * adaptive_dt + noise
    h_diag = np.full_like(mean, -(1.0 / noise_scale ** 2 + 1.0))
    factor = np.exp((-2 * h_diag + 2 * D) * dt)
    min_eig = max(min_variance, 1e-8)
    new_var = np.clip(variance * factor, min_eig, max_variance)
    return new_mean, new_var, 1.0 / new_var


def belief_update_full(mean: np.ndarray, cov: np.ndarray, observation: np.ndarray, score: np.ndarray,
                       eps: np.ndarray, *, dt: float, D: float, lr: float, noise_scale: float,
                       min_variance: float):
    """Full-covariance branch (:139-150, 268-294, 296-339): Σ ← E Σ Eᵀ, E = expm((-H-Hᵀ+2D I)dt),
    H = -(1/ns²+1) I, eigen-clamp at max(min_variance,1e-8), condition-number regularisation,
    precision = inv(Σ + min_eig I).  float64 via torch.linalg."""
    mean_t = torch.as_tensor(mean, dtype=torch.float64)
    L = mean_t.shape[0]
    new_mean, _, _ = belief_update_diag(mean, np.ones(L), observation, score, eps, dt=dt, D=D, lr=lr,
                                        noise_scale=noise_scale, min_variance=min_variance,
                                        max_variance=np.inf)
    H = -(1.0 / noise_scale ** 2 + 1.0) * torch.eye(L, dtype=torch.float64)
    drift = -H - H.T + 2 * D * torch.eye(L, dtype=torch.float64)
    E = torch.matrix_exp(drift * dt)
    S = E @ torch.as_tensor(cov, dtype=torch.float64) @ E.T
    min_eig = max(min_variance, 1e-8)
    w, V = torch.linalg.eigh(S)
    w = torch.clamp(w, min=min_eig)
    if (w.max() / w.min()) > 1e6:
        w = w + w.mean() * 1e-6
    S = V @ torch.diag(w) @ V.T
    P = torch.linalg.inv(S + min_eig * torch.eye(L, dtype=torch.float64))
    return new_mean, S.numpy(), P.numpy()


# --------------------------------------------------------------------------
# DrQ-v2 visual encoder — encoder/visual_encoders.py, eval mode
# --------------------------------------------------------------------------

def spectral_weight(p: Params, name: str) -> torch.Tensor:
    """Weight of a spectral-normed conv in eval mode: weight_orig / (u . (W_mat v)) with the stored
    u, v and no power iteration (torch.nn.utils.spectral_norm, applied at visual_encoders.py:70-71).
    A conv built without spectral norm has a plain `weight`."""
    if name + ".weight_orig" not in p:
        return p[name + ".weight"]
    w = p[name + ".weight_orig"]
    sigma = torch.dot(p[name + ".weight_u"], torch.mv(w.reshape(w.shape[0], -1), p[name + ".weight_v"]))
    return w / sigma


def encoder_num_layers(p: Params) -> int:
    n = 0
    while f"norms.{n}.weight" in p:
        n += 1
    return n


def spatial_attention(p: Params, x: torch.Tensor) -> torch.Tensor:
    """visual_encoders.py:210-224: x + x * sigmoid(conv7x7([mean_c, max_c]) / temperature)."""
    pooled = torch.cat([x.mean(dim=1, keepdim=True), x.max(dim=1, keepdim=True)[0]], dim=1)
    logits = F.conv2d(pooled, p["attention.spatial_conv.weight"], p["attention.spatial_conv.bias"], padding=3)
    return x + x * torch.sigmoid(logits / p["attention.temperature"])


def encoder_canonical_input(x: torch.Tensor, base_channels: int, frame_stack: int) -> torch.Tensor:
    """Input conventions of visual_encoders.py:149-166 (frame axis folding, single-frame repeat,
    uint8 -> [0,1])."""
    if x.dim() == 5:
        b, t, c, h, w = x.shape
        x = x.reshape(b, t * c, h, w)
    elif x.dim() == 4:
        if x.shape[1] == base_channels and frame_stack > 1:
            x = x.repeat(1, frame_stack, 1, 1)
    elif x.dim() == 3:
        x = x.unsqueeze(0)
    if x.dtype == torch.uint8:
        x = x.float() / 255.0
    return x


def encoder_forward(p: Params, x: torch.Tensor, return_intermediates: bool = False):
    """DrQV2Encoder.forward, visual_encoders.py:136-189, on an already canonical float input
    [B, C*frames, H, W]: conv(3x3, stride 2 then 1, no bias) -> GroupNorm -> Mish per layer
    (Dropout2d is the identity in eval), spatial attention, flatten (channel-major), LayerNorm,
    Linear -> LayerNorm -> Mish -> Linear -> LayerNorm -> tanh."""
    inter = {}
    n_layers = encoder_num_layers(p)
    for i in range(n_layers):
        x = F.conv2d(x, spectral_weight(p, f"convs.{i}"), None, stride=2 if i == 0 else 1, padding=1)
        gw = p[f"norms.{i}.weight"]
        x = F.mish(F.group_norm(x, min(32, gw.shape[0] // 4), gw, p[f"norms.{i}.bias"], 1e-5))
        inter[f"conv{i}"] = x
    if "attention.spatial_conv.weight" in p:
        x = spatial_attention(p, x)
    x = layer_norm(p, "ln", x.reshape(x.shape[0], -1))
    inter["ln"] = x
    x = F.mish(layer_norm(p, "output_layers.1", linear(p, "output_layers.0", x)))
    x = torch.tanh(layer_norm(p, "output_layers.5", linear(p, "output_layers.4", x)))
    return (x, inter) if return_intermediates else x


# --------------------------------------------------------------------------
# lambda-returns — core/active_inference.py:638-707
# --------------------------------------------------------------------------

def lambda_returns(rewards: torch.Tensor, next_values: torch.Tensor, dones: torch.Tensor, gamma: float,
                   lam: float = 0.95, n_steps: int = 5, exclude_immediate_rewards: bool = False) -> torch.Tensor:
    """Dreamer-style lambda-returns along the batch axis, restated with the reference's arithmetic
    (0-dim fp32 tensor ops in the same order, Python-float weights) so the result is bit-identical:
    for every index the n-step returns n = 1..min(n_steps, B-1-i) (discounted reward prefix, the
    discount killed by a terminal flag, bootstrap from next_values[i+n] unless step i+n-1 was
    terminal) are mixed with (1-lam) lam^(n-1), the last taking lam^(N-1), and normalised by the
    weight sum + 1e-8 (:686-700); without any n-step return the one-step TD target is used (:701-705)."""
    B = rewards.shape[0]
    out = torch.zeros_like(rewards)
    alive = 1 - dones.float()
    for i in range(B):
        N = min(n_steps, B - 1 - i)
        if N <= 0:
            boot = gamma * alive[i] * next_values[i]
            out[i] = boot if exclude_immediate_rewards else rewards[i] + boot
            continue
        prefix, disc, rets = 0, 1.0, []
        for n in range(1, N + 1):
            k = n - 1
            if not (exclude_immediate_rewards and k == 0):
                prefix = prefix + disc * rewards[i + k]
            disc = disc * (gamma * alive[i + k])
            ret = prefix
            if i + n < B and not bool(dones[i + n - 1]):
                ret = ret + disc * next_values[i + n]
            rets.append(ret)
        total, wsum = 0, 0
        for j, ret in enumerate(rets[:-1]):
            wj = (1 - lam) * (lam ** j)
            total = total + wj * ret
            wsum += wj
        w_last = lam ** (len(rets) - 1)
        total = total + w_last * rets[-1]
        wsum += w_last
        out[i] = total / (wsum + 1e-8)
    return out
