"""TEST INFRASTRUCTURE — hand-derived first- and second-order backpropagation through the trunk of
the score network, in plain torch.  It is the specification the native training kernels
(csrc/train.inc: aid_dsm_forward / aid_gp_forward_backward / aid_dsm_backward) are written against
and the checker for their intermediate tensors; only tests may import it.

What the training loss needs from the score net (core/active_inference.py:584-606, 709-729):

    s   = s_theta(z, t, o)                                  score-matching term
    g   = d(sum s)/dz   (create_graph=True)                 gradient penalty  mean((|g|_2 - 1)^2)

and the gradients of a scalar loss L(s, g) w.r.t. the parameters, the modulations (conditioning
path) and z.  The reference evaluates the network twice (once per term); both evaluations see the
same numbers, so ONE forward serves both here (the penalty's input is detached: its gradient must
not reach z -- the two cotangent streams "D" (score matching) and "P" (penalty) are therefore kept
apart on the input-gradient side and summed only for the weight gradients).

Passes (F = FLOPs of one trunk forward; autograd on the reference formulation spends 9 F):
  1. forward                                  F      saves LayerNorm stats, pre-activations, operands
  2. VJP with cotangent 1 ("c stream")        F      g, saves the stream's cotangents per layer
  3. adjoint of the VJP ("c^ stream")        2F      runs in FORWARD direction; weight gradients of
                                                     the VJP's GEMMs + second-order terms (LayerNorm,
                                                     GELU'', SiLU'') injected into the forward graph
  4. backward through the forward graph      3F      D and P input-gradient streams (stacked rows in the
                                                     kernels) + ONE weight-gradient GEMM per layer on D+P
                                              = 7F

Network (models/score_networks.py:151-171, attention folded to W_f = W_o W_v):
    h = z W_lp^T + b_lp
    per block:  h += adaLN1(h) W_f^T + b_f ;  h += GELU(adaLN2(h) W_1^T + b_1) W_2^T + b_2
    r = SiLU(adaLN_f(h) W_o0^T + b_o0) W_o2^T ;  s = clamp(r, +-10) * mult * tw
    adaLN(h) = LN(h) * s1 + sh        (s1 = 1 + scale)
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


def ln_stats(h: torch.Tensor, eps: float = 1e-5):
    mu = h.mean(-1, keepdim=True)
    var = ((h - mu) ** 2).mean(-1, keepdim=True)
    rstd = torch.rsqrt(var + eps)
    return (h - mu) * rstd, rstd


def ln_bwd(c: torch.Tensor, n: torch.Tensor, rstd: torch.Tensor) -> torch.Tensor:
    """LayerNorm input gradient for a cotangent c on the normalised output (no affine).  As a linear
    map of c it is symmetric, so the adjoint of the VJP stream uses the same function."""
    return rstd * (c - c.mean(-1, keepdim=True) - n * (c * n).mean(-1, keepdim=True))


def ln_second(a: torch.Tensor, c: torch.Tensor, n: torch.Tensor, rstd: torch.Tensor) -> torch.Tensor:
    """d/dh <a, ln_bwd(c; h)>: the LayerNorm second-order term (a = adjoint of the VJP stream after the
    node, c = the VJP stream's cotangent on the normalised output)."""
    H = n.shape[-1]
    m_cn = (c * n).mean(-1, keepdim=True)
    m_an = (a * n).mean(-1, keepdim=True)
    phi0 = (a * c).sum(-1, keepdim=True) - H * a.mean(-1, keepdim=True) * c.mean(-1, keepdim=True) - H * m_an * m_cn
    k = -rstd * (a * m_cn + c * m_an)
    return ln_bwd(k, n, rstd) - phi0 * rstd * rstd * n / H


def gelu_d1(u):
    return 0.5 * (1 + torch.erf(u / math.sqrt(2))) + u * torch.exp(-0.5 * u * u) / math.sqrt(2 * math.pi)


def gelu_d2(u):
    return torch.exp(-0.5 * u * u) / math.sqrt(2 * math.pi) * (2 - u * u)


def silu_d1(p):
    s = torch.sigmoid(p)
    return s * (1 + p * (1 - s))


def silu_d2(p):
    s = torch.sigmoid(p)
    return s * (1 - s) * (2 + p * (1 - 2 * s))


class Trunk:
    """Weights of the trunk as plain tensors.  blocks[i] = dict(Wf, bf, W1, b1, W2, b2)."""

    def __init__(self, W_lp, b_lp, blocks: List[Dict[str, torch.Tensor]], Wo0, bo0, Wo2, mult):
        self.W_lp, self.b_lp, self.blocks, self.Wo0, self.bo0, self.Wo2, self.mult = W_lp, b_lp, blocks, Wo0, bo0, Wo2, mult

    @staticmethod
    def from_state_dict(p: Dict[str, torch.Tensor]) -> "Trunk":
        nb = 0
        while f"transformer_blocks.{nb}.mlp.0.weight" in p:
            nb += 1
        H = p["latent_proj.weight"].shape[0]
        blocks = []
        for i in range(nb):
            pre = f"transformer_blocks.{i}."
            w_v, b_v = p[pre + "attention.in_proj_weight"][2 * H:], p[pre + "attention.in_proj_bias"][2 * H:]
            w_o, b_o = p[pre + "attention.out_proj.weight"], p[pre + "attention.out_proj.bias"]
            blocks.append(dict(Wf=w_o @ w_v, bf=w_o @ b_v + b_o, W1=p[pre + "mlp.0.weight"], b1=p[pre + "mlp.0.bias"],
                               W2=p[pre + "mlp.2.weight"], b2=p[pre + "mlp.2.bias"]))
        return Trunk(p["latent_proj.weight"], p["latent_proj.bias"], blocks, p["output_proj.0.weight"],
                     p["output_proj.0.bias"], p["output_proj.2.weight"], p["output_proj_multiplier"]
                     if "output_proj_multiplier" in p else p["output_multiplier"])


def forward_and_vjp(T: Trunk, z, mods, tw):
    """Passes 1 and 2.  mods: list of (s1, sh) per LayerNorm site in forward order (2 per block + final);
    tw [B,1] or None.  Returns (s, g, saved)."""
    sv = {"z": z, "mods": mods, "tw": tw, "blk": []}
    h = z @ T.W_lp.t() + T.b_lp
    for i, b in enumerate(T.blocks):
        s1, sh = mods[2 * i]
        n1, r1 = ln_stats(h)
        a1 = n1 * s1 + sh
        h1 = h + a1 @ b["Wf"].t() + b["bf"]
        s1b, shb = mods[2 * i + 1]
        n2, r2 = ln_stats(h1)
        a2 = n2 * s1b + shb
        u = a2 @ b["W1"].t() + b["b1"]
        v = F.gelu(u)
        h2 = h1 + v @ b["W2"].t() + b["b2"]
        sv["blk"].append(dict(n1=n1, r1=r1, a1=a1, n2=n2, r2=r2, a2=a2, u=u, v=v))
        h = h2
    s1f, shf = mods[-1]
    nf, rf = ln_stats(h)
    af = nf * s1f + shf
    p = af @ T.Wo0.t() + T.bo0
    q = F.silu(p)
    r = q @ T.Wo2.t()
    mask = ((r >= -10) & (r <= 10)).to(r.dtype)
    scale = T.mult * (tw if tw is not None else 1.0)
    s = torch.clamp(r, -10, 10) * scale
    sv.update(nf=nf, rf=rf, af=af, p=p, q=q, r=r, mask=mask, scale=scale)
    # ---- pass 2: cotangent 1 on every element of s, back to z
    c_r = mask * scale
    c_q = c_r @ T.Wo2
    c_p = c_q * silu_d1(p)
    c_af = c_p @ T.Wo0
    c_h = ln_bwd(c_af * s1f, nf, rf)
    sv.update(c_r=c_r, c_q=c_q, c_p=c_p, c_af=c_af, c_h_top=c_h)
    for i in reversed(range(len(T.blocks))):
        b, k = T.blocks[i], sv["blk"][i]
        s1, _ = mods[2 * i]
        s1b, _ = mods[2 * i + 1]
        k["c_hB"] = c_h                                   # residual cotangent entering the MLP node
        c_v = c_h @ b["W2"]
        c_u = c_v * gelu_d1(k["u"])
        c_a2 = c_u @ b["W1"]
        c_h = c_h + ln_bwd(c_a2 * s1b, k["n2"], k["r2"])
        k["c_hA"] = c_h                                   # ... entering the attention node
        c_a1 = c_h @ b["Wf"]
        c_h = c_h + ln_bwd(c_a1 * s1, k["n1"], k["r1"])
        k.update(c_v=c_v, c_u=c_u, c_a2=c_a2, c_a1=c_a1)
    sv["c_h0"] = c_h
    g = c_h @ T.W_lp
    return s, g, sv


def backward(T: Trunk, sv, s_bar, g_bar):
    """Passes 3 and 4.  Returns dict(z=z_bar (D stream only), mods=[(ds1, dsh)], and the weight
    gradients keyed like the Trunk fields: W_lp, b_lp, blocks[i][...], Wo0, bo0, Wo2, mult)."""
    mods, nb = sv["mods"], len(T.blocks)
    G = dict(blocks=[dict() for _ in range(nb)], mods=[None] * len(mods))
    gp_h = [dict() for _ in range(nb)]      # second-order terms injected into the forward graph
    # ---- pass 3: adjoint of the VJP, forward direction
    G["W_lp"] = sv["c_h0"].t() @ g_bar
    a = g_bar @ T.W_lp.t()                                # adjoint of the residual cotangent c_h
    for i, b in enumerate(T.blocks):
        k = sv["blk"][i]
        s1, _ = mods[2 * i]
        s1b, _ = mods[2 * i + 1]
        # attention node: c_h' = c_h + ln_bwd((c_h Wf) * s1)
        a_n = ln_bwd(a, k["n1"], k["r1"])
        gp_h[i]["h_in"] = ln_second(a, k["c_a1"] * s1, k["n1"], k["r1"])
        ds1_attn = a_n * k["c_a1"]
        a_a = a_n * s1
        G["blocks"][i]["Wf"] = k["c_hA"].t() @ a_a
        a = a + a_a @ b["Wf"].t()
        # MLP node: c_h' = c_h + ln_bwd(((c_h W2 * gelu'(u)) W1) * s1b)
        a_n = ln_bwd(a, k["n2"], k["r2"])
        gp_h[i]["h_mid"] = ln_second(a, k["c_a2"] * s1b, k["n2"], k["r2"])
        ds1_mlp = a_n * k["c_a2"]
        a_a2 = a_n * s1b
        G["blocks"][i]["W1"] = k["c_u"].t() @ a_a2
        a_u = a_a2 @ b["W1"].t()
        gp_h[i]["u"] = a_u * k["c_v"] * gelu_d2(k["u"])
        a_v = a_u * gelu_d1(k["u"])
        G["blocks"][i]["W2"] = k["c_hB"].t() @ a_v
        a = a + a_v @ b["W2"].t()
        G["mods"][2 * i] = [ds1_attn, None]
        G["mods"][2 * i + 1] = [ds1_mlp, None]
    s1f, _ = mods[-1]
    a_n = ln_bwd(a, sv["nf"], sv["rf"])
    gp_hf = ln_second(a, sv["c_af"] * s1f, sv["nf"], sv["rf"])
    G["mods"][-1] = [a_n * sv["c_af"], None]
    a_af = a_n * s1f
    G["Wo0"] = sv["c_p"].t() @ a_af
    a_p = a_af @ T.Wo0.t()
    gp_p = a_p * sv["c_q"] * silu_d2(sv["p"])
    a_q = a_p * silu_d1(sv["p"])
    G["Wo2"] = sv["c_r"].t() @ a_q
    a_r = a_q @ T.Wo2.t()
    tw = sv["tw"] if sv["tw"] is not None else 1.0
    G["mult"] = (a_r * sv["mask"] * tw).sum().reshape(1)
    # ---- pass 4: backward through the forward graph; D = score-matching stream, P = penalty stream
    G["mult"] = G["mult"] + (s_bar * torch.clamp(sv["r"], -10, 10) * tw).sum().reshape(1)
    rD = s_bar * sv["mask"] * sv["scale"]
    G["Wo2"] = G["Wo2"] + rD.t() @ sv["q"]
    pD = (rD @ T.Wo2) * silu_d1(sv["p"])
    pP = gp_p
    G["Wo0"] = G["Wo0"] + (pD + pP).t() @ sv["af"]
    G["bo0"] = (pD + pP).sum(0)
    afD, afP = pD @ T.Wo0, pP @ T.Wo0
    G["mods"][-1][0] = G["mods"][-1][0] + (afD + afP) * sv["nf"]
    G["mods"][-1][1] = afD + afP
    hD = ln_bwd(afD * s1f, sv["nf"], sv["rf"])
    hP = ln_bwd(afP * s1f, sv["nf"], sv["rf"]) + gp_hf
    for i in reversed(range(nb)):
        b, k, gb = T.blocks[i], sv["blk"][i], G["blocks"][i]
        s1, _ = mods[2 * i]
        s1b, _ = mods[2 * i + 1]
        # h2 = h1 + gelu(a2 W1^T + b1) W2^T + b2
        gb["W2"] = gb["W2"] + (hD + hP).t() @ k["v"]
        gb["b2"] = (hD + hP).sum(0)
        uD = (hD @ b["W2"]) * gelu_d1(k["u"])
        uP = (hP @ b["W2"]) * gelu_d1(k["u"]) + gp_h[i]["u"]
        gb["W1"] = gb["W1"] + (uD + uP).t() @ k["a2"]
        gb["b1"] = (uD + uP).sum(0)
        aD, aP = uD @ b["W1"], uP @ b["W1"]
        G["mods"][2 * i + 1][0] = G["mods"][2 * i + 1][0] + (aD + aP) * k["n2"]
        G["mods"][2 * i + 1][1] = aD + aP
        hD = hD + ln_bwd(aD * s1b, k["n2"], k["r2"])
        hP = hP + ln_bwd(aP * s1b, k["n2"], k["r2"]) + gp_h[i]["h_mid"]
        # h1 = h + a1 Wf^T + bf
        gb["Wf"] = gb["Wf"] + (hD + hP).t() @ k["a1"]
        gb["bf"] = (hD + hP).sum(0)
        aD, aP = hD @ b["Wf"], hP @ b["Wf"]
        G["mods"][2 * i][0] = G["mods"][2 * i][0] + (aD + aP) * k["n1"]
        G["mods"][2 * i][1] = aD + aP
        hD = hD + ln_bwd(aD * s1, k["n1"], k["r1"])
        hP = hP + ln_bwd(aP * s1, k["n1"], k["r1"]) + gp_h[i]["h_in"]
    G["W_lp"] = G["W_lp"] + (hD + hP).t() @ sv["z"]
    G["b_lp"] = (hD + hP).sum(0)
    G["z"] = hD @ T.W_lp                                  # the penalty's input is detached
    G["z_penalty"] = hP @ T.W_lp                          # (kept for the checker only)
    G["dbg"] = dict(gp_h=gp_h, gp_hf=gp_hf, gp_p=gp_p)    # second-order terms (kernel intermediates)
    return G
