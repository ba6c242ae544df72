"""CPU: the hand-derived first/second-order backward of the score-net trunk (oracle/manual_score_grad.py,
the specification of csrc/train.inc) against autograd on the oracle restatement of
core/active_inference.py:584-606,709-729 in fp64: score, input gradient g, and the gradient of
L = mean_b w_b |s - target|^2 + 0.1 mean_b (|g|_2 - 1)^2 w.r.t. every parameter and w.r.t. z (the
penalty's input is detached, so dL/dz carries the score-matching stream only)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import manual_score_grad as M
from oracle import restatement as R
from tests.util import gen, make_score_net, rel_l2


def trunk_inputs(p, t, obs, B, NB):
    cond, tw = R.score_conditioning(p, t, obs, B)
    names = [f"transformer_blocks.{i}.{n}" for i in range(NB) for n in ("norm1", "norm2")] + ["norm_final"]
    mods = []
    for n_ in names:
        m = F.linear(F.silu(cond), p[n_ + ".adaLN_modulation.1.weight"], p[n_ + ".adaLN_modulation.1.bias"])
        scale, shift = m.chunk(2, dim=-1)
        mods.append((1 + scale, shift))
    return mods, tw


@pytest.mark.parametrize("continuous", [True, False])
def test_manual_backward_matches_autograd_fp64(continuous):
    L, O, H, NB, B = 16, 5, 32, 2, 9
    _, params = make_score_net(L, O, H, NB)
    mk = lambda: {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in params.items()}
    p, p2 = mk(), mk()
    g = gen(4)
    z = torch.randn(B, L, generator=g, dtype=torch.float64)
    obs = torch.randn(B, O, generator=g, dtype=torch.float64)
    t = torch.rand(B, generator=g, dtype=torch.float64) if continuous else torch.full((B,), 7.0, dtype=torch.float64)
    tgt = torch.randn(B, L, generator=g, dtype=torch.float64)
    w = torch.rand(B, generator=g, dtype=torch.float64)

    def loss_of(s, gx):
        return (w * ((s - tgt) ** 2).sum(1)).mean() + 0.1 * ((gx.norm(2, dim=1) - 1) ** 2).mean()

    zz = z.clone().requires_grad_(True)
    s_ref = R.score_forward(p, zz, t, obs)
    x = z.clone().requires_grad_(True)
    (g_ref,) = torch.autograd.grad(R.score_forward(p, x, t, obs).sum(), x, create_graph=True)
    loss_of(s_ref, g_ref).backward()

    mods, tw = trunk_inputs(p2, t, obs, B, NB)
    T = M.Trunk.from_state_dict(p2)
    det = lambda v: v.detach()
    Td = M.Trunk(det(T.W_lp), det(T.b_lp), [{k: det(v) for k, v in b.items()} for b in T.blocks], det(T.Wo0),
                 det(T.bo0), det(T.Wo2), det(T.mult))
    s, gx, sv = M.forward_and_vjp(Td, z, [(det(a), det(b)) for a, b in mods], None if tw is None else det(tw))
    assert rel_l2(s, s_ref.detach()) < 1e-12 and rel_l2(gx, g_ref.detach()) < 1e-12
    s_l, g_l = s.clone().requires_grad_(True), gx.clone().requires_grad_(True)
    loss_of(s_l, g_l).backward()
    G = M.backward(Td, sv, s_l.grad, g_l.grad)
    outs = [T.W_lp, T.b_lp, T.Wo0, T.bo0, T.Wo2, T.mult]
    grads = [G["W_lp"], G["b_lp"], G["Wo0"], G["bo0"], G["Wo2"], G["mult"]]
    for b, gb in zip(T.blocks, G["blocks"]):
        for k in ("Wf", "bf", "W1", "b1", "W2", "b2"):
            outs.append(b[k]); grads.append(gb[k])
    for (a, b), (ga, gb_) in zip(mods, G["mods"]):
        outs += [a, b]; grads += [ga, gb_]
    torch.autograd.backward(outs, grads)          # fold W_o W_v and the conditioning path: autograd's chain rule
    assert rel_l2(G["z"], zz.grad) < 1e-10
    checked = 0
    for k, v in p.items():
        if not v.is_floating_point() or v.grad is None or float(v.grad.abs().max()) == 0.0:
            continue
        assert p2[k].grad is not None, k
        assert rel_l2(p2[k].grad, v.grad) < 1e-9, (k, rel_l2(p2[k].grad, v.grad))
        checked += 1
    assert checked >= 30
