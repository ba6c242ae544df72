"""The oracle restatement against the LIVE reference (fresh seeds, not the committed goldens).
Runs only where /root/reference exists (the build container); skipped on the GPU box.  CPU only."""
import pytest
import torch

from oracle import restatement as R
from oracle.harness import RecordingRNG, perturb_state_dict
from oracle.ref_import import import_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference is not present")


@pytest.fixture(scope="module")
def ref():
    import_reference()
    from active_inference_diffusion.configs.config import DiffusionConfig
    from active_inference_diffusion.core.diffusion import LatentDiffusionProcess
    from active_inference_diffusion.models.score_networks import LatentScoreNetwork
    return dict(DiffusionConfig=DiffusionConfig, LatentDiffusionProcess=LatentDiffusionProcess,
                LatentScoreNetwork=LatentScoreNetwork)


@pytest.mark.parametrize("seed,L,O,H,NB", [(3, 32, 17, 128, 2), (4, 64, 40, 192, 3)])
def test_score_forward_every_branch(ref, seed, L, O, H, NB):
    torch.manual_seed(seed)
    net = ref["LatentScoreNetwork"](L, O, H, num_layers=NB).eval()
    net.load_state_dict(perturb_state_dict(net.state_dict(), seed + 100, 0.1))
    p = {k: v.clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(seed)
    B = 9
    z, obs = torch.randn(B, L, generator=g), torch.randn(B, O, generator=g)
    for t in (torch.full((B,), 5.0), torch.zeros(B), torch.ones(B), torch.rand(B, generator=g),
              torch.tensor([0.0, 2.0] + [3.0] * (B - 2))):          # last: batch-global branch quirk
        with torch.no_grad():
            want = net(z, t, obs)
            got = R.score_forward(p, z, t, obs)
        assert torch.equal(got, want)


@pytest.mark.parametrize("sched,T", [("cosine", 6), ("linear", 5)])
def test_reverse_diffusion_with_recorded_draws(ref, sched, T):
    L, O, H, NB, B = 32, 17, 128, 2, 5
    torch.manual_seed(9)
    net = ref["LatentScoreNetwork"](L, O, H, num_layers=NB).eval()
    net.load_state_dict(perturb_state_dict(net.state_dict(), 77, 0.1))
    p = {k: v.clone() for k, v in net.state_dict().items()}
    dp = ref["LatentDiffusionProcess"](ref["DiffusionConfig"](num_diffusion_steps=T, beta_schedule=sched), latent_dim=L)
    obs = torch.randn(B, O)
    torch.manual_seed(21)
    with torch.no_grad(), RecordingRNG() as rec:
        traj = dp.generate_latent_trajectory(net, B, obs)
    zT, noise = rec.of("randn")[0], rec.of("randn_like")
    assert len(noise) == T - 1                                     # no draw at t == 0 (core/diffusion.py:232)
    sch = R.make_schedule(T, sched)
    for k in ("betas", "alphas_cumprod", "posterior_variance"):
        assert torch.equal(sch[k], getattr(dp, k))
    with torch.no_grad():
        got = R.generate_latent_trajectory(p, sch, zT, obs, noise)
    assert len(got) == len(traj)
    for a, b in zip(got, traj):
        assert torch.equal(a, b)


@pytest.mark.parametrize("shape,stack,F,nf,att,sn", [((3, 16, 20), 2, 12, 8, True, True), ((1, 9, 9), 1, 10, 16, False, False)])
def test_encoder_restatement_and_mirror_init(shape, stack, F, nf, att, sn):
    """DrQV2Encoder (encoder/visual_encoders.py): fresh seed, eval forward of the restatement, and the
    mirror's constructor against the reference's (same seed -> same state_dict)."""
    import_reference()
    from active_inference_diffusion.encoder.visual_encoders import DrQV2Encoder as RefEncoder
    from active_inference_diffusion_b200 import DrQV2Encoder
    torch.manual_seed(13)
    ref = RefEncoder(shape, feature_dim=F, frame_stack=stack, num_filters=nf, use_attention=att, use_spectral_norm=sn)
    torch.manual_seed(13)
    mine = DrQV2Encoder(shape, feature_dim=F, frame_stack=stack, num_filters=nf, use_attention=att, use_spectral_norm=sn)
    rsd, msd = ref.state_dict(), mine.state_dict()
    assert list(rsd.keys()) == list(msd.keys())
    for k in rsd:
        assert torch.equal(rsd[k], msd[k]), k
    assert mine.conv_out_dim == ref.conv_out_dim
    ref.eval()
    x = torch.randint(0, 256, (4, stack * shape[0], shape[1], shape[2]), dtype=torch.uint8)
    with torch.no_grad():
        want = ref(x)
        got = R.encoder_forward({k: v.clone() for k, v in rsd.items()}, R.encoder_canonical_input(x, shape[0], stack))
    assert torch.allclose(got, want, rtol=0, atol=2e-6)
