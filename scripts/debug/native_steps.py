"""Debug driver: the four passes of the native training path one call at a time with syncs in between."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from active_inference_diffusion_b200 import _lib, autograd_path as AP, train_native as TN
from tests.util import make_score_net, gen

operand = sys.argv[1] if len(sys.argv) > 1 else "f16"
L, O, H, NB, B = 32, 17, 128, 2, 200
net, params = make_score_net(L, O, H, NB, device="cuda")
dev = torch.device("cuda", 0)
g = gen(5)
z = torch.randn(B, L, generator=g).cuda(); obs = torch.randn(B, O, generator=g).cuda(); t = torch.rand(B, generator=g).cuda()
with torch.no_grad(), AP.precision("bf16x3"):
    cond, tw = AP.score_cond_embedding(net, t, obs, B, True)
    folds = AP.fold_attention(net)
    plist = [p.detach().contiguous() for p in TN.trunk_parameters(net, folds)]
l = _lib.lib(operand)
d = TN._dims(net)
pb, wb = l.aid_train_packed_bytes(ctypes.byref(d)), l.aid_train_workspace_bytes(ctypes.byref(d), B)
packed = torch.empty(pb, dtype=torch.uint8, device=dev); ws = torch.empty(wb, dtype=torch.uint8, device=dev)
s = torch.empty(B, L, device=dev); gg = torch.empty(B, L, device=dev)
tw1 = tw.reshape(-1).contiguous()
st = _lib.stream_ptr(dev)
def step(name, rc):
    torch.cuda.synchronize()
    print(name, rc, l.aid_last_error() if rc else "", flush=True)
step("pack", l.aid_train_pack(ctypes.byref(d), TN._table(plist), len(plist), packed.data_ptr(), pb, st))
step("fwd", l.aid_dsm_forward(ctypes.byref(d), packed.data_ptr(), ws.data_ptr(), wb, B, z.data_ptr(), cond.data_ptr(), tw1.data_ptr(), s.data_ptr(), st))
step("gp0", l.aid_gp_forward_backward(ctypes.byref(d), packed.data_ptr(), ws.data_ptr(), wb, B, 0, tw1.data_ptr(), gg.data_ptr(), None, None, None, st))
print("s", float(s.abs().mean()), "g", float(gg.abs().mean()), flush=True)
grads = [torch.zeros_like(p) for p in plist]
table = TN._table(grads)
sb = torch.randn(B, L, device=dev) * 1e-3; gb = torch.randn(B, L, device=dev) * 1e-4
step("gp1", l.aid_gp_forward_backward(ctypes.byref(d), packed.data_ptr(), ws.data_ptr(), wb, B, 1, tw1.data_ptr(), None, gb.data_ptr(), sb.data_ptr(), table, st))
dz = torch.empty(B, L, device=dev); dc = torch.empty(B, H, device=dev)
for stage in range(NB + 2):
    step(f"bwd stage {stage}", l.aid_dsm_backward(ctypes.byref(d), packed.data_ptr(), ws.data_ptr(), wb, B, sb.data_ptr(), tw1.data_ptr(), cond.data_ptr(), 1, table, dz.data_ptr(), dc.data_ptr(), stage, stage + 1, st))
print("grads", [float(x.abs().mean()) for x in grads][:8], float(dz.abs().mean()), float(dc.abs().mean()))
