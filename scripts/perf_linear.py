"""Device-timed single GEMM through aid_linear (developer tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from active_inference_diffusion_b200 import _lib
M, N, K = [int(a) for a in sys.argv[1:4]]
act = int(sys.argv[4]) if len(sys.argv) > 4 else 0
via = int(sys.argv[5]) if len(sys.argv) > 5 else 0
x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
l = _lib.lib()
ws_bytes = l.aid_linear_workspace_bytes(M, N, K)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda"); y = torch.empty(M, N, device="cuda")
def run():
    _lib.check(l.aid_linear(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, N, K, act, via, ws.data_ptr(), ws_bytes, torch.cuda.current_stream().cuda_stream), "lin")
for _ in range(3): run()
torch.cuda.synchronize()
# time only the gemm: total minus pack is not separable here, so report total of (pack x + pack w + gemm [+unpack])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"AID_DEBUG={os.environ.get('AID_DEBUG','0')} M={M} N={N} K={K} act={act} via={via}: {ms*1e3:.1f} us/call  {2*M*N*K/ms/1e9:.1f} TFLOP/s (incl. packing)")
