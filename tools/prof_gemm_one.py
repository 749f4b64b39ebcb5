"""One GEMM configuration, a few launches: target for `ncu -k regex:gemm -c 1 --launch-skip 3` (see profiles/)."""
import ctypes, os, sys
sys.path.insert(0, os.getcwd())
import torch
import depthmap_b200._lib as L
lib = L.load()
dev = torch.device('cuda')
M, N, K = (int(v) for v in sys.argv[1:4])
mode = sys.argv[4]
A = (torch.randn(M, K, device=dev) * 0.5).half()
W = (torch.randn(N, K, device=dev) * 0.05).half()
bias = torch.randn(N, device=dev)
gamma = torch.ones(N, device=dev)
d = L.GemmDesc()
d.M, d.N, d.K = M, N, K
C = torch.empty(M, N, dtype=torch.float16, device=dev)
X = torch.zeros(M, N, dtype=torch.float32, device=dev)
if mode == 'gelu':
    d.epi, d.act, d.C, d.ldc, d.bias = L.EPI_STORE_F16, 1, C.data_ptr(), N, bias.data_ptr()
elif mode == 'bias':
    d.epi, d.act, d.C, d.ldc, d.bias = L.EPI_STORE_F16, 0, C.data_ptr(), N, bias.data_ptr()
else:
    d.epi, d.X, d.ldx, d.bias, d.gamma = L.EPI_RESID_F32, X.data_ptr(), N, bias.data_ptr(), gamma.data_ptr()
for _ in range(6):
    L.check(lib.dm_gemm_ex(A.data_ptr(), K, W.data_ptr(), K, ctypes.byref(d), L.stream_ptr()))
torch.cuda.synchronize()
