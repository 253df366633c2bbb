"""Raw GEMM probe of the wide path cores against torch (debug tool)."""
import sys, ctypes as C
import torch
sys.path.insert(0, ".")
from jaxili_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
torch.manual_seed(0)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(variant, core, M, N, K, structured=False):
    if variant == 0:   A = torch.randn(M, K, device=dev); B = torch.randn(K, N, device=dev); ref = A.double() @ B.double()
    elif variant == 1: A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); ref = A.double() @ B.double().T
    else:              A = torch.randn(K, M, device=dev); B = torch.randn(K, N, device=dev); ref = A.double().T @ B.double()
    Cm = torch.zeros(M, N, device=dev)
    zn = torch.zeros(N, device=dev); zmn = torch.zeros(M, N, device=dev); mask = torch.ones(M, N, dtype=torch.uint8, device=dev)
    rc = lib.mafb200_debug_gemm(variant, core, A.data_ptr(), B.data_ptr(), Cm.data_ptr(), M, N, K, zn.data_ptr(), zmn.data_ptr(), mask.data_ptr(), st)
    torch.cuda.synchronize()
    if rc: return "ERR " + lib.mafb200_last_error().decode()
    err = (Cm.double() - ref).abs().max().item() / ref.abs().max().item()
    return f"{err:.2e}"
for (M, N, K) in [(128, 128, 32), (128, 128, 64), (128, 128, 512), (256, 256, 96), (300, 64, 160), (200, 32, 512), (128, 512, 160)]:
    print((M, N, K), {v: {c: run(v, c, M, N, K) for c in (0, 1)} for v in (0, 1, 2)}, flush=True)
print("big K:", {v: run(v, 1, 640, 512, 2048) for v in (0, 1, 2)}, {v: run(v, 1, 640, 512, 8192) for v in (0, 1, 2)})
