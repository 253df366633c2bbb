"""Large raw GEMM for profiling the tcgen05 core (debug tool): python tests/gemm_prof.py variant M N K"""
import sys, ctypes as C
import torch
sys.path.insert(0, ".")
from jaxili_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
variant, M, N, K = [int(a) for a in sys.argv[1:5]]
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
if variant == 0:   A = torch.randn(M, K, device=dev); B = torch.randn(K, N, device=dev)
elif variant == 1: A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev)
else:              A = torch.randn(K, M, device=dev); B = torch.randn(K, N, device=dev)
Cm = torch.zeros(M, N, device=dev); zn = torch.zeros(N, device=dev); zmn = torch.zeros(M, N, device=dev)
mask = torch.ones(M, N, dtype=torch.uint8, device=dev)
for core in (1,):
    for _ in range(3):
        lib.mafb200_debug_gemm(variant, core, A.data_ptr(), B.data_ptr(), Cm.data_ptr(), M, N, K, zn.data_ptr(), zmn.data_ptr(), mask.data_ptr(), st)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.mafb200_debug_gemm(variant, core, A.data_ptr(), B.data_ptr(), Cm.data_ptr(), M, N, K, zn.data_ptr(), zmn.data_ptr(), mask.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"variant {variant} core {core} {M}x{N}x{K}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s")
