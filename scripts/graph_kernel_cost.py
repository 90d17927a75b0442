"""Per-launch cost of each kernel type inside a CUDA graph of R back-to-back launches (kernel + boundary)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pipeedge_b200 import _lib, ops
M, H, I = 1576, 768, 3072
dev = 'cuda'
x32 = torch.randn(M, H, device=dev); g = torch.ones(H, device=dev); b = torch.zeros(H, device=dev)
a16 = torch.randn(M, H, device=dev).half(); i16 = torch.randn(M, I, device=dev).half()
wq = (torch.randn(3 * H, H, device=dev) * .02).half(); wo = (torch.randn(H, H, device=dev) * .02).half()
w1 = (torch.randn(I, H, device=dev) * .02).half(); w2 = (torch.randn(H, I, device=dev) * .02).half()
bq = torch.zeros(3 * H, device=dev); bo = torch.zeros(H, device=dev); b1 = torch.zeros(I, device=dev)
qkv = torch.randn(M, 3 * H, device=dev).half()
o_qkv = torch.empty(M, 3 * H, device=dev, dtype=torch.float16); o_h = torch.empty(M, H, device=dev)
o_i = torch.empty(M, I, device=dev, dtype=torch.float16)
from pipeedge_b200._lib import LIB, check
def ln(): check(LIB.pe_layernorm(x32.data_ptr(), g.data_ptr(), b.data_ptr(), 1e-12, None, a16.data_ptr(), M, H, torch.cuda.current_stream().cuda_stream))
def att(): check(LIB.pe_attention(qkv.data_ptr(), a16.data_ptr(), 8, 197, 12, 64, torch.cuda.current_stream().cuda_stream))
kernels = {
 'layernorm': ln,
 'gemm_qkv': lambda: ops.linear(a16, wq, bq, _lib.PE_EPI_F16, out=o_qkv),
 'attention': att,
 'gemm_out': lambda: ops.linear(a16, wo, bo, _lib.PE_EPI_F32, out=o_h),
 'gemm_fc1': lambda: ops.linear(a16, w1, b1, _lib.PE_EPI_GELU_F16, out=o_i),
 'gemm_fc2': lambda: ops.linear(i16, w2, bo, _lib.PE_EPI_F32, out=o_h),
 'empty_kernel': lambda: check(LIB.pe_cast_f32_to_f16(x32.data_ptr(), a16.data_ptr(), 4, torch.cuda.current_stream().cuda_stream)),
}
R = 48
for name, fn in kernels.items():
    for _ in range(3): fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(R): fn()
    graph.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): graph.replay()
    e.record(); torch.cuda.synchronize()
    print(f"{name}: {s.elapsed_time(e) / (10 * R) * 1e3:.2f} us per launch in-graph", flush=True)
