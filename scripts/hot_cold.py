"""Per-GEMM cost inside a CUDA graph with weights streaming from HBM (36 distinct sets, > L2) vs L2-resident
(one set), for the ViT-B ubatch-8 shapes. Run with PE_NO_PDL=1 for the plain-launch comparison."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pipeedge_b200 import _lib, ops  # noqa: E402

SHAPES = {'qkv': (1576, 2304, 768, _lib.PE_EPI_F16), 'out': (1576, 768, 768, _lib.PE_EPI_F32),
          'fc1': (1576, 3072, 768, _lib.PE_EPI_GELU_F16), 'fc2': (1576, 768, 3072, _lib.PE_EPI_F32)}
N = 36
print(f"PDL {'off' if os.environ.get('PE_NO_PDL') == '1' else 'on'}")
for name, (m, n, k, epi) in SHAPES.items():
    a = torch.randn(m, k, device='cuda').half()
    ws = [torch.randn(n, k, device='cuda').half() * 0.05 for _ in range(N)]
    bias = torch.randn(n, device='cuda')
    out = torch.empty(m, n, device='cuda', dtype=torch.float16 if epi in (_lib.PE_EPI_F16, _lib.PE_EPI_GELU_F16) else torch.float32)
    res = {}
    for mode in ('cold', 'hot'):
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for i in range(3):
                ops.linear(a, ws[i], bias, epi, out=out, static_w=True)
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for i in range(N):
                    ops.linear(a, ws[i if mode == 'cold' else 0], bias, epi, out=out, static_w=True)
        torch.cuda.synchronize()
        for _ in range(3):
            g.replay()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        res[mode] = s.elapsed_time(e) * 1e3 / (10 * N)
    print(f"{name}: cold {res['cold']:.2f} us  hot {res['hot']:.2f} us per launch")
