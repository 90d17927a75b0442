"""In-graph cost per launch of each ViT-B GEMM under forced tile plans (36 distinct weight sets: weights stream from HBM)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pipeedge_b200 import _lib, ops  # noqa: E402

SHAPES = {'qkv': (1576, 2304, 768, _lib.PE_EPI_F16), 'out': (1576, 768, 768, _lib.PE_EPI_F32),
          'fc1': (1576, 3072, 768, _lib.PE_EPI_GELU_F16), 'fc2': (1576, 768, 3072, _lib.PE_EPI_F32)}
PLANS = {'qkv': [256, 224, 192, 160, 128, 96], 'out': [256, 128, 96, 64], 'fc1': [256, 224, 192, 160, 128, 96],
         'fc2': [256, 192, 128, 96, 64]}
N = 36
only = [a for a in sys.argv[1:] if not a.startswith('--')] or list(SHAPES)
if '--stages' in sys.argv:   # pipeline-depth sensitivity at the chosen plans
    PLANS = {'qkv': [224], 'out': [96], 'fc1': [160], 'fc2': [96, 128]}
for name in only:
    m, n, k, epi = SHAPES[name]
    a = torch.randn(m, k, device='cuda').half()
    ws = [torch.randn(n, k, device='cuda').half() * 0.05 for _ in range(N)]
    bias = torch.randn(n, device='cuda')
    out = torch.empty(m, n, device='cuda', dtype=torch.float16 if epi in (_lib.PE_EPI_F16, _lib.PE_EPI_GELU_F16) else torch.float32)
    for bn in [0] + PLANS[name]:
        if bn:
            os.environ['PE_GEMM_FORCE'] = f"1,1,{bn}"
        else:
            os.environ.pop('PE_GEMM_FORCE', None)
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for i in range(3):
                ops.linear(a, ws[i], bias, epi, out=out, static_w=True)
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for i in range(N):
                    ops.linear(a, ws[i], bias, epi, out=out, static_w=True)
        torch.cuda.synchronize()
        for _ in range(3):
            g.replay()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        tiles = ((m + 127) // 128) * ((n + bn - 1) // bn) if bn else 0
        print(f"{name} bn {bn or 'auto'} tiles {tiles} stages<={os.environ.get('PE_GEMM_STAGES', '-')}: "
              f"{s.elapsed_time(e) * 1e3 / (10 * N):.2f} us per launch", flush=True)
