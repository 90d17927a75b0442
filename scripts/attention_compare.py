"""In-graph time per launch of the two attention kernels at the ViT-B ubatch-8 shape (next-round experiment)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pipeedge_b200 import ops  # noqa: E402

batch, tokens, heads = 8, 197, 12
qkv = (torch.randn(batch * tokens, 3 * heads * 64, device='cuda') * 1.5).half()
for mode in ('mma.sync', 'tcgen05'):
    if mode == 'tcgen05':
        os.environ['PE_ATTN_TCGEN05'] = '1'
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            ops.attention(qkv, batch, tokens, heads)
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(24):
                ops.attention(qkv, batch, tokens, heads)
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    print(f"{mode}: {s.elapsed_time(e) * 1e3 / 240:.2f} us per launch")
