"""In-graph time per launch of the two attention kernels at BASELINE shapes. The library reads PE_ATTN_TCGEN05 once per
process, so each kernel is timed in its own child process:  python scripts/attention_compare.py"""
import os
import subprocess
import sys

CHILD = """
import os, sys, torch
sys.path.insert(0, sys.argv[1])
from pipeedge_b200 import ops
for batch, tokens, heads in ((8, 197, 12), (16, 197, 16), (32, 128, 12), (32, 198, 12)):
    qkv = (torch.randn(batch * tokens, 3 * heads * 64, device='cuda') * 1.5).half()
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
    us = s.elapsed_time(e) * 1e3 / 240
    flops = 4.0 * batch * tokens * tokens * heads * 64
    print(f"{os.environ.get('PE_ATTN_TCGEN05', '0')} batch {batch} S {tokens} heads {heads}: {us:.2f} us per launch, "
          f"{flops / us / 1e6:.1f} TFLOP/s", flush=True)
"""
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for mode in ('0', '1'):
    print('mma.sync' if mode == '0' else 'tcgen05')
    subprocess.run([sys.executable, '-c', CHILD, root], env=dict(os.environ, PE_ATTN_TCGEN05=mode), check=False)
