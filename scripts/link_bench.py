"""Device time of a link's send + receive pair on a loop-back link under different synchronisation settings
(PE_LINK_SYNC, PE_LINK_GRID_CAP), each in its own child process; 24 pairs per CUDA graph, CUDA events."""
import os
import subprocess
import sys

CHILD = """
import ctypes, os, sys, torch
sys.path.insert(0, sys.argv[1])
from pipeedge_b200 import _lib
from pipeedge_b200._lib import LIB, check
dev = torch.device('cuda', 0)
for label, shape, bit, with_b in (('raw 8x197x768', (8, 197, 768), 0, False), ('raw+add 8x197x768', (8, 197, 768), 0, True),
                                  ('q8 32x198x768 a+b', (32, 198, 768), 8, True), ('raw 32x128x768', (32, 128, 768), 0, False)):
    a = torch.randn(shape, device=dev)
    b = torch.randn(shape, device=dev)
    dst = torch.empty_like(a)
    items, n = shape[0], shape[1] * shape[2]
    h = ctypes.c_void_p()
    check(LIB.pe_link_open_local(a.numel() * 4 + 4096, 4, bit, ctypes.byref(h)))
    side = torch.cuda.Stream()
    def pair():
        s = torch.cuda.current_stream().cuda_stream
        check(LIB.pe_link_put(h, a.data_ptr(), b.data_ptr() if with_b else None, n, None, None, 0, items, bit,
                              _lib.PE_CLAMP_AUTO if bit else 0, s))
        check(LIB.pe_link_get(h, dst.data_ptr(), None, items, n, 0, s))
    with torch.cuda.stream(side):
        for _ in range(3):
            pair()
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(24):
                pair()
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
    print(f"  {label}: {us:.2f} us per put + wait + get", flush=True)
    LIB.pe_link_close(h)
"""
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get('PE_ONLY_DEFAULT'):     # the library's defaults only
    print("library defaults (PE_LINK_SYNC=2, copies on half the SMs)", flush=True)
    env = {k: v for k, v in os.environ.items() if k not in ('PE_LINK_SYNC', 'PE_LINK_GRID_CAP')}
    subprocess.run([sys.executable, '-c', CHILD, root], env=env, check=False)
    sys.exit(0)
for sync in ('0', '1', '2', '3'):
    for cap in ('0', '74', '37'):
        print(f"PE_LINK_SYNC={sync} PE_LINK_GRID_CAP={cap}", flush=True)
        subprocess.run([sys.executable, '-c', CHILD, root], env=dict(os.environ, PE_LINK_SYNC=sync, PE_LINK_GRID_CAP=cap),
                       check=False)
