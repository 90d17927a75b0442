"""Sweep (CM, CN, BN) for the model's GEMM shapes with PE_GEMM_FORCE and print microseconds per launch
(CUDA events over back-to-back launches rotating over enough weight copies to defeat L2 reuse of W)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pipeedge_b200 import _lib, ops  # noqa: E402

SHAPES = {'qkv': (1576, 2304, 768, _lib.PE_EPI_F16), 'out': (1576, 768, 768, _lib.PE_EPI_RESID_F32),
          'fc1': (1576, 3072, 768, _lib.PE_EPI_GELU_F16), 'fc2': (1576, 768, 3072, _lib.PE_EPI_RESID_F32)}
if len(sys.argv) > 1 and sys.argv[1] == 'big':
    SHAPES = {'bert_qkv': (4096, 2304, 768, _lib.PE_EPI_F16), 'bert_fc2': (4096, 768, 3072, _lib.PE_EPI_RESID_F32),
              'deit_fc1': (6336, 3072, 768, _lib.PE_EPI_GELU_F16), 'vitl_fc1': (3152, 4096, 1024, _lib.PE_EPI_GELU_F16)}


def bench(m, n, k, epi, force):
    if force:
        os.environ['PE_GEMM_FORCE'] = force
    else:
        os.environ.pop('PE_GEMM_FORCE', None)
    copies = max(2, int(160e6 // (n * k * 2)))
    a = torch.randn(m, k, device='cuda').half()
    ws = [torch.randn(n, k, device='cuda').half() * 0.05 for _ in range(copies)]
    bias = torch.randn(n, device='cuda')
    resid = torch.randn(m, n, device='cuda')
    out = torch.empty(m, n, device='cuda', dtype=torch.float16 if epi in (0, 1) else torch.float32)
    for i in range(3):
        ops.linear(a, ws[i % copies], bias, epi, resid=resid, out=out)
    torch.cuda.synchronize()
    iters = 40
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(iters):
        ops.linear(a, ws[i % copies], bias, epi, resid=resid, out=out)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


for name, (m, n, k, epi) in SHAPES.items():
    base = bench(m, n, k, epi, None)
    print(f"{name}: heuristic {base:.1f} us ({2 * m * n * k / base / 1e6:.0f} TF/s)", flush=True)
    res = []
    for cm, cn in ((1, 1), (2, 1), (1, 2), (2, 2), (4, 1)):
        for bn in (64, 96, 128, 160, 192, 224, 256):
            if (bn // cm) % 8 or bn % cm:
                continue
            try:
                us = bench(m, n, k, epi, f"{cm},{cn},{bn}")
            except Exception as exc:  # noqa
                print('  fail', cm, cn, bn, str(exc)[:80])
                continue
            res.append((us, cm, cn, bn))
    res.sort()
    for us, cm, cn, bn in res[:6]:
        print(f"   {us:7.1f} us  cm={cm} cn={cn} bn={bn}  ({2 * m * n * k / us / 1e6:.0f} TF/s)")
    print("   worst:", res[-1], flush=True)
