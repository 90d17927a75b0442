"""Per-CTA timeline of the tcgen05 GEMM (clock64 stamps) for the model's shapes and a few tile plans."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pipeedge_b200 import _lib, ops  # noqa: E402
from pipeedge_b200._lib import LIB  # noqa: E402

SHAPES = {'qkv': (1576, 2304, 768, _lib.PE_EPI_F16), 'out': (1576, 768, 768, _lib.PE_EPI_F32),
          'fc1': (1576, 3072, 768, _lib.PE_EPI_GELU_F16), 'fc2': (1576, 768, 3072, _lib.PE_EPI_F32)}
PLANS = {'qkv': ['1,1,224'], 'out': ['1,1,96'], 'fc1': ['1,1,160', '1,1,256'], 'fc2': ['1,1,96', '1,1,256']}
if os.environ.get('TRACE_ONLY'):
    SHAPES = {k: v for k, v in SHAPES.items() if k in os.environ['TRACE_ONLY'].split(',')}
MODES = os.environ.get('TRACE_MODES', '0').split(',')   # bit0 no MMA, bit1 no TMA, bit3 no fence, bit4 plain arrive, bit5 no full wait
SLOTS = 32
trace = torch.zeros(148 * SLOTS, dtype=torch.int64, device='cuda')
for name, (m, n, k, epi) in SHAPES.items():
    a = torch.randn(m, k, device='cuda').half()
    w = torch.randn(n, k, device='cuda').half() * 0.05
    bias = torch.randn(n, device='cuda')
    resid = torch.randn(m, n, device='cuda')
    for plan, mode in [(pl, md) for pl in PLANS[name] for md in MODES]:
        os.environ['PE_GEMM_FORCE'] = plan
        os.environ['PE_GEMM_DEBUG_MODE'] = mode
        for _ in range(2):
            ops.linear(a, w, bias, epi, resid=resid, static_w=True)
        torch.cuda.synchronize()
        trace.zero_()
        LIB.pe_debug_gemm_trace(trace.data_ptr())
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.linear(a, w, bias, epi, resid=resid, static_w=True)
        e.record()
        torch.cuda.synchronize()
        LIB.pe_debug_gemm_trace(None)
        t = trace.view(148, SLOTS).cpu()
        used = t[:, 0] > 0
        t = t[used]
        rel = t[:, :7] - t[:, :1]     # clock64 is per SM: only differences within a CTA mean anything (int64 maths)
        names = ['start', 'setup', 'first_full', 'mma_issued', 'acc_ready', 'epi_done', 'exit']
        print(f"{name} plan {plan} mode {mode}: mainloop med {int((t[:, 3] - t[:, 2]).median())} cycles; ctas {int(used.sum())} event_us {s.elapsed_time(e) * 1e3:.1f}")
        print("   " + "  ".join(f"{nm}: med {int(rel[:, i].median())} max {int(rel[:, i].max())}" for i, nm in enumerate(names)))
