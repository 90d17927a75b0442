"""Per-forward time of the CPU port in a fresh process (why a short CPU arm and a long one disagree)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pipeedge_b200.synth import MODEL_SPECS
if len(sys.argv) > 1 and sys.argv[1] == 'cuda':
    torch.cuda.init(); _p = torch.empty(1 << 20).pin_memory(); _d = torch.zeros(8, device='cuda'); torch.cuda.synchronize()
    print('cuda initialised first', flush=True)
spec = MODEL_SPECS['google/vit-base-patch16-224']
fwd, cores, cal = bench.cpu_forward_timer(spec, 8, 0)
print('threads', cores, cal, flush=True)
ts = []
for i in range(40):
    t0 = time.perf_counter(); fwd(); ts.append(time.perf_counter() - t0)
print(' '.join(f'{t:.3f}' for t in ts))
