"""Is the resident-input loop host-bound? Host enqueue time vs CUDA-event time per step, for the whole shard
forward and for the captured stage graph alone."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pipeedge_b200.synth import MODEL_SPECS, synth_input, synth_weights
spec = MODEL_SPECS['google/vit-base-patch16-224']
shard = bench.make_shard(spec, synth_weights(spec, 0), 1, spec.layers)
shard.use_cuda_graph = True
xs = [synth_input(spec, 8, seed=i).cuda() for i in range(8)]
stream = torch.cuda.Stream()
def run(fn, n=200, tag=''):
    with torch.cuda.stream(stream):
        for i in range(10): fn(i)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); s.record()
        for i in range(n): fn(i)
        e.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
    print(f"{tag}: host enqueue {1e3*(t1-t0)/n:.3f} ms/step, gpu {s.elapsed_time(e)/n:.3f} ms/step", flush=True)
run(lambda i: shard(xs[i % 8]), tag='shard forward (embed + graph + head)')
emb = shard.vit._embed(xs[0])
st = shard.vit.stage
outs = [(torch.empty(8, 197, 768, device='cuda'), None) for _ in range(4)]
run(lambda i: st.forward(emb, out=outs[i % 4], use_graph=True), tag='stage graph only')
run(lambda i: st.forward(emb, out=outs[i % 4], use_graph=False), tag='stage eager')
run(lambda i: shard.vit._embed(xs[i % 8]), tag='embed only')
