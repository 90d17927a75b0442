"""GPU x2: a real 2-stage pipeline over NCCL through the reference-shaped API (DistP2pContext +
DistP2pPipelineStage + shard + QuantPipe hooks), checked against the CPU oracle. Skipped with < 2 GPUs."""
import os
import socket
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, cuts, qbits, n_ubatch, ubatch, out_q):
    import faulthandler
    import sys
    import threading
    faulthandler.enable()   # a native crash prints every thread's Python stack into the test log
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    import runtime as rt
    from pipeedge_b200.comm.p2p import DistP2pContext, DistP2pPipelineStage
    from pipeedge_b200.models import ModuleShardConfig
    from pipeedge_b200.models.transformers import bert, deit, vit
    from pipeedge_b200.synth import MODEL_SPECS, hf_config, synth_input, synth_weights
    spec = MODEL_SPECS[name]
    classes = {'vit': vit.ViTShardForImageClassification, 'deit': deit.DeiTShardForImageClassification,
               'bert': bert.BertShardForSequenceClassification}
    lo = 1 if rank == 0 else cuts[rank - 1] + 1
    hi = cuts[rank]
    cfg = ModuleShardConfig(layer_start=lo, layer_end=hi, is_first=lo == 1, is_last=hi == spec.layers)
    shard = classes[spec.family](hf_config(spec), cfg, synth_weights(spec, seed=0))
    shard.use_cuda_graph = True
    shard.register_buffer('quant_bit', torch.tensor(qbits[rank]), persistent=False)
    if rank != world - 1:
        shard.register_forward_hook(rt.forward_hook_quant_encode)
    if rank != 0:
        shard.register_forward_pre_hook(rt.forward_pre_hook_quant_decode)
    stop = threading.Event()
    results = []
    done = threading.Event()

    def results_cb(t):
        results.append(t.cpu())
        if len(results) == n_ubatch:
            done.set()

    with DistP2pContext(('gloo',), {'world_size': world, 'rank': rank}, lambda c, t: stop.set() if c == 0 else None) as ctx:
        src = world - 1 if rank == 0 else rank - 1
        dst = 0 if rank == world - 1 else rank + 1
        with DistP2pPipelineStage(src, dst, shard, results_cb if rank == 0 else None) as stage:
            if rank == 0:
                for i in range(n_ubatch):
                    stage.enqueue_tensor(synth_input(spec, ubatch, seed=10 + i, seq_len=32))
                assert done.wait(300), "results did not arrive"
                stage.check_workers()
                ctx.cmd_broadcast(0)
                out_q.put([r.numpy() for r in results])
            else:
                assert stop.wait(420)
                stage.check_workers()
    # leave without exit-time teardown of the CUDA / NCCL libraries (see bench.py: it can crash after a clean shutdown)
    out_q.close()
    out_q.join_thread()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def _local_reference(name, cuts, qbits, n_ubatch, ubatch):
    """The same two shards + QuantPipe hooks run back to back in THIS process on cuda:0 (no communication)."""
    import sys
    sys.path.insert(0, ROOT)
    import runtime as rt
    from pipeedge_b200.models import ModuleShardConfig
    from pipeedge_b200.models.transformers import bert, deit, vit
    from pipeedge_b200.synth import MODEL_SPECS, hf_config, synth_input, synth_weights
    spec = MODEL_SPECS[name]
    classes = {'vit': vit.ViTShardForImageClassification, 'deit': deit.DeiTShardForImageClassification,
               'bert': bert.BertShardForSequenceClassification}
    weights = synth_weights(spec, seed=0)
    shards = []
    for rank in range(2):
        lo = 1 if rank == 0 else cuts[rank - 1] + 1
        cfg = ModuleShardConfig(layer_start=lo, layer_end=cuts[rank], is_first=lo == 1, is_last=cuts[rank] == spec.layers)
        shard = classes[spec.family](hf_config(spec), cfg, weights)
        shard.register_buffer('quant_bit', torch.tensor(qbits[rank]), persistent=False)
        shards.append(shard)
    shards[0].register_forward_hook(rt.forward_hook_quant_encode)
    shards[1].register_forward_pre_hook(rt.forward_pre_hook_quant_decode)
    return [shards[1](shards[0](synth_input(spec, ubatch, seed=10 + i, seq_len=32))).cpu().numpy()
            for i in range(n_ubatch)]


@pytest.mark.parametrize('name,cuts,qbits', [('test/vit-tiny', (6, 12), (0, 0)), ('test/vit-tiny', (5, 12), (8, 0)),
                                             ('test/bert-tiny', (7, 12), (4, 0))])
def test_two_stage_pipeline_over_nccl(name, cuts, qbits):
    """Results arrive in FIFO order and are BIT-IDENTICAL to running the same shards and hooks locally (the hop
    moves bytes, it must not change them); the unquantised case is also checked against the CPU oracle."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np
    from oracle import shards as osh
    from pipeedge_b200.synth import MODEL_SPECS, synth_input, synth_weights
    n_ubatch, ubatch = 7, 3
    ctx = mp.get_context('spawn')
    out_q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, cuts, qbits, n_ubatch, ubatch, out_q)) for r in range(2)]
    for p in procs:
        p.start()
    got = out_q.get(timeout=600)   # a fresh box pages torch in for a minute per process
    for r, p in enumerate(procs):
        p.join(120)
        assert p.exitcode == 0, f"rank {r} exited with {p.exitcode}"
    assert len(got) == n_ubatch
    local = _local_reference(name, cuts, qbits, n_ubatch, ubatch)
    for i, (logits, want) in enumerate(zip(got, local)):
        np.testing.assert_array_equal(logits, want, err_msg=f"micro-batch {i}")
    if qbits[0] == 0:
        spec = MODEL_SPECS[name]
        w = synth_weights(spec, seed=0)
        for i, logits in enumerate(got):
            x = synth_input(spec, ubatch, seed=10 + i, seq_len=32)
            want = osh.shard_forward(spec, w, 1, spec.layers, x).numpy()
            assert np.abs(logits - want).max() <= 4e-3 * np.abs(want).max(), f"ubatch {i}"


@pytest.mark.parametrize('policy', ['HEURISTIC', 'HEURISTIC2', 'CONTROLLER'])
def test_runtime_cli_adaptive_quant(policy, tmp_path):
    """`runtime.py` on 2 ranks with an adaptive QuantPipe policy and a send-rate constraint no hop can meet: the policy
    reads the hop's device-side transfer times and must move the stage's bit-width off 'no quantization', and the run
    must still deliver every result (`runtime.py:121-216,465-477` of the reference)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess
    import sys
    port = _free_port()
    env = dict(os.environ, ADAPTIVE_QUANT=policy, SEND_CONSTRAINT='1e9', WINDOW_SIZE='2', PYTHONUNBUFFERED='1',
               MONITORING='1' if policy == 'CONTROLLER' else '0')   # one run also with the opt-in device-timed heartbeats
    cmd = [sys.executable, os.path.join(ROOT, 'runtime.py'), None, '2', '--port', str(port), '-m',
           'facebook/deit-tiny-distilled-patch16-224', '-b', '64', '-u', '8', '-pt', '1,24,25,48', '-q', '0,0']
    procs = []
    for rank in (1, 0):
        argv = list(cmd)
        argv[2] = str(rank)
        procs.append(subprocess.Popen(argv, cwd=str(tmp_path), env=dict(env, LOCAL_RANK=str(rank)), stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            p.kill()
            out, _ = p.communicate()
        outs.append(out)
    for p, out in zip(procs, outs):
        assert p.returncode == 0, out[-3000:]
    rank0 = outs[1]
    assert 'throughput is' in rank0, rank0[-3000:]
    import re
    bits = [int(b) for b in re.findall(r'Adaptive quantization \(\w+\): bitwidth1?=(\d+)', rank0)]
    assert bits, rank0[-3000:]
    assert any(0 < b < 32 for b in bits), bits
    if policy == 'CONTROLLER':
        for key in ('shard', 'quant_encode', 'output', 'send'):
            assert f'{key}: Global Time' in rank0, rank0[-3000:]
