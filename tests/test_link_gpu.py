"""GPU: peer-memory links (csrc/link.cu) on a loop-back link - both ends in this process, same kernels and flags as a
hop between two GPUs. The fused quantise-and-send kernel must write exactly the reference's wire bytes (goldens produced
by the reference's own `tensor_encode_outerdim` + clamp, `oracle/make_goldens.py`), the receive kernel exactly its decoded
values; raw payloads must arrive unchanged; the ring's flags must order producers and consumers on different streams."""
import ctypes
import os
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
QG = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'quant.npz'))
HDR = 16384


@pytest.fixture(scope='module')
def lib():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pipeedge_b200 import _lib
    return _lib


class Loop:
    """A loop-back link plus the helpers the tests need."""

    def __init__(self, lib, payload_bytes, slots=4):
        self.lib = lib
        self.h = ctypes.c_void_p()
        lib.check(lib.LIB.pe_link_open_local(payload_bytes, slots, 8, ctypes.byref(self.h)))

    def close(self):
        self.lib.LIB.pe_link_close(self.h)

    def put(self, a, b=None, bit=0, clamp=0, a1=None, b1=None, stream=None):
        items = a.shape[0]
        n0 = a.numel() // items
        n1 = 0 if a1 is None else a1.numel() // items
        s = (stream or torch.cuda.current_stream()).cuda_stream
        self.lib.check(self.lib.LIB.pe_link_put(self.h, a.data_ptr(), None if b is None else b.data_ptr(), n0,
                                                None if a1 is None else a1.data_ptr(),
                                                None if b1 is None else b1.data_ptr(), n1, items, bit, clamp, s))

    def get(self, shape, shape1=None, stream=None):
        items = shape[0]
        d0 = torch.empty(shape, dtype=torch.float32, device='cuda')
        d1 = None if shape1 is None else torch.empty(shape1, dtype=torch.float32, device='cuda')
        s = (stream or torch.cuda.current_stream()).cuda_stream
        self.lib.check(self.lib.LIB.pe_link_get(self.h, d0.data_ptr(), None if d1 is None else d1.data_ptr(), items,
                                                d0.numel() // items, 0 if d1 is None else d1.numel() // items, s))
        return d0 if d1 is None else (d0, d1)

    def read(self, seq, offset, nbytes):
        buf = (ctypes.c_uint8 * nbytes)()
        self.lib.check(self.lib.LIB.pe_link_debug_read(self.h, seq, offset, buf, nbytes))
        return np.frombuffer(buf, dtype=np.uint8).copy()

    def check(self):
        self.lib.check(self.lib.LIB.pe_link_check(self.h))


def test_raw_payload_round_trip_and_deferred_add(lib):
    link = Loop(lib, 8 << 20)
    try:
        g = torch.Generator(device='cuda').manual_seed(3)
        a = torch.randn(8, 197, 768, device='cuda', generator=g)
        b = torch.randn(8, 197, 768, device='cuda', generator=g)
        link.put(a)
        got = link.get(a.shape)
        link.put(a, b)                    # a stage that deferred its final residual add
        got_sum = link.get(a.shape)
        torch.cuda.synchronize()
        link.check()
        assert torch.equal(got, a)
        assert torch.equal(got_sum, a + b)
    finally:
        link.close()


def test_two_tensor_payload_ragged_sizes(lib):
    """A mid-block cut ships (data, skip) with different widths; odd element counts take the scalar tails."""
    link = Loop(lib, 1 << 20)
    try:
        g = torch.Generator(device='cuda').manual_seed(4)
        a = torch.randn(3, 7, 33, device='cuda', generator=g)      # 231 per item: not a multiple of 4
        s = torch.randn(3, 7, 9, device='cuda', generator=g)
        link.put(a, a1=s)
        ga, gs = link.get(a.shape, s.shape)
        torch.cuda.synchronize()
        link.check()
        assert torch.equal(ga, a) and torch.equal(gs, s)
    finally:
        link.close()


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
@pytest.mark.parametrize('bit', [2, 3, 4, 5, 6, 8, 10, 16])
def test_fused_quant_send_matches_reference_goldens(lib, tag, bit):
    """pe_quant_encode_send (clamp + per-item quantise + pack + store into the consumer's slot) vs the bytes the
    reference produced for the same fp32 input; pe_link_get vs the reference's decoded tensor. Tags a / c take the
    one-kernel path for bit in {2,4,8,16} (n % 16 == 0), everything else the staged path."""
    x = torch.from_numpy(QG[f"x_{tag}"]).cuda()
    items = x.shape[0]
    n = x.numel() // items
    link = Loop(lib, 1 << 20)
    try:
        lib.check(lib.LIB.pe_quant_encode_send(link.h, x.data_ptr(), None, items, n, bit, lib.PE_CLAMP_AUTO,
                                               torch.cuda.current_stream().cuda_stream))
        want_codes = QG[f"comm_{tag}_{bit}"]
        codes = link.read(0, HDR, want_codes.size).reshape(want_codes.shape)
        scale = link.read(0, 256, 4 * items).view(np.float32)
        shift = link.read(0, 256 + 2048, 4 * items).view(np.float32)
        np.testing.assert_array_equal(codes, want_codes)
        np.testing.assert_array_equal(scale, QG[f"scale_{tag}_{bit}"])
        np.testing.assert_array_equal(shift, QG[f"shift_{tag}_{bit}"])
        alpha = link.read(0, 16 + 24, 4).view(np.float32)[0]        # LinkHeader.t[0].alpha
        from oracle import quant as oq
        want_alpha, kind = oq.clamp_alpha(x.cpu(), bit)             # the oracle's threshold is pinned by the goldens
        assert kind == 'laplace' and alpha == np.float32(want_alpha)
        dec = link.get(x.shape)
        torch.cuda.synchronize()
        link.check()
        np.testing.assert_array_equal(dec.cpu().numpy(), QG[f"dec_{tag}_{bit}"])
    finally:
        link.close()


@pytest.mark.parametrize('shape,bit', [((32, 198, 768), 8), ((8, 197, 768), 4), ((200, 4, 64), 8), ((5, 1, 16), 2)])
def test_fused_quant_send_equals_standalone_kernels(lib, shape, bit):
    """BASELINE config 5's hop ([32,198,768], 8 bit: the shared-memory-resident slice), a many-items case (several
    segments per CTA) and tiny items: the fused kernel's bytes == the three stand-alone kernels' (themselves pinned by
    the goldens and the NumPy oracle), also when the payload is the deferred sum a + b."""
    from pipeedge_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(11)
    a = torch.randn(shape, device='cuda', generator=g) * 1.7
    b = torch.randn(shape, device='cuda', generator=g) * 0.3
    items = shape[0]
    n = a.numel() // items
    link = Loop(lib, items * n * 4 + 4096)
    try:
        for seq, (p, q) in enumerate(((a, None), (a, b))):
            x = p if q is None else p + q
            codes_w, scale_w, shift_w, _ = ops.quant_encode(x, bit, True)
            link.put(p, q, bit=bit, clamp=lib.PE_CLAMP_AUTO)
            got = link.read(seq, HDR, codes_w.numel()).reshape(codes_w.shape)
            np.testing.assert_array_equal(got, codes_w.cpu().numpy())
            np.testing.assert_array_equal(link.read(seq, 256, 4 * items).view(np.float32), scale_w.cpu().numpy())
            np.testing.assert_array_equal(link.read(seq, 256 + 2048, 4 * items).view(np.float32), shift_w.cpu().numpy())
            dec = link.get(shape)
            want = ops.quant_decode(codes_w, shape[1:], bit, scale_w, shift_w)
            torch.cuda.synchronize()
            assert torch.equal(dec, want)
        link.check()
    finally:
        link.close()


def test_ring_flags_order_producer_and_consumer_streams(lib):
    """12 payloads through a 3-slot ring with the producer and the consumer on different streams and the producer
    enqueued first: puts 4.. must wait (on the device) for the gets that free their slots; order and contents survive."""
    link = Loop(lib, 1 << 20, slots=3)
    try:
        prod, cons = torch.cuda.Stream(), torch.cuda.Stream()
        g = torch.Generator(device='cuda').manual_seed(5)
        xs = [torch.randn(4, 64, 128, device='cuda', generator=g) for _ in range(12)]
        torch.cuda.synchronize()
        outs = []
        for i, x in enumerate(xs):
            link.put(x, bit=8 if i % 2 else 0, clamp=lib.PE_CLAMP_AUTO if i % 2 else 0, stream=prod)
        for _ in xs:
            with torch.cuda.stream(cons):
                outs.append(link.get(xs[0].shape, stream=cons))
        torch.cuda.synchronize()
        link.check()
        from pipeedge_b200 import ops
        for i, (x, y) in enumerate(zip(xs, outs)):
            if i % 2 == 0:
                assert torch.equal(x, y), i
            else:
                c, sc, sh, _ = ops.quant_encode(x, 8, True)
                assert torch.equal(y, ops.quant_decode(c, x.shape[1:], 8, sc, sh)), i
    finally:
        link.close()


def test_host_fed_link_and_f16_wire(lib, monkeypatch):
    """The data rank's input ring (host copy + flag on a side stream, raw get) and the optional fp16 wire format."""
    h = ctypes.c_void_p()
    lib.check(lib.LIB.pe_link_open_host(1 << 20, 2, ctypes.byref(h)))
    try:
        copy = torch.cuda.Stream()
        srcs = [torch.arange(1000 + i, dtype=torch.int64).pin_memory() + i for i in range(5)]
        outs = []
        for s in srcs:     # 5 payloads through 2 slots: feed blocks on the host until the get released the slot
            nbytes = s.numel() * 8
            d = torch.empty(s.numel(), dtype=torch.int64, device='cuda')
            lib.check(lib.LIB.pe_link_feed(h, s.data_ptr(), nbytes, 1, copy.cuda_stream))
            lib.check(lib.LIB.pe_link_get_raw(h, d.data_ptr(), nbytes, torch.cuda.current_stream().cuda_stream))
            outs.append(d)
        torch.cuda.synchronize()
        for s, d in zip(srcs, outs):
            assert torch.equal(d.cpu(), s)
    finally:
        lib.LIB.pe_link_close(h)
    monkeypatch.setenv('PIPEEDGE_WIRE_F16', '1')
    link = Loop(lib, 1 << 20)
    try:
        x = torch.randn(4, 50, 64, device='cuda')
        link.put(x)
        y = link.get(x.shape)
        torch.cuda.synchronize()
        assert torch.equal(y, x.half().float())
    finally:
        link.close()
