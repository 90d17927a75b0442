"""GPU: QuantPipe kernels vs the reference-generated goldens and the NumPy oracle. Bit-exact target:
uint8 code bytes, fp32 scale/shift, decoded fp32 values and the Laplace clamp threshold."""
import os
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
QG = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'quant.npz'))


@pytest.fixture(scope='module')
def q():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pipeedge_b200.quantization import basic_op, clamp_op
    return basic_op, clamp_op


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
@pytest.mark.parametrize('bit', [2, 3, 4, 5, 6, 8, 10, 16])
def test_encode_decode_vs_reference_goldens(q, tag, bit):
    """Same seeded fp32 input as the reference run: the hook's clamp + encode, then decode."""
    basic_op, _ = q
    x = torch.from_numpy(QG[f"x_{tag}"]).cuda()
    enc = basic_op.tensor_encode_outerdim(x, bit, clamp=True)
    for nm, t in zip(('comm', 'shape', 'scale', 'shift', 'bit'), enc):
        want = QG[f"{nm}_{tag}_{bit}"]
        got = t.cpu().numpy()
        assert got.dtype == want.dtype, nm
        np.testing.assert_array_equal(got, want, err_msg=nm)
    dec = basic_op.tensor_decode_outerdim(enc)
    np.testing.assert_array_equal(dec.cpu().numpy(), QG[f"dec_{tag}_{bit}"])


@pytest.mark.parametrize('bit', [4, 8])
def test_gelu_clamp_branch(q, bit):
    """All values >= 0.2 selects the GeLU clamp; its fp32 `sum(x^2)` follows torch's CPU summation order, so the
    threshold is held to 2 ulp and the codes to exact equality only when the threshold matches bit for bit."""
    basic_op, clamp_op = q
    from pipeedge_b200 import ops
    from oracle import quant as oq
    x = torch.from_numpy(QG['x_g']).cuda()
    _, _, _, alpha = ops.quant_encode(x, bit, True)
    want_alpha, kind = oq.clamp_alpha(x.cpu(), bit)
    assert kind == 'gelu'
    got_alpha = np.float32(alpha.item())
    assert abs(float(got_alpha) - float(want_alpha)) <= 2 * np.spacing(want_alpha)
    enc = basic_op.tensor_encode_outerdim(x, bit, clamp=True)
    if got_alpha == want_alpha or float(x.abs().max()) < min(float(got_alpha), float(want_alpha)):
        # same threshold, or a threshold that clamps nothing: the reference's codes must be reproduced exactly
        np.testing.assert_array_equal(enc[0].cpu().numpy(), QG[f"comm_g_{bit}"])
        np.testing.assert_array_equal(enc[2].cpu().numpy(), QG[f"scale_g_{bit}"])
    clamped = clamp_op.clamp_banner2019_gelu(x, bit)
    assert float(clamped.max()) <= float(got_alpha)


def test_bit0_passthrough(q):
    basic_op, _ = q
    x = torch.from_numpy(QG['x_g']).cuda()
    enc = basic_op.tensor_encode_outerdim(x, 0)
    assert enc[0] is x
    for nm, t in zip(('shape', 'scale', 'shift', 'bit'), enc[1:]):
        want = QG[f"{nm}_g_0"]
        assert t.numpy().dtype == want.dtype, nm
        np.testing.assert_array_equal(t.numpy(), want, err_msg=nm)
    assert basic_op.tensor_decode_outerdim(enc) is x


@pytest.mark.parametrize('shape,bit', [((32, 198, 768), 8), ((8, 197, 768), 8), ((4, 197, 3072), 4),
                                       ((2, 128, 768), 16), ((3, 50, 77), 6), ((1, 7), 2), ((5, 1000), 5)])
def test_full_size_against_oracle(q, shape, bit):
    """BASELINE config C5's hop tensor and friends: CUDA vs the NumPy oracle on the same seeded input."""
    from oracle import quant as oq
    basic_op, _ = q
    gen = torch.Generator().manual_seed(sum(shape) + bit)
    x = torch.randn(*shape, generator=gen) * 1.3 + 0.1
    x[0].view(-1)[:3] = torch.tensor([40.0, -35.0, 12.0])    # outliers so that the clamp bites
    want = oq.hook_encode(x, bit)
    enc = basic_op.tensor_encode_outerdim(x.cuda(), bit, clamp=True)
    for nm, g, w in zip(('comm', 'shape', 'scale', 'shift', 'bit'), enc, want):
        np.testing.assert_array_equal(g.cpu().numpy(), w.numpy(), err_msg=nm)
    dec = basic_op.tensor_decode_outerdim(enc).cpu()
    np.testing.assert_array_equal(dec.numpy(), oq.hook_decode(want).numpy())
    # size-independent property: decode(encode(x)) is within half a quantisation step of clamp(x)
    alpha, _ = oq.clamp_alpha(x, bit)
    xc = x.clamp(-float(alpha), float(alpha))
    step = enc[2].cpu().view(-1, *([1] * (x.dim() - 1))) / ((1 << bit) - 1)
    slack = 8 * float(np.finfo(np.float32).eps) * float(x.abs().max())   # fp32 rounding of scale/shift arithmetic
    assert torch.all((dec - xc).abs() <= 0.5 * step * 1.0001 + slack)


def test_no_clamp_is_bare_tensor_encode_outerdim(q):
    from oracle import quant as oq
    basic_op, _ = q
    x = torch.randn(3, 33, 64, generator=torch.Generator().manual_seed(3))
    want = oq.tensor_encode_outerdim(x, 8)
    enc = basic_op.tensor_encode_outerdim(x.cuda(), 8, clamp=False)
    for g, w in zip(enc, want):
        np.testing.assert_array_equal(g.cpu().numpy(), w.numpy())


def test_laplace_clamp_function(q):
    from oracle import quant as oq
    _, clamp_op = q
    x = torch.randn(4, 100, 64, generator=torch.Generator().manual_seed(9)) * 3
    alpha, kind = oq.clamp_alpha(x, 4)
    assert kind == 'laplace'
    got = clamp_op.clamp_banner2019_laplace(x.cuda(), 4).cpu()
    np.testing.assert_array_equal(got.numpy(), x.clamp(-float(alpha), float(alpha)).numpy())


def test_rejects_bad_bit(q):
    basic_op, _ = q
    from pipeedge_b200._lib import PipeEdgeB200Error
    with pytest.raises((PipeEdgeB200Error, ValueError)):
        basic_op.tensor_encode_outerdim(torch.randn(2, 8).cuda(), 17)
