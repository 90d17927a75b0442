"""Tensor-level wrappers over the C-ABI: device memory and streams come from PyTorch, the math does not.

Each function takes contiguous CUDA tensors, enqueues the library's kernels on the current CUDA stream
and returns without synchronising. dtype/shape mistakes raise before anything is launched.
"""
from typing import Optional, Tuple
import torch
from . import _lib
from ._lib import LIB, check


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _want(t: torch.Tensor, dtype: torch.dtype, name: str) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (pipeedge_b200 has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous tensor")


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, want_f32: bool = True,
              want_f16: bool = False, out_f32: Optional[torch.Tensor] = None,
              out_f16: Optional[torch.Tensor] = None) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """LayerNorm over the last dim of an fp32 tensor; returns (f32 or None, f16 or None). `out_f32` / `out_f16`
    supply persistent output buffers (CUDA-graph capture)."""
    _want(x, torch.float32, 'x')
    _want(gamma, torch.float32, 'gamma')
    _want(beta, torch.float32, 'beta')
    hidden = x.shape[-1]
    rows = x.numel() // hidden
    o32 = (out_f32 if out_f32 is not None else torch.empty_like(x)) if want_f32 else None
    o16 = (out_f16 if out_f16 is not None else torch.empty(x.shape, dtype=torch.float16, device=x.device)) \
        if want_f16 else None
    if o32 is not None:
        _want(o32, torch.float32, 'out_f32')
    if o16 is not None:
        _want(o16, torch.float16, 'out_f16')
    check(LIB.pe_layernorm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), float(eps), _ptr(o32), _ptr(o16), rows,
                           hidden, _stream()))
    return o32, o16


def residual_layernorm(y: torch.Tensor, resid: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
                       want_sum: bool = True, want_f32: bool = False, want_f16: bool = True):
    """t = y + resid; returns (t or None, LayerNorm(t) f32 or None, LayerNorm(t) f16 or None)."""
    for name, t in (('y', y), ('resid', resid), ('gamma', gamma), ('beta', beta)):
        _want(t, torch.float32, name)
    hidden = y.shape[-1]
    rows = y.numel() // hidden
    tsum = torch.empty_like(y) if want_sum else None
    o32 = torch.empty_like(y) if want_f32 else None
    o16 = torch.empty(y.shape, dtype=torch.float16, device=y.device) if want_f16 else None
    check(LIB.pe_residual_layernorm(y.data_ptr(), resid.data_ptr(), gamma.data_ptr(), beta.data_ptr(), float(eps),
                                    _ptr(tsum), _ptr(o32), _ptr(o16), rows, hidden, _stream()))
    return tsum, o32, o16


_EPI_OUT = {_lib.PE_EPI_F16: torch.float16, _lib.PE_EPI_GELU_F16: torch.float16, _lib.PE_EPI_RESID_F32: torch.float32,
            _lib.PE_EPI_F32: torch.float32, _lib.PE_EPI_TANH_F32: torch.float32}


def linear(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], epilogue: int,
           resid: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
           debug_simt: bool = False, static_w: bool = False) -> torch.Tensor:
    """`epilogue(a @ w.T + bias)`; a f16 [..., k], w f16 [n, k] (nn.Linear layout). `static_w`: w is a model weight,
    not written by work still pending on the stream (lets the kernel fetch it ahead of the preceding kernel's end)."""
    _want(a, torch.float16, 'a')
    _want(w, torch.float16, 'w')
    k = a.shape[-1]
    m = a.numel() // k
    n = w.shape[0]
    if w.shape[1] != k:
        raise ValueError(f"linear: a[..., {k}] vs w[{n}, {w.shape[1]}]")
    if bias is not None:
        _want(bias, torch.float32, 'bias')
    if resid is not None:
        _want(resid, torch.float32, 'resid')
    if out is None:
        out = torch.empty(a.shape[:-1] + (n,), dtype=_EPI_OUT[epilogue], device=a.device)
    else:
        _want(out, _EPI_OUT[epilogue], 'out')
    fn = LIB.pe_debug_linear_simt if debug_simt else LIB.pe_linear
    flags = _lib.PE_EPI_STATIC_W if (static_w and not debug_simt) else 0
    check(fn(a.data_ptr(), w.data_ptr(), _ptr(bias), _ptr(resid), out.data_ptr(), m, n, k, epilogue | flags, _stream()))
    return out


def linear_residual_layernorm(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, resid: torch.Tensor,
                              gamma: torch.Tensor, beta: torch.Tensor, eps: float, f32_is_ln: bool = False,
                              want_f32: bool = True, want_f16: bool = True, out_f32: Optional[torch.Tensor] = None):
    """v = a @ w.T + bias + resid, LayerNorm(v) - one kernel (cluster epilogue). Returns (f32 tensor or None, f16 or
    None): the f32 output is v, or LayerNorm(v) when `f32_is_ln`; `out_f32` may be `resid` itself (in place)."""
    _want(a, torch.float16, 'a')
    _want(w, torch.float16, 'w')
    for name, t in (('bias', bias), ('resid', resid), ('gamma', gamma), ('beta', beta)):
        _want(t, torch.float32, name)
    k = a.shape[-1]
    m = a.numel() // k
    n = w.shape[0]
    if LIB.pe_linear_ln_cluster(n) <= 0:
        raise ValueError(f"linear_residual_layernorm: width {n} is not supported by the fused kernel")
    o32 = (out_f32 if out_f32 is not None else torch.empty((m, n), dtype=torch.float32, device=a.device)) if want_f32 else None
    o16 = torch.empty((m, n), dtype=torch.float16, device=a.device) if want_f16 else None
    check(LIB.pe_linear_residual_layernorm(a.data_ptr(), w.data_ptr(), bias.data_ptr(), resid.data_ptr(), gamma.data_ptr(),
                                           beta.data_ptr(), float(eps), _ptr(o32), 1 if f32_is_ln else 0, _ptr(o16), m, n, k,
                                           1, _stream()))
    return o32, o16


def attention(qkv: torch.Tensor, batch: int, tokens: int, heads: int, head_dim: int = 0) -> torch.Tensor:
    """Unmasked MHA over fused qkv f16 [batch*tokens, 3*H]; returns merged-head ctx f16 [batch*tokens, H].
    `head_dim` defaults to H / heads and is only checked against it."""
    _want(qkv, torch.float16, 'qkv')
    hidden = qkv.shape[-1] // 3
    if head_dim and head_dim * heads != hidden:
        raise ValueError(f"attention: heads={heads} x head_dim={head_dim} != hidden={hidden}")
    ctx = torch.empty((batch * tokens, hidden), dtype=torch.float16, device=qkv.device)
    check(LIB.pe_attention(qkv.data_ptr(), ctx.data_ptr(), batch, tokens, heads, hidden // heads, _stream()))
    return ctx


def quant_encode(x: torch.Tensor, bit: int, clamp: bool):
    """QuantPipe encode of an fp32 `[items, ...]` tensor. Returns (codes u8 [items, 4*words], scale f32
    [items], shift f32 [items], alpha f32 [1]); nothing is synchronised."""
    _want(x, torch.float32, 'x')
    items = x.shape[0]
    n = x.numel() // items
    words = LIB.pe_quant_words(n, bit)
    if words == 0:
        raise ValueError(f"quant_encode: unsupported bit width {bit}")
    dev = x.device
    codes = torch.empty((items, 4 * words), dtype=torch.uint8, device=dev)
    scale = torch.empty(items, dtype=torch.float32, device=dev)
    shift = torch.empty(items, dtype=torch.float32, device=dev)
    alpha = torch.empty(1, dtype=torch.float32, device=dev)
    work = torch.empty(LIB.pe_quant_workspace_bytes(items, n), dtype=torch.uint8, device=dev)
    check(LIB.pe_quant_encode(x.data_ptr(), items, n, bit, _lib.PE_CLAMP_AUTO if clamp else _lib.PE_CLAMP_NONE,
                              codes.data_ptr(), scale.data_ptr(), shift.data_ptr(), alpha.data_ptr(), work.data_ptr(),
                              _stream()))
    return codes, scale, shift, alpha


def quant_decode(codes: torch.Tensor, item_shape, bit: int, scale: torch.Tensor, shift: torch.Tensor) -> torch.Tensor:
    """Inverse of `quant_encode`: fp32 `[items, *item_shape]`."""
    _want(codes, torch.uint8, 'codes')
    _want(scale, torch.float32, 'scale')
    _want(shift, torch.float32, 'shift')
    items = codes.shape[0]
    n = 1
    for d in item_shape:
        n *= int(d)
    if codes.shape[1] != 4 * LIB.pe_quant_words(n, bit):
        raise ValueError(f"quant_decode: {codes.shape[1]} bytes per item do not hold {n} codes of {bit} bits")
    out = torch.empty((items, *[int(d) for d in item_shape]), dtype=torch.float32, device=codes.device)
    check(LIB.pe_quant_decode(codes.data_ptr(), items, n, bit, scale.data_ptr(), shift.data_ptr(), out.data_ptr(),
                              _stream()))
    return out


def launch_count() -> int:
    """Kernels launched by the library since it was loaded."""
    return int(LIB.pe_launch_count())
