"""Device QuantPipe encode/decode - drop-in for `pipeedge.quantization.basic_op` (reference `basic_op.py`).

`tensor_encode_outerdim` / `tensor_decode_outerdim` keep the reference's 5-tensor wire format per encoded
tensor, `[comm u8, shape, scale f32, shift f32, quant_bit i8]` (`basic_op.py:114-176`); the integer codes, scales
and shifts are bit-identical to the NumPy implementation for the same fp32 input. `comm`, `scale` and
`shift` live on the GPU; `shape` and `quant_bit` are host-known and stay on the CPU so that reading them
(`runtime.py:103`) never synchronises the device.
"""
from typing import List
import torch
from .. import ops


def compression_factor(quant_bit: torch.Tensor) -> torch.Tensor:
    """Data-size improvement for quantization bit widths > 0 (`basic_op.py:109-111`)."""
    return torch.div(32, quant_bit)


_PASSTHROUGH_META = {}


def _passthrough_meta(items: int, item_shape: tuple):
    """The four host-side tensors of an un-quantised payload; constant per shape, so built once (read-only)."""
    key = (items, item_shape)
    meta = _PASSTHROUGH_META.get(key)
    if meta is None:
        meta = (torch.tensor(item_shape).expand(items, -1).clone(), torch.ones(items), torch.zeros(items),
                torch.zeros(items, dtype=torch.int8))
        _PASSTHROUGH_META[key] = meta
    return meta


_QUANT_META = {}


def _quant_meta(items: int, item_shape: tuple, bit: int):
    """Host-side `shape` (int32) and `quant_bit` (int8) tensors of an encoded payload: constant per (shape, bit),
    so the same read-only objects are returned every time (the hop then skips re-sending them)."""
    key = (items, item_shape, bit)
    meta = _QUANT_META.get(key)
    if meta is None:
        meta = (torch.tensor(item_shape, dtype=torch.int32).expand(items, -1).clone(),
                torch.full((items,), bit, dtype=torch.int8))
        _QUANT_META[key] = meta
    return meta


def tensor_encode_outerdim(batched_tensor: torch.Tensor, quant_bit: int, clamp: bool = False) -> List[torch.Tensor]:
    """Per-item quantisation of a micro-batched fp32 CUDA tensor (`basic_op.py:166-170`).

    `clamp=True` fuses the Banner-2019 clamp of `forward_hook_quant_encode` (`runtime.py:83-85`) into the same
    pass over the data."""
    items = batched_tensor.shape[0]
    item_shape = tuple(batched_tensor.shape[1:])
    if quant_bit == 0:   # passthrough layout of `tensor_encode` (`basic_op.py:120-122`)
        return [batched_tensor, *_passthrough_meta(items, item_shape)]
    comm, scale, shift, _alpha = ops.quant_encode(batched_tensor, int(quant_bit), clamp)
    shape, bits = _quant_meta(items, item_shape, int(quant_bit))
    return [comm, shape, scale, shift, bits]


def tensor_decode_outerdim(batched_encodings: List[torch.Tensor]) -> torch.Tensor:
    """Inverse of `tensor_encode_outerdim` (`basic_op.py:173-176`)."""
    comm, shape, scale, shift, quant_bit = batched_encodings
    bit = int(quant_bit[0])
    if bit == 0:
        return comm
    return ops.quant_decode(comm, shape[0].tolist(), bit, scale, shift)
