"""Common machinery of the ViT / DeiT / BERT shard classes: weight-file handling, payload staging on the
device, output-buffer rings for CUDA-graph replay, and the last-stage heads."""
from collections.abc import Mapping
from typing import Optional, Tuple, Union
import numpy as np
import torch
from ... import _lib, ops
from ..._lib import LIB, check
from .. import ModuleShard, ModuleShardConfig
from . import TransformerShardData
from ._stage import EncoderStage, _dev


class GpuTransformerShard(ModuleShard):
    """A shard whose encoder blocks (and edges) run through `libpipeedge_b200.so` on the current device.

    Constructor contract = the reference's (`vit.py:192-204`): `(config, shard_config, model_weights)` with
    `model_weights` an `.npz` path or an already-loaded mapping. `forward` takes/returns a tensor or the
    `(data, skip)` tuple of a mid-block cut; outputs are fp32 CUDA tensors.

    Attributes that tune the device path (not part of the reference API):
      use_cuda_graph  replay the stage's kernel sequence as one CUDA graph (needs stable buffers)
      num_slots       ring of persistent output buffers used when `use_cuda_graph` is on; an output is
                      overwritten `num_slots` forwards later, which matches the reference's pipeline depth of
                      queued + in-flight payloads (`p2p/__init__.py:374-376`)
    """
    FAMILY = ''

    def __init__(self, config, shard_config: ModuleShardConfig, model_weights: Union[str, Mapping]):
        super().__init__(config, shard_config)
        self.use_cuda_graph = False
        self.num_slots = 4
        self._slot = 0
        self._rings = {}
        self._stream_warm = False
        if isinstance(model_weights, str):
            with np.load(model_weights) as weights:
                self._build_shard(weights)
        else:
            self._build_shard(model_weights)

    # ------------------------------------------------------------------ construction
    def _build_shard(self, weights: Mapping) -> None:
        raise NotImplementedError

    def _make_stage(self, weights: Mapping, tokens: int, max_ubatch: int = 64) -> EncoderStage:
        return EncoderStage(self.FAMILY, self.config, self.shard_config.layer_start, self.shard_config.layer_end,
                            weights, tokens, max_ubatch)

    @staticmethod
    def _f16(arr, device) -> torch.Tensor:
        return _dev(arr, torch.float16, device)

    @staticmethod
    def _f32(arr, device) -> torch.Tensor:
        return _dev(arr, torch.float32, device)

    # ------------------------------------------------------------------ payload staging
    def _to_device(self, data: TransformerShardData, dtype=torch.float32) -> TransformerShardData:
        dev = self.stage.device
        if isinstance(data, torch.Tensor):
            return data.to(device=dev, dtype=dtype, non_blocking=True).contiguous()
        return tuple(t.to(device=dev, dtype=dtype, non_blocking=True).contiguous() for t in data)

    def _ring(self, name: str, shape, dtype=torch.float32) -> torch.Tensor:
        """Persistent buffer `name` of the current slot (allocated on first use per shape)."""
        key = (name, tuple(shape), dtype, self._slot)
        buf = self._rings.get(key)
        if buf is None:
            buf = torch.empty(tuple(shape), dtype=dtype, device=self.stage.device)
            self._rings[key] = buf
        return buf

    def _stage_out(self, ubatch: int) -> Optional[Tuple[torch.Tensor, Optional[torch.Tensor]]]:
        if not self.use_cuda_graph:
            return None
        s0, s1 = self.stage.out_shapes(ubatch)
        return self._ring('out0', s0), (None if s1 is None else self._ring('out1', s1))

    def _run_blocks(self, data: TransformerShardData) -> TransformerShardData:
        in0 = data[0] if isinstance(data, tuple) else data
        out = self._stage_out(in0.shape[0])
        res = self.stage.forward(data, out=out, use_graph=self.use_cuda_graph)
        if self.use_cuda_graph:
            self._slot = (self._slot + 1) % max(1, self.num_slots)
        return res

    # ------------------------------------------------------------------ heads
    def _cls_rows(self, data: torch.Tensor) -> torch.Tensor:
        return data[:, 0, :].contiguous()

    def _classify(self, a16: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
        return ops.linear(a16, weight, bias, _lib.PE_EPI_F32)
