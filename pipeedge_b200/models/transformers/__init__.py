"""Transformer shards (`src/pipeedge/models/transformers/__init__.py`)."""
from typing import Tuple, Type, Union
from torch import Tensor

TransformerShardData: Type = Union[Tensor, Tuple[Tensor, Tensor]]
"""A transformer shard input/output type: a tensor, or the (data, skip) tuple of a mid-block cut."""
