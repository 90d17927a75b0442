"""Model specs and seeded synthetic weights in the reference's three npz key layouts.

There is no network, so neither `AutoConfig.from_pretrained` (reference `model_cfg.py:60`) nor the
reference's weight download tools (`save_model_weights.py`) can run. This module supplies:

* `MODEL_SPECS` - explicit architecture numbers for every name in the reference registry
  (`model_cfg.py:24-43`) plus two tiny test models, and `hf_config(spec)` which builds the
  matching HuggingFace config object without touching the network.
* `synth_weights(spec, seed)` - a dict with exactly the keys and array layouts the reference
  loaders read: Google/JAX ViT layout (`vit.py:120-159`), timm DeiT layout (`deit.py:119-156`) and
  the HF BERT state_dict under a `bert.` prefix (`bert.py:104-140,191-201`).

The same dict (or an `.npz` written from it) feeds the reference shards, the CPU oracle and the GPU
shards, which is what makes parity runs "identical inputs".
"""
from dataclasses import dataclass
from typing import Dict
import numpy as np


@dataclass(frozen=True)
class ModelSpec:
    """Architecture numbers for one registry entry."""
    name: str
    family: str            # 'vit' | 'deit' | 'bert'
    hidden: int
    blocks: int            # transformer blocks; the reference's "layers" = 4 * blocks
    heads: int
    inter: int
    num_labels: int
    image_size: int = 224
    patch: int = 16
    channels: int = 3
    vocab: int = 30522
    max_pos: int = 512
    type_vocab: int = 2
    eps: float = 1e-12
    classify: bool = True  # False for the bare `BertModelShard` registry entries

    @property
    def layers(self) -> int:
        """Sub-layer count as the reference's registry counts it (`model_cfg.py:23`)."""
        return 4 * self.blocks

    @property
    def tokens(self) -> int:
        """Sequence length of the image families (BERT's is the input's)."""
        n = (self.image_size // self.patch) ** 2
        return n + (2 if self.family == 'deit' else 1)


MODEL_SPECS: Dict[str, ModelSpec] = {s.name: s for s in [
    ModelSpec('google/vit-base-patch16-224', 'vit', 768, 12, 12, 3072, 1000),
    ModelSpec('google/vit-large-patch16-224', 'vit', 1024, 24, 16, 4096, 1000),
    ModelSpec('google/vit-huge-patch14-224-in21k', 'vit', 1280, 32, 16, 5120, 21843, patch=14),
    ModelSpec('bert-base-uncased', 'bert', 768, 12, 12, 3072, 0, classify=False),
    ModelSpec('bert-large-uncased', 'bert', 1024, 24, 16, 4096, 0, classify=False),
    ModelSpec('textattack/bert-base-uncased-CoLA', 'bert', 768, 12, 12, 3072, 2),
    ModelSpec('facebook/deit-base-distilled-patch16-224', 'deit', 768, 12, 12, 3072, 1000),
    ModelSpec('facebook/deit-small-distilled-patch16-224', 'deit', 384, 12, 6, 1536, 1000),
    ModelSpec('facebook/deit-tiny-distilled-patch16-224', 'deit', 192, 12, 3, 768, 1000),
    # Tiny models (not in the reference registry) so complete tensors fit in committed fixtures.
    ModelSpec('test/vit-tiny', 'vit', 128, 3, 2, 512, 10, image_size=64),
    ModelSpec('test/deit-tiny', 'deit', 128, 2, 2, 512, 10, image_size=64),
    ModelSpec('test/bert-tiny', 'bert', 128, 3, 2, 512, 2, vocab=1000, max_pos=64),
    # ViT-Huge's geometry in small: head_dim 80, patch 14 (K = 588 is not a multiple of 64), odd label count
    ModelSpec('test/vit-huge-tiny', 'vit', 160, 2, 2, 320, 11, image_size=56, patch=14),
]}


def hf_config(spec: ModelSpec):
    """Build the HuggingFace config the reference shard constructors expect (no network)."""
    # pylint: disable=import-outside-toplevel
    from transformers import BertConfig, DeiTConfig, ViTConfig
    if spec.family == 'bert':
        return BertConfig(hidden_size=spec.hidden, num_hidden_layers=spec.blocks,
                          num_attention_heads=spec.heads, intermediate_size=spec.inter,
                          vocab_size=spec.vocab, max_position_embeddings=spec.max_pos,
                          type_vocab_size=spec.type_vocab, layer_norm_eps=spec.eps,
                          num_labels=max(spec.num_labels, 1))
    cls = DeiTConfig if spec.family == 'deit' else ViTConfig
    return cls(hidden_size=spec.hidden, num_hidden_layers=spec.blocks,
               num_attention_heads=spec.heads, intermediate_size=spec.inter,
               image_size=spec.image_size, patch_size=spec.patch, num_channels=spec.channels,
               layer_norm_eps=spec.eps, num_labels=spec.num_labels)


class _Gen:
    """Seeded generator: N(0, std) matrices, LayerNorm scales 1+N(0, std)."""

    def __init__(self, seed: int, std: float):
        self.rng = np.random.default_rng(seed)
        self.std = std

    def normal(self, *shape) -> np.ndarray:
        out = self.rng.standard_normal(shape, dtype=np.float32)
        out *= np.float32(self.std)
        return out

    def ln_scale(self, n) -> np.ndarray:
        return (1.0 + self.normal(n)).astype(np.float32)


def _synth_vit(spec: ModelSpec, g: _Gen) -> Dict[str, np.ndarray]:
    h, nh, inter = spec.hidden, spec.heads, spec.inter
    d = h // nh
    w = {}
    w["cls"] = g.normal(1, 1, h)
    w["Transformer/posembed_input/pos_embedding"] = g.normal(1, spec.tokens, h)
    # JAX conv kernel is [kh, kw, in, out]; the loader transposes [3,2,0,1] (vit.py:125-128)
    w["embedding/kernel"] = g.normal(spec.patch, spec.patch, spec.channels, h)
    w["embedding/bias"] = g.normal(h)
    for i in range(spec.blocks):
        root = f"Transformer/encoderblock_{i}/"
        att = root + "MultiHeadDotProductAttention_1/"
        w[root + "LayerNorm_0/scale"] = g.ln_scale(h)
        w[root + "LayerNorm_0/bias"] = g.normal(h)
        for name in ("query", "key", "value"):
            w[att + name + "/kernel"] = g.normal(h, nh, d)
            w[att + name + "/bias"] = g.normal(nh, d)
        w[att + "out/kernel"] = g.normal(nh, d, h)
        w[att + "out/bias"] = g.normal(h)
        w[root + "LayerNorm_2/scale"] = g.ln_scale(h)
        w[root + "LayerNorm_2/bias"] = g.normal(h)
        w[root + "MlpBlock_3/Dense_0/kernel"] = g.normal(h, inter)
        w[root + "MlpBlock_3/Dense_0/bias"] = g.normal(inter)
        w[root + "MlpBlock_3/Dense_1/kernel"] = g.normal(inter, h)
        w[root + "MlpBlock_3/Dense_1/bias"] = g.normal(h)
    w["Transformer/encoder_norm/scale"] = g.ln_scale(h)
    w["Transformer/encoder_norm/bias"] = g.normal(h)
    w["head/kernel"] = g.normal(h, spec.num_labels)
    w["head/bias"] = g.normal(spec.num_labels)
    return w


def _synth_deit(spec: ModelSpec, g: _Gen) -> Dict[str, np.ndarray]:
    h, inter = spec.hidden, spec.inter
    w = {}
    w["cls_token"] = g.normal(1, 1, h)
    # present in real timm checkpoints but never read by the reference (deit.py:119-124)
    w["dist_token"] = g.normal(1, 1, h)
    w["pos_embed"] = g.normal(1, spec.tokens, h)
    w["patch_embed.proj.weight"] = g.normal(h, spec.channels, spec.patch, spec.patch)
    w["patch_embed.proj.bias"] = g.normal(h)
    for i in range(spec.blocks):
        root = f"blocks.{i}."
        w[root + "norm1.weight"] = g.ln_scale(h)
        w[root + "norm1.bias"] = g.normal(h)
        w[root + "attn.qkv.weight"] = g.normal(3 * h, h)
        w[root + "attn.qkv.bias"] = g.normal(3 * h)
        w[root + "attn.proj.weight"] = g.normal(h, h)
        w[root + "attn.proj.bias"] = g.normal(h)
        w[root + "norm2.weight"] = g.ln_scale(h)
        w[root + "norm2.bias"] = g.normal(h)
        w[root + "mlp.fc1.weight"] = g.normal(inter, h)
        w[root + "mlp.fc1.bias"] = g.normal(inter)
        w[root + "mlp.fc2.weight"] = g.normal(h, inter)
        w[root + "mlp.fc2.bias"] = g.normal(h)
    w["norm.weight"] = g.ln_scale(h)
    w["norm.bias"] = g.normal(h)
    w["head.weight"] = g.normal(spec.num_labels, h)
    w["head.bias"] = g.normal(spec.num_labels)
    return w


def _synth_bert(spec: ModelSpec, g: _Gen) -> Dict[str, np.ndarray]:
    h, inter = spec.hidden, spec.inter
    # `BertShardForSequenceClassification` strips a 'bert.' prefix (bert.py:191-196);
    # the bare `BertModelShard` registry entries read un-prefixed keys.
    p = "bert." if spec.classify else ""
    w = {}
    w[p + "embeddings.position_ids"] = np.arange(spec.max_pos, dtype=np.int64)[None, :]
    w[p + "embeddings.word_embeddings.weight"] = g.normal(spec.vocab, h)
    w[p + "embeddings.position_embeddings.weight"] = g.normal(spec.max_pos, h)
    w[p + "embeddings.token_type_embeddings.weight"] = g.normal(spec.type_vocab, h)
    w[p + "embeddings.LayerNorm.weight"] = g.ln_scale(h)
    w[p + "embeddings.LayerNorm.bias"] = g.normal(h)
    for i in range(spec.blocks):
        root = p + f"encoder.layer.{i}."
        for name in ("query", "key", "value"):
            w[root + f"attention.self.{name}.weight"] = g.normal(h, h)
            w[root + f"attention.self.{name}.bias"] = g.normal(h)
        w[root + "attention.output.dense.weight"] = g.normal(h, h)
        w[root + "attention.output.dense.bias"] = g.normal(h)
        w[root + "attention.output.LayerNorm.weight"] = g.ln_scale(h)
        w[root + "attention.output.LayerNorm.bias"] = g.normal(h)
        w[root + "intermediate.dense.weight"] = g.normal(inter, h)
        w[root + "intermediate.dense.bias"] = g.normal(inter)
        w[root + "output.dense.weight"] = g.normal(h, inter)
        w[root + "output.dense.bias"] = g.normal(h)
        w[root + "output.LayerNorm.weight"] = g.ln_scale(h)
        w[root + "output.LayerNorm.bias"] = g.normal(h)
    w[p + "pooler.dense.weight"] = g.normal(h, h)
    w[p + "pooler.dense.bias"] = g.normal(h)
    if spec.classify:
        w["classifier.weight"] = g.normal(spec.num_labels, h)
        w["classifier.bias"] = g.normal(spec.num_labels)
    return w


def synth_weights(spec: ModelSpec, seed: int = 0, std: float = 0.02) -> Dict[str, np.ndarray]:
    """Seeded synthetic weights (fp32) in the reference npz layout of `spec.family`."""
    g = _Gen(seed, std)
    if spec.family == 'vit':
        return _synth_vit(spec, g)
    if spec.family == 'deit':
        return _synth_deit(spec, g)
    if spec.family == 'bert':
        return _synth_bert(spec, g)
    raise ValueError(f"unknown family: {spec.family}")


def synth_input(spec: ModelSpec, ubatch: int, seed: int = 1, seq_len: int = 128):
    """Seeded synthetic stage-0 input: images `[B,3,H,W]` fp32 or token ids `[B,S]` int64."""
    # pylint: disable=import-outside-toplevel
    import torch
    gen = torch.Generator().manual_seed(seed)
    if spec.family == 'bert':
        return torch.randint(0, spec.vocab, (ubatch, min(seq_len, spec.max_pos)), generator=gen)
    return torch.randn(ubatch, spec.channels, spec.image_size, spec.image_size, generator=gen)
