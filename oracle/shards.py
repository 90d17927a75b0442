"""CPU oracle: ViT / DeiT / BERT shard forward in fp32 - TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates, with plain torch CPU fp32 ops, what the reference computes when it runs
`{ViT,DeiT,Bert}Shard...forward` (`vit.py:55-70,161-170,220-226`, `deit.py:54-69,158-167,220-226`,
`bert.py:41-52,142-151,203-209`) over the HuggingFace modules it instantiates. Weights are read from
the reference's own npz key layouts (`vit.py:120-159`, `deit.py:119-156`, `bert.py:104-140`).
"""
import math
from typing import Mapping, Tuple, Union
import numpy as np
import torch
import torch.nn.functional as F

ShardData = Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]


def _t(arr) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(arr)).float()


def block_params(family: str, weights: Mapping, layer_id: int, hidden: int) -> dict:
    """Canonical per-block tensors (`nn.Linear` layout: weight `[out,in]`) from an npz layout.

    Restates `_load_weights_layer`: ViT `vit.py:138-159` (JAX kernels are `[in,out]` and per-head,
    hence `.view(H,H).t()`), DeiT `deit.py:131-156` (fused qkv split by rows), BERT `bert.py:118-140`.
    """
    p = {}
    if family == 'vit':
        root = f"Transformer/encoderblock_{layer_id}/"
        att = root + "MultiHeadDotProductAttention_1/"
        p['ln1_w'], p['ln1_b'] = _t(weights[root + "LayerNorm_0/scale"]), _t(weights[root + "LayerNorm_0/bias"])
        for k, name in (('q', 'query'), ('k', 'key'), ('v', 'value')):
            p['w' + k] = _t(weights[att + name + "/kernel"]).view(hidden, hidden).t().contiguous()
            p['b' + k] = _t(weights[att + name + "/bias"]).view(-1)
        p['wo'] = _t(weights[att + "out/kernel"]).view(hidden, hidden).t().contiguous()
        p['bo'] = _t(weights[att + "out/bias"]).view(-1)
        p['ln2_w'], p['ln2_b'] = _t(weights[root + "LayerNorm_2/scale"]), _t(weights[root + "LayerNorm_2/bias"])
        p['w1'] = _t(weights[root + "MlpBlock_3/Dense_0/kernel"]).t().contiguous()
        p['b1'] = _t(weights[root + "MlpBlock_3/Dense_0/bias"])
        p['w2'] = _t(weights[root + "MlpBlock_3/Dense_1/kernel"]).t().contiguous()
        p['b2'] = _t(weights[root + "MlpBlock_3/Dense_1/bias"])
    elif family == 'deit':
        root = f"blocks.{layer_id}."
        p['ln1_w'], p['ln1_b'] = _t(weights[root + "norm1.weight"]), _t(weights[root + "norm1.bias"])
        qkv_w, qkv_b = _t(weights[root + "attn.qkv.weight"]), _t(weights[root + "attn.qkv.bias"])
        for i, k in enumerate('qkv'):
            p['w' + k] = qkv_w[i * hidden:(i + 1) * hidden, :]
            p['b' + k] = qkv_b[i * hidden:(i + 1) * hidden]
        p['wo'], p['bo'] = _t(weights[root + "attn.proj.weight"]), _t(weights[root + "attn.proj.bias"])
        p['ln2_w'], p['ln2_b'] = _t(weights[root + "norm2.weight"]), _t(weights[root + "norm2.bias"])
        p['w1'], p['b1'] = _t(weights[root + "mlp.fc1.weight"]), _t(weights[root + "mlp.fc1.bias"])
        p['w2'], p['b2'] = _t(weights[root + "mlp.fc2.weight"]), _t(weights[root + "mlp.fc2.bias"])
    elif family == 'bert':
        root = f"encoder.layer.{layer_id}."
        for k, name in (('q', 'query'), ('k', 'key'), ('v', 'value')):
            p['w' + k] = _t(weights[root + f"attention.self.{name}.weight"])
            p['b' + k] = _t(weights[root + f"attention.self.{name}.bias"])
        p['wo'], p['bo'] = _t(weights[root + "attention.output.dense.weight"]), _t(weights[root + "attention.output.dense.bias"])
        p['ln1_w'] = _t(weights[root + "attention.output.LayerNorm.weight"])
        p['ln1_b'] = _t(weights[root + "attention.output.LayerNorm.bias"])
        p['w1'], p['b1'] = _t(weights[root + "intermediate.dense.weight"]), _t(weights[root + "intermediate.dense.bias"])
        p['w2'], p['b2'] = _t(weights[root + "output.dense.weight"]), _t(weights[root + "output.dense.bias"])
        p['ln2_w'] = _t(weights[root + "output.LayerNorm.weight"])
        p['ln2_b'] = _t(weights[root + "output.LayerNorm.bias"])
    else:
        raise ValueError(family)
    return p


def self_attention(x: torch.Tensor, p: dict, heads: int) -> torch.Tensor:
    """Unmasked, non-causal multi-head attention; returns the merged-head context `[B,S,H]`.

    HF `ViTSelfAttention.forward` / `BertSelfAttention.forward` + `eager_attention_forward`:
    `softmax(Q K^T * d^-0.5) V`; the reference passes no mask (`vit.py:60`, `bert.py:45`).
    """
    b, s, h = x.shape
    d = h // heads
    q = F.linear(x, p['wq'], p['bq']).view(b, s, heads, d).transpose(1, 2)
    k = F.linear(x, p['wk'], p['bk']).view(b, s, heads, d).transpose(1, 2)
    v = F.linear(x, p['wv'], p['bv']).view(b, s, heads, d).transpose(1, 2)
    scores = torch.matmul(q, k.transpose(2, 3)) * (d ** -0.5)
    probs = F.softmax(scores, dim=-1)
    ctx = torch.matmul(probs, v).transpose(1, 2).contiguous()
    return ctx.reshape(b, s, h)


def block_sublayers(family: str, data: ShardData, p: dict, heads: int, eps: float,
                    sub_start: int, sub_end: int) -> ShardData:
    """Sub-layers `sub_start..sub_end` (0-based, inclusive) of one transformer block.

    ViT/DeiT are pre-LN (`vit.py:55-70`): 0 = LN_before + attention -> (ctx, skip); 1 = out-proj
    `+= skip`; 2 = LN_after + FC1 + exact-erf GELU -> (inter, skip); 3 = FC2 + skip.
    BERT is post-LN (`bert.py:41-52`): 0 = attention -> (ctx, x); 1 = LN(out-proj + x);
    2 = GELU(FC1) -> (inter, x); 3 = LN(FC2 + x).
    """
    has = lambda i: sub_start <= i <= sub_end  # noqa: E731
    h = p['wo'].shape[0]
    if family in ('vit', 'deit'):
        if has(0):
            data = (self_attention(F.layer_norm(data, (h,), p['ln1_w'], p['ln1_b'], eps), p, heads), data)
        if has(1):
            data = F.linear(data[0], p['wo'], p['bo']) + data[1]
        if has(2):
            norm = F.layer_norm(data, (h,), p['ln2_w'], p['ln2_b'], eps)
            data = (F.gelu(F.linear(norm, p['w1'], p['b1'])), data)
        if has(3):
            data = F.linear(data[0], p['w2'], p['b2']) + data[1]
    else:
        if has(0):
            data = (self_attention(data, p, heads), data)
        if has(1):
            data = F.layer_norm(F.linear(data[0], p['wo'], p['bo']) + data[1], (h,), p['ln1_w'], p['ln1_b'], eps)
        if has(2):
            data = (F.gelu(F.linear(data, p['w1'], p['b1'])), data)
        if has(3):
            data = F.layer_norm(F.linear(data[0], p['w2'], p['b2']) + data[1], (h,), p['ln2_w'], p['ln2_b'], eps)
    return data


def sublayer_ranges(layer_start: int, layer_end: int):
    """Map the 1-based inclusive `[layer_start, layer_end]` onto (block id, sub_start, sub_end).

    Restates the loop in `ViTModelShard._build_shard` (`vit.py:99-113`; same in deit/bert).
    """
    out = []
    cur = layer_start
    while cur <= layer_end:
        block = math.ceil(cur / 4) - 1
        s0 = (cur - 1) % 4
        s1 = (layer_end - 1) % 4 if block == math.ceil(layer_end / 4) - 1 else 3
        out.append((block, s0, s1))
        cur += s1 - s0 + 1
    return out


def embedding_params(spec, weights: Mapping) -> dict:
    """Stage-0 embedding tensors converted once (`_load_weights_first`: `vit.py:120-130`, `deit.py:119-124`,
    `bert.py:104-111`)."""
    if spec.family == 'vit':
        return {'conv_w': _t(np.transpose(weights["embedding/kernel"], [3, 2, 0, 1])), 'conv_b': _t(weights["embedding/bias"]),
                'cls': _t(weights["cls"]), 'pos': _t(weights["Transformer/posembed_input/pos_embedding"])}
    if spec.family == 'deit':
        # the reference never loads the distillation token: it stays zeros (deit.py:119-124)
        return {'conv_w': _t(weights["patch_embed.proj.weight"]), 'conv_b': _t(weights["patch_embed.proj.bias"]),
                'cls': _t(weights["cls_token"]), 'pos': _t(weights["pos_embed"])}
    return {'pos_ids': torch.from_numpy(np.asarray(weights["embeddings.position_ids"])),
            'word': _t(weights["embeddings.word_embeddings.weight"]),
            'type': _t(weights["embeddings.token_type_embeddings.weight"]),
            'pos': _t(weights["embeddings.position_embeddings.weight"]),
            'ln_w': _t(weights["embeddings.LayerNorm.weight"]), 'ln_b': _t(weights["embeddings.LayerNorm.bias"])}


def embeddings(spec, e: dict, data: torch.Tensor) -> torch.Tensor:
    """Stage-0 embeddings: HF `ViTEmbeddings` / `DeiTEmbeddings` / `BertEmbeddings` (eval mode) over the
    tensors of `embedding_params`."""
    if spec.family in ('vit', 'deit'):
        x = F.conv2d(data, e['conv_w'], e['conv_b'], stride=spec.patch).flatten(2).transpose(1, 2)
        cls = e['cls'].expand(x.shape[0], -1, -1)
        prefix = (cls,) if spec.family == 'vit' else (cls, torch.zeros_like(cls))
        return torch.cat(prefix + (x,), dim=1) + e['pos']
    s = data.shape[1]
    x = F.embedding(data, e['word']) + e['type'][0]
    x = x + F.embedding(e['pos_ids'][:, :s], e['pos'])
    return F.layer_norm(x, (spec.hidden,), e['ln_w'], e['ln_b'], spec.eps)


def _bert_inner(spec, weights: Mapping) -> Mapping:
    if spec.family == 'bert' and spec.classify:
        return {k[len('bert.'):]: v for k, v in weights.items() if k.startswith('bert.')}
    return weights


class PreparedShard:
    """A shard with its weights converted once (what the reference's constructors do at load time), so that
    repeated forwards cost only the arithmetic. Used for the timed CPU baseline in bench.py."""

    def __init__(self, spec, weights: Mapping, layer_start: int, layer_end: int):
        self.spec = spec
        self.layer_start, self.layer_end = layer_start, layer_end
        self.is_first, self.is_last = layer_start == 1, layer_end == spec.layers   # model_cfg.py:87-90
        inner = _bert_inner(spec, weights)
        self.ranges = sublayer_ranges(layer_start, layer_end)
        self.params = {block: block_params(spec.family, inner, block, spec.hidden) for block, _, _ in self.ranges}
        self.embed_weights = embedding_params(spec, inner) if self.is_first else None
        self.head = None
        if self.is_last:
            if spec.family == 'vit':
                self.head = (_t(weights["Transformer/encoder_norm/scale"]), _t(weights["Transformer/encoder_norm/bias"]),
                             _t(np.transpose(weights["head/kernel"])), _t(weights["head/bias"]))
            elif spec.family == 'deit':
                self.head = (_t(weights["norm.weight"]), _t(weights["norm.bias"]), _t(weights["head.weight"]),
                             _t(weights["head.bias"]))
            else:
                self.head = (_t(inner["pooler.dense.weight"]), _t(inner["pooler.dense.bias"]),
                             _t(weights["classifier.weight"]) if spec.classify else None,
                             _t(weights["classifier.bias"]) if spec.classify else None)

    @torch.no_grad()
    def forward(self, data: ShardData, boundaries: dict = None) -> ShardData:
        """Forward of sub-layers `[layer_start, layer_end]`; see `shard_forward`."""
        spec = self.spec
        if self.is_first:
            data = embeddings(spec, self.embed_weights, data)
        for block, s0, s1 in self.ranges:
            p = self.params[block]
            for sub in range(s0, s1 + 1):
                data = block_sublayers(spec.family, data, p, spec.heads, spec.eps, sub, sub)
                if boundaries is not None:
                    boundaries[block * 4 + sub + 1] = data
        if self.is_last:
            if spec.family in ('vit', 'deit'):
                ln_w, ln_b, head_w, head_b = self.head
                data = F.layer_norm(data, (spec.hidden,), ln_w, ln_b, spec.eps)
                data = F.linear(data[:, 0, :], head_w, head_b)
            else:
                pool_w, pool_b, cls_w, cls_b = self.head
                data = torch.tanh(F.linear(data[:, 0], pool_w, pool_b))
                if cls_w is not None:
                    data = F.linear(data, cls_w, cls_b)
        return data


def shard_forward(spec, weights: Mapping, layer_start: int, layer_end: int, data: ShardData,
                  boundaries: dict = None) -> ShardData:
    """Forward of the shard covering sub-layers `[layer_start, layer_end]` (1-based, inclusive).

    `is_first`/`is_last` follow `model_cfg.py:87-90`. If `boundaries` is a dict, the value after
    every sub-layer `l` is recorded under key `l` (tuples kept as tuples).
    """
    return PreparedShard(spec, weights, layer_start, layer_end).forward(data, boundaries)
