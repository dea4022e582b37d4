"""Generator checkpoint contract: geometries, shape inference, key remap, synthetic checkpoints.

The on-disk layout is the reference trainers' ``torch.save({"model": state_dict, "vocab": tok2id})``
(reference train/train_large.py:159-164, train/train_large2.py:100-110).  The engine consumes the
*remapped* key names that reference api_cache.py:118-134 (``remap_state_dict``) produces.

No trained weights ship with the reference, so ``make_checkpoint``
builds a random-init state dict of exactly the trainer's shapes/keys for tests and benchmarks.
"""
from __future__ import annotations

import dataclasses
import hashlib
import re
from typing import Dict, Optional

import torch
import torch.nn as nn


@dataclasses.dataclass(frozen=True)
class Geometry:
    """Model geometry; ``d_ff`` is always ``4 * d_model`` in every reference trainer."""

    vocab_size: int
    pos_rows: int
    d_model: int
    n_head: int
    n_layer: int

    @property
    def d_ff(self) -> int:
        return 4 * self.d_model

    @property
    def head_dim(self) -> int:
        return self.d_model // self.n_head


# Table G of SURVEY.md section 8 (reference train/*.py).  vocab_size is data dependent for the first
# two; 8324 is the only code-determined size (train/train_large2.py:23-28) and is the nominal one.
GEOMETRIES: Dict[str, Geometry] = {
    "train_mini": Geometry(8324, 512, 256, 4, 2),       # train/train_mini.py:10,54-60
    "train_large": Geometry(8324, 255, 256, 8, 4),      # train/train_large.py:5,11-13,125-136
    "train_large_pos512": Geometry(8324, 512, 256, 8, 4),  # config 4: 256-token prompts need >255 rows
    "train_large2": Geometry(8324, 511, 512, 8, 6),     # train/train_large2.py:10-12,83-94
    "train_no_inst": Geometry(8579, 511, 512, 8, 6),    # train/train_no_inst.py:9-13,95-105
    "tiny": Geometry(96, 24, 64, 2, 2),                 # test-only: small enough for exhaustive checks
    "tiny_hd64": Geometry(128, 40, 128, 2, 2),          # test-only: head_dim 64 path
}


def infer_geometry(state_dict: Dict[str, torch.Tensor], n_head: int) -> Geometry:
    """Shape inference rule of reference api_cache.py:31-37 (``n_head`` cannot be inferred: :112)."""
    layer_keys = [k for k in state_dict if "tr.layers." in k]
    if not layer_keys:
        raise KeyError("state dict has no 'tr.layers.N.*' tensors")
    n_layer = max(int(k.split(".")[2]) for k in layer_keys) + 1
    pos = state_dict["pos"]
    vocab_size = state_dict["emb.weight"].shape[0]
    d_model = int(pos.shape[1])
    if d_model % n_head:
        raise ValueError(f"d_model {d_model} is not divisible by n_head {n_head}")
    return Geometry(int(vocab_size), int(pos.shape[0]), d_model, int(n_head), int(n_layer))


# Trainer name -> KV-cache-model name.  Rule set of reference api_cache.py:118-134: the top-level
# tensors are renamed outright, per-layer tensors keep their layer index and leaf name.
_TOP_LEVEL_RENAMES = {"emb.weight": "tok_emb.weight", "pos": "pos_emb",
                      "fc.weight": "head.weight", "fc.bias": "head.bias"}
_LAYER_MODULE_RENAMES = {"self_attn": "attn", "norm1": "ln1", "norm2": "ln2",
                         "linear1": "mlp.0", "linear2": "mlp.2"}
_LAYER_KEY = re.compile(r"^tr\.layers\.(\d+)\.(self_attn|norm1|norm2|linear1|linear2)\.(.+)$")


def remap_key(key: str) -> str:
    """Trainer key -> KV-cache model key (reference api_cache.py:118-134)."""
    if key in _TOP_LEVEL_RENAMES:
        return _TOP_LEVEL_RENAMES[key]
    m = _LAYER_KEY.match(key)
    if m is None:
        return key
    return f"layers.{m.group(1)}.{_LAYER_MODULE_RENAMES[m.group(2)]}.{m.group(3)}"


def remap_state_dict(old_sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {remap_key(k): v for k, v in old_sd.items()}


def expected_keys(geo: Geometry) -> Dict[str, tuple]:
    """Remapped key -> shape, the full set the engine requires (4 + 12 * n_layer tensors)."""
    d, v, f = geo.d_model, geo.vocab_size, geo.d_ff
    out = {
        "tok_emb.weight": (v, d),
        "pos_emb": (geo.pos_rows, d),
        "head.weight": (v, d),
        "head.bias": (v,),
    }
    for i in range(geo.n_layer):
        p = f"layers.{i}."
        out.update({
            p + "attn.in_proj_weight": (3 * d, d), p + "attn.in_proj_bias": (3 * d,),
            p + "attn.out_proj.weight": (d, d), p + "attn.out_proj.bias": (d,),
            p + "mlp.0.weight": (f, d), p + "mlp.0.bias": (f,),
            p + "mlp.2.weight": (d, f), p + "mlp.2.bias": (d,),
            p + "ln1.weight": (d,), p + "ln1.bias": (d,),
            p + "ln2.weight": (d,), p + "ln2.bias": (d,),
        })
    return out


class _TrainerShapedGPT(nn.Module):
    """Module with the trainers' parameter names (emb / pos / tr.layers.N.* / fc); init only."""

    def __init__(self, geo: Geometry):
        super().__init__()
        self.emb = nn.Embedding(geo.vocab_size, geo.d_model)
        self.pos = nn.Parameter(torch.zeros(geo.pos_rows, geo.d_model))
        blk = nn.TransformerEncoderLayer(geo.d_model, geo.n_head, geo.d_ff, batch_first=True)
        self.tr = nn.TransformerEncoder(blk, geo.n_layer, enable_nested_tensor=False)
        self.fc = nn.Linear(geo.d_model, geo.vocab_size)


def make_state_dict(geo: Geometry, seed: int = 0, aux_std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Random-init trainer-shaped fp32 state dict.

    PyTorch default inits, then ``pos``, LayerNorm affine and all biases are perturbed with
    N(0, aux_std): the reference initialises ``pos`` to zeros and LN to identity, which would leave
    the pos_emb[0]-on-decode quirk and the bias paths untested (SURVEY.md section 8c "known traps").
    """
    gen_state = torch.random.get_rng_state()
    try:
        torch.manual_seed(seed)
        m = _TrainerShapedGPT(geo)
        # nn.TransformerEncoder deep-copies one layer: every layer starts identical.  Re-draw each
        # layer's matrices so the layers differ (a trained checkpoint's would).
        g = torch.Generator().manual_seed(seed + 7919)
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        for k, v in sd.items():
            if k == "pos":
                v.copy_(torch.randn(v.shape, generator=g) * aux_std)
            elif k.endswith("norm1.weight") or k.endswith("norm2.weight"):
                v.copy_(1.0 + torch.randn(v.shape, generator=g) * aux_std)
            elif v.dim() == 1:
                v.copy_(torch.randn(v.shape, generator=g) * aux_std)
            elif k.startswith("tr.layers."):
                bound = (6.0 / (v.shape[0] + v.shape[1])) ** 0.5 if "in_proj" in k else (1.0 / v.shape[1]) ** 0.5
                v.copy_((torch.rand(v.shape, generator=g) * 2 - 1) * bound)
    finally:
        torch.random.set_rng_state(gen_state)
    return sd


def make_checkpoint(geo: Geometry, seed: int = 0, vocab: Optional[Dict[str, int]] = None) -> dict:
    """Checkpoint dict in the trainers' layout: ``{"model": state_dict, "vocab": tok2id}``."""
    from .vocab import build_synthetic_vocab

    if vocab is None:
        vocab = build_synthetic_vocab(geo.vocab_size)
    if len(vocab) != geo.vocab_size:
        raise ValueError("vocab size does not match geometry")
    return {"model": make_state_dict(geo, seed), "vocab": vocab}


def state_dict_digest(sd: Dict[str, torch.Tensor]) -> str:
    """Order-independent sha256 over names, shapes and fp32 bytes; pins golden fixtures to weights."""
    h = hashlib.sha256()
    for k in sorted(sd):
        t = sd[k].detach().to(torch.float32).contiguous()
        h.update(k.encode())
        h.update(str(tuple(t.shape)).encode())
        h.update(t.numpy().tobytes())
    return h.hexdigest()
