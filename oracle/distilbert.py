"""ORACLE (test infrastructure, not product code): CPU restatement of the classifier forward.

The reference's arithmetic for this path lives in a third-party, un-vendored dependency:
``transformers==4.40.0`` + ``peft==0.10.0`` (reference requirements.txt:6-7), reached from
emotion_analysis/modeling.py:14-21 and emotion_analysis/inference.py:16-20.  This file restates the
published DistilBertForSequenceClassification algorithm:

  embeddings  : word_embeddings[id] + position_embeddings[t] -> LayerNorm(eps=1e-12)
  6 x block   : q/k/v_lin, softmax(q k^T / sqrt(64) + padding mask) v, out_lin, LN(sa + x),
                lin2(gelu_erf(lin1(.))), LN(ffn + sa_out)                       (post-LN)
  head        : hidden[:, 0] -> pre_classifier -> ReLU -> classifier -> argmax (inference.py:20)
  LoRA        : PEFT merge rule W' = W + (lora_alpha / r) * B @ A on q_lin and v_lin
                (Scripts/finetuneDistillBert.ipynb:790-795: r=8, lora_alpha=16 -> scale 2)

Pin: validated against the installed transformers DistilBertForSequenceClassification on the same
random weights (tests/test_oracle_cpu.py; fixtures from oracle/make_golden.py).  The real
fine-tuned weights / tokenizer / peft are unreachable offline, so parity w.r.t. the deployed model is
UNPINNED.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

from .gpt_kv import gelu_erf, layer_norm

BERT_LN_EPS = 1e-12
N_LAYERS = 6


def merge_lora(weight: torch.Tensor, lora_a: torch.Tensor, lora_b: torch.Tensor, alpha: float = 16.0,
               r: int = 8) -> torch.Tensor:
    """W [out,in], A [r,in], B [out,r] -> W + (alpha/r) B A."""
    return weight + (alpha / r) * (lora_b @ lora_a)


@torch.no_grad()
def forward(sd: Dict[str, torch.Tensor], input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
            n_heads: int = 12, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """HF-layout state dict, ids [N,T] -> logits [N, num_labels]."""
    g = lambda k: sd[k].to(dtype)
    N, T = input_ids.shape
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    x = g("distilbert.embeddings.word_embeddings.weight")[input_ids] \
        + g("distilbert.embeddings.position_embeddings.weight")[:T]
    x = layer_norm(x, g("distilbert.embeddings.LayerNorm.weight"), g("distilbert.embeddings.LayerNorm.bias"),
                   BERT_LN_EPS)
    d = x.shape[-1]
    hd = d // n_heads
    neg = torch.finfo(dtype).min
    key_mask = (attention_mask == 0).view(N, 1, 1, T)
    n_layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("distilbert.transformer.layer."))
    for i in range(n_layers):
        p = f"distilbert.transformer.layer.{i}."
        lin = lambda name, t: t @ g(p + name + ".weight").T + g(p + name + ".bias")
        q = lin("attention.q_lin", x).view(N, T, n_heads, hd).transpose(1, 2)
        k = lin("attention.k_lin", x).view(N, T, n_heads, hd).transpose(1, 2)
        v = lin("attention.v_lin", x).view(N, T, n_heads, hd).transpose(1, 2)
        s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
        s = s.masked_fill(key_mask, neg)
        ctx = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(N, T, d)
        sa = lin("attention.out_lin", ctx)
        x = layer_norm(sa + x, g(p + "sa_layer_norm.weight"), g(p + "sa_layer_norm.bias"), BERT_LN_EPS)
        f = lin("ffn.lin2", gelu_erf(lin("ffn.lin1", x)))
        x = layer_norm(f + x, g(p + "output_layer_norm.weight"), g(p + "output_layer_norm.bias"), BERT_LN_EPS)
    pooled = torch.relu(x[:, 0] @ g("pre_classifier.weight").T + g("pre_classifier.bias"))
    return pooled @ g("classifier.weight").T + g("classifier.bias")


def predict_ids(sd: Dict[str, torch.Tensor], input_ids: torch.Tensor,
                attention_mask: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(label ids [N], logits [N,28]) -- argmax rule of emotion_analysis/inference.py:20."""
    logits = forward(sd, input_ids, attention_mask)
    return torch.argmax(logits, dim=1), logits
