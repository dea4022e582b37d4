"""Classifier checkpoint contract: HF DistilBertForSequenceClassification layout + PEFT LoRA merge.

Key names / shapes are the HF layout the reference loads through
``AutoModelForSequenceClassification.from_pretrained(REPO_ID, num_labels=28)``
(reference emotion_analysis/modeling.py:14-21, emotion_analysis/config.py:3,36); the adapter is
LoRA r=8, alpha=16 on q_lin / v_lin (reference Scripts/finetuneDistillBert.ipynb:790-795), merged as
``W + (alpha / r) * B @ A`` at load time.  The fine-tuned weights live on the HF hub and are not
reachable offline, so ``make_bert_state_dict`` builds a random-init set of the same shapes.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, Tuple

import torch


@dataclasses.dataclass(frozen=True)
class BertGeometry:
    vocab_size: int = 30522
    max_pos: int = 512
    dim: int = 768
    n_heads: int = 12
    n_layers: int = 6
    hidden_dim: int = 3072
    num_labels: int = 28

    @property
    def head_dim(self) -> int:
        return self.dim // self.n_heads


DISTILBERT_BASE = BertGeometry()
TINY_BERT = BertGeometry(vocab_size=200, max_pos=32, dim=128, n_heads=2, n_layers=2, hidden_dim=512, num_labels=28)

# reference emotion_analysis/config.py:5-34
ID2LABEL = {
    0: "admiration", 1: "amusement", 2: "anger", 3: "annoyance", 4: "approval", 5: "caring",
    6: "confusion", 7: "curiosity", 8: "desire", 9: "disappointment", 10: "disapproval",
    11: "disgust", 12: "embarrassment", 13: "excitement", 14: "fear", 15: "gratitude", 16: "grief",
    17: "joy", 18: "love", 19: "nervousness", 20: "optimism", 21: "pride", 22: "realization",
    23: "relief", 24: "remorse", 25: "sadness", 26: "surprise", 27: "neutral",
}
LORA_R, LORA_ALPHA = 8, 16


def expected_bert_keys(geo: BertGeometry) -> Dict[str, Tuple[int, ...]]:
    d, f = geo.dim, geo.hidden_dim
    out = {
        "distilbert.embeddings.word_embeddings.weight": (geo.vocab_size, d),
        "distilbert.embeddings.position_embeddings.weight": (geo.max_pos, d),
        "distilbert.embeddings.LayerNorm.weight": (d,),
        "distilbert.embeddings.LayerNorm.bias": (d,),
        "pre_classifier.weight": (d, d), "pre_classifier.bias": (d,),
        "classifier.weight": (geo.num_labels, d), "classifier.bias": (geo.num_labels,),
    }
    for i in range(geo.n_layers):
        p = f"distilbert.transformer.layer.{i}."
        for lin in ("q_lin", "k_lin", "v_lin", "out_lin"):
            out[p + f"attention.{lin}.weight"] = (d, d)
            out[p + f"attention.{lin}.bias"] = (d,)
        out[p + "sa_layer_norm.weight"] = (d,)
        out[p + "sa_layer_norm.bias"] = (d,)
        out[p + "ffn.lin1.weight"] = (f, d)
        out[p + "ffn.lin1.bias"] = (f,)
        out[p + "ffn.lin2.weight"] = (d, f)
        out[p + "ffn.lin2.bias"] = (d,)
        out[p + "output_layer_norm.weight"] = (d,)
        out[p + "output_layer_norm.bias"] = (d,)
    return out


def infer_bert_geometry(sd: Dict[str, torch.Tensor], n_heads: int = 12) -> BertGeometry:
    we = sd["distilbert.embeddings.word_embeddings.weight"]
    n_layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("distilbert.transformer.layer."))
    return BertGeometry(
        vocab_size=we.shape[0], max_pos=sd["distilbert.embeddings.position_embeddings.weight"].shape[0],
        dim=we.shape[1], n_heads=n_heads, n_layers=n_layers,
        hidden_dim=sd["distilbert.transformer.layer.0.ffn.lin1.weight"].shape[0],
        num_labels=sd["classifier.weight"].shape[0])


def make_bert_state_dict(geo: BertGeometry = DISTILBERT_BASE, seed: int = 0, with_lora: bool = True,
                         head_std: float = 2.0) -> Dict[str, torch.Tensor]:
    """Random-init HF-layout state dict (+ ``*.lora_A.weight`` / ``*.lora_B.weight`` adapters).

    Encoder matrices ~ N(0, 1/fan_in) (unit gain, activations stay O(1) through six layers); the two
    head matrices use the wider ``head_std`` so that class logits are separated by far more than
    bf16 rounding noise (the label-identity test would otherwise be a coin flip on random weights).
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for k, shape in expected_bert_keys(geo).items():
        if k.endswith("LayerNorm.weight") or k.endswith("layer_norm.weight"):
            sd[k] = 1.0 + 0.05 * torch.randn(shape, generator=g)
        elif len(shape) == 1:
            sd[k] = 0.02 * torch.randn(shape, generator=g)
        elif k.startswith("pre_classifier") or k.startswith("classifier"):
            sd[k] = head_std / (shape[1] ** 0.5) * torch.randn(shape, generator=g)
        elif "embeddings" in k:
            sd[k] = 0.5 * torch.randn(shape, generator=g)
        else:
            sd[k] = (1.0 / shape[1] ** 0.5) * torch.randn(shape, generator=g)
    if with_lora:
        for i in range(geo.n_layers):
            for lin in ("q_lin", "v_lin"):
                base = f"distilbert.transformer.layer.{i}.attention.{lin}"
                sd[base + ".lora_A.weight"] = 0.02 * torch.randn((LORA_R, geo.dim), generator=g)
                sd[base + ".lora_B.weight"] = 0.02 * torch.randn((geo.dim, LORA_R), generator=g)
    return sd


def merge_lora_state_dict(sd: Dict[str, torch.Tensor], alpha: float = LORA_ALPHA, r: int = LORA_R) -> Dict[str, torch.Tensor]:
    """Fold ``*.lora_A.weight`` / ``*.lora_B.weight`` into their base weights; returns plain HF keys.

    Accepts every spelling the reference's load path can produce (emotion_analysis/modeling.py:14-21: a base model wrapped
    by ``PeftModel.from_pretrained``; the adapter was trained with ``TaskType.SEQ_CLS``, Scripts/finetuneDistillBert.ipynb:790-795,
    so PEFT also saves the two classification heads as ``modules_to_save``):
      * adapter file on disk:      ``base_model.model.<m>.lora_A.weight``, ``base_model.model.classifier.weight``
      * live ``PeftModel.state_dict()``: ``base_model.model.<m>.base_layer.weight``, ``<m>.lora_A.default.weight``,
        ``classifier.original_module.weight`` (the UNTRAINED base head) and ``classifier.modules_to_save.default.weight``
        (the fine-tuned head)
      * a plain dict merge ``{**base, **adapter}`` or ``{**adapter, **base}``: the adapter's head wins in both orders.
    Priority for one target key: modules_to_save > adapter-file spelling (had the ``base_model.model.`` prefix) > plain key;
    ``original_module`` entries are dropped.
    """
    clean: Dict[str, torch.Tensor] = {}
    rank: Dict[str, int] = {}
    loras: Dict[str, Dict[str, torch.Tensor]] = {}
    for k0, v in sd.items():
        k = k0
        prio = 0
        if k.startswith("base_model.model."):
            k = k[len("base_model.model."):]
            prio = 1
        if ".original_module." in k:
            continue                                         # the frozen copy of a modules_to_save head
        if ".modules_to_save." in k:
            head, tail = k.split(".modules_to_save.", 1)     # tail = "<adapter>.weight" or "weight"
            k = head + "." + tail.split(".")[-1]
            prio = 2
        k = k.replace(".base_layer.", ".")
        for ab in ("lora_A", "lora_B"):                      # "<m>.lora_A.<adapter>.weight" -> "<m>.lora_A.weight"
            if f".{ab}." in k:
                k = k.split(f".{ab}.")[0] + f".{ab}.weight"
        if k.endswith(".lora_A.weight"):
            loras.setdefault(k[: -len(".lora_A.weight")], {})["A"] = v
        elif k.endswith(".lora_B.weight"):
            loras.setdefault(k[: -len(".lora_B.weight")], {})["B"] = v
        elif k not in clean or prio >= rank[k]:
            clean[k] = v
            rank[k] = prio
    for base, ab in loras.items():
        if "A" not in ab or "B" not in ab:
            raise KeyError(f"incomplete LoRA pair for {base}")
        w = clean[base + ".weight"].to(torch.float32)
        clean[base + ".weight"] = w + (alpha / r) * (ab["B"].to(torch.float32) @ ab["A"].to(torch.float32))
    return clean
