"""ORACLE (test infrastructure, not product code): CPU restatement of the reference NO-cache path.

Follows reference generate_music/generate.py model (A) (same code at api.py:41-92,
generate_adi.py:29-63, generate_music/generate2.py:11-41):
  * GPT.forward  generate.py:25-35   fc(tr(emb(x) + pos[:T])) with nn.TransformerEncoder defaults:
                                     POST-LN, ReLU, no attention mask (bidirectional), true positions.
  * sample       generate.py:46-61   full recompute of the whole sequence every step; same sampler as
                                     the KV path except ``if top_k:`` (0/None disables top-k).

(A) and (B) are different functions of the same checkpoint; this oracle backs the
engine's "recompute mode" and the secondary CPU baseline.  Weight names are the remapped (B) names
so one weight set serves both oracles: layers.N.ln1/ln2 are the trainer's norm1/norm2.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch

from .gpt_kv import layer_norm, topk_probs


class NoCacheModelOracle:
    def __init__(self, sd: Dict[str, torch.Tensor], n_head: int, dtype: torch.dtype = torch.float32):
        self.sd = {k: v.detach().to(dtype) for k, v in sd.items()}
        self.n_head = n_head
        self.d_model = self.sd["pos_emb"].shape[1]
        self.pos_rows = self.sd["pos_emb"].shape[0]
        self.n_layer = 1 + max(int(k.split(".")[1]) for k in self.sd if k.startswith("layers."))
        self.head_dim = self.d_model // n_head

    def forward(self, idx: torch.Tensor) -> torch.Tensor:
        sd, d, H, hd = self.sd, self.d_model, self.n_head, self.head_dim
        B, T = idx.shape
        if T > self.pos_rows:
            raise RuntimeError(f"sequence of {T} tokens exceeds the {self.pos_rows}-row position table")
        x = sd["tok_emb.weight"][idx] + sd["pos_emb"][:T]
        for i in range(self.n_layer):
            p = f"layers.{i}."
            w_in, b_in = sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]
            qkv = x @ w_in.T + b_in
            q, k, v = qkv.split(d, dim=-1)
            qh = q.view(B, T, H, hd).transpose(1, 2)
            kh = k.view(B, T, H, hd).transpose(1, 2)
            vh = v.view(B, T, H, hd).transpose(1, 2)
            att = torch.softmax((qh @ kh.transpose(-1, -2)) / math.sqrt(hd), dim=-1)   # unmasked
            o = (att @ vh).transpose(1, 2).reshape(B, T, d)
            sa = o @ sd[p + "attn.out_proj.weight"].T + sd[p + "attn.out_proj.bias"]
            x = layer_norm(x + sa, sd[p + "ln1.weight"], sd[p + "ln1.bias"])           # post-LN
            ff = torch.relu(x @ sd[p + "mlp.0.weight"].T + sd[p + "mlp.0.bias"])
            ff = ff @ sd[p + "mlp.2.weight"].T + sd[p + "mlp.2.bias"]
            x = layer_norm(x + ff, sd[p + "ln2.weight"], sd[p + "ln2.bias"])
        return x @ sd["head.weight"].T + sd["head.bias"]


@torch.no_grad()
def sample_ids(model: NoCacheModelOracle, prompt_ids: Sequence[int], max_len: int = 512,
               temperature: float = 1.0, top_k: Optional[int] = 50, eos_id: int = -1,
               generator: Optional[torch.Generator] = None, return_logits: bool = False):
    """Loop of generate.py:46-61 on integer ids (``top_k`` of 0/None disables the mask)."""
    ids = torch.tensor(list(prompt_ids), dtype=torch.long).unsqueeze(0)
    step_logits = []
    for _ in range(max_len - len(prompt_ids)):
        row = model.forward(ids)[0, -1]
        if return_logits:
            step_logits.append(row.clone())
        probs = topk_probs(row, temperature, top_k if top_k else None)
        nxt = torch.multinomial(probs, 1, generator=generator)
        ids = torch.cat([ids, nxt.view(1, 1)], dim=1)
        if int(nxt) == eos_id:
            break
    out = ids[0].tolist()
    return (out, torch.stack(step_logits)) if return_logits else out
