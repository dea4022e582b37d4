"""ORACLE (test infrastructure, not product code): CPU restatement of the reference KV-cache path.

Follows reference api_cache.py model (B):
  * GPTBlock.forward   api_cache.py:51-74   pre-LN block; the "cache" holds LN1(x) (width d_model,
                                            the same tensor for K and V) and is re-projected through
                                            W_k / W_v on every step; attention has NO mask.
  * GPTWithKV.forward  api_cache.py:87-106  x = tok_emb[idx] + pos_emb[:T]  (so every decode step,
                                            T == 1, uses pos_emb[0]); no final LayerNorm; head bias.
  * sample_kvcache     api_cache.py:159-184 prefill (logits discarded), then each iteration re-feeds
                                            the last token (first iteration: the last PROMPT token
                                            again), /temperature, top-k -1e10 additive mask, softmax,
                                            multinomial, stop on EOS.

Pinned against the reference's own classes executed from /root/reference (see oracle/refload.py and
oracle/make_golden.py; fixtures in tests/golden/).  The reference has no tests or golden vectors of
its own, so those generated fixtures are the pin.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  Plain torch CPU ops, fp32 by default (fp64 selectable as a "truth" variant).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

LN_EPS = 1e-5          # nn.LayerNorm default used by api_cache.py:42,44
MASK_VALUE = -1e10     # api_cache.py:173


def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = LN_EPS) -> torch.Tensor:
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)          # biased variance, like nn.LayerNorm
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    """Exact GELU: nn.GELU() default approximate='none' (api_cache.py:47)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


class KVModelOracle:
    """Weights are the *remapped* names (tok_emb.weight, pos_emb, layers.N.attn.*, ..., head.*)."""

    def __init__(self, sd: Dict[str, torch.Tensor], n_head: int, dtype: torch.dtype = torch.float32):
        self.dtype = dtype
        self.sd = {k: v.detach().to(dtype) for k, v in sd.items()}
        self.n_head = n_head
        self.d_model = self.sd["pos_emb"].shape[1]
        self.pos_rows = self.sd["pos_emb"].shape[0]
        self.vocab_size = self.sd["tok_emb.weight"].shape[0]
        self.n_layer = 1 + max(int(k.split(".")[1]) for k in self.sd if k.startswith("layers."))
        self.head_dim = self.d_model // n_head

    # -- one block: api_cache.py:51-74 -------------------------------------------------------------
    def _block(self, i: int, x: torch.Tensor, past: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        sd, d, H, hd = self.sd, self.d_model, self.n_head, self.head_dim
        p = f"layers.{i}."
        xn = layer_norm(x, sd[p + "ln1.weight"], sd[p + "ln1.bias"])
        cache = xn if past is None else torch.cat([past, xn], dim=1)      # [B, T_all, d]
        w_in, b_in = sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]
        q = xn @ w_in[:d].T + b_in[:d]                                     # rows [0:d] -> q
        k = cache @ w_in[d:2 * d].T + b_in[d:2 * d]                        # rows [d:2d] -> k (whole cache)
        v = cache @ w_in[2 * d:].T + b_in[2 * d:]                          # rows [2d:3d] -> v (whole cache)
        B, Tq, Tk = q.shape[0], q.shape[1], k.shape[1]
        qh = q.view(B, Tq, H, hd).transpose(1, 2)
        kh = k.view(B, Tk, H, hd).transpose(1, 2)
        vh = v.view(B, Tk, H, hd).transpose(1, 2)
        att = (qh @ kh.transpose(-1, -2)) * (1.0 / math.sqrt(hd))          # no mask (api_cache.py:68)
        att = torch.softmax(att, dim=-1)
        o = (att @ vh).transpose(1, 2).reshape(B, Tq, d)
        x = x + (o @ sd[p + "attn.out_proj.weight"].T + sd[p + "attn.out_proj.bias"])
        h = layer_norm(x, sd[p + "ln2.weight"], sd[p + "ln2.bias"])
        h = gelu_erf(h @ sd[p + "mlp.0.weight"].T + sd[p + "mlp.0.bias"])
        x = x + (h @ sd[p + "mlp.2.weight"].T + sd[p + "mlp.2.bias"])
        return x, cache

    # -- whole model: api_cache.py:87-106 ----------------------------------------------------------
    def forward(self, idx: torch.Tensor, past: Optional[List[torch.Tensor]] = None,
                true_positions: bool = False) -> Tuple[torch.Tensor, List[torch.Tensor]]:
        """``true_positions=True`` is the COUNTER-test (what the reference does NOT do)."""
        B, T = idx.shape
        if T > self.pos_rows:
            raise RuntimeError(f"sequence of {T} tokens exceeds the {self.pos_rows}-row position table")
        if true_positions and past is not None:
            start = past[0].shape[1]
            pos = self.sd["pos_emb"][start:start + T]
        else:
            pos = self.sd["pos_emb"][:T]
        x = self.sd["tok_emb.weight"][idx] + pos
        presents = []
        for i in range(self.n_layer):
            x, c = self._block(i, x, None if past is None else past[i])
            presents.append(c)
        logits = x @ self.sd["head.weight"].T + self.sd["head.bias"]
        return logits, presents


def topk_probs(logits_row: torch.Tensor, temperature: float, top_k: Optional[int]) -> torch.Tensor:
    """Sampling distribution of api_cache.py:169-177 for one row of logits ([V] -> [V] probs)."""
    z = logits_row / temperature
    if top_k is not None:
        if top_k > z.numel():
            raise RuntimeError("selected index k out of range")
        _, idxs = z.topk(top_k)
        mask = torch.full_like(z, MASK_VALUE)
        mask[idxs] = 0.0
        z = z + mask
    return torch.softmax(z, dim=-1)


@torch.no_grad()
def sample_ids(model: KVModelOracle, prompt_ids: Sequence[int], max_len: int = 512, temperature: float = 1.0,
               top_k: Optional[int] = 50, eos_id: int = -1, generator: Optional[torch.Generator] = None,
               return_logits: bool = False):
    """Batch-1 loop of api_cache.py:159-184 on integer ids.  Returns all ids including the prompt."""
    ids = torch.tensor(list(prompt_ids), dtype=torch.long).unsqueeze(0)
    _, past = model.forward(ids)                       # prefill; logits discarded (:163)
    step_logits = []
    for _ in range(max_len - ids.shape[1]):
        logits, past = model.forward(ids[:, -1:], past)    # re-feeds the last token (:167-168)
        row = logits[0, -1]
        if return_logits:
            step_logits.append(row.clone())
        probs = topk_probs(row, temperature, top_k)
        nxt = torch.multinomial(probs, 1, generator=generator)
        ids = torch.cat([ids, nxt.view(1, 1)], dim=1)
        if int(nxt) == eos_id:
            break
    out = ids[0].tolist()
    return (out, torch.stack(step_logits)) if return_logits else out


def sample_tokens(model: KVModelOracle, tok2id: Dict[str, int], prompt: Sequence[str], max_len: int = 512,
                  temperature: float = 1.0, top_k: Optional[int] = 50,
                  generator: Optional[torch.Generator] = None) -> List[str]:
    """String-token form, same call shape as the reference's sample_kvcache."""
    id2tok = {i: t for t, i in tok2id.items()}
    ids = sample_ids(model, [tok2id[t] for t in prompt], max_len, temperature, top_k,
                     tok2id.get("[END_SEQUENCE]", -1), generator)
    return [id2tok[i] for i in ids]


@torch.no_grad()
def teacher_forced_logits(model: KVModelOracle, prompt_ids: Sequence[int], forced_ids: Sequence[int],
                          n_steps: int) -> torch.Tensor:
    """Logits of decode steps 0..n_steps-1 when the sampled tokens are replaced by ``forced_ids``.

    Step 0 feeds the last prompt token (the reference's duplicate feed); step i>0 feeds
    forced_ids[i-1].  Shape [n_steps, V].
    """
    ids = torch.tensor(list(prompt_ids), dtype=torch.long).unsqueeze(0)
    _, past = model.forward(ids)
    out = []
    feed = int(prompt_ids[-1])
    for i in range(n_steps):
        logits, past = model.forward(torch.tensor([[feed]]), past)
        out.append(logits[0, -1].clone())
        if i < len(forced_ids):
            feed = int(forced_ids[i])
    return torch.stack(out)


@torch.no_grad()
def batched_decode_step_time_port(model: KVModelOracle, prompts: Sequence[Sequence[int]], n_steps: int,
                                  temperature: float = 1.0, top_k: Optional[int] = 40,
                                  generator: Optional[torch.Generator] = None) -> List[List[int]]:
    """Batched restatement of the loop (``.item()`` removed) for equal-length prompts.

    Used only as a labelled CPU baseline: the reference sampler itself is
    batch-1.  Same per-step cost structure as the reference (full-cache re-projection + cat).
    """
    lens = {len(p) for p in prompts}
    if len(lens) != 1:
        raise ValueError("batched port needs equal-length prompts")
    ids = torch.tensor([list(p) for p in prompts], dtype=torch.long)
    _, past = model.forward(ids)
    for _ in range(n_steps):
        logits, past = model.forward(ids[:, -1:], past)
        z = logits[:, -1, :] / temperature
        if top_k is not None:
            _, idxs = z.topk(top_k)
            mask = torch.full_like(z, MASK_VALUE)
            mask.scatter_(1, idxs, 0.0)
            z = z + mask
        nxt = torch.multinomial(torch.softmax(z, dim=-1), 1, generator=generator)
        ids = torch.cat([ids, nxt], dim=1)
    return ids.tolist()


@torch.no_grad()
def teacher_forced_logits_projected(model: KVModelOracle, prompt_ids: Sequence[int], forced_ids: Sequence[int], n_steps: int,
                                    want_steps: Sequence[int]) -> torch.Tensor:
    """Same numbers as ``teacher_forced_logits`` at the steps in ``want_steps``, but caching the PROJECTED K / V rows.

    The reference caches LN1(x) and re-projects the whole cache through W_k / W_v on every step (api_cache.py:60-68),
    O(T d^2) per layer and step -- minutes of CPU time at the cache lengths of BASELINE configs 3 / 4 (1030 / 4352).  K and V of
    a cached row never change (same LN1(x), same weights), so caching K = LN1(x) W_k^T + b_k and V likewise is the same
    arithmetic with the row-wise matmul done once (``tests/test_oracle_cpu.py`` pins this function against
    ``teacher_forced_logits`` in fp64).  Shape [len(want_steps), V].
    """
    sd, d, H, hd, L = model.sd, model.d_model, model.n_head, model.head_dim, model.n_layer
    want = {int(s): i for i, s in enumerate(want_steps)}
    out = torch.empty((len(want), model.vocab_size), dtype=model.dtype)
    ids = torch.tensor(list(prompt_ids), dtype=torch.long).unsqueeze(0)
    _, past = model.forward(ids)                                   # prefill exactly as the reference (bidirectional)
    Ks, Vs = [], []
    cap = past[0].shape[1] + n_steps
    for i in range(L):
        p = f"layers.{i}."
        w_in, b_in = sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]
        k = past[i][0] @ w_in[d:2 * d].T + b_in[d:2 * d]
        v = past[i][0] @ w_in[2 * d:].T + b_in[2 * d:]
        K = torch.empty((cap, d), dtype=model.dtype); V = torch.empty((cap, d), dtype=model.dtype)
        K[:k.shape[0]] = k; V[:v.shape[0]] = v
        Ks.append(K); Vs.append(V)
    T = past[0].shape[1]
    feed = int(prompt_ids[-1])
    scale = 1.0 / math.sqrt(hd)
    for step in range(n_steps):
        x = sd["tok_emb.weight"][feed] + sd["pos_emb"][0]          # pos_emb[0] on every decode step (api_cache.py:99)
        for i in range(L):
            p = f"layers.{i}."
            w_in, b_in = sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]
            xn = layer_norm(x, sd[p + "ln1.weight"], sd[p + "ln1.bias"])
            qkv = w_in @ xn + b_in
            Ks[i][T] = qkv[d:2 * d]; Vs[i][T] = qkv[2 * d:]
            q = qkv[:d].view(H, hd)
            Kh = Ks[i][:T + 1].view(T + 1, H, hd); Vh = Vs[i][:T + 1].view(T + 1, H, hd)
            att = torch.softmax(torch.einsum("hd,thd->ht", q, Kh) * scale, dim=-1)
            o = torch.einsum("ht,thd->hd", att, Vh).reshape(d)
            x = x + (sd[p + "attn.out_proj.weight"] @ o + sd[p + "attn.out_proj.bias"])
            h = layer_norm(x, sd[p + "ln2.weight"], sd[p + "ln2.bias"])
            h = gelu_erf(sd[p + "mlp.0.weight"] @ h + sd[p + "mlp.0.bias"])
            x = x + (sd[p + "mlp.2.weight"] @ h + sd[p + "mlp.2.bias"])
        if step in want:
            out[want[step]] = sd["head.weight"] @ x + sd["head.bias"]
        T += 1
        if step < len(forced_ids):
            feed = int(forced_ids[step])
    return out
