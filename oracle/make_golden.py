"""Generate tests/golden/*.npz by running the REFERENCE'S OWN code (in-container only).

    python -m oracle.make_golden            # needs /root/reference; writes tests/golden/

The reference has no tests, golden vectors or fixtures.  These files are the
pin for oracle/ and for the CUDA engine: outputs of the reference classes (AST-loaded from
/root/reference by oracle/refload.py) on seeded random-init checkpoints from
music-generation-emotion-adaptive_b200/checkpoint.py.  Weights are not stored; each fixture carries
the sha256 digest of the state dict it was produced from, and tests regenerate the weights from
(geometry, seed) and verify the digest first.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mgea_b200 as mg                                    # noqa: E402
from mgea_b200 import bert_checkpoint as bc               # noqa: E402
from oracle import refload                                # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _prompt_ids(tok2id, seed, n=1):
    if len(tok2id) >= 8324:
        return [mg.encode(tok2id, p) for p in mg.synthetic_prompts(tok2id, n, seed=seed)]
    g = torch.Generator().manual_seed(1000 + seed)       # tiny vocabularies: raw ids
    return [torch.randint(0, len(tok2id), (int(torch.randint(3, 7, (1,), generator=g)),), generator=g).tolist()
            for _ in range(n)]


def _ref_greedy(ns, model, tok2id, ids, max_len):
    id2tok = {i: t for t, i in tok2id.items()}
    ns["tok2id"], ns["id2tok"] = tok2id, id2tok
    toks = ns["sample_kvcache"](model, [id2tok[i] for i in ids], max_len=max_len, temperature=1.0, top_k=1)
    return [tok2id[t] for t in toks]


def kv_greedy():
    """Greedy (top_k=1) token sequences from reference api_cache.py:160-184."""
    ns = refload.load_kv_reference()
    cases = [("tiny", s, 24) for s in range(3)] + [("tiny_hd64", s, 40) for s in range(2)] \
        + [("train_mini", s, 512) for s in range(3)] + [("train_large", s, 255) for s in range(3)] \
        + [("train_large2", 0, 160)]
    out, meta = {}, []
    for name, seed, max_len in cases:
        geo = mg.GEOMETRIES[name]
        ck = mg.make_checkpoint(geo, seed)
        model = refload.build_kv_model(ns, ck["model"], geo.n_head)
        # remove the EOS token from the vocab view so every run goes to max_len (api_cache.py:181 -> -1)
        vocab = {t: i for t, i in ck["vocab"].items()}
        eos = vocab.get("[END_SEQUENCE]")
        prompts = _prompt_ids(vocab, seed, n=2)
        for j, ids in enumerate(prompts):
            key = f"{name}_s{seed}_p{j}"
            v2 = dict(vocab)
            if j == 0 and eos is not None:
                # run 0: EOS disabled (rename it) -> full length
                v2 = {("[EOS_DISABLED]" if t == "[END_SEQUENCE]" else t): i for t, i in vocab.items()}
            toks = _ref_greedy(ns, model, v2, ids, max_len)
            out[key + "_prompt"] = np.asarray(ids, np.int32)
            out[key + "_tokens"] = np.asarray(toks, np.int32)
            meta.append({"key": key, "geometry": name, "seed": seed, "max_len": max_len,
                         "eos_id": -1 if (j == 0 or eos is None) else int(eos),
                         "digest": mg.state_dict_digest(ck["model"])})
            print(key, len(toks))
    np.savez_compressed(os.path.join(OUT, "kv_greedy.npz"), meta=json.dumps(meta), **out)


@torch.no_grad()
def kv_logits():
    """Teacher-forced per-step logits (fp32 reference run, and the same module cast to fp64)."""
    ns = refload.load_kv_reference()
    cases = [("tiny", 0, 12), ("tiny_hd64", 1, 12), ("train_large", 0, 6), ("train_mini", 1, 6)]
    out, meta = {}, []
    for name, seed, n_steps in cases:
        geo = mg.GEOMETRIES[name]
        ck = mg.make_checkpoint(geo, seed)
        model = refload.build_kv_model(ns, ck["model"], geo.n_head)
        model64 = refload.build_kv_model(ns, ck["model"], geo.n_head).double()
        B = 3
        prompts = _prompt_ids(ck["vocab"], 50 + seed, n=B)
        g = torch.Generator().manual_seed(77 + seed)
        forced = torch.randint(0, geo.vocab_size, (B, n_steps), generator=g)
        key = f"{name}_s{seed}"
        for b, ids in enumerate(prompts):
            for tag, m in (("f32", model), ("f64", model64)):
                x = torch.tensor([ids])
                _, past = m(x)
                feed, rows = ids[-1], []
                for i in range(n_steps):
                    lg, past = m(torch.tensor([[feed]]), past)
                    rows.append(lg[0, -1].to(torch.float64 if tag == "f64" else torch.float32).numpy())
                    feed = int(forced[b, i])
                out[f"{key}_b{b}_logits_{tag}"] = np.stack(rows)
            out[f"{key}_b{b}_prompt"] = np.asarray(ids, np.int32)
        out[f"{key}_forced"] = forced.numpy().astype(np.int32)
        meta.append({"key": key, "geometry": name, "seed": seed, "n_steps": n_steps, "batch": B,
                     "digest": mg.state_dict_digest(ck["model"])})
        print(key, "logits ok")
    np.savez_compressed(os.path.join(OUT, "kv_logits.npz"), meta=json.dumps(meta), **out)


@torch.no_grad()
def kv_layers():
    """Per-layer cache tensors (= LN1(x), api_cache.py:60-70) of one prefill and one decode step."""
    ns = refload.load_kv_reference()
    geo = mg.GEOMETRIES["tiny"]
    ck = mg.make_checkpoint(geo, 0)
    model = refload.build_kv_model(ns, ck["model"], geo.n_head)
    ids = _prompt_ids(ck["vocab"], 5, n=1)[0]
    lg0, past = model(torch.tensor([ids]))
    lg1, past1 = model(torch.tensor([[ids[-1]]]), past)
    out = {"prompt": np.asarray(ids, np.int32), "prefill_logits": lg0[0].numpy(), "decode_logits": lg1[0, -1].numpy()}
    for i, (k, _) in enumerate(past1):
        out[f"cache_l{i}"] = k[0].numpy()
    np.savez_compressed(os.path.join(OUT, "kv_layers.npz"),
                        meta=json.dumps({"geometry": "tiny", "seed": 0, "digest": mg.state_dict_digest(ck["model"])}),
                        **out)
    print("kv_layers ok")


@torch.no_grad()
def nocache_greedy():
    """Model (A): greedy tokens + last-position logits from reference generate.py:25-35,46-61."""
    ns = refload.load_nocache_reference()
    out, meta = {}, []
    for name, seed, max_len in [("tiny", 0, 20), ("tiny_hd64", 0, 30), ("train_mini", 0, 48)]:
        geo = mg.GEOMETRIES[name]
        ck = mg.make_checkpoint(geo, seed)
        model = refload.build_nocache_model(ns, ck["model"], geo.n_head)
        ids = _prompt_ids(ck["vocab"], seed, n=1)[0]
        x = torch.tensor([ids])
        first_logits = model(x)[0, -1].numpy()
        for _ in range(max_len - len(ids)):          # control flow of generate.py:50-58 with top_k=1
            nxt = int(torch.argmax(model(x)[0, -1]))
            x = torch.cat([x, torch.tensor([[nxt]])], dim=1)
        key = f"{name}_s{seed}"
        out[key + "_prompt"] = np.asarray(ids, np.int32)
        out[key + "_tokens"] = x[0].numpy().astype(np.int32)
        out[key + "_first_logits"] = first_logits
        meta.append({"key": key, "geometry": name, "seed": seed, "max_len": max_len,
                     "digest": mg.state_dict_digest(ck["model"])})
        print(key, "nocache ok")
    np.savez_compressed(os.path.join(OUT, "nocache_greedy.npz"), meta=json.dumps(meta), **out)


@torch.no_grad()
def topk_distribution():
    """Top-k sampling distributions computed with the op sequence of api_cache.py:169-177."""
    g = torch.Generator().manual_seed(3)
    out = {}
    for name, V, k, temp in [("v8324_k40", 8324, 40, 1.0), ("v8324_k50_t08", 8324, 50, 0.8), ("v96_k5", 96, 5, 1.3),
                             ("v8324_full", 8324, None, 1.0)]:
        logits = (torch.randn(1, V, generator=g) * 2.0)
        z = logits / temp
        if k is not None:
            vals, idxs = z.topk(k)
            mask = torch.full_like(z, -1e10)
            mask.scatter_(1, idxs, 0.0)
            z = z + mask
        probs = torch.softmax(z, dim=-1)
        out[name + "_logits"] = logits[0].numpy()
        out[name + "_probs"] = probs[0].numpy()
        out[name + "_k"] = np.asarray(-1 if k is None else k)
        out[name + "_temp"] = np.asarray(temp, np.float32)
    np.savez_compressed(os.path.join(OUT, "topk_probs.npz"), **out)
    print("topk ok")


@torch.no_grad()
def distilbert():
    """Logits of the installed transformers DistilBertForSequenceClassification (third-party dep)."""
    from transformers import DistilBertConfig, DistilBertForSequenceClassification

    out, meta = {}, []
    for name, geo, N, T, seed in [("tiny", bc.TINY_BERT, 6, 16, 0), ("base", bc.DISTILBERT_BASE, 4, 24, 0)]:
        sd = bc.make_bert_state_dict(geo, seed)
        merged = bc.merge_lora_state_dict(sd)
        cfg = DistilBertConfig(vocab_size=geo.vocab_size, max_position_embeddings=geo.max_pos, dim=geo.dim,
                               n_heads=geo.n_heads, n_layers=geo.n_layers, hidden_dim=geo.hidden_dim,
                               num_labels=geo.num_labels, dropout=0.0, attention_dropout=0.0, seq_classif_dropout=0.0)
        model = DistilBertForSequenceClassification(cfg).eval()
        missing = model.load_state_dict(merged, strict=True)
        g = torch.Generator().manual_seed(11)
        ids = torch.randint(1, geo.vocab_size, (N, T), generator=g)
        mask = torch.ones(N, T, dtype=torch.long)
        mask[1, T - 5:] = 0                       # ragged rows exercise the padding mask
        mask[2, T // 2:] = 0
        logits = model(input_ids=ids, attention_mask=mask).logits
        out[name + "_ids"] = ids.numpy().astype(np.int32)
        out[name + "_mask"] = mask.numpy().astype(np.uint8)
        out[name + "_logits"] = logits.numpy()
        meta.append({"key": name, "seed": seed, "N": N, "T": T, "digest": mg.state_dict_digest(sd)})
        print(name, "distilbert ok", missing)
    np.savez_compressed(os.path.join(OUT, "distilbert.npz"), meta=json.dumps(meta), **out)


@torch.no_grad()
def distilbert_margin():
    """tests/golden/distilbert_margin.npz: 48 inputs whose top-2 logit margin under the installed transformers
    DistilBertForSequenceClassification (fp32, LoRA merged) is the largest of a pool of 384 random rows -- every one of them far
    above the bf16 engine's logit error, so the parity test can demand label identity on ALL rows."""
    from transformers import DistilBertConfig, DistilBertForSequenceClassification

    geo, seed, N, T, keep = bc.DISTILBERT_BASE, 0, 384, 32, 48
    sd = bc.make_bert_state_dict(geo, seed)
    cfg = DistilBertConfig(vocab_size=geo.vocab_size, max_position_embeddings=geo.max_pos, dim=geo.dim,
                           n_heads=geo.n_heads, n_layers=geo.n_layers, hidden_dim=geo.hidden_dim,
                           num_labels=geo.num_labels, dropout=0.0, attention_dropout=0.0, seq_classif_dropout=0.0)
    model = DistilBertForSequenceClassification(cfg).eval()
    model.load_state_dict(bc.merge_lora_state_dict(sd), strict=True)
    g = torch.Generator().manual_seed(23)
    ids = torch.randint(1000, 30000, (N, T), generator=g)
    ids[:, 0], ids[:, T - 1] = 101, 102
    logits = torch.cat([model(input_ids=ids[i:i + 64]).logits for i in range(0, N, 64)])
    top2 = torch.topk(logits, 2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    sel = torch.argsort(margin, descending=True)[:keep]
    meta = {"seed": seed, "N": keep, "T": T, "pool": N, "digest": mg.state_dict_digest(sd), "min_margin": float(margin[sel].min())}
    np.savez_compressed(os.path.join(OUT, "distilbert_margin.npz"), meta=json.dumps(meta), ids=ids[sel].numpy().astype(np.int32),
                        logits=logits[sel].numpy(), margin=margin[sel].numpy())
    print("distilbert margin fixture ok", meta)


def eats_table():
    """tests/golden/eats_table.json: the emotion -> music-parameter table the REFERENCE builds at import
    (emotion_analysis/EATS.py:7-19 from its lookup_table.csv) and the outputs of its get_music_params for all 28 labels under
    random.seed(0) -- the pin for pipeline.eats_music_params (BASELINE config 5 uses the reference's own mapping)."""
    import importlib.util
    import random
    path = os.path.join(refload.REFERENCE_ROOT, "emotion_analysis", "EATS.py")
    spec = importlib.util.spec_from_file_location("ref_eats", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                              # reads lookup_table.csv next to it (pandas)
    random.seed(0)
    labels = list(bc.ID2LABEL.values())
    calls = [mod.get_music_params(lab) for lab in labels]
    with open(os.path.join(OUT, "eats_table.json"), "w") as f:
        json.dump({"table": mod.EATS, "labels": labels, "seed": 0, "get_music_params": calls}, f, indent=1, sort_keys=True)
    print("eats table ok", len(mod.EATS))


if __name__ == "__main__":
    if not refload.reference_available():
        sys.exit("reference tree not found; golden fixtures can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    kv_greedy()
    kv_logits()
    kv_layers()
    nocache_greedy()
    topk_distribution()
    distilbert()
    distilbert_margin()
    eats_table()
