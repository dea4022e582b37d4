"""ORACLE SUPPORT (in-container only): execute the reference's OWN hot-path classes.

/root/reference exists only in the build container (never on the GPU box), so nothing in the
``-m gpu`` tests, smoke() or bench.py may call this at run time.  It is used by
oracle/make_golden.py to generate tests/golden/ fixtures and by ``-m "not gpu"`` tests that skip
when the reference tree is absent.

Nothing is copied: the class bodies are compiled straight from the read-only reference files by
selecting AST nodes (recipe proven in SURVEY.md section 8c).
"""
from __future__ import annotations

import ast
import os
import re
from typing import Dict, Iterable

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("MG_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "api_cache.py"))


def _exec_nodes(path: str, wanted: Iterable[str], ns: Dict) -> Dict:
    src = open(path).read()
    wanted = set(wanted)
    body = [n for n in ast.parse(src).body if getattr(n, "name", None) in wanted]
    missing = wanted - {n.name for n in body}
    if missing:
        raise RuntimeError(f"reference nodes not found in {path}: {sorted(missing)}")
    exec(compile(ast.Module(body=body, type_ignores=[]), os.path.basename(path) + ":slice", "exec"), ns)
    return ns


def load_kv_reference() -> Dict:
    """GPTBlock / GPTWithKV / remap_state_dict / sample_kvcache of reference api_cache.py:39-184."""
    ns = {"torch": torch, "nn": nn, "re": re}
    return _exec_nodes(os.path.join(REFERENCE_ROOT, "api_cache.py"),
                       {"GPTBlock", "GPTWithKV", "remap_state_dict", "sample_kvcache"}, ns)


def load_nocache_reference() -> Dict:
    """GPT of reference generate_music/generate.py:25-35 (its ``sample`` closes over module globals)."""
    ns = {"torch": torch, "nn": nn}
    return _exec_nodes(os.path.join(REFERENCE_ROOT, "generate_music", "generate.py"), {"GPT", "sample"}, ns)


def build_kv_model(ns: Dict, trainer_sd: Dict[str, torch.Tensor], n_head: int):
    pos = trainer_sd["pos"]
    n_layer = 1 + max(int(k.split(".")[2]) for k in trainer_sd if "tr.layers." in k)
    model = ns["GPTWithKV"](trainer_sd["emb.weight"].shape[0], pos.shape[0], pos.shape[1], n_head, n_layer)
    model.load_state_dict(ns["remap_state_dict"](trainer_sd))
    return model.eval()


def build_nocache_model(ns: Dict, trainer_sd: Dict[str, torch.Tensor], n_head: int):
    pos = trainer_sd["pos"]
    n_layer = 1 + max(int(k.split(".")[2]) for k in trainer_sd if "tr.layers." in k)
    # generate.py:28 allocates pos with seq_len-1 rows
    model = ns["GPT"](trainer_sd["emb.weight"].shape[0], pos.shape[0] + 1, pos.shape[1], n_head=n_head, n_layer=n_layer)
    model.load_state_dict(trainer_sd, strict=True)
    return model.eval()
