"""The request path of the reference service with both hot calls on the engine:
classify(text) -> emotion label -> music parameters -> prompt tokens -> generate   (reference api_cache.py:187-204).

The emotion -> music-parameter table (reference emotion_analysis/EATS.py + lookup_table.csv) is host-side glue that stays
in the reference; pass its ``get_music_params`` as ``params_fn``.  ``synthetic_music_params`` is a deterministic stand-in
with the same return shape for benchmarks on boxes where the reference tree does not exist.
"""
from __future__ import annotations

import random
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from .bert_checkpoint import ID2LABEL
from .vocab import FAMILY_TO_INSTRUMENTS, KEY_NAMES, build_prompt, encode


def synthetic_music_params(label: str, rng: Optional[random.Random] = None) -> Dict:
    """Same keys as reference EATS.get_music_params (EATS.py:21-42): emotion, bpm, key, scale_type, inst_family, all_families."""
    rng = rng or random.Random(0)
    h = sum(ord(c) for c in label)
    lo = 60 + (h % 6) * 20
    key = KEY_NAMES[h % len(KEY_NAMES)]
    fams = list(FAMILY_TO_INSTRUMENTS)
    k = 1 + h % len(fams)
    all_fams = [fams[(h + i) % len(fams)] for i in range(k)]
    return {"emotion": label, "bpm": rng.randint(lo, lo + 20), "key": key.title(), "scale_type": key.split()[1],
            "inst_family": rng.choice(all_fams), "all_families": all_fams}


def load_eats_table(path: str) -> Dict[str, Dict]:
    """The reference's emotion table, from its own ``lookup_table.csv`` (emotion_analysis/EATS.py:7-19: emotion, bpm_min,
    bpm_max, key, scale_type, JSON list of instrument families) or from the JSON dump of it (tests/golden/eats_table.json)."""
    import csv
    import json
    if path.endswith(".json"):
        with open(path) as f:
            j = json.load(f)
        return j.get("table", j)
    with open(path, newline="") as f:
        return {r["emotion"]: {"bpm_min": int(r["bpm_min"]), "bpm_max": int(r["bpm_max"]), "key": r["key"],
                               "scale_type": r["scale_type"], "instrument_families": json.loads(r["instrument_families"])}
                for r in csv.DictReader(f)}


def eats_music_params(table: Dict[str, Dict]) -> Callable[[str], Dict]:
    """``get_music_params`` of the reference for one label (emotion_analysis/EATS.py:21-42) over a loaded table: same keys, same
    ValueError for an unknown emotion, and the same two draws from the module-level ``random`` in the same order (bpm first,
    family second), so ``random.seed(0)`` reproduces the reference's choices (pinned by tests/golden/eats_table.json)."""
    def params(label: str) -> Dict:
        lc = label.lower()
        if lc not in table:
            raise ValueError(f"Emotion '{label}' not in lookup table")
        e = table[lc]
        bpm = random.randint(e["bpm_min"], e["bpm_max"])
        fam = random.choice(e["instrument_families"])
        return {"emotion": lc, "bpm": bpm, "key": e["key"], "scale_type": e["scale_type"], "inst_family": fam,
                "all_families": e["instrument_families"]}
    return params


def classify_prompt_generate(clf, gen, tok2id: Dict[str, int], input_ids, attention_mask=None,
                             params_fn: Callable[[str], Dict] = synthetic_music_params, max_len: int = 512,
                             temperature: float = 1.0, top_k: Optional[int] = 50, seed: Optional[int] = None, batch: int = 128,
                             seq_index_base: int = 0) -> List[List[int]]:
    """N requests -> N token-id lists (prompt included), exactly the per-request flow of api_cache.py:187-204.
    ``seed=None``: one fresh seed for the whole call (requests differ by their Philox sequence index)."""
    if seed is None:
        from .engine import fresh_seed
        seed = fresh_seed()
    ids = np.asarray(input_ids)
    labels: List[int] = []
    step = max(1, clf_capacity(clf) // ids.shape[1])
    for lo in range(0, ids.shape[0], step):
        m = None if attention_mask is None else np.asarray(attention_mask)[lo:lo + step]
        lab, _ = clf.classify(ids[lo:lo + step], m)
        labels.extend(int(x) for x in lab)
    prompts = []
    for lab in labels:
        prm = params_fn(ID2LABEL[lab])                                  # api_cache.py:190
        fams = prm.get("all_families") or [prm["inst_family"]]
        prompts.append(encode(tok2id, build_prompt(tok2id, prm["bpm"], prm["key"], fams)))   # :194-203, KeyError on OOV (:162)
    eos = tok2id.get("[END_SEQUENCE]", -1)
    out: List[List[int]] = []
    for lo in range(0, len(prompts), batch):
        chunk = prompts[lo:lo + batch]
        out.extend(gen.generate(chunk, [max(0, max_len - len(p)) for p in chunk], temperature, top_k, eos_id=eos, seed=seed,
                                seq_index_base=seq_index_base + lo))
    return out


def clf_capacity(clf) -> int:
    """Token budget (N * T) of one ``classify`` call: the ``max_tokens`` the classifier's workspace was sized for."""
    return int(getattr(clf, "max_tokens", 16384))
