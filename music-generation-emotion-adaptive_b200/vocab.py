"""Synthetic vocabulary and the prompt-token contract of the reference service.

The real vocabularies are data dependent and not shipped.  The synthetic
one satisfies everything the unchanged prompt builder looks up (reference api_cache.py:142-156,
194-203): ``[START_SEQUENCE]``, ``[END_SEQUENCE]``, ``[PAD]``, ``[BPM] 20.0`` .. ``[BPM] 250.0``,
the 24 ``[KEY_SIGNATURE]`` names of reference dataparsing/analysis_output.txt:2-25, the three
``[INSTRUMENT]`` names of api_cache.py:152-156, then train_large.py-style atomic note tokens
(train/train_large.py:36-55); ids are assigned by ``sorted()`` like train/train_large.py:77.
"""
from __future__ import annotations

import re
from typing import Dict, List, Sequence

KEY_NAMES = [
    "C major", "G major", "F major", "D major", "A minor", "A major", "D minor", "B- major",
    "E minor", "E- major", "E major", "G minor", "C minor", "A- major", "B minor", "F minor",
    "C# major", "B major", "F# minor", "F# major", "E- minor", "B- minor", "C# minor", "G# minor",
]
INSTRUMENT_NAMES = ["Violin", "Acoustic Grand Piano", "Flute"]
# reference api_cache.py:152-156
FAMILY_TO_INSTRUMENTS = {"Strings": ["Violin"], "Piano": ["Acoustic Grand Piano"], "Woodwind": ["Flute"]}
EOS_TOKEN = "[END_SEQUENCE]"

_PITCH_LETTERS = ["C", "C#", "D", "E-", "E", "F", "F#", "G", "G#", "A", "B-", "B"]


def _fixed_tokens() -> List[str]:
    toks = ["[START_SEQUENCE]", EOS_TOKEN, "[PAD]"]
    toks += [f"[BPM] {float(b)}" for b in range(20, 251)]
    toks += [f"[KEY_SIGNATURE] {k}" for k in KEY_NAMES]
    toks += [f"[INSTRUMENT] {n}" for n in INSTRUMENT_NAMES]
    toks += ["[NOTE]", "[PITCH]", "[START_T]", "[END_T]", "[DUR_T]"]
    toks += [f"{p}{o}" for o in range(0, 9) for p in _PITCH_LETTERS]   # 108 pitch names
    return toks


def build_synthetic_vocab(vocab_size: int = 8324) -> Dict[str, int]:
    """token string -> id, exactly ``vocab_size`` entries, ids by sorted() order."""
    toks = _fixed_tokens()
    if vocab_size < 8:
        raise ValueError("vocab_size too small")
    if vocab_size < len(toks):
        # tiny test vocabularies: keep the specials that the sampler contract needs first
        toks = toks[:vocab_size]
    else:
        tick = 0
        have = set(toks)
        while len(toks) < vocab_size:     # tick integers, as strings (train_large.py:49-53)
            s = str(tick)
            if s not in have:
                toks.append(s)
                have.add(s)
            tick += 1
    assert len(set(toks)) == vocab_size
    return {t: i for i, t in enumerate(sorted(toks))}


def closest_bpm_token(tok2id: Dict[str, int], val: float) -> str:
    """Nearest ``[BPM] x`` vocabulary entry (behaviour of reference api_cache.py:142-144)."""
    cands = [t for t in tok2id if t.startswith("[BPM]")]
    if not cands:
        raise ValueError("vocabulary has no [BPM] tokens")
    return min(cands, key=lambda s: abs(float(s.split()[-1]) - val))


def normalize_key_signature(key_string: str) -> str:
    """``"D Major"`` -> ``"[KEY_SIGNATURE] D major"`` (behaviour of reference api_cache.py:145-151)."""
    s = key_string.replace("♭", "-").replace("♯", "#")
    parts = s.strip().split()
    if len(parts) == 2:
        return f"[KEY_SIGNATURE] {parts[0]} {parts[1].lower()}"
    return f"[KEY_SIGNATURE] {s}"


def build_prompt(tok2id: Dict[str, int], bpm: float, key: str, families: Sequence[str]) -> List[str]:
    """Prompt token strings in the service's order (reference api_cache.py:194-203)."""
    instruments: List[str] = []
    for fam in families:
        instruments.extend(FAMILY_TO_INSTRUMENTS.get(fam, []))
    return (["[START_SEQUENCE]", closest_bpm_token(tok2id, bpm), normalize_key_signature(key)]
            + [f"[INSTRUMENT] {i}" for i in instruments])


def encode(tok2id: Dict[str, int], tokens: Sequence[str]) -> List[int]:
    """Raises KeyError for an out-of-vocabulary token, like reference api_cache.py:162."""
    return [tok2id[t] for t in tokens]


def synthetic_prompts(tok2id: Dict[str, int], n: int, seed: int = 0) -> List[List[str]]:
    """``n`` production-shaped prompts (3..6 tokens) drawn deterministically."""
    import random

    rng = random.Random(seed)
    fams = list(FAMILY_TO_INSTRUMENTS)
    out = []
    for _ in range(n):
        k = rng.randint(0, 3)
        chosen = rng.sample(fams, k)
        # lookup_table.csv writes keys as "D Major"; title() reproduces that spelling
        out.append(build_prompt(tok2id, rng.randint(20, 250), rng.choice(KEY_NAMES).title(), chosen))
    return out


# ---- device-side detokenisation table (reference api_cache.py:157,208-221) ------------------------------------------
NOTE_RE = re.compile(r"\[NOTE\] \[PITCH:(.+?)\] \[START:(.+?)\] \[END:(.+?)\] \[DURATION:(.+?)\]")   # api_cache.py:157
_NOTE_NAME_RE = re.compile(r"^(?P<n>[A-Ga-g])(?P<off>[#b!]?)(?P<oct>[+-]?\d+)$")
_PITCH_CLASS = {"C": 0, "D": 2, "E": 4, "F": 5, "G": 7, "A": 9, "B": 11}
# General MIDI programs of the instruments the service can ask for (api_cache.py:152-156); any other name goes through the
# caller's ``instrument_program`` (pretty_midi.instrument_name_to_program on the service box) or falls to 0 like api_cache.py:212
GM_PROGRAMS = {"Acoustic Grand Piano": 0, "Violin": 40, "Flute": 73}


def note_name_to_number(name: str) -> int:
    """pretty_midi.note_name_to_number (third-party, pinned by the reference's requirements.txt): letter, optional # / b / !,
    signed octave; ``12 * (octave + 1) + pitch class + accidental``; ValueError for anything else."""
    m = _NOTE_NAME_RE.match(name)
    if not m:
        raise ValueError(f"Improper note format: {name}")
    return 12 * (int(m.group("oct")) + 1) + _PITCH_CLASS[m.group("n").upper()] + {"#": 1, "": 0, "b": -1, "!": -1}[m.group("off")]


def note_table(tok2id: Dict[str, int], vocab_size: int, instrument_program=None, note_number=None):
    """One record per token id for ``mg_set_note_table``: (kind, value, start, end) as int32 / int32 / float32 / float32 arrays.

    Everything the reference's loop does per generated token -- ``startswith("[INSTRUMENT]")``, the name -> program look-up
    (0 when the name is unknown), ``note_re.match``, ``note_name_to_number``, two ``float()`` -- is done here once per vocabulary
    entry.  A NOTE token whose pitch or times do not parse would raise inside the reference's loop (HTTP 500 when it is ever
    generated): the table builder raises the same ValueError up front."""
    import numpy as np
    prog = instrument_program or (lambda name: GM_PROGRAMS.get(name, 0))
    num = note_number or note_name_to_number
    kind = np.zeros(vocab_size, np.int32)
    value = np.zeros(vocab_size, np.int32)
    start = np.zeros(vocab_size, np.float32)
    end = np.zeros(vocab_size, np.float32)
    for tok, i in tok2id.items():
        if not 0 <= i < vocab_size:
            continue
        if tok.startswith("[INSTRUMENT]"):                              # api_cache.py:209-214
            kind[i], value[i] = 1, int(prog(tok.split("]", 1)[1].strip()))
        else:
            m = NOTE_RE.match(tok)                                      # api_cache.py:215-220
            if m:
                kind[i], value[i] = 2, int(num(m.group(1)))
                start[i], end[i] = float(m.group(2)), float(m.group(3))
    return kind, value, start, end
