"""ctypes binding of the C ABI in include/mg_engine.h and the host-side mirrors of the reference calls.

* ``Generator``       engine-level form of reference api_cache.py:108-138 (model build + weight load)
                      and :159-184 (``sample_kvcache``): ``generate(prompt_tokens, max_new_tokens,
                      temperature, top_k)``.
* ``sample_kvcache``  drop-in with the reference's own signature (api_cache.py:160).
* ``sample``          drop-in for the no-cache twin (generate_music/generate.py:46).
* ``Classifier``      engine-level form of reference emotion_analysis/modeling.py:8-25 +
                      inference.py:12-22: ``classify(texts-as-ids)`` / ``predict``.

There is NO CPU fallback: every constructor raises if ``libmgea_b200.so`` is missing or no B200 is
visible.  Errors of the C ABI are re-raised as the exception types the reference raises at the same
place (KeyError for OOV tokens comes from ``vocab.encode``; RuntimeError for an over-long prompt or
``top_k`` > vocab; ValueError for bad arguments).
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .bert_checkpoint import ID2LABEL, BertGeometry, expected_bert_keys, infer_bert_geometry, merge_lora_state_dict
from .checkpoint import Geometry, expected_keys, infer_geometry, remap_state_dict

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmgea_b200.so")

MG_OK = 0
MG_E_SHAPE, MG_E_PROMPT_TOO_LONG, MG_E_TOPK, MG_E_CUDA, MG_E_OOM, MG_E_STATE, MG_E_ARG, MG_E_TOKEN = range(-1, -9, -1)
MG_DTYPE_FP32, MG_DTYPE_BF16 = 0, 1
_DTYPES = {"fp32": MG_DTYPE_FP32, "float32": MG_DTYPE_FP32, "bf16": MG_DTYPE_BF16, "bfloat16": MG_DTYPE_BF16}

# every symbol include/mg_engine.h declares (tests/test_host_cpu.py checks the list against the header)
EXPORTED_SYMBOLS = [
    "mg_abi_version", "mg_last_error", "mg_device_count", "mg_engine_create", "mg_engine_destroy", "mg_load_weight",
    "mg_engine_finalize", "mg_generate", "mg_upload_prompts", "mg_run", "mg_download", "mg_synchronize",
    "mg_engine_stream", "mg_step_logits", "mg_step_logits_at", "mg_generate_nocache", "mg_forward_nocache", "mg_sample_logits",
    "mg_engine_stats", "mg_slots_begin", "mg_slots_admit", "mg_slots_step", "mg_slots_fetch", "mg_slots_fetch_many", "mg_slots_end", "mg_set_note_table", "mg_note_events", "mg_last_run_timing", "mg_last_step_times", "mg_last_decode_path", "mg_bert_create", "mg_bert_destroy", "mg_bert_load_weight",
    "mg_bert_finalize", "mg_classify", "mg_bert_upload", "mg_bert_run", "mg_bert_download", "mg_bert_synchronize",
    "mg_bert_stream", "mg_bert_stats", "mg_test_gemm_bf16", "mg_test_grid_plan",
]


class _Geometry(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("vocab_size", "pos_rows", "d_model", "n_head", "n_layer", "d_ff")]


class _BertGeometry(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("vocab_size", "max_pos", "dim", "n_heads", "n_layers", "hidden_dim",
                                              "num_labels")]


_lib = None


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    """Load the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.isfile(p):
        raise RuntimeError(f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           f"(or `make -C {os.path.join(_HERE, 'csrc')}`); there is no CPU fallback")
    lib = ctypes.CDLL(p)
    c = ctypes
    i32p, f32p, u8p, i64p = c.POINTER(c.c_int32), c.POINTER(c.c_float), c.POINTER(c.c_uint8), c.POINTER(c.c_int64)
    vp, u64, u64p = c.c_void_p, c.c_uint64, c.POINTER(c.c_uint64)
    sig = {
        "mg_abi_version": (c.c_int, []),
        "mg_last_error": (c.c_char_p, []),
        "mg_device_count": (c.c_int, []),
        "mg_engine_create": (c.c_int, [c.POINTER(_Geometry), c.c_int, c.c_int, c.c_int, c.c_int, c.POINTER(vp)]),
        "mg_engine_destroy": (None, [vp]),
        "mg_load_weight": (c.c_int, [vp, c.c_char_p, f32p, i64p, c.c_int]),
        "mg_engine_finalize": (c.c_int, [vp]),
        "mg_generate": (c.c_int, [vp, i32p, i32p, c.c_int, c.c_int, i32p, c.c_float, c.c_int, c.c_int, u64, u64, i32p,
                                  c.c_int, i32p]),
        "mg_upload_prompts": (c.c_int, [vp, i32p, i32p, c.c_int, c.c_int, i32p]),
        "mg_run": (c.c_int, [vp, c.c_float, c.c_int, c.c_int, u64, u64]),
        "mg_download": (c.c_int, [vp, i32p, c.c_int, i32p]),
        "mg_synchronize": (c.c_int, [vp]),
        "mg_engine_stream": (vp, [vp]),
        "mg_step_logits": (c.c_int, [vp, i32p, i32p, c.c_int, i32p, c.c_int, f32p]),
        "mg_step_logits_at": (c.c_int, [vp, i32p, i32p, c.c_int, i32p, c.c_int, i32p, c.c_int, f32p]),
        "mg_generate_nocache": (c.c_int, [vp, i32p, i32p, c.c_int, c.c_int, c.c_float, c.c_int, c.c_int, u64, u64,
                                          i32p, c.c_int, i32p]),
        "mg_forward_nocache": (c.c_int, [vp, i32p, i32p, c.c_int, f32p]),
        "mg_sample_logits": (c.c_int, [vp, f32p, c.c_int, c.c_int, c.c_float, c.c_int, u64, u64, c.c_uint32, i32p]),
        "mg_engine_stats": (c.c_int, [vp, u64p, u64p, u64p]),
        "mg_last_run_timing": (c.c_int, [vp, f32p, f32p, f32p, c.POINTER(c.c_int)]),
        "mg_last_step_times": (c.c_int, [vp, f32p, c.c_int, c.POINTER(c.c_int)]),
        "mg_slots_begin": (c.c_int, [vp, c.c_int, c.c_int, c.c_float, c.c_int, c.c_int, u64]),
        "mg_slots_admit": (c.c_int, [vp, c.c_int, i32p, i32p, i32p, i32p, i32p]),
        "mg_slots_step": (c.c_int, [vp, c.c_int, u8p, i32p]),
        "mg_slots_fetch": (c.c_int, [vp, c.c_int, i32p, c.c_int, c.POINTER(c.c_int)]),
        "mg_slots_fetch_many": (c.c_int, [vp, c.c_int, i32p, i32p, c.c_int, i32p]),
        "mg_slots_end": (c.c_int, [vp]),
        "mg_set_note_table": (c.c_int, [vp, i32p, i32p, f32p, f32p, c.c_int]),
        "mg_note_events": (c.c_int, [vp, c.c_int, c.c_int, i32p, i32p, i32p, i32p, i32p, i32p, f32p, f32p]),
        "mg_last_decode_path": (c.c_int, [vp]),
        "mg_bert_create": (c.c_int, [c.POINTER(_BertGeometry), c.c_int, c.c_int, c.POINTER(vp)]),
        "mg_bert_destroy": (None, [vp]),
        "mg_bert_load_weight": (c.c_int, [vp, c.c_char_p, f32p, i64p, c.c_int]),
        "mg_bert_finalize": (c.c_int, [vp]),
        "mg_classify": (c.c_int, [vp, i32p, u8p, c.c_int, c.c_int, f32p, i32p]),
        "mg_bert_upload": (c.c_int, [vp, i32p, u8p, c.c_int, c.c_int]),
        "mg_bert_run": (c.c_int, [vp]),
        "mg_bert_download": (c.c_int, [vp, f32p, i32p]),
        "mg_bert_synchronize": (c.c_int, [vp]),
        "mg_bert_stream": (vp, [vp]),
        "mg_bert_stats": (c.c_int, [vp, u64p, u64p, u64p]),
        "mg_test_gemm_bf16": (c.c_int, [c.c_int, f32p, f32p, f32p, c.c_int, c.c_int, c.c_int, c.c_int, f32p]),
        "mg_test_grid_plan": (c.c_int, [c.c_int] * 6 + [i32p, c.POINTER(c.c_int16), i32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if path is None:
        _lib = lib
    return lib


def _check(lib, rc: int) -> None:
    """Map C status codes to the exception types the reference raises at the same point."""
    if rc == MG_OK:
        return
    msg = (lib.mg_last_error() or b"").decode(errors="replace")
    if rc in (MG_E_ARG, MG_E_TOKEN):
        raise ValueError(msg)                      # reference: bad call arguments -> ValueError / IndexError
    if rc == MG_E_OOM:
        raise MemoryError(msg)
    raise RuntimeError(f"[mg status {rc}] {msg}")  # reference: torch RuntimeError (shape / top-k / device)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a: Optional[np.ndarray], ctype):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctype))


def fresh_seed() -> int:
    """A new 64-bit Philox key per call.  The reference draws from torch's global RNG (api_cache.py:178), so two
    requests with the same prompt give two different pieces; a fixed default seed would repeat the same one."""
    return int.from_bytes(os.urandom(8), "little")


def _seed(seed: Optional[int]) -> int:
    return fresh_seed() if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF


def _pack_prompts(prompts: Sequence[Sequence[int]]) -> Tuple[np.ndarray, np.ndarray]:
    if len(prompts) == 0:
        raise ValueError("empty batch")
    lens = [len(p) for p in prompts]
    offs = np.zeros(len(prompts) + 1, np.int32)
    offs[1:] = np.cumsum(lens)
    flat = np.fromiter((t for p in prompts for t in p), dtype=np.int64, count=int(offs[-1]))
    if flat.size and (flat.min() < -2**31 or flat.max() >= 2**31):
        raise ValueError("token id does not fit int32")
    return flat.astype(np.int32), offs


class Generator:
    """One MIDI-token generator replica on one GPU (engine-level mirror of api_cache.py:108-184)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], n_head: int = 8, dtype: str = "bf16", max_batch: int = 64,
                 max_seq: int = 1088, device: int = 0, already_remapped: bool = False):
        self.lib = load_library()
        self._h = ctypes.c_void_p()
        if dtype not in _DTYPES:
            raise ValueError(f"dtype must be one of {sorted(_DTYPES)}")
        if already_remapped:
            sd = dict(state_dict)
            pos = sd["pos_emb"]
            n_layer = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("layers."))
            self.geometry = Geometry(int(sd["tok_emb.weight"].shape[0]), int(pos.shape[0]), int(pos.shape[1]),
                                     int(n_head), n_layer)
        else:
            self.geometry = infer_geometry(state_dict, n_head)        # api_cache.py:31-37 (+ :112 n_head)
            sd = remap_state_dict(state_dict)                          # api_cache.py:118-134
        g = self.geometry
        self.dtype, self.max_batch, self.max_seq, self.device = dtype, int(max_batch), int(max_seq), int(device)
        cg = _Geometry(g.vocab_size, g.pos_rows, g.d_model, g.n_head, g.n_layer, g.d_ff)
        _check(self.lib, self.lib.mg_engine_create(ctypes.byref(cg), device, _DTYPES[dtype], max_batch, max_seq,
                                                   ctypes.byref(self._h)))
        want = expected_keys(g)
        missing = sorted(set(want) - set(sd))
        if missing:
            raise RuntimeError(f"Error(s) in loading state_dict: missing keys {missing[:4]}...")   # load_state_dict raises
        for name, shape in want.items():
            t = sd[name].detach().to(torch.float32).contiguous().cpu()
            if tuple(t.shape) != tuple(shape):
                raise RuntimeError(f"size mismatch for {name}: checkpoint {tuple(t.shape)} vs model {tuple(shape)}")
            arr = t.numpy()
            shp = (ctypes.c_int64 * arr.ndim)(*arr.shape)
            _check(self.lib, self.lib.mg_load_weight(self._h, name.encode(), _ptr(arr, ctypes.c_float), shp, arr.ndim))
        _check(self.lib, self.lib.mg_engine_finalize(self._h))

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.mg_engine_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def from_checkpoint(cls, ckpt: dict, n_head: int = 8, **kw) -> "Generator":
        """``ckpt`` = the trainers' ``{"model": state_dict, "vocab": tok2id}`` (train_large.py:159-164)."""
        return cls(ckpt["model"], n_head=n_head, **kw)

    # -- the decode path --------------------------------------------------------------------------
    def generate(self, prompt_tokens: Sequence[Sequence[int]], max_new_tokens, temperature: float = 1.0,
                 top_k: Optional[int] = 50, eos_id: int = -1, seed: Optional[int] = None,
                 seq_index_base: int = 0, as_arrays: bool = False) -> List[List[int]]:
        """Row b of the result == a batch-1 reference ``sample_kvcache`` run on prompt b (ids incl. prompt).

        ``max_new_tokens`` is an int or one int per sequence (negative = 0: the reference's ``range(max_len - Tp)`` is
        empty and the prompt comes back unchanged); ``top_k=None`` disables the top-k mask; ``seed=None`` draws a fresh
        64-bit seed per call (the reference samples from torch's global RNG), an explicit seed makes the call reproducible.
        ``as_arrays=True`` returns int32 numpy rows (views of the one host buffer the engine filled) instead of lists.
        """
        flat, offs = _pack_prompts(prompt_tokens)
        B = len(prompt_tokens)
        per = None
        if isinstance(max_new_tokens, (list, tuple, np.ndarray)):
            per = np.maximum(_i32(max_new_tokens), 0)
            if per.shape != (B,):
                raise ValueError("max_new_tokens must have one entry per sequence")
            mx = int(per.max())
        else:
            mx = max(int(max_new_tokens), 0)
        if mx == 0:                                               # api_cache.py:166: empty range, prompt unchanged
            for p in prompt_tokens:
                if len(p) > self.geometry.pos_rows:               # the prefill at :163 still runs (and raises) first
                    raise RuntimeError(f"[mg status {MG_E_PROMPT_TOO_LONG}] prompt of {len(p)} tokens exceeds the "
                                       f"{self.geometry.pos_rows}-row position table")
            return [list(map(int, p)) for p in prompt_tokens]
        stride = int(max(len(p) for p in prompt_tokens)) + max(mx, 0)
        out = np.zeros((B, max(stride, 1)), np.int32)
        lens = np.zeros(B, np.int32)
        self._last_B = B
        _check(self.lib, self.lib.mg_generate(self._h, _ptr(flat, ctypes.c_int32), _ptr(offs, ctypes.c_int32), B, mx,
                                              _ptr(per, ctypes.c_int32), float(temperature), 0 if top_k is None else int(top_k),
                                              int(eos_id), _seed(seed), int(seq_index_base), _ptr(out, ctypes.c_int32),
                                              out.shape[1], _ptr(lens, ctypes.c_int32)))
        if as_arrays:
            return [out[b, :lens[b]] for b in range(B)]
        return [out[b, :lens[b]].tolist() for b in range(B)]

    def upload(self, prompt_tokens: Sequence[Sequence[int]], max_new_tokens: int) -> None:
        flat, offs = _pack_prompts(prompt_tokens)
        self._last_B = len(prompt_tokens)
        self._last_stride = int(max(len(p) for p in prompt_tokens)) + int(max_new_tokens)
        _check(self.lib, self.lib.mg_upload_prompts(self._h, _ptr(flat, ctypes.c_int32), _ptr(offs, ctypes.c_int32),
                                                    self._last_B, int(max_new_tokens), None))

    def run(self, temperature: float = 1.0, top_k: Optional[int] = 50, eos_id: int = -1, seed: Optional[int] = None,
            seq_index_base: int = 0) -> None:
        _check(self.lib, self.lib.mg_run(self._h, float(temperature), 0 if top_k is None else int(top_k), int(eos_id),
                                         _seed(seed), int(seq_index_base)))

    def synchronize(self) -> None:
        _check(self.lib, self.lib.mg_synchronize(self._h))

    def download(self) -> List[List[int]]:
        out = np.zeros((self._last_B, self._last_stride), np.int32)
        lens = np.zeros(self._last_B, np.int32)
        _check(self.lib, self.lib.mg_download(self._h, _ptr(out, ctypes.c_int32), out.shape[1], _ptr(lens, ctypes.c_int32)))
        return [out[b, :lens[b]].tolist() for b in range(self._last_B)]

    def step_logits(self, prompt_tokens: Sequence[Sequence[int]], forced_ids, n_steps: int) -> np.ndarray:
        """Teacher-forced logits [n_steps, B, V] along the reference loop (api_cache.py:87-106,167-168)."""
        flat, offs = _pack_prompts(prompt_tokens)
        B = len(prompt_tokens)
        forced = _i32(forced_ids) if forced_ids is not None else np.zeros((B, n_steps), np.int32)
        if forced.shape != (B, n_steps):
            raise ValueError("forced_ids must be [B, n_steps]")
        out = np.empty((n_steps, B, self.geometry.vocab_size), np.float32)
        _check(self.lib, self.lib.mg_step_logits(self._h, _ptr(flat, ctypes.c_int32), _ptr(offs, ctypes.c_int32), B,
                                                 _ptr(forced, ctypes.c_int32), int(n_steps), _ptr(out, ctypes.c_float)))
        return out

    def step_logits_at(self, prompt_tokens: Sequence[Sequence[int]], forced_ids, n_steps: int, want_steps: Sequence[int]) -> np.ndarray:
        """The same teacher-forced run, keeping only ``want_steps``: [len(want_steps), B, V].  For parity checks at the
        cache lengths of BASELINE configs 3 / 4, where all steps x batch x vocab would be gigabytes."""
        flat, offs = _pack_prompts(prompt_tokens)
        B = len(prompt_tokens)
        forced = _i32(forced_ids)
        if forced.shape != (B, n_steps):
            raise ValueError("forced_ids must be [B, n_steps]")
        want = _i32(want_steps)
        out = np.empty((len(want), B, self.geometry.vocab_size), np.float32)
        _check(self.lib, self.lib.mg_step_logits_at(self._h, _ptr(flat, ctypes.c_int32), _ptr(offs, ctypes.c_int32), B,
                                                    _ptr(forced, ctypes.c_int32), int(n_steps), _ptr(want, ctypes.c_int32),
                                                    len(want), _ptr(out, ctypes.c_float)))
        return out

    def sample_logits(self, logits: np.ndarray, temperature: float = 1.0, top_k: Optional[int] = 50, seed: int = 0,
                      seq_index_base: int = 0, step: int = 0) -> np.ndarray:
        """The fused sampler alone (api_cache.py:169-178) on caller-provided logits [rows, V]."""
        lg = np.ascontiguousarray(logits, dtype=np.float32)
        out = np.zeros(lg.shape[0], np.int32)
        _check(self.lib, self.lib.mg_sample_logits(self._h, _ptr(lg, ctypes.c_float), lg.shape[0], lg.shape[1],
                                                   float(temperature), 0 if top_k is None else int(top_k), int(seed),
                                                   int(seq_index_base), int(step), _ptr(out, ctypes.c_int32)))
        return out

    # -- recompute mode: model (A), generate_music/generate.py:25-61 ---------------------------------
    def generate_nocache(self, prompt_tokens: Sequence[Sequence[int]], max_new_tokens: int, temperature: float = 1.0,
                         top_k: Optional[int] = 50, eos_id: int = -1, seed: Optional[int] = None,
                         seq_index_base: int = 0):
        flat, offs = _pack_prompts(prompt_tokens)
        B = len(prompt_tokens)
        if int(max_new_tokens) <= 0:                               # generate.py:50: empty range
            return [list(map(int, p)) for p in prompt_tokens]
        stride = int(max(len(p) for p in prompt_tokens)) + max(int(max_new_tokens), 0)
        out = np.zeros((B, max(stride, 1)), np.int32)
        lens = np.zeros(B, np.int32)
        _check(self.lib, self.lib.mg_generate_nocache(self._h, _ptr(flat, ctypes.c_int32), _ptr(offs, ctypes.c_int32), B,
                                                      int(max_new_tokens), float(temperature),
                                                      0 if not top_k else int(top_k),      # generate.py:54 `if top_k:`
                                                      int(eos_id), _seed(seed), int(seq_index_base),
                                                      _ptr(out, ctypes.c_int32), out.shape[1], _ptr(lens, ctypes.c_int32)))
        return [out[b, :lens[b]].tolist() for b in range(B)]

    def forward_nocache(self, prompt_tokens: Sequence[Sequence[int]]) -> np.ndarray:
        """Last-position logits [B, V] of one full ``GPT.forward`` (generate.py:34-35)."""
        flat, offs = _pack_prompts(prompt_tokens)
        B = len(prompt_tokens)
        out = np.empty((B, self.geometry.vocab_size), np.float32)
        _check(self.lib, self.lib.mg_forward_nocache(self._h, _ptr(flat, ctypes.c_int32), _ptr(offs, ctypes.c_int32), B,
                                                     _ptr(out, ctypes.c_float)))
        return out

    # -- counters -----------------------------------------------------------------------------------
    def stats(self) -> Dict[str, int]:
        a, b, c = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        _check(self.lib, self.lib.mg_engine_stats(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"kernel_launches": a.value, "h2d_bytes": b.value, "d2h_bytes": c.value}

    DECODE_PATHS = {0: "step_graph", 1: "cluster_kernel", 2: "flow_kernel", 3: "grid_kernel"}

    def last_decode_path(self) -> str:
        """Which CUDA decode path served the last run: step_graph / cluster_kernel (decode_mega.cu) / flow_kernel (decode_flow.cu) /
        grid_kernel (decode_grid.cu)."""
        return self.DECODE_PATHS[int(self.lib.mg_last_decode_path(self._h))]

    # -- slot session: continuous batching (caller side: api_cache.py:186-204) --------------------
    def slots_begin(self, n_slots: int, max_len: int, temperature: float = 1.0, top_k: Optional[int] = 50, eos_id: int = -1,
                    seed: Optional[int] = None) -> None:
        self._n_slots = int(n_slots)
        self._slots_max_len = int(max_len)
        _check(self.lib, self.lib.mg_slots_begin(self._h, int(n_slots), int(max_len), float(temperature),
                                                 0 if top_k is None else int(top_k), int(eos_id), _seed(seed)))

    def slots_admit(self, slots: Sequence[int], prompt_tokens: Sequence[Sequence[int]], max_new: Sequence[int],
                    seq_index: Sequence[int]) -> None:
        flat, offs = _pack_prompts(prompt_tokens)
        sl, mn, si = _i32(slots), _i32(max_new), _i32(seq_index)
        if not (len(sl) == len(mn) == len(si) == len(prompt_tokens)):
            raise ValueError("slots, prompts, max_new and seq_index must have the same length")
        i32 = ctypes.c_int32
        _check(self.lib, self.lib.mg_slots_admit(self._h, len(sl), _ptr(sl, i32), _ptr(flat, i32), _ptr(offs, i32), _ptr(mn, i32),
                                                 _ptr(si, i32)))

    def slots_step(self, n_steps: int) -> Tuple[np.ndarray, np.ndarray]:
        """Up to ``n_steps`` decode steps for every slot in flight -> (finished[n_slots] bool, out_len[n_slots])."""
        fin, ln = np.zeros(self._n_slots, np.uint8), np.zeros(self._n_slots, np.int32)
        _check(self.lib, self.lib.mg_slots_step(self._h, int(n_steps), _ptr(fin, ctypes.c_uint8), _ptr(ln, ctypes.c_int32)))
        return fin.astype(bool), ln

    def slots_fetch(self, slot: int, cap: int = 8192) -> List[int]:
        buf, n = np.zeros(cap, np.int32), ctypes.c_int()
        _check(self.lib, self.lib.mg_slots_fetch(self._h, int(slot), _ptr(buf, ctypes.c_int32), cap, ctypes.byref(n)))
        return buf[:n.value].tolist()

    def slots_fetch_many(self, slots: Sequence[int], as_arrays: bool = False):
        """Token rows of several finished slots with one synchronisation."""
        sl = _i32(slots)
        if len(sl) == 0:
            return []
        stride = (int(self._slots_max_len) + 7) & ~7
        buf, lens = np.zeros((len(sl), stride), np.int32), np.zeros(len(sl), np.int32)
        _check(self.lib, self.lib.mg_slots_fetch_many(self._h, len(sl), _ptr(sl, ctypes.c_int32), _ptr(buf, ctypes.c_int32), stride,
                                                      _ptr(lens, ctypes.c_int32)))
        return [buf[j, :lens[j]] if as_arrays else buf[j, :lens[j]].tolist() for j in range(len(sl))]

    def slots_end(self) -> None:
        _check(self.lib, self.lib.mg_slots_end(self._h))

    # -- device-side detokenisation (api_cache.py:157,208-221) -------------------------------------
    def set_note_table(self, tok2id: Dict[str, int], instrument_program=None, note_number=None) -> None:
        """Upload the token-id -> (instrument | note) table built by ``vocab.note_table``; pass pretty_midi's
        ``instrument_name_to_program`` / ``note_name_to_number`` where it is installed (the service box)."""
        from .vocab import note_table
        kind, value, start, end = note_table(tok2id, self.geometry.vocab_size, instrument_program, note_number)
        _check(self.lib, self.lib.mg_set_note_table(self._h, _ptr(kind, ctypes.c_int32), _ptr(value, ctypes.c_int32),
                                                    _ptr(start, ctypes.c_float), _ptr(end, ctypes.c_float), len(kind)))
        self._id2tok = {i: t for t, i in tok2id.items()}

    def note_events(self, max_inst: int = 16, max_notes: int = 1088) -> List[List[Dict]]:
        """Note events of the last generation, assembled on the device from the token ids in HBM: per sequence the list of
        instruments in order of appearance, ``{"name", "program", "notes": [(pitch, start, end), ...]}`` -- what the loop of
        api_cache.py:208-221 feeds into pretty_midi (velocity is the constant 100 there)."""
        B = int(self._last_B)
        n_inst, n_notes = np.zeros(B, np.int32), np.zeros(B, np.int32)
        ip, it = np.zeros((B, max_inst), np.int32), np.zeros((B, max_inst), np.int32)
        ni, npit = np.zeros((B, max_notes), np.int32), np.zeros((B, max_notes), np.int32)
        ns, ne = np.zeros((B, max_notes), np.float32), np.zeros((B, max_notes), np.float32)
        i32, f32 = ctypes.c_int32, ctypes.c_float
        _check(self.lib, self.lib.mg_note_events(self._h, max_inst, max_notes, _ptr(n_inst, i32), _ptr(ip, i32), _ptr(it, i32),
                                                 _ptr(n_notes, i32), _ptr(ni, i32), _ptr(npit, i32), _ptr(ns, f32), _ptr(ne, f32)))
        if (n_inst > max_inst).any() or (n_notes > max_notes).any():
            raise ValueError("note_events: more instruments / notes than the given capacity")
        out = []
        for b in range(B):
            insts = [{"name": self._id2tok[int(it[b, i])].split("]", 1)[1].strip(), "program": int(ip[b, i]), "notes": []}
                     for i in range(int(n_inst[b]))]
            for j in range(int(n_notes[b])):
                insts[int(ni[b, j])]["notes"].append((int(npit[b, j]), float(ns[b, j]), float(ne[b, j])))
            out.append(insts)
        return out

    def last_step_times_us(self) -> np.ndarray:
        """Per-token latencies (microseconds between consecutive tokens of sequence 0) of the last run, device-stamped."""
        cap = 1 << 16
        buf = np.empty(cap, np.float32)
        n = ctypes.c_int()
        _check(self.lib, self.lib.mg_last_step_times(self._h, _ptr(buf, ctypes.c_float), cap, ctypes.byref(n)))
        return buf[:n.value].copy()

    def last_timing(self) -> Dict[str, float]:
        t, p, d, s = ctypes.c_float(), ctypes.c_float(), ctypes.c_float(), ctypes.c_int()
        _check(self.lib, self.lib.mg_last_run_timing(self._h, ctypes.byref(t), ctypes.byref(p), ctypes.byref(d), ctypes.byref(s)))
        return {"total_ms": t.value, "prefill_ms": p.value, "decode_ms": d.value, "steps": s.value}


class KVModel:
    """What the reference passes around as ``model`` plus the vocabulary its sampler reads from module
    globals (``tok2id`` / ``id2tok``, api_cache.py:140-141)."""

    def __init__(self, ckpt: dict, n_head: int = 8, dtype: str = "fp32", max_batch: int = 1, max_seq: Optional[int] = None,
                 device: int = 0, coalesce_ms: Optional[float] = None):
        """``coalesce_ms``: when set (and ``max_batch`` > 1), concurrent ``sample_kvcache`` calls -- the reference endpoint
        runs on a threadpool, api_cache.py:186-187 -- that arrive within that many milliseconds are decoded as one batch
        (``batcher.RequestBatcher``)."""
        self.tok2id: Dict[str, int] = ckpt["vocab"]
        self.id2tok = {i: t for t, i in self.tok2id.items()}
        pos_rows = int(ckpt["model"]["pos"].shape[0])
        self.seq_len = pos_rows                                 # SEQ_LEN of api_cache.py:36
        self.engine = Generator(ckpt["model"], n_head=n_head, dtype=dtype, max_batch=max_batch,
                                max_seq=max_seq or max(2 * pos_rows, 1088), device=device)
        self.batcher = None
        if coalesce_ms is not None and max_batch > 1:
            from .batcher import RequestBatcher
            self.batcher = RequestBatcher(self.engine, max_batch=max_batch, max_wait_ms=coalesce_ms)

    def to(self, device):            # the reference calls model.to(device).eval() (api_cache.py:161)
        return self

    def eval(self):
        return self


def sample_kvcache(model: KVModel, prompt: Sequence[str], max_len: int = 512, temperature: float = 1.0,
                   top_k: Optional[int] = 50, device: str = "cpu", seed: Optional[int] = None) -> List[str]:
    """Drop-in for reference api_cache.py:160-184 (same arguments; ``device`` is accepted and ignored:
    the engine already lives on its GPU).  Returns ALL tokens including the prompt, as strings.  ``seed`` is an
    addition: None (default) = a fresh seed per call, like the reference's draws from the global torch RNG."""
    ids = [model.tok2id[t] for t in prompt]                      # KeyError on OOV, like :162
    eos = model.tok2id.get("[END_SEQUENCE]", -1)                 # :181
    max_new = max(0, max_len - len(ids))                         # :166 range() of a negative number is empty
    if getattr(model, "batcher", None) is not None and seed is None:   # coalesced with the other requests in flight
        out = model.batcher.generate(ids, max_new, temperature, top_k, eos)
    else:
        out = model.engine.generate([ids], max_new, temperature, top_k, eos_id=eos, seed=seed)[0]
    return [model.id2tok[i] for i in out]


def sample(model: KVModel, prompt: Sequence[str], max_len: int = 512, temperature: float = 1.0, top_k: Optional[int] = 50,
           device: str = "cpu", seed: Optional[int] = None) -> List[str]:
    """Drop-in for the no-cache ``sample`` of reference generate_music/generate.py:46-61."""
    ids = [model.tok2id[t] for t in prompt]
    eos = model.tok2id.get("[END_SEQUENCE]", -1)
    out = model.engine.generate_nocache([ids], max(0, max_len - len(ids)), temperature, top_k, eos_id=eos, seed=seed)[0]
    return [model.id2tok[i] for i in out]


class Classifier:
    """DistilBERT emotion classifier replica (mirror of emotion_analysis/modeling.py:8-25, inference.py:12-22).

    ``state_dict``: HF DistilBertForSequenceClassification tensors, optionally with PEFT LoRA adapter
    tensors (merged here as W + (alpha/r) B A before upload).
    """

    def __init__(self, state_dict: Dict[str, torch.Tensor], n_heads: int = 12, max_tokens: int = 16384, device: int = 0,
                 tokenizer=None):
        self.lib = load_library()
        self._h = ctypes.c_void_p()
        sd = merge_lora_state_dict(state_dict)
        self.geometry: BertGeometry = infer_bert_geometry(sd, n_heads)
        g = self.geometry
        self.tokenizer = tokenizer
        self.max_tokens = int(max_tokens)
        cg = _BertGeometry(g.vocab_size, g.max_pos, g.dim, g.n_heads, g.n_layers, g.hidden_dim, g.num_labels)
        _check(self.lib, self.lib.mg_bert_create(ctypes.byref(cg), device, int(max_tokens), ctypes.byref(self._h)))
        for name, shape in expected_bert_keys(g).items():
            if name not in sd:
                raise RuntimeError(f"Error(s) in loading state_dict: missing key {name}")
            arr = sd[name].detach().to(torch.float32).contiguous().cpu().numpy()
            if tuple(arr.shape) != tuple(shape):
                raise RuntimeError(f"size mismatch for {name}: checkpoint {arr.shape} vs model {tuple(shape)}")
            shp = (ctypes.c_int64 * arr.ndim)(*arr.shape)
            _check(self.lib, self.lib.mg_bert_load_weight(self._h, name.encode(), _ptr(arr, ctypes.c_float), shp, arr.ndim))
        _check(self.lib, self.lib.mg_bert_finalize(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.mg_bert_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def classify(self, input_ids, attention_mask=None) -> Tuple[np.ndarray, np.ndarray]:
        """ids [N, T] (+ mask, 1 = token) -> (label ids [N], logits [N, num_labels])."""
        ids = _i32(input_ids)
        if ids.ndim != 2:
            raise ValueError("input_ids must be [N, T]")
        N, T = ids.shape
        mask = None if attention_mask is None else np.ascontiguousarray(attention_mask, dtype=np.uint8)
        if mask is not None and mask.shape != ids.shape:
            raise ValueError("attention_mask shape differs from input_ids")
        logits = np.empty((N, self.geometry.num_labels), np.float32)
        labels = np.empty(N, np.int32)
        _check(self.lib, self.lib.mg_classify(self._h, _ptr(ids, ctypes.c_int32), _ptr(mask, ctypes.c_uint8), N, T,
                                              _ptr(logits, ctypes.c_float), _ptr(labels, ctypes.c_int32)))
        return labels, logits

    def upload(self, input_ids, attention_mask=None) -> None:
        ids = _i32(input_ids)
        mask = None if attention_mask is None else np.ascontiguousarray(attention_mask, dtype=np.uint8)
        self._N = ids.shape[0]
        _check(self.lib, self.lib.mg_bert_upload(self._h, _ptr(ids, ctypes.c_int32), _ptr(mask, ctypes.c_uint8),
                                                 ids.shape[0], ids.shape[1]))

    def run(self) -> None:
        _check(self.lib, self.lib.mg_bert_run(self._h))

    def synchronize(self) -> None:
        _check(self.lib, self.lib.mg_bert_synchronize(self._h))

    def download(self) -> Tuple[np.ndarray, np.ndarray]:
        logits = np.empty((self._N, self.geometry.num_labels), np.float32)
        labels = np.empty(self._N, np.int32)
        _check(self.lib, self.lib.mg_bert_download(self._h, _ptr(logits, ctypes.c_float), _ptr(labels, ctypes.c_int32)))
        return labels, logits

    def stats(self) -> Dict[str, int]:
        a, b, c = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        _check(self.lib, self.lib.mg_bert_stats(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"kernel_launches": a.value, "h2d_bytes": b.value, "d2h_bytes": c.value}

    # -- the reference's call shape ---------------------------------------------------------------------
    def predict(self, text: str) -> str:
        """``inference.predict(text) -> label`` (emotion_analysis/inference.py:12-22); needs a tokenizer."""
        if self.tokenizer is None:
            raise RuntimeError("Classifier.predict needs the HF tokenizer the reference loads (modeling.py:14)")
        enc = self.tokenizer(text, return_tensors="np", truncation=True, padding=True)      # inference.py:16
        labels, _ = self.classify(enc["input_ids"], enc.get("attention_mask"))
        return ID2LABEL[int(labels[0])]

    def predict_ids(self, input_ids, attention_mask=None) -> List[str]:
        labels, _ = self.classify(input_ids, attention_mask)
        return [ID2LABEL[int(i)] for i in labels]

    # The three other read-outs of the same forward (emotion_analysis/inference.py:26-80): softmax over the logits of ONE
    # text, probabilities rounded to 4 decimals.  ``*_ids`` variants take token ids (no tokenizer needed).
    def _encode(self, text: str):
        if self.tokenizer is None:
            raise RuntimeError("this call needs the HF tokenizer the reference loads (modeling.py:14)")
        enc = self.tokenizer(text, return_tensors="np", truncation=True, padding=True)      # inference.py:30,45,66
        return enc["input_ids"], enc.get("attention_mask")

    def probabilities_ids(self, input_ids, attention_mask=None) -> np.ndarray:
        """softmax(logits, dim=1) in fp32, [N, num_labels] (inference.py:33, 49, 70)."""
        _, logits = self.classify(input_ids, attention_mask)
        return torch.softmax(torch.from_numpy(logits), dim=1).numpy()

    def predict_all_labels_ids(self, input_ids, attention_mask=None) -> Dict[str, float]:
        probs = self.probabilities_ids(input_ids, attention_mask)[0]
        return {ID2LABEL[i]: round(float(p), 4) for i, p in enumerate(probs)}               # inference.py:36-38

    def predict_top_k_labels_ids(self, input_ids, attention_mask=None, k: int = 3):
        probs = torch.from_numpy(self.probabilities_ids(input_ids, attention_mask))
        top_p, top_i = torch.topk(probs, k)                                                 # inference.py:52 (raises if k > labels)
        return [(ID2LABEL[int(i)], round(float(p), 4)) for i, p in zip(top_i[0], top_p[0])]  # :55-59

    def predict_labels_above_threshold_ids(self, input_ids, attention_mask=None, threshold: float = 0.2):
        probs = self.probabilities_ids(input_ids, attention_mask)[0]
        return [(ID2LABEL[i], round(float(p), 4)) for i, p in enumerate(probs) if float(p) > threshold]   # :73-79

    def predict_all_labels(self, text: str) -> Dict[str, float]:
        """``inference.predict_all_labels(text)`` (emotion_analysis/inference.py:26-38)."""
        return self.predict_all_labels_ids(*self._encode(text))

    def predict_top_k_labels(self, text: str, k: int = 3):
        """``inference.predict_top_k_labels(text, k)`` (emotion_analysis/inference.py:41-60)."""
        return self.predict_top_k_labels_ids(*self._encode(text), k=k)

    def predict_labels_above_threshold(self, text: str, threshold: float = 0.2):
        """``inference.predict_labels_above_threshold(text, threshold)`` (emotion_analysis/inference.py:62-80)."""
        return self.predict_labels_above_threshold_ids(*self._encode(text), threshold=threshold)


def tc_gemm(A: np.ndarray, W: np.ndarray, bias: Optional[np.ndarray] = None, act: int = 0, device: int = 0) -> np.ndarray:
    """C = act(A W^T + bias) through the tcgen05 / TMA GEMM kernel (kernel-level test hook)."""
    lib = load_library()
    A = np.ascontiguousarray(A, np.float32)
    W = np.ascontiguousarray(W, np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, np.float32)
    M, K = A.shape
    N = W.shape[0]
    C = np.empty((M, N), np.float32)
    _check(lib, lib.mg_test_gemm_bf16(device, _ptr(A, ctypes.c_float), _ptr(W, ctypes.c_float), _ptr(b, ctypes.c_float), M,
                                      N, K, act, _ptr(C, ctypes.c_float)))
    return C


def grid_plan(d_model: int, d_ff: int, n_layer: int, vocab: int, B: int, n_cta: int = 148):
    """Work plan of the grid-synchronous decode kernel (decode_grid.cu) for a batch of B sequences: returns
    (tn, ks, items) with tn / ks = n-tiles per item / k-splits per phase kind (qkv, attention, out_proj, mlp.0, mlp.2, head) and
    items[c] = [(phase, row_tile, sequence_group), ...] of CTA c in phase order.  Host-only test hook (no GPU needed)."""
    lib = load_library()
    tn_ks = np.zeros(16, np.int32)
    items = np.zeros((n_cta, 96, 4), np.int16)
    n_items = np.zeros(n_cta, np.int32)
    _check(lib, lib.mg_test_grid_plan(int(d_model), int(d_ff), int(n_layer), int(vocab), int(B), int(n_cta), _ptr(tn_ks, ctypes.c_int32),
                                      items.ctypes.data_as(ctypes.POINTER(ctypes.c_int16)), _ptr(n_items, ctypes.c_int32)))
    per_cta = [[tuple(int(v) for v in items[c, i, :3]) for i in range(int(n_items[c]))] for c in range(n_cta)]
    return tn_ks[:8].tolist(), tn_ks[8:].tolist(), per_cta
