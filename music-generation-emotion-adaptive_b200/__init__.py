"""B200-native engine for the reference's one hot path (import name: ``mgea_b200``).

KV-cached autoregressive MIDI-token decoding (reference api_cache.py:39-106,159-184; no-cache twin
generate_music/generate.py:25-61) and the DistilBERT emotion-classifier forward (reference
emotion_analysis/inference.py:12-22), as hand-written sm_100a CUDA behind the C ABI declared in
include/mg_engine.h.  There is no CPU fallback: importing works anywhere, but constructing an engine
needs the built library and a GPU.
"""
from .checkpoint import (GEOMETRIES, Geometry, expected_keys, infer_geometry, make_checkpoint,  # noqa: F401
                         make_state_dict, remap_state_dict, state_dict_digest)
from .vocab import (build_prompt, build_synthetic_vocab, closest_bpm_token, encode,  # noqa: F401
                    normalize_key_signature, synthetic_prompts)

from .bert_checkpoint import (DISTILBERT_BASE, ID2LABEL, TINY_BERT, BertGeometry, make_bert_state_dict,  # noqa: F401
                              merge_lora_state_dict)
from .engine import (Classifier, Generator, KVModel, grid_plan, load_library, sample, sample_kvcache, tc_gemm)  # noqa: F401
from .batcher import ContinuousBatcher, RequestBatcher  # noqa: F401
from .pipeline import (classify_prompt_generate, clf_capacity, eats_music_params, load_eats_table,  # noqa: F401
                       synthetic_music_params)
from .replicas import gather_token_lists, shard, shard_range  # noqa: F401

__all__ = [
    "Classifier", "Generator", "KVModel", "RequestBatcher", "ContinuousBatcher", "load_library", "sample", "sample_kvcache", "tc_gemm", "grid_plan", "gather_token_lists",
    "shard", "shard_range", "classify_prompt_generate", "clf_capacity", "synthetic_music_params", "eats_music_params", "load_eats_table", "DISTILBERT_BASE", "ID2LABEL", "TINY_BERT", "BertGeometry", "make_bert_state_dict",
    "merge_lora_state_dict",
    "GEOMETRIES", "Geometry", "expected_keys", "infer_geometry", "make_checkpoint", "make_state_dict",
    "remap_state_dict", "state_dict_digest", "build_prompt", "build_synthetic_vocab",
    "closest_bpm_token", "encode", "normalize_key_signature", "synthetic_prompts",
]
