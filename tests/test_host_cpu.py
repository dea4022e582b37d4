"""CPU: host-side logic and the C-ABI surface (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import mgea_b200 as mg
from mgea_b200 import engine as eng
from conftest import ROOT, has_gpu


def test_geometry_inference_and_remap_follow_the_reference_rules():
    geo = mg.GEOMETRIES["train_large"]
    sd = mg.make_state_dict(geo, 0)
    assert len(sd) == 4 + 12 * geo.n_layer
    assert mg.infer_geometry(sd, 8) == geo                       # api_cache.py:31-37 (+ n_head :112)
    r = mg.remap_state_dict(sd)
    assert set(r) == set(mg.expected_keys(geo))
    assert r["layers.3.mlp.2.weight"].shape == (256, 1024) and r["pos_emb"].shape == (255, 256)
    assert r["layers.0.attn.in_proj_weight"] is sd["tr.layers.0.self_attn.in_proj_weight"]   # tensors untouched
    with pytest.raises(ValueError):
        mg.infer_geometry(sd, 7)


def test_synthetic_vocab_satisfies_the_prompt_builder_contract():
    v = mg.build_synthetic_vocab(8324)
    assert len(v) == 8324 and sorted(v.values()) == list(range(8324))
    assert [t for t, _ in sorted(v.items(), key=lambda kv: kv[1])] == sorted(v)      # ids by sorted() order
    p = mg.build_prompt(v, 121.7, "E♭ Major", ["Strings", "Piano"])
    assert p == ["[START_SEQUENCE]", "[BPM] 122.0", "[KEY_SIGNATURE] E- major", "[INSTRUMENT] Violin",
                 "[INSTRUMENT] Acoustic Grand Piano"]
    assert all(t in v for t in p)
    with pytest.raises(KeyError):
        mg.encode(v, ["[NOT A TOKEN]"])                          # api_cache.py:162 raises KeyError
    prompts = mg.synthetic_prompts(v, 64, seed=0)
    assert all(3 <= len(q) <= 6 for q in prompts) and prompts == mg.synthetic_prompts(v, 64, seed=0)


def test_shard_range_is_a_balanced_partition():
    for n in (0, 1, 7, 64, 512, 513):
        for w in (1, 2, 4, 8):
            parts = [mg.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        mg.shard_range(4, 2, 2)


def test_header_symbols_are_all_exported_by_the_library():
    header = open(os.path.join(ROOT, "include", "mg_engine.h")).read()
    declared = sorted(set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(eng.EXPORTED_SYMBOLS)
    lib = mg.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mg_abi_version() == 1


def test_pack_prompts_and_status_mapping():
    flat, offs = eng._pack_prompts([[1, 2, 3], [4], [5, 6]])
    assert flat.tolist() == [1, 2, 3, 4, 5, 6] and offs.tolist() == [0, 3, 4, 6]
    with pytest.raises(ValueError):
        eng._pack_prompts([])
    lib = mg.load_library()
    for rc, exc in ((eng.MG_E_ARG, ValueError), (eng.MG_E_TOKEN, ValueError), (eng.MG_E_OOM, MemoryError),
                    (eng.MG_E_TOPK, RuntimeError), (eng.MG_E_PROMPT_TOO_LONG, RuntimeError), (eng.MG_E_CUDA, RuntimeError)):
        with pytest.raises(exc):
            eng._check(lib, rc)


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_engine_creation_fails_loudly_without_a_gpu():
    ck = mg.make_checkpoint(mg.GEOMETRIES["tiny"], 0)
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        mg.Generator(ck["model"], n_head=2, dtype="fp32")
    with pytest.raises(RuntimeError):
        mg.Classifier(mg.make_bert_state_dict(mg.TINY_BERT, 0), n_heads=2)
    with pytest.raises(RuntimeError, match="not found"):
        mg.load_library("/nonexistent/libmgea_b200.so")


def test_missing_library_is_an_error_not_a_fallback(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mg.load_library(str(tmp_path / "libmgea_b200.so"))


def test_bert_geometry_inference_and_key_count():
    sd = mg.merge_lora_state_dict(mg.make_bert_state_dict(mg.DISTILBERT_BASE, 0))
    assert len(sd) == 104                                          # 104 tensors
    assert sum(v.numel() for v in sd.values()) == 66_975_004
    from mgea_b200.bert_checkpoint import infer_bert_geometry
    assert infer_bert_geometry(sd) == mg.DISTILBERT_BASE
    assert mg.ID2LABEL[27] == "neutral" and len(mg.ID2LABEL) == 28


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "music-generation-emotion-adaptive_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("test oracle", ""), f


def test_pipeline_prompt_building_uses_the_service_order():
    v = mg.build_synthetic_vocab(8324)
    prm = mg.synthetic_music_params("joy")
    assert set(prm) == {"emotion", "bpm", "key", "scale_type", "inst_family", "all_families"}     # EATS.py:29-37
    p = mg.build_prompt(v, prm["bpm"], prm["key"], prm["all_families"])
    assert p[0] == "[START_SEQUENCE]" and p[1].startswith("[BPM]") and p[2].startswith("[KEY_SIGNATURE]")
    assert all(t in v for t in p) and 3 <= len(p) <= 6


def test_transposed_v_block_layout_matches_the_mma_fragments():
    """The persistent kernel stores V per 32-key block as [dim][32 positions] with key 8j + 2t + e at position
    8t + 2j + e (csrc/decode_mega.cu, attn_tc / mega_relayout_kv_kernel).  Restated here: the map is a bijection, lane t
    of the score MMAs (keys 8j + 2t + {0, 1} of MMA j) ends up owning positions [8t, 8t + 8) -- the 16 bytes it loads for
    the P.V MMAs -- and the two k-steps of that MMA pair (positions 8t + 4u + {0..3}) see keys of score MMAs 2u, 2u + 1."""
    def pos_of(key):                      # as in the kernels
        return 8 * ((key >> 1) & 3) + 2 * (key >> 3) + (key & 1)
    assert sorted(pos_of(k) for k in range(32)) == list(range(32))
    for t in range(4):
        owned = sorted(pos_of(8 * j + 2 * t + e) for j in range(4) for e in range(2))
        assert owned == list(range(8 * t, 8 * t + 8))
        for u in range(2):
            keys = [8 * j + 2 * t + e for j in (2 * u, 2 * u + 1) for e in range(2)]
            assert sorted(pos_of(k) for k in keys) == list(range(8 * t + 4 * u, 8 * t + 4 * u + 4))


class _FakeEngine:
    """Stands in for Generator on the CPU: row b = prompt b followed by max_new[b] copies of (sum(prompt) % 97)."""

    def __init__(self, delay=0.01):
        import threading
        self.calls, self.delay, self.lock = [], delay, threading.Lock()

    def generate(self, prompts, max_new, temperature=1.0, top_k=50, eos_id=-1, seed=0, seq_index_base=0):
        import time
        for p in prompts:
            if len(p) > 8:
                raise RuntimeError("prompt longer than the position table")
        time.sleep(self.delay)
        with self.lock:
            self.calls.append((len(prompts), temperature, top_k, seq_index_base))
        return [list(p) + [sum(p) % 97] * n for p, n in zip(prompts, max_new)]


def test_request_batcher_coalesces_concurrent_requests_and_routes_results():
    import threading
    eng = _FakeEngine()
    b = mg.RequestBatcher(eng, max_batch=8, max_wait_ms=50.0)
    results = {}

    def worker(i):
        results[i] = b.generate([i, i + 1, 2], 3 + i % 4, 1.0, 50 if i % 5 else 7, -1)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(20)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    b.close()
    for i in range(20):
        assert results[i] == [i, i + 1, 2] + [(2 * i + 3) % 97] * (3 + i % 4)          # every caller got ITS row
    assert sum(c[0] for c in eng.calls) == 20 and len(eng.calls) < 20                  # requests were coalesced ...
    assert all(c[0] <= 8 for c in eng.calls)                                           # ... within max_batch ...
    assert all(len({c[2]}) == 1 for c in eng.calls)                                    # ... one sampling setting per call
    bases = [c[3] for c in eng.calls]
    assert len(set(bases)) == len(bases)                                               # distinct Philox stream ranges


def test_request_batcher_isolates_a_failing_request():
    import threading
    eng = _FakeEngine()
    b = mg.RequestBatcher(eng, max_batch=4, max_wait_ms=100.0)
    out, err = {}, {}

    def worker(i, prompt):
        try:
            out[i] = b.generate(prompt, 2)
        except RuntimeError as e:
            err[i] = str(e)

    threads = [threading.Thread(target=worker, args=(0, [1, 2])), threading.Thread(target=worker, args=(1, list(range(9)))),
               threading.Thread(target=worker, args=(2, [5]))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    b.close()
    assert out == {0: [1, 2, 3, 3], 2: [5, 5, 5]} and list(err) == [1] and "position table" in err[1]
    with pytest.raises(RuntimeError):
        b.submit([1], 1)


def test_request_batcher_retry_keeps_distinct_streams_and_fails_whole_batch_on_device_errors():
    """the per-request retry after a failed batch must give request i the Philox sequence index base + i (not
    base for all of them), and a device failure (MG_E_CUDA / MG_E_OOM) must fail the batch instead of N serial re-runs."""
    import threading
    eng = _FakeEngine()
    b = mg.RequestBatcher(eng, max_batch=4, max_wait_ms=100.0, seed=5)
    threads = [threading.Thread(target=lambda p=p: _swallow(b.generate, p, 2)) for p in ([1, 2], list(range(9)), [5], [7, 7])]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    singles = [c for c in eng.calls if c[0] == 1]
    assert len(singles) == 3 and len({c[3] for c in singles}) == 3                     # three survivors, three streams

    class _Broken(_FakeEngine):
        def generate(self, prompts, *a, **k):
            with self.lock:
                self.calls.append(len(prompts))
            raise RuntimeError("[mg status -4] cudaStreamSynchronize: an illegal memory access was encountered")
    bad = _Broken()
    b2 = mg.RequestBatcher(bad, max_batch=4, max_wait_ms=100.0)
    errs = []
    threads = [threading.Thread(target=lambda i=i: errs.append(_swallow(b2.generate, [i], 2))) for i in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    b.close(); b2.close()
    assert len(errs) == 3 and all(isinstance(e, RuntimeError) for e in errs)
    assert sum(bad.calls) == 3                                                          # no re-run of a failed batch


def _swallow(fn, *a):
    try:
        return fn(*a)
    except Exception as e:                # noqa: BLE001 - the tests inspect the exception object
        return e


def test_batcher_and_engine_default_to_fresh_seeds():
    """seed=None must not mean seed 0 (identical prompts would always give the identical piece)."""
    from mgea_b200 import engine as eng
    assert eng._seed(7) == 7 and eng._seed(None) != eng._seed(None)
    b1, b2 = mg.RequestBatcher(_FakeEngine(), seed=None), mg.RequestBatcher(_FakeEngine(), seed=None)
    assert b1._seed != b2._seed
    b1.close(); b2.close()
    import inspect
    for fn in (mg.sample_kvcache, mg.sample, mg.Generator.generate, mg.Generator.run, mg.classify_prompt_generate):
        assert inspect.signature(fn).parameters["seed"].default is None, fn


def test_peft_spellings_of_the_finetuned_classifier_load():
    """SURVEY 8(f) rank 4 / ADVICE r1: the adapter was trained with TaskType.SEQ_CLS (Scripts/finetuneDistillBert.ipynb:790-795),
    so pre_classifier / classifier are PEFT modules_to_save.  Every spelling modeling.load_model() can yield must map to the
    plain HF keys with the FINE-TUNED heads and W + 2 B A on q_lin / v_lin."""
    import torch
    geo = mg.TINY_BERT
    base = mg.make_bert_state_dict(geo, 0, with_lora=False)
    g = torch.Generator().manual_seed(1)
    tuned = {k: torch.randn(v.shape, generator=g) for k, v in base.items() if k.split(".")[0] in ("pre_classifier", "classifier")}
    lora = {}
    for i in range(geo.n_layers):
        for lin in ("q_lin", "v_lin"):
            m = f"distilbert.transformer.layer.{i}.attention.{lin}"
            lora[m] = (0.05 * torch.randn(8, geo.dim, generator=g), 0.05 * torch.randn(geo.dim, 8, generator=g))

    want = dict(base)
    want.update(tuned)
    for m, (A, B) in lora.items():
        want[m + ".weight"] = base[m + ".weight"] + 2.0 * (B @ A)

    def check(sd):
        got = mg.merge_lora_state_dict(sd)
        assert set(got) == set(want), sorted(set(got) ^ set(want))[:4]
        for k in want:
            assert torch.allclose(got[k], want[k], atol=1e-6), k

    # (1) adapter_model.safetensors as PEFT writes it + the base checkpoint, merged in BOTH dict orders
    disk = {f"base_model.model.{m}.lora_A.weight": A for m, (A, B) in lora.items()}
    disk.update({f"base_model.model.{m}.lora_B.weight": B for m, (A, B) in lora.items()})
    disk.update({f"base_model.model.{k}": v for k, v in tuned.items()})
    check({**base, **disk})
    check({**disk, **base})
    # (2) live PeftModel.state_dict(): base_layer wrappers, .default adapters, original_module + modules_to_save.default heads
    live = {}
    for k, v in base.items():
        mod = k.rsplit(".", 1)[0]
        if mod in lora:
            live[f"base_model.model.{mod}.base_layer.{k.rsplit('.', 1)[1]}"] = v
        elif k.split(".")[0] in ("pre_classifier", "classifier"):
            leaf = k.split(".", 1)[1]
            live[f"base_model.model.{k.split('.')[0]}.original_module.{leaf}"] = v
            live[f"base_model.model.{k.split('.')[0]}.modules_to_save.default.{leaf}"] = tuned[k]
        else:
            live[f"base_model.model.{k}"] = v
    for m, (A, B) in lora.items():
        live[f"base_model.model.{m}.lora_A.default.weight"] = A
        live[f"base_model.model.{m}.lora_B.default.weight"] = B
    check(live)
    check(dict(reversed(list(live.items()))))


def test_classifier_readouts_follow_inference_py():
    """predict_all_labels / predict_top_k_labels / predict_labels_above_threshold (emotion_analysis/inference.py:26-80):
    softmax of one text's logits, rounded to 4 decimals, k default 3, threshold default 0.2 and strict."""
    import numpy as np
    import torch
    from mgea_b200.engine import Classifier
    logits = torch.randn(1, 28, generator=torch.Generator().manual_seed(3)).numpy() * 2.0

    class _Tok:
        def __call__(self, text, return_tensors, truncation, padding):
            assert (return_tensors, truncation, padding) == ("np", True, True)
            return {"input_ids": np.array([[101, 7, 102]]), "attention_mask": np.array([[1, 1, 1]])}

    class _Stub(Classifier):
        def __init__(self):
            self.tokenizer = _Tok()

        def classify(self, ids, mask=None):
            return logits.argmax(1).astype(np.int32), logits

        def __del__(self):
            pass

    c = _Stub()
    p = torch.softmax(torch.from_numpy(logits), 1)[0]
    allp = c.predict_all_labels("x")
    assert list(allp) == [mg.ID2LABEL[i] for i in range(28)] and allp == {mg.ID2LABEL[i]: round(float(p[i]), 4) for i in range(28)}
    top = c.predict_top_k_labels("x")
    order = torch.argsort(p, descending=True)[:3]
    assert top == [(mg.ID2LABEL[int(i)], round(float(p[i]), 4)) for i in order]
    assert len(c.predict_top_k_labels("x", k=5)) == 5
    thr = c.predict_labels_above_threshold("x")
    assert thr == [(mg.ID2LABEL[i], round(float(p[i]), 4)) for i in range(28) if float(p[i]) > 0.2]
    assert c.predict("x") == mg.ID2LABEL[int(p.argmax())]
    with pytest.raises(RuntimeError):
        c.predict_top_k_labels("x", k=29)                       # torch.topk raises in the reference too


def test_generate_with_no_room_returns_the_prompt_like_the_reference():
    """api_cache.py:166: range(max_len - Tp) is empty when max_len <= len(prompt): the prompt comes back unchanged."""
    from mgea_b200.engine import Generator
    g = Generator.__new__(Generator)
    g.geometry = mg.GEOMETRIES["tiny"]
    assert Generator.generate(g, [[1, 2, 3]], -4) == [[1, 2, 3]] and Generator.generate(g, [[1], [2, 3]], [0, -1]) == [[1], [2, 3]]
    with pytest.raises(RuntimeError):
        Generator.generate(g, [list(range(g.geometry.pos_rows + 1))], 0)
    assert mg.clf_capacity(type("C", (), {"max_tokens": 4096})()) == 4096


def test_eats_music_params_reproduce_the_reference_mapping():
    """pipeline.eats_music_params over the reference's own table (tests/golden/eats_table.json, dumped from
    emotion_analysis/EATS.py by oracle/make_golden.py) returns what the reference's get_music_params returned for all 28
    labels under random.seed(0); every prompt it leads to is encodable in the synthetic vocabulary (api_cache.py:194-203)."""
    import json
    import os
    import random
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eats_table.json")
    with open(path) as f:
        gold = json.load(f)
    fn = mg.eats_music_params(mg.load_eats_table(path))
    random.seed(gold["seed"])
    got = [fn(lab) for lab in gold["labels"]]
    assert got == gold["get_music_params"]
    with pytest.raises(ValueError):
        fn("not-an-emotion")
    tok2id = mg.build_synthetic_vocab(8324)
    for prm in got:
        ids = mg.encode(tok2id, mg.build_prompt(tok2id, prm["bpm"], prm["key"], prm["all_families"]))
        assert 3 <= len(ids) <= 6


def test_note_table_follows_the_reference_token_rules():
    """vocab.note_table = the per-token work of api_cache.py:208-221 done once per vocabulary entry."""
    from mgea_b200 import vocab as V
    tok2id = {"[START_SEQUENCE]": 0, "[INSTRUMENT] Violin": 1, "[INSTRUMENT] Kazoo": 2,
              "[NOTE] [PITCH:C#4] [START:1.5] [END:2.25] [DURATION:0.75]": 3, "[NOTE]": 4, "[PITCH]": 5,
              "[NOTE] [PITCH:B-3] [START:0] [END:1] [DURATION:1]": 6}
    kind, value, start, end = V.note_table(tok2id, 8)
    assert kind.tolist() == [0, 1, 1, 2, 0, 0, 2, 0]
    assert value[1] == 40 and value[2] == 0                              # unknown instrument name -> program 0 (api_cache.py:212)
    assert value[3] == 61 and (start[3], end[3]) == (1.5, 2.25)
    assert value[6] == 12 * (-3 + 1) + 11                                # pretty_midi reads "B-3" as B, octave -3
    assert V.note_name_to_number("A4") == 69 and V.note_name_to_number("Eb2") == 39 and V.note_name_to_number("c0") == 12
    with pytest.raises(ValueError):
        V.note_table({"[NOTE] [PITCH:H2] [START:0] [END:1] [DURATION:1]": 0}, 1)
    from oracle import detok as odetok
    toks = ["[NOTE] [PITCH:C#4] [START:1.5] [END:2.25] [DURATION:0.75]", "[INSTRUMENT] Violin", "[PITCH]",
            "[NOTE] [PITCH:C#4] [START:1.5] [END:2.25] [DURATION:0.75]", "[INSTRUMENT] Kazoo", "[INSTRUMENT] Violin",
            "[NOTE] [PITCH:B-3] [START:0] [END:1] [DURATION:1]"]
    got = odetok.tokens_to_instruments(toks, lambda n: V.GM_PROGRAMS.get(n, 0), V.note_name_to_number)
    assert [(g["name"], g["program"], g["notes"]) for g in got] == [("Violin", 40, [(61, 1.5, 2.25)]), ("Kazoo", 0, []),
                                                                     ("Violin", 40, [(-13, 0.0, 1.0)])]


class _FakeSlotEngine:
    """Host-logic stand-in for the slot-session calls of Generator: a slot emits prompt[-1] + 1, + 2, ... one token per step."""

    def __init__(self):
        self.rows, self.left, self.log = {}, {}, []

    def slots_begin(self, n_slots, max_len, temperature, top_k, eos_id, seed):
        self.n = n_slots

    def slots_admit(self, slots, prompts, max_new, seq_index):
        for p in prompts:
            if any(t < 0 for t in p):
                raise ValueError("prompt token id outside [0, vocab)")
        for s, p, m, i in zip(slots, prompts, max_new, seq_index):
            assert s not in self.rows
            self.rows[s], self.left[s] = list(p), m
            self.log.append((s, i))

    def slots_step(self, n_steps):
        import numpy as np
        for s in self.rows:
            k = min(n_steps, self.left[s])
            self.rows[s] += [self.rows[s][-1] + j + 1 for j in range(k)]
            self.left[s] -= k
        fin = np.array([s not in self.rows or self.left[s] == 0 for s in range(self.n)])
        return fin, np.zeros(self.n, np.int32)

    def slots_fetch(self, slot, cap):
        self.left.pop(slot)
        return self.rows.pop(slot)

    def slots_end(self):
        self.ended = True


def test_continuous_batcher_admits_between_chunks_reuses_slots_and_isolates_bad_requests():
    eng = _FakeSlotEngine()
    cb = mg.ContinuousBatcher(eng, n_slots=3, max_len=40, chunk_steps=4, seed=0, first_seq_index=100)
    futs = [cb.submit([10 * i, 10 * i + 1], 3 + 2 * i) for i in range(8)]
    bad = cb.submit([5, -1], 4)
    long_one = cb.submit([1], 100)
    res = [f.result(timeout=30) for f in futs]
    with pytest.raises(ValueError):
        bad.result(timeout=30)
    with pytest.raises(ValueError):
        long_one.result(timeout=30)
    assert cb.generate([7, 8], 0) == [7, 8]
    cb.close()
    for i, r in enumerate(res):
        n = 3 + 2 * i
        assert r[:2] == [10 * i, 10 * i + 1] and len(r) == 2 + n
    assert sorted(i for _, i in eng.log) == list(range(100, 108))          # one Philox stream index per admitted request
    assert max(s for s, _ in eng.log) <= 2 and len(eng.log) == 8           # 8 requests through 3 slots
    assert len(cb.admissions) >= 3 and eng.ended


@pytest.mark.parametrize("geo_name,B", [("train_large", 64), ("train_large", 16), ("train_large", 1), ("train_large2", 64), ("train_large2", 9),
                                        ("train_mini", 3)])
def test_grid_kernel_plan_covers_every_weight_tile_exactly_once(geo_name, B):
    """Host logic of the grid-synchronous decode kernel (decode_grid.cu: grid_plan): per phase every 16-row tile of the phase's
    matrix meets every sequence group exactly once, items of a CTA come in phase order, n-tiles x k-splits = the 8 warps of a CTA,
    a warp never holds more than 16 k-step pairs, and the items are spread evenly over the CTAs."""
    geo = mg.GEOMETRIES[geo_name]
    n_cta = 148
    tn, ks, items = mg.grid_plan(geo.d_model, geo.d_ff, geo.n_layer, geo.vocab_size, B, n_cta)
    K_QKV, K_OUT, K_MLP1, K_MLP2 = 0, 2, 3, 4
    nt = (B + 7) // 8
    for kind, K in ((K_OUT, geo.d_model), (K_MLP2, geo.d_ff)):
        assert tn[kind] * ks[kind] == 8 and (K // 32) % ks[kind] == 0 and (K // 32) // ks[kind] <= 16
    want = {}
    for l in range(geo.n_layer):
        want[5 * l + K_QKV] = (3 * geo.d_model // 16, 1)
        want[5 * l + K_OUT] = (geo.d_model // 16, -(-nt // tn[K_OUT]))
        want[5 * l + K_MLP1] = (geo.d_ff // 16, 1)
        want[5 * l + K_MLP2] = (geo.d_model // 16, -(-nt // tn[K_MLP2]))
    want[5 * geo.n_layer] = (-(-geo.vocab_size // 16), 1)
    seen = {}
    for c, lst in enumerate(items):
        assert [it[0] for it in lst] == sorted(it[0] for it in lst), c                 # phase order inside a CTA
        for ph, rt, g in lst:
            key = (ph, rt, g)
            assert key not in seen, key
            seen[key] = c
    for ph, (tiles, groups) in want.items():
        got = sorted((rt, g) for (p, rt, g) in seen if p == ph)
        assert got == [(rt, g) for rt in range(tiles) for g in range(groups)], ph
    assert set(p for (p, _, _) in seen) == set(want)
    counts = [len(lst) for lst in items]
    assert max(counts) - min(counts) <= 1 and max(counts) <= 96
