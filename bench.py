#!/usr/bin/env python
"""Benchmark of the hot path: batched KV-cache MIDI-token decoding (BASELINE.json config 3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one whole generation job of the workload: prefill of 64 production-shaped prompts +
1024 decode steps with top-k-40 Philox sampling, bf16 weights + bf16 KV cache, on ONE GPU
(65,536 new tokens per step per GPU).  At N > 1 every rank is a full replica running its own batch of
64 (weak scaling; no collective on the decode path); only the end-to-end leg gathers the
finished token lists.

  value    tokens/s, device time (CUDA events on the engine's own stream), prompts already resident in HBM
  e2e      the same metric through the public host call (host prompt buffers -> host token buffers)
  roofline algorithmic HBM bytes of the decode steps (SURVEY.md 8d formula) / duration of the persistent
           decode kernel (one launch per job), against MEASURED_PEAKS.json
  cpu_baseline / --impl reference : the reference's CPU loop (oracle port of api_cache.py:159-184; the
           reference tree itself does not exist on the GPU box) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

GEOMETRY = "train_large"          # d 256, 8 heads (hd 32), 4 layers, 255 position rows, V 8324
BATCH, NEW_TOKENS, TOP_K, TEMPERATURE = 64, 1024, 40, 1.0
WORKLOAD = ("config3: MIDI GPT-2 (train_large.py size) top-k=40 sampling, 1024 new tokens, batch 64, "
            "bf16 weights + bf16 KV cache")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), float(j.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def sustained_tflops():
    """Driver-measured cuBLAS bf16 throughput back to back for seconds (MEASURED_PEAKS.json), None when absent."""
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
    except Exception:
        return None


def algorithmic_bytes(geo, prompt_lens, n_steps, s_w=2, s_kv=2):
    """bytes(step) = W + sum_b (T_b + 1) kappa + B kappa + B d s_w."""
    d, L, V, B = geo.d_model, geo.n_layer, geo.vocab_size, len(prompt_lens)
    W = (L * (12 * d * d + 13 * d) + V * d + V) * s_w
    kappa = L * 2 * d * s_kv
    total = 0
    for i in range(n_steps):
        kv_read = sum((tp + i + 1) * kappa for tp in prompt_lens)
        total += W + kv_read + B * kappa + B * d * s_w
    return total, W, kappa


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            pass

    def run(self):
        while self.ok and not self.stop_flag:
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                mask = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if self.ok and self.is_alive():
            self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": int(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def ncu_traffic_bytes():
    """DRAM bytes (read + write) per launch of the persistent decode kernel from the committed `ncu --set full`
    capture of this same workload (the newest of profiles/r2n_ / r2m_ / r1j_decode_mega_full_raw.csv); None when none is there."""
    import csv
    path = next((q for q in (os.path.join(ROOT, "profiles", f"{r}_decode_mega_full_raw.csv") for r in ("r2n", "r2m", "r1j"))
                 if os.path.exists(q)), "")
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        tot = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(name)
            tot += float(rows[2][i].replace(",", "")) * scale[units[i]]
        return tot
    except Exception:
        return None


def host_threads():
    """Host cores this process may actually use (cgroup / affinity aware), capped at 32 for the small GEMMs."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, 32))


def config_dict(geo, prompt_lens, world):
    """The `config` object both arms print (the reference arm measures a bounded sample of the SAME workload)."""
    return {"workload": WORKLOAD, "geometry": GEOMETRY, "d_model": geo.d_model, "n_head": geo.n_head,
            "n_layer": geo.n_layer, "vocab": geo.vocab_size, "batch_per_gpu": BATCH, "new_tokens": NEW_TOKENS,
            "top_k": TOP_K, "temperature": TEMPERATURE, "prompt_tokens": f"{min(prompt_lens)}-{max(prompt_lens)}",
            "parallelism": f"replicas x{world} (batch sharded, no decode-path collective)",
            "weights": "random-init (seed 0), synthetic 8324-token vocab",
            "l2": "KV working set grows to 270 MB per GPU (> 126 MB L2), rewritten every step: inputs larger than L2, no flush"}


def build_workload(mg):
    geo = mg.GEOMETRIES[GEOMETRY]
    ck = mg.make_checkpoint(geo, 0)
    prompts = [mg.encode(ck["vocab"], p) for p in mg.synthetic_prompts(ck["vocab"], BATCH, seed=0)]
    return geo, ck, prompts


# ---------------------------------------------------------------------------------------------------
# CPU reference loop (oracle port of the reference's sample_kvcache; batch-1 like the reference)
# ---------------------------------------------------------------------------------------------------
def cpu_reference_sample(ck, geo, prompt, n_new, threads):
    from oracle import gpt_kv
    import mgea_b200 as mg
    torch.set_num_threads(threads)
    model = gpt_kv.KVModelOracle(mg.remap_state_dict(ck["model"]), geo.n_head)
    g = torch.Generator().manual_seed(0)
    t0 = time.perf_counter()
    out = gpt_kv.sample_ids(model, prompt, max_len=len(prompt) + n_new, temperature=TEMPERATURE, top_k=TOP_K, eos_id=-1,
                            generator=g)
    dt = time.perf_counter() - t0
    assert len(out) == len(prompt) + n_new
    return n_new / dt, dt


def run_reference_arm(args):
    """--impl reference: the reference's own CPU path (batch-1 loop, api_cache.py:159-184) on host cores.  One step = ONE of
    the 64 prompts of the workload run for all 1024 new tokens (the reference sampler is batch-1 only: a whole job is 64 such
    runs back to back, so tokens/s of one run IS the job's tokens/s)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import mgea_b200 as mg
    geo, ck, prompts = build_workload(mg)
    threads = host_threads()
    n_new = NEW_TOKENS
    for _ in range(min(args.warmup, 2)):
        cpu_reference_sample(ck, geo, prompts[0], 64, threads)
    t_all = 0.0
    for i in range(args.steps):
        _, dt = cpu_reference_sample(ck, geo, prompts[i % len(prompts)], n_new, threads)
        t_all += dt
    value = args.steps * n_new / t_all
    sample = (f"oracle port of the reference batch-1 sample_kvcache loop: 1 of the 64 prompts per step, all {n_new} new tokens "
              f"(cache length to 1030), top-k 40, fp32, torch CPU, {threads} threads")
    line = {
        "impl": "reference", "metric": "midi_decode_tokens_per_sec", "value": value, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(geo, [len(p) for p in prompts], max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the batch-1 latency and classifier side measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch.distributed as dist
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"                     # keep NCCL's version banner out of stdout: ONE JSON line
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.warmup < 3:
        args.warmup = 3                                           # timing rule: W >= 3

    import mgea_b200 as mg
    geo, ck, prompts = build_workload(mg)
    prompt_lens = [len(p) for p in prompts]
    eng = mg.Generator(ck["model"], n_head=geo.n_head, dtype="bf16", max_batch=BATCH, max_seq=1088, device=local_rank)
    hbm_peak, tf_peak, peak_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.synchronize()

    def device_step(seed):
        eng.upload(prompts, NEW_TOKENS)                           # inputs resident in HBM before the timed region
        eng.synchronize()
        eng.run(TEMPERATURE, TOP_K, eos_id=-1, seed=seed, seq_index_base=rank * BATCH)
        eng.synchronize()
        return eng.last_timing()

    for i in range(args.warmup):
        device_step(100 + i)
    launches0 = eng.stats()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    wall0 = time.perf_counter()
    tot_ms = dec_ms = 0.0
    dec_steps = 0
    for i in range(args.steps):
        t = device_step(i)
        tot_ms += t["total_ms"]
        dec_ms += t["decode_ms"]
        dec_steps += t["steps"]
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.result()
    launches = eng.stats()["kernel_launches"] - launches0
    tokens_out = eng.download()
    assert all(len(o) == len(p) + NEW_TOKENS for o, p in zip(tokens_out, prompts))

    # ---- end to end: host prompt buffers -> host token buffers through the public call ----
    st0 = eng.stats()
    barrier()
    e0 = time.perf_counter()
    for i in range(args.steps):
        out = eng.generate(prompts, NEW_TOKENS, TEMPERATURE, TOP_K, eos_id=-1, seed=i, seq_index_base=rank * BATCH, as_arrays=True)
        if world > 1:
            mg.gather_token_lists(out, BATCH * world, as_arrays=True, max_len=max(prompt_lens) + NEW_TOKENS)   # the only exchange of the path
    barrier()
    e2e_s = time.perf_counter() - e0
    st1 = eng.stats()

    t_dev = torch.tensor([tot_ms, dec_ms, e2e_s * 1e3], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    tot_ms, dec_ms, e2e_ms = [float(x) for x in t_dev.tolist()]

    tokens_per_step = BATCH * NEW_TOKENS * world
    value = tokens_per_step * args.steps / (tot_ms * 1e-3)
    e2e_value = tokens_per_step * args.steps / (e2e_ms * 1e-3)
    alg_bytes, W_bytes, kappa = algorithmic_bytes(geo, prompt_lens, NEW_TOKENS)
    achieved = alg_bytes * args.steps / (dec_ms * 1e-3) / 1e9     # GB/s per GPU over the decode loop
    line = {
        "metric": "midi_decode_tokens_per_sec", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": config_dict(geo, prompt_lens, world),
        "e2e": {"value": e2e_value, "unit": "tokens/s",
                "h2d_bytes_per_step": (st1["h2d_bytes"] - st0["h2d_bytes"]) // args.steps,
                "d2h_bytes_per_step": (st1["d2h_bytes"] - st0["d2h_bytes"]) // args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": ncu_traffic_bytes(), "traffic_note": "DRAM read+write bytes per launch (= per job), ncu --set full, profiles/r*_decode_mega_full_raw.csv (newest)",
                     "peak_source": peak_src,
                     "kernel": ("decode_mega_kernel: ONE persistent cluster launch per job runs all 1024 decode steps "
                                "(embedding, 4 blocks, head, top-k sampler); achieved = algorithmic bytes of the decode "
                                "steps / that launch's duration (CUDA events on the engine stream)"),
                     "algorithmic_bytes_per_job": alg_bytes, "weight_bytes_per_step": W_bytes, "kv_bytes_per_position": kappa,
                     "frac_of_nominal_8TBs": achieved / 8000.0, "decode_ms_per_step": dec_ms / max(dec_steps, 1)},
        "wall_s_timed_region": wall,
    }

    secondary = {}
    if not args.no_extras:
        if world == 1:
            line["batch1"] = batch1_latency(mg)
            line["classifier"] = classifier_throughput(mg, tf_peak, peak_src)
            line["long_context"] = long_context(mg, hbm_peak)
            line["continuous_batching"] = continuous_batching(mg)
            line["production_geometry"] = production_geometry(mg, hbm_peak)
            secondary.update({"batch1_p50_ms_per_token_bf16": line["batch1"]["bf16"]["p50_ms_per_token"],
                              "batch1_p50_ms_per_token_fp32": line["batch1"]["fp32"]["p50_ms_per_token"],
                              "classifier_ms": line["classifier"]["ms"], "classifier_tflops": line["classifier"]["tflops"],
                              "classifier_frac_of_tensor_peak": line["classifier"]["frac_of_tensor_peak"],
                              "long_context_us_per_step": line["long_context"]["decode_us_per_step"],
                              "long_context_frac_of_measured_hbm": line["long_context"]["frac_of_measured_hbm"],
                              "continuous_batching_tokens_per_s": line["continuous_batching"]["continuous_tokens_per_s"],
                              "static_batching_tokens_per_s": line["continuous_batching"]["static_batches_tokens_per_s"],
                              "train_large2_us_per_step": line["production_geometry"]["decode_us_per_step"],
                              "train_large2_frac_of_measured_hbm": line["production_geometry"]["frac_of_measured_hbm"]})
        eng.close()
        eng = None
        pipe = pipeline_512(mg, rank, world, local_rank, dist)          # every rank: 512 / N requests (strong scaling)
        if rank == 0:
            line["pipeline"] = pipe
            secondary.update({"pipeline_requests": pipe["requests"], "pipeline_wall_s": pipe["wall_s"],
                              "pipeline_tokens_per_s": pipe["tokens_per_s"]})
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        n_new = 1024
        v, dt = cpu_reference_sample(ck, geo, prompts[0], n_new, threads)
        bt0 = time.perf_counter()
        from oracle import gpt_kv
        eq = [p for p in prompts if len(p) == 5][:16]
        gpt_kv.batched_decode_step_time_port(gpt_kv.KVModelOracle(mg.remap_state_dict(ck["model"]), geo.n_head), eq, 96,
                                             TEMPERATURE, TOP_K, torch.Generator().manual_seed(0))
        batched = len(eq) * 96 / (time.perf_counter() - bt0)
        line["cpu_baseline"] = {
            "value": v, "unit": "tokens/s", "cores": threads, "kind": "port",
            "sample": (f"oracle port of the reference batch-1 sample_kvcache loop on 1 of the 64 prompts, all {n_new} new "
                       f"tokens, top-k 40, fp32 torch CPU ({dt:.1f} s); batch-1 only: a 64-prompt job = 64 such runs"),
            "batched_restatement_tokens_per_s": batched,
            "batched_restatement_note": (f"NOT the reference's loop: GPTWithKV.forward fed idx [{len(eq)},1], .item() stop "
                                         "removed (SURVEY 8d), 96 steps"),
            "config1": config1_cpu(mg, threads), "distilbert_cpu": distilbert_cpu(mg, threads)}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0 and secondary:
        line["roofline"]["secondary"] = secondary                      # compact mirror of the side measurements
    if rank == 0:
        print(json.dumps(line), flush=True)
    if eng is not None:
        eng.close()
    if world > 1:
        dist.destroy_process_group()


def batch1_latency(mg):
    """BASELINE metric part 2: p50 ms/token @ batch 1 (config 1 shape: train_mini, 5-token prompt, 507 new tokens, greedy).
    p50 over the PER-TOKEN latencies (device %globaltimer stamp of every token, mg_last_step_times), all timed runs pooled."""
    geo = mg.GEOMETRIES["train_mini"]
    ck = mg.make_checkpoint(geo, 0)
    prompt = mg.encode(ck["vocab"], mg.synthetic_prompts(ck["vocab"], 1, seed=1)[0])
    out = {}
    for dtype in ("fp32", "bf16"):
        eng = mg.Generator(ck["model"], n_head=geo.n_head, dtype=dtype, max_batch=1, max_seq=1088)
        per_tok, run_means = [], []
        n_new = 512 - len(prompt)
        for i in range(7):
            eng.upload([prompt], n_new)
            eng.run(1.0, 1, eos_id=-1)
            eng.synchronize()
            t = eng.last_timing()
            if i >= 2:
                per_tok.append(eng.last_step_times_us())
                run_means.append(t["decode_ms"] / max(t["steps"], 1))
        lat = np.concatenate(per_tok) * 1e-3
        out[dtype] = {"p50_ms_per_token": float(np.percentile(lat, 50)), "p99_ms_per_token": float(np.percentile(lat, 99)),
                      "mean_ms_per_token": float(statistics.mean(run_means)), "tokens_timed": int(lat.size), "new_tokens": n_new,
                      "path": eng.last_decode_path(),
                      "note": "percentiles over per-token latencies (device-stamped), 5 runs pooled; greedy (top_k=1)"}
        eng.close()
    return out


def _with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def long_context(mg, hbm_peak):
    """BASELINE config 4: 256-token prompt prefill + 4096 new tokens, batch 16 (train_large blocks, 512-row position table).
    Default path = the grid kernel (all SMs, every (sequence, head) split over key ranges on different SMs); the cluster kernel
    (16 clusters x 4 CTAs = 64 SMs) is timed beside it on the same box."""
    geo = mg.GEOMETRIES["train_large_pos512"]
    ck = mg.make_checkpoint(geo, 0)
    rng = np.random.default_rng(0)
    prompts = [rng.integers(0, geo.vocab_size, 256).tolist() for _ in range(16)]
    alg, _, _ = algorithmic_bytes(geo, [256] * 16, 4096)

    def run(reps):
        eng = mg.Generator(ck["model"], n_head=geo.n_head, dtype="bf16", max_batch=16, max_seq=4352)
        best = None
        for i in range(reps):
            eng.upload(prompts, 4096)
            eng.run(1.0, 40, eos_id=-1, seed=i)
            eng.synchronize()
            t = eng.last_timing()
            best = t if best is None or t["total_ms"] < best["total_ms"] else best
        path = eng.last_decode_path()
        eng.close()
        return best, path

    best, path = run(2)                                                # default policy: few sequences, long caches -> grid kernel
    cbest, cpath = _with_env({"MG_GRID": "0"}, lambda: run(2))
    return {"workload": "config4: 256-token prompt + 4096 new tokens, batch 16, bf16", "path": path,
            "tokens_per_s": 16 * 4096 / (best["total_ms"] * 1e-3),
            "prefill_ms": best["prefill_ms"], "decode_us_per_step": 1e3 * best["decode_ms"] / 4096,
            "hbm_gbs": alg / (best["decode_ms"] * 1e-3) / 1e9, "frac_of_measured_hbm": alg / (best["decode_ms"] * 1e-3) / 1e9 / hbm_peak,
            "note": "grid-synchronous kernel on all 148 SMs: attention split over key ranges on different SMs, one grid barrier per phase",
            "cluster_kernel": {"path": cpath, "decode_us_per_step": 1e3 * cbest["decode_ms"] / 4096,
                               "frac_of_measured_hbm": alg / (cbest["decode_ms"] * 1e-3) / 1e9 / hbm_peak,
                               "note": "MG_GRID=0: 16 sequences -> 16 clusters x 4 CTAs = 64 of 148 SMs busy (one sequence per cluster)"}}


def pipeline_512(mg, rank, world, local_rank, dist):
    """BASELINE config 5: 512 requests classify -> emotion -> music parameters -> prompt -> generate (1024 tokens total each,
    top-k 40), batch-sharded over the replicas: every rank serves 512 / N requests (strong scaling); wall time = max over ranks.
    The emotion -> parameter mapping is the reference's own (EATS.py table via tests/golden/eats_table.json, random.seed(0))."""
    import random
    geo = mg.GEOMETRIES["train_large_pos512"]
    ck = mg.make_checkpoint(geo, 0)
    vocab = {t: i for t, i in ck["vocab"].items() if t != "[END_SEQUENCE]"}       # EOS disabled: every request runs to max_len
    vocab["[EOS_DISABLED]"] = ck["vocab"]["[END_SEQUENCE]"]
    gen = mg.Generator(ck["model"], n_head=geo.n_head, dtype="bf16", max_batch=128, max_seq=1088, device=local_rank)
    clf = mg.Classifier(mg.make_bert_state_dict(mg.DISTILBERT_BASE, 0), n_heads=12, max_tokens=16384, device=local_rank)
    params_fn = mg.eats_music_params(mg.load_eats_table(os.path.join(ROOT, "tests", "golden", "eats_table.json")))
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(1000, 30000, (512, 64), generator=g)
    ids[:, 0], ids[:, 63] = 101, 102
    lo, hi = mg.shard_range(512, rank, world)
    mine = ids.numpy()[lo:hi]
    random.seed(0)
    mg.classify_prompt_generate(clf, gen, vocab, mine[:min(128, len(mine))], params_fn=params_fn, max_len=64, top_k=40)   # warm-up
    random.seed(0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = mg.classify_prompt_generate(clf, gen, vocab, mine, params_fn=params_fn, max_len=1024, temperature=1.0, top_k=40,
                                      batch=128, seq_index_base=lo)
    dt = time.perf_counter() - t0
    assert len(out) == hi - lo and all(len(o) == 1024 for o in out)
    prompt_toks = sum(len(o) for o in out)
    t_dev = torch.tensor([dt, float(prompt_toks)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        t_max = t_dev.clone()
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_dev, op=dist.ReduceOp.SUM)
        dt, total = float(t_max[0]), float(t_dev[1])
    else:
        total = float(prompt_toks)
    gen.close()
    clf.close()
    return {"workload": "config5: 512 requests classify -> EATS mapping -> prompt -> generate to 1024 tokens, top-k 40",
            "requests": 512, "replicas": world, "requests_per_replica": hi - lo, "wall_s": dt, "requests_per_s": 512 / dt,
            "tokens_per_s": total / dt, "scaling": "strong",
            "note": "host wall clock, max over ranks: tokenised-text H2D, classifier, reference EATS table, prompt building, "
                    "generation in batches of 128, D2H of tokens (prompt tokens included in the count)"}


def production_geometry(mg, hbm_peak):
    """The geometry the paper's production model was trained with (train/train_large2.py:10-15: d 512, 8 heads of 64, 6 layers,
    511 position rows, V 8324), batch 64, generation to max_len = SEQ_LEN like the service call (api_cache.py:204).  The cluster
    kernel does not take it (d_model 256 only); default path = the grid-synchronous persistent kernel (decode_grid.cu), with the
    step graph (46 launches per step, batched decode GEMMs on tcgen05) timed beside it on the same box."""
    geo = mg.GEOMETRIES["train_large2"]
    ck = mg.make_checkpoint(geo, 0)
    prompts = [mg.encode(ck["vocab"], p) for p in mg.synthetic_prompts(ck["vocab"], BATCH, seed=0)]
    n_new = geo.pos_rows - max(len(p) for p in prompts)

    def run():
        eng = mg.Generator(ck["model"], n_head=geo.n_head, dtype="bf16", max_batch=BATCH, max_seq=544)
        best = None
        for i in range(3):
            eng.upload(prompts, n_new)
            eng.run(TEMPERATURE, TOP_K, eos_id=-1, seed=i)
            eng.synchronize()
            t = eng.last_timing()
            best = t if best is None or t["decode_ms"] < best["decode_ms"] else best
        path = eng.last_decode_path()
        eng.close()
        return best, path

    best, path = run()
    sbest, spath = _with_env({"MG_GRID": "0"}, run)
    alg, W, kappa = algorithmic_bytes(geo, [len(p) for p in prompts], n_new)
    gbs = alg / (best["decode_ms"] * 1e-3) / 1e9
    return {"workload": f"train_large2 geometry (d 512, L 6, hd 64), batch 64, {n_new} new tokens, top-k 40, bf16", "path": path,
            "tokens_per_s": BATCH * n_new / (best["total_ms"] * 1e-3), "decode_us_per_step": 1e3 * best["decode_ms"] / n_new,
            "hbm_gbs": gbs, "frac_of_measured_hbm": gbs / hbm_peak, "weight_bytes_per_step": W, "kv_bytes_per_position": kappa,
            "step_graph": {"path": spath, "decode_us_per_step": 1e3 * sbest["decode_ms"] / n_new,
                           "frac_of_measured_hbm": alg / (sbest["decode_ms"] * 1e-3) / 1e9 / hbm_peak}}


def continuous_batching(mg):
    """a stream of 640 requests with ragged budgets (32..512 new tokens) through a 64-slot session (chunks of 32
    decode steps, finished slots refilled between chunks) against static batches of 64 that run to their longest member."""
    geo = mg.GEOMETRIES[GEOMETRY]
    ck = mg.make_checkpoint(geo, 0)
    n_req = 640
    prompts = [mg.encode(ck["vocab"], p) for p in mg.synthetic_prompts(ck["vocab"], n_req, seed=3)]
    rng = np.random.default_rng(0)
    budgets = rng.integers(32, 513, n_req).tolist()
    total_new = int(sum(budgets))
    eng = mg.Generator(ck["model"], n_head=geo.n_head, dtype="bf16", max_batch=BATCH, max_seq=1088)
    out = {"workload": f"{n_req} requests, 32..512 new tokens each (uniform, seed 0), top-k 40, train_large bf16, one GPU"}
    for rep in range(2):                                              # first pass = warm-up
        t0 = time.perf_counter()
        for lo in range(0, n_req, BATCH):
            eng.generate(prompts[lo:lo + BATCH], budgets[lo:lo + BATCH], TEMPERATURE, TOP_K, eos_id=-1, seed=rep, as_arrays=True)
        t_static = time.perf_counter() - t0
        t0 = time.perf_counter()
        eng.slots_begin(BATCH, 1088, TEMPERATURE, TOP_K, eos_id=-1, seed=rep)
        nxt, slot_req, done, chunks = 0, [None] * BATCH, 0, 0
        while done < n_req:
            free = [b for b in range(BATCH) if slot_req[b] is None]
            take = list(range(nxt, min(nxt + len(free), n_req)))
            if take:
                eng.slots_admit(free[:len(take)], [prompts[r] for r in take], [budgets[r] for r in take], take)
                for b, r in zip(free, take):
                    slot_req[b] = r
                nxt += len(take)
            fin, _ = eng.slots_step(32)
            chunks += 1
            fin_slots = [b for b in range(BATCH) if fin[b] and slot_req[b] is not None]
            for b, row in zip(fin_slots, eng.slots_fetch_many(fin_slots, as_arrays=True)):
                assert len(row) == len(prompts[slot_req[b]]) + budgets[slot_req[b]]
                slot_req[b], done = None, done + 1
        eng.slots_end()
        t_cont = time.perf_counter() - t0
    out.update({"static_batches_tokens_per_s": total_new / t_static, "continuous_tokens_per_s": total_new / t_cont,
                "speedup": t_static / t_cont, "chunks": chunks, "slots": BATCH, "chunk_steps": 32,
                "note": "host wall clock incl. admission prefills, per-chunk status D2H and per-request token D2H"})
    eng.close()
    return out


def config1_cpu(mg, threads):
    """BASELINE config 1 on the host cores: train_mini, greedy, 507 new tokens from one 5-token prompt, fp32 -- BOTH reference
    loops (KV-cache sample_kvcache, api_cache.py:159-184; no-cache sample, generate.py:46-61) as oracle ports, at all host
    threads and at 1 thread (the no-cache loop is O(T^2): bounded to 192 new tokens at 1 thread)."""
    from oracle import gpt_kv, gpt_nocache
    geo = mg.GEOMETRIES["train_mini"]
    ck = mg.make_checkpoint(geo, 0)
    prompt = mg.encode(ck["vocab"], mg.synthetic_prompts(ck["vocab"], 1, seed=1)[0])
    sd = mg.remap_state_dict(ck["model"])
    kv, nc = gpt_kv.KVModelOracle(sd, geo.n_head), gpt_nocache.NoCacheModelOracle(sd, geo.n_head)
    out = {"workload": "config1: train_mini greedy, 5-token prompt, fp32, batch 1, CPU"}
    for th in (threads, 1):
        torch.set_num_threads(th)
        n_kv, n_nc = 512 - len(prompt), (512 - len(prompt) if th > 1 else 192)
        t0 = time.perf_counter()
        gpt_kv.sample_ids(kv, prompt, max_len=len(prompt) + n_kv, temperature=1.0, top_k=1)
        t_kv = time.perf_counter() - t0
        t0 = time.perf_counter()
        gpt_nocache.sample_ids(nc, prompt, max_len=len(prompt) + n_nc, temperature=1.0, top_k=1)
        t_nc = time.perf_counter() - t0
        out[f"threads_{th}"] = {"kv_loop_ms_per_token": 1e3 * t_kv / n_kv, "kv_new_tokens": n_kv,
                                "nocache_loop_ms_per_token": 1e3 * t_nc / n_nc, "nocache_new_tokens": n_nc}
    torch.set_num_threads(threads)
    return out


def distilbert_cpu(mg, threads):
    """BASELINE config 2 on the host cores: the installed HF DistilBertForSequenceClassification (the third-party class the
    reference calls, emotion_analysis/modeling.py:14-21), fp32, ids [256, 64], random-init weights; None if transformers is absent."""
    try:
        from transformers import DistilBertConfig, DistilBertForSequenceClassification
    except Exception:
        return None
    geo = mg.DISTILBERT_BASE
    torch.set_num_threads(threads)
    cfg = DistilBertConfig(vocab_size=geo.vocab_size, max_position_embeddings=geo.max_pos, dim=geo.dim, n_heads=geo.n_heads,
                           n_layers=geo.n_layers, hidden_dim=geo.hidden_dim, num_labels=geo.num_labels)
    torch.manual_seed(0)
    model = DistilBertForSequenceClassification(cfg).eval()
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(1000, 30000, (256, 64), generator=g)
    ids[:, 0], ids[:, 63] = 101, 102
    with torch.no_grad():
        model(input_ids=ids[:32])
        t0 = time.perf_counter()
        model(input_ids=ids)
        dt = time.perf_counter() - t0
    return {"workload": "config2: HF DistilBertForSequenceClassification fp32, 256 x 64, torch CPU", "seconds": dt,
            "texts_per_s": 256 / dt, "threads": threads}


def classifier_throughput(mg, tf_peak, peak_src):
    """BASELINE config 2: DistilBERT-base, 256 synthetic texts x 64 tokens, bf16 (device time, ids resident)."""
    geo = mg.DISTILBERT_BASE
    sd = mg.make_bert_state_dict(geo, 0)
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(1000, 30000, (256, 64), generator=g)
    ids[:, 0], ids[:, 63] = 101, 102
    clf = mg.Classifier(sd, n_heads=12, max_tokens=16384)
    ids_np = ids.numpy()
    clf.upload(ids_np)
    for _ in range(3):
        clf.run()
    clf.synchronize()
    times = []
    for _ in range(10):
        clf.upload(ids_np)
        clf.synchronize()
        t0 = time.perf_counter()
        clf.run()
        clf.synchronize()
        times.append(time.perf_counter() - t0)
    t = statistics.median(times)
    e0 = time.perf_counter()
    for _ in range(5):
        clf.classify(ids_np)
    e2e = (time.perf_counter() - e0) / 5
    flop = 16384 * 6 * (4 * 2 * 768 * 768 + 2 * 2 * 768 * 3072 + 4 * 64 * 768) + 256 * 2 * (768 * 768 + 768 * 28)
    clf.close()
    return {"workload": "config2: DistilBERT-base, 256 texts x 64 tokens, bf16", "texts_per_s": 256 / t, "ms": t * 1e3,
            "e2e_texts_per_s": 256 / e2e, "tflops": flop / t / 1e12, "tensor_peak_tflops": tf_peak,
            "frac_of_tensor_peak": flop / t / 1e12 / tf_peak, "peak_source": peak_src,
            "tensor_peak_sustained_tflops": sustained_tflops(),
            "frac_of_sustained_tensor_peak": (flop / t / 1e12 / sustained_tflops()) if sustained_tflops() else None,
            "peak_note": ("frac_of_tensor_peak is against the BURST cuBLAS figure; the pass is 50 back-to-back kernels, for "
                          "which the driver's sustained figure (SM clocks drop under tensor load) is the fairer denominator"),
            "timing": "host wall clock around run + stream sync, median of 10"}


if __name__ == "__main__":
    main()
