"""Same-box A/B of the decode paths on the benchmarked shapes: cluster kernel / step graph (default) vs grid kernel (MG_GRID=1).
usage: python tools/grid_ab.py [cases: c3,c4,l2,b1] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mgea_b200 as mg

CASES = {
    "c3": ("train_large", 64, 6, 1024, 1088),          # BASELINE config 3
    "c4": ("train_large_pos512", 16, 256, 4096, 4352),  # BASELINE config 4
    "l2": ("train_large2", 64, 6, 512, 576),            # the paper's production geometry
    "b1": ("train_mini", 1, 6, 506, 512),               # batch-1 latency shape
    "c3b128": ("train_large", 128, 6, 1024, 1088),
}
want = (sys.argv[1] if len(sys.argv) > 1 else "c3,c4,l2,b1").split(",")
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for name in want:
    geo_name, B, tp, new, max_seq = CASES[name]
    geo = mg.GEOMETRIES[geo_name]
    ck = mg.make_checkpoint(geo, 0)
    rng = np.random.default_rng(0)
    prompts = [rng.integers(0, geo.vocab_size, tp).tolist() for _ in range(B)]
    envs = ({"MG_GRID": "0"}, {"MG_GRID": "1"})
    if os.environ.get("GRID_ONLY"): envs = ({"MG_GRID": "1"},)
    for env in envs:
        os.environ.update(env)
        eng = mg.Generator(ck["model"], n_head=geo.n_head, dtype="bf16", max_batch=B, max_seq=max_seq)
        best = None
        for r in range(reps + 1):
            eng.upload(prompts, new); eng.run(1.0, 40 if name != "b1" else 1, eos_id=-1, seed=r); eng.synchronize()
            t = eng.last_timing()
            us = 1e3 * t["decode_ms"] / t["steps"]
            if r > 0 or reps == 0: best = us if best is None else min(best, us)
        print(f"{name:7s} {geo_name:20s} B {B:3d} new {new:5d} MG_GRID={env['MG_GRID']} path {eng.last_decode_path():14s} "
              f"{best:8.2f} us/step  {B * 1e6 / best / 1e3:9.1f} k tokens/s", flush=True)
        eng.close()
