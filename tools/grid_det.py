import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mgea_b200 as mg
geo = mg.GEOMETRIES["train_large"]
ck = mg.make_checkpoint(geo, 0)
prompts = [mg.encode(ck["vocab"], p) for p in mg.synthetic_prompts(ck["vocab"], 24, seed=4)]
os.environ["MG_GRID"] = "1"
e = mg.Generator(ck["model"], n_head=geo.n_head, dtype="bf16", max_batch=64, max_seq=320)
for topk in (40, 1):
    runs = [e.generate(prompts, 48, 1.0, topk, seed=7) for _ in range(6)]
    bad = [(i, b, next(j for j, (x, y) in enumerate(zip(runs[0][b], runs[i][b])) if x != y)) for i in range(1, 6) for b in range(24) if runs[i][b] != runs[0][b]]
    print("fence", os.environ.get("MG_GRID_FENCE"), "top_k", topk, "path", e.last_decode_path(), "mismatches (run, row, first pos):", bad[:12], flush=True)
