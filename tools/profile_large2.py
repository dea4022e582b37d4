"""Short deterministic run of the train_large2 geometry (d 512, L 6, hd 64; grid kernel, MG_GRID=0: step graph) for ncu / timing.

    python tools/profile_large2.py [new_tokens] [batch]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mgea_b200 as mg  # noqa: E402

new_tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 64
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
geo = mg.GEOMETRIES["train_large2"]
ck = mg.make_checkpoint(geo, 0)
prompts = [mg.encode(ck["vocab"], p) for p in mg.synthetic_prompts(ck["vocab"], batch, seed=0)]
eng = mg.Generator(ck["model"], n_head=geo.n_head, dtype="bf16", max_batch=batch, max_seq=544)
for _ in range(2):
    eng.upload(prompts, new_tokens)
    eng.run(1.0, 40, eos_id=-1, seed=0)
    eng.synchronize()
t = eng.last_timing()
print("profile_large2 ok", eng.last_decode_path(), t, "us/step %.1f" % (1e3 * t["decode_ms"] / max(t["steps"], 1)), eng.stats())
