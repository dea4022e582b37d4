"""Timing of the persistent decode kernel at BASELINE config 3 (B=64, 1024 new tokens, top-k 40)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mgea_b200 as mg
new_tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
k = int(sys.argv[2]) if len(sys.argv) > 2 else 40
geo = mg.GEOMETRIES["train_large"]
ck = mg.make_checkpoint(geo, 0)
prompts = [mg.encode(ck["vocab"], p) for p in mg.synthetic_prompts(ck["vocab"], 64, seed=0)]
eng = mg.Generator(ck["model"], n_head=geo.n_head, dtype="bf16", max_batch=64, max_seq=1088)
for i in range(3):
    eng.upload(prompts, new_tokens); eng.run(1.0, k, eos_id=-1, seed=i); eng.synchronize()
    t = eng.last_timing()
    print("run", i, t, "us/step %.1f" % (1e3 * t["decode_ms"] / t["steps"]), "tok/s %.0f" % (64 * new_tokens / (t["total_ms"] * 1e-3)), flush=True)
out = eng.download()
assert all(len(o) == len(p) + new_tokens for o, p in zip(out, prompts))
