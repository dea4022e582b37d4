"""p50 ms/token at batch 1 (config 1 shape: train_mini, greedy, 507 new tokens) through the bf16 persistent kernel."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mgea_b200 as mg
geo = mg.GEOMETRIES["train_mini"]
ck = mg.make_checkpoint(geo, 0)
prompt = mg.encode(ck["vocab"], mg.synthetic_prompts(ck["vocab"], 1, seed=1)[0])
eng = mg.Generator(ck["model"], n_head=geo.n_head, dtype="bf16", max_batch=1, max_seq=1088)
n_new = 512 - len(prompt)
per_tok = []
for i in range(7):
    eng.upload([prompt], n_new); eng.run(1.0, int(sys.argv[1]) if len(sys.argv) > 1 else 1, eos_id=-1); eng.synchronize()
    t = eng.last_timing()
    if i >= 2: per_tok.append(t["decode_ms"] / t["steps"])
print("batch-1 bf16 p50 ms/token %.5f" % statistics.median(per_tok))
