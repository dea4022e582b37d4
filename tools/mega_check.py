"""Debug driver for the persistent cluster decode kernel: teacher-forced logits vs the oracle, then timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mgea_b200 as mg
from oracle import gpt_kv

geo_name = sys.argv[1] if len(sys.argv) > 1 else "train_large"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = int(sys.argv[3]) if len(sys.argv) > 3 else 6
geo = mg.GEOMETRIES[geo_name]
ck = mg.make_checkpoint(geo, 0)
prompts = [mg.encode(ck["vocab"], p) for p in mg.synthetic_prompts(ck["vocab"], B, seed=3)]
forced = np.random.default_rng(0).integers(0, geo.vocab_size, (B, n)).astype(np.int32)
eng = mg.Generator(ck["model"], n_head=geo.n_head, dtype="bf16", max_batch=max(B, 64), max_seq=1088)
t0 = time.time()
lg = eng.step_logits(prompts, forced, n)
print("step_logits done in %.2fs" % (time.time() - t0), "path", eng.last_decode_path(), flush=True)
ora = gpt_kv.KVModelOracle(mg.remap_state_dict(ck["model"]), geo.n_head, torch.float64)
for b in range(min(B, 4)):
    want = gpt_kv.teacher_forced_logits(ora, prompts[b], forced[b].tolist(), n).numpy()
    err = np.abs(lg[:, b, :] - want).max(axis=1)
    print("seq", b, "max abs err per step", np.array2string(err, precision=4), flush=True)
if B >= 8:
    for k in (1, 40):
        out = eng.generate(prompts, 64, 1.0, k, seed=5)
        print("generate top_k", k, "ok", eng.last_decode_path(), [len(o) for o in out[:4]], eng.last_timing(), flush=True)
