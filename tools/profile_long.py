"""Short run of BASELINE config 4 (256-token prompts + N new tokens, batch 16) through the persistent decode kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mgea_b200 as mg
new_tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
geo = mg.GEOMETRIES["train_large_pos512"]
ck = mg.make_checkpoint(geo, 0)
rng = np.random.default_rng(0)
prompts = [rng.integers(0, geo.vocab_size, 256).tolist() for _ in range(16)]
eng = mg.Generator(ck["model"], n_head=geo.n_head, dtype="bf16", max_batch=16, max_seq=4352)
eng.upload(prompts, new_tokens); eng.run(1.0, 40, eos_id=-1, seed=0); eng.synchronize()
t = eng.last_timing()
print("profile_long ok", t, "us/step %.1f" % (1e3 * t["decode_ms"] / t["steps"]))
