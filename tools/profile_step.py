"""Short, deterministic run of the config-3 decode path for ncu (no timing is taken from it).

    python tools/profile_step.py [new_tokens] [batch]

Launch-list recipe (B200_PROFILING.md):  run it plain first, then
    ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <n> --csv --log-file gpurun_out/launches.csv \
        python tools/profile_step.py 600
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mgea_b200 as mg  # noqa: E402

new_tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 600
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
geo = mg.GEOMETRIES["train_large"]
ck = mg.make_checkpoint(geo, 0)
prompts = [mg.encode(ck["vocab"], p) for p in mg.synthetic_prompts(ck["vocab"], batch, seed=0)]
eng = mg.Generator(ck["model"], n_head=geo.n_head, dtype="bf16", max_batch=batch, max_seq=1088)
eng.upload(prompts, new_tokens)
eng.run(1.0, 40, eos_id=-1, seed=0)
eng.synchronize()
t = eng.last_timing()
out = eng.download()
assert all(len(o) == len(p) + new_tokens for o, p in zip(out, prompts))
print("profile_step ok", eng.last_decode_path(), t, eng.stats())
