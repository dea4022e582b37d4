"""One DistilBERT-base classify pass at BASELINE config 2 (256 texts x 64 tokens, bf16) for ncu / timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mgea_b200 as mg
geo = mg.DISTILBERT_BASE
sd = mg.make_bert_state_dict(geo, 0)
ids = torch.randint(1000, 30000, (256, 64), generator=torch.Generator().manual_seed(0))
ids[:, 0], ids[:, 63] = 101, 102
clf = mg.Classifier(sd, n_heads=12, max_tokens=16384)
clf.upload(ids.numpy())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ts = []
for _ in range(n):
    t0 = time.perf_counter(); clf.run(); clf.synchronize(); ts.append(time.perf_counter() - t0)
ts = sorted(ts[min(5, n - 1):])
print("classifier pass %.3f ms (median of %d; min %.3f)" % (ts[len(ts) // 2] * 1e3, len(ts), ts[0] * 1e3), clf.stats())
