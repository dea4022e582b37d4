"""Per-kernel counts of the SASS mnemonics that tell Blackwell-native code from recompiled legacy code
(B200_PROFILING.md): UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG (TMA tensor copies), UBLKCP (cp.async.bulk),
HMMA (mma.sync), LDSM (ldmatrix), SYNCS (mbarrier), LDG/STG.

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "music-generation-emotion-adaptive_b200", "libmgea_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
pat = {"UTC*MMA": r"\bUTC[A-Z]*MMA", "LDTM": r"\bLDTM", "STTM": r"\bSTTM", "UTMALDG": r"\bUTMALDG", "UTMASTG": r"\bUTMASTG",
       "UBLKCP": r"\bUBLKCP", "UTCBAR": r"\bUTCBAR", "HMMA": r"\bHMMA", "LDSM": r"\bLDSM", "SYNCS": r"\bSYNCS", "LDG": r"\bLDG", "STG": r"\bSTG",
       "LDS": r"\bLDS\b|\bLDS\.", "STS": r"\bSTS\b|\bSTS\."}
counts, order, cur = collections.defaultdict(collections.Counter), [], None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "").replace("void ", "", 1)
        cur = re.sub(r"\(.*", "", cur)
        order.append(cur)
        continue
    if cur is None:
        continue
    for k, p in pat.items():
        if re.search(p, line):
            counts[cur][k] += 1
print("# cuobjdump -sass libmgea_b200.so (sm_100a): instruction counts per kernel (static, not executed counts)")
print(f"{'kernel':70s} " + " ".join(f"{k:>8s}" for k in pat))
for name in order:
    c = counts[name]
    print(f"{name[:70]:70s} " + " ".join(f"{c[k]:8d}" for k in pat))
