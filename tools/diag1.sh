set -x
python tools/profile_step.py 1024 64 2>&1 | tail -2
MG_MEGA_SKIP_LOADS=1 python tools/profile_step.py 1024 64 2>&1 | tail -1
MG_MEGA_ATTN_HOT=1 python tools/profile_step.py 1024 64 2>&1 | tail -1
MG_MEGA_ATTN_HOT=1 MG_MEGA_SKIP_LOADS=1 python tools/profile_step.py 1024 64 2>&1 | tail -1
python tools/profile_step.py 1024 32 2>&1 | tail -1
python tools/profile_step.py 1024 16 2>&1 | tail -1
python tools/profile_step.py 1024 2 2>&1 | tail -1
MG_MEGA_SKIP_LOADS=1 python tools/profile_step.py 1024 2 2>&1 | tail -1
MG_MEGA_PROF_STEP=500 python tools/profile_step.py 1024 64 2>&1 | tail -4
MG_MEGA_PROF_STEP=500 python tools/profile_step.py 1024 2 2>&1 | tail -4
MG_MEGA_PROF_STEP=500 MG_MEGA_SKIP_LOADS=1 python tools/profile_step.py 1024 64 2>&1 | tail -4
