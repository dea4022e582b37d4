MG_MEGA_PROF_STEP=500 python tools/profile_step.py 1024 64 2>&1 | grep "mega prof"
MG_MEGA_PROF_STEP=500 python tools/profile_step.py 1024 2 2>&1 | grep "mega prof"
MG_MEGA_PROF_STEP=501 MG_MEGA_SKIP_LOADS=1 python tools/profile_step.py 1024 64 2>&1 | grep "mega prof"
