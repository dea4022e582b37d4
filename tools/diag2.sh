set -x
MG_MEGA_DEBUG=1 python tools/mega_check.py train_large 3 6 2>&1 | tail -8
MG_MEGA_VARIANT=1 python tools/profile_step.py 1024 64 2>&1 | tail -1
MG_MEGA_VARIANT=0 python tools/profile_step.py 1024 64 2>&1 | tail -1
MG_MEGA_VARIANT=0 python tools/profile_step.py 1024 32 2>&1 | tail -1
MG_MEGA_VARIANT=0 python tools/profile_step.py 1024 2 2>&1 | tail -1
MG_MEGA_VARIANT=0 MG_MEGA_PROF_STEP=500 python tools/profile_step.py 1024 64 2>&1 | grep "mega prof\] step"
