python tools/mega_check.py train_large 3 6 2>&1 | tail -4
python tools/mega_check.py train_mini 2 4 2>&1 | tail -3
MG_MEGA_PROF_STEP=500 python tools/profile_step.py 1024 64 2>&1 | grep "mega prof\|profile_step"
MG_MEGA_PROF_STEP=1000 python tools/profile_step.py 1024 64 2>&1 | grep "mega prof"
