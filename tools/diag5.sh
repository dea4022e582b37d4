python tools/mega_check.py train_large 3 6 2>&1 | tail -3
python tools/mega_check.py train_mini 2 4 2>&1 | tail -2
for st in 40 500 1000; do MG_MEGA_PROF_STEP=$st python tools/profile_step.py 1024 64 2>&1 | grep "prof\] step"; done
python tools/profile_step.py 1024 64 2>&1 | grep profile_step
