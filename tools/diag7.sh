run() { echo "groups=$1 ns=$2: $(MG_MEGA_STAGGER_GROUPS=$1 MG_MEGA_STAGGER_NS=$2 timeout 100 python tools/profile_step.py 1024 64 2>&1 | grep profile_step | cut -c1-120)"; }
run 0 0
run 2 25000
run 2 12000
run 4 12000
run 4 6000
run 8 6000
run 32 1500
run 0 0
