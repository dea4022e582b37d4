for m in 0 3; do echo "GEMM_DBG=$m"; MG_MEGA_GEMM_DBG=$m MG_MEGA_PROF_STEP=40 python tools/profile_step.py 1024 64 2>&1 | grep "prof\] step"; done
python tools/profile_step.py 1024 64 2>&1 | grep profile_step
