cd "$(dirname "$0")/.."
LIB=music-generation-emotion-adaptive_b200/libmgea_b200.so
cp $LIB /tmp/lib_new.so
timeout 120 python tools/mega_check.py train_large 3 6 2>&1 | tail -3; timeout 120 python tools/mega_check.py train_mini 2 4 2>&1 | tail -2
timeout 400 python -m pytest tests -x -q -m gpu -k "persistent or long_caches or config3" 2>&1 | tail -2
for rep in 1 2; do
  cp tools/lib_before.so $LIB; echo "[before] $(MG_MEGA_PROF_STEP=40 timeout 100 python tools/profile_step.py 1024 64 2>&1 | grep 'prof\] step\|profile_step' | sed 's/.mega prof. step 40 .ns since first stamp.://' | cut -c1-190 | tr '\n' ' ')"
  cp /tmp/lib_new.so $LIB; echo "[after ] $(MG_MEGA_PROF_STEP=40 timeout 100 python tools/profile_step.py 1024 64 2>&1 | grep 'prof\] step\|profile_step' | sed 's/.mega prof. step 40 .ns since first stamp.://' | cut -c1-190 | tr '\n' ' ')"
done
cp /tmp/lib_new.so $LIB
