cd "$(dirname "$0")/.."
LIB=music-generation-emotion-adaptive_b200/libmgea_b200.so
cp tools/ab_libs/new.so $LIB
timeout 180 python -m pytest tests -x -q -m gpu -k "tc_gemm or classifier" 2>&1 | tail -5
for rep in 1 2 3; do
  cp tools/ab_libs/old.so $LIB
  echo "[old       ] $(timeout 120 python tools/profile_classifier.py 20 2>&1 | grep 'classifier pass\|rror' | cut -c1-40)"
  cp tools/ab_libs/new.so $LIB
  echo "[new, 1cta ] $(MG_GEMM_2CTA=0 timeout 120 python tools/profile_classifier.py 20 2>&1 | grep 'classifier pass\|rror' | cut -c1-40)"
  echo "[new, 2cta ] $(timeout 120 python tools/profile_classifier.py 20 2>&1 | grep 'classifier pass\|rror' | cut -c1-40)"
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/clf_new.csv python tools/profile_classifier.py 2 > gpurun_out/ncu_clf.log 2>&1
