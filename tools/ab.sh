# A/B timing of kernel build variants on ONE box: bash tools/ab.sh "<EXTRA flags A>" "<EXTRA flags B>" ...
cd "$(dirname "$0")/.."
LIB=music-generation-emotion-adaptive_b200/libmgea_b200.so
i=0
for flags in "$@"; do
  touch music-generation-emotion-adaptive_b200/csrc/decode_mega.cu music-generation-emotion-adaptive_b200/csrc/engine.cu
  rm -rf music-generation-emotion-adaptive_b200/csrc/build; make -C music-generation-emotion-adaptive_b200/csrc EXTRA="$flags" -j8 > /dev/null 2>&1 || echo "build failed: $flags"
  cp $LIB /tmp/lib_$i.so; i=$((i+1))
done
cp /tmp/lib_0.so $LIB; timeout 120 python tools/mega_check.py train_large 3 6 2>&1 | tail -3; timeout 120 python tools/mega_check.py train_mini 2 4 2>&1 | tail -2
for rep in 1 2; do
  i=0
  for flags in "$@"; do
    cp /tmp/lib_$i.so $LIB; i=$((i+1))
    echo "[$flags] c3: $(MG_MEGA_PROF_STEP=40 timeout 120 python tools/profile_step.py 1024 64 2>&1 | grep 'prof\] step\|profile_step' | sed 's/.mega prof. step 40 .ns since first stamp.://' | cut -c1-200 | tr '\n' ' ') | c4 $(timeout 120 python tools/profile_long.py 1024 2>&1 | grep profile_long | sed 's/.*us.step/us\/step/')"
  done
done
rm -rf music-generation-emotion-adaptive_b200/csrc/build; make -C music-generation-emotion-adaptive_b200/csrc -j8 > /dev/null 2>&1
