# A/B timing of kernel build variants on ONE box: bash tools/ab.sh "<EXTRA flags A>" "<EXTRA flags B>" ...
cd "$(dirname "$0")/.."
LIB=music-generation-emotion-adaptive_b200/libmgea_b200.so
i=0
for flags in "$@"; do
  touch music-generation-emotion-adaptive_b200/csrc/decode_mega.cu
  make -C music-generation-emotion-adaptive_b200/csrc EXTRA="$flags" -j8 > /dev/null 2>&1 || echo "build failed: $flags"
  cp $LIB /tmp/lib_$i.so; i=$((i+1))
done
cp /tmp/lib_0.so $LIB; timeout 120 python tools/mega_check.py train_large 3 6 2>&1 | tail -3; timeout 120 python tools/mega_check.py train_mini 2 4 2>&1 | tail -2
for rep in 1 2; do
  i=0
  for flags in "$@"; do
    cp /tmp/lib_$i.so $LIB; i=$((i+1))
    echo "[$flags] c3: $(timeout 120 python tools/profile_step.py 1024 64 2>&1 | grep 'profile_step' | cut -c20-110) | c4: $(timeout 120 python tools/profile_long.py 2048 2>&1 | grep profile_long | sed 's/.*us.step/us\/step/')"
  done
done
touch music-generation-emotion-adaptive_b200/csrc/decode_mega.cu
make -C music-generation-emotion-adaptive_b200/csrc -j8 > /dev/null 2>&1
