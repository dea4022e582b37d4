# Same-box A/B of prebuilt libraries: bash tools/ab_libs.sh tools/ab_libs/a.so tools/ab_libs/b.so ...   (last one stays installed)
cd "$(dirname "$0")/.."
LIB=music-generation-emotion-adaptive_b200/libmgea_b200.so
for rep in 1 2 3; do
  for lib in "$@"; do
    cp "$lib" $LIB
    echo "[$lib] c3: $(timeout 120 python tools/profile_step.py 1024 64 2>&1 | grep profile_step | cut -c1-120) | c4 $(timeout 120 python tools/profile_long.py 4096 2>&1 | grep profile_long | sed 's/.*us.step/us\/step/')"
  done
done
