# Same-box A/B of the K/V load flavours of the grid kernel (tools/ab_libs/kvld{0..3}.so): determinism probe + timing
cd "$(dirname "$0")/.."
LIB=music-generation-emotion-adaptive_b200/libmgea_b200.so
cp $LIB /tmp/lib_keep.so
for v in 0 1 2 3; do
  cp tools/ab_libs/kvld$v.so $LIB
  echo "== kvld$v (0 volatile, 1 ld.cg, 2 relaxed.gpu no_allocate, 3 weak no_allocate; appends by st.cg in all)"
  timeout 200 python tools/grid_det.py 2>&1 | tail -2
  GRID_ONLY=1 timeout 300 python tools/grid_ab.py c3,c4 1 2>&1 | tail -2
done
cp /tmp/lib_keep.so $LIB
