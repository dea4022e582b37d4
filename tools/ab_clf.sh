# Same-box A/B of prebuilt libraries on the classifier pass: bash tools/ab_clf.sh a.so b.so ...   (the installed library is restored)
cd "$(dirname "$0")/.."
LIB=music-generation-emotion-adaptive_b200/libmgea_b200.so
cp $LIB /tmp/lib_keep.so
for rep in 1 2 3; do
  for lib in "$@"; do
    cp "$lib" $LIB
    echo "[$lib] $(timeout 120 python tools/profile_classifier.py 200 2>&1 | grep 'classifier pass' | cut -c1-70)"
  done
done
cp /tmp/lib_keep.so $LIB
