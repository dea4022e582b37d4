# fine-grained phase trace of the persistent kernel: rebuild with -DMG_MEGA_TRACE, run, restore the product build
# usage (on the GPU box): bash tools/trace.sh <step> [batch]
set -e
cd "$(dirname "$0")/.."
touch music-generation-emotion-adaptive_b200/csrc/decode_mega.cu
make -C music-generation-emotion-adaptive_b200/csrc EXTRA=-DMG_MEGA_TRACE -j8 > /dev/null
MG_MEGA_PROF_STEP=${1:-500} python tools/profile_step.py 1024 ${2:-64} 2>&1 | grep "mega prof"
touch music-generation-emotion-adaptive_b200/csrc/decode_mega.cu
make -C music-generation-emotion-adaptive_b200/csrc -j8 > /dev/null
