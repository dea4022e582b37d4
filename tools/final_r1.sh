set -x
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_final.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
# launch list of one bench step (serialised, cold cache: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1j.csv python tools/profile_step.py 1024 64 > gpurun_out/ncu_launch.log 2>&1
# full capture of the dominant kernel
timeout 700 ncu --set full --import-source on --clock-control none -k regex:decode_mega -o gpurun_out/mega_final python tools/profile_step.py 1024 64 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/mega_final.ncu-rep --page raw --csv > gpurun_out/mega_final_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_classifier_r1j.csv python tools/profile_classifier.py 2 > gpurun_out/ncu_clf.log 2>&1
