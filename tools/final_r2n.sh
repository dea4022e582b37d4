# Round-2 final evidence run (one gpurun call): smoke, bench + reference arm, ncu launch lists, full captures of the two persistent decode kernels.
set -x
cd "$(dirname "$0")/.."
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; tail -1 gpurun_out/r2n_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -2 gpurun_out/r2n_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2n_bench_ref.json 2>> gpurun_out/r2n_bench.err
# launch lists (serialised, cold cache: shares only): one job of the benchmarked workload, one of the train_large2 geometry
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches_persistent.csv python tools/profile_step.py 1024 64 > gpurun_out/r2n_ncu_launch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches_large2_grid.csv python tools/profile_large2.py 505 > gpurun_out/r2n_ncu_launch_l2.log 2>&1
# full captures
timeout 900 ncu --set full --import-source on --clock-control none -k regex:decode_mega --launch-skip 0 --launch-count 1 -o gpurun_out/r2n_mega_full python tools/profile_step.py 1024 64 > gpurun_out/r2n_ncu_full.log 2>&1
ncu -i gpurun_out/r2n_mega_full.ncu-rep --page raw --csv > gpurun_out/r2n_decode_mega_full_raw.csv 2>/dev/null
timeout 900 ncu --set full --import-source on --clock-control none -k regex:decode_grid --launch-skip 0 --launch-count 1 -o gpurun_out/r2n_grid_full python tools/profile_large2.py 505 > gpurun_out/r2n_ncu_full_grid.log 2>&1
ncu -i gpurun_out/r2n_grid_full.ncu-rep --page raw --csv > gpurun_out/r2n_decode_grid_full_raw.csv 2>/dev/null
timeout 300 env MG_BERT_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2n_launches_classifier.csv python tools/profile_classifier.py 2 > gpurun_out/r2n_ncu_clf.log 2>&1
ls -la gpurun_out | tail -14
