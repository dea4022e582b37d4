# Round-2 evidence run (one gpurun call): bench + reference arm, then the ncu launch lists and the full capture of the dominant kernel.
set -x
cd "$(dirname "$0")/.."
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; tail -2 gpurun_out/r2m_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m_bench_ref.json 2>> gpurun_out/r2m_bench.err
# launch list of one job of the benchmarked workload (serialised, cold cache: shares only)
timeout 300 python tools/profile_step.py 1024 64 > gpurun_out/r2m_plain.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2m_launches_persistent.csv python tools/profile_step.py 1024 64 > gpurun_out/r2m_ncu_launch.log 2>&1
# full capture of the dominant kernel
timeout 900 ncu --set full --import-source on --clock-control none -k regex:decode_mega --launch-skip 0 --launch-count 1 -o gpurun_out/r2m_mega_full python tools/profile_step.py 1024 64 > gpurun_out/r2m_ncu_full.log 2>&1
ncu -i gpurun_out/r2m_mega_full.ncu-rep --page raw --csv > gpurun_out/r2m_decode_mega_full_raw.csv 2>/dev/null
# classifier pass, eager launches
timeout 300 env MG_BERT_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2m_launches_classifier.csv python tools/profile_classifier.py 2 > gpurun_out/r2m_ncu_clf.log 2>&1
ls -la gpurun_out | tail -12
