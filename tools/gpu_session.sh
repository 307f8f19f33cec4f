set -x
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?
tail -3 gpurun_out/bench_n1.err
timeout -s KILL 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?; cut -c1-300 gpurun_out/bench_ref.json
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r2g.csv python bench.py --steps 3 --warmup 3 --skip-cpu > gpurun_out/bench_ncu.json 2> gpurun_out/bench_ncu.err; echo rc=$?
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_flat_agg -s 1 -c 1 -o gpurun_out/prof_agg_bench_r2g python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --skip-c2 > /dev/null 2> gpurun_out/ncu_agg.log; tail -2 gpurun_out/ncu_agg.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_flat_filter -s 4 -c 1 -o gpurun_out/prof_filter_bench_r2g python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e > /dev/null 2> gpurun_out/ncu_filter.log; tail -2 gpurun_out/ncu_filter.log
export PROBE_PARITY=0
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_flat_agg -s 4 -c 1 -o gpurun_out/prof_agg_c4_r2g python tests/scripts/perf_probe.py 384 3 "C4" > gpurun_out/ncu_agg2.log 2>&1; tail -2 gpurun_out/ncu_agg2.log
timeout -s KILL 300 python tests/scripts/perf_probe.py 384 20 "" 2>&1 | grep p50
