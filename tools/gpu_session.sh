set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout -s KILL 100 python tests/scripts/grid_stress.py 2>&1 | tail -2
P="timeout -s KILL 300 python tests/scripts/perf_probe.py 384 20"
export PROBE_PARITY=0
echo "== default (back-off)"; $P "C" 2>&1 | grep p50
echo "== hint wait"; PQB_LIB=$PWD/parseable_b200/libparseable_b200_hint.so $P "C" 2>&1 | grep p50
for c in 2 3 4 5; do echo "== CTAS $c"; PQB_FILTER_CTAS=$c $P "C2 filter -> row ids" 2>&1 | grep p50; done
echo "== rest"; $P "G" 2>&1 | grep p50; $P "global" 2>&1 | grep p50
echo "== e2e C4, LZ4_RAW files"; PQB_BENCH_CODEC=LZ4 timeout -s KILL 400 python tests/scripts/open_probe.py 480 2>&1 | grep -E "step|table open" | tail -9
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_flat_filter -s 4 -c 1 -o gpurun_out/prof_filter_r2d python tests/scripts/perf_probe.py 384 3 "C2 filter -> row ids" > gpurun_out/ncu_filter.log 2>&1; tail -2 gpurun_out/ncu_filter.log
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_flat_agg -s 4 -c 1 -o gpurun_out/prof_agg_c4_r2d python tests/scripts/perf_probe.py 384 3 "C4" > gpurun_out/ncu_agg.log 2>&1; tail -2 gpurun_out/ncu_agg.log
