set -x
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_flat_agg -s 1 -c 1 -o gpurun_out/prof_agg_bench_r2e python bench.py --steps 3 --warmup 3 --skip-cpu --skip-e2e --skip-c2 > /dev/null 2> gpurun_out/ncu_agg.log; tail -2 gpurun_out/ncu_agg.log
export PROBE_PARITY=0
P="timeout -s KILL 300 python tests/scripts/perf_probe.py 384 20"
echo "== agg default"; $P "C" 2>&1 | grep p50 | grep -v C2
echo "== agg NOWORK"; PQB_AGG_NOWORK=1 $P "C" 2>&1 | grep p50 | grep -v C2
