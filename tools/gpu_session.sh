set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -60 > gpurun_out/pytest.log; tail -5 gpurun_out/pytest.log
for cfg in "8 0" "8 1" "6 1" "4 1" "4 0" "2 1" "0 1"; do set -- $cfg; echo "== SMEM_SHARE=$1 F64_GLOBAL=$2"; PROBE_PARITY=0 PQB_SMEM_SHARE=$1 PQB_F64_GLOBAL=$2 timeout -s KILL 200 python tests/scripts/perf_probe.py 96 20 "GROUP BY host" 2>&1 | grep "p50"; done > gpurun_out/probe_sweep.log 2>&1
cat gpurun_out/probe_sweep.log
PROBE_PARITY=0 timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_flat_agg -s 4 -c 1 -o gpurun_out/prof_agg_c4 python tests/scripts/perf_probe.py 96 3 "C4" > gpurun_out/ncu_agg.log 2>&1
tail -3 gpurun_out/ncu_agg.log
