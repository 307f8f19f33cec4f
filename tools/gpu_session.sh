set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for t in 0 4 8 16; do echo "== LANE_SLOTS $t"; PQB_LANE_SLOTS=$t PROBE_PARITY=0 timeout -s KILL 300 python tests/scripts/perf_probe.py 384 20 "C" 2>&1 | grep p50 | grep -v C2; done
echo "== DIRECT8 off"; PQB_AGG_DIRECT8=0 PROBE_PARITY=0 timeout -s KILL 300 python tests/scripts/perf_probe.py 384 20 "C4" 2>&1 | grep p50
echo "== DIRECT8 off LANE 0"; PQB_LANE_SLOTS=0 PQB_AGG_DIRECT8=0 PROBE_PARITY=0 timeout -s KILL 300 python tests/scripts/perf_probe.py 384 20 "C4" 2>&1 | grep p50
PROBE_PARITY=0 timeout -s KILL 400 python tests/scripts/perf_probe.py 384 20 "" 2>&1 | grep "p50"
timeout -s KILL 100 python tests/scripts/grid_stress.py 2>&1 | tail -2
timeout -s KILL 400 python tests/scripts/c5_probe.py 384 20 2>&1 | tail -8
