set -x
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -60 > gpurun_out/pytest.log; tail -30 gpurun_out/pytest.log
PROBE_PARITY=1 timeout -s KILL 300 python tests/scripts/perf_probe.py 96 20 > gpurun_out/probe_default.log 2>&1
grep -v "^\[pqb\]" gpurun_out/probe_default.log | tail -12
timeout -s KILL 100 python tests/scripts/grid_stress.py > gpurun_out/grid.log 2>&1; tail -3 gpurun_out/grid.log
