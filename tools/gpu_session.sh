set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -60 > gpurun_out/pytest.log; tail -30 gpurun_out/pytest.log
timeout -s KILL 300 python tests/scripts/open_probe.py 480 > gpurun_out/open_gather.log 2>&1; grep -v "^\[pqb\] +\|k_flat\|k_scan" gpurun_out/open_gather.log | tail -20
PQB_UPLOAD=memcpy timeout -s KILL 300 python tests/scripts/open_probe.py 480 > gpurun_out/open_memcpy.log 2>&1; grep "^step" gpurun_out/open_memcpy.log
