set -x
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -60 > gpurun_out/pytest.log; tail -30 gpurun_out/pytest.log
