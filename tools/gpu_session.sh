set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tail -1
( time timeout -s KILL 900 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/gputests.log 2>&1; tail -25 gpurun_out/gputests.log
