set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 100 python tests/scripts/grid_stress.py 2>&1 | tail -2
export PROBE_PARITY=0
timeout -s KILL 300 python tests/scripts/perf_probe.py 384 20 "" 2>&1 | grep p50
timeout -s KILL 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?
tail -3 gpurun_out/bench_n1.err
