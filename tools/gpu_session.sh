set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export PROBE_PARITY=0
timeout -s KILL 300 python tests/scripts/perf_probe.py 384 20 "C" 2>&1 | grep p50
echo "== e2e C4"; timeout -s KILL 400 python tests/scripts/open_probe.py 480 2>&1 | grep -E "step|table open|\+" | tail -14
timeout -s KILL 900 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-cpu > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; echo rc=$?
tail -3 gpurun_out/bench_n1b.err
