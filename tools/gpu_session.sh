set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 100 python tests/scripts/grid_stress.py 2>&1 | tail -2
echo "== e2e C4 (spans + DMA)"; timeout -s KILL 400 python tests/scripts/open_probe.py 480 2>&1 | grep -E "step|table open" | tail -9
echo "== e2e C4 (spans, gather kernel)"; PQB_UPLOAD=gather timeout -s KILL 400 python tests/scripts/open_probe.py 480 2>&1 | grep -E "step" | tail -4
timeout -s KILL 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?
tail -4 gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/bench_n1.json
