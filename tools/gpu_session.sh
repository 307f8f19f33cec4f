set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout -s KILL 120 python tests/scripts/dbg_datebin.py 2>&1 | tail -6
timeout -s KILL 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?
tail -4 gpurun_out/bench_n1.err
