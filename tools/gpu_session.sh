set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest.log; tail -4 gpurun_out/pytest.log
timeout -s KILL 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo rc=$?
tail -6 gpurun_out/bench_n1.err
timeout -s KILL 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?
tail -3 gpurun_out/bench_ref.err; head -c 1200 gpurun_out/bench_ref.json
