set -x
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 200 bash tests/scripts/mgpu_run.sh 150 2>&1 | tail -4
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --skip-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?
tail -4 gpurun_out/bench_n2.err; cut -c1-300 gpurun_out/bench_n2.json
