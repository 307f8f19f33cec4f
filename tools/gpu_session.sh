set -x
mkdir -p gpurun_out
nvidia-smi -L
bash tests/scripts/mgpu_run.sh 150 > gpurun_out/mgpu.log 2>&1; tail -12 gpurun_out/mgpu.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?
grep -v "^W\|^\*\*\*" gpurun_out/bench_n2.err | tail -12; head -c 1500 gpurun_out/bench_n2.json
