# End-of-round GPU session: parity suite, the bench line, the launch list of the same command, codec probes.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tail -1
( time timeout -s KILL 600 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/gputests_final.log 2>&1; tail -12 gpurun_out/gputests_final.log
( time timeout -s KILL 900 python bench.py ) > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; tail -c 600 gpurun_out/bench_n1_final.json; tail -5 gpurun_out/bench_n1_final.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2k.csv python bench.py --steps 3 --warmup 3 --skip-cpu > gpurun_out/launches_r2k.log 2>&1; tail -2 gpurun_out/launches_r2k.log | cut -c1-300
for codec in ZSTD GZIP LZ4; do
  echo "== open probe $codec"
  PQB_BENCH_CODEC=$codec timeout -s KILL 240 python tests/scripts/open_probe.py 48 2>&1 | grep -E "^step|decompress|generated" | tail -6
done
