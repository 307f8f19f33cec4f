set -x
mkdir -p gpurun_out
PQB_BENCH_CODEC=LZ4 timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_lz4.csv python tests/scripts/open_probe.py 480 > gpurun_out/lz4_probe.log 2>&1; tail -3 gpurun_out/lz4_probe.log
