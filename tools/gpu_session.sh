# last GPU session of the round: the parity suite on the final build, then the ZSTD open probe with the windowed bit reader
set -x
mkdir -p gpurun_out
( time timeout -s KILL 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/gputests_r2l.log 2>&1; tail -6 gpurun_out/gputests_r2l.log
PQB_BENCH_CODEC=ZSTD timeout -s KILL 120 python tests/scripts/open_probe.py 48 2>&1 | grep -E "^step|generated" | tail -5
