# memcheck of the kernels added in the last sessions (ZSTD / GZIP decoders, hashed and row-key GROUP BY, JSON egress)
set -x
mkdir -p gpurun_out
( time timeout -s KILL 420 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q \
   -k "compressed_pages_decoded_on_gpu or hashed or without_dictionary or json_egress or plain_byte_array or garbled" ) > gpurun_out/memcheck_r2k.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/memcheck_r2k.log | head -20; tail -5 gpurun_out/memcheck_r2k.log
