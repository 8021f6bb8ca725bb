python -m pytest tests/test_gpu_stream.py tests/test_gpu_parity.py tests/test_gpu_cli.py -x -q 2>&1 | tail -8
python tools/e2e_ab.py > gpurun_out/r2f_ab.log 2>&1
for route in packed ascii; do
  echo "######## route=$route" >> gpurun_out/r2f_trace.log
  ORPHEUM_B200_TRACE=1 python tools/e2e_ab.py run $route 2>&1 | tail -45 >> gpurun_out/r2f_trace.log
done
ORPHEUM_B200_TRACE=1 python tools/stream_bench.py 10000000 0 > gpurun_out/r2f_stream_bench.log 2>&1
grep -v "orph trace" gpurun_out/r2f_stream_bench.log; cat gpurun_out/r2f_ab.log
