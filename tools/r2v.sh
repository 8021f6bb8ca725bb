python -m pytest tests/test_gpu_cli.py -x -q 2>&1 | tail -3
python tools/cli_throughput.py 2000000 > gpurun_out/r2v_cli.json 2>/dev/null; cat gpurun_out/r2v_cli.json
python tools/cli_throughput.py 8000000 > gpurun_out/r2v_cli8.json 2>/dev/null; cat gpurun_out/r2v_cli8.json
timeout 400 python tests/scripts/fuzz_stream.py 200 31 > gpurun_out/r2v_fuzz_stream.log 2>&1; tail -3 gpurun_out/r2v_fuzz_stream.log
