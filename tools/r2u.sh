python -m pytest tests/test_gpu_cli.py tests/test_gpu_stream.py -x -q 2>&1 | tail -4
python tools/cli_throughput.py 2000000 --profile > gpurun_out/r2u_cli.json 2> gpurun_out/r2u_cli.prof; cat gpurun_out/r2u_cli.json
python tools/cli_throughput.py 8000000 > gpurun_out/r2u_cli8.json 2>/dev/null; cat gpurun_out/r2u_cli8.json
