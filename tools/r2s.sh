python -m pytest tests/test_gpu_cli.py tests/test_gpu_stream.py -x -q 2>&1 | tail -4
python tools/cli_throughput.py 2000000 --profile > gpurun_out/r2s_cli.json 2> gpurun_out/r2s_cli.prof; cat gpurun_out/r2s_cli.json
python tools/e2e_ab.py 10000000 ascii > gpurun_out/r2s_ab.log 2>&1; cat gpurun_out/r2s_ab.log
