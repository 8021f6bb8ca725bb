python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/cli_throughput.py 2000000 > gpurun_out/r2t_cli.json 2>/dev/null; cat gpurun_out/r2t_cli.json
python tools/cli_throughput.py 4000000 > gpurun_out/r2t_cli4.json 2>/dev/null; cat gpurun_out/r2t_cli4.json
python tools/e2e_ab.py 10000000 ascii > gpurun_out/r2t_ab.log 2>&1; cat gpurun_out/r2t_ab.log
