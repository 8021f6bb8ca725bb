python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/cli_throughput.py 2000000 --profile > gpurun_out/r2j_cli.json 2> gpurun_out/r2j_cli.prof; cat gpurun_out/r2j_cli.json
python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -c 1200 gpurun_out/r2j_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2j_ref.json 2> gpurun_out/r2j_ref.err; cat gpurun_out/r2j_ref.json; tail -c 600 gpurun_out/r2j_ref.err
