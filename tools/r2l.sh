python -m pytest tests/test_gpu_cli.py tests/test_gpu_stream.py -x -q 2>&1 | tail -4
python tools/cli_throughput.py 2000000 --profile > gpurun_out/r2l_cli.json 2> gpurun_out/r2l_cli.prof; cat gpurun_out/r2l_cli.json
python tools/cli_throughput.py 2000000 --gz > gpurun_out/r2l_cli_gz.json 2>/dev/null; cat gpurun_out/r2l_cli_gz.json
python bench.py --no-configs > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; tail -c 600 gpurun_out/r2l_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_bench.json'))
print(json.dumps(d['e2e']['from_fastq'])[:1500])
PY
