python -m pytest tests/test_gpu_parity.py tests/test_gpu_stream.py -x -q 2>&1 | tail -5
python tests/scripts/fuzz_parity.py 90 44 2>&1 | tail -1
T0=$(date +%s); python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench wall $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/r2x_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2x_bench.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "packed", d["e2e"]["packed_input"]["ms_per_step"])
for k,v in d["configs"].items(): print(k, v.get("value"), v.get("gather_frac"), v.get("kernel_ms_per_step"), v.get("byte_path_reads"), v.get("input_format"), v.get("error"))
PY
