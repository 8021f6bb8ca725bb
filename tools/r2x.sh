python -m pytest tests/test_gpu_parity.py -x -q -k "with_n_on or mixed_vs" 2>&1 | tail -5
/usr/bin/time -v python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; grep -a "Elapsed (wall" gpurun_out/r2x_bench.err; grep -av "^\s" gpurun_out/r2x_bench.err | tail -5
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2x_bench.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "packed", d["e2e"]["packed_input"]["ms_per_step"])
print("fastq", {k:v for k,v in d["e2e"]["from_fastq"].items() if k in ("value","ms")}, d["e2e"]["from_fastq"].get("gzip",{}).get("value"), d["e2e"]["from_fastq"].get("bgzf",{}).get("value"))
for k,v in d["configs"].items(): print(k, v.get("value"), v.get("gather_frac"), v.get("kernel_ms_per_step"), v.get("byte_path_reads"))
print(d["roofline"]["frac"], d["cpu_baseline"])
PY
