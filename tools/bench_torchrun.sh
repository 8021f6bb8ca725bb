# usage: bash tools/bench_torchrun.sh N [extra bench.py flags]  -- bench.py on N GPUs of one box the way the driver launches it; summary of the line
python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $1 --steps 10 --warmup 3 $2 > gpurun_out/bench_n$1.json 2> gpurun_out/bench_n$1.err
tail -c 1200 gpurun_out/bench_n$1.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n$1.json") if l.startswith("{")][-1])
e=d["e2e"]
print("value", d["value"], "e2e", e["value"], e["ms_per_step"], "pcie_only", e["ascii_over_pcie_only"]["value"], "packed", e["packed_input"]["value"], e["packed_input"]["ms_per_step"])
print("from_fastq", e["from_fastq"])
print("output_parity", d["output_parity"])
print("configs", list((d.get("configs") or {}).keys()))
PY
