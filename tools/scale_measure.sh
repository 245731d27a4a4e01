set -x
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python tools/h2d_scaling_probe.py; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/h2d_scaling_probe.py; fi
done > gpurun_out/r02_h2d_scaling_probe.jsonl 2> gpurun_out/h2d_probe.err
cat gpurun_out/r02_h2d_scaling_probe.jsonl
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_bench_${n}gpu.json 2> gpurun_out/bench${n}.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_${n}gpu.json"))
print($n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("runs"), d["gather_validated"], d["gather_bus_GBps"])
PY
done
