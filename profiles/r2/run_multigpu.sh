#!/bin/bash
# usage: run_multigpu.sh N  — PCIe probe at 1..N GPUs, then the bench the way the driver launches it at N GPUs
N=${1:-2}
mkdir -p gpurun_out
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  CUDA_VISIBLE_DEVICES=$(seq -s, 0 $((n-1))) python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n \
      profiles/r2/pcie_probe.py 2>/dev/null | grep '^{' | tee -a gpurun_out/r2_pcie_probe.jsonl
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 2000 --warmup 100 \
    > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -c 600 gpurun_out/r2_bench_${N}gpu.err
python - $N <<'PY'
import json, sys
d = json.loads([l for l in open(f"gpurun_out/r2_bench_{sys.argv[1]}gpu.json") if l.startswith("{")][-1])
print("N", d["n_gpus"], "value %.4g" % d["value"], "us/step %.2f" % (d["ms_per_step"] * 1e3), "e2e %.4g" % d["e2e"]["value"], d["e2e"]["step_ms"], d["config"]["collective"])
for k, v in d["secondary"].items():
    print(" ", k, "value %.4g" % v["value"], v["config"]["workload"][:70], {kk: "%.4g" % vv["value"] for kk, vv in v.items() if isinstance(vv, dict) and "value" in vv})
PY
