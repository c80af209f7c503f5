#!/bin/bash
# one bench line at N ranks into gpurun_out/<tag>_bench_n<N>.json ; usage: scale_run.sh <tag> <N> [extra bench args]
tag=$1; n=$2; shift 2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus $n "$@" > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err
echo "rc=$?"; tail -c 400 gpurun_out/${tag}_bench_n${n}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench_n${n}.json").read().strip().splitlines()[-1])
print("N=$n ms/step", d["ms_per_step"], "value", d["value"], d["config"]["launch"], "eager", d["config"]["eager_ms_per_step"], "bwd", d["config"]["backward"])
print("kernels", {k:round(v["ms_per_launch"],4) for k,v in d["kernels"].items()})
print("whole_step", d["roofline"]["whole_step"]); print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"]); print("parity", d["parity"]); print("clocks", d["clocks"])
r=d.get("retrieval")
if r: print("retr", r["value"], r["ms_per_step"], r["device_only_ms_per_step"], r.get("roofline",{}).get("frac"), r["two_kernel_form"].get("topk_roofline",{}).get("frac"), r.get("query_sharded_comparison"), r["e2e"]["value"])
PY
