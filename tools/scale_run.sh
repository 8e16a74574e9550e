#!/bin/bash
# Weak-scaling bench lines under torchrun on N GPUs of one box:  tools/scale_run.sh N [full]
# (full: also BASELINE config 2 and the NCCL all-reduce form of the prototype exchange)
mkdir -p gpurun_out/r2s
N=$1
run() { # name, extra args
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 20 --warmup 5 $2 2> gpurun_out/r2s/$1.err | grep "^{" > gpurun_out/r2s/$1.json
}
run bench_n${N}_c5 ""
if [ "$2" = full ]; then
  run bench_n${N}_c2 "--config 2 --no-e2e"
  C3D_PEER_EXCHANGE=0 run bench_n${N}_c2_nccl "--config 2 --no-e2e --min-seconds 0.5"
  C3D_PEER_EXCHANGE=0 run bench_n${N}_c5_nccl "--no-e2e --min-seconds 0.5"
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2s/bench_n${N}_*.json")):
    try:
        j=json.load(open(f)); e=j.get("e2e") or {}
        print(f.split("/")[-1], "%.0f scans/s"%j["value"], "%.1f us"%(1e3*j["ms_per_step"]), j["run"]["banks_identical_across_ranks"], "e2e", e.get("value"))
    except Exception as ex: print(f, "ERR", ex)
PY
