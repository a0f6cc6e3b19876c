#!/usr/bin/env bash
# Strong-scaling line at N GPUs (peer reducer), with e2e.  usage: gpurun --gpus N -- 'bash tools/r2_scale.sh N'
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --steps 50 --warmup 5 ${EXTRA:-} > gpurun_out/bench_scale_n$N.json 2> gpurun_out/bench_scale_n$N.err
echo "rc=$?"; tail -3 gpurun_out/bench_scale_n$N.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_scale_n$N.json") if l.startswith("{")][-1])
    print("N=$N ms/step", round(d["ms_per_step"], 4), "value %.4g" % d["value"], d["scaling"])
    print({k: round(v["ms"], 4) for k, v in d["kernels"].items()})
    print("peer phases (us):", d["kernels"].get("peer_reduce_expand", {}).get("phases_us_rank0_last_call"))
    print("e2e", d["e2e"] and round(d["e2e"]["ms_per_step"], 3), "active", d["e2e_active_voxels"] and round(d["e2e_active_voxels"]["ms_per_step"], 3), "verify", d["verify"])
except Exception as e:
    print("no bench line:", e)
PY
