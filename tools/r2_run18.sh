#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "adjoint or session" 2>&1 | tail -4
for ts in 1 2 4 8; do
echo "== emulate 1/8, tsplit=$ts"
IONO_PADJ_TSPLIT=$ts timeout 600 python bench.py --steps 50 --warmup 5 --emulate-shard 8 --no-cpu-baseline --no-e2e > gpurun_out/bench_emul8_ts$ts.json 2> gpurun_out/bench_emul8_ts$ts.err; echo "rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_emul8_ts$ts.json") if l.startswith("{")][-1])
print("ms/step", round(d["ms_per_step"], 4), {k: round(v["ms"], 4) for k, v in d["kernels"].items()})
PY
done
echo "== emulate 1/8 auto"
timeout 600 python bench.py --steps 50 --warmup 5 --emulate-shard 8 --no-cpu-baseline --no-e2e > gpurun_out/bench_emul8.json 2> gpurun_out/bench_emul8.err; echo "rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_emul8.json") if l.startswith("{")][-1])
print("ms/step", round(d["ms_per_step"], 4), {k: round(v["ms"], 4) for k, v in d["kernels"].items()})
PY
echo "== emulate 1/2 auto"
timeout 600 python bench.py --steps 50 --warmup 5 --emulate-shard 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_emul2.json 2> gpurun_out/bench_emul2.err; echo "rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_emul2.json") if l.startswith("{")][-1])
print("ms/step", round(d["ms_per_step"], 4), {k: round(v["ms"], 4) for k, v in d["kernels"].items()})
PY
