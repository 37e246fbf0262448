#!/bin/bash
# One gpurun call for A/B runs: GPU parity tests, then a short bench per "ENV=VAL,ENV2=VAL" variant given on the command line.
# usage: tools/gpu_ab.sh <tag> [variant ...]      (variant "-" = no extra environment)
tag=$1; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$tag.log
i=0
for v in "$@"; do
  i=$((i+1))
  envs=$(echo "$v" | tr ',' ' '); [ "$v" = "-" ] && envs=""
  env $envs python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_${tag}_$i.json 2> gpurun_out/bench_${tag}_$i.err; echo "bench[$v] rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${tag}_$i.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("  value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "serial", round(d["roofline"]["single_stream_ms_per_step"], 3), "e2e", round(e["value"]), "probe", d["parity_probe"]["result"])
    print("  ", {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
except Exception as ex:
    print("ERR", ex)
PY
done
