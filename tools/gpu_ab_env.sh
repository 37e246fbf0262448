#!/bin/bash
# GPU parity suite on the in-tree library, then short bench lines for environment variants.
# usage: tools/gpu_ab_env.sh <tag> "NAME=ENV=VAL[,ENV=VAL]" ...      (NAME=BPC_DUMMY=1 for the default)
tag=$1; shift
mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - t0 ))s"; tail -4 gpurun_out/pytest_$tag.log
for spec in "$@"; do
  name=${spec%%=*}; envs=${spec#*=}
  env $(echo $envs | tr ',' ' ') timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/b_${tag}_$name.json 2> gpurun_out/b_${tag}_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/b_${tag}_$name.json").read().strip().splitlines()[-1])
    k = d["roofline"]["kernel_ms_per_step"]
    print("$name", "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "serial", round(d["roofline"]["single_stream_ms_per_step"], 3),
          "probe", d["parity_probe"]["result"], {n: round(v, 3) for n, v in k.items()}, "t=$(( $(date +%s) - t0 ))s")
except Exception as ex:
    print("$name ERR", ex)
PY
done
