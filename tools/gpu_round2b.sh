#!/bin/bash
# e2e host-path sweep + GPU tests (no -x) + short bench.  usage: tools/gpu_round2b.sh <tag>
tag=$1
mkdir -p gpurun_out
python tools/e2e_sweep.py > gpurun_out/e2e_sweep_$tag.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/e2e_sweep_$tag.log | tail -22
python -m pytest tests -m gpu -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_$tag.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(e["value"]), "full", round(e["full_layout"]["value"]),
          "d2h", round(e["d2h_gbs"], 1), "ceiling", round(e["d2h_ceiling_gbs"], 1), "probe", d["parity_probe"]["result"], "check", e["matches_device_path"])
    print("extras config2", d["extras"].get("config2_logmel"))
except Exception as ex:
    print("ERR", ex)
PY
