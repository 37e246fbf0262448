#!/bin/bash
# One gpurun call for a round summary: GPU parity tests, full bench line, ncu launch list and one full ncu capture of
# exactly one step (the launch count of a step is measured first).  usage: tools/gpu_round3.sh <tag>
tag=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "serial", round(d["roofline"]["single_stream_ms_per_step"], 3),
          "e2e", round(e["value"]), "sync", round(e["synchronous_call"]["value"]), "full", round(e["full_layout"]["value"]),
          "d2h", round(e["d2h_gbs"], 1), "ceiling", round(e["d2h_ceiling_gbs"], 1), "probe", d["parity_probe"]["result"])
    print({k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
    print("extras", {k: (v if not isinstance(v, dict) else {a: b for a, b in list(v.items())[:6]}) for k, v in d.get("extras", {}).items()})
except Exception as ex:
    print("ERR", ex)
PY
N=$(python tools/profile_step.py --steps 1 --batch 4096 | awk '{print $2}'); echo "launches per step: $N"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_l_$tag.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on --launch-skip $N -c $N -f -o gpurun_out/prof_$tag \
    python tools/profile_step.py --steps 2 --batch 4096 > gpurun_out/ncu_f_$tag.log 2>&1; echo "ncu full rc=$?"
