#!/bin/bash
# One gpurun call for the end-of-round summary, most important stage first (a clamped call keeps what finished):
# GPU parity tests, the bench line with the driver's own arguments, the ncu launch list of the bench command, one full
# ncu capture of exactly one step, then the parity tests again on the index-asserting library (make checked).
# usage: tools/gpu_round4.sh <tag>
tag=$1
mkdir -p gpurun_out
t0=$(date +%s)
timeout 480 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - t0 ))s"; tail -3 gpurun_out/pytest_$tag.log
timeout 420 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$? t=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "serial", round(d["roofline"]["single_stream_ms_per_step"], 3),
          "e2e", round(e["value"]), "steps", e["steps"], "sync", round(e["synchronous_call"]["value"]), "full", round(e["full_layout"]["value"]),
          "d2h", round(e["d2h_gbs"], 1), "ceiling", round(e["d2h_ceiling_gbs"], 1), "probe", d["parity_probe"]["result"])
    print({k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
    print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
    print("extras", {k: (v if not isinstance(v, dict) else {a: b for a, b in list(v.items())[:6]}) for k, v in d.get("extras", {}).items()})
except Exception as ex:
    print("ERR", ex)
PY
N=$(python tools/profile_step.py --steps 1 --batch 4096 | awk '{print $2}'); echo "launches per step: $N"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_l_$tag.log 2>&1; echo "ncu list rc=$? t=$(( $(date +%s) - t0 ))s"
timeout 300 ncu --set full --clock-control none --import-source on --launch-skip $N -c $N -f -o gpurun_out/prof_$tag \
    python tools/profile_step.py --steps 2 --batch 4096 > gpurun_out/ncu_f_$tag.log 2>&1; echo "ncu full rc=$? t=$(( $(date +%s) - t0 ))s"
if [ -f gpurun_variants/lib_checked.so ]; then
  BPC_LIB=$PWD/gpurun_variants/lib_checked.so timeout 300 python -m pytest tests -m gpu -q -x -k "not long_segments and not real_fixture" > gpurun_out/pytest_checked_$tag.log 2>&1
  echo "checked pytest rc=$? t=$(( $(date +%s) - t0 ))s"; tail -2 gpurun_out/pytest_checked_$tag.log
fi
