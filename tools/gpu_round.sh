#!/bin/bash
# One gpurun call: GPU parity tests, bench (side streams on / off), ncu launch list and one full ncu capture.
# usage: tools/gpu_round.sh <tag> [full]
tag=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
BPC_STREAMS=0 BPC_COMPACT_D2H=0 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu > gpurun_out/bench_${tag}_serial.json 2> gpurun_out/bench_${tag}_serial.err
python - <<PY
import json
for f in ("gpurun_out/bench_$tag.json", "gpurun_out/bench_${tag}_serial.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"].get("single_stream_ms_per_step"))
        for k, v in d["roofline"]["kernel_ms_per_step"].items(): print("  ", k, round(v, 3))
    except Exception as e: print(f, "ERR", e)
PY
if [ "$2" = "full" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_l_$tag.log 2>&1; echo "ncu list rc=$?"
  ncu --set full --clock-control none --import-source on --launch-skip 13 -c 13 -f -o gpurun_out/prof_$tag \
      python tools/profile_step.py --steps 2 --batch 4096 > gpurun_out/ncu_f_$tag.log 2>&1; echo "ncu full rc=$?"
fi
