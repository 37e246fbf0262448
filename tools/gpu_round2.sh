#!/bin/bash
# One gpurun call of round 2: GPU parity tests, smoke, bench (with the host-path trace).  usage: tools/gpu_round2.sh <tag> [ncu]
tag=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$tag.log
BPC_HOST_TRACE=1 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(e["value"]), "full", round(e["full_layout"]["value"]),
          "d2h", round(e["d2h_gbs"], 1), "ceiling", round(e["d2h_ceiling_gbs"], 1), "probe", d["parity_probe"]["result"], "check", e["matches_device_path"])
    for k, v in d["roofline"]["kernel_ms_per_step"].items(): print("  ", k, round(v, 3))
    print("fp64", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["roofline"]["fp64"].items() if k in ("achieved_tflops", "frac", "pipe_busy_frac")})
    print("extras config2", d["extras"].get("config2_logmel"))
except Exception as ex:
    print("ERR", ex)
PY
if [ "$2" = "ncu" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_l_$tag.log 2>&1; echo "ncu list rc=$?"
  ncu --set full --clock-control none --import-source on --launch-skip 13 -c 13 -f -o gpurun_out/prof_$tag \
      python tools/profile_step.py --steps 2 --batch 4096 > gpurun_out/ncu_f_$tag.log 2>&1; echo "ncu full rc=$?"
fi
