#!/bin/bash
# Short end-of-round evidence: the bench line with the driver's arguments, then the ncu launch list of the bench command.
# usage: tools/gpu_final.sh <tag>
tag=$1
mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$? t=$(( $(date +%s) - t0 ))s"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "serial", round(d["roofline"]["single_stream_ms_per_step"], 3),
          "e2e", round(e["value"]), "steps", e["steps"], "sync", round(e["synchronous_call"]["value"]), "full", round(e["full_layout"]["value"]),
          "d2h", round(e["d2h_gbs"], 1), "ceiling", round(e["d2h_ceiling_gbs"], 1), "probe", d["parity_probe"]["result"])
    print({k: round(v, 3) for k, v in d["roofline"]["kernel_ms_per_step"].items()})
    print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "fp64", d["roofline"]["fp64"]["pipe_busy_frac"])
    print("sweep", {k: round(v["audio_seconds_per_s"]) for k, v in d["extras"]["config4_long_segment_sweep"].items() if isinstance(v, dict)})
    print("cfg2", d["extras"]["config2_logmel"]["ms_per_step"])
except Exception as ex:
    print("ERR", ex)
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_l_$tag.log 2>&1; echo "ncu list rc=$? t=$(( $(date +%s) - t0 ))s"
