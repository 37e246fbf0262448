#!/bin/bash
# Launch size vs step time vs DRAM traffic (BPC_CHUNK): short bench + one ncu pass with the two DRAM byte counters per chunk size.
# --cache-control none: ncu's default flushes the L2 before every kernel, which hides exactly the reuse between
# consecutive kernels that a smaller launch is meant to create.
# usage: tools/gpu_chunks.sh <tag> <chunk> [chunk ...]
tag=$1; shift
mkdir -p gpurun_out
for c in "$@"; do
  BPC_CHUNK=$c python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_${tag}_c$c.json 2> gpurun_out/bench_${tag}_c$c.err
  BPC_CHUNK=$c ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none --csv \
      --log-file gpurun_out/dram_${tag}_c$c.csv python tools/profile_step.py --steps 2 --batch 4096 > gpurun_out/ncu_${tag}_c$c.log 2>&1
  python - <<PY
import csv, json, collections
d = json.loads(open("gpurun_out/bench_${tag}_c$c.json").read().strip().splitlines()[-1])
rows = [r for r in csv.reader(open("gpurun_out/dram_${tag}_c$c.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit"); ii = hdr.index("ID")
ids = sorted({int(r[ii]) for r in rows[1:]})
half = ids[len(ids) // 2:]                      # second of the two steps
tot = collections.Counter(); per = collections.Counter()
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in rows[1:]:
    if int(r[ii]) not in half or "dram__bytes" not in r[mi]: continue
    v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1)
    tot[r[mi]] += v; per[r[ki].split("(")[0][-28:]] += v
tb = sum(tot.values())
print("chunk $c: value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "| DRAM GB/step", round(tb / 1e9, 3), "=", round(tb / (4096 * 354448), 2), "x algorithmic; launches", len(half))
PY
done
