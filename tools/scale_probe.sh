#!/bin/bash
# One multi-GPU gpurun call: box topology, the concurrent D2H ceiling and bench.py at N = 1, 2, 4, 8 (as many as the box has).
tag=$1; shift
NS="${@:-1 2 4 8}"                 # usage: tools/scale_probe.sh <tag> [N ...]   (default: 1 2 4 8)
mkdir -p gpurun_out
{
  nvidia-smi topo -m
  lscpu | grep -E "^CPU\(s\)|Model name|Socket|NUMA|Thread"
  for n in /sys/devices/system/node/node*; do echo "$n cpus $(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
  nproc; free -g | head -2
} > gpurun_out/topo_$tag.txt 2>&1
NG=$(nvidia-smi -L | wc -l)
for N in $NS; do
  [ $N -gt $NG ] && break
  if [ $N -eq 1 ]; then
    D2H_K=4 timeout 300 python tools/d2h_ceiling.py > gpurun_out/d2h_${tag}_n$N.json 2> gpurun_out/d2h_${tag}_n$N.err
    BPC_HOST_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err
  else
    D2H_K=4 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      tools/d2h_ceiling.py > gpurun_out/d2h_${tag}_n$N.json 2> gpurun_out/d2h_${tag}_n$N.err
    BPC_HOST_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
      bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err
  fi
  echo "N=$N rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/d2h_${tag}_n$N.json").read().strip().splitlines()[-1])
    for k, v in d["results"].items(): print("  d2h", k, v["aggregate_gbs"], v["per_rank_gbs"])
    print("  ranks", [(r["pci"], r["gpu_numa_node"], r.get("pages_default"), r.get("pages_numa")) for r in d["ranks"]])
except Exception as ex: print("  d2h ERR", ex)
try:
    d = json.loads(open("gpurun_out/bench_${tag}_n$N.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("  bench value", round(d["value"]), "e2e", round(e["value"]), "full", round(e["full_layout"]["value"]), "d2h", round(e["d2h_gbs"], 1), "ceiling", round(e["d2h_ceiling_gbs"], 1), "probe", d["parity_probe"]["result"])
except Exception as ex: print("  bench ERR", ex)
PY
done
