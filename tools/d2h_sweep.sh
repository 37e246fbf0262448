#!/bin/bash
# One gpurun --gpus 8 call: box topology + the concurrent D2H ceiling at N = 1, 2, 4, 8 (tools/d2h_ceiling.py).
mkdir -p gpurun_out
{
  nvidia-smi topo -m
  lscpu | head -30
  for n in /sys/devices/system/node/node*; do echo "$n cpus $(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
  nproc
} > gpurun_out/topo_$1.txt 2>&1
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  [ $N -gt $NG ] && break
  if [ $N -eq 1 ]; then
    timeout 300 python tools/d2h_ceiling.py > gpurun_out/d2h_$1_n$N.json 2> gpurun_out/d2h_$1_n$N.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      tools/d2h_ceiling.py > gpurun_out/d2h_$1_n$N.json 2> gpurun_out/d2h_$1_n$N.err
  fi
  echo "N=$N rc=$?"; tail -c 600 gpurun_out/d2h_$1_n$N.json
done
