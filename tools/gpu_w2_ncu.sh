#!/bin/bash
# ncu --set full of the STFT-2048 frame kernel alone, for the in-tree library and prebuilt variants
# usage: tools/gpu_w2_ncu.sh <tag> [variant ...]
tag=$1; shift
mkdir -p gpurun_out
cap() {  # name, env...
  name=$1; shift
  env "$@" timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_frame2048 --launch-skip 1 --launch-count 1 -f \
      -o gpurun_out/prof_${tag}_$name python tools/profile_step.py --steps 2 --batch 4096 > gpurun_out/ncu_${tag}_$name.log 2>&1
  echo "$name rc=$?"
}
cap intree BPC_DUMMY=1
for v in "$@"; do cap $v BPC_LIB=$PWD/gpurun_variants/lib_$v.so; done
