#!/bin/bash
# A/B of the two-warps-per-frame STFT-2048 kernel (k_frame2048_w2): parity suite on the in-tree library, then short
# bench lines for the in-tree library (w2), the one-warp kernel (BPC_F2_W2=0) and prebuilt variants.
# usage: tools/gpu_w2.sh <tag> [variant ...]     (variants: gpurun_variants/lib_<name>.so)
tag=$1; shift
mkdir -p gpurun_out
t0=$(date +%s)
timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - t0 ))s"; tail -4 gpurun_out/pytest_$tag.log
one() {  # name, env assignments...
  name=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/b_${tag}_$name.json 2> gpurun_out/b_${tag}_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/b_${tag}_$name.json").read().strip().splitlines()[-1])
    k = d["roofline"]["kernel_ms_per_step"]
    print("$name", "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "serial", round(d["roofline"]["single_stream_ms_per_step"], 3),
          "probe", d["parity_probe"]["result"], "frame2048", round(k.get("k_frame2048", 0), 3), "t=$(( $(date +%s) - t0 ))s")
except Exception as ex:
    print("$name ERR", ex)
PY
}
one w2 BPC_DUMMY=1
one w1 BPC_F2_W2=0
for v in "$@"; do one $v BPC_LIB=$PWD/gpurun_variants/lib_$v.so; done
