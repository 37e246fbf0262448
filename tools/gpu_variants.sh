#!/bin/bash
# A/B of prebuilt library variants (gpurun_variants/lib_<name>.so): each is copied over the in-tree library and benched.
# usage: tools/gpu_variants.sh <name> [name ...]
cp breathing-phase-classifier_b200/bpc_b200/libbpc_b200.so /tmp/lib_orig.so
for v in "$@"; do
  cp gpurun_variants/lib_$v.so breathing-phase-classifier_b200/bpc_b200/libbpc_b200.so
  python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > /tmp/b_$v.json 2> /tmp/b_$v.err
  python - <<PY
import json
try:
    d = json.loads(open("/tmp/b_$v.json").read().strip().splitlines()[-1])
    k = d["roofline"]["kernel_ms_per_step"]
    print("$v", "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "probe", d["parity_probe"]["result"], {n: round(x, 3) for n, x in k.items() if "cens" in n or "frame" in n or "stft" in n})
except Exception as ex:
    print("$v ERR", ex)
PY
done
cp /tmp/lib_orig.so breathing-phase-classifier_b200/bpc_b200/libbpc_b200.so
