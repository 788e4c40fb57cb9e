#!/bin/bash
# bench the train step under several environment variants in one box round trip:
#   scripts/gpu_variants.sh tag "VAR=1 VAR2=x" "VAR3=y" ...   ("-" = no variables)
tag=$1; shift
mkdir -p gpurun_out
B="--steps 30 --warmup 5 --no-cpu-baseline --no-spectral --no-eager-baseline --no-extras"
i=0
for v in "$@"; do
  [ "$v" = "-" ] && v=""
  env $v python bench.py $B > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  echo -n "[$v] "; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${tag}_$i.json").read().strip().splitlines()[-1]); print(round(d["value"],1), round(d["ms_per_step"],3), round(d.get("e2e",{}).get("value",0),1))
except Exception as e: print("ERR", e)
P
  i=$((i+1))
done
