#!/bin/bash
# A/B of one environment switch on the train-step bench + the GPU suite, one box round trip.
#   scripts/gpu_ab.sh <tag> <ENVVAR>      (bench with ENVVAR unset, then ENVVAR=0)
tag=${1:-ab}; var=${2:-STARCOP_PDL}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
B="--steps 30 --warmup 5 --no-cpu-baseline --no-spectral --no-eager-baseline --no-extras"
python bench.py $B > gpurun_out/${tag}_on.json 2> gpurun_out/${tag}_on.err
env $var=0 python bench.py $B > gpurun_out/${tag}_off.json 2> gpurun_out/${tag}_off.err
python bench.py $B > gpurun_out/${tag}_on2.json 2>> gpurun_out/${tag}_on.err
python bench.py --dtype f32 --steps 5 --warmup 3 --no-cpu-baseline --no-spectral --no-eager-baseline --no-extras > gpurun_out/${tag}_f32.json 2> gpurun_out/${tag}_f32.err
for f in on off on2 f32; do echo -n "$f: "; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${tag}_$f.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"))
except Exception as e: print("ERR", e)
P
done
tail -3 gpurun_out/${tag}_on.err
