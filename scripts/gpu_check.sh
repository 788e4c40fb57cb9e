#!/bin/bash
# One GPU-box round trip: the -m gpu suite, the bench line, optionally the ncu launch list of one step.
#   scripts/gpu_check.sh <tag> [launches]
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=12 2>&1 | tail -60 > gpurun_out/${tag}_pytest.log
tail -25 gpurun_out/${tag}_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cut -c1-300 gpurun_out/${tag}_bench.json
tail -4 gpurun_out/${tag}_bench.err
if [ "$2" = "launches" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile --no-graph --no-spectral > gpurun_out/${tag}_prof.log 2>&1
  python scripts/launch_summary.py gpurun_out/${tag}_launches.csv 150 > gpurun_out/${tag}_launch_summary.txt 2>&1
  head -40 gpurun_out/${tag}_launch_summary.txt
fi
