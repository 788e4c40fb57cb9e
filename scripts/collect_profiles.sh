#!/bin/bash
# Round profiles in one GPU-box round trip -> gpurun_out/<tag>_*  (summarised into profiles/ afterwards):
#   scripts/collect_profiles.sh r02
tag=${1:-r02}
o=gpurun_out
mkdir -p $o
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $o/${tag}_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $o/${tag}_bench.json 2> $o/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference.json 2>> $o/${tag}_bench.err
# launch list of the same step the graph replays (eager launches), all kernels
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --profile --no-graph --no-spectral > $o/${tag}_prof.log 2>&1
python scripts/launch_summary.py $o/${tag}_launches.csv 100 > $o/${tag}_launch_summary.txt 2>&1
gzip -f $o/${tag}_launches.csv
# full captures: the step's heaviest kernel families inside a real step, the largest GEMM, the spectral kernels
ncu --set full --clock-control none --import-source on -k regex:"bn_bwd_apply|bn_bwd_reduce|tc_conv3x3_halo|tc_conv_wgrad_kernel" -c 14 \
    -o $o/${tag}_step_kernels python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile --no-graph --no-spectral > $o/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tc_conv_fprop" -s 2 -c 1 -o $o/${tag}_conv python scripts/profile_kernels.py dominant > $o/${tag}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"mag1c_tc" -s 1 -c 2 -o $o/${tag}_mag1c python scripts/profile_kernels.py mag1c > $o/${tag}_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ratio_cluster16" -s 1 -c 1 -o $o/${tag}_ratio python scripts/profile_kernels.py ratio > $o/${tag}_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"srf" -s 1 -c 1 -o $o/${tag}_srf python scripts/srf_bench.py > $o/${tag}_ncu5.log 2>&1
python scripts/layer_bench.py > $o/${tag}_layer_bench.txt 2>&1
python scripts/dw_bench.py > $o/${tag}_dw_bench.txt 2>&1
python scripts/ratio_bench.py > $o/${tag}_ratio_bench.txt 2>&1
python scripts/diag_mag1c_tc.py > $o/${tag}_mag1c_diag.txt 2>&1
python scripts/step_profile.py > $o/${tag}_step_profile.txt 2>&1
for b in fp64_rate smem_rate lds_rate imad_rate; do echo "== $b"; ./scratch/$b; done > $o/${tag}_microbench.txt 2>&1
cp $o/parity512.json $o/${tag}_parity512.json
tail -3 $o/${tag}_pytest_gpu.log; cut -c1-200 $o/${tag}_bench.json; ls -la $o | grep ${tag}_ | awk '{print $5, $9}'
