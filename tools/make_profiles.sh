#!/bin/bash
# GPU box: the evidence set of a round (bench JSONs, ncu launch list, ncu --set full kernel table) into gpurun_out/prof/.
#   bash tools/make_profiles.sh r02
set -u
R=${1:-r02}
O=gpurun_out/prof
mkdir -p $O
python bench.py > $O/${R}_bench_bf16_n1.json 2> $O/bench_n1.err
python bench.py --precision fp32 --steps 5 --no-extra --no-eager --no-cpu-baseline --no-latency --no-train > $O/${R}_bench_fp32_n1.json 2> $O/bench_fp32.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference_arm.json 2> $O/bench_ref.err
python tools/tail_bench.py 37 10 > $O/${R}_tail_bench.txt 2>&1
python tools/profile_target.py bf16 37 > $O/pt.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_bf16_launches.csv python tools/profile_target.py bf16 37 > $O/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"l2d_front|dsconv_tc|bottleneck_s|ppm_|ffm_t|upsample_argmax" -s 21 -c 21 -o $O/${R}_full -f python tools/profile_target.py bf16 37 > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
