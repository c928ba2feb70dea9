#!/bin/bash
# Round-2 profiling recipe (run under gpurun, 1 GPU).  Outputs land in gpurun_out/.
#  1. launch list of a short bench (shares of the step per kernel)
#  2. ncu --set full of k_wave IN SITU: two warm-up plies of 1200 playouts first (deep trees), then eager waves
#  3. ncu --set full of the small-batch trunk kernel (cz_tower.cu)
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 38000 -c 340 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --kwave-capture --warmup 2 --profile-waves 24 --no-graph --legs none > gpurun_out/r02_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wave -s 2415 -c 3 -o gpurun_out/r02_kwave \
    python bench.py --kwave-capture --warmup 2 --profile-waves 40 --no-graph --legs none > gpurun_out/r02_kwave_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tower_small -s 30 -c 2 -o gpurun_out/r02_tower \
    python tools/latency_bench.py 400 4 7 1 > gpurun_out/r02_tower_bench.log 2>&1
