#!/bin/bash
# Round-1 final profiling pass (board mode + native network ends).  ONE GPU.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 300 --csv --log-file gpurun_out/r01b_launches.csv \
    python bench.py --steps 1 --warmup 1 --playouts 48 --no-graph --no-cpu-baseline --profile-waves 4 > gpurun_out/r01b_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_wave -s 60 -c 2 -o gpurun_out/r01b_kwave \
    python bench.py --steps 1 --warmup 1 --playouts 48 --no-graph --no-cpu-baseline --profile-waves 4 > gpurun_out/r01b_kwave_bench.log 2>&1
ncu --set full --clock-control none -k regex:"k_first_conv|k_head_conv|k_value_mlp|k_policy_fc" -s 80 -c 4 -o gpurun_out/r01b_ends \
    python bench.py --steps 1 --warmup 1 --playouts 48 --no-graph --no-cpu-baseline --profile-waves 4 > gpurun_out/r01b_ends_bench.log 2>&1
