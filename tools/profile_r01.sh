#!/bin/bash
# Round-1 profiling pass (run under gpurun, ONE GPU).  Outputs go to gpurun_out/.
set -x
mkdir -p gpurun_out
# 1) launch list: every kernel with its device time (serialised, cold cache -> compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 420 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 1 --warmup 1 --playouts 48 --no-graph --no-cpu-baseline --profile-waves 4 > gpurun_out/r01_launches_bench.log 2>&1
# 2) full capture of our tree kernel
ncu --set full --clock-control none --import-source on -k regex:k_wave -s 60 -c 3 -o gpurun_out/r01_kwave \
    python bench.py --steps 1 --warmup 1 --playouts 48 --no-graph --no-cpu-baseline --profile-waves 4 > gpurun_out/r01_kwave_bench.log 2>&1
# 3) full capture of the dominant network kernel (library convolution) for tensor-pipe utilisation
ncu --set full --clock-control none -k regex:"conv|gemm|cutlass|cudnn|xmma|sm100|sm90|sm80" -s 200 -c 4 -o gpurun_out/r01_nnconv \
    python bench.py --steps 1 --warmup 1 --playouts 48 --no-graph --no-cpu-baseline --profile-waves 4 > gpurun_out/r01_nnconv_bench.log 2>&1
ls -la gpurun_out
