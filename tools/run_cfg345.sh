#!/bin/sh
# Configs 3, 4, 5 (BASELINE.json) on one B200: bench lines with cpu_baseline + parity slice (--verify builds the CPU oracle's
# table, minutes), then one `ncu --set full` capture of the dominant kernel of each and a launch list. Outputs: gpurun_out/r02_cfg*.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --workload cfg3 --verify --cpu-pairs 1048576 --steps 5 > gpurun_out/r02_cfg3_bench.json 2> gpurun_out/r02_cfg3_bench.err
python bench.py --workload cfg4 --verify --cpu-pairs 262144 --steps 5 > gpurun_out/r02_cfg4_bench.json 2> gpurun_out/r02_cfg4_bench.err
python bench.py --workload cfg5 --steps 5 > gpurun_out/r02_cfg5_bench.json 2> gpurun_out/r02_cfg5_bench.err
ncu --set full --clock-control none --import-source on -k regex:bbduk_direct_kernel -s 3 -c 1 -f -o gpurun_out/r02_cfg3_direct \
    python bench.py --workload cfg3 --steps 1 --cpu-pairs 0 --pairs-per-step 2097152 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:bbduk_direct_kernel -s 3 -c 1 -f -o gpurun_out/r02_cfg4_direct \
    python bench.py --workload cfg4 --steps 1 --cpu-pairs 0 --pairs-per-step 2097152 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:kcount_kernel -s 3 -c 1 -f -o gpurun_out/r02_cfg5_kcount \
    python bench.py --workload cfg5 --steps 1 --cpu-pairs 0 --pairs-per-step 2097152 > /dev/null 2>&1
ls -la gpurun_out/r02_cfg*
