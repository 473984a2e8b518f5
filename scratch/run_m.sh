set -x
python -m pytest tests/test_parity_gpu.py tests/test_tool_gpu.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t_m.txt
python scratch/decomp.py 2>&1 | tail -6 > gpurun_out/decomp_m.txt
ncu --set full --import-source on --clock-control none -k regex:bbduk_fast_kernel -s 4 -c 1 -o gpurun_out/prof_fast_m -f python bench.py --steps 2 --warmup 1 --cpu-pairs 0 --pairs-per-step 2097152 > /dev/null 2>&1
cat gpurun_out/t_m.txt gpurun_out/decomp_m.txt
ls -la gpurun_out/prof_fast_m.ncu-rep
