python -m pytest tests/test_parity_gpu.py tests/test_tool_gpu.py tests/test_direct_gpu.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t_e.txt
python scratch/decomp.py 2>&1 | tail -5 > gpurun_out/decomp_e.txt
cat gpurun_out/t_e.txt gpurun_out/decomp_e.txt
