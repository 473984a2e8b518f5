python -m pytest tests/test_parity_gpu.py tests/test_tool_gpu.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t_n.txt
python scratch/decomp.py 2>&1 | tail -6 > gpurun_out/decomp_n.txt
cat gpurun_out/t_n.txt gpurun_out/decomp_n.txt
