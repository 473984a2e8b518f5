python scratch/decomp.py 2>&1 | tail -5
python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -1
