cp scratch/lib_dense.so bbtools_b200/libbbduk_b200.so
python scratch/decomp.py 2>&1 | tail -5 | head -2
python scratch/decomp.py 2>&1 | tail -5 | head -2
