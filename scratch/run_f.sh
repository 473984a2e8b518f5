cp scratch/lib_dbg.so bbtools_b200/libbbduk_b200.so
python scratch/dbgcount.py 2>&1 | tail -3 | tee gpurun_out/dbgcount.txt
