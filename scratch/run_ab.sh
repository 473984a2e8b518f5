rm -f gpurun_out/ab.txt
for i in 1 2; do
for v in a b; do
cp scratch/lib_$v.so bbtools_b200/libbbduk_b200.so
echo "== $v" >> gpurun_out/ab.txt
python scratch/decomp.py 2>&1 | tail -5 | head -2 >> gpurun_out/ab.txt
done
done
cp scratch/lib_b.so bbtools_b200/libbbduk_b200.so
python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -2 >> gpurun_out/ab.txt
cat gpurun_out/ab.txt
