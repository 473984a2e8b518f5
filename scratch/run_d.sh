rm -f gpurun_out/drain.txt
for d in 32 24 16 8 4 1; do
echo "== drain_min $d" >> gpurun_out/drain.txt
BBDUK_B200_DRAIN_MIN=$d python scratch/decomp.py 2>&1 | tail -5 | head -4 >> gpurun_out/drain.txt
done
cat gpurun_out/drain.txt
