set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/gpu_tests_m.txt
python bench.py > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01m.csv python bench.py --steps 2 --warmup 1 --cpu-pairs 0 > gpurun_out/b_under_ncu_m.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:bbduk_fast_kernel -s 4 -c 1 -o gpurun_out/prof_fast_m2 -f python bench.py --steps 2 --warmup 1 --cpu-pairs 0 --pairs-per-step 2097152 > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_m.txt 2>&1
cat gpurun_out/gpu_tests_m.txt gpurun_out/smoke_m.txt; head -c 600 gpurun_out/bench_r.json
