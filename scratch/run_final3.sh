set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/gpu_tests_o.txt
python bench.py > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01o.csv python bench.py --steps 2 --warmup 1 --cpu-pairs 0 > gpurun_out/b_under_ncu_o.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:bbduk_fast_kernel -s 4 -c 1 -o gpurun_out/prof_fast_o -f python bench.py --steps 2 --warmup 1 --cpu-pairs 0 --pairs-per-step 2097152 > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_o.txt 2>&1
cat gpurun_out/gpu_tests_o.txt gpurun_out/smoke_o.txt; tail -2 gpurun_out/bench_t.err; python -c "
import json; d=json.load(open('gpurun_out/bench_t.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['config']['chain_e2e_kmer_tbo_qtrim']['reads_per_s'], d['config']['kmer_block_plus_tbo'], d['config']['parity_vs_oracle_on_timed_batch'])"
