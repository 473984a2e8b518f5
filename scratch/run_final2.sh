python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/gpu_tests_n.txt
python bench.py > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_n.txt 2>&1
cat gpurun_out/gpu_tests_n.txt gpurun_out/smoke_n.txt; tail -2 gpurun_out/bench_s.err; python -c "
import json; d=json.load(open('gpurun_out/bench_s.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['config']['chain_e2e_kmer_tbo_qtrim']['reads_per_s'], d['config']['parity_vs_oracle_on_timed_batch'])"
