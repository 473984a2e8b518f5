python -m pytest tests/test_tool_gpu.py -m gpu -x -q 2>&1 | tail -2 > gpurun_out/t_c2.txt
python bench.py --cpu-pairs 0 --steps 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
cat gpurun_out/t_c2.txt; tail -3 gpurun_out/bench_c2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print(d['value'], d['e2e']['value'], d['config']['chain_e2e_kmer_tbo_qtrim'])"
