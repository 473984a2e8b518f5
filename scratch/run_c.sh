python -m pytest tests/test_tool_gpu.py tests/test_qtrim_gpu.py tests/test_tbo_gpu.py tests/test_entropy_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t_c.txt
python bench.py --cpu-pairs 0 --steps 5 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
cat gpurun_out/t_c.txt; python -c "
import json; d=json.load(open('gpurun_out/bench_c.json')); print(d['value'], d['e2e']['value'], d['config']['chain_e2e_kmer_tbo_qtrim'])"
