set -x
for c in 2 4 8 20; do BBDUK_B200_CAND_CAP=$c python bench.py --cpu-pairs 0 --steps 6 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('CAND_CAP', $c, d['value'], d['config']['parity_vs_oracle_on_timed_batch'])"; done > gpurun_out/candcap.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01h.csv python bench.py --steps 2 --warmup 1 --cpu-pairs 0 > gpurun_out/b_under_ncu_h.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:tbo_kernel -c 3 -o gpurun_out/prof_tbo4 -f python bench.py --steps 2 --warmup 1 --cpu-pairs 0 > /dev/null 2>&1
python bench.py > gpurun_out/bench_h2.json 2> gpurun_out/bench_h2.err
cat gpurun_out/candcap.txt
