P='import json,sys; d=json.loads(sys.stdin.read()); print({k:round(v["ms_median"],3) for k,v in d.items() if isinstance(v,dict)})'
for v in u4 u2; do echo $v; BBDUK_B200_LIB=$PWD/bbtools_b200/csrc/build/ab/lib$v.so python tools/time_kmer_block.py --check 2000 2>&1 | tail -1 | python -c "$P"; done
echo cur; python tools/time_kmer_block.py --check 0 2>&1 | tail -1 | python -c "$P"
