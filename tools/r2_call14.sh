#!/bin/bash
# Round 2, GPU call 14: decoder cross-attention in-graph (few-query kernels vs tcgen05), S-first issue order (_sfirst build) on the
# T5 attention kernels, ncu of the T5 attention kernels on the representative bias.
set -u
O=gpurun_out
mkdir -p $O
( MRB_ATTN_BENCH_DROP=1 timeout 300 python tools/attn_bench.py cross ) > $O/c14_cross.log 2>&1
cat $O/c14_cross.log | cut -c1-120
for v in "" _sfirst "" _sfirst; do
  ( MRB_LIB_VARIANT=$v MRB_ATTN_BENCH_DROP=1 timeout 200 python tools/attn_bench.py t5enc tc ) > $O/c14_attn_bench$v.log 2>&1
  echo "variant [$v]"; cat $O/c14_attn_bench$v.log | cut -c1-120
done
( MRB_LIB_VARIANT=_sfirst timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -3 ) > $O/c14_pytest_sfirst.log 2>&1
tail -2 $O/c14_pytest_sfirst.log
( timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_.*tc -s 3 -c 3 -o $O/c14_ncu_attn_t5 -f python tools/attn_one.py ) > $O/c14_ncu_attn.log 2>&1
tail -2 $O/c14_ncu_attn.log
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c14_cross_kernels.csv python tools/attn_bench.py cross ) > $O/c14_cross_ncu.log 2>&1
python - <<'PY'
import csv,re,collections
rows=[]
with open('gpurun_out/c14_cross_kernels.csv') as f:
    for line in f:
        if line.startswith('"ID"'): break
    for r in csv.reader(f):
        if len(r)>=15: rows.append((re.sub(r'\(.*','',r[4]).replace('void ',''), r[8], float(r[14])/1e3))
agg=collections.defaultdict(list)
for n,g,t in rows: agg[(n[:60],g)].append(t)
for k,v in sorted(agg.items(), key=lambda x:-sum(x[1]))[:16]: print('%-62s %-14s n=%4d avg=%7.1f us'%(k[0],k[1],len(v),sum(v)/len(v)))
PY
