#!/bin/bash
# does the A->B->C scratch hand-off hit the L2 when the row groups are small?  DRAM bytes per phase kernel with the
# caches left alone between kernels (ncu --cache-control none), group budgets 16 / 32 / 64 MB vs 1 GB
mkdir -p gpurun_out
for mb in 16 32 64 1024; do
  HY_L2_MB=$mb timeout 600 ncu --cache-control none --clock-control none \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
    -k regex:'k_row_conv|k_col' -s 0 -c 12 --csv --log-file gpurun_out/l2probe_spec_$mb.csv python tools/prof_conv.py 1000000 128 1 bf16 1 > gpurun_out/l2probe_$mb.log 2>&1
  # forward kernels: skip the spectrum launches of the first run() (2 per group)
  ng=$(( (128 * 8 + mb - 1) / mb )); if [ $ng -lt 1 ]; then ng=1; fi
  HY_L2_MB=$mb timeout 600 ncu --cache-control none --clock-control none \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
    -k regex:'k_row_conv|k_col' -s $(( 2 * ng )) -c 12 --csv --log-file gpurun_out/l2probe_fwd_$mb.csv python tools/prof_conv.py 1000000 128 1 bf16 1 >> gpurun_out/l2probe_$mb.log 2>&1
done
for f in gpurun_out/l2probe_*.csv; do echo "== $f"; grep -v "^==" $f | python -c "
import csv,sys
rows=list(csv.DictReader(sys.stdin))
agg={}
for r in rows:
    k=(r['ID'],r['Kernel Name'][:40]); agg.setdefault(k,{})[r['Metric Name']]=r['Metric Value']+' '+r['Metric Unit']
for k,v in agg.items(): print(k, v)
"; done
