#!/bin/bash
# per-kernel durations + DRAM bytes + instruction counts of one long-conv fwd+bwd at the headline shape (256 rows)
mkdir -p gpurun_out
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
  -k regex:'k_row_conv|k_col|k_shortconv' -s 44 -c 22 --csv --log-file gpurun_out/conv_kernels.csv python tools/prof_conv.py 1000000 256 1 bf16 1 > gpurun_out/conv_kernels.log 2>&1
grep -v "^==" gpurun_out/conv_kernels.csv | python -c "
import csv,sys
rows=list(csv.DictReader(sys.stdin))
agg={}
for r in rows:
    k=(int(r['ID']),r['Kernel Name'][:52]); agg.setdefault(k,{})[r['Metric Name']]=float(r['Metric Value'].replace(',',''))
tot=0
for k,v in sorted(agg.items()):
    t=v['gpu__time_duration.sum']/1e3; tot+=t
    print(f\"{k[0]:3d} {k[1]:52s} {t:8.1f} us  rd {v['dram__bytes_read.sum']/1e9:5.2f} wr {v['dram__bytes_write.sum']/1e9:5.2f} GB  {(v['dram__bytes_read.sum']+v['dram__bytes_write.sum'])/t/1e6:6.2f} TB/s  inst {v['smsp__inst_executed.sum']/1e6:7.1f}M issue {v['smsp__issue_active.avg.pct_of_peak_sustained_active']:5.1f}%\")
print('total us', tot)
" | tee gpurun_out/conv_kernels.txt
