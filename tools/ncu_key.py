"""Key counters of one kernel from `ncu --page raw --csv` + top source lines / opcode mix from `--page source --csv`.
usage: python tools/ncu_key.py raw.csv [source.csv]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d['Kernel Name'][:60])
    for k in want:
        if k in d: print('   ', k, d[k], rows[1][hdr.index(k)])
    st = sorted(((float(d[k].replace(',', '') or 0), k.split('stalled_')[1].split('_per')[0]) for k in hdr
                 if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k), reverse=True)[:7]
    print('    stalls/issue:', ' '.join(f'{b}={a:.2f}' for a, b in st))
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2], errors='replace')))
    hdr = next(r for r in rows if 'Source' in r and 'Instructions Executed' in r)
    ist = hdr.index("Warp Stall Sampling (All Samples)"); iex = hdr.index("Instructions Executed")
    byline = collections.defaultdict(lambda: [0, 0]); src = {}; ops = collections.Counter(); tot = [0, 0]; cur = None
    for r in rows:
        if len(r) <= iex or r is hdr: continue
        if r[0].isdigit():
            cur = int(r[0]); src[cur] = r[1]; continue      # per-source-line aggregate row: skip, use the SASS rows
        if cur is None or not r[2].startswith('0x'): continue
        try: s_ = float(r[ist]); e = float(r[iex])
        except ValueError: continue
        byline[cur][0] += s_; byline[cur][1] += e; tot[0] += s_; tot[1] += e
        t = r[3].split(); op = t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else '')
        ops[op.split('.')[0]] += e
    print('total samples %d, warp instructions %d' % tuple(tot))
    for ln, (s_, e) in sorted(byline.items(), key=lambda kv: -kv[1][0])[:22]:
        print(f"{ln:4d} stall {100*s_/max(tot[0],1):5.1f}% inst {100*e/max(tot[1],1):5.1f}%  {src.get(ln,'')[:120]}")
    print(' '.join(f'{o}={100*c/max(tot[1],1):.1f}%' for o, c in ops.most_common(16)))
