#!/bin/bash
# usage: ncu_export.sh <report.ncu-rep> <out-prefix>  -> <out-prefix>_details.txt, <out-prefix>_raw.csv (selected metrics), <out-prefix>_opmix.txt
rep=$1; out=$2
ncu -i $rep --page details > ${out}_details.txt 2>/dev/null
ncu -i $rep --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h,u=rows[0],rows[1]
keep=('gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput','gpu__dram_throughput','sm__pipe_tensor','sm__inst_executed_pipe','sm__warps_active','launch__','smsp__inst_executed.sum','smsp__issue_active','sm__throughput','lts__t_bytes.sum','l1tex__t_bytes.sum','lts__throughput','smsp__warp_issue_stalled','sm__cycles_active.avg','tcgen05','sm__ops_path_tensor')
w=csv.writer(sys.stdout)
for k,r in enumerate(rows[2:]):
    for i,n in enumerate(h):
        if any(n.startswith(p) or p in n for p in keep) and ('pct' in n or n.endswith('.sum') or n.startswith('launch__') or 'per_cycle' in n or 'avg' in n) and not n.endswith('peak_sustained') :
            w.writerow([k,n,u[i],r[i]])
" > ${out}_raw.csv
ncu -i $rep --page source --csv 2>/dev/null | python3 -c "
import csv,sys,re,collections
rows=list(csv.reader(sys.stdin))
hdr=rows[1]; si,ii,st=hdr.index('Source'),hdr.index('Instructions Executed'),hdr.index('Warp Stall Sampling (All Samples)')
ops=collections.Counter(); stall=collections.Counter(); tot=0
for r in rows[2:]:
    try: n=int(r[ii])
    except: continue
    m=re.match(r'\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[si]); op=(m.group(1) if m else r[si][:16]).split('.')[0]
    ops[op]+=n; tot+=n
    try: stall[op]+=int(r[st])
    except: pass
print('total warp-instructions', tot)
for k,v in ops.most_common(24): print('%-10s %14d %5.1f%%  stall-samples %d'%(k,v,100.0*v/max(tot,1),stall[k]))
" > ${out}_opmix.txt
