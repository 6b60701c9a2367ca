#!/bin/bash
# Turn gpurun_out/*.ncu-rep / launches csv into the tracked summaries under profiles/ (run here, no GPU)
set -e
R=${1:-r01}
REP=gpurun_out/prof_step.ncu-rep
mkdir -p profiles
ncu -i $REP --page details --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h={n:i for i,n in enumerate(rows[0])}
keep=('Duration','Elapsed Cycles','SM Frequency','DRAM Throughput','Memory Throughput','L1/TEX Hit Rate','L2 Hit Rate','Executed Ipc Active','Issue Slots Busy','No Eligible','Active Warps Per Scheduler','Warp Cycles Per Issued Instruction','Registers Per Thread','Achieved Active Warps Per SM','Theoretical Active Warps per SM','Block Size','Grid Size','Dynamic Shared Memory Per Block','Static Shared Memory Per Block','Issued Instructions')
print('kernel_id,kernel,section,metric,unit,value')
for r in rows[1:]:
    if r[h['Metric Name']] in keep: print(','.join([r[h['ID']],'\"'+r[h['Kernel Name']][:60]+'\"',r[h['Section Name']],r[h['Metric Name']],r[h['Metric Unit']],r[h['Metric Value']].replace(',','')]))
" > profiles/${R}_step_kernel_details.csv
ncu -i $REP --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]; units=rows[1]
want=[i for i,n in enumerate(hdr) if n in ('ID','Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','smsp__inst_executed.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum')]
w=csv.writer(sys.stdout)
w.writerow([hdr[i] for i in want]); w.writerow([units[i] for i in want])
for r in rows[2:]: w.writerow([r[i][:70] for i in want])
" > profiles/${R}_step_kernel_raw.csv
ncu -i $REP --page source --csv --print-source sass 2>/dev/null > /tmp/src_sass.csv
python3 - <<'PY' > profiles/${R}_step_kernel_stalls.txt
import csv
rows=list(csv.reader(open('/tmp/src_sass.csv')))
name=rows[0][1]
hdr=rows[1]; h={n:i for i,n in enumerate(hdr)}
stalls=[n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
tot={s:0 for s in stalls}; total=0; n=0
for r in rows[2:]:
    if not r or r[0] in ('Kernel Name','Address'): break
    n+=1; total+=int(r[h['# Samples']])
    for s in stalls: tot[s]+=int(r[h[s]])
print('kernel:',name); print('SASS instructions:',n,'  warp-stall samples:',total)
for s,v in sorted(tot.items(), key=lambda x:-x[1]):
    if v: print(f'  {s:28s}{v:8d} {100*v/total:5.1f}%')
PY
cat profiles/${R}_step_kernel_raw.csv; cat profiles/${R}_step_kernel_stalls.txt
