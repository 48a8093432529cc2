import csv, collections, re, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
idx_k = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__cycles_active.avg','sass__inst_executed_local_loads','sass__inst_executed_local_stores']
want += [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h]
ix={h:i for i,h in enumerate(hdr)}
sel=[r for r in rows[2:] if kern in r[ix['Kernel Name']]]
r=sel[idx_k]
print(r[ix['Kernel Name']])
for w in want:
    if w in ix:
        v=r[ix[w]]
        try:
            if float(v.replace(',',''))==0: continue
        except: pass
        print(f"  {w.replace('smsp__average_warps_issue_stalled_','stall:').replace('_per_issue_active.ratio',''):70s} {v[:18]:>18s} {units[ix[w]]}")
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name","regex:"+kern,"--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines())); hdr=rows[1]; data=[]
for r in rows[2:]:
    if len(r)!=len(hdr) or r[0]=='Address': break
    data.append(r)
ia=hdr.index('Instructions Executed'); isrc=hdr.index('Source'); isamp=hdr.index('# Samples')
st=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot=sum(int(r[ia]) for r in data); print("total warp instr", tot, "static", len(data))
seg=0; segs=collections.defaultdict(lambda:[0,0,collections.Counter(),collections.Counter()])
for r in data:
    if 'BAR.SYNC' in r[isrc]: seg+=1
    segs[seg][0]+=int(r[ia]); segs[seg][1]+=int(r[isamp])
    for h in st: segs[seg][2][h]+=int(r[hdr.index(h)])
    m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[isrc]); o=m.group(2).split('.')[0] if m else '?'
    segs[seg][3][o]+=int(r[ia])
for s,(n,sm,stc,ops) in segs.items(): print(f" seg{s}: {n:11d} {100*n/tot:5.1f}% instr; samples {sm:6d}", dict(stc.most_common(4)), dict(ops.most_common(5)))
top=sorted(data,key=lambda r:-int(r[isamp]))[:14]
for r in top: print("  ", r[isamp], r[ia], r[isrc][:80], {h:r[hdr.index(h)] for h in st if int(r[hdr.index(h)])>100})
