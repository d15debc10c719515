"""Region-level stall profile of an .ncu-rep (read here, no GPU): consecutive SASS instructions with the same execution count\nform a region; prints instructions, executions, share of the stall samples, stall mix and opcode mix per region.\nusage: python tools/ncu_regions.py gpurun_out/prof.ncu-rep"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols = [(i,h) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
regs=[]; cur=None
tot_s=0
for k,rr in enumerate(rows[2:]):
    if rr and rr[0]=='Kernel Name': break
    if len(rr)<len(hdr): continue
    e=int(rr[iE]); s=int(rr[iSm]); tot_s+=s
    if cur is None or abs(e-cur['e'])>0.02*max(e,cur['e'],1):
        cur={'e':e,'n':0,'s':0,'start':k,'st':collections.Counter(),'ops':collections.Counter(),'first':rr[iS].strip()}
        regs.append(cur)
    cur['n']+=1; cur['s']+=s
    op=rr[iS].strip().split()[0]
    if op.startswith('@'): op=rr[iS].strip().split()[1]
    cur['ops'][op.split('.')[0] if not op.startswith('MUFU') else op]+=1
    for i,h in stall_cols:
        v=int(rr[i]) if rr[i].isdigit() else 0
        if v: cur['st'][h[6:]]+=v
print('total samples',tot_s)
for r in regs:
    if r['s']<0.004*tot_s: continue
    st=' '.join(f'{k}:{100*v/r["s"]:.0f}' for k,v in r['st'].most_common(5))
    ops=' '.join(f'{k}:{v}' for k,v in r['ops'].most_common(6))
    print(f"@{r['start']:5d} n={r['n']:4d} exec={r['e']:8d} samples={100*r['s']/tot_s:5.1f}%  [{st}]  {ops}")
