import csv, sys, subprocess
from collections import Counter
def summarize(rep):
    out = subprocess.run(['ncu','-i',rep,'--page','details'],capture_output=True,text=True).stdout
    keys=["Duration","Registers Per","Theoretical Occ","Achieved Occ","Executed Ipc Active","Issue Slots Busy","L1/TEX Hit","L2 Hit","Warp Cycles Per Issued","Avg. Active Threads","Avg. Not Predicated","Branch Eff","No Eligible","Eligible Warps","Local Load","Local Store"]
    for l in out.splitlines():
        if any(k in l for k in keys): print('   ', ' '.join(l.split()))
    src = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
    rows=list(csv.reader(src.splitlines()))
    hdr=rows[1]; data=rows[2:]; ix={h:i for i,h in enumerate(hdr)}
    stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot=Counter()
    for r in data:
        for s in stalls:
            try: tot[s]+=int(r[ix[s]])
            except: pass
    T=sum(tot.values())
    print('    stalls:', ', '.join(f'{s[6:]} {100*v/T:.1f}%' for s,v in tot.most_common(9)))
    inst=sum(int(r[ix['Instructions Executed']]) for r in data)
    tinst=sum(int(r[ix['Thread Instructions Executed']]) for r in data)
    pinst=sum(int(r[ix['Predicated-On Thread Instructions Executed']]) for r in data)
    print(f'    warp insts {inst/1e9:.2f} G, thread insts {tinst/1e9:.1f} G, pred-on {pinst/1e9:.1f} G')
    c=Counter()
    for r in data:
        t=r[ix['Source']].split()
        op=t[1] if t[0].startswith('@') else t[0]
        c[op.split('.')[0]]+=int(r[ix['Instructions Executed']])
    print('    ops:', ', '.join(f'{op} {100*v/inst:.1f}%' for op,v in c.most_common(16)))
    return inst
for rep in sys.argv[1:]:
    print(rep); summarize(rep)
