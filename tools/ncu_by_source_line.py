import csv, sys, subprocess
rep=sys.argv[1]
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
cur=None; hdr=None; out=[]
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>=2 and r[0]=='Line No': hdr=r; ix={}; 
    if hdr is not None and r is hdr:
        for i,h in enumerate(hdr):
            if h not in ix: ix[h]=i
        continue
    if hdr and r and r[0].isdigit():
        try:
            out.append((cur,int(r[0]),r[1].strip()[:90],int(r[ix['Instructions Executed']]),int(r[ix['# Samples']]), int(r[ix['Thread Instructions Executed']])))
        except Exception as e: pass
tot=sum(o[3] for o in out); ts=sum(o[4] for o in out)
print('total warp insts %.2f G'%(tot/1e9))
for o in sorted(out,key=lambda x:-x[3])[:int(sys.argv[2]) if len(sys.argv)>2 else 40]:
    print(f'{100*o[3]/tot:5.1f}% inst {100*o[4]/ts:5.1f}% smpl  thr/inst {o[5]/max(o[3],1):5.1f}  {o[0]}:{o[1]}  {o[2]}')
