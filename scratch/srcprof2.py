import csv, collections, subprocess, sys
rep=sys.argv[1]; skip=sys.argv[2]; top=int(sys.argv[3]) if len(sys.argv)>3 else 45
txt=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass","--launch-skip",skip,"--launch-count","1"],capture_output=True,text=True).stdout
rows=list(csv.reader(txt.splitlines()))
cur_file=None
tot=collections.Counter(); thr=collections.Counter(); samples=collections.Counter(); src={}
hdr=None
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r[0]=='Line No': hdr=r; continue
    if r[0]=='Function Name': continue
    if r[0] and r[0].isdigit() and hdr:
        try:
            key=(cur_file,int(r[0]))
            ie=int(r[hdr.index('Instructions Executed')]); te=int(r[hdr.index('Thread Instructions Executed')])
            sm=int(r[hdr.index('# Samples')])
        except Exception: continue
        tot[key]+=ie; thr[key]+=te; samples[key]+=sm; src[key]=r[1]
T=sum(tot.values()); S=sum(samples.values())
print("total instr", T, "samples", S)
for key,v in tot.most_common(top):
    print(f"{key[0]:14s}:{key[1]:4d} {100*v/T:5.1f}% inst {100*samples[key]/S:5.1f}% smp  act={thr[key]/max(v,1):4.1f}  {src[key][:90]}")
