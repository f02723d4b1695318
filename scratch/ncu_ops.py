import csv,re,collections,subprocess,sys
out=subprocess.run(["ncu","-i",sys.argv[1],"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=None; k=0; data=[]
for r in rows:
    if r and r[0]=="Kernel Name":
        k+=1; continue
    if r and r[0]=="Address": hdr=r; continue
    if k==1 and hdr and len(r)==len(hdr): data.append(r)
iS=hdr.index("Source"); iN=hdr.index("# Samples"); iE=hdr.index("Instructions Executed")
tot=sum(int(r[iN]) for r in data); totE=sum(int(r[iE]) for r in data)
print("instrs",len(data),"samples",tot,"executed",totE)
by=collections.Counter(); ex=collections.Counter()
for r in data:
    t=r[iS].split()
    op=t[1] if t[0].startswith('@') else t[0]
    op=op.split('.')[0]
    by[op]+=int(r[iN]); ex[op]+=int(r[iE])
print("  ".join("%s s%.0f%% x%.0f%%"%(op,100*c/tot,100*ex[op]/totE) for op,c in by.most_common(12)))
print("  ".join("%s %.0f%%"%(n[6:],100*sum(int(r[hdr.index(n)]) for r in data)/tot) for n in ["stall_barrier","stall_math","stall_wait","stall_short_sb","stall_long_sb","stall_mio","stall_not_selected","stall_selected","stall_dispatch","stall_lg","stall_no_inst","stall_branch_resolving"]))
