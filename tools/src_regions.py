import csv, gzip, sys
f = sys.argv[1]; step = int(sys.argv[2]) if len(sys.argv) > 2 else 250
rows=list(csv.reader(gzip.open(f,'rt')))
h=rows[1]
si=h.index('# Samples'); ii=h.index('Instructions Executed'); src=h.index('Source')
data=[(k,int(r[ii]),int(r[si]),r[src].strip()) for k,r in enumerate(rows[2:]) if len(r)>ii and r[ii].isdigit()]
tot=sum(d[1] for d in data); ts=sum(d[2] for d in data)
print(len(data), 'instrs; warp-inst', tot, 'samples', ts)
def op(t):
    w=t.split()
    return (w[1] if w[0].startswith('@') else w[0]).split('.')[0]
for lo in range(0,len(data),step):
    seg=data[lo:lo+step]
    n=sum(d[1] for d in seg); s=sum(d[2] for d in seg)
    fp=sum(d[1] for d in seg if op(d[3]).startswith(('FFMA','FMUL','FADD')))
    if n: print(f"#{lo:5d}-{lo+step:5d}: exec {100*n/tot:5.1f}%  samples {100*s/ts:5.1f}%  fp-share {100*fp/max(n,1):4.0f}%  maxexec {max(d[1] for d in seg)}")
hist={}
for d in data:
    e=hist.setdefault(op(d[3]),[0,0]); e[0]+=d[1]; e[1]+=d[2]
print("opcode exec% samples%")
for o,(n,s) in sorted(hist.items(), key=lambda x:-x[1][0])[:16]: print(f"  {o:8s} {100*n/tot:5.1f} {100*s/ts:5.1f}")
print("top by samples")
for d in sorted(data,key=lambda d:-d[2])[:int(sys.argv[3]) if len(sys.argv)>3 else 20]: print(f"{100*d[2]/ts:5.2f}% exec {d[1]:>9d} #{d[0]:5d} {d[3][:100]}")
