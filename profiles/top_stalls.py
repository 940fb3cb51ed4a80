import csv,sys,subprocess
rep=sys.argv[1]; thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.01
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[1]
ia=hdr.index("Address"); isrc=hdr.index("Source"); ins=hdr.index("# Samples")
names=["stall_long_sb","stall_barrier","stall_math","stall_wait","stall_short_sb","stall_membar","stall_branch_resolving","stall_no_inst","stall_dispatch","stall_mio","stall_lg","stall_not_selected","stall_selected"]
idx={n:hdr.index(n) for n in names}
tot=sum(int(r[ins] or 0) for r in rows[2:])
agg={n:sum(int(r[idx[n]] or 0) for r in rows[2:]) for n in names}
print("total samples",tot, {k:round(100*v/tot,1) for k,v in agg.items() if v>tot*0.005})
for r in rows[2:]:
    n=int(r[ins] or 0)
    if n>=tot*thr:
        print(r[ia][-5:], f"{n:6d} {100*n/tot:5.1f}% ", " ".join(f"{k[6:]}={r[idx[k]]}" for k in names if int(r[idx[k]] or 0)>n*0.15), " ", r[isrc][:80])
