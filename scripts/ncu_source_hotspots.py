import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
want=sys.argv[2] if len(sys.argv)>2 else "ae_bwd_f2_kernel<(int)0"
# sections: "File Path" row, "Function Name" row, header, data...
i=0; tot=collections.Counter(); per=collections.defaultdict(collections.Counter); src={}; inst=collections.Counter()
fn=None; fp=None; hdr=None; ix=None
for r in rows:
    if not r: continue
    if r[0]=="File Path": fp=r[1].split('/')[-1]; continue
    if r[0]=="Function Name": fn=r[1]; continue
    if r[0]=="Line No": hdr=r; ix={h:i for i,h in enumerate(hdr)}; stallcols=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]; continue
    if fn is None or want not in fn: continue
    if r[0] in ('','-'): continue
    try: ln=int(r[0])
    except: continue
    extra=len(r)-len(hdr)
    if extra>0: r=[r[0],','.join(r[1:2+extra])]+r[2+extra:]
    key=(fp,ln); src[key]=r[1]
    v0=r[ix["# Samples"]]; n=int(v0) if v0 not in ("","-") else 0; tot[key]+=n
    v1=r[ix["Instructions Executed"]]; inst[key]+=int(v1) if v1 not in ("","-") else 0
    for c in stallcols:
        v=r[ix[c]]
        if v and v!='-': per[key][c]+=int(v)
T=sum(tot.values())
print("total samples",T,"total inst",sum(inst.values()))
allst=collections.Counter()
for k in per: allst.update(per[k])
print({c[6:]:v for c,v in allst.most_common(8)})
for key,n in tot.most_common(int(sys.argv[3]) if len(sys.argv)>3 else 30):
    top=", ".join(f"{c[6:]}:{v}" for c,v in per[key].most_common(3))
    print(f"{key[0][:14]}:{key[1]:4d} {100*n/T:5.1f}% inst={inst[key]:9d} [{top}] {src[key].strip()[:80]}")
