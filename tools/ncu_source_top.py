import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
secs=[i for i,r in enumerate(rows) if r and r[0]=="Kernel Name"]
secs.append(len(rows))
which=[int(a) for a in sys.argv[2:]] or range(len(secs)-1)
for si in which:
    a,b=secs[si],secs[si+1]
    hdr=rows[a+1]; body=[r for r in rows[a+2:b] if len(r)>=len(hdr)-2]
    ci={h:i for i,h in enumerate(hdr)}
    ns=ci["# Samples"]; 
    tot=sum(int(r[ns] or 0) for r in body)
    print("=== section",si,rows[a][1][:60],"lines",len(body),"samples",tot)
    stalls=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg={h:sum(int(r[ci[h]] or 0) for r in body) for h in stalls}
    print("  ",{k:v for k,v in sorted(agg.items(),key=lambda kv:-kv[1])[:8]})
    top=sorted(body,key=lambda r:-int(r[ns] or 0))[:22]
    for r in top:
        st={h:int(r[ci[h]] or 0) for h in stalls}
        top2=sorted(st.items(),key=lambda kv:-kv[1])[:2]
        print(f"  {int(r[ns]):6d} {r[ci['Address']][-6:] if 'Address' in ci else ''} {r[ci['Source']][:110]} | {top2}")
