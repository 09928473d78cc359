"""Per-source-line instruction counts, active lanes and stall samples from an ncu report."""
import csv, subprocess, sys

def main(rep, regex, top=30):
    out = subprocess.run(["ncu","-i",rep,"--page","source","--print-source","cuda,sass","--csv","--kernel-name","regex:"+regex],capture_output=True,text=True).stdout
    rows=list(csv.reader(out.splitlines()))
    hdr=None; cur=""; agg={}
    for r in rows:
        if len(r)==2 and r[0]=="File Path": cur=r[1].split("/")[-1]; continue
        if len(r)>4 and r[0]=="Line No": hdr=r; continue
        if hdr is None or len(r)<8 or not r[0]: continue
        ie=hdr.index("Instructions Executed"); te=hdr.index("Thread Instructions Executed")
        try: a=int(r[ie]); t=int(r[te]); st=int(r[4])
        except: continue
        k=(cur,int(r[0]),r[1].strip()[:100])
        x=agg.get(k,[0,0,0]); x[0]+=a; x[1]+=t; x[2]+=st; agg[k]=x
    tot=sum(v[0] for v in agg.values()); tt=sum(v[1] for v in agg.values()); ts=sum(v[2] for v in agg.values())
    print("warp-instr %d  thread-instr %d  avg active lanes %.1f  stall samples %d"%(tot,tt,tt/max(tot,1),ts))
    for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:top]:
        print("%9d %5.1f%% lanes %4.1f stall %5.1f%%  %s:%d %s"%(v[0],100*v[0]/tot,v[1]/max(v[0],1),100*v[2]/max(ts,1),k[0],k[1],k[2]))

if __name__=="__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv)>3 else 30)
