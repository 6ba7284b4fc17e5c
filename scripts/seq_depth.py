"""Development aid (CPU): sequence statistics of the bench corpora from the decoder's own records (tests/hostsim dumps
them): lengths, distances, and the depth of the dependency chains inside a step of N consecutive sequences -- what bounds
the CTA-per-frame execute kernel (known limits).   python scripts/seq_depth.py"""
import sys, ctypes
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, corpus, hostsim
def stats(kind, fs, lvl, streamed, n=1):
    c = corpus.build(kind, n, fs, lvl, False, streamed=streamed)
    hostsim.run([], [])
    L = hostsim._lib
    frames=[bytes(c.comp[int(c.offsets[i]):int(c.offsets[i]+c.sizes[i])]) for i in range(n)]
    blob=b"".join(frames); comp=np.frombuffer(blob,dtype=np.uint8).copy()
    sizes=np.array([len(f) for f in frames],dtype=np.uint64); offs=np.zeros(n,np.uint64); offs[1:]=np.cumsum(sizes)[:-1]
    caps=np.full(n,fs,np.uint64); doffs=np.zeros(n,np.uint64); doffs[1:]=np.cumsum(caps)[:-1]
    dst=np.zeros(int(caps.sum())+64,np.uint8); status=np.zeros(n,np.uint32); osz=np.zeros(n,np.uint64)
    cap=fs*n//3+4096; seq=np.zeros(cap,np.uint64)
    L.kzs_set_exec(0,8192)
    r=L.kzs_decompress_batch(comp.ctypes.data,int(sizes.sum()),offs.ctypes.data,sizes.ctypes.data,n,dst.ctypes.data,doffs.ctypes.data,caps.ctypes.data,status.ctypes.data,osz.ctypes.data,seq.ctypes.data,cap,None,0)
    assert r==0 and (status==0).all()
    # records are per block, running totals; reconstruct ll, ml
    rec=seq
    lit_end=((rec>>np.uint64(28))&np.uint64(0x3FFFF)).astype(np.int64); out_end=(rec>>np.uint64(46)).astype(np.int64); off=(rec&np.uint64(0xFFFFFFF)).astype(np.int64)
    nz=np.nonzero(rec)[0]; ns=nz[-1]+1
    lit_end=lit_end[:ns]; out_end=out_end[:ns]; off=off[:ns]
    d_out=np.diff(out_end,prepend=0); newblk=d_out<0
    pl=np.where(newblk,0,np.concatenate([[0],lit_end[:-1]])); po=np.where(newblk,0,np.concatenate([[0],out_end[:-1]]))
    ll=lit_end-pl; ml=out_end-po-ll
    sym=(off&0x8000000)!=0
    print(f"{kind} fs={fs} l{lvl}: seqs={ns} per64K={ns/(n*fs/65536):.0f} avg ll={ll.mean():.2f} ml={ml.mean():.2f} seqbytes={(ll+ml).mean():.1f}  ll==0: {(ll==0).mean():.2f}")
    for q in (50,75,90,95,99): print(f"   p{q}: ml={np.percentile(ml,q):.0f} ll={np.percentile(ll,q):.0f} off={np.percentile(off[~sym],q):.0f}")
    o=off[~sym]
    for t in (464,976,2000,4096,16384,65536,131072,1<<20): print(f"   off>{t}: {(o>t).mean():.3f}", end='')
    print()
    print(f"   ml>16 {(ml>16).mean():.3f} ml>32 {(ml>32).mean():.3f} ml>64 {(ml>64).mean():.3f} ml>256 {(ml>256).mean():.4f}; bytes in ml>64: {ml[ml>64].sum()/ (ll+ml).sum():.3f}")
stats("logs",65536,3,False,8)
stats("json",65536,3,False,8)
stats("mixed",4<<20,19,True,1)

def depth(kind, fs, lvl, streamed, STEP=512):
    import bisect
    n=1
    c = corpus.build(kind, n, fs, lvl, False, streamed=streamed)
    hostsim.run([], [])
    L = hostsim._lib
    frames=[bytes(c.comp[int(c.offsets[i]):int(c.offsets[i]+c.sizes[i])]) for i in range(n)]
    blob=b"".join(frames); comp=np.frombuffer(blob,dtype=np.uint8).copy()
    sizes=np.array([len(f) for f in frames],dtype=np.uint64); offs=np.zeros(n,np.uint64)
    caps=np.full(n,fs,np.uint64); doffs=np.zeros(n,np.uint64)
    dst=np.zeros(int(caps.sum())+64,np.uint8); status=np.zeros(n,np.uint32); osz=np.zeros(n,np.uint64)
    cap=fs*n//3+4096; seq=np.zeros(cap,np.uint64)
    L.kzs_set_exec(0,8192)
    r=L.kzs_decompress_batch(comp.ctypes.data,int(sizes.sum()),offs.ctypes.data,sizes.ctypes.data,n,dst.ctypes.data,doffs.ctypes.data,caps.ctypes.data,status.ctypes.data,osz.ctypes.data,seq.ctypes.data,cap,None,0)
    rec=seq; ns=np.nonzero(rec)[0][-1]+1
    lit_end=((rec>>np.uint64(28))&np.uint64(0x3FFFF)).astype(np.int64)[:ns]; out_end=(rec>>np.uint64(46)).astype(np.int64)[:ns]; off=(rec&np.uint64(0xFFFFFFF)).astype(np.int64)[:ns]
    # split into blocks
    starts=[0]+[i for i in range(1,ns) if out_end[i]<out_end[i-1]]+[ns]
    depths=[]; pend=[]
    for bi in range(len(starts)-1):
        s,e=starts[bi],starts[bi+1]
        oe=out_end[s:e]; le=lit_end[s:e]; of=off[s:e]
        po=np.concatenate([[0],oe[:-1]]); pl=np.concatenate([[0],le[:-1]])
        ll=le-pl; mdst=po+ll; ml=oe-mdst
        for s0 in range(0,e-s,STEP):
            m=min(STEP,e-s-s0); lev=[0]*m; S0=po[s0]; mx=0; np_=0
            oel=list(oe[s0:s0+m])
            for i in range(m):
                gi=s0+i
                o=of[gi]
                if o & 0x8000000: o=200  # symbolic: unknown; assume near
                sp=mdst[gi]-o; need=sp+ml[gi] if o>=ml[gi] else mdst[gi]
                if ml[gi]==0 or need<=S0: continue
                a=max(sp,S0); b=need
                ja=bisect.bisect_right(oel,a,0,i); jb=min(bisect.bisect_left(oel,b,0,i),i-1)
                if ja>jb: continue
                np_+=1
                lv=1+max(lev[ja:jb+1]); lev[i]=lv; mx=max(mx,lv)
            depths.append(mx); pend.append(np_/m)
    d=np.array(depths)
    print(f"{kind} l{lvl} step={STEP}: steps={len(d)} depth mean={d.mean():.1f} p50={np.percentile(d,50):.0f} p90={np.percentile(d,90):.0f} max={d.max()} pending frac={np.mean(pend):.2f}")
depth("mixed",4<<20,19,True,512)
depth("mixed",4<<20,19,True,256)
depth("mixed",4<<20,19,True,1024)
depth("logs",65536,3,False,512)
