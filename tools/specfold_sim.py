"""Simulation of fold32 (csrc/nc_fwbw.cu): blocks of 32 terms folded by speculating the table indices from the running value at the
start of the block, verified against the sequential p7_FLogsum chain of the oracle; prints rounds needed and a cycle model.
Needs oracle/libnc_oracle.so (test infrastructure)."""
import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import oracle_lib
from nanocall_b200 import models, synth
port = oracle_lib.port()
tbl = port.flogsum_table()
tblp = np.concatenate([tbl[:15999], np.zeros(1, np.float32)]).astype(np.float32)
f32=np.float32
def flogsum(a,b):
    mx=max(a,b); mn=min(a,b)
    if mn==-np.inf or f32(mx-mn)>=f32(15.999): return f32(mx)
    return f32(mx+tbl[int(f32(f32(mx-mn)*f32(1000.)))])
def idx_of(acc,x):
    d=np.abs((acc-x).astype(np.float32)) if isinstance(x,np.ndarray) else abs(f32(acc-x))
    d=np.minimum(d,f32(15.999))
    d=np.where(np.isnan(d),f32(15.999),d)
    return (d*f32(1000.)).astype(np.float32).astype(np.int32)
table = models.builtin_model("r73.t.006.ont.model")["table"]
rng=np.random.default_rng(5)
true=synth.random_params(rng,1)[0]
rd=synth.make_read(rng,table,100,tuple(true))
kmers=port.st_train_kmers()
def terms_for(pm, st=(0.1,0.3)):
    fb=port.fwbw(table,np.array(pm,np.float32),st[0],st[1],rd["mean"],rd["stdv"],rd["start"])
    al,be,lz=fb["alpha"],fb["beta"],fb["log_pr_data"]
    return (al[:-1][:,kmers]+be[:-1][:,kmers]-lz).astype(np.float32)  # denom terms approx (order of ops differs slightly; fine for stats)
def sim(terms, maxit=3):
    acc=f32(-np.inf); stats=dict(blocks=0,dead=0,it=[0]*6,seq=0,live=0,cyc=0)
    T=terms.reshape(-1)
    n=len(T)//32*32
    ref=f32(-np.inf)
    for b in range(0,n,32):
        x=T[b:b+32]
        # reference sequential
        r=acc
        for v in x: r=flogsum(r,v)
        stats['blocks']+=1
        live=~((x==-np.inf)|((acc>x)&((acc-x).astype(np.float32)>=f32(15.999))))
        L=int(live.sum())
        if L==0:
            stats['dead']+=1; stats['cyc']+=40; assert r==acc; continue
        stats['live']+=L
        xs=x[live]
        ok=False
        if acc>-np.inf and np.all(xs<acc):
            idx=idx_of(acc,xs)
            for it in range(1,maxit+1):
                t=tblp[idx]
                a=np.empty(L+1,np.float32); a[0]=acc
                for k in range(L): a[k+1]=f32(a[k]+t[k])
                idx2=idx_of(a[:L],xs)   # elementwise with each a_k
                stats['cyc']+=40+4*L+40
                if np.array_equal(idx2,idx):
                    ok=True; stats['it'][it]+=1; res=a[L]; break
                idx=idx2
        if not ok:
            stats['seq']+=1; stats['cyc']+=50*L
            res=r
        assert res==r, (res,r)
        acc=r
    return stats
for name,pm in (("true",true),("init",[1.0,0.0,0,1.0,1,1]),("off",[true[0]*1.03,true[1]+1.0,0,true[3]*1.2,true[4],true[5]])):
    t=terms_for(pm)
    s=sim(t)
    ev=t.shape[0]
    print(name, s, "cycles/event", s['cyc']/ev, "live/event", s['live']/ev)
