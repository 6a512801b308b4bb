"""Folds per thread and column of every warp of fwbw_kernel (forward: by self slot; backward: by path type), and a pairing of
backward items (lo range, rho) of complementary cost (profiles/r2_training_kernels.md, what comes next)."""
def code(t,k):
    j=t+512*k; f=k&1
    tb=(t&255)<<4; ob=(t+512*f)<<2
    oIn=(ob>>4)==(tb>>4); oBef=(not oIn) and ob<tb
    sInT=(j>>4)==(tb>>4); sInO=(j>>2)==(ob>>2)
    if oIn or sInT or sInO: return 0
    pos=(1 if j>tb else 0)+(1 if j>ob else 0)
    return 1 if pos==2 else ((2 if oBef else 3) if pos==1 else 4)
cost={0:23,1:1,2:17,3:5,4:20}
tot=[]
for w in range(16):
    c=2*19  # main chains
    for k in range(8):
        cs=set(code(32*w+l,k) for l in range(32))
        c+=cost[cs.pop()] if len(cs)==1 else 23
    tot.append(c)
print(tot, sum(tot)/16, max(tot))
# forward
fw=[]
for s in range(16):
    c=s>>2
    pre=sum(2 if (x&3)==c else 1 for x in range(s))
    suf=sum(2 if (x&3)==c else 1 for x in range(s+1,16))
    selfc=5 if (s&3)==c else 2
    fw.append(2*pre+8*(suf+selfc))
print(fw, sum(fw)/16, max(fw))
print("---- balanced backward: thread t: family A=(lo, rho=p), family B=(255-lo, rho=p+2)")
def codej(j, lo, rho):
    tb=lo<<4; ob=(lo+256*rho)<<2   # (j&1023)<<2 with j&1023 = lo+256*rho
    oIn=(ob>>4)==(tb>>4); oBef=(not oIn) and ob<tb
    sInT=(j>>4)==(tb>>4); sInO=(j>>2)==(ob>>2)
    if oIn or sInT or sInO: return 0
    pos=(1 if j>tb else 0)+(1 if j>ob else 0)
    return 1 if pos==2 else ((2 if oBef else 3) if pos==1 else 4)
tot=[]
for w in range(16):
    p=w>>3; c=2*19
    for fam in range(2):
        for q in range(4):
            cs=set()
            for l in range(32):
                lo=(32*(w&7)+l) if fam==0 else 255-(32*(w&7)+l)
                rho=p if fam==0 else p+2
                j=lo+256*(rho+4*q)
                cs.add(codej(j,lo,rho))
            c+=cost[cs.pop()] if len(cs)==1 else 23
    tot.append(c)
print(tot,sum(tot)/16,max(tot))
print("---- item costs (lo-range r, rho)")
items=[]
for r in range(8):
    for rho in range(4):
        c=19
        for q in range(4):
            cs=set(codej((32*r+l)+256*(rho+4*q),32*r+l,rho) for l in range(32))
            c+=cost[cs.pop()] if len(cs)==1 else 23
        items.append((c,r,rho))
items.sort()
print(items)
pairs=[(items[i],items[31-i]) for i in range(16)]
print([a[0]+b[0] for a,b in pairs])
