"""Minimum number of p7_FLogsum folds per column when chains that start with the same (state, weight) prefix share it:
a trie over the ordered predecessor (forward) / successor (backward) lists of all 4096 states.  (profiles/r2_training_kernels.md)"""
import numpy as np
S=4096
def mask(i,j):
    m = 1 if i==j else 0
    for l in range(1,6):
        k=6-l
        if (i & ((1<<(2*k))-1)) == (j >> (2*(6-k))): m |= 1<<l
    return m
def preds(j):
    s={j}
    for b in range(4): s.add((b<<10)|(j>>2))
    for bb in range(16): s.add((bb<<8)|(j>>4))
    return sorted(s)
def succs(j):
    s={j}
    for b in range(4): s.add(((j&1023)<<2)|b)
    for bb in range(16): s.add(((j&255)<<4)|bb)
    return sorted(s)
def trie_count(lists):
    seen=set()
    tot=0
    for L in lists:
        pre=()
        for x in L:
            pre=pre+(x,)
            if pre not in seen:
                seen.add(pre); tot+=1
    return tot
fw=[[(p,mask(p,j)) for p in preds(j)] for j in range(S)]
bw=[[(v,mask(j,v)) for v in succs(j)] for j in range(S)]
print("edges fw", sum(map(len,fw)), "trie fw", trie_count(fw))
print("edges bw", sum(map(len,bw)), "trie bw", trie_count(bw))
