import numpy as np, sys
N=4096
e=np.fromfile(sys.argv[1] if len(sys.argv)>1 else "energies.bin")
x=np.sin(0.001*np.arange(N)*1.0)
bad=0; worst=0
for p in range(2,1025):
    S=np.zeros(p); np.add.at(S, np.arange(N)%p, x)
    cnt=np.bincount(np.arange(N)%p, minlength=p)
    ref=float((S*S/cnt).sum())
    if e[p]==0: 
        bad+=1; 
        if bad<10: print("missing",p)
        continue
    rel=abs(e[p]-ref)/ref; worst=max(worst,rel)
    if rel>1e-11:
        bad+=1
        if bad<10: print("mismatch",p,e[p],ref)
print("bad",bad,"worst rel",worst)
