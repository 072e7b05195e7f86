import sys, numpy as np
d=np.loadtxt(sys.argv[1],dtype=np.int64)
# slot-pass duration in steady state: difference between slot0 and slot1 wait starts in stages 2-4
it0=d[d[:,0]==1]
vals=[]
for st in (2,3,4):
    a=it0[(it0[:,1]==st)&(it0[:,2]==0)][0][3]; b=it0[(it0[:,1]==st)&(it0[:,2]==1)][0][3]
    c=it0[(it0[:,1]==st+1)&(it0[:,2]==0)][0][3]
    vals.append((b-a, c-b))
print(sys.argv[1], "slot passes (cycles):", vals, " -> per MMA %.1f cycles" % (np.mean(vals)/16))
tile=d[(d[:,0]==2)&(d[:,1]==0)&(d[:,2]==0)][0][3]-d[(d[:,0]==1)&(d[:,1]==0)&(d[:,2]==0)][0][3]
print("   tile period", tile, "cycles (ideal 10 stages: %d)"%(2*(16*4+4+20+18)*128))
