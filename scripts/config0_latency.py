"""Developer measurement script (run from the repo root on a B200); numbers quoted in profiles/r01d_eigensolver.md / DESIGN.md."""
import sys, time, threading; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
from oracle import synth, gpcsd_oracle as O
from helpers import engine_from_oracle, hp_from_oracle
x,t = synth.geometry_1d(24,50); om = synth.model_1d(x,t); lfp = synth.matched_lfp(om,50,1)
eng,hp = engine_from_oracle(om,lfp)
for _ in range(5): eng.loglik_grad(hp)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(50): eng.loglik_grad(hp)
dt=(time.perf_counter()-t0)/50; print("cfg1 serial loglik+grad: %.3f ms/eval -> %.0f evals/s"%(dt*1e3,1/dt))
t0=time.perf_counter()
for _ in range(50): eng.loglik(hp)
dt=(time.perf_counter()-t0)/50; print("cfg1 serial loglik: %.3f ms/eval"%(dt*1e3))
t0=time.perf_counter(); 
for _ in range(20): O.loglik_and_grad(om,lfp)
print("cpu oracle loglik+grad %.3f ms"%((time.perf_counter()-t0)/20*1e3))
for nth in (2,4,8,16):
    engs=[engine_from_oracle(om,lfp)[0] for _ in range(nth)]
    streams=[torch.cuda.Stream() for _ in range(nth)]
    def work(i, n):
        with torch.cuda.stream(streams[i]):
            for _ in range(n): engs[i].loglik_grad(hp)
    ths=[threading.Thread(target=work,args=(i,3)) for i in range(nth)]; [a.start() for a in ths]; [a.join() for a in ths]
    t0=time.perf_counter()
    ths=[threading.Thread(target=work,args=(i,20)) for i in range(nth)]; [a.start() for a in ths]; [a.join() for a in ths]
    dt=time.perf_counter()-t0; print("threads %d: %.0f evals/s"%(nth, nth*20/dt))
