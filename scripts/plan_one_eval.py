"""A few evaluations of one configuration through the native plan in eager mode (for ncu launch lists)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "plan_timing.py")).read().split("for name in sys.argv")[0])
name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
eng, hp = case(name)
eng.loglik_grad(hp())
list(eng._plans.values())[0].set_graph(False)
for _ in range(3):
    eng.loglik_grad(hp())
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.loglik_grad(hp())
torch.cuda.synchronize()
torch.cuda.profiler.stop()
