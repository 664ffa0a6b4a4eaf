import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0]]
import numpy as np, torch
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "plan_timing.py")).read().split("for name in sys.argv")[0])
for name in ["cfg0", "cfg1"]:
    eng, hp = case(name)
    eng.loglik_grad(hp())
    pl = list(eng._plans.values())[0]
    pl.set_graph(False)
    for _ in range(6):
        eng.loglik_grad(hp())
