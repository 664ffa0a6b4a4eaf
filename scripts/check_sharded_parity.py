"""Trial-sharded loglik / loglik+grad on N GPUs (NCCL) vs the same evaluation unsharded on rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_sharded_parity.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import faulthandler
    faulthandler.dump_traceback_later(150, exit=True)
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from helpers import engine_from_oracle
    from oracle import synth
    worst = 0.0
    for vec_noise in (False, True):
        x, t = synth.geometry_1d(24, 120, ms_grid=True)
        rng = np.random.default_rng(7)
        sig = 1e-2 * np.exp(0.3 * rng.standard_normal(24)) if vec_noise else 1e-2
        om = synth.model_1d(x, t, a=-200.0, b=2600.0, sig2n=sig)
        lfp = synth.matched_lfp(om, 301, 11)
        sharded, hp = engine_from_oracle(om, lfp, group=True)
        ll_s, g_s = sharded.loglik_grad(hp)
        l_s = sharded.loglik(hp)
        if rank == 0:
            full, _ = engine_from_oracle(om, lfp)
            ll_f, g_f = full.loglik_grad(hp)
            l_f = full.loglik(hp)
            e1, e2 = abs(ll_s - ll_f) / abs(ll_f), float(np.max(np.abs(g_s - g_f) / np.maximum(np.abs(g_f), 1e-300)))
            e3 = abs(l_s - l_f) / abs(l_f)
            print("vector noise %s: world %d  loglik+grad rel err %.1e / %.1e, loglik %.1e" % (vec_noise, dist.get_world_size(), e1, e2, e3))
            worst = max(worst, e1, e2, e3)
    dist.barrier()
    if rank == 0:
        assert worst < 1e-9, worst
        print("OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
