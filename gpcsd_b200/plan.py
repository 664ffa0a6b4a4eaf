"""Python handle of the native evaluation plan (csrc/gpcsd_plan.cu, include/gpcsd_b200.h "gpcsd_plan"): ONE C-ABI call per
loglik+gradient evaluation, batched over restarts, replayed from a CUDA graph.

The reference evaluates ``obj_fun`` and ``grad(obj_fun)`` once per L-BFGS iteration per restart from Python
(gpcsd1d.py:193-211); ``EvalPlan.evaluate(thetas)`` is the replacement for R restarts at once.  PyTorch supplies the workspace
allocation, the stream and (trial-sharded models) the all-reduce of the device result; nothing else.
"""
import ctypes

import numpy as np
import torch

from . import _lib as L

F64 = torch.float64
_PD = ctypes.POINTER(ctypes.c_double)
_PI = ctypes.POINTER(ctypes.c_int)


def _pd(a):
    return a.ctypes.data_as(_PD)


class EvalPlan:
    """One plan = one model geometry x one hyperparameter structure (temporal kernel kinds, scalar or per-electrode noise)."""

    def __init__(self, eng, kinds, nsig, eps=0.0, max_restarts=1):
        L.load()
        self.eng = eng
        self.kinds = tuple(int(k) for k in kinds)
        self.nsig = int(nsig)
        self.rmax = int(max_restarts)
        x = np.ascontiguousarray(eng.x_host, dtype=np.float64).reshape(-1)
        t = np.ascontiguousarray(eng.t_host, dtype=np.float64)
        q = eng.quad_host
        if eng.dim == 1:
            g1, w1 = q["gl_x"], q["gl_w"]
            g2 = w2 = np.zeros(1)
            G1, G2 = len(g1), 0
        else:
            g1, w1, g2, w2 = q["gl_x1"], q["gl_w1"], q["gl_x2"], q["gl_w2"]
            G1, G2 = len(g1), len(g2)
        kinds_c = (ctypes.c_int * len(self.kinds))(*self.kinds)
        pairs = eng.s_pairs_host
        ra = rb = None
        if pairs is not None:
            ra = np.ascontiguousarray(pairs[0], dtype=np.int32).ctypes.data_as(_PI)
            rb = np.ascontiguousarray(pairs[1], dtype=np.int32).ctypes.data_as(_PI)
            self._keep = pairs
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(eng.device):
            L.call("gpcsd_plan_create", ctypes.byref(self.handle), eng.dim, eng.nx, eng.nt, _pd(x), _pd(t), G1,
                   _pd(np.ascontiguousarray(g1, dtype=np.float64)), _pd(np.ascontiguousarray(w1, dtype=np.float64)), G2,
                   _pd(np.ascontiguousarray(g2, dtype=np.float64)), _pd(np.ascontiguousarray(w2, dtype=np.float64)),
                   len(self.kinds), kinds_c, self.nsig, float(eng.jitter), float(eps), 1 if eng.t_uniform else 0,
                   int(eng.FOLD_MIN_NT), ra, rb, self.rmax)
        self.P = int(L.load().gpcsd_plan_num_params(self.handle))
        self.outw = self.P + 4
        self.ws = None
        self._bound = None
        self._mailbox = None
        self._mailbox_tried = False
        self.n_replays = 0

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value:
                L.load().gpcsd_plan_destroy(self.handle)
                self.handle = ctypes.c_void_p()
        except Exception:
            pass

    def ws_bytes(self):
        eng = self.eng
        return int(L.query("gpcsd_plan_ws_bytes", self.handle, int(eng.ldn), int(eng.ntrials)))

    def bind(self):
        """(Re)bind the engine's uploaded LFP slab and a workspace; cheap no-op when nothing changed."""
        eng = self.eng
        if eng.Y is None:
            raise RuntimeError("no LFP uploaded: call set_lfp first")
        key = (eng.Y.data_ptr(), eng.ldn, eng.ntrials, float(eng.ntrials_total), eng.shard.det_fraction())
        if self._bound is not None and self._bound[0] == key:
            if self._bound[1] != eng._lfp_version:               # same buffer, new contents
                L.call("gpcsd_plan_touch_lfp", self.handle)
                self._bound = (key, eng._lfp_version)
            return
        nbytes = self.ws_bytes()
        if self.ws is None or self.ws.numel() * 8 < nbytes:
            self.ws = None
            self.ws = torch.empty((nbytes + 7) // 8, dtype=F64, device=eng.device)
        L.call("gpcsd_plan_set_lfp", self.handle, eng.Y.data_ptr(), int(eng.ldn), int(eng.ntrials), float(eng.ntrials_total),
               float(eng.shard.det_fraction()), self.ws.data_ptr(), self.ws.numel() * 8)
        self._bound = (key, eng._lfp_version)

    def setup_mailbox(self):
        """Trial-sharded model with all ranks on one node: create / attach the shared host segment of the result mailbox
        (gpcsd_plan_set_mailbox) so that the per-evaluation all-reduce needs neither NCCL nor a device->host copy.  Collective
        over the shard's process group (every rank must call it).  Returns True when the mailbox is active."""
        import mmap
        import os
        import socket
        import uuid
        import torch.distributed as dist
        sh = self.eng.shard
        if self._mailbox is not None or not (sh.enabled and sh.world > 1) or os.environ.get("GPCSD_MAILBOX", "1") == "0":
            return self._mailbox is not None
        hosts = [None] * sh.world
        dist.all_gather_object(hosts, socket.gethostname(), group=sh.group)
        if len(set(hosts)) != 1:
            return False                                   # multi-node: stay on the NCCL path
        nbytes = int(L.query("gpcsd_plan_mailbox_bytes", self.handle, sh.world))
        box = [None]
        path = None
        if sh.rank == 0:
            path = "/dev/shm/gpcsd_b200_%s" % uuid.uuid4().hex
            fd = os.open(path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
            os.ftruncate(fd, nbytes)                       # zero-filled
            box[0] = path
        src = dist.get_global_rank(sh.group, 0) if sh.group is not None else 0
        dist.broadcast_object_list(box, src=src, group=sh.group)
        if sh.rank != 0:
            fd = os.open(box[0], os.O_RDWR)
        mm = mmap.mmap(fd, nbytes, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
        os.close(fd)
        addr = ctypes.addressof(ctypes.c_char.from_buffer(mm))
        with torch.cuda.device(self.eng.device):
            L.call("gpcsd_plan_set_mailbox", self.handle, addr, nbytes, sh.world, sh.rank)
        dist.barrier(group=sh.group)                       # every rank has mapped the segment: the name can go
        if sh.rank == 0:
            os.unlink(path)
        self._mailbox = mm
        return True

    def device_result(self, R):
        """The plan's device result [R][P+4] as a torch view into the workspace (for the trial-shard all-reduce)."""
        ptr = L.load().gpcsd_plan_device_result(self.handle)
        off = (int(ptr) - self.ws.data_ptr()) // 8
        return self.ws[off: off + R * self.outw]

    def evaluate(self, thetas, want_grad=True):
        """thetas: (R, P) natural-unit hyperparameters -> host array (R, P+4): loglik, gradient, solver flag, checksum pair
        (already all-reduced over the trial shards when the engine is sharded)."""
        eng = self.eng
        thetas = np.ascontiguousarray(thetas, dtype=np.float64).reshape(-1, self.P)
        R = thetas.shape[0]
        if R > self.rmax:
            raise ValueError("plan was created for at most %d restarts" % self.rmax)
        self.bind()
        stream = torch.cuda.current_stream(eng.device)
        if eng._y_ready is not None:
            stream.wait_event(eng._y_ready)
        out = np.empty((R, self.outw), dtype=np.float64)
        sharded = eng.shard.enabled and eng.shard.world > 1
        if sharded and self._mailbox is None and not self._mailbox_tried:
            self._mailbox_tried = True
            self.setup_mailbox()
        if not sharded or self._mailbox is not None:
            # single rank, or the ranks' results meet in the host-mapped mailbox inside gpcsd_plan_finish
            L.call("gpcsd_plan_loglik_grad", self.handle, R, _pd(thetas), 1 if want_grad else 0, _pd(out), stream.cuda_stream)
        else:
            L.call("gpcsd_plan_enqueue", self.handle, R, _pd(thetas), 1 if want_grad else 0, stream.cuda_stream)
            eng.shard.allreduce_inplace(self.device_result(R))
            L.call("gpcsd_plan_finish", self.handle, R, _pd(out), stream.cuda_stream)
        if sharded:
            mean, meansq = out[:, self.P + 2], out[:, self.P + 3]
            if np.any(np.abs(meansq - mean * mean) > 1e-9 * np.maximum(np.abs(meansq), 1e-300)):
                raise RuntimeError("trial-sharded evaluation: the ranks hold different hyperparameters; seed numpy identically "
                                   "on every rank or broadcast the parameters")
        self.n_replays += 1
        return out

    def evaluate_with_factors(self, theta, QsT, ls, QtT, lt, want_grad=True):
        """Kernel-level entry: one evaluation with caller-supplied device factors (rows = eigenvectors)."""
        eng = self.eng
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(self.P)
        self.bind()
        stream = torch.cuda.current_stream(eng.device)
        if eng._y_ready is not None:
            stream.wait_event(eng._y_ready)
        out = np.empty((1, self.outw), dtype=np.float64)
        L.call("gpcsd_plan_loglik_grad_factors", self.handle, _pd(theta), QsT.data_ptr(), ls.data_ptr(), QtT.data_ptr(),
               lt.data_ptr(), 1 if want_grad else 0, _pd(out), stream.cuda_stream)
        return out

    def last_launches(self):
        return int(L.load().gpcsd_plan_last_launches(self.handle))

    def set_graph(self, enable):
        L.call("gpcsd_plan_set_graph", self.handle, 1 if enable else 0)
