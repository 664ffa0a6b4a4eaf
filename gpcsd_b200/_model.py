"""Shared implementation of the GPCSD1D / GPCSD2D model classes on top of the CUDA engine.

The public surface (constructor arguments, attribute names, hyperparameter dictionaries, method names
and side effects) follows the reference classes (gpcsd1d.py:19-309, gpcsd2d.py:18-360); the arithmetic
is delegated to ``engine.KronEngine``.  Dimension-specific pieces (parameter names, bounds, quadrature)
live in the two subclasses.
"""
import numpy as np
import scipy.optimize
from tqdm import tqdm

from . import devops
from .engine import HyperParams, KronEngine
from .parallel import RestartShard

np.seterr(all='ignore')  # gpcsd1d.py:7 -- NaN/inf propagate as values


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def _is_scalar(v):
    return np.isscalar(v) or np.ndim(v) == 0


def _dlpdf(prior, v):
    """d lpdf / d value.  The reference's prior protocol is lpdf + sample only (priors.py:14-54; autograd differentiates
    lpdf): a user-supplied prior without the closed-form `dlpdf` gets a 4th-order central difference of its lpdf."""
    f = getattr(prior, "dlpdf", None)
    if f is not None:
        return f(v)
    h = 1e-4 * abs(v)
    return (-prior.lpdf(v + 2 * h) + 8 * prior.lpdf(v + h) - 8 * prior.lpdf(v - h) + prior.lpdf(v - 2 * h)) / (12 * h)


class GPCSDModelBase:
    DIM = None
    JITTER = None
    SPATIAL_ELL_KEYS = ()       # ('ell',) or ('ell1', 'ell2')
    DEFAULT_MAXITER = 1000

    # ------------------------------------------------------------------ engine plumbing
    def _quadrature(self):
        raise NotImplementedError

    def _get_engine(self):
        eng = getattr(self, "_engine", None)
        geom_key = (id(self.x), id(self.t), id(self.spatial_cov))
        if eng is None or self._engine_geom != geom_key:
            eng = KronEngine(self.DIM, self.x, self.t, self._quadrature(), group=getattr(self, "_group", None),
                             jitter=self.JITTER)
            if getattr(self, "_collective_order", None) is not None:
                eng.shard.set_order(*self._collective_order)      # models evaluated from several host threads (parallel.py)
            self._engine, self._engine_geom, self._engine_lfp, self._engine_fp = eng, geom_key, None, None
        fp = self._lfp_fingerprint()
        if self._engine_lfp is not self.lfp or self._engine_fp != fp:
            eng.set_lfp(self.lfp, local=getattr(self, "lfp_is_local", False))
            self._engine_lfp, self._engine_fp = self.lfp, fp
        return eng

    def _lfp_fingerprint(self):
        """Cheap content fingerprint of ``self.lfp`` (address, shape and 64 strided samples): the reference re-reads the array
        on every call, so an in-place edit such as ``model.lfp -= model.lfp.mean(...)`` must trigger a new upload even though
        the object is the same.  Edits that miss all the samples need ``invalidate()``."""
        a = self.lfp
        try:
            flat = a.reshape(-1)
            n = flat.shape[0]
            step = max(1, n // 64)
            return (a.__array_interface__['data'][0], a.shape, float(np.sum(flat[::step][:64])), float(flat[n - 1]) if n else 0.0)
        except Exception:
            return None

    def invalidate(self):
        """Public: the data (or geometry arrays) were modified in place -- re-upload / rebuild at the next evaluation."""
        self._engine_lfp = None
        self._engine_geom = None

    def _invalidate_lfp(self):
        """Force a fresh host->device upload at the next evaluation (called by update_lfp)."""
        self._engine_lfp = None

    def _hyperparams(self):
        sig = self.sig2n['value']
        sig = float(sig) if _is_scalar(sig) else np.asarray(sig, dtype=np.float64)
        return HyperParams(
            R=float(self.R['value']),
            ells=tuple(float(self.spatial_cov.params[k]['value']) for k in self.SPATIAL_ELL_KEYS),
            temporal=[(tc.KIND, float(tc.params['ell']['value']), float(tc.params['sigma2']['value']))
                      for tc in self.temporal_cov_list],
            sig2n=sig, eps=float(getattr(self, "eps", 0.0) or 0.0))

    # ------------------------------------------------------------------ parameter (de)serialisation
    def _temporal_lists(self):
        return ([tc.params['ell']['value'] for tc in self.temporal_cov_list],
                [tc.params['sigma2']['value'] for tc in self.temporal_cov_list])

    def _restore_temporal(self, params):
        if len(self.temporal_cov_list) != len(params['temporal_ell_list']):
            print('different number of temporal covariance functions! stopping.')
            return
        for i, tc in enumerate(self.temporal_cov_list):
            tc.params['ell']['value'] = params['temporal_ell_list'][i]
            tc.params['sigma2']['value'] = params['temporal_sigma2_list'][i]

    def _str_temporal(self):
        s = ""
        for i, tc in enumerate(self.temporal_cov_list):
            s += "Temporal covariance %d class name: %s\n" % (i + 1, type(tc).__name__)
            s += "Temporal covariance %d ell prior: %s\n" % (i + 1, str(tc.params['ell']['prior']))
            s += "Temporal covariance %d ell value %0.4g\n" % (i + 1, tc.params['ell']['value'])
            s += "Temporal covariance %d sigma2 prior: %s\n" % (i + 1, str(tc.params['sigma2']['prior']))
            s += "Temporal covariance %d sigma2 value %0.4g\n" % (i + 1, tc.params['sigma2']['value'])
        return s

    # ------------------------------------------------------------------ likelihood
    def loglik(self):
        """Marginal log-likelihood of all trials at the current hyperparameters
        (gpcsd1d.py:113-128 / gpcsd2d.py:136-151)."""
        return np.float64(self._get_engine().loglik(self._hyperparams()))

    def loglik_and_grad(self):
        """(loglik, gradient w.r.t. R, spatial ell(s), (ell_t, sigma2_t) per temporal kernel, sig2n[...])
        in natural units -- one fused device evaluation (replaces loglik + autograd.grad)."""
        return self._get_engine().loglik_grad(self._hyperparams())

    # ------------------------------------------------------------------ fit objective in log space
    def _param_slots(self):
        """[(dict, scale)] in the reference's tparams order (gpcsd1d.py:160-174 / gpcsd2d.py:185-199):
        value = exp(tparam) * scale."""
        slots = [(self.R, 100.0)] + [(self.spatial_cov.params[k], 100.0) for k in self.SPATIAL_ELL_KEYS]
        for tc in self.temporal_cov_list:
            slots += [(tc.params['ell'], 1.0), (tc.params['sigma2'], 1.0)]
        return slots

    def _noise_is_scalar(self):
        return _is_scalar(self.sig2n['value'])

    def _bounds(self):
        b = [(np.log(d['min'] / sc), np.log(d['max'] / sc)) for d, sc in self._param_slots()]
        if self._noise_is_scalar():
            b.append((np.log(self.sig2n['min']), np.log(self.sig2n['max'])))
        else:
            b += [(np.log(lo), np.log(hi)) for lo, hi in zip(self.sig2n['min'], self.sig2n['max'])]
        return b

    def _set_tparams(self, tparams, fix_R):
        slots = self._param_slots()
        for k, (d, sc) in enumerate(slots):
            if k == 0 and fix_R:
                continue
            d['value'] = np.exp(tparams[k]) * sc
        p = len(slots)
        self.sig2n['value'] = np.exp(tparams[p]) if self._noise_is_scalar() else np.exp(np.asarray(tparams[p:]))

    def _sample_tparams0(self, fix_R):
        """Random start from the priors, same draw order as gpcsd1d.py:194-208."""
        t0 = []
        for k, (d, sc) in enumerate(self._param_slots()):
            v = d['value'] if (k == 0 and fix_R) else d['prior'].sample()
            t0.append(np.log(v) - np.log(sc))
        if self._noise_is_scalar():
            t0.append(np.log(self.sig2n['prior'].sample()))
        else:
            t0 += [np.log(p.sample()) for p in self.sig2n['prior']]
        return np.array(t0)

    def _prior_terms(self):
        """(sum of lpdf, d lpdf / d value per tparams slot) at the current values (gpcsd1d.py:177-186)."""
        vals, pris = [], []
        for d, _ in self._param_slots():
            vals.append(d['value'])
            pris.append(d['prior'])
        if self._noise_is_scalar():
            vals.append(self.sig2n['value'])
            pris.append(self.sig2n['prior'])
        else:
            vals += list(self.sig2n['value'])
            pris += list(self.sig2n['prior'])
        lp = 0.0
        for p, v in zip(pris, vals):
            lp = lp + p.lpdf(v)
        dlp = np.array([_dlpdf(p, v) if v > 0 else 0.0 for p, v in zip(pris, vals)])
        return lp, dlp, np.array(vals, dtype=np.float64)

    def obj_fun(self, tparams, fix_R=False):
        """Negative log posterior at log-space ``tparams``; writes the values into the model like the
        reference's closure (gpcsd1d.py:153-191)."""
        self._set_tparams(tparams, fix_R)
        lp, _, _ = self._prior_terms()
        try:
            llik = self.loglik()
        except np.linalg.LinAlgError:
            if self.DIM == 1:
                raise
            llik = -np.inf                                             # gpcsd2d.py:215-219
        return -1.0 * (llik + lp)

    def obj_fun_and_grad(self, tparams, fix_R=False):
        """(nll, d nll / d tparams): the pair scipy's L-BFGS-B consumes (jac=True)."""
        self._set_tparams(tparams, fix_R)
        lp, dlp, vals = self._prior_terms()
        try:
            ll, g = self.loglik_and_grad()
        except np.linalg.LinAlgError:
            if self.DIM == 1:
                raise
            return np.inf, np.zeros(len(vals))
        grad = -(np.asarray(g) + dlp) * vals                           # chain rule of value = exp(tparam) * scale
        if fix_R:
            grad[0] = 0.0
        return -1.0 * (ll + lp), grad

    def _pure_objective(self, engine, fix_R):
        """(nll, d nll / d tparams) as a function of tparams that does NOT write into the model's dictionaries, bound to
        its own engine: what the concurrent restarts of fit() optimise.  Same arithmetic as obj_fun_and_grad."""
        slots = self._param_slots()
        scales = np.array([sc for _, sc in slots], dtype=np.float64)
        priors = [d['prior'] for d, _ in slots]
        noise_scalar = self._noise_is_scalar()
        priors += [self.sig2n['prior']] if noise_scalar else list(self.sig2n['prior'])
        nslots, nsp = len(slots), len(self.SPATIAL_ELL_KEYS)
        kinds = [tc.KIND for tc in self.temporal_cov_list]
        R_fixed = float(self.R['value'])
        eps = float(getattr(self, "eps", 0.0) or 0.0)
        dim = self.DIM

        def fun(tparams):
            with np.errstate(all='ignore'):        # numpy's error state is per thread; gpcsd1d.py:7 semantics in the workers too
                return _fun(tparams)

        def _fun(tparams):
            tp = np.asarray(tparams, dtype=np.float64)
            vals = np.exp(tp)
            vals[:nslots] *= scales
            if fix_R:
                vals[0] = R_fixed
            lp = 0.0
            for p, v in zip(priors, vals):
                lp = lp + p.lpdf(v)
            dlp = np.array([_dlpdf(p, v) if v > 0 else 0.0 for p, v in zip(priors, vals)])
            temporal = [(kinds[k], float(vals[1 + nsp + 2 * k]), float(vals[2 + nsp + 2 * k])) for k in range(len(kinds))]
            sig = float(vals[nslots]) if noise_scalar else np.array(vals[nslots:])
            hp = HyperParams(R=float(vals[0]), ells=tuple(float(v) for v in vals[1:1 + nsp]), temporal=temporal, sig2n=sig, eps=eps)
            try:
                ll, g = engine.loglik_grad(hp)
            except np.linalg.LinAlgError:
                if dim == 1:
                    raise
                return np.inf, np.zeros(len(vals))                  # gpcsd2d.py:215-219
            grad = -(np.asarray(g) + dlp) * vals
            if fix_R:
                grad[0] = 0.0
            return -1.0 * (ll + lp), grad
        return fun

    # ------------------------------------------------------------------ fit
    def _batched_objective(self, engine, fix_R):
        """fun(X (b, n), idx) -> (nll (b,), d nll / d tparams (b, n)) for b restarts in ONE native call (the restart-batched
        plan); same arithmetic as obj_fun_and_grad, nothing written into the model's dictionaries.  A restart whose
        eigendecomposition fails (numpy would raise LinAlgError, gpcsd1d.py:219 / gpcsd2d.py:215-219) gets nll = +inf."""
        slots = self._param_slots()
        scales = np.array([sc for _, sc in slots], dtype=np.float64)
        priors = [d['prior'] for d, _ in slots]
        noise_scalar = self._noise_is_scalar()
        priors += [self.sig2n['prior']] if noise_scalar else list(self.sig2n['prior'])
        nslots, nsp = len(slots), len(self.SPATIAL_ELL_KEYS)
        kinds = [tc.KIND for tc in self.temporal_cov_list]
        R_fixed = float(self.R['value'])
        eps = float(getattr(self, "eps", 0.0) or 0.0)

        # closed-form priors of this package evaluate vectorised over the restarts; anything else (user-supplied priors with
        # the reference's scalar lpdf protocol) falls back to a per-value loop
        from .priors import GPCSDHalfNormalPrior, GPCSDInvGammaPrior
        builtin = all(type(p) in (GPCSDHalfNormalPrior, GPCSDInvGammaPrior) for p in priors)
        if builtin:
            is_ig = np.array([type(p) is GPCSDInvGammaPrior for p in priors])
            pa = np.array([float(p.alpha) if type(p) is GPCSDInvGammaPrior else 0.0 for p in priors])
            pb = np.array([float(p.beta) if type(p) is GPCSDInvGammaPrior else 0.0 for p in priors])
            psd = np.array([float(p.sd) if type(p) is GPCSDHalfNormalPrior else 1.0 for p in priors])
        template = self._hyperparams()

        def prior_terms(vals):
            if builtin:
                pos = vals > 0
                v = np.where(pos, vals, 1.0)
                lp = np.where(is_ig, -(pa + 1.0) * np.log(v) - pb / v, -0.5 * np.square(v / psd))       # priors.py:27, :50
                dlp = np.where(is_ig, -(pa + 1.0) / v + pb / (v * v), -v / (psd * psd))
                lp = np.where(pos, lp, -np.inf)
                return np.sum(lp, axis=1), np.where(pos, dlp, 0.0)
            lps, dlps = [], []
            for v in vals:
                lp = 0.0
                for p, x in zip(priors, v):
                    lp = lp + p.lpdf(x)
                lps.append(lp)
                dlps.append([_dlpdf(p, x) if x > 0 else 0.0 for p, x in zip(priors, v)])
            return np.array(lps), np.array(dlps)

        def fun(X, idx):
            with np.errstate(all='ignore'):
                X = np.atleast_2d(np.asarray(X, dtype=np.float64))
                vals = np.exp(X)
                vals[:, :nslots] *= scales[None, :]
                if fix_R:
                    vals[:, 0] = R_fixed
                lps, dlps = prior_terms(vals)
                # natural-unit theta rows in the plan's order are exactly `vals` (R, ell(s), (ell_t, sigma2_t)..., sig2n[...])
                ll, g, flag = engine.loglik_grad_thetas(vals, template)
                nll = -(ll + lps)
                grad = -(g + dlps) * vals
                if fix_R:
                    grad[:, 0] = 0.0
                bad = flag != 0
                nll[bad] = np.inf
                grad[bad] = 0.0
                return nll, grad
        return fun

    def _fit_lockstep(self, n_restarts, fix_R, verbose, options):
        """Multi-start MAP fit with ALL restarts advancing in lock step: every optimiser step is one
        restart-batched loglik+gradient call on the GPU.  Same bounds, prior-sampled starts, stopping tests (gtol, ftol,
        maxiter) and best-finite-restart selection as the reference's sequential loop (gpcsd1d.py:137-246)."""
        from .batched_opt import batched_lbfgsb
        bounds = self._bounds()
        shard = RestartShard(getattr(self, "_restart_group", None))
        starts = [self._sample_tparams0(fix_R) for _ in range(n_restarts)]
        starts = self._broadcast_from_rank0(starts, fix_R)
        mine = [i for i in range(n_restarts) if shard.mine(i)]
        local = {}
        if mine:
            eng = self._get_engine()
            res = batched_lbfgsb(self._batched_objective(eng, fix_R), np.array([starts[i] for i in mine]), bounds=bounds,
                                 maxiter=int(options.get('maxiter', self.DEFAULT_MAXITER)), gtol=float(options.get('gtol', 1e-5)),
                                 ftol=float(options.get('ftol', 1e7 * np.finfo(float).eps)))
            self._last_fit_info = {"nit": res["nit"], "batched_calls": res["nfev"], "status": list(res["status"])}
            for k, i in enumerate(mine):
                if res["status"][k] == "nonfinite-start":      # the reference's restart dies with an exception: no entry
                    print("restart %d: objective not finite at the starting point" % i)
                    continue
                local[i] = (float(res["fun"][k]), np.asarray(res["x"][k], dtype=np.float64), str(res["status"][k]))
        return self._select_best(shard.gather(local), fix_R, verbose)

    def _select_best(self, merged, fix_R, verbose):
        nll_values, params, term_msg = [], [], []
        for i in sorted(merged):
            nll_values.append(merged[i][0])
            params.append(merged[i][1])
            term_msg.append(merged[i][2])
        nll_values = np.array(nll_values)
        if len(nll_values) < 1 or not np.any(np.isfinite(nll_values)):
            print('problem with optimization!')
            return
        finite = np.isfinite(nll_values)
        best_ind = np.argmin(nll_values[finite])
        params = [p for p, ok in zip(params, finite) if ok]
        if verbose:
            print('\nNeg log lik values across different initializations:')
            print(nll_values)
            print('Best index termination message')
            print(term_msg[best_ind])
        self._set_tparams(params[best_ind], fix_R)

    def _fit(self, n_restarts, method, fix_R, verbose, options, n_workers=2, lockstep=None):
        if lockstep is None:
            lockstep = (method == 'L-BFGS-B')
        if lockstep:
            return self._fit_lockstep(n_restarts, fix_R, verbose, options)
        bounds = self._bounds()
        options = dict(options)
        if method == 'L-BFGS-B' and not options.get('disp', False):
            options.pop('disp', None)       # the reference passes disp=False; newer scipy warns on the key
        # Every rank draws ALL starting points in order (same RNG stream as an unsharded run, so restart i starts
        # from the same point whatever the world size) and optimises only its own share of them.
        shard = RestartShard(getattr(self, "_restart_group", None))
        starts = [self._sample_tparams0(fix_R) for _ in range(n_restarts)]
        # Distributed runs: rank 0's starting points (and its R when fix_R) are THE starting points -- ranks whose numpy RNG
        # state differs would otherwise optimise different trajectories (trial sharding: different numbers of collectives
        # per rank, i.e. a hang; restart sharding: restart i not reproducible across world sizes).
        starts = self._broadcast_from_rank0(starts, fix_R)
        mine = [i for i in range(n_restarts) if shard.mine(i)]
        # Restarts are independent: run up to n_workers of them concurrently, one host thread + CUDA stream + engine
        # workspace each (all sharing the uploaded LFP), so one restart's latency-bound syevd overlaps another's GEMMs.
        # Trial-sharded models keep one worker: collectives of concurrent evaluations must not interleave.
        if getattr(self, "_group", None) is not None:
            n_workers = 1
        n_workers = max(1, min(int(n_workers), len(mine)))
        base = self._get_engine()
        engines = [base]
        for _ in range(n_workers - 1):
            e = KronEngine(self.DIM, self.x, self.t, self._quadrature(), group=getattr(self, "_group", None), jitter=self.JITTER)
            e.share_data_with(base)
            engines.append(e)
        import queue
        import threading
        import torch
        free = queue.Queue()
        for e in engines:
            free.put((e, torch.cuda.Stream(device=e.device) if n_workers > 1 else None))
        local, lock = {}, threading.Lock()
        bar = tqdm(total=n_restarts, desc="Restarts", disable=shard.rank != 0)

        def run(i):
            eng, stream = free.get()
            try:
                if stream is not None:
                    torch.cuda.set_device(eng.device)       # worker threads start on device 0
                ctx = torch.cuda.stream(stream) if stream is not None else _nullcontext()
                with ctx:
                    res = scipy.optimize.minimize(self._pure_objective(eng, fix_R), starts[i], jac=True, method=method,
                                                  options=options, bounds=bounds)
                with lock:
                    local[i] = (float(res.fun), np.asarray(res.x, dtype=np.float64), str(res.message))
            except (ValueError, np.linalg.LinAlgError) as e:
                print(e)
                if self.DIM == 2:
                    print('\nrestarting optimization...')
            finally:
                free.put((eng, stream))
                with lock:
                    bar.update(1)

        if n_workers == 1:
            for i in mine:
                run(i)
        else:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=n_workers) as pool:
                list(pool.map(run, mine))
        bar.update(n_restarts - len(mine))
        bar.close()
        return self._select_best(shard.gather(local), fix_R, verbose)

    def _broadcast_from_rank0(self, starts, fix_R):
        import torch.distributed as dist
        for grp in (getattr(self, "_group", None), getattr(self, "_restart_group", None)):
            if grp is None or grp is False or not (dist.is_available() and dist.is_initialized()):
                continue
            g = None if grp is True else grp
            if dist.get_world_size(g) == 1:
                continue
            box = [(starts, float(self.R['value']))]
            dist.broadcast_object_list(box, src=dist.get_global_rank(g, 0) if g is not None else 0, group=g)
            starts, r0 = box[0]
            if fix_R:
                self.R['value'] = r0
        return starts

    # ------------------------------------------------------------------ prediction
    def predict(self, z, t, type="csd"):
        """Posterior mean CSD and/or LFP at locations ``z`` and times ``t``; results are stored as attributes
        csd_pred, csd_pred_list, lfp_pred, lfp_pred_list, t_pred, x_pred (gpcsd1d.py:248-293)."""
        out = self._get_engine().predict(self._hyperparams(), z, t, kind=type)
        if type == "both" or type == "csd":
            self.csd_pred_list = out["csd_pred_list"]
            self.csd_pred = out["csd_pred"]
        if type == "both" or type == "lfp":
            self.lfp_pred_list = out["lfp_pred_list"]
            self.lfp_pred = out["lfp_pred"]
        self.t_pred = t
        self.x_pred = z

    # ------------------------------------------------------------------ per-trial evoked shifts (the step after predict)
    def trial_shift_objective(self, mu, tau, mutau=0.0, sigtau=10.0, want_grad=True):
        """nll_r(tau_r) (and its gradient) of auditory_lfp/fit_mean_function.py:311-321 for every trial of ``self.lfp`` in one
        batched device evaluation.  mu: (nx, nt, nseg+1) background + evoked components, tau: (ntrials, nseg)."""
        return self._get_engine().shift_objective(self._hyperparams(), mu, tau, mutau, sigtau, want_grad)

    def fit_trial_shifts(self, mu, tau0=None, mutau=0.0, sigtau=10.0, maxiter=200, gtol=1e-5,
                         ftol=1e7 * np.finfo(float).eps):
        """Per-trial time shifts of the evoked components (fit_mean_function.py:323-335: one scipy L-BFGS-B per trial on
        joblib workers in the reference).  Here all trials advance in lock step (batched_opt.batched_lbfgsb), each step one
        batched evaluation of the objective and its closed-form gradient on the GPU.
        Returns (tau_hat (ntrials, nseg), success (ntrials,) bool, status (ntrials,))."""
        from .batched_opt import batched_lbfgsb
        eng = self._get_engine()
        hp = self._hyperparams()
        N, nseg = eng.ntrials, np.asarray(mu).shape[2] - 1
        T0 = np.zeros((N, nseg)) if tau0 is None else np.broadcast_to(np.asarray(tau0, dtype=np.float64), (N, nseg)).copy()
        full = T0.copy()

        def fun(T, idx):
            full[idx] = T
            f, g = eng.shift_objective(hp, mu, full, mutau, sigtau)
            return f[idx], g[idx]
        res = batched_lbfgsb(fun, T0, maxiter=maxiter, gtol=gtol, ftol=ftol)
        ok = np.array([s in ("gtol", "ftol") for s in res["status"]])
        return res["x"], ok, res["status"]

    # ------------------------------------------------------------------ prior sampling helpers
    def _kt_total(self):
        nt = self.t.shape[0]
        Kt = np.zeros((nt, nt))
        for tc in self.temporal_cov_list:
            Kt += tc.compute_Kt()
        return Kt

