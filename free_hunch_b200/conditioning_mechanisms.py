"""`conditioning_mechanism=free_hunch`: host-side mirror of the Free Hunch pieces of
conditioning_utils/conditioning_mechanisms.py (reference :16-50, :190-294, :307-323, :562-600).

Same registry function, constructor kwargs and call contract as the reference's `online_covariance`
(BFGSOnlineUpdate); the covariance model, the Sigma.v products, the measurement operators and the CG guidance
solve all run in the sm_100a library.  The denoiser (`model`) and its VJP stay in PyTorch.
"""
import math
from warnings import warn

import torch

from . import engine
from .covariance import CovarianceHessianBFGS, CovarianceHessianBFGSDCT


class PhaseTimer:
    """CUDA-event timers around the phases the reference times with time.time() prints
    (conditioning_mechanisms.py:242-271): no host sync inside the loop, elapsed times are read at the end."""

    def __init__(self, enabled=False):
        self.enabled = enabled
        self.pairs = []

    def __call__(self, name):
        return _Phase(self, name)

    def totals(self):
        """{phase: (calls, total_ms)}; synchronises."""
        out = {}
        if self.pairs:
            torch.cuda.synchronize()
        for name, a, b in self.pairs:
            c, t = out.get(name, (0, 0.0))
            out[name] = (c + 1, t + a.elapsed_time(b))
        return out


class _Phase:
    def __init__(self, timer, name):
        self.timer, self.name = timer, name

    def __enter__(self):
        if self.timer.enabled:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if self.timer.enabled:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            self.timer.pairs.append((self.name, self.a, b))
        return False


def choose_conditioning_mechanism(name):
    """conditioning_mechanisms.py:16-36 with the new `free_hunch` entry (F1 of SURVEY.md: the reference calls the
    same mechanism `online_covariance`).  The other entries are baselines, not Free Hunch, and are not provided."""
    if name in ("free_hunch", "online_covariance"):
        return FreeHunch
    if name in ("dps", "pigdm", "pigdm_videodiff_schedule", "peng_convert", "peng_analytic", "tmpd", "diffpir"):
        raise ValueError(f"conditioning mechanism '{name}' is a baseline of the reference and outside free_hunch_b200")
    if name == "ddnm":
        raise ValueError("DDNM conditioning mechanism not implemented in this branch of the codebase")
    raise ValueError(f"Unknown conditioning mechanism: {name}")


class ConditioningMechanism:
    """conditioning_mechanisms.py:38-50."""

    def __init__(self, cond_scaling, forward_operator, clip_x0_mean, init_denoiser_variance=None,
                 init_noise_variance=None, data_dim=None, pigdm_posthoc_scaling=False, **argv):
        self.cond_scaling = cond_scaling
        self.forward_operator = forward_operator
        self.clip_x0_mean = clip_x0_mean

    def __call__(self, x_t, x_0_mean, y, sigma):
        x_0_mean_new = self.x0_mean_update(x_t, x_0_mean, y, sigma)
        if self.clip_x0_mean:
            x_0_mean_new = x_0_mean_new.clip(-1, 1)
        return x_0_mean_new


def rtol_func(sigma, rtol_max=1e0, rtol_min=1e-14):
    """conditioning_mechanisms.py:307-323, including the clamp expression of :313 as written."""
    sigma_min, sigma_max = 0.1, 80.0
    sigma = max(min(sigma, sigma_max), max(sigma_min, sigma))
    p = 0.1
    log_factor = ((math.log10(sigma) - math.log10(sigma_min)) / (math.log10(sigma_max) - math.log10(sigma_min))) ** p
    log_rtol = log_factor * (math.log10(rtol_max) - math.log10(rtol_min)) + math.log10(rtol_min)
    return 10 ** log_rtol


def choose_solver(operator_name, operator, y, x0_mean, theta0_var=None, covariance_model=None, method="customcuda",
                  max_rtol=1, ortho_tf=None, sigma_t=None, use_rtol_func=False, maxiter=5000, return_info=False,
                  fixed_iters=0, defer=False):
    """conditioning_mechanisms.py:562-600 for method='customcuda' (the reference default, config.yaml:121): returns
    mat = A^T (A Sigma A^T + sigma_y^2 I)^-1 (y - A x0_mean) from the CG solve with x0 = b and
    rtol = rtol_func(sigma_t, max_rtol).  The scipy paths of the reference are CPU baselines and are not provided."""
    if operator_name not in ("gaussian_blur", "super_resolution", "motion_blur", "inpainting"):
        raise ValueError("Invalid operator name. Please choose 'gaussian_blur', 'super_resolution', 'motion_blur', or 'inpainting'.")
    if method != "customcuda":
        raise ValueError(f"solver_type '{method}' is a CPU path of the reference; free_hunch_b200 provides 'customcuda' only")
    # sigma_s is a 1-element float32 tensor; clip and square in float32 like the reference (:387,:491,:642)
    sy2 = getattr(operator, "_fh_sy2", None)
    if sy2 is None or operator._fh_sy2_src is not operator.sigma_s:
        sig = operator.sigma_s.detach().to(torch.float32).cpu().clip(min=0.001)        # one device read per operator, not per call
        if operator_name == "super_resolution":
            sig = sig.clip(min=1e-2)
        sy2 = float((sig ** 2).item())
        try:
            operator._fh_sy2, operator._fh_sy2_src = sy2, operator.sigma_s
        except AttributeError:
            pass
    res = None
    if defer:
        # no host read-back: info is an engine.DeferredCgInfo (decoded by FreeHunch.finalize); None = not available here
        res = engine.cg_solve_async(covariance_model.state, operator.op, y, x0_mean, sy2, rtol_func(sigma_t, max_rtol),
                                    maxiter=maxiter, fixed_iters=fixed_iters)
    if res is not None:
        mat, info = res
    else:
        mat, info = engine.cg_solve(covariance_model.state, operator.op, y, x0_mean, sy2, rtol_func(sigma_t, max_rtol),
                                    maxiter=maxiter, fixed_iters=fixed_iters)
        if any(n >= maxiter for n in info["niter"]):
            warn("CG not converge.")
    mat = mat.to(x0_mean.dtype)
    return (mat, info) if return_info else mat


class FreeHunch(ConditioningMechanism):
    """The Free Hunch mechanism = the reference's BFGSOnlineUpdate (conditioning_mechanisms.py:190-294).
    One instance per sampled image batch; stateful (history of sigma, x, denoiser mean kept ON DEVICE)."""

    def __init__(self, cond_scaling, forward_operator, clip_x0_mean, init_denoiser_variance, init_noise_variance, data_dim,
                 pigdm_posthoc_scaling=False, **argv):
        super().__init__(cond_scaling, forward_operator, clip_x0_mean)
        argv.setdefault("solver_type", "customcuda")
        argv.setdefault("max_rtol", 1.0)
        argv.setdefault("use_rtol_func", True)
        argv.setdefault("project_to_diagonal", False)
        argv.setdefault("image_base_covariance", "dct_diagonal")
        argv.setdefault("max_vector_count", None)
        argv.setdefault("do_space_updates", True)
        argv.setdefault("denoiser_mean_error_threshold", 0.2)
        argv.setdefault("use_analytical_score_time_update", True)
        argv.setdefault("space_step_update_threshold", 10.0)
        argv.setdefault("space_step_update_lower_threshold", 1.0)
        argv.setdefault("use_analytic_var_at_end", False)
        # defer_host_reads: run x0_mean_update without any host synchronisation (see x0_mean_update); the caller must
        # call finalize() after the last step.  Off by default: direct callers get the reference's blocking behaviour.
        argv.setdefault("defer_host_reads", False)
        if argv["use_analytic_var_at_end"]:
            raise ValueError("use_analytic_var_at_end (scipy baseline branch, conditioning_mechanisms.py:273-278) is not provided")
        self.argv = argv
        self.batch = int(argv.get("batch", 1))
        device = argv.get("device", getattr(forward_operator, "device", None))
        nv = init_noise_variance.item() if torch.is_tensor(init_noise_variance) else float(init_noise_variance)
        base = argv["image_base_covariance"]
        mvc = argv["max_vector_count"]
        if mvc is not None and mvc >= engine.MAX_RANK // 2:
            mvc = None                        # the reference default (100000) never drops; capacity is MAX_RANK/2 pairs
        if base == "identity":
            shape = tuple(getattr(forward_operator, "in_shape", (1, 1, 1, data_dim))[-3:])
            self.covariance_model = CovarianceHessianBFGS(init_denoiser_variance, nv, data_dim, max_vector_count=mvc,
                                                          project_to_diagonal=argv["project_to_diagonal"], device=device,
                                                          batch=self.batch, image_shape=shape)
        elif base in ("dct_diagonal", "dct_diagonal_noinfo"):
            self.covariance_model = CovarianceHessianBFGSDCT(argv.get("data_dir"), nv, data_dim, max_vector_count=mvc,
                                                             dct_variance=argv.get("dct_variance"),
                                                             use_precalculated_info=(base == "dct_diagonal"),
                                                             project_to_diagonal=argv["project_to_diagonal"], device=device,
                                                             batch=self.batch)
        else:
            raise ValueError(f"image_base_covariance '{base}' is not available (pca_dct_diagonal is undefined in the reference too)")
        self.do_space_updates = argv["do_space_updates"]
        self.init_denoiser_variance = init_denoiser_variance
        self.init_noise_variance = init_noise_variance
        self.data_dim = data_dim
        self.sigmas, self.xs, self.denoiser_means = [], [], []
        self.denoiser_mean_error_threshold = argv["denoiser_mean_error_threshold"]
        self.use_analytical_score_time_update = argv["use_analytical_score_time_update"]
        self.space_step_update_threshold = argv["space_step_update_threshold"]
        self.space_step_update_lower_threshold = argv["space_step_update_lower_threshold"]
        self.log = []
        self.timer = argv.get("timer") or PhaseTimer(False)
        self.defer = bool(argv["defer_host_reads"])
        self._deferred = []                   # (log record, DeferredCgInfo | None, same-x flag | None, fallback flag | None)

    def mat_solver(self, y, x0_mean, covariance_model=None, sigma_t=None, return_info=False):
        return choose_solver(self.forward_operator.name, self.forward_operator, y, x0_mean, None,
                             covariance_model or self.covariance_model, self.argv["solver_type"], self.argv["max_rtol"],
                             None, sigma_t, use_rtol_func=self.argv["use_rtol_func"], return_info=return_info,
                             defer=self.defer)

    def update_time_step(self, x_t, sigma_t, sigma_tnext, score_t):
        self.covariance_model.update_time_step(x_t, sigma_t, sigma_tnext, score_t)

    def update_space_step(self, denoiser_mean_at_x, denoiser_mean_at_xnext, sigma_t, x, xnext):
        self.covariance_model.update_space_step(denoiser_mean_at_x, denoiser_mean_at_xnext, sigma_t, x, xnext)

    def x0_mean_update(self, x_t, model, y, sigma):
        """conditioning_mechanisms.py:237-294.

        The reference reads three device values on the host in every call: `torch.allclose(x_t, xs[-1])` (:250), the
        `.std() > threshold` fallback test (:283) and the CG loop's convergence flags.  With `defer_host_reads` the call
        queues its work without waiting for any of them: the fallback is selected on the device (both candidates are
        computed, `torch.where` picks), the CG outcome stays in device memory, and the x-changed test is ASSUMED true
        (a sampler never evaluates the mechanism twice at the same point) and verified in finalize(), which raises if
        the assumption was wrong.  The arithmetic is the same in both modes."""
        sigma_f = float(sigma.item()) if torch.is_tensor(sigma) else float(sigma)
        x_t = x_t.requires_grad_()
        # one library call per piece of element-wise glue (fh_mech.cuh) where the tensors are what the sampler hands over
        # (float64 on the device, DCT basis); anything else takes the PyTorch expressions of the reference
        fused = bool(self.argv.get("fused_glue", True)) and x_t.is_cuda and x_t.dtype == torch.float64 and \
            self.covariance_model.state.basis == engine.BASIS_DCT
        if fused and y.dtype != torch.float64:                      # widen the measurement once per image, not per call
            if getattr(self, "_y_src", None) is not y:
                self._y_src, self._y64 = y, y.to(torch.float64)
            y = self._y64
        with self.timer("denoiser"):
            x_0_mean, _ = model(x_t, sigma)
        cov = self.covariance_model
        rec = {"sigma": sigma_f, "time_update": False, "space_update": False}
        if self.do_space_updates:
            prev_mean_at_sigma = None
            if len(self.sigmas) != 0 and sigma_f != self.sigmas[-1]:                                     # :241-244
                with self.timer("time_update"):
                    if fused:
                        prev_mean_at_sigma, _ = cov.state.time_update_from_mean(self.xs[-1], self.denoiser_means[-1], self.sigmas[-1], sigma_f)
                        cov._sigma = sigma_f
                        cov._track("time_update")                  # factor history of drop_vectors (max_vector_count only)
                    else:
                        score_previous = (self.denoiser_means[-1] - self.xs[-1]) / self._scalar(self.sigmas[-1] ** 2, x_t)
                        prev_mean_at_sigma, _ = cov.update_time_step(self.xs[-1], self.sigmas[-1], sigma_f, score_previous)
                rec["time_update"] = True
            elif len(self.sigmas) != 0:                                                                  # :248-249
                prev_mean_at_sigma = self.denoiser_means[-1]
            same_x = None
            if len(self.xs) != 0:
                if self.defer:
                    # the test only decides whether the space update below runs: outside its sigma window (and with the
                    # analytical score transfer) nothing depends on it -- at small sigma consecutive Euler iterates DO
                    # agree to allclose's tolerance, exactly as in the reference
                    in_window = self.space_step_update_lower_threshold < sigma_f < self.space_step_update_threshold
                    if in_window or not self.use_analytical_score_time_update:
                        if fused:
                            same_x = engine.mech_notclose(cov.ctx, x_t.detach(), self.xs[-1]).sum() == 0
                        else:
                            same_x = torch.isclose(x_t.detach(), self.xs[-1]).all()   # device flag, checked in finalize()
                    x_changed = True
                elif fused:
                    x_changed = bool(engine.mech_notclose(cov.ctx, x_t.detach(), self.xs[-1]).any())
                else:
                    x_changed = not torch.allclose(x_t.detach(), self.xs[-1])
            rec["_same_x"] = same_x
            if len(self.xs) != 0 and x_changed:                                                          # :250
                if not self.use_analytical_score_time_update:                                            # :252-254
                    with torch.no_grad():
                        prev_mean_at_sigma, _ = model(self.xs[-1], sigma)
                if self.space_step_update_lower_threshold < sigma_f < self.space_step_update_threshold:  # :256
                    with self.timer("space_update"):
                        cov.update_space_step(prev_mean_at_sigma, x_0_mean.detach(), sigma_f, self.xs[-1], x_t.detach())
                    rec["space_update"] = True
        elif len(self.sigmas) != 0 and sigma_f != self.sigmas[-1]:                                       # :261-263
            score_previous = (self.denoiser_means[-1] - self.xs[-1]) / self.sigmas[-1] ** 2
            with self.timer("time_update"):
                cov.update_time_step(self.xs[-1], self.sigmas[-1], sigma_f, score_previous, only_covariance=True)
            rec["time_update"] = True

        with self.timer("mat_solver"):
            mat, info = self.mat_solver(y, x_0_mean.detach(), sigma_t=sigma_f, return_info=True)        # :268
        s2 = sigma_f ** 2
        with self.timer("denoiser"):
            # :280 is grad((mat.detach() * x_0_mean).sum(), x_t): the same VJP with mat as the cotangent -- the sum's
            # backward multiplies mat by 1.0, so the result is bit-identical and four element-wise launches shorter
            p_y_xt_grad = torch.autograd.grad(x_0_mean, x_t, grad_outputs=mat.detach().to(x_0_mean.dtype))[0]
        # :283-288.  The reference runs batch 1; with B independent images per call the VJP-vs-Sigma.mat decision is taken
        # PER IMAGE (a whole-batch statistic would make an image's guidance depend on its batch mates).
        if fused:
            # Sigma.mat (float64) is formed only for the images that take that branch, as in the reference's `if`
            x_0_mean_new, flag = engine.mech_combine_lazy(cov.state, x_0_mean.detach(), p_y_xt_grad, mat.detach(), s2,
                                                          self.denoiser_mean_error_threshold, self.cond_scaling)
            if self.defer:
                rec["fallback"] = flag
            else:
                flags = [bool(v) for v in flag.tolist()]
                rec["fallback"] = flags[0] if len(flags) == 1 else flags
        else:
            s2d = self._scalar(s2, x_t)                  # divisor as a 0-dim device tensor: a true division like the reference's
            use_cov = (p_y_xt_grad * s2).flatten(1).std(dim=1) > self.denoiser_mean_error_threshold      # (B,), on the device
            if self.defer:
                alt = cov.denoiser_cov_vector_dot(mat.detach(), use_cuda=True) * self.cond_scaling / s2d
                p_y_xt_grad = torch.where(use_cov.view(-1, 1, 1, 1), alt, p_y_xt_grad * self.cond_scaling)
                rec["fallback"] = use_cov
            else:
                flags = [bool(v) for v in use_cov.tolist()]                                              # host read, like :283
                if any(flags):
                    alt = cov.denoiser_cov_vector_dot(mat.detach(), use_cuda=True) * self.cond_scaling / s2d
                    p_y_xt_grad = alt if all(flags) else torch.where(use_cov.view(-1, 1, 1, 1), alt, p_y_xt_grad * self.cond_scaling)
                else:
                    p_y_xt_grad = p_y_xt_grad * self.cond_scaling
                rec["fallback"] = flags[0] if len(flags) == 1 else flags
            x_0_mean_new = x_0_mean + p_y_xt_grad * s2
        # :290-292 -- history stays ON DEVICE; only the last entry is ever read, so only it is kept
        self.sigmas.append(sigma_f)
        self.xs = [x_t.detach()]
        self.denoiser_means = [x_0_mean.detach()]
        rec["niter"] = info["niter"] if isinstance(info, dict) else info
        rec["k"] = cov.vector_count
        self.log.append(rec)
        return x_0_mean_new

    @staticmethod
    def _scalar(v, like):
        """A divisor as a 0-dim tensor on `like`'s device.  The reference divides by 0-dim device tensors (sigma.pow(2),
        t_hat), which PyTorch evaluates as a true division; dividing a CUDA tensor by a Python float is evaluated as a
        multiplication with the rounded reciprocal, one ulp off -- enough to move a loosely converged CG stop."""
        return torch.full((), float(v), dtype=like.dtype, device=like.device)

    def finalize(self):
        """Deferred mode: ONE synchronisation for the whole run -- decode the CG outcomes and device-side flags into the
        log, warn about unconverged solves and verify the x-changed assumption.  No-op otherwise."""
        pending = [r for r in self.log if not isinstance(r.get("niter"), list) or torch.is_tensor(r.get("fallback"))
                   or r.get("_same_x") is not None]
        if not pending:
            return self.log
        maxiter = self.argv.get("maxiter", 5000)
        for r in pending:
            if not isinstance(r["niter"], list):
                info = r["niter"].fetch()
                r["niter"] = info["niter"]
                if any(n >= maxiter for n in info["niter"]):
                    warn("CG not converge.")
            if torch.is_tensor(r.get("fallback")):
                flags = [bool(v) for v in r["fallback"].reshape(-1).tolist()]
                r["fallback"] = flags[0] if len(flags) == 1 else flags
            same = r.pop("_same_x", None)
            if same is not None and bool(same.item()):
                raise RuntimeError("defer_host_reads: the mechanism was evaluated twice at the same x_t (sigma = "
                                   f"{r['sigma']}); the reference skips the space update there -- rerun with defer_host_reads=False")
        return self.log


# the reference's name for the same class
BFGSOnlineUpdate = FreeHunch
