#!/usr/bin/env python
"""Benchmark of the Free Hunch guided-sampling hot path on B200 (contract: see DESIGN.md "Measurement").

One *step* = one guided sampling of a micro-batch of B synthetic 256x256 images through the whole
N-evaluation EDM loop (30-step Heun = 59 mechanism calls by default), every call running the covariance
time/space updates, the CG guidance solve and the measurement operators in libfreehunch_b200.so.

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference ...                   the CPU restatement of the reference on host cores

Prints ONE JSON line (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same through host
buffers (pinned y / noise -> device, final image -> host inside the timed region); `roofline` = the dominant
kernel's algorithmic bytes / CUDA-event time; `cpu_baseline` = oracle/fh_oracle.py on a bounded sample.
"""
import os
import sys

# The CPU legs use every host core: torchrun exports OMP_NUM_THREADS=1 for its children, which would silently time the
# reference on one BLAS thread.  Must happen before numpy / scipy load their thread pools.
if "reference" in sys.argv:
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import argparse
import json
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "free_hunch_images_per_s_256x256"
UNIT = "images/s"
D_VEC_BYTES = 3 * 256 * 256 * 4          # one float32 image vector (786 432 B)

OPERATORS = {"gaussian_blur": "gaussian_ks61_std3", "motion_blur": "motion_ks61_std05",
             "super_resolution": "bicubic_x4", "inpainting": None}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("FH_BENCH_BATCH", "1")),
                    help="images sampled concurrently per GPU (independent covariance states)")
    ap.add_argument("--operator", default="gaussian_blur", choices=list(OPERATORS))
    ap.add_argument("--num-steps", type=int, default=30, help="EDM sigma steps (30 Heun / 100 Euler in BASELINE)")
    ap.add_argument("--solver", default="heun", choices=["heun", "euler"])
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="budget of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-unet-timing", action="store_true", help="skip the separate timing of the random-init PyTorch UNet")
    ap.add_argument("--blocking-host-reads", action="store_true",
                    help="run the mechanism with the reference's per-call host reads (allclose / std / CG flags) instead of deferring them")
    ap.add_argument("--batched", type=int, default=16, help="also report a B-images-per-GPU step (HBM-bound regime); 0 = off")
    ap.add_argument("--images", type=int, default=0,
                    help="sweep mode (BASELINE configs[4]): sample this many images, sharded round-robin over the ranks, with the "
                         "final NCCL gather of the uint8 results inside the timed region; one step = the whole sweep")
    ap.add_argument("--no-unet-in-loop", action="store_true", help="skip the measured image with the random-init UNet as the denoiser")
    return ap.parse_args()


def bench_config(a):
    """The `config` object both arms print (identical keys and values: the workload is the same by construction)."""
    return {"workload": workload_name(a), "operator": a.operator, "num_steps": a.num_steps, "solver": a.solver,
            "micro_batch": a.batch, "denoiser": "analytic DCT-Gaussian stand-in (exact denoiser of the DCT-diagonal prior)",
            "images": a.images}


def workload_name(a):
    cfg = {("gaussian_blur", "heun", 30, 1): "configs[1]", ("inpainting", "heun", 30, 1): "configs[2] (one of the 8 shards)",
           ("super_resolution", "heun", 30, 16): "configs[3] (memory-efficient update)",
           ("motion_blur", "euler", 100, 1): "configs[4] (one image of the sweep)"}.get((a.operator, a.solver, a.num_steps, a.batch), "variant")
    return (f"{cfg}: {a.operator} sigma_y=0.1 ks=61, Free Hunch (dct_diagonal, space updates 1<sigma<10), "
            f"{a.num_steps}-step {a.solver}, 3x256x256, micro-batch {a.batch}/GPU")


# ------------------------------------------------------------------------------------------------ inputs
def synth_inputs(assets, operator, n_images, seed0):
    """Synthetic clean images drawn from the DCT-diagonal Gaussian prior, measurement noise and
    initial noise, all seeded by the global image index so results do not depend on the world size."""
    import scipy.fft as sfft
    var = assets["dct_variance"].astype(np.float64)
    xs, noises, ynoise = [], [], []
    for i in range(n_images):
        rs = np.random.RandomState(seed0 + i)
        z = rs.standard_normal(var.shape) * np.sqrt(var)
        xs.append(np.clip(sfft.idctn(z, type=2, norm="ortho", axes=(-2, -1)), -1, 1))
        noises.append(rs.standard_normal(var.shape))
        ynoise.append(rs.standard_normal(var.shape))
    return np.stack(xs), np.stack(noises), np.stack(ynoise)


def random_mask(seed):
    """measurements.py:287-299 semantics (random p in [0.6, 0.8]), seeded per image."""
    rs = np.random.RandomState(seed)
    prob = rs.uniform(0.6, 0.8)
    m = np.ones(256 * 256, np.float32)
    m[rs.choice(256 * 256, int(256 * 256 * prob), replace=False)] = 0
    return m.reshape(256, 256)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  In-process NVML (nvidia_ml_py) from a thread when it is
    importable: the `nvidia-smi -lms` loop it replaces stalled the GPU for tens of ms per sample on this pool (and for
    whole seconds while starting up), which is a large fraction of an 80 ms step.  Falls back to nvidia-smi."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_s=0.25):
        self.index, self.rows, self.proc, self.first = index, [], None, 0
        self.period, self.how, self._stop, self._thread = period_s, None, threading.Event(), None

    def mark(self):
        """Samples taken before this call (start-up, warm-up steps) are not reported."""
        self.first = len(self.rows)

    def _nvml_loop(self, nv, h):
        smax = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        names = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except AttributeError:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(self.index), str(sm), str(smax), ""] + ["Active" if mask & bit else "Not Active" for _, bit in names])
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        if os.environ.get("FH_BENCH_NO_CLOCKS"):          # diagnostics only: measure the sampler's own interference
            return
        if not os.environ.get("FH_BENCH_NVIDIA_SMI"):
            try:
                import pynvml as nv
                nv.nvmlInit()
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
                h = nv.nvmlDeviceGetHandleByIndex(idx)
                self.how = "nvml"
                self._thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
                self._thread.start()
                return
            except Exception:
                self.how = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "250"], stdout=subprocess.PIPE, text=True)
            self.how = "nvidia-smi"
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        elif self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampler unavailable"]}
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, smax, reasons = [], None, set()
        for r in self.rows[self.first:]:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1])); smax = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "sampler": self.how}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_sample(assets, a, budget_s, threads):
    """The reference algorithm (oracle/fh_oracle.py: float64 CG, complex128 covariance state -- the reference's own
    precision) on the host cores for a BOUNDED prefix of the same workload: mechanism calls of image 0 are run until
    `budget_s` seconds are spent.  Returns per-call records so the caller can scale to images/s."""
    import contextlib
    import scipy.fft as sfft
    import torch
    from oracle import fh_oracle as fo
    torch.set_num_threads(threads)
    pools = contextlib.ExitStack()
    pools.enter_context(sfft.set_workers(threads))                 # scipy.fft (DCTs, OTF products) is 1 worker by default
    try:
        from threadpoolctl import threadpool_limits
        pools.enter_context(threadpool_limits(limits=threads))     # BLAS pools, whatever OMP_NUM_THREADS said at import
    except ImportError:
        pass
    var = assets["dct_variance"]
    x_clean, noise, ynoise = synth_inputs(assets, a.operator, 1, 1000)
    if a.operator == "inpainting":
        mask = np.broadcast_to(random_mask(1000), (1, 3, 256, 256))
        op = fo.InpaintingOracle(mask, 0.1)
        y = (x_clean + 0.1 * ynoise) * mask
    elif a.operator == "super_resolution":
        op = fo.SuperResolutionOracle(assets["bicubic_x4"], 0.1, 4)
        y = op.A(x_clean) + 0.1 * ynoise[..., :64, :64]
    else:
        cls = fo.GaussianBlurOracle if a.operator == "gaussian_blur" else fo.MotionBlurOracle
        op = cls(assets[OPERATORS[a.operator]], 0.1)
        y = op.A(x_clean) + 0.1 * ynoise
    t_steps = fo.edm_sigma_steps(a.num_steps)
    mech = fo.FreeHunchOracle(op, fo.CovarianceOracleDCT(var, t_steps[0] ** 2, var.size))
    model = fo.analytic_dct_gaussian_denoiser(var)
    calls = []
    t_begin = time.perf_counter()

    class Budget(Exception):
        pass

    def timed(x_t, mdl, yy, s):
        t0 = time.perf_counter()
        out = mech(x_t, mdl, yy, s)
        calls.append({"sigma": float(s), "seconds": time.perf_counter() - t0, "niter": mech.log[-1]["niter"]})
        if time.perf_counter() - t_begin > budget_s:
            raise Budget()
        return out
    try:
        fo.conditional_sampler(timed, model, y, noise, t_steps, solver=a.solver)
        complete = True
    except Budget:
        complete = False
    finally:
        pools.close()
    return calls, complete, time.perf_counter() - t_begin


def total_calls(a):
    return 2 * a.num_steps - 1 if a.solver == "heun" else a.num_steps


def cpu_images_per_s(a, calls, secs):
    """Scale the bounded prefix to whole images: by CG-iteration share when the oracle's full iteration profile of this
    workload is committed (tests/golden/bench_oracle_iters.json, tools/make_bench_iters.py), else by call share."""
    key = f"{a.operator}/{a.num_steps}/{a.solver}"
    try:
        prof = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_oracle_iters.json")))[key]
        total = float(sum(prof["niters"]))
        done = float(sum(c["niter"] for c in calls))
        return (done / total) / secs, f"{int(done)} of {int(total)} CG iterations of one image (first {len(calls)} of {total_calls(a)} mechanism calls)"
    except (OSError, KeyError):
        return (len(calls) / total_calls(a)) / secs, f"first {len(calls)} of {total_calls(a)} mechanism calls of one image, scaled by call count"


def run_reference_arm(a, assets):
    """The reference algorithm on the host cores (oracle/fh_oracle.py: the reference is a set of Python scripts, not an
    installable package, and /root/reference does not travel to the GPU box -- kind "port").  Step 1 samples a WHOLE
    image when a short probe says it fits ~150 s (no extrapolation); further steps are bounded prefixes of the same
    image scaled by the CG-iteration share of step 1's own complete profile."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(5.0, min(a.cpu_seconds, 240.0 / max(1, a.steps)))
    est_total = None
    if a.warmup > 0:
        calls, complete, secs = cpu_reference_sample(assets, a, 3.0, cores)
        v0, _ = cpu_images_per_s(a, calls, secs)
        est_total = 1.0 / max(v0, 1e-9)
    vals, samples, profile, profile_secs = [], [], None, 0.0
    t0 = time.perf_counter()
    for step in range(a.steps):
        # whole images while they fit ~4 minutes in total; otherwise a bounded prefix scaled by what the SAME calls took
        # inside step 1's whole image (per-call seconds, not iteration counts: early iterations are the cheap ones)
        spent = time.perf_counter() - t0
        whole = (est_total is None or est_total <= 150.0) and (step == 0 or (profile is not None and spent + profile_secs <= 240.0))
        calls, complete, secs = cpu_reference_sample(assets, a, 1e9 if whole else per_step, cores)
        if complete:
            profile = [c["seconds"] for c in calls]
            profile_secs = secs
            vals.append(1.0 / secs)
            samples.append(f"a whole image: {len(calls)} mechanism calls, {sum(c['niter'] for c in calls)} CG iterations in {secs:.1f} s")
        elif profile is not None:
            share = sum(profile[:len(calls)]) / sum(profile)
            vals.append(share / secs)
            samples.append(f"first {len(calls)} mechanism calls of the same image in {secs:.1f} s, scaled by the time share ({share:.3f}) those calls had in step 1")
        else:
            v_, how = cpu_images_per_s(a, calls, secs)
            vals.append(v_)
            samples.append(how)
    ms = (time.perf_counter() - t0) / max(1, a.steps) * 1e3
    v = float(np.mean(vals))
    sample = f"oracle/fh_oracle.py (float64 CG, complex128 state) on {cores} threads; step 1: {samples[0]}" + (f"; later steps: {samples[-1]}" if len(samples) > 1 else "")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/c128",
        "data": "synthetic", "config": bench_config(a), "values_per_step": [float(x) for x in vals],
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------ B200 arm
def algorithmic_bytes_per_launch(kernel, m_rows, B):
    """Algorithmic bytes of ONE launch of a fused CG kernel at Z rank m_rows (DESIGN.md "Kernels"): every operand
    touched once, unit = one float32 image vector (786 432 B) per image."""
    u = D_VEC_BYTES * B
    table = {
        "fh_k_cg_fwd_rows<RR>": 4 * u,                 # r, p in; p, T out
        "fh_k_cg_fwd_cols<RR>": (3 + m_rows) * u,      # T, delta, Z rows in; w_hat out
        "fh_k_cg_inv_cols<RR>": (3 + m_rows) * u,      # w_hat, delta, Z rows in; T' out
        "fh_k_cg_inv_rows<RR>": 3 * u,                 # T', p in; q out
        "fh_k_cg_xr<RR>": 6 * u,                       # x, p, r, q in; x, r out
    }
    return table.get(kernel)


def algorithmic_bytes_per_cg_iteration(m_rows, operator="gaussian_blur"):
    """one blur / inpainting CG iteration touches p, x, r (read), x, r, p (write), the diagonal of Sigma
    and the 2k = m low-rank vectors once: (7 + m) image vectors (the mask / OTF extra is not counted); x4 super-resolution
    has 1/16-size measurement vectors: (1 + m) image vectors."""
    return ((1 if operator == "super_resolution" else 7) + m_rows) * D_VEC_BYTES


def algorithmic_bytes_of_step(log, operator):
    """SURVEY.md 8(d), summed over one sampled image: every CG iteration at the rank it ran at, every time update
    (8k + 12 vectors) and space update (10k + 19) -- the numerator of the STEP-level roofline fraction."""
    tot, k_prev = 0, 0
    for r in log:
        k = int(r["k"])
        if r.get("time_update"):
            tot += (8 * k_prev + 12) * D_VEC_BYTES
        if r.get("space_update"):
            tot += (10 * k_prev + 19) * D_VEC_BYTES
        tot += max(r["niter"]) * algorithmic_bytes_per_cg_iteration(2 * k, operator)
        k_prev = k
    return tot

def run_b200(a, assets):
    import torch
    import torch.distributed as dist
    from free_hunch_b200 import engine as E
    from free_hunch_b200.conditioning_mechanisms import choose_conditioning_mechanism, PhaseTimer
    from free_hunch_b200.measurements import get_operator
    from free_hunch_b200.sampler import AnalyticDCTGaussianDenoiser, conditional_sampler, edm_sigma_steps

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: free_hunch_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    var = assets["dct_variance"]
    t_steps = edm_sigma_steps(a.num_steps)

    def make_problem(step, B, seed0=None):
        # every step samples NEW images: global image index = (step * world + rank) * B + b  (weak scaling)
        if seed0 is None:
            seed0 = 1000 + (step * world + rank) * B
        x_clean, noise, ynoise = synth_inputs(assets, a.operator, B, seed0)
        kw = dict(in_shape=(B, 3, 256, 256), sigma_s=0.1, device=dev)
        if a.operator == "inpainting":
            masks = np.stack([random_mask(seed0 + b) for b in range(B)])
            op = get_operator("inpainting", mask=torch.from_numpy(masks)[:, None].repeat(1, 3, 1, 1), **kw)
            y = (x_clean + 0.1 * ynoise) * masks[:, None]
        else:
            sf = 4 if a.operator == "super_resolution" else 1
            op = get_operator(a.operator, kernel=assets[OPERATORS[a.operator]], scale_factor=sf, **kw)
            xa = op.A(torch.from_numpy(x_clean).float().to(dev)).double().cpu().numpy()
            y = xa + 0.1 * ynoise[..., :256 // sf, :256 // sf]
        return op, y.astype(np.float32), noise.astype(np.float32)

    def make_mechanism(op, B, timer=None):
        return choose_conditioning_mechanism("free_hunch")(
            1.0, op, False, init_denoiser_variance=1, init_noise_variance=torch.tensor(float(t_steps[0]) ** 2, dtype=torch.float64),
            data_dim=3 * 256 * 256, max_vector_count=100000, dct_variance=var, image_base_covariance="dct_diagonal",
            denoiser_mean_error_threshold=0.2, use_analytical_score_time_update=True, project_to_diagonal=False,
            space_step_update_threshold=10.0, space_step_update_lower_threshold=1.0, max_rtol=1.0, do_space_updates=True,
            use_analytic_var_at_end=False, solver_type="customcuda", use_rtol_func=True, device=dev, batch=B, timer=timer,
            defer_host_reads=not a.blocking_host_reads)

    nets = {}

    host_trace = []

    def sample(op, y_dev, noise_dev, timer=None):
        B = y_dev.shape[0]
        t_mk = time.perf_counter()
        mech = make_mechanism(op, B, timer)
        if os.environ.get("FH_BENCH_TRACE_HOST"):            # diagnostics: host seconds per mechanism call
            inner = mech.x0_mean_update
            host_trace.clear()
            host_trace.append(("make_mechanism", time.perf_counter() - t_mk))

            def traced(*args, **kw):
                t0 = time.perf_counter()
                out = inner(*args, **kw)
                host_trace.append((f"call{len(host_trace)}", time.perf_counter() - t0))
                return out
            mech.x0_mean_update = traced
            fin = mech.finalize

            def traced_fin():
                t0 = time.perf_counter()
                out = fin()
                host_trace.append(("finalize", time.perf_counter() - t0))
                return out
            mech.finalize = traced_fin
        ctx = mech.covariance_model.ctx
        if id(ctx) not in nets:
            nets[id(ctx)] = AnalyticDCTGaussianDenoiser(var, ctx)
        x = conditional_sampler(mech, nets[id(ctx)], y_dev, noise_dev, t_steps, solver=a.solver)
        return x, mech

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    import gc
    B = a.batch
    n_iter = a.warmup + a.steps + 2                # warm-up, timed steps (both arms sample the same images), phase-timed step, kernel-profiled step
    problems = [make_problem(s, B) for s in range(n_iter)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # 256 MiB > 126 MB L2
    it = 0
    # nvidia-smi is started BEFORE the warm-up: its start-up (NVML attach, ~1-3 s) stalls the GPU for whole seconds when it
    # falls into a timed step; only samples taken after clocks.mark() are reported
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    mech = None
    for _ in range(a.warmup):
        op, y, noise = problems[it]; it += 1
        # like the timed loop, keep the previous step's mechanism alive while the next one is built: the library parks the
        # buffers of destroyed covariance states and a step that finds the pool empty pays ~100 ms of cudaMalloc
        # (that was the "second timed step is slow" stall: the warm-up dropped every mechanism at once, the timed loop did not)
        _, mech = sample(op, torch.from_numpy(y).to(dev), torch.from_numpy(noise).to(dev))
    # ---- device-resident arm: inputs already in HBM, CUDA events per step, L2 flushed between steps
    if rank == 0 and not os.environ.get("FH_BENCH_NO_CLOCKS"):
        t_wait = time.perf_counter()                       # let the sampler get past its start-up: a few complete query
        while len(clocks.rows) < 3 and time.perf_counter() - t_wait < 3.0:   # cycles (first NVML calls on a fresh box are slow
            time.sleep(0.05)                               # and contend with launches for driver locks)
    # a cyclic-GC pass over the interpreter's objects is tens of ms of host time in the middle of a 80 ms step: collect
    # now, then keep the collector off while steps are being timed
    gc.collect()
    gc.freeze()
    gc.disable()
    barrier()
    clocks.mark()
    E.launch_count(reset=True)
    ms_steps, cg_iters, ranks_k = [], [], []
    for _ in range(a.steps):
        op, y, noise = problems[it]; it += 1
        y_dev, noise_dev = torch.from_numpy(y).to(dev), torch.from_numpy(noise).to(dev)
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dbg = PhaseTimer(True) if os.environ.get("FH_BENCH_TRACE_PHASES") else None
        e0.record()
        x, mech = sample(op, y_dev, noise_dev, dbg)
        e1.record()
        e1.synchronize()
        ms_steps.append(e0.elapsed_time(e1))
        if host_trace:
            top = sorted(host_trace, key=lambda kv: -kv[1])[:4]
            print(f"[host] step {len(ms_steps)} {ms_steps[-1]:.1f} ms; host total {sum(v for _, v in host_trace) * 1e3:.1f} ms; top {[(n, round(v * 1e3, 2)) for n, v in top]}", file=sys.stderr, flush=True)
        if dbg is not None:                                  # diagnostics: where did a slow step spend its time?
            torch.cuda.synchronize()
            durs = sorted(((p0_.elapsed_time(p1_), nm, i) for i, (nm, p0_, p1_) in enumerate(dbg.pairs)), reverse=True)
            gaps = sorted(((dbg.pairs[i][2].elapsed_time(dbg.pairs[i + 1][1]), dbg.pairs[i][0] + "->" + dbg.pairs[i + 1][0], i) for i in range(len(dbg.pairs) - 1)), reverse=True)
            print(f"[phases] step {len(ms_steps)} {ms_steps[-1]:.1f} ms; longest phases {[(round(d, 2), n, i) for d, n, i in durs[:4]]}; "
                  f"longest gaps {[(round(d, 2), n, i) for d, n, i in gaps[:4]]}", file=sys.stderr, flush=True)
        cg_iters.append(sum(max(r["niter"]) for r in mech.log))
        ranks_k.append(mech.covariance_model.vector_count)
        assert bool(torch.isfinite(x).all()), "non-finite sample"
    barrier()
    launches = E.launch_count()
    total_ms = reduce_max(float(sum(ms_steps)))
    value = world * B * a.steps / (total_ms * 1e-3)
    # ---- end-to-end arm: pinned host buffers in, host image out, wall clock of the user-facing call
    e2e_ms = []
    for k in range(-1, a.steps):                               # k = -1: one untimed pass over the pinned-memory path
        op, y, noise = problems[a.warmup + max(k, 0)]          # the same images as the device-resident arm
        y_pin, n_pin = torch.from_numpy(y).pin_memory(), torch.from_numpy(noise).pin_memory()
        out_pin = torch.empty((B, 3, 256, 256), dtype=torch.float64).pin_memory()
        flush.fill_(1)
        barrier()
        t0 = time.perf_counter()
        # `mech` is rebound like in the device-resident arm: holding the last mechanism of that arm AND the previous pass's
        # would keep three covariance states alive and make the first timed pass allocate one (cudaMalloc in the loop)
        x, mech = sample(op, y_pin.to(dev, non_blocking=True), n_pin.to(dev, non_blocking=True))
        out_pin.copy_(x, non_blocking=True)
        torch.cuda.synchronize()
        if k >= 0:
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
        if host_trace:
            top = sorted(host_trace, key=lambda kv: -kv[1])[:4]
            print(f"[host-e2e] pass {k} {(time.perf_counter() - t0) * 1e3:.1f} ms; host total {sum(v for _, v in host_trace) * 1e3:.1f} ms; top {[(n, round(v * 1e3, 2)) for n, v in top]}", file=sys.stderr, flush=True)
    gc.enable()
    e2e_total = reduce_max(float(sum(e2e_ms)))
    clk = clocks.stop() if rank == 0 else None
    e2e_value = world * B * a.steps / (e2e_total * 1e-3)
    h2d = int(problems[0][1].nbytes + problems[0][2].nbytes)
    d2h = int(B * 3 * 256 * 256 * 8)
    # ---- sweep mode (BASELINE configs[4]): N images round-robin over the ranks, no collective in the loop, ONE all-gather of
    #      the decoded uint8 images at the end (generate_conditional.py:328-340, 557-569) inside the timed region
    sweep = None
    if a.images > 0:
        from free_hunch_b200.outputs import decode_uint8
        from free_hunch_b200.sharding import gather_images, shard_indices
        idxs = shard_indices(a.images, rank, world)
        local = []
        for gi in idxs:                                      # inputs resident in HBM before the clock starts
            op_i, y_i, n_i = make_problem(0, 1, seed0=50000 + gi)
            local.append((op_i, torch.from_numpy(y_i).to(dev), torch.from_numpy(n_i).to(dev)))
        flush.fill_(1)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        imgs, its = [], 0
        for op_i, y_i, n_i in local:
            x_i, mech_i = sample(op_i, y_i, n_i)
            imgs.append(decode_uint8(x_i))
            its += sum(max(r["niter"]) for r in mech_i.log)
        mine = torch.cat(imgs) if imgs else torch.empty((0, 3, 256, 256), dtype=torch.uint8, device=dev)
        allimgs = gather_images(mine, idxs, a.images)
        s1.record(); s1.synchronize()
        sweep_ms = reduce_max(s0.elapsed_time(s1))
        sweep = {"images": a.images, "n_gpus": world, "images_per_s": a.images / (sweep_ms * 1e-3), "ms_total": sweep_ms,
                 "images_this_rank": len(idxs), "cg_iterations_this_rank": its, "gathered_bytes": int(allimgs.numel()),
                 "gather": "torch.distributed.all_gather over NCCL (uint8, 196 608 B per image), inside the timed region" if world > 1 else "single rank: local placement",
                 "checksum": int(allimgs.to(torch.int64).sum().item())}
    # ---- one step with CUDA-event phase timers (the reference's own phase boundaries, conditioning_mechanisms.py:242-271)
    op, y, noise = problems[it]; it += 1
    timer = PhaseTimer(True)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    x, mech = sample(op, torch.from_numpy(y).to(dev), torch.from_numpy(noise).to(dev), timer)
    p1.record(); p1.synchronize()
    prof_ms = p0.elapsed_time(p1)
    phases = timer.totals()
    # ---- one step with an event pair around every kernel launch: per-kernel SHARES (each pair adds ~6 us and breaks the
    #      programmatic-dependent-launch overlap, so absolutes are inflated; the roofline uses fh_cg_kernel_time instead)
    op, y, noise = problems[it]; it += 1
    E.profile_kernels(True)
    sample(op, torch.from_numpy(y).to(dev), torch.from_numpy(noise).to(dev))
    torch.cuda.synchronize()
    E.profile_kernels(False)
    kprof = E.kernel_profile()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    n_calls = len(mech.log)
    n_cg = sum(max(r["niter"]) for r in mech.log)
    hot_ms = sum(phases.get(p, (0, 0.0))[1] for p in ("time_update", "space_update", "mat_solver"))
    kern_total = sum(ms for _, ms in kprof.values())
    cg_names = [k for k in kprof if k.split("<")[0] in E.CG_KERNELS]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))

    def kernel_roofline(state, op_handle, Bk):
        """Dominant fused-CG kernel at the state's final rank: back-to-back launches between two CUDA events."""
        m_rows = state.rank
        times = {n: E.cg_kernel_time(state, op_handle, i, reps=200) for i, n in enumerate(E.CG_KERNELS)}
        top = max(times, key=times.get)
        ab = algorithmic_bytes_per_launch(top + "<RR>", m_rows, Bk)
        ach = ab / (times[top] * 1e-3) / 1e9
        return top, times, ab, ach, m_rows

    def resident_roofline(state, op_handle, iters=200):
        """The resident solve is ONE kernel: `iters` iterations per launch, (7 + m) image vectors of algorithmic bytes each."""
        m_rows = state.rank
        ms = E.cg_resident_time(state, op_handle, iters=iters, reps=5)
        ab = iters * algorithmic_bytes_per_cg_iteration(m_rows, a.operator)
        return ms, ab, ab / (ms * 1e-3) / 1e9, m_rows

    roofline = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None}
    fold = a.operator in ("gaussian_blur", "inpainting")
    ctx0 = mech.covariance_model.ctx
    resident = ctx0.cg_last_mode == E.CG_RESIDENT
    step_bytes = algorithmic_bytes_of_step(mech.log, a.operator)
    step_frac = step_bytes / (hot_ms * 1e-3) / 1e9 / peak if hot_ms > 0 else None
    if resident:
        iters = 200
        ms, ab, ach, m_rows = resident_roofline(mech.covariance_model.state, mech.forward_operator.op, iters)
        # the same kernel at rank 0: most iterations of a run happen before the first space update (sigma > 10)
        st0 = E.CovState(ctx0, 1, E.BASIS_DCT)
        st0.init(torch.from_numpy(var), float(t_steps[0]) ** 2)
        ms0 = E.cg_resident_time(st0, mech.forward_operator.op, iters=iters, reps=5)
        ab0 = iters * algorithmic_bytes_per_cg_iteration(0, a.operator)
        st0.close()
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary_r02_res64.json")))["kernels"][E.CG_RES64_KERNEL]["dram_bytes_per_launch"]
        except (OSError, KeyError):
            pass
        share = sum(v[1] for k, v in kprof.items() if k.startswith(E.CG_RES64_KERNEL)) / kern_total
        stimes = {}
        if fold:
            ctx0.set_cg_mode(E.CG_STREAMING)
            try:
                _, stimes, _, _, _ = kernel_roofline(mech.covariance_model.state, mech.forward_operator.op, B)
            finally:
                ctx0.set_cg_mode(E.CG_AUTO)
        roofline.update({"kernel": E.CG_RES64_KERNEL, "achieved": ach, "frac": ach / peak, "traffic": traffic,
                         "rank0": {"us_per_cg_iteration": ms0 * 1e3 / iters, "achieved": ab0 / (ms0 * 1e-3) / 1e9, "frac": ab0 / (ms0 * 1e-3) / 1e9 / peak},
                         "step": {"algorithmic_bytes": step_bytes, "cov_plus_cg_ms": hot_ms, "achieved": step_bytes / (hot_ms * 1e-3) / 1e9,
                                  "frac": step_frac, "note": "all CG iterations at the rank they ran at + time / space updates (SURVEY 8d) over "
                                  "the phase-timed covariance + CG time of one sampled image"},
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                         "avg_us": ms * 1e3, "iterations_per_launch": iters, "us_per_cg_iteration": ms * 1e3 / iters,
                         "algorithmic_bytes_per_launch": ab, "algorithmic_bytes_per_iteration": ab // iters, "Z_rows": m_rows,
                         "share_of_kernel_time_in_step": share,
                         "streaming_path_kernels_us": {k: round(v * 1e3, 2) for k, v in stimes.items()},
                         "note": "one cooperative launch per solve, float64 arithmetic (the reference's CG precision): Z (m x 786 KB) stays "
                                 "in shared memory and x/r/p in registers, so the kernel moves ~0 bytes of DRAM; `achieved` divides the "
                                 "contract's algorithmic bytes ((7+m) image vectors per iteration, SURVEY 8d) by the measured time. "
                                 "The kernel is bound by its grid-wide exchanges, the FP64 pipe and shared-memory latency, not by HBM"})
    elif fold:
        top, times, ab, ach, m_rows = kernel_roofline(mech.covariance_model.state, mech.forward_operator.op, B)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary_r01.json")))["kernels"][top]["dram_bytes_per_launch"]
        except (OSError, KeyError):
            pass
        share = sum(kprof[k][1] for k in cg_names if k.startswith(top)) / kern_total
        roofline.update({"kernel": top, "achieved": ach, "frac": ach / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                         "avg_us": times[top] * 1e3, "algorithmic_bytes_per_launch": ab, "Z_rows": m_rows,
                         "all_kernels_us": {k: round(v * 1e3, 2) for k, v in times.items()},
                         "share_of_kernel_time_in_step": share,
                         "note": f"micro-batch {B}: " + ("the working set (<= 36 MB) is L2-resident, so the kernel is bound by "
                                 "L2 / SM latency rather than HBM; `batched` has the HBM-bound regime" if B < 4 else "HBM-bound regime")})
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": total_ms / a.steps, "ms_steps": [round(v, 3) for v in ms_steps], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 CG solve (vectors, DCTs, operators, reductions); f32 storage of the covariance state", "data": "synthetic",
        "config": bench_config(a),
        "run": {"l2": "256 MiB flush between steps; every step samples new images", "mechanism_calls": n_calls,
                "host_reads": "per call (reference behaviour)" if a.blocking_host_reads else "deferred: one synchronisation per sampled image (FreeHunch.finalize)",
                "cg": "float64 resident kernel (fh_k_cg_res64)" if resident else "float32 streaming iteration",
                "cg_iterations_per_image": float(np.mean(cg_iters)), "cg_iterations_steps": cg_iters, "final_bfgs_pairs": ranks_k[-1]},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_total / a.steps, "ms_steps": [round(v, 3) for v in e2e_ms]},
        "gpu_launches": int(launches),
        "sweep": sweep,
        "clocks": clk,
        "roofline": roofline,
        "hot_path": {"phase_timed_step_ms": prof_ms, "cov_plus_cg_ms_per_mechanism_call": hot_ms / n_calls,
                     "cov_plus_cg_us_per_cg_iteration": hot_ms / max(1, n_cg) * 1e3, "cg_iterations": n_cg,
                     "phases_ms": {k: round(v[1], 3) for k, v in phases.items()},
                     "kernels_ms_event_pairs": {k: [v[0], round(v[1], 3)] for k, v in sorted(kprof.items(), key=lambda kv: -kv[1][1])}},
    }
    # ---- the HBM-bound regime: B independent images per GPU (one warm-up step, one timed step)
    if fold and a.batched > 1 and B == 1 and world == 1:
        Bb = a.batched
        res, res_resident = [], []
        for stp in range(4):
            # steps 0,1: streaming CG (Z streamed from HBM, all images per launch; what FH_CG_AUTO picks at this batch);
            # steps 2,3: resident CG forced (image by image)
            opb, yb, nb = make_problem(100 + stp % 2, Bb)
            yd, nd = torch.from_numpy(yb).to(dev), torch.from_numpy(nb).to(dev)
            flush.fill_(1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            mechb = make_mechanism(opb, Bb)
            ctxb = mechb.covariance_model.ctx
            ctxb.set_cg_mode(E.CG_STREAMING if stp < 2 else E.CG_RESIDENT)
            if id(ctxb) not in nets:
                nets[id(ctxb)] = AnalyticDCTGaussianDenoiser(var, ctxb)
            xb = conditional_sampler(mechb, nets[id(ctxb)], yd, nd, t_steps, solver=a.solver)
            e1.record(); e1.synchronize()
            (res if stp < 2 else res_resident).append(e0.elapsed_time(e1))
        ctxb.set_cg_mode(E.CG_STREAMING)
        top, times, ab, ach, m_rows = kernel_roofline(mechb.covariance_model.state, mechb.forward_operator.op, Bb)
        ctxb.set_cg_mode(E.CG_AUTO)
        traffic_b = None
        try:
            traffic_b = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary_r01_batch16.json")))["kernels"][top]["dram_bytes_per_launch"] if Bb == 16 else None
        except (OSError, KeyError):
            pass
        out["batched"] = {"micro_batch": Bb, "cg": "streaming (the FH_CG_AUTO choice above 4 images): all images per launch, Z streamed from HBM",
                          "images_per_s": Bb / (res[-1] * 1e-3), "ms_per_step": res[-1],
                          "images_per_s_resident_cg": Bb / (res_resident[-1] * 1e-3), "ms_per_step_resident_cg": res_resident[-1],
                          "cg_iterations_max_over_batch": sum(max(r["niter"]) for r in mechb.log),
                          "roofline": {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                       "traffic": traffic_b, "avg_us": times[top] * 1e3, "algorithmic_bytes_per_launch": ab, "Z_rows": m_rows,
                                       "all_kernels_us": {k: round(v * 1e3, 2) for k, v in times.items()}}}
    # ---- the denoiser the reference would run (random-init guided-diffusion UNet, PyTorch), timed separately
    if not a.no_unet_timing and world == 1:
        try:
            from free_hunch_b200.denoiser import build_denoiser
            den = build_denoiser("imagenet256", dev)
            xin = torch.randn(1, 3, 256, 256, dtype=torch.float64, device=dev)
            sig = torch.tensor(5.1405, dtype=torch.float64, device=dev)
            w = torch.randn(1, 3, 256, 256, dtype=torch.float64, device=dev)

            def call():
                xg = xin.clone().requires_grad_()
                d, _ = den(xg, sig)
                return torch.autograd.grad((w * d).sum(), xg)[0]
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            u0.record()
            for _ in range(10):
                call()
            u1.record(); u1.synchronize()
            unet_ms = u0.elapsed_time(u1) / 10
            standin_ms = phases.get("denoiser", (0, 0.0))[1]
            step_ms = total_ms / a.steps
            out["unet"] = {"config": "imagenet256 guided-diffusion UNet, random init, fp16 autocast, batch 1 (free_hunch_b200/denoiser.py)",
                           "params_M": sum(p.numel() for p in den.parameters()) / 1e6, "fwd_plus_vjp_ms_per_call": unet_ms,
                           "calls_per_image": n_calls}
            # CUDA-graph capture of forward + VJP
            try:
                from free_hunch_b200.denoiser import GraphedDenoiser
                gden = GraphedDenoiser(den, xin)

                def gcall():
                    xg = xin.clone().requires_grad_()
                    d, _ = gden(xg, 5.1405)
                    return torch.autograd.grad((w * d).sum(), xg)[0]
                g_ref, g_got = call(), gcall()
                for _ in range(3):
                    gcall()
                torch.cuda.synchronize()
                u0.record()
                for _ in range(10):
                    gcall()
                u1.record(); u1.synchronize()
                out["unet"]["fwd_plus_vjp_ms_per_call_cuda_graph"] = u0.elapsed_time(u1) / 10
                out["unet"]["cuda_graph_vjp_rel_diff_vs_eager"] = float((g_got - g_ref).norm() / g_ref.norm())
            except Exception as e:
                gden = None
                out["unet"]["cuda_graph_error"] = repr(e)[:300]
            # one image MEASURED with the random-init UNet as the denoiser inside the guided loop (BASELINE: "images/s on
            # random-init UNet weights").  A random-init net is not a denoiser: the CG systems it induces are arbitrary, so
            # the iteration count is reported next to the time.
            if not a.no_unet_in_loop and B == 1:
                try:
                    op, y, noise = problems[a.warmup]
                    for use, name in ((gden, "cuda_graph"), (den, "eager")):
                        if use is None:
                            continue
                        torch.cuda.synchronize()
                        u0.record()
                        mech_u = make_mechanism(op, 1)
                        x_u = conditional_sampler(mech_u, use, torch.from_numpy(y).to(dev), torch.from_numpy(noise).to(dev), t_steps, solver=a.solver)
                        u1.record(); u1.synchronize()
                        ms_u = u0.elapsed_time(u1)
                        out["unet"][f"in_loop_{name}"] = {"images_per_s": 1e3 / ms_u, "ms_per_image": ms_u,
                                                         "cg_iterations": sum(max(r["niter"]) for r in mech_u.log),
                                                         "finite": bool(torch.isfinite(x_u).all())}
                except Exception as e:
                    out["unet"]["in_loop_error"] = repr(e)[:300]
            del den
        except Exception as e:                                   # the UNet timing is auxiliary: never fail the bench on it
            out["unet"] = {"error": repr(e)[:200]}
    if not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        calls, complete, secs = cpu_reference_sample(assets, a, a.cpu_seconds, cores)
        v, how = cpu_images_per_s(a, calls, secs)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"oracle/fh_oracle.py (float64 CG, complex128 state) for {secs:.1f} s: {how}"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    assets = np.load(os.path.join(ROOT, "tests", "golden", "assets.npz"))
    if a.impl == "reference":
        run_reference_arm(a, assets)
    else:
        run_b200(a, assets)


if __name__ == "__main__":
    main()
