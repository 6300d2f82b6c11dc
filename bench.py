#!/usr/bin/env python
"""bench.py -- filter+smoother state-steps/sec (fp64) on the BASELINE.json batched sweep.

    python bench.py --gpus N --steps K --warmup W            # B200 arm (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on host cores

Workload (config.workload = "c5"): BASELINE.json config 5 -- 65,536 independent series x 10,000
steps, Matern-7/2 state blocks (state dim d = 4 * nblk, default d = 4), scalar Gaussian observations
(m = 1, sigma^2 = 0.1), per-series lengthscales ~ LogU, shared irregular time grid, 5 % observations
missing.  WEAK scaling: every rank owns 65,536 series of its own (series are independent, no data-path
collective); each rank walks them in sub-batches of `--sub-batch` series whose full-state outputs
(filtered m, P and smoothed m, P, each step's d x d block in reference element order, batch stored
time-major) are materialised in HBM.

A "step" = one pass of the hot path (filter kernel + smoother kernel) over the rank's 65,536 x 10,000
job.  `value` = N * series * T * K / (max over ranks of the CUDA-event time of the K timed steps), inputs
resident in HBM.  `e2e` = the same job through the reference-shaped host API with HOST buffers: per
sub-batch a pinned-host -> device copy of Y, filter + smoother, and a device -> pinned-host read of the
user-facing result (smoothed mean / variance of f and the per-series log marginal likelihood).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SERIES_TOTAL = 65536
T_STEPS = 10000
NOISE_VAR = 0.1
NAN_FRAC = 0.05
DT0 = 0.1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c5", "c1", "c3", "cvi", "c2", "c2cvi", "c3cvi", "grad", "spatial"],
                    help="c5 (default, the headline metric): batched sweep; c3: one long series, parallel-in-time "
                         "scan, time-sharded over the ranks; cvi: CVI ELBO + natural-gradient step (config 4)")
    ap.add_argument("--chunk-len", type=int, default=256, help="c3: steps per scan chunk")
    ap.add_argument("--filter-type", default="b200_auto", help="cvi: filter_type of the surrogate SDE_GP")
    ap.add_argument("--obs-dim", type=int, default=0, help="c3: observation dim (0 = full state, m = d)")
    ap.add_argument("--state-dim", type=int, default=None, help="4 * nblk Matern-7/2 blocks (default 4; c3: 8)")
    ap.add_argument("--series", type=int, default=None, help="series per GPU (default 65536; c3: 1; cvi: 1000)")
    ap.add_argument("--T", type=int, default=None, help="steps (default 10000; c3: 1000000)")
    ap.add_argument("--sub-batch", type=int, default=32768)
    ap.add_argument("--e2e-sub-batch", type=int, default=4096)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="c5: weak = --series per GPU (default); strong = --series in TOTAL, split over the ranks")
    ap.add_argument("--post-sub-batch", type=int, default=None,
                    help="series per launch of the packed posterior-only timing (default: the whole batch when its "
                         "workspace fits)")
    ap.add_argument("--no-post", action="store_true",
                    help="skip the posterior-only (packed hand-over) timing beside the c5 d = 4 line")
    ap.add_argument("--no-sweep", action="store_true",
                    help="c5 default line: skip the d = 8 / 16 / 32 sweep and the CVI-step section")
    ap.add_argument("--cpu-sample-series", type=int, default=0, help="series in the CPU sample (0 = auto)")
    a = ap.parse_args()
    dflt = {"c5": (SERIES_TOTAL, T_STEPS, 4), "c1": (1, 10000, 2), "c3": (1, 1000000, 8), "cvi": (1000, T_STEPS, 2),
            "c2": (200, 5000, 400), "c2cvi": (200, 5000, 400), "c3cvi": (1, 1000000, 4), "grad": (32768, T_STEPS, 4),
            "spatial": (200, 5000, 200)}[a.workload]
    a.series = dflt[0] if a.series is None else a.series
    a.T = dflt[1] if a.T is None else a.T
    a.state_dim = dflt[2] if a.state_dim is None else a.state_dim
    return a


def algorithmic_bytes(d, m):
    """Bytes per state-step: filter in y,R,dt + out (m,P); smoother re-read + out."""
    filt = 8 * (d * d + d + m * m + m + 1)
    smooth = 8 * 2 * (d * d + d)
    return filt, smooth


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def recorded_traffic(d):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("d%d" % d)
        except Exception:
            return None
    return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index, self.skip = [], None, index, 0

    def start(self):
        """Starts `nvidia-smi -lms 20` and returns once it is delivering rows (it takes a few hundred ms to come
        up -- longer than a short timed region), so that every row kept is from the timed region itself."""
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
            t0 = time.perf_counter()
            while not self.rows and time.perf_counter() - t0 < 3.0:
                time.sleep(0.005)
            self.skip = len(self.rows)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows[self.skip:] or self.rows
        for r in rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[0])), mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ inputs
def make_hypers(series, nblk, seed=0):
    """Per-series lengthscales ~ LogU(0.5, 2) x (10 dt0), unit variances."""
    rng = np.random.default_rng(seed)
    ls = np.exp(rng.uniform(np.log(0.5), np.log(2.0), (series, nblk))) * (10 * DT0)
    steps = rng.uniform(0.5, 1.5, T_STEPS) * DT0
    return ls, steps


def device_observations(n, T, dev, seed):
    """Seeded synthetic observations generated on the device: smooth signal + noise, 5 % NaN. [n, T, 1]"""
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    k = torch.arange(T, device=dev, dtype=torch.float64)[None, :]
    phase = torch.rand((n, 1), generator=g, device=dev, dtype=torch.float64) * 6.283185307179586
    freq = 0.01 + 0.04 * torch.rand((n, 1), generator=g, device=dev, dtype=torch.float64)
    Y = torch.sin(freq * k + phase) + 0.3 * torch.randn((n, T), generator=g, device=dev, dtype=torch.float64)
    miss = torch.rand((n, T), generator=g, device=dev) < NAN_FRAC
    Y[miss] = float("nan")
    # time-major in memory ([T, n, 1]), logical shape [n, T, 1]: the B200 layout
    return Y.t().contiguous()[..., None].transpose(0, 1)


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_port_rate(d, T, sample_series, seed=0, nthreads=None, budget_s=10.0):
    """state-steps/s of the C oracle port (oracle/ssm_oracle.c, OpenMP over series) on host cores."""
    from oracle import c_oracle
    from physs_gp_b200 import sdes
    c_oracle.build()
    if nthreads is None:
        nthreads = _host_threads()
    nblk = d // 4
    rng = np.random.default_rng(seed)
    ls, steps = make_hypers(sample_series, nblk, seed)
    t = np.cumsum(steps[:T])
    prior = sdes.BatchedMaternSDE(4, ls)
    k = np.arange(T)[None, :]
    Y = np.sin(rng.uniform(0.01, 0.05, (sample_series, 1)) * k + rng.uniform(0, 6.28, (sample_series, 1)))
    Y = Y + 0.3 * rng.normal(size=Y.shape)
    Y[rng.uniform(size=Y.shape) < NAN_FRAC] = np.nan
    args = (4, prior.lam(), prior.P_inf(), prior.H(), t, Y[..., None], np.array([[NOISE_VAR]]))
    # repeat the call on the same sample until ~`budget_s` seconds of CPU work have been timed (outputs of
    # one call: sample_series * T * 320 B, so the sample itself stays small)
    c_oracle.filter_smooth(*args, jitter=1e-5, full_state=True, keep_filtered=True, nthreads=nthreads)   # warm-up
    el, reps, out = 0.0, 0, None
    while el < budget_s and reps < 200:
        t0 = time.perf_counter()
        out = c_oracle.filter_smooth(*args, jitter=1e-5, full_state=True, keep_filtered=True, nthreads=nthreads)
        el += time.perf_counter() - t0
        reps += 1
    return sample_series * T * reps / el, out["threads"], el, reps


def _host_threads():
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank)
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


CPU_KIND_NOTE = ("restatement, not the JAX reference (jax/jaxlib are not installable here): oracle/ssm_oracle.c, the "
                 "numpy oracle's C port, OpenMP over series")


def cpu_c3_rate(d, m, T_sample, budget_s=10.0):
    """state-steps/s of the C oracle port on ONE long series (config 3 shape, m = d full-state sites): the
    sequential recursion cannot use more than one core per series."""
    from oracle import c_oracle
    from physs_gp_b200 import sdes
    c_oracle.build()
    rng = np.random.default_rng(0)
    nblk = d // 4
    steps = rng.uniform(0.5, 1.5, T_sample) * DT0
    prior = sdes.BatchedMaternSDE(4, np.exp(rng.uniform(np.log(0.5), np.log(2.0), (1, nblk))) * (10 * DT0),
                                  full_state_obs=(m == d))
    Y = np.sin(0.01 * np.arange(T_sample))[None, :, None] + 0.3 * rng.normal(size=(1, T_sample, m))
    H = np.eye(d) if m == d else prior.H()
    args = (4, prior.lam(), prior.P_inf(), H, np.cumsum(steps), Y, NOISE_VAR * np.eye(m))
    c_oracle.filter_smooth(*args, jitter=1e-5, full_state=True, keep_filtered=True, nthreads=1)
    el, reps = 0.0, 0
    while el < budget_s and reps < 50:
        t0 = time.perf_counter()
        c_oracle.filter_smooth(*args, jitter=1e-5, full_state=True, keep_filtered=True, nthreads=1)
        el += time.perf_counter() - t0
        reps += 1
    return T_sample * reps / el, el, reps


def cpu_cvi_step_ms(B_full, T, sample_blocks, budget_s=10.0, beta=0.1, K=20):
    """ms per CVI iteration (natural-gradient update + ELBO) of the CPU restatement oracle/cvi_vec.py (C port of
    the filter / smoother + vectorised numpy site algebra) on `sample_blocks` blocks, scaled linearly to B_full."""
    from oracle import c_oracle, cvi_vec
    from physs_gp_b200 import sdes
    c_oracle.build()
    nthreads = _host_threads()
    rng = np.random.default_rng(0)
    n = sample_blocks
    t = np.cumsum(rng.uniform(0.5, 1.5, T) * DT0)
    prior = sdes.BatchedMaternSDE(2, np.exp(rng.uniform(np.log(0.5), np.log(2.0), (n, 1))) * (10 * DT0))
    rate = np.exp(0.5 * np.sin(0.02 * np.arange(T))[None, :] + 0.3 * rng.normal(size=(n, 1)))
    Y = rng.poisson(rate).astype(np.float64)
    Y[rng.uniform(size=Y.shape) < NAN_FRAC] = np.nan
    pa = (2, prior.lam(), prior.P_inf(), prior.H())
    Yt, Vt = np.full((n, T), 1e-5), np.ones((n, T))
    Yt, Vt, _ = cvi_vec.cvi_iteration(pa, t, Y, Yt, Vt, "poisson", beta, K=K, nthreads=nthreads)       # warm-up
    el, reps = 0.0, 0
    while el < budget_s and reps < 100:
        t0 = time.perf_counter()
        Yt, Vt, elbo = cvi_vec.cvi_iteration(pa, t, Y, Yt, Vt, "poisson", beta, K=K, nthreads=nthreads)
        el += time.perf_counter() - t0
        reps += 1
    assert np.isfinite(elbo).all()
    ms_sample = 1e3 * el / reps
    return ms_sample * B_full / n, nthreads, el, reps, ms_sample


def cpu_c2_rate(Ns, T_sample):
    """state-steps/s of the numpy oracle (LAPACK through numpy, its own threading) on the config-2 shape."""
    from oracle import filters as of
    from oracle import sde as osde
    rng = np.random.default_rng(0)
    Xs = rng.uniform(size=(Ns, 2))
    D2 = ((Xs[:, None, :] - Xs[None, :, :]) ** 2).sum(-1)
    Ks = np.exp(-0.5 * D2 / 0.2 ** 2) + 1e-6 * np.eye(Ns)
    prior = osde.LTI_SDE([osde.SpaceTimeSeparable(osde.Matern32(10 * DT0, 1.0), Ks)])
    t = DT0 * np.arange(1, T_sample + 1)
    Y = rng.normal(size=(T_sample, Ns))
    R = np.tile(NOISE_VAR * np.eye(Ns), [T_sample, 1, 1])
    t0 = time.perf_counter()
    lml, mf, Pf, _ = of.filter_sequential(prior, t, Y, R, 1e-5)
    of.smoother_sequential(prior, t, mf, Pf, full_state=False, jitter=1e-5)
    el = time.perf_counter() - t0
    return T_sample / el, el


def run_reference(a):
    """CPU arm: the reference's algorithm on the host cores for the SAME workload / metric / unit as the B200 arm.
    The JAX reference is not installable here, so this times the oracle restatement (kind "port") on a bounded
    sample of the workload and says so in config.workload and cpu_baseline.sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    d, T = a.state_dim, a.T
    cores = os.cpu_count() or 1
    base = {"impl": "reference", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "gpu_launches": 0}
    if a.workload == "cvi":
        n = a.cpu_sample_series or max(cores, 8)
        vals = []
        for i in range(a.warmup + a.steps):
            ms_full, threads, el, reps, ms_s = cpu_cvi_step_ms(a.series, T, n, budget_s=2.0)
            if i >= a.warmup:
                vals.append(ms_full)
        value = float(np.mean(vals))
        sample = ("%d of %d blocks x %d steps, ~2 s per bench step, scaled linearly to %d blocks (oracle/cvi_vec.py: "
                  "C port of filter + smoother, vectorised numpy site algebra / Gauss-Hermite K=20) -- %s"
                  % (n, a.series, T, a.series, CPU_KIND_NOTE))
        cfg = cvi_config(a.series, T)
        cfg["workload"] += " [CPU arm: %s]" % sample
        line = dict(base, metric="CVI ELBO+natgrad step time", value=value, unit="ms", ms_per_step=value,
                    higher_is_better=False, scaling="weak", config=cfg,
                    cpu_baseline={"value": value, "unit": "ms", "cores": threads, "kind": "port", "sample": sample},
                    e2e={"value": value, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line), flush=True)
        return
    if a.workload == "c1":
        vals = [cpu_c1_ms(T, budget_s=1.0)[0] for _ in range(a.warmup + a.steps)][a.warmup:]
        ms = float(np.mean(vals))
        sample = "the whole workload (1 series x %d steps), ONE core -- %s" % (T, CPU_KIND_NOTE)
        line = dict(base, metric="filter+smoother state-steps/sec (fp64)", value=T / (ms * 1e-3), unit="state-steps/s",
                    ms_per_step=ms, higher_is_better=True, scaling="weak",
                    config={"workload": "c1: ONE 1-D temporal Matern-3/2 series (state dim 2) x %d steps [CPU arm: %s]"
                                        % (T, sample)},
                    cpu_baseline={"value": T / (ms * 1e-3), "unit": "state-steps/s", "cores": 1, "kind": "port",
                                  "sample": sample},
                    e2e={"value": T / (ms * 1e-3), "unit": "state-steps/s", "h2d_bytes_per_step": 0,
                         "d2h_bytes_per_step": 0})
        print(json.dumps(line), flush=True)
        return
    if a.workload in ("c3", "c3cvi"):
        m = a.obs_dim or d
        Ts = min(T, 100000)
        vals = []
        for i in range(a.warmup + a.steps):
            r, el, reps = cpu_c3_rate(d, m, Ts, budget_s=2.0)
            if i >= a.warmup:
                vals.append((r, el))
        value = float(np.mean([r for r, _ in vals]))
        sample = ("first %d of %d steps of the series, ~2 s per bench step, ONE core (the sequential recursion of a "
                  "single series does not thread) -- %s" % (Ts, T, CPU_KIND_NOTE))
        line = dict(base, metric="filter+smoother state-steps/sec (fp64)", value=value, unit="state-steps/s",
                    ms_per_step=1e3 * a.series * T / value, higher_is_better=True, scaling="strong",
                    config={"workload": "c3: %d series x %d steps, state dim %d, obs dim %d [CPU arm: %s]"
                                        % (a.series, T, d, m, sample), "series": a.series, "T": T, "state_dim": d,
                            "obs_dim": m},
                    cpu_baseline={"value": value, "unit": "state-steps/s", "cores": 1, "kind": "port", "sample": sample},
                    e2e={"value": value, "unit": "state-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line), flush=True)
        return
    if a.workload == "c2":
        Ns = a.series
        vals = []
        for i in range(a.warmup + a.steps):
            r, el = cpu_c2_rate(Ns, 8)
            if i >= a.warmup:
                vals.append(r)
        value = float(np.mean(vals))
        sample = ("8 of %d time steps (numpy oracle oracle/filters.py, LAPACK threading as numpy configures it) -- "
                  "restatement, not the JAX reference" % T)
        line = dict(base, metric="filter+smoother state-steps/sec (fp64)", value=value, unit="state-steps/s",
                    ms_per_step=1e3 * T / value, higher_is_better=True, scaling="weak",
                    config={"workload": "c2: separable Matern-3/2 x RBF, %d spatial x %d time points, state dim %d "
                                        "[CPU arm: %s]" % (Ns, T, 2 * Ns, sample), "spatial_points": Ns, "T": T},
                    cpu_baseline={"value": value, "unit": "state-steps/s", "cores": cores, "kind": "port", "sample": sample},
                    e2e={"value": value, "unit": "state-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line), flush=True)
        return
    if a.workload == "c2cvi":
        Ns = a.series
        vals = [cpu_c2cvi_ms(Ns, 4) * T / 4 for i in range(a.warmup + a.steps)][a.warmup:]
        value = float(np.mean(vals))
        sample = ("4 of %d time steps, scaled linearly (numpy oracle: oracle/filters.py + oracle/cvi.py, LAPACK threading as "
                  "numpy configures it) -- restatement, not the JAX reference" % T)
        line = dict(base, metric="CVI ELBO+natgrad step time", value=value, unit="ms", ms_per_step=value,
                    higher_is_better=False, scaling="weak",
                    config={"workload": "c2cvi: CVI iteration of a separable Matern-3/2 x RBF model, %d spatial x %d time "
                                        "points, one D = %d site block per step [CPU arm: %s]" % (Ns, T, Ns, sample),
                            "spatial_points": Ns, "T": T},
                    cpu_baseline={"value": value, "unit": "ms", "cores": cores, "kind": "port", "sample": sample},
                    e2e={"value": value, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line), flush=True)
        return
    if a.workload == "spatial":
        from oracle import dense_gp
        M, N = a.state_dim, a.series
        rng = np.random.default_rng(5)
        X, XS = rng.uniform(size=[M, 2]), rng.uniform(size=[N, 2])

        def gram(A_, B_):
            r = np.sqrt(((A_[:, None, :] - B_[None, :, :]) ** 2).sum(-1)) * np.sqrt(3.0) / 0.3
            return (1.0 + r) * np.exp(-r)
        Kzz, Ksz, Kss = gram(X, X), gram(XS, X), gram(XS, XS)
        Ts = 40
        Bm = rng.normal(size=[Ts, M, 16]) * 0.1
        P = 0.2 * Kzz[None] + Bm @ np.swapaxes(Bm, 1, 2) + 0.01 * np.eye(M)
        mm = rng.normal(size=[Ts, M, 1])
        vals = []
        for i in range(a.warmup + a.steps):
            t0 = time.perf_counter()
            dense_gp.spatial_conditional(Kzz, Ksz, Kss, np.full(Ts, 0.9), mm, P, 1e-6)
            if i >= a.warmup:
                vals.append(Ts / (time.perf_counter() - t0))
        value = float(np.mean(vals))
        sample = ("%d of %d time steps (oracle/dense_gp.py:spatial_conditional, numpy / LAPACK threading as numpy configures "
                  "it) -- restatement, not the JAX reference" % (Ts, T))
        line = dict(base, metric="spatial conditional time-steps/sec (fp64)", value=value, unit="time-steps/s",
                    ms_per_step=1e3 * T / value, higher_is_better=True, scaling="weak",
                    config={"workload": "spatial: posterior at M = %d spatial points x %d time steps carried to N = %d new "
                                        "points, full N x N covariance blocks [CPU arm: %s]" % (M, T, N, sample),
                            "M": M, "N": N, "T": T},
                    cpu_baseline={"value": value, "unit": "time-steps/s", "cores": cores, "kind": "port", "sample": sample},
                    e2e={"value": value, "unit": "time-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line), flush=True)
        return
    n = a.cpu_sample_series or max(cores * 8, 64)
    if d > 4:
        n = max(cores * 2, 16)
    rates = []
    for i in range(a.warmup + a.steps):
        r, threads, el, reps = cpu_port_rate(d, T, n, seed=i, budget_s=3.0)
        if i >= a.warmup:
            rates.append((r, el))
    value = float(np.mean([r for r, _ in rates]))
    sample = ("%d of %d series x %d steps, repeated for ~3 s per bench step, rate scaled linearly, d=%d, m=1 -- %s"
              % (n, a.series, T, d, CPU_KIND_NOTE))
    cfg = workload_config(a, sub_batch=None)
    cfg["workload"] += " [CPU arm: %s]" % sample
    line = dict(base, metric="filter+smoother state-steps/sec (fp64)", value=value, unit="state-steps/s",
                ms_per_step=1e3 * float(np.mean([el for _, el in rates])), higher_is_better=True,
                scaling=a.scaling, config=cfg,
                cpu_baseline={"value": value, "unit": "state-steps/s", "cores": threads, "kind": "port",
                              "sample": sample},
                e2e={"value": value, "unit": "state-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads to the CPU cores NVML reports as local to its GPU, BEFORE any pinned host
    buffer is allocated: under torchrun the ranks float over both sockets, and a pinned staging buffer that
    lands on the far socket sends every host<->device byte of the e2e path across the inter-socket link.
    Returns a short description for the JSON line (None when NVML / affinity is unavailable)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return "rank threads bound to the %d cores local to GPU %d" % (len(allowed), local)
    except Exception:
        pass
    return None


def series_split(a, world):
    """(series per GPU, series in total) under --scaling weak (default: --series per GPU) / strong (--series total)."""
    if a.scaling == "strong":
        if a.series % world:
            raise SystemExit("--scaling strong: --series must be divisible by the number of GPUs")
        return a.series // world, a.series
    return a.series, a.series * world


def workload_config(a, sub_batch, world=1):
    per, total = series_split(a, world)
    return {"workload": "c5: %d independent series x %d steps %s, Matern-7/2 x %d (state dim %d), m=1, "
                        "Gaussian noise, per-series lengthscales, 5%% missing" % (
                            a.series, a.T, "per GPU" if a.scaling == "weak" else "in total, split over the GPUs",
                            a.state_dim // 4, a.state_dim),
            "series_per_gpu": per, "series_total": total, "T": a.T,
            "state_dim": a.state_dim, "obs_dim": 1, "sub_batch": sub_batch,
            "outputs": "filtered (m, P) + smoothed (m, P) full state, fp64, every step materialised in HBM",
            "layout": "time-major batch [T][B][d*d] (step strides (1, B) of the C ABI)",
            "l2": "inputs+outputs per launch >> 126 MB L2 (no flush needed)",
            "parallelism": "independent series sharded over ranks (%s scaling), no data-path collective" % a.scaling}


def cvi_config(B, T):
    return {"workload": "cvi (config 4): %d blocks x %d steps per GPU, Matern-3/2 (d=2), Poisson exp-link counts, "
                        "Gauss-Hermite K=20, beta=0.1, 5%% missing" % (B, T),
            "blocks_per_gpu": B, "T": T, "state_dim": 2, "site_dim": 1, "quad_points": 20,
            "parallelism": "independent blocks per rank, no collective"}


# ------------------------------------------------------------------------------------------ GPU arm
def run_b200(a):
    import torch
    import torch.distributed as dist
    from physs_gp_b200 import ops, sdes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (b200 arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    d, T = a.state_dim, a.T
    if d % 4:
        raise SystemExit("--state-dim must be a multiple of 4 (Matern-7/2 blocks)")
    nblk = d // 4
    # weak scaling (default): every rank owns `--series` series of its own; strong: `--series` in total, split
    # over the ranks.  Series are independent: no data-path collective either way.
    n_local, n_total = series_split(a, world)
    lo = rank * n_local
    sub = min(a.sub_batch, n_local)
    starts = list(range(0, n_local, sub))

    ls_all, steps = make_hypers(n_total, nblk)
    steps = steps[:T] if T <= T_STEPS else np.resize(steps, T)
    prior = sdes.BatchedMaternSDE(4, ls_all[lo:lo + n_local])
    lam = torch.as_tensor(prior.lam(), device=dev)
    Pinf = torch.as_tensor(prior.P_inf(), device=dev)
    H = torch.as_tensor(prior.H(), device=dev)
    m0 = torch.zeros((1, d), dtype=torch.float64, device=dev)
    dt_f = torch.as_tensor(np.hstack([0.0, steps[1:]]), device=dev)     # dt[k] = t_k - t_{k-1}, dt[0] = 0
    dt_s = torch.as_tensor(np.hstack([steps[1:], 0.0]), device=dev)     # dt[k] = t_{k+1} - t_k, dt[T-1] = 0
    R = torch.full((1, 1, 1, 1), NOISE_VAR, dtype=torch.float64, device=dev)
    Ys = [device_observations(min(sub, n_local - s), T, dev, seed=1000 + lo + s) for s in starts]

    def bufs(n):
        return (ops.empty_steps(n, T, (d,), dev, True), ops.empty_steps(n, T, (d, d), dev, True),
                ops.empty_steps(n, T, (d,), dev, True), ops.empty_steps(n, T, (d, d), dev, True))
    out_full = bufs(sub)
    out_tail = bufs(n_local - starts[-1]) if n_local - starts[-1] != sub else out_full
    lml_all = torch.empty((n_local,), dtype=torch.float64, device=dev)

    ev = {"f": [], "s": []}

    def one_step(record):
        for i, s in enumerate(starts):
            n = min(sub, n_local - s)
            mf, Pf, ms, Ps = out_full if n == sub else out_tail
            disc = ops.Disc.matern(nblk, lam[s:s + n], Pinf[s:s + n])
            if record:
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
            lml, _, _ = ops.kf_filter(dt_f, Ys[i], R, H, m0, Pinf[s:s + n], disc, jitter=1e-5, out=(mf, Pf))
            if record:
                e1.record()
            ops.rts_smooth(dt_s, mf, Pf, disc, Hout=None, jitter=1e-5, out=(ms, Ps))
            if record:
                e2.record()
                ev["f"].append((e0, e1, n)), ev["s"].append((e1, e2, n))
            lml_all[s:s + n] = lml

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        one_step(False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()                                 # rank 0 waits for nvidia-smi to come up: keep the ranks together
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(a.steps):
        one_step(True)
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = torch.tensor([t_start.elapsed_time(t_end)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    elapsed_ms = float(elapsed_ms.item())
    value = n_total * T * a.steps / (elapsed_ms * 1e-3)
    assert torch.isfinite(lml_all).all(), "non-finite log marginal likelihood in the bench run"

    # per-kernel durations (this rank), roofline of the dominant kernel
    fb, sb = algorithmic_bytes(d, 1)
    f_ms = [e0.elapsed_time(e1) for e0, e1, _ in ev["f"]]
    s_ms = [e0.elapsed_time(e1) for e0, e1, _ in ev["s"]]
    f_units = [n * T for _, _, n in ev["f"]]
    s_units = [n * T for _, _, n in ev["s"]]
    peak, peak_src = measured_peak_gbs()
    kern = {}
    # which kernel family the C ABI dispatches this shape to (physs_api.cu: prefer_seq / rt_supported)
    if d <= 4:
        # time-major plain mode: the software-pipelined smoother for even d (physs_seq_impl.cuh: launch_smooth)
        kn = ("seq_filter_kernel<%d>" % d, ("seq_smooth_pipe_kernel<%d>" if d % 2 == 0 else "seq_smooth_kernel<%d>") % d)
    elif d <= 32:
        kn = ("rt_filter_kernel<%d>" % d, "rt_smooth_kernel<%d>" % d)
    else:
        kn = ("grp_filter_kernel", "grp_smooth_kernel")
    for name, msl, units, bpu in ((kn[0], f_ms, f_units, fb), (kn[1], s_ms, s_units, sb)):
        avg_ms = float(np.mean(msl))
        avg_bytes = float(np.mean(units)) * bpu
        kern[name] = {"avg_ms": avg_ms, "bytes_per_launch": avg_bytes,
                      "achieved_gbs": avg_bytes / (avg_ms * 1e-3) / 1e9, "total_ms": float(np.sum(msl))}
    dom = max(kern, key=lambda k: kern[k]["total_ms"])
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": kern[dom]["achieved_gbs"] / peak, "peak_source": peak_src,
                "traffic": recorded_traffic(d),
                "bytes_per_state_step": {"filter": fb, "smoother": sb},
                "kernels": kern,
                "whole_step_frac": (fb + sb) * (value / world) / 1e9 / peak,
                "fp64_peak_tflops_measured": ops.fp64_peak_tflops(dev)}
    # the other regime of SURVEY 8d: F(d, m) flop per state-step against the FP64 pipe measured on this box
    flops = 14.3 * d ** 3 + 4 * d * d + 6 * d + 0.67
    tf = flops * (value / world) / 1e12
    roofline["fp64"] = {"flops_per_state_step": flops, "achieved_tflops": tf,
                        "frac": tf / roofline["fp64_peak_tflops_measured"]}

    # ---------------------------------------------------------------- e2e through the host API
    e2e = None
    del out_full, out_tail
    torch.cuda.empty_cache()
    # ---- the posterior-only call (filter_and_smooth(full_state=False): smoothed mean / variance of f + lml, no
    # filtered output): packed hand-over in a workspace (physs_kf_filter_smooth_packed_f64) beside the two-output
    # call, same inputs, device-timed.  Not part of `value` (which materialises all four full-state outputs).
    post = None
    if d == 4 and not a.no_post:
        post = {}
        k_post = max(2, min(a.steps, 5))
        ws = None
        # the packed call needs 112 B per series-step of workspace instead of 320 B of filtered outputs: the whole
        # batch fits one launch (twice the resident warps of a 32,768-series sub-batch) where the two-output call
        # needs the sub-batches of the main line
        ws_need = lambda n: int(ops._lib.load().physs_kf_filter_smooth_packed_ws_bytes(n, T, n, d))   # noqa: E731
        psub = n_local if (a.post_sub_batch is None and ws_need(n_local) <= 90e9) else min(a.post_sub_batch or sub, n_local)
        Yp = Ys
        if psub != sub:
            Y_all = torch.cat([y.transpose(0, 1) for y in Ys], dim=1).transpose(0, 1)    # time-major [T][n_local][1]
            Yp = [Y_all[s:s + psub] for s in range(0, n_local, psub)]
            if psub != n_local:
                Yp = [y.transpose(0, 1).contiguous().transpose(0, 1) for y in Yp]
        for name in ("packed", "two_output"):
            ysrc, nsub = (Yp, psub) if name == "packed" else (Ys, sub)
            pstarts = list(range(0, n_local, nsub))

            def call(i, s, n):
                disc = ops.Disc.matern(nblk, lam[s:s + n], Pinf[s:s + n])
                if name == "packed":
                    return ops.kf_filter_smooth_packed(dt_f, dt_s, ysrc[i], R, H, m0, Pinf[s:s + n], disc, Hout=H,
                                                       jitter=1e-5, ws=ws)[0]
                return ops.kf_filter_smooth(dt_f, dt_s, ysrc[i], R, H, m0, Pinf[s:s + n], disc, Hout=H, jitter=1e-5)[0]
            if name == "packed":
                ws = torch.empty((ws_need(nsub),), dtype=torch.uint8, device=dev)
            for i, s in enumerate(pstarts):                                 # warm-up
                call(i, s, min(nsub, n_local - s))
            barrier()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for _ in range(k_post):
                for i, s in enumerate(pstarts):
                    lml_p = call(i, s, min(nsub, n_local - s))
            p1.record()
            barrier()
            ms_p = torch.tensor([p0.elapsed_time(p1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms_p, op=dist.ReduceOp.MAX)
            assert torch.isfinite(lml_p).all()
            rate = n_total * T * k_post / (float(ms_p.item()) * 1e-3)
            # y, R, dt in + hand-over out and back + (mean, variance) out
            bpss = 8 * (3 + 2 * (14 if name == "packed" else d * d + d) + 2)
            post[name] = {"value": rate, "unit": "state-steps/s", "steps": k_post, "series_per_launch": nsub,
                          "ms_per_step": float(ms_p.item()) / k_post, "bytes_per_state_step": bpss,
                          "achieved_gbs": bpss * (rate / world) / 1e9, "frac": bpss * (rate / world) / 1e9 / peak}
            ws = None
            torch.cuda.empty_cache()
        Yp = None
        post["api"] = ("ops.kf_filter_smooth_packed vs ops.kf_filter_smooth(Hout=H): inputs resident in HBM, "
                       "outputs = smoothed mean / variance of f [B, T] + lml [B]")
    if not a.no_e2e:
        e2e = run_e2e(a, dev, world, rank, prior, steps, starts, sub, n_local, Ys, numa)

    cpu = None
    if rank == 0 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = a.cpu_sample_series or (max(cores * 8, 64) if d <= 4 else max(cores * 2, 16))
        r, threads, el, reps = cpu_port_rate(d, T, n, budget_s=12.0)
        cpu = {"value": r, "unit": "state-steps/s", "cores": threads, "kind": "port",
               "sample": "%d of %d series x %d steps x %d repeats = %.1f s of CPU work, rate scaled linearly -- %s"
                         % (n, n_local, T, reps, el, CPU_KIND_NOTE)}

    # ------------------------------------------- the rest of BASELINE config 5 (d = 8, 16, 32) and metric (2)
    sweep, cvi_sec, c2_sec = None, None, None
    if not a.no_sweep and d == 4:
        Ys = None
        torch.cuda.empty_cache()
        sweep = {}
        for dd in (8, 16, 32):
            sweep["d%d" % dd] = sweep_point(a, dev, world, rank, dd, n_local, n_total, lo,
                                            cpu=(rank == 0 and not a.no_cpu_baseline))
            torch.cuda.empty_cache()
        cvi_sec = cvi_measure(a, dev, world, rank, local, 1000, T_STEPS, steps=max(3, min(a.steps, 10)),
                              warmup=max(3, a.warmup), with_clocks=False,
                              cpu=(rank == 0 and not a.no_cpu_baseline))
        torch.cuda.empty_cache()
        # BASELINE config 2 (one series, d = 400, m = 200, T = 5000) on the hand-written separable-prior kernels:
        # every rank runs its own replica ("replicas only"), two timed passes
        c2_sec = c2_measure(a, 200, 5000, steps=2, warmup=1, with_clocks=False,
                            cpu_base=(rank == 0 and not a.no_cpu_baseline), init_dist=False)

    if rank == 0:
        launches = 2 * len(starts) * a.steps
        line = {
            "metric": "filter+smoother state-steps/sec (fp64)", "value": value, "unit": "state-steps/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": elapsed_ms / a.steps,
            "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(a, sub, world),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": launches,
        }
        if post is not None:
            line["posterior_only"] = post
        if sweep is not None:
            line["sweep"] = sweep
            line["cvi"] = cvi_sec
            line["c2"] = c2_sec
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def wave_series(d):
    """Series the library keeps resident on the whole GPU for this state dim (one full wave of the smoother
    kernel); sub-batches are sized in whole waves so that no launch ends on a half-empty GPU."""
    from physs_gp_b200 import ops
    return ops.kf_wave_series(d, 1, d // 4)


def sweep_point(a, dev, world, rank, d, n_local, n_total, lo, cpu):
    """One timed pass of filter + smoother over this rank's n_local series at state dim d (BASELINE config 5,
    d = 8 / 16 / 32), in sub-batches of whole waves that fit HBM; per-launch CUDA events; max over ranks."""
    import torch
    import torch.distributed as dist
    from physs_gp_b200 import ops, sdes
    T, nblk = a.T, d // 4
    per_series = T * (d * d + d) * 8 * 2                       # filtered + smoothed full-state outputs
    wave = wave_series(d)
    cap = int(135e9 // per_series)
    sub = wave * max(1, cap // wave) if cap >= wave else max(32, wave // -(-wave // cap))
    sub = min(sub, n_local)
    starts = list(range(0, n_local, sub))
    ls_all, steps = make_hypers(n_total, nblk)
    steps = steps[:T] if T <= T_STEPS else np.resize(steps, T)
    prior = sdes.BatchedMaternSDE(4, ls_all[lo:lo + n_local])
    lam = torch.as_tensor(prior.lam(), device=dev)
    Pinf = torch.as_tensor(prior.P_inf(), device=dev)
    H = torch.as_tensor(prior.H(), device=dev)
    m0 = torch.zeros((1, d), dtype=torch.float64, device=dev)
    dt_f = torch.as_tensor(np.hstack([0.0, steps[1:]]), device=dev)
    dt_s = torch.as_tensor(np.hstack([steps[1:], 0.0]), device=dev)
    R = torch.full((1, 1, 1, 1), NOISE_VAR, dtype=torch.float64, device=dev)
    Y = device_observations(sub, T, dev, seed=2000 + lo)        # the same observations for every sub-batch
    # one flat allocation per output, viewed time-major for whatever batch size a sub-batch has (the ragged last
    # one re-uses the same memory: no second set of buffers)
    flat = [torch.empty((T * sub * k,), dtype=torch.float64, device=dev) for k in (d, d * d, d, d * d)]

    def out_views(n):
        shp = ((d,), (d, d), (d,), (d, d))
        return tuple(f[:T * n * int(np.prod(sh))].view((T, n) + sh).transpose(0, 1) for f, sh in zip(flat, shp))
    mf, Pf, ms, Ps = out_views(sub)
    tail = n_local - starts[-1]
    tail_bufs = None
    if tail != sub:
        tail_bufs = (device_observations(tail, T, dev, seed=2001 + lo),) + out_views(tail)

    def one_pass(record):
        ev, lmls = [], []
        for s0 in starts:
            n = min(sub, n_local - s0)
            Yv, mfv, Pfv, msv, Psv = (Y, mf, Pf, ms, Ps) if n == sub else tail_bufs
            disc = ops.Disc.matern(nblk, lam[s0:s0 + n], Pinf[s0:s0 + n])
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if record else None
            if record:
                e[0].record()
            lml, _, _ = ops.kf_filter(dt_f, Yv, R, H, m0, Pinf[s0:s0 + n], disc, jitter=1e-5, out=(mfv, Pfv))
            if record:
                e[1].record()
            ops.rts_smooth(dt_s, mfv, Pfv, disc, Hout=None, jitter=1e-5, out=(msv, Psv))
            if record:
                e[2].record()
                ev.append((e, n))
            lmls.append(lml)
        return ev, lmls

    # warm-up: the first sub-batch once (kernel load, clocks), then ONE timed pass over all n_local series
    disc0 = ops.Disc.matern(nblk, lam[:min(sub, 256)], Pinf[:min(sub, 256)])
    w = min(sub, 256)
    for _ in range(2):
        ops.kf_filter(dt_f[:200], Y[:w, :200], R, H, m0, Pinf[:w], disc0, jitter=1e-5)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    ev, lmls = one_pass(True)
    t1.record()
    torch.cuda.synchronize()
    assert all(bool(torch.isfinite(x).all()) for x in lmls), "non-finite lml in the sweep (d=%d)" % d
    el = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    ms_pass = float(el.item())
    value = n_total * T / (ms_pass * 1e-3)
    f_ms = float(np.sum([e[0].elapsed_time(e[1]) for e, _ in ev]))
    s_ms = float(np.sum([e[1].elapsed_time(e[2]) for e, _ in ev]))
    fb, sb = algorithmic_bytes(d, 1)
    peak, _ = measured_peak_gbs()
    fp64_peak = ops.fp64_peak_tflops(dev)
    flops = 14.3 * d ** 3 + 4 * d * d + 6 * d + 0.67
    per_gpu = value / world
    out = {"value": value, "unit": "state-steps/s", "state_dim": d, "series_per_gpu": n_local, "T": T,
           "sub_batch": sub, "wave_series": wave, "passes_timed": 1, "ms_per_pass": ms_pass,
           "filter_ms": f_ms, "smoother_ms": s_ms, "gpu_launches": 2 * len(starts),
           "hbm": {"bytes_per_state_step": fb + sb, "achieved_gbs": (fb + sb) * per_gpu / 1e9,
                   "frac": (fb + sb) * per_gpu / 1e9 / peak},
           "fp64": {"flops_per_state_step": flops, "achieved_tflops": flops * per_gpu / 1e12,
                    "peak_tflops_measured": fp64_peak, "frac": flops * per_gpu / 1e12 / fp64_peak}}
    out["roofline"] = {"bound": "hbm" if out["hbm"]["frac"] >= out["fp64"]["frac"] else "fp64",
                       "frac": max(out["hbm"]["frac"], out["fp64"]["frac"])}
    del Y, mf, Pf, ms, Ps, tail_bufs, flat
    if cpu:
        cores = os.cpu_count() or 1
        n = max(cores * 2, 16)
        r, threads, elc, reps = cpu_port_rate(d, T, n, budget_s=3.0)
        out["cpu_baseline"] = {"value": r, "unit": "state-steps/s", "cores": threads, "kind": "port",
                               "sample": "%d of %d series x %d steps x %d repeats = %.1f s -- %s"
                                         % (n, n_local, T, reps, elc, CPU_KIND_NOTE)}
    return out


def run_e2e(a, dev, world, rank, prior, steps, starts, sub, n_local, Ys_dev, numa=None):
    """Same job through the reference-shaped API (SDE_GP.filter_and_smooth) with HOST buffers.

    Every sub-batch is one user-level call sequence on its own CUDA stream (eight streams round-robin):
    pinned-host -> device copy of its observations, filter + smoother through the host API -> C ABI, and a
    device -> pinned-host read of the result.  The API is stream-ordered and never synchronises the host, so
    the H2D of later sub-batches, the kernels of the current ones and the D2H of earlier ones overlap on the two
    copy engines and the SMs; every byte still crosses PCIe inside the timed region.  Measured on this box: PCIe
    55 GB/s H2D, 57 GB/s D2H, 50 + 50 GB/s concurrently (tools/pcie.py): the 16 B per state-step of the
    read-back bound this number at 3.1e9 state-steps/s; sub-batch / stream count swept in tools/e2e_probe2.py."""
    import torch
    import torch.distributed as dist
    from physs_gp_b200 import data, likelihood, models, sdes

    T, d = a.T, a.state_dim
    t_host = np.cumsum(steps)
    Y_host = torch.empty((n_local, T, 1), dtype=torch.float64, pin_memory=True)
    for s0, Yd in zip(starts, Ys_dev):
        Y_host[s0:s0 + Yd.shape[0]].copy_(Yd)
    del Ys_dev
    out_mu = torch.empty((n_local, T, 1), dtype=torch.float64, pin_memory=True)
    out_var = torch.empty((n_local, T, 1), dtype=torch.float64, pin_memory=True)
    out_lml = torch.empty((n_local,), dtype=torch.float64, pin_memory=True)
    lik = likelihood.Gaussian(NOISE_VAR)
    esub = min(a.e2e_sub_batch, n_local)
    estarts = list(range(0, n_local, esub))
    streams = [torch.cuda.Stream(device=dev) for _ in range(8)]
    torch.cuda.synchronize()

    def step(readback=True):
        for i, s in enumerate(estarts):
            n = min(esub, n_local - s)
            with torch.cuda.stream(streams[i % len(streams)]):
                sub_prior = sdes.BatchedMaternSDE(4, prior.ls[s:s + n], prior.var[s:s + n])
                dat = data.TemporalData(t_host, Y_host[s:s + n, :, :, None])
                model = models.SDE_GP(dat, sub_prior, lik)
                lml, mu, var = model.filter_and_smooth(full_state=False, return_lml=True)
                if readback:
                    out_mu[s:s + n].copy_(mu[..., 0], non_blocking=True)
                    out_var[s:s + n].copy_(var[..., 0], non_blocking=True)
                out_lml[s:s + n].copy_(lml, non_blocking=True)
        torch.cuda.synchronize()

    def timed(readback):
        step(readback)                        # warm-up (allocator pools per stream, page-locking of first touch)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        k = max(1, min(a.steps, 2))
        t0 = time.perf_counter()
        for _ in range(k):
            step(readback)
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        assert bool(torch.isfinite(out_lml).all())
        return k, float(el.item())

    k, el = timed(True)
    k2, el2 = timed(False)
    per_rank_in = n_local * T * 8 + 2 * T * 8
    per_rank_out = n_local * T * 16 + n_local * 8
    return {"value": n_local * world * T * k / el, "unit": "state-steps/s", "steps": k, "host_affinity": numa,
            "h2d_bytes_per_step": per_rank_in * world, "d2h_bytes_per_step": per_rank_out * world,
            "sub_batch": esub, "streams": len(streams),
            "api": "SDE_GP.filter_and_smooth(full_state=False, return_lml=True) per sub-batch, pinned host "
                   "buffers, one CUDA stream per in-flight sub-batch",
            "result": "smoothed mean/variance of f [B,T] + lml [B] read back to pinned host memory",
            # the same calls when only the loss (lml per series) is read back and the posterior stays in HBM
            # for the next consumer (a CVI step, predict_f), as the reference's device arrays would
            "loss_only": {"value": n_local * world * T * k2 / el2, "unit": "state-steps/s", "steps": k2,
                          "h2d_bytes_per_step": per_rank_in * world, "d2h_bytes_per_step": n_local * 8 * world,
                          "result": "lml [B] read back; smoothed mean/variance left on the device"}}


# ------------------------------------------------------------------ c3: one long series, parallel in time
def run_c3(a):
    """BASELINE config 3 shape: B long series (default 1) x T steps (default 1M), derivative-augmented state
    d (default 8 = 2 x Matern-7/2), full-state Gaussian pseudo-observations (m = d, the CVI site
    configuration) -- filter + smoother by the chunked associative scan; with N ranks the series is sharded
    in TIME (one all-gather of range summaries per pass, physs_gp_b200/timeshard.py)."""
    import torch
    import torch.distributed as dist
    from physs_gp_b200 import ops, sdes, timeshard

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        comm = timeshard.TorchDist()
    else:
        comm = timeshard.SingleProcess()
    B, T, d = a.series, a.T, a.state_dim
    m = a.obs_dim or d
    nblk, L, jitter = d // 4, a.chunk_len, 1e-5
    rng = np.random.default_rng(0)
    steps = rng.uniform(0.5, 1.5, T) * DT0
    dt_f, dt_s = np.hstack([0.0, steps[1:]]), np.hstack([steps[1:], 0.0])
    prior = sdes.BatchedMaternSDE(4, np.exp(rng.uniform(np.log(0.5), np.log(2.0), (B, nblk))) * (10 * DT0),
                                  full_state_obs=(m == d))
    t0, t1 = timeshard.time_ranges(T, world)[rank]
    L = ops.even_chunk_len(t1 - t0, L)            # a chunk length that divides the range: no ragged launch
    tt = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)   # noqa: E731
    Yh = np.sin(0.01 * np.arange(t0, t1))[None, :, None] + 0.3 * rng.normal(size=(B, t1 - t0, m))
    Y = tt(Yh)
    lam, Pinf = tt(prior.lam()), tt(prior.P_inf())
    H = None if m == d else tt(prior.H())
    disc = ops.Disc.matern(nblk, lam, Pinf)
    m0 = torch.zeros((1, d), dtype=torch.float64, device=dev)
    R = NOISE_VAR * torch.eye(m, dtype=torch.float64, device=dev)[None, None]
    ws = ops.pscan_workspace(B, t1 - t0, d, L, dev)
    args = (tt(dt_f[t0:t1]), tt(dt_s[t0:t1]), Y, R, H, m0, Pinf, disc, disc)

    def step():
        return timeshard.filter_smooth(comm, ops, *args, chunk_len=L, jitter=jitter, ws=ws,
                                       cross_rank_polish=world > 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(a.warmup):
        out = step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()                                 # rank 0 waits for nvidia-smi to come up: keep the ranks together
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    el = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    ms = float(el.item()) / a.steps
    assert torch.isfinite(out[0]).all() and int(out[-1].item()) == 0, "non-finite lml or unconverged fix-up"
    value = B * T / (ms * 1e-3)
    fb, sb = algorithmic_bytes(d, m)
    flops = 14.3 * d ** 3 + 4 * m * d * d + 6 * m * m * d + 0.67 * m ** 3          # SURVEY 8d, sequential count
    peak, peak_src = measured_peak_gbs()
    fp64 = ops.fp64_peak_tflops(dev)
    # e2e: host buffers in, smoothed mean / marginal variance of every state out
    e2e = None
    if not a.no_e2e:
        Y_host = torch.empty(Y.shape, dtype=torch.float64, pin_memory=True); Y_host.copy_(Y)
        o_m = torch.empty((B, t1 - t0, d), dtype=torch.float64, pin_memory=True)
        o_v = torch.empty((B, t1 - t0, d), dtype=torch.float64, pin_memory=True)
        o_l = torch.empty((B,), dtype=torch.float64, pin_memory=True)

        def e2e_step():
            Yd = Y_host.to(dev, non_blocking=True)
            lml, mf, Pf, ms_, Ps_, st = timeshard.filter_smooth(comm, ops, args[0], args[1], Yd, *args[3:],
                                                                chunk_len=L, jitter=jitter, ws=ws,
                                                                cross_rank_polish=world > 1)
            o_m.copy_(ms_, non_blocking=True)
            o_v.copy_(torch.diagonal(Ps_, dim1=-2, dim2=-1), non_blocking=True)
            o_l.copy_(lml, non_blocking=True)
            torch.cuda.synchronize()
        e2e_step()
        barrier()
        tw = time.perf_counter()
        for _ in range(a.steps):
            e2e_step()
        elw = torch.tensor([time.perf_counter() - tw], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(elw, op=dist.ReduceOp.MAX)
        e2e = {"value": B * T * a.steps / float(elw.item()), "unit": "state-steps/s",
               "h2d_bytes_per_step": B * T * m * 8, "d2h_bytes_per_step": B * T * 2 * d * 8 + B * 8,
               "api": "timeshard.filter_smooth (pscan local/fold/finish over the C ABI), pinned host buffers",
               "result": "smoothed mean + marginal variances of the full state [B,T,d] + lml"}
    cpu = None
    if rank == 0 and not a.no_cpu_baseline:
        Ts = min(T, 100000)
        r, elc, reps = cpu_c3_rate(d, m, Ts, budget_s=8.0)
        cpu = {"value": r, "unit": "state-steps/s", "cores": 1, "kind": "port",
               "sample": "first %d of %d steps x %d repeats = %.1f s on ONE core (a single series' sequential recursion "
                         "does not thread) -- %s" % (Ts, T, reps, elc, CPU_KIND_NOTE)}
    if rank == 0:
        line = {"metric": "filter+smoother state-steps/sec (fp64)", "value": value, "unit": "state-steps/s",
                "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "c3: %d series x %d steps, state dim %d (Matern-7/2 x %d), obs dim %d, "
                                       "parallel-in-time chunked scan (chunk %d), jitter 1e-5" % (B, T, d, nblk, m, L),
                           "series": B, "T": T, "state_dim": d, "obs_dim": m, "chunk_len": L,
                           "l2": "per-pass working set %.1f GB >> 126 MB L2" % ((fb + sb) * B * T / 1e9),
                           "parallelism": "time-sharded over ranks: all-gather of range summaries (NCCL)"},
                "roofline": {"bound": "hbm", "achieved": (fb + sb) * (value / world) / 1e9, "peak": peak,
                             "unit": "GB/s", "frac": (fb + sb) * (value / world) / 1e9 / peak, "peak_source": peak_src,
                             "traffic": None,
                             "fp64": {"achieved_tflops": flops * (value / world) / 1e12, "peak_tflops": fp64,
                                      "frac": flops * (value / world) / 1e12 / fp64,
                                      "flops_per_state_step": flops, "peak_source": "physs_fp64_probe, this run"},
                             "note": "against the SEQUENTIAL algorithmic bytes / flops per state-step "
                                     "(SURVEY 8d): the scan's extra passes are overhead, not credit"},
                "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------ cvi: ELBO + natgrad step
def cvi_measure(a, dev, world, rank, local, B, T, steps, warmup, with_clocks, cpu):
    """BASELINE config 4 shape: B spatial blocks (default 1000) x T steps (default 10k), Matern-3/2 state
    (d = 2), scalar Poisson counts (exp link), Gauss-Hermite K = 20, beta = 0.1.  One step = one natural-gradient
    site update (filter + smoother, ELL gradients, theta <-> lambda, block update) + one ELBO evaluation (filter +
    smoother on the new sites, data ELL, surrogate ELL) -- vgp.py:274-282,148-157.  Returns the metric-(2) record
    (ms per step, roofline, e2e, CPU restatement beside it)."""
    import torch
    import torch.distributed as dist
    from physs_gp_b200 import cvi, sdes

    rng = np.random.default_rng(rank)
    t = np.cumsum(rng.uniform(0.5, 1.5, T) * DT0)
    prior = sdes.BatchedMaternSDE(2, np.exp(rng.uniform(np.log(0.5), np.log(2.0), (B, 1))) * (10 * DT0))
    rate = np.exp(0.5 * np.sin(0.02 * np.arange(T))[None, :] + 0.3 * rng.normal(size=(B, 1)))
    Yh = rng.poisson(rate).astype(np.float64)[..., None]
    Yh[rng.uniform(size=Yh.shape) < NAN_FRAC] = np.nan
    q = cvi.FullConjugateGaussian(t, prior, 1, B=B, device=dev, filter_type=a.filter_type)
    model = cvi.VGP(Yh, cvi.PoissonLik(1.0), q, ell_quad_points=20)

    graphed = not os.environ.get("PHYSS_CVI_EAGER")
    if graphed:
        model.compile_step(0.1)                  # natgrad + ELBO as ONE CUDA graph (VGP.compile_step)

    def step():
        if graphed:
            return model.step()
        model.natural_gradient_update(0.1)
        return model.elbo()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(warmup):
        elbo = step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0 and with_clocks:
        sampler.start()
    barrier()                                 # rank 0 waits for nvidia-smi to come up: keep the ranks together
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        elbo = step()
    e1.record()
    barrier()
    clocks = sampler.stop() if (rank == 0 and with_clocks) else None
    el = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    ms = float(el.item()) / steps
    assert torch.isfinite(elbo).all()
    # e2e: data from pinned host memory every step, ELBO read back
    Y_host = torch.empty(Yh.shape, dtype=torch.float64, pin_memory=True); Y_host.copy_(torch.as_tensor(Yh))
    o_elbo = torch.empty((B,), dtype=torch.float64, pin_memory=True)

    pipelined = graphed and not os.environ.get("PHYSS_CVI_E2E_SERIAL")

    def e2e_step():
        model.set_data(Y_host)
        o_elbo.copy_(step(), non_blocking=True)
        torch.cuda.synchronize()
    e2e_step(); barrier()
    tw = time.perf_counter()
    if pipelined:
        # every step's data still comes from pinned host memory; the upload of step i + 1 runs on a copy stream while
        # step i computes (VGP.stage_data / commit_data), the ELBO of every step is read back before the next commit
        model.stage_data(Y_host)
        for i in range(steps):
            model.commit_data()
            if i + 1 < steps:
                model.stage_data(Y_host)
            o_elbo.copy_(step(), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        torch.cuda.synchronize()
    else:
        for _ in range(steps):
            e2e_step()
    elw = torch.tensor([time.perf_counter() - tw], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elw, op=dist.ReduceOp.MAX)
    peak, peak_src = measured_peak_gbs()
    d, D = 2, 1
    fb, sb = algorithmic_bytes(d, 1)
    # two posterior passes (natgrad + ELBO) each with time-varying site noise R_k (+8 B) ; site update: sites in/out
    # + posterior read; ELLs: posterior read twice + data
    byt = 2 * (fb + 8 + sb) + 8 * (2 * (D * D + D) + (D * D + D)) + 8 * (2 * (D * D + D) + 1 + (D * D + D))
    cfg = cvi_config(B, T)
    cfg["l2"] = "per-step working set %.1f GB >> 126 MB L2" % (byt * B * T / 1e9)
    rec = {"metric": "CVI ELBO+natgrad step time", "value": ms, "unit": "ms", "ms_per_step": ms, "steps": steps,
           "warmup": warmup, "higher_is_better": False, "config": cfg,
           "state_steps_per_s": world * B * T / (ms * 1e-3),
           "roofline": {"bound": "hbm", "achieved": byt * B * T / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": byt * B * T / (ms * 1e-3) / 1e9 / peak, "peak_source": peak_src, "traffic": None,
                        "bytes_per_state_step": byt},
           "cpu_baseline": None,
           "e2e": {"value": 1e3 * float(elw.item()) / steps, "unit": "ms", "h2d_bytes_per_step": B * T * 8 * world,
                   "d2h_bytes_per_step": B * 8 * world,
                   "api": ("VGP.compile_step(0.1) once, then VGP.set_data + VGP.step() (one CUDA graph per iteration)"
                           if graphed else "VGP.natural_gradient_update(0.1) + VGP.elbo()") +
                          (", data from pinned host memory; the upload of step i + 1 (VGP.stage_data, copy stream) overlaps "
                           "step i, VGP.commit_data swaps it in" if pipelined else ", data from pinned host memory")},
           "clocks": clocks}
    if graphed and world == 1 and not os.environ.get("PHYSS_CVI_NO_REUSE"):
        # beside the standard step (NOT the headline value): the training-loop form that serves the posterior of iteration
        # i's ELBO to iteration i + 1's natural-gradient step -- same sites, same numbers, one posterior pass per iteration
        q2 = cvi.FullConjugateGaussian(t, prior, 1, B=B, device=dev, filter_type=a.filter_type)
        m2 = cvi.VGP(Yh, cvi.PoissonLik(1.0), q2, ell_quad_points=20)
        m2.compile_step(0.1, reuse_posterior=True)
        for _ in range(warmup):
            m2.step()
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(steps):
            eb = m2.step()
        r1.record()
        torch.cuda.synchronize()
        assert torch.isfinite(eb).all()
        rec["reuse_posterior"] = {"ms_per_step": r0.elapsed_time(r1) / steps, "unit": "ms",
                                  "api": "VGP.compile_step(0.1, reuse_posterior=True): the filter + smoother pass of the "
                                         "ELBO of iteration i feeds the natural-gradient step of iteration i + 1 (the "
                                         "reference runs it twice per iteration, vgp.py:274-282,148-157); bitwise the "
                                         "same sites and ELBOs (tests/test_gpu_cvi.py)"}
        del m2, q2
    if cpu:
        n = a.cpu_sample_series or max(os.cpu_count() or 1, 8)
        ms_full, threads, elc, reps, ms_s = cpu_cvi_step_ms(B, T, n, budget_s=6.0)
        rec["cpu_baseline"] = {"value": ms_full, "unit": "ms", "cores": threads, "kind": "port",
                               "sample": "%d of %d blocks x %d steps x %d iterations = %.1f s (%.1f ms per iteration of "
                                         "the sample, scaled linearly to %d blocks; oracle/cvi_vec.py: C port of filter + "
                                         "smoother, vectorised numpy site algebra) -- restatement, not the JAX reference"
                                         % (n, B, T, reps, elc, ms_s, B)}
    return rec


def run_cvi(a):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rec = cvi_measure(a, dev, world, rank, local, a.series, a.T, a.steps, a.warmup, with_clocks=True,
                      cpu=(rank == 0 and not a.no_cpu_baseline))
    if rank == 0:
        line = dict(rec, n_gpus=world, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                    gpu_launches=None)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------- c2: separable spatio-temporal CVI (d = 2 Ns)
def c2_measure(a, Ns, T, steps, warmup, with_clocks, cpu_base, init_dist):
    """BASELINE config 2 shape: Matern-3/2 (time) x RBF (space), Ns = --series spatial points (default 200) x T
    time points (default 5000): state d = 2 Ns, observations / CVI sites over f at the Ns points (m = Ns) with a
    full time-varying site covariance R_k [m, m].  One step = the posterior pass every CVI iteration runs
    (SDE_GP.posterior_blocks of the surrogate: filter + smoother, large-block path).  The D = 200 site-update /
    ELBO kernels are not built yet (DESIGN.md status table), so this workload reports the first metric
    (state-steps/s), not the CVI step time.  A single series: every rank runs its own replica."""
    import torch
    import torch.distributed as dist
    from physs_gp_b200 import data, kernels as K, likelihood, models, ops, sdes

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and init_dist:
        dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(0)
    Xs = rng.uniform(size=(Ns, 2))
    D2 = ((Xs[:, None, :] - Xs[None, :, :]) ** 2).sum(-1)
    Ks = np.exp(-0.5 * D2 / 0.2 ** 2) + 1e-6 * np.eye(Ns)
    prior = sdes.LTI_SDE(sdes.Independent([K.SpatioTemporalSeperableKernel(K.Matern32(10 * DT0, 1.0), Ks)]))
    t = DT0 * np.arange(1, T + 1)                            # regular grid: one distinct (A, Q) pair
    Yh = np.sin(0.02 * np.arange(T))[:, None] * np.cos(3 * Xs[:, 0])[None, :] + 0.3 * rng.normal(size=(T, Ns))
    Yh[rng.uniform(size=Yh.shape) < NAN_FRAC] = np.nan
    Y = torch.as_tensor(Yh, device=dev)
    G = 0.05 * torch.randn((T, Ns, Ns), dtype=torch.float64, device=dev, generator=torch.Generator(dev).manual_seed(1))
    R = G @ G.transpose(-1, -2) + NOISE_VAR * torch.eye(Ns, dtype=torch.float64, device=dev)   # site covariances

    from physs_gp_b200 import settings as _settings
    if os.environ.get("PHYSS_C2_LIBRARY"):
        _settings.kron_kernels = False                       # A-B run: the cuBLAS / cuSOLVER-backed library path
    kron = bool(_settings.kron_kernels)

    def step(Yd):
        model = models.SDE_GP(data.TemporalData(t, Yd[:, :, None]), prior, likelihood.BlockDiagonalGaussian(R))
        return model.filter_and_smooth(full_state=False, return_lml=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(warmup):
        out = step(Y)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0 and with_clocks:
        sampler.start()
    barrier()                                 # rank 0 waits for nvidia-smi to come up: keep the ranks together
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step(Y)
    e1.record()
    barrier()
    clocks = sampler.stop() if (rank == 0 and with_clocks) else None
    el = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    ms = float(el.item()) / steps
    assert torch.isfinite(out[0]).all()
    d, m = 2 * Ns, Ns
    flops = 14.3 * d ** 3 + 4 * m * d * d + 6 * m * m * d + 0.67 * m ** 3
    fp64 = ops.fp64_peak_tflops(dev)
    Y_host = torch.empty((T, Ns), dtype=torch.float64, pin_memory=True); Y_host.copy_(Y)
    o_mu = torch.empty((T, Ns), dtype=torch.float64, pin_memory=True)
    o_var = torch.empty((T, Ns), dtype=torch.float64, pin_memory=True)

    def e2e_step():
        lml, mu, var = step(Y_host.to(dev, non_blocking=True))
        o_mu.copy_(mu[..., 0], non_blocking=True)
        o_var.copy_(torch.diagonal(var, dim1=-2, dim2=-1), non_blocking=True)
        torch.cuda.synchronize()
    e2e_step(); barrier()
    tw = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    elw = (time.perf_counter() - tw) / steps
    value = world * T / (ms * 1e-3)
    cpu = None
    if rank == 0 and cpu_base:
        r, elc = cpu_c2_rate(Ns, 8)
        cpu = {"value": r, "unit": "state-steps/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "8 of %d time steps = %.1f s (numpy oracle oracle/filters.py, LAPACK threading as numpy "
                         "configures it) -- restatement, not the JAX reference" % (T, elc)}
    if rank == 0:
        line = {"metric": "filter+smoother state-steps/sec (fp64)", "value": value, "unit": "state-steps/s",
                "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "c2: separable Matern-3/2 x RBF, %d spatial x %d time points, state dim %d, "
                                       "obs dim %d, full time-varying site covariance, 5%% missing" % (Ns, T, d, m),
                           "spatial_points": Ns, "T": T, "state_dim": d, "obs_dim": m,
                           "l2": "filtered covariances %.1f GB >> 126 MB L2" % (T * d * d * 8 / 1e9),
                           "parallelism": "replicas only (one series; no collective)"},
                "roofline": {"bound": "tensor", "achieved": flops * (value / world) / 1e12, "peak": fp64,
                             "unit": "TFLOP/s", "frac": flops * (value / world) / 1e12 / fp64,
                             "peak_source": "physs_fp64_probe (FP64 FMA pipe), this run", "traffic": None,
                             "note": "dense sequential flop count per state-step (SURVEY 8d: the reference multiplies "
                                     "the d x d Kronecker matrices out); " + (
                                         "hand-written persistent cooperative kernels (csrc/physs_kron.cu: Kronecker-"
                                         "structured predict, shared-memory Cholesky, DMMA tile GEMMs), no library calls"
                                         if kron else "products and factorisations are cuBLAS / cuSOLVER fp64 calls "
                                         "enqueued per step (PHYSS_C2_LIBRARY=1)")},
                "cpu_baseline": cpu,
                "e2e": {"value": world * T / elw, "unit": "state-steps/s", "h2d_bytes_per_step": T * Ns * 8 * world,
                        "d2h_bytes_per_step": 2 * T * Ns * 8 * world,
                        "api": "SDE_GP.filter_and_smooth(full_state=False, return_lml=True), pinned host buffers"},
                "clocks": clocks,
                # kernels of libphyss_b200.so per step: the persistent filter + last-step emit + one (gain, recursion)
                # pair per smoother time chunk of 4 x SM-count steps
                "gpu_launches": (steps * (2 + 2 * -(-(T - 1) // (4 * torch.cuda.get_device_properties(dev).multi_processor_count)))
                                 if kron else None)}
        return line
    return None
    if world > 1:
        dist.destroy_process_group()


def cpu_c2cvi_ms(Ns, T_sample):
    """ms per CVI iteration of the numpy oracle on T_sample steps of the config-2 shape (natgrad + ELBO)."""
    from oracle import cvi as ocvi
    from oracle import filters as of
    from oracle import sde as osde
    rng = np.random.default_rng(0)
    Xs = rng.uniform(size=(Ns, 2))
    D2 = ((Xs[:, None, :] - Xs[None, :, :]) ** 2).sum(-1)
    Ks = np.exp(-0.5 * D2 / 0.2 ** 2) + 1e-6 * np.eye(Ns)
    prior = osde.LTI_SDE([osde.SpaceTimeSeparable(osde.Matern32(10 * DT0, 1.0), Ks)])
    t = DT0 * np.arange(1, T_sample + 1)
    Y = rng.normal(size=(T_sample, Ns))
    Yt = 1e-5 * np.ones((T_sample, Ns)); Vt = np.tile(np.eye(Ns), [T_sample, 1, 1])
    R = NOISE_VAR * np.eye(Ns)
    t0 = time.perf_counter()
    _, qm, qv = of.filter_and_smooth(prior, t, Yt, Vt)
    g = [ocvi.gaussian_ell_and_grads(Y[i], R, None, qm[i][:, 0], qv[i]) for i in range(T_sample)]
    Yt, Vt = ocvi.cvi_step(Yt, Vt, qm[:, :, 0], qv, np.stack([x[1] for x in g]), np.stack([x[2] for x in g]), 0.5)
    lml, qm, qv = of.filter_and_smooth(prior, t, Yt, Vt)
    ell = sum(ocvi.gaussian_ell_and_grads(Y[i], R, None, qm[i][:, 0], qv[i])[0] for i in range(T_sample))
    ocvi.elbo(ell, ocvi.surrogate_ell(Yt, Vt, qm[:, :, 0], qv), lml)
    return 1e3 * (time.perf_counter() - t0)


def run_c2cvi(a):
    """BASELINE config 2 as the reference runs it: the CVI iteration of a spatio-temporal separable Matern-3/2 x RBF
    model, 200 spatial x 5000 time points, ONE D = 200 site block per time step, Gaussian likelihood, 5 % missing.
    One step = natural-gradient site update (posterior pass on the separable-prior kernels, theta <-> lambda and block
    update on the large-block site kernels) + ELBO (posterior pass on the new sites, data ELL, surrogate ELL).
    Every rank runs its own replica ("replicas only")."""
    import torch
    import torch.distributed as dist
    from physs_gp_b200 import cvi, kernels as K, ops, sdes

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    Ns, T = a.series, a.T
    rng = np.random.default_rng(0)
    Xs = rng.uniform(size=(Ns, 2))
    D2 = ((Xs[:, None, :] - Xs[None, :, :]) ** 2).sum(-1)
    Ks = np.exp(-0.5 * D2 / 0.2 ** 2) + 1e-6 * np.eye(Ns)
    prior = sdes.LTI_SDE(sdes.Independent([K.SpatioTemporalSeperableKernel(K.Matern32(10 * DT0, 1.0), Ks)]))
    t = DT0 * np.arange(1, T + 1)
    Yh = np.sin(0.02 * np.arange(T))[:, None] * np.cos(3 * Xs[:, 0])[None, :] + 0.3 * rng.normal(size=(T, Ns))
    Yh[rng.uniform(size=Yh.shape) < NAN_FRAC] = np.nan
    q = cvi.FullConjugateGaussian(t, prior, Ns, B=1, device=dev)
    model = cvi.VGP(Yh[None], cvi.GaussianLik(NOISE_VAR * np.eye(Ns)), q)

    def step():
        model.natural_gradient_update(0.5)
        return model.elbo()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(a.warmup):
        elbo = step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        elbo = step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    el = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    ms = float(el.item()) / a.steps
    assert torch.isfinite(elbo).all()
    Y_host = torch.empty((1, T, Ns), dtype=torch.float64, pin_memory=True); Y_host.copy_(torch.as_tensor(Yh)[None])
    o_elbo = torch.empty((1,), dtype=torch.float64, pin_memory=True)

    def e2e_step():
        model.set_data(Y_host)
        o_elbo.copy_(step(), non_blocking=True)
        torch.cuda.synchronize()
    e2e_step(); barrier()
    tw = time.perf_counter()
    for _ in range(a.steps):
        e2e_step()
    elw = 1e3 * (time.perf_counter() - tw) / a.steps
    reuse = None
    if world == 1 and not os.environ.get("PHYSS_CVI_NO_REUSE"):
        # beside the standard iteration (not in `value`): posterior of iteration i's ELBO served to iteration i + 1
        model.reuse_posterior = True
        step(); torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(a.steps):
            eb = step()
        r1.record()
        torch.cuda.synchronize()
        assert torch.isfinite(eb).all()
        reuse = {"ms_per_step": r0.elapsed_time(r1) / a.steps, "unit": "ms",
                 "api": "VGP.reuse_posterior = True: one filter + smoother pass per iteration instead of the reference's two "
                        "(vgp.py:274-282,148-157), same sites and ELBOs"}
        model.reuse_posterior = False
        model.invalidate()
    d, m = 2 * Ns, Ns
    # two posterior passes + (2 + 1) SPD inverses of D x D per step (chol D^3/3 + inverse 2 D^3/3 ... = D^3 each, FMA = 2)
    flops = 2 * (14.3 * d ** 3 + 4 * m * d * d + 6 * m * m * d + 0.67 * m ** 3) + 3 * 2.0 * m ** 3
    fp64 = ops.fp64_peak_tflops(dev)
    cpu = None
    if rank == 0 and not a.no_cpu_baseline:
        ms_s = cpu_c2cvi_ms(Ns, 4)
        cpu = {"value": ms_s * T / 4, "unit": "ms", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "4 of %d time steps = %.1f s (numpy oracle: oracle/filters.py + oracle/cvi.py, LAPACK threading "
                         "as numpy configures it), scaled linearly -- restatement, not the JAX reference" % (T, ms_s / 1e3)}
    if rank == 0:
        print(json.dumps({
            "metric": "CVI ELBO+natgrad step time", "value": ms, "unit": "ms", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "c2cvi: CVI iteration (natgrad + ELBO) of a separable Matern-3/2 x RBF model, %d spatial x "
                                   "%d time points, one D = %d site block per time step, Gaussian likelihood, 5%% missing"
                                   % (Ns, T, Ns), "spatial_points": Ns, "T": T, "state_dim": d, "site_block": m,
                       "l2": "filtered covariances %.1f GB >> 126 MB L2" % (T * d * d * 8 / 1e9),
                       "parallelism": "replicas only (one series; no collective)"},
            "state_steps_per_s": world * T / (ms * 1e-3),
            "roofline": {"bound": "tensor", "achieved": flops * T / (ms * 1e-3) / 1e12, "peak": fp64, "unit": "TFLOP/s",
                         "frac": flops * T / (ms * 1e-3) / 1e12 / fp64, "traffic": None,
                         "peak_source": "physs_fp64_probe (FP64 FMA pipe), this run",
                         "note": "dense flop count of two posterior passes (SURVEY 8d) + three D x D SPD inverses per block"},
            "cpu_baseline": cpu, "reuse_posterior": reuse,
            "e2e": {"value": elw, "unit": "ms", "h2d_bytes_per_step": T * Ns * 8 * world, "d2h_bytes_per_step": 8 * world,
                    "api": "VGP.set_data + VGP.natural_gradient_update(0.5) + VGP.elbo(), pinned host buffers"},
            "clocks": clocks, "gpu_launches": None}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_c2(a):
    line = c2_measure(a, a.series, a.T, a.steps, a.warmup, with_clocks=True, cpu_base=not a.no_cpu_baseline,
                      init_dist=True)
    if line is not None:
        print(json.dumps(line), flush=True)


# ------------------------------------------------ c3cvi: physics-informed CVI step on one long series
def run_c3cvi(a):
    """BASELINE config 3, the CVI side: damped-oscillator collocation model on ONE series of T steps (default
    1M) -- Matern-7/2 state (x, x_t, x_tt, x_ttt), full-state sites (m = d = 4), sparse noisy observations of x,
    collocation residual x_tt + (g/l) sin x + b x_t = 0 at every step, Laplace-Gauss-Newton curvature.
    One step = natural_gradient_update (posterior by the parallel-in-time scan + closed-form collocation ELL
    gradients + site update) + ELBO (second posterior pass + data / surrogate ELLs)."""
    import torch
    from physs_gp_b200 import cvi, sdes, settings

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    T = a.T
    # full-state sites carry 1 / ng_jitter variances on the unobserved derivatives, so the surrogate filter mixes
    # slowly in those directions: longer chunks and more fix-up passes than the defaults
    settings.pscan_chunk_len = a.chunk_len if a.chunk_len != 256 else 512
    settings.pscan_polish = 8
    rng = np.random.default_rng(0)
    t = 0.02 * np.arange(1, T + 1)
    x_true = 1.0 * np.exp(-0.01 * t) * np.cos(3.0 * t)                  # decaying oscillation (stand-in signal)
    Y = np.full((1, T, 2), np.nan)
    obs = np.arange(0, T, 10)
    Y[0, obs, 0] = x_true[obs] + 0.05 * rng.normal(size=len(obs))
    Y[0, :, 1] = 0.0
    prior = sdes.BatchedMaternSDE(4, np.array([[0.6]]), np.array([[2.0]]), full_state_obs=True)
    q = cvi.FullConjugateGaussian(t, prior, 4, B=1, device=dev, filter_type="b200_parallel")
    lik = cvi.DampedPendulumLik(g=9.81, l=1.0, b=0.3, var_obs=0.05 ** 2, var_col=0.5 ** 2)
    model = cvi.VGP(Y, lik, q)

    def step():
        model.natural_gradient_update(0.5, enforce_psd_type='laplace_gauss_newton_delta_u')
        return model.elbo()
    for _ in range(a.warmup):
        elbo = step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        elbo = step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / a.steps
    assert torch.isfinite(elbo).all()
    Y_host = torch.empty((1, T, 2), dtype=torch.float64, pin_memory=True); Y_host.copy_(torch.as_tensor(Y))
    o = torch.empty((1,), dtype=torch.float64, pin_memory=True)

    def e2e_step():
        model.set_data(Y_host)
        o.copy_(step(), non_blocking=True)
        torch.cuda.synchronize()
    e2e_step()
    tw = time.perf_counter()
    for _ in range(a.steps):
        e2e_step()
    elw = (time.perf_counter() - tw) / a.steps
    peak, peak_src = measured_peak_gbs()
    d = D = 4
    fb, sb = algorithmic_bytes(d, d)
    byt = 2 * (fb + sb) + 8 * (2 * (D * D + D) + (D * D + D) + 2 + (D * D + D)) + 8 * (2 * (D * D + D) + 2 + (D * D + D))
    line = {"metric": "CVI ELBO+natgrad step time", "value": ms, "unit": "ms", "n_gpus": 1, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "c3cvi (config 3): damped-oscillator collocation CVI, 1 series x %d steps, "
                                   "Matern-7/2 derivative state d=4, full-state sites, Gauss-Newton curvature, "
                                   "parallel-in-time scan" % T, "T": T, "state_dim": 4, "site_dim": 4,
                       "l2": "per-step working set %.1f GB >> 126 MB L2" % (byt * T / 1e9),
                       "parallelism": "single GPU (time sharding: --workload c3)"},
            "state_steps_per_s": T / (ms * 1e-3),
            "roofline": {"bound": "hbm", "achieved": byt * T / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": byt * T / (ms * 1e-3) / 1e9 / peak, "peak_source": peak_src, "traffic": None,
                         "bytes_per_state_step": byt},
            "cpu_baseline": None,
            "e2e": {"value": 1e3 * elw, "unit": "ms", "h2d_bytes_per_step": T * 16, "d2h_bytes_per_step": 8,
                    "api": "VGP.natural_gradient_update(0.5, 'laplace_gauss_newton_delta_u') + VGP.elbo()"},
            "clocks": clocks, "gpu_launches": None}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ grad: lml value + hyper-parameter gradient
def run_grad(a):
    """value and gradient of the log marginal likelihood for a batch of series (the C5
    shape: Matern-7/2, d = 4, m = 1, per-series lengthscales, 5 % missing) -- one forward filter launch and one
    reverse launch (physs_kf_filter_vjp_f64).  Metric: state-steps/s of the (value, gradient) pair."""
    import torch
    from physs_gp_b200 import ops, sdes
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        raise SystemExit("--workload grad is a single-GPU workload")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    d, T, B = a.state_dim, a.T, a.series
    nblk = d // 4
    ls_all, steps = make_hypers(B, nblk)
    steps = steps[:T] if T <= T_STEPS else np.resize(steps, T)
    prior = sdes.BatchedMaternSDE(4, ls_all)
    lam = torch.as_tensor(prior.lam(), device=dev)
    Pinf = torch.as_tensor(prior.P_inf(), device=dev)
    H = torch.as_tensor(prior.H(), device=dev)
    m0 = torch.zeros((1, d), dtype=torch.float64, device=dev)
    dt_f = torch.as_tensor(np.hstack([0.0, steps[1:]]), device=dev)
    R = torch.full((1, 1, 1, 1), NOISE_VAR, dtype=torch.float64, device=dev)
    Y = device_observations(B, T, dev, seed=1000)
    disc = ops.Disc.matern(nblk, lam, Pinf)
    mf = ops.empty_steps(B, T, (d,), dev, True)
    Pf = ops.empty_steps(B, T, (d, d), dev, True)
    ev = []

    def step(record):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if record else None
        if record:
            e[0].record()
        lml, _, _ = ops.kf_filter(dt_f, Y, R, H, m0, Pinf, disc, jitter=1e-5, out=(mf, Pf))
        if record:
            e[1].record()
        g = ops.kf_filter_vjp(dt_f, Y, R, H, m0, Pinf, disc, mf, Pf, jitter=1e-5)
        if record:
            e[2].record()
            ev.append(e)
        return lml, g

    for _ in range(a.warmup):
        step(False)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(a.steps):
        lml, g = step(True)
    t1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = t0.elapsed_time(t1)
    assert torch.isfinite(lml).all() and torch.isfinite(g["glam"]).all()
    value = B * T * a.steps / (ms * 1e-3)
    f_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    v_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    peak, peak_src = measured_peak_gbs()
    fb = 8 * (d * d + d + 3)                       # filter: y, R, dt in; (m, P) out
    vb = 8 * (d * d + d + 3)                       # reverse: (m, P)[k-1], y, R, dt in; reductions stay on chip
    kern = {"seq_filter_kernel<%d>" % d: {"avg_ms": f_ms, "achieved_gbs": B * T * fb / (f_ms * 1e-3) / 1e9},
            "kf_vjp_kernel<%d>" % d: {"avg_ms": v_ms, "achieved_gbs": B * T * vb / (v_ms * 1e-3) / 1e9}}
    dom = max(kern, key=lambda k: kern[k]["avg_ms"])
    line = {"metric": "lml value+gradient state-steps/sec (fp64)", "value": value, "unit": "state-steps/s",
            "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "grad: %d series x %d steps, Matern-7/2 (state dim %d), m=1, 5%% missing: lml and its "
                                   "gradient w.r.t. (lam, Pinf, H, R, m0, P0) per series" % (B, T, d),
                       "series": B, "T": T, "state_dim": d, "layout": "time-major",
                       "l2": "per-launch working set >> 126 MB L2"},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_gbs"], "peak": peak,
                         "unit": "GB/s", "frac": kern[dom]["achieved_gbs"] / peak, "peak_source": peak_src,
                         "traffic": None, "bytes_per_state_step": {"filter": fb, "reverse": vb}, "kernels": kern},
            "cpu_baseline": None, "e2e": None, "clocks": clocks, "gpu_launches": 2 * a.steps}
    print(json.dumps(line), flush=True)


def run_spatial(a):
    """SURVEY section 8 row f3 (second half): the spatial conditional behind the smoother at the config-2 shape -- the
    posterior at M = 200 spatial points of each of T = 5000 steps carried to N new points
    (physs_spatial_conditional_f64; --series = N, --state-dim = M).  Metric: conditioned time steps per second, full
    N x N covariance blocks; the diagonal-only variant beside it.  Roofline: FP64 (DMMA) -- 2 M M N + N (N + 32) M flops
    per step (lower triangle of the second product in 32-wide tiles) against the measured FP64 peak."""
    import torch
    from physs_gp_b200 import spatial
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        raise SystemExit("--workload spatial is a single-GPU workload (time steps are independent: shard T)")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    M, N, T = a.state_dim, a.series, a.T
    rng = np.random.default_rng(5)
    X, XS = rng.uniform(size=[M, 2]), rng.uniform(size=[N, 2])

    def gram(A, B):
        r = np.sqrt(((A[:, None, :] - B[None, :, :]) ** 2).sum(-1)) * np.sqrt(3.0) / 0.3
        return (1.0 + r) * np.exp(-r)
    Kzz, Ksz, Kss = gram(X, X), gram(XS, X), gram(XS, XS)
    g = torch.Generator(device=dev).manual_seed(3)
    Bm = torch.randn((T, M, 16), dtype=torch.float64, device=dev, generator=g) * 0.1
    P = 0.2 * torch.as_tensor(Kzz, device=dev)[None] + Bm @ Bm.transpose(1, 2) + 0.01 * torch.eye(M, dtype=torch.float64, device=dev)
    m = torch.randn((T, M, 1), dtype=torch.float64, device=dev, generator=g)
    del Bm

    def timed(diagonal):
        for _ in range(a.warmup):
            spatial.spatial_conditional_block(Kzz, Ksz, Kss, 0.9, m, P, diagonal=diagonal, jitter=1e-6)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(a.steps):
            mu, var = spatial.spatial_conditional_block(Kzz, Ksz, Kss, 0.9, m, P, diagonal=diagonal, jitter=1e-6)
        t1.record()
        torch.cuda.synchronize()
        assert torch.isfinite(var).all()
        return t0.elapsed_time(t1) / a.steps

    sampler = ClockSampler(0)
    sampler.start()
    ms_full = timed(False)
    clocks = sampler.stop()
    ms_diag = timed(True)
    nt = (N + 31) // 32
    flops_full = T * (2.0 * M * M * N + 2.0 * M * 32 * 32 * nt * (nt + 1) / 2)
    flops_diag = T * (2.0 * M * M * N + 2.0 * M * N)
    fp64_peak = 34.8
    try:
        fp64_peak = float(json.load(open(os.path.join(ROOT, "profiles", "fp64_peak.json")))["tflops"])
    except Exception:
        pass
    cpu = None
    if not a.no_cpu_baseline:
        from oracle import dense_gp
        Ts = 40
        t0 = time.perf_counter()
        dense_gp.spatial_conditional(Kzz, Ksz, Kss, np.full(Ts, 0.9), m[:Ts].cpu().numpy(), P[:Ts].cpu().numpy(), 1e-6)
        dt = time.perf_counter() - t0
        cpu = {"value": Ts / dt, "unit": "time-steps/s", "cores": _host_threads(), "kind": "port",
               "sample": "%d of %d steps (%.1f s): oracle/dense_gp.py:spatial_conditional, numpy / LAPACK -- "
                         "restatement, not the JAX reference" % (Ts, T, dt)}
    line = {"metric": "spatial conditional time-steps/sec (fp64)", "value": T / (ms_full * 1e-3), "unit": "time-steps/s",
            "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_full, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "spatial: posterior at M = %d spatial points x %d time steps carried to N = %d new "
                                   "points, full N x N covariance blocks (config-2 shape)" % (M, T, N),
                       "M": M, "N": N, "T": T, "l2": "P [T, M, M] = %.1f GB >> 126 MB L2" % (T * M * M * 8 / 1e9)},
            "roofline": {"bound": "tensor", "kernel": "kron_spatial_cond_kernel", "achieved": flops_full / (ms_full * 1e-3) / 1e12,
                         "peak": fp64_peak, "unit": "TFLOP/s", "frac": flops_full / (ms_full * 1e-3) / 1e12 / fp64_peak,
                         "peak_source": "FP64 DFMA probe of round 1 (tools / physs_fp64_probe), not the bf16 figure of "
                                        "MEASURED_PEAKS.json: the path computes in f64", "traffic": None,
                         "hbm_gbs": T * 8.0 * (M * M + M + N * N + N) / (ms_full * 1e-3) / 1e9},
            "diagonal_only": {"ms": ms_diag, "value": T / (ms_diag * 1e-3), "tflops": flops_diag / (ms_diag * 1e-3) / 1e12,
                              "hbm_gbs": T * 8.0 * (M * M + M + 2 * N) / (ms_diag * 1e-3) / 1e9},
            "cpu_baseline": cpu, "e2e": None, "clocks": clocks, "gpu_launches": 2 * a.steps}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ c1: BASELINE configs[0]
def _c1_problem(T):
    rng = np.random.default_rng(0)
    t = np.cumsum(rng.uniform(0.5, 1.5, T) * DT0)
    y = np.sin(0.5 * t) + 0.3 * rng.normal(size=T)
    y[rng.uniform(size=T) < 0.05] = np.nan
    return t, y, 10 * DT0, 1.0            # lengthscale, variance


def cpu_c1_ms(T, budget_s=5.0):
    """ms per exact lml + smoothed posterior of ONE Matern-3/2 series (C port of the oracle, one core)."""
    from oracle import c_oracle
    from physs_gp_b200 import sdes
    c_oracle.build()
    t, y, ls, var = _c1_problem(T)
    prior = sdes.BatchedMaternSDE(2, np.array([[ls]]), np.array([[var]]))
    args = (2, prior.lam(), prior.P_inf(), prior.H(), t, y[None, :, None], NOISE_VAR * np.eye(1))
    c_oracle.filter_smooth(*args, jitter=1e-5, full_state=False, nthreads=1)
    el, reps = 0.0, 0
    while el < budget_s and reps < 2000:
        t0 = time.perf_counter()
        c_oracle.filter_smooth(*args, jitter=1e-5, full_state=False, nthreads=1)
        el += time.perf_counter() - t0
        reps += 1
    return el / reps * 1e3, el, reps


def run_c1(a):
    """BASELINE configs[0] (the reference's own CPU-runnable case): ONE 1-D temporal Matern-3/2 GP, Gaussian
    likelihood, N = 10k, exact marginal likelihood + smoothed posterior through the host API
    (`SDE_GP.filter_and_smooth(return_lml=True)`).  One series cannot fill a GPU: the sequential kernel walks it with
    one thread (filter_type='b200'), the chunked scan spreads it over the chip ('b200_parallel'); both are timed,
    the faster one is `value`.  A parity-test shape, not a throughput shape -- reported for completeness."""
    import torch
    from physs_gp_b200 import data, kernels, likelihood, models, sdes
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("--workload c1 is one series: single GPU")
    torch.cuda.set_device(0)
    T = a.T
    t, y, ls, var = _c1_problem(T)
    res = {}
    for ft in ("b200", "b200_parallel"):
        prior = sdes.LTI_SDE(sdes.Independent([kernels.Matern32(ls, var)]))
        model = models.SDE_GP(data.TemporalData(t, y[:, None, None]), prior, likelihood.Gaussian(NOISE_VAR),
                              filter_type=ft)
        for _ in range(max(3, a.warmup)):
            lml, mu, v = model.filter_and_smooth(return_lml=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(a.steps):
            lml, mu, v = model.filter_and_smooth(return_lml=True)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / a.steps * 1e3
        assert bool(torch.isfinite(lml).all()) and bool(torch.isfinite(mu).all())
        res[ft] = {"ms_per_call": e0.elapsed_time(e1) / a.steps, "wall_ms_per_call": wall, "lml": float(lml)}
    best = min(res, key=lambda k: res[k]["ms_per_call"])
    ms = res[best]["ms_per_call"]
    cpu = None
    if not a.no_cpu_baseline:
        cms, el, reps = cpu_c1_ms(T)
        cpu = {"value": T / (cms * 1e-3), "unit": "state-steps/s", "ms_per_call": cms, "cores": 1, "kind": "port",
               "sample": "the whole workload (1 series x %d steps) x %d repeats = %.1f s -- %s" % (T, reps, el, CPU_KIND_NOTE)}
    line = {"metric": "filter+smoother state-steps/sec (fp64)", "value": T / (ms * 1e-3), "unit": "state-steps/s",
            "n_gpus": 1, "steps": a.steps, "warmup": max(3, a.warmup), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "c1: ONE 1-D temporal Matern-3/2 series (state dim 2) x %d steps, Gaussian likelihood, 5%% "
                                   "missing: exact lml + smoothed posterior through SDE_GP.filter_and_smooth; filter_type=%s" % (T, best),
                       "l2": "working set %.1f MB < L2: a latency-bound single series, no flush" % (T * 8 * 16 / 1e6),
                       "parallelism": "single GPU"},
            "by_filter_type": res,
            "roofline": {"bound": "hbm", "achieved": T * 8 * (3 + 2 * 6 + 2) / (ms * 1e-3) / 1e9, "peak": measured_peak_gbs()[0],
                         "unit": "GB/s", "frac": T * 8 * (3 + 2 * 6 + 2) / (ms * 1e-3) / 1e9 / measured_peak_gbs()[0],
                         "traffic": None, "note": "one series x 10k steps is launch / latency bound by construction"},
            "cpu_baseline": cpu, "e2e": {"value": T / (res[best]["wall_ms_per_call"] * 1e-3), "unit": "state-steps/s",
                                         "h2d_bytes_per_step": T * 8 * 2, "d2h_bytes_per_step": 0,
                                         "note": "host wall clock per call, numpy inputs uploaded by the host API every call; results left on the device"},
            "clocks": None, "gpu_launches": None}
    print(json.dumps(line), flush=True)


def main():
    a = parse()
    # stdout carries the ONE JSON line and nothing else: native libraries that write to file descriptor 1 (NCCL's
    # version banner at N > 1) are sent to stderr, Python's own stdout keeps the original descriptor
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(keep, "w")
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "c1":
        run_c1(a)
    elif a.workload == "c3cvi":
        run_c3cvi(a)
    elif a.workload == "c2":
        run_c2(a)
    elif a.workload == "c2cvi":
        run_c2cvi(a)
    elif a.workload == "c3":
        run_c3(a)
    elif a.workload == "cvi":
        run_cvi(a)
    elif a.workload == "grad":
        run_grad(a)
    elif a.workload == "spatial":
        run_spatial(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
