#!/usr/bin/env python
"""bench.py -- SGD rating-updates/s of the matrix-factorisation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload netflix|ml20m|ml100k] [--k K] [--iters-per-step T]

Workload (BASELINE.json configs[2], the one the metric is quoted on): Netflix-shape synthetic
low-rank-plus-noise ratings (480 189 users x 17 770 items x ~100.5 M ratings, 90/10 split),
k = 128, reference hyper-parameters (lr 0.01, all regularisers 0.02, check_error 500).
A "step" is one check_error segment of the reference loop: T = 500 reference iterations
(T x U rating updates, one sampled rating per user per iteration, sgd.cu:27-37) followed by
the train+test loss check (training.cu:118-158).

  value  : updates/s over the timed steps with ratings and model resident in HBM
           (device time of the whole enqueued loop: sampler + SGD + loss kernels).
  e2e    : the same metric through the C ABI with HOST (pinned) buffers: session create (H2D of
           both rating matrices and the initial model) + T iterations + download (D2H of
           P, Q, biases) + destroy, wall clock.
  roofline: the SGD kernel alone, algorithmic bytes (16k+12 per update) / its CUDA-event time,
           against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline: the UNMODIFIED reference mf_cpu (oracle/_ref/mf_cpu, single-threaded) on a
           bounded user-prefix sample of the same workload, timed by its own clock() line.

--impl reference times that CPU reference alone and prints the same JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# torchrun pins OMP_NUM_THREADS=1; the host-side workload generator / CSV code is OpenMP code.
# Give every rank its share of the host cores (must happen before libgomp initialises).
if int(os.environ.get("WORLD_SIZE", "1")) > 1:
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // int(os.environ.get("LOCAL_WORLD_SIZE", os.environ["WORLD_SIZE"]))))

WORKLOADS = {
    # name: (users, items, ratings, integer_ratings, default k)
    "netflix": (480189, 17770, 100480507, True, 128),
    "ml20m": (138493, 26744, 20000263, False, 64),
    "ml100k": (943, 1682, 100000, False, 32),
    # one DSGD item block of the netflix shape at 8 GPUs (all users, 1/8 of the items and ratings):
    # single-GPU stand-in for the item-popularity concentration a rank sees inside a sub-epoch
    "nfblock8": (480189, 2221, 12560000, True, 128),
    # the same for 4 and 2 GPUs (the regime in which round 1's 4-GPU run diverged)
    "nfblock4": (480189, 4442, 25120000, True, 128),
    "nfblock2": (480189, 8885, 50240000, True, 128),
    # one (user block, item block) cell at 8 / 4 GPUs: with CU2B_DSGD_ROUND = 64 / world this is the sub-epoch kernel
    # of a real rank, launch for launch (users per rank, draws per run, item popularity of the block)
    "nfcell8": (60024, 2221, 1570000, True, 128),
    "nfcell4": (120047, 4442, 6280000, True, 128),
}
DATA_SEED = 20240607
METRIC = "sgd_rating_updates_per_sec"
UNIT = "updates/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons DURING the timed region. NVML is polled from a thread every few
    milliseconds (the timed region of a multi-GPU run lasts tens of milliseconds: `nvidia-smi -lms` would return
    nothing); nvidia-smi is the fallback when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device = device
        self.proc = None
        self.lines = []
        self.samples = []
        self.nvml = None
        self._stop = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x for x in vis.split(",") if x.strip()]
            if self.device < len(ids) and ids[self.device].strip().isdigit():
                return int(ids[self.device])
        return self.device

    def _poll(self):
        nv = self.nvml
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1e3
                rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, mx, pw, rs))
            except Exception:
                pass
            time.sleep(0.004)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.t.join(timeout=1)
            nv = self.nvml
            names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            reasons = sorted(n for n, bit in names.items() if any(s[3] & bit for s in self.samples))
            sm = [s[0] for s in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                    "sm_max_mhz": max(s[1] for s in self.samples) if self.samples else None,
                    "power_w_max": max(s[2] for s in self.samples) if self.samples else None, "samples": len(sm),
                    "reasons": reasons, "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvidia-smi"}


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------
_WORKLOAD_CACHE = {}


def make_workload(name, users_prefix=None):
    import cu2rec_b200 as cu
    U, I, R, integer, _ = WORKLOADS[name]
    if name not in _WORKLOAD_CACHE:
        t0 = time.time()
        _WORKLOAD_CACHE[name] = cu.synth_ratings(U, I, R, integer_ratings=integer, seed=DATA_SEED)
        log("[bench] generated %s: %d train + %d test ratings in %.1fs" % (
            name, len(_WORKLOAD_CACHE[name][0]), len(_WORKLOAD_CACHE[name][1]), time.time() - t0))
    tr, te = _WORKLOAD_CACHE[name]
    if users_prefix is not None and users_prefix < U:
        tr = tr[: int(np.searchsorted(tr["user"], users_prefix))]
        te = te[: int(np.searchsorted(te["user"], users_prefix))]
        U = users_prefix
    return tr, te, U, I


class _Keep(np.ndarray):
    pass


def pin(a):
    try:
        import torch
        t = torch.empty(max(1, a.nbytes), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        out = t.numpy()[: a.nbytes].view(a.dtype).reshape(a.shape).view(_Keep)
        out._owner = t
        out[...] = a
        return out
    except Exception as e:  # pragma: no cover
        log("[bench] pinned allocation failed (%s); using pageable memory" % e)
        return np.ascontiguousarray(a)


# ---------------------------------------------------------------------------------------------
# the reference CPU arm (oracle/_ref/mf_cpu, else the oracle port)
# ---------------------------------------------------------------------------------------------
def write_csv(path, r):
    with open(path, "w") as f:
        f.write("userId,itemId,rating\n")
        u, i, x = r["user"] + 1, r["item"] + 1, r["rating"]
        step = 1 << 18
        for s in range(0, len(r), step):
            f.write("".join("%d,%d,%.1f\n" % t for t in zip(u[s:s + step].tolist(), i[s:s + step].tolist(), x[s:s + step].tolist())))


class CpuReference:
    """Times the reference's own CPU implementation on a bounded sample of the workload."""

    def __init__(self, workload, k, sample_users=10000, budget_s=12.0):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle as O
        self.O = O
        self.k = k
        self.workload = workload
        self.sample_users = min(sample_users, WORKLOADS[workload][0])
        self.budget_s = budget_s
        self.binary = O.ref_binary("mf_cpu")
        self.kind = "reference" if self.binary else "port"
        self.tmp = tempfile.mkdtemp(prefix="cu2b_ref_")
        # mf_sequential.cu:111 draws from the INCLUSIVE range [low, high]: for the last user the
        # reference reads one element past its rating arrays with probability 1/(degree+1) per
        # iteration and can segfault on what it finds there. End the sample on the heaviest user
        # near the requested cut to make that rare, and retry on a crash (run_once).
        tr_all, _, U_all, _ = make_workload(workload)
        if self.sample_users < U_all:
            lo = max(1, int(self.sample_users * 0.9))
            deg = np.bincount(tr_all["user"][: int(np.searchsorted(tr_all["user"], self.sample_users))],
                              minlength=self.sample_users)
            self.sample_users = lo + int(np.argmax(deg[lo:self.sample_users])) + 1
        self.tr, self.te, self.U, self.I = make_workload(workload, users_prefix=self.sample_users)
        if self.binary:
            write_csv(os.path.join(self.tmp, "train.csv"), self.tr)
            write_csv(os.path.join(self.tmp, "test.csv"), self.te)
        # ~21 us per update for the reference (std::random_device per update, mf_sequential.cu:109)
        per_update = 21e-6 if self.binary else 0.4e-6 * max(1.0, k / 32)
        self.iters = int(max(2, min(2000, budget_s / (per_update * self.U))))

    def sample_desc(self):
        return ("first %d users of the %s-shape workload (%d train ratings, all %d items), k=%d, %d iterations = "
                "%d updates per step, incl. the reference's loss checks at iterations 1 and %d" %
                (self.U, self.workload, len(self.tr), self.I, self.k, self.iters, self.iters * self.U, self.iters))

    def run_once(self):
        """-> (updates, seconds, final test rmse)"""
        updates = self.iters * self.U
        if self.binary:
            cfg = os.path.join(self.tmp, "c.cfg")
            open(cfg, "w").write("0 %d %d 0.01 42 0.02 0.02 0.02 0.02" % (self.iters, self.k))
            for attempt in range(6):
                p = subprocess.run([self.binary, "-c", cfg, os.path.join(self.tmp, "train.csv"), os.path.join(self.tmp, "test.csv")],
                                   capture_output=True, text=True)
                if p.returncode == 0:
                    break
                log("[bench] reference mf_cpu died with rc=%d (its out-of-range sample, mf_sequential.cu:111); retrying" % p.returncode)
            else:
                raise RuntimeError("reference mf_cpu crashed 6 times in a row")
            out = p.stdout
            secs = float(re.search(r"Time taken for \d+ of iterations is ([0-9.eE+-]+)", out).group(1))
            rmse = float([l for l in out.splitlines() if l.startswith("TEST:")][-1].split()[-1])
            return updates, secs, rmse
        return self._run_port(self.iters)

    def _run_port(self, iters):
        """The oracle restatement (oracle/mf_oracle.cpp: seeded counter-based sampler, half-open range) on the sample."""
        O, k = self.O, self.k
        import cu2rec_b200 as cu
        mtr, mte = cu.createSparseMatrix(self.tr, self.U, self.I), cu.createSparseMatrix(self.te, self.U, self.I)
        mu = np.float32(self.tr["rating"].astype(np.float64).mean())
        init = lambda n: O.init_normal(n, k)
        P, Q, ub, ib = init(self.U * k), init(self.I * k), init(self.U), init(self.I)
        t0 = time.perf_counter()
        *_, lg = O.train((mtr.indptr, mtr.indices, mtr.data), (mte.indptr, mte.indices, mte.data), P, Q, ub, ib, mu,
                         O.hyper(k), 42, iters, use_decay=False)
        return iters * self.U, time.perf_counter() - t0, lg[-1]["test_rmse"]

    def port_fixed_rng(self, budget_s=3.0):
        """~21 us of every mf_cpu update is the std::random_device + mt19937 construction
        (mf_sequential.cu:109), so the restatement with a seeded RNG is reported beside it -- labelled, not a substitute."""
        iters = int(max(2, min(2000, budget_s / (0.4e-6 * max(1.0, self.k / 32) * self.U))))
        u, s, rmse = self._run_port(iters)
        return {"value": u / s, "unit": UNIT, "cores": 1, "kind": "port", "seconds": s, "final_test_rmse": rmse,
                "what": "CPU oracle (fixed RNG): the restatement oracle/mf_oracle.cpp, same sample, %d iterations, updates on one "
                        "thread (its two loss checks use the host threads; they are < 2 %% of the time); "
                        "not the reference and not the headline baseline" % iters}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # under torchrun only rank 0 measures the CPU reference
    k = args.k or WORKLOADS[args.workload][4]
    total = max(1, args.steps + args.warmup)
    ref = CpuReference(args.workload, k, budget_s=min(12.0, args.cpu_budget, 150.0 / total))
    for _ in range(args.warmup):
        ref.run_once()
    ups, secs, rmse = 0, 0.0, None
    for _ in range(args.steps):
        u, s, rmse = ref.run_once()
        ups += u
        secs += s
    value = ups / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the arm's step is a bounded sample of the workload (the whole step would take hours on one core)
        "config": dict(config_dict(args, k), iters_per_step=ref.iters, updates_per_step=ref.iters * ref.U,
                       users_in_sample=ref.U, train_ratings_in_sample=int(len(ref.tr)),
                       step="%d reference iterations on the first %d users (bounded sample; per-update cost does not depend "
                            "on the sample size) incl. the reference's loss checks" % (ref.iters, ref.U)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": ref.kind, "sample": ref.sample_desc(),
                         "host_cores_available": os.cpu_count(), "final_test_rmse": rmse,
                         "port_fixed_rng": ref.port_fixed_rng() if ref.kind == "reference" else None},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def config_dict(args, k):
    U, I, R, _, _ = WORKLOADS[args.workload]
    return {"workload": "%s-shape synthetic low-rank+noise (%d users x %d items x ~%d ratings, 90/10 split), k=%d, "
                        "hogwild, per_user sampler (one sampled rating per user per iteration) fused into the "
                        "update kernel, iteration-tiled schedule (32 iterations per user round)" % (args.workload, U, I, R, k),
            "n_factors": k, "iters_per_step": args.iters_per_step, "updates_per_step": args.iters_per_step * U,
            "step": "T reference iterations + one train/test loss check (training.cu:118)",
            "l2": "inputs larger than L2 (P %d MB + rating/update streams >> 126 MB)" % (U * k * 4 >> 20),
            "lr": 0.01, "reg": 0.02, "check_error": args.iters_per_step,
            # experiment switches that change what runs (none set = the defaults DESIGN.md describes)
            **{"env": {n: os.environ[n] for n in ("CU2B_DSGD_THIN", "CU2B_DSGD_THIN_BIAS", "CU2B_DSGD_ROUND", "CU2B_INFLIGHT_LR", "CU2B_ROUND",
                                                  "CU2B_TILE_PIPE", "CU2B_PLACEMENT", "CU2B_IB_STRIDE", "CU2B_DSGD_FUSED", "CU2B_DSGD_GRID") if n in os.environ}}}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def l2_row_peak():
    """Measured peak of the update kernels' binding resource: a warp reads a 512-byte row from L2 and adds a 512-byte
    step to it with red.global.add.v4.f32, uniformly random rows (no popularity skew, no bias side traffic).
    Live from tools/micro/l2_rows (a few seconds on this box) when the binary is there, else the committed
    profile of the same tool. -> (G rows/s, L2 TB/s, source)"""
    exe = os.path.join(ROOT, "tools", "micro", "l2_rows")
    rows = []
    src = None
    if os.path.exists(exe):
        try:
            out = subprocess.run([exe, "peak"], capture_output=True, text=True, timeout=120).stdout
            rows = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
            src = "measured in this run (tools/micro/l2_rows peak)"
        except Exception as exc:  # pragma: no cover
            log("[bench] l2_rows peak failed: %r" % (exc,))
    if not rows:
        prof = os.path.join(ROOT, "profiles", "r2_l2_rows_micro.jsonl")
        if os.path.exists(prof):
            rows = [json.loads(l) for l in open(prof) if l.startswith("{")]
            src = "profiles/r2_l2_rows_micro.jsonl (measured on this pool's B200, round 2)"
    best = [r for r in rows if r["mode"] == "gather_red" and r["dist"] == "uniform" and r["rows"] == 17770]
    if not best:
        return None, None, "unavailable"
    top = max(best, key=lambda r: r["G_rows_per_s"])
    return top["G_rows_per_s"], top["l2_TB_per_s"], src


def traffic_from_profiles(kernel, k):
    """DRAM bytes per update of the SGD kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "sgd_traffic.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))[kernel][str(k)]["bytes_per_update"])
        except Exception:
            return None
    return None


def run_ours(args):
    import cu2rec_b200 as cu
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        return run_ours_dsgd(args, rank, world)
    if args.gpus != 1:
        raise SystemExit("--gpus %d needs a torchrun launch (python -m torch.distributed.run --nproc-per-node %d ...)"
                         % (args.gpus, args.gpus))
    k = args.k or WORKLOADS[args.workload][4]
    T = args.iters_per_step
    info = cu.device_info(0)
    log("[bench] device: %s, %d SMs" % (info["name"], info["sm_count"]))
    tr, te, U, I = make_workload(args.workload)
    mtr, mte = cu.createSparseMatrix(tr, U, I), cu.createSparseMatrix(te, U, I)
    mu = np.float32(tr["rating"].astype(np.float64).sum() / len(tr))
    init = lambda n: cu.initialize_normal_array(n, k)
    P0, Q0, ub0, ib0 = init(U * k), init(I * k), init(U), init(I)
    total_iters = (args.steps + args.warmup) * T
    cfg = cu.Config(total_iterations=total_iters, n_factors=k, check_error=T)

    # ---- resident-data throughput ------------------------------------------------------------
    sess = cu.Session(mtr, mte, cfg, P0, Q0, ub0, ib0, mu)
    for _ in range(args.warmup):
        sess.run(T)
    sess.stats(reset=True)
    clocks = ClockSampler()
    clocks.start()
    t0 = time.perf_counter()
    step_ms, prev = [], 0.0
    for _ in range(args.steps):
        sess.run(T)  # each call ends with a stream synchronize; raises on a non-finite loss check (CU2B_ERR_DIVERGED)
        now = sess.stats()["total_ms"]
        step_ms.append(now - prev)
        prev = now
    wall = time.perf_counter() - t0
    clk = clocks.stop()
    st = sess.stats()
    lg = sess.log()
    sess.close()
    if not all(np.isfinite(r["test_rmse"]) and np.isfinite(r["train_rmse"]) for r in lg):
        raise SystemExit("[bench] non-finite RMSE in the training log: %r" % ([r["test_rmse"] for r in lg],))
    updates = st["updates"]
    dev_s = st["total_ms"] / 1e3
    value = updates / dev_s
    spread = {"steps_ms": [round(x, 3) for x in step_ms], "min_ms": min(step_ms), "median_ms": float(np.median(step_ms)),
              "max_ms": max(step_ms), "value_best_step": T * U / (min(step_ms) / 1e3), "value_worst_step": T * U / (max(step_ms) / 1e3)}
    bytes_per_update = 16 * k + 12
    sgd_gbs = updates * bytes_per_update / (st["sgd_ms"] / 1e3) / 1e9
    peak, peak_src = peaks()
    launches = max(1, int(st["sgd_launches"]))
    kernel_name = ("mf_sgd_hogwild" if cfg.round_iters <= 1 else
                   "mf_sgd_user_tiles" if os.environ.get("CU2B_TILE_PIPE") == "tma" else "mf_sgd_user_rounds")
    bpu = traffic_from_profiles(kernel_name, k)
    kernel_ups = updates / (st["sgd_ms"] / 1e3)
    # The kernel keeps a user's P row in registers for a whole round, so HBM sees ~80 B/update; what binds is the
    # L2: every update reads an item row (4 k bytes) and adds a step to it with 128-bit atomics (4 k bytes), plus
    # 4 + 4 bytes of item bias. Bound = that traffic / the measured peak of exactly this access pattern with
    # uniformly random rows (tools/micro/l2_rows; no popularity skew, which only the data decides).
    l2_bytes_per_update = 8 * k + 8
    peak_rows, peak_tbs, peak_l2_src = l2_row_peak()
    l2_achieved = kernel_ups * l2_bytes_per_update / 1e9
    l2_peak = None if peak_tbs is None else peak_tbs * 1e3 * (512.0 / 512.0)
    roofline = {"bound": "l2_atomic", "kernel": kernel_name, "achieved": l2_achieved, "peak": l2_peak, "unit": "GB/s",
                "frac": None if not l2_peak else l2_achieved / l2_peak, "peak_source": peak_l2_src,
                "what": "L2 bytes moved by the item side of the updates (row read + 128-bit atomic add of the row + bias) per "
                        "second vs the measured peak of 'read a 512-byte row, red.global.add.v4.f32 a 512-byte step' on "
                        "uniformly random rows",
                "l2_bytes_per_update": l2_bytes_per_update, "peak_G_rows_per_s": peak_rows,
                # per launch, like `achieved`: ncu dram__bytes_read+write per update x updates per launch
                "traffic": None if bpu is None else bpu * updates / launches,
                "traffic_bytes_per_update": bpu,
                "launches": launches, "kernel_ms_per_launch": st["sgd_ms"] / launches,
                "kernel_ms_per_step": st["sgd_ms"] / args.steps, "kernel_updates_per_s": kernel_ups,
                # the north star's accounting, kept as a labelled secondary: 16k+12 algorithmic bytes per update
                # against the HBM copy peak. It exceeds 1 because P rows live in registers for 32 updates and Q in L2.
                "hbm_algorithmic": {"bytes_per_update": bytes_per_update, "achieved": sgd_gbs, "peak": peak, "unit": "GB/s",
                                    "frac": sgd_gbs / peak, "peak_source": peak_src,
                                    "note": "not a bound for this kernel: physical DRAM traffic is `traffic`"}}

    # ---- the reference's iteration-synchronous order (round_iters = 1), for comparison ----------
    variants = {}
    if not args.no_variants:
        cfg1 = cu.Config(total_iterations=(2 + args.warmup) * T, n_factors=k, check_error=T, round_iters=1)
        with cu.Session(mtr, mte, cfg1, P0, Q0, ub0, ib0, mu) as s1:
            for _ in range(args.warmup):
                s1.run(T)
            s1.stats(reset=True)
            s1.run(T)
            s1.run(T)
            st1, lg1 = s1.stats(), s1.log()
        variants["iteration_synchronous_round1"] = {
            "kernel": "mf_sgd_hogwild (TMA triplet stream, warp per rating, per-user ordering gate)",
            "value": st1["updates"] / (st1["total_ms"] / 1e3), "kernel_updates_per_s": st1["updates"] / (st1["sgd_ms"] / 1e3),
            "algorithmic_gbs": st1["updates"] * bytes_per_update / (st1["sgd_ms"] / 1e3) / 1e9,
            "test_rmse": [round(r["test_rmse"], 5) for r in lg1]}

    # ---- end to end through the C ABI with pinned host buffers --------------------------------
    hp = {n: pin(getattr(mtr, n)) for n in ("indptr", "indices", "data")}
    hq = {n: pin(getattr(mte, n)) for n in ("indptr", "indices", "data")}
    ptr = cu.CSRMatrix(U, I, hp["indptr"], hp["indices"], hp["data"])
    pte = cu.CSRMatrix(U, I, hq["indptr"], hq["indices"], hq["data"])
    hP, hQ, hub, hib = pin(P0), pin(Q0), pin(ub0), pin(ib0)
    cfg_e = cu.Config(total_iterations=T, n_factors=k, check_error=T)
    h2d = sum(a.nbytes for a in list(hp.values()) + list(hq.values()) + [hP, hQ, hub, hib])
    d2h = sum(a.nbytes for a in (hP, hQ, hub, hib))
    e2e_steps = max(3, min(args.steps + 1, 5))

    oP, oQ, oub, oib = pin(P0), pin(Q0), pin(ub0), pin(ib0)  # page-locked result buffers

    def e2e_cold_once():  # everything a first train() call pays: allocation, set-up, teardown
        with cu.Session(ptr, pte, cfg_e, hP, hQ, hub, hib, mu) as s:
            out = s.run_download(T, out=(oP, oQ, oub, oib))  # what cu2b_train does: the D2H overlaps the last loss check
            rm = s.log()[-1]["test_rmse"]
        return out, rm

    def median_ms(fn, reps):
        fn()  # warm-up
        ts, last = [], None
        for _ in range(reps):
            t0 = time.perf_counter()
            last = fn()
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts)), [round(1e3 * t, 2) for t in ts], last

    # e2e: a resident session (buffers allocated once), and per step the H2D of that step's inputs
    # (both rating matrices + the initial model, pinned), T iterations, the D2H of the result
    sess_e = cu.Session(ptr, pte, cfg_e, hP, hQ, hub, hib, mu)

    def e2e_once():
        sess_e.reload(ptr, pte, hP, hQ, hub, hib, mu)
        sess_e.run(T)
        out = sess_e.download(out=(oP, oQ, oub, oib))
        return out, sess_e.log()[-1]["test_rmse"]

    e2e_s, e2e_all, (_, e2e_rmse) = median_ms(e2e_once, e2e_steps)  # >= 3 repeats, median

    # The same steps with the input feed double-buffered, as a training loop prefetches its next batch: two
    # resident sessions; while one runs step i and hands back its result, a second host thread uploads step
    # i + 1's inputs into the other (ctypes releases the GIL inside the library). Every step still pays its
    # full H2D, its 500 iterations and its D2H inside the timed region; the first upload is not overlapped.
    pipelined = None
    try:
        sess_f = cu.Session(ptr, pte, cfg_e, hP, hQ, hub, hib, mu)
        pair = (sess_e, sess_f)

        def upload(sx, box):
            try:
                sx.reload(ptr, pte, hP, hQ, hub, hib, mu)
            except Exception as exc:  # surfaced on the main thread
                box.append(exc)

        def pipelined_once(n_steps):
            errs, rmses = [], []
            t0 = time.perf_counter()
            th = threading.Thread(target=upload, args=(pair[0], errs))
            th.start()
            for i in range(n_steps):
                th.join()
                if errs:
                    raise errs[0]
                if i + 1 < n_steps:
                    th = threading.Thread(target=upload, args=(pair[(i + 1) % 2], errs))
                    th.start()
                cur = pair[i % 2]
                cur.run(T)
                cur.download(out=(oP, oQ, oub, oib))
                rmses.append(cur.log()[-1]["test_rmse"])
            return (time.perf_counter() - t0) / n_steps, rmses

        pipelined_once(2)  # warm-up
        p_s, p_rmse = pipelined_once(e2e_steps)
        if all(abs(r - e2e_rmse) / e2e_rmse < 0.005 for r in p_rmse):
            pipelined = {"value": T * U / p_s, "ms_per_step": 1e3 * p_s, "steps": e2e_steps,
                         "test_rmse": [round(r, 5) for r in p_rmse],
                         "what": "two resident sessions, double-buffered input feed: step i+1's cu2b_session_reload (H2D of "
                                 "both rating matrices + initial model, pinned) is issued from a second host thread while "
                                 "step i runs its %d iterations and downloads (D2H of P, Q, biases); wall clock of %d "
                                 "consecutive steps / %d, the first upload not overlapped" % (T, e2e_steps, e2e_steps)}
        else:
            log("[bench] pipelined e2e: RMSE mismatch %s vs %s; not reported" % (p_rmse, e2e_rmse))
        sess_f.close()
    except Exception as exc:
        log("[bench] pipelined e2e not measured: %r" % (exc,))
    sess_e.close()
    cold_s, cold_all, _ = median_ms(e2e_cold_once, 3)
    seq_what = ("resident session; per step: cu2b_session_reload (H2D of both rating matrices + initial model from "
                "pinned host memory) + %d iterations + download (D2H of P, Q, biases), one step after the other" % T)
    sequential = {"value": T * U / e2e_s, "ms_per_step": 1e3 * e2e_s, "steps": e2e_steps, "ms_all_steps": e2e_all,
                  "what": seq_what}
    # headline = the single call a user of train() / bin/mf makes: create + run + download + destroy, every step.
    # The resident-session feeds of the same per-step work are listed beside it as named variants.
    e2e = {"value": T * U / cold_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": 1e3 * cold_s, "steps": 3, "ms_all_steps": cold_all, "feed": "single_call",
           "what": "cu2b_session_create (allocation, H2D of both rating matrices + initial model from pinned host memory, "
                   "set-up) + cu2b_session_run_download (%d iterations with their loss checks; the D2H of P, Q, biases runs "
                   "while the last check evaluates the final model) + destroy; median of 3" % T,
           "resident_sequential": sequential, "resident_pipelined": pipelined}

    # ---- CPU baseline (rank 0, bounded sample) -----------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        ref = CpuReference(args.workload, k, budget_s=args.cpu_budget)
        u, s, r_rmse = ref.run_once()
        cpu = {"value": u / s, "unit": UNIT, "cores": 1, "kind": ref.kind, "sample": ref.sample_desc(),
               "host_cores_available": os.cpu_count(), "seconds": s, "final_test_rmse": r_rmse,
               "port_fixed_rng": ref.port_fixed_rng() if ref.kind == "reference" else None}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dev_s / args.steps, "wall_ms_per_step": 1e3 * wall / args.steps, "spread": spread,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, k), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(st["kernel_launches"]), "clocks": clk,
        "breakdown_ms_per_step": {"sgd": st["sgd_ms"] / args.steps, "sampler": st["sampler_ms"] / args.steps,
                                  "loss_check": st["loss_ms"] / args.steps},
        "test_rmse": [round(r["test_rmse"], 5) for r in lg], "e2e_test_rmse": e2e_rmse,
        "epochs_per_step": T * U / float(mtr.nonzeros), "variants": variants,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_ours_dsgd(args, rank, world):
    """N > 1: one process per GPU, DSGD 2-D stratification, item blocks rotated through peer memory.
    torch.distributed (NCCL) is only plumbing here: handle exchange, barriers, max-over-ranks."""
    import torch
    import torch.distributed as dist
    import cu2rec_b200 as cu
    local = int(os.environ.get("LOCAL_RANK", rank))
    # NCCL prints its version banner on stdout; the contract is ONE JSON line there
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    k = args.k or WORKLOADS[args.workload][4]
    T = args.iters_per_step
    tr, te, U, I = make_workload(args.workload)
    mu = np.float32(tr["rating"].astype(np.float64).sum() / len(tr))
    part = cu.dsgd_partition(tr, U, I, world)
    init = lambda n: cu.initialize_normal_array(n, k)
    inp = cu.dsgd_rank_inputs(tr, te, U, I, part, rank, init(U * k), init(I * k), init(U), init(I))
    n_train_rows = int(len(tr))
    del tr, te
    _WORKLOAD_CACHE.clear()

    def exchange(blob):
        mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
        allb = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allb, mine)
        return [bytes(b.cpu().numpy().tobytes()) for b in allb]

    def make(total_iters, inputs):
        cfg = cu.Config(total_iterations=total_iters, n_factors=k, check_error=T)
        d = cu.Dsgd(rank, world, inputs, part, cfg, mu, device=local)
        d.connect(exchange(d.handle))
        return d

    def maxreduce(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def sumreduce(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return t.item()

    d = make((args.steps + args.warmup) * T, inp)
    for _ in range(args.warmup):
        d.run(T)
    d.stats(reset=True)
    dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        d.run(T)
    torch.cuda.synchronize()
    dist.barrier()
    wall = time.perf_counter() - t0
    clk = clocks.stop() if rank == 0 else None
    st = d.stats()
    lg = d.log()
    if not all(np.isfinite(r["test_rmse"]) and np.isfinite(r["train_rmse"]) for r in lg):
        raise SystemExit("[bench] rank %d: non-finite RMSE in the DSGD training log: %r" % (rank, [r["test_rmse"] for r in lg]))
    # NCCL all-reduce of the rank-local loss sums must reproduce the peer-memory combine
    sums = torch.tensor(d.local_sums(), dtype=torch.float64, device="cuda")
    dist.all_reduce(sums)
    nccl_rmse = float(np.sqrt(sums[0].item() / n_train_rows))
    d.close()
    dev_s = maxreduce(st["total_ms"]) / 1e3           # slowest rank, device time
    updates = sumreduce(st["updates"])
    sgd_ms_max = maxreduce(st["sgd_ms"])
    sampler_ms_max, loss_ms_max = maxreduce(st["sampler_ms"]), maxreduce(st["loss_ms"])
    wait_ms_max, send_ms_max = maxreduce(st["wait_ms"]), maxreduce(st["send_ms"])
    wait_ms_min = -maxreduce(-st["wait_ms"])
    launches = sumreduce(st["kernel_launches"])
    value = updates / dev_s
    bytes_per_update = 16 * k + 12
    peak, peak_src = peaks()
    my_gbs = st["updates"] * bytes_per_update / (st["sgd_ms"] / 1e3) / 1e9
    my_ups = st["updates"] / (st["sgd_ms"] / 1e3)
    peak_rows, peak_tbs, peak_l2_src = l2_row_peak() if rank == 0 else (None, None, None)
    l2_bytes_per_update = 8 * k + 8
    roofline = {"bound": "l2_atomic", "kernel": "mf_sgd_user_runs (fused wait + sub-epoch + hand-off)", "scope": "rank 0, per GPU",
                "achieved": my_ups * l2_bytes_per_update / 1e9, "peak": None if peak_tbs is None else peak_tbs * 1e3, "unit": "GB/s",
                "frac": None if not peak_tbs else my_ups * l2_bytes_per_update / 1e9 / (peak_tbs * 1e3), "peak_source": peak_l2_src,
                "l2_bytes_per_update": l2_bytes_per_update, "kernel_updates_per_s": my_ups, "traffic": None,
                "kernel_ms_per_step_max_rank": sgd_ms_max / args.steps,
                "note": "the kernel time of a linked sub-epoch includes waiting for the upstream rank's hand-off; a rank sees 1/N "
                        "of the catalogue at a time, so item popularity concentrates its L2 traffic on fewer slices than on one GPU",
                "hbm_algorithmic": {"bytes_per_update": bytes_per_update, "achieved": my_gbs, "peak": peak, "unit": "GB/s",
                                    "frac": my_gbs / peak, "peak_source": peak_src}}

    # end to end: per rank H2D of its strips + model, T iterations, download, destroy
    pinp = cu.api.DsgdRankInputs(
        cu.CSRMatrix(inp.train.rows, inp.train.cols, pin(inp.train.indptr), pin(inp.train.indices), pin(inp.train.data)),
        cu.CSRMatrix(inp.test.rows, inp.test.cols, pin(inp.test.indptr), pin(inp.test.indices), pin(inp.test.data)),
        pin(inp.P), pin(inp.Q), pin(inp.user_bias), pin(inp.item_bias), inp.user_ids, inp.n_train_global,
        inp.n_test_global, inp.n_active_global)
    h2d = sum(a.nbytes for a in (pinp.train.indptr, pinp.train.indices, pinp.train.data, pinp.test.indptr,
                                 pinp.test.indices, pinp.test.data, pinp.P, pinp.Q, pinp.user_bias, pinp.item_bias))
    d2h = sum(a.nbytes for a in (pinp.P, pinp.Q, pinp.user_bias, pinp.item_bias))

    outs = tuple(pin(a) for a in (inp.P, inp.Q, inp.user_bias, inp.item_bias))  # page-locked result buffers

    # e2e: resident contexts (buffers, IPC mappings and flags set up once); per step every rank
    # uploads its strips + the initial model from pinned host memory, the ranks synchronise, run T
    # iterations and download their strips
    de = make(T, pinp)

    def e2e_once():
        de.reload(pinp, mu)
        dist.barrier()
        de.run(T)
        de.download(out=outs)
        return de.log()[-1]["test_rmse"]

    e2e_once()
    e2e_steps = max(3, min(args.steps + 1, 5))
    times = []
    for _ in range(e2e_steps):
        dist.barrier()
        t0 = time.perf_counter()
        e2e_rmse = e2e_once()
        dist.barrier()
        times.append(maxreduce(time.perf_counter() - t0))
    e2e_s = float(np.median(times))
    # the same steps with the input feed double-buffered over two resident contexts per rank: while context A runs step
    # i and downloads, a second host thread reloads step i + 1's inputs into context B (own buffers, own peer mappings);
    # the ranks meet on a host barrier between "every rank has reloaded B" and "anyone runs B"
    pipelined = None
    try:
        de2 = make(T, pinp)
        pair = (de, de2)
        errs = []

        def upload(ctx):
            try:
                ctx.reload(pinp, mu)
            except Exception as exc:  # surfaced on the main thread
                errs.append(exc)

        def pipelined_steps(n_steps):
            rm = []
            dist.barrier()
            t0 = time.perf_counter()
            th = threading.Thread(target=upload, args=(pair[0],))
            th.start()
            for i in range(n_steps):
                th.join()
                if errs:
                    raise errs[0]
                dist.barrier()  # every rank's copy of this step's inputs is in place
                if i + 1 < n_steps:
                    th = threading.Thread(target=upload, args=(pair[(i + 1) % 2],))
                    th.start()
                cur = pair[i % 2]
                cur.run(T)
                cur.download(out=outs)
                rm.append(cur.log()[-1]["test_rmse"])
            dist.barrier()
            return maxreduce(time.perf_counter() - t0) / n_steps, rm

        pipelined_steps(2)
        p_s, p_rmse = pipelined_steps(e2e_steps)
        if all(np.isfinite(r) and abs(r - e2e_rmse) / e2e_rmse < 0.005 for r in p_rmse):
            pipelined = {"value": T * U / p_s, "ms_per_step": 1e3 * p_s, "steps": e2e_steps, "test_rmse": [round(r, 5) for r in p_rmse],
                         "what": "two resident DSGD contexts per rank, double-buffered input feed: step i+1's cu2b_dsgd_reload is "
                                 "issued from a second host thread while step i runs and downloads; wall clock of %d consecutive "
                                 "steps / %d, max over ranks, the first upload not overlapped" % (e2e_steps, e2e_steps)}
        de2.close()
    except Exception as exc:
        log("[bench] rank %d: pipelined DSGD e2e not measured: %r" % (rank, exc))
    de.close()
    h2d_all, d2h_all = sumreduce(h2d), sumreduce(d2h)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / args.steps, "wall_ms_per_step": 1e3 * wall / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config_dict(args, k), parallelism="dsgd%d (2-D stratification, item blocks rotated through "
                           "peer memory over NVLink, loss partials combined in rank order)" % world),
            "roofline": roofline, "cpu_baseline": None,
            "e2e": {"value": T * U / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                    "ms_per_step": 1e3 * e2e_s, "steps": e2e_steps,
                    "ms_all_steps": [round(1e3 * t, 2) for t in times],
                    "what": "resident DSGD contexts; per step and rank: cu2b_dsgd_reload (H2D of the rank's strips + "
                            "initial model, pinned) + barrier + %d iterations + download; max over ranks" % T,
                    "resident_pipelined": pipelined},
            "gpu_launches": int(launches), "clocks": clk,
            "breakdown_ms_per_step_max_rank": {"sgd_subepochs": sgd_ms_max / args.steps, "sampler": sampler_ms_max / args.steps,
                                               "loss_check_incl_gather": loss_ms_max / args.steps,
                                               "handoff_waits_and_sends": (dev_s * 1e3 - sgd_ms_max - sampler_ms_max - loss_ms_max) / args.steps,
                                               "wait_kernels_max_rank": wait_ms_max / args.steps,
                                               "wait_kernels_min_rank": wait_ms_min / args.steps,
                                               "send_kernels_max_rank": send_ms_max / args.steps},
            "test_rmse": [round(r["test_rmse"], 5) for r in lg], "e2e_test_rmse": e2e_rmse,
            "loss_allreduce_check": {"nccl_train_rmse": nccl_rmse, "peer_memory_train_rmse": lg[-1]["train_rmse"]},
            "block_nnz_imbalance": float(part.block_nnz.max() / part.block_nnz.mean()),
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    dist.barrier()
    dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="netflix", choices=sorted(WORKLOADS))  # nfblock8: tools only
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--iters-per-step", type=int, default=500)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("[bench] note: fewer than 3 warm-up steps requested")
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
