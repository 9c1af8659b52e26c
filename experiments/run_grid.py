#!/usr/bin/env python
"""Experiment grid of the reference (experiments/cu2rec.sh:8-16, cu2rec_prof.sh:17) for this build:
data set x iterations x factors through the drop-in CLI `bin/mf -c exp.cfg train.csv test.csv`,
`ncu` in place of `nvprof`, and a roofline table in place of eyeballing the log.

    python experiments/run_grid.py                         # the reference's grid on synthetic data
    python experiments/run_grid.py --datasets ml-100k --iterations 100 500 --factors 50
    python experiments/run_grid.py --prof --datasets ml-100k --iterations 100 --factors 50

Data: `<data-dir>/<dataset>/ratings_mapped_{train,test}.csv` (the reference's layout). Real files
are used when present; otherwise a synthetic low-rank-plus-noise set of the same shape is generated
once (cu2b_synth_ratings + cu2b_write_ratings_csv). Every run appends the CLI's own stdout plus
wall time to results/<date>-<commit>.txt like the reference's script does, and one row to
results/<date>-<commit>.jsonl / .md: updates/s from the CLI's "Time taken for N of iterations"
line (device time of the training loop), epochs (run_surprise.py:20-23: U * iterations / R), the
algorithmic byte rate (16k+12 B per update) against the HBM roofline, end-to-end wall time (CSV
parse + H2D + training + 5 output CSVs), final TEST RMSE.
--prof wraps each run in `ncu --metrics gpu__time_duration.sum` and stores the launch list under
results/prof/<dataset>-<iterations>-<factors>.csv (a number printed under ncu is not a bench value;
the table then only carries the per-kernel time shares)."""
import argparse
import collections
import csv
import datetime
import json
import math
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = {  # users, items, ratings, integer ratings
    "ml-100k": (943, 1682, 100000, False),
    "ml-20m": (138493, 26744, 20000263, False),
    "netflix": (480189, 17770, 100480507, True),
}


def ensure_dataset(data_dir, name):
    d = os.path.join(data_dir, name)
    tr, te = os.path.join(d, "ratings_mapped_train.csv"), os.path.join(d, "ratings_mapped_test.csv")
    if os.path.exists(tr) and os.path.exists(te):
        return tr, te, "files"
    import cu2rec_b200 as cu
    U, I, R, integer = SHAPES[name]
    os.makedirs(d, exist_ok=True)
    t0 = time.time()
    a, b = cu.synth_ratings(U, I, R, integer_ratings=integer, seed=20240607)
    cu.write_ratings_csv(tr, a)
    cu.write_ratings_csv(te, b)
    print("[grid] generated synthetic %s (%d + %d ratings) in %.1fs" % (name, len(a), len(b), time.time() - t0), file=sys.stderr)
    return tr, te, "synthetic"


def dataset_counts(train_csv):
    import cu2rec_b200 as cu
    r, rows, cols, _ = cu.readCSV(train_csv)
    import numpy as np
    return int(len(r)), int(np.count_nonzero(np.bincount(r["user"], minlength=rows))), rows, cols


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback"


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--datasets", nargs="+", default=["ml-100k", "ml-20m", "netflix"], choices=sorted(SHAPES))
    ap.add_argument("--iterations", nargs="+", type=int, default=[100, 500, 1000, 5000, 10000])  # cu2rec.sh:10
    ap.add_argument("--factors", nargs="+", type=int, default=[50, 300])                          # cu2rec.sh:11
    ap.add_argument("--data-dir", default=os.path.join(ROOT, "data", "datasets"))
    ap.add_argument("--results-dir", default=os.path.join(ROOT, "experiments", "results"))
    ap.add_argument("--prof", action="store_true", help="cu2rec_prof.sh: one ncu launch list per run")
    ap.add_argument("--tag", default=None)
    args = ap.parse_args()

    import cu2rec_b200 as cu
    mf = os.path.join(ROOT, "bin", "mf")
    if not os.path.exists(mf):
        raise SystemExit("bin/mf is not built: python -c 'import __graft_entry__ as g; g.build()'")
    os.makedirs(os.path.join(args.results_dir, "prof"), exist_ok=True)
    try:
        commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip() or "nogit"
    except OSError:
        commit = "nogit"
    stamp = args.tag or "%s-%s" % (datetime.datetime.now().strftime("%Y-%m-%d-%H-%M-%S"), commit)
    log_path = os.path.join(args.results_dir, stamp + ".txt")
    peak, peak_src = hbm_peak()
    rows_out = []
    cfg_path = os.path.join(args.results_dir, "exp.cfg")
    for ds in args.datasets:
        train_csv, test_csv, origin = ensure_dataset(args.data_dir, ds)
        n_train, n_active, U, I = dataset_counts(train_csv)
        for it in args.iterations:
            for k in args.factors:
                cu.create_config(cfg_path, num_iterations=it, num_factors=k)  # create_config.py exp.cfg -n it -f k
                cmd = [mf, "-c", cfg_path, train_csv, test_csv]
                prof_csv = None
                if args.prof:
                    prof_csv = os.path.join(args.results_dir, "prof", "%s-%d-%d.csv" % (ds, it, k))
                    cmd = ["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "--csv", "--log-file", prof_csv] + cmd
                t0 = time.perf_counter()
                p = subprocess.run(cmd, capture_output=True, text=True)
                wall = time.perf_counter() - t0
                with open(log_path, "a") as f:
                    f.write(p.stdout + p.stderr + "\nreal\t%.3fs\n" % wall)
                    f.write("Done with %d factors with %d iterations on %s\n" % (k, it, ds))  # cu2rec.sh:17
                print("Done with %d factors with %d iterations on %s" % (k, it, ds))
                if p.returncode != 0:
                    rows_out.append(dict(dataset=ds, iterations=it, factors=k, error=(p.stderr or p.stdout)[-300:]))
                    continue
                m = re.search(r"Time taken for (\d+) of iterations is ([0-9.eE+-]+)", p.stdout)
                secs = float(m.group(2)) if m else float("nan")
                tests = [l for l in p.stdout.splitlines() if l.startswith("TEST:")]
                rmse = float(tests[-1].split()[-1]) if tests else float("nan")
                updates = it * n_active
                row = dict(dataset=ds, data=origin, users=U, items=I, train_ratings=n_train, iterations=it, factors=k,
                           epochs=round(updates / n_train, 3), epochs_ceil=math.ceil(U * it / n_train),  # run_surprise.py:20-23
                           train_seconds=secs, wall_seconds=round(wall, 3), test_rmse=rmse)
                if args.prof and prof_csv and os.path.exists(prof_csv):
                    share = collections.Counter()
                    with open(prof_csv) as f:
                        lines = [l for l in f if not l.startswith("==")]
                    for rec in csv.DictReader(lines):
                        if rec.get("Metric Name") == "gpu__time_duration.sum":
                            name = rec["Kernel Name"].replace("<unnamed>::", "").split("(")[0]
                            share[re.sub(r"<.*", "", name).split("::")[-1].replace("void ", "")] += float(rec["Metric Value"])
                    tot = sum(share.values()) or 1.0
                    row["kernel_time_share"] = {kname: round(v / tot, 4) for kname, v in share.most_common(6)}
                    row["note"] = "timed under ncu: only the shares are meaningful"
                else:
                    ups = updates / secs if secs > 0 else float("nan")
                    gbs = ups * (16 * k + 12) / 1e9
                    row.update(updates_per_s=ups, algorithmic_GBps=round(gbs, 1), hbm_peak_GBps=peak, roofline_frac=round(gbs / peak, 3))
                rows_out.append(row)
                with open(os.path.join(args.results_dir, stamp + ".jsonl"), "a") as f:
                    f.write(json.dumps(row) + "\n")
    # roofline table
    md = ["# cu2rec_b200 experiment grid %s (HBM peak %.0f GB/s, %s)" % (stamp, peak, peak_src), "",
          "| data set | iterations | epochs | k | train s (device loop) | wall s (CLI end to end) | updates/s | algorithmic GB/s (16k+12) | frac of HBM peak | test RMSE |",
          "|---|---|---|---|---|---|---|---|---|---|"]
    for r in rows_out:
        if "error" in r:
            md.append("| %s | %d | | %d | failed: %s |" % (r["dataset"], r["iterations"], r["factors"], r["error"].replace("\n", " ")))
        elif "updates_per_s" in r:
            md.append("| %s (%s) | %d | %.2f | %d | %.4f | %.2f | %.3g | %.0f | %.2f | %.4f |" % (
                r["dataset"], r["data"], r["iterations"], r["epochs"], r["factors"], r["train_seconds"], r["wall_seconds"],
                r["updates_per_s"], r["algorithmic_GBps"], r["roofline_frac"], r["test_rmse"]))
        else:
            md.append("| %s (%s) | %d | %.2f | %d | (ncu) | %.2f | | | | %.4f | %s" % (
                r["dataset"], r["data"], r["iterations"], r["epochs"], r["factors"], r["wall_seconds"], r["test_rmse"],
                json.dumps(r.get("kernel_time_share", {}))))
    with open(os.path.join(args.results_dir, stamp + ".md"), "w") as f:
        f.write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
