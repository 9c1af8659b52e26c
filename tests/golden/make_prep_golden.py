"""Generates tests/golden/prep/*: inputs and the outputs of the REFERENCE's own preprocessing scripts
(/root/reference/preprocessing/*.py, run unmodified with this interpreter) on them. Run in the
build container only (the reference is not available on the GPU box); the fixtures are committed.

    python tests/golden/make_prep_golden.py
"""
import os
import random
import shutil
import subprocess
import sys
import tempfile

REF = "/root/reference/preprocessing"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "prep")


def run(script, *args, cwd=None):
    subprocess.run([sys.executable, os.path.join(REF, script), *map(str, args)], check=True, cwd=cwd)


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = random.Random(2024)
    # raw MovieLens-like file: sparse ids, users not grouped, a 4th column, ratings in 0.5 steps
    users = rng.sample(range(10, 5000), 37)
    items = rng.sample(range(1, 200000), 90)
    seen, rows = set(), []
    while len(rows) < 600:
        u, i = rng.choice(users), rng.choice(items)
        if (u, i) in seen:
            continue
        seen.add((u, i))
        rows.append((u, i, rng.choice([0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0, 4.5, 5.0]), 964982703 + len(rows)))
    raw = os.path.join(OUT, "ratings.csv")
    with open(raw, "w") as f:
        f.write("userId,movieId,rating,timestamp\n")
        for u, i, r, t in rows:
            f.write("%d,%d,%s,%d\n" % (u, i, ("%g" % r) if rng.random() < 0.3 else repr(r), t))
    tmp = tempfile.mkdtemp()
    try:
        work = os.path.join(tmp, "ratings.csv")
        shutil.copy(raw, work)
        run("map_items.py", work)
        shutil.copy(os.path.join(tmp, "ratings_mapped.csv"), os.path.join(OUT, "ratings_mapped.csv"))
        mapped = os.path.join(tmp, "ratings_mapped.csv")
        for ratio, seed in ((0.1, 42), (0.25, 7), (0.5, 0)):
            run("split_to_test_train.py", mapped, ratio, "-s", seed)
            for part in ("train", "test"):
                shutil.copy(os.path.join(tmp, "ratings_mapped_%s.csv" % part),
                            os.path.join(OUT, "ratings_mapped_%s_r%s_s%d.csv" % (part, ratio, seed)))
        # sort_ratings on the split's training part of seed 42 (rows of a user are in shuffled order there)
        run("split_to_test_train.py", mapped, 0.1, "-s", 42)
        run("sort_ratings.py", os.path.join(tmp, "ratings_mapped_train.csv"))
        shutil.copy(os.path.join(tmp, "ratings_mapped_train_sorted.csv"), os.path.join(OUT, "ratings_mapped_train_sorted.csv"))
        # Netflix-style text files: "user item  rating" (two spaces), test has unseen users / items
        nf = os.path.join(tmp, "data", "datasets", "netflix")
        os.makedirs(nf)
        cwd = os.path.join(tmp, "preprocessing")
        os.makedirs(cwd)
        tr_rows = [(rng.choice(users), rng.choice(items[:60]), rng.randint(1, 5)) for _ in range(300)]
        te_rows = [(rng.choice(users + [7, 8]), rng.choice(items), rng.randint(1, 5)) for _ in range(120)]
        for name, rr in (("netflix_train.txt", tr_rows), ("netflix_test.txt", te_rows)):
            with open(os.path.join(nf, name), "w") as f:
                for u, i, r in rr:
                    f.write("%d %d  %d\n" % (u, i, r))
            shutil.copy(os.path.join(nf, name), os.path.join(OUT, name))
        run("map_netflix.py", cwd=cwd)
        for name in ("ratings_mapped_train.csv", "ratings_mapped_test.csv"):
            shutil.copy(os.path.join(nf, name), os.path.join(OUT, "netflix_" + name))
        # create_config
        run("create_config.py", os.path.join(tmp, "a.cfg"))
        run("create_config.py", os.path.join(tmp, "b.cfg"), "-n", 250, "-f", 128, "-l", 0.005, "-s", 7, "-p", 0.05, "-q", 0.03,
            "-u", 0.01, "-i", 0.125)
        shutil.copy(os.path.join(tmp, "a.cfg"), os.path.join(OUT, "config_defaults.cfg"))
        shutil.copy(os.path.join(tmp, "b.cfg"), os.path.join(OUT, "config_custom.cfg"))
        # convert_to_np.py on factor files as writeCSV produces them and on the shapes genfromtxt squeezes
        npy_inputs = {
            "np_matrix.csv": "".join(",".join("%f" % rng.gauss(0, 0.3) for _ in range(5)) + "\n" for _ in range(24)),
            "np_column.csv": "".join("%f\n" % rng.gauss(0, 1) for _ in range(17)),
            "np_scalar.csv": "3.529860\n",
            "np_row.csv": "1.5,-2.25,1e-3,4\n",
            "np_odd.csv": "# header comment\n1.0, 2.5 ,abc\n\n4,,6.0  # trailing comment\r\n  \n7,8,nan\n-inf,0x10,1e400",
            "np_empty_lines_only.csv": "\n\n",
        }
        for name, text in npy_inputs.items():
            with open(os.path.join(OUT, name), "w", newline="") as f:
                f.write(text)
            work = shutil.copy(os.path.join(OUT, name), os.path.join(tmp, name))
            run("convert_to_np.py", work)
            shutil.copy(os.path.splitext(work)[0] + ".npy", os.path.join(OUT, os.path.splitext(name)[0] + ".npy"))
    finally:
        shutil.rmtree(tmp)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
