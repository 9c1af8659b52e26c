"""Host side of the drop-in boundary (libcu2b host functions; no GPU needed): config file,
ratings CSV ingest, CSR build, matrix reader/writer, initialisation -- against the reference's
golden values, the reference's own outputs (tests/golden) and the CPU oracle."""
import json
import os

import numpy as np
import pytest

import cu2rec_b200 as cu
import oracle as O


@pytest.fixture(scope="module")
def ref_tests(golden_dir):
    return json.load(open(os.path.join(golden_dir, "reference_tests.json")))


def test_read_csv_reference_goldens(golden_dir, fixtures_dir, ref_tests):
    gold = json.load(open(os.path.join(golden_dir, "ref_read_csv.json")))
    for name, g in gold.items():
        r, rows, cols, gb = cu.readCSV(os.path.join(fixtures_dir, name))
        assert (len(r), rows, cols) == (g["n"], g["rows"], g["cols"]), name
        assert int(np.float32(gb).view(np.uint32)) == g["global_bias_bits"], name
        assert r["user"].tolist() == g["users"] and r["item"].tolist() == g["items"]
        assert r["rating"].tolist() == g["ratings"]
    g = ref_tests["test_util.cu:28-31"]  # tests/test_util.cu:28-31
    r, rows, cols, gb = cu.readCSV(os.path.join(fixtures_dir, "test_ratings.csv"))
    assert rows == 6 and cols == 5 and len(r) == 18 and abs(gb - g["global_bias"]) < g["tol"]


@pytest.mark.parametrize("fname,key", [("test_ratings.csv", "test_util.cu:123-125"),
                                       ("test_missing_user_ratings.csv", "test_util.cu:170-172")])
def test_create_sparse_matrix_goldens(fixtures_dir, ref_tests, fname, key):
    r, rows, cols, _ = cu.readCSV(os.path.join(fixtures_dir, fname))
    m = cu.createSparseMatrix(r, rows, cols)
    g = ref_tests[key]
    assert m.indptr.tolist() == g["indptr"]
    assert m.indices.tolist() == g["indices"]
    assert m.data.tolist() == [float(x) for x in g["data"]]


def test_csr_rejects_unsorted_and_out_of_range():
    bad = np.array([(1, 0, 1.0), (0, 0, 1.0)], dtype=cu.RATING_DTYPE)
    with pytest.raises(cu._lib.Cu2bError):
        cu.createSparseMatrix(bad, 2, 1)
    with pytest.raises(cu._lib.Cu2bError):
        cu.createSparseMatrix(np.array([(5, 0, 1.0)], dtype=cu.RATING_DTYPE), 2, 1)
    empty = cu.createSparseMatrix(np.zeros(0, dtype=cu.RATING_DTYPE), 3, 2)
    assert empty.indptr.tolist() == [0, 0, 0, 0]


def _random_csv(path, rng, n, style):
    users = np.sort(rng.randint(1, 400, n))
    items = rng.randint(1, 5000, n)
    with open(path, "w") as f:
        f.write("userId,itemId,rating\n")
        for u, i in zip(users, items):
            if style == "half":
                r = "%.1f" % (rng.randint(1, 11) / 2)
            elif style == "int":
                r = "%d" % rng.randint(1, 6)
            elif style == "long":
                r = repr(float(rng.rand() * 5))
            else:
                r = "%.3e" % (rng.rand() * 5)
            sep = rng.choice([",", ", ", " ,", "\t", ";"]) if style == "messy" else ","
            f.write("%d%s%d%s%s%s" % (u, sep, i, sep, r, "\r\n" if style == "messy" else "\n"))


@pytest.mark.parametrize("style", ["half", "int", "long", "exp", "messy"])
def test_read_csv_matches_oracle_on_random_files(tmp_path, style):
    rng = np.random.RandomState(hash(style) % 1000)
    p = tmp_path / "r.csv"
    _random_csv(p, rng, 3000, style)
    a, rows, cols, gb = cu.readCSV(p)
    b, rows2, cols2, gb2 = O.read_csv(p)
    assert (rows, cols, len(a)) == (rows2, cols2, len(b))
    assert a.tobytes() == b.tobytes()
    assert np.float32(gb).view(np.uint32) == np.float32(gb2).view(np.uint32)


def test_read_csv_multithreaded_path_matches_oracle(tmp_path):
    rng = np.random.RandomState(5)
    p = tmp_path / "big.csv"
    n = 200000  # > 1 MiB => the parallel tokenizer is used
    users = np.sort(rng.randint(1, 20000, n))
    items = rng.randint(1, 5000, n)
    r = rng.randint(1, 11, n) / 2
    with open(p, "w") as f:
        f.write("userId,itemId,rating\n")
        f.write("".join("%d,%d,%.1f\n" % t for t in zip(users, items, r)))
    assert os.path.getsize(p) > (1 << 20)
    a, rows, cols, gb = cu.readCSV(p)
    b, rows2, cols2, gb2 = O.read_csv(p)
    assert len(a) == n and a.tobytes() == b.tobytes() and (rows, cols) == (rows2, cols2)
    assert np.float32(gb).view(np.uint32) == np.float32(gb2).view(np.uint32)


def _big_csv(path, rng, n, fmt, extra=lambda k: ""):
    users = np.sort(rng.randint(1, 20000, n))
    items = rng.randint(1, 5000, n)
    with open(path, "w") as f:
        f.write("userId,itemId,rating\n")
        f.write("".join(fmt(u, i) + extra(k) for k, (u, i) in enumerate(zip(users, items))))
    assert os.path.getsize(path) > (1 << 20)


@pytest.mark.parametrize("case", ["blank_lines", "two_per_line", "inexact_sum", "stops_midway", "crlf"])
def test_read_csv_parallel_slices_compaction_and_fallbacks(tmp_path, case):
    """The threaded reader parses into per-chunk slices of one array (lines bound the records):
    gaps are closed when lines hold no record, several records per line fall back to the serial
    reader, the mean is summed as exact integers only when that cannot differ from util.cu:34's
    sequential double sum, and a malformed record ends the stream like a failed operator>>."""
    rng = np.random.RandomState(17)
    p = tmp_path / (case + ".csv")
    n = 150000
    if case == "blank_lines":
        _big_csv(p, rng, n, lambda u, i: "%d,%d,%.1f\n" % (u, i, rng.randint(1, 11) / 2), lambda k: "\n" if k % 7 == 0 else "")
    elif case == "two_per_line":
        _big_csv(p, rng, n, lambda u, i: "%d,%d,%d" % (u, i, rng.randint(1, 6)), lambda k: " " if k % 2 == 0 else "\n")
    elif case == "inexact_sum":
        _big_csv(p, rng, n, lambda u, i: "%d,%d,%s\n" % (u, i, repr(float(rng.rand() * 5 + 1e-9))))
    elif case == "stops_midway":
        _big_csv(p, rng, n, lambda u, i: "%d,%d,%.1f\n" % (u, i, rng.randint(1, 11) / 2), lambda k: "oops\n" if k == 100000 else "")
    else:
        _big_csv(p, rng, n, lambda u, i: "%d,%d,%.1f\r\n" % (u, i, rng.randint(1, 11) / 2))
    a, rows, cols, gb = cu.readCSV(p)
    b, rows2, cols2, gb2 = O.read_csv(p)
    assert len(a) == len(b) == (100001 if case == "stops_midway" else n)
    assert a.tobytes() == b.tobytes() and (rows, cols) == (rows2, cols2)
    assert np.float32(gb).view(np.uint32) == np.float32(gb2).view(np.uint32)


@pytest.fixture
def tiny_chunks(monkeypatch):
    """Forces the chunked readers onto every file, however small (8 chunks of a 200-byte file)."""
    monkeypatch.setenv("CU2B_IO_PARALLEL_MIN_BYTES", "0")


def test_read_csv_chunked_reader_fuzz_against_oracle(tmp_path, tiny_chunks):
    """Chunk cuts land anywhere in small messy files: records straddling a cut, blank lines, several
    records on a line, garbage that ends the stream, no final newline, chunks without a record."""
    rng = np.random.RandomState(123)
    seps = [",", ", ", " ,", " , ", "\t", ";", "|"]
    for case in range(250):
        n = rng.randint(0, 60)
        parts, u = ["userId,itemId,rating" + ("\n" if rng.rand() < 0.95 else "")], 1
        for r in range(n):
            u += rng.randint(0, 3)
            sep = seps[rng.randint(len(seps))]
            val = rng.choice(["%d" % rng.randint(1, 6), "%.1f" % (rng.randint(1, 11) / 2), "%.3f" % (rng.rand() * 5),
                              repr(float(rng.rand() * 5)), "%.2e" % (rng.rand() * 5)])
            rec = "%s%d%s%d%s%s" % (" " * rng.randint(0, 3), u, sep, rng.randint(1, 5000), sep, val)
            x = rng.rand()
            end = "\n" if x < 0.7 else "\r\n" if x < 0.8 else " " if x < 0.9 else "\n\n  \n"
            if rng.rand() < 0.01:
                rec = "oops" + rec
            parts.append(rec + end)
        text = "".join(parts)
        if rng.rand() < 0.3:
            text = text.rstrip("\n ")
        p = tmp_path / ("f%d.csv" % case)
        p.write_text(text, newline="")
        a, rows, cols, gb = cu.readCSV(p)
        b, rows2, cols2, gb2 = O.read_csv(p)
        assert a.tobytes() == b.tobytes() and (rows, cols) == (rows2, cols2), (case, text)
        if len(a):
            assert np.float32(gb).view(np.uint32) == np.float32(gb2).view(np.uint32), (case, text)


def test_read_array_and_convert_to_np_chunked_readers_fuzz(tmp_path, tiny_chunks):
    rng = np.random.RandomState(77)
    for case in range(150):
        rows, cols = rng.randint(1, 30), rng.randint(1, 7)
        lines = []
        for r in range(rows):
            toks = []
            for c in range(cols):
                v = rng.standard_normal() * 10.0 ** rng.randint(-4, 5)
                t = rng.choice(["%f" % v, "%d" % int(v), "%.3e" % v, repr(float(np.float32(v)))])
                toks.append(" " * rng.randint(0, 2) + t + " " * rng.randint(0, 2))
            lines.append(",".join(toks))
        text = "\n".join(lines) + ("\n" if rng.rand() < 0.7 else "")
        p = tmp_path / ("m%d.csv" % case)
        p.write_text(text)
        arr, r, c = cu.read_array(p)
        want = np.array([np.float32(t) for ln in lines for t in ln.split(",")], np.float32)
        assert (r, c) == (rows, rows * cols) and arr.tobytes() == want.tobytes(), (case, text)
        info = cu.convert_to_np(str(p))
        got = np.load(info["out"])
        assert np.array_equal(got, np.squeeze(np.array([[float(t) for t in ln.split(",")] for ln in lines]))), (case, text)
        assert (tmp_path / ("m%d.npy" % case)).read_bytes()[:6] == b"\x93NUMPY"


def test_build_csr_parallel_path_matches_numpy():
    rng = np.random.RandomState(3)
    n, U, I = 300000, 50000, 700  # > 2^16 ratings => the threaded fill; ~0.2 % of the users are missing
    r = np.zeros(n, dtype=cu.RATING_DTYPE)
    r["user"], r["item"], r["rating"] = np.sort(rng.randint(0, U - 5, n)), rng.randint(0, I, n), rng.randint(1, 6, n)
    m = cu.createSparseMatrix(r, U, I)
    assert m.indptr.tolist() == np.searchsorted(r["user"], np.arange(U + 1), side="left").tolist()
    assert np.array_equal(m.indices, r["item"]) and np.array_equal(m.data, r["rating"])
    bad = r.copy()
    bad["user"][200000] = bad["user"][199999] - 1  # first violation decides the message
    bad["user"][250000] = U + 3
    with pytest.raises(cu._lib.Cu2bError, match="rating 200000: ratings are not grouped"):
        cu.createSparseMatrix(bad, U, I)
    bad["user"][100] = -1
    with pytest.raises(cu._lib.Cu2bError, match="rating 100: user id -1 outside"):
        cu.createSparseMatrix(bad, U, I)


def test_read_csv_edge_cases(tmp_path):
    # a fourth column ends the parse after the first row (util.cu:30)
    p = tmp_path / "four.csv"
    p.write_text("u,i,r,ts\n1,2,3.0,978300760\n1,3,4.0,978300761\n")
    a, rows, cols, gb = cu.readCSV(p)
    b, *_ = O.read_csv(p)
    assert len(a) == len(b) == 1 and a.tobytes() == b.tobytes()
    # header only / empty file
    p2 = tmp_path / "hdr.csv"
    p2.write_text("userId,itemId,rating\n")
    a, rows, cols, gb = cu.readCSV(p2)
    assert len(a) == 0 and rows == 0 and cols == 0
    (tmp_path / "empty.csv").write_text("")
    a, rows, cols, gb = cu.readCSV(tmp_path / "empty.csv")
    assert len(a) == 0
    with pytest.raises(cu._lib.Cu2bError):
        cu.readCSV(tmp_path / "does_not_exist.csv")
    # no trailing newline, like the reference fixtures
    p3 = tmp_path / "nonl.csv"
    p3.write_text("h\n1,1,1.0\n2,2,5.0")
    a, rows, cols, gb = cu.readCSV(p3)
    assert len(a) == 2 and rows == 2 and a["rating"].tolist() == [1.0, 5.0]


def test_config_goldens_and_round_trip(tmp_path, golden_dir, fixtures_dir, ref_tests):
    cfg = cu.Config()
    assert (cfg.total_iterations, cfg.n_factors, cfg.check_error, cfg.n_threads) == (5000, 50, 500, 32)  # config.h:25-45
    assert cfg.learning_rate == np.float32(0.01) and cfg.patience == 2.0 and cfg.learning_rate_decay == np.float32(0.2)
    cfg.read_config(os.path.join(fixtures_dir, "test_config.cfg"))
    g = ref_tests["test_config.cu:14-15"]
    assert cfg.total_iterations == g["total_iterations"] and abs(cfg.P_reg - g["P_reg"]) < g["tol"]
    ref = json.load(open(os.path.join(golden_dir, "ref_read_config.json")))["fields"]
    got = [cfg.cur_iterations, cfg.total_iterations, cfg.n_factors, cfg.learning_rate, cfg.seed, cfg.P_reg, cfg.Q_reg,
           cfg.user_bias_reg, cfg.item_bias_reg]
    for a, b in zip(got, ref):
        assert np.float32(a) == np.float32(float(b))
    # oracle agrees
    n, v = O.read_config(os.path.join(fixtures_dir, "test_config.cfg"))
    assert [np.float32(x) for x in v] == [np.float32(x) for x in got]
    # tests/test_config.cu:19-26 save -> load
    c2 = cu.Config(total_iterations=100, P_reg=0.2)
    c2.write_config(tmp_path / "gen.cfg")
    c3 = cu.Config()
    c3.read_config(tmp_path / "gen.cfg")
    assert c3.total_iterations == 100 and abs(c3.P_reg - 0.2) < 1e-4
    assert len((tmp_path / "gen.cfg").read_text().split()) == 9
    # short file keeps defaults for the missing fields; optional extension fields parse
    (tmp_path / "short.cfg").write_text("0 77 12")
    c4 = cu.Config()
    c4.read_config(tmp_path / "short.cfg")
    assert (c4.total_iterations, c4.n_factors, c4.learning_rate) == (77, 12, np.float32(0.01))
    (tmp_path / "ext.cfg").write_text("0 10 8 0.05 7 0.1 0.2 0.3 0.4 64 3 0.5 25 1 1 16 2 4")
    c5 = cu.Config()
    c5.read_config(tmp_path / "ext.cfg")
    assert (c5.n_threads, c5.patience, c5.check_error, c5.mode, c5.sampler, c5.n_blocks, c5.n_gpus) == (64, 3.0, 25, 1, 1, 16, 2)
    assert c5.round_iters == 4 and cu.Config().round_iters == 32
    with pytest.raises(cu._lib.Cu2bError):
        cu.Config().read_config(tmp_path / "missing.cfg")


def test_print_config_format():
    txt = cu.Config().format()  # config.cu:50-64
    assert txt.splitlines() == [
        "Hyperparameters:", "total_iterations: 5000", "n_factors: 50", "learning_rate: 0.010000", "P_reg: 0.020000",
        "Q_reg: 0.020000", "user_bias_reg: 0.020000", "item_bias_reg: 0.020000", "is_train: true", "n_threads: 32",
        "check_error: 500", "patience: 2.000000", "learning_rate_decay: 0.200000"]


@pytest.mark.parametrize("size,k", [(64, 2), (1000, 32), (257, 128), (50, 50)])
def test_initialize_normal_array_bits(golden_dir, size, k):
    want = np.fromfile(os.path.join(golden_dir, "ref_init_normal_%d_%d.bin" % (size, k)), dtype=np.float32)
    assert cu.initialize_normal_array(size, k).view(np.uint32).tolist() == want.view(np.uint32).tolist()


@pytest.mark.parametrize("size,k,mean", [(1 << 18, 128, 0.0), ((1 << 18) + 1, 50, 0.0), (3_000_001, 64, 0.25), (9_000_000, 128, 0.0)])
def test_initialize_normal_array_parallel_path_is_the_sequential_stream(size, k, mean):
    """>= 2^18 values take the threaded path (attempts of the polar method evaluated at their fixed
    positions of the mt19937 stream); it must reproduce std::normal_distribution bit for bit."""
    got = cu.initialize_normal_array(size, k, mean=mean)
    want = O.init_normal(size, k, mean=mean)
    assert got.view(np.uint32).tobytes() == want.view(np.uint32).tobytes()


def test_read_array_parallel_path_and_errors(tmp_path):
    rng = np.random.RandomState(9)
    mat = (rng.standard_normal((30000, 16)) * 10.0 ** rng.randint(-3, 5, (30000, 16))).astype(np.float32)
    cu.writeCSV(tmp_path / "m.csv", mat.ravel(), 30000, 16)
    assert os.path.getsize(tmp_path / "m.csv") > (1 << 20)
    arr, r, c = cu.read_array(tmp_path / "m.csv")
    want = np.array([np.float32(t) for t in (tmp_path / "m.csv").read_text().replace("\n", ",").split(",")[:-1]], np.float32)
    assert (r, c) == (30000, 30000 * 16) and arr.tobytes() == want.tobytes()
    # trailing comma yields no extra piece, an empty line is a row without values, no final newline
    (tmp_path / "odd.csv").write_text("1.5,2.5,\n\n 3e1 ,0x10,7")
    arr, r, c = cu.read_array(tmp_path / "odd.csv")
    assert (r, c) == (3, 5) and arr.tolist() == [1.5, 2.5, 30.0, 16.0, 7.0]
    # std::stof would throw on these (util.cu:63)
    for txt in ("1.0,,2.0\n", "1.0,abc\n", "1.0\n\r\n"):
        (tmp_path / "bad.csv").write_text(txt)
        with pytest.raises(cu._lib.Cu2bError, match="not a number"):
            cu.read_array(tmp_path / "bad.csv")
    big_bad = (tmp_path / "m.csv").read_text().replace("\n", "\nx,", 20000).replace("\nx,", "\n", 19999)
    (tmp_path / "bigbad.csv").write_text(big_bad)
    with pytest.raises(cu._lib.Cu2bError, match="not a number: 'x'"):
        cu.read_array(tmp_path / "bigbad.csv")


def test_write_csv_byte_identical_to_reference(tmp_path, golden_dir):
    mat = np.fromfile(os.path.join(golden_dir, "write_csv_input.bin"), dtype=np.float32)
    cu.writeCSV(tmp_path / "o.csv", mat, 24, 5)
    assert (tmp_path / "o.csv").read_bytes() == open(os.path.join(golden_dir, "ref_write_csv.csv"), "rb").read()


def test_write_csv_matches_printf_on_random_floats(tmp_path):
    rng = np.random.RandomState(11)
    bits = rng.randint(0, 2 ** 32, 200000, dtype=np.uint64).astype(np.uint32)
    vals = bits.view(np.float32)
    vals = vals[np.isfinite(vals)]
    small = (rng.standard_normal(100000) * 10.0 ** rng.randint(-9, 6, 100000)).astype(np.float32)
    ties = (np.arange(1, 4001, dtype=np.float64) / 2 ** 12).astype(np.float32)  # exact binary fractions
    allv = np.concatenate([vals, small, ties]).astype(np.float32)
    allv = allv[: (len(allv) // 7) * 7]
    cu.writeCSV(tmp_path / "r.csv", allv, len(allv) // 7, 7)
    got = (tmp_path / "r.csv").read_text().replace("\n", ",").split(",")[:-1]
    want = ["%f" % float(v) for v in allv]
    assert got == want


def test_write_to_file_naming_and_read_array(tmp_path, golden_dir, fixtures_dir, ref_tests):
    # tests/test_util.cu:49-92 writes five components; we also read them back (util.cu:52-76)
    P = np.ones(12, np.float32)
    cu.writeToFile(tmp_path, "test_ratings", "csv", "p", P, 6, 2, 2)
    cu.writeToFile(tmp_path, "test_ratings", "csv", "global_bias", np.array([3.5], np.float32), 1, 1, 2)
    assert (tmp_path / "test_ratings_f2_p.csv").read_text() == "1.000000,1.000000\n" * 6
    assert (tmp_path / "test_ratings_f2_global_bias.csv").read_text() == "3.500000\n"
    arr, r, c = cu.read_array(tmp_path / "test_ratings_f2_p.csv")
    assert r == 6 and c == 12 and arr.tolist() == [1.0] * 12
    arr, r, c = cu.read_array(os.path.join(fixtures_dir, "test_Q.csv"))
    g = json.load(open(os.path.join(golden_dir, "ref_read_array.json")))
    assert (r, c) == (g["n_rows"], g["n_cols"]) and arr.tolist() == g["values"]
    assert np.allclose(arr[:10], ref_tests["test_util.cu:43"]["read_array_first10"], atol=1e-3)
    with pytest.raises(cu._lib.Cu2bError):
        cu.read_array(tmp_path / "nope.csv")


def test_synth_ratings_contract():
    tr, te = cu.synth_ratings(943, 1682, 100000, integer_ratings=False)
    assert 90000 < len(tr) + len(te) < 110000
    assert 0.07 < len(te) / (len(tr) + len(te)) < 0.13
    for part in (tr, te):
        assert np.all(np.diff(part["user"]) >= 0)  # grouped by ascending user
        assert part["item"].min() >= 0 and part["item"].max() < 1682
        assert part["rating"].min() >= 0.5 and part["rating"].max() <= 5.0
        assert np.all(part["rating"] * 2 == np.round(part["rating"] * 2))
    assert np.array_equal(np.unique(tr["user"]), np.arange(943))  # every user has a train rating
    allr = np.concatenate([tr, te])
    key = allr["user"].astype(np.int64) * 1682 + allr["item"]
    assert len(np.unique(key)) == len(key)  # (user, item) pairs without replacement
    tr2, te2 = cu.synth_ratings(943, 1682, 100000, integer_ratings=False)
    assert tr2.tobytes() == tr.tobytes() and te2.tobytes() == te.tobytes()
    # popularity and activity are skewed
    pop = np.sort(np.bincount(allr["item"], minlength=1682))[::-1]
    act = np.sort(np.bincount(allr["user"], minlength=943))[::-1]
    assert pop[0] > 8 * np.median(pop) and act[0] > 5 * np.median(act)
    tri, _ = cu.synth_ratings(200, 300, 5000, integer_ratings=True)
    assert set(np.unique(tri["rating"])) <= {1.0, 2.0, 3.0, 4.0, 5.0}


@pytest.mark.parametrize("B", [1, 2, 7, 16, 64])
def test_block_schedule_matches_oracle_and_is_conflict_free(B):
    rng = np.random.RandomState(B)
    U, I, n = 97, 61, 5000
    coo = np.zeros(n, dtype=cu.RATING_DTYPE)
    coo["user"], coo["item"], coo["rating"] = np.sort(rng.randint(0, U, n)), rng.randint(0, I, n), rng.randint(1, 6, n)
    order, used = cu.block_schedule_order(coo, U, I, B)
    assert used == B and sorted(order.tolist()) == list(range(n))
    assert order.tolist() == O.block_schedule_order(coo, U, I, B).tolist()
    # blocks that share a round never share a user block or an item block
    ubs, ibs = -(-U // B), -(-I // B)
    ub, ib = coo["user"][order] // ubs, coo["item"][order] // ibs
    rnd = (ib - ub) % B
    assert np.all(np.diff(rnd) >= 0)  # round-major
    for s in range(B):
        sel = rnd == s
        pairs = set(zip(ub[sel].tolist(), ib[sel].tolist()))
        assert len({a for a, _ in pairs}) == len(pairs) == len({b for _, b in pairs})
    _, auto = cu.block_schedule_order(coo, U, I, 0)
    assert auto == min(U, I)


def test_read_csv_malformed_rows_match_the_reference_stream_semantics(tmp_path, golden_dir):
    """Rows a strtof-based reader would swallow (inf, nan, hex, 1e50, dangling exponents) and ids that overflow int:
    the reference reads with operator>> and stops at the first row the stream rejects. Goldens produced by the
    unmodified reference readCSV (tests/golden/make_malformed_golden.py)."""
    cases = json.load(open(os.path.join(golden_dir, "ref_read_csv_malformed.json")))
    assert len(cases) >= 20
    for name, want in cases.items():
        path = tmp_path / (name + ".csv")
        path.write_text(want["text"])
        got, rows, cols, gb = cu.readCSV(str(path))
        assert (len(got), rows, cols) == (want["n"], want["rows"], want["cols"]), name
        assert got["user"].tolist() == want["users"] and got["item"].tolist() == want["items"], name
        assert got["rating"].view(np.uint32).tolist() == want["rating_bits"], name
        assert int(np.float32(gb).view(np.uint32)) == want["global_bias_bits"], name
        assert np.all(np.isfinite(got["rating"])), name


def test_read_csv_through_the_binary_sidecar(tmp_path):
    """the sidecar returns exactly what parsing returns, belongs to one state of its source (size + mtime),
    is rewritten when the source changes, and a damaged or foreign file at its place is ignored."""
    import os
    rng = np.random.RandomState(5)
    n = 20000
    r = np.zeros(n, dtype=cu.RATING_DTYPE)
    r["user"] = np.sort(rng.randint(0, 700, n))
    r["item"] = rng.randint(0, 300, n)
    r["rating"] = rng.randint(1, 11, n) * 0.5
    path = tmp_path / "ratings.csv"
    cu.write_ratings_csv(path, r)
    want = cu.readCSV(path)
    side = str(path) + ".cu2bcache"

    def same(a, b):
        return (a[0].tobytes() == b[0].tobytes() and a[1:3] == b[1:3]
                and np.float32(a[3]).tobytes() == np.float32(b[3]).tobytes())
    first = cu.readCSV(path, cache=True)
    assert not cu.readCSV.last_hit and os.path.exists(side) and same(first, want)
    assert os.path.getsize(side) == 48 + 12 * n
    again = cu.readCSV(path, cache=True)
    assert cu.readCSV.last_hit and same(again, want)
    # the source changes (one more rating): the stale sidecar is not used and is replaced
    with open(path, "a") as f:
        f.write("701,5,4.5\n")
    st = os.stat(path)
    os.utime(path, ns=(st.st_atime_ns, st.st_mtime_ns + 1_000_000_000))
    grown = cu.readCSV(path, cache=True)
    assert not cu.readCSV.last_hit and len(grown[0]) == n + 1 and same(grown, cu.readCSV(path))
    assert cu.readCSV(path, cache=True)[0].tobytes() == grown[0].tobytes() and cu.readCSV.last_hit
    # same size, later mtime: still a different state of the file
    os.utime(path, ns=(st.st_atime_ns, st.st_mtime_ns + 5_000_000_000))
    cu.readCSV(path, cache=True)
    assert not cu.readCSV.last_hit
    # a truncated sidecar and a foreign file are ignored (and replaced by a good one)
    for damage in (lambda: open(side, "r+b").truncate(48 + 12 * 100), lambda: open(side, "wb").write(b"not a sidecar" * 10)):
        damage()
        got = cu.readCSV(path, cache=True)
        assert not cu.readCSV.last_hit and same(got, grown)
        cu.readCSV(path, cache=True)
        assert cu.readCSV.last_hit
    # an explicit sidecar path; an unwritable one is not an error of the read
    other = tmp_path / "elsewhere.bin"
    assert same(cu.readCSV(path, cache=other), grown) and other.exists()
    assert same(cu.readCSV(path, cache=other), grown) and cu.readCSV.last_hit
    assert same(cu.readCSV(path, cache=tmp_path / "no" / "such" / "dir" / "x.bin"), grown) and not cu.readCSV.last_hit
    # a missing source reports what the plain reader reports
    with pytest.raises(cu._lib.Cu2bError) as err:
        cu.readCSV(tmp_path / "absent.csv", cache=True)
    assert err.value.status == 2


def test_shim_read_csv_uses_the_sidecar_when_asked(tmp_path):
    """CU2B_CSV_CACHE=1 routes the shim's readCSV (what bin/mf calls) through the sidecar: same CSR either way."""
    import os, subprocess
    r = np.zeros(500, dtype=cu.RATING_DTYPE)
    r["user"] = np.sort(np.arange(500) % 40)
    r["item"] = (np.arange(500) * 7) % 60
    r["rating"] = 1 + (np.arange(500) % 9) * 0.5
    path = tmp_path / "r.csv"
    cu.write_ratings_csv(path, r)
    src = tmp_path / "t.cpp"
    src.write_text("""#include "cu2rec_shim.h"
#include <cstdio>
int main(int argc, char **argv) { int rows, cols; float gb; auto v = readCSV(argv[1], &rows, &cols, &gb);
  double s = 0; for (auto &x : v) s += x.rating * (x.userID + 1) + x.itemID;
  printf("%zu %d %d %.9g %.9g\\n", v.size(), rows, cols, gb, s); return 0; }
""")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "t"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(root, "include"), "-I" + os.path.join(root, "cu2rec_b200", "csrc"),
                    str(src), "-o", str(exe), "-L" + os.path.join(root, "cu2rec_b200", "lib"), "-lcu2b",
                    "-Wl,-rpath," + os.path.join(root, "cu2rec_b200", "lib")], check=True)
    plain = subprocess.run([str(exe), str(path)], check=True, capture_output=True, text=True).stdout
    assert not os.path.exists(str(path) + ".cu2bcache")
    env = dict(os.environ, CU2B_CSV_CACHE="1")
    miss = subprocess.run([str(exe), str(path)], check=True, capture_output=True, text=True, env=env).stdout
    assert os.path.exists(str(path) + ".cu2bcache")
    hit = subprocess.run([str(exe), str(path)], check=True, capture_output=True, text=True, env=env).stdout
    assert plain == miss == hit and plain.split()[0] == "500"
