"""ctypes binding of oracle/liboracle.so -- the CPU restatement of the reference's algorithm.
TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ODIR, "liboracle.so")
REF_DIR = os.path.join(ODIR, "_ref")

TRIPLET = np.dtype([("user", np.int32), ("item", np.int32), ("rating", np.float32)])
FLAVOUR_REF, FLAVOUR_KERNEL, FLAVOUR_LOSS_KERNEL = 0, 1, 2


class Hyper(C.Structure):
    _fields_ = [("n_factors", C.c_int), ("learning_rate", C.c_float), ("P_reg", C.c_float), ("Q_reg", C.c_float),
                ("user_bias_reg", C.c_float), ("item_bias_reg", C.c_float), ("is_train", C.c_int)]


class LogRow(C.Structure):
    _fields_ = [("iteration", C.c_int), ("train_mae", C.c_float), ("train_rmse", C.c_float), ("test_mae", C.c_float),
                ("test_rmse", C.c_float), ("learning_rate", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(ODIR, "mf_oracle.cpp")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.run(["make", "-C", ODIR, "oracle"], check=True, capture_output=True)
        _lib = C.CDLL(LIB)
        _lib.orc_predict.restype = C.c_float
        _lib.orc_sgd_update_one.restype = C.c_float
        _lib.orc_read_csv.restype = C.c_long
        _lib.orc_sample_per_user.restype = C.c_long
    return _lib


def ref_binary(name):
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None


def run_mf_cpu(cfg, train_csv, test_csv, attempts=8):
    """Runs the unmodified reference mf_cpu and returns its stdout. mf_sequential.cu:111 samples from
    the inclusive range [low, high], so the last user occasionally reads one element past the
    rating arrays and the process can die on what it finds there: retry."""
    for _ in range(attempts):
        p = subprocess.run([ref_binary("mf_cpu"), "-c", str(cfg), str(train_csv), str(test_csv)], capture_output=True, text=True)
        if p.returncode == 0:
            return p.stdout
    raise RuntimeError("reference mf_cpu crashed %d times in a row (rc=%d)" % (attempts, p.returncode))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def hyper(k, lr=0.01, P_reg=0.02, Q_reg=0.02, ub_reg=0.02, ib_reg=0.02, is_train=1):
    return Hyper(k, lr, P_reg, Q_reg, ub_reg, ib_reg, is_train)


def hyper_from_cfg(cfg):
    return Hyper(cfg.n_factors, cfg.learning_rate, cfg.P_reg, cfg.Q_reg, cfg.user_bias_reg, cfg.item_bias_reg,
                 cfg.is_train)


def init_normal(size, n_factors, mean=0.0, stddev=1.0, seed=42):
    out = np.empty(size, dtype=np.float32)
    lib().orc_init_normal(_p(out), size, n_factors, C.c_float(mean), C.c_float(stddev), seed)
    return out


def read_config(path):
    v = [C.c_int(0), C.c_int(5000), C.c_int(50), C.c_float(0.01), C.c_int(42), C.c_float(0.02), C.c_float(0.02),
         C.c_float(0.02), C.c_float(0.02)]
    n = lib().orc_read_config(str(path).encode(), *[C.byref(x) for x in v])
    return n, [x.value for x in v]


def read_csv(path):
    rows, cols, gb = C.c_int(), C.c_int(), C.c_float()
    n = lib().orc_read_csv(str(path).encode(), None, C.c_long(0), C.byref(rows), C.byref(cols), C.byref(gb))
    if n < 0:
        raise IOError(path)
    out = np.empty(n, dtype=TRIPLET)
    lib().orc_read_csv(str(path).encode(), _p(out), C.c_long(n), C.byref(rows), C.byref(cols), C.byref(gb))
    return out, rows.value, cols.value, np.float32(gb.value)


def build_csr(ratings, rows):
    ratings = np.ascontiguousarray(ratings, dtype=TRIPLET)
    n = ratings.shape[0]
    indptr = np.zeros(rows + 1, dtype=np.int32)
    indices = np.empty(n, dtype=np.int32)
    data = np.empty(n, dtype=np.float32)
    lib().orc_build_csr(_p(ratings), C.c_long(n), rows, _p(indptr), _p(indices), _p(data))
    return indptr, indices, data


def predict(p, q, ub, ib, mu, flavour=FLAVOUR_REF):
    p = np.ascontiguousarray(p, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    return np.float32(lib().orc_predict(_p(p), _p(q), p.shape[0], C.c_float(ub), C.c_float(ib), C.c_float(mu), flavour))


def sgd_apply_stream(stream, P, Q, ub, ib, mu, h, flavour=FLAVOUR_REF):
    stream = np.ascontiguousarray(stream, dtype=TRIPLET)
    P, Q, ub, ib = (np.array(np.ascontiguousarray(x, np.float32), copy=True) for x in (P, Q, ub, ib))
    lib().orc_sgd_apply_stream(_p(stream), C.c_long(stream.shape[0]), _p(P), _p(Q), _p(ub), _p(ib), C.c_float(mu),
                               C.byref(h), flavour)
    return P, Q, ub, ib


def residuals(indptr, indices, data, P, Q, ub, ib, mu, k, flavour=FLAVOUR_REF):
    P, Q, ub, ib = (np.ascontiguousarray(x, np.float32) for x in (P, Q, ub, ib))
    err = np.empty(indices.shape[0], dtype=np.float32)
    lib().orc_residuals(indptr.shape[0] - 1, _p(indptr), _p(indices), _p(data), _p(P), _p(Q), _p(ub), _p(ib),
                        C.c_float(mu), k, _p(err), flavour)
    return err


def error_metrics(err):
    err = np.ascontiguousarray(err, np.float32)
    mae, rmse = C.c_float(), C.c_float()
    lib().orc_error_metrics(_p(err), C.c_long(err.shape[0]), C.byref(mae), C.byref(rmse))
    return np.float32(mae.value), np.float32(rmse.value)


def loss(indptr, indices, data, P, Q, ub, ib, mu, k, flavour=FLAVOUR_REF):
    P, Q, ub, ib = (np.ascontiguousarray(x, np.float32) for x in (P, Q, ub, ib))
    mae, rmse, sse, sae = C.c_float(), C.c_float(), C.c_double(), C.c_double()
    lib().orc_loss(indptr.shape[0] - 1, _p(indptr), _p(indices), _p(data), _p(P), _p(Q), _p(ub), _p(ib),
                   C.c_float(mu), k, C.byref(mae), C.byref(rmse), C.byref(sse), C.byref(sae), flavour)
    return np.float32(mae.value), np.float32(rmse.value), sse.value, sae.value


def sample_per_user(indptr, indices, data, seed, iter0, n_iter):
    rows = indptr.shape[0] - 1
    n_active = int(np.count_nonzero(np.diff(indptr)))
    out = np.empty(n_iter * n_active, dtype=TRIPLET)
    n = lib().orc_sample_per_user(rows, _p(indptr), _p(indices), _p(data), seed, iter0, n_iter, _p(out))
    assert n == out.shape[0]
    return out


def sample_per_rating(indptr, indices, data, seed, first, count):
    out = np.empty(count, dtype=TRIPLET)
    lib().orc_sample_per_rating(indptr.shape[0] - 1, _p(indptr), _p(indices), _p(data), seed, C.c_longlong(first),
                                C.c_longlong(count), _p(out))
    return out


def train(tr, te, P, Q, ub, ib, mu, h, seed, total_iterations, check_error=500, patience=2.0, lr_decay=0.2,
          use_decay=True, flavour=FLAVOUR_REF, iter0=0):
    """tr / te = (indptr, indices, data). Arrays are updated in place (copies returned)."""
    P, Q, ub, ib = (np.array(np.ascontiguousarray(x, np.float32), copy=True) for x in (P, Q, ub, ib))
    cap = total_iterations // check_error + 8
    log = (LogRow * cap)()
    n = lib().orc_train(tr[0].shape[0] - 1, _p(tr[0]), _p(tr[1]), _p(tr[2]), te[0].shape[0] - 1, _p(te[0]), _p(te[1]),
                        _p(te[2]), _p(P), _p(Q), _p(ub), _p(ib), C.c_float(mu), C.byref(h), seed, iter0,
                        total_iterations, check_error, C.c_float(patience), C.c_float(lr_decay), int(use_decay),
                        flavour, log, cap)
    rows = [dict(iteration=r.iteration, train_mae=r.train_mae, train_rmse=r.train_rmse, test_mae=r.test_mae,
                 test_rmse=r.test_rmse, learning_rate=r.learning_rate) for r in log[:min(n, cap)]]
    return P, Q, ub, ib, rows


def block_schedule_order(coo, rows, cols, B):
    coo = np.ascontiguousarray(coo, dtype=TRIPLET)
    order = np.empty(coo.shape[0], dtype=np.int64)
    lib().orc_block_schedule_order(_p(coo), C.c_long(coo.shape[0]), rows, cols, B, _p(order))
    return order


def predict_topk(P, Q, ub, ib, mu, topk, exclude=None):
    """exclude = (indptr, indices) or None -> (items [U, topk], scores [U, topk])."""
    P, Q, ub, ib = (np.ascontiguousarray(x, np.float32) for x in (P, Q, ub, ib))
    rows, cols = ub.shape[0], ib.shape[0]
    k = P.size // rows
    items = np.empty((rows, topk), dtype=np.int32)
    scores = np.empty((rows, topk), dtype=np.float32)
    if exclude is None:
        lib().orc_predict_topk(rows, cols, k, _p(P), _p(Q), _p(ub), _p(ib), C.c_float(mu), None, None, 0, topk, _p(items), _p(scores))
    else:
        ip, ii = np.ascontiguousarray(exclude[0], np.int32), np.ascontiguousarray(exclude[1], np.int32)
        lib().orc_predict_topk(rows, cols, k, _p(P), _p(Q), _p(ub), _p(ib), C.c_float(mu), _p(ip), _p(ii), ip.shape[0] - 1, topk,
                               _p(items), _p(scores))
    return items, scores
