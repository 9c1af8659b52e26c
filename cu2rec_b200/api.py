"""Python harness over the C ABI, named after the reference's interface for this path so the
parity tests read like the reference's own tests (util.h, config.h, loss.h, training.h).
Everything computes inside libcu2b.so; numpy only carries the buffers."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import Config as _CConfig, Csr, Metrics, Rating, Stats, check

RATING_DTYPE = np.dtype([("user", np.int32), ("item", np.int32), ("rating", np.float32)])

MODE_HOGWILD, MODE_DETERMINISTIC = 0, 1
SAMPLER_PER_USER, SAMPLER_PER_RATING = 0, 1


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ------------------------------------------------------------------------------------------
# config::Config (config.h:20-58)
# ------------------------------------------------------------------------------------------
class Config:
    """Mirror of config::Config. Attribute names are the reference's field names."""

    def __init__(self, **kw):
        self.c = _CConfig()
        _lib.load().cu2b_config_default(C.byref(self.c))
        for k, v in kw.items():
            setattr(self, k, v)

    def __getattr__(self, name):
        c = object.__getattribute__(self, "c")
        if name in dict(_CConfig._fields_):
            return getattr(c, name)
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name == "c":
            object.__setattr__(self, name, value)
        elif name in dict(_CConfig._fields_):
            setattr(self.c, name, value)
        else:
            raise AttributeError(name)

    def read_config(self, path):  # config.cu:7-13
        check(_lib.load().cu2b_config_read(str(path).encode(), C.byref(self.c)))
        return True

    def write_config(self, path):  # config.cu:15-22
        check(_lib.load().cu2b_config_write(str(path).encode(), C.byref(self.c)))
        return True

    def format(self):  # config.cu:50-64 print_config text
        buf = C.create_string_buffer(1024)
        _lib.load().cu2b_config_format(C.byref(self.c), buf, 1024)
        return buf.value.decode()

    def print_config(self):
        print(self.format(), end="")

    def copy(self):
        other = Config()
        C.memmove(C.byref(other.c), C.byref(self.c), C.sizeof(_CConfig))
        return other


# ------------------------------------------------------------------------------------------
# util.h
# ------------------------------------------------------------------------------------------
def readCSV(path, cache=None):
    """util.cu:17-45 -> (ratings[RATING_DTYPE], rows, cols, global_bias). cache: None = parse the text; True = through the
    binary sidecar "<path>.cu2bcache"; a path = through that sidecar (cu2b_read_csv_cached). readCSV.last_hit tells
    whether the last cached read used the sidecar."""
    lib = _lib.load()
    p = C.POINTER(Rating)()
    n = C.c_int64()
    rows, cols, gb = C.c_int(), C.c_int(), C.c_float()
    if cache is None or cache is False:
        check(lib.cu2b_read_csv(str(path).encode(), C.byref(p), C.byref(n), C.byref(rows), C.byref(cols), C.byref(gb)))
    else:
        hit = C.c_int()
        where = None if cache is True else str(cache).encode()
        check(lib.cu2b_read_csv_cached(str(path).encode(), where, C.byref(p), C.byref(n), C.byref(rows), C.byref(cols),
                                       C.byref(gb), C.byref(hit)))
        readCSV.last_hit = bool(hit.value)
    try:
        out = np.empty(n.value, dtype=RATING_DTYPE)
        if n.value:
            C.memmove(out.ctypes.data, p, n.value * RATING_DTYPE.itemsize)
    finally:
        lib.cu2b_free(p)
    return out, rows.value, cols.value, np.float32(gb.value)


readCSV.last_hit = False


@dataclass
class CSRMatrix:
    """Host-side CSR view (matrix.h:11-19); the device copy lives inside a session."""
    rows: int
    cols: int
    indptr: np.ndarray
    indices: np.ndarray
    data: np.ndarray

    @property
    def nonzeros(self):
        return int(self.indices.shape[0])

    def c(self):
        m = Csr()
        m.rows, m.cols, m.nonzeros = self.rows, self.cols, self.nonzeros
        m.indptr, m.indices, m.data = _ptr(self.indptr), _ptr(self.indices), _ptr(self.data)
        m.on_device = 0
        return m


def createSparseMatrix(ratings, rows, cols):
    """util.cu:152-179 (host part)."""
    ratings = np.ascontiguousarray(ratings, dtype=RATING_DTYPE)
    n = ratings.shape[0]
    indptr = np.empty(rows + 1, dtype=np.int32)
    indices = np.empty(n, dtype=np.int32)
    data = np.empty(n, dtype=np.float32)
    check(_lib.load().cu2b_build_csr(_ptr(ratings), n, rows, _ptr(indptr), _ptr(indices), _ptr(data)))
    return CSRMatrix(rows, cols, indptr, indices, data)


def read_array(path):
    """util.cu:52-76 -> (flat float32 array, n_rows, n_cols as accumulated by the reference)."""
    lib = _lib.load()
    p = C.POINTER(C.c_float)()
    r, c = C.c_int(), C.c_int()
    check(lib.cu2b_read_array(str(path).encode(), C.byref(p), C.byref(r), C.byref(c)))
    try:
        out = np.empty(c.value, dtype=np.float32)
        if c.value:
            C.memmove(out.ctypes.data, p, c.value * 4)
    finally:
        lib.cu2b_free(p)
    return out, r.value, c.value


def writeCSV(path, data, rows, cols):  # util.cu:86-97
    data = _f32(data)
    check(_lib.load().cu2b_write_csv(str(path).encode(), _ptr(data), rows, cols))


def writeToFile(parent_dir, base_filename, extension, component, data, rows, cols, factors):  # util.cu:99-103
    data = _f32(data)
    check(_lib.load().cu2b_write_component(str(parent_dir).encode(), base_filename.encode(), extension.encode(),
                                           component.encode(), _ptr(data), rows, cols, factors))


def initialize_normal_array(size, n_factors, mean=0.0, stddev=1.0, seed=42):  # util.cu:124-144
    out = np.empty(size, dtype=np.float32)
    _lib.load().cu2b_init_normal(_ptr(out), size, n_factors, mean, stddev, seed)
    return out


def synth_ratings(users, items, target_ratings, rank=16, noise=0.5, integer_ratings=True, test_fraction=0.1,
                  seed=20240607):
    """Synthetic low-rank-plus-noise ratings -> (train, test) RATING_DTYPE arrays."""
    lib = _lib.load()
    ntr, nte = C.c_int64(), C.c_int64()
    args = (users, items, target_ratings, rank, noise, int(integer_ratings), test_fraction, seed)
    check(lib.cu2b_synth_ratings(*args, None, C.byref(ntr), None, C.byref(nte)))
    train = np.empty(ntr.value, dtype=RATING_DTYPE)
    test = np.empty(nte.value, dtype=RATING_DTYPE)
    check(lib.cu2b_synth_ratings(*args, _ptr(train), C.byref(ntr), _ptr(test), C.byref(nte)))
    return train, test


# ------------------------------------------------------------------------------------------
# loss.h
# ------------------------------------------------------------------------------------------
def calculate_loss_gpu(P, Q, n_factors, matrix, user_bias, item_bias, global_bias):
    """loss.cu:40-49 -> residual vector error[i] = data[i] - prediction."""
    P, Q, ub, ib = _f32(P), _f32(Q), _f32(user_bias), _f32(item_bias)
    err = np.empty(matrix.nonzeros, dtype=np.float32)
    m = matrix.c()
    check(_lib.load().cu2b_residuals(C.byref(m), _ptr(P), _ptr(Q), _ptr(ub), _ptr(ib), float(global_bias), n_factors, _ptr(err)))
    return err


def get_error_metrics_gpu(errors):
    """loss.cu:196-200 -> (mae, rmse)."""
    errors = _f32(errors)
    mae, rmse = C.c_float(), C.c_float()
    check(_lib.load().cu2b_error_metrics(_ptr(errors), errors.shape[0], C.byref(mae), C.byref(rmse)))
    return np.float32(mae.value), np.float32(rmse.value)


def loss(P, Q, n_factors, matrix, user_bias, item_bias, global_bias):
    """Fused single pass: (mae, rmse) of the model on `matrix`."""
    P, Q, ub, ib = _f32(P), _f32(Q), _f32(user_bias), _f32(item_bias)
    mae, rmse = C.c_float(), C.c_float()
    m = matrix.c()
    check(_lib.load().cu2b_loss(C.byref(m), _ptr(P), _ptr(Q), _ptr(ub), _ptr(ib), float(global_bias), n_factors,
                                C.byref(mae), C.byref(rmse)))
    return np.float32(mae.value), np.float32(rmse.value)


# ------------------------------------------------------------------------------------------
# sgd.h
# ------------------------------------------------------------------------------------------
def sample_per_user(matrix, seed, iter0, n_iter):
    """sgd.cu:27-37 sampling step -> triplet stream [n_iter * active_users]."""
    lib = _lib.load()
    m = matrix.c()
    n = C.c_int64()
    check(lib.cu2b_sample_per_user(C.byref(m), seed, iter0, n_iter, None, C.byref(n)))
    out = np.empty(n.value, dtype=RATING_DTYPE)
    check(lib.cu2b_sample_per_user(C.byref(m), seed, iter0, n_iter, _ptr(out), C.byref(n)))
    return out


def sample_per_rating(matrix, seed, first_update, n_updates):
    """per_rating sampler: updates [first_update, first_update + n_updates) of a run (shuffled passes over the ratings)."""
    m = matrix.c()
    out = np.empty(n_updates, dtype=RATING_DTYPE)
    check(_lib.load().cu2b_sample_per_rating(C.byref(m), seed, first_update, n_updates, _ptr(out)))
    return out


def sgd_apply(stream, P, Q, user_bias, item_bias, global_bias, cfg, order=0):
    """sgd.cu:40-72 arithmetic over an explicit stream; returns updated copies."""
    stream = np.ascontiguousarray(stream, dtype=RATING_DTYPE)
    P, Q, ub, ib = (np.array(_f32(x), copy=True) for x in (P, Q, user_bias, item_bias))
    k = cfg.n_factors
    rows, cols = P.size // k, Q.size // k
    check(_lib.load().cu2b_sgd_apply(_ptr(stream), stream.shape[0], _ptr(P), rows, _ptr(Q), cols, _ptr(ub), _ptr(ib),
                                     float(global_bias), C.byref(cfg.c), order))
    return P, Q, ub, ib


def sgd_blocked(coo, P, Q, user_bias, item_bias, global_bias, cfg, n_blocks, n_passes=1):
    """Deterministic conflict-free pass(es) over `coo`; returns updated copies."""
    coo = np.ascontiguousarray(coo, dtype=RATING_DTYPE)
    P, Q, ub, ib = (np.array(_f32(x), copy=True) for x in (P, Q, user_bias, item_bias))
    k = cfg.n_factors
    rows, cols = P.size // k, Q.size // k
    check(_lib.load().cu2b_sgd_blocked(_ptr(coo), coo.shape[0], _ptr(P), rows, _ptr(Q), cols, _ptr(ub), _ptr(ib),
                                       float(global_bias), C.byref(cfg.c), n_blocks, n_passes))
    return P, Q, ub, ib


def block_schedule_order(coo, rows, cols, n_blocks=0):
    """Canonical sequential order of the deterministic schedule -> (permutation, B used)."""
    coo = np.ascontiguousarray(coo, dtype=RATING_DTYPE)
    order = np.empty(coo.shape[0], dtype=np.int64)
    used = C.c_int()
    check(_lib.load().cu2b_block_schedule_order(_ptr(coo), coo.shape[0], rows, cols, n_blocks, _ptr(order), C.byref(used)))
    return order, used.value


# ------------------------------------------------------------------------------------------
# training.h
# ------------------------------------------------------------------------------------------
def _metrics_rows(buf, n):
    return [dict(iteration=m.iteration, train_mae=m.train_mae, train_rmse=m.train_rmse, test_mae=m.test_mae,
                 test_rmse=m.test_rmse, learning_rate=m.learning_rate) for m in buf[:n]]


def _stats_dict(st):
    return {name: getattr(st, name) for name, _ in Stats._fields_}


def train(train_matrix, test_matrix, cfg, global_bias, Q=None, item_bias=None):
    """training.h:12-15. With Q/item_bias None this is the 8-argument overload (they are
    initialised inside, training.cu:208-217); otherwise the 10-argument one. cfg is updated
    in place (learning_rate, cur_iterations). Returns dict(P, Q, losses, user_bias, item_bias,
    log, stats)."""
    lib = _lib.load()
    k = cfg.n_factors
    rows, cols = train_matrix.rows, train_matrix.cols
    init_item = Q is None
    P = np.empty(rows * k, dtype=np.float32)
    ub = np.empty(rows, dtype=np.float32)
    Q = np.empty(cols * k, dtype=np.float32) if init_item else np.array(_f32(Q).reshape(-1), copy=True)
    ib = np.empty(cols, dtype=np.float32) if item_bias is None else np.array(_f32(item_bias), copy=True)
    losses = np.empty(max(1, cfg.total_iterations), dtype=np.float32)
    cap = cfg.total_iterations // max(1, cfg.check_error) + 8
    log = (Metrics * cap)()
    n_log = C.c_int()
    st = Stats()
    tm, te = train_matrix.c(), test_matrix.c()
    check(lib.cu2b_train(C.byref(tm), C.byref(te), C.byref(cfg.c), _ptr(P), _ptr(Q), _ptr(ub), _ptr(ib),
                         float(global_bias), int(init_item), _ptr(losses), log, cap, C.byref(n_log), C.byref(st)))
    return dict(P=P.reshape(rows, k), Q=Q.reshape(cols, k), losses=losses[:cfg.total_iterations], user_bias=ub,
                item_bias=ib, log=_metrics_rows(log, min(cap, n_log.value)), stats=_stats_dict(st))


class Session:
    """Resident-data training session (cu2b_session_*)."""

    def __init__(self, train_matrix, test_matrix, cfg, P, Q, user_bias, item_bias, global_bias, device=0):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        self.k = cfg.n_factors
        self.rows, self.cols = train_matrix.rows, train_matrix.cols
        self.cfg = cfg
        P, Q, ub, ib = _f32(P), _f32(Q), _f32(user_bias), _f32(item_bias)
        tm, te = train_matrix.c(), test_matrix.c()
        check(self.lib.cu2b_session_create(C.byref(self.h), device, C.byref(tm), C.byref(te), C.byref(cfg.c),
                                           _ptr(P), _ptr(Q), _ptr(ub), _ptr(ib), float(global_bias)))

    def run(self, n_iterations):
        check(self.lib.cu2b_session_run(self.h, n_iterations))

    def run_download(self, n_iterations, out=None):
        """run() + download() as one call: the model's D2H overlaps the run's last loss check. -> (P, Q, user_bias, item_bias)"""
        if out is None:
            out = (np.empty((self.rows, self.k), dtype=np.float32), np.empty((self.cols, self.k), dtype=np.float32),
                   np.empty(self.rows, dtype=np.float32), np.empty(self.cols, dtype=np.float32))
        P, Q, ub, ib = out
        check(self.lib.cu2b_session_run_download(self.h, n_iterations, _ptr(P), _ptr(Q), _ptr(ub), _ptr(ib)))
        return P, Q, ub, ib

    def reload(self, train_matrix, test_matrix, P, Q, user_bias, item_bias, global_bias):
        """Start over on same-shaped data with a new initial model (cu2b_session_reload)."""
        P, Q, ub, ib = _f32(P), _f32(Q), _f32(user_bias), _f32(item_bias)
        tm, te = train_matrix.c(), test_matrix.c()
        check(self.lib.cu2b_session_reload(self.h, C.byref(tm), C.byref(te), _ptr(P), _ptr(Q), _ptr(ub), _ptr(ib),
                                           float(global_bias)))

    def eval(self):
        v = [C.c_float() for _ in range(4)]
        check(self.lib.cu2b_session_eval(self.h, *[C.byref(x) for x in v]))
        return dict(train_mae=v[0].value, train_rmse=v[1].value, test_mae=v[2].value, test_rmse=v[3].value)

    def log(self):
        cap = self.cfg.total_iterations // max(1, self.cfg.check_error) + 8
        buf = (Metrics * cap)()
        n = C.c_int()
        check(self.lib.cu2b_session_log(self.h, buf, cap, C.byref(n)))
        return _metrics_rows(buf, min(cap, n.value))

    def download(self, out=None):
        """-> (P, Q, user_bias, item_bias). `out` = four preallocated float32 arrays (e.g. page-locked)."""
        if out is None:
            out = (np.empty((self.rows, self.k), dtype=np.float32), np.empty((self.cols, self.k), dtype=np.float32),
                   np.empty(self.rows, dtype=np.float32), np.empty(self.cols, dtype=np.float32))
        P, Q, ub, ib = out
        assert P.size == self.rows * self.k and Q.size == self.cols * self.k and ub.size == self.rows and ib.size == self.cols
        check(self.lib.cu2b_session_download(self.h, _ptr(P), _ptr(Q), _ptr(ub), _ptr(ib)))
        return P, Q, ub, ib

    def config(self):
        out = Config()
        check(self.lib.cu2b_session_get_config(self.h, C.byref(out.c)))
        return out

    def stats(self, reset=False):
        st = Stats()
        check(self.lib.cu2b_session_stats(self.h, C.byref(st), int(reset)))
        return _stats_dict(st)

    def close(self):
        if self.h:
            self.lib.cu2b_session_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------
# multi-GPU DSGD
# ------------------------------------------------------------------------------------------
@dataclass
class DsgdPartition:
    world: int
    user_block: np.ndarray       # [U] block of original user
    user_local: np.ndarray       # [U] index inside its block
    users_per_block: np.ndarray  # [G]
    item_new: np.ndarray         # [I] renumbered item id
    item_block_ptr: np.ndarray   # [G+1]
    block_nnz: np.ndarray        # [G, G]


def dsgd_partition(train, rows, cols, world):
    """LPT-balanced user / item blocks for DSGD (host only)."""
    train = np.ascontiguousarray(train, dtype=RATING_DTYPE)
    ub, ul = np.empty(rows, np.int32), np.empty(rows, np.int32)
    upb = np.empty(world, np.int32)
    inew, iptr = np.empty(cols, np.int32), np.empty(world + 1, np.int32)
    nnz = np.empty(world * world, np.int64)
    check(_lib.load().cu2b_dsgd_partition(_ptr(train), train.shape[0], rows, cols, world, _ptr(ub), _ptr(ul), _ptr(upb),
                                          _ptr(inew), _ptr(iptr), _ptr(nnz)))
    return DsgdPartition(world, ub, ul, upb, inew, iptr, nnz.reshape(world, world))


def dsgd_extract_strip(ratings, part, rank):
    """Ratings of one rank: local user ids, renumbered items, original per-user order."""
    lib = _lib.load()
    ratings = np.ascontiguousarray(ratings, dtype=RATING_DTYPE)
    n = C.c_int64()
    args = (_ptr(ratings), ratings.shape[0], _ptr(part.user_block), _ptr(part.user_local), _ptr(part.item_new), rank)
    check(lib.cu2b_dsgd_extract_strip(*args, None, C.byref(n)))
    out = np.empty(n.value, dtype=RATING_DTYPE)
    check(lib.cu2b_dsgd_extract_strip(*args, _ptr(out), C.byref(n)))
    return out


def dsgd_item_keep(train_strip, item_block_ptr, learning_rate, groups_in_flight, budget):
    """Item-step thinning fractions of one rank's strip (experimental; cu2b_dsgd_item_keep, host only)."""
    keep = np.empty(train_strip.cols, dtype=np.float32)
    iptr = None if item_block_ptr is None else np.ascontiguousarray(item_block_ptr, dtype=np.int32)
    world = 1 if iptr is None else len(iptr) - 1
    m = train_strip.c()
    check(_lib.load().cu2b_dsgd_item_keep(C.byref(m), None if iptr is None else _ptr(iptr), world, learning_rate,
                                          groups_in_flight, budget, _ptr(keep)))
    return keep


@dataclass
class DsgdRankInputs:
    train: "CSRMatrix"
    test: "CSRMatrix"
    P: np.ndarray
    Q: np.ndarray
    user_bias: np.ndarray
    item_bias: np.ndarray
    user_ids: np.ndarray
    n_train_global: int
    n_test_global: int
    n_active_global: int


def dsgd_rank_inputs(train, test, rows, cols, part, rank, P, Q, user_bias, item_bias):
    """Slices the global problem (original ids) into what rank `rank` owns."""
    k = P.size // rows
    users = np.flatnonzero(part.user_block == rank).astype(np.int32)  # ascending == local order
    tr, te = dsgd_extract_strip(train, part, rank), dsgd_extract_strip(test, part, rank)
    n_local = len(users)
    inv = np.empty(cols, np.int64)
    inv[part.item_new] = np.arange(cols)
    n_active = int(np.count_nonzero(np.bincount(np.ascontiguousarray(train)["user"], minlength=rows)))
    return DsgdRankInputs(createSparseMatrix(tr, n_local, cols), createSparseMatrix(te, n_local, cols),
                          _f32(P).reshape(rows, k)[users].copy(), _f32(Q).reshape(cols, k)[inv].copy(),
                          _f32(user_bias)[users].copy(), _f32(item_bias)[inv].copy(), users,
                          int(len(train)), int(len(test)), n_active)


class Dsgd:
    """One DSGD rank (cu2b_dsgd_*). Exchange `handle` between ranks, then connect() and run()."""

    def __init__(self, rank, world, inputs, part, cfg, global_bias, device=0):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        self.rank, self.world, self.cfg = rank, world, cfg
        self.k = cfg.n_factors
        self.rows, self.cols = inputs.train.rows, inputs.train.cols
        self._keep = inputs
        blob = C.create_string_buffer(_lib.DSGD_HANDLE_BYTES)
        tm, te = inputs.train.c(), inputs.test.c()
        iptr = np.ascontiguousarray(part.item_block_ptr, dtype=np.int32)
        check(self.lib.cu2b_dsgd_create(C.byref(self.h), device, rank, world, C.byref(tm), C.byref(te), C.byref(cfg.c),
                                        _ptr(inputs.P), _ptr(inputs.Q), _ptr(inputs.user_bias), _ptr(inputs.item_bias),
                                        float(global_bias), _ptr(inputs.user_ids), _ptr(iptr), inputs.n_train_global,
                                        inputs.n_test_global, inputs.n_active_global, blob))
        self.handle = blob.raw

    def connect(self, handles):
        """handles: list of `world` blobs in rank order."""
        buf = b"".join(handles)
        assert len(buf) == self.world * _lib.DSGD_HANDLE_BYTES
        check(self.lib.cu2b_dsgd_connect(self.h, buf))

    def run(self, n_iterations):
        check(self.lib.cu2b_dsgd_run(self.h, n_iterations))

    def reload(self, inputs, global_bias):
        """Same-shaped strips + a new initial model into the existing context (cu2b_dsgd_reload).
        Synchronise the ranks (a host barrier) before the next run()."""
        self._keep = inputs
        tm, te = inputs.train.c(), inputs.test.c()
        check(self.lib.cu2b_dsgd_reload(self.h, C.byref(tm), C.byref(te), _ptr(inputs.P), _ptr(inputs.Q),
                                        _ptr(inputs.user_bias), _ptr(inputs.item_bias), float(global_bias)))

    def _session(self):
        return C.c_void_p(self.lib.cu2b_dsgd_session(self.h))

    def log(self):
        cap = self.cfg.total_iterations // max(1, self.cfg.check_error) + 8
        buf = (Metrics * cap)()
        n = C.c_int()
        check(self.lib.cu2b_session_log(self._session(), buf, cap, C.byref(n)))
        return _metrics_rows(buf, min(cap, n.value))

    def stats(self, reset=False):
        st = Stats()
        check(self.lib.cu2b_session_stats(self._session(), C.byref(st), int(reset)))
        return _stats_dict(st)

    def local_sums(self):
        out = (C.c_double * 4)()
        check(self.lib.cu2b_dsgd_local_sums(self.h, out))
        return list(out)

    def download(self, out=None):
        """-> (P strip, Q in renumbered item order, user_bias strip, item_bias renumbered)."""
        if out is None:
            out = (np.empty((self.rows, self.k), dtype=np.float32), np.empty((self.cols, self.k), dtype=np.float32),
                   np.empty(self.rows, dtype=np.float32), np.empty(self.cols, dtype=np.float32))
        P, Q, ub, ib = out
        check(self.lib.cu2b_session_download(self._session(), _ptr(P), _ptr(Q), _ptr(ub), _ptr(ib)))
        return P, Q, ub, ib

    def close(self):
        if self.h:
            self.lib.cu2b_dsgd_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def predict_topk(P, Q, user_bias, item_bias, global_bias, topk, exclude=None):
    """Top-k unrated items for every user -> (items [U, topk] int32, scores [U, topk] float32, timing dict)."""
    P, Q, ub, ib = _f32(P), _f32(Q), _f32(user_bias), _f32(item_bias)
    rows, cols = ub.shape[0], ib.shape[0]
    k = P.size // rows
    items = np.empty((rows, topk), dtype=np.int32)
    scores = np.empty((rows, topk), dtype=np.float32)
    ms = (C.c_double * 2)()
    ex = exclude.c() if exclude is not None else None
    check(_lib.load().cu2b_predict_topk(_ptr(P), rows, _ptr(Q), cols, _ptr(ub), _ptr(ib), float(global_bias), k,
                                        C.byref(ex) if ex is not None else None, topk, _ptr(items), _ptr(scores), ms))
    return items, scores, dict(candidates_ms=ms[0], rescore_ms=ms[1])


def device_info(device=0):
    lib = _lib.load()
    name = C.create_string_buffer(256)
    sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
    fb, tb = C.c_int64(), C.c_int64()
    check(lib.cu2b_device_info(device, name, 256, C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(fb), C.byref(tb)))
    return dict(name=name.value.decode(), sm_count=sm.value, cc=(maj.value, mnr.value), free_bytes=fb.value,
                total_bytes=tb.value)


# ---------------------------------------------------------------------------------------------
# preprocessing (host only) -- the reference's preprocessing/ scripts, byte-compatible outputs
# ---------------------------------------------------------------------------------------------
def _mapped_name(path, suffix):
    root, ext = os.path.splitext(path)
    return "%s_%s%s" % (root, suffix, ext)


def map_items(file_ratings, file_out=None, delimiter=",", has_header=True, rating_col=2, second=None, second_out=None):
    """preprocessing/map_items.py (process_file) / map_netflix.py. Returns a dict of counts."""
    lib = _lib.load()
    file_out = file_out or _mapped_name(file_ratings, "mapped")
    v = [C.c_int64() for _ in range(6)]
    check(lib.cu2b_prep_map(os.fsencode(file_ratings), os.fsencode(file_out), delimiter.encode(), int(has_header),
                            rating_col, os.fsencode(second) if second else None,
                            os.fsencode(second_out) if second_out else None, *[C.byref(x) for x in v]))
    keys = ("rows", "rows_second", "users", "items", "skipped_users", "skipped_items")
    return dict(zip(keys, (x.value for x in v)), out=file_out)


def sort_ratings(file_ratings, file_out=None):
    """preprocessing/sort_ratings.py: by userId, then itemId -> <name>_sorted<ext>."""
    lib = _lib.load()
    file_out = file_out or _mapped_name(file_ratings, "sorted")
    n = C.c_int64()
    check(lib.cu2b_prep_sort(os.fsencode(file_ratings), os.fsencode(file_out), C.byref(n)))
    return dict(rows=n.value, out=file_out)


def split_to_test_train(file_ratings, test_ratio, seed=42, train_out=None, test_out=None):
    """preprocessing/split_to_test_train.py (split_true) -> <name>_train<ext>, <name>_test<ext>."""
    lib = _lib.load()
    train_out = train_out or _mapped_name(file_ratings, "train")
    test_out = test_out or _mapped_name(file_ratings, "test")
    a, b = C.c_int64(), C.c_int64()
    check(lib.cu2b_prep_split(os.fsencode(file_ratings), os.fsencode(train_out), os.fsencode(test_out),
                              float(test_ratio), int(seed), C.byref(a), C.byref(b)))
    return dict(train=a.value, test=b.value, train_out=train_out, test_out=test_out)


def create_config(filename, num_iterations=1000, num_factors=100, learning_rate=0.01, seed=42, p_reg=0.02, q_reg=0.02,
                  user_bias_reg=0.02, item_bias_reg=0.02):
    """preprocessing/create_config.py (same defaults)."""
    check(_lib.load().cu2b_prep_create_config(os.fsencode(filename), num_iterations, num_factors, learning_rate, seed,
                                              p_reg, q_reg, user_bias_reg, item_bias_reg))


def convert_to_np(filename_in, filename_out=None):
    """preprocessing/convert_to_np.py: CSV float matrix -> .npy as np.save(np.genfromtxt(..., delimiter=',')) writes it."""
    if filename_out is None:
        filename_out = os.path.splitext(filename_in)[0] + ".npy"
    r, c = C.c_int64(), C.c_int64()
    check(_lib.load().cu2b_prep_convert_to_np(os.fsencode(filename_in), os.fsencode(filename_out), C.byref(r), C.byref(c)))
    return dict(rows=r.value, cols=c.value, out=filename_out)


def write_ratings_csv(path, ratings):
    """Ratings triplets (0-based ids) -> "userId,itemId,rating" CSV with 1-based ids (readCSV's input)."""
    ratings = np.ascontiguousarray(ratings, dtype=RATING_DTYPE)
    check(_lib.load().cu2b_write_ratings_csv(os.fsencode(path), ratings.ctypes.data_as(C.POINTER(Rating)), len(ratings)))
