/* cu2b.h -- C ABI of libcu2b.so: a B200-native (sm_100a) drop-in for the SGD
 * matrix-factorisation training path of nickgreenquist/cu2rec.
 *
 * Plain C: pointers and sizes only, integer status codes, no exceptions and no
 * allocation ownership crossing the boundary except through cu2b_free(). Each entry
 * point cites the reference interface it replaces as file:line relative to the
 * reference repository's matrix_factorization/ directory.
 *
 * Conventions
 *   - every function returning cu2b_status returns CU2B_OK (0) on success; on failure
 *     cu2b_last_error() returns a thread-local human readable message
 *     (reference: CHECK_CUDA throws std::runtime_error, util.h:27-34);
 *   - "host" pointers are ordinary process memory; entry points ending in _dev take
 *     device pointers on the current CUDA device;
 *   - factor matrices are dense row-major [rows x n_factors] float32 (matrix.h:20-28),
 *     ids are 0-based int32 (util.cu:31), ratings float32;
 *   - there is NO CPU fallback: if the CUDA device or the sm_100a image is unavailable the
 *     call fails with CU2B_ERR_CUDA.
 */
#ifndef CU2B_H_
#define CU2B_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CU2B_VERSION 100

typedef enum cu2b_status {
    CU2B_OK = 0,
    CU2B_ERR_INVALID = 1,     /* bad argument */
    CU2B_ERR_IO = 2,          /* file could not be opened / parsed */
    CU2B_ERR_CUDA = 3,        /* CUDA runtime error, no device, wrong architecture */
    CU2B_ERR_NOMEM = 4,
    CU2B_ERR_UNSUPPORTED = 5, /* e.g. n_factors > 512 */
    CU2B_ERR_DIVERGED = 6     /* a loss check saw a non-finite RMSE: the run's results are void (no
                                 reference counterpart: training.cu:134-158 prints nan and carries on) */
} cu2b_status;

const char *cu2b_last_error(void);
int cu2b_version(void);
void cu2b_free(void *p); /* releases memory returned by cu2b_read_csv / cu2b_read_array */

/* ------------------------------------------------------------------------------------
 * Hyper-parameters. Mirrors config::Config (config.h:20-58) field for field; the
 * __constant__ mirror (config.h:9-18, config.cu:24-48) is replaced by kernel parameters.
 * Fields after learning_rate_decay are extensions with reference-neutral defaults.
 * ---------------------------------------------------------------------------------- */
enum { CU2B_MODE_HOGWILD = 0, CU2B_MODE_DETERMINISTIC = 1 };
enum { CU2B_SAMPLER_PER_USER = 0, CU2B_SAMPLER_PER_RATING = 1 };

typedef struct cu2b_config {
    int cur_iterations;        /* config.h:23 */
    int total_iterations;      /* config.h:25  default 5000 */
    int n_factors;             /* config.h:27  default 50 */
    float learning_rate;       /* config.h:29  default 0.01 */
    int seed;                  /* config.h:31  default 42 (sampler seed, sgd.cu:11-16) */
    float P_reg;               /* config.h:33  default 0.02 */
    float Q_reg;               /* config.h:35 */
    float user_bias_reg;       /* config.h:37 */
    float item_bias_reg;       /* config.h:39 */
    int is_train;              /* config.h:41  0 => Q and item_bias frozen (predict.cu:105) */
    int n_threads;             /* config.h:43  kept for file/print compatibility; unused */
    int check_error;           /* config.h:45  default 500 */
    float patience;            /* config.h:48  default 2 */
    float learning_rate_decay; /* config.h:51  default 0.2 */
    /* extensions */
    int mode;     /* CU2B_MODE_*: Hogwild (default) or deterministic conflict-free blocks */
    int sampler;  /* CU2B_SAMPLER_*: per_user = the reference's one-rating-per-user-per-
                     iteration distribution (sgd.cu:27-37, default; the parity mode); per_rating =
                     shuffled passes over the rating list, an iteration still being one update per
                     active user (single GPU, iteration-synchronous kernel; it converges to a LOWER
                     test RMSE than the reference's distribution at equal updates -- SURVEY 7.1 --
                     so it is offered, not compared two-sidedly) */
    int n_blocks; /* deterministic mode: B of the BxB block grid, 0 = automatic */
    int n_gpus;   /* DSGD width, 0/1 = single GPU */
    int round_iters; /* Hogwild schedule on one GPU: this many consecutive iterations of a user are
                        applied back to back with the user's row kept in registers (default 32;
                        1 = the reference's iteration-synchronous order) */
} cu2b_config;

void cu2b_config_default(cu2b_config *cfg);
/* config.cu:7-13: nine whitespace separated positional tokens; a short file keeps defaults.
 * Optional tokens 10.. (n_threads patience learning_rate_decay check_error mode sampler
 * n_blocks n_gpus round_iters) are an add-only extension (config.h TODO / create_config.py:16-17). */
cu2b_status cu2b_config_read(const char *path, cu2b_config *cfg);
/* config.cu:15-22: writes the nine reference tokens on one line. */
cu2b_status cu2b_config_write(const char *path, const cu2b_config *cfg);
/* config.cu:50-64 print_config: formats the same block of lines into buf; returns the
 * length that was (or would have been) written. */
int cu2b_config_format(const cu2b_config *cfg, char *buf, int cap);

/* ------------------------------------------------------------------------------------
 * Ratings IO (host side).
 * ---------------------------------------------------------------------------------- */
typedef struct cu2b_rating { /* util.h:19-24 struct Rating; also the device stream element */
    int32_t user;
    int32_t item;
    float rating;
} cu2b_rating;

/* util.cu:17-45 readCSV: header line skipped, rows "int <c> int <c> float", 1-based ids stored
 * 0-based, *rows = max userId, *cols = max itemId, *global_bias = mean rating (double sum).
 * *ratings is allocated by the library (cu2b_free). */
cu2b_status cu2b_read_csv(const char *path, cu2b_rating **ratings, int64_t *n, int *rows,
                          int *cols, float *global_bias);
/* The same read through a binary sidecar (no reference counterpart: util.cu:17-45 re-parses the text on
 * every run). `cache_path` NULL = "<path>.cu2bcache". If that file is a sidecar written for exactly this state of
 * `path` (same size and modification time), its triplets, dimensions and mean are returned without parsing and
 * *hit = 1; otherwise the CSV is parsed as by cu2b_read_csv and the sidecar is (re)written, best effort and
 * atomically (temporary file + rename). The results are identical either way. `hit` may be NULL. */
cu2b_status cu2b_read_csv_cached(const char *path, const char *cache_path, cu2b_rating **ratings, int64_t *n, int *rows,
                                 int *cols, float *global_bias, int *hit);
/* util.cu:152-179 createSparseMatrix (host part): ratings grouped by ascending user -> CSR;
 * missing users repeat indptr. indptr has rows+1 entries. */
cu2b_status cu2b_build_csr(const cu2b_rating *ratings, int64_t n, int rows, int *indptr,
                           int *indices, float *data);
/* util.cu:52-76 read_array: comma separated float matrix; *n_cols accumulates over all rows
 * exactly like the reference (util.cu:61-66), *n_elems is the element count. */
cu2b_status cu2b_read_array(const char *path, float **data, int *n_rows, int *n_cols);
/* util.cu:86-97 writeCSV: "%f" text, comma separated, one matrix row per line. */
cu2b_status cu2b_write_csv(const char *path, const float *data, int rows, int cols);
/* util.cu:99-103 writeToFile: <parent>/<base>_f<factors>_<component>.<ext>. */
cu2b_status cu2b_write_component(const char *parent_dir, const char *base, const char *ext,
                                 const char *component, const float *data, int rows, int cols,
                                 int factors);
/* util.cu:124-144 initialize_normal_array: mt19937(seed), N(mean, stddev / n_factors). */
void cu2b_init_normal(float *out, int64_t size, int n_factors, float mean, float stddev,
                      int seed);

/* Synthetic low-rank-plus-noise ratings of a given shape; used by the
 * tests, smoke and bench only. Rows come out grouped by ascending user, ids 0-based, every
 * user has >= 1 train rating. Two calls: with train == NULL it only reports the counts. */
cu2b_status cu2b_synth_ratings(int users, int items, int64_t target_ratings, int rank,
                               float noise, int integer_ratings, float test_fraction,
                               uint64_t seed, cu2b_rating *train, int64_t *n_train,
                               cu2b_rating *test, int64_t *n_test);

/* ------------------------------------------------------------------------------------
 * CSR view (matrix.h:11-19 CudaCSRMatrix). on_device selects host or device pointers.
 * ---------------------------------------------------------------------------------- */
typedef struct cu2b_csr {
    int rows, cols, nonzeros;
    const int *indptr;  /* rows + 1 */
    const int *indices; /* nonzeros */
    const float *data;  /* nonzeros */
    int on_device;
} cu2b_csr;

/* ------------------------------------------------------------------------------------
 * Kernel-level entry points, host buffers in / host buffers out (copies inside).
 * ---------------------------------------------------------------------------------- */
/* loss.cu:40-49 calculate_loss_gpu + loss.cu:196-200 get_error_metrics_gpu in ONE pass:
 * mae = sum|err|/n, rmse = sqrt(sum err^2/n), double accumulation, float results. */
cu2b_status cu2b_loss(const cu2b_csr *m, const float *P, const float *Q, const float *user_bias,
                      const float *item_bias, float global_bias, int n_factors, float *mae,
                      float *rmse);
/* loss.cu:40-49 calculate_loss_gpu: err[i] = data[i] - prediction(i). */
cu2b_status cu2b_residuals(const cu2b_csr *m, const float *P, const float *Q,
                           const float *user_bias, const float *item_bias, float global_bias,
                           int n_factors, float *err);
/* loss.cu:196-200 get_error_metrics_gpu on an explicit error vector. */
cu2b_status cu2b_error_metrics(const float *err, int64_t n, float *mae, float *rmse);
/* sgd.cu:27-37 sampling step only: for iterations [iter0, iter0 + n_iter) one uniformly drawn
 * rating per user that has any, ascending user order. out must hold n_iter * (active users). */
cu2b_status cu2b_sample_per_user(const cu2b_csr *m, int seed, int iter0, int n_iter,
                                 cu2b_rating *out, int64_t *n_out);
/* The per_rating sampler's stream (CU2B_SAMPLER_PER_RATING; no reference counterpart): updates
 * [first_update, first_update + n_updates) of a run = a shuffled pass over the rating list, pass after pass
 * (update q applies rating perm_e(q mod nnz) of pass e = q / nnz; perm_e is a keyed bijection of [0, nnz)). */
cu2b_status cu2b_sample_per_rating(const cu2b_csr *m, int seed, int64_t first_update, int64_t n_updates,
                                   cu2b_rating *out);
/* sgd.cu:40-72 update arithmetic (in place Q / item_bias as mf_sequential.cu:114-141) applied
 * to an explicit stream of ratings. order: 0 = Hogwild (parallel, racy by design),
 * 1 = strictly sequential in stream order (one update in flight; for bit-exact checks). */
cu2b_status cu2b_sgd_apply(const cu2b_rating *stream, int64_t n, float *P, int rows, float *Q,
                           int cols, float *user_bias, float *item_bias, float global_bias,
                           const cu2b_config *cfg, int order);
/* Deterministic conflict-free mode: one pass over the ratings in the block-diagonal schedule
 * (B x B blocks, round s runs blocks (b,(b+s) mod B) concurrently, each block serially). The
 * result equals a sequential replay in round-major / user-block / original order. */
cu2b_status cu2b_sgd_blocked(const cu2b_rating *coo, int64_t n, float *P, int rows, float *Q,
                             int cols, float *user_bias, float *item_bias, float global_bias,
                             const cu2b_config *cfg, int n_blocks, int n_passes);

/* Host-only helper: the canonical sequential order of that schedule (a permutation of [0,n));
 * n_blocks == 0 selects the automatic B, reported through n_blocks_used. No GPU needed. */
cu2b_status cu2b_block_schedule_order(const cu2b_rating *coo, int64_t n, int rows, int cols,
                                      int n_blocks, int64_t *order, int *n_blocks_used);

/* ------------------------------------------------------------------------------------
 * Training (training.h:12-15 train()).
 * ---------------------------------------------------------------------------------- */
typedef struct cu2b_metrics { /* one row per loss check (training.cu:118-158) */
    int iteration;            /* 1-based, as printed by training.cu:135 */
    float train_mae, train_rmse, test_mae, test_rmse;
    float learning_rate;      /* after the patience / decay step of this check */
} cu2b_metrics;

typedef struct cu2b_stats {
    double sgd_ms;          /* device time inside SGD kernels (CUDA events) */
    double loss_ms;         /* device time inside loss kernels */
    double sampler_ms;      /* device time inside sampler kernels */
    double total_ms;        /* device time of the whole enqueued loop */
    int64_t updates;        /* rating updates performed */
    int64_t kernel_launches;
    int64_t sgd_launches;
    double wait_ms;         /* DSGD: device time spent waiting for a peer's item block / loss sums */
    double send_ms;         /* DSGD: device time of the peer-memory hand-off kernels */
} cu2b_stats;

/* One call = the reference's train(): uploads the matrices, trains cfg->total_iterations
 * iterations with the loss cadence / patience / decay of training.cu:107-170 entirely on the
 * device (no per-iteration host round trip), downloads the result.
 *   init_item_side != 0 : training.cu:208-217 (8-argument overload) -- Q and item_bias are
 *                         initialised here (N(0,1/k), seed 42);
 *   init_item_side == 0 : training.cu:21 (10-argument overload) -- Q / item_bias are inputs.
 * P [rows x k] and user_bias [rows] are always initialised inside (training.cu:28,54).
 * cfg is in-out (learning_rate, cur_iterations; training.cu:151,170). losses may be NULL; else
 * it has total_iterations floats: validation RMSE at check iterations, NaN elsewhere
 * (training.cu:158). log may be NULL. */
cu2b_status cu2b_train(const cu2b_csr *train, const cu2b_csr *test, cu2b_config *cfg, float *P,
                       float *Q, float *user_bias, float *item_bias, float global_bias,
                       int init_item_side, float *losses, cu2b_metrics *log, int log_cap,
                       int *n_log, cu2b_stats *stats);

/* Resident-data session: the same loop with explicit lifetime, for callers that keep the
 * model on the device between calls (bench, multi-segment training). */
typedef struct cu2b_session cu2b_session;
cu2b_status cu2b_session_create(cu2b_session **out, int device, const cu2b_csr *train,
                                const cu2b_csr *test, const cu2b_config *cfg, const float *P,
                                const float *Q, const float *user_bias, const float *item_bias,
                                float global_bias);
/* Runs n_iterations more iterations (continuing cur_iterations); blocks until done. The
 * "last iteration" check of training.cu:118 fires at cfg.total_iterations. */
cu2b_status cu2b_session_run(cu2b_session *s, int n_iterations);
/* cu2b_session_run followed by cu2b_session_download as ONE call (what train() does at its end, training.cu:180-185):
 * when the run ends with a loss check -- it does whenever it reaches total_iterations -- the device->host copies
 * of the final model run on a second stream while that check evaluates it. Any output pointer may be NULL. */
cu2b_status cu2b_session_run_download(cu2b_session *s, int n_iterations, float *P, float *Q, float *user_bias,
                                      float *item_bias);
cu2b_status cu2b_session_eval(cu2b_session *s, float *train_mae, float *train_rmse,
                              float *test_mae, float *test_rmse);
cu2b_status cu2b_session_log(cu2b_session *s, cu2b_metrics *out, int cap, int *n);
cu2b_status cu2b_session_download(cu2b_session *s, float *P, float *Q, float *user_bias,
                                  float *item_bias);
cu2b_status cu2b_session_get_config(cu2b_session *s, cu2b_config *out);
cu2b_status cu2b_session_stats(cu2b_session *s, cu2b_stats *out, int reset);
/* Starts over in an existing session: uploads a problem of the SAME shape (rows, cols, rating
 * counts, number of users with ratings) and a new initial model into the buffers the session
 * already owns, and resets the schedule state (learning rate, patience, log, iteration counter,
 * statistics) to the configuration the session was created with. A repeated train() on
 * same-shaped data (hyper-parameter sweeps, retraining on a refreshed snapshot) then pays the
 * host->device copies but no allocation or set-up. Hogwild mode only. The item draw weights are recomputed
 * on the device from the reloaded matrix and the bound on concurrently applied updates that keeps
 * asynchronous SGD stable (it depends on the most popular item's share of the draws and on the learning
 * rate) is re-derived from them, as are a DSGD rank's per-item step fractions; only the internal item
 * placement (a performance heuristic) stays the one chosen at creation. */
cu2b_status cu2b_session_reload(cu2b_session *s, const cu2b_csr *train, const cu2b_csr *test,
                                const float *P, const float *Q, const float *user_bias,
                                const float *item_bias, float global_bias);
void cu2b_session_destroy(cu2b_session *s);

/* ------------------------------------------------------------------------------------
 * Multi-GPU: DSGD 2-D stratification on one NVLink/NVSwitch box (no reference counterpart;
 * cu2rec is single GPU). Users and items are cut into `world` blocks; rank g owns P / user_bias
 * of user block g and that block's ratings; sub-epoch s of a round runs rating block
 * (g, (g+s) mod world) on rank g; item blocks (Q rows + item_bias) rotate g -> g-1 through peer
 * memory (P2P stores over NVLink + a release flag), never through the host.
 * ---------------------------------------------------------------------------------- */
/* Host only. LPT-balanced block assignment by rating count. user_block[u] / user_local[u]: block
 * and index inside the block of original user u; item_new[i]: renumbered item id such that item
 * block b is the contiguous range [item_block_ptr[b], item_block_ptr[b+1]). block_nnz (optional)
 * receives the world x world rating counts. */
cu2b_status cu2b_dsgd_partition(const cu2b_rating *train, int64_t n, int rows, int cols, int world,
                                int *user_block, int *user_local, int *users_per_block,
                                int *item_new, int *item_block_ptr, int64_t *block_nnz);
/* Host only. The ratings of one rank (local user ids, renumbered items, original order kept).
 * out == NULL only counts. */
cu2b_status cu2b_dsgd_extract_strip(const cu2b_rating *ratings, int64_t n, const int *user_block,
                                    const int *user_local, const int *item_new, int rank,
                                    cu2b_rating *out, int64_t *n_out);

/* Host only. Item-step thinning fractions (a DSGD rank applies them to the bias steps of the
 * popular items by default, CU2B_DSGD_THIN_BIAS, and to the row steps on request, CU2B_DSGD_THIN): keep[i] in (0, 1] such that
 * lr x (item i's share of the draws of its item block under per-user sampling) x groups_in_flight x
 * keep[i] <= budget. Items under the budget keep 1. item_block_ptr == NULL: one block. */
cu2b_status cu2b_dsgd_item_keep(const cu2b_csr *train_strip, const int *item_block_ptr, int world,
                                float learning_rate, int groups_in_flight, double budget, float *keep);

#define CU2B_DSGD_HANDLE_BYTES 512
typedef struct cu2b_dsgd cu2b_dsgd;
/* One context per rank (one process per GPU, or several contexts in one process). The strips use
 * local user ids and renumbered item ids; Q / item_bias are full size in the renumbered order;
 * user_ids[rows] are the ORIGINAL user ids of the strip's users (they key the per-user sampler so
 * that every rank draws exactly what a single GPU would draw). handle_out receives an opaque
 * blob (CU2B_DSGD_HANDLE_BYTES) to be exchanged between ranks by any host transport. */
cu2b_status cu2b_dsgd_create(cu2b_dsgd **out, int device, int rank, int world,
                             const cu2b_csr *train_strip, const cu2b_csr *test_strip,
                             const cu2b_config *cfg, const float *P_strip, const float *Q,
                             const float *user_bias_strip, const float *item_bias, float global_bias,
                             const int *user_ids, const int *item_block_ptr, int64_t n_train_global,
                             int64_t n_test_global, int64_t n_active_global, void *handle_out);
/* handles: world blobs in rank order (own blob included). Maps the peers' buffers (CUDA IPC
 * across processes, direct peer access inside one process). */
cu2b_status cu2b_dsgd_connect(cu2b_dsgd *d, const void *handles);
/* Collective: every rank calls it with the same n_iterations. Blocks until this rank is done. */
cu2b_status cu2b_dsgd_run(cu2b_dsgd *d, int n_iterations);
/* The rank-local session (log / stats / download of the strip / config); owned by the context. */
cu2b_session *cu2b_dsgd_session(cu2b_dsgd *d);
/* This rank's latest local loss sums {train sse, train sae, test sse, test sae} (for
 * cross-checking the peer-memory combine against an NCCL / gloo all-reduce). */
cu2b_status cu2b_dsgd_local_sums(cu2b_dsgd *d, double out[4]);
/* cu2b_session_reload for a DSGD rank: same-shaped strips and a new initial model into the
 * existing context; peer mappings, flags and buffers stay. Every rank calls it, and the ranks
 * must synchronise (any host barrier) between their reloads and the next cu2b_dsgd_run, because a
 * running peer writes item blocks into this rank's Q. */
cu2b_status cu2b_dsgd_reload(cu2b_dsgd *d, const cu2b_csr *train_strip, const cu2b_csr *test_strip,
                             const float *P_strip, const float *Q, const float *user_bias_strip,
                             const float *item_bias, float global_bias);
void cu2b_dsgd_destroy(cu2b_dsgd *d);

/* ------------------------------------------------------------------------------------
 * Batched predict + top-k (predict.cu:17-29 predict_ratings, :49-63 get_recommendations, for ALL
 * users at once): for every user the topk items with the highest predicted rating
 * mu + b_u + b_i + p_u.q_i among the items NOT present in `exclude` (may be NULL). Candidate
 * generation runs on the tensor cores (tcgen05, TF32); the returned scores are exact fp32 in the
 * reference's summation order and the order is (score descending, item ascending). Slots that
 * cannot be filled hold item -1 / score NaN. n_factors in [1, 512] (rows are zero-padded on the device
 * to whole 128-byte rows; the reference's default 50, config.h:27, and the 300 of
 * experiments/cu2rec.sh:10 included), topk <= 128 (16 per pass over the catalogue). ms_out (optional)
 * receives {candidate kernels ms, rescore + ordering kernels ms}.
 * ---------------------------------------------------------------------------------- */
cu2b_status cu2b_predict_topk(const float *P, int rows, const float *Q, int cols, const float *user_bias,
                              const float *item_bias, float global_bias, int n_factors,
                              const cu2b_csr *exclude, int topk, int32_t *out_items, float *out_scores,
                              double *ms_out);

/* Sessions allocate from a stream-ordered memory pool that the library owns (one per device) and keeps warm
 * between calls, so that repeated train() calls do not pay for allocating several GB each time (the
 * reference allocates with cudaMalloc per call, matrix.cu:12-40, training.cu:34-88). This returns the cached
 * device memory to the driver; live sessions are not affected. CU2B_NO_MEMPOOL=1 in the environment makes
 * sessions use plain cudaMalloc / cudaFree instead. */
cu2b_status cu2b_release_cache(void);

/* Device introspection used by bench / CLI ("Free memory: %ld", mf.cu:35-37). */
cu2b_status cu2b_device_info(int device, char *name, int name_cap, int *sm_count,
                             int *cc_major, int *cc_minor, int64_t *free_bytes,
                             int64_t *total_bytes);

/* ------------------------------------------------------------------------------------
 * Preprocessing (host only): native counterparts of the reference's preprocessing/ scripts,
 * byte-compatible with their output files.
 * ------------------------------------------------------------------------------------ */
/* map_items.py:21-96 (and map_netflix.py:9-27 when in2_path / out2_path are given): user and item
 * ids -> 1-based sequential ids in first-appearance order, rows grouped by ascending user with the
 * input order kept inside a user, written as "userId,itemId,rating". rating_col = 0-based field of
 * the rating (2; 3 for the Netflix text files whose rating follows two spaces). The second file is
 * mapped with the first file's tables; its rows with unknown users / items are skipped. */
cu2b_status cu2b_prep_map(const char *in_path, const char *out_path, char delimiter, int has_header,
                          int rating_col, const char *in2_path, const char *out2_path,
                          int64_t *n_rows, int64_t *n_rows2, int64_t *n_users, int64_t *n_items,
                          int64_t *skipped_users, int64_t *skipped_items);
/* sort_ratings.py:29-37: by userId, then itemId (stable). */
cu2b_status cu2b_prep_sort(const char *in_path, const char *out_path, int64_t *n_rows);
/* split_to_test_train.py:39-49,71-82 (split_true): one shuffle of all rows with the stream of
 * Python's random.seed(seed) / random.shuffle, the first int(n * (1 - test_ratio)) rows are the
 * training set, both parts stably sorted by user. */
cu2b_status cu2b_prep_split(const char *in_path, const char *train_path, const char *test_path,
                            double test_ratio, int64_t seed, int64_t *n_train, int64_t *n_test);
/* Ratings triplets (0-based ids) -> the reference's input CSV (header, 1-based ids; util.cu:17-45). */
cu2b_status cu2b_write_ratings_csv(const char *path, const cu2b_rating *ratings, int64_t n);
/* create_config.py:10-19: "0 <iterations> <factors> <lr> <seed> <p_reg> <q_reg> <ub_reg> <ib_reg>". */
cu2b_status cu2b_prep_create_config(const char *path, int num_iterations, int num_factors,
                                    double learning_rate, int seed, double p_reg, double q_reg,
                                    double user_bias_reg, double item_bias_reg);
/* convert_to_np.py:6-13: a float matrix in CSV form -> <out>.npy exactly as
 * np.save(out, np.genfromtxt(in, delimiter=',')) writes it (float64, squeezed shape). */
cu2b_status cu2b_prep_convert_to_np(const char *in_path, const char *out_path, int64_t *n_rows,
                                    int64_t *n_cols);

#ifdef __cplusplus
}
#endif
#endif /* CU2B_H_ */
