// cu2rec_shim.h -- header-only C++ layer that gives the C ABI (include/cu2b.h) the names,
// argument meaning, ownership and error behaviour of the reference's in-process interface, so
// that code written against cu2rec's headers (mf.cu, predict.cu, tests/*.cu) ports 1:1.
//
//   reference                                         here
//   config::Config (config.h:20-58)                   config::Config (same fields / methods)
//   Rating (util.h:19-24)                             Rating
//   cu2rec::CudaCSRMatrix (matrix.h:11-19)            cu2rec::CudaCSRMatrix (host-resident view;
//                                                       the device copy lives inside train())
//   readCSV / read_array / writeCSV / writeToFile /   same signatures (util.h:37-49)
//   initialize_normal_array / createSparseMatrix
//   train(...) 8- and 10-argument (training.h:12-15)  same signatures; callee new[]s the outputs
//   CHECK_CUDA -> std::runtime_error (util.h:27-34)   CU2B_CHECK -> std::runtime_error
//
// Everything that computes runs inside libcu2b.so; this file only marshals.
#ifndef CU2REC_SHIM_H_
#define CU2REC_SHIM_H_

#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "cu2b.h"

#define CU2B_CHECK(expr)                                                     \
    do {                                                                     \
        if ((expr) != CU2B_OK) throw std::runtime_error(cu2b_last_error()); \
    } while (0)

namespace config {
class Config : public cu2b_config {
   public:
    Config() { cu2b_config_default(this); }
    bool read_config(const std::string &path) {
        // config.cu:7-13 never checks the stream: an unreadable file silently keeps the defaults
        cu2b_config_read(path.c_str(), this);
        return true;
    }
    bool write_config(const std::string &path) { return cu2b_config_write(path.c_str(), this) == CU2B_OK; }
    // Hyper-parameters travel as kernel arguments; there is no __constant__ mirror to refresh.
    bool set_cuda_variables() { return true; }
    bool get_cuda_variables() { return true; }
    void print_config() {
        char buf[1024];
        cu2b_config_format(this, buf, (int)sizeof buf);
        fputs(buf, stdout);
    }
};
}  // namespace config

struct Rating {
    int userID;
    int itemID;
    float rating;
};
static_assert(sizeof(Rating) == sizeof(cu2b_rating), "Rating must alias cu2b_rating");

namespace cu2rec {
struct CudaCSRMatrix {
    CudaCSRMatrix(int rows_, int cols_, int nonzeros_, const int *indptr_, const int *indices_, const float *data_)
        : rows(rows_), cols(cols_), nonzeros(nonzeros_) {
        indptr = new int[rows + 1];
        indices = new int[nonzeros > 0 ? nonzeros : 1];
        data = new float[nonzeros > 0 ? nonzeros : 1];
        memcpy(indptr, indptr_, sizeof(int) * (rows + 1));
        if (nonzeros > 0) {
            memcpy(indices, indices_, sizeof(int) * nonzeros);
            memcpy(data, data_, sizeof(float) * nonzeros);
        }
    }
    // Allocates without filling: createSparseMatrix() builds the arrays in place.
    CudaCSRMatrix(int rows_, int cols_, int nonzeros_) : rows(rows_), cols(cols_), nonzeros(nonzeros_) {
        indptr = new int[rows + 1];
        indices = new int[nonzeros > 0 ? nonzeros : 1];
        data = new float[nonzeros > 0 ? nonzeros : 1];
    }
    ~CudaCSRMatrix() {
        delete[] indptr;
        delete[] indices;
        delete[] data;
    }
    CudaCSRMatrix(const CudaCSRMatrix &) = delete;
    CudaCSRMatrix &operator=(const CudaCSRMatrix &) = delete;
    cu2b_csr view() const {
        cu2b_csr v;
        v.rows = rows; v.cols = cols; v.nonzeros = nonzeros;
        v.indptr = indptr; v.indices = indices; v.data = data;
        v.on_device = 0;
        return v;
    }
    int *indptr, *indices;
    float *data;
    int rows, cols, nonzeros;
};
}  // namespace cu2rec

inline std::vector<Rating> readCSV(const std::string &filename, int *rows, int *cols, float *global_bias) {
    cu2b_rating *r = nullptr;
    int64_t n = 0;
    std::vector<Rating> out;
    // CU2B_CSV_CACHE=1: keep / use a binary sidecar "<file>.cu2bcache" (same results; a second run skips the parse)
    const char *cache = getenv("CU2B_CSV_CACHE");
    const bool cached = cache && *cache && strcmp(cache, "0") != 0;
    const cu2b_status rc = cached ? cu2b_read_csv_cached(filename.c_str(), nullptr, &r, &n, rows, cols, global_bias, nullptr)
                                  : cu2b_read_csv(filename.c_str(), &r, &n, rows, cols, global_bias);
    if (rc != CU2B_OK) {
        fprintf(stderr, "ERROR: The file isnt open.\n");  // util.cu:42
        return out;
    }
    // one pass (no zero-fill before the copy): the file may hold 10^8 ratings
    const Rating *first = reinterpret_cast<const Rating *>(r);
    if (n) out.assign(first, first + n);
    cu2b_free(r);
    return out;
}

inline float *read_array(const char *file_path, int *n_rows_ptr, int *n_cols_ptr) {
    float *tmp = nullptr;
    int r = 0, c = 0;
    if (cu2b_read_array(file_path, &tmp, &r, &c) != CU2B_OK) return nullptr;  // util.cu:69-71
    float *out = new float[c > 0 ? c : 1];
    if (c) memcpy(out, tmp, sizeof(float) * (size_t)c);
    cu2b_free(tmp);
    *n_rows_ptr = r;
    *n_cols_ptr = c;
    return out;
}
inline float *read_array(const char *file_path) {
    int r, c;
    return read_array(file_path, &r, &c);
}

inline void writeCSV(char *file_path, float *data, int rows, int cols) { CU2B_CHECK(cu2b_write_csv(file_path, data, rows, cols)); }
inline void writeToFile(const std::string &parent_dir, const std::string &base_filename, const std::string &extension,
                        const std::string &component, float *data, int rows, int cols, int factors) {
    CU2B_CHECK(cu2b_write_component(parent_dir.c_str(), base_filename.c_str(), extension.c_str(), component.c_str(), data,
                                    rows, cols, factors));
}

inline float *initialize_normal_array(int size, int n_factors, float mean, float stddev, int seed) {
    float *a = new float[size > 0 ? size : 1];
    cu2b_init_normal(a, size, n_factors, mean, stddev, seed);
    return a;
}
inline float *initialize_normal_array(int size, int n_factors, float mean, float stddev) {
    return initialize_normal_array(size, n_factors, mean, stddev, 42);
}
inline float *initialize_normal_array(int size, int n_factors, int seed) { return initialize_normal_array(size, n_factors, 0, 1, seed); }
inline float *initialize_normal_array(int size, int n_factors) { return initialize_normal_array(size, n_factors, 0, 1); }

inline cu2rec::CudaCSRMatrix *createSparseMatrix(std::vector<Rating> *ratings, int rows, int cols) {
    const int64_t n = (int64_t)ratings->size();
    if (n > INT32_MAX) throw std::runtime_error("createSparseMatrix: more than 2^31-1 ratings");
    cu2rec::CudaCSRMatrix *m = new cu2rec::CudaCSRMatrix(rows, cols, (int)n);
    if (cu2b_build_csr(reinterpret_cast<const cu2b_rating *>(ratings->data()), n, rows, m->indptr, m->indices, m->data) !=
        CU2B_OK) {
        delete m;
        throw std::runtime_error(cu2b_last_error());
    }
    return m;
}

inline size_t getFreeBytes(const int where, size_t *total_bytes) {
    int64_t f = 0, t = 0;
    if (cu2b_device_info(0, nullptr, 0, nullptr, nullptr, nullptr, &f, &t) != CU2B_OK) {
        printf("getFreeBytes: call index %d: cudaMemGetInfo returned the error: %s\n", where, cu2b_last_error());
        exit(1);  // util.cu:188-193
    }
    *total_bytes = (size_t)t;
    return (size_t)f;
}

namespace cu2rec_detail {
// training.cu:21-204 surface: allocates the outputs with new[], trains, prints the reference's
// progress lines (emitted after the device-resident loop from its metric log).
inline void train_impl(cu2rec::CudaCSRMatrix *train_matrix, cu2rec::CudaCSRMatrix *test_matrix, config::Config *cfg,
                       float **P_ptr, float *Q, float **losses_ptr, float **user_bias_ptr, float *item_bias,
                       float global_bias, int init_item_side) {
    const int user_count = train_matrix->rows, k = cfg->n_factors;
    const int total = cfg->total_iterations;
    float *P = new float[(size_t)(user_count > 0 ? user_count : 1) * k];
    float *losses = new float[total > 0 ? total : 1];
    float *user_bias = new float[user_count > 0 ? user_count : 1];
    *P_ptr = P;
    *losses_ptr = losses;
    *user_bias_ptr = user_bias;
    const int cap = total / (cfg->check_error > 0 ? cfg->check_error : 1) + 8;
    std::vector<cu2b_metrics> log((size_t)cap);
    int n_log = 0;
    cu2b_stats st;
    memset(&st, 0, sizeof st);
    const float lr0 = cfg->learning_rate;
    cu2b_csr tr = train_matrix->view(), te = test_matrix->view();
    CU2B_CHECK(cu2b_train(&tr, &te, cfg, P, Q, user_bias, item_bias, global_bias, init_item_side, losses, log.data(), cap,
                          &n_log, &st));
    float lr_prev = lr0;
    for (int r = 0; r < n_log && r < cap; ++r) {
        printf("TRAIN: Iteration %d GPU MAE: %f RMSE: %f\n", log[r].iteration, log[r].train_mae, log[r].train_rmse);
        printf("TEST: Iteration %d GPU MAE: %f RMSE: %f\n", log[r].iteration, log[r].test_mae, log[r].test_rmse);
        if (log[r].learning_rate != lr_prev) {
            printf("New Learning Rate: %f\n: ", log[r].learning_rate);  // training.cu:154
            lr_prev = log[r].learning_rate;
        }
    }
    printf("Time taken for %d of iterations is %lf\n", total, st.total_ms / 1e3);  // training.cu:177
    printf("cu2b: %lld rating updates, %.3f G updates/s in the SGD kernels, %lld kernel launches\n",
           (long long)st.updates, st.sgd_ms > 0 ? st.updates / st.sgd_ms / 1e6 : 0.0, (long long)st.kernel_launches);
}
}  // namespace cu2rec_detail

// training.h:12-13 (10-argument form: Q and item_bias are inputs, updated in place)
inline void train(cu2rec::CudaCSRMatrix *train_matrix, cu2rec::CudaCSRMatrix *test_matrix, config::Config *cfg,
                  float **P_ptr, float **Q_ptr, float *Q, float **losses_ptr, float **user_bias_ptr,
                  float **item_bias_ptr, float *item_bias, float global_bias) {
    (void)Q_ptr;
    (void)item_bias_ptr;
    cu2rec_detail::train_impl(train_matrix, test_matrix, cfg, P_ptr, Q, losses_ptr, user_bias_ptr, item_bias, global_bias, 0);
}

// training.h:14-15 (8-argument form: Q and item_bias are initialised here, training.cu:208-217)
inline void train(cu2rec::CudaCSRMatrix *train_matrix, cu2rec::CudaCSRMatrix *test_matrix, config::Config *cfg,
                  float **P_ptr, float **Q_ptr, float **losses_ptr, float **user_bias_ptr, float **item_bias_ptr,
                  float global_bias) {
    const int item_count = train_matrix->cols, k = cfg->n_factors;
    float *Q = new float[(size_t)(item_count > 0 ? item_count : 1) * k];
    float *item_bias = new float[item_count > 0 ? item_count : 1];
    *Q_ptr = Q;
    *item_bias_ptr = item_bias;
    cu2rec_detail::train_impl(train_matrix, test_matrix, cfg, P_ptr, Q, losses_ptr, user_bias_ptr, item_bias, global_bias, 1);
}

#endif  // CU2REC_SHIM_H_
