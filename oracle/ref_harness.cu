// oracle/ref_harness.cu -- TEST INFRASTRUCTURE. Our own driver around the UNMODIFIED reference.
//
// oracle/Makefile concatenates the reference's sources from where they lie under
// /root/reference/matrix_factorization (config.h matrix.h util.h loss.h sgd.h training.h
// config.cu matrix.cu util.cu loss.cu sgd.cu training.cu -- the same single-translation-unit
// recipe as the reference's own makefile:8) followed by this file, and compiles the result to
// oracle/_ref/ref_harness. Nothing of the reference is copied into this repository; this file
// only CALLS the reference's public functions so that tests can obtain outputs of the
// reference itself on chosen inputs (golden-vector generation on CPU here, GPU oracles on the
// B200 box).
//
// Sub-commands (raw little-endian binary files, layouts documented per command):
//   init_normal <size> <k> <out.bin>           util.cu:124-144   (CPU)
//   read_csv <ratings.csv> <out.bin>           util.cu:17-45     (CPU)
//   read_config <file.cfg>                     config.cu:7-13    (CPU, prints the 9 fields)
//   write_csv <in.bin> <rows> <cols> <out.csv> util.cu:86-97     (CPU)
//   read_array <file.csv> <out.bin>            util.cu:52-76     (CPU)
//   loss_gpu <in.bin> <out.bin>                loss.cu:40-49,196-200 (GPU)
//   sgd_gpu <in.bin> <out.bin>                 sgd.cu:11-16,22-75 as launched by training.cu:88,110 (GPU)
//   total_loss <n> <grid> <block>              loss.cu:196-200 on an all-ones vector (GPU)

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static std::vector<char> slurp(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf(n);
    if (fread(buf.data(), 1, n, f) != (size_t)n) { fprintf(stderr, "short read\n"); exit(2); }
    fclose(f);
    return buf;
}

struct Cursor {
    const char *p;
    template <typename T> T get() { T v; memcpy(&v, p, sizeof(T)); p += sizeof(T); return v; }
    template <typename T> const T *arr(size_t n) { const T *r = (const T *)p; p += n * sizeof(T); return r; }
};

static void dump(FILE *f, const void *p, size_t bytes) {
    if (fwrite(p, 1, bytes, f) != bytes) { fprintf(stderr, "short write\n"); exit(2); }
}

// in.bin: int rows, cols, nnz, k; float mu; int indptr[rows+1]; int indices[nnz]; float data[nnz];
//         float P[rows*k]; float Q[cols*k]; float ub[rows]; float ib[cols]
// out.bin: float err[nnz]; float mae; float rmse
static int cmd_loss_gpu(const char *in, const char *out) {
    std::vector<char> buf = slurp(in);
    Cursor c{buf.data()};
    int rows = c.get<int>(), cols = c.get<int>(), nnz = c.get<int>(), k = c.get<int>();
    float mu = c.get<float>();
    const int *indptr = c.arr<int>(rows + 1);
    const int *indices = c.arr<int>(nnz);
    const float *data = c.arr<float>(nnz);
    const float *P = c.arr<float>((size_t)rows * k);
    const float *Q = c.arr<float>((size_t)cols * k);
    const float *ub = c.arr<float>(rows);
    const float *ib = c.arr<float>(cols);

    cu2rec::CudaCSRMatrix matrix(rows, cols, nnz, indptr, indices, data);
    cu2rec::CudaDenseMatrix P_d(rows, k, P), Q_d(cols, k, Q);
    config::Config cfg;
    cfg.n_factors = k;
    float *err_d, *ub_d, *ib_d;
    CHECK_CUDA(cudaMalloc(&err_d, nnz * sizeof(float)));
    CHECK_CUDA(cudaMalloc(&ub_d, rows * sizeof(float)));
    CHECK_CUDA(cudaMalloc(&ib_d, cols * sizeof(float)));
    CHECK_CUDA(cudaMemcpy(ub_d, ub, rows * sizeof(float), cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(ib_d, ib, cols * sizeof(float), cudaMemcpyHostToDevice));
    calculate_loss_gpu(&P_d, &Q_d, &cfg, rows, cols, nnz, &matrix, err_d, ub_d, ib_d, mu);
    const int grid = 256, block = 2 * cfg.n_threads;  // training.cu:75-76
    double *blk_d;
    std::vector<double> blk_h(grid);
    CHECK_CUDA(cudaMalloc(&blk_d, grid * sizeof(double)));
    float mae, rmse;
    std::tie(mae, rmse) = get_error_metrics_gpu(err_d, blk_d, blk_h.data(), nnz, grid, block);
    std::vector<float> err(nnz);
    CHECK_CUDA(cudaMemcpy(err.data(), err_d, nnz * sizeof(float), cudaMemcpyDeviceToHost));
    FILE *f = fopen(out, "wb");
    dump(f, err.data(), nnz * sizeof(float));
    dump(f, &mae, sizeof(float));
    dump(f, &rmse, sizeof(float));
    fclose(f);
    printf("mae %.9g rmse %.9g\n", mae, rmse);
    cudaFree(err_d); cudaFree(ub_d); cudaFree(ib_d); cudaFree(blk_d);
    return 0;
}

// in.bin: as loss_gpu, followed by float lr, P_reg, Q_reg, ub_reg, ib_reg; int seed; int start_user
// One initCurand + one sgd_update launch with the grid/block of training.cu:73-74, then the
// Q / item_bias swap of training.cu:164-165.
// out.bin: float P[rows*k]; float Q[cols*k]; float ub[rows]; float ib[cols]
static int cmd_sgd_gpu(const char *in, const char *out) {
    std::vector<char> buf = slurp(in);
    Cursor c{buf.data()};
    int rows = c.get<int>(), cols = c.get<int>(), nnz = c.get<int>(), k = c.get<int>();
    float mu = c.get<float>();
    const int *indptr = c.arr<int>(rows + 1);
    const int *indices = c.arr<int>(nnz);
    const float *data = c.arr<float>(nnz);
    const float *P = c.arr<float>((size_t)rows * k);
    const float *Q = c.arr<float>((size_t)cols * k);
    const float *ub = c.arr<float>(rows);
    const float *ib = c.arr<float>(cols);
    config::Config cfg;
    cfg.n_factors = k;
    cfg.learning_rate = c.get<float>();
    cfg.P_reg = c.get<float>();
    cfg.Q_reg = c.get<float>();
    cfg.user_bias_reg = c.get<float>();
    cfg.item_bias_reg = c.get<float>();
    cfg.seed = c.get<int>();
    int start_user = c.get<int>();
    cfg.set_cuda_variables();

    cu2rec::CudaCSRMatrix matrix(rows, cols, nnz, indptr, indices, data);
    cu2rec::CudaDenseMatrix P_d(rows, k, P), Q_d(cols, k, Q), Q_t(cols, k, Q);
    float *ub_d, *ib_d, *ib_t;
    CHECK_CUDA(cudaMalloc(&ub_d, rows * sizeof(float)));
    CHECK_CUDA(cudaMalloc(&ib_d, cols * sizeof(float)));
    CHECK_CUDA(cudaMalloc(&ib_t, cols * sizeof(float)));
    CHECK_CUDA(cudaMemcpy(ub_d, ub, rows * sizeof(float), cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(ib_d, ib, cols * sizeof(float), cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(ib_t, ib, cols * sizeof(float), cudaMemcpyHostToDevice));
    bool *flags;
    CHECK_CUDA(cudaMalloc(&flags, cols * sizeof(bool)));
    CHECK_CUDA(cudaMemset(flags, 0, cols * sizeof(bool)));
    curandState *st;
    CHECK_CUDA(cudaMalloc(&st, rows * sizeof(curandState)));
    dim3 block(cfg.n_threads), grid(rows / cfg.n_threads + 1);
    initCurand<<<grid, block>>>(st, cfg.seed, rows);
    sgd_update<<<grid, block>>>(matrix.indptr, matrix.indices, matrix.data, P_d.data, Q_d.data,
                                Q_t.data, rows, ub_d, ib_d, ib_t, st, mu, start_user, flags);
    CHECK_CUDA(cudaGetLastError());
    CHECK_CUDA(cudaDeviceSynchronize());
    std::vector<float> Po((size_t)rows * k), Qo((size_t)cols * k), ubo(rows), ibo(cols);
    P_d.to_host(Po.data());
    Q_t.to_host(Qo.data());  // post-swap "current" Q
    CHECK_CUDA(cudaMemcpy(ubo.data(), ub_d, rows * sizeof(float), cudaMemcpyDeviceToHost));
    CHECK_CUDA(cudaMemcpy(ibo.data(), ib_t, cols * sizeof(float), cudaMemcpyDeviceToHost));
    FILE *f = fopen(out, "wb");
    dump(f, Po.data(), Po.size() * sizeof(float));
    dump(f, Qo.data(), Qo.size() * sizeof(float));
    dump(f, ubo.data(), ubo.size() * sizeof(float));
    dump(f, ibo.data(), ibo.size() * sizeof(float));
    fclose(f);
    cudaFree(ub_d); cudaFree(ib_d); cudaFree(ib_t); cudaFree(flags); cudaFree(st);
    return 0;
}

static int cmd_total_loss(int n, int grid, int block) {
    std::vector<float> ones(n, 1.0f);
    float *e_d;
    double *b_d;
    std::vector<double> b_h(grid);
    CHECK_CUDA(cudaMalloc(&e_d, n * sizeof(float)));
    CHECK_CUDA(cudaMalloc(&b_d, grid * sizeof(double)));
    CHECK_CUDA(cudaMemcpy(e_d, ones.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    float mae, rmse;
    std::tie(mae, rmse) = get_error_metrics_gpu(e_d, b_d, b_h.data(), n, grid, block);
    printf("mae %.9g rmse %.9g\n", mae, rmse);
    cudaFree(e_d); cudaFree(b_d);
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: ref_harness <command> ...\n"); return 2; }
    std::string cmd = argv[1];
    if (cmd == "init_normal" && argc == 5) {
        int size = atoi(argv[2]), k = atoi(argv[3]);
        float *a = initialize_normal_array(size, k);
        FILE *f = fopen(argv[4], "wb");
        dump(f, a, (size_t)size * sizeof(float));
        fclose(f);
        delete[] a;
        return 0;
    }
    if (cmd == "read_csv" && argc == 4) {
        int rows, cols;
        float gb;
        std::vector<Rating> r = readCSV(argv[2], &rows, &cols, &gb);
        int n = (int)r.size();
        FILE *f = fopen(argv[3], "wb");
        dump(f, &n, 4); dump(f, &rows, 4); dump(f, &cols, 4); dump(f, &gb, 4);
        dump(f, r.data(), r.size() * sizeof(Rating));
        fclose(f);
        printf("n %d rows %d cols %d global_bias %.9g\n", n, rows, cols, gb);
        return 0;
    }
    if (cmd == "read_config" && argc == 3) {
        config::Config cfg;
        cfg.read_config(argv[2]);
        printf("%d %d %d %.9g %d %.9g %.9g %.9g %.9g\n", cfg.cur_iterations, cfg.total_iterations,
               cfg.n_factors, cfg.learning_rate, cfg.seed, cfg.P_reg, cfg.Q_reg, cfg.user_bias_reg,
               cfg.item_bias_reg);
        return 0;
    }
    if (cmd == "write_csv" && argc == 6) {
        std::vector<char> buf = slurp(argv[2]);
        int rows = atoi(argv[3]), cols = atoi(argv[4]);
        writeCSV(argv[5], (float *)buf.data(), rows, cols);
        return 0;
    }
    if (cmd == "read_array" && argc == 4) {
        int r, c;
        float *a = read_array(argv[2], &r, &c);
        if (!a) return 3;
        FILE *f = fopen(argv[3], "wb");
        dump(f, &r, 4); dump(f, &c, 4);
        dump(f, a, (size_t)c * sizeof(float));  // util.cu:61-66: n_cols accumulates over all rows
        fclose(f);
        printf("rows %d cols %d\n", r, c);
        return 0;
    }
    if (cmd == "loss_gpu" && argc == 4) return cmd_loss_gpu(argv[2], argv[3]);
    if (cmd == "sgd_gpu" && argc == 4) return cmd_sgd_gpu(argv[2], argv[3]);
    if (cmd == "total_loss" && argc == 5) return cmd_total_loss(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]));
    fprintf(stderr, "unknown command / wrong arity\n");
    return 2;
}
