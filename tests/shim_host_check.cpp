// Host-side check of the C++ shim (cu2rec_shim.h): the calls the reference's tests/test_util.cu and
// tests/test_config.cu make, through the reference's own names, printing what they assert on.
// Compiled and run by tests/test_shim.py (no GPU: nothing here reaches a kernel).
//   usage: shim_host_check <fixtures dir> <scratch dir>
#include <cinttypes>
#include <cstdio>
#include <string>

#include "cu2rec_shim.h"

static void print_matrix(const char *tag, cu2rec::CudaCSRMatrix *m) {
    printf("%s rows=%d cols=%d nonzeros=%d\n", tag, m->rows, m->cols, m->nonzeros);
    printf("%s indptr", tag);
    for (int i = 0; i <= m->rows; ++i) printf(" %d", m->indptr[i]);
    printf("\n%s indices", tag);
    for (int i = 0; i < m->nonzeros; ++i) printf(" %d", m->indices[i]);
    printf("\n%s data", tag);
    for (int i = 0; i < m->nonzeros; ++i) printf(" %g", m->data[i]);
    printf("\n");
}

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    const std::string dir = argv[1], scratch = argv[2];
    int rows, cols;
    float global_bias;
    // test_util.cu:20-34
    std::vector<Rating> ratings = readCSV(dir + "/test_ratings.csv", &rows, &cols, &global_bias);
    printf("read_csv rows=%d cols=%d n=%zu global_bias=%.6f\n", rows, cols, ratings.size(), global_bias);
    // test_util.cu:98-143
    cu2rec::CudaCSRMatrix *m = createSparseMatrix(&ratings, rows, cols);
    print_matrix("sparse", m);
    delete m;
    // test_util.cu:146-190
    std::vector<Rating> missing = readCSV(dir + "/test_missing_user_ratings.csv", &rows, &cols, &global_bias);
    m = createSparseMatrix(&missing, rows, cols);
    print_matrix("missing", m);
    delete m;
    // test_util.cu:36-47
    int n_rows = 0, n_cols = 0;
    float *arr = read_array((dir + "/test_Q.csv").c_str(), &n_rows, &n_cols);
    printf("read_array n_rows=%d n_cols=%d first", n_rows, n_cols);
    for (int i = 0; i < 10; ++i) printf(" %g", arr[i]);
    printf("\n");
    delete[] arr;
    printf("read_array_missing %s\n", read_array((scratch + "/nope.csv").c_str()) == nullptr ? "nullptr" : "pointer");
    // util.cu:124-144 through every overload
    float *a = initialize_normal_array(64, 2), *b = initialize_normal_array(64, 2, 42), *c = initialize_normal_array(64, 2, 0, 1, 42);
    printf("init_normal bits");
    for (int i = 0; i < 64; ++i) {
        uint32_t u;
        memcpy(&u, a + i, 4);
        printf(" %08" PRIx32, u);
    }
    printf("\ninit_normal overloads_agree=%d\n", !memcmp(a, b, 256) && !memcmp(a, c, 256));
    delete[] a;
    delete[] b;
    delete[] c;
    // test_util.cu:50-95
    float ones[12];
    for (float &v : ones) v = 1.0f;
    float gb[1] = {3.5f};
    writeToFile(scratch, "test_ratings", "csv", "p", ones, 6, 2, 2);
    writeToFile(scratch, "test_ratings", "csv", "global_bias", gb, 1, 1, 2);
    // test_config.cu:9-27
    config::Config cfg;
    printf("config defaults total_iterations=%d n_factors=%d check_error=%d\n", cfg.total_iterations, cfg.n_factors, cfg.check_error);
    cfg.read_config(dir + "/test_config.cfg");
    printf("config read total_iterations=%d P_reg=%.6f\n", cfg.total_iterations, cfg.P_reg);
    cfg.total_iterations = 250;
    cfg.P_reg = 0.3f;
    cfg.write_config(scratch + "/gen.cfg");
    config::Config back;
    back.read_config(scratch + "/gen.cfg");
    printf("config round_trip total_iterations=%d P_reg=%.6f\n", back.total_iterations, back.P_reg);
    back.read_config(scratch + "/does_not_exist.cfg");  // config.cu:7-13: silently keeps the values
    printf("config unreadable_keeps total_iterations=%d\n", back.total_iterations);
    fflush(stdout);
    cfg.print_config();
    // error behaviour: missing ratings file -> message + empty vector (util.cu:41-44); bad input -> std::runtime_error
    std::vector<Rating> none = readCSV(scratch + "/nope.csv", &rows, &cols, &global_bias);
    printf("read_csv_missing n=%zu\n", none.size());
    std::vector<Rating> unsorted = {{2, 1, 1.0f}, {1, 1, 1.0f}};
    for (Rating &r : unsorted) { r.userID -= 1; r.itemID -= 1; }
    try {
        createSparseMatrix(&unsorted, 2, 1);
        printf("unsorted accepted\n");
    } catch (const std::runtime_error &e) {
        printf("unsorted threw: %s\n", e.what());
    }
    return 0;
}
