// bin/prep -- the reference's preprocessing/ scripts as one native tool, same output
// file names and bytes:
//   prep map_items <ratings.csv>                     -> <ratings>_mapped.csv        (map_items.py)
//   prep map_netflix <train.txt> <test.txt> <train_out.csv> <test_out.csv>          (map_netflix.py)
//   prep sort_ratings <ratings.csv>                  -> <ratings>_sorted.csv        (sort_ratings.py)
//   prep split_to_test_train <ratings.csv> <test_ratio> [-s seed]
//                                                    -> <ratings>_train.csv, <ratings>_test.csv
//   prep create_config <file> [-n iters] [-f factors] [-l lr] [-s seed] [-p p_reg] [-q q_reg]
//                      [-u user_bias_reg] [-i item_bias_reg]                        (create_config.py)
//   prep convert_to_np <matrix.csv> [...]            -> <matrix>.npy                (convert_to_np.py)
//   prep dsgd_partition <train.csv> <n_gpus>         -> <train>_dsgd<G>_users.csv, <train>_dsgd<G>_items.csv
//        (no reference counterpart: the block permutation multi-GPU training uses, cu2b_dsgd_partition)
// Host only (no GPU needed).
#include <getopt.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "cu2b.h"

static std::string with_suffix(const std::string &path, const char *suffix) {  // os.path.splitext
    const size_t slash = path.find_last_of('/');
    const size_t dot = path.find_last_of('.');
    const bool has_ext = dot != std::string::npos && (slash == std::string::npos || dot > slash + 1) && dot != 0;
    const std::string root = has_ext ? path.substr(0, dot) : path, ext = has_ext ? path.substr(dot) : "";
    return root + "_" + suffix + ext;
}

static int fail() {
    fprintf(stderr, "prep: %s\n", cu2b_last_error());
    return 1;
}

static int usage() {
    fprintf(stderr,
            "usage: prep map_items <ratings.csv>\n"
            "       prep map_netflix <train.txt> <test.txt> <train_out.csv> <test_out.csv>\n"
            "       prep sort_ratings <ratings.csv>\n"
            "       prep split_to_test_train <ratings.csv> <test_ratio> [-s seed]\n"
            "       prep create_config <file> [-n N] [-f F] [-l LR] [-s SEED] [-p P] [-q Q] [-u UB] [-i IB]\n"
            "       prep convert_to_np <matrix.csv> [...]\n"
            "       prep dsgd_partition <train.csv> <n_gpus>\n");
    return 2;
}

int main(int argc, char **argv) {
    if (argc < 3) return usage();
    const std::string cmd = argv[1];
    if (cmd == "map_items") {
        int64_t rows = 0, users = 0, items = 0;
        const std::string out = with_suffix(argv[2], "mapped");
        if (cu2b_prep_map(argv[2], out.c_str(), ',', 1, 2, nullptr, nullptr, &rows, nullptr, &users, &items, nullptr, nullptr) != CU2B_OK)
            return fail();
        fprintf(stderr, "%lld rows, %lld users, %lld items -> %s\n", (long long)rows, (long long)users, (long long)items, out.c_str());
        return 0;
    }
    if (cmd == "map_netflix") {
        if (argc < 6) return usage();
        int64_t rows = 0, rows2 = 0, su = 0, si = 0;
        if (cu2b_prep_map(argv[2], argv[4], ' ', 0, 3, argv[3], argv[5], &rows, &rows2, nullptr, nullptr, &su, &si) != CU2B_OK)
            return fail();
        if (su > 0) printf("Skipped %lld rows because of missing users\n", (long long)su);  // map_items.py:55-58
        if (si > 0) printf("Skipped %lld rows because of missing items\n", (long long)si);
        return 0;
    }
    if (cmd == "sort_ratings") {
        printf("sorting ratings...\n");  // sort_ratings.py:33-35
        const std::string out = with_suffix(argv[2], "sorted");
        if (cu2b_prep_sort(argv[2], out.c_str(), nullptr) != CU2B_OK) return fail();
        printf("done sorting ratings...\n");
        return 0;
    }
    if (cmd == "split_to_test_train") {
        if (argc < 4) return usage();
        long long seed = 42;
        for (int a = 4; a + 1 < argc; ++a)
            if (!strcmp(argv[a], "-s") || !strcmp(argv[a], "--seed")) seed = atoll(argv[a + 1]);
        const std::string tr = with_suffix(argv[2], "train"), te = with_suffix(argv[2], "test");
        if (cu2b_prep_split(argv[2], tr.c_str(), te.c_str(), atof(argv[3]), seed, nullptr, nullptr) != CU2B_OK) return fail();
        return 0;
    }
    if (cmd == "create_config") {
        int n = 1000, f = 100, s = 42;  // create_config.py:24-32 defaults
        double l = 0.01, p = 0.02, q = 0.02, u = 0.02, i = 0.02;
        const char *file = argv[2];
        optind = 3;
        int o;
        while ((o = getopt(argc, argv, "n:f:l:s:p:q:u:i:t:a:d:")) != -1) {
            switch (o) {
                case 'n': n = atoi(optarg); break;
                case 'f': f = atoi(optarg); break;
                case 'l': l = atof(optarg); break;
                case 's': s = atoi(optarg); break;
                case 'p': p = atof(optarg); break;
                case 'q': q = atof(optarg); break;
                case 'u': u = atof(optarg); break;
                case 'i': i = atof(optarg); break;
                case 't': case 'a': case 'd': break;  // accepted and not written, like the script
                default: return usage();
            }
        }
        if (cu2b_prep_create_config(file, n, f, l, s, p, q, u, i) != CU2B_OK) return fail();
        return 0;
    }
    if (cmd == "convert_to_np") {
        for (int a = 2; a < argc; ++a) {
            // convert_to_np.py:11-13: os.path.splitext(filename)[0] + ".npy"
            const std::string in = argv[a];
            const size_t slash = in.find_last_of('/'), dot = in.find_last_of('.');
            const bool has_ext = dot != std::string::npos && dot != 0 && (slash == std::string::npos || dot > slash + 1);
            const std::string out = (has_ext ? in.substr(0, dot) : in) + ".npy";
            if (cu2b_prep_convert_to_np(in.c_str(), out.c_str(), nullptr, nullptr) != CU2B_OK) return fail();
        }
        return 0;
    }
    if (cmd == "dsgd_partition") {
        // The user / item block permutation of the DSGD trainer for this file and GPU count, as two CSVs with the
        // 1-based ids of the ratings file: userId,block,local_index and itemId,block,row (row = position of the item
        // in the renumbered catalogue; block b is the contiguous row range the trainer rotates as one unit).
        if (argc < 4) return usage();
        const int world = atoi(argv[3]);
        cu2b_rating *r = nullptr;
        int64_t n = 0;
        int rows = 0, cols = 0;
        float mean = 0.f;
        if (cu2b_read_csv(argv[2], &r, &n, &rows, &cols, &mean) != CU2B_OK) return fail();
        if (world < 1 || rows < 1 || cols < 1) {
            cu2b_free(r);
            fprintf(stderr, "prep: dsgd_partition needs n_gpus >= 1 and a non-empty ratings file\n");
            return 1;
        }
        std::vector<int> ublock((size_t)rows), ulocal((size_t)rows), per_block((size_t)world), inew((size_t)cols), iptr((size_t)world + 1);
        std::vector<int64_t> nnz((size_t)world * world);
        const cu2b_status rc = cu2b_dsgd_partition(r, n, rows, cols, world, ublock.data(), ulocal.data(), per_block.data(), inew.data(),
                                                   iptr.data(), nnz.data());
        cu2b_free(r);
        if (rc != CU2B_OK) return fail();
        const std::string tag = "dsgd" + std::to_string(world);
        const std::string uf = with_suffix(argv[2], (tag + "_users").c_str()), itf = with_suffix(argv[2], (tag + "_items").c_str());
        FILE *f = fopen(uf.c_str(), "w");
        if (!f) { fprintf(stderr, "prep: cannot write %s\n", uf.c_str()); return 1; }
        fprintf(f, "userId,block,local_index\n");
        for (int u = 0; u < rows; ++u) fprintf(f, "%d,%d,%d\n", u + 1, ublock[u], ulocal[u]);
        fclose(f);
        f = fopen(itf.c_str(), "w");
        if (!f) { fprintf(stderr, "prep: cannot write %s\n", itf.c_str()); return 1; }
        fprintf(f, "itemId,block,row\n");
        for (int i = 0; i < cols; ++i) {
            int b = 0;
            while (b + 1 < world && inew[i] >= iptr[b + 1]) ++b;
            fprintf(f, "%d,%d,%d\n", i + 1, b, inew[i]);
        }
        fclose(f);
        int64_t lo = INT64_MAX, hi = 0;
        for (int64_t v : nnz) { lo = v < lo ? v : lo; hi = v > hi ? v : hi; }
        printf("%d x %d rating blocks, %lld ratings: smallest block %lld, largest %lld (x%.4f of the mean)\n", world, world, (long long)n,
               (long long)lo, (long long)hi, n > 0 ? (double)hi * world * world / (double)n : 0.0);
        return 0;
    }
    return usage();
}
