// bin/mf -- drop-in for the reference trainer CLI (mf.cu:16-99):
//   mf [-c <config>] <train.csv> <test.csv>
// Same flags, same stdout line formats, same five output files next to the training file
// (<base>_f<k>_{p,q,user_bias,item_bias,global_bias}.csv, "%f" text). Training runs on the GPU
// through libcu2b.so; there is no CPU path.
#include <getopt.h>

#include <algorithm>
#include <iostream>

#include "cu2rec_shim.h"

using namespace cu2rec;
using std::string;

int main(int argc, char **argv) {
    if (argc < 2) return -1;  // mf.cu:17-19
    string filename_config;
    int o;
    while ((o = getopt(argc, argv, "c:")) != -1) {
        switch (o) {
            case 'c':
                filename_config = optarg;
                break;
            default:
                std::cout << "Unknown option.\n";  // mf.cu:28-29
                return 1;
        }
    }
    if (argc - optind < 2) {
        std::cerr << "usage: mf [-c config] train.csv test.csv\n";
        return -1;
    }
    try {
        size_t total_bytes;
        const long free_bytes_before = (long)getFreeBytes(0, &total_bytes);
        printf("Free memory: %ld\n\n", free_bytes_before);  // mf.cu:37

        string file_path_train = argv[optind++];
        int rows, cols;
        float global_bias;
        std::vector<Rating> train_ratings = readCSV(file_path_train, &rows, &cols, &global_bias);
        string file_path_test = argv[optind++];
        int r, c;
        float gb;
        std::vector<Rating> test_ratings = readCSV(file_path_test, &r, &c, &gb);
        // The reference sizes the test matrix from the test file alone (mf.cu:50-51) and then
        // reads P/Q out of bounds if it is larger; we size the model with max(train, test).
        rows = std::max(rows, r);
        cols = std::max(cols, c);
        CudaCSRMatrix *train_matrix = createSparseMatrix(&train_ratings, rows, cols);
        CudaCSRMatrix *test_matrix = createSparseMatrix(&test_ratings, rows, cols);

        config::Config *cfg = new config::Config();
        if (!filename_config.empty()) cfg->read_config(filename_config);
        cfg->print_config();

        float *P, *Q, *losses, *user_bias, *item_bias;
        train(train_matrix, test_matrix, cfg, &P, &Q, &losses, &user_bias, &item_bias, global_bias);

        // mf.cu:65-77: outputs go next to the training file
        size_t dir_index = file_path_train.find_last_of("/");
        string parent_dir, filename;
        if (dir_index != string::npos) {
            parent_dir = file_path_train.substr(0, dir_index);
            filename = file_path_train.substr(dir_index + 1);
        } else {
            parent_dir = ".";
            filename = file_path_train;
        }
        string basename = filename.substr(0, filename.find_last_of("."));
        float global_bias_array[1] = {global_bias};
        writeToFile(parent_dir, basename, "csv", "p", P, rows, cfg->n_factors, cfg->n_factors);
        writeToFile(parent_dir, basename, "csv", "q", Q, cols, cfg->n_factors, cfg->n_factors);
        writeToFile(parent_dir, basename, "csv", "user_bias", user_bias, rows, 1, cfg->n_factors);
        writeToFile(parent_dir, basename, "csv", "item_bias", item_bias, cols, 1, cfg->n_factors);
        writeToFile(parent_dir, basename, "csv", "global_bias", global_bias_array, 1, 1, cfg->n_factors);

        delete cfg;
        delete train_matrix;
        delete test_matrix;
        delete[] P;
        delete[] Q;
        delete[] losses;
        delete[] user_bias;
        delete[] item_bias;
    } catch (const std::exception &e) {
        // the reference lets std::runtime_error escape to std::terminate (util.h:27-34)
        std::cerr << "terminate called after throwing an instance of 'std::runtime_error'\n  what():  " << e.what() << "\n";
        return 134;
    }
    return 0;
}
