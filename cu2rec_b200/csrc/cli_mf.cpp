// bin/mf -- the trainer executable of the drop-in surface.
//
//   mf [-c <config>] <train.csv> <test.csv>
//
// Contract taken from the reference CLI (mf.cu:16-99; this file shares no code with it):
//   * no arguments -> exit status 255 without output; an unknown option prints "Unknown option."
//     and exits 1; -c is optional (defaults of config.h:23-51);
//   * stdout: "Free memory: <bytes>" + blank line, the hyper-parameter block, the TRAIN / TEST
//     lines of every loss check, the "Time taken" line;
//   * five "%f" CSV files next to the training file: <stem>_f<k>_{p,q,user_bias,item_bias,global_bias}.csv,
//     the set bin/predict (ours or the reference's) loads.
// With config token 17 (n_gpus) > 1 the same call trains by DSGD on that many GPUs of the box.
// Everything numeric happens on the GPU inside libcu2b.so; there is no CPU training path.
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <exception>
#include <filesystem>
#include <memory>
#include <string>
#include <vector>

#include "cu2rec_shim.h"

namespace fs = std::filesystem;

namespace {

enum class ArgStatus { ok, nothing_given, bad_option, missing_files };

struct CommandLine {
    std::string config_file;  // empty: built-in defaults
    fs::path train_file, test_file;
};

ArgStatus read_command_line(int argc, char **argv, CommandLine *cl) {
    if (argc <= 1) return ArgStatus::nothing_given;
    for (int opt; (opt = getopt(argc, argv, "c:")) != -1;) {
        if (opt != 'c') return ArgStatus::bad_option;
        cl->config_file = optarg;
    }
    if (argc - optind < 2) return ArgStatus::missing_files;
    cl->train_file = argv[optind];
    cl->test_file = argv[optind + 1];
    return ArgStatus::ok;
}

// One ratings CSV in memory together with what the reader derives from it.
struct RatingSet {
    std::vector<Rating> triplets;
    int n_users = 0, n_items = 0;
    float mean = 0.f;
    explicit RatingSet(const fs::path &file) { triplets = readCSV(file.string(), &n_users, &n_items, &mean); }
};

// Output files sit beside the training file and reuse its name without the extension.
struct OutputPlace {
    std::string directory, stem;
    explicit OutputPlace(const fs::path &train_file) {
        const fs::path parent = train_file.parent_path();
        directory = parent.empty() ? std::string(".") : parent.string();
        stem = train_file.stem().string();
    }
    void put(const char *component, float *values, int rows, int cols, int n_factors) const {
        writeToFile(directory, stem, "csv", component, values, rows, cols, n_factors);
    }
};

struct TrainedModel {  // owns what train() hands back through its pointer-to-pointer outputs
    std::unique_ptr<float[]> user_factors, item_factors, user_bias, item_bias, validation_curve;
};

TrainedModel fit(cu2rec::CudaCSRMatrix &train_m, cu2rec::CudaCSRMatrix &test_m, config::Config &cfg, float mean_rating) {
    float *p = nullptr, *q = nullptr, *curve = nullptr, *bu = nullptr, *bi = nullptr;
    train(&train_m, &test_m, &cfg, &p, &q, &curve, &bu, &bi, mean_rating);
    TrainedModel m;
    m.user_factors.reset(p);
    m.item_factors.reset(q);
    m.validation_curve.reset(curve);
    m.user_bias.reset(bu);
    m.item_bias.reset(bi);
    return m;
}

int run(const CommandLine &cl) {
    size_t device_total = 0;
    std::printf("Free memory: %ld\n\n", (long)getFreeBytes(0, &device_total));

    RatingSet train_set(cl.train_file), test_set(cl.test_file);
    // The model must cover every id that occurs in either file. (The reference sizes each matrix
    // from its own file, mf.cu:43-51, and indexes past P / Q when the test file reaches further.)
    const int n_users = std::max(train_set.n_users, test_set.n_users);
    const int n_items = std::max(train_set.n_items, test_set.n_items);
    std::unique_ptr<cu2rec::CudaCSRMatrix> train_m(createSparseMatrix(&train_set.triplets, n_users, n_items));
    std::unique_ptr<cu2rec::CudaCSRMatrix> test_m(createSparseMatrix(&test_set.triplets, n_users, n_items));

    config::Config cfg;
    if (!cl.config_file.empty()) cfg.read_config(cl.config_file);
    cfg.print_config();

    TrainedModel model = fit(*train_m, *test_m, cfg, train_set.mean);

    const OutputPlace out(cl.train_file);
    const int k = cfg.n_factors;
    out.put("p", model.user_factors.get(), n_users, k, k);
    out.put("q", model.item_factors.get(), n_items, k, k);
    out.put("user_bias", model.user_bias.get(), n_users, 1, k);
    out.put("item_bias", model.item_bias.get(), n_items, 1, k);
    float mean_cell = train_set.mean;
    out.put("global_bias", &mean_cell, 1, 1, k);
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    CommandLine cl;
    switch (read_command_line(argc, argv, &cl)) {
        case ArgStatus::nothing_given:
            return -1;
        case ArgStatus::bad_option:
            std::puts("Unknown option.");
            return 1;
        case ArgStatus::missing_files:
            std::fputs("usage: mf [-c config] train.csv test.csv\n", stderr);
            return -1;
        case ArgStatus::ok:
            break;
    }
    try {
        return run(cl);
    } catch (const std::exception &err) {
        // An uncaught std::runtime_error is how the reference ends on a CUDA failure (util.h:27-34):
        // same message shape on stderr, same abort status.
        std::fprintf(stderr, "terminate called after throwing an instance of 'std::runtime_error'\n  what():  %s\n", err.what());
        return 134;
    }
}
