// bin/predict -- the recommender executable of the drop-in surface.
//
//   predict -c <config> -i <item_bias.csv> -g <global_bias.csv> -q <Q.csv> <user_ratings.csv>
//
// Contract taken from the reference CLI (predict.cu:72-146; this file shares no code with it): the
// ratings file describes ONE new user; that user's factors and bias are fitted against the trained,
// frozen item side (the reference means to freeze it, predict.cu:105, but its device constant is
// never refreshed, config.cu:24-35), every item is scored, the items the user has not rated are
// listed best first. No arguments -> exit status 2; unknown option -> "Unknown option.", status 1.
//
// Extension (add-only): with -p <P.csv> -u <user_bias.csv> [-k <topk>] [-x <train.csv>] the positional
// file is not needed; ALL users of P are scored against the catalogue on the tensor cores
// (cu2b_predict_topk, tcgen05) and the top-k unrated items of each user are printed.
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <iostream>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "cu2rec_shim.h"

namespace {

struct CommandLine {
    std::string config_file, item_bias_file, global_bias_file, item_factor_file, ratings_file;
    std::string user_factor_file, user_bias_file, seen_file;  // batched extension
    int topk = 10;
    bool batched() const { return !user_factor_file.empty(); }
};

enum class ArgStatus { ok, nothing_given, bad_option, missing_file };

ArgStatus read_command_line(int argc, char **argv, CommandLine *cl) {
    if (argc <= 1) return ArgStatus::nothing_given;
    for (int opt; (opt = getopt(argc, argv, "c:i:g:q:p:u:k:x:")) != -1;) {
        switch (opt) {
            case 'c': cl->config_file = optarg; break;
            case 'i': cl->item_bias_file = optarg; break;
            case 'g': cl->global_bias_file = optarg; break;
            case 'q': cl->item_factor_file = optarg; break;
            case 'p': cl->user_factor_file = optarg; break;
            case 'u': cl->user_bias_file = optarg; break;
            case 'x': cl->seen_file = optarg; break;
            case 'k': cl->topk = std::atoi(optarg); break;
            default: return ArgStatus::bad_option;
        }
    }
    if (optind < argc) cl->ratings_file = argv[optind];
    if (cl->ratings_file.empty() && !cl->batched()) return ArgStatus::missing_file;
    return ArgStatus::ok;
}

// A float matrix file as the shared reader returns it: a flat array, its row count and its TOTAL
// element count (the reader's "columns" accumulate over the rows, util.cu:61-66).
struct FloatFile {
    std::unique_ptr<float[]> values;
    int n_rows = 0, n_values = 0;
    explicit FloatFile(const std::string &path) {
        values.reset(read_array(path.c_str(), &n_rows, &n_values));
        if (!values) throw std::runtime_error("cannot read " + path);
    }
    int width() const { return n_rows > 0 ? n_values / n_rows : 0; }
};

struct ItemSide {
    FloatFile factors, bias, mean;
    int n_items, n_factors;
    ItemSide(const CommandLine &cl, const config::Config &cfg)
        : factors(cl.item_factor_file), bias(cl.item_bias_file), mean(cl.global_bias_file),
          n_items(factors.n_rows), n_factors(factors.width()) {
        if (n_factors != cfg.n_factors || bias.n_rows != n_items || mean.n_values < 1)
            throw std::runtime_error("model files do not match the config (n_factors / item count)");
    }
    float global_mean() const { return mean.values[0]; }
};

void print_ranked(const std::vector<float> &score, const std::vector<char> &seen) {
    // unrated items, best first; equal scores keep ascending item order
    std::vector<int> order;
    for (int item = 0; item < (int)score.size(); ++item)
        if (!seen[item]) order.push_back(item);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return score[a] > score[b]; });
    std::cout << "Recommendations:" << std::endl;
    int place = 0;
    for (int item : order) std::printf("Rank: %d\tItem: %d\tEstimated rating: %f\n", ++place, item, score[item]);
}

int recommend_for_new_user(const CommandLine &cl, config::Config &cfg) {
    cfg.is_train = 0;  // the item side is an input here, not a result
    ItemSide items(cl, cfg);

    int seen_users = 0, seen_items = 0;
    float unused_mean = 0.f;
    std::vector<Rating> given = readCSV(cl.ratings_file, &seen_users, &seen_items, &unused_mean);
    std::vector<char> seen((size_t)items.n_items, 0);
    for (Rating &r : given) {
        if (r.itemID < 0 || r.itemID >= items.n_items) throw std::runtime_error("rated item id beyond the trained catalogue");
        r.userID = 0;  // whatever id the file used, this is user 0 of a one-user problem
        seen[r.itemID] = 1;
    }
    std::unique_ptr<cu2rec::CudaCSRMatrix> one_user(createSparseMatrix(&given, 1, items.n_items));

    float *fitted = nullptr, *curve = nullptr, *fitted_bias = nullptr, *q_same = nullptr, *bi_same = nullptr;
    train(one_user.get(), one_user.get(), &cfg, &fitted, &q_same, items.factors.values.get(), &curve, &fitted_bias, &bi_same,
          items.bias.values.get(), items.global_mean());
    std::unique_ptr<float[]> user_factors(fitted), validation_curve(curve), user_bias(fitted_bias);

    // score of item i for this user: mean + b_u + b_i + <q_i, p>, accumulated in factor order in fp32
    std::vector<float> score((size_t)items.n_items);
    const float base = items.global_mean() + user_bias[0];
    for (int i = 0; i < items.n_items; ++i) {
        const float *q = items.factors.values.get() + (size_t)i * items.n_factors;
        float s = base + items.bias.values[i];
        for (int f = 0; f < items.n_factors; ++f) s += q[f] * user_factors[f];
        score[i] = s;
    }
    std::cout << "Predictions: \n[";
    for (float s : score) std::cout << s << ", ";
    std::cout << "]\n";
    print_ranked(score, seen);
    return 0;
}

// Extension: top-k for every user of a trained model through the tcgen05 candidate kernel.
int recommend_for_all_users(const CommandLine &cl, config::Config &cfg) {
    ItemSide items(cl, cfg);
    FloatFile users(cl.user_factor_file), user_bias(cl.user_bias_file);
    if (users.width() != items.n_factors || user_bias.n_rows != users.n_rows)
        throw std::runtime_error("user-side files do not match the item side");
    std::unique_ptr<cu2rec::CudaCSRMatrix> seen;
    cu2b_csr seen_view;
    if (!cl.seen_file.empty()) {
        int r = 0, c = 0;
        float m = 0.f;
        std::vector<Rating> known = readCSV(cl.seen_file, &r, &c, &m);
        if (r > users.n_rows || c > items.n_items) throw std::runtime_error("-x file reaches beyond the model");
        seen.reset(createSparseMatrix(&known, users.n_rows, items.n_items));
        seen_view = seen->view();
    }
    const int k = std::max(1, cl.topk);
    std::vector<int32_t> best((size_t)users.n_rows * k);
    std::vector<float> best_score((size_t)users.n_rows * k);
    double ms[2] = {0, 0};
    CU2B_CHECK(cu2b_predict_topk(users.values.get(), users.n_rows, items.factors.values.get(), items.n_items,
                                 user_bias.values.get(), items.bias.values.get(), items.global_mean(), items.n_factors,
                                 seen ? &seen_view : nullptr, k, best.data(), best_score.data(), ms));
    for (int u = 0; u < users.n_rows; ++u)
        for (int j = 0; j < k && best[(size_t)u * k + j] >= 0; ++j)
            std::printf("User: %d\tRank: %d\tItem: %d\tEstimated rating: %f\n", u, j + 1, best[(size_t)u * k + j],
                        best_score[(size_t)u * k + j]);
    std::fprintf(stderr, "cu2b: %d users x %d items scored in %.3f ms (candidates) + %.3f ms (exact rescoring)\n", users.n_rows,
                 items.n_items, ms[0], ms[1]);
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    CommandLine cl;
    switch (read_command_line(argc, argv, &cl)) {
        case ArgStatus::nothing_given:
            return 2;
        case ArgStatus::bad_option:
            std::puts("Unknown option.");
            return 1;
        case ArgStatus::missing_file:
            std::fputs("usage: predict -c cfg -i item_bias -g global_bias -q Q ratings.csv\n"
                       "       predict -c cfg -i item_bias -g global_bias -q Q -p P -u user_bias [-k topk] [-x seen.csv]\n", stderr);
            return 2;
        case ArgStatus::ok:
            break;
    }
    try {
        config::Config cfg;
        cfg.read_config(cl.config_file);
        return cl.batched() ? recommend_for_all_users(cl, cfg) : recommend_for_new_user(cl, cfg);
    } catch (const std::exception &err) {
        std::fprintf(stderr, "terminate called after throwing an instance of 'std::runtime_error'\n  what():  %s\n", err.what());
        return 134;
    }
}
