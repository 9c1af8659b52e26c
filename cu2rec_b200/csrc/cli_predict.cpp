// bin/predict -- drop-in for the reference's recommender CLI (predict.cu:72-146):
//   predict -c <config> -i <item_bias.csv> -g <global_bias.csv> -q <Q.csv> <user_ratings.csv>
// Partial fit of ONE new user against a trained Q / item_bias (which stay frozen -- the
// reference intends this, predict.cu:105, but never copies is_train to the device,
// config.cu:24-35), then a score for every item and the ranked list of unrated items.
#include <getopt.h>

#include <algorithm>
#include <iostream>
#include <set>

#include "cu2rec_shim.h"

using namespace cu2rec;
using std::cout;
using std::string;
using std::vector;

typedef std::pair<float, int> rated_item;

int main(int argc, char **argv) {
    if (argc < 2) return 2;  // predict.cu:73-75
    string filename_config, filename_item_bias, filename_global_bias, filename_Q;
    int o;
    while ((o = getopt(argc, argv, "c:i:g:q:")) != -1) {
        switch (o) {
            case 'c': filename_config = optarg; break;
            case 'i': filename_item_bias = optarg; break;
            case 'g': filename_global_bias = optarg; break;
            case 'q': filename_Q = optarg; break;
            default:
                cout << "Unknown option.\n";
                return 1;
        }
    }
    if (optind >= argc) {
        std::cerr << "usage: predict -c cfg -i item_bias -g global_bias -q Q ratings.csv\n";
        return 2;
    }
    try {
        config::Config *cfg = new config::Config();
        cfg->read_config(filename_config);
        cfg->is_train = 0;
        int n_items = 0, n_factors = 0, tmp_r = 0, tmp_c = 0;
        float *item_bias = read_array(filename_item_bias.c_str(), &tmp_r, &tmp_c);
        float *global_bias_arr = read_array(filename_global_bias.c_str());
        float *Q = read_array(filename_Q.c_str(), &n_items, &n_factors);
        if (!item_bias || !global_bias_arr || !Q) throw std::runtime_error("cannot read model files");
        const float global_bias = global_bias_arr[0];
        // read_array's column count accumulates over rows (util.cu:61-66): n_factors = total / rows
        if (n_items > 0) n_factors /= n_items;
        if (n_factors != cfg->n_factors || tmp_r != n_items)
            throw std::runtime_error("model files do not match the config (n_factors / item count)");

        string filename_user_ratings = argv[optind++];
        int rows, cols;
        float user_mean;
        vector<Rating> ratings = readCSV(filename_user_ratings, &rows, &cols, &user_mean);
        for (Rating &r : ratings) {
            r.userID = 0;  // predict.cu:120-122
            if (r.itemID >= n_items) throw std::runtime_error("rated item id beyond the trained catalogue");
        }
        CudaCSRMatrix *matrix = createSparseMatrix(&ratings, 1, n_items);

        float *P, *losses, *user_bias;
        train(matrix, matrix, cfg, &P, &Q, Q, &losses, &user_bias, &item_bias, item_bias, global_bias);

        // predict.cu:17-29: score every item for the fitted user
        vector<float> predictions((size_t)n_items);
        for (int i = 0; i < n_items; i++) {
            const float *Q_i = &Q[(size_t)i * n_factors];
            float pred = global_bias + user_bias[0] + item_bias[i];
            for (int f = 0; f < n_factors; f++) pred += Q_i[f] * P[f];
            predictions[i] = pred;
        }
        cout << "Predictions: \n[";  // predict.cu:31-38
        for (int i = 0; i < n_items; i++) cout << predictions[i] << ", ";
        cout << "]\n";
        // predict.cu:49-63: drop the items the user rated, sort the rest high to low
        std::set<int> rated;
        for (const Rating &r : ratings) rated.insert(r.itemID);
        vector<rated_item> items;
        for (int item = 0; item < n_items; ++item)
            if (!rated.count(item)) items.push_back(rated_item(predictions[item], item));
        std::stable_sort(items.begin(), items.end(), [](const rated_item &l, const rated_item &r) { return l.first > r.first; });
        cout << "Recommendations:" << std::endl;  // predict.cu:65-70
        for (size_t i = 0; i < items.size(); ++i)
            printf("Rank: %d\tItem: %d\tEstimated rating: %f\n", (int)i + 1, items[i].second, items[i].first);

        delete cfg;
        delete matrix;
        delete[] losses;
        delete[] P;
        delete[] Q;
        delete[] user_bias;
        delete[] item_bias;
        delete[] global_bias_arr;
    } catch (const std::exception &e) {
        std::cerr << "terminate called after throwing an instance of 'std::runtime_error'\n  what():  " << e.what() << "\n";
        return 134;
    }
    return 0;
}
