// shim_selftest.cc — drives the C++ mirror of the reference interface exactly the way the reference would
// (cv::Mat CV_32FC1 in, vector<pair<int,int>> out) and prints the match lists so that tests/test_gpu_parity.py can
// check them.  Input: a raw file of two float32 matrices; output: text on stdout.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "feature_matching_b200.h"

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s <desc.f32> <rows1> <rows2>\n", argv[0]);
        return 2;
    }
    const int n1 = atoi(argv[2]), n2 = atoi(argv[3]);
    std::vector<float> buf((size_t)(n1 + n2) * 128);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(buf.data(), sizeof(float), buf.size(), f) != buf.size()) {
        fprintf(stderr, "cannot read %s\n", argv[1]);
        return 2;
    }
    fclose(f);
    cv::Mat d1(n1, 128, CV_32FC1, buf.data()), d2(n2, 128, CV_32FC1, buf.data() + (size_t)n1 * 128);
    std::vector<cv::KeyPoint> kp1(n1), kp2(n2);

    objectsfm::FeatureMatchingB200 matcher;
    if (!matcher.ok()) {
        fprintf(stderr, "matcher: %s\n", matcher.last_error().c_str());
        return 3;
    }
    std::vector<std::pair<int, int>> matches;
    const bool ok = matcher.KNNMatching(kp1, d1, kp2, d2, matches);
    printf("KNNMatching %d %zu\n", ok ? 1 : 0, matches.size());
    for (auto &m : matches) printf("%d %d\n", m.first, m.second);

    std::vector<int> id((size_t)n2 * 2);
    std::vector<float> dis((size_t)n2 * 2);
    const bool ok2 = matcher.KNN2(d1, d2, id.data(), dis.data());
    printf("KNN2 %d %d\n", ok2 ? 1 : 0, n2);
    for (int i = 0; i < n2; ++i) printf("%d %d %.1f %.1f\n", id[2 * i], id[2 * i + 1], dis[2 * i], dis[2 * i + 1]);

    objectsfm::MatchGraphB200 graph(0, 2, n1 + n2);
    std::vector<std::vector<int>> init = {{1}, {0}};
    std::vector<std::vector<objectsfm::MatchGraphB200::PairMatches>> out;
    const bool ok3 = graph.ok() && graph.AddImage(0, d1) && graph.AddImage(1, d2) && graph.MatchPairs(init, out);
    printf("MatchPairs %d\n", ok3 ? 1 : 0);
    if (ok3)
        for (size_t i = 0; i < out.size(); ++i)
            for (auto &pm : out[i]) {
                printf("pair %zu ok %d n %zu\n", i, pm.ok ? 1 : 0, pm.matches_all.size());
                for (size_t k = 0; k < pm.matches_all.size(); ++k)
                    printf("%d %d %d\n", pm.matches_all[k].first, pm.matches_all[k].second, (int)pm.is_good[k]);
            }
    return 0;
}
