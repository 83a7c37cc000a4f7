// shim_selftest.cc — drives the C++ mirror of the reference interface exactly the way the reference would
// (cv::Mat CV_32FC1 in, vector<pair<int,int>> out) and prints the match lists so that tests/test_gpu_parity.py can
// check them.  Input: a raw file of two float32 matrices; output: text on stdout.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "feature_matching_b200.h"

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s <desc.f32> <rows1> <rows2> [<keypoints.f32: xy of image 1 then image 2>]\n", argv[0]);
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

    if (argc > 4) {  // keypoints given: the verified matcher of feature_matching.cpp:67-150
        std::vector<float> xy((size_t)(n1 + n2) * 2);
        FILE *fk = fopen(argv[4], "rb");
        if (!fk || fread(xy.data(), sizeof(float), xy.size(), fk) != xy.size()) {
            fprintf(stderr, "cannot read %s\n", argv[4]);
            return 2;
        }
        fclose(fk);
        for (int i = 0; i < n1; ++i) { kp1[i].pt.x = xy[2 * i]; kp1[i].pt.y = xy[2 * i + 1]; }
        for (int i = 0; i < n2; ++i) { kp2[i].pt.x = xy[2 * (n1 + i)]; kp2[i].pt.y = xy[2 * (n1 + i) + 1]; }
        {   // least-squares homography of the unverified ratio matches (the quantity the degeneracy gate looks at)
            std::vector<cv::Point2f> a, b;
            for (auto &m : matches) { a.push_back(kp1[m.first].pt); b.push_back(kp2[m.second].pt); }
            double Ha[9] = {0};
            const bool oka = objectsfm::FindHomographyDLT(a, b, Ha);
            printf("HomographyAll %d", oka ? 1 : 0);
            for (int i = 0; i < 9; ++i) printf(" %.9g", Ha[i]);
            printf("\n");
        }
        std::vector<std::pair<int, int>> verified;
        const bool okv = matcher.KNNMatchingWithGeoVerify(kp1, d1, kp2, d2, verified);
        printf("GeoVerify %d %zu\n", okv ? 1 : 0, verified.size());
        for (auto &m : verified) printf("%d %d\n", m.first, m.second);
        std::vector<cv::Point2f> p1, p2;
        for (auto &m : verified) { p1.push_back(kp1[m.first].pt); p2.push_back(kp2[m.second].pt); }
        double H[9] = {0};
        const bool okh = objectsfm::FindHomographyDLT(p1, p2, H);
        printf("Homography %d", okh ? 1 : 0);
        for (int i = 0; i < 9; ++i) printf(" %.9g", H[i]);
        printf("\n");
    }

    {   // FeatureMatchingCudaSift::Run-shaped entry (mutual best match) and the persistent-index overloads
        std::vector<std::pair<int, int>> rm;
        const bool okr = matcher.Run(kp1, d1, kp2, d2, rm);
        printf("Run %d %zu\n", okr ? 1 : 0, rm.size());
        for (auto &m : rm) printf("%d %d\n", m.first, m.second);
        objectsfm::KDIndexB200 *idx1 = nullptr, *idx2 = nullptr;
        const bool okg = matcher.GenerateKDIndex(d1, &idx1) && matcher.GenerateKDIndex(d2, &idx2);
        for (int rep = 0; rep < 2; ++rep) {  // the index is built once and queried again: nothing is re-uploaded for it
            std::vector<std::pair<int, int>> m1, m2;
            const bool ok1 = okg && matcher.MatchAgainstIndex(idx1, d2, true, m1);   // index on image 1, queries = image 2
            const bool ok2i = okg && matcher.MatchAgainstIndex(idx2, d1, false, m2);  // index on image 2, queries = image 1
            printf("Index1 %d %zu\n", ok1 ? 1 : 0, m1.size());
            for (auto &m : m1) printf("%d %d\n", m.first, m.second);
            printf("Index2 %d %zu\n", ok2i ? 1 : 0, m2.size());
            for (auto &m : m2) printf("%d %d\n", m.first, m.second);
        }
        matcher.ReleaseKDIndex(idx1);
        matcher.ReleaseKDIndex(idx2);
        // SiftMatchGPU-shaped front end: u8 rows, distmax off-by-default semantics checked with a wide bound
        objectsfm::SiftMatchB200 sm(4096);
        std::vector<unsigned char> u1((size_t)n1 * 128), u2((size_t)n2 * 128);
        for (size_t i = 0; i < u1.size(); ++i) u1[i] = (unsigned char)buf[i];
        for (size_t i = 0; i < u2.size(); ++i) u2[i] = (unsigned char)buf[(size_t)n1 * 128 + i];
        sm.SetDescriptors(0, n1, u1.data(), 7);
        sm.SetDescriptors(1, n2, u2.data(), 8);
        sm.SetDescriptors(1, n2, u1.data(), 8);  // same id: ignored, set 1 keeps image 2
        std::vector<int> mb((size_t)n1 * 2);
        const int full = sm.ok() ? sm.GetSiftMatch(n1, reinterpret_cast<int(*)[2]>(mb.data()), 3.2f, 0.8f, 1) : -1;
        printf("SiftMatch %d\n", full);
        for (int k = 0; k < full; ++k) printf("%d %d\n", mb[2 * k], mb[2 * k + 1]);
        const int capped = sm.ok() ? sm.GetSiftMatch(5, reinterpret_cast<int(*)[2]>(mb.data()), 3.2f, 0.8f, 0) : -1;
        printf("SiftMatchCapped %d\n", capped);
        for (int k = 0; k < capped; ++k) printf("%d %d\n", mb[2 * k], mb[2 * k + 1]);
        const int gated = sm.ok() ? sm.GetSiftMatch(n1, reinterpret_cast<int(*)[2]>(mb.data()), 0.2f, 0.8f, 0) : -1;
        printf("SiftMatchGated %d\n", gated);
        for (int k = 0; k < gated; ++k) printf("%d %d\n", mb[2 * k], mb[2 * k + 1]);
    }

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
