// C++ caller of include/pgmatch.hpp, written like a caller of the reference's C# class.
// argv[1] = case file written by tests/test_cpp_mirror.py:
//   int32 n1, n2, nbytes; then n1*nbytes + n2*nbytes descriptor bytes (little-endian BigInteger bytes);
//   int32 count; then count * 3 int32 expected (i, j, distance) triples in the reference's order.
// Exit codes: 0 ok, 2 no device (printed), anything else = failure.
#include <cstdio>
#include <fstream>
#include <iostream>
#include "pgmatch.hpp"

using namespace ImageProcessing;

int main(int argc, char **argv) {
    if (argc < 2) return 64;
    std::ifstream f(argv[1], std::ios::binary);
    std::int32_t n1 = 0, n2 = 0, nb = 0, count = 0;
    f.read((char *)&n1, 4); f.read((char *)&n2, 4); f.read((char *)&nb, 4);
    std::vector<Keypoint> k1((std::size_t)n1), k2((std::size_t)n2);
    for (auto *v : {&k1, &k2})
        for (auto &k : *v) { k.BriefDescriptor.resize((std::size_t)nb); f.read((char *)k.BriefDescriptor.data(), nb); }
    f.read((char *)&count, 4);
    std::vector<std::int32_t> exp((std::size_t)count * 3);
    f.read((char *)exp.data(), (std::streamsize)exp.size() * 4);
    if (!f) return 65;
    try {
        KeypointMatching matching;                                   // new KeypointMatching()
        const auto pairs = matching.MatchKeypoints(k1, k2);          // List<KeypointPair>
        if ((int)pairs.size() != count || (int)pairs.size() != n1) return 10;
        for (int k = 0; k < count; k++) {
            if (pairs[k].Keypoint1 != &k1[exp[3 * k]] || pairs[k].Keypoint2 != &k2[exp[3 * k + 1]] ||
                pairs[k].Distance != exp[3 * k + 2]) {
                std::printf("mismatch at %d\n", k);
                return 11;
            }
        }
        if (n1 > n2 && n2 > 0 &&
            (pairs.back().Keypoint1 != &k1[0] || pairs.back().Keypoint2 != &k2[0] || pairs.back().Distance != INT_MAX))
            return 12;                                               // KeypointMatching.cs:38-42, 57-62
        bool threw = false;
        try { matching.MatchKeypoints(k1, {}); } catch (const std::out_of_range &) { threw = true; }   // :61
        if (!threw && n1 > 0) return 13;
        if (!matching.MatchKeypoints({}, k2).empty()) return 14;
        std::puts("ok");
        return 0;
    } catch (const PgmatchError &e) {
        if (e.Status == PGM_E_NO_DEVICE) { std::puts("no-device"); return 2; }
        std::printf("error %d: %s\n", e.Status, e.what());
        return 20;
    }
}
