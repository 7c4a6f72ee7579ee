// pgmatch.hpp -- C++17 host-side mirror of the reference's matcher surface over the C ABI (pgmatch.h).
//
// The reference is managed code (C#); where its toolchain is absent the host side above the C ABI is written in
// C++ with the reference's names, argument meaning and error behaviour, so that a caller written against
//
//     ImageProcessing.KeypointMatching.MatchKeypoints(List<Keypoint>, List<Keypoint>) : List<KeypointPair>
//         (dotnet_src/ImageProcessing/KeypointMatching.cs:14-69)
//
// reads the same here.  Header-only; link with -lpgmatch.  The same marshalling is done by
// dotnet/GpuKeypointMatching.cs (P/Invoke) and photogrammetry_b200/keypoint_matching.py (ctypes).
//
//   Coordinate    dotnet_src/Math/LinearAlgebra/Coordinate.cs:3-17
//   Keypoint      dotnet_src/ImageProcessing.Abstractions/Keypoint.cs:9-15  (BriefDescriptor : BigInteger)
//   KeypointPair  dotnet_src/ImageProcessing.Abstractions/KeypointPair.cs:3-8
//
// BriefDescriptor is the little-endian byte string of the non-negative BigInteger
// (BigInteger.ToByteArray(isUnsigned: true, isBigEndian: false)); leading zero bytes may be omitted, exactly as
// .NET drops them.  There is no CPU fallback: the constructor throws when no sm_100 device is usable.
#pragma once

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "pgmatch.h"

namespace ImageProcessing {

struct Coordinate {
    int X = 0, Y = 0;
};

struct Keypoint {
    ::ImageProcessing::Coordinate Coordinate{};
    int FastScore = 0;
    std::vector<std::uint8_t> BriefDescriptor;      // little-endian magnitude bytes of the reference's BigInteger

    // number of significant bits (BigInteger.GetBitLength)
    int GetBitLength() const {
        for (std::size_t k = BriefDescriptor.size(); k-- > 0;)
            if (BriefDescriptor[k]) {
                int b = 8;
                while (!(BriefDescriptor[k] >> (b - 1))) b--;
                return (int)k * 8 + b;
            }
        return 0;
    }
};

// Holds references into the two input lists, like the C# record (KeypointMatching.cs:57-62).
struct KeypointPair {
    const Keypoint *Keypoint1 = nullptr;
    const Keypoint *Keypoint2 = nullptr;
    int Distance = 0;
};

class PgmatchError : public std::runtime_error {
public:
    PgmatchError(int status, const std::string &what) : std::runtime_error(what), Status(status) {}
    int Status;
};

class KeypointMatching {
public:
    // descBits: KeypointDetectionOptions.NumGaussianPairs (appsettings.json:23), 256 by default
    explicit KeypointMatching(int deviceOrdinal = 0, int descBits = 256) : descBits_(descBits) {
        const int rc = pgm_create(deviceOrdinal, &handle_);
        if (rc != PGM_OK) throw PgmatchError(rc, std::string("libpgmatch: ") + pgm_status_string(rc));
    }
    ~KeypointMatching() { if (handle_) pgm_destroy(handle_); }
    KeypointMatching(const KeypointMatching &) = delete;
    KeypointMatching &operator=(const KeypointMatching &) = delete;

    // Same contract as KeypointMatching.MatchKeypoints (KeypointMatching.cs:14): keypoints1.size() pairs in the
    // reference's (distance, i, j) order; when keypoints1 is longer than keypoints2 the remaining entries are
    // (keypoints1[0], keypoints2[0], int.MaxValue) (:38-42, :57-62); an empty keypoints2 with a non-empty keypoints1
    // throws std::out_of_range, the counterpart of the ArgumentOutOfRangeException raised at :61.
    std::vector<KeypointPair> MatchKeypoints(const std::vector<Keypoint> &keypoints1,
                                             const std::vector<Keypoint> &keypoints2) {
        const int n1 = (int)keypoints1.size(), n2 = (int)keypoints2.size();
        int bits = descBits_;
        for (const auto &k : keypoints1) bits = std::max(bits, k.GetBitLength());
        for (const auto &k : keypoints2) bits = std::max(bits, k.GetBitLength());
        const int stride = (bits + 127) / 128 * 16;
        const std::vector<std::uint8_t> q = Pack(keypoints1, stride), t = Pack(keypoints2, stride);
        std::vector<std::int32_t> qi(std::max(n1, 1)), tj(std::max(n1, 1)), dd(std::max(n1, 1));
        std::int32_t count = 0;
        const int rc = pgm_match_hamming_greedy(handle_, q.data(), n1, t.data(), n2, bits, stride, qi.data(), tj.data(),
                                                dd.data(), n1, &count, PGM_FLAG_REFERENCE_COMPAT_TAIL);
        if (rc == PGM_E_EMPTY_TRAIN) throw std::out_of_range("index");          // what keypoints2[0] throws upstream
        if (rc != PGM_OK) throw PgmatchError(rc, std::string("libpgmatch: ") + pgm_last_error(handle_));
        std::vector<KeypointPair> pairs;
        pairs.reserve((std::size_t)count);
        for (int k = 0; k < count; k++) pairs.push_back(KeypointPair{&keypoints1[qi[k]], &keypoints2[tj[k]], dd[k]});
        return pairs;
    }

private:
    static std::vector<std::uint8_t> Pack(const std::vector<Keypoint> &kps, int stride) {
        std::vector<std::uint8_t> out(std::max<std::size_t>(kps.size(), 1) * (std::size_t)stride, 0);
        for (std::size_t i = 0; i < kps.size(); i++) {
            const auto &d = kps[i].BriefDescriptor;
            std::size_t n = d.size();
            while (n > 0 && d[n - 1] == 0) n--;                                 // leading zero bytes carry no value
            if (n) std::memcpy(out.data() + i * (std::size_t)stride, d.data(), n);
        }
        return out;
    }

    pgm_handle *handle_ = nullptr;
    int descBits_;
};

}  // namespace ImageProcessing
