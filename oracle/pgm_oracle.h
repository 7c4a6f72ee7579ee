/*
 * oracle/pgm_oracle.h -- CPU restatement of the Photogrammetry descriptor matcher.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and there only as the checker or the
 * timed CPU baseline.  The product (libpgmatch.so) never links or calls it.
 *
 * What it restates (all paths relative to the upstream repository):
 *   dotnet_src/ImageProcessing/KeypointMatching.cs:14-69   MatchKeypoints
 *   dotnet_src/ImageProcessing/KeypointMatching.cs:71-82   CountOnes
 *   dotnet_src/ImageProcessing.Abstractions/Keypoint.cs:14,29-57  descriptor format
 *   python_src/photogrammetry/image_processing/keypoint_matching.py:7-40
 *
 * PARITY STATUS: the reference has no test, golden vector or known-answer
 * for this path, and its C# cannot be compiled here
 * (no dotnet/mono).  The oracle is pinned instead on
 *   (1) the two frozen descriptor sets the reference ships
 *       (data/feature_matching_test/lego_space_1_from_{left,right}_keypoints.dat),
 *   (2) distances produced by the reference's own Python `hamming_distance`
 *       and `match_keypoints`, imported from the reference tree by
 *       tests/golden/make_golden.py, and
 *   (3) an independent numpy restatement of the greedy loop (same script).
 * The greedy assignment itself (C#-only) is therefore "parity unpinned" in
 * the strict sense: it is checked against two independent restatements of
 * the source, not against an execution of the reference.
 *
 * Descriptor layout everywhere: uint8[n][stride], the little-endian bytes of
 * the non-negative BigInteger (BigInteger.ToByteArray(isUnsigned: true)),
 * zero padded to `stride` bytes.  Bit order inside the integer is irrelevant
 * to the Hamming distance.
 */
#ifndef PGM_ORACLE_H
#define PGM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_OK 0
#define ORC_E_INVALID_ARG (-1)
#define ORC_E_NOMEM (-2)
#define ORC_E_EMPTY_TRAIN (-5) /* keypoints2[0] on an empty list: KeypointMatching.cs:61 */

#define ORC_TAIL_DISTANCE 2147483647 /* int.MaxValue, KeypointMatching.cs:40 */

/* CountOnes(a XOR b), restated literally: Kernighan loop over a multi-limb
 * integer (KeypointMatching.cs:71-82). */
int orc_count_ones_kernighan(const uint8_t *a, const uint8_t *b, int stride);
/* Same value through the hardware popcount. */
int orc_hamming(const uint8_t *a, const uint8_t *b, int stride);

/* Full distance matrix, row-major int32[n1][n2] (KeypointMatching.cs:17-31;
 * keypoint_matching.py:8-13).  kernighan != 0 uses the literal CountOnes. */
int orc_distance_matrix(const uint8_t *q, int n1, const uint8_t *t, int n2,
                        int stride, int kernighan, int32_t *out);

/* MatchKeypoints restated literally: full matrix, then n1 global argmin
 * scans over the live rows x live columns in ascending index order with a
 * strict '<' update, including the degenerate (0,0,int.MaxValue) tail when
 * n1 > n2 (KeypointMatching.cs:33-68).  Writes exactly n1 triples.
 * Returns ORC_E_EMPTY_TRAIN when n1 > 0 and n2 == 0 (the reference throws). */
int orc_match_literal(const uint8_t *q, int n1, const uint8_t *t, int n2,
                      int stride, int kernighan,
                      int32_t *out_qi, int32_t *out_tj, int32_t *out_dist);

/* Same output through a counting-sort sweep over all (d,i,j) edges.  O(n1*n2)
 * time and memory.  Used to validate bigger cases than the literal loop can
 * reach (n up to ~16k). */
int orc_match_sweep(const uint8_t *q, int n1, const uint8_t *t, int n2,
                    int stride,
                    int32_t *out_qi, int32_t *out_tj, int32_t *out_dist);

/* Same output through iterated mutual-nearest-neighbour rounds, distances
 * recomputed every round, OpenMP over rows.  O(n1+n2) memory; the large-n
 * checker (200k x 200k).  *out_rounds (optional) receives the round count. */
int orc_match_rounds(const uint8_t *q, int n1, const uint8_t *t, int n2,
                     int stride,
                     int32_t *out_qi, int32_t *out_tj, int32_t *out_dist,
                     int32_t *out_rounds);

/* Per-query best and second-best train index under the (d, j) order.
 * Missing entries (n2 < 1 / n2 < 2) are -1.  No reference counterpart except
 * column 0/1 of keypoint_matching.py's sorted rows (up to tie order). */
int orc_knn2(const uint8_t *q, int n1, const uint8_t *t, int n2, int stride,
             int32_t *best_j, int32_t *best_d, int32_t *second_j, int32_t *second_d);

/* Ratio test + mutual cross-check on top of orc_knn2 (north_star extension;
 * NOT in the reference -- semantics defined here):
 *   keep (i, j1, d1) iff  [n2 < 2  or  (float)d1 < ratio * (float)d2]
 *                    and  [!cross_check or argmin_i' (d(i',j1), i') == i].
 * ratio <= 0 disables the ratio test.  Output ordered by i ascending.
 * max_dist >= 0 additionally requires d1 <= max_dist (the `dist <= 75`
 * filter of python_src/scripts/match_keypoints.py:126-128). */
int orc_match_ratio_crosscheck(const uint8_t *q, int n1, const uint8_t *t, int n2,
                               int stride, float ratio, int cross_check, int max_dist,
                               int32_t *out_qi, int32_t *out_tj, int32_t *out_dist,
                               int32_t *out_count);

/* keypoint_matching.py:7-33 restated: out int64[n1][n2][2] = (idx2, dist),
 * each row sorted by dist.  numpy's default argsort is not stable, so the
 * order among equal distances is unspecified upstream; the oracle uses the
 * stable (dist, idx2) order and tests compare modulo ties. */
int orc_python_twin(const uint8_t *q, int n1, const uint8_t *t, int n2,
                    int stride, int64_t *out);

/* Float-descriptor extension (north_star; NOT in the reference): exact
 * squared-L2 top-2, accumulated in double, ties by smaller j. */
int orc_l2_knn2(const float *q, int n1, const float *t, int n2, int dim,
                int32_t *best_j, float *best_d, int32_t *second_j, float *second_d);

/* Seeded synthetic descriptors.
 * splitmix64 stream; identical generator in photogrammetry_b200/synthetic.py. */
void orc_gen_uniform(uint64_t seed, int n, int desc_bits, int stride, uint8_t *out);
/* train = permuted copy of query, every bit flipped w.p. flip_p, then
 * outlier_p of the rows replaced by uniform noise. */
void orc_gen_noisy_copy(uint64_t seed, const uint8_t *query, int n, int desc_bits,
                        int stride, double flip_p, double outlier_p, uint8_t *out);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
