/*
 * oracle/pgm_oracle.c -- CPU restatement of the Photogrammetry descriptor
 * matcher.  TEST INFRASTRUCTURE ONLY (see pgm_oracle.h for the rules and
 * the parity status).  Build: `make -C oracle` -> oracle/liborc.so.
 *
 * Every function cites the upstream lines it follows.  Nothing here is
 * copied from upstream: the C# uses BigInteger/Dictionary/HashSet, this file
 * restates the arithmetic on packed byte rows and plain arrays.
 */
#include "pgm_oracle.h"

#include <limits.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* distances                                                            */
/* ------------------------------------------------------------------ */

/* KeypointMatching.cs:71-82 -- `while (value != 0) { value &= value - 1; n++ }`
 * on the arbitrary-precision XOR.  value-1 borrows through the zero low
 * limbs; ANDing clears exactly the lowest set bit and leaves the (zero) low
 * limbs zero, which is what the limb loop below does, one bit per iteration
 * like the original. */
int orc_count_ones_kernighan(const uint8_t *a, const uint8_t *b, int stride)
{
    uint64_t limb[64];
    int nl = (stride + 7) / 8, n = 0;
    if (nl > 64) return -1;
    for (int k = 0; k < nl; k++) {
        uint64_t va = 0, vb = 0;
        int nb = stride - 8 * k < 8 ? stride - 8 * k : 8;
        memcpy(&va, a + 8 * k, (size_t)nb);
        memcpy(&vb, b + 8 * k, (size_t)nb);
        limb[k] = va ^ vb; /* k1.BriefDescriptor ^ k2.BriefDescriptor, :28 */
    }
    for (;;) {
        int k = 0;
        while (k < nl && limb[k] == 0) k++;
        if (k == nl) break;           /* value == 0 */
        limb[k] &= limb[k] - 1;       /* value &= value - 1 */
        n++;                          /* numOnes += 1 */
    }
    return n;
}

/* keypoint_matching.py:38-40 -- bin(a ^ b).count("1") */
int orc_hamming(const uint8_t *a, const uint8_t *b, int stride)
{
    int n = 0, k = 0;
    for (; k + 8 <= stride; k += 8) {
        uint64_t va, vb;
        memcpy(&va, a + k, 8);
        memcpy(&vb, b + k, 8);
        n += __builtin_popcountll(va ^ vb);
    }
    for (; k < stride; k++) n += __builtin_popcount((unsigned)(a[k] ^ b[k]));
    return n;
}

static inline int ham_fast(const uint8_t *a, const uint8_t *b, int stride)
{
    if (stride == 32) {
        uint64_t x[4], y[4];
        memcpy(x, a, 32);
        memcpy(y, b, 32);
        return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) +
               __builtin_popcountll(x[2] ^ y[2]) + __builtin_popcountll(x[3] ^ y[3]);
    }
    return orc_hamming(a, b, stride);
}

/* KeypointMatching.cs:17-31 (k1ToK2ToDistance) / keypoint_matching.py:8-13 */
int orc_distance_matrix(const uint8_t *q, int n1, const uint8_t *t, int n2,
                        int stride, int kernighan, int32_t *out)
{
    if (n1 < 0 || n2 < 0 || stride <= 0 || stride > 512) return ORC_E_INVALID_ARG;
    for (int i = 0; i < n1; i++)
        for (int j = 0; j < n2; j++)
            out[(size_t)i * n2 + j] =
                kernighan ? orc_count_ones_kernighan(q + (size_t)i * stride, t + (size_t)j * stride, stride)
                          : ham_fast(q + (size_t)i * stride, t + (size_t)j * stride, stride);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* MatchKeypoints, literal                                              */
/* ------------------------------------------------------------------ */

/* KeypointMatching.cs:14-69.  The two HashSet<int> (:35-36) are built from
 * Enumerable.Range and only ever shrink, so they enumerate in ascending
 * order; they are kept here as ascending arrays of live indices. */
int orc_match_literal(const uint8_t *q, int n1, const uint8_t *t, int n2,
                      int stride, int kernighan,
                      int32_t *out_qi, int32_t *out_tj, int32_t *out_dist)
{
    if (n1 < 0 || n2 < 0 || stride <= 0 || stride > 512) return ORC_E_INVALID_ARG;
    if (n1 == 0) return ORC_OK;                      /* while (0 < 0) never runs */
    if (n2 == 0) return ORC_E_EMPTY_TRAIN;           /* keypoints2[0] throws, :61 */

    int32_t *D = (int32_t *)malloc((size_t)n1 * n2 * sizeof(int32_t));
    int32_t *a1 = (int32_t *)malloc((size_t)n1 * sizeof(int32_t));
    int32_t *a2 = (int32_t *)malloc((size_t)n2 * sizeof(int32_t));
    if (!D || !a1 || !a2) { free(D); free(a1); free(a2); return ORC_E_NOMEM; }
    orc_distance_matrix(q, n1, t, n2, stride, kernighan, D);   /* :20-31 */
    int m1 = n1, m2 = n2;
    for (int i = 0; i < n1; i++) a1[i] = i;                    /* :35 */
    for (int j = 0; j < n2; j++) a2[j] = j;                    /* :36 */

    for (int count = 0; count < n1; count++) {                 /* :38 */
        int smallest = INT_MAX, si = 0, sj = 0;                /* :40-42 */
        int pi = -1, pj = -1;
        for (int x = 0; x < m1; x++) {                         /* :44 */
            const int32_t *row = D + (size_t)a1[x] * n2;       /* :46 */
            for (int y = 0; y < m2; y++) {                     /* :47 */
                int d = row[a2[y]];
                if (smallest <= d) continue;                   /* :49-50 */
                smallest = d; si = a1[x]; sj = a2[y];          /* :51-53 */
                pi = x; pj = y;
            }
        }
        out_dist[count] = smallest;                            /* :57-62 */
        out_qi[count] = si;
        out_tj[count] = sj;
        if (pi >= 0) {                                         /* :64-65 */
            memmove(a1 + pi, a1 + pi + 1, (size_t)(m1 - pi - 1) * sizeof(int32_t)); m1--;
            memmove(a2 + pj, a2 + pj + 1, (size_t)(m2 - pj - 1) * sizeof(int32_t)); m2--;
        } else {
            /* nothing beat int.MaxValue: Remove(0) on both sets; removing row 0
             * or column 0 (if still live) cannot change any later output since
             * one of the sets is already empty. */
            if (m1 > 0 && a1[0] == 0) { memmove(a1, a1 + 1, (size_t)(m1 - 1) * sizeof(int32_t)); m1--; }
            if (m2 > 0 && a2[0] == 0) { memmove(a2, a2 + 1, (size_t)(m2 - 1) * sizeof(int32_t)); m2--; }
        }
    }
    free(D); free(a1); free(a2);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* MatchKeypoints through a sorted-edge sweep                           */
/* ------------------------------------------------------------------ */

/* Equivalent formulation of KeypointMatching.cs:38-66: retiring a row and a
 * column never reorders the remaining (d,i,j) triples, so the n-th global
 * argmin is the n-th edge of the lexicographically sorted edge list whose
 * endpoints are both still free.  d <= 8*stride, so a counting sort by d
 * that keeps row-major order inside a bucket is the full (d,i,j) order. */
int orc_match_sweep(const uint8_t *q, int n1, const uint8_t *t, int n2,
                    int stride,
                    int32_t *out_qi, int32_t *out_tj, int32_t *out_dist)
{
    if (n1 < 0 || n2 < 0 || stride <= 0 || stride > 512) return ORC_E_INVALID_ARG;
    if (n1 == 0) return ORC_OK;
    if (n2 == 0) return ORC_E_EMPTY_TRAIN;
    size_t ne = (size_t)n1 * n2;
    if (ne > 0xFFFFFFFFull) return ORC_E_INVALID_ARG;
    int nb = 8 * stride + 1;
    uint16_t *D = (uint16_t *)malloc(ne * sizeof(uint16_t));
    uint32_t *E = (uint32_t *)malloc(ne * sizeof(uint32_t));
    size_t *start = (size_t *)calloc((size_t)nb + 1, sizeof(size_t));
    uint8_t *used1 = (uint8_t *)calloc((size_t)n1, 1), *used2 = (uint8_t *)calloc((size_t)n2, 1);
    if (!D || !E || !start || !used1 || !used2) { free(D); free(E); free(start); free(used1); free(used2); return ORC_E_NOMEM; }

#pragma omp parallel for schedule(static)
    for (int i = 0; i < n1; i++)
        for (int j = 0; j < n2; j++)
            D[(size_t)i * n2 + j] = (uint16_t)ham_fast(q + (size_t)i * stride, t + (size_t)j * stride, stride);
    for (size_t e = 0; e < ne; e++) start[D[e] + 1]++;
    for (int b = 0; b < nb; b++) start[b + 1] += start[b];
    {
        size_t *fill = (size_t *)malloc((size_t)nb * sizeof(size_t));
        memcpy(fill, start, (size_t)nb * sizeof(size_t));
        for (size_t e = 0; e < ne; e++) E[fill[D[e]]++] = (uint32_t)e;  /* row-major inside a bucket */
        free(fill);
    }
    int m = n1 < n2 ? n1 : n2, count = 0;
    for (size_t k = 0; k < ne && count < m; k++) {
        uint32_t e = E[k];
        int i = (int)(e / (uint32_t)n2), j = (int)(e % (uint32_t)n2);
        if (used1[i] || used2[j]) continue;
        used1[i] = used2[j] = 1;
        out_qi[count] = i; out_tj[count] = j; out_dist[count] = D[e];
        count++;
    }
    for (; count < n1; count++) {            /* degenerate tail, :38-42,57-62 */
        out_qi[count] = 0; out_tj[count] = 0; out_dist[count] = ORC_TAIL_DISTANCE;
    }
    free(D); free(E); free(start); free(used1); free(used2);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* MatchKeypoints through mutual-nearest-neighbour rounds               */
/* ------------------------------------------------------------------ */

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

/* Third formulation of KeypointMatching.cs:38-66:
 * an edge that is the minimum under (d,i,j) among all edges touching its row
 * or its column is part of the greedy matching; accepting all such edges and
 * repeating on the remainder yields exactly the greedy matching.  Output
 * order is recovered by sorting the accepted triples by (d,i,j). */
int orc_match_rounds(const uint8_t *q, int n1, const uint8_t *t, int n2,
                     int stride,
                     int32_t *out_qi, int32_t *out_tj, int32_t *out_dist,
                     int32_t *out_rounds)
{
    if (n1 < 0 || n2 < 0 || stride <= 0 || stride > 512) return ORC_E_INVALID_ARG;
    if (out_rounds) *out_rounds = 0;
    if (n1 == 0) return ORC_OK;
    if (n2 == 0) return ORC_E_EMPTY_TRAIN;
    if (n1 >= (1 << 20) || n2 >= (1 << 20)) return ORC_E_INVALID_ARG;   /* 20-bit index packing below */
    int nthreads = orc_num_threads();
    int32_t *lr = (int32_t *)malloc((size_t)n1 * 4), *lc = (int32_t *)malloc((size_t)n2 * 4);
    uint64_t *rb = (uint64_t *)malloc((size_t)n1 * 8);
    uint64_t *cb = (uint64_t *)malloc((size_t)n2 * 8 * (size_t)nthreads);
    uint64_t *acc = (uint64_t *)malloc((size_t)(n1 < n2 ? n1 : n2) * 8);
    int32_t *mj = (int32_t *)malloc((size_t)n1 * 4);
    uint8_t *dead2 = (uint8_t *)calloc((size_t)n2, 1);
    if (!lr || !lc || !rb || !cb || !acc || !mj || !dead2) return ORC_E_NOMEM;
    int m1 = n1, m2 = n2, nacc = 0, rounds = 0;
    for (int i = 0; i < n1; i++) { lr[i] = i; mj[i] = -1; }
    for (int j = 0; j < n2; j++) lc[j] = j;

    while (m1 > 0 && m2 > 0) {
        rounds++;
#pragma omp parallel num_threads(nthreads)
        {
#ifdef _OPENMP
            int tid = omp_get_thread_num();
#else
            int tid = 0;
#endif
            uint64_t *mycb = cb + (size_t)tid * n2;
            for (int y = 0; y < m2; y++) mycb[lc[y]] = ~0ull;
#pragma omp for schedule(dynamic, 16)
            for (int x = 0; x < m1; x++) {
                int i = lr[x];
                const uint8_t *qi = q + (size_t)i * stride;
                uint64_t best = ~0ull;
                for (int y = 0; y < m2; y++) {
                    int j = lc[y];
                    uint64_t d = (uint64_t)ham_fast(qi, t + (size_t)j * stride, stride);
                    uint64_t rk = (d << 32) | (uint32_t)j;   /* (d, j) for fixed i */
                    uint64_t ck = (d << 32) | (uint32_t)i;   /* (d, i) for fixed j */
                    if (rk < best) best = rk;
                    if (ck < mycb[j]) mycb[j] = ck;
                }
                rb[i] = best;
            }
        }
        for (int y = 0; y < m2; y++) {
            int j = lc[y];
            uint64_t b = cb[j];
            for (int th = 1; th < nthreads; th++)
                if (cb[(size_t)th * n2 + j] < b) b = cb[(size_t)th * n2 + j];
            cb[j] = b;
        }
        int k1 = 0;
        for (int x = 0; x < m1; x++) {
            int i = lr[x];
            int j = (int)(uint32_t)rb[i];
            if ((int)(uint32_t)cb[j] == i) {                 /* mutual: accept */
                mj[i] = j; dead2[j] = 1;
                acc[nacc++] = ((rb[i] >> 32) << 40) | ((uint64_t)i << 20) | (uint64_t)j;
            } else lr[k1++] = i;
        }
        int k2 = 0;
        for (int y = 0; y < m2; y++) if (!dead2[lc[y]]) lc[k2++] = lc[y];
        m1 = k1; m2 = k2;
    }
    qsort(acc, (size_t)nacc, 8, cmp_u64);                     /* (d,i,j) order */
    int count = 0;
    for (; count < nacc; count++) {
        out_dist[count] = (int32_t)(acc[count] >> 40);
        out_qi[count] = (int32_t)((acc[count] >> 20) & 0xFFFFF);
        out_tj[count] = (int32_t)(acc[count] & 0xFFFFF);
    }
    for (; count < n1; count++) { out_qi[count] = 0; out_tj[count] = 0; out_dist[count] = ORC_TAIL_DISTANCE; }
    if (out_rounds) *out_rounds = rounds;
    free(lr); free(lc); free(rb); free(cb); free(acc); free(mj); free(dead2);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* nearest / second nearest, ratio test, cross check                    */
/* ------------------------------------------------------------------ */

int orc_knn2(const uint8_t *q, int n1, const uint8_t *t, int n2, int stride,
             int32_t *best_j, int32_t *best_d, int32_t *second_j, int32_t *second_d)
{
    if (n1 < 0 || n2 < 0 || stride <= 0 || stride > 512) return ORC_E_INVALID_ARG;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n1; i++) {
        uint64_t b = ~0ull, s = ~0ull;
        for (int j = 0; j < n2; j++) {
            uint64_t k = ((uint64_t)ham_fast(q + (size_t)i * stride, t + (size_t)j * stride, stride) << 32) | (uint32_t)j;
            if (k < b) { s = b; b = k; } else if (k < s) s = k;
        }
        best_j[i] = b == ~0ull ? -1 : (int32_t)(uint32_t)b;
        best_d[i] = b == ~0ull ? -1 : (int32_t)(b >> 32);
        second_j[i] = s == ~0ull ? -1 : (int32_t)(uint32_t)s;
        second_d[i] = s == ~0ull ? -1 : (int32_t)(s >> 32);
    }
    return ORC_OK;
}

int orc_match_ratio_crosscheck(const uint8_t *q, int n1, const uint8_t *t, int n2,
                               int stride, float ratio, int cross_check, int max_dist,
                               int32_t *out_qi, int32_t *out_tj, int32_t *out_dist,
                               int32_t *out_count)
{
    if (n1 < 0 || n2 < 0 || stride <= 0 || stride > 512) return ORC_E_INVALID_ARG;
    *out_count = 0;
    if (n1 == 0 || n2 == 0) return ORC_OK;
    int32_t *bj = (int32_t *)malloc((size_t)n1 * 4 * 4), *bd = bj + n1, *sj = bd + n1, *sd = sj + n1;
    uint64_t *cb = (uint64_t *)malloc((size_t)n2 * 8);
    if (!bj || !cb) { free(bj); free(cb); return ORC_E_NOMEM; }
    orc_knn2(q, n1, t, n2, stride, bj, bd, sj, sd);
    for (int j = 0; j < n2; j++) {
        uint64_t b = ~0ull;
        for (int i = 0; i < n1; i++) {
            uint64_t k = ((uint64_t)ham_fast(q + (size_t)i * stride, t + (size_t)j * stride, stride) << 32) | (uint32_t)i;
            if (k < b) b = k;
        }
        cb[j] = b;
    }
    int c = 0;
    for (int i = 0; i < n1; i++) {
        if (ratio > 0.0f && sj[i] >= 0 && !((float)bd[i] < ratio * (float)sd[i])) continue;
        if (cross_check && (int32_t)(uint32_t)cb[bj[i]] != i) continue;
        if (max_dist >= 0 && bd[i] > max_dist) continue;
        out_qi[c] = i; out_tj[c] = bj[i]; out_dist[c] = bd[i]; c++;
    }
    *out_count = c;
    free(bj); free(cb);
    return ORC_OK;
}

/* keypoint_matching.py:7-33 */
int orc_python_twin(const uint8_t *q, int n1, const uint8_t *t, int n2,
                    int stride, int64_t *out)
{
    if (n1 < 0 || n2 < 0 || stride <= 0 || stride > 512) return ORC_E_INVALID_ARG;
    uint64_t *keys = (uint64_t *)malloc((size_t)(n2 > 0 ? n2 : 1) * 8);
    if (!keys) return ORC_E_NOMEM;
    for (int i = 0; i < n1; i++) {
        for (int j = 0; j < n2; j++)   /* :9-13 */
            keys[j] = ((uint64_t)ham_fast(q + (size_t)i * stride, t + (size_t)j * stride, stride) << 32) | (uint32_t)j;
        qsort(keys, (size_t)n2, 8, cmp_u64);   /* :28-31, stable-by-construction */
        for (int j = 0; j < n2; j++) {
            out[((size_t)i * n2 + j) * 2 + 0] = (int64_t)(uint32_t)keys[j];
            out[((size_t)i * n2 + j) * 2 + 1] = (int64_t)(keys[j] >> 32);
        }
    }
    free(keys);
    return ORC_OK;
}

int orc_l2_knn2(const float *q, int n1, const float *t, int n2, int dim,
                int32_t *best_j, float *best_d, int32_t *second_j, float *second_d)
{
    if (n1 < 0 || n2 < 0 || dim <= 0) return ORC_E_INVALID_ARG;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n1; i++) {
        double bd = 0, sd = 0; int bj = -1, sj = -1;
        for (int j = 0; j < n2; j++) {
            double acc = 0;
            for (int k = 0; k < dim; k++) {
                double df = (double)q[(size_t)i * dim + k] - (double)t[(size_t)j * dim + k];
                acc += df * df;
            }
            if (bj < 0 || acc < bd) { sd = bd; sj = bj; bd = acc; bj = j; }
            else if (sj < 0 || acc < sd) { sd = acc; sj = j; }
        }
        best_j[i] = bj; best_d[i] = bj < 0 ? -1.0f : (float)bd;
        second_j[i] = sj; second_d[i] = sj < 0 ? -1.0f : (float)sd;
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* synthetic descriptors                                                */
/* ------------------------------------------------------------------ */

#define GAMMA 0x9E3779B97F4A7C15ull
static inline uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* counter-based splitmix64: the idx-th output of the stream seeded `seed` */
static inline uint64_t sm64(uint64_t seed, uint64_t idx) { return mix64(seed + (idx + 1) * GAMMA); }

static void put_row(uint8_t *row, int desc_bits, int stride, uint64_t seed, uint64_t r)
{
    int nw = (desc_bits + 63) / 64;
    memset(row, 0, (size_t)stride);
    for (int w = 0; w < nw; w++) {
        uint64_t v = sm64(seed, r * (uint64_t)nw + (uint64_t)w);
        int bits = desc_bits - 64 * w;
        if (bits < 64) v &= (1ull << bits) - 1;
        int nb = stride - 8 * w < 8 ? stride - 8 * w : 8;
        for (int b = 0; b < nb; b++) row[8 * w + b] = (uint8_t)(v >> (8 * b));
    }
}

void orc_gen_uniform(uint64_t seed, int n, int desc_bits, int stride, uint8_t *out)
{
    for (int r = 0; r < n; r++) put_row(out + (size_t)r * stride, desc_bits, stride, seed, (uint64_t)r);
}

typedef struct { uint64_t key; int32_t idx; } perm_t;
static int cmp_perm(const void *a, const void *b)
{
    const perm_t *x = (const perm_t *)a, *y = (const perm_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : x->idx > y->idx;
}

void orc_gen_noisy_copy(uint64_t seed, const uint8_t *query, int n, int desc_bits,
                        int stride, double flip_p, double outlier_p, uint8_t *out)
{
    uint64_t s_perm = mix64(seed + 1 * GAMMA), s_flip = mix64(seed + 2 * GAMMA);
    uint64_t s_out = mix64(seed + 3 * GAMMA), s_noise = mix64(seed + 4 * GAMMA);
    perm_t *p = (perm_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(perm_t));
    for (int r = 0; r < n; r++) { p[r].key = sm64(s_perm, (uint64_t)r); p[r].idx = r; }
    qsort(p, (size_t)n, sizeof(perm_t), cmp_perm);
    unsigned thr = (unsigned)(flip_p * 256.0);
    int nbytes = (desc_bits + 7) / 8;                 /* one random byte per descriptor bit */
    int words_per_row = (desc_bits + 7) / 8;          /* u64 draws per row: 8 bits each */
    for (int r = 0; r < n; r++) {
        uint8_t *row = out + (size_t)r * stride;
        double u = (double)(sm64(s_out, (uint64_t)r) >> 11) * (1.0 / 9007199254740992.0);
        if (u < outlier_p) { put_row(row, desc_bits, stride, s_noise, (uint64_t)r); continue; }
        memcpy(row, query + (size_t)p[r].idx * stride, (size_t)stride);
        for (int by = 0; by < nbytes; by++) {
            uint64_t v = sm64(s_flip, (uint64_t)r * (uint64_t)words_per_row + (uint64_t)by);
            uint8_t m = 0;
            for (int b = 0; b < 8; b++) {
                if (8 * by + b >= desc_bits) break;
                if (((v >> (8 * b)) & 0xFF) < thr) m |= (uint8_t)(1u << b);
            }
            row[by] ^= m;
        }
    }
    free(p);
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
