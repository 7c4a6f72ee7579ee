// pgm_detect.cuh -- the producer of the matcher's inputs on the GPU:
//
//   FAST-12 corner test      dotnet_src/ImageProcessing/KeypointDetection.cs:42-138
//   BRIEF descriptor         dotnet_src/ImageProcessing.Abstractions/Keypoint.cs:29-57
//   non-maximum suppression  dotnet_src/ImageProcessing/RedundantKeypointEliminator.cs:16-39
//
// Images are float32 [H][W] arrays of Grayscale.K (Grayscale.cs:19-23), pixel (x, y) at img[y*W + x].
// All three stages reproduce the reference bit for bit, quirks included (the ring table's last entry
// repeats {-3, 1}; a fifth in-threshold ring pixel rejects the candidate; out-of-image BRIEF pairs give
// a 0 bit; NMS drops distance <= radius).  These kernels are HBM/latency-bound byte work: coalesced
// loads, no tensor cores.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace pgm_det {

// Ring tables as (dx, dy).  [0] = KeypointDetection.cs:15-19; entry 15 is upstream's typo ({-3, 1} twice), kept.
// [1] = the Python generation's BRESENHAM_CIRCLE_3 (python_src/.../keypoint_detection.py:12-29), whose entries
// are (height_offset, width_offset) = (dy, dx) and whose entry 15 is the correct {-3, -1}.
__constant__ int c_ring[2][16][2] = {
    {{-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}, {0, 3}, {1, 3}, {2, 2}, {3, 1},
     {3, 0}, {3, -1}, {2, -2}, {1, -3}, {0, -3}, {-1, -3}, {-2, -2}, {-3, 1}},
    {{0, -3}, {1, -3}, {2, -2}, {3, -1}, {3, 0}, {3, 1}, {2, 2}, {1, 3},
     {0, 3}, {-1, 3}, {-2, 2}, {-3, 1}, {-3, 0}, {-3, -1}, {-2, -2}, {-1, -3}}};

// score map: 0 = no keypoint, else the longest circular run (12..16) = Keypoint.FastScore
template <int RING>
__global__ void fast_score_kernel(const float *__restrict__ img, int w, int h, float threshold,
                                  uint8_t *__restrict__ score) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    img += (size_t)blockIdx.z * w * h; score += (size_t)blockIdx.z * w * h;      // batched form: blockIdx.z = image
    uint8_t out = 0;
    if (x >= 3 && x < w - 3 && y >= 3 && y < h - 3) {                 // Detect's loop bounds, :45-47
        const float *p = img + (size_t)y * w + x;
        const float c = __ldg(p);
        const float lo = c - threshold, hi = c + threshold;            // InThreshold, :135-138
        unsigned inside = 0;                                           // bit k: ring pixel k is inside the threshold
        // IsPotentialKeypoint (:116-133): the compass points are ring entries 0, 4, 8, 12; at most one may be
        // inside.  Most pixels of a natural image fail here and never touch the other twelve taps.
#pragma unroll
        for (int k = 0; k < 16; k += 4) {
            const float t = __ldg(p + c_ring[RING][k][1] * w + c_ring[RING][k][0]);
            inside |= (unsigned)(t > lo && t < hi) << k;
        }
        if (__popc(inside) <= 1) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if ((k & 3) == 0) continue;
                const float t = __ldg(p + c_ring[RING][k][1] * w + c_ring[RING][k][0]);
                inside |= (unsigned)(t > lo && t < hi) << k;
            }
            // GetIntensityValueIfKeypoint (:65-114): longest circular run of ring pixels OUTSIDE the threshold,
            // accepted from 12 up.  (Its early exit on a fifth inside pixel cannot change the outcome: five
            // inside pixels leave at most eleven outside.)  On the doubled ring a circular run is a linear run;
            // bit i of `r` survives the shifts iff bits i .. i+11 are all set.
            const unsigned o16 = ~inside & 0xFFFFu;
            if (o16 == 0xFFFFu) out = 16;
            else {
                unsigned r = o16 | (o16 << 16);
                r &= r >> 1; r &= r >> 2; r &= r >> 4; r &= r >> 4;
                if (r) {
                    int len = 12;
                    while ((r &= r >> 1) != 0u) len++;
                    out = (uint8_t)len;
                }
            }
        }
    }
    score[(size_t)y * w + x] = out;
}

// per-row keypoint counts (one block per row)
__global__ void row_count_kernel(const uint8_t *__restrict__ score, int w, int32_t *__restrict__ rowcnt) {
    const int y = blockIdx.x;
    score += (size_t)blockIdx.y * w * gridDim.x; rowcnt += (size_t)blockIdx.y * gridDim.x;   // batched form: blockIdx.y = image
    int c = 0;
    for (int x = threadIdx.x; x < w; x += blockDim.x) c += score[(size_t)y * w + x] != 0;
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s, c);
    __syncthreads();
    if (threadIdx.x == 0) rowcnt[y] = s;
}

// exclusive scan of n <= 65536 ints by one block of 1024 threads; total written to out[n]
__global__ void scan_kernel(const int32_t *__restrict__ in, int n, int32_t *__restrict__ out) {
    __shared__ int s_w[32];
    in += (size_t)blockIdx.x * n; out += (size_t)blockIdx.x * (n + 1);          // batched form: blockIdx.x = image
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (n + 1023) / 1024, p0 = min(tid * per, n), p1 = min(p0 + per, n);
    int sum = 0;
    for (int k = p0; k < p1; k++) sum += in[k];
    int inc = sum;
    for (int o = 1; o < 32; o <<= 1) { const int a = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += a; }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    int off = inc - sum, tot = 0;
    for (int k = 0; k < 32; k++) { if (k < wid) off += s_w[k]; tot += s_w[k]; }
    for (int k = p0; k < p1; k++) { out[k] = off; off += in[k]; }
    if (tid == 0) out[n] = tot;
}

// ordered emission of one row (one warp per row, ballot prefix): Detect's row-major list order
__global__ void emit_kernel(const uint8_t *__restrict__ score, int w, int h, const int32_t *__restrict__ rowoff,
                            int capacity, int32_t *__restrict__ xy, int32_t *__restrict__ sc) {
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (y >= h) return;
    // batched form: blockIdx.y = image; every image owns `capacity` output slots
    score += (size_t)blockIdx.y * w * h; rowoff += (size_t)blockIdx.y * (h + 1);
    xy += (size_t)blockIdx.y * capacity * 2; sc += (size_t)blockIdx.y * capacity;
    int base = rowoff[y];
    for (int x0 = 0; x0 < w; x0 += 32) {
        const int x = x0 + lane;
        const uint8_t v = x < w ? score[(size_t)y * w + x] : 0;
        const unsigned m = __ballot_sync(0xffffffffu, v != 0);
        if (v) {
            const int pos = base + __popc(m & ((1u << lane) - 1u));
            if (pos < capacity) { xy[2 * pos] = x; xy[2 * pos + 1] = y; sc[pos] = v; }
        }
        base += __popc(m);
    }
}

// BRIEF: one warp per keypoint.  pair p sets bit (n_pairs - 1 - p) of the descriptor integer
// (`descriptor <<= 1` per pair, first pair = MSB), stored little-endian in `stride` bytes.
__global__ void brief_kernel(const float *__restrict__ img, int w, int h, const int32_t *__restrict__ xy, int n,
                             const int32_t *__restrict__ pairs /*[n_pairs][4]*/, int n_pairs, int stride_words,
                             int lsb_first, uint32_t *__restrict__ desc, const int32_t *__restrict__ rowoff = nullptr) {
    extern __shared__ uint32_t s_desc[];           // [warps][stride_words]
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + wid;
    if (rowoff) {                                  // batched form: blockIdx.y = image, n = capacity per image, and the
        const int z = blockIdx.y;                  // image's keypoint count sits behind its row offsets on the device
        const int found = __ldg(rowoff + (size_t)z * (h + 1) + h);
        img += (size_t)z * w * h; xy += (size_t)z * n * 2; desc += (size_t)z * n * stride_words;
        n = min(n, found);
    }
    uint32_t *mine = s_desc + wid * stride_words;
    for (int k = lane; k < stride_words; k += 32) mine[k] = 0u;
    __syncwarp();
    if (i < n) {
        const int x = xy[2 * i], y = xy[2 * i + 1];
        for (int p = lane; p < n_pairs; p += 32) {
            const int4 pr = __ldg(reinterpret_cast<const int4 *>(pairs) + p);
            const int x1 = x + pr.x, y1 = y + pr.y, x2 = x + pr.z, y2 = y + pr.w;
            bool bit = false;
            if (x1 >= 0 && x1 < w && y1 >= 0 && y1 < h && x2 >= 0 && x2 < w && y2 >= 0 && y2 < h)   // Keypoint.cs:39-45
                bit = __ldg(img + (size_t)y1 * w + x1) < __ldg(img + (size_t)y2 * w + x2);          // :47-53
            const int pos = lsb_first ? p : n_pairs - 1 - p;   // models/keypoint.py:49 `des += 2**idx` vs Keypoint.cs:36 `<<= 1`
            if (bit) atomicOr(&mine[pos >> 5], 1u << (pos & 31));
        }
        __syncwarp();
        for (int k = lane; k < stride_words; k += 32) desc[(size_t)i * stride_words + k] = mine[k];
    }
}

// ---- NMS ------------------------------------------------------------------------------------
// rank[i] = position of keypoint i in the stable order by score descending (LINQ OrderByDescending is
// stable, :20): the number of keypoints with a larger score plus the earlier ones with the same score.
__global__ void nms_rank_kernel(const int32_t *__restrict__ sc, int n, int32_t *__restrict__ rank,
                                int32_t *__restrict__ order) {
    __shared__ int s_s[256];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int mine = i < n ? sc[i] : 0;
    int r = 0;
    for (int j0 = 0; j0 < n; j0 += 256) {
        __syncthreads();
        if (j0 + (int)threadIdx.x < n) s_s[threadIdx.x] = sc[j0 + threadIdx.x];
        __syncthreads();
        const int lim = min(256, n - j0);
        for (int k = 0; k < lim; k++) r += (s_s[k] > mine) || (s_s[k] == mine && j0 + k < i);
    }
    if (i < n) { rank[i] = r; order[r] = i; }
}

// One round of the parallel form of the sequential suppression loop (:23-32).  In rank order, a keypoint
// is kept iff no KEPT keypoint of smaller rank lies within the radius.  state: 0 undecided, 1 kept, 2 dropped.
// An undecided keypoint is dropped as soon as a kept higher-priority neighbour exists, kept once every
// higher-priority neighbour is dropped; repeating to a fixed point gives exactly the sequential answer.
__global__ void nms_round_kernel(const int32_t *__restrict__ xy, const int32_t *__restrict__ rank, int n,
                                 long long radius2, const uint8_t *__restrict__ state_in, uint8_t *__restrict__ state_out,
                                 int32_t *__restrict__ n_undecided) {
    __shared__ int s_x[256], s_y[256], s_r[256];
    __shared__ uint8_t s_s[256];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int x = 0, y = 0, r = 0;
    uint8_t st = 2;
    if (i < n) { x = xy[2 * i]; y = xy[2 * i + 1]; r = rank[i]; st = state_in[i]; }
    bool any_kept = false, any_undecided = false;
    for (int j0 = 0; j0 < n; j0 += 256) {
        const int j = j0 + threadIdx.x;
        __syncthreads();
        if (j < n) { s_x[threadIdx.x] = xy[2 * j]; s_y[threadIdx.x] = xy[2 * j + 1]; s_r[threadIdx.x] = rank[j]; s_s[threadIdx.x] = state_in[j]; }
        __syncthreads();
        if (st == 0) {
            const int lim = min(256, n - j0);
            for (int k = 0; k < lim; k++) {
                if (s_r[k] >= r || s_s[k] == 2) continue;
                const long long dx = s_x[k] - x, dy = s_y[k] - y;
                if (dx * dx + dy * dy > radius2) continue;          // IsAcceptableDistance: distance > radius survives
                if (s_s[k] == 1) any_kept = true; else any_undecided = true;
            }
        }
    }
    if (i < n) {
        uint8_t ns = st;
        if (st == 0) {
            if (any_kept) ns = 2;
            else if (!any_undecided) ns = 1;
            if (ns == 0) atomicAdd(n_undecided, 1);
        }
        state_out[i] = ns;
    }
}

// ---- spatially binned form of the same rounds -------------------------------------------------------
// Cells of side max(radius, 1): every keypoint within `radius` of (x, y) lies in the 3 x 3 cells around it.
// Keypoints are sorted by cell id (64-bit radix sort); a round then visits only the 9 neighbouring runs
// instead of all n keypoints, so the cost is O(n . neighbours) per round instead of O(n^2).
__global__ void nms_cell_kernel(const int32_t *__restrict__ xy, int n, int cs, int minx, int miny, long long ncx,
                                unsigned long long *__restrict__ cell, int32_t *__restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long cx = ((long long)xy[2 * i] - minx) / cs, cy = ((long long)xy[2 * i + 1] - miny) / cs;
    cell[i] = (unsigned long long)(cy * ncx + cx);
    idx[i] = i;
}

// key[i] = (inverted score << 32) | i: an ascending sort of the keys is the stable order by score descending
__global__ void nms_score_key_kernel(const int32_t *__restrict__ sc, int n, unsigned long long *__restrict__ key) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) key[i] = ((unsigned long long)(~((unsigned)sc[i] ^ 0x80000000u)) << 32) | (unsigned)i;
}

// order[r] = keypoint of rank r, rank[order[r]] = r
__global__ void nms_order_from_keys_kernel(const unsigned long long *__restrict__ sorted_key, int n,
                                           int32_t *__restrict__ order, int32_t *__restrict__ rank) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int i = (int)(unsigned)sorted_key[r];
    order[r] = i; rank[i] = r;
}

__device__ __forceinline__ int lower_bound_u64(const unsigned long long *__restrict__ a, int n, unsigned long long key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void nms_round_binned_kernel(const int32_t *__restrict__ xy, const int32_t *__restrict__ rank, int n,
                                        long long radius2, int cs, int minx, int miny, long long ncx, long long ncy,
                                        const unsigned long long *__restrict__ sorted_cell,
                                        const int32_t *__restrict__ sorted_idx, const uint8_t *__restrict__ state_in,
                                        uint8_t *__restrict__ state_out, int32_t *__restrict__ n_undecided) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t st = state_in[i];
    if (st != 0) { state_out[i] = st; return; }
    const int x = xy[2 * i], y = xy[2 * i + 1], r = rank[i];
    const long long cx = ((long long)x - minx) / cs, cy = ((long long)y - miny) / cs;
    bool any_kept = false, any_undecided = false;
    for (int dy = -1; dy <= 1 && !any_kept; dy++) {
        const long long yy = cy + dy;
        if (yy < 0 || yy >= ncy) continue;
        // the three cells (cx-1 .. cx+1) of one cell row are consecutive ids: one search, one run
        const long long x0 = cx > 0 ? cx - 1 : 0, x1 = cx + 1 < ncx ? cx + 1 : ncx - 1;
        const unsigned long long c0 = (unsigned long long)(yy * ncx + x0), c1 = (unsigned long long)(yy * ncx + x1);
        for (int k = lower_bound_u64(sorted_cell, n, c0); k < n && sorted_cell[k] <= c1; k++) {
            const int j = sorted_idx[k];
            if (rank[j] >= r) continue;
            const uint8_t sj = state_in[j];
            if (sj == 2) continue;
            const long long ddx = (long long)xy[2 * j] - x, ddy = (long long)xy[2 * j + 1] - y;
            if (ddx * ddx + ddy * ddy > radius2) continue;       // IsAcceptableDistance: distance > radius survives
            if (sj == 1) { any_kept = true; break; }
            any_undecided = true;
        }
    }
    uint8_t ns = 0;
    if (any_kept) ns = 2;
    else if (!any_undecided) ns = 1;
    if (ns == 0) atomicAdd(n_undecided, 1);
    state_out[i] = ns;
}

// gathers (xy, score) of the kept keypoints, in output order
__global__ void gather_kept_kernel(const int32_t *__restrict__ xy, const int32_t *__restrict__ sc,
                                   const int32_t *__restrict__ kept, int n_kept, int32_t *__restrict__ oxy,
                                   int32_t *__restrict__ osc) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_kept) return;
    const int i = kept[k];
    oxy[2 * k] = xy[2 * i]; oxy[2 * k + 1] = xy[2 * i + 1]; osc[k] = sc[i];
}

// kept keypoints in rank order = the reference's output order (acceptableList, :21-26)
__global__ void nms_emit_kernel(const int32_t *__restrict__ order, const uint8_t *__restrict__ state, int n,
                                int32_t *__restrict__ kept, int32_t *__restrict__ n_kept) {
    // single block, ordered compaction
    __shared__ int s_w[32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nt = blockDim.x;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int r0 = 0; r0 < n; r0 += nt) {
        const int r = r0 + tid;
        const int i = r < n ? order[r] : 0;
        const bool hit = r < n && state[i] == 1;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_w[wid] = __popc(m);
        __syncthreads();
        int off = s_base;
        for (int k = 0; k < wid; k++) off += s_w[k];
        if (hit) kept[off + __popc(m & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (tid == 0) { int t = 0; for (int k = 0; k < (nt >> 5); k++) t += s_w[k]; s_base += t; }
        __syncthreads();
    }
    if (tid == 0) *n_kept = s_base;
}

}  // namespace pgm_det
