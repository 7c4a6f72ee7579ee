// pgm_ransac.cuh -- hypothesis scoring of the fundamental-matrix RANSAC, the consumer of the match list
// (SURVEY.md section 8 row f3): the inner loops of CameraPoseEstimation.GetFundamentalMatrix
// (dotnet_src/ImageProcessing/CameraPoseEstimation.cs:41-88).
//
// For every hypothesis F (row-major 3x3 float) and every pair (Keypoint1 = (x1, y1), Keypoint2 = (x2, y2)):
//     result = (F . (x2, y2, 1)) . (x1, y1, 1)              :66-71
//     inlier  <=>  result <= threshold                      :73 (signed, as upstream -- no absolute value)
// and the winner is the first hypothesis whose inlier count exceeds every earlier one (:79-84); a hypothesis
// flagged invalid (upstream: rank != 2, :46-51) is skipped.  Arithmetic is float32 in MathNet.Numerics'
// (5.0.0) managed order -- each 3-term sum accumulated left to right, no fused multiply-add.
// Trivially parallel integer/float work: one CTA per hypothesis, coordinates streamed from L2.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace pgm_ransac {

__device__ __forceinline__ float residual(const float (&f)[9], float x1, float y1, float x2, float y2) {
    // F.Multiply(kp1Mat): r_i = (f[i][0]*x2 + f[i][1]*y2) + f[i][2]*1   -- __f*_rn forbids FMA contraction
    const float r0 = __fadd_rn(__fadd_rn(__fmul_rn(f[0], x2), __fmul_rn(f[1], y2)), f[2]);
    const float r1 = __fadd_rn(__fadd_rn(__fmul_rn(f[3], x2), __fmul_rn(f[4], y2)), f[5]);
    const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(f[6], x2), __fmul_rn(f[7], y2)), f[8]);
    // .DotProduct(kp2Mat)
    return __fadd_rn(__fadd_rn(__fmul_rn(r0, x1), __fmul_rn(r1, y1)), r2);
}

__global__ void score_kernel(const float *__restrict__ F, const uint8_t *__restrict__ valid, int n_hyp,
                             const int32_t *__restrict__ xy1, const int32_t *__restrict__ xy2, int n, float threshold,
                             int32_t *__restrict__ counts) {
    const int hyp = blockIdx.x;
    if (hyp >= n_hyp) return;
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int c = 0;
    if (!valid || valid[hyp]) {
        float f[9];
#pragma unroll
        for (int k = 0; k < 9; k++) f[k] = __ldg(F + (size_t)hyp * 9 + k);
        for (int p = threadIdx.x; p < n; p += blockDim.x) {
            const int2 a = __ldg(reinterpret_cast<const int2 *>(xy1) + p), b = __ldg(reinterpret_cast<const int2 *>(xy2) + p);
            c += residual(f, (float)a.x, (float)a.y, (float)b.x, (float)b.y) <= threshold;
        }
    }
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) counts[hyp] = (!valid || valid[hyp]) ? s_cnt : -1;
}

// first index attaining the maximum count (> 0), else -1; then the winner's inlier mask
__global__ void best_kernel(const float *__restrict__ F, const int32_t *__restrict__ counts, int n_hyp,
                            const int32_t *__restrict__ xy1, const int32_t *__restrict__ xy2, int n, float threshold,
                            int32_t *__restrict__ best, uint8_t *__restrict__ mask) {
    __shared__ unsigned long long s_key;
    if (threadIdx.x == 0) s_key = 0ull;
    __syncthreads();
    unsigned long long k = 0ull;
    for (int hI = threadIdx.x; hI < n_hyp; hI += blockDim.x) {
        const int c = counts[hI];
        if (c > 0) {   // larger count wins, then the smaller index
            const unsigned long long key = ((unsigned long long)c << 32) | (unsigned)(0x7FFFFFFF - hI);
            k = key > k ? key : k;
        }
    }
    atomicMax(&s_key, k);
    __syncthreads();
    const int b = s_key ? 0x7FFFFFFF - (int)(s_key & 0xFFFFFFFFull) : -1;
    if (threadIdx.x == 0) *best = b;
    if (mask) {
        float f[9];
#pragma unroll
        for (int q = 0; q < 9; q++) f[q] = b >= 0 ? F[(size_t)b * 9 + q] : 0.f;
        for (int p = threadIdx.x; p < n; p += blockDim.x)
            mask[p] = b >= 0 && residual(f, (float)xy1[2 * p], (float)xy1[2 * p + 1], (float)xy2[2 * p], (float)xy2[2 * p + 1]) <= threshold;
    }
}

}  // namespace pgm_ransac
