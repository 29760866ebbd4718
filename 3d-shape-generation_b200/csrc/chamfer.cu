// Chamfer nearest-neighbour core (metrics.py:7-47) as tiled shared-memory pairwise-distance
// kernels: per-cloud cube normalisation, directional NN search with register-resident
// queries and smem-staged targets (direct differences: exact fp32 distances, no
// |x|^2+|y|^2-2xy cancellation), sqrt deferred until after the min, deterministic means.
#include <cstdint>
#include <cuda_runtime.h>

namespace pcd {

// ------------------------------------------------------------------------------------------
// normalize_to_cube (metrics.py:7-21): centre = (max+min)/2 per axis, one scalar scale per
// cloud = max over points and axes of |p - centre|; output packed as float4 (w unused).
// One CTA per cloud.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(256) cloud_norm_kernel(const float* __restrict__ pts, int N, float4* __restrict__ out) {
    __shared__ float red[6][8];
    __shared__ float cen[3];
    __shared__ float sscale;
    const float* p = pts + static_cast<long long>(blockIdx.x) * N * 3;
    float4* o = out + static_cast<long long>(blockIdx.x) * N;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = tid; i < N; i += 256)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = p[i * 3 + c];
            mn[c] = fminf(mn[c], v); mx[c] = fmaxf(mx[c], v);
        }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float a = warp_min(mn[c]), b = warp_max(mx[c]);
        if (lane == 0) { red[c][warp] = a; red[3 + c][warp] = b; }
    }
    __syncthreads();
    if (tid < 3) {
        float a = red[tid][0], b = red[3 + tid][0];
        for (int w = 1; w < 8; ++w) { a = fminf(a, red[tid][w]); b = fmaxf(b, red[3 + tid][w]); }
        cen[tid] = __fdiv_rn(__fadd_rn(b, a), 2.0f);
    }
    __syncthreads();
    const float c0 = cen[0], c1 = cen[1], c2 = cen[2];
    float am = 0.f;
    for (int i = tid; i < N; i += 256)
        am = fmaxf(am, fmaxf(fabsf(__fsub_rn(p[i * 3], c0)), fmaxf(fabsf(__fsub_rn(p[i * 3 + 1], c1)), fabsf(__fsub_rn(p[i * 3 + 2], c2)))));
    am = warp_max(am);
    __syncthreads();
    if (lane == 0) red[0][warp] = am;
    __syncthreads();
    if (tid == 0) {
        float a = red[0][0];
        for (int w = 1; w < 8; ++w) a = fmaxf(a, red[0][w]);
        sscale = a;
    }
    __syncthreads();
    const float sc = sscale;   // 0 for a degenerate cloud -> NaN, as in the reference
    for (int i = tid; i < N; i += 256)
        o[i] = make_float4(__fdiv_rn(__fsub_rn(p[i * 3], c0), sc), __fdiv_rn(__fsub_rn(p[i * 3 + 1], c1), sc),
                           __fdiv_rn(__fsub_rn(p[i * 3 + 2], c2), sc), 0.f);
}

// ------------------------------------------------------------------------------------------
// Directional nearest neighbour: for every query point of cloud qi, min / argmin over all
// points of cloud ti.  Each thread keeps R queries in registers; targets stream through smem
// (one broadcast LDS.128 feeds R distance evaluations).  grid = (ceil(Nq/(128R)), pairs).
// pair -> (qi, ti): matrix mode (n_inner > 0) qi = pair / n_inner, ti = pair % n_inner (or swapped).
// ------------------------------------------------------------------------------------------
constexpr int kChamferTile = 1024;

template <int R, bool IDX>
__global__ void __launch_bounds__(128) chamfer_dir_kernel(const float4* __restrict__ Q, const float4* __restrict__ T, int Nq,
                                                          int Nt, float* __restrict__ mind, int* __restrict__ idx) {
    __shared__ float4 st[kChamferTile];
    const int pair = blockIdx.y;
    const float4* q = Q + static_cast<long long>(pair) * Nq;
    const float4* t = T + static_cast<long long>(pair) * Nt;
    float qx[R], qy[R], qz[R], best[R];
    int bi[R];
    const int q0 = blockIdx.x * (128 * R) + threadIdx.x;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int qi = q0 + r * 128;
        const float4 v = qi < Nq ? q[qi] : make_float4(0.f, 0.f, 0.f, 0.f);
        qx[r] = v.x; qy[r] = v.y; qz[r] = v.z;
        best[r] = 3.0e38f; bi[r] = 0;
    }
    for (int t0 = 0; t0 < Nt; t0 += kChamferTile) {
        const int cnt = min(kChamferTile, Nt - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += 128) st[i] = t[t0 + i];
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const float4 tv = st[j];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float dx = qx[r] - tv.x, dy = qy[r] - tv.y, dz = qz[r] - tv.z;
                const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                if (IDX) {
                    if (d2 < best[r]) { best[r] = d2; bi[r] = t0 + j; }   // strict <: first index wins ties
                } else {
                    best[r] = fminf(best[r], d2);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int qi = q0 + r * 128;
        if (qi < Nq) {
            // a degenerate cloud (all points equal) normalises to 0/0 = NaN everywhere (metrics.py:19-20) and
            // torch.min / mean propagate it; fminf would silently drop it, so re-inject it here
            const bool nan_in = (qx[r] != qx[r]) || (t[0].x != t[0].x);
            mind[static_cast<long long>(pair) * Nq + qi] = nan_in ? __int_as_float(0x7fc00000) : sqrtf(best[r]);   // L2, not squared (metrics.py:41)
            if (IDX) idx[static_cast<long long>(pair) * Nq + qi] = bi[r];
        }
    }
}

// cd[pair] = scaling * (mean_i min_j d + mean_j min_i d)   (metrics.py:43-47), fixed-order sums
__global__ void __launch_bounds__(256) chamfer_reduce_kernel(const float* __restrict__ dxy, const float* __restrict__ dyx, int N,
                                                             int M, float scaling, float* __restrict__ cd) {
    __shared__ float red[2][256];
    const int pair = blockIdx.x, tid = threadIdx.x;
    float a = 0.f, b = 0.f;
    for (int i = tid; i < N; i += 256) a += dxy[static_cast<long long>(pair) * N + i];
    for (int i = tid; i < M; i += 256) b += dyx[static_cast<long long>(pair) * M + i];
    red[0][tid] = a; red[1][tid] = b;
    __syncthreads();
    for (int s = 128; s; s >>= 1) {
        if (tid < s) { red[0][tid] += red[0][tid + s]; red[1][tid] += red[1][tid + s]; }
        __syncthreads();
    }
    if (tid == 0) cd[pair] = (red[0][0] / static_cast<float>(N) + red[1][0] / static_cast<float>(M)) * scaling;
}

// ------------------------------------------------------------------------------------------
// All-pairs matrix, one direction per launch: one CTA per (query cloud qi, target cloud ti),
// sum_i min_j |q_i - t_j| accumulated with a fixed-order block reduction (deterministic).
// swap = 0: queries from A[qi = pair / nB], targets from B[ti = pair % nB]  -> acc[pair]  = sum
// swap = 1: queries from B[ti],             targets from A[qi]              -> acc[pair] += sum, then scaled
// ------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256) chamfer_matrix_dir_kernel(const float4* __restrict__ A, const float4* __restrict__ B,
                                                                 int nB, int N, int swap, float scale_over_n,
                                                                 float* __restrict__ acc) {
    __shared__ float4 st[kChamferTile];
    __shared__ float red[256];
    const long long pair = blockIdx.x;
    const int ai = static_cast<int>(pair / nB), bi = static_cast<int>(pair % nB);
    const float4* q = swap ? B + static_cast<long long>(bi) * N : A + static_cast<long long>(ai) * N;
    const float4* t = swap ? A + static_cast<long long>(ai) * N : B + static_cast<long long>(bi) * N;
    float total = 0.f;
    for (int qbase = 0; qbase < N; qbase += 256 * R) {
        float qx[R], qy[R], qz[R], best[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int qi = qbase + r * 256 + threadIdx.x;
            const float4 v = qi < N ? q[qi] : make_float4(0.f, 0.f, 0.f, 0.f);
            qx[r] = v.x; qy[r] = v.y; qz[r] = v.z; best[r] = 3.0e38f;
        }
        for (int t0 = 0; t0 < N; t0 += kChamferTile) {
            const int cnt = min(kChamferTile, N - t0);
            __syncthreads();
            for (int i = threadIdx.x; i < cnt; i += 256) st[i] = t[t0 + i];
            __syncthreads();
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                const float4 tv = st[j];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float dx = qx[r] - tv.x, dy = qy[r] - tv.y, dz = qz[r] - tv.z;
                    best[r] = fminf(best[r], fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (qbase + r * 256 + threadIdx.x < N)
                total += ((qx[r] != qx[r]) || (t[0].x != t[0].x)) ? __int_as_float(0x7fc00000) : sqrtf(best[r]);
    }
    red[threadIdx.x] = total;
    __syncthreads();
    for (int s = 128; s; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (swap) acc[pair] = (acc[pair] + red[0]) * scale_over_n;
        else acc[pair] = red[0];
    }
}

cudaError_t launch_chamfer_matrix(const float4* G, int nG, const float4* Rc, int nR, int N, float scaling, float* out,
                                  cudaStream_t stream) {
    const long long pairs = static_cast<long long>(nG) * nR;
    if (pairs > 0x7fffffffLL) return cudaErrorInvalidValue;
    chamfer_matrix_dir_kernel<8><<<static_cast<unsigned>(pairs), 256, 0, stream>>>(G, Rc, nR, N, 0, 0.f, out);
    chamfer_matrix_dir_kernel<8><<<static_cast<unsigned>(pairs), 256, 0, stream>>>(G, Rc, nR, N, 1, scaling / static_cast<float>(N), out);
    return cudaGetLastError();
}

cudaError_t launch_cloud_norm(const float* pts, int clouds, int N, float4* out, cudaStream_t stream) {
    cloud_norm_kernel<<<clouds, 256, 0, stream>>>(pts, N, out);
    return cudaGetLastError();
}

cudaError_t launch_chamfer_dir(const float4* Q, const float4* T, int pairs, int Nq, int Nt, float* mind, int* idx,
                               cudaStream_t stream) {
    constexpr int R = 4;
    dim3 grid((Nq + 128 * R - 1) / (128 * R), pairs);
    if (idx) chamfer_dir_kernel<R, true><<<grid, 128, 0, stream>>>(Q, T, Nq, Nt, mind, idx);
    else chamfer_dir_kernel<R, false><<<grid, 128, 0, stream>>>(Q, T, Nq, Nt, mind, idx);
    return cudaGetLastError();
}

cudaError_t launch_chamfer_reduce(const float* dxy, const float* dyx, int pairs, int N, int M, float scaling, float* cd,
                                  cudaStream_t stream) {
    chamfer_reduce_kernel<<<pairs, 256, 0, stream>>>(dxy, dyx, N, M, scaling, cd);
    return cudaGetLastError();
}

}  // namespace pcd
