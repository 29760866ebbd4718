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

// ------------------------------------------------------------------------------------------
// Fused pair kernel (values only): ONE pass over the Na x Nb distance matrix of a cloud pair feeds
// both directional minima, so every distance is evaluated once instead of twice.
//   * both clouds live in shared memory as SoA (padded to 128 with far-away sentinels),
//   * 256 threads = 16 x 16; a thread owns an 8-query x 8-target register block per 128 x 128 tile:
//     64 direct-difference distances update 8 row minima (registers, live across the target loop)
//     and 8 column minima (registers, reduced over the two query groups of the warp by one shuffle,
//     then merged into a shared column-min array with atomicMin on the float bits, d2 >= 0),
//   * row minima are reduced over the 16 target groups with shuffles; sums use fixed-order trees.
// pair -> (ai, bi) = (pair / nB, pair % nB) in matrix mode, (pair, pair) in pair mode (nB == 0).
// ------------------------------------------------------------------------------------------
constexpr float kFar = 1.0e18f;   // sentinel coordinate: (kFar - x)^2 ~ 1e36 < FLT_MAX, never a minimum

__global__ void __launch_bounds__(256) chamfer_fused_kernel(const float4* __restrict__ A, const float4* __restrict__ B, int nB,
                                                            int Na, int Nb, float scaling, float* __restrict__ out) {
    extern __shared__ float sm[];
    const int Nap = (Na + 127) & ~127, Nbp = (Nb + 127) & ~127;
    float* ax = sm; float* ay = ax + Nap; float* az = ay + Nap;
    float* bx = az + Nap; float* by = bx + Nbp; float* bz = by + Nbp;
    unsigned* cminb = reinterpret_cast<unsigned*>(bz + Nbp);
    __shared__ float red[2][256];
    const long long pair = blockIdx.x;
    const long long ai = nB > 0 ? pair / nB : pair, bi = nB > 0 ? pair % nB : pair;
    const float4* a = A + ai * Na;
    const float4* b = B + bi * Nb;
    const int tid = threadIdx.x;
    for (int i = tid; i < Nap; i += 256) {
        const float4 v = i < Na ? a[i] : make_float4(kFar, kFar, kFar, 0.f);
        ax[i] = v.x; ay[i] = v.y; az[i] = v.z;
    }
    for (int i = tid; i < Nbp; i += 256) {
        const float4 v = i < Nb ? b[i] : make_float4(-kFar, -kFar, -kFar, 0.f);
        bx[i] = v.x; by[i] = v.y; bz[i] = v.z;
        cminb[i] = 0x7f7fffffu;   // FLT_MAX
    }
    __syncthreads();
    const bool nan_in = (ax[0] != ax[0]) || (bx[0] != bx[0]);   // degenerate cloud -> NaN (metrics.py:19-20)
    const int ty = tid >> 4, tx = tid & 15, lane = tid & 31;
    float rowsum = 0.f;
    for (int q0 = ty * 8; q0 < Nap; q0 += 128) {
        float qx[8], qy[8], qz[8], rmin[8];
        {
            const float4 x0 = *reinterpret_cast<const float4*>(ax + q0), x1 = *reinterpret_cast<const float4*>(ax + q0 + 4);
            const float4 y0 = *reinterpret_cast<const float4*>(ay + q0), y1 = *reinterpret_cast<const float4*>(ay + q0 + 4);
            const float4 z0 = *reinterpret_cast<const float4*>(az + q0), z1 = *reinterpret_cast<const float4*>(az + q0 + 4);
            qx[0] = x0.x; qx[1] = x0.y; qx[2] = x0.z; qx[3] = x0.w; qx[4] = x1.x; qx[5] = x1.y; qx[6] = x1.z; qx[7] = x1.w;
            qy[0] = y0.x; qy[1] = y0.y; qy[2] = y0.z; qy[3] = y0.w; qy[4] = y1.x; qy[5] = y1.y; qy[6] = y1.z; qy[7] = y1.w;
            qz[0] = z0.x; qz[1] = z0.y; qz[2] = z0.z; qz[3] = z0.w; qz[4] = z1.x; qz[5] = z1.y; qz[6] = z1.z; qz[7] = z1.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) rmin[i] = 3.0e38f;
#pragma unroll 1
        for (int t0 = tx * 8; t0 < Nbp; t0 += 128) {
            float tx8[8], ty8[8], tz8[8], cmin[8];
            {
                const float4 x0 = *reinterpret_cast<const float4*>(bx + t0), x1 = *reinterpret_cast<const float4*>(bx + t0 + 4);
                const float4 y0 = *reinterpret_cast<const float4*>(by + t0), y1 = *reinterpret_cast<const float4*>(by + t0 + 4);
                const float4 z0 = *reinterpret_cast<const float4*>(bz + t0), z1 = *reinterpret_cast<const float4*>(bz + t0 + 4);
                tx8[0] = x0.x; tx8[1] = x0.y; tx8[2] = x0.z; tx8[3] = x0.w; tx8[4] = x1.x; tx8[5] = x1.y; tx8[6] = x1.z; tx8[7] = x1.w;
                ty8[0] = y0.x; ty8[1] = y0.y; ty8[2] = y0.z; ty8[3] = y0.w; ty8[4] = y1.x; ty8[5] = y1.y; ty8[6] = y1.z; ty8[7] = y1.w;
                tz8[0] = z0.x; tz8[1] = z0.y; tz8[2] = z0.z; tz8[3] = z0.w; tz8[4] = z1.x; tz8[5] = z1.y; tz8[6] = z1.z; tz8[7] = z1.w;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) cmin[j] = 3.0e38f;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float dx = qx[i] - tx8[j], dy = qy[i] - ty8[j], dz = qz[i] - tz8[j];
                    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    rmin[i] = fminf(rmin[i], d2);
                    cmin[j] = fminf(cmin[j], d2);
                }
#pragma unroll
            for (int j = 0; j < 8; ++j) cmin[j] = fminf(cmin[j], __shfl_xor_sync(0xffffffffu, cmin[j], 16));
            if (lane < 16) {
#pragma unroll
                for (int j = 0; j < 8; ++j) atomicMin(&cminb[t0 + j], __float_as_uint(cmin[j]));
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int o = 8; o; o >>= 1) rmin[i] = fminf(rmin[i], __shfl_xor_sync(0xffffffffu, rmin[i], o));
            if (tx == 0 && q0 + i < Na) rowsum += sqrtf(rmin[i]);
        }
    }
    __syncthreads();
    float colsum = 0.f;
    for (int j = tid; j < Nb; j += 256) colsum += sqrtf(__uint_as_float(cminb[j]));
    red[0][tid] = rowsum; red[1][tid] = colsum;
    __syncthreads();
    for (int s2 = 128; s2; s2 >>= 1) {
        if (tid < s2) { red[0][tid] += red[0][tid + s2]; red[1][tid] += red[1][tid + s2]; }
        __syncthreads();
    }
    if (tid == 0) {
        const float v = (red[0][0] / static_cast<float>(Na) + red[1][0] / static_cast<float>(Nb)) * scaling;
        out[pair] = nan_in ? __int_as_float(0x7fc00000) : v;
    }
}

static size_t chamfer_fused_smem(int Na, int Nb) {
    const size_t Nap = (Na + 127) & ~127, Nbp = (Nb + 127) & ~127;
    return (3 * Nap + 4 * Nbp) * sizeof(float);
}

// true if the fused kernel can hold both clouds in shared memory
bool chamfer_fused_fits(int Na, int Nb) { return chamfer_fused_smem(Na, Nb) <= 200 * 1024; }

cudaError_t launch_chamfer_fused(const float4* A, const float4* B, long long pairs, int nB, int Na, int Nb, float scaling,
                                 float* out, cudaStream_t stream) {
    if (pairs > 0x7fffffffLL) return cudaErrorInvalidValue;
    const size_t smem = chamfer_fused_smem(Na, Nb);
    cudaError_t e = cudaFuncSetAttribute(chamfer_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    chamfer_fused_kernel<<<static_cast<unsigned>(pairs), 256, smem, stream>>>(A, B, nB, Na, Nb, scaling, out);
    return cudaGetLastError();
}

cudaError_t launch_chamfer_matrix(const float4* G, int nG, const float4* Rc, int nR, int N, float scaling, float* out,
                                  cudaStream_t stream) {
    const long long pairs = static_cast<long long>(nG) * nR;
    if (pairs > 0x7fffffffLL) return cudaErrorInvalidValue;
    if (chamfer_fused_fits(N, N)) return launch_chamfer_fused(G, Rc, pairs, nR, N, N, scaling, out, stream);
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
