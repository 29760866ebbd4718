// Micro-benchmark: issue rates of the FP32 instruction forms the Chamfer kernels are built from (packed FADD2 / FMUL2 / FFMA2,
// scalar FADD / FMUL / FFMA, three-input FMNMX3) alone and in the kernel's mix, on one SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_fp32 tools/ubench_fp32.cu ; ./tools/ubench_fp32
// Prints warp-instructions per clock per SM sub-partition (SMSP) and the FP32 lane-operations per clock per SM they amount to
// (peak = 128).  Timing: clock64() inside one CTA per SM, W warps per SMSP.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int kIters = 2048;

template <int MODE>
__global__ void __launch_bounds__(1024) bench(float* out, long long* cycles, float seed) {
    // 8 query rows x (2 targets packed): the register picture of chamfer_fused_kernel's inner block
    float2 qx[8], qy[8], qz[8];
    float2 tx[4], ty[4], tz[4];
    float rmin[8], cmin[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        qx[i] = make_float2(seed + i + threadIdx.x, seed + i + 1); qy[i] = make_float2(seed * i, seed - i); qz[i] = make_float2(seed * 3 + i, seed);
        rmin[i] = 3e38f; cmin[i] = 3e38f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { tx[j] = make_float2(seed * j, seed + j); ty[j] = make_float2(seed - j, seed * 2 + j); tz[j] = make_float2(seed + 2 * j, j); }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
        if (MODE == 0) {          // scalar FFMA only: 48 per iteration
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                qx[i].x = fmaf(qx[i].x, tx[0].x, ty[0].x); qx[i].y = fmaf(qx[i].y, tx[1].x, ty[1].x);
                qy[i].x = fmaf(qy[i].x, tx[2].x, ty[2].x); qy[i].y = fmaf(qy[i].y, tx[3].x, ty[3].x);
                qz[i].x = fmaf(qz[i].x, tx[0].y, ty[0].y); qz[i].y = fmaf(qz[i].y, tx[1].y, ty[1].y);
            }
        } else if (MODE == 1) {   // FFMA2 only: 24 per iteration
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                qx[i] = __ffma2_rn(qx[i], tx[i & 3], ty[i & 3]);
                qy[i] = __ffma2_rn(qy[i], ty[i & 3], tz[i & 3]);
                qz[i] = __ffma2_rn(qz[i], tz[i & 3], tx[i & 3]);
            }
        } else if (MODE == 2) {   // FADD2 only: 24 per iteration
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                qx[i] = __fadd2_rn(qx[i], tx[i & 3]);
                qy[i] = __fadd2_rn(qy[i], ty[i & 3]);
                qz[i] = __fadd2_rn(qz[i], tz[i & 3]);
            }
        } else if (MODE == 3) {   // FMUL2 only (x * x form): 24 per iteration
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                qx[i] = __fmul2_rn(qx[i], qx[i]);
                qy[i] = __fmul2_rn(qy[i], qy[i]);
                qz[i] = __fmul2_rn(qz[i], qz[i]);
            }
        } else if (MODE == 4 || MODE == 5) {   // the kernel's block: 8 rows x 4 target pairs, packed; MODE 5 without the minima
#pragma unroll
            for (int i = 0; i < 8; i += 2)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 dx = __fadd2_rn(qx[i], tx[j]), dy = __fadd2_rn(qy[i], ty[j]), dz = __fadd2_rn(qz[i], tz[j]);
                    const float2 d = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                    const float2 ex = __fadd2_rn(qx[i + 1], tx[j]), ey = __fadd2_rn(qy[i + 1], ty[j]), ez = __fadd2_rn(qz[i + 1], tz[j]);
                    const float2 e = __ffma2_rn(ez, ez, __ffma2_rn(ey, ey, __fmul2_rn(ex, ex)));
                    if (MODE == 4) {
                        rmin[i] = fminf(rmin[i], fminf(d.x, d.y));
                        rmin[i + 1] = fminf(rmin[i + 1], fminf(e.x, e.y));
                        cmin[2 * j] = fminf(cmin[2 * j], fminf(d.x, e.x));
                        cmin[2 * j + 1] = fminf(cmin[2 * j + 1], fminf(d.y, e.y));
                    } else {
                        rmin[i] += d.x + d.y; rmin[i + 1] += e.x + e.y;     // keeps the results live with 4 scalar adds per 4 distances... counted
                    }
                }
            tx[0].x += 1.0f;      // the targets change every step
        } else if (MODE == 6) {   // scalar form of the block: 64 distances = 192 FADD + 64 FMUL + 128 FFMA + 64 FMNMX3
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float dx0 = qx[i].x + tx[j].x, dy0 = qy[i].x + ty[j].x, dz0 = qz[i].x + tz[j].x;
                    const float dx1 = qx[i].x + tx[j].y, dy1 = qy[i].x + ty[j].y, dz1 = qz[i].x + tz[j].y;
                    const float d0 = fmaf(dz0, dz0, fmaf(dy0, dy0, dx0 * dx0)), d1 = fmaf(dz1, dz1, fmaf(dy1, dy1, dx1 * dx1));
                    rmin[i] = fminf(rmin[i], fminf(d0, d1));
                    cmin[2 * j] = fminf(cmin[2 * j], d0);
                    cmin[2 * j + 1] = fminf(cmin[2 * j + 1], d1);
                }
            tx[0].x += 1.0f;
        } else if (MODE == 7) {   // hybrid: differences scalar (FADD), squares packed
#pragma unroll
            for (int i = 0; i < 8; i += 2)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 dx, dy, dz, ex, ey, ez;
                    dx.x = qx[i].x + tx[j].x; dx.y = qx[i].y + tx[j].y; dy.x = qy[i].x + ty[j].x; dy.y = qy[i].y + ty[j].y;
                    dz.x = qz[i].x + tz[j].x; dz.y = qz[i].y + tz[j].y;
                    ex.x = qx[i + 1].x + tx[j].x; ex.y = qx[i + 1].y + tx[j].y; ey.x = qy[i + 1].x + ty[j].x; ey.y = qy[i + 1].y + ty[j].y;
                    ez.x = qz[i + 1].x + tz[j].x; ez.y = qz[i + 1].y + tz[j].y;
                    const float2 d = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                    const float2 e = __ffma2_rn(ez, ez, __ffma2_rn(ey, ey, __fmul2_rn(ex, ex)));
                    rmin[i] = fminf(rmin[i], fminf(d.x, d.y));
                    rmin[i + 1] = fminf(rmin[i + 1], fminf(e.x, e.y));
                    cmin[2 * j] = fminf(cmin[2 * j], fminf(d.x, e.x));
                    cmin[2 * j + 1] = fminf(cmin[2 * j + 1], fminf(d.y, e.y));
                }
            tx[0].x += 1.0f;
        } else if (MODE == 8) {   // FMNMX3 only: 32 per iteration
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                rmin[i] = fminf(rmin[i], fminf(qx[i].x, qy[i].x));
                cmin[i] = fminf(cmin[i], fminf(qx[i].y, qy[i].y));
                qx[i].x = fminf(qx[i].x, fminf(qz[i].x, tx[i & 3].x));
                qy[i].y = fminf(qy[i].y, fminf(qz[i].y, tx[i & 3].y));
            }
        }
    }
    const long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += qx[i].x + qx[i].y + qy[i].x + qy[i].y + qz[i].x + qz[i].y + rmin[i] + cmin[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double fp_inst_per_iter, double lane_ops_per_iter, double other_per_iter, int warps_per_smsp) {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int threads = 128 * warps_per_smsp;
    float* out; long long* cyc;
    CK(cudaMalloc(&out, sizeof(float) * sms * threads));
    CK(cudaMalloc(&cyc, sizeof(long long) * sms));
    bench<MODE><<<sms, threads>>>(out, cyc, 0.5f);
    bench<MODE><<<sms, threads>>>(out, cyc, 0.5f);
    CK(cudaDeviceSynchronize());
    long long h[256];
    CK(cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += static_cast<double>(h[i]);
    mean /= sms;
    const double per_iter = mean / kIters;                         // cycles per iteration with W warps per SMSP interleaved
    printf("{\"mode\": \"%s\", \"warps_per_smsp\": %d, \"cycles_per_iter_per_warp\": %.2f, \"fp_inst_per_clk_smsp\": %.3f, "
           "\"all_inst_per_clk_smsp\": %.3f, \"fp32_lane_ops_per_clk_sm\": %.1f}\n",
           name, warps_per_smsp, per_iter / warps_per_smsp, fp_inst_per_iter * warps_per_smsp / per_iter,
           (fp_inst_per_iter + other_per_iter) * warps_per_smsp / per_iter, lane_ops_per_iter * 32 * warps_per_smsp * 4 / per_iter);
    CK(cudaFree(out)); CK(cudaFree(cyc));
}

int main() {
    for (int w : {1, 2, 4, 6}) {
        run<0>("ffma_scalar", 48, 48, 0, w);
        run<1>("ffma2", 24, 48, 0, w);
        run<2>("fadd2", 24, 48, 0, w);
        run<3>("fmul2_xx", 24, 48, 0, w);
        run<8>("fmnmx3", 0, 0, 32, w);
        run<5>("block_packed_no_min", 192 + 64, 384 + 64, 0, w);
        run<4>("block_packed", 192, 384, 64, w);
        run<6>("block_scalar", 384, 384, 96, w);
        run<7>("block_hybrid_fadd_scalar", 192 + 96, 384, 64, w);
    }
    return 0;
}
