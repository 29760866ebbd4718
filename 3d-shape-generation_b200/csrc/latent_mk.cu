// Persistent kernel for the latent-diffusion path (BASELINE config 4): the WHOLE reverse loop of
// LatentDiffusion.sample / sample2 / sample3 (reference diffusion.py:575-707) over SimpleLatentUNetPointNet
// (networks.py:962-1106), and SimplePointNetVAE.decode (networks.py:1144-1154, 1219-1231), as one cooperative launch.
//
// Why a persistent kernel: rows are samples (M = B = 128 per GPU at config 4), so a reverse step is 4.9 GFLOP over 76 MB of
// L2-resident fp32 weights -- a few microseconds of math behind ~40 dependent launches.  The CUDA-graph version spent
// 336 us per step almost entirely on launch/drain latency.  Here one CTA per SM walks a small "program" of phases; phases are
// separated by a grid barrier (one atomic + one polled load), nothing returns to the host between the S steps.
//
// Arithmetic: every Linear is a 128 x 64 x K tile job on the warp-level tensor path with the 3xTF32 split
// (x = hi + lo, both TF32; D += lo*hi + hi*lo + hi*hi, fp32 accumulate): 2^-21 relative per product, i.e. the results
// sit inside the fp32 parity bound (2e-5 per forward) that the CUDA-core path was held to, at 1/3 of the TF32 rate instead
// of the FP32-FMA rate.  tcgen05 is deliberately not used here: the operands are fp32 in global memory, would have to be
// split into 16-bit planes first (2x the weight bytes: no longer L2 resident) and the path is latency bound, not tensor bound.
//
// Split-K partial sums are written to an fp32 workspace and reduced in a FIXED order by the GroupNorm phase, and the split
// count depends on the layer shape only, so a row's result does not depend on the batch it is in (sharded == unsharded).
#include "pcd_sampler.cuh"
#include "pcd_types.h"

namespace pcd {

namespace {

constexpr int BM = 128, BN = 64, BK = 32, NST = 4;
constexpr int A_TILE = BM * BK, W_TILE = BN * BK, STAGE = A_TILE + W_TILE;   // floats

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// 16-byte L2-only async copy; src_bytes = 0 zero-fills (rows past the last sample)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// element (row, k) of a [rows][32] fp32 tile whose 16-byte chunks are XOR-swizzled by the row: fragment loads
// (8 rows x 4 consecutive k) hit 32 distinct banks
__device__ __forceinline__ int sw_idx(int row, int k) { return row * BK + ((((k >> 2) ^ row) & 7) << 2) + (k & 3); }

__device__ __forceinline__ void grid_sync(unsigned* counter, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

struct StepCtx {
    int step;       // reverse-loop step (0 in the prologue and in forward mode)
    int forward;    // 1: per-sample t, eps written to eps_out; 0: sampler
};

__device__ __forceinline__ const float* bias_row(const LtOp& op, int row, const StepCtx& cx) {
    if (op.bias_mode == 0) return op.bias;
    return op.bias + static_cast<long long>(cx.forward ? row : cx.step) * op.bias_ld;
}

// one (m_tile, n_tile, split) job: acc[128 x 64] = A[m0.., k-range] * W[n0.., k-range]^T
__device__ void gemm_item(const LtOp& op, const LatentCall& c, const StepCtx& cx, int rows, int m_tile, int n_tile, int split,
                          float* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp & 3, wn = warp >> 2;
    const int m0 = m_tile * BM, n0 = n_tile * BN;
    const int nchunks = op.chunks_per_split;
    const int kbase = split * nchunks * BK;

    auto issue = [&](int ci) {
        float* sA = smem + (ci % NST) * STAGE;
        float* sW = sA + A_TILE;
        const int kg = kbase + ci * BK;
        const float* src;
        int ld;
        if (kg < op.K0) { src = (op.A0 ? op.A0 : c.z) + kg; ld = op.lda0; }
        else { src = op.A1 + (kg - op.K0); ld = op.lda1; }
#pragma unroll
        for (int i = 0; i < (A_TILE / 4) / 256; ++i) {
            const int ch = tid + i * 256, row = ch >> 3, cc = ch & 7;
            const int gr = m0 + row;
            const bool ok = gr < rows;
            cp_async16(smem_u32(sA + row * BK + ((cc ^ (row & 7)) << 2)), src + static_cast<long long>(ok ? gr : 0) * ld + cc * 4,
                       ok ? 16 : 0);
        }
        const float* wsrc = op.W + static_cast<long long>(n0) * op.ldw + kg;
#pragma unroll
        for (int i = 0; i < (W_TILE / 4) / 256; ++i) {
            const int ch = tid + i * 256, row = ch >> 3, cc = ch & 7;
            cp_async16(smem_u32(sW + row * BK + ((cc ^ (row & 7)) << 2)), wsrc + static_cast<long long>(row) * op.ldw + cc * 4, 16);
        }
    };

    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

#pragma unroll
    for (int s = 0; s < NST - 1; ++s) {
        if (s < nchunks) issue(s);
        cp_async_commit();
    }
    for (int ci = 0; ci < nchunks; ++ci) {
        cp_async_wait<NST - 2>();
        __syncthreads();
        if (ci + NST - 1 < nchunks) issue(ci + NST - 1);
        cp_async_commit();
        const float* sA = smem + (ci % NST) * STAGE;
        const float* sW = sA + A_TILE;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 8) {
            uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                const int r = wm * 32 + mi * 16 + g;
                split_tf32(sA[sw_idx(r, kk + t)], ah[mi][0], al[mi][0]);
                split_tf32(sA[sw_idx(r + 8, kk + t)], ah[mi][1], al[mi][1]);
                split_tf32(sA[sw_idx(r, kk + t + 4)], ah[mi][2], al[mi][2]);
                split_tf32(sA[sw_idx(r + 8, kk + t + 4)], ah[mi][3], al[mi][3]);
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int n = wn * 32 + ni * 8 + g;
                split_tf32(sW[sw_idx(n, kk + t)], bh[ni][0], bl[ni][0]);
                split_tf32(sW[sw_idx(n, kk + t + 4)], bh[ni][1], bl[ni][1]);
            }
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    mma_tf32(acc[mi][ni], al[mi], bh[ni]);     // small terms first
                    mma_tf32(acc[mi][ni], ah[mi], bl[ni]);
                    mma_tf32(acc[mi][ni], ah[mi], bh[ni]);
                }
        }
    }
    cp_async_wait<0>();
    __syncthreads();   // every warp is done with the ring before the next job refills it

    // epilogue: thread holds rows {r, r+8} x columns {n, n+1} of each 16 x 8 block
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = m0 + wm * 32 + mi * 16 + g + half * 8;
            if (r >= rows) continue;
            const float* brow = op.epi == LT_PARTIAL ? nullptr : bias_row(op, r, cx);
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int n = n0 + wn * 32 + ni * 8 + 2 * t;
                float v0 = acc[mi][ni][half * 2], v1 = acc[mi][ni][half * 2 + 1];
                if (op.epi == LT_PARTIAL) {
                    float* dst = op.out + (static_cast<long long>(split) * rows + r) * op.N + n;
                    __stcg(reinterpret_cast<float2*>(dst), make_float2(v0, v1));
                    continue;
                }
                v0 += __ldg(brow + n); v1 += __ldg(brow + n + 1);
                if (op.epi == LT_BIAS_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                else if (op.epi == LT_BIAS_SILU) { v0 = v0 / (1.f + expf(-v0)); v1 = v1 / (1.f + expf(-v1)); }
                if (op.epi != LT_FINAL) {
                    __stcg(reinterpret_cast<float2*>(op.out + static_cast<long long>(r) * op.ldo + n), make_float2(v0, v1));
                    continue;
                }
                // LT_FINAL: eps = output.2(...) ; z0 = (z - n*eps)/s ; z <- s_next*z0 + n_next*eps + cz*noise
                // (diffusion.py:586-606, 637-645), or eps_out <- eps in forward mode
                const long long i = static_cast<long long>(r) * c.D + n;
                if (cx.forward) {
                    c.eps_out[i] = v0; c.eps_out[i + 1] = v1;
                    continue;
                }
                const float* sr = c.sched + static_cast<long long>(cx.step) * kSchedRow;
                const float nr = sr[0], sg = sr[1], s2 = sr[2], n2 = sr[3], cz = sr[4];
                const float e[2] = {v0, v1};
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float zt = __ldcg(c.z + i + q);
                    const float z0 = __fdiv_rn(__fsub_rn(zt, __fmul_rn(nr, e[q])), sg);
                    float zn = __fadd_rn(__fmul_rn(s2, z0), __fmul_rn(n2, e[q]));
                    if (cz != 0.f) {
                        float w;
                        if (c.noise) w = c.noise[static_cast<long long>(cx.step) * c.noise_step_stride + i + q];
                        else {
                            float w1, w2;
                            philox_normal3(c.seed, c.sample_offset + r, static_cast<uint32_t>(cx.step), static_cast<uint32_t>(n + q), w,
                                           w1, w2);
                        }
                        zn = __fadd_rn(zn, __fmul_rn(cz, w));
                    }
                    __stcg(c.z + i + q, zn);
                }
            }
        }
}

// y[row, group] <- relu(GroupNorm(bias + sum_s partial[s])) with the statistics of nn.GroupNorm(8, C) on [B, C]
// (biased variance, eps 1e-5); gamma == nullptr: plain bias (+ ReLU).  One warp per (row, group), values stay in registers.
__device__ void norm_phase(const LtOp& op, const LatentCall& c, const StepCtx& cx, int rows) {
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const int G = op.C >> 3;
    const long long plane = static_cast<long long>(rows) * op.C;
    for (int wi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); wi < rows * 8; wi += nwarps) {
        const int b = wi >> 3, grp = wi & 7;
        const float* brow = bias_row(op, b, cx) + grp * G;
        const long long off = static_cast<long long>(b) * op.C + grp * G;
        float* out = (op.out ? op.out : c.eps_out) + static_cast<long long>(b) * op.ldo + grp * G;
        for (int base = 0; base < G; base += 512) {     // G <= 512 for every GroupNorm layer: a single trip
            float v[16];
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int i = base + lane + 32 * j;
                v[j] = 0.f;
                if (i < G) {
                    float a = __ldg(brow + i);
                    for (int sp = 0; sp < op.nsplit; ++sp) a += __ldcg(op.partial + sp * plane + off + i);
                    v[j] = a;
                    s += a;
                }
            }
            if (op.gamma == nullptr) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int i = base + lane + 32 * j;
                    if (i < G) __stcg(out + i, op.act ? fmaxf(v[j], 0.f) : v[j]);
                }
                continue;
            }
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s / static_cast<float>(G);
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int i = base + lane + 32 * j;
                if (i < G) { const float d = v[j] - mean; q = fmaf(d, d, q); }
            }
            for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            const float rstd = rsqrtf(q / static_cast<float>(G) + 1e-5f);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int i = base + lane + 32 * j;
                if (i < G) {
                    const int ch = grp * G + i;
                    __stcg(out + i, fmaxf((v[j] - mean) * rstd * __ldg(op.gamma + ch) + __ldg(op.beta + ch), 0.f));
                }
            }
        }
    }
}

// sinusoidal timestep embedding (networks.py:1088-1106): emb[r] = [sin(t_r f_j), cos(t_r f_j)], one row per time row
__device__ void emb_phase(const LtOp& op, const LatentCall& c, const StepCtx& cx, int rows) {
    const long long n = static_cast<long long>(rows) * 256;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i >> 8), j = static_cast<int>(i & 255);
        const float tt = cx.forward ? c.t_in[r] : c.sched[static_cast<long long>(r) * kSchedRow + 5];
        const float a = tt * __ldg(op.bias + (j & 127));     // op.bias = frequency table [128]
        __stcg(op.out + i, j < 128 ? sinf(a) : cosf(a));
    }
}

__device__ void run_op(const LtOp& op, const LatentCall& c, const StepCtx& cx, int R, float* smem) {
    const int rows = op.rows_mode ? R : c.B;
    if (op.kind == LT_GEMM) {
        const int m_tiles = (rows + BM - 1) / BM, n_tiles = op.N / BN;
        const int items = m_tiles * n_tiles * op.ks;
        for (int it = blockIdx.x; it < items; it += gridDim.x) {
            const int split = it % op.ks, rest = it / op.ks;
            gemm_item(op, c, cx, rows, rest / n_tiles, rest % n_tiles, split, smem);
        }
    } else if (op.kind == LT_NORM) {
        norm_phase(op, c, cx, rows);
    } else {
        emb_phase(op, c, cx, rows);
    }
}

}  // namespace

__global__ void __launch_bounds__(256, 1) latent_mk_kernel(const LtProgram* __restrict__ prog, const LatentCall* __restrict__ callp,
                                                           int S, int R, int forward, unsigned* bar) {
    extern __shared__ __align__(128) float lt_smem[];
    const LatentCall c = *callp;
    unsigned target = 0;
    StepCtx cx{0, forward};
    const int n_pre = prog->n_pre, n_loop = prog->n_loop;
    for (int i = 0; i < n_pre; ++i) {
        run_op(prog->ops[i], c, cx, R, lt_smem);
        grid_sync(bar, target);
    }
    for (int step = 0; step < S; ++step) {
        cx.step = step;
        for (int i = 0; i < n_loop; ++i) {
            run_op(prog->ops[n_pre + i], c, cx, R, lt_smem);
            grid_sync(bar, target);
        }
    }
}

// C[i][col0 + j] = sum_m Wd[i][col0 + m] * Wr[m][j]  (double accumulation), cbias[i] = bd[i] + sum_m Wd[i][col0 + m] * br[m]:
// decK(cat([prev, refineK(x_k)])) = Wd[:, :col0] prev + (Wd[:, col0:] Wr) x_k + cbias   (networks.py:1080-1083)
__global__ void compose_refine_kernel(const float* __restrict__ Wd, int ldd, int col0, const float* __restrict__ Wr, int kr,
                                      const float* __restrict__ bd, const float* __restrict__ br, float* __restrict__ C,
                                      float* __restrict__ cbias, int cout) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (i >= cout) return;
    const float* wrow = Wd + static_cast<long long>(i) * ldd + col0;
    if (j < kr) {
        double a = 0.0;
        for (int m = 0; m < kr; ++m) a += static_cast<double>(wrow[m]) * static_cast<double>(Wr[static_cast<long long>(m) * kr + j]);
        C[static_cast<long long>(i) * ldd + col0 + j] = static_cast<float>(a);
    }
    if (j == 0) {
        double a = bd[i];
        for (int m = 0; m < kr; ++m) a += static_cast<double>(wrow[m]) * static_cast<double>(br[m]);
        cbias[i] = static_cast<float>(a);
    }
}

cudaError_t launch_compose_refine(const float* Wd, int ldd, int col0, const float* Wr, int kr, const float* bd, const float* br,
                                  float* C, float* cbias, int cout, cudaStream_t s) {
    dim3 grid((kr + 127) / 128, cout);
    compose_refine_kernel<<<grid, 128, 0, s>>>(Wd, ldd, col0, Wr, kr, bd, br, C, cbias, cout);
    return cudaGetLastError();
}

static constexpr int kLtSmemBytes = NST * STAGE * static_cast<int>(sizeof(float));

cudaError_t latent_mk_grid(int num_sms, int* grid_out) {
    cudaError_t e = cudaFuncSetAttribute(latent_mk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLtSmemBytes);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, latent_mk_kernel, 256, kLtSmemBytes);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    *grid_out = num_sms;      // one CTA per SM: every phase is sized for it
    return cudaSuccess;
}

cudaError_t launch_latent_mk(const LtProgram* prog, const LatentCall* call, int S, int R, int forward, unsigned* bar, int grid,
                             cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(bar, 0, sizeof(unsigned), stream);
    if (e != cudaSuccess) return e;
    void* args[] = {&prog, &call, &S, &R, &forward, &bar};
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(latent_mk_kernel), dim3(grid), dim3(256), args, kLtSmemBytes,
                                       stream);
}

}  // namespace pcd
