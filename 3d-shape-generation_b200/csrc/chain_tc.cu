// Fused chains of narrow per-point layers (sm_100a): several consecutive `Conv1d(k=1) -> BatchNorm1d(eval) -> ReLU` layers of the
// denoiser (networks.py:46-48, 800-803, 812-816) whose intermediates are at most 128 channels wide run as ONE persistent kernel
// per chain.  A 128-point row tile walks the whole chain inside a CTA: the intermediate activations never leave shared memory
// (they are written by the epilogue in exactly the K-major SWIZZLE_128B layout tcgen05.mma reads its A operand from), only the
// chain's inputs, its skip outputs (x1, x2) and its result touch HBM, and 6-7 launches become one.
//
//   chain A : enc1.conv1 (K = 3, CUDA cores, from x_t) -> enc1.conv2 -> enc1.conv3 (= x1, to HBM) -> enc2.conv1 -> enc2.conv2
//             -> enc2.conv3 (= x2, to HBM only)
//   chain D': dec1.conv1 ([d2 | x1] streamed from HBM) -> dec1.conv2 -> dec1.conv3 -> output.0 -> output.3 + sampler update
//
// Why these chains: their layers are HBM bound on their own (profiles/ncu_per_layer_r2_summary.txt: 73-89 % of the copy rate,
// 13-52 % tensor pipe) and, at small batch, pure launch / drain latency (10-12 us per launch for microseconds of work).
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-9 = epilogue (two warps per TMEM
// lane quarter on alternate 32-column sub-groups).  One ring of operand slots (a slot = one 128-row x 64-column k-block, all
// planes) feeds the MMAs with weight tiles and, for a chain's first layer, streamed activation tiles.  Arithmetic, operand planes
// (NP = 1 or 3) and rounding are those of gemm_tc.cu's layers, so results are bit-identical to the unfused kernels.
#include "pcd_launch.h"
#include "pcd_ptx.cuh"
#include "pcd_sampler.cuh"
#include "pcd_types.h"

namespace pcd {

namespace {

constexpr int kCM = 128;                 // rows per tile
constexpr int kKB = 16384;               // bytes of one 128 x 64 16-bit k-block tile (one plane)
constexpr int kChainThreads = 320;
constexpr int kActKb = 2;                // k-blocks of the on-chip activation buffer (128 channels)

__host__ __device__ constexpr int chain_slots(int pl) { return pl == 2 ? 4 : 6; }
__host__ __device__ constexpr int chain_smem_bytes(int pl) {
    return kActKb * pl * kKB + chain_slots(pl) * pl * kKB + 8 * 2048 /*store staging*/ + 4096 /*aux*/ + 1024 /*alignment*/;
}

// D[tmem] (+)= A[smem] * B[smem]^T, one CTA, M = 128
__device__ __forceinline__ void mma1(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    tc_mma_bf16(d, da, db, idesc, acc);
}

}  // namespace

template <int PL, int F16>
__global__ void __launch_bounds__(kChainThreads, 1)
chain_tc_kernel(const __grid_constant__ ChainMaps maps, const ChainParams p) {
    constexpr int NS = chain_slots(PL);
    constexpr int SLOT = PL * kKB;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* act = smem;                                   // [kActKb][PL] k-block tiles
    uint8_t* ring = act + kActKb * PL * kKB;               // [NS] slots of PL tiles
    uint8_t* stage_out = ring + NS * SLOT;                 // 8 warps x 2 KB
    uint8_t* aux = stage_out + 8 * 2048;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);             // [NS]
    uint64_t* empty_bar = full_bar + NS;                               // [NS]
    uint64_t* tfull_bar = empty_bar + NS;                              // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                              // [2]
    uint64_t* act_bar = tempty_bar + 2;                                // the activation buffer holds the next layer's input
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(act_bar + 1);
    float* sbias = reinterpret_cast<float*>(aux + 256);                // [256]
    float* sw = sbias + 256;                                           // enc1.conv1: Wx [64][3] | final: w3 [3][64] + b3 [3]
    float* sxe = sw + 196;                                             // final: partial eps [128][3]   (384 floats)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int l = 0; l < p.nlayers; ++l) prefetch_tensormap(&maps.w[l]);
        for (int i = 0; i < NS; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
        mbar_init(act_bar, 8);
        fence_mbar_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch();
    pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer: weight tiles (+ streamed activation tiles of the first layer) =====================
        if (lane == 0) {
            int slot = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_m_blocks; tile += gridDim.x) {
                for (int l = 0; l < p.nlayers; ++l) {
                    const ChainLayer& L = p.L[l];
                    const int bnc = L.n < 128 ? L.n : 128, chunks = L.n / bnc;
                    const int wplanes = L.np == 3 ? 2 : 1;
                    for (int c = 0; c < chunks; ++c)
                        for (int kb = 0; kb < L.kb; ++kb) {
                            if (L.kb_ext0 + L.kb_ext1 > 0) {           // streamed A operand: this tile's rows, every plane the layer reads
                                mbar_wait(&empty_bar[slot], phase ^ 1);
                                mbar_arrive_expect_tx(&full_bar[slot], wplanes * kKB);
                                const CUtensorMap* tm = kb < L.kb_ext0 ? &maps.ext[0] : &maps.ext[1];
                                const int col = (kb < L.kb_ext0 ? kb : kb - L.kb_ext0) * 64;
                                for (int pl = 0; pl < wplanes; ++pl)
                                    tma_load_2d(ring + slot * SLOT + pl * kKB, tm, &full_bar[slot], col, tile * kCM + pl * p.a_plane_rows);
                                if (++slot == NS) { slot = 0; phase ^= 1; }
                            }
                            mbar_wait(&empty_bar[slot], phase ^ 1);
                            mbar_arrive_expect_tx(&full_bar[slot], wplanes * bnc * 128);
                            for (int pl = 0; pl < wplanes; ++pl)
                                tma_load_2d(ring + slot * SLOT + pl * kKB, &maps.w[l], &full_bar[slot], kb * 64, c * bnc + pl * L.n);
                            if (++slot == NS) { slot = 0; phase ^= 1; }
                        }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int slot = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            uint32_t act_phase = 0;
            for (int tile = blockIdx.x; tile < p.num_m_blocks; tile += gridDim.x) {
                for (int l = 0; l < p.nlayers; ++l) {
                    const ChainLayer& L = p.L[l];
                    const int bnc = L.n < 128 ? L.n : 128, chunks = L.n / bnc;
                    const bool streamed = L.kb_ext0 + L.kb_ext1 > 0;
                    const uint32_t idesc = make_idesc(128, bnc, F16);
                    if (!streamed) {                       // the previous layer's epilogue (or enc1.conv1) has filled the activation buffer
                        mbar_wait(act_bar, act_phase);
                        act_phase ^= 1;
                        tc_fence_after();
                    }
                    for (int c = 0; c < chunks; ++c) {
                        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + acc * 128;
                        for (int kb = 0; kb < L.kb; ++kb) {
                            uint32_t a_addr;
                            int a_slot = -1;
                            if (streamed) {
                                mbar_wait(&full_bar[slot], phase);
                                a_slot = slot;
                                a_addr = smem_u32(ring + slot * SLOT);
                                if (++slot == NS) { slot = 0; phase ^= 1; }
                            } else {
                                a_addr = smem_u32(act + kb * PL * kKB);
                            }
                            mbar_wait(&full_bar[slot], phase);
                            tc_fence_after();
                            const uint64_t da = make_sw128_kmajor_desc(a_addr);
                            const uint64_t db = make_sw128_kmajor_desc(smem_u32(ring + slot * SLOT));
                            constexpr uint64_t LO = kKB >> 4;              // second plane of a tile pair: 16 KB further
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                mma1(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                                if (L.np == 3) {
                                    mma1(d_tmem, da + 2 * k, db + LO + 2 * k, idesc, 1u);          // hi * lo
                                    mma1(d_tmem, da + LO + 2 * k, db + 2 * k, idesc, 1u);          // lo * hi
                                }
                            }
                            if (a_slot >= 0) tc_commit(&empty_bar[a_slot]);
                            tc_commit(&empty_bar[slot]);
                            if (++slot == NS) { slot = 0; phase ^= 1; }
                        }
                        tc_commit(&tfull_bar[acc]);
                        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                    }
                }
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int ew = warp - 2, q = warp & 3, half = ew >> 2;
        const int epi_tid = threadIdx.x - 64;          // 0..255
        const int row_in_tile = q * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        uint8_t* stg = stage_out + ew * 2048;
        const int swz64 = (lane >> 1) & 3;
        const SamplerArgs& sa = p.call->s;

        for (int tile = blockIdx.x; tile < p.num_m_blocks; tile += gridDim.x) {
            const long long row0 = static_cast<long long>(tile) * kCM;
            const int sample = static_cast<int>(row0 / p.rows_per_sample);
            if (p.first_from_x) {
                // enc1.conv1 (xyz columns; the temb columns are the per-step bias row): 8 threads per row, 8 channels each,
                // same expressions as enc1_first_kernel -> the activation buffer's k-block 0
                for (int i = epi_tid; i < 192; i += 256) sw[i] = p.Wx[i];
                if (epi_tid < 64)
                    sbias[epi_tid] = p.call->bias1_steps ? p.call->bias1_steps[static_cast<long long>(*sa.step_ptr) * 64 + epi_tid]
                                                         : p.bias1[static_cast<long long>(sample) * 64 + epi_tid];
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const int c8 = epi_tid & 7;
#pragma unroll 1
                for (int it = 0; it < 4; ++it) {
                    const int r = it * 32 + (epi_tid >> 3);
                    const int n = static_cast<int>(row0 - static_cast<long long>(sample) * p.rows_per_sample) + r;
                    float x0 = 0.f, x1 = 0.f, x2 = 0.f;
                    if (n < sa.N) {
                        const float* xp = sa.x + (static_cast<long long>(sample) * sa.N + n) * 3;
                        x0 = xp[0]; x1 = xp[1]; x2 = xp[2];
                    }
                    uint32_t pk[4], pl2[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = c8 * 8 + 2 * j;
                        float a = fmaf(sw[c * 3 + 2], x2, fmaf(sw[c * 3 + 1], x1, fmaf(sw[c * 3], x0, sbias[c])));
                        float d = fmaf(sw[c * 3 + 5], x2, fmaf(sw[c * 3 + 4], x1, fmaf(sw[c * 3 + 3], x0, sbias[c + 1])));
                        a = fmaxf(a, 0.f); d = fmaxf(d, 0.f);
                        pk[j] = pack16x2(a, d, F16);
                        const float2 back = unpack16x2(pk[j], F16);
                        pl2[j] = pack16x2(a - back.x, d - back.y, F16);
                    }
                    const uint32_t off = r * 128 + ((c8 ^ (r & 7)) << 4);
                    *reinterpret_cast<uint4*>(act + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    if (PL == 2) *reinterpret_cast<uint4*>(act + kKB + off) = make_uint4(pl2[0], pl2[1], pl2[2], pl2[3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(act_bar);
                asm volatile("bar.sync 1, 256;" ::: "memory");     // sw / sbias are rewritten below
            }
            for (int l = 0; l < p.nlayers; ++l) {
                const ChainLayer& L = p.L[l];
                const int bnc = L.n < 128 ? L.n : 128, chunks = L.n / bnc;
                // this layer's bias row (shared + per-sample part)
                {
                    const float* bsrc = L.bias + static_cast<long long>(sample) * L.bias_sample_stride;
                    for (int i = epi_tid; i < L.n; i += 256) sbias[i] = bsrc[i];
                    if (L.final) {
                        for (int i = epi_tid; i < 192; i += 256) sw[i] = sa.w3[i];
                        if (epi_tid < 3) sw[192 + epi_tid] = sa.b3[epi_tid];
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
                for (int c = 0; c < chunks; ++c) {
                    mbar_wait(&tfull_bar[acc], acc_phase);
                    tc_fence_after();
                    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 128;
                    if (L.final) {
                        // output.0 accumulators -> ReLU -> output.3 (64 -> 3) in ONE sequential fmaf chain per point (channels 0..31 on the
                        // first warp of the quarter, 32..63 continued by the second: the order of gemm_tc.cu's final epilogue)
                        uint32_t v[32];
                        tmem_ld_32x32(t_addr + half * 32, v);
                        tc_wait_ld();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                        float e0, e1, e2;
                        if (half == 0) { e0 = sw[192]; e1 = sw[193]; e2 = sw[194]; }
                        else {
                            asm volatile("bar.sync 2, 256;" ::: "memory");
                            e0 = sxe[row_in_tile * 3]; e1 = sxe[row_in_tile * 3 + 1]; e2 = sxe[row_in_tile * 3 + 2];
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float h = fmaxf(__uint_as_float(v[j]) + sbias[half * 32 + j], 0.f);
                            e0 = fmaf(sw[half * 32 + j], h, e0);
                            e1 = fmaf(sw[64 + half * 32 + j], h, e1);
                            e2 = fmaf(sw[128 + half * 32 + j], h, e2);
                        }
                        if (half == 0) {
                            sxe[row_in_tile * 3] = e0; sxe[row_in_tile * 3 + 1] = e1; sxe[row_in_tile * 3 + 2] = e2;
                            asm volatile("bar.sync 2, 256;" ::: "memory");
                        } else {
                            sampler_apply(sa, row0 + row_in_tile, e0, e1, e2);
                        }
                    } else {
#pragma unroll 1
                        for (int sg = half; sg < bnc / 32; sg += 2) {
                            uint32_t v[32];
                            tmem_ld_32x32(t_addr + sg * 32, v);
                            tc_wait_ld();
                            const float4* sb4 = reinterpret_cast<const float4*>(sbias + c * bnc + sg * 32);
                            uint4 pk[4], pk2[4];
#pragma unroll
                            for (int cc = 0; cc < 4; ++cc) {
                                const float4 b0 = sb4[2 * cc], b1 = sb4[2 * cc + 1];
                                float f[8] = {__uint_as_float(v[8 * cc]) + b0.x,     __uint_as_float(v[8 * cc + 1]) + b0.y,
                                              __uint_as_float(v[8 * cc + 2]) + b0.z, __uint_as_float(v[8 * cc + 3]) + b0.w,
                                              __uint_as_float(v[8 * cc + 4]) + b1.x, __uint_as_float(v[8 * cc + 5]) + b1.y,
                                              __uint_as_float(v[8 * cc + 6]) + b1.z, __uint_as_float(v[8 * cc + 7]) + b1.w};
#pragma unroll
                                for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
                                const uint32_t h0 = pack16x2(f[0], f[1], F16), h1 = pack16x2(f[2], f[3], F16);
                                const uint32_t h2 = pack16x2(f[4], f[5], F16), h3 = pack16x2(f[6], f[7], F16);
                                pk[cc] = make_uint4(h0, h1, h2, h3);
                                if (PL == 2) {
                                    const float2 r0 = unpack16x2(h0, F16), r1 = unpack16x2(h1, F16);
                                    const float2 r2 = unpack16x2(h2, F16), r3 = unpack16x2(h3, F16);
                                    pk2[cc] = make_uint4(pack16x2(f[0] - r0.x, f[1] - r0.y, F16), pack16x2(f[2] - r1.x, f[3] - r1.y, F16),
                                                         pack16x2(f[4] - r2.x, f[5] - r2.y, F16), pack16x2(f[6] - r3.x, f[7] - r3.y, F16));
                                }
                            }
                            if (L.to_act) {
                                // next layer's A operand: k-block (c * bnc + sg * 32) / 64, 16-byte chunks ((sg & 1) * 4 + cc) ^ (row & 7)
                                const int gsg = c * (bnc / 32) + sg;
                                uint8_t* dst = act + (gsg >> 1) * PL * kKB + row_in_tile * 128;
#pragma unroll
                                for (int cc = 0; cc < 4; ++cc) {
                                    const uint32_t o = ((((gsg & 1) * 4 + cc) ^ (row_in_tile & 7)) << 4);
                                    *reinterpret_cast<uint4*>(dst + o) = pk[cc];
                                    if (PL == 2) *reinterpret_cast<uint4*>(dst + kKB + o) = pk2[cc];
                                }
                            }
                            if (L.to_hbm >= 0) {
                                const CUtensorMap* tm = &maps.out[L.to_hbm];
                                if (lane == 0) tma_store_wait_read<0>();
                                __syncwarp();
#pragma unroll
                                for (int cc = 0; cc < 4; ++cc) *reinterpret_cast<uint4*>(stg + lane * 64 + ((cc ^ swz64) << 4)) = pk[cc];
                                fence_proxy_async_smem();
                                __syncwarp();
                                if (lane == 0) {
                                    tma_store_2d(tm, stg, c * bnc + sg * 32, static_cast<int>(row0) + q * 32);
                                    tma_store_commit();
                                }
                                if (PL == 2) {
                                    if (lane == 0) tma_store_wait_read<0>();
                                    __syncwarp();
#pragma unroll
                                    for (int cc = 0; cc < 4; ++cc) *reinterpret_cast<uint4*>(stg + lane * 64 + ((cc ^ swz64) << 4)) = pk2[cc];
                                    fence_proxy_async_smem();
                                    __syncwarp();
                                    if (lane == 0) {
                                        tma_store_2d(tm, stg, c * bnc + sg * 32, static_cast<int>(row0) + q * 32 + p.a_plane_rows);
                                        tma_store_commit();
                                    }
                                }
                            }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                    }
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                if (L.to_act) {                // the whole layer output is in the activation buffer: hand it to the MMA issuer
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(act_bar);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");     // sbias / sw are rewritten by the next layer
            }
        }
        if (lane == 0) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

cudaError_t configure_chain_tc() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(chain_tc_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain_smem_bytes(1))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(chain_tc_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain_smem_bytes(1))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(chain_tc_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain_smem_bytes(2))) != cudaSuccess) return e;
    return cudaFuncSetAttribute(chain_tc_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, chain_smem_bytes(2));
}

cudaError_t launch_chain_tc(int planes, int f16, const ChainMaps& maps, const ChainParams& p, int num_sms, cudaStream_t stream) {
    const int grid = p.num_m_blocks < num_sms ? p.num_m_blocks : num_sms;
    const size_t smem = chain_smem_bytes(planes);
    if (planes == 2) {
        if (f16) return launch_pdl(chain_tc_kernel<2, 1>, dim3(grid), dim3(kChainThreads), smem, stream, maps, p);
        return launch_pdl(chain_tc_kernel<2, 0>, dim3(grid), dim3(kChainThreads), smem, stream, maps, p);
    }
    if (f16) return launch_pdl(chain_tc_kernel<1, 1>, dim3(grid), dim3(kChainThreads), smem, stream, maps, p);
    return launch_pdl(chain_tc_kernel<1, 0>, dim3(grid), dim3(kChainThreads), smem, stream, maps, p);
}

}  // namespace pcd
