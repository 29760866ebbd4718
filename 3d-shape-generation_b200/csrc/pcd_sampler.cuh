// Device-side sampler update (x0-prediction + re-noise) and the counter-based noise source.
#pragma once
#include "pcd_types.h"

namespace pcd {

// Philox4x32-10 (Salmon et al. 2011).  counter = (point, step, sample_lo, sample_hi),
// key = (seed_lo, seed_hi): a noise value depends only on (seed, GLOBAL sample index,
// step, point), so 1/2/4/8-GPU shardings draw identical noise.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float u32_to_unit(uint32_t x) {
    // (x >> 8 + 0.5) * 2^-24: exactly representable, never 0 or 1
    return (static_cast<float>(x >> 8) + 0.5f) * 5.9604644775390625e-8f;
}

// three N(0,1) draws for one point (Box-Muller on two uniform pairs; 4th value discarded)
__device__ __forceinline__ void philox_normal3(unsigned long long seed, unsigned long long sample, uint32_t step,
                                               uint32_t point, float& z0, float& z1, float& z2) {
    uint32_t r[4];
    philox4x32_10(point, step, static_cast<uint32_t>(sample), static_cast<uint32_t>(sample >> 32),
                  static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
    const float u1 = u32_to_unit(r[0]), u2 = u32_to_unit(r[1]), u3 = u32_to_unit(r[2]), u4 = u32_to_unit(r[3]);
    const float ra = sqrtf(-2.0f * logf(u1)), rb = sqrtf(-2.0f * logf(u3));
    const float two_pi = 6.283185307179586f;
    z0 = ra * cosf(two_pi * u2);
    z1 = ra * sinf(two_pi * u2);
    z2 = rb * cosf(two_pi * u4);
}

// Apply the per-point tail of one step.  `row` is the padded row index b*Npad + n.
//   x0 = (x - n*eps)/s                         (remove_noise, diffusion.py:167)
//   x' = s_next*x0 + n_next*eps + cz*z          (DDIM :287 has cz=0; DDPM :255 has n_next=0,
//                                                cz = sqrt(n_prev/n)*n; last step: s_next=1,
//                                                n_next=0, cz=0 so x' = x0 exactly)
// Explicit _rn intrinsics keep nvcc from contracting into FMAs: same op order as the reference.
__device__ __forceinline__ void sampler_apply(const SamplerArgs& s, long long row, float e0, float e1, float e2) {
    const int b = static_cast<int>(row / s.Npad);
    const int n = static_cast<int>(row - static_cast<long long>(b) * s.Npad);
    if (n >= s.N) return;
    const long long xi = (static_cast<long long>(b) * s.N + n) * 3;
    if (s.mode == 0) {
        s.eps_out[xi + 0] = e0; s.eps_out[xi + 1] = e1; s.eps_out[xi + 2] = e2;
        return;
    }
    const int step = *s.step_ptr;
    const float* r = s.sched + (static_cast<long long>(step) * s.sched_rows + (s.sched_rows > 1 ? b : 0)) * kSchedRow;
    const float nr = r[0], sr = r[1], s2 = r[2], n2 = r[3], cz = r[4];
    float z0 = 0.f, z1 = 0.f, z2 = 0.f;
    if (cz != 0.f) {
        if (s.noise != nullptr) {
            const float* zp = s.noise + static_cast<long long>(step) * s.noise_step_stride + xi;
            z0 = zp[0]; z1 = zp[1]; z2 = zp[2];
        } else {
            philox_normal3(s.seed, s.sample_offset + b, static_cast<uint32_t>(step), static_cast<uint32_t>(n), z0, z1,
                           z2);
        }
    }
    const float e[3] = {e0, e1, e2};
    const float z[3] = {z0, z1, z2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float xt = s.x[xi + c];
        const float x0 = __fdiv_rn(__fsub_rn(xt, __fmul_rn(nr, e[c])), sr);
        float xn = __fadd_rn(__fmul_rn(s2, x0), __fmul_rn(n2, e[c]));
        if (cz != 0.f) xn = __fadd_rn(xn, __fmul_rn(cz, z[c]));
        s.x[xi + c] = xn;
    }
}

// post-ReLU values are >= 0, so signed-int ordering of the bit patterns equals float ordering
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}

}  // namespace pcd
