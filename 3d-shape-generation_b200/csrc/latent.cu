// Latent-diffusion path (BASELINE config 4): SimpleLatentUNetPointNet (reference networks.py:962-1106)
// = Linear + GroupNorm(8) + ReLU blocks on [B, C] rows, the latent reverse-loop update
// (reference diffusion.py:575-707) and SimplePointNetVAE.decode (networks.py:1144-1154, 1219-1231).
// Rows are samples (M = B), so this path is weight-bandwidth / latency bound; the
// GEMMs run on the fp32 CUDA-core kernel of gemm_simt.cu, GroupNorm+ReLU is one warp per
// (sample, group), and the whole step is replayed as a CUDA graph.
#include "pcd_sampler.cuh"
#include "pcd_types.h"

namespace pcd {

// y[b, g*G .. (g+1)*G) <- relu( (y - mean) * rsqrt(var + 1e-5) * gamma + beta ), biased variance,
// statistics over the G = C/8 channels of one group of one sample (nn.GroupNorm(8, C) on [B, C]).
// If nsplit > 0 the pre-norm value is first assembled from split-K partial sums in a fixed order:
// y = bias + sum_s partial[s]  (deterministic), otherwise y already holds Linear(x) + bias.
__global__ void __launch_bounds__(256) groupnorm_relu_kernel(float* __restrict__ y, const float* __restrict__ partial, int nsplit,
                                                             const float* __restrict__ bias, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, int B, int C) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= B * 8) return;
    const int b = warp >> 3, g = warp & 7;
    const int G = C >> 3;
    float* row = y + static_cast<long long>(b) * C + g * G;
    if (nsplit > 0) {
        const long long n = static_cast<long long>(B) * C, off = static_cast<long long>(b) * C + g * G;
        for (int i = lane; i < G; i += 32) {
            float v = bias[g * G + i];
            for (int sp = 0; sp < nsplit; ++sp) v += partial[sp * n + off + i];
            row[i] = v;
        }
        __syncwarp();
    }
    float s = 0.f;
    for (int i = lane; i < G; i += 32) s += row[i];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / static_cast<float>(G);
    float v = 0.f;
    for (int i = lane; i < G; i += 32) { const float d = row[i] - mean; v = fmaf(d, d, v); }
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v / static_cast<float>(G) + 1e-5f);
    for (int i = lane; i < G; i += 32) {
        const int c = g * G + i;
        row[i] = fmaxf((row[i] - mean) * rstd * gamma[c] + beta[c], 0.f);
    }
}

cudaError_t launch_groupnorm_relu(float* y, const float* partial, int nsplit, const float* bias, const float* gamma, const float* beta,
                                  int B, int C, cudaStream_t stream) {
    const int warps = B * 8;
    groupnorm_relu_kernel<<<(warps * 32 + 255) / 256, 256, 0, stream>>>(y, partial, nsplit, bias, gamma, beta, B, C);
    return cudaGetLastError();
}

// z [B, D] update:  z0 = (z - n*eps)/s ;  z <- s_next*z0 + n_next*eps + cz*noise   (diffusion.py:586-606, 637-645)
// mode 0: copy eps to eps_out (forward-only hook).
__global__ void __launch_bounds__(256) latent_update_kernel(const float* __restrict__ eps, const LatentCall* __restrict__ ca) {
    const LatentCall c = *ca;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(c.B) * c.D) return;
    const float e = eps[i];
    if (c.mode == 0) { c.eps_out[i] = e; return; }
    const int step = *c.step_ptr;
    const float* r = c.sched + static_cast<long long>(step) * kSchedRow;
    const float nr = r[0], sr = r[1], s2 = r[2], n2 = r[3], cz = r[4];
    const float z0 = __fdiv_rn(__fsub_rn(c.z[i], __fmul_rn(nr, e)), sr);
    float zn = __fadd_rn(__fmul_rn(s2, z0), __fmul_rn(n2, e));
    if (cz != 0.f) {
        float w;
        if (c.noise) w = c.noise[static_cast<long long>(step) * c.noise_step_stride + i];
        else {
            const int b = static_cast<int>(i / c.D), j = static_cast<int>(i - static_cast<long long>(b) * c.D);
            float w1, w2;
            philox_normal3(c.seed, c.sample_offset + b, static_cast<uint32_t>(step), static_cast<uint32_t>(j), w, w1, w2);
        }
        zn = __fadd_rn(zn, __fmul_rn(cz, w));
    }
    c.z[i] = zn;
}

cudaError_t launch_latent_update(const float* eps, const LatentCall* ca, int B, int D, cudaStream_t stream) {
    const long long n = static_cast<long long>(B) * D;
    latent_update_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, stream>>>(eps, ca);
    return cudaGetLastError();
}

__global__ void latent_philox_fill_kernel(float* out, unsigned long long seed, unsigned long long sample_offset, int step,
                                          int B, int D) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(B) * D) return;
    const int b = static_cast<int>(i / D), j = static_cast<int>(i - static_cast<long long>(b) * D);
    float w, w1, w2;
    philox_normal3(seed, sample_offset + b, static_cast<uint32_t>(step), static_cast<uint32_t>(j), w, w1, w2);
    out[i] = w;
}
cudaError_t launch_latent_philox_fill(float* out, unsigned long long seed, unsigned long long sample_offset, int step, int B,
                                      int D, cudaStream_t stream) {
    const long long n = static_cast<long long>(B) * D;
    latent_philox_fill_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, stream>>>(out, seed, sample_offset, step, B, D);
    return cudaGetLastError();
}

// time MLP for the latent model: temb[b] = W2 silu(W1 emb(t_b) + b1) + b2 (networks.py:977-981, 1064-1065)
__global__ void __launch_bounds__(256) latent_time_kernel(const LatentCall* __restrict__ ca, const float* __restrict__ freqs,
                                                          const float* __restrict__ W1T, const float* __restrict__ b1,
                                                          const float* __restrict__ W2T, const float* __restrict__ b2,
                                                          float* __restrict__ temb_out) {
    __shared__ float e[256], h[256];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float t = ca->t_in ? ca->t_in[b] : ca->sched[static_cast<long long>(*ca->step_ptr) * kSchedRow + 5];
    const float a = t * freqs[tid & 127];
    e[tid] = tid < 128 ? sinf(a) : cosf(a);
    __syncthreads();
    float s = b1[tid];
    for (int k = 0; k < 256; ++k) s = fmaf(W1T[k * 256 + tid], e[k], s);
    h[tid] = s / (1.f + expf(-s));
    __syncthreads();
    s = b2[tid];
    for (int k = 0; k < 256; ++k) s = fmaf(W2T[k * 256 + tid], h[k], s);
    temb_out[b * 256 + tid] = s;
}
cudaError_t launch_latent_time(int B, const LatentCall* ca, const float* freqs, const float* W1T, const float* b1,
                               const float* W2T, const float* b2, float* temb_out, cudaStream_t stream) {
    latent_time_kernel<<<B, 256, 0, stream>>>(ca, freqs, W1T, b1, W2T, b2, temb_out);
    return cudaGetLastError();
}

}  // namespace pcd
