// dense_misc.cu -- the non-GEMM kernels of the VAE-encoder / UNet / ControlNet path (rows a7, a8):
// GroupNorm(+SiLU) forward/backward, LayerNorm, GEGLU, nearest 2x upsample, strided copies,
// transposes, row softmax forward/backward, layout / dtype conversion at the fp32 boundaries,
// posterior sampling, add_noise and the sinusoidal timestep embedding.
// NHWC activations, fp16 or bf16 storage, fp32 math.  All are HBM-bound streaming kernels:
// 16-byte vector accesses, grid sized in multiples of the SM count for the strided loops.
//
// Reference call sites (diffusers modules reached from models/guidance/dreammat_guidance.py:218-292):
// torch.nn.GroupNorm/SiLU in ResnetBlock2D, LayerNorm + GEGLU in BasicTransformerBlock,
// Upsample2D (nearest), AutoencoderKL DiagonalGaussianDistribution.sample (:291),
// DDIMScheduler.add_noise (:463), Timesteps / get_timestep_embedding.
#include <cuda_bf16.h>
#include "common.cuh"
#include "dense_hp.cuh"

namespace {

template <typename T> struct H;
template <> struct H<__half> {
    static __device__ __forceinline__ float f(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half t(float v) { return __float2half_rn(v); }
};
template <> struct H<__nv_bfloat16> {
    static __device__ __forceinline__ float f(__nv_bfloat16 v) { return __bfloat162float(v); }
    static __device__ __forceinline__ __nv_bfloat16 t(float v) { return __float2bfloat16_rn(v); }
};

template <> struct H<float> {
    static __device__ __forceinline__ float f(float v) { return v; }
    static __device__ __forceinline__ float t(float v) { return v; }
};

// storage dtype selector of every dense-path entry point: 0 fp16, 1 bf16, 2 fp32 (high-precision mode, dense_hp.cu)
#define DM_DISPATCH_T(bf16, ...)                                  \
    do {                                                          \
        if (bf16) { using T = __nv_bfloat16; __VA_ARGS__; }       \
        else { using T = __half; __VA_ARGS__; }                   \
    } while (0)
// element-wise kernels (no 16-bit vector packing) also run on fp32 storage
#define DM_DISPATCH_T3(dt, ...)                                   \
    do {                                                          \
        if ((dt) == 2) { using T = float; __VA_ARGS__; }          \
        else if (dt) { using T = __nv_bfloat16; __VA_ARGS__; }    \
        else { using T = __half; __VA_ARGS__; }                   \
    } while (0)
#define DM_DTYPE_OK(dt) DM_REQUIRE((dt) >= 0 && (dt) <= 2, "dtype selector: 0 fp16, 1 bf16, 2 fp32")

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
__device__ __forceinline__ float silu_grad(float x) {
    float s = __fdividef(1.0f, 1.0f + __expf(-x));
    return s * (1.0f + x * (1.0f - s));
}

// ----------------------------------------------------------------------------------- GroupNorm
// Thread mapping shared by the four kernels: the channel axis is cut into slabs of `vslab` 8-channel vectors
// (vslab = largest divisor of C/8 that is <= 32, i.e. 512 contiguous bytes per pixel), grid = (pixel chunks,
// images, slabs).  A thread owns ONE vector of the slab and strides over the chunk's pixels with `lanes` =
// blockDim / vslab pixel lanes, so per-channel constants live in registers and the loop body is pure
// load / FMA / store with several 16-byte loads in flight.  Small feature maps (8x8 .. 32x32 latents) get their
// parallelism from the slab axis instead of idling most of the machine.
struct GnIdx {
    int v, q, lanes, p0, p1;
    bool active;
};
__device__ __forceinline__ GnIdx gn_index(int HW, int vslab, int chunks) {
    GnIdx m;
    m.lanes = (int)blockDim.x / vslab;
    m.q = (int)threadIdx.x / vslab;
    m.v = (int)blockIdx.z * vslab + (int)threadIdx.x % vslab;
    m.active = m.q < m.lanes;
    const int px_per_chunk = (HW + chunks - 1) / chunks;
    m.p0 = (int)blockIdx.x * px_per_chunk;
    m.p1 = min(HW, m.p0 + px_per_chunk);
    return m;
}

// stats[(img*G + g)*2 + {0,1}] += (sum, sumsq) over the group's channels and the pixel chunk.  Channel PAIRS never
// straddle a group (cpg is even for every SD layer), so each thread keeps 4 (sum, sumsq) pairs.
template <typename T>
__global__ void __launch_bounds__(256) gn_stats_kernel(const T* __restrict__ x, int HW, int C, int ld, int G,
                                                       int chunks, int vslab, float* __restrict__ stats) {
    extern __shared__ float s_acc[];  // [G*2]
    const int img = blockIdx.y, cpg = C / G;
    griddep_launch_dependents();
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    griddep_wait();
    const GnIdx m = gn_index(HW, vslab, chunks);
    if (m.active) {
        float s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
        const T* base = x + ((int64_t)img * HW) * ld + 8 * m.v;
        const int lanes = m.lanes;
        auto acc = [&](const uint4& u) {
            const T* h = reinterpret_cast<const T*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float a = H<T>::f(h[2 * j]), b = H<T>::f(h[2 * j + 1]);
                s[j] += a + b; ss[j] += a * a + b * b;
            }
        };
        int p = m.p0 + m.q;
        for (; p + 3 * lanes < m.p1; p += 4 * lanes) {
            uint4 u[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) u[k] = *reinterpret_cast<const uint4*>(base + (int64_t)(p + k * lanes) * ld);
#pragma unroll
            for (int k = 0; k < 4; ++k) acc(u[k]);
        }
        for (; p < m.p1; p += lanes) acc(*reinterpret_cast<const uint4*>(base + (int64_t)p * ld));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int g = (8 * m.v + 2 * j) / cpg;
            atomicAdd(&s_acc[2 * g], s[j]); atomicAdd(&s_acc[2 * g + 1], ss[j]);
        }
    }
    __syncthreads();
    // only the groups this slab touches
    const int g0 = (8 * (int)blockIdx.z * vslab) / cpg, g1 = (8 * ((int)blockIdx.z + 1) * vslab - 1) / cpg;
    for (int i = 2 * g0 + threadIdx.x; i <= 2 * g1 + 1; i += blockDim.x) atomicAdd(stats + ((int64_t)img * G) * 2 + i, s_acc[i]);
}

// y = act(x * A[c] + B[c]) with A = rstd*gamma, B = beta - mean*rstd*gamma (16 registers per thread).
template <typename T>
__global__ void __launch_bounds__(256) gn_apply_kernel(const T* __restrict__ x, int HW, int C, int ld, int ldy, int G,
                                                       int chunks, int vslab, const float* __restrict__ stats,
                                                       const T* __restrict__ gamma, const T* __restrict__ beta,
                                                       float eps, int silu, T* __restrict__ y) {
    griddep_launch_dependents();
    griddep_wait();
    const int img = blockIdx.y, cpg = C / G;
    const float inv_cnt = 1.0f / ((float)HW * (float)cpg);
    const GnIdx m = gn_index(HW, vslab, chunks);
    if (!m.active) return;
    const int v = m.v, lanes = m.lanes;
    float A[8], Bc[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int g = (8 * v + 2 * j) / cpg;
        float s = stats[((int64_t)img * G + g) * 2], ss = stats[((int64_t)img * G + g) * 2 + 1];
        float mean = s * inv_cnt, rstd = rsqrtf(fmaxf(ss * inv_cnt - mean * mean, 0.f) + eps);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            float gm = H<T>::f(gamma[8 * v + 2 * j + k]);
            A[2 * j + k] = rstd * gm;
            Bc[2 * j + k] = H<T>::f(beta[8 * v + 2 * j + k]) - mean * rstd * gm;
        }
    }
    const T* xb = x + ((int64_t)img * HW) * ld + 8 * v;
    T* yb = y + ((int64_t)img * HW) * ldy + 8 * v;
    auto body = [&](uint4 u, int p) {
        const T* h = reinterpret_cast<const T*>(&u);
        uint4 o; T* oh = reinterpret_cast<T*>(&o);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float val = fmaf(H<T>::f(h[k]), A[k], Bc[k]);
            if (silu) val = silu_f(val);
            oh[k] = H<T>::t(val);
        }
        *reinterpret_cast<uint4*>(yb + (int64_t)p * ldy) = o;
    };
    int p = m.p0 + m.q;
    for (; p + 3 * lanes < m.p1; p += 4 * lanes) {
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = *reinterpret_cast<const uint4*>(xb + (int64_t)(p + k * lanes) * ld);
#pragma unroll
        for (int k = 0; k < 4; ++k) body(u[k], p + k * lanes);
    }
    for (; p < m.p1; p += lanes) body(*reinterpret_cast<const uint4*>(xb + (int64_t)p * ld), p);
}

// backward pass 1: bstats[(img*G+g)*2] += sum(gamma*dy'), += sum(gamma*dy'*xhat), dy' = dz * act'(y)
template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_stats_kernel(const T* __restrict__ x, const T* __restrict__ dz, int HW,
                                                           int C, int G, int chunks, int vslab,
                                                           const float* __restrict__ stats, const T* __restrict__ gamma,
                                                           const T* __restrict__ beta, float eps, int silu,
                                                           float* __restrict__ bstats) {
    extern __shared__ float s_acc[];
    griddep_launch_dependents();
    const int img = blockIdx.y, cpg = C / G;
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    griddep_wait();
    const float inv_cnt = 1.0f / ((float)HW * (float)cpg);
    const GnIdx m = gn_index(HW, vslab, chunks);
    if (m.active) {
        const int v = m.v, lanes = m.lanes;
        // xhat*gamma + beta = x*A + Bc;  xhat = x*rs + ms
        float A[8], Bc[8], gm[8], rs[4], ms[4], a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int g = (8 * v + 2 * j) / cpg;
            float s = stats[((int64_t)img * G + g) * 2], ss = stats[((int64_t)img * G + g) * 2 + 1];
            float mean = s * inv_cnt;
            rs[j] = rsqrtf(fmaxf(ss * inv_cnt - mean * mean, 0.f) + eps);
            ms[j] = -mean * rs[j];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            gm[k] = H<T>::f(gamma[8 * v + k]);
            A[k] = rs[k >> 1] * gm[k];
            Bc[k] = H<T>::f(beta[8 * v + k]) + ms[k >> 1] * gm[k];
        }
        const int64_t base = ((int64_t)img * HW) * C + 8 * v;
        auto acc = [&](const uint4& ux, const uint4& ud) {
            const T* hx = reinterpret_cast<const T*>(&ux); const T* hd = reinterpret_cast<const T*>(&ud);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float xv = H<T>::f(hx[k]);
                float d = H<T>::f(hd[k]) * gm[k];
                if (silu) d *= silu_grad(fmaf(xv, A[k], Bc[k]));
                a1[k >> 1] += d;
                a2[k >> 1] = fmaf(d, fmaf(xv, rs[k >> 1], ms[k >> 1]), a2[k >> 1]);
            }
        };
        int p = m.p0 + m.q;
        for (; p + lanes < m.p1; p += 2 * lanes) {
            uint4 ux0 = *reinterpret_cast<const uint4*>(x + base + (int64_t)p * C);
            uint4 ud0 = *reinterpret_cast<const uint4*>(dz + base + (int64_t)p * C);
            uint4 ux1 = *reinterpret_cast<const uint4*>(x + base + (int64_t)(p + lanes) * C);
            uint4 ud1 = *reinterpret_cast<const uint4*>(dz + base + (int64_t)(p + lanes) * C);
            acc(ux0, ud0); acc(ux1, ud1);
        }
        for (; p < m.p1; p += lanes)
            acc(*reinterpret_cast<const uint4*>(x + base + (int64_t)p * C), *reinterpret_cast<const uint4*>(dz + base + (int64_t)p * C));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int g = (8 * v + 2 * j) / cpg;
            atomicAdd(&s_acc[2 * g], a1[j]); atomicAdd(&s_acc[2 * g + 1], a2[j]);
        }
    }
    __syncthreads();
    const int g0 = (8 * (int)blockIdx.z * vslab) / cpg, g1 = (8 * ((int)blockIdx.z + 1) * vslab - 1) / cpg;
    for (int i = 2 * g0 + threadIdx.x; i <= 2 * g1 + 1; i += blockDim.x) atomicAdd(bstats + ((int64_t)img * G) * 2 + i, s_acc[i]);
}

// backward pass 2: dx = rstd * (gamma*dy' - (S1 + xhat*S2)/cnt)  (+ dx_add if given); same mapping as gn_apply
template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dz, int HW,
                                                           int C, int G, int chunks, int vslab,
                                                           const float* __restrict__ stats, const float* __restrict__ bstats,
                                                           const T* __restrict__ gamma, const T* __restrict__ beta,
                                                           float eps, int silu, const T* __restrict__ dx_add,
                                                           T* __restrict__ dx) {
    griddep_launch_dependents();
    griddep_wait();
    const int img = blockIdx.y, cpg = C / G;
    const float inv_cnt = 1.0f / ((float)HW * (float)cpg);
    const GnIdx m = gn_index(HW, vslab, chunks);
    if (!m.active) return;
    const int v = m.v, lanes = m.lanes;
    // act'(y) needs y = x*A + Bc (A = rs*gm, Bc = beta + ms*gm, ms = -mean*rs); with xhat = x*rs + ms
    // dx = rs*(gm*d' - S1 - xhat*S2) = d'*A - (x*P + Q),  P = rs^2*S2,  Q = rs*(S1 + ms*S2)
    float A[8], Bc[8], rs[4], ms[4], Pq[4], Qq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int g = (8 * v + 2 * j) / cpg;
        int64_t si = ((int64_t)img * G + g) * 2;
        float s = stats[si], ss = stats[si + 1];
        float mean = s * inv_cnt;
        rs[j] = rsqrtf(fmaxf(ss * inv_cnt - mean * mean, 0.f) + eps);
        ms[j] = -mean * rs[j];
        float S1 = bstats[si] * inv_cnt, S2 = bstats[si + 1] * inv_cnt;
        Pq[j] = rs[j] * rs[j] * S2;
        Qq[j] = rs[j] * (S1 + ms[j] * S2);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float gm = H<T>::f(gamma[8 * v + k]);
        A[k] = rs[k >> 1] * gm;
        Bc[k] = H<T>::f(beta[8 * v + k]) + ms[k >> 1] * gm;
    }
    const int64_t base = ((int64_t)img * HW) * C + 8 * v;
    auto body = [&](const uint4& ux, const uint4& ud, const uint4& ua, int64_t off) {
        const T* hx = reinterpret_cast<const T*>(&ux); const T* hd = reinterpret_cast<const T*>(&ud);
        const T* ha = reinterpret_cast<const T*>(&ua);
        uint4 o; T* oh = reinterpret_cast<T*>(&o);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float xv = H<T>::f(hx[k]);
            float d = H<T>::f(hd[k]);
            if (silu) d *= silu_grad(fmaf(xv, A[k], Bc[k]));
            float val = fmaf(d, A[k], -fmaf(xv, Pq[k >> 1], Qq[k >> 1]));
            if (dx_add) val += H<T>::f(ha[k]);
            oh[k] = H<T>::t(val);
        }
        *reinterpret_cast<uint4*>(dx + off) = o;
    };
    const uint4 zero = make_uint4(0, 0, 0, 0);
    int p = m.p0 + m.q;
    for (; p + lanes < m.p1; p += 2 * lanes) {
        const int64_t o0 = base + (int64_t)p * C, o1 = base + (int64_t)(p + lanes) * C;
        uint4 ux0 = *reinterpret_cast<const uint4*>(x + o0), ud0 = *reinterpret_cast<const uint4*>(dz + o0);
        uint4 ux1 = *reinterpret_cast<const uint4*>(x + o1), ud1 = *reinterpret_cast<const uint4*>(dz + o1);
        uint4 ua0 = dx_add ? *reinterpret_cast<const uint4*>(dx_add + o0) : zero;
        uint4 ua1 = dx_add ? *reinterpret_cast<const uint4*>(dx_add + o1) : zero;
        body(ux0, ud0, ua0, o0); body(ux1, ud1, ua1, o1);
    }
    for (; p < m.p1; p += lanes) {
        const int64_t o0 = base + (int64_t)p * C;
        body(*reinterpret_cast<const uint4*>(x + o0), *reinterpret_cast<const uint4*>(dz + o0),
             dx_add ? *reinterpret_cast<const uint4*>(dx_add + o0) : zero, o0);
    }
}

// ----------------------------------------------------------------------------------- LayerNorm (warp per row)
template <typename T>
__global__ void __launch_bounds__(256) layernorm_kernel(const T* __restrict__ x, int64_t M, int C,
                                                        const T* __restrict__ gamma, const T* __restrict__ beta,
                                                        float eps, T* __restrict__ y) {
    // warp per row, 16-byte vectors (C % 8 == 0, C <= 1280 -> at most 5 vectors per lane kept in registers);
    // two-pass mean / centred variance on the register copy
    griddep_launch_dependents();
    griddep_wait();
    int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31, nv = C >> 3;
    const T* xr = x + row * C;
    float vals[40];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int v = lane + 32 * k;
        if (v < nv) {
            uint4 u = *reinterpret_cast<const uint4*>(xr + 8 * v);
            const T* h = reinterpret_cast<const T*>(&u);
#pragma unroll
            for (int j = 0; j < 8; ++j) { float f = H<T>::f(h[j]); vals[8 * k + j] = f; s += f; }
        }
    }
    s = warp_sum(s);
    const float mean = s / (float)C;
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        if (lane + 32 * k < nv) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { float d = vals[8 * k + j] - mean; ss += d * d; }
        }
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss / (float)C + eps);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int v = lane + 32 * k;
        if (v < nv) {
            uint4 ug = *reinterpret_cast<const uint4*>(gamma + 8 * v), ub = *reinterpret_cast<const uint4*>(beta + 8 * v);
            const T* hg = reinterpret_cast<const T*>(&ug); const T* hb = reinterpret_cast<const T*>(&ub);
            uint4 o; T* oh = reinterpret_cast<T*>(&o);
#pragma unroll
            for (int j = 0; j < 8; ++j) oh[j] = H<T>::t((vals[8 * k + j] - mean) * rstd * H<T>::f(hg[j]) + H<T>::f(hb[j]));
            *reinterpret_cast<uint4*>(y + row * C + 8 * v) = o;
        }
    }
}

// GEGLU: out[m, j] = h[m, j] * gelu(h[m, D + j])   (diffusers GEGLU: hidden, gate = chunk(2); hidden * gelu(gate))
template <typename T>
__global__ void __launch_bounds__(256) geglu_kernel(const T* __restrict__ h, int64_t M, int D, T* __restrict__ out) {
    int64_t n = M * (int64_t)(D / 8);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t m = i / (D / 8); int j = (int)(i - m * (D / 8)) * 8;
        uint4 a = *reinterpret_cast<const uint4*>(h + m * 2 * D + j), g = *reinterpret_cast<const uint4*>(h + m * 2 * D + D + j);
        const T* ha = reinterpret_cast<const T*>(&a); const T* hg = reinterpret_cast<const T*>(&g);
        uint4 o; T* oh = reinterpret_cast<T*>(&o);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float gv = H<T>::f(hg[k]);
            oh[k] = H<T>::t(H<T>::f(ha[k]) * (0.5f * gv * (1.0f + erff(gv * 0.70710678118654752f))));
        }
        *reinterpret_cast<uint4*>(out + m * D + j) = o;
    }
}

// nearest-neighbour 2x upsample, NHWC
template <typename T>
__global__ void __launch_bounds__(256) upsample2x_kernel(const T* __restrict__ x, int n, int Hh, int W, int C,
                                                         T* __restrict__ y) {
    int64_t nv = (int64_t)n * 2 * Hh * 2 * W * (C / 8);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % (C / 8)); int64_t r = i / (C / 8);
        int xo = (int)(r % (2 * W)); r /= 2 * W; int yo = (int)(r % (2 * Hh)); int img = (int)(r / (2 * Hh));
        const uint4* src = reinterpret_cast<const uint4*>(x + (((int64_t)img * Hh + yo / 2) * W + xo / 2) * C) + c;
        reinterpret_cast<uint4*>(y)[i] = *src;
    }
}

// zero-insertion 2x (transpose of a stride-2 gather): y[2i,2j] = x[i,j], else 0
template <typename T>
__global__ void __launch_bounds__(256) zero_insert2x_kernel(const T* __restrict__ x, int n, int Hh, int W, int C,
                                                            T* __restrict__ y) {
    int64_t nv = (int64_t)n * 2 * Hh * 2 * W * (C / 8);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % (C / 8)); int64_t r = i / (C / 8);
        int xo = (int)(r % (2 * W)); r /= 2 * W; int yo = (int)(r % (2 * Hh)); int img = (int)(r / (2 * Hh));
        uint4 v = make_uint4(0, 0, 0, 0);
        if (((xo | yo) & 1) == 0) v = reinterpret_cast<const uint4*>(x + (((int64_t)img * Hh + yo / 2) * W + xo / 2) * C)[c];
        reinterpret_cast<uint4*>(y)[i] = v;
    }
}

// dst[r, 0:cols] = a*src[r, 0:cols] (+ b*src2[r,0:cols]) with independent row strides (16-byte vectors)
template <typename T>
__global__ void __launch_bounds__(256) axpby2d_kernel(const T* __restrict__ s1, int64_t ld1, float a,
                                                      const T* __restrict__ s2, int64_t ld2, float b, int64_t rows,
                                                      int cols, T* __restrict__ dst, int64_t ldd) {
    int vpr = cols / 8;
    int64_t nv = rows * vpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / vpr; int c = (int)(i - r * vpr) * 8;
        uint4 u = *reinterpret_cast<const uint4*>(s1 + r * ld1 + c);
        if (!s2 && a == 1.0f) { *reinterpret_cast<uint4*>(dst + r * ldd + c) = u; continue; }
        uint4 w = make_uint4(0, 0, 0, 0);
        if (s2) w = *reinterpret_cast<const uint4*>(s2 + r * ld2 + c);
        const T* hu = reinterpret_cast<const T*>(&u); const T* hw = reinterpret_cast<const T*>(&w);
        uint4 o; T* oh = reinterpret_cast<T*>(&o);
#pragma unroll
        for (int k = 0; k < 8; ++k) oh[k] = H<T>::t(a * H<T>::f(hu[k]) + (s2 ? b * H<T>::f(hw[k]) : 0.f));
        *reinterpret_cast<uint4*>(dst + r * ldd + c) = o;
    }
}

// batched transpose [B, R, C] -> [B, C, R] through a 32x33 shared tile
template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ x, int R, int C, int64_t ldx,
                                                        int64_t bsx, T* __restrict__ y, int64_t ldy, int64_t bsy) {
    // 64 x 64 tile of 16-bit elements moved as 32-bit pairs on both sides (128-byte warp requests): the load packs two
    // neighbouring columns, the store two neighbouring rows of the input.  Even R, C, ld (checked on the host).
    __shared__ uint32_t tile[64][33];     // [row][column pair]
    const T* xb = x + (int64_t)blockIdx.z * bsx; T* yb = y + (int64_t)blockIdx.z * bsy;
    const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 64; j += 8) {
        int r = r0 + j, c = c0 + 2 * tx;
        uint32_t v = 0;
        if (r < R && c < C) v = *reinterpret_cast<const uint32_t*>(xb + (int64_t)r * ldx + c);
        tile[j][tx] = v;
    }
    __syncthreads();
    for (int j = ty; j < 64; j += 8) {       // output row c0 + j holds input column c0 + j; this thread writes rows r0+2tx, +1
        int c = c0 + j, r = r0 + 2 * tx;
        if (c < C && r < R) {
            uint32_t lo = tile[2 * tx][j >> 1], hi = tile[2 * tx + 1][j >> 1];
            uint32_t a = (j & 1) ? (lo >> 16) : (lo & 0xffffu), bq = (j & 1) ? (hi >> 16) : (hi & 0xffffu);
            *reinterpret_cast<uint32_t*>(yb + (int64_t)c * ldy + r) = a | (bq << 16);
        }
    }
}

// row softmax: y = softmax(scale * x) over `cols`; one block per row
template <typename T>
__global__ void __launch_bounds__(256) softmax_rows_kernel(const T* __restrict__ x, int cols, int64_t ld, float scale,
                                                           T* __restrict__ y) {
    __shared__ float red[8];
    const T* xr = x + (int64_t)blockIdx.x * ld; T* yr = y + (int64_t)blockIdx.x * ld;
    float mx = -3e38f;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) mx = fmaxf(mx, H<T>::f(xr[c]) * scale);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) mx = fmaxf(mx, red[k]);
    __syncthreads();
    float s = 0.f;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) s += __expf(H<T>::f(xr[c]) * scale - mx);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    s = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += red[k];
    float inv = 1.0f / s;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) yr[c] = H<T>::t(__expf(H<T>::f(xr[c]) * scale - mx) * inv);
}

// dS = scale * P * (dP - rowsum(P * dP))
template <typename T>
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const T* __restrict__ P, const T* __restrict__ dP, int cols,
                                                          int64_t ld, float scale, T* __restrict__ dS) {
    __shared__ float red[8];
    const T* pr = P + (int64_t)blockIdx.x * ld; const T* dr = dP + (int64_t)blockIdx.x * ld; T* o = dS + (int64_t)blockIdx.x * ld;
    float s = 0.f;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) s += H<T>::f(pr[c]) * H<T>::f(dr[c]);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    s = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += red[k];
    for (int c = threadIdx.x; c < cols; c += blockDim.x) o[c] = H<T>::t(scale * H<T>::f(pr[c]) * (H<T>::f(dr[c]) - s));
}

// fp32 [rows, cin] -> T [rows, cpad]: y = x*scale + shift on the real channels, 0 on the padding
template <typename T>
__global__ void __launch_bounds__(256) pad_convert_kernel(const float* __restrict__ x, int64_t rows, int cin, int cpad,
                                                          float scale, float shift, T* __restrict__ y) {
    // one 16-byte store (8 channels) per thread; cpad % 8 == 0
    const int vpr = cpad >> 3;
    const int64_t n = rows * vpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / vpr; const int c0 = (int)(i - r * vpr) * 8;
        uint4 o; T* oh = reinterpret_cast<T*>(&o);
#pragma unroll
        for (int k = 0; k < 8; ++k) oh[k] = H<T>::t((c0 + k < cin) ? fmaf(x[r * cin + c0 + k], scale, shift) : 0.f);
        *reinterpret_cast<uint4*>(y + r * cpad + c0) = o;
    }
}
// N1: the pre-rendered condition maps stay on the device as the bytes the PNG decode produced
// (data/uncond.py:532-582): depth fp32 [V, HW], normal u8 [V, HW, 3], light u8 [V, E, HW, 18].  One pass gathers the
// (view, env) pairs of the batch, de-quantises (x / 255, the reference's arithmetic) and writes the channel-padded
// ControlNet condition [B, HW, cpad] in the storage dtype: channel order depth | normal | 6 x RGB light (:581-582,:802).
template <typename T>
__global__ void __launch_bounds__(256) cond_gather_kernel(const float* __restrict__ depth, const uint8_t* __restrict__ normal,
                                                          const uint8_t* __restrict__ light, int E, int64_t HW,
                                                          const int32_t* __restrict__ view_ids, const int32_t* __restrict__ env_ids,
                                                          int B, int cpad, T* __restrict__ out) {
    const int vpr = cpad >> 3;
    const int64_t n = (int64_t)B * HW * vpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % vpr) * 8; const int64_t r = i / vpr;
        const int64_t px = r % HW; const int b = (int)(r / HW);
        const int v = view_ids[b], e = env_ids[b];
        const uint8_t* nr = normal + ((int64_t)v * HW + px) * 3;
        const uint8_t* lt = light + (((int64_t)v * E + e) * HW + px) * 18;
        uint4 o; T* oh = reinterpret_cast<T*>(&o);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = c0 + k;
            float f = 0.f;
            if (c == 0) f = depth[(int64_t)v * HW + px];
            else if (c < 4) f = (float)nr[c - 1] / 255.0f;
            else if (c < 22) f = (float)lt[c - 4] / 255.0f;
            oh[k] = H<T>::t(f);
        }
        *reinterpret_cast<uint4*>(out + r * cpad + c0) = o;
    }
}

// T [rows, ld] (first cout channels) -> fp32 [rows, cout] * scale
template <typename T>
__global__ void __launch_bounds__(256) unpad_convert_kernel(const T* __restrict__ x, int64_t rows, int ld, int cout,
                                                            float scale, float* __restrict__ y) {
    int64_t n = rows * cout;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / cout; int c = (int)(i - r * cout);
        y[i] = H<T>::f(x[r * ld + c]) * scale;
    }
}
// NHWC T [n, HW, ld] (first C channels) -> NCHW fp32 [n, C, HW]
template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_f32_kernel(const T* __restrict__ x, int n, int HW, int ld, int Cc,
                                                               float* __restrict__ y) {
    int64_t tot = (int64_t)n * Cc * HW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
        int p = (int)(i % HW); int64_t r = i / HW; int c = (int)(r % Cc); int img = (int)(r / Cc);
        y[i] = H<T>::f(x[((int64_t)img * HW + p) * ld + c]);
    }
}

// posterior sample (DiagonalGaussianDistribution): z = (mean + exp(0.5*clamp(logvar,-30,20)) * eps) * sf
// moments NHWC [n, HW, ld] (mean = ch 0..3, logvar = ch 4..7); z, eps NCHW fp32 [n,4,HW]
template <typename T>
__global__ void __launch_bounds__(256) vae_sample_kernel(const T* __restrict__ mom, int n, int HW, int ld,
                                                         const float* __restrict__ eps, float sf, float* __restrict__ z) {
    int64_t tot = (int64_t)n * 4 * HW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
        int p = (int)(i % HW); int64_t r = i / HW; int c = (int)(r % 4); int img = (int)(r / 4);
        const T* m = mom + ((int64_t)img * HW + p) * ld;
        float mean = H<T>::f(m[c]), lv = fminf(fmaxf(H<T>::f(m[4 + c]), -30.f), 20.f);
        // diffusers computes the sample in the weights dtype: round the pre-scale sample to T
        float smp = H<T>::f(H<T>::t(mean + H<T>::f(H<T>::t(__expf(0.5f * lv))) * H<T>::f(H<T>::t(eps[i]))));
        z[i] = H<T>::f(H<T>::t(smp * sf));
    }
}
// backward: dmom[.., c] = dz*sf ; dmom[.., 4+c] = dz*sf*eps*0.5*std (0 outside the clamp); padding channels 0
template <typename T>
__global__ void __launch_bounds__(256) vae_sample_bwd_kernel(const T* __restrict__ mom, int n, int HW, int ld,
                                                             const float* __restrict__ eps, float sf,
                                                             const float* __restrict__ dz, T* __restrict__ dmom) {
    int64_t tot = (int64_t)n * HW * ld;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
        int ch = (int)(i % ld); int64_t r = i / ld; int p = (int)(r % HW); int img = (int)(r / HW);
        float v = 0.f;
        if (ch < 4) v = dz[((int64_t)img * 4 + ch) * HW + p] * sf;
        else if (ch < 8) {
            int c = ch - 4;
            float lvr = H<T>::f(mom[i]);
            if (lvr >= -30.f && lvr <= 20.f) {
                float sd = __expf(0.5f * lvr);
                int64_t zi = ((int64_t)img * 4 + c) * HW + p;
                v = dz[zi] * sf * eps[zi] * 0.5f * sd;
            }
        }
        dmom[i] = H<T>::t(v);
    }
}

// add_noise + CFG replication: out[(k*B + b), p, 0:4] = sqrt(ac[t_b]) z + sqrt(1-ac[t_b]) noise, k = 0..rep-1; pad 0
template <typename T>
__global__ void __launch_bounds__(256) add_noise_kernel(const float* __restrict__ z, const float* __restrict__ noise,
                                                        const float* __restrict__ sqrt_ac,
                                                        const float* __restrict__ sqrt_1mac, int B, int HW, int cpad,
                                                        int rep, T* __restrict__ out) {
    int64_t tot = (int64_t)rep * B * HW * cpad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(i % cpad); int64_t r = i / cpad; int p = (int)(r % HW); int bk = (int)(r / HW); int b = bk % B;
        float v = 0.f;
        if (c < 4) { int64_t zi = ((int64_t)b * 4 + c) * HW + p; v = sqrt_ac[b] * z[zi] + sqrt_1mac[b] * noise[zi]; }
        out[i] = H<T>::t(v);
    }
}

// get_timestep_embedding(t, dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos(t f_k) | sin(t f_k)]
template <typename T>
__global__ void timestep_embed_kernel(const float* __restrict__ t, int n, int dim, T* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int half = dim / 2;
    if (i >= n * half) return;
    int b = i / half, k = i % half;
    float f = expf(-9.210340371976184f * (float)k / (float)half);
    float a = t[b] * f;
    out[b * dim + k] = H<T>::t(cosf(a));
    out[b * dim + half + k] = H<T>::t(sinf(a));
}

template <typename T>
__global__ void __launch_bounds__(256) silu_kernel(const T* __restrict__ x, int64_t n, T* __restrict__ y) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = H<T>::t(silu_f(H<T>::f(x[i])));
}

inline int grid_for(int64_t work, int threads = 256) {
    int64_t b = dm_ceil_div(work, threads);
    int64_t cap = (int64_t)DM_NUM_SMS * 8;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}
inline int gn_vslab(int vpp) {
    for (int d = 32; d > 1; --d) if (vpp % d == 0) return d;
    return 1;
}
// pixel chunks so that chunks * n_img * slabs ~ 8 CTAs per SM, each thread keeping >= 4 pixels to stream
inline int gn_chunks(int n_img, int HW, int slabs, int lanes) {
    int c = (DM_NUM_SMS * 8) / ((n_img > 0 ? n_img : 1) * slabs);
    if (c < 1) c = 1;
    int maxc = HW / (4 * lanes); if (maxc < 1) maxc = 1;
    return c < maxc ? c : maxc;
}

}  // namespace

extern "C" int dm_groupnorm(int bf16, const void* x, int n_img, int HW, int C, int ld, int G, const void* gamma,
                            const void* beta, float eps, int silu, void* y, int ldy, float* stats, void* stream) {
    DM_DTYPE_OK(bf16);
    if (bf16 == 2) return hp_groupnorm((const float*)x, n_img, HW, C, ld, G, (const float*)gamma, (const float*)beta, eps, silu, (float*)y, ldy, stats, stream);
    DM_REQUIRE(x && gamma && beta && y && stats, "null pointer");
    DM_REQUIRE(C % 8 == 0 && C % G == 0 && (C / G) % 2 == 0 && ld % 8 == 0 && ldy % 8 == 0, "channel layout");
    cudaStream_t st = (cudaStream_t)stream;
    DM_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * n_img * G, st));
    const int vslab = gn_vslab(C / 8), slabs = (C / 8) / vslab;
    const int chunks = gn_chunks(n_img, HW, slabs, 256 / vslab);
    const dim3 grid(chunks, n_img, slabs);
    DM_DISPATCH_T(bf16, DM_CHECK_CUDA(dm_launch(gn_stats_kernel<T>, grid, dim3(256), 2 * G * sizeof(float), st, (const T*)x, HW, C, ld, G, chunks, vslab, stats)));
    DM_CHECK_LAUNCH();
    DM_DISPATCH_T(bf16, DM_CHECK_CUDA(dm_launch(gn_apply_kernel<T>, grid, dim3(256), 0, st, (const T*)x, HW, C, ld, ldy, G, chunks, vslab, (const float*)stats,
                                                (const T*)gamma, (const T*)beta, eps, silu, (T*)y)));
    DM_CHECK_LAUNCH();
    return DM_OK;
}

extern "C" int dm_groupnorm_bwd(int bf16, const void* x, const void* dz, int n_img, int HW, int C, int G,
                                const void* gamma, const void* beta, float eps, int silu, const float* stats,
                                float* bstats, const void* dx_add, void* dx, void* stream) {
    DM_DTYPE_OK(bf16);
    if (bf16 == 2) return hp_groupnorm_bwd((const float*)x, (const float*)dz, n_img, HW, C, G, (const float*)gamma, (const float*)beta, eps, silu, stats, (const float*)dx_add, (float*)dx, stream);
    DM_REQUIRE(x && dz && gamma && beta && stats && bstats && dx, "null pointer");
    DM_REQUIRE(C % 8 == 0 && C % G == 0 && (C / G) % 2 == 0, "channel layout");
    cudaStream_t st = (cudaStream_t)stream;
    DM_CHECK_CUDA(cudaMemsetAsync(bstats, 0, sizeof(float) * 2 * n_img * G, st));
    const int vslab = gn_vslab(C / 8), slabs = (C / 8) / vslab;
    const int chunks = gn_chunks(n_img, HW, slabs, 256 / vslab);
    const dim3 grid(chunks, n_img, slabs);
    DM_DISPATCH_T(bf16, DM_CHECK_CUDA(dm_launch(gn_bwd_stats_kernel<T>, grid, dim3(256), 2 * G * sizeof(float), st,
                            (const T*)x, (const T*)dz, HW, C, G, chunks, vslab, stats, (const T*)gamma, (const T*)beta, eps, silu, bstats)));
    DM_CHECK_LAUNCH();
    DM_DISPATCH_T(bf16, DM_CHECK_CUDA(dm_launch(gn_bwd_apply_kernel<T>, grid, dim3(256), 0, st, (const T*)x, (const T*)dz, HW, C, G, chunks, vslab, stats,
                                                (const float*)bstats, (const T*)gamma, (const T*)beta, eps, silu, (const T*)dx_add, (T*)dx)));
    DM_CHECK_LAUNCH();
    return DM_OK;
}

extern "C" int dm_layernorm(int bf16, const void* x, int64_t M, int C, const void* gamma, const void* beta, float eps,
                            void* y, void* stream) {
    DM_DTYPE_OK(bf16);
    if (bf16 == 2) return hp_layernorm((const float*)x, M, C, (const float*)gamma, (const float*)beta, eps, (float*)y, stream);
    DM_REQUIRE(x && gamma && beta && y, "null pointer");
    DM_REQUIRE(C <= 1280 && C % 8 == 0, "C <= 1280, multiple of 8");
    if (M == 0) return DM_OK;
    DM_DISPATCH_T(bf16, DM_CHECK_CUDA(dm_launch(layernorm_kernel<T>, dim3((unsigned)dm_ceil_div(M, 8)), dim3(256), 0, (cudaStream_t)stream,
                                                (const T*)x, M, C, (const T*)gamma, (const T*)beta, eps, (T*)y)));
    DM_CHECK_LAUNCH();
    return DM_OK;
}

extern "C" int dm_geglu(int bf16, const void* h, int64_t M, int D, void* out, void* stream) {
    DM_DTYPE_OK(bf16);
    if (bf16 == 2) return hp_geglu((const float*)h, M, D, (float*)out, stream);
    DM_REQUIRE(h && out && D % 8 == 0, "bad args");
    DM_DISPATCH_T(bf16, geglu_kernel<T><<<grid_for(M * (D / 8)), 256, 0, (cudaStream_t)stream>>>((const T*)h, M, D, (T*)out));
    DM_CHECK_LAUNCH();
    return DM_OK;
}

extern "C" int dm_upsample2x(int bf16, const void* x, int n, int H, int W, int C, int zero_insert, void* y, void* stream) {
    DM_DTYPE_OK(bf16);
    if (bf16 == 2) return dm_upsample2x(0, x, n, H, W, 2 * C, zero_insert, y, stream);
    DM_REQUIRE(x && y && C % 8 == 0, "bad args");
    int64_t nv = (int64_t)n * 4 * H * W * (C / 8);
    if (zero_insert) DM_DISPATCH_T(bf16, zero_insert2x_kernel<T><<<grid_for(nv), 256, 0, (cudaStream_t)stream>>>((const T*)x, n, H, W, C, (T*)y));
    else DM_DISPATCH_T(bf16, upsample2x_kernel<T><<<grid_for(nv), 256, 0, (cudaStream_t)stream>>>((const T*)x, n, H, W, C, (T*)y));
    DM_CHECK_LAUNCH();
    return DM_OK;
}

extern "C" int dm_axpby2d(int bf16, const void* s1, int64_t ld1, float a, const void* s2, int64_t ld2, float b,
                          int64_t rows, int cols, void* dst, int64_t ldd, void* stream) {
    DM_DTYPE_OK(bf16);
    if (bf16 == 2) return hp_axpby2d((const float*)s1, ld1, a, (const float*)s2, ld2, b, rows, cols, (float*)dst, ldd, stream);
    DM_REQUIRE(s1 && dst && cols % 8 == 0 && ld1 % 8 == 0 && ldd % 8 == 0 && (!s2 || ld2 % 8 == 0), "bad args");
    if (rows == 0) return DM_OK;
    DM_DISPATCH_T(bf16, axpby2d_kernel<T><<<grid_for(rows * (cols / 8)), 256, 0, (cudaStream_t)stream>>>((const T*)s1, ld1, a, (const T*)s2, ld2, b,
                                                                                                     rows, cols, (T*)dst, ldd));
    DM_CHECK_LAUNCH();
    return DM_OK;
}

extern "C" int dm_transpose(int bf16, const void* x, int batch, int R, int C, int64_t ldx, int64_t bsx, void* y,
                            int64_t ldy, int64_t bsy, void* stream) {
    DM_DTYPE_OK(bf16);
    if (bf16 == 2) return hp_transpose((const float*)x, batch, R, C, ldx, bsx, (float*)y, ldy, bsy, stream);
    DM_REQUIRE(x && y, "null pointer");
    DM_REQUIRE(((R | C) & 1) == 0 && ((ldx | ldy | bsx | bsy) & 1) == 0 && ((((uintptr_t)x) | ((uintptr_t)y)) & 3) == 0,
               "transpose moves 32-bit pairs: even extents / strides, 4-byte aligned bases");
    dim3 grid((unsigned)dm_ceil_div(C, 64), (unsigned)dm_ceil_div(R, 64), (unsigned)batch);
    DM_DISPATCH_T(bf16, transpose_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, R, C, ldx, bsx, (T*)y, ldy, bsy));
    DM_CHECK_LAUNCH();
    return DM_OK;
}

extern "C" int dm_softmax_rows(int bf16, const void* x, int64_t rows, int cols, int64_t ld, float scale, void* y,
                               void* stream) {
    DM_DTYPE_OK(bf16);
    DM_REQUIRE(x && y, "null pointer");
    if (rows == 0) return DM_OK;
    DM_DISPATCH_T3(bf16, softmax_rows_kernel<T><<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>((const T*)x, cols, ld, scale, (T*)y));
    DM_CHECK_LAUNCH();
    return DM_OK;
}

extern "C" int dm_softmax_bwd(int bf16, const void* P, const void* dP, int64_t rows, int cols, int64_t ld, float scale,
                              void* dS, void* stream) {
    DM_DTYPE_OK(bf16);
    DM_REQUIRE(P && dP && dS, "null pointer");
    if (rows == 0) return DM_OK;
    DM_DISPATCH_T3(bf16, softmax_bwd_kernel<T><<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>((const T*)P, (const T*)dP, cols, ld, scale, (T*)dS));
    DM_CHECK_LAUNCH();
    return DM_OK;
}

extern "C" int dm_pad_convert(int bf16, const float* x, int64_t rows, int cin, int cpad, float scale, float shift,
                              void* y, void* stream) {
    DM_DTYPE_OK(bf16);
    if (bf16 == 2) return hp_pad_convert(x, rows, cin, cpad, scale, shift, (float*)y, stream);
    DM_REQUIRE(x && y && cpad >= cin && cpad % 8 == 0, "bad args (cpad must be a multiple of 8)");
    DM_DISPATCH_T(bf16, pad_convert_kernel<T><<<grid_for(rows * (cpad / 8)), 256, 0, (cudaStream_t)stream>>>(x, rows, cin, cpad, scale, shift, (T*)y));
    DM_CHECK_LAUNCH();
    return DM_OK;
}

extern "C" int dm_cond_gather(int bf16, const float* depth, const uint8_t* normal, const uint8_t* light, int n_env, int64_t HW,
                              const int32_t* view_ids, const int32_t* env_ids, int B, int cpad, void* out, void* stream) {
    DM_REQUIRE(bf16 == 0 || bf16 == 1, "16-bit storage (the fp32 mode takes the float condition map through dm_pad_convert)");
    DM_REQUIRE(depth && normal && light && view_ids && env_ids && out && cpad >= 22 && cpad % 8 == 0 && n_env > 0, "bad args");
    if (B == 0 || HW == 0) return DM_OK;
    DM_DISPATCH_T(bf16, cond_gather_kernel<T><<<grid_for((int64_t)B * HW * (cpad / 8)), 256, 0, (cudaStream_t)stream>>>(
                            depth, normal, light, n_env, HW, view_ids, env_ids, B, cpad, (T*)out));
    DM_CHECK_LAUNCH();
    return DM_OK;
}

extern "C" int dm_unpad_convert(int bf16, const void* x, int64_t rows, int ld, int cout, float scale, float* y,
                                void* stream) {
    DM_DTYPE_OK(bf16);
    DM_REQUIRE(x && y && ld >= cout, "bad args");
    DM_DISPATCH_T3(bf16, unpad_convert_kernel<T><<<grid_for(rows * cout), 256, 0, (cudaStream_t)stream>>>((const T*)x, rows, ld, cout, scale, y));
    DM_CHECK_LAUNCH();
    return DM_OK;
}

extern "C" int dm_nhwc_to_nchw_f32(int bf16, const void* x, int n, int HW, int ld, int C, float* y, void* stream) {
    DM_DTYPE_OK(bf16);
    DM_REQUIRE(x && y, "null pointer");
    DM_DISPATCH_T3(bf16, nhwc_to_nchw_f32_kernel<T><<<grid_for((int64_t)n * C * HW), 256, 0, (cudaStream_t)stream>>>((const T*)x, n, HW, ld, C, y));
    DM_CHECK_LAUNCH();
    return DM_OK;
}

extern "C" int dm_vae_sample(int bf16, const void* moments, int n, int HW, int ld, const float* eps, float scaling,
                             float* z, void* stream) {
    DM_DTYPE_OK(bf16);
    DM_REQUIRE(moments && eps && z && ld >= 8, "bad args");
    DM_DISPATCH_T3(bf16, vae_sample_kernel<T><<<grid_for((int64_t)n * 4 * HW), 256, 0, (cudaStream_t)stream>>>((const T*)moments, n, HW, ld, eps, scaling, z));
    DM_CHECK_LAUNCH();
    return DM_OK;
}

extern "C" int dm_vae_sample_bwd(int bf16, const void* moments, int n, int HW, int ld, const float* eps, float scaling,
                                 const float* dz, void* dmoments, void* stream) {
    DM_DTYPE_OK(bf16);
    DM_REQUIRE(moments && eps && dz && dmoments && ld >= 8, "bad args");
    DM_DISPATCH_T3(bf16, vae_sample_bwd_kernel<T><<<grid_for((int64_t)n * HW * ld), 256, 0, (cudaStream_t)stream>>>((const T*)moments, n, HW, ld, eps,
                                                                                                               scaling, dz, (T*)dmoments));
    DM_CHECK_LAUNCH();
    return DM_OK;
}

extern "C" int dm_add_noise(int bf16, const float* z, const float* noise, const float* sqrt_ac, const float* sqrt_1mac,
                            int B, int HW, int cpad, int rep, void* out, void* stream) {
    DM_DTYPE_OK(bf16);
    DM_REQUIRE(z && noise && sqrt_ac && sqrt_1mac && out && cpad >= 4, "bad args");
    DM_DISPATCH_T3(bf16, add_noise_kernel<T><<<grid_for((int64_t)rep * B * HW * cpad), 256, 0, (cudaStream_t)stream>>>(z, noise, sqrt_ac, sqrt_1mac, B, HW,
                                                                                                                  cpad, rep, (T*)out));
    DM_CHECK_LAUNCH();
    return DM_OK;
}

extern "C" int dm_timestep_embedding(int bf16, const float* t, int n, int dim, void* out, void* stream) {
    DM_DTYPE_OK(bf16);
    DM_REQUIRE(t && out && dim % 2 == 0, "bad args");
    DM_DISPATCH_T3(bf16, timestep_embed_kernel<T><<<(unsigned)dm_ceil_div((int64_t)n * dim / 2, 128), 128, 0, (cudaStream_t)stream>>>(t, n, dim, (T*)out));
    DM_CHECK_LAUNCH();
    return DM_OK;
}

extern "C" int dm_silu(int bf16, const void* x, int64_t n, void* y, void* stream) {
    DM_DTYPE_OK(bf16);
    DM_REQUIRE(x && y, "null pointer");
    DM_DISPATCH_T3(bf16, silu_kernel<T><<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const T*)x, n, (T*)y));
    DM_CHECK_LAUNCH();
    return DM_OK;
}
