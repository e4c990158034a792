// BatchNorm (+ReLU, +2x2 max-pool, +feature flatten) forward and backward on
// padded-planes tensors.  All kernels touch VALID pixels only, so the zero
// pads of act / pooled / dLin survive from the one-time memset.
// Reference: lib/layer_types.py:109-110 (pool), :196-199 (ReLU), :219-249 (BN).
#include "common.cuh"
#include "../../include/mpnn.h"

// ------------------------------------------------------------------ finalize
__global__ void bn_finalize_kernel(const float* __restrict__ partials, int n_parts, int C, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ m_avg, float* __restrict__ v_avg,
                                   float d, float eps, int train,
                                   float* __restrict__ ss, float* __restrict__ mr) {
    // one warp per channel: lanes stride over the partial rows (fixed order -> deterministic)
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= C) return;
    float mean, var;
    if (train) {
        double s = 0.0, s2 = 0.0;
        for (int i = lane; i < n_parts; i += 32) {
            s += (double)partials[((size_t)i * 2) * C + c];
            s2 += (double)partials[((size_t)i * 2 + 1) * C + c];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (lane != 0) return;
        double m = s / count;
        double v = s2 / count - m * m;
        if (v < 0.0) v = 0.0;
        mean = (float)m; var = (float)v;
        if (m_avg) {
            m_avg[c] = d * m_avg[c] + (1.f - d) * mean;
            v_avg[c] = d * v_avg[c] + (1.f - d) * var;
        }
    } else {
        if (lane != 0) return;
        mean = m_avg[c]; var = v_avg[c];
    }
    float rstd = 1.0f / sqrtf(var + eps);
    float a = gamma[c] * rstd;
    ss[c] = a;
    ss[C + c] = beta[c] - mean * a;
    mr[c] = mean;
    mr[C + c] = rstd;
}

extern "C" int mpnn_bn_finalize(const float* partials, int n_parts, int C, double count,
                                const float* gamma, const float* beta, float* m_avg, float* v_avg,
                                float d, float eps, int train, float* ss, float* mr, void* stream) {
    MPNN_REQUIRE(C > 0 && (train == 0 || (partials && n_parts > 0)), "bn_finalize: args");
    bn_finalize_kernel<<<ceil_div(C, 4), 128, 0, (cudaStream_t)stream>>>(
        partials, n_parts, C, count, gamma, beta, m_avg, v_avg, d, eps, train, ss, mr);
    return mpnn_check_launch("bn_finalize");
}

// ------------------------------------------------------------------ forward
template <typename T, bool POOL>
__global__ void bn_relu_pool_fwd_kernel(const T* __restrict__ lin, int C, Geom g,
                                        const float* __restrict__ ss, T* __restrict__ act,
                                        T* __restrict__ pooled, Geom gp,
                                        T* __restrict__ feat, int Balloc) {
    const int KG = C / 8;
    const int HH = POOL ? g.H / 2 : g.H, WW = POOL ? g.W / 2 : g.W;
    const long long total = (long long)KG * g.B * HH * WW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int w = i % WW;
        long long r = i / WW;
        int h = r % HH; r /= HH;
        int n = r % g.B;
        int kg = r / g.B;
        float a[8], c[8];
        if (ss && (act || feat)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { a[j] = __ldg(ss + kg * 8 + j); c[j] = __ldg(ss + C + kg * 8 + j); }
        }
        if (POOL) {
            float mx[8];
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    int p = row_of(g, n, 2 * h + dy, 2 * w + dx);
                    float v[8];
                    Row8<T>::load(plane_row(lin, kg, g.P, p), v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) mx[j] = (dy == 0 && dx == 0) ? v[j] : fmaxf(mx[j], v[j]);
                    if (act && ss) {
                        float o[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = fmaxf(fmaf(a[j], v[j], c[j]), 0.f);
                        Row8<T>::store(plane_row(act, kg, g.P, p), o);
                    }
                }
            Row8<T>::store(plane_row(pooled, kg, gp.P, row_of(gp, n, h, w)), mx);
        } else {
            int p = row_of(g, n, h, w);
            float v[8], o[8];
            Row8<T>::load(plane_row(lin, kg, g.P, p), v);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaxf(fmaf(a[j], v[j], c[j]), 0.f);
            if (act) Row8<T>::store(plane_row(act, kg, g.P, p), o);
            if (feat) Row8<T>::store(plane_row(feat, (h * g.W + w) * KG + kg, Balloc, n), o);
        }
    }
}

extern "C" int mpnn_bn_relu_pool_fwd(const void* lin, int C, int B, int H, int W, int G, int P,
                                     const float* ss, void* act, void* pooled, int Pp,
                                     void* feat, int Balloc, int dtype, void* stream) {
    MPNN_REQUIRE(C % 8 == 0, "bn_relu_pool_fwd: C=%d", C);
    MPNN_REQUIRE(!(pooled && feat), "bn_relu_pool_fwd: pooled and feat are exclusive");
    MPNN_REQUIRE(pooled || ss, "bn_relu_pool_fwd: nothing to do");
    MPNN_REQUIRE(!pooled || (H % 2 == 0 && W % 2 == 0), "bn_relu_pool_fwd: odd size");
    Geom g = make_geom(B, H, W, G, P);
    Geom gp = make_geom(B, H / 2, W / 2, G, Pp);
    long long total = (long long)(C / 8) * B * (pooled ? (H / 2) * (W / 2) : H * W);
    int grid = (int)((total + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid < 1) grid = 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (pooled) {
        MPNN_DISPATCH_DTYPE(dtype, (bn_relu_pool_fwd_kernel<T, true><<<grid, 256, 0, st>>>(
            (const T*)lin, C, g, ss, (T*)act, (T*)pooled, gp, (T*)feat, Balloc)));
    } else {
        MPNN_DISPATCH_DTYPE(dtype, (bn_relu_pool_fwd_kernel<T, false><<<grid, 256, 0, st>>>(
            (const T*)lin, C, g, ss, (T*)act, (T*)pooled, gp, (T*)feat, Balloc)));
    }
    return mpnn_check_launch("bn_relu_pool_fwd");
}

