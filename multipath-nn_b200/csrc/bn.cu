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

// ------------------------------------------------------------------ backward
// dy'[j] for one row: relu-masked incoming gradient; also xhat.
template <typename T>
__device__ __forceinline__ void bn_row_grad(const T* __restrict__ lin, const T* __restrict__ dAct,
                                            const T* __restrict__ dFeat, int Balloc, int C,
                                            const Geom& g, int kg, int n, int h, int w,
                                            const float* a, const float* c, const float* mean,
                                            const float* rstd, float dy[8], float xh[8], float lv[8]) {
    const int KG = C / 8;
    int p = row_of(g, n, h, w);
    Row8<T>::load(plane_row(lin, kg, g.P, p), lv);
    float d1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) d1[j] = 0.f;
    if (dAct) Row8<T>::load(plane_row(dAct, kg, g.P, p), d1);
    if (dFeat) {
        float d2[8];
        Row8<T>::load(plane_row(dFeat, (h * g.W + w) * KG + kg, Balloc, n), d2);
#pragma unroll
        for (int j = 0; j < 8; ++j) d1[j] += d2[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float ypre = fmaf(a[j], lv[j], c[j]);
        dy[j] = ypre > 0.f ? d1[j] : 0.f;
        xh[j] = (lv[j] - mean[j]) * rstd[j];
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const T* __restrict__ lin, const T* __restrict__ dAct, const T* __restrict__ dFeat,
                     int Balloc, const float* __restrict__ ss, const float* __restrict__ mr, int C, Geom g,
                     float* __restrict__ partials) {
    const int kg = blockIdx.y;
    float a[8], c[8], mean[8], rstd[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        a[j] = ss[kg * 8 + j]; c[j] = ss[C + kg * 8 + j];
        mean[j] = mr[kg * 8 + j]; rstd[j] = mr[C + kg * 8 + j];
    }
    float s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s0[j] = 0.f; s1[j] = 0.f; }
    const int total = g.B * g.H * g.W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int w = i % g.W;
        int r = i / g.W;
        int h = r % g.H;
        int n = r / g.H;
        float dy[8], xh[8], lv[8];
        bn_row_grad<T>(lin, dAct, dFeat, Balloc, C, g, kg, n, h, w, a, c, mean, rstd, dy, xh, lv);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s0[j] += dy[j]; s1[j] += dy[j] * xh[j]; }
    }
    __shared__ float red[8][16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float t0 = warp_sum(s0[j]), t1 = warp_sum(s1[j]);
        if (lane == 0) { red[warp][j] = t0; red[warp][8 + j] = t1; }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        float t = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) t += red[wv][threadIdx.x];
        int which = threadIdx.x / 8, j = threadIdx.x % 8;
        partials[((size_t)blockIdx.x * 2 + which) * C + kg * 8 + j] = t;
    }
}

extern "C" int mpnn_bn_bwd_reduce(const void* lin, const void* dAct, const void* dFeat, int Balloc,
                                  const float* ss, const float* mr, int C,
                                  int B, int H, int W, int G, int P,
                                  float* partials, int cap, int* n_parts, int dtype, void* stream) {
    MPNN_REQUIRE(C % 8 == 0 && cap > 0, "bn_bwd_reduce: args");
    Geom g = make_geom(B, H, W, G, P);
    int total = B * H * W;
    int gx = ceil_div(total, 256 * 4);
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    if (n_parts) *n_parts = gx;
    dim3 grid(gx, C / 8);
    MPNN_DISPATCH_DTYPE(dtype, (bn_bwd_reduce_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const T*)lin, (const T*)dAct, (const T*)dFeat, Balloc, ss, mr, C, g, partials)));
    return mpnn_check_launch("bn_bwd_reduce");
}

__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partials, int n_parts, int C,
                                       float* __restrict__ sums, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // warp per channel
    const int lane = threadIdx.x & 31;
    if (c >= C) return;
    double s0 = 0.0, s1 = 0.0;
    for (int i = lane; i < n_parts; i += 32) {
        s0 += (double)partials[((size_t)i * 2) * C + c];
        s1 += (double)partials[((size_t)i * 2 + 1) * C + c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane != 0) return;
    sums[c] = (float)s0;
    sums[C + c] = (float)s1;
    if (dgamma) dgamma[c] += (float)s1;
    if (dbeta) dbeta[c] += (float)s0;
}

extern "C" int mpnn_bn_bwd_finalize(const float* partials, int n_parts, int C,
                                    float* sums, float* dgamma, float* dbeta, void* stream) {
    bn_bwd_finalize_kernel<<<ceil_div(C, 4), 128, 0, (cudaStream_t)stream>>>(
        partials, n_parts, C, sums, dgamma, dbeta);
    return mpnn_check_launch("bn_bwd_finalize");
}

template <typename T, bool POOL>
__global__ void bn_relu_pool_bwd_kernel(const T* __restrict__ lin, const T* __restrict__ dAct,
                                        const T* __restrict__ dFeat, int Balloc,
                                        const T* __restrict__ dPooled, Geom gp,
                                        const float* __restrict__ ss, const float* __restrict__ mr,
                                        const float* __restrict__ sums, float inv_count,
                                        int C, Geom g, T* __restrict__ dLin, float* __restrict__ dbias) {
    // grid: x strides over pixels (or 2x2 blocks), y = 8-channel plane
    const int kg = blockIdx.y;
    const int HH = POOL ? g.H / 2 : g.H, WW = POOL ? g.W / 2 : g.W;
    const int total = g.B * HH * WW;
    float a[8], c[8], mean[8], rstd[8], m0[8], m1[8], bs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bs[j] = 0.f;
    if (ss) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a[j] = __ldg(ss + kg * 8 + j); c[j] = __ldg(ss + C + kg * 8 + j);
            mean[j] = __ldg(mr + kg * 8 + j); rstd[j] = __ldg(mr + C + kg * 8 + j);
            m0[j] = __ldg(sums + kg * 8 + j) * inv_count;
            m1[j] = __ldg(sums + C + kg * 8 + j) * inv_count;
        }
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int w = i % WW;
        int r = i / WW;
        int h = r % HH;
        int n = r / HH;
        if (POOL) {
            float out[4][8], lv[4][8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int hh = 2 * h + (k >> 1), ww = 2 * w + (k & 1);
                if (ss) {
                    float dy[8], xh[8];
                    bn_row_grad<T>(lin, dAct, dFeat, Balloc, C, g, kg, n, hh, ww, a, c, mean, rstd, dy, xh, lv[k]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) out[k][j] = a[j] * (dy[j] - m0[j] - xh[j] * m1[j]);
                } else {
                    Row8<T>::load(plane_row(lin, kg, g.P, row_of(g, n, hh, ww)), lv[k]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) out[k][j] = 0.f;
                }
            }
            if (dPooled) {
                float dp[8];
                Row8<T>::load(plane_row(dPooled, kg, gp.P, row_of(gp, n, h, w)), dp);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    int best = 0; float bv = lv[0][j];
#pragma unroll
                    for (int k = 1; k < 4; ++k) if (lv[k][j] > bv) { bv = lv[k][j]; best = k; }
#pragma unroll
                    for (int k = 0; k < 4; ++k) if (k == best) out[k][j] += dp[j];
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int hh = 2 * h + (k >> 1), ww = 2 * w + (k & 1);
                Row8<T>::store(plane_row(dLin, kg, g.P, row_of(g, n, hh, ww)), out[k]);
#pragma unroll
                for (int j = 0; j < 8; ++j) bs[j] += out[k][j];
            }
        } else {
            float dy[8], xh[8], lv[8], out[8];
            bn_row_grad<T>(lin, dAct, dFeat, Balloc, C, g, kg, n, h, w, a, c, mean, rstd, dy, xh, lv);
#pragma unroll
            for (int j = 0; j < 8; ++j) { out[j] = a[j] * (dy[j] - m0[j] - xh[j] * m1[j]); bs[j] += out[j]; }
            Row8<T>::store(plane_row(dLin, kg, g.P, row_of(g, n, h, w)), out);
        }
    }
    if (dbias) {
        // conv bias gradient = column sums of dLin (layer_types.py:181-185: b_k is added before BN)
        __shared__ float red[8][8];
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float t = warp_sum(bs[j]);
            if (lane == 0) red[warp][j] = t;
        }
        __syncthreads();
        if (threadIdx.x < 8) {
            float t = 0.f;
            for (int wv = 0; wv < 8; ++wv) t += red[wv][threadIdx.x];
            atomicAdd(dbias + kg * 8 + threadIdx.x, t);
        }
    }
}

extern "C" int mpnn_bn_relu_pool_bwd(const void* lin, const void* dAct, const void* dFeat, int Balloc,
                                     const void* dPooled, int Pp,
                                     const float* ss, const float* mr, const float* sums, double count,
                                     int C, int B, int H, int W, int G, int P,
                                     void* dLin, float* dbias, int dtype, void* stream) {
    MPNN_REQUIRE(C % 8 == 0, "bn_relu_pool_bwd: C=%d", C);
    MPNN_REQUIRE(ss || dPooled, "bn_relu_pool_bwd: nothing to do");
    MPNN_REQUIRE(!ss || (mr && sums), "bn_relu_pool_bwd: missing stats");
    Geom g = make_geom(B, H, W, G, P);
    Geom gp = make_geom(B, H / 2, W / 2, G, Pp);
    const bool pool = dPooled != nullptr;
    MPNN_REQUIRE(!pool || (H % 2 == 0 && W % 2 == 0), "bn_relu_pool_bwd: odd size");
    long long total = (long long)B * (pool ? (H / 2) * (W / 2) : H * W);
    int gx = (int)((total + 255) / 256);
    int cap = 148 * 16 / (C / 8);
    if (cap < 148) cap = 148;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    dim3 grid(gx, C / 8);
    float inv = (float)(1.0 / count);
    cudaStream_t st = (cudaStream_t)stream;
    if (pool) {
        MPNN_DISPATCH_DTYPE(dtype, (bn_relu_pool_bwd_kernel<T, true><<<grid, 256, 0, st>>>(
            (const T*)lin, (const T*)dAct, (const T*)dFeat, Balloc, (const T*)dPooled, gp, ss, mr, sums,
            inv, C, g, (T*)dLin, dbias)));
    } else {
        MPNN_DISPATCH_DTYPE(dtype, (bn_relu_pool_bwd_kernel<T, false><<<grid, 256, 0, st>>>(
            (const T*)lin, (const T*)dAct, (const T*)dFeat, Balloc, (const T*)dPooled, gp, ss, mr, sums,
            inv, C, g, (T*)dLin, dbias)));
    }
    return mpnn_check_launch("bn_relu_pool_bwd");
}
