// BatchNorm (+ReLU, +2x2 max-pool) backward on padded-planes tensors.
// Reference: TF autodiff of lib/layer_types.py:109-110 (pool), :196-199 (ReLU),
// :219-249 (train-mode BN: gradient flows through the batch moments).
//
//   y = relu(a*x + c),  a = gamma*rstd,  c = beta - mean*a,  xhat = (x - mean)*rstd
//   dy' = dy * [a*x + c > 0]
//   dx  = a*(dy' - mean(dy') - xhat*mean(dy'*xhat))  + unpool(dPooled)
//       = a*dy' + p*x + q,   p = -a*rstd*m1,  q = -a*m0 + a*rstd*mean*m1     (m0,m1 = batch means)
// Both kernels are pure HBM streams, so they are written for occupancy: per-thread
// state is four 8-vectors of constants plus the rows in flight (<= 64 registers).
#include "common.cuh"
#include "../../include/mpnn.h"
#include "bn_fuse.cuh"

// incoming gradient of one pixel row: dAct + dFeat (either may be absent)
template <typename T>
__device__ __forceinline__ void load_dy(const T* __restrict__ dAct, const T* __restrict__ dFeat, int Balloc,
                                        int KG, const Geom& g, int kg, int n, int h, int w, int p, float d[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = 0.f;
    if (dAct) Row8<T>::load(plane_row(dAct, kg, g.P, p), d);
    if (dFeat) {
        float d2[8];
        Row8<T>::load(plane_row(dFeat, (h * g.W + w) * KG + kg, Balloc, n), d2);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] += d2[j];
    }
}

// pass 1: per-CTA partial sums of dy' and dy'*(x - mean) (scaled by rstd into the xhat form in finalize)
template <typename T>
__global__ void __launch_bounds__(256, 4)
bn_bwd_reduce_kernel(const T* __restrict__ lin, const T* __restrict__ dAct, const T* __restrict__ dFeat,
                     int Balloc, const float* __restrict__ ss, int C, Geom g, float* __restrict__ partials,
                     const float* __restrict__ mr, const mpnn_bn_bwd_fuse f) {
    pdl_launch_dependents();
    pdl_wait();
    const int kg = blockIdx.y, KG = C / 8;
    float a[8], c[8], mu[8], s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        a[j] = ss[kg * 8 + j]; c[j] = ss[C + kg * 8 + j];
        mu[j] = mr[kg * 8 + j];       // the second sum is taken of dy'*(x - mean): no cancellation against mean * sum dy'
        s0[j] = 0.f; s1[j] = 0.f;
    }
    const int total = g.B * g.H * g.W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int w = i % g.W;
        const int r = i / g.W;
        const int h = r % g.H;
        const int n = r / g.H;
        const int p = row_of(g, n, h, w);
        float lv[8], d[8];
        Row8<T>::load(plane_row(lin, kg, g.P, p), lv);
        load_dy<T>(dAct, dFeat, Balloc, KG, g, kg, n, h, w, p, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float dy = fmaf(a[j], lv[j], c[j]) > 0.f ? d[j] : 0.f;
            s0[j] += dy;
            s1[j] = fmaf(dy, lv[j] - mu[j], s1[j]);
        }
    }
    __shared__ float red[8][16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float t0 = warp_sum(s0[j]), t1 = warp_sum(s1[j]);
        if (lane == 0) { red[warp][j] = t0; red[warp][8 + j] = t1; }
    }
    __syncthreads();
    if (partials && threadIdx.x < 16) {
        float t = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) t += red[wv][threadIdx.x];
        const int which = threadIdx.x / 8, j = threadIdx.x % 8;
        partials[((size_t)blockIdx.x * 2 + which) * C + kg * 8 + j] = t;
    }
    if (f.acc) {
        // fused finalisation (see bn_fuse.cuh): totals of dy' and dy'*x -> sums in the xhat form
        const bool last = mpnn_acc_and_ticket(f.acc, C, kg * 8, 8, gridDim.x * gridDim.y, [&](int i) {
            float t = 0.f;
#pragma unroll
            for (int wv = 0; wv < 8; ++wv) t += red[wv][i];
            return t; });
        if (last) mpnn_bn_bwd_finalize_last(f, mr, C);
    }
}

static int bn_bwd_reduce_impl(const void* lin, const void* dAct, const void* dFeat, int Balloc,
                              const float* ss, const float* mr, int C,
                              int B, int H, int W, int G, int P,
                              float* partials, int cap, int* n_parts, const mpnn_bn_bwd_fuse* fp,
                              int dtype, void* stream) {
    MPNN_REQUIRE(C % 8 == 0 && ss && mr && (fp || cap > 0), "bn_bwd_reduce: args");
    MPNN_REQUIRE(!fp || (fp->acc && fp->sums), "bn_bwd_reduce_fused: incomplete mpnn_bn_bwd_fuse");
    Geom g = make_geom(B, H, W, G, P);
    int total = B * H * W;
    // latency-bound when small: one pixel per thread as long as that is at most ~4 CTAs per SM, else four
    int gx = ceil_div(total, 256);
    if ((long long)gx * (C / 8) > 148 * 4) gx = ceil_div(total, 256 * 4);
    int lim = 148 * 8 / (C / 8);
    if (lim < 74) lim = 74;
    if (gx > lim) gx = lim;
    if (partials && gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    if (n_parts) *n_parts = gx;
    mpnn_bn_bwd_fuse f = {};
    if (fp) f = *fp;
    dim3 grid(gx, C / 8);
    MPNN_DISPATCH_DTYPE(dtype, (mpnn_launch_pdl(bn_bwd_reduce_kernel<T>, grid, dim3(256), 0, (cudaStream_t)stream,
        (const T*)lin, (const T*)dAct, (const T*)dFeat, Balloc, ss, C, g, partials, mr, f)));
    return mpnn_check_launch("bn_bwd_reduce");
}

extern "C" int mpnn_bn_bwd_reduce(const void* lin, const void* dAct, const void* dFeat, int Balloc,
                                  const float* ss, const float* mr, int C,
                                  int B, int H, int W, int G, int P,
                                  float* partials, int cap, int* n_parts, int dtype, void* stream) {
    MPNN_REQUIRE(partials, "bn_bwd_reduce: partials is NULL");
    return bn_bwd_reduce_impl(lin, dAct, dFeat, Balloc, ss, mr, C, B, H, W, G, P, partials, cap, n_parts,
                              nullptr, dtype, stream);
}

extern "C" int mpnn_bn_bwd_reduce_fused(const void* lin, const void* dAct, const void* dFeat, int Balloc,
                                        const float* ss, const float* mr, int C,
                                        int B, int H, int W, int G, int P,
                                        const mpnn_bn_bwd_fuse* f, int dtype, void* stream) {
    MPNN_REQUIRE(f, "bn_bwd_reduce_fused: f is NULL");
    return bn_bwd_reduce_impl(lin, dAct, dFeat, Balloc, ss, mr, C, B, H, W, G, P, nullptr, 0, nullptr, f,
                              dtype, stream);
}

// partials hold sum(dy') and sum(dy'*x); sums[0] = sum dy', sums[1] = sum dy'*xhat
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partials, int n_parts, int C,
                                       const float* __restrict__ mr, float* __restrict__ sums,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
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
    const double rstd = mr[C + c];
    const double sx = rstd * s1;                         // sum dy'*xhat (s1 is already centred)
    sums[c] = (float)s0;
    sums[C + c] = (float)sx;
    if (dgamma) dgamma[c] += (float)sx;
    if (dbeta) dbeta[c] += (float)s0;
}

extern "C" int mpnn_bn_bwd_finalize(const float* partials, int n_parts, int C, const float* mr,
                                    float* sums, float* dgamma, float* dbeta, void* stream) {
    bn_bwd_finalize_kernel<<<ceil_div(C, 4), 128, 0, (cudaStream_t)stream>>>(
        partials, n_parts, C, mr, sums, dgamma, dbeta);
    return mpnn_check_launch("bn_bwd_finalize");
}

// pass 2
template <typename T, bool POOL>
__global__ void __launch_bounds__(256, POOL ? 3 : 4)
bn_relu_pool_bwd_kernel(const T* __restrict__ lin, const T* __restrict__ dAct,
                        const T* __restrict__ dFeat, int Balloc,
                        const T* __restrict__ dPooled, Geom gp,
                        const float* __restrict__ ss, const float* __restrict__ mr,
                        const float* __restrict__ sums, float inv_count,
                        int C, Geom g, T* __restrict__ dLin, float* __restrict__ dbias) {
    // grid: x strides over pixels (or 2x2 blocks), y = 8-channel plane
    pdl_launch_dependents();
    pdl_wait();
    const int kg = blockIdx.y, KG = C / 8;
    const int HH = POOL ? g.H / 2 : g.H, WW = POOL ? g.W / 2 : g.W;
    const int total = g.B * HH * WW;
    float a[8], c[8], pp[8], qq[8], bs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { bs[j] = 0.f; a[j] = 0.f; c[j] = 0.f; pp[j] = 0.f; qq[j] = 0.f; }
    if (ss) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a[j] = __ldg(ss + kg * 8 + j); c[j] = __ldg(ss + C + kg * 8 + j);
            const float mean = __ldg(mr + kg * 8 + j), rstd = __ldg(mr + C + kg * 8 + j);
            const float m0 = __ldg(sums + kg * 8 + j) * inv_count, m1 = __ldg(sums + C + kg * 8 + j) * inv_count;
            pp[j] = -a[j] * rstd * m1;
            qq[j] = -a[j] * m0 + a[j] * rstd * mean * m1;
        }
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int w = i % WW;
        const int r = i / WW;
        const int h = r % HH;
        const int n = r / HH;
        if (POOL) {
            // which of the 2x2 pixels holds the (first) maximum, per channel: 2 bits each
            unsigned best = 0;
            float dp[8];
            if (dPooled) {
                float bv[8];
                Row8<T>::load(plane_row(lin, kg, g.P, row_of(g, n, 2 * h, 2 * w)), bv);
#pragma unroll
                for (int k = 1; k < 4; ++k) {
                    float v[8];
                    Row8<T>::load(plane_row(lin, kg, g.P, row_of(g, n, 2 * h + (k >> 1), 2 * w + (k & 1))), v);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (v[j] > bv[j]) { bv[j] = v[j]; best = (best & ~(3u << (2 * j))) | ((unsigned)k << (2 * j)); }
                }
                Row8<T>::load(plane_row(dPooled, kg, gp.P, row_of(gp, n, h, w)), dp);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int hh = 2 * h + (k >> 1), ww = 2 * w + (k & 1);
                const int p = row_of(g, n, hh, ww);
                float out[8];
                if (ss) {
                    float lv[8], d[8];
                    Row8<T>::load(plane_row(lin, kg, g.P, p), lv);      // L1 hit: read above for the argmax
                    load_dy<T>(dAct, dFeat, Balloc, KG, g, kg, n, hh, ww, p, d);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float dy = fmaf(a[j], lv[j], c[j]) > 0.f ? d[j] : 0.f;
                        out[j] = fmaf(a[j], dy, fmaf(pp[j], lv[j], qq[j]));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) out[j] = 0.f;
                }
                if (dPooled) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (((best >> (2 * j)) & 3u) == (unsigned)k) out[j] += dp[j];
                }
                Row8<T>::store(plane_row(dLin, kg, g.P, p), out);
#pragma unroll
                for (int j = 0; j < 8; ++j) bs[j] += out[j];
            }
        } else {
            const int p = row_of(g, n, h, w);
            float lv[8], d[8], out[8];
            Row8<T>::load(plane_row(lin, kg, g.P, p), lv);
            load_dy<T>(dAct, dFeat, Balloc, KG, g, kg, n, h, w, p, d);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float dy = fmaf(a[j], lv[j], c[j]) > 0.f ? d[j] : 0.f;
                out[j] = fmaf(a[j], dy, fmaf(pp[j], lv[j], qq[j]));
                bs[j] += out[j];
            }
            Row8<T>::store(plane_row(dLin, kg, g.P, p), out);
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

// pass 2, v2 (bf16): every row the thread needs (lin, dAct of the 2x2 block, dPooled) is requested
// before the first use and lin is kept in registers instead of being re-read for the outputs; pixel
// index split by shifts for power-of-two sizes.  v1 ran at 49 % of DRAM peak, latency-bound.
struct PixSplitB { int logW, logHH; };
__device__ __forceinline__ uint4 ldg16b(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void unpack8b(const uint4& r, float v[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

__global__ void __launch_bounds__(256, 2)
bn_relu_pool_bwd_v2_kernel(const __nv_bfloat16* __restrict__ lin, const __nv_bfloat16* __restrict__ dAct,
                           const __nv_bfloat16* __restrict__ dFeat, int Balloc,
                           const __nv_bfloat16* __restrict__ dPooled, Geom gp,
                           const float* __restrict__ ss, const float* __restrict__ mr,
                           const float* __restrict__ sums, float inv_count,
                           int C, Geom g, PixSplitB ps, __nv_bfloat16* __restrict__ dLin, float* __restrict__ dbias) {
    typedef __nv_bfloat16 T;
    pdl_launch_dependents();
    pdl_wait();
    const int kg = blockIdx.y;
    const int HH = g.H / 2, WW = g.W / 2;            // one thread per 2x2 block (this kernel is the pooled case)
    const int total = g.B * HH * WW;
    float a[8], c[8], pp[8], qq[8], bs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { bs[j] = 0.f; a[j] = 0.f; c[j] = 0.f; pp[j] = 0.f; qq[j] = 0.f; }
    if (ss) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a[j] = __ldg(ss + kg * 8 + j); c[j] = __ldg(ss + C + kg * 8 + j);
            const float mean = __ldg(mr + kg * 8 + j), rstd = __ldg(mr + C + kg * 8 + j);
            const float m0 = __ldg(sums + kg * 8 + j) * inv_count, m1 = __ldg(sums + C + kg * 8 + j) * inv_count;
            pp[j] = -a[j] * rstd * m1;
            qq[j] = -a[j] * m0 + a[j] * rstd * mean * m1;
        }
    }
    auto split = [&](int i, int& n, int& h, int& w) {
        if (ps.logW >= 0) { w = i & (WW - 1); const int r = i >> ps.logW; h = r & (HH - 1); n = r >> ps.logHH; }
        else { w = i % WW; const int r = i / WW; h = r % HH; n = r / HH; }
    };
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
            int n, h, w;
            split(i, n, h, w);
            const int p00 = row_of(g, n, 2 * h, 2 * w);
            const int pk[4] = {p00, p00 + 1, p00 + g.Wp, p00 + g.Wp + 1};
            uint4 Lr[4], Dr[4], Pr = zero4;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                Lr[k] = ldg16b(plane_row(lin, kg, g.P, pk[k]));
                Dr[k] = (ss && dAct) ? ldg16b(plane_row(dAct, kg, g.P, pk[k])) : zero4;
            }
            if (dPooled) Pr = ldg16b(plane_row(dPooled, kg, gp.P, row_of(gp, n, h, w)));
            // which of the 2x2 pixels holds the (first) maximum, per channel: 2 bits each
            unsigned best = 0;
            float dp[8];
            if (dPooled) {
                float bv[8];
                unpack8b(Lr[0], bv);
#pragma unroll
                for (int k = 1; k < 4; ++k) {
                    float v[8];
                    unpack8b(Lr[k], v);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (v[j] > bv[j]) { bv[j] = v[j]; best = (best & ~(3u << (2 * j))) | ((unsigned)k << (2 * j)); }
                }
                unpack8b(Pr, dp);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float out[8];
                if (ss) {
                    float lv[8], d[8];
                    unpack8b(Lr[k], lv);
                    unpack8b(Dr[k], d);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float dy = fmaf(a[j], lv[j], c[j]) > 0.f ? d[j] : 0.f;
                        out[j] = fmaf(a[j], dy, fmaf(pp[j], lv[j], qq[j]));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) out[j] = 0.f;
                }
                if (dPooled) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (((best >> (2 * j)) & 3u) == (unsigned)k) out[j] += dp[j];
                }
                Row8<T>::store(plane_row(dLin, kg, g.P, pk[k]), out);
#pragma unroll
                for (int j = 0; j < 8; ++j) bs[j] += out[j];
            }
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

// ---------------------------------------------------------------------------
// Small tensors (B*H*W <= 8192 pixels, no pooling branch: the coarsest scale of a stage at the reference's
// batch): reduction AND gradient in ONE launch.  The two-pass pair above is a chain -- partial sums, fp64
// atomics, fence, ticket, last-CTA finalise, a second launch that reads lin / dAct again -- of ~13 us for
// a tensor of <= 0.5 MB, eight times on the critical path of the backward pass.  Here one 8-CTA thread-block
// cluster owns an 8-channel plane: every thread loads its <= 4 pixels once and keeps lin and dy' in
// registers, the per-channel sums go warp -> CTA -> cluster (distributed shared memory, fixed rank order,
// fp64 across CTAs) behind ONE cluster barrier, every CTA derives the constants and writes its part of dLin.
// ---------------------------------------------------------------------------
#include <cooperative_groups.h>
namespace cgb = cooperative_groups;
constexpr int kSmallCL = 8, kSmallT = 256, kSmallRPT = 4;

template <typename T, int RPT>
__global__ void __cluster_dims__(kSmallCL, 1, 1) __launch_bounds__(kSmallT)
bn_bwd_small_kernel(const T* __restrict__ lin, const T* __restrict__ dAct, const T* __restrict__ dFeat, int Balloc,
                    const float* __restrict__ ss, const float* __restrict__ mr, int C, Geom g,
                    float* __restrict__ sums, float* __restrict__ dgamma, float* __restrict__ dbeta,
                    float inv_count, T* __restrict__ dLin, float* __restrict__ dbias) {
    cgb::cluster_group cluster = cgb::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int kg = blockIdx.y, KG = C / 8, tid = threadIdx.x;
    const int total = g.B * g.H * g.W;
    const int per = (total + kSmallCL - 1) / kSmallCL;
    __shared__ float red[kSmallT / 32][16];
    __shared__ float part[16];
    __shared__ double tot[16];
    float a[8], c[8], mu[8], rstd[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        a[j] = __ldg(ss + kg * 8 + j); c[j] = __ldg(ss + C + kg * 8 + j);
        mu[j] = __ldg(mr + kg * 8 + j); rstd[j] = __ldg(mr + C + kg * 8 + j);
    }
    float lv[RPT][8], dy[RPT][8];
    int prow[RPT];
    float s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s0[j] = 0.f; s1[j] = 0.f; }
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
        const int k = u * kSmallT + tid, i = rank * per + k;
        prow[u] = -1;
        if (k < per && i < total) {
            const int w = i % g.W, r = i / g.W, h = r % g.H, n = r / g.H;
            const int p = row_of(g, n, h, w);
            prow[u] = p;
            float d[8];
            Row8<T>::load(plane_row(lin, kg, g.P, p), lv[u]);
            load_dy<T>(dAct, dFeat, Balloc, KG, g, kg, n, h, w, p, d);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float v = fmaf(a[j], lv[u][j], c[j]) > 0.f ? d[j] : 0.f;
                dy[u][j] = v;
                s0[j] += v;
                s1[j] = fmaf(v, lv[u][j] - mu[j], s1[j]);       // centred: no cancellation against mean * sum dy'
            }
        }
    }
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float t0 = warp_sum(s0[j]), t1 = warp_sum(s1[j]);
        if (lane == 0) { red[warp][j] = t0; red[warp][8 + j] = t1; }
    }
    __syncthreads();
    if (tid < 16) {
        float t = 0.f;
#pragma unroll
        for (int wv = 0; wv < kSmallT / 32; ++wv) t += red[wv][tid];
        part[tid] = t;
    }
    cluster.sync();
    if (tid < 16) {
        double t = 0.0;
        for (int k = 0; k < kSmallCL; ++k) t += (double)cluster.map_shared_rank(&part[0], k)[tid];
        tot[tid] = t;
    }
    __syncthreads();
    float pp[8], qq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double t0 = tot[j], sx = (double)rstd[j] * tot[8 + j];        // sum dy', sum dy'*xhat
        const float m0 = (float)t0 * inv_count, m1 = (float)sx * inv_count;
        pp[j] = -a[j] * rstd[j] * m1;
        qq[j] = -a[j] * m0 + a[j] * rstd[j] * mu[j] * m1;
    }
    if (rank == 0 && tid < 8) {
        const double t0 = tot[tid], sx = (double)__ldg(mr + C + kg * 8 + tid) * tot[8 + tid];
        if (sums) { sums[kg * 8 + tid] = (float)t0; sums[C + kg * 8 + tid] = (float)sx; }
        if (dgamma) dgamma[kg * 8 + tid] += (float)sx;
        if (dbeta) dbeta[kg * 8 + tid] += (float)t0;
    }
    float bs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bs[j] = 0.f;
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
        if (prow[u] >= 0) {
            float out[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                out[j] = fmaf(a[j], dy[u][j], fmaf(pp[j], lv[u][j], qq[j]));
                bs[j] += out[j];
            }
            Row8<T>::store(plane_row(dLin, kg, g.P, prow[u]), out);
        }
    }
    if (dbias) {
        // conv bias gradient = column sums of dLin (layer_types.py:181-185: b_k is added before BN)
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float t = warp_sum(bs[j]);
            if (lane == 0) red[warp][j] = t;
        }
        __syncthreads();
        if (tid < 8) {
            float t = 0.f;
#pragma unroll
            for (int wv = 0; wv < kSmallT / 32; ++wv) t += red[wv][tid];
            atomicAdd(dbias + kg * 8 + tid, t);
        }
    }
    cluster.sync();          // keep this CTA's shared memory alive until every peer has read it
}

static_assert(kSmallCL * kSmallT * kSmallRPT == MPNN_BN_SMALL_MAX_PIXELS, "bn_bwd_small capacity");

extern "C" int mpnn_bn_bwd_small(const void* lin, const void* dAct, const void* dFeat, int Balloc,
                                 const float* ss, const float* mr, int C, int B, int H, int W, int G, int P,
                                 float* sums, float* dgamma, float* dbeta, double count,
                                 void* dLin, float* dbias, int dtype, void* stream) {
    MPNN_REQUIRE(C % 8 == 0 && lin && ss && mr && dLin && (dAct || dFeat) && count > 0, "bn_bwd_small: args");
    const long long total = (long long)B * H * W;
    MPNN_REQUIRE(total <= (long long)kSmallCL * kSmallT * kSmallRPT,
                 "bn_bwd_small: %lld pixels (at most %d: use mpnn_bn_bwd_reduce_fused + mpnn_bn_relu_pool_bwd)", total,
                 kSmallCL * kSmallT * kSmallRPT);
    Geom g = make_geom(B, H, W, G, P);
    const int per = ((int)total + kSmallCL - 1) / kSmallCL;
    const int rpt = (per + kSmallT - 1) / kSmallT;
    dim3 grid(kSmallCL, C / 8);
    const float inv = (float)(1.0 / count);
    cudaStream_t st = (cudaStream_t)stream;
#define MPNN_BN_SMALL_LAUNCH(R)                                                                             \
    MPNN_DISPATCH_DTYPE(dtype, (bn_bwd_small_kernel<T, R><<<grid, kSmallT, 0, st>>>(                        \
        (const T*)lin, (const T*)dAct, (const T*)dFeat, Balloc, ss, mr, C, g, sums, dgamma, dbeta, inv,     \
        (T*)dLin, dbias)))
    if (rpt <= 1) { MPNN_BN_SMALL_LAUNCH(1); }
    else if (rpt <= 2) { MPNN_BN_SMALL_LAUNCH(2); }
    else { MPNN_BN_SMALL_LAUNCH(4); }
#undef MPNN_BN_SMALL_LAUNCH
    return mpnn_check_launch("bn_bwd_small");
}

static inline int ilog2_exact_b(int v) {
    if (v <= 0 || (v & (v - 1))) return -1;
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
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
    static const int v2 = getenv("MPNN_BN_V2") ? atoi(getenv("MPNN_BN_V2")) : 1;
    if (v2 && dtype == MPNN_BF16 && pool && !dFeat) {
        PixSplitB ps = {ilog2_exact_b(W / 2), ilog2_exact_b(H / 2)};
        if (ps.logHH < 0) ps.logW = -1;
        typedef __nv_bfloat16 T;
        mpnn_launch_pdl(bn_relu_pool_bwd_v2_kernel, grid, dim3(256), 0, st,
            (const T*)lin, (const T*)dAct, (const T*)dFeat, Balloc, (const T*)dPooled, gp, ss, mr, sums,
            inv, C, g, ps, (T*)dLin, dbias);
        return mpnn_check_launch("bn_relu_pool_bwd");
    }
    if (pool) {
        MPNN_DISPATCH_DTYPE(dtype, (mpnn_launch_pdl(bn_relu_pool_bwd_kernel<T, true>, grid, dim3(256), 0, st,
            (const T*)lin, (const T*)dAct, (const T*)dFeat, Balloc, (const T*)dPooled, gp, ss, mr, sums,
            inv, C, g, (T*)dLin, dbias)));
    } else {
        MPNN_DISPATCH_DTYPE(dtype, (mpnn_launch_pdl(bn_relu_pool_bwd_kernel<T, false>, grid, dim3(256), 0, st,
            (const T*)lin, (const T*)dAct, (const T*)dFeat, Balloc, (const T*)dPooled, gp, ss, mr, sums,
            inv, C, g, (T*)dLin, dbias)));
    }
    return mpnn_check_launch("bn_relu_pool_bwd");
}
