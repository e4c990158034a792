// ActivityError (lib/layer_types.py:287-293): c_mod[b] = alpha * sum over (h, w, c) of x^2, a PER-EXAMPLE cost on the
// activations of a layer.  Forward: the cost (for the objective value); backward: its gradient added to the
// gradient wrt the activations, weighted like every other cost of the node (1/B, times the node's p_tr in a
// dynamically-routed net).
#include "common.cuh"
#include "../../include/mpnn.h"

// one CTA per example: the rows of an image are contiguous in every plane
template <typename T>
__global__ void __launch_bounds__(256)
activity_fwd_kernel(const T* __restrict__ x, int C, Geom g, float alpha, float* __restrict__ cost) {
    const int n = blockIdx.x, KG = C / 8;
    float s = 0.f;
    for (int i = threadIdx.x; i < KG * g.H * g.W; i += blockDim.x) {
        const int px = i % (g.H * g.W), kg = i / (g.H * g.W);
        float v[8];
        Row8<T>::load(plane_row(x, kg, g.P, row_of(g, n, px / g.W, px % g.W)), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) s = fmaf(v[j], v[j], s);
    }
    __shared__ float red[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        cost[n] = alpha * t;
    }
}

extern "C" int mpnn_activity_fwd(const void* x, int C, int B, int H, int W, int G, int P, float alpha,
                                 float* cost, int dtype, void* stream) {
    MPNN_REQUIRE(x && cost && C % 8 == 0 && B >= 1, "activity_fwd: args");
    Geom g = make_geom(B, H, W, G, P);
    MPNN_DISPATCH_DTYPE(dtype, (activity_fwd_kernel<T><<<B, 256, 0, (cudaStream_t)stream>>>((const T*)x, C, g, alpha, cost)));
    return mpnn_check_launch("activity_fwd");
}

// dx (+)= scale * coef[n] * x over the pixels of the images (coef == NULL: 1); acc == 0 writes, acc != 0 adds
template <typename T>
__global__ void __launch_bounds__(256)
activity_bwd_kernel(const T* __restrict__ x, int C, Geom g, const float* __restrict__ coef, float scale,
                    T* __restrict__ dx, int acc) {
    const int kg = blockIdx.y;
    const int total = g.B * g.H * g.W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int px = i % (g.H * g.W), n = i / (g.H * g.W);
        const int p = row_of(g, n, px / g.W, px % g.W);
        const float k = scale * (coef ? coef[n] : 1.f);
        float v[8], d[8];
        Row8<T>::load(plane_row(x, kg, g.P, p), v);
        if (acc) {
            Row8<T>::load(plane_row(dx, kg, g.P, p), d);
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = fmaf(k, v[j], d[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = k * v[j];
        }
        Row8<T>::store(plane_row(dx, kg, g.P, p), d);
    }
}

extern "C" int mpnn_activity_bwd(const void* x, int C, int B, int H, int W, int G, int P, const float* coef,
                                 float scale, void* dx, int acc, int dtype, void* stream) {
    MPNN_REQUIRE(x && dx && C % 8 == 0, "activity_bwd: args");
    Geom g = make_geom(B, H, W, G, P);
    const long long total = (long long)B * H * W;
    int gx = (int)((total + 255) / 256);
    if (gx > 148 * 8) gx = 148 * 8;
    MPNN_DISPATCH_DTYPE(dtype, (activity_bwd_kernel<T><<<dim3(gx, C / 8), 256, 0, (cudaStream_t)stream>>>(
        (const T*)x, C, g, coef, scale, (T*)dx, acc)));
    return mpnn_check_launch("activity_bwd");
}

// ---------------------------------------------------------------------------------------------------------
// Dropout (lib/layer_types.py:212-217): tf.nn.dropout(x, keep_prob) = x * m / keep_prob with m ~ Bernoulli(keep_prob)
// per element, redrawn at every evaluation and in every mode (the reference does not gate it on 'tr').
// TF's random stream cannot be reproduced; the mask here is a counter-based hash of (seed, draw, element index)
//   u = mix(mix(index * 0x9E3779B9 + seed) ^ (draw * 0x85EBCA6B)),  keep <=> u < keep_prob * 2^32
// with mix = the 32-bit finaliser  x ^= x >> 16; x *= 0x7feb352d; x ^= x >> 15; x *= 0x846ca68b; x ^= x >> 16
// (tests/util.py restates it in numpy for the oracle).  index = ((n * H + h) * W + w) * C + c, so a planes tensor
// and its flattened feature copy see the same mask; `draw` is read from hyp[MPNN_HYP_DRAW] (bits of a uint32).
// The same launch serves the backward pass (scaling by m / keep_prob is its own adjoint).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t drop_mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ bool drop_keep(uint32_t idx, uint32_t seed, uint32_t draw, uint32_t thresh) {
    return drop_mix(drop_mix(idx * 0x9E3779B9u + seed) ^ (draw * 0x85EBCA6Bu)) < thresh;
}

// x: planes [C/8][P][8] (feat == 0) or flattened features [H*W*C/8][Balloc][8] (feat != 0), scaled in place
template <typename T>
__global__ void __launch_bounds__(256)
dropout_kernel(T* __restrict__ x, int C, Geom g, int feat, int Balloc, float keep, uint32_t seed,
               const float* __restrict__ hyp) {
    const int kg = blockIdx.y, KG = C / 8;
    const uint32_t draw = __float_as_uint(hyp[MPNN_HYP_DRAW]);
    const uint32_t thresh = keep >= 1.f ? 0xFFFFFFFFu : (uint32_t)((double)keep * 4294967296.0);
    const float inv = 1.f / keep;
    const int total = g.B * g.H * g.W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int px = i % (g.H * g.W), n = i / (g.H * g.W);
        T* row = feat ? plane_row(x, px * KG + kg, Balloc, n) : plane_row(x, kg, g.P, row_of(g, n, px / g.W, px % g.W));
        float v[8];
        Row8<T>::load(row, v);
        const uint32_t base = ((uint32_t)i * (uint32_t)C) + (uint32_t)kg * 8u;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = drop_keep(base + j, seed, draw, thresh) ? v[j] * inv : 0.f;
        Row8<T>::store(row, v);
    }
}

extern "C" int mpnn_dropout(void* x, int C, int B, int H, int W, int G, int P, int feat, int Balloc,
                            float keep, unsigned seed, const float* hyp, int dtype, void* stream) {
    MPNN_REQUIRE(x && hyp && C % 8 == 0 && keep > 0.f && keep <= 1.f, "dropout: C=%d keep=%g", C, (double)keep);
    MPNN_REQUIRE((long long)B * H * W * C < (1ll << 32), "dropout: tensor too large for the 32-bit element counter");
    Geom g = make_geom(B, H, W, G, P);
    const long long total = (long long)B * H * W;
    int gx = (int)((total + 255) / 256);
    if (gx > 148 * 8) gx = 148 * 8;
    MPNN_DISPATCH_DTYPE(dtype, (dropout_kernel<T><<<dim3(gx, C / 8), 256, 0, (cudaStream_t)stream>>>(
        (T*)x, C, g, feat, Balloc, keep, (uint32_t)seed, hyp)));
    return mpnn_check_launch("dropout");
}
