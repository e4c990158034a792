// "Last CTA finalises": BatchNorm batch statistics without a second launch.
//
// A producer kernel (conv forward, BN-backward reduce) ends every CTA with
//   1. fp64 atomicAdd of the CTA's per-channel partial sums into acc[2][C],
//   2. __threadfence + one ticket (acc[2*C] reinterpreted as unsigned),
// and the CTA that draws the last ticket reads the totals back from L2, writes
// the derived per-channel constants, and zeroes acc and the ticket for the next
// launch (self-cleaning: the buffer is memset once at allocation).  fp64 sums of
// a few hundred fp32 partials are order-independent far below fp32 resolution,
// so the fp32 results are reproducible although the atomics are unordered.
#pragma once
#include "common.cuh"

// val(i), i in [0, 2*nb): partial sum `i / nb` of column n0 + i % nb held by this CTA.
// Returns true in every thread of the CTA that drew the last of n_ctas tickets.
template <typename F>
__device__ __forceinline__ bool mpnn_acc_and_ticket(double* acc, int C, int n0, int nb, unsigned n_ctas, F val) {
    __shared__ int s_last;
    for (int i = threadIdx.x; i < 2 * nb; i += blockDim.x)
        atomicAdd(acc + (size_t)(i / nb) * C + n0 + i % nb, (double)val(i));
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(reinterpret_cast<unsigned*>(acc + 2 * C), 1u);
        s_last = (t == n_ctas - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}

// deferred variant: only the fp64 adds; the consumer kernel derives the constants from the totals
template <typename F>
__device__ __forceinline__ void mpnn_acc_only(double* acc, int C, int n0, int nb, F val) {
    for (int i = threadIdx.x; i < 2 * nb; i += blockDim.x)
        atomicAdd(acc + (size_t)(i / nb) * C + n0 + i % nb, (double)val(i));
}

// scale / shift of channel c from the accumulated totals (the arithmetic of mpnn_bn_fwd_finalize_last)
__device__ __forceinline__ void mpnn_bn_consts_from_acc(const mpnn_bn_fuse& f, int C, int c, float& a, float& sh,
                                                        float& mean, float& var, float& rstd) {
    const double s = f.acc[c], s2 = f.acc[C + c];
    const double m = s / f.count;
    double v = s2 / f.count - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m; var = (float)v;
    rstd = 1.0f / sqrtf(var + f.eps);
    a = f.gamma[c] * rstd;
    sh = f.beta[c] - mean * a;
}

// forward finalisation by the last CTA (lib/layer_types.py:219-249): batch moments ->
// scale/shift (ss), mean/rstd (mr), running averages; then reset acc + ticket.
__device__ __forceinline__ void mpnn_bn_fwd_finalize_last(const mpnn_bn_fuse& f, int C) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double s = __ldcg(f.acc + c), s2 = __ldcg(f.acc + C + c);
        f.acc[c] = 0.0; f.acc[C + c] = 0.0;
        const double m = s / f.count;
        double v = s2 / f.count - m * m;
        if (v < 0.0) v = 0.0;
        const float mean = (float)m, var = (float)v;
        if (f.m_avg) {
            f.m_avg[c] = f.d * f.m_avg[c] + (1.f - f.d) * mean;
            f.v_avg[c] = f.d * f.v_avg[c] + (1.f - f.d) * var;
        }
        const float rstd = 1.0f / sqrtf(var + f.eps);
        const float a = f.gamma[c] * rstd;
        f.ss[c] = a;
        f.ss[C + c] = f.beta[c] - mean * a;
        f.mr[c] = mean;
        f.mr[C + c] = rstd;
    }
    if (threadIdx.x == 0) *reinterpret_cast<unsigned*>(f.acc + 2 * C) = 0u;
}

// backward finalisation by the last CTA: totals of dy' and dy'*(x - mean) -> sums in the xhat form
// (sums[0] = sum dy', sums[1] = sum dy'*xhat), dgamma / dbeta accumulated; then reset acc + ticket.
__device__ __forceinline__ void mpnn_bn_bwd_finalize_last(const mpnn_bn_bwd_fuse& f, const float* mr, int C) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double t0 = __ldcg(f.acc + c), t1 = __ldcg(f.acc + C + c);
        f.acc[c] = 0.0; f.acc[C + c] = 0.0;
        const double rstd = mr[C + c];
        const double sx = rstd * t1;                         // sum dy'*xhat (t1 = sum dy'*(x - mean))
        f.sums[c] = (float)t0;
        f.sums[C + c] = (float)sx;
        if (f.dgamma) f.dgamma[c] += (float)sx;
        if (f.dbeta) f.dbeta[c] += (float)t0;
    }
    if (threadIdx.x == 0) *reinterpret_cast<unsigned*>(f.acc + 2 * C) = 0u;
}
