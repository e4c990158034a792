// minimize_expectation + MomentumOptimizer as one fused multi-tensor kernel
// over the flat parameter / gradient / momentum buffers.
// Reference: lib/net_types.py:24-37 (TALR scaling), :96-97,178-181 (momentum,
// non-Nesterov: a <- mu*a + g ; theta <- theta - lr*a), L2 terms
// lib/layer_types.py:52,72,186-188 weighted by sg(p_tr) (net_types.py:171-173).
#include "common.cuh"
#include "../../include/mpnn.h"

__global__ void talr_momentum_kernel(float* __restrict__ theta, const float* __restrict__ grad,
                                     float* __restrict__ accum, int n,
                                     const int* __restrict__ seg_start, const int* __restrict__ seg_node,
                                     const float* __restrict__ seg_mult, const float* __restrict__ seg_l2,
                                     int n_seg, const float* __restrict__ node_stats, int talr,
                                     const float* __restrict__ hyp) {
    const float lr = hyp[MPNN_HYP_LR], mu = hyp[MPNN_HYP_MU], grad_scale = hyp[MPNN_HYP_GSCALE];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        int lo = 0, hi = n_seg;                 // seg_start[lo] <= e < seg_start[hi]
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (seg_start[mid] <= e) lo = mid; else hi = mid;
        }
        const int s = lo;
        float coef = 1.f, scale = seg_mult[s];
        if (node_stats) {
            const int nd = seg_node[s];
            // the moments ride in the all-reduced gradient tail: same 1/world scaling
            coef = node_stats[nd * 2 + 1] * grad_scale;
            if (talr) scale *= 1.0f / sqrtf(node_stats[nd * 2] * grad_scale);
        }
        float th = theta[e];
        float g = grad[e] * grad_scale + 2.f * seg_l2[s] * coef * th;
        g *= scale;
        float a = mu * accum[e] + g;
        accum[e] = a;
        theta[e] = th - lr * a;
    }
}

extern "C" int mpnn_talr_momentum_step(float* theta, const float* grad, float* accum, int n,
                                       const int* seg_start, const int* seg_node, const float* seg_mult,
                                       const float* seg_l2, int n_seg, const float* node_stats, int talr,
                                       const float* hyp, void* stream) {
    MPNN_REQUIRE(n > 0 && n_seg > 0, "talr_momentum_step: n=%d n_seg=%d", n, n_seg);
    int grid = ceil_div(n, 256);
    if (grid > 148 * 8) grid = 148 * 8;
    talr_momentum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        theta, grad, accum, n, seg_start, seg_node, seg_mult, seg_l2, n_seg, node_stats, talr, hyp);
    return mpnn_check_launch("talr_momentum_step");
}
