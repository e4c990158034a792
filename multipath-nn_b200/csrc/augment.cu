// Training-batch augmentation on the device (reference: scripts/lib/data.py:24-34).
//
//   out[i][r][c][:] = x[j_i][r + du_i][flip_i ? W-1-(c + dv_i) : c + dv_i][:]   inside the image,
//                     mean over (h, w) of image j_i                              outside,
//   y_out[i] = y[j_i].
// The random draws (sample index, flip, shifts) stay on the host in the reference's order, so a
// seeded run reproduces the NumPy batches; the per-example Python loop -- four orders of magnitude
// slower than the training step -- is what moves to the GPU.  One CTA per output image.
#include "common.cuh"
#include "../../include/mpnn.h"

namespace {

constexpr int kMaxC = 8;

__global__ void __launch_bounds__(256)
augment_batch_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W, int C, int n_cls,
                     const int* __restrict__ idx, const int* __restrict__ flip, const int* __restrict__ du,
                     const int* __restrict__ dv, float* __restrict__ x_out, float* __restrict__ y_out) {
    const int i = blockIdx.x, tid = threadIdx.x;
    const size_t img = (size_t)H * W * C;
    const float* src = x + (size_t)idx[i] * img;
    float* dst = x_out + (size_t)i * img;
    // per-channel mean in fp64 (data.py averages the float64 copy of the image)
    double s[kMaxC];
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) s[c] = 0.0;
    for (int p = tid; p < H * W; p += blockDim.x)
        for (int c = 0; c < C; ++c) s[c] += (double)src[(size_t)p * C + c];
    __shared__ double red[8][kMaxC];
    __shared__ float mean[kMaxC];
    const int warp = tid >> 5, lane = tid & 31;
    for (int c = 0; c < C; ++c) {
        double t = s[c];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) red[warp][c] = t;
    }
    __syncthreads();
    if (tid < C) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w][tid];
        mean[tid] = (float)(t / (double)(H * W));
    }
    __syncthreads();
    const int a = du[i], b = dv[i], f = flip[i];
    for (int e = tid; e < H * W * C; e += blockDim.x) {
        const int c = e % C, p = e / C, col = p % W, row = p / W;
        const int sr = row + a, sc = col + b;
        float v = mean[c];
        if (sr >= 0 && sr < H && sc >= 0 && sc < W) v = src[((size_t)sr * W + (f ? W - 1 - sc : sc)) * C + c];
        dst[e] = v;
    }
    for (int k = tid; k < n_cls; k += blockDim.x) y_out[(size_t)i * n_cls + k] = y[(size_t)idx[i] * n_cls + k];
}

}  // namespace

extern "C" int mpnn_augment_batch(const float* x, const float* y, int N, int H, int W, int C, int n_cls,
                                  const int* idx, const int* flip, const int* du, const int* dv, int B,
                                  float* x_out, float* y_out, void* stream) {
    MPNN_REQUIRE(N > 0 && B > 0 && H > 0 && W > 0 && C >= 1 && C <= kMaxC && n_cls >= 1,
                 "augment_batch: N=%d B=%d H=%d W=%d C=%d", N, B, H, W, C);
    MPNN_REQUIRE(x && y && idx && flip && du && dv && x_out && y_out, "augment_batch: null pointer");
    augment_batch_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, y, H, W, C, n_cls, idx, flip, du, dv, x_out, y_out);
    return mpnn_check_launch("augment_batch");
}
