// Shared device/host helpers for libmpnn_sm100 (sm_100a only).
//
// Activation layout ("padded planes"), used by every kernel in this library:
//   tensor[C/8][P_alloc][8]   element type T in {float, bf16}
// where a row p is one pixel of a zero-interleaved flat image stack:
//   p = G + n*(H+1)*(W+1) + (h+1)*(W+1) + (w+1)
// Row 0 and column 0 of every (H+1)x(W+1) image block are zero pads shared
// with the neighbouring row/image, and G front / >=192 back guard rows are
// zero, so a 3x3 SAME convolution is a 9-tap 1-D stencil over p with offsets
// dh*(W+1)+dw and every operand tile is a contiguous run of 16/32-byte rows.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define MPNN_OK 0
#define MPNN_ERR_ARG (-1)
#define MPNN_ERR_CUDA (-2)
#define MPNN_ERR_UNSUPPORTED (-3)

void mpnn_set_error(const char* fmt, ...);
int mpnn_check_launch(const char* what);

#define MPNN_REQUIRE(cond, ...)                         \
    do {                                                \
        if (!(cond)) {                                  \
            mpnn_set_error(__VA_ARGS__);                \
            return MPNN_ERR_ARG;                        \
        }                                               \
    } while (0)

enum { MPNN_F32 = 0, MPNN_BF16 = 1 };

struct Geom {
    int B, H, W, G;      // batch, image size, front guard rows
    int Wp, S;           // W+1, (H+1)*(W+1)
    int rows;            // B*S  (logical rows after the guard)
    int P;               // allocated rows per plane
    uint32_t mS, mWp;    // floor(2^32/S)+1, floor(2^32/Wp)+1: division by multiply-high
};

static inline Geom make_geom(int B, int H, int W, int G, int P) {
    Geom g;
    g.B = B; g.H = H; g.W = W; g.G = G;
    g.Wp = W + 1; g.S = (H + 1) * (W + 1);
    g.rows = B * g.S; g.P = P;
    g.mS = g.S > 1 ? (uint32_t)((1ull << 32) / (uint64_t)g.S) + 1u : 0xFFFFFFFFu;
    g.mWp = g.Wp > 1 ? (uint32_t)((1ull << 32) / (uint64_t)g.Wp) + 1u : 0xFFFFFFFFu;
    return g;
}

// row index (relative to the guard) -> validity; also yields n,h,w
__device__ __forceinline__ bool row_valid(const Geom& g, int q, int& n, int& h, int& w) {
    if (q < 0 || q >= g.rows) return false;
    // __umulhi(q, m) with m = floor(2^32/d)+1 is q/d or q/d+1 for any q < 2^32: one fix-up, no division
    n = (int)__umulhi((uint32_t)q, g.mS);
    int r = q - n * g.S;
    if (r < 0) { r += g.S; --n; } else if (r >= g.S) { r -= g.S; ++n; }   // second arm: degenerate S == 1 only
    int hr = (int)__umulhi((uint32_t)r, g.mWp);
    int wc = r - hr * g.Wp;
    if (wc < 0) { wc += g.Wp; --hr; } else if (wc >= g.Wp) { wc -= g.Wp; ++hr; }
    h = hr - 1; w = wc - 1;
    return hr >= 1 && wc >= 1;
}
__device__ __forceinline__ int row_of(const Geom& g, int n, int h, int w) {
    return g.G + n * g.S + (h + 1) * g.Wp + (w + 1);
}

// ---- 8-channel row load/store ------------------------------------------- //
template <typename T> struct Row8;
template <> struct Row8<float> {
    static __device__ __forceinline__ void load(const float* p, float v[8]) {
        float4 a = __ldg(reinterpret_cast<const float4*>(p));
        float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float v[8]) {
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
};
template <> struct Row8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float v[8]) {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __bfloat1622float2(h[i]);
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float v[8]) {
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = u;
    }
};

template <typename T>
__device__ __forceinline__ T* plane_row(T* base, int kg, int P, int p) {
    return base + ((size_t)kg * P + p) * 8;
}
template <typename T>
__device__ __forceinline__ const T* plane_row(const T* base, int kg, int P, int p) {
    return base + ((size_t)kg * P + p) * 8;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch (PDL) ---------------------------------- //
// The conv -> BN -> conv chain of one pyramid scale is a sequence of short dependent launches on
// one stream.  Kernels of that chain are launched with the programmatic-stream-serialization
// attribute: the next kernel's CTAs are scheduled (and run their prologue: barrier init, TMEM
// allocation, shared-memory setup) while the previous kernel drains, and block in
// griddepcontrol.wait -- which returns only when the previous grid has COMPLETED and its memory is
// visible -- before touching global memory.  MPNN_PDL=0 disables the attribute.
#include <cstdlib>
static inline bool mpnn_pdl_enabled() {
    static const int v = getenv("MPNN_PDL") ? atoi(getenv("MPNN_PDL")) : 1;
    return v != 0;
}
// Only grids of more than one CTA per SM take the attribute (MPNN_PDL_MIN_CTAS=n overrides): the early-scheduled
// CTAs of a small dependent grid sit on shared memory / TMEM that the kernels of the other lanes could be using
// (measured, B200, cifar10-ac: B = 128 0.637 -> 0.592 ms / step, B = 256 0.708 -> 0.676, B >= 1024 unchanged,
// B = 512 +2 %).
static inline unsigned mpnn_pdl_min_ctas() {
    static const int v = getenv("MPNN_PDL_MIN_CTAS") ? atoi(getenv("MPNN_PDL_MIN_CTAS")) : 149;
    return (unsigned)v;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t mpnn_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                          cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = (mpnn_pdl_enabled() && grid.x * grid.y * grid.z >= mpnn_pdl_min_ctas()) ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
// device side: let the dependent grid start scheduling, then wait for the grid we depend on
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

#define MPNN_DISPATCH_DTYPE(dtype, ...)                                  \
    do {                                                                 \
        if ((dtype) == MPNN_F32) { typedef float T; __VA_ARGS__; }       \
        else if ((dtype) == MPNN_BF16) { typedef __nv_bfloat16 T; __VA_ARGS__; } \
        else { mpnn_set_error("bad dtype %d", (int)(dtype)); return MPNN_ERR_ARG; } \
    } while (0)
