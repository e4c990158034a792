// tcgen05 weight gradient of the stencil GEMM (conv wgrad) for sm_100a.
//
//   dW[tap][k][n] += sum_p A[p + off(tap)][k] * G[p][n]
//
// The reduction runs over pixels p, which in the padded-planes layout are the
// ROWS of both operands: element (channel c, pixel p) sits at plane c/8, row p,
// i.e. exactly the UMMA MN-major interleaved core-matrix layout (8 pixels x 16 B
// contiguous per core matrix; LBO = 128 B between 8-pixel groups, SBO = plane
// stride between 8-channel groups).  So both operands are staged with the same
// one-bulk-copy-per-plane producer as the forward kernel and need no transpose:
//   D_tap[m = input channel][n = output channel] += A_tap^T (MN-major, shifted
//   start address per tap) x G (MN-major),  K = 16 pixels per tcgen05.mma.
// Input channels go on M (padded to 128 by reading whatever follows the staged
// planes -- rows of D are independent, rows >= K0+K1 are never read back), the
// output channels on N, so the tensor time per tap is proportional to C_out.
//
// A CTA owns a group of TG taps (all nine when 9*N TMEM columns fit, else one
// kernel row of three) and walks pixel chunks of 128 rows, accumulating in
// TMEM across the whole walk; partial sums from different CTAs are combined
// with fp32 red.global.add into the flat gradient buffer (which the optimiser
// kernel consumes).
#include "common.cuh"
#include "umma.cuh"
#include "../../include/mpnn.h"

namespace {

struct WgradArgs {
    const __nv_bfloat16* A0; const __nv_bfloat16* A1; const __nv_bfloat16* Gd;
    float* dW0; float* dW1;
    Geom g;
    int K0, K1, K0real, K1real, N, Nreal, ntaps, TG, n_chunks, nstage, rowsA, halo;
};

constexpr int kWThreads = 192;
constexpr int kChunk = 128;          // pixels per stage

__global__ void __launch_bounds__(kWThreads, 1)
stencil_wgrad_umma_kernel(const WgradArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KG = (a.K0 + a.K1) >> 3, KG0 = a.K0 >> 3, NG = a.N >> 3;
    const uint32_t PSA = (uint32_t)a.rowsA * 16;         // plane stride of the A stage
    const uint32_t PSG = (uint32_t)kChunk * 16;          // plane stride of the G stage
    const uint32_t stageA = PSA * KG, stageG = PSG * NG;
    uint8_t* sA = smem;
    uint8_t* sG = smem + (size_t)a.nstage * stageA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sG + (size_t)a.nstage * stageG);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * a.nstage, done = empty0 + 8 * a.nstage;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * a.nstage + 1);
    const int tap0 = blockIdx.y * a.TG;
    const uint32_t ncols = tmem_cols_pow2(a.TG * a.N);

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.nstage; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const bool has_work = (int)blockIdx.x < a.n_chunks;

    if (warp == 0) {
        if (lane == 0) {
            int it = 0;
            for (int c = blockIdx.x; c < a.n_chunks; c += gridDim.x, ++it) {
                const int s = it % a.nstage;
                const uint32_t ph = (uint32_t)(it / a.nstage) & 1u;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                mbar_expect_tx(full0 + 8 * s, stageA + stageG);
                const size_t p0 = (size_t)a.g.G + (size_t)c * kChunk;
                const uint32_t dA = smem_u32(sA) + (uint32_t)s * stageA;
                const uint32_t dG = smem_u32(sG) + (uint32_t)s * stageG;
                for (int kg = 0; kg < KG; ++kg) {
                    const __nv_bfloat16* src = kg < KG0 ? a.A0 + ((size_t)kg * a.g.P + p0 - a.halo) * 8
                                                        : a.A1 + ((size_t)(kg - KG0) * a.g.P + p0 - a.halo) * 8;
                    bulk_g2s(dA + (uint32_t)kg * PSA, src, PSA, full0 + 8 * s);
                }
                for (int ng = 0; ng < NG; ++ng)
                    bulk_g2s(dG + (uint32_t)ng * PSG, a.Gd + ((size_t)ng * a.g.P + p0) * 8, PSG, full0 + 8 * s);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0 && has_work) {
            const uint32_t idesc = make_idesc(a.N, 1, 1);        // both operands MN-major
            int it = 0;
            for (int c = blockIdx.x; c < a.n_chunks; c += gridDim.x, ++it) {
                const int s = it % a.nstage;
                const uint32_t ph = (uint32_t)(it / a.nstage) & 1u;
                mbar_wait(full0 + 8 * s, ph);
                tc_fence_after();
                const uint32_t aB = smem_u32(sA) + (uint32_t)s * stageA;
                const uint32_t gB = smem_u32(sG) + (uint32_t)s * stageG;
                for (int t = 0; t < a.TG; ++t) {
                    const int tap = tap0 + t;
                    const int off = a.ntaps == 9 ? (tap / 3 - 1) * a.g.Wp + (tap % 3 - 1) : 0;
                    const uint32_t arow = aB + (uint32_t)(a.halo + off) * 16;
                    const uint32_t dcol = tmem_base + (uint32_t)t * a.N;
                    for (int ks = 0; ks < kChunk / 16; ++ks) {
                        const uint64_t ad = make_desc(arow + (uint32_t)ks * 256, 128, PSA);
                        const uint64_t bd = make_desc(gB + (uint32_t)ks * 256, 128, PSG);
                        tc_mma(dcol, ad, bd, idesc, (it | ks) != 0);
                    }
                }
                tc_commit(empty0 + 8 * s);
            }
            tc_commit(done);
        }
        __syncwarp();
    } else if (has_work) {
        // epilogue: TMEM row m = input channel, columns = (tap, output channel)
        const int quad = warp & 3;
        const int m = quad * 32 + lane;
        mbar_wait(done, 0);
        tc_fence_after();
        const bool in0 = m < a.K0real;
        const bool in1 = m >= a.K0 && (m - a.K0) < a.K1real;
        for (int t = 0; t < a.TG; ++t) {
            const int tap = tap0 + t;
            float* dst = nullptr;
            if (in0) dst = a.dW0 + ((size_t)tap * a.K0real + m) * a.Nreal;
            else if (in1) dst = a.dW1 + ((size_t)tap * a.K1real + (m - a.K0)) * a.Nreal;
            for (int c = 0; c < a.N; c += 16) {
                float v[16];
                tc_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * a.N + c), v);
                if (dst) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c + i < a.Nreal) atomicAdd(dst + c + i, v[i]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// dbias[n] += sum over rows of G[p][n] (pad rows are zero)
__global__ void __launch_bounds__(256)
colsum_planes_kernel(const __nv_bfloat16* __restrict__ Gd, Geom g, int Nreal, float* __restrict__ dbias) {
    const int kg = blockIdx.y;
    float s[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < g.rows; q += gridDim.x * blockDim.x) {
        float v[8];
        Row8<__nv_bfloat16>::load(plane_row(Gd, kg, g.P, g.G + q), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += v[j];
    }
    __shared__ float red[8][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float t = warp_sum(s[j]);
        if (lane == 0) red[warp][j] = t;
    }
    __syncthreads();
    if (threadIdx.x < 8 && kg * 8 + threadIdx.x < Nreal) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        atomicAdd(dbias + kg * 8 + threadIdx.x, t);
    }
}

}  // namespace

int mpnn_stencil_wgrad_umma(const void* A0, int K0, int K0real, float* dW0, const void* A1, int K1,
                            int K1real, float* dW1, const void* Gd, int N, int Nreal, float* dbias,
                            int ntaps, Geom g, cudaStream_t st) {
    MPNN_REQUIRE(K0 % 16 == 0 && K1 % 16 == 0 && N % 16 == 0, "stencil_wgrad(tcgen05): K0=%d K1=%d N=%d", K0, K1, N);
    MPNN_REQUIRE(K0 + K1 <= 128 && N <= 256, "stencil_wgrad(tcgen05): K=%d > 128 or N=%d > 256", K0 + K1, N);
    const int KG = (K0 + K1) / 8, NG = N / 8;
    const int halo = ntaps == 9 ? g.Wp + 1 : 0;
    const int rowsA = kChunk + 2 * halo;
    int TG = ntaps;                                  // taps per CTA, bounded by 512 TMEM columns
    if (TG * N > 512) TG = ntaps == 9 ? 3 : 1;
    MPNN_REQUIRE(TG * N <= 512 && ntaps % TG == 0, "stencil_wgrad(tcgen05): N=%d too wide", N);
    const size_t stageA = (size_t)rowsA * 16 * KG, stageG = (size_t)kChunk * 16 * NG;
    const size_t kMax = 227 * 1024 - 1024;
    int nstage = (int)((kMax - 256) / (stageA + stageG));
    if (nstage > 4) nstage = 4;
    MPNN_REQUIRE(nstage >= 2, "stencil_wgrad(tcgen05): K=%d N=%d does not fit shared memory", K0 + K1, N);
    size_t smem = (size_t)nstage * (stageA + stageG) + 256;
    // the M=128 descriptor reads 16 planes from the start of an A stage: keep that inside the allocation
    const size_t reach = (size_t)(nstage - 1) * stageA + (size_t)16 * rowsA * 16 + 64;
    if (smem < reach) smem = reach;
    MPNN_REQUIRE(smem <= kMax + 1024, "stencil_wgrad(tcgen05): shared memory reach %zu", smem);
    int ncols = 32;
    while (ncols < TG * N) ncols <<= 1;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm > 512 / ncols) per_sm = 512 / ncols;
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    WgradArgs a;
    a.A0 = (const __nv_bfloat16*)A0; a.A1 = (const __nv_bfloat16*)A1; a.Gd = (const __nv_bfloat16*)Gd;
    a.dW0 = dW0; a.dW1 = dW1; a.g = g; a.K0 = K0; a.K1 = K1; a.K0real = K0real; a.K1real = K1real;
    a.N = N; a.Nreal = Nreal; a.ntaps = ntaps; a.TG = TG; a.n_chunks = ceil_div(g.rows, kChunk);
    a.nstage = nstage; a.rowsA = rowsA; a.halo = halo;
    const int groups = ntaps / TG;
    int gx = 148 * per_sm / groups;
    if (gx > a.n_chunks) gx = a.n_chunks;
    if (gx < 1) gx = 1;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(stencil_wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kMax + 1024);
        if (e != cudaSuccess) { mpnn_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return MPNN_ERR_CUDA; }
        attr_set = true;
    }
    stencil_wgrad_umma_kernel<<<dim3(gx, groups), kWThreads, smem, st>>>(a);
    int rc = mpnn_check_launch("stencil_wgrad_umma");
    if (rc) return rc;
    if (dbias) {
        int bx = ceil_div(g.rows, 256 * 8);
        if (bx > 148) bx = 148;
        if (bx < 1) bx = 1;
        colsum_planes_kernel<<<dim3(bx, NG), 256, 0, st>>>((const __nv_bfloat16*)Gd, g, Nreal, dbias);
        rc = mpnn_check_launch("colsum_planes");
    }
    return rc;
}
