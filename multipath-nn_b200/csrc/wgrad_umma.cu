// tcgen05 weight gradient of the stencil GEMM (conv wgrad) for sm_100a.
//
//   dW[tap][k][n] += sum_p A[p + off(tap)][k] * G[p][n]
//
// The reduction runs over pixels p, which in the padded-planes layout are the
// ROWS of both operands: element (channel c, pixel p) sits at plane c/8, row p,
// i.e. exactly the UMMA MN-major interleaved core-matrix layout (8 pixels x 16 B
// contiguous per core matrix; LBO = 128 B between 8-pixel groups, SBO = plane
// stride between 8-channel groups).  Both operands are staged with one bulk
// copy per plane and need no transpose:
//   D[m = input channel][n = output channel] += A^T (MN-major, start address
//   shifted by the tap's row offset) x G (MN-major),  K = 16 pixels per MMA.
//
// tcgen05.mma always fetches M >= 64 operand rows from shared memory, and for
// thin layers (16..32 input channels) that fetch -- not the tensor pipe, not HBM
// -- is the bottleneck (128 B/clk/SM).  Two measures keep the rows useful:
//   * M = 64 instead of 128 whenever the stacked rows fit in 64;
//   * "dw stacking" (CP = 3) for K0+K1 <= 40: each plane is staged three times,
//     shifted by dw = -1, 0, +1 rows, as consecutive 8-row groups of M, so ONE
//     MMA per kernel row dh covers three taps (rows m = (plane*3 + dw)*8 + c).
//   * for 16 output channels additionally "row stacking" along N (GS): the G descriptor's group
//     stride is set to one IMAGE ROW of pixels instead of a plane, so group j of a staged G plane is
//     that plane displaced by j image rows; with the dw-stacked A this makes ONE MMA per G plane
//     cover all nine taps (16 instead of 24 MMAs per 128 pixels, no extra staging).
// Without stacking (CP = 1) a CTA owns TG taps (9, or 3 when 9*N exceeds the
// 512 TMEM columns) and issues one MMA per tap.
//
// A CTA walks pixel chunks of 128 rows accumulating in TMEM across the whole
// walk; partial sums from different CTAs are combined with fp32
// red.global.add into the flat gradient buffer the optimiser kernel consumes.
#include <cstdlib>
#include "common.cuh"
#include "umma.cuh"
#include "../../include/mpnn.h"

namespace {

// destination of one (row segment, column segment) block of the accumulator
struct WDst { float* p; int ld, rows, cols; };

struct WgradArgs {
    const __nv_bfloat16* A0; const __nv_bfloat16* A1; const __nv_bfloat16* Gd;
    WDst dst[2][2];        // [row segment: k < K0 | k >= K0][column segment: n < Nsplit | n >= Nsplit]
    Geom g;
    int K0, K1, N, Nsplit, ntaps;
    int CP;        // staged copies per plane: 1, or 3 (dw stacking)
    int NM;        // MMAs per 16-pixel step per CTA (taps, or kernel rows when stacked)
    int M;         // 64 or 128
    int GS;        // with CP = 3: kernel ROWS stacked along N by giving the G descriptor a group stride of
                   // (W+1) pixel rows -- group j of one staged G plane is that plane displaced by j image
                   // rows -- so one MMA per G plane and K step covers all nine taps (GS = groups: 3, or 4
                   // when M = 128 needs N % 16 == 0; the 4th group is ignored)
    int n_chunks, nstage, rowsA, rowsG, halo;
    int vec4;      // gradient rows are 16-byte aligned: use vector reductions
};

constexpr int kWThreads = 192;
constexpr int kChunk = 128;          // pixels per stage

// NMT = MMAs per 16-pixel step (taps, or kernel rows when stacked: 9, 3 or 1), GST = row stacking along N
// (then NMT = 1 and N = 16).  Compile-time so that the issue loop is straight-line code with the tap
// offsets hoisted: with a run-time tap loop the single issuing warp spent ~100 cycles per MMA
// (ncu: tensor pipe 38 % active, one CTA per SM), which bounded the 32..128-channel layers.
template <int NMT, bool GST>
__global__ void __launch_bounds__(kWThreads, 1)
stencil_wgrad_umma_kernel(const WgradArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KGall = (a.K0 + a.K1) >> 3, KG0 = a.K0 >> 3, NG = a.N >> 3;
    const int kgb = blockIdx.z * 16;                     // first plane of this CTA's M block
    const int KG = min(16, KGall - kgb);                 // planes staged by this CTA
    const uint32_t PSA = (uint32_t)a.rowsA * 16;         // stride between 8-row groups of the A stage
    const uint32_t PSG = (uint32_t)a.rowsG * 16;         // plane stride of the G stage
    const uint32_t stageA = PSA * KG * a.CP, stageG = PSG * NG;
    uint8_t* sA = smem;
    uint8_t* sG = smem + (size_t)a.nstage * stageA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sG + (size_t)a.nstage * stageG);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * a.nstage, done = empty0 + 8 * a.nstage;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * a.nstage + 1);
    const int t0 = blockIdx.y * a.NM;                    // first tap (CP=1) / kernel row (CP=3) of this CTA
    const int NW = a.GS * 8;                             // GS: accumulator columns per G plane
    const uint32_t ncols = tmem_cols_pow2(a.GS ? NG * NW : a.NM * a.N);

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

    if (warp == 0) {
        {
            // issue loops run with the whole warp in uniform control flow, one elected lane issuing
            // (see elect_one in umma.cuh); running ring counters, hoisted bases
            const bool leader = elect_one();
            int s = 0; uint32_t ph = 0;
            const size_t plane = (size_t)a.g.P * 8;
            // planes of this CTA's M block: [kgb, kgb + KG) of the concatenation A0 | A1
            const int nA0 = max(0, min(KG, KG0 - kgb)), nA1 = KG - nA0;
            const __nv_bfloat16* base0 = a.A0 + (size_t)kgb * plane + ((size_t)a.g.G - a.halo) * 8;
            const __nv_bfloat16* base1 = a.A1 ? a.A1 + (size_t)max(0, kgb - KG0) * plane + ((size_t)a.g.G - a.halo) * 8 : nullptr;
            const __nv_bfloat16* baseG = a.Gd + ((size_t)a.g.G - (GST ? (size_t)a.g.Wp : 0)) * 8;
            const uint32_t cpy = a.CP == 3 ? 3 * PSA : PSA;
            for (int c = blockIdx.x; c < a.n_chunks; c += gridDim.x) {
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                if (leader) mbar_expect_tx(full0 + 8 * s, stageA + stageG);
                const size_t coff = (size_t)c * (kChunk * 8);
                uint32_t dA = smem_u32(sA) + (uint32_t)s * stageA;
                uint32_t dG = smem_u32(sG) + (uint32_t)s * stageG;
                const uint32_t fb = full0 + 8 * s;
                const __nv_bfloat16* pl = base0 + coff;
                for (int kg = 0; kg < nA0; ++kg, pl += plane, dA += cpy) {
                    if (a.CP == 3) {                  // copy cp holds the plane shifted by dw = cp-1 rows
                        if (leader) { bulk_g2s(dA, pl - 8, PSA, fb); bulk_g2s(dA + PSA, pl, PSA, fb); bulk_g2s(dA + 2 * PSA, pl + 8, PSA, fb); }
                    } else if (leader) bulk_g2s(dA, pl, PSA, fb);
                }
                pl = base1 + coff;
                for (int kg = 0; kg < nA1; ++kg, pl += plane, dA += cpy) {
                    if (a.CP == 3) {
                        if (leader) { bulk_g2s(dA, pl - 8, PSA, fb); bulk_g2s(dA + PSA, pl, PSA, fb); bulk_g2s(dA + 2 * PSA, pl + 8, PSA, fb); }
                    } else if (leader) bulk_g2s(dA, pl, PSA, fb);
                }
                // GS: the G stage starts one image row early (group j is read j image rows further down)
                const __nv_bfloat16* gp = baseG + coff;
                for (int ng = 0; ng < NG; ++ng, gp += plane, dG += PSG)
                    if (leader) bulk_g2s(dG, gp, PSG, fb);
                if (++s == a.nstage) { s = 0; ph ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        {
            const bool leader = elect_one();
            const uint32_t idesc = make_idesc(a.GS ? NW : a.N, 1, 1, a.M);   // both operands MN-major
            // descriptor lo word = (address >> 4) | (LBO = 128 B) << 16; hi = SBO (plane stride) | version
            const uint32_t a_hi = (PSA >> 4) | (1u << 14);
            const uint32_t g_hi = (a.GS ? (uint32_t)a.g.Wp : (PSG >> 4)) | (1u << 14);
            const uint32_t a_lo0 = (((smem_u32(sA) + (uint32_t)a.halo * 16) & 0x3FFFFu) >> 4) | (8u << 16);
            const uint32_t g_lo0 = ((smem_u32(sG) & 0x3FFFFu) >> 4) | (8u << 16);
            const uint32_t a_stage = stageA >> 4, g_stage = stageG >> 4;
            int s = 0; uint32_t ph = 0;
            uint32_t accf = 0;                        // 0 on the CTA's first chunk: the MMAs overwrite the accumulator
            // row shift (16 B units) of each MMA's A operand, hoisted out of the chunk loop
            uint32_t offs[NMT];
#pragma unroll
            for (int t = 0; t < NMT; ++t) {
                int off;
                if (a.CP == 3) off = (t0 + t - 1) * a.g.Wp;
                else off = a.ntaps == 9 ? ((t0 + t) / 3 - 1) * a.g.Wp + ((t0 + t) % 3 - 1) : 0;
                offs[t] = (uint32_t)off;
            }
            const uint32_t N_ = (uint32_t)a.N, gplane = PSG >> 4;
            for (int c = blockIdx.x; c < a.n_chunks; c += gridDim.x) {
                mbar_wait(full0 + 8 * s, ph);
                tc_fence_after();
                const uint32_t aB = a_lo0 + (uint32_t)s * a_stage;
                const uint32_t gB = g_lo0 + (uint32_t)s * g_stage;
                if (GST) {
                    // D[(plane, dw, c)][(G plane j, dh' , n)]: group dh' of G plane j = rows displaced by dh' image
                    // rows = kernel row dh = 1 - dh'   (N = 16: two G planes)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t gj = gB + (uint32_t)j * gplane;
                        const uint32_t dcol = tmem_base + (uint32_t)(j * NW);
#pragma unroll
                        for (int ks = 0; ks < kChunk / 16; ++ks) {
                            if (leader) tc_mma2(dcol, aB + ks * 16, a_hi, gj + ks * 16, g_hi, idesc, ks ? 1u : accf);
                        }
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < NMT; ++t) {
                        const uint32_t arow = aB + offs[t];
                        const uint32_t dcol = tmem_base + (uint32_t)t * N_;
#pragma unroll
                        for (int ks = 0; ks < kChunk / 16; ++ks) {   // 16 pixels = 256 B per K step
                            if (leader) tc_mma2(dcol, arow + ks * 16, a_hi, gB + ks * 16, g_hi, idesc, ks ? 1u : accf);
                        }
                    }
                }
                accf = 1;
                if (leader) tc_commit(empty0 + 8 * s);
                if (++s == a.nstage) { s = 0; ph ^= 1u; }
            }
            if (leader) tc_commit(done);
        }
        __syncwarp();
    } else {
        // epilogue: TMEM row m = stacked input-channel row, columns = (MMA index, output channel)
        const int quad = warp & 3;
        const int m = a.M == 64 ? quad * 16 + lane : quad * 32 + lane;
        const bool row_ok = a.M == 64 ? lane < 16 : true;
        mbar_wait(done, 0);
        tc_fence_after();
        const int grp = m >> 3, cc = m & 7;
        const int kg = a.CP == 3 ? grp / 3 : grp;
        const int dwi = a.CP == 3 ? grp % 3 : 0;
        const int k = (kgb + kg) * 8 + cc;               // input channel in the concatenated K space
        const int rs = k >= a.K0 ? 1 : 0;
        const int kr = rs ? k - a.K0 : k;
        const bool row_live = row_ok && kg < KG;
        if (a.GS) {
            for (int j = 0; j < NG; ++j) {
                const int nn = j * 8;
                const int cs = nn >= a.Nsplit ? 1 : 0;
                const WDst d = a.dst[rs][cs];
                const int n = cs ? nn - a.Nsplit : nn;
#pragma unroll
                for (int dh2 = 0; dh2 < 3; ++dh2) {
                    float v[8];
                    tc_ld8(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(j * NW + dh2 * 8), v);
                    const int tap = (2 - dh2) * 3 + dwi;
                    if (row_live && d.p && kr < d.rows) {
                        float* dst = d.p + ((size_t)tap * d.rows + kr) * d.ld + n;
                        if (a.vec4 && (d.ld & 3) == 0) {
#pragma unroll
                            for (int i = 0; i < 8; i += 4)
                                if (n + i < d.cols)
                                    atomicAdd(reinterpret_cast<float4*>(dst + i),
                                              make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (n + i < d.cols) atomicAdd(dst + i, v[i]);
                        }
                    }
                }
            }
        } else
        for (int t = 0; t < a.NM; ++t) {
            const int tap = a.CP == 3 ? (t0 + t) * 3 + dwi : t0 + t;
            for (int c = 0; c < a.N; c += 16) {
                float v[16];
                tc_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * a.N + c), v);
                const int cs = c >= a.Nsplit ? 1 : 0;
                const WDst d = a.dst[rs][cs];
                const int n = cs ? c - a.Nsplit : c;
                if (row_live && d.p && kr < d.rows) {
                    float* dst = d.p + ((size_t)tap * d.rows + kr) * d.ld + n;
                    if (a.vec4 && (d.ld & 3) == 0) {
                        // one 16-byte red.global.add.v4.f32 per 4 columns (sm_90+)
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            if (n + i < d.cols)
                                atomicAdd(reinterpret_cast<float4*>(dst + i),
                                          make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (n + i < d.cols) atomicAdd(dst + i, v[i]);
                    }
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

// shared launcher: fills the tiling fields of `a` (operands / destinations / K0,K1,N,Nsplit,ntaps,g set by caller)
static int launch_wgrad(WgradArgs& a, const float* dbias_src_unused, cudaStream_t st) {
    const int KGall = (a.K0 + a.K1) / 8, NG = a.N / 8;
    const int mblocks = ceil_div(KGall, 16);
    const int KG = KGall < 16 ? KGall : 16;          // planes per CTA (largest block)
    a.CP = (a.ntaps == 9 && mblocks == 1 && 3 * KG <= 16) ? 3 : 1;
    a.M = KG * a.CP <= 8 ? 64 : 128;
    const int n_units = a.CP == 3 ? 3 : a.ntaps;     // MMAs per 16-pixel step over all CTAs of a chunk
    a.NM = n_units;
    if (a.NM * a.N > 512) a.NM = a.ntaps == 9 ? 3 : 1;
    MPNN_REQUIRE(a.NM * a.N <= 512 && n_units % a.NM == 0, "wgrad(tcgen05): N=%d too wide", a.N);
    a.n_chunks = ceil_div(a.g.rows, kChunk);
    a.GS = 0;
    if (a.CP == 3 && NG <= 2) {      // one MMA per G plane instead of one per kernel row: fewer only for N = 16
        const int gs = a.M == 64 ? 3 : 4;
        // the G stage of the last chunk ends (gs - 2) image rows past the chunk
        if (a.g.G >= a.g.Wp && (long long)a.g.G + 128ll * a.n_chunks + (long long)(gs - 2) * a.g.Wp <= a.g.P &&
            NG * gs * 8 <= 512)
            a.GS = gs;
    }
    a.halo = a.ntaps == 9 ? (a.GS ? 0 : (a.CP == 3 ? a.g.Wp : a.g.Wp + 1)) : 0;
    a.rowsA = kChunk + 2 * a.halo;
    a.rowsG = kChunk + (a.GS ? (a.GS - 1) * a.g.Wp : 0);
    if (a.GS) a.NM = 1;
    const size_t stageA = (size_t)a.rowsA * 16 * KG * a.CP, stageG = (size_t)a.rowsG * 16 * NG;
    const size_t kMax = 227 * 1024 - 1024;
    int nstage = (int)((kMax - 256) / (stageA + stageG));
    if (nstage > 4) nstage = 4;
    MPNN_REQUIRE(nstage >= 2, "wgrad(tcgen05): K=%d N=%d does not fit shared memory", a.K0 + a.K1, a.N);
    int ncols = 32;
    while (ncols < (a.GS ? NG * a.GS * 8 : a.NM * a.N)) ncols <<= 1;
    // residency before ring depth: the single issuing warp of a CTA is latency-bound, a second (third)
    // resident CTA interleaves its MMAs; take the deepest ring that still gives the most CTAs per SM
    // (bounded by the TMEM columns: a blocked tcgen05.alloc would stall the extra CTA)
    static const int tune_wps = getenv("MPNN_TUNE_WGRAD_PER_SM") ? atoi(getenv("MPNN_TUNE_WGRAD_PER_SM")) : 0;
    int cap = 512 / ncols;
    if (cap > 3) cap = 3;
    if (tune_wps && cap > tune_wps) cap = tune_wps;
    if (cap < 1) cap = 1;
    size_t smem = 0;
    int per_sm = 0;
    for (int ns = nstage; ns >= 2; --ns) {
        size_t sm = (size_t)ns * (stageA + stageG) + 256;
        // the descriptor reads M/8 groups from the start of an A stage: keep that inside the allocation
        const size_t reach = (size_t)(ns - 1) * stageA + (size_t)(a.M / 8) * a.rowsA * 16 + 64;
        if (sm < reach) sm = reach;
        int fit = (int)((227 * 1024) / (sm + 1024));
        if (fit > cap) fit = cap;
        if (fit < 1) fit = 1;
        if (fit > per_sm) { per_sm = fit; smem = sm; nstage = ns; }
    }
    MPNN_REQUIRE(smem <= kMax + 1024, "wgrad(tcgen05): shared memory reach %zu", smem);
    a.nstage = nstage;
    a.vec4 = 1;
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 2; ++c)
            if (a.dst[r][c].p && ((uintptr_t)a.dst[r][c].p % 16 != 0)) a.vec4 = 0;
    const int groups = a.GS ? 1 : n_units / a.NM;
    int gx = 148 * per_sm / (groups * mblocks);
    // every CTA ends with a full-size reduction of its accumulators into dW: give each
    // at least 8 chunks of pixels so that the reduction traffic stays small next to the MMAs
    if (gx > ceil_div(a.n_chunks, 8)) gx = ceil_div(a.n_chunks, 8);
    if (gx < 1) gx = 1;
    typedef void (*Kern)(const WgradArgs);
    static const Kern kerns[4] = {stencil_wgrad_umma_kernel<9, false>, stencil_wgrad_umma_kernel<3, false>,
                                  stencil_wgrad_umma_kernel<1, false>, stencil_wgrad_umma_kernel<1, true>};
    MPNN_REQUIRE(a.GS ? NG == 2 : (a.NM == 9 || a.NM == 3 || a.NM == 1), "wgrad(tcgen05): NM=%d GS=%d NG=%d", a.NM, a.GS, NG);
    const Kern kern = a.GS ? kerns[3] : (a.NM == 9 ? kerns[0] : (a.NM == 3 ? kerns[1] : kerns[2]));
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {          // per-device function attribute
        for (int i = 0; i < 4; ++i) {
            cudaError_t e = cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMax + 1024);
            if (e != cudaSuccess) { mpnn_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return MPNN_ERR_CUDA; }
            if (!getenv("MPNN_TUNE_NO_CARVEOUT")) cudaFuncSetAttribute(kerns[i], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        }
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    kern<<<dim3(gx, groups, mblocks), kWThreads, smem, st>>>(a);
    return mpnn_check_launch("stencil_wgrad_umma");
}

int mpnn_stencil_wgrad_umma(const void* A0, int K0, int K0real, float* dW0, const void* A1, int K1,
                            int K1real, float* dW1, const void* Gd, int N, int Nreal, float* dbias,
                            int ntaps, Geom g, cudaStream_t st) {
    MPNN_REQUIRE(K0 % 16 == 0 && K1 % 16 == 0 && N % 16 == 0, "stencil_wgrad(tcgen05): K0=%d K1=%d N=%d", K0, K1, N);
    MPNN_REQUIRE(N <= 256, "stencil_wgrad(tcgen05): N=%d > 256", N);
    WgradArgs a;
    a.A0 = (const __nv_bfloat16*)A0; a.A1 = (const __nv_bfloat16*)A1; a.Gd = (const __nv_bfloat16*)Gd;
    a.g = g; a.K0 = K0; a.K1 = K1; a.N = N; a.Nsplit = N; a.ntaps = ntaps;
    a.dst[0][0] = WDst{dW0, Nreal, K0real, Nreal};
    a.dst[1][0] = WDst{dW1, Nreal, K1real, Nreal};
    a.dst[0][1] = WDst{nullptr, 0, 0, 0};
    a.dst[1][1] = WDst{nullptr, 0, 0, 0};
    int rc = launch_wgrad(a, nullptr, st);
    if (rc) return rc;
    if (dbias) {
        int bx = ceil_div(g.rows, 256 * 8);
        if (bx > 148) bx = 148;
        if (bx < 1) bx = 1;
        colsum_planes_kernel<<<dim3(bx, N / 8), 256, 0, st>>>((const __nv_bfloat16*)Gd, g, Nreal, dbias);
        rc = mpnn_check_launch("colsum_planes");
    }
    return rc;
}

// Weight gradient of the fully-connected heads sharing one feature matrix X:
//   dWa[f][j] += sum_b X[f][b] * dZ[b][j]            (j <  na, f < Fa)   -- LogReg
//   dWb[f][j] += sum_b X[f][b] * dZ[b][Nsplit + j]   (j <  nb, f < Fb)   -- first router FC
// X: feature planes [F/8][Balloc][8] bf16, dZ: planes [N/8][Balloc][8] bf16.
extern "C" int mpnn_fc_wgrad(const void* X, int F, int Balloc, int B, const void* dZ, int N, int Nsplit,
                             float* dWa, int Fa, int na, float* dWb, int Fb, int nb, void* stream) {
    MPNN_REQUIRE(F % 16 == 0 && N % 16 == 0 && Nsplit % 16 == 0 && N <= 256, "fc_wgrad: F=%d N=%d Nsplit=%d", F, N, Nsplit);
    MPNN_REQUIRE(Balloc >= ceil_div(B, kChunk) * kChunk, "fc_wgrad: Balloc=%d must cover whole 128-row chunks", Balloc);
    WgradArgs a;
    a.A0 = (const __nv_bfloat16*)X; a.A1 = nullptr; a.Gd = (const __nv_bfloat16*)dZ;
    a.g = make_geom(B, 0, 0, 0, Balloc);
    a.K0 = F; a.K1 = 0; a.N = N; a.Nsplit = Nsplit; a.ntaps = 1;
    a.dst[0][0] = WDst{dWa, na, Fa, na};
    a.dst[0][1] = WDst{dWb, nb, Fb, nb};
    a.dst[1][0] = WDst{nullptr, 0, 0, 0};
    a.dst[1][1] = WDst{nullptr, 0, 0, 0};
    return launch_wgrad(a, nullptr, (cudaStream_t)stream);
}
