// Data-parallel step tail as ONE kernel over NVLink peer memory: gradient reduce-scatter, TALR + momentum on
// the owned slice, all-gather of the new parameters -- instead of ncclAllReduce followed by the optimiser.
//
// Every rank keeps [flags | grad | theta | accum] in one cudaMalloc'ed exchange buffer that all other ranks of
// the node map through CUDA IPC (mpnn_p2p_export / mpnn_p2p_import).  The kernel on rank r
//   1. tells every peer "my gradients are complete" (release store of the step's epoch into the peer's flag
//      word) and waits for the same word from every peer;
//   2. sums the per-node TALR moments of all ranks (a few dozen floats, every rank does it: fixed rank order,
//      so every replica scales its learning rates identically);
//   3. for its slice of the parameter vector: loads the gradient from ALL ranks (peer loads over NVLink, summed
//      in rank order), applies minimize_expectation + momentum (the arithmetic of csrc/optim.cu,
//      lib/net_types.py:24-37,96-97,178-181) and stores the new theta / momentum (and the reduced gradient)
//      into EVERY rank's buffers (peer stores) -- replicas are bit-identical by construction, since each
//      element is computed exactly once;
//   4. the last CTA tells every peer "my slice is written" and waits for the same from every peer, so the
//      kernel ends only when the local theta is complete and nobody reads the local gradients any more.
// Per rank and step: (N-1)/N of the vector read and 3 (N-1)/N written over NVLink, two flag round trips, no
// separate optimiser launch.  Waits are bounded (a peer that never arrives trips a status word instead of
// hanging the GPU).
#include <cstring>
#include "common.cuh"
#include "../../include/mpnn.h"

namespace {

// flag words (32-bit) at the head of an exchange buffer: one set per channel, so that two calls of a step (the
// deep-stage bucket under the backward pass, the rest at its end) never share a word
enum { F_READY = 0, F_DONE = MPNN_P2P_MAX, F_EPOCH = 2 * MPNN_P2P_MAX, F_TICKET, F_CHANNEL = 64, F_STATUS = 512 };
constexpr int kMaxStats = 1024;
constexpr unsigned long long kSpinLimitNs = 8000000000ull;       // 8 s

struct P2PArgs {
    mpnn_p2p_desc d;
    const int* seg_start; const int* seg_node; const float* seg_mult; const float* seg_l2;
    int n_seg, use_stats, talr, write_back;
    const float* hyp;
    long long lo4, hi4;     // range of the call in units of four parameters
    int channel;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_sys_f4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_sys_f(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// wait until *flag has reached epoch e (wrap-safe); gives up after kSpinLimitNs and records it
__device__ __forceinline__ void spin_until(const unsigned* flag, unsigned e, unsigned* status) {
    const unsigned long long t0 = globaltimer_ns();
    while ((int)(ld_acquire_sys(flag) - e) < 0) {
        if (globaltimer_ns() - t0 > kSpinLimitNs) { atomicExch(status, 1u); break; }
        __nanosleep(40);
    }
}

__global__ void __launch_bounds__(256)
allreduce_talr_p2p_kernel(const P2PArgs a) {
    __shared__ unsigned s_epoch;
    __shared__ int s_last;
    __shared__ float s_stats[kMaxStats];
    const int W = a.d.world, me = a.d.rank;
    unsigned* fl = reinterpret_cast<unsigned*>(a.d.base[me]) + a.channel * F_CHANNEL;
    unsigned* status = reinterpret_cast<unsigned*>(a.d.base[me]) + F_STATUS;
    if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned*>(fl + F_EPOCH) + 1u;
    __syncthreads();
    const unsigned e = s_epoch;
    // 1. my gradients are complete (they were written by earlier kernels of this stream) -> every peer
    if (blockIdx.x == 0 && threadIdx.x < W) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned*>(a.d.base[threadIdx.x]) + a.channel * F_CHANNEL + F_READY + me, e);
    }
    if (threadIdx.x < W) spin_until(fl + F_READY + threadIdx.x, e, status);
    __syncthreads();
    const float* gp[MPNN_P2P_MAX];
#pragma unroll
    for (int p = 0; p < MPNN_P2P_MAX; ++p)
        gp[p] = p < W ? reinterpret_cast<const float*>(static_cast<const char*>(a.d.base[p]) + a.d.off_grad) : nullptr;
    // 2. per-node moments summed over the ranks, in rank order
    if (a.use_stats) {
        for (int j = threadIdx.x; j < a.d.g0; j += blockDim.x) {
            float s = 0.f;
#pragma unroll
            for (int p = 0; p < MPNN_P2P_MAX; ++p)
                if (p < W) s += ld_sys_f(gp[p] + j);
            s_stats[j] = s;
        }
        __syncthreads();
    }
    // 3. my slice, four parameters per thread and iteration
    const float lr = a.hyp[MPNN_HYP_LR], mu = a.hyp[MPNN_HYP_MU], grad_scale = a.hyp[MPNN_HYP_GSCALE];
    const long long n4 = a.hi4 - a.lo4;
    const long long lo = a.lo4 + n4 * me / W, hi = a.lo4 + n4 * (me + 1) / W;
    float* th_l = reinterpret_cast<float*>(static_cast<char*>(a.d.base[me]) + a.d.off_theta);
    float* ac_l = reinterpret_cast<float*>(static_cast<char*>(a.d.base[me]) + a.d.off_accum);
    // (two independent items per thread and iteration: both sets of peer loads are in flight before the first use --
    //  the loop is a chain of NVLink round trips otherwise)
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < hi; i0 += 2 * stride) {
        const long long i1 = i0 + stride;
        const bool two = i1 < hi;
        const long long ea = i0 << 2, eb = (two ? i1 : i0) << 2;
        float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga;
#pragma unroll
        for (int p = 0; p < MPNN_P2P_MAX; ++p) {
            if (p < W) {
                const float4 va = ld_sys_f4(gp[p] + a.d.g0 + ea);
                const float4 vb = ld_sys_f4(gp[p] + a.d.g0 + eb);
                ga.x += va.x; ga.y += va.y; ga.z += va.z; ga.w += va.w;
                gb.x += vb.x; gb.y += vb.y; gb.z += vb.z; gb.w += vb.w;
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) break;
            const long long e0 = u ? eb : ea;
            const float4 g = u ? gb : ga;
            int slo = 0, shi = a.n_seg;             // seg_start[slo] <= e0 < seg_start[shi]; tensors start on multiples of 4
            while (shi - slo > 1) {
                const int mid = (slo + shi) >> 1;
                if (a.seg_start[mid] <= e0) slo = mid; else shi = mid;
            }
            float coef = 1.f, scale = a.seg_mult[slo];
            if (a.use_stats) {
                const int nd = a.seg_node[slo];
                coef = s_stats[nd * 2 + 1] * grad_scale;
                if (a.talr) scale *= 1.0f / sqrtf(s_stats[nd * 2] * grad_scale);
            }
            const float l2 = 2.f * a.seg_l2[slo] * coef;
            const float4 th = *reinterpret_cast<const float4*>(th_l + e0);
            const float4 ac = *reinterpret_cast<const float4*>(ac_l + e0);
            float4 an, tn;
            an.x = mu * ac.x + (g.x * grad_scale + l2 * th.x) * scale;  tn.x = th.x - lr * an.x;
            an.y = mu * ac.y + (g.y * grad_scale + l2 * th.y) * scale;  tn.y = th.y - lr * an.y;
            an.z = mu * ac.z + (g.z * grad_scale + l2 * th.z) * scale;  tn.z = th.z - lr * an.z;
            an.w = mu * ac.w + (g.w * grad_scale + l2 * th.w) * scale;  tn.w = th.w - lr * an.w;
#pragma unroll
            for (int p = 0; p < MPNN_P2P_MAX; ++p) {
                if (p < W) {
                    char* b = static_cast<char*>(a.d.base[p]);
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(b + a.d.off_theta) + e0) = tn;
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(b + a.d.off_accum) + e0) = an;
                    if (a.write_back) *reinterpret_cast<float4*>(reinterpret_cast<float*>(b + a.d.off_grad) + a.d.g0 + e0) = g;
                }
            }
        }
    }
    // 4. my slice is written everywhere -> every peer; leave when theirs is
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(fl + F_TICKET, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if (threadIdx.x < W) {
        st_release_sys(reinterpret_cast<unsigned*>(a.d.base[threadIdx.x]) + a.channel * F_CHANNEL + F_DONE + me, e);
        spin_until(fl + F_DONE + threadIdx.x, e, status);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        fl[F_TICKET] = 0u;
        *reinterpret_cast<volatile unsigned*>(fl + F_EPOCH) = e;
        __threadfence();
    }
}

}  // namespace

extern "C" int mpnn_p2p_alloc(void** ptr, long long bytes) {
    MPNN_REQUIRE(ptr && bytes > 0, "p2p_alloc: bytes=%lld", bytes);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { mpnn_set_error("p2p_alloc: %s", cudaGetErrorString(e)); return MPNN_ERR_CUDA; }
    *ptr = p;
    return MPNN_OK;
}

extern "C" int mpnn_p2p_free(void* ptr) {
    cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) { mpnn_set_error("p2p_free: %s", cudaGetErrorString(e)); return MPNN_ERR_CUDA; }
    return MPNN_OK;
}

extern "C" int mpnn_p2p_export(void* ptr, void* handle64) {
    MPNN_REQUIRE(ptr && handle64, "p2p_export: args");
    static_assert(sizeof(cudaIpcMemHandle_t) == MPNN_P2P_HANDLE_BYTES, "IPC handle size");
    cudaError_t e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), ptr);
    if (e != cudaSuccess) { mpnn_set_error("p2p_export (cudaIpcGetMemHandle): %s", cudaGetErrorString(e)); return MPNN_ERR_CUDA; }
    return MPNN_OK;
}

extern "C" int mpnn_p2p_import(const void* handle64, void** ptr) {
    MPNN_REQUIRE(ptr && handle64, "p2p_import: args");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { mpnn_set_error("p2p_import (cudaIpcOpenMemHandle): %s", cudaGetErrorString(e)); return MPNN_ERR_CUDA; }
    *ptr = p;
    return MPNN_OK;
}

extern "C" int mpnn_p2p_close(void* ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) { mpnn_set_error("p2p_close: %s", cudaGetErrorString(e)); return MPNN_ERR_CUDA; }
    return MPNN_OK;
}

extern "C" int mpnn_p2p_status(const void* local_base, int* status) {
    MPNN_REQUIRE(local_base && status, "p2p_status: args");
    unsigned s = 0;
    cudaError_t e = cudaMemcpy(&s, static_cast<const unsigned*>(local_base) + F_STATUS, sizeof(s), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { mpnn_set_error("p2p_status: %s", cudaGetErrorString(e)); return MPNN_ERR_CUDA; }
    *status = (int)s;
    return MPNN_OK;
}

extern "C" int mpnn_allreduce_talr_p2p(const mpnn_p2p_desc* d, const int* seg_start, const int* seg_node,
                                       const float* seg_mult, const float* seg_l2, int n_seg, int use_stats,
                                       int talr, const float* hyp, int write_back, int lo, int hi, int channel,
                                       void* stream) {
    MPNN_REQUIRE(d && d->world >= 2 && d->world <= MPNN_P2P_MAX && d->rank >= 0 && d->rank < d->world,
                 "allreduce_talr_p2p: world / rank");
    MPNN_REQUIRE(d->n > 0 && n_seg > 0 && d->g0 % 4 == 0 && d->g0 <= kMaxStats, "allreduce_talr_p2p: n=%d g0=%d", d->n, d->g0);
    MPNN_REQUIRE(d->off_grad >= MPNN_P2P_FLAG_BYTES && d->off_grad % 16 == 0 && d->off_theta % 16 == 0 && d->off_accum % 16 == 0,
                 "allreduce_talr_p2p: the buffers start behind the flag block, 16-byte aligned");
    for (int p = 0; p < d->world; ++p) MPNN_REQUIRE(d->base[p], "allreduce_talr_p2p: base[%d] is NULL", p);
    if (hi <= 0) hi = (d->n + 3) / 4 * 4;
    MPNN_REQUIRE(lo >= 0 && lo < hi && lo % 4 == 0 && hi % 4 == 0 && hi <= (d->n + 3) / 4 * 4 && channel >= 0 && channel < 4,
                 "allreduce_talr_p2p: range [%d, %d) of %d parameters, channel %d", lo, hi, d->n, channel);
    P2PArgs a;
    a.d = *d;
    a.seg_start = seg_start; a.seg_node = seg_node; a.seg_mult = seg_mult; a.seg_l2 = seg_l2;
    a.n_seg = n_seg; a.use_stats = use_stats; a.talr = talr; a.write_back = write_back; a.hyp = hyp;
    a.lo4 = lo / 4; a.hi4 = hi / 4; a.channel = channel;
    const long long n4 = a.hi4 - a.lo4, per_rank = ceil_div((int)n4, d->world);
    int grid = ceil_div((int)per_rank, 256);
    if (grid > 296) grid = 296;             // (no wait inside the kernel depends on another CTA of the grid being resident)
    if (grid < 1) grid = 1;
    allreduce_talr_p2p_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    return mpnn_check_launch("allreduce_talr_p2p");
}
