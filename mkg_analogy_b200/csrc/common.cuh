// Shared device/host helpers for the b200mart kernels (sm_100a only).
// PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM) and
// small numeric helpers. Nothing here is derived from the reference (it has no native code).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- error plumbing (host)
extern thread_local char g_b200mart_err[512];
int b200mart_set_err(int code, const char* fmt, ...);

#define CUDA_TRY(expr)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return b200mart_set_err(3, "%s:%d CUDA error %s: %s", __FILE__, __LINE__,      \
                                    cudaGetErrorName(_e), cudaGetErrorString(_e));         \
    } while (0)

// every kernel launch site: check the launch and count it (b200mart_launch_count, for bench.py's gpu_launches)
extern unsigned long long g_b200mart_launches;
#define B2_LAUNCHED()                                                                      \
    do {                                                                                   \
        CUDA_TRY(cudaGetLastError());                                                      \
        ++g_b200mart_launches;                                                             \
    } while (0)

#define B2_CHECK(cond, ...)                                                                \
    do {                                                                                   \
        if (!(cond)) return b200mart_set_err(1, __VA_ARGS__);                              \
    } while (0)

#define B2_TRY(expr)                                                                       \
    do {                                                                                   \
        int _r = (expr);                                                                   \
        if (_r != 0) return _r;                                                            \
    } while (0)

// ---------------------------------------------------------------- dropout (counter-based, stateless)
// keep(element) = hash32(index ^ site_seed) >= thresh, thresh = p * 2^32; kept values are scaled by 1 / (1 - p).  The mask
// of a site is a pure function of (seed, site, element index), so backward regenerates it instead of storing it.
// Sites of the BERT stream (MarT/models/modeling_unimo.py:185, :362, :389, :476): 0 = embeddings; per text layer i:
// 1 + 3 i = attention probabilities, 2 + 3 i = attention-output dense, 3 + 3 i = FFN-output dense.
struct DropCfg {
    uint32_t site_seed;
    uint32_t thresh;   // 0 = dropout off
    float scale;
    // optional per-step salt in DEVICE memory, XORed into site_seed by the kernels: lets a captured CUDA graph (whose kernel
    // arguments are frozen) draw fresh masks at every replay (b200mart_set_step_block)
    const uint32_t* salt = nullptr;
};
__host__ __device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
inline DropCfg make_drop(float p, uint32_t seed, int site, const uint32_t* salt = nullptr) {
    DropCfg d;
    d.salt = (p > 0.f) ? salt : nullptr;
    d.site_seed = hash32(seed + 0x9E3779B9U * (uint32_t)(site + 1));
    d.thresh = (p > 0.f) ? (uint32_t)((double)p * 4294967296.0) : 0u;
    d.scale = (p > 0.f) ? 1.0f / (1.0f - p) : 1.0f;
    return d;
}

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// Every kernel of the two dependent chains is launched with cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs
// may become resident (and run their prologue: barrier / tensor-memory set-up, index arithmetic) while the previous
// kernel of the stream is still draining; pdl_wait() (griddepcontrol.wait) blocks until that kernel has completed and
// its writes are visible, and MUST precede the kernel's first access to global memory.  pdl_launch_dependents() lets
// the next kernel's CTAs be scheduled once every CTA of this grid has called it (or exited).  B200MART_PDL=0 turns the
// attribute off (A/B runs); both instructions are no-ops for a kernel launched without it.
int b200mart_pdl_enabled();
#ifdef __CUDACC__
#include <utility>
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = b200mart_pdl_enabled();
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
// multiplier of element `idx` at a dropout site: 0 (dropped) or 1 / (1 - p)
__device__ __forceinline__ uint32_t drop_seed(const DropCfg& d) {      // site seed of this launch (salted when a step block is set)
    return d.salt ? (d.site_seed ^ __ldg(d.salt)) : d.site_seed;
}
__device__ __forceinline__ float drop_mult(const DropCfg& d, uint32_t idx) {
    return (hash32(idx ^ drop_seed(d)) >= d.thresh) ? d.scale : 0.f;
}


// Spin limit for every mbarrier wait: a protocol bug must trap (CUDA error on the host)
// instead of hanging the GPU box.
#ifndef B2_SPIN_LIMIT
#define B2_SPIN_LIMIT (1u << 26)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > B2_SPIN_LIMIT) {
            printf("b200mart: mbarrier wait timed out (block %d,%d thread %d bar %u parity %u)\n",
                   blockIdx.x, blockIdx.y, threadIdx.x, bar, parity);
            __trap();
        }
    }
}

// 2D TMA tile load: coordinates are (c0 = innermost element index, c1 = row index).
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
        "%4}], [%2];" ::"r"(smem_dst),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// ---- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
// Warp-collective TMEM -> registers: 32 lanes x 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
          "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
          "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
          "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- CTA-pair (cta_group::2) variants: two CTAs of a 2-CTA cluster on one TPC drive ONE 256-row MMA; each CTA holds its
// own 128 rows of A, HALF of the B tile and its 128 accumulator lanes.  PTX forms as in CUTLASS (cute/arch/copy_sm100_tma.hpp,
// cutlass/arch/barrier.h): shared::cta addresses of the same offset in the two CTAs differ in bit 24 ("peer bit"); clearing it
// addresses the leader (rank 0) CTA's copy.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA tile load issued by either CTA of the pair into ITS OWN shared memory; the transaction bytes are signalled on the
// LEADER CTA's mbarrier (peer bit cleared).
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1)
        : "memory");
}
// D[tmem of both CTAs] (+)= A * B over the CTA pair (M = 256), issued by ONE thread of the leader CTA.
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier at this shared-memory offset in BOTH CTAs once the previously issued pair-MMAs have completed
__device__ __forceinline__ void umma_commit_2cta_mc(uint32_t bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
// plain arrive on the LEADER CTA's mbarrier from either CTA
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_BIT_MASK) : "memory");
}

// ---- shared-memory matrix descriptors for tcgen05.mma, 128-byte swizzle, bf16.
// Bit layout per the PTX "shared memory descriptor" (version 1 = Blackwell):
//   [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version, [61,64) swizzle (2 = 128B).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                    uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor, kind::f16, bf16 x bf16 -> fp32.
//   [4,6) D fmt (1 = f32), [7,10) A fmt (1 = bf16), [10,13) B fmt, bit15 A MN-major, bit16 B MN-major,
//   [17,23) N>>3, [24,29) M>>4.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- numerics
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
    float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
    float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}
__device__ __forceinline__ float quick_gelu(float x) { return x / (1.0f + __expf(-1.702f * x)); }
__device__ __forceinline__ float quick_gelu_grad(float x) {
    float s = 1.0f / (1.0f + __expf(-1.702f * x));
    return s + 1.702f * x * s * (1.0f - s);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
#endif  // __CUDACC__
