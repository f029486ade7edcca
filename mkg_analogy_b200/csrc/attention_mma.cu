// Tensor-core fused attention (head_dim 64) for the MKGformer step, forward + two-pass backward.
//   text   — BertSelfAttention.forward, MarT/models/modeling_unimo.py:317-377 (adaptive block scaling :342-349
//            applied BEFORE the additive -1e4 padding mask :351-355)
//   vision — CLIPAttention.forward :212-272 with the text K/V prefix (:227-229, unmasked)
// One CTA (4 warps) per (head, sample).  Scores never leave the SM: each warp owns 16 query rows (forward, rows pass)
// or 16 keys (keys pass) and streams the other side from shared memory in 64-wide chunks with an online softmax
// (forward) or the saved log-sum-exp (backward).  The per-(sample, head) tiles are tiny (<= 99 x 163 x 64 at base
// size), so the matrix products are warp-level mma.sync.m16n8k16 bf16 -> fp32 with ldmatrix fragments: a 128-row
// tcgen05 tile would be >= 50 % padding for the 64-row text tiles and needs a CTA-wide TMEM round trip per product.
// (The 97 % of the step's FLOPs that are large GEMMs run on tcgen05, gemm_kernel.cuh.)
// Backward = rows pass (dQ, dW0/dW1, delta) + keys pass (dK, dV; fp32 for prefix keys), both recomputing P.
#include "kernels.cuh"
#include <stdlib.h>

namespace {

constexpr int D = 64;
constexpr int LDS = 72;         // padded smem row (bf16 elements): 144 B keeps ldmatrix rows on distinct banks
constexpr int WARPS = 4;
constexpr float NEG_BIG = -1.0e30f;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// rows [row0, row0+16) x 64 columns of a bf16 global matrix as mma A fragments (4 k-steps); rows >= nrows_total read 0
__device__ __forceinline__ void load_a_bf16(const bf16* base, long ld, int row0, int nrows_total, int lane, uint32_t (&a)[4][4]) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {       // row g / g+8
            const int row = row0 + g + 8 * h;
            const bool ok = row < nrows_total;
            const bf16* p = base + (long)row * ld + ks * 16 + 2 * t;
            a[ks][h] = ok ? *reinterpret_cast<const uint32_t*>(p) : 0u;
            a[ks][h + 2] = ok ? *reinterpret_cast<const uint32_t*>(p + 8) : 0u;
        }
    }
}
// same from an fp32 matrix (rounded to bf16); also returns the fp32 values' dot with a bf16 matrix `o` (for delta)
__device__ __forceinline__ void load_a_f32(const float* base, long ld, int row0, int nrows_total, int lane, uint32_t (&a)[4][4],
                                           const bf16* o, long ldo, float (&dot)[2]) {
    const int g = lane >> 2, t = lane & 3;
    dot[0] = dot[1] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int row = row0 + g + 8 * h;
            const bool ok = row < nrows_total;
            float2 v0 = make_float2(0.f, 0.f), v1 = v0;
            if (ok) {
                const float* p = base + (long)row * ld + ks * 16 + 2 * t;
                v0 = *reinterpret_cast<const float2*>(p);
                v1 = *reinterpret_cast<const float2*>(p + 8);
                if (o) {
                    const bf16* q = o + (long)row * ldo + ks * 16 + 2 * t;
                    const float2 o0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(q));
                    const float2 o1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(q + 8));
                    dot[h] += v0.x * o0.x + v0.y * o0.y + v1.x * o1.x + v1.y * o1.y;
                }
            }
            a[ks][h] = pack_bf16x2(v0.x, v0.y);
            a[ks][h + 2] = pack_bf16x2(v1.x, v1.y);
        }
    }
}

// cooperative copy of `rows` x 64 bf16 (global stride ld) into padded smem rows; rows [rows, rows_pad) are zeroed
__device__ __forceinline__ void stage_bf16(bf16* dst, const bf16* src, long ld, int rows, int rows_pad, int tid, int nthreads) {
    for (int i = tid; i < rows_pad * 8; i += nthreads) {
        const int r = i >> 3, c = i & 7;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (r < rows) q = *reinterpret_cast<const uint4*>(src + (long)r * ld + c * 8);
        *reinterpret_cast<uint4*>(dst + r * LDS + c * 8) = q;
    }
}
__device__ __forceinline__ void stage_f32_as_bf16(bf16* dst, const float* src, long ld, int rows, int rows_pad, int tid, int nthreads) {
    for (int i = tid; i < rows_pad * 16; i += nthreads) {
        const int r = i >> 4, c = i & 15;
        uint2 q = make_uint2(0, 0);
        if (r < rows) {
            const float4 v = *reinterpret_cast<const float4*>(src + (long)r * ld + c * 4);
            q = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
        }
        *reinterpret_cast<uint2*>(dst + r * LDS + c * 4) = q;
    }
}

// asynchronous variant (cp.async, 16 B per request, zero fill for the padding rows): every request of a thread is in
// flight at once; finish with cp_async_wait_all() + __syncthreads()
__device__ __forceinline__ void stage_bf16_async(bf16* dst, const bf16* src, long ld, int rows, int rows_pad, int tid, int nthreads) {
    for (int i = tid; i < rows_pad * 8; i += nthreads) {
        const int r = i >> 3, c = i & 7;
        const bool ok = r < rows;
        const uint32_t d = smem_u32(dst + r * LDS + c * 8);
        const bf16* g = src + (long)(ok ? r : 0) * ld + c * 8;
        const int bytes = ok ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(g), "r"(bytes) : "memory");
    }
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// dO (fp32, global) -> bf16 rows in shared memory, and delta[r] = sum_d dO[r][d] * O[r][d] (O = forward context, bf16) in
// fp32.  16 consecutive lanes cover one row (4 columns each); loads are issued four rows ahead of their use.
__device__ __forceinline__ void stage_do_delta(bf16* dst, float* delta_s, const float* dsrc, long ldd, const bf16* osrc, long ldo, int rows,
                                               int rows_pad, int tid, int nthreads) {
    const int total = rows_pad * 16;
    for (int base = 0; base < total; base += 4 * nthreads) {      // uniform trip count: the shuffles below need whole warps
        const int i0 = base + tid;
        float4 v[4];
        uint2 o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * nthreads, r = i >> 4, c = i & 15;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            o[u] = make_uint2(0u, 0u);
            if (i < total && r < rows) {
                v[u] = *reinterpret_cast<const float4*>(dsrc + (long)r * ldd + c * 4);
                o[u] = *reinterpret_cast<const uint2*>(osrc + (long)r * ldo + c * 4);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * nthreads, r = i >> 4, c = i & 15;
            const float2 o0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&o[u].x));
            const float2 o1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&o[u].y));
            float part = v[u].x * o0.x + v[u].y * o0.y + v[u].z * o1.x + v[u].w * o1.y;
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            part += __shfl_xor_sync(0xffffffffu, part, 4);
            part += __shfl_xor_sync(0xffffffffu, part, 8);
            if (i < total) {
                *reinterpret_cast<uint2*>(dst + r * LDS + c * 4) = make_uint2(pack_bf16x2(v[u].x, v[u].y), pack_bf16x2(v[u].z, v[u].w));
                if (c == 0) delta_s[r] = part;
            }
        }
    }
}
// rows [row0, row0 + 16) x 64 columns of a padded shared-memory matrix as mma A fragments (4 k-steps)
__device__ __forceinline__ void load_a_smem(const bf16* As, int row0, int lane, uint32_t (&a)[4][4]) {
    const uint32_t base = smem_u32(As);
    const int lr = lane & 7, sel = lane >> 3;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
        ldsm_x4(base + (uint32_t)((row0 + lr + 8 * (sel & 1)) * LDS + ks * 16 + 8 * (sel >> 1)) * 2u, a[ks]);
}

// C[16 x 64 keys] = A[16 x 64] * Bs[key][d]^T for the 64 "n" rows starting at n0 of a padded smem matrix (n x k row-major)
__device__ __forceinline__ void mma_a_bt(float (&c)[8][4], const uint32_t (&a)[4][4], const bf16* Bs, int n0, int lane) {
    const uint32_t base = smem_u32(Bs);
    const int lr = lane & 7, sel = lane >> 3;          // ldmatrix address lanes: matrix sel = lane / 8
#pragma unroll
    for (int np = 0; np < 4; ++np) {                  // pairs of 8-wide n tiles
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t b[4];
            const int row = n0 + np * 16 + (sel >> 1) * 8 + lr;
            const int col = ks * 16 + (sel & 1) * 8;
            ldsm_x4(base + (uint32_t)(row * LDS + col) * 2u, b);
            mma16816(c[2 * np], a[ks], b[0], b[1]);
            mma16816(c[2 * np + 1], a[ks], b[2], b[3]);
        }
    }
}
// C[16 x 64 d] += P[16 x 64 k] * Bs[k][d]  (k x n row-major smem rows k0..k0+63), P given as packed A fragments
__device__ __forceinline__ void mma_p_b(float (&c)[8][4], const uint32_t (&pa)[4][4], const bf16* Bs, int k0, int lane) {
    const uint32_t base = smem_u32(Bs);
    const int lr = lane & 7, sel = lane >> 3;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b[4];
            const int row = k0 + ks * 16 + (sel & 1) * 8 + lr;
            const int col = np * 16 + (sel >> 1) * 8;
            ldsm_x4_t(base + (uint32_t)(row * LDS + col) * 2u, b);
            mma16816(c[2 * np], pa[ks], b[0], b[1]);
            mma16816(c[2 * np + 1], pa[ks], b[2], b[3]);
        }
    }
}
// The same three helpers for a chunk of NT n-tiles (8 NT keys, NT even) instead of 64 keys.  The backward kernels walk the keys
// in chunks of 16 / 32 (NT = 2 / 4): their per-chunk element-wise code is fully unrolled over the accumulator fragments, and at
// 64 keys the straight-line body of attn_bwd_fused was ~8,000 SASS instructions (128 KB) that every warp runs once or twice -
// ncu attributes 25 % (vision) to 50 % (text) of its stall samples to instruction fetch ("no_inst",
// profiles/attention_r2.txt).  A quarter of the body, looped four times as often, stays in the instruction caches.
template <int NT>
__device__ __forceinline__ void mma_a_bt_t(float (&c)[NT][4], const uint32_t (&a)[4][4], const bf16* Bs, int n0, int lane) {
    const uint32_t base = smem_u32(Bs);
    const int lr = lane & 7, sel = lane >> 3;
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t b[4];
            const int row = n0 + np * 16 + (sel >> 1) * 8 + lr;
            const int col = ks * 16 + (sel & 1) * 8;
            ldsm_x4(base + (uint32_t)(row * LDS + col) * 2u, b);
            mma16816(c[2 * np], a[ks], b[0], b[1]);
            mma16816(c[2 * np + 1], a[ks], b[2], b[3]);
        }
    }
}
template <int NT>
__device__ __forceinline__ void mma_p_b_t(float (&c)[8][4], const uint32_t (&pa)[NT / 2][4], const bf16* Bs, int k0, int lane) {
    const uint32_t base = smem_u32(Bs);
    const int lr = lane & 7, sel = lane >> 3;
#pragma unroll
    for (int ks = 0; ks < NT / 2; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b[4];
            const int row = k0 + ks * 16 + (sel & 1) * 8 + lr;
            const int col = np * 16 + (sel >> 1) * 8;
            ldsm_x4_t(base + (uint32_t)(row * LDS + col) * 2u, b);
            mma16816(c[2 * np], pa[ks], b[0], b[1]);
            mma16816(c[2 * np + 1], pa[ks], b[2], b[3]);
        }
    }
}
template <int NT>
__device__ __forceinline__ void acc_to_a_t(const float (&c)[NT][4], uint32_t (&a)[NT / 2][4]) {
#pragma unroll
    for (int ks = 0; ks < NT / 2; ++ks) {
        a[ks][0] = pack_bf16x2(c[2 * ks][0], c[2 * ks][1]);
        a[ks][1] = pack_bf16x2(c[2 * ks][2], c[2 * ks][3]);
        a[ks][2] = pack_bf16x2(c[2 * ks + 1][0], c[2 * ks + 1][1]);
        a[ks][3] = pack_bf16x2(c[2 * ks + 1][2], c[2 * ks + 1][3]);
    }
}
// accumulator tiles (16 x 64) -> packed bf16 A fragments (4 k-steps of 16)
__device__ __forceinline__ void acc_to_a(const float (&c)[8][4], uint32_t (&a)[4][4]) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        a[ks][0] = pack_bf16x2(c[2 * ks][0], c[2 * ks][1]);
        a[ks][1] = pack_bf16x2(c[2 * ks][2], c[2 * ks][3]);
        a[ks][2] = pack_bf16x2(c[2 * ks + 1][0], c[2 * ks + 1][1]);
        a[ks][3] = pack_bf16x2(c[2 * ks + 1][2], c[2 * ks + 1][3]);
    }
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    return v;
}

struct Adapt { int split, row_lo; float w0c, w1c; };
__device__ __forceinline__ Adapt load_adapt(const AttnFwd& a, int b) {
    Adapt ad;
    ad.split = -1; ad.row_lo = 0; ad.w0c = ad.w1c = 1.f;
    if (a.sep_idx) {
        ad.split = (int)a.sep_idx[(long)b * a.sep_stride];
        ad.row_lo = a.row_lo;
        ad.w0c = fminf(fmaxf(*a.w0, 0.f), 0.5f);
        ad.w1c = fminf(fmaxf(*a.w1, 0.5f), 1.0f);
    }
    return ad;
}
// factor of score (query row r, own-key index j2 = key - n1); prefix keys (j2 < 0) are never scaled
__device__ __forceinline__ float afactor(const Adapt& ad, int r, int j2) {
    if (ad.split < 0 || j2 < ad.split) return 1.0f;
    if (r >= ad.split) return ad.w1c;
    return (r >= ad.row_lo) ? ad.w0c : 1.0f;
}

// shared: K/V staging of all keys (prefix first) + additive mask per key (NEG_BIG for padding keys >= n)
__device__ __forceinline__ void stage_keys(const AttnFwd& a, int b, int h, bf16* Ks, bf16* Vs, float* madd, int n_pad, int tid) {
    const int n1 = a.n1, n2 = a.n2, n = n1 + n2;
    if (n1 > 0) {
        stage_bf16(Ks, a.k1 + (long)b * n1 * a.ld1 + h * D, a.ld1, n1, n1, tid, WARPS * 32);
        stage_bf16(Vs, a.v1 + (long)b * n1 * a.ld1 + h * D, a.ld1, n1, n1, tid, WARPS * 32);
    }
    stage_bf16(Ks + n1 * LDS, a.k2 + (long)b * n2 * a.ld2 + h * D, a.ld2, n2, n_pad - n1, tid, WARPS * 32);
    stage_bf16(Vs + n1 * LDS, a.v2 + (long)b * n2 * a.ld2 + h * D, a.ld2, n2, n_pad - n1, tid, WARPS * 32);
    for (int j = tid; j < n_pad; j += WARPS * 32) {
        float m = 0.f;
        if (j >= n) m = NEG_BIG;
        else if (a.key_mask && j >= n1) m = (1.0f - (float)a.key_mask[(long)b * n2 + (j - n1)]) * -10000.0f;
        madd[j] = m;
    }
}

// the same with cp.async (every request of the CTA in flight at once); finish with cp_async_wait_all() + __syncthreads()
__device__ __forceinline__ void stage_keys_async(const AttnFwd& a, int b, int h, bf16* Ks, bf16* Vs, float* madd, int n_pad, int tid, int nthreads) {
    const int n1 = a.n1, n2 = a.n2, n = n1 + n2;
    if (n1 > 0) {
        stage_bf16_async(Ks, a.k1 + (long)b * n1 * a.ld1 + h * D, a.ld1, n1, n1, tid, nthreads);
        stage_bf16_async(Vs, a.v1 + (long)b * n1 * a.ld1 + h * D, a.ld1, n1, n1, tid, nthreads);
    }
    stage_bf16_async(Ks + n1 * LDS, a.k2 + (long)b * n2 * a.ld2 + h * D, a.ld2, n2, n_pad - n1, tid, nthreads);
    stage_bf16_async(Vs + n1 * LDS, a.v2 + (long)b * n2 * a.ld2 + h * D, a.ld2, n2, n_pad - n1, tid, nthreads);
    for (int j = tid; j < n_pad; j += nthreads) {
        float m = 0.f;
        if (j >= n) m = NEG_BIG;
        else if (a.key_mask && j >= n1) m = (1.0f - (float)a.key_mask[(long)b * n2 + (j - n1)]) * -10000.0f;
        madd[j] = m;
    }
}

// partial row sums of P_ij * mult_ij * dP_ij over one 16 x 64 chunk (s = raw QK^T accumulators, dp = dO V^T accumulators):
// the softmax-backward delta that is consistent with the dP values the caller goes on to use
template <int NT>
__device__ __forceinline__ void consistent_delta(const AttnFwd& a, const Adapt& ad, const float (&s)[NT][4], const float (&dp)[NT][4],
                                                 const float* madd, const float (&lse)[2], int kc, int r0, int g, int t, int n1, int n,
                                                 int Lq, int b, int h, float (&acc)[2]) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int col = kc + nt * 8 + 2 * t + (e & 1);
            const int row = r0 + g + 8 * (e >> 1);
            const float f = afactor(ad, row, col - n1);
            float p = (row < Lq) ? __expf(s[nt][e] * a.scale * f + madd[col] - lse[e >> 1]) : 0.f;
            if (a.drop.thresh) p *= drop_mult(a.drop, (uint32_t)((((long)b * a.nh + h) * Lq + row) * n + col));
            acc[e >> 1] = fmaf(p, dp[nt][e], acc[e >> 1]);
        }
    }
}

// ------------------------------------------------------------------------------------------ forward
// NW warps of 16 query rows each: 4 for the text stream (L <= 64: one round), 8 for the vision stream (99 rows: one round
// instead of two over the same staged K / V)
template <int NW>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 2 : 3)
attn_fwd_mma(const AttnFwd a, float* __restrict__ out32) {
    extern __shared__ __align__(16) uint8_t smem[];
    pdl_wait();
    pdl_launch_dependents();
    const int h = blockIdx.x, b = blockIdx.y;
    const int n1 = a.n1, n = a.n1 + a.n2, n_pad = (n + 63) & ~63;
    bf16* Ks = reinterpret_cast<bf16*>(smem);
    bf16* Vs = Ks + (size_t)n_pad * LDS;
    float* madd = reinterpret_cast<float*>(Vs + (size_t)n_pad * LDS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    stage_keys_async(a, b, h, Ks, Vs, madd, n_pad, tid, NW * 32);
    const Adapt ad = load_adapt(a, b);
    const bf16* qbase = a.q + (long)b * a.Lq * a.ldq + h * D;
    cp_async_wait_all();
    __syncthreads();

    for (int r0 = warp * 16; r0 < a.Lq; r0 += NW * 16) {
        uint32_t qa[4][4];
        load_a_bf16(qbase, a.ldq, r0, a.Lq, lane, qa);
        float o[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
        float mrun[2] = {NEG_BIG, NEG_BIG}, lrun[2] = {0.f, 0.f};
        for (int kc = 0; kc < n_pad; kc += 64) {
            float s[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
            mma_a_bt(s, qa, Ks, kc, lane);
            float mx[2] = {NEG_BIG, NEG_BIG};
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int col = kc + nt * 8 + 2 * t + (e & 1);
                    const int row = r0 + g + 8 * (e >> 1);
                    const float v = s[nt][e] * a.scale * afactor(ad, row, col - n1) + madd[col];
                    s[nt][e] = v;
                    mx[e >> 1] = fmaxf(mx[e >> 1], v);
                }
            }
            float alpha[2];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const float mnew = fmaxf(mrun[hh], quad_max(mx[hh]));
                alpha[hh] = __expf(mrun[hh] - mnew);
                mrun[hh] = mnew;
                lrun[hh] *= alpha[hh];
            }
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float p = __expf(s[nt][e] - mrun[e >> 1]);
                    lrun[e >> 1] += p;
                    if (a.drop.thresh) {   // dropout on the normalised probabilities (:362): the row sum stays undropped
                        const int col = kc + nt * 8 + 2 * t + (e & 1);
                        const int row = r0 + g + 8 * (e >> 1);
                        p *= drop_mult(a.drop, (uint32_t)((((long)b * a.nh + h) * a.Lq + row) * n + col));
                    }
                    s[nt][e] = p;
                }
                o[nt][0] *= alpha[0]; o[nt][1] *= alpha[0]; o[nt][2] *= alpha[1]; o[nt][3] *= alpha[1];
            }
            uint32_t pa[4][4];
            acc_to_a(s, pa);
            mma_p_b(o, pa, Vs, kc, lane);
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int row = r0 + g + 8 * hh;
            const float l = quad_sum(lrun[hh]);
            const float inv = 1.0f / l;
            if (row < a.Lq) {
                if (t == 0 && a.lse) a.lse[((long)b * a.nh + h) * a.Lq + row] = mrun[hh] + __logf(l);
                const long off = ((long)b * a.Lq + row) * a.ldo + h * D;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const float x0 = o[nt][2 * hh] * inv, x1 = o[nt][2 * hh + 1] * inv;
                    *reinterpret_cast<uint32_t*>(a.out + off + nt * 8 + 2 * t) = pack_bf16x2(x0, x1);
                    if (out32) *reinterpret_cast<float2*>(out32 + off + nt * 8 + 2 * t) = make_float2(x0, x1);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ backward, rows pass
// per 16 query rows: delta = rowsum(dO * O); over key chunks: P, dP = dO V^T, dS = P (dP - delta); dQ += (dS f) K;
// dW0/dW1 partial sums of dS * S_raw over the scaled blocks.
__global__ void __launch_bounds__(WARPS * 32)
attn_bwd_rows_mma(const AttnBwd gr, const bf16* __restrict__ ctx, long ldctx, float* __restrict__ delta) {
    extern __shared__ __align__(16) uint8_t smem[];
    const AttnFwd& a = gr.f;
    const int h = blockIdx.x, b = blockIdx.y;
    const int n1 = a.n1, n = a.n1 + a.n2, n_pad = (n + 63) & ~63;
    bf16* Ks = reinterpret_cast<bf16*>(smem);
    bf16* Vs = Ks + (size_t)n_pad * LDS;
    float* madd = reinterpret_cast<float*>(Vs + (size_t)n_pad * LDS);
    __shared__ float wred[2][WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    stage_keys(a, b, h, Ks, Vs, madd, n_pad, tid);
    const Adapt ad = load_adapt(a, b);
    __syncthreads();

    const bf16* qbase = a.q + (long)b * a.Lq * a.ldq + h * D;
    const float* dobase = gr.dout + (long)b * a.Lq * gr.lddo + h * D;
    const bf16* obase = ctx + (long)b * a.Lq * ldctx + h * D;
    float dw0 = 0.f, dw1 = 0.f;
    for (int r0 = warp * 16; r0 < a.Lq; r0 += WARPS * 16) {
        uint32_t qa[4][4], doa[4][4];
        float dl[2];
        load_a_bf16(qbase, a.ldq, r0, a.Lq, lane, qa);
        load_a_f32(dobase, gr.lddo, r0, a.Lq, lane, doa, obase, ldctx, dl);
        float lse[2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            dl[hh] = quad_sum(dl[hh]);
            const int row = r0 + g + 8 * hh;
            lse[hh] = (row < a.Lq) ? a.lse[((long)b * a.nh + h) * a.Lq + row] : 0.f;
            if (t == 0 && row < a.Lq) delta[((long)b * a.nh + h) * a.Lq + row] = dl[hh];
        }
        if (ad.split >= 0) {   // adaptive scaling: delta consistent with this kernel's own dP (see attn_bwd_fused); the keys pass reads it
            float acc[2] = {0.f, 0.f};
            for (int kc = 0; kc < n_pad; kc += 64) {
                float s[8][4], dp[8][4];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
                    dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
                }
                mma_a_bt(s, qa, Ks, kc, lane);
                mma_a_bt(dp, doa, Vs, kc, lane);
                consistent_delta<8>(a, ad, s, dp, madd, lse, kc, r0, g, t, n1, n, a.Lq, b, h, acc);
            }
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                dl[hh] = quad_sum(acc[hh]);
                const int row = r0 + g + 8 * hh;
                if (t == 0 && row < a.Lq) delta[((long)b * a.nh + h) * a.Lq + row] = dl[hh];
            }
        }
        float dq[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
        for (int kc = 0; kc < n_pad; kc += 64) {
            float s[8][4], dp[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
                dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
            }
            mma_a_bt(s, qa, Ks, kc, lane);
            mma_a_bt(dp, doa, Vs, kc, lane);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int col = kc + nt * 8 + 2 * t + (e & 1);
                    const int row = r0 + g + 8 * (e >> 1);
                    const int j2 = col - n1;
                    const float f = afactor(ad, row, j2);
                    const float sraw = s[nt][e] * a.scale;
                    const float p = (row < a.Lq) ? __expf(sraw * f + madd[col] - lse[e >> 1]) : 0.f;
                    float dpe = dp[nt][e];
                    if (a.drop.thresh) dpe *= drop_mult(a.drop, (uint32_t)((((long)b * a.nh + h) * a.Lq + row) * n + col));
                    const float ds = p * (dpe - dl[e >> 1]);
                    if (ad.split >= 0 && j2 >= ad.split && col < n) {
                        if (row >= ad.split) dw1 = fmaf(ds, sraw, dw1);
                        else if (row >= ad.row_lo) dw0 = fmaf(ds, sraw, dw0);
                    }
                    s[nt][e] = ds * f;
                }
            }
            uint32_t da[4][4];
            acc_to_a(s, da);
            mma_p_b(dq, da, Ks, kc, lane);
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int row = r0 + g + 8 * hh;
            if (row < a.Lq) {
                const long off = ((long)b * a.Lq + row) * gr.ldd + h * D;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
                    *reinterpret_cast<uint32_t*>(gr.dq + off + nt * 8 + 2 * t) =
                        pack_bf16x2(dq[nt][2 * hh] * a.scale, dq[nt][2 * hh + 1] * a.scale);
            }
        }
    }
    if (a.sep_idx) {
        dw0 = warp_sum(dw0);
        dw1 = warp_sum(dw1);
        if (lane == 0) { wred[0][warp] = dw0; wred[1][warp] = dw1; }
        __syncthreads();
        if (tid == 0) {
            float s0 = 0.f, s1 = 0.f;
            for (int w = 0; w < WARPS; ++w) { s0 += wred[0][w]; s1 += wred[1][w]; }
            // torch.clamp passes the gradient at both (inclusive) bounds
            const float w0 = *a.w0, w1 = *a.w1;
            if (w0 >= 0.f && w0 <= 0.5f) atomicAdd(gr.dw0, s0);
            if (w1 >= 0.5f && w1 <= 1.0f) atomicAdd(gr.dw1, s1);
        }
    }
}

// ------------------------------------------------------------------------------------------ backward, keys pass
// per 16 keys: over query-row chunks: S^T = K Q^T, dP^T = V dO^T, P^T, dS^T; dV += P^T dO; dK += (dS^T f) Q.
__global__ void __launch_bounds__(WARPS * 32)
attn_bwd_keys_mma(const AttnBwd gr, const float* __restrict__ delta) {
    extern __shared__ __align__(16) uint8_t smem[];
    const AttnFwd& a = gr.f;
    const int h = blockIdx.x, b = blockIdx.y;
    const int n1 = a.n1, n2 = a.n2, n = n1 + n2, Lq = a.Lq, Lq_pad = (Lq + 63) & ~63;
    bf16* Qs = reinterpret_cast<bf16*>(smem);
    bf16* Os = Qs + (size_t)Lq_pad * LDS;
    float* lse_s = reinterpret_cast<float*>(Os + (size_t)Lq_pad * LDS);
    float* dl_s = lse_s + Lq_pad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    stage_bf16(Qs, a.q + (long)b * Lq * a.ldq + h * D, a.ldq, Lq, Lq_pad, tid, WARPS * 32);
    stage_f32_as_bf16(Os, gr.dout + (long)b * Lq * gr.lddo + h * D, gr.lddo, Lq, Lq_pad, tid, WARPS * 32);
    for (int r = tid; r < Lq_pad; r += WARPS * 32) {
        lse_s[r] = (r < Lq) ? a.lse[((long)b * a.nh + h) * Lq + r] : 0.f;
        dl_s[r] = (r < Lq) ? delta[((long)b * a.nh + h) * Lq + r] : 0.f;
    }
    const Adapt ad = load_adapt(a, b);
    __syncthreads();

    for (int j0 = warp * 16; j0 < n; j0 += WARPS * 16) {
        // this block's 16 keys may straddle the prefix / own boundary: fragments are gathered row by row
        uint32_t ka[4][4], va[4][4];
        float madd[2];
        int j2v[2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int j = j0 + g + 8 * hh;
            const bool ok = j < n, prefix = j < n1;
            const bf16* kr = prefix ? a.k1 + ((long)b * n1 + j) * a.ld1 + h * D : a.k2 + ((long)b * n2 + (j - n1)) * a.ld2 + h * D;
            const bf16* vr = prefix ? a.v1 + ((long)b * n1 + j) * a.ld1 + h * D : a.v2 + ((long)b * n2 + (j - n1)) * a.ld2 + h * D;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                ka[ks][hh] = ok ? *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 2 * t) : 0u;
                ka[ks][hh + 2] = ok ? *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 2 * t + 8) : 0u;
                va[ks][hh] = ok ? *reinterpret_cast<const uint32_t*>(vr + ks * 16 + 2 * t) : 0u;
                va[ks][hh + 2] = ok ? *reinterpret_cast<const uint32_t*>(vr + ks * 16 + 2 * t + 8) : 0u;
            }
            j2v[hh] = j - n1;
            madd[hh] = 0.f;
            if (!ok) madd[hh] = NEG_BIG;
            else if (a.key_mask && !prefix) madd[hh] = (1.0f - (float)a.key_mask[(long)b * n2 + (j - n1)]) * -10000.0f;
        }
        float dk[8][4], dv[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
            dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
        }
        for (int rc = 0; rc < Lq_pad; rc += 64) {
            float st[8][4], dpt[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
                dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
            }
            mma_a_bt(st, ka, Qs, rc, lane);
            mma_a_bt(dpt, va, Os, rc, lane);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int r = rc + nt * 8 + 2 * t + (e & 1);      // query row (column of the transposed tile)
                    const int hh = e >> 1;                            // key row g / g+8
                    const float f = afactor(ad, r, j2v[hh]);
                    const float p = (r < Lq) ? __expf(st[nt][e] * a.scale * f + madd[hh] - lse_s[r]) : 0.f;
                    float mult = 1.f;
                    if (a.drop.thresh)
                        mult = drop_mult(a.drop, (uint32_t)((((long)b * a.nh + h) * Lq + r) * n + (j0 + g + 8 * hh)));
                    const float ds = p * (dpt[nt][e] * mult - dl_s[r]);
                    dpt[nt][e] = p * mult;    // dropped P^T
                    st[nt][e] = ds * f;       // dS^T * f
                }
            }
            uint32_t pa[4][4], da[4][4];
            acc_to_a(dpt, pa);
            acc_to_a(st, da);
            mma_p_b(dv, pa, Os, rc, lane);
            mma_p_b(dk, da, Qs, rc, lane);
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int j = j0 + g + 8 * hh;
            if (j >= n) continue;
            if (j < n1) {
                const long off = ((long)b * n1 + j) * gr.ldd1 + h * D;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    *reinterpret_cast<float2*>(gr.dk1 + off + nt * 8 + 2 * t) = make_float2(dk[nt][2 * hh] * a.scale, dk[nt][2 * hh + 1] * a.scale);
                    *reinterpret_cast<float2*>(gr.dv1 + off + nt * 8 + 2 * t) = make_float2(dv[nt][2 * hh], dv[nt][2 * hh + 1]);
                }
            } else {
                const long row = (long)b * n2 + (j - n1);
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    float k0 = dk[nt][2 * hh] * a.scale, k1 = dk[nt][2 * hh + 1] * a.scale;
                    float v0 = dv[nt][2 * hh], v1 = dv[nt][2 * hh + 1];
                    if (gr.dk2_add) {
                        const float2 ak = *reinterpret_cast<const float2*>(gr.dk2_add + row * gr.ldadd + h * D + nt * 8 + 2 * t);
                        const float2 av = *reinterpret_cast<const float2*>(gr.dv2_add + row * gr.ldadd + h * D + nt * 8 + 2 * t);
                        k0 += ak.x; k1 += ak.y; v0 += av.x; v1 += av.y;
                    }
                    *reinterpret_cast<uint32_t*>(gr.dk2 + row * gr.ldd + h * D + nt * 8 + 2 * t) = pack_bf16x2(k0, k1);
                    *reinterpret_cast<uint32_t*>(gr.dv2 + row * gr.ldd + h * D + nt * 8 + 2 * t) = pack_bf16x2(v0, v1);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ backward, tiled two-pass
// The two passes above for LONG sequences (FLAVA 197 / 458 tokens, ViLT 418, VisualBERT 200, MKGformer-large 513 + prefix), where
// one CTA of four warps per (head, sample) with every key (or query row) resident in shared memory left 2 CTAs per SM walking
// 4 - 9 rounds of 16-row blocks through a ~8,000-instruction straight-line body (CUPTI: 40 % of the FLAVA / ViLT / VisualBERT
// steps).  Here the (head, sample) work is split over gridDim.z CTAs - warp w of CTA z owns 16-row block w * Z + z, so every
// CTA carries the same number of blocks (+- 1) - the other side is streamed through shared memory in 64-row chunks (18 KB instead of
// 74 - 166 KB: shared memory no longer bounds residency) and each chunk is consumed in sub-chunks of 8 NT keys by a loop that
// stays in the instruction caches (as attn_bwd_fused).
constexpr int TC = 64;      // rows of the streamed side per staged chunk

// rows [kc, kc + TC) of the concatenated [prefix ; own] keys -> Ks / Vs (padded rows); rows >= n are zero-filled
__device__ __forceinline__ void stage_kv_chunk_async(const AttnFwd& a, int b, int h, int kc, bf16* Ks, bf16* Vs, int tid, int nthreads) {
    const int n1 = a.n1, n = a.n1 + a.n2;
    for (int i = tid; i < TC * 8; i += nthreads) {
        const int r = i >> 3, c = i & 7, j = kc + r;
        const bool ok = j < n, prefix = j < n1;
        const long row = ok ? (prefix ? (long)b * n1 + j : (long)b * a.n2 + (j - n1)) : 0;
        const bf16* kg = (prefix ? a.k1 + row * a.ld1 : a.k2 + row * a.ld2) + h * D + c * 8;
        const bf16* vg = (prefix ? a.v1 + row * a.ld1 : a.v2 + row * a.ld2) + h * D + c * 8;
        const int bytes = ok ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(Ks + r * LDS + c * 8)), "l"(kg), "r"(bytes) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(Vs + r * LDS + c * 8)), "l"(vg), "r"(bytes) : "memory");
    }
}

template <int NT, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
attn_bwd_rows_tiled(const AttnBwd gr, const bf16* __restrict__ ctx, long ldctx, float* __restrict__ delta) {
    extern __shared__ __align__(16) uint8_t smem[];
    const AttnFwd& a = gr.f;
    const int h = blockIdx.x, b = blockIdx.y, z = blockIdx.z, Z = gridDim.z;
    const int n1 = a.n1, n = a.n1 + a.n2, n_pad = (n + 63) & ~63;
    bf16* KV = reinterpret_cast<bf16*>(smem);                  // two stages of [K chunk ; V chunk]
    float* madd = reinterpret_cast<float*>(KV + (size_t)4 * TC * LDS);
    __shared__ float wred[2][WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    for (int j = tid; j < n_pad; j += WARPS * 32) {
        float m = 0.f;
        if (j >= n) m = NEG_BIG;
        else if (a.key_mask && j >= n1) m = (1.0f - (float)a.key_mask[(long)b * a.n2 + (j - n1)]) * -10000.0f;
        madd[j] = m;
    }
    const Adapt ad = load_adapt(a, b);
    const int r0 = (warp * Z + z) * 16;          // this warp's 16 query rows
    const bool active = r0 < a.Lq;
    uint32_t qa[4][4], doa[4][4];
    float dl[2] = {0.f, 0.f}, lse[2] = {0.f, 0.f};
    if (active) {
        load_a_bf16(a.q + (long)b * a.Lq * a.ldq + h * D, a.ldq, r0, a.Lq, lane, qa);
        load_a_f32(gr.dout + (long)b * a.Lq * gr.lddo + h * D, gr.lddo, r0, a.Lq, lane, doa, ctx + (long)b * a.Lq * ldctx + h * D, ldctx, dl);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            dl[hh] = quad_sum(dl[hh]);
            const int row = r0 + g + 8 * hh;
            lse[hh] = (row < a.Lq) ? a.lse[((long)b * a.nh + h) * a.Lq + row] : 0.f;
        }
    }
    float dw0 = 0.f, dw1 = 0.f;
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    // pass 0 (adaptive scaling only): delta consistent with this kernel's own dP (see attn_bwd_fused); pass 1: dS, dQ, dW0 / dW1
    for (int pass = (ad.split >= 0) ? 0 : 1; pass < 2; ++pass) {
        float acc[2] = {0.f, 0.f};
        __syncthreads();                         // both stages are free (first pass: madd is complete)
        stage_kv_chunk_async(a, b, h, 0, KV, KV + (size_t)TC * LDS, tid, WARPS * 32);
        cp_async_commit();
        for (int kc = 0, it = 0; kc < n_pad; kc += TC, ++it) {
            const bf16* Ks = KV + (size_t)(it & 1) * 2 * TC * LDS;
            const bf16* Vs = Ks + (size_t)TC * LDS;
            if (kc + TC < n_pad) {               // the next chunk streams in under this chunk's products
                bf16* Kn = KV + (size_t)((it + 1) & 1) * 2 * TC * LDS;
                stage_kv_chunk_async(a, b, h, kc + TC, Kn, Kn + (size_t)TC * LDS, tid, WARPS * 32);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();                     // chunk `it` has landed for every thread
            if (active) {
#pragma unroll 1
            for (int sub = 0; sub < TC; sub += 8 * NT) {
                float s[NT][4], dp[NT][4];
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
                    dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
                }
                mma_a_bt_t<NT>(s, qa, Ks, sub, lane);
                mma_a_bt_t<NT>(dp, doa, Vs, sub, lane);
                if (pass == 0) {
                    consistent_delta<NT>(a, ad, s, dp, madd, lse, kc + sub, r0, g, t, n1, n, a.Lq, b, h, acc);
                    continue;
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int col = kc + sub + nt * 8 + 2 * t + (e & 1);
                        const int row = r0 + g + 8 * (e >> 1);
                        const int j2 = col - n1;
                        const float f = afactor(ad, row, j2);
                        const float sraw = s[nt][e] * a.scale;
                        const float p = (row < a.Lq) ? __expf(sraw * f + madd[col] - lse[e >> 1]) : 0.f;
                        float dpe = dp[nt][e];
                        if (a.drop.thresh) dpe *= drop_mult(a.drop, (uint32_t)((((long)b * a.nh + h) * a.Lq + row) * n + col));
                        const float ds = p * (dpe - dl[e >> 1]);
                        if (ad.split >= 0 && j2 >= ad.split && col < n) {
                            if (row >= ad.split) dw1 = fmaf(ds, sraw, dw1);
                            else if (row >= ad.row_lo) dw0 = fmaf(ds, sraw, dw0);
                        }
                        s[nt][e] = ds * f;
                    }
                }
                uint32_t da[NT / 2][4];
                acc_to_a_t<NT>(s, da);
                mma_p_b_t<NT>(dq, da, Ks, sub, lane);
            }
            }
            __syncthreads();                     // stage it & 1 may be refilled (by the trip after next)
        }
        if (pass == 0 && active) {
            dl[0] = quad_sum(acc[0]);
            dl[1] = quad_sum(acc[1]);
        }
    }
    if (active) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int row = r0 + g + 8 * hh;
            if (row < a.Lq) {
                if (t == 0) delta[((long)b * a.nh + h) * a.Lq + row] = dl[hh];      // read by the keys pass
                const long off = ((long)b * a.Lq + row) * gr.ldd + h * D;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
                    *reinterpret_cast<uint32_t*>(gr.dq + off + nt * 8 + 2 * t) =
                        pack_bf16x2(dq[nt][2 * hh] * a.scale, dq[nt][2 * hh + 1] * a.scale);
            }
        }
    }
    if (a.sep_idx) {
        dw0 = warp_sum(dw0);
        dw1 = warp_sum(dw1);
        if (lane == 0) { wred[0][warp] = dw0; wred[1][warp] = dw1; }
        __syncthreads();
        if (tid == 0) {
            float s0 = 0.f, s1 = 0.f;
            for (int w = 0; w < WARPS; ++w) { s0 += wred[0][w]; s1 += wred[1][w]; }
            const float w0 = *a.w0, w1 = *a.w1;      // torch.clamp passes the gradient at both (inclusive) bounds
            if (w0 >= 0.f && w0 <= 0.5f) atomicAdd(gr.dw0, s0);
            if (w1 >= 0.5f && w1 <= 1.0f) atomicAdd(gr.dw1, s1);
        }
    }
}

template <int NT, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
attn_bwd_keys_tiled(const AttnBwd gr, const float* __restrict__ delta) {
    extern __shared__ __align__(16) uint8_t smem[];
    const AttnFwd& a = gr.f;
    const int h = blockIdx.x, b = blockIdx.y, z = blockIdx.z, Z = gridDim.z;
    const int n1 = a.n1, n2 = a.n2, n = n1 + n2, Lq = a.Lq, Lq_pad = (Lq + 63) & ~63;
    bf16* QO = reinterpret_cast<bf16*>(smem);                  // two stages of [Q chunk ; dO chunk (bf16)]
    float* lse_s = reinterpret_cast<float*>(QO + (size_t)4 * TC * LDS);
    float* dl_s = lse_s + Lq_pad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    for (int r = tid; r < Lq_pad; r += WARPS * 32) {
        lse_s[r] = (r < Lq) ? a.lse[((long)b * a.nh + h) * Lq + r] : 0.f;
        dl_s[r] = (r < Lq) ? delta[((long)b * a.nh + h) * Lq + r] : 0.f;
    }
    const Adapt ad = load_adapt(a, b);
    const int j0 = (warp * Z + z) * 16;          // this warp's 16 keys
    const bool active = j0 < n;
    // the block's 16 keys may straddle the prefix / own boundary: fragments are gathered row by row
    uint32_t ka[4][4], va[4][4];
    float madd[2] = {NEG_BIG, NEG_BIG};
    int j2v[2] = {0, 0};
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int j = j0 + g + 8 * hh;
        const bool ok = j < n, prefix = j < n1;
        const long row = ok ? (prefix ? (long)b * n1 + j : (long)b * n2 + (j - n1)) : 0;
        const bf16* kr = (prefix ? a.k1 + row * a.ld1 : a.k2 + row * a.ld2) + h * D;
        const bf16* vr = (prefix ? a.v1 + row * a.ld1 : a.v2 + row * a.ld2) + h * D;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            ka[ks][hh] = ok ? *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 2 * t) : 0u;
            ka[ks][hh + 2] = ok ? *reinterpret_cast<const uint32_t*>(kr + ks * 16 + 2 * t + 8) : 0u;
            va[ks][hh] = ok ? *reinterpret_cast<const uint32_t*>(vr + ks * 16 + 2 * t) : 0u;
            va[ks][hh + 2] = ok ? *reinterpret_cast<const uint32_t*>(vr + ks * 16 + 2 * t + 8) : 0u;
        }
        j2v[hh] = j - n1;
        if (ok) madd[hh] = (a.key_mask && !prefix) ? (1.0f - (float)a.key_mask[(long)b * n2 + (j - n1)]) * -10000.0f : 0.f;
    }
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
        dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    const bf16* qbase = a.q + (long)b * Lq * a.ldq + h * D;
    const float* dobase = gr.dout + (long)b * Lq * gr.lddo + h * D;
    const bf16* do16 = gr.dout16 ? gr.dout16 + (long)b * Lq * gr.lddo16 + h * D : nullptr;
    // chunk rc -> stage st: Q rows through cp.async; dO through cp.async when the caller supplied a bf16 copy, else converted here
    auto stage = [&](int rc, int st) {
        bf16* Qn = QO + (size_t)st * 2 * TC * LDS;
        bf16* On = Qn + (size_t)TC * LDS;
        const int rows = min(TC, Lq - rc);
        stage_bf16_async(Qn, qbase + (long)rc * a.ldq, a.ldq, rows, TC, tid, WARPS * 32);
        if (do16) stage_bf16_async(On, do16 + (long)rc * gr.lddo16, gr.lddo16, rows, TC, tid, WARPS * 32);
        else stage_f32_as_bf16(On, dobase + (long)rc * gr.lddo, gr.lddo, rows, TC, tid, WARPS * 32);
        cp_async_commit();
    };
    stage(0, 0);
    for (int rc = 0, it = 0; rc < Lq_pad; rc += TC, ++it) {
        const bf16* Qs = QO + (size_t)(it & 1) * 2 * TC * LDS;
        const bf16* Os = Qs + (size_t)TC * LDS;
        if (rc + TC < Lq_pad) {                  // the next chunk streams in under this chunk's products
            stage(rc + TC, (it + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();                         // chunk `it` has landed for every thread (first trip: lse_s / dl_s are complete)
        if (active) {
#pragma unroll 1
        for (int sub = 0; sub < TC; sub += 8 * NT) {
            float st[NT][4], dpt[NT][4];
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
                dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
            }
            mma_a_bt_t<NT>(st, ka, Qs, sub, lane);
            mma_a_bt_t<NT>(dpt, va, Os, sub, lane);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int r = rc + sub + nt * 8 + 2 * t + (e & 1);      // query row (column of the transposed tile)
                    const int hh = e >> 1;                                  // key row g / g + 8
                    const float f = afactor(ad, r, j2v[hh]);
                    const float p = (r < Lq) ? __expf(st[nt][e] * a.scale * f + madd[hh] - lse_s[r]) : 0.f;
                    float mult = 1.f;
                    if (a.drop.thresh)
                        mult = drop_mult(a.drop, (uint32_t)((((long)b * a.nh + h) * Lq + r) * n + (j0 + g + 8 * hh)));
                    const float ds = p * (dpt[nt][e] * mult - dl_s[r]);
                    dpt[nt][e] = p * mult;    // dropped P^T
                    st[nt][e] = ds * f;       // dS^T * f
                }
            }
            uint32_t pa[NT / 2][4], da[NT / 2][4];
            acc_to_a_t<NT>(dpt, pa);
            acc_to_a_t<NT>(st, da);
            mma_p_b_t<NT>(dv, pa, Os, sub, lane);
            mma_p_b_t<NT>(dk, da, Qs, sub, lane);
        }
        }
        __syncthreads();                         // stage it & 1 may be refilled (by the trip after next)
    }
    if (!active) return;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int j = j0 + g + 8 * hh;
        if (j >= n) continue;
        if (j < n1) {
            const long off = ((long)b * n1 + j) * gr.ldd1 + h * D;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                *reinterpret_cast<float2*>(gr.dk1 + off + nt * 8 + 2 * t) = make_float2(dk[nt][2 * hh] * a.scale, dk[nt][2 * hh + 1] * a.scale);
                *reinterpret_cast<float2*>(gr.dv1 + off + nt * 8 + 2 * t) = make_float2(dv[nt][2 * hh], dv[nt][2 * hh + 1]);
            }
        } else {
            const long row = (long)b * n2 + (j - n1);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                float k0 = dk[nt][2 * hh] * a.scale, k1 = dk[nt][2 * hh + 1] * a.scale;
                float v0 = dv[nt][2 * hh], v1 = dv[nt][2 * hh + 1];
                if (gr.dk2_add) {
                    const float2 ak = *reinterpret_cast<const float2*>(gr.dk2_add + row * gr.ldadd + h * D + nt * 8 + 2 * t);
                    const float2 av = *reinterpret_cast<const float2*>(gr.dv2_add + row * gr.ldadd + h * D + nt * 8 + 2 * t);
                    k0 += ak.x; k1 += ak.y; v0 += av.x; v1 += av.y;
                }
                *reinterpret_cast<uint32_t*>(gr.dk2 + row * gr.ldd + h * D + nt * 8 + 2 * t) = pack_bf16x2(k0, k1);
                *reinterpret_cast<uint32_t*>(gr.dv2 + row * gr.ldd + h * D + nt * 8 + 2 * t) = pack_bf16x2(v0, v1);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ forward, tiled
// attn_fwd_mma for long sequences (Lq > 128): warp w of CTA z owns 16-row block w * Z + z and the keys stream through two 64-key
// stages (as attn_bwd_rows_tiled) instead of being resident (418 keys = 129 KB = one CTA per SM).
template <int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
attn_fwd_tiled(const AttnFwd a, float* __restrict__ out32) {
    extern __shared__ __align__(16) uint8_t smem[];
    pdl_wait();
    pdl_launch_dependents();
    const int h = blockIdx.x, b = blockIdx.y, z = blockIdx.z, Z = gridDim.z;
    const int n1 = a.n1, n = a.n1 + a.n2, n_pad = (n + 63) & ~63;
    bf16* KV = reinterpret_cast<bf16*>(smem);
    float* madd = reinterpret_cast<float*>(KV + (size_t)4 * TC * LDS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    stage_kv_chunk_async(a, b, h, 0, KV, KV + (size_t)TC * LDS, tid, WARPS * 32);
    cp_async_commit();
    for (int j = tid; j < n_pad; j += WARPS * 32) {
        float m = 0.f;
        if (j >= n) m = NEG_BIG;
        else if (a.key_mask && j >= n1) m = (1.0f - (float)a.key_mask[(long)b * a.n2 + (j - n1)]) * -10000.0f;
        madd[j] = m;
    }
    const Adapt ad = load_adapt(a, b);
    const int r0 = (warp * Z + z) * 16;
    const bool active = r0 < a.Lq;
    uint32_t qa[4][4];
    if (active) load_a_bf16(a.q + (long)b * a.Lq * a.ldq + h * D, a.ldq, r0, a.Lq, lane, qa);
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float mrun[2] = {NEG_BIG, NEG_BIG}, lrun[2] = {0.f, 0.f};
    for (int kc = 0, it = 0; kc < n_pad; kc += TC, ++it) {
        const bf16* Ks = KV + (size_t)(it & 1) * 2 * TC * LDS;
        const bf16* Vs = Ks + (size_t)TC * LDS;
        if (kc + TC < n_pad) {
            bf16* Kn = KV + (size_t)((it + 1) & 1) * 2 * TC * LDS;
            stage_kv_chunk_async(a, b, h, kc + TC, Kn, Kn + (size_t)TC * LDS, tid, WARPS * 32);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();                         // chunk `it` has landed (first trip: madd is complete)
        if (active) {
            float s[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
            mma_a_bt(s, qa, Ks, 0, lane);
            float mx[2] = {NEG_BIG, NEG_BIG};
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int col = kc + nt * 8 + 2 * t + (e & 1);
                    const int row = r0 + g + 8 * (e >> 1);
                    const float v = s[nt][e] * a.scale * afactor(ad, row, col - n1) + madd[col];
                    s[nt][e] = v;
                    mx[e >> 1] = fmaxf(mx[e >> 1], v);
                }
            }
            float alpha[2];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const float mnew = fmaxf(mrun[hh], quad_max(mx[hh]));
                alpha[hh] = __expf(mrun[hh] - mnew);
                mrun[hh] = mnew;
                lrun[hh] *= alpha[hh];
            }
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float p = __expf(s[nt][e] - mrun[e >> 1]);
                    lrun[e >> 1] += p;
                    if (a.drop.thresh) {   // dropout on the normalised probabilities: the row sum stays undropped
                        const int col = kc + nt * 8 + 2 * t + (e & 1);
                        const int row = r0 + g + 8 * (e >> 1);
                        p *= drop_mult(a.drop, (uint32_t)((((long)b * a.nh + h) * a.Lq + row) * n + col));
                    }
                    s[nt][e] = p;
                }
                o[nt][0] *= alpha[0]; o[nt][1] *= alpha[0]; o[nt][2] *= alpha[1]; o[nt][3] *= alpha[1];
            }
            uint32_t pa[4][4];
            acc_to_a(s, pa);
            mma_p_b(o, pa, Vs, 0, lane);
        }
        __syncthreads();                         // stage it & 1 may be refilled
    }
    if (!active) return;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int row = r0 + g + 8 * hh;
        const float l = quad_sum(lrun[hh]);
        const float inv = 1.0f / l;
        if (row < a.Lq) {
            if (t == 0 && a.lse) a.lse[((long)b * a.nh + h) * a.Lq + row] = mrun[hh] + __logf(l);
            const long off = ((long)b * a.Lq + row) * a.ldo + h * D;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float x0 = o[nt][2 * hh] * inv, x1 = o[nt][2 * hh + 1] * inv;
                *reinterpret_cast<uint32_t*>(a.out + off + nt * 8 + 2 * t) = pack_bf16x2(x0, x1);
                if (out32) *reinterpret_cast<float2*>(out32 + off + nt * 8 + 2 * t) = make_float2(x0, x1);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ backward, one kernel
// Both passes in ONE CTA per (head, sample) when P and dS of the whole (head, sample) fit in shared memory (bf16,
// [Lq][n]): phase A = the rows pass above (warp per 16 query rows: S, dP, P, dS, dQ, dW0/dW1) which also leaves P (with
// its dropout mask) and dS f in shared memory; phase B = the keys pass (warp per 16 keys) reading them back TRANSPOSED
// with ldmatrix.trans as the A operand of dV = P^T dO and dK = (dS f)^T Q, instead of recomputing S^T and dP^T from
// global memory.  5 matrix products instead of 7, one launch, no delta round trip, Q / dO staged once over the K / V
// buffers phase A is done with.  (The two-kernel path stays for shapes whose P does not fit: MKGformer-large vision.)
template <int NW, int NTK>
__global__ void __launch_bounds__(NW * 32, NW == 4 ? 3 : 1)     // text tiles: 3 CTAs per SM hold all B * heads CTAs in one wave
attn_bwd_fused(const AttnBwd gr, const bf16* __restrict__ ctx, long ldctx) {
    extern __shared__ __align__(16) uint8_t smem[];
    pdl_wait();
    pdl_launch_dependents();
    const AttnFwd& a = gr.f;
    const int h = blockIdx.x, b = blockIdx.y;
    const int n1 = a.n1, n2 = a.n2, n = n1 + n2, n_pad = (n + 63) & ~63, Lq = a.Lq, Lq16 = (Lq + 15) & ~15;
    const int LDP = n_pad + 8;                      // row of P / dS (bf16 elements): 16 B past a multiple of 128 B
    bf16* Ks = reinterpret_cast<bf16*>(smem);       // [n_pad][LDS]
    bf16* Vs = Ks + (size_t)n_pad * LDS;
    bf16* Qs = Vs + (size_t)n_pad * LDS;            // [Lq16][LDS]
    bf16* Os = Qs + (size_t)Lq16 * LDS;             // [Lq16][LDS]  dO (bf16)
    bf16* Ps = Os + (size_t)Lq16 * LDS;             // [Lq16][LDP]  P (dropped)
    bf16* Ds = Ps + (size_t)Lq16 * LDP;             // [Lq16][LDP]  dS * f
    float* madd = reinterpret_cast<float*>(Ds + (size_t)Lq16 * LDP);   // [n_pad]
    float* dl_s = madd + n_pad;                     // [Lq16] delta = rowsum(dO * O)
    __shared__ float wred[2][NW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const bf16* qbase = a.q + (long)b * Lq * a.ldq + h * D;
    const float* dobase = gr.dout + (long)b * Lq * gr.lddo + h * D;
    const bf16* obase = ctx + (long)b * Lq * ldctx + h * D;
    {   // every operand of the (head, sample) goes to shared memory once; all requests in flight together
        if (n1 > 0) {
            stage_bf16_async(Ks, a.k1 + (long)b * n1 * a.ld1 + h * D, a.ld1, n1, n1, tid, NW * 32);
            stage_bf16_async(Vs, a.v1 + (long)b * n1 * a.ld1 + h * D, a.ld1, n1, n1, tid, NW * 32);
        }
        stage_bf16_async(Ks + n1 * LDS, a.k2 + (long)b * n2 * a.ld2 + h * D, a.ld2, n2, n_pad - n1, tid, NW * 32);
        stage_bf16_async(Vs + n1 * LDS, a.v2 + (long)b * n2 * a.ld2 + h * D, a.ld2, n2, n_pad - n1, tid, NW * 32);
        stage_bf16_async(Qs, qbase, a.ldq, Lq, Lq16, tid, NW * 32);
        stage_do_delta(Os, dl_s, dobase, gr.lddo, obase, ldctx, Lq, Lq16, tid, NW * 32);
        for (int j = tid; j < n_pad; j += NW * 32) {
            float m = 0.f;
            if (j >= n) m = NEG_BIG;
            else if (a.key_mask && j >= n1) m = (1.0f - (float)a.key_mask[(long)b * n2 + (j - n1)]) * -10000.0f;
            madd[j] = m;
        }
        cp_async_wait_all();
    }
    const Adapt ad = load_adapt(a, b);
    __syncthreads();

    // ---------------- phase A: query-row blocks, keys in chunks of KC = 8 NTK
    constexpr int KC = 8 * NTK;
    float dw0 = 0.f, dw1 = 0.f;
#pragma unroll 1
    for (int r0 = warp * 16; r0 < Lq16; r0 += NW * 16) {
        uint32_t qa[4][4], doa[4][4];
        load_a_smem(Qs, r0, lane, qa);
        load_a_smem(Os, r0, lane, doa);
        float dl[2], lse[2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int row = r0 + g + 8 * hh;
            dl[hh] = dl_s[row];
            lse[hh] = (row < Lq) ? a.lse[((long)b * a.nh + h) * Lq + row] : 0.f;
        }
        float dq[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
        // Adaptive scaling: dW0 / dW1 = sum dS * S_raw is a heavily cancelling sum (sum_j dS_ij = 0 in exact arithmetic), so
        // delta must be CONSISTENT with the dP this kernel computes (bf16 dO, bf16 V): delta_i = sum_j P_ij dP_ij from the same
        // products, not rowsum(dO * O) with its independently rounded O - otherwise (delta error) x (row mean of S_raw) leaks
        // into the two scalars.  A pre-pass over the key chunks accumulates it.
        const bool adaptive = ad.split >= 0;
        if (adaptive) {
            float acc[2] = {0.f, 0.f};
#pragma unroll 1
            for (int kc = 0; kc < n_pad; kc += KC) {
                float s[NTK][4], dp[NTK][4];
#pragma unroll
                for (int i = 0; i < NTK; ++i) {
                    s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
                    dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
                }
                mma_a_bt_t<NTK>(s, qa, Ks, kc, lane);
                mma_a_bt_t<NTK>(dp, doa, Vs, kc, lane);
                consistent_delta<NTK>(a, ad, s, dp, madd, lse, kc, r0, g, t, n1, n, Lq, b, h, acc);
            }
            dl[0] = quad_sum(acc[0]); dl[1] = quad_sum(acc[1]);
        }
#pragma unroll 1
        for (int kc = 0; kc < n_pad; kc += KC) {
            float s[NTK][4], dp[NTK][4];
#pragma unroll
            for (int i = 0; i < NTK; ++i) {
                s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
                dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
            }
            mma_a_bt_t<NTK>(s, qa, Ks, kc, lane);
            mma_a_bt_t<NTK>(dp, doa, Vs, kc, lane);
#pragma unroll
            for (int nt = 0; nt < NTK; ++nt) {
                float pd[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int col = kc + nt * 8 + 2 * t + (e & 1);
                    const int row = r0 + g + 8 * (e >> 1);
                    const int j2 = col - n1;
                    const float f = afactor(ad, row, j2);
                    const float sraw = s[nt][e] * a.scale;
                    const float p = (row < Lq) ? __expf(sraw * f + madd[col] - lse[e >> 1]) : 0.f;
                    float mult = 1.f;
                    if (a.drop.thresh) mult = drop_mult(a.drop, (uint32_t)((((long)b * a.nh + h) * Lq + row) * n + col));
                    const float ds = p * (dp[nt][e] * mult - dl[e >> 1]);
                    if (ad.split >= 0 && j2 >= ad.split && col < n) {
                        if (row >= ad.split) dw1 = fmaf(ds, sraw, dw1);
                        else if (row >= ad.row_lo) dw0 = fmaf(ds, sraw, dw0);
                    }
                    pd[e] = p * mult;
                    s[nt][e] = ds * f;
                }
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int off = (r0 + g + 8 * hh) * LDP + kc + nt * 8 + 2 * t;
                    *reinterpret_cast<uint32_t*>(Ps + off) = pack_bf16x2(pd[2 * hh], pd[2 * hh + 1]);
                    *reinterpret_cast<uint32_t*>(Ds + off) = pack_bf16x2(s[nt][2 * hh], s[nt][2 * hh + 1]);
                }
            }
            uint32_t da[NTK / 2][4];
            acc_to_a_t<NTK>(s, da);
            mma_p_b_t<NTK>(dq, da, Ks, kc, lane);
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int row = r0 + g + 8 * hh;
            if (row < Lq) {
                const long off = ((long)b * Lq + row) * gr.ldd + h * D;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
                    *reinterpret_cast<uint32_t*>(gr.dq + off + nt * 8 + 2 * t) =
                        pack_bf16x2(dq[nt][2 * hh] * a.scale, dq[nt][2 * hh + 1] * a.scale);
            }
        }
    }
    __syncthreads();            // P / dS complete

    // ---------------- phase B: key blocks
    const int lr = lane & 7, sel = lane >> 3;
    const uint32_t ps_u = smem_u32(Ps), ds_u = smem_u32(Ds), qs_u = smem_u32(Qs), os_u = smem_u32(Os);
#pragma unroll 1
    for (int j0 = warp * 16; j0 < n; j0 += NW * 16) {
        float dk[8][4], dv[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
            dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
        }
#pragma unroll 1
        for (int r16 = 0; r16 < Lq16; r16 += 16) {
            // A = P^T / (dS f)^T tile (16 keys x 16 query rows): transposed 8x8 blocks of the row-major [row][key] buffers
            uint32_t pa[4], da[4];
            const uint32_t aoff = (uint32_t)((r16 + lr + 8 * (sel >> 1)) * LDP + j0 + 8 * (sel & 1)) * 2u;
            ldsm_x4_t(ps_u + aoff, pa);
            ldsm_x4_t(ds_u + aoff, da);
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                const uint32_t boff = (uint32_t)((r16 + (sel & 1) * 8 + lr) * LDS + np * 16 + (sel >> 1) * 8) * 2u;
                uint32_t bo[4], bq[4];
                ldsm_x4_t(os_u + boff, bo);
                ldsm_x4_t(qs_u + boff, bq);
                mma16816(dv[2 * np], pa, bo[0], bo[1]);
                mma16816(dv[2 * np + 1], pa, bo[2], bo[3]);
                mma16816(dk[2 * np], da, bq[0], bq[1]);
                mma16816(dk[2 * np + 1], da, bq[2], bq[3]);
            }
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int j = j0 + g + 8 * hh;
            if (j >= n) continue;
            if (j < n1) {
                const long off = ((long)b * n1 + j) * gr.ldd1 + h * D;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    *reinterpret_cast<float2*>(gr.dk1 + off + nt * 8 + 2 * t) = make_float2(dk[nt][2 * hh] * a.scale, dk[nt][2 * hh + 1] * a.scale);
                    *reinterpret_cast<float2*>(gr.dv1 + off + nt * 8 + 2 * t) = make_float2(dv[nt][2 * hh], dv[nt][2 * hh + 1]);
                }
            } else {
                const long row = (long)b * n2 + (j - n1);
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    float k0 = dk[nt][2 * hh] * a.scale, k1 = dk[nt][2 * hh + 1] * a.scale;
                    float v0 = dv[nt][2 * hh], v1 = dv[nt][2 * hh + 1];
                    if (gr.dk2_add) {
                        const float2 ak = *reinterpret_cast<const float2*>(gr.dk2_add + row * gr.ldadd + h * D + nt * 8 + 2 * t);
                        const float2 av = *reinterpret_cast<const float2*>(gr.dv2_add + row * gr.ldadd + h * D + nt * 8 + 2 * t);
                        k0 += ak.x; k1 += ak.y; v0 += av.x; v1 += av.y;
                    }
                    *reinterpret_cast<uint32_t*>(gr.dk2 + row * gr.ldd + h * D + nt * 8 + 2 * t) = pack_bf16x2(k0, k1);
                    *reinterpret_cast<uint32_t*>(gr.dv2 + row * gr.ldd + h * D + nt * 8 + 2 * t) = pack_bf16x2(v0, v1);
                }
            }
        }
    }
    if (a.sep_idx) {
        dw0 = warp_sum(dw0);
        dw1 = warp_sum(dw1);
        if (lane == 0) { wred[0][warp] = dw0; wred[1][warp] = dw1; }
        __syncthreads();
        if (tid == 0) {
            float s0 = 0.f, s1 = 0.f;
            for (int w = 0; w < NW; ++w) { s0 += wred[0][w]; s1 += wred[1][w]; }
            const float w0 = *a.w0, w1 = *a.w1;     // torch.clamp passes the gradient at both (inclusive) bounds
            if (w0 >= 0.f && w0 <= 0.5f) atomicAdd(gr.dw0, s0);
            if (w1 >= 0.5f && w1 <= 1.0f) atomicAdd(gr.dw1, s1);
        }
    }
}

int check_attn(const AttnFwd& a) {
    B2_CHECK(a.B > 0 && a.Lq > 0 && a.nh > 0 && a.n2 > 0 && a.n1 >= 0, "attention: bad sizes B=%d Lq=%d nh=%d n1=%d n2=%d", a.B,
             a.Lq, a.nh, a.n1, a.n2);
    B2_CHECK(a.ldq % 8 == 0 && a.ld2 % 8 == 0 && (a.n1 == 0 || a.ld1 % 8 == 0) && a.ldo % 2 == 0,
             "attention: leading dimensions must be multiples of 8 elements");
    B2_CHECK(a.sep_idx == nullptr || (a.w0 && a.w1 && a.n1 == 0), "attention: adaptive scaling needs w0/w1 and no key prefix");
    return 0;
}

// B200MART_ATTN_OCC4: the tiled forward / rows kernels compiled for 4 CTAs per SM (128 registers, a few spilled words) instead of 3
int tiled_occ4() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("B200MART_ATTN_OCC4"); v = e ? atoi(e) : 0; }
    return v;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    B2_CHECK(bytes <= 227 * 1024, "attention: needs %zu bytes of shared memory (> 227 KB)", bytes);
    CUDA_TRY(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

}  // namespace

int launch_attention_fwd(const AttnFwd& a, float* out32, cudaStream_t s) {
    B2_TRY(check_attn(a));
    B2_CHECK(a.out, "attention: output missing");
    const int n_pad = (a.n1 + a.n2 + 63) & ~63;
    const size_t smem = (size_t)2 * n_pad * LDS * 2 + (size_t)n_pad * 4;
    if (a.Lq > 64) {
        static int ftiled = -1;    // B200MART_ATTN_FWD_TILED=0: the resident kernel for every shape
        if (ftiled < 0) { const char* e = getenv("B200MART_ATTN_FWD_TILED"); ftiled = e ? atoi(e) : 1; }
        if (ftiled && a.Lq > 128) {     // more than one round of 8 x 16 rows: rows split over grid.z, keys streamed in 64-key chunks
            const int zf = ((a.Lq + 15) / 16 + WARPS - 1) / WARPS;
            const size_t sm = (size_t)4 * TC * LDS * 2 + (size_t)n_pad * 4;
            if (tiled_occ4()) {
                B2_TRY(set_smem(attn_fwd_tiled<4>, sm));
                CUDA_TRY(launch_pdl(attn_fwd_tiled<4>, dim3(a.nh, a.B, zf), dim3(WARPS * 32), sm, s, a, out32));
            } else {
                B2_TRY(set_smem(attn_fwd_tiled<3>, sm));
                CUDA_TRY(launch_pdl(attn_fwd_tiled<3>, dim3(a.nh, a.B, zf), dim3(WARPS * 32), sm, s, a, out32));
            }
            B2_LAUNCHED();
            return 0;
        }
        B2_TRY(set_smem(attn_fwd_mma<8>, smem));
        CUDA_TRY(launch_pdl(attn_fwd_mma<8>, dim3(a.nh, a.B), dim3(8 * 32), smem, s, a, out32));
    } else {
        B2_TRY(set_smem(attn_fwd_mma<4>, smem));
        CUDA_TRY(launch_pdl(attn_fwd_mma<4>, dim3(a.nh, a.B), dim3(4 * 32), smem, s, a, out32));
    }
    B2_LAUNCHED();
    return 0;
}

int launch_attention_bwd(const AttnBwd& g, const bf16* ctx, long ldctx, float* delta, cudaStream_t s) {
    const AttnFwd& a = g.f;
    B2_TRY(check_attn(a));
    B2_CHECK(g.dout && g.dq && g.dk2 && g.dv2 && delta && a.lse && ctx, "attention bwd: missing buffers");
    B2_CHECK(a.n1 == 0 || (g.dk1 && g.dv1), "attention bwd: prefix gradients need dk1/dv1");
    B2_CHECK(g.lddo % 4 == 0 && g.ldd % 2 == 0 && ldctx % 2 == 0, "attention bwd: bad leading dimensions");
    const int n_pad = (a.n1 + a.n2 + 63) & ~63;
    {   // one-kernel path when P and dS of a (head, sample) fit in shared memory next to the operands
        static int allow = -1;
        if (allow < 0) { const char* e = getenv("B200MART_ATTN_FUSED_BWD"); allow = e ? atoi(e) : 1; }
        const int lq16 = (a.Lq + 15) & ~15;
        const size_t smemf = (size_t)2 * (n_pad + lq16) * LDS * 2 + (size_t)2 * lq16 * (n_pad + 8) * 2 + (size_t)(n_pad + lq16) * 4;
        if (allow && smemf <= 200 * 1024) {
            static int ntk = -1;       // B200MART_ATTN_KC = keys per chunk of the rows pass: 16 / 32 / 64 (default 32)
            if (ntk < 0) { const char* e = getenv("B200MART_ATTN_KC"); const int kc = e ? atoi(e) : 32; ntk = kc == 16 ? 2 : (kc == 64 ? 8 : 4); }
            auto go = [&](auto kernel, int nw) -> int {
                B2_TRY(set_smem(kernel, smemf));
                CUDA_TRY(launch_pdl(kernel, dim3(a.nh, a.B), dim3(nw * 32), smemf, s, g, ctx, ldctx));
                return 0;
            };
            if (a.Lq <= 64 && a.n1 + a.n2 <= 64) {
                B2_TRY(ntk == 2 ? go(attn_bwd_fused<4, 2>, 4) : (ntk == 8 ? go(attn_bwd_fused<4, 8>, 4) : go(attn_bwd_fused<4, 4>, 4)));
            } else {
                B2_TRY(ntk == 2 ? go(attn_bwd_fused<8, 2>, 8) : (ntk == 8 ? go(attn_bwd_fused<8, 8>, 8) : go(attn_bwd_fused<8, 4>, 8)));
            }
            B2_LAUNCHED();
            return 0;
        }
    }
    static int tiled = -1;     // B200MART_ATTN_TILED=0: the resident two-pass kernels (one CTA per (head, sample)) instead of the tiled ones
    if (tiled < 0) { const char* e = getenv("B200MART_ATTN_TILED"); tiled = e ? atoi(e) : 1; }
    if (tiled) {
        const int lq_pad = (a.Lq + 63) & ~63;
        const int zr = ((a.Lq + 15) / 16 + WARPS - 1) / WARPS, zk = ((a.n1 + a.n2 + 15) / 16 + WARPS - 1) / WARPS;
        const size_t sm_r = (size_t)4 * TC * LDS * 2 + (size_t)n_pad * 4, sm_k = (size_t)4 * TC * LDS * 2 + (size_t)2 * lq_pad * 4;
        B2_CHECK(!g.dout16 || g.lddo16 % 8 == 0, "attention bwd: the bf16 copy of dout needs a leading dimension that is a multiple of 8");
        if (tiled_occ4()) {
            B2_TRY(set_smem(attn_bwd_rows_tiled<4, 4>, sm_r));
            attn_bwd_rows_tiled<4, 4><<<dim3(a.nh, a.B, zr), WARPS * 32, sm_r, s>>>(g, ctx, ldctx, delta);
        } else {
            B2_TRY(set_smem(attn_bwd_rows_tiled<4, 3>, sm_r));
            attn_bwd_rows_tiled<4, 3><<<dim3(a.nh, a.B, zr), WARPS * 32, sm_r, s>>>(g, ctx, ldctx, delta);
        }
        B2_LAUNCHED();
        static int kv = -1;    // B200MART_ATTN_KEYS_VARIANT: 0 = 32-row inner chunks, 2 CTAs / SM (222 registers); 1 (default) = 16-row chunks, 3 CTAs / SM; 2 = 32-row, 3 CTAs
        if (kv < 0) { const char* e = getenv("B200MART_ATTN_KEYS_VARIANT"); kv = e ? atoi(e) : 1; }
        if (kv == 1) {
            B2_TRY(set_smem(attn_bwd_keys_tiled<2, 3>, sm_k));
            attn_bwd_keys_tiled<2, 3><<<dim3(a.nh, a.B, zk), WARPS * 32, sm_k, s>>>(g, delta);
        } else if (kv == 2) {
            B2_TRY(set_smem(attn_bwd_keys_tiled<4, 3>, sm_k));
            attn_bwd_keys_tiled<4, 3><<<dim3(a.nh, a.B, zk), WARPS * 32, sm_k, s>>>(g, delta);
        } else {
            B2_TRY(set_smem(attn_bwd_keys_tiled<4, 2>, sm_k));
            attn_bwd_keys_tiled<4, 2><<<dim3(a.nh, a.B, zk), WARPS * 32, sm_k, s>>>(g, delta);
        }
        B2_LAUNCHED();
        return 0;
    }
    const size_t smem1 = (size_t)2 * n_pad * LDS * 2 + (size_t)n_pad * 4;
    B2_TRY(set_smem(attn_bwd_rows_mma, smem1));
    attn_bwd_rows_mma<<<dim3(a.nh, a.B), WARPS * 32, smem1, s>>>(g, ctx, ldctx, delta);
    B2_LAUNCHED();
    const int lq_pad = (a.Lq + 63) & ~63;
    const size_t smem2 = (size_t)2 * lq_pad * LDS * 2 + (size_t)2 * lq_pad * 4;
    B2_TRY(set_smem(attn_bwd_keys_mma, smem2));
    attn_bwd_keys_mma<<<dim3(a.nh, a.B), WARPS * 32, smem2, s>>>(g, delta);
    B2_LAUNCHED();
    return 0;
}
