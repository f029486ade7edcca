// BertFusion (MarT/models/modeling_unimo.py:400-414): out = softmax(ctx vh^T) vh over the vision tokens, no scale,
// no mask; forward + two-pass backward on the tensor cores at fp32-grade precision.
//
// The scores are raw H-dim dot products (O(sqrt(H)) magnitude, near one-hot softmax): bf16 operand rounding would move
// probabilities by several percent (SURVEY.md §7 hard part 4), so every product here is a 3xTF32 split product
// (x = hi + lo, acc += lo*hi' + hi*lo' + hi*hi', warp-level mma.sync.m16n8k8 with fp32 accumulation): ~2^-21 relative
// operand error.  The op is 0.3 % of the step's FLOPs but sits on the serial part of the step (the text chain waits
// for the vision chain at every fused layer), so what matters is latency: 16 query rows (or 16 keys) per CTA, the K
// dimension split over the 8 warps of the CTA, operands read straight from L2 with 16-byte loads (the k / n index of a
// fragment is permuted so that each lane's elements are contiguous in memory).  (Self-attention: attention_mma.cu.)
#include "kernels.cuh"

namespace {

constexpr int FROWS = 16;     // query rows (forward, backward rows pass) or keys (backward keys pass) per CTA = one m16 tile
constexpr int FTHREADS = 256;
constexpr int FWARPS = FTHREADS / 32;
constexpr int NTG = 16;       // n-tiles (8 keys each) whose score accumulators a warp holds at once
constexpr int TB = 4;         // n-tiles per load / multiply batch
constexpr int GPW = 3;        // 32-column groups of the H dimension a warp accumulates at once (H = 768: all of them)

// x = hi + lo with hi = x truncated to TF32 (a full-rate bit mask; cvt.rna.tf32 issues at 1/8 rate and would dominate the
// kernel) and lo = x - hi (exact in fp32; the tensor core ignores its low 13 bits): ~2^-21 relative error per operand
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// acc[j] += A * B_j for NB independent accumulators at fp32-grade precision: x = hi + lo per operand, small cross terms
// first.  Term-major order: consecutive mma.sync never share an accumulator (dependent issue costs 21 cycles, independent 8).
template <int NB>
__device__ __forceinline__ void mma_3x_batch(float (*c[NB])[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4],
                                             const float (&b0)[NB], const float (&b1)[NB]) {
    uint32_t bh0[NB], bl0[NB], bh1[NB], bl1[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        split_tf32(b0[j], bh0[j], bl0[j]);
        split_tf32(b1[j], bh1[j], bl1[j]);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) mma_tf32(*c[j], alo, bh0[j], bh1[j]);
#pragma unroll
    for (int j = 0; j < NB; ++j) mma_tf32(*c[j], ahi, bl0[j], bl1[j]);
#pragma unroll
    for (int j = 0; j < NB; ++j) mma_tf32(*c[j], ahi, bh0[j], bh1[j]);
}

// shared-memory row strides (fp32 words) that keep the A-fragment reads conflict-free
__host__ __device__ inline int x_ld(int H) { return H + 2; }                         // 64-bit reads: ld / 2 = 1 (mod 16)
__host__ __device__ inline int s_ld(int n) { return ((n + 31) / 32) * 32 + 4; }      // 32-bit reads: ld = 4 (mod 32)

// FROWS rows of a [*, H] fp32 matrix -> xs (rows >= nrows are zero)
__device__ __forceinline__ void stage_rows(float* xs, int ldx, const float* __restrict__ src, int nrows, int H, int tid) {
    const int q = H >> 2;
    for (int i = tid; i < FROWS * q; i += FTHREADS) {
        const int r = i / q, c = (i - r * q) << 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nrows) v = __ldg(reinterpret_cast<const float4*>(src + (long)r * H + c));
        float* d = xs + r * ldx + c;
        *reinterpret_cast<float2*>(d) = make_float2(v.x, v.y);
        *reinterpret_cast<float2*>(d + 2) = make_float2(v.z, v.w);
    }
}

// S[r][j] = sum_k xs[r][k] * vb[j][k]   (r < 16, j < Nv; columns up to the next multiple of 8 hold the clamped last key:
// callers mask them).  Split-K: warp w owns the 32-wide k chunks w, w + 8, ...; every warp writes its partial sums of a
// group of NTG n-tiles (128 keys) to its own slab, and the CTA then adds the 8 slabs into S (shared-memory atomics cost
// ~64 cycles per warp instruction: more than the products themselves).  Contains __syncthreads: call from all threads.
// Fragment index permutation: within a chunk, lane t of a quad covers k = 8 t .. 8 t + 7 (mma step s uses 8 t + 2 s and
// 8 t + 2 s + 1 as the hardware's k = t and k = t + 4), so B comes in as two 16-byte loads and A as 8-byte smem reads.
constexpr int SLAB_LD = NTG * 8 + 8;   // 64-bit stores of (row g, column pair) are conflict-free when ld = 8 (mod 32)
__device__ __forceinline__ void scores_splitk(const float* xs, int ldx, const float* __restrict__ vb, int Nv, int H, float* S,
                                              int ldS, float* slab, int tid) {
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int NT = (Nv + 7) >> 3;
    float* my = slab + warp * (FROWS * SLAB_LD);
    for (int nt0 = 0; nt0 < NT; nt0 += NTG) {
        float acc[NTG][4];
#pragma unroll
        for (int i = 0; i < NTG; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
        for (int kc = warp * 32; kc < H; kc += FWARPS * 32) {
            uint32_t ahi[4][4], alo[4][4];
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const float2 x0 = *reinterpret_cast<const float2*>(xs + g * ldx + kc + 8 * t + 2 * s);
                const float2 x1 = *reinterpret_cast<const float2*>(xs + (g + 8) * ldx + kc + 8 * t + 2 * s);
                split_tf32(x0.x, ahi[s][0], alo[s][0]);
                split_tf32(x1.x, ahi[s][1], alo[s][1]);
                split_tf32(x0.y, ahi[s][2], alo[s][2]);
                split_tf32(x1.y, ahi[s][3], alo[s][3]);
            }
            const float* bcol = vb + kc + 8 * t;
            auto brow = [&](int nt) { const int j = nt * 8 + g; return bcol + (long)(j < Nv ? j : Nv - 1) * H; };
            // software pipeline over batches of TB n-tiles: the 2 TB 16-byte loads of the next batch are in flight while
            // this batch is multiplied (one CTA of 8 warps per SM: memory-level parallelism has to come from each warp)
            float4 nx[TB][2];
            auto load_batch = [&](int ntb) {
#pragma unroll
                for (int j = 0; j < TB; ++j) {
                    const float4* q = reinterpret_cast<const float4*>(brow(ntb + j < NT ? ntb + j : NT - 1));
                    nx[j][0] = __ldg(q);
                    nx[j][1] = __ldg(q + 1);
                }
            };
            load_batch(nt0);
#pragma unroll
            for (int bt = 0; bt < NTG / TB; ++bt) {
                if (nt0 + bt * TB < NT) {
                    float4 cu[TB][2];
#pragma unroll
                    for (int j = 0; j < TB; ++j) { cu[j][0] = nx[j][0]; cu[j][1] = nx[j][1]; }
                    if (bt + 1 < NTG / TB && nt0 + (bt + 1) * TB < NT) load_batch(nt0 + (bt + 1) * TB);
                    float (*cp[TB])[4];
#pragma unroll
                    for (int j = 0; j < TB; ++j) cp[j] = &acc[bt * TB + j];
                    float b0[TB], b1[TB];
#pragma unroll
                    for (int j = 0; j < TB; ++j) { b0[j] = cu[j][0].x; b1[j] = cu[j][0].y; }
                    mma_3x_batch<TB>(cp, ahi[0], alo[0], b0, b1);
#pragma unroll
                    for (int j = 0; j < TB; ++j) { b0[j] = cu[j][0].z; b1[j] = cu[j][0].w; }
                    mma_3x_batch<TB>(cp, ahi[1], alo[1], b0, b1);
#pragma unroll
                    for (int j = 0; j < TB; ++j) { b0[j] = cu[j][1].x; b1[j] = cu[j][1].y; }
                    mma_3x_batch<TB>(cp, ahi[2], alo[2], b0, b1);
#pragma unroll
                    for (int j = 0; j < TB; ++j) { b0[j] = cu[j][1].z; b1[j] = cu[j][1].w; }
                    mma_3x_batch<TB>(cp, ahi[3], alo[3], b0, b1);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NTG; ++i) {
            float* p = my + g * SLAB_LD + i * 8 + 2 * t;
            *reinterpret_cast<float2*>(p) = make_float2(acc[i][0], acc[i][1]);
            *reinterpret_cast<float2*>(p + 8 * SLAB_LD) = make_float2(acc[i][2], acc[i][3]);
        }
        __syncthreads();
        const int ncols = min(NTG, NT - nt0) * 8;
        for (int e = tid; e < FROWS * NTG * 8; e += FTHREADS) {
            const int r = e / (NTG * 8), c = e % (NTG * 8);
            if (c < ncols) {
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < FWARPS; ++w) v += slab[w * (FROWS * SLAB_LD) + r * SLAB_LD + c];
                S[r * ldS + nt0 * 8 + c] = v;
            }
        }
        __syncthreads();
    }
}

// acc[i] (16 x 32 columns at cb[i]) += As[16][Kp] * Bg[K][H]   for the GPW column groups of this warp (cb[i] < 0: none).
// As: shared, row stride ldA, zero beyond the valid K; Bg rows >= Kvalid are never read (clamped; they meet zeros of As).
// Column permutation: the hardware's n = g of n-tile jn is column cb + 4 g + jn, so each lane reads B as one 16-byte
// load per k row and ends up with 8 consecutive output columns (cb + 8 t .. + 7) of rows g and g + 8.
__device__ __forceinline__ void mma_cols(float (&acc)[GPW][4][4], const float* As, int ldA, int Kp, const float* __restrict__ Bg,
                                         int Kvalid, int H, const int (&cb)[GPW], int lane) {
    const int g = lane >> 2, t = lane & 3;
    float4 na[GPW], nb[GPW];
    auto load_step = [&](int k0) {
        const int ka = (k0 + t < Kvalid) ? k0 + t : Kvalid - 1;
        const int kb = (k0 + t + 4 < Kvalid) ? k0 + t + 4 : Kvalid - 1;
#pragma unroll
        for (int i = 0; i < GPW; ++i) {
            const int c = (cb[i] >= 0 ? cb[i] : 0) + 4 * g;
            na[i] = __ldg(reinterpret_cast<const float4*>(Bg + (long)ka * H + c));
            nb[i] = __ldg(reinterpret_cast<const float4*>(Bg + (long)kb * H + c));
        }
    };
    load_step(0);
    for (int k0 = 0; k0 < Kp; k0 += 8) {
        float4 va[GPW], vb[GPW];
#pragma unroll
        for (int i = 0; i < GPW; ++i) { va[i] = na[i]; vb[i] = nb[i]; }
        if (k0 + 8 < Kp) load_step(k0 + 8);      // next k step's rows are in flight during this step's products
        uint32_t ahi[4], alo[4];
        split_tf32(As[g * ldA + k0 + t], ahi[0], alo[0]);
        split_tf32(As[(g + 8) * ldA + k0 + t], ahi[1], alo[1]);
        split_tf32(As[g * ldA + k0 + t + 4], ahi[2], alo[2]);
        split_tf32(As[(g + 8) * ldA + k0 + t + 4], ahi[3], alo[3]);
#pragma unroll
        for (int i = 0; i < GPW; ++i) {
            if (cb[i] >= 0) {
                float (*cp[4])[4] = {&acc[i][0], &acc[i][1], &acc[i][2], &acc[i][3]};
                const float b0[4] = {va[i].x, va[i].y, va[i].z, va[i].w};
                const float b1[4] = {vb[i].x, vb[i].y, vb[i].z, vb[i].w};
                mma_3x_batch<4>(cp, ahi, alo, b0, b1);
            }
        }
    }
}
__device__ __forceinline__ void zero_acc(float (&acc)[GPW][4][4]) {
#pragma unroll
    for (int i = 0; i < GPW; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
}
// column groups of warp `warp` in pass `pass` (a pass covers FWARPS * GPW groups of 32 columns)
__device__ __forceinline__ void col_groups(int (&cb)[GPW], int pass, int warp, int H) {
#pragma unroll
    for (int i = 0; i < GPW; ++i) {
        const int q = (pass * GPW + i) * FWARPS + warp;
        cb[i] = (q * 32 < H) ? q * 32 : -1;
    }
}
// the 8 consecutive values a lane holds for row half hh (0: row g, 1: row g + 8) of group i: columns cb + 8 t .. + 7
__device__ __forceinline__ void lane_row(const float (&acc)[GPW][4][4], int i, int hh, float4& lo, float4& hi) {
    lo = make_float4(acc[i][0][2 * hh], acc[i][1][2 * hh], acc[i][2][2 * hh], acc[i][3][2 * hh]);
    hi = make_float4(acc[i][0][2 * hh + 1], acc[i][1][2 * hh + 1], acc[i][2][2 * hh + 1], acc[i][3][2 * hh + 1]);
}

// ------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(FTHREADS, 1)
fusion_fwd_kernel(const float* __restrict__ ctx, const float* __restrict__ vh, bf16* __restrict__ out, float* __restrict__ lse,
                  int L, int Nv, int H) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(16) float fs[];
    const int ldx = x_ld(H), Nvp = (Nv + 7) & ~7, ldS = s_ld(Nvp);
    float* xs = fs;                    // [FROWS][ldx]
    float* S = xs + FROWS * ldx;       // [FROWS][ldS]
    float* slab = S + FROWS * ldS;     // [FWARPS][FROWS][SLAB_LD]
    const int b = blockIdx.y, r0 = blockIdx.x * FROWS;
    const int nrows = min(FROWS, L - r0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const float* vb = vh + (long)b * Nv * H;
    stage_rows(xs, ldx, ctx + ((long)b * L + r0) * H, nrows, H, tid);
    __syncthreads();
    scores_splitk(xs, ldx, vb, Nv, H, S, ldS, slab, tid);
    for (int r = warp; r < FROWS; r += FWARPS) {   // softmax, warp per row
        float* Sr = S + r * ldS;
        float mx = -INFINITY;
        for (int j = lane; j < Nv; j += 32) mx = fmaxf(mx, Sr[j]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < Nv; j += 32) {
            const float e = __expf(Sr[j] - mx);
            Sr[j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        for (int j = lane; j < Nvp; j += 32) Sr[j] = (j < Nv) ? Sr[j] * inv : 0.f;
        if (lane == 0 && r < nrows) lse[(long)b * L + r0 + r] = mx + __logf(sum);
    }
    __syncthreads();
    for (int pass = 0; pass * FWARPS * GPW * 32 < H; ++pass) {
        int cb[GPW];
        col_groups(cb, pass, warp, H);
        float acc[GPW][4][4];
        zero_acc(acc);
        mma_cols(acc, S, ldS, Nvp, vb, Nv, H, cb, lane);
#pragma unroll
        for (int i = 0; i < GPW; ++i) {
            if (cb[i] < 0) continue;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int r = g + 8 * hh;
                if (r >= nrows) continue;
                float4 lo, hi;
                lane_row(acc, i, hh, lo, hi);
                uint4 q = make_uint4(pack_bf16x2(lo.x, lo.y), pack_bf16x2(lo.z, lo.w), pack_bf16x2(hi.x, hi.y), pack_bf16x2(hi.z, hi.w));
                *reinterpret_cast<uint4*>(out + ((long)b * L + r0 + r) * H + cb[i] + 8 * t) = q;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ backward, rows pass
// per 16 query rows: S (recomputed) and dP = dG vh^T, P = exp(S - lse), dS = P (dP - rowsum(P dP)); P, dS -> pbuf / dsbuf
// [B, L, Nv] for the keys pass; dctx += dS vh
__global__ void __launch_bounds__(FTHREADS, 1)
fusion_bwd_rows_kernel(const float* __restrict__ dg, const float* __restrict__ ctx, const float* __restrict__ vh,
                       const float* __restrict__ lse, float* __restrict__ pbuf, float* __restrict__ dsbuf,
                       float* __restrict__ dctx, int L, int Nv, int H) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(16) float fs[];
    const int ldx = x_ld(H), Nvp = (Nv + 7) & ~7, ldS = s_ld(Nvp);
    float* xs = fs;                    // [FROWS][ldx]  ctx rows, later dG rows
    float* S = xs + FROWS * ldx;       // [FROWS][ldS]  scores -> P
    float* dP = S + FROWS * ldS;       // [FROWS][ldS]  dP -> dS
    float* slab = dP + FROWS * ldS;    // [FWARPS][FROWS][SLAB_LD]
    const int b = blockIdx.y, r0 = blockIdx.x * FROWS;
    const int nrows = min(FROWS, L - r0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const float* vb = vh + (long)b * Nv * H;
    stage_rows(xs, ldx, ctx + ((long)b * L + r0) * H, nrows, H, tid);
    __syncthreads();
    scores_splitk(xs, ldx, vb, Nv, H, S, ldS, slab, tid);      // ends with a barrier: xs may be overwritten
    stage_rows(xs, ldx, dg + ((long)b * L + r0) * H, nrows, H, tid);
    __syncthreads();
    scores_splitk(xs, ldx, vb, Nv, H, dP, ldS, slab, tid);
    for (int r = warp; r < FROWS; r += FWARPS) {
        float* Sr = S + r * ldS;
        float* Dr = dP + r * ldS;
        const bool live = r < nrows;
        const float l = live ? lse[(long)b * L + r0 + r] : 0.f;
        float part = 0.f;
        for (int j = lane; j < Nv; j += 32) {
            const float p = live ? __expf(Sr[j] - l) : 0.f;
            Sr[j] = p;
            part = fmaf(p, Dr[j], part);
        }
        const float dl = warp_sum(part);
        for (int j = lane; j < Nvp; j += 32) {
            const float ds = (j < Nv) ? Sr[j] * (Dr[j] - dl) : 0.f;
            Dr[j] = ds;
            if (live && j < Nv) {
                pbuf[((long)b * L + r0 + r) * Nv + j] = Sr[j];
                dsbuf[((long)b * L + r0 + r) * Nv + j] = ds;
            }
        }
    }
    __syncthreads();
    for (int pass = 0; pass * FWARPS * GPW * 32 < H; ++pass) {
        int cb[GPW];
        col_groups(cb, pass, warp, H);
        float acc[GPW][4][4];
        zero_acc(acc);
        mma_cols(acc, dP, ldS, Nvp, vb, Nv, H, cb, lane);
#pragma unroll
        for (int i = 0; i < GPW; ++i) {
            if (cb[i] < 0) continue;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int r = g + 8 * hh;
                if (r >= nrows) continue;
                float4 lo, hi;
                lane_row(acc, i, hh, lo, hi);
                float4* d = reinterpret_cast<float4*>(dctx + ((long)b * L + r0 + r) * H + cb[i] + 8 * t);
                const float4 o0 = d[0], o1 = d[1];
                d[0] = make_float4(o0.x + lo.x, o0.y + lo.y, o0.z + lo.z, o0.w + lo.w);
                d[1] = make_float4(o1.x + hi.x, o1.y + hi.y, o1.z + hi.z, o1.w + hi.w);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ backward, keys pass
// per 16 keys: dvh[j] = dvh_in[j] + sum_r dS[r][j] ctx[r] + P[r][j] dG[r]
__global__ void __launch_bounds__(FTHREADS, 1)
fusion_bwd_keys_kernel(const float* __restrict__ dg, const float* __restrict__ ctx, const float* __restrict__ pbuf,
                       const float* __restrict__ dsbuf, const float* __restrict__ dvh_in, float* __restrict__ dvh32,
                       bf16* __restrict__ dvh16, int L, int Nv, int H) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(16) float fs[];
    const int Lp = (L + 7) & ~7, ldA = s_ld(Lp);
    float* Pk = fs;                     // [FROWS keys][ldA]  P^T
    float* Dk = Pk + FROWS * ldA;       // [FROWS keys][ldA]  dS^T
    const int b = blockIdx.y, j0 = blockIdx.x * FROWS;
    const int nkeys = min(FROWS, Nv - j0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    for (int i = tid; i < Lp * FROWS; i += FTHREADS) {
        const int r = i / FROWS, k = i % FROWS;
        const bool ok = (k < nkeys) && (r < L);
        const long o = ((long)b * L + r) * Nv + j0 + k;
        Pk[k * ldA + r] = ok ? pbuf[o] : 0.f;
        Dk[k * ldA + r] = ok ? dsbuf[o] : 0.f;
    }
    __syncthreads();
    const float* xb = ctx + (long)b * L * H;
    const float* gb = dg + (long)b * L * H;
    for (int pass = 0; pass * FWARPS * GPW * 32 < H; ++pass) {
        int cb[GPW];
        col_groups(cb, pass, warp, H);
        float acc[GPW][4][4];
        zero_acc(acc);
        mma_cols(acc, Dk, ldA, Lp, xb, L, H, cb, lane);
        mma_cols(acc, Pk, ldA, Lp, gb, L, H, cb, lane);
#pragma unroll
        for (int i = 0; i < GPW; ++i) {
            if (cb[i] < 0) continue;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int k = g + 8 * hh;
                if (k >= nkeys) continue;
                float4 lo, hi;
                lane_row(acc, i, hh, lo, hi);
                const long o = ((long)b * Nv + j0 + k) * H + cb[i] + 8 * t;
                if (dvh_in) {
                    const float4 a0 = *reinterpret_cast<const float4*>(dvh_in + o), a1 = *reinterpret_cast<const float4*>(dvh_in + o + 4);
                    lo.x += a0.x; lo.y += a0.y; lo.z += a0.z; lo.w += a0.w;
                    hi.x += a1.x; hi.y += a1.y; hi.z += a1.z; hi.w += a1.w;
                }
                if (dvh32) {
                    *reinterpret_cast<float4*>(dvh32 + o) = lo;
                    *reinterpret_cast<float4*>(dvh32 + o + 4) = hi;
                }
                if (dvh16)
                    *reinterpret_cast<uint4*>(dvh16 + o) =
                        make_uint4(pack_bf16x2(lo.x, lo.y), pack_bf16x2(lo.z, lo.w), pack_bf16x2(hi.x, hi.y), pack_bf16x2(hi.z, hi.w));
            }
        }
    }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    B2_CHECK(bytes <= 227 * 1024, "fusion: needs %zu bytes of shared memory (> 227 KB)", bytes);
    CUDA_TRY(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

int check_fusion(const void* a, const void* b, const void* c, int B, int L, int Nv, int H) {
    B2_CHECK(B > 0 && L > 0 && Nv > 0 && H > 0 && H % 32 == 0, "fusion: bad sizes B=%d L=%d Nv=%d H=%d (H must be a multiple of 32)", B, L, Nv, H);
    B2_CHECK(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0,
             "fusion: operands must be 16-byte aligned");
    return 0;
}

}  // namespace

int launch_fusion_fwd(const float* ctx, const float* vh, bf16* out, float* lse, int B, int L, int Nv, int H, cudaStream_t s) {
    B2_TRY(check_fusion(ctx, vh, out, B, L, Nv, H));
    const size_t smem = (size_t)(FROWS * x_ld(H) + FROWS * s_ld((Nv + 7) & ~7) + FWARPS * FROWS * SLAB_LD) * 4;
    B2_TRY(set_smem(fusion_fwd_kernel, smem));
    CUDA_TRY(launch_pdl(fusion_fwd_kernel, dim3((L + FROWS - 1) / FROWS, B), dim3(FTHREADS), smem, s, ctx, vh, out, lse, L, Nv, H));
    B2_LAUNCHED();
    return 0;
}

int launch_fusion_bwd_rows(const float* dg, const float* ctx, const float* vh, const float* lse, float* pbuf, float* dsbuf,
                           float* dctx, int B, int L, int Nv, int H, cudaStream_t s) {
    B2_TRY(check_fusion(dg, ctx, vh, B, L, Nv, H));
    B2_CHECK((reinterpret_cast<uintptr_t>(dctx) & 15) == 0, "fusion: dctx must be 16-byte aligned");
    const size_t smem = (size_t)(FROWS * x_ld(H) + 2 * FROWS * s_ld((Nv + 7) & ~7) + FWARPS * FROWS * SLAB_LD) * 4;
    B2_TRY(set_smem(fusion_bwd_rows_kernel, smem));
    CUDA_TRY(launch_pdl(fusion_bwd_rows_kernel, dim3((L + FROWS - 1) / FROWS, B), dim3(FTHREADS), smem, s, dg, ctx, vh, lse, pbuf, dsbuf, dctx, L, Nv, H));
    B2_LAUNCHED();
    return 0;
}

int launch_fusion_bwd_keys(const float* dg, const float* ctx, const float* pbuf, const float* dsbuf, const float* dvh_in,
                           float* dvh32, bf16* dvh16, int B, int L, int Nv, int H, cudaStream_t s) {
    B2_TRY(check_fusion(dg, ctx, dvh_in, B, L, Nv, H));
    B2_CHECK(((reinterpret_cast<uintptr_t>(dvh32) | reinterpret_cast<uintptr_t>(dvh16)) & 15) == 0, "fusion: dvh outputs must be 16-byte aligned");
    const size_t smem = (size_t)2 * FROWS * s_ld((L + 7) & ~7) * 4;
    B2_TRY(set_smem(fusion_bwd_keys_kernel, smem));
    CUDA_TRY(launch_pdl(fusion_bwd_keys_kernel, dim3((Nv + FROWS - 1) / FROWS, B), dim3(FTHREADS), smem, s, dg, ctx, pbuf, dsbuf, dvh_in, dvh32, dvh16, L, Nv, H));
    B2_LAUNCHED();
    return 0;
}
