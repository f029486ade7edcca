// Fused single-pass attention kernels of the MKGformer step (head_dim = 64), forward and backward:
//   * text self-attention  — BertSelfAttention.forward, MarT/models/modeling_unimo.py:317-377:
//       S = Q K^T / 8; adaptive analogy scaling of the (rows<s | rows>=s) x (cols>=s) blocks by
//       clamp(w0,0,.5) / clamp(w1,.5,1) (:342-349) BEFORE the additive -1e4 padding mask (:351-355); softmax; P V.
//   * vision self-attention with the text K/V prefix — CLIPAttention.forward, :212-272: keys/values are
//       [text K/V of the previous text layer ; own], no mask (padded text keys ARE attended, :227-229).
//   * BertFusion, :400-414: softmax(ctx vh^T) vh, no scale, no mask, fp32 throughout (raw 768-dim dot products).
// One CTA per (head, sample); the whole K/V of that (head, sample) lives in shared memory, scores never
// touch HBM.  Backward is two deterministic passes (per query row: dQ, dW0/dW1; per key: dK, dV) that
// recompute P from the saved log-sum-exp.
//
// v1 arithmetic: fp32 FMA on CUDA cores from bf16 shared-memory tiles (rows padded to 33 words so that
// "lane = key" and "lane = column pair" accesses are both bank-conflict free).
#include "kernels.cuh"

namespace {

constexpr int D = 64;          // head dim
constexpr int KPAD = 66;       // padded bf16 row length (33 words)
constexpr int ATT_WARPS = 8;
constexpr int MAX_KEY_ITERS = 20;  // keys per lane: supports up to 640 keys (large config: 64 + 513 = 577)

__device__ __forceinline__ float2 bf2_to_f2(uint32_t w) {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}

// copy `rows` rows of 64 bf16 (global row stride ld elements) into padded shared rows
__device__ __forceinline__ void load_rows_padded(bf16* dst, const bf16* src, long ld, int rows, int tid, int nthreads) {
    // 8 x 16-byte chunks per row; padded destination rows are only 4-byte aligned -> store words
    for (int i = tid; i < rows * 8; i += nthreads) {
        const int r = i >> 3, c = i & 7;
        const uint4 q = *reinterpret_cast<const uint4*>(src + (long)r * ld + c * 8);
        uint32_t* d = reinterpret_cast<uint32_t*>(dst + r * KPAD + c * 8);
        d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
    }
}

struct RowCtx {
    int split;        // adaptive split column s (or -1)
    float w0c, w1c;   // clamped adaptive weights
};

// scale factor of score (r, j) for own-key index j2 (= j - n1); prefix keys are never scaled
__device__ __forceinline__ float adaptive_factor(const RowCtx& c, int r, int j2) {
    if (c.split < 0 || j2 < c.split) return 1.0f;
    return (r < c.split) ? c.w0c : c.w1c;
}

// ------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(ATT_WARPS * 32)
attention_fwd_kernel(const AttnFwd a, bf16* __restrict__ out, float* __restrict__ out32) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int h = blockIdx.x, b = blockIdx.y;
    const int n1 = a.n1, n2 = a.n2, n = n1 + n2;
    bf16* Ks = reinterpret_cast<bf16*>(smem);          // [n][KPAD]
    bf16* Vs = Ks + (size_t)n * KPAD;                   // [n][KPAD]
    float* Ps = reinterpret_cast<float*>(Vs + (size_t)n * KPAD);  // [ATT_WARPS][n]
    float* maskadd = Ps + (size_t)ATT_WARPS * n;       // [n]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (n1 > 0) {
        load_rows_padded(Ks, a.k1 + (long)b * n1 * a.ld1 + h * D, a.ld1, n1, tid, blockDim.x);
        load_rows_padded(Vs, a.v1 + (long)b * n1 * a.ld1 + h * D, a.ld1, n1, tid, blockDim.x);
    }
    load_rows_padded(Ks + (size_t)n1 * KPAD, a.k2 + (long)b * n2 * a.ld2 + h * D, a.ld2, n2, tid, blockDim.x);
    load_rows_padded(Vs + (size_t)n1 * KPAD, a.v2 + (long)b * n2 * a.ld2 + h * D, a.ld2, n2, tid, blockDim.x);
    for (int j = tid; j < n; j += blockDim.x) {
        float m = 0.f;
        if (a.key_mask && j >= n1) m = (1.0f - (float)a.key_mask[(long)b * n2 + (j - n1)]) * -10000.0f;
        maskadd[j] = m;
    }
    RowCtx rc;
    rc.split = -1; rc.w0c = rc.w1c = 1.f;
    if (a.sep_idx) {
        rc.split = (int)a.sep_idx[(long)b * 6 + 2];
        rc.w0c = fminf(fmaxf(*a.w0, 0.f), 0.5f);
        rc.w1c = fminf(fmaxf(*a.w1, 0.5f), 1.0f);
    }
    __syncthreads();

    float* P = Ps + (size_t)warp * n;
    for (int r = warp; r < a.Lq; r += ATT_WARPS) {
        // whole q row in registers (bf16x2 words), same for every lane (broadcast load)
        uint32_t qw[32];
        const uint4* qp = reinterpret_cast<const uint4*>(a.q + ((long)b * a.Lq + r) * a.ldq + h * D);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 t = __ldg(qp + i);
            qw[4 * i] = t.x; qw[4 * i + 1] = t.y; qw[4 * i + 2] = t.z; qw[4 * i + 3] = t.w;
        }
        float mx = -INFINITY;
        for (int j = lane; j < n; j += 32) {
            const uint32_t* kr = reinterpret_cast<const uint32_t*>(Ks + (size_t)j * KPAD);
            float acc = 0.f;
#pragma unroll
            for (int w = 0; w < 32; ++w) {
                const float2 kf = bf2_to_f2(kr[w]), qf = bf2_to_f2(qw[w]);
                acc = fmaf(qf.x, kf.x, acc);
                acc = fmaf(qf.y, kf.y, acc);
            }
            const float s = acc * a.scale * adaptive_factor(rc, r, j - n1) + maskadd[j];
            P[j] = s;
            mx = fmaxf(mx, s);
        }
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < n; j += 32) {
            const float e = __expf(P[j] - mx);
            P[j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        if (lane == 0 && a.lse) a.lse[((long)b * a.nh + h) * a.Lq + r] = mx + __logf(sum);
        __syncwarp();
        float o0 = 0.f, o1 = 0.f;
        for (int j = 0; j < n; ++j) {
            const float p = P[j];
            const float2 vf = bf2_to_f2(reinterpret_cast<const uint32_t*>(Vs + (size_t)j * KPAD)[lane]);
            o0 = fmaf(p, vf.x, o0);
            o1 = fmaf(p, vf.y, o1);
        }
        o0 *= inv; o1 *= inv;
        const long orow = (long)b * a.Lq + r;
        *reinterpret_cast<uint32_t*>(out + orow * a.ldo + h * D + 2 * lane) = pack_bf16x2(o0, o1);
        if (out32) *reinterpret_cast<float2*>(out32 + orow * a.ldo + h * D + 2 * lane) = make_float2(o0, o1);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------ backward, pass 1 (rows)
// per query row r: recompute P, dP = dO V^T, delta_r = sum_j P dP, dS = P (dP - delta); dQ = scale * (dS*f) K;
// adaptive-weight partials: dW = sum dS * S_raw over the block.  delta_r is written for pass 2.
template <int KI>
__global__ void __launch_bounds__(ATT_WARPS * 32)
attention_bwd_rows_kernel(const AttnBwd g, float* __restrict__ delta) {
    extern __shared__ __align__(16) uint8_t smem[];
    const AttnFwd& a = g.f;
    const int h = blockIdx.x, b = blockIdx.y;
    const int n1 = a.n1, n2 = a.n2, n = n1 + n2;
    bf16* Ks = reinterpret_cast<bf16*>(smem);
    bf16* Vs = Ks + (size_t)n * KPAD;
    float* Ps = reinterpret_cast<float*>(Vs + (size_t)n * KPAD);   // [ATT_WARPS][n] : dS * f
    float* maskadd = Ps + (size_t)ATT_WARPS * n;
    __shared__ float wred[2][ATT_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (n1 > 0) {
        load_rows_padded(Ks, a.k1 + (long)b * n1 * a.ld1 + h * D, a.ld1, n1, tid, blockDim.x);
        load_rows_padded(Vs, a.v1 + (long)b * n1 * a.ld1 + h * D, a.ld1, n1, tid, blockDim.x);
    }
    load_rows_padded(Ks + (size_t)n1 * KPAD, a.k2 + (long)b * n2 * a.ld2 + h * D, a.ld2, n2, tid, blockDim.x);
    load_rows_padded(Vs + (size_t)n1 * KPAD, a.v2 + (long)b * n2 * a.ld2 + h * D, a.ld2, n2, tid, blockDim.x);
    for (int j = tid; j < n; j += blockDim.x) {
        float m = 0.f;
        if (a.key_mask && j >= n1) m = (1.0f - (float)a.key_mask[(long)b * n2 + (j - n1)]) * -10000.0f;
        maskadd[j] = m;
    }
    RowCtx rc;
    rc.split = -1; rc.w0c = rc.w1c = 1.f;
    if (a.sep_idx) {
        rc.split = (int)a.sep_idx[(long)b * 6 + 2];
        rc.w0c = fminf(fmaxf(*a.w0, 0.f), 0.5f);
        rc.w1c = fminf(fmaxf(*a.w1, 0.5f), 1.0f);
    }
    __syncthreads();

    float* P = Ps + (size_t)warp * n;
    float dw0 = 0.f, dw1 = 0.f;
    for (int r = warp; r < a.Lq; r += ATT_WARPS) {
        const long grow = (long)b * a.Lq + r;
        uint32_t qw[32];
        const uint4* qp = reinterpret_cast<const uint4*>(a.q + grow * a.ldq + h * D);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 t = __ldg(qp + i);
            qw[4 * i] = t.x; qw[4 * i + 1] = t.y; qw[4 * i + 2] = t.z; qw[4 * i + 3] = t.w;
        }
        float dov[64];
        const float4* dp = reinterpret_cast<const float4*>(g.dout + grow * g.lddo + h * D);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float4 t = __ldg(dp + i);
            dov[4 * i] = t.x; dov[4 * i + 1] = t.y; dov[4 * i + 2] = t.z; dov[4 * i + 3] = t.w;
        }
        const float lse = a.lse[((long)b * a.nh + h) * a.Lq + r];
        float part = 0.f;
        // first sweep: P and dP per key (kept in registers across the delta reduction)
        float pj[KI], dpj[KI], sraw[KI];
#pragma unroll
        for (int t = 0; t < KI; ++t) {
            const int j = lane + 32 * t;
            pj[t] = 0.f; dpj[t] = 0.f; sraw[t] = 0.f;
            if (j < n) {
                const uint32_t* kr = reinterpret_cast<const uint32_t*>(Ks + (size_t)j * KPAD);
                const uint32_t* vr = reinterpret_cast<const uint32_t*>(Vs + (size_t)j * KPAD);
                float acc = 0.f, dacc = 0.f;
#pragma unroll
                for (int w = 0; w < 32; ++w) {
                    const float2 kf = bf2_to_f2(kr[w]), qf = bf2_to_f2(qw[w]), vf = bf2_to_f2(vr[w]);
                    acc = fmaf(qf.x, kf.x, acc);
                    acc = fmaf(qf.y, kf.y, acc);
                    dacc = fmaf(dov[2 * w], vf.x, dacc);
                    dacc = fmaf(dov[2 * w + 1], vf.y, dacc);
                }
                const float sr = acc * a.scale;
                const float s = sr * adaptive_factor(rc, r, j - n1) + maskadd[j];
                const float p = __expf(s - lse);
                pj[t] = p; dpj[t] = dacc; sraw[t] = sr;
                part = fmaf(p, dacc, part);
            }
        }
        const float dl = warp_sum(part);
        if (lane == 0) delta[((long)b * a.nh + h) * a.Lq + r] = dl;
#pragma unroll
        for (int t = 0; t < KI; ++t) {
            const int j = lane + 32 * t;
            if (j < n) {
                const float ds = pj[t] * (dpj[t] - dl);
                const int j2 = j - n1;
                if (rc.split >= 0 && j2 >= rc.split) {
                    if (r < rc.split) dw0 = fmaf(ds, sraw[t], dw0); else dw1 = fmaf(ds, sraw[t], dw1);
                }
                P[j] = ds * adaptive_factor(rc, r, j2);
            }
        }
        __syncwarp();
        float q0 = 0.f, q1 = 0.f;
        for (int j = 0; j < n; ++j) {
            const float ds = P[j];
            const float2 kf = bf2_to_f2(reinterpret_cast<const uint32_t*>(Ks + (size_t)j * KPAD)[lane]);
            q0 = fmaf(ds, kf.x, q0);
            q1 = fmaf(ds, kf.y, q1);
        }
        *reinterpret_cast<uint32_t*>(g.dq + grow * g.ldd + h * D + 2 * lane) = pack_bf16x2(q0 * a.scale, q1 * a.scale);
        __syncwarp();
    }
    if (a.sep_idx) {
        dw0 = warp_sum(dw0);
        dw1 = warp_sum(dw1);
        if (lane == 0) { wred[0][warp] = dw0; wred[1][warp] = dw1; }
        __syncthreads();
        if (tid == 0) {
            float s0 = 0.f, s1 = 0.f;
            for (int w = 0; w < ATT_WARPS; ++w) { s0 += wred[0][w]; s1 += wred[1][w]; }
            // torch.clamp passes the gradient at both (inclusive) bounds
            const float w0 = *a.w0, w1 = *a.w1;
            if (w0 >= 0.f && w0 <= 0.5f) atomicAdd(g.dw0, s0);
            if (w1 >= 0.5f && w1 <= 1.0f) atomicAdd(g.dw1, s1);
        }
    }
}

// ------------------------------------------------------------------------------------------ backward, pass 2 (keys)
// per key j: dV_j = sum_r P_rj dO_r ; dK_j = scale * sum_r dS_rj f_rj Q_r.  Q and dO (bf16-rounded) in shared memory.
__global__ void __launch_bounds__(ATT_WARPS * 32)
attention_bwd_keys_kernel(const AttnBwd g, const float* __restrict__ delta) {
    extern __shared__ __align__(16) uint8_t smem[];
    const AttnFwd& a = g.f;
    const int h = blockIdx.x, b = blockIdx.y;
    const int n1 = a.n1, n2 = a.n2, n = n1 + n2, Lq = a.Lq;
    bf16* Qs = reinterpret_cast<bf16*>(smem);            // [Lq][KPAD]
    bf16* Os = Qs + (size_t)Lq * KPAD;                    // [Lq][KPAD]  dO rounded to bf16
    float* Ps = reinterpret_cast<float*>(Os + (size_t)Lq * KPAD);  // [ATT_WARPS][2][Lq]
    float* lse_s = Ps + (size_t)ATT_WARPS * 2 * Lq;      // [Lq]
    float* dl_s = lse_s + Lq;                            // [Lq]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    load_rows_padded(Qs, a.q + (long)b * Lq * a.ldq + h * D, a.ldq, Lq, tid, blockDim.x);
    for (int i = tid; i < Lq * 32; i += blockDim.x) {
        const int r = i >> 5, c = i & 31;
        const float2 v = *reinterpret_cast<const float2*>(g.dout + ((long)b * Lq + r) * g.lddo + h * D + 2 * c);
        reinterpret_cast<uint32_t*>(Os + (size_t)r * KPAD)[c] = pack_bf16x2(v.x, v.y);
    }
    for (int r = tid; r < Lq; r += blockDim.x) {
        lse_s[r] = a.lse[((long)b * a.nh + h) * Lq + r];
        dl_s[r] = delta[((long)b * a.nh + h) * Lq + r];
    }
    RowCtx rc;
    rc.split = -1; rc.w0c = rc.w1c = 1.f;
    if (a.sep_idx) {
        rc.split = (int)a.sep_idx[(long)b * 6 + 2];
        rc.w0c = fminf(fmaxf(*a.w0, 0.f), 0.5f);
        rc.w1c = fminf(fmaxf(*a.w1, 0.5f), 1.0f);
    }
    __syncthreads();

    float* Pp = Ps + (size_t)warp * 2 * Lq;   // P_rj for this key
    float* Dp = Pp + Lq;                      // dS_rj * f_rj
    for (int j = warp; j < n; j += ATT_WARPS) {
        const bool prefix = j < n1;
        const int j2 = j - n1;
        const bf16* krow = prefix ? a.k1 + ((long)b * n1 + j) * a.ld1 + h * D : a.k2 + ((long)b * n2 + j2) * a.ld2 + h * D;
        const bf16* vrow = prefix ? a.v1 + ((long)b * n1 + j) * a.ld1 + h * D : a.v2 + ((long)b * n2 + j2) * a.ld2 + h * D;
        uint32_t kw[32], vw[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 t = __ldg(reinterpret_cast<const uint4*>(krow) + i);
            kw[4 * i] = t.x; kw[4 * i + 1] = t.y; kw[4 * i + 2] = t.z; kw[4 * i + 3] = t.w;
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(vrow) + i);
            vw[4 * i] = u.x; vw[4 * i + 1] = u.y; vw[4 * i + 2] = u.z; vw[4 * i + 3] = u.w;
        }
        float madd = 0.f;
        if (a.key_mask && !prefix) madd = (1.0f - (float)a.key_mask[(long)b * n2 + j2]) * -10000.0f;
        for (int r = lane; r < Lq; r += 32) {
            const uint32_t* qr = reinterpret_cast<const uint32_t*>(Qs + (size_t)r * KPAD);
            const uint32_t* orow = reinterpret_cast<const uint32_t*>(Os + (size_t)r * KPAD);
            float acc = 0.f, dacc = 0.f;
#pragma unroll
            for (int w = 0; w < 32; ++w) {
                const float2 qf = bf2_to_f2(qr[w]), kf = bf2_to_f2(kw[w]), of = bf2_to_f2(orow[w]), vf = bf2_to_f2(vw[w]);
                acc = fmaf(qf.x, kf.x, acc);
                acc = fmaf(qf.y, kf.y, acc);
                dacc = fmaf(of.x, vf.x, dacc);
                dacc = fmaf(of.y, vf.y, dacc);
            }
            const float f = adaptive_factor(rc, r, j2);
            const float p = __expf(acc * a.scale * f + madd - lse_s[r]);
            Pp[r] = p;
            Dp[r] = p * (dacc - dl_s[r]) * f;
        }
        __syncwarp();
        float k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
        for (int r = 0; r < Lq; ++r) {
            const float p = Pp[r], ds = Dp[r];
            const float2 qf = bf2_to_f2(reinterpret_cast<const uint32_t*>(Qs + (size_t)r * KPAD)[lane]);
            const float2 of = bf2_to_f2(reinterpret_cast<const uint32_t*>(Os + (size_t)r * KPAD)[lane]);
            k0 = fmaf(ds, qf.x, k0); k1 = fmaf(ds, qf.y, k1);
            v0 = fmaf(p, of.x, v0);  v1 = fmaf(p, of.y, v1);
        }
        k0 *= a.scale; k1 *= a.scale;
        if (prefix) {
            const long o = ((long)b * n1 + j) * g.ldd1 + h * D + 2 * lane;
            *reinterpret_cast<float2*>(g.dk1 + o) = make_float2(k0, k1);
            *reinterpret_cast<float2*>(g.dv1 + o) = make_float2(v0, v1);
        } else {
            const long row = (long)b * n2 + j2;
            if (g.dk2_add) {
                const float2 ak = *reinterpret_cast<const float2*>(g.dk2_add + row * g.ldadd + h * D + 2 * lane);
                const float2 av = *reinterpret_cast<const float2*>(g.dv2_add + row * g.ldadd + h * D + 2 * lane);
                k0 += ak.x; k1 += ak.y; v0 += av.x; v1 += av.y;
            }
            *reinterpret_cast<uint32_t*>(g.dk2 + row * g.ldd + h * D + 2 * lane) = pack_bf16x2(k0, k1);
            *reinterpret_cast<uint32_t*>(g.dv2 + row * g.ldd + h * D + 2 * lane) = pack_bf16x2(v0, v1);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------ BertFusion (fp32)
constexpr int FROWS = 8;      // query rows (forward, backward rows pass) or keys (backward keys pass) per CTA
constexpr int FTHREADS = 256;

// scores of FROWS query rows (in shared memory, fp32 [FROWS][H]) against all Nv rows of vh: warp per key.
__device__ __forceinline__ void fusion_scores(const float* __restrict__ xs, const float* __restrict__ vh, float* S, int Nv,
                                              int H, int ldS, int lane, int warp) {
    for (int j = warp; j < Nv; j += FTHREADS / 32) {
        float acc[FROWS];
#pragma unroll
        for (int r = 0; r < FROWS; ++r) acc[r] = 0.f;
        for (int c = lane * 4; c < H; c += 128) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(vh + (long)j * H + c));
#pragma unroll
            for (int r = 0; r < FROWS; ++r) {
                const float4 x = *reinterpret_cast<const float4*>(xs + r * H + c);
                acc[r] = fmaf(x.x, v.x, fmaf(x.y, v.y, fmaf(x.z, v.z, fmaf(x.w, v.w, acc[r]))));
            }
        }
#pragma unroll
        for (int r = 0; r < FROWS; ++r) {
            const float s = warp_sum(acc[r]);
            if (lane == 0) S[r * ldS + j] = s;
        }
    }
}

__global__ void __launch_bounds__(FTHREADS)
fusion_fwd_kernel(const float* __restrict__ ctx, const float* __restrict__ vh, bf16* __restrict__ out, float* __restrict__ lse,
                  int L, int Nv, int H) {
    extern __shared__ __align__(16) float fs[];
    float* xs = fs;                    // [FROWS][H]
    float* S = xs + FROWS * H;         // [FROWS][Nv]
    const int b = blockIdx.y, r0 = blockIdx.x * FROWS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* vb = vh + (long)b * Nv * H;
    for (int i = tid; i < FROWS * H / 4; i += FTHREADS) {
        const int r = (i * 4) / H, c = (i * 4) % H;
        float4 v = make_float4(0, 0, 0, 0);
        if (r0 + r < L) v = *reinterpret_cast<const float4*>(ctx + ((long)b * L + r0 + r) * H + c);
        *reinterpret_cast<float4*>(xs + r * H + c) = v;
    }
    __syncthreads();
    fusion_scores(xs, vb, S, Nv, H, Nv, lane, warp);
    __syncthreads();
    {   // softmax: warp per row (FROWS == number of warps)
        const int r = warp;
        float mx = -INFINITY;
        for (int j = lane; j < Nv; j += 32) mx = fmaxf(mx, S[r * Nv + j]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int j = lane; j < Nv; j += 32) {
            const float e = __expf(S[r * Nv + j] - mx);
            S[r * Nv + j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        for (int j = lane; j < Nv; j += 32) S[r * Nv + j] *= inv;
        if (lane == 0 && r0 + r < L) lse[(long)b * L + r0 + r] = mx + __logf(sum);
    }
    __syncthreads();
    for (int c = tid; c < H; c += FTHREADS) {
        float acc[FROWS];
#pragma unroll
        for (int r = 0; r < FROWS; ++r) acc[r] = 0.f;
        for (int j = 0; j < Nv; ++j) {
            const float v = __ldg(vb + (long)j * H + c);
#pragma unroll
            for (int r = 0; r < FROWS; ++r) acc[r] = fmaf(S[r * Nv + j], v, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < FROWS; ++r)
            if (r0 + r < L) out[((long)b * L + r0 + r) * H + c] = __float2bfloat16(acc[r]);
    }
}

// rows pass: P, dS to global; dctx += dS vh
__global__ void __launch_bounds__(FTHREADS)
fusion_bwd_rows_kernel(const float* __restrict__ dg, const float* __restrict__ ctx, const float* __restrict__ vh,
                       const float* __restrict__ lse, float* __restrict__ pbuf, float* __restrict__ dsbuf,
                       float* __restrict__ dctx, int L, int Nv, int H) {
    extern __shared__ __align__(16) float fs[];
    float* xs = fs;                    // [FROWS][H]  ctx rows, later dG rows
    float* S = xs + FROWS * H;         // [FROWS][Nv] P
    float* dP = S + FROWS * Nv;        // [FROWS][Nv] dP -> dS
    const int b = blockIdx.y, r0 = blockIdx.x * FROWS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* vb = vh + (long)b * Nv * H;
    auto load_rows = [&](const float* src) {
        for (int i = tid; i < FROWS * H / 4; i += FTHREADS) {
            const int r = (i * 4) / H, c = (i * 4) % H;
            float4 v = make_float4(0, 0, 0, 0);
            if (r0 + r < L) v = *reinterpret_cast<const float4*>(src + ((long)b * L + r0 + r) * H + c);
            *reinterpret_cast<float4*>(xs + r * H + c) = v;
        }
    };
    load_rows(ctx);
    __syncthreads();
    fusion_scores(xs, vb, S, Nv, H, Nv, lane, warp);
    __syncthreads();
    load_rows(dg);
    __syncthreads();
    fusion_scores(xs, vb, dP, Nv, H, Nv, lane, warp);
    __syncthreads();
    {
        const int r = warp;
        const float l = (r0 + r < L) ? lse[(long)b * L + r0 + r] : 0.f;
        float part = 0.f;
        for (int j = lane; j < Nv; j += 32) {
            const float p = __expf(S[r * Nv + j] - l);
            S[r * Nv + j] = p;
            part = fmaf(p, dP[r * Nv + j], part);
        }
        const float dl = warp_sum(part);
        for (int j = lane; j < Nv; j += 32) {
            const float ds = S[r * Nv + j] * (dP[r * Nv + j] - dl);
            dP[r * Nv + j] = ds;
            if (r0 + r < L) {
                pbuf[((long)b * L + r0 + r) * Nv + j] = S[r * Nv + j];
                dsbuf[((long)b * L + r0 + r) * Nv + j] = ds;
            }
        }
    }
    __syncthreads();
    for (int c = tid; c < H; c += FTHREADS) {
        float acc[FROWS];
#pragma unroll
        for (int r = 0; r < FROWS; ++r) acc[r] = 0.f;
        for (int j = 0; j < Nv; ++j) {
            const float v = __ldg(vb + (long)j * H + c);
#pragma unroll
            for (int r = 0; r < FROWS; ++r) acc[r] = fmaf(dP[r * Nv + j], v, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < FROWS; ++r)
            if (r0 + r < L) dctx[((long)b * L + r0 + r) * H + c] += acc[r];
    }
}

// keys pass: dvh[j] = dvh_in[j] + sum_r dS[r][j] ctx[r] + P[r][j] dG[r]
__global__ void __launch_bounds__(FTHREADS)
fusion_bwd_keys_kernel(const float* __restrict__ dg, const float* __restrict__ ctx, const float* __restrict__ pbuf,
                       const float* __restrict__ dsbuf, const float* __restrict__ dvh_in, float* __restrict__ dvh32,
                       bf16* __restrict__ dvh16, int L, int Nv, int H) {
    extern __shared__ __align__(16) float fs[];
    float* Pk = fs;                 // [L][FROWS]
    float* Dk = Pk + L * FROWS;     // [L][FROWS]
    const int b = blockIdx.y, j0 = blockIdx.x * FROWS;
    const int tid = threadIdx.x;
    for (int i = tid; i < L * FROWS; i += FTHREADS) {
        const int r = i / FROWS, k = i % FROWS;
        const bool ok = j0 + k < Nv;
        Pk[i] = ok ? pbuf[((long)b * L + r) * Nv + j0 + k] : 0.f;
        Dk[i] = ok ? dsbuf[((long)b * L + r) * Nv + j0 + k] : 0.f;
    }
    __syncthreads();
    for (int c = tid; c < H; c += FTHREADS) {
        float acc[FROWS];
#pragma unroll
        for (int k = 0; k < FROWS; ++k) acc[k] = 0.f;
        for (int r = 0; r < L; ++r) {
            const float x = __ldg(ctx + ((long)b * L + r) * H + c);
            const float d = __ldg(dg + ((long)b * L + r) * H + c);
#pragma unroll
            for (int k = 0; k < FROWS; ++k) acc[k] = fmaf(Dk[r * FROWS + k], x, fmaf(Pk[r * FROWS + k], d, acc[k]));
        }
#pragma unroll
        for (int k = 0; k < FROWS; ++k)
            if (j0 + k < Nv) {
                const long o = ((long)b * Nv + j0 + k) * H + c;
                const float v = acc[k] + (dvh_in ? dvh_in[o] : 0.f);
                if (dvh32) dvh32[o] = v;
                if (dvh16) dvh16[o] = __float2bfloat16(v);
            }
    }
}

int check_attn(const AttnFwd& a) {
    B2_CHECK(a.B > 0 && a.Lq > 0 && a.nh > 0 && a.n2 > 0 && a.n1 >= 0, "attention: bad sizes B=%d Lq=%d nh=%d n1=%d n2=%d", a.B,
             a.Lq, a.nh, a.n1, a.n2);
    B2_CHECK(a.n1 + a.n2 <= 32 * MAX_KEY_ITERS, "attention: %d keys exceed the supported %d", a.n1 + a.n2, 32 * MAX_KEY_ITERS);
    B2_CHECK(a.ldq % 8 == 0 && a.ld2 % 8 == 0 && (a.n1 == 0 || a.ld1 % 8 == 0), "attention: leading dimensions must be multiples of 8");
    B2_CHECK(a.sep_idx == nullptr || (a.w0 && a.w1 && a.n1 == 0), "attention: adaptive scaling needs w0/w1 and no key prefix");
    return 0;
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
    const int n = a.n1 + a.n2;
    const size_t smem = (size_t)2 * n * KPAD * 2 + (size_t)(ATT_WARPS + 1) * n * 4;
    B2_TRY(set_smem(attention_fwd_kernel, smem));
    attention_fwd_kernel<<<dim3(a.nh, a.B), ATT_WARPS * 32, smem, s>>>(a, a.out, out32);
    B2_LAUNCHED();
    return 0;
}

int launch_attention_bwd(const AttnBwd& g, float* delta, cudaStream_t s) {
    const AttnFwd& a = g.f;
    B2_TRY(check_attn(a));
    B2_CHECK(g.dout && g.dq && g.dk2 && g.dv2 && delta && a.lse, "attention bwd: missing buffers");
    B2_CHECK(a.n1 == 0 || (g.dk1 && g.dv1), "attention bwd: prefix gradients need dk1/dv1");
    B2_CHECK(g.lddo % 4 == 0, "attention bwd: lddo must be a multiple of 4");
    const int n = a.n1 + a.n2;
    const size_t smem1 = (size_t)2 * n * KPAD * 2 + (size_t)(ATT_WARPS + 1) * n * 4;
    if (n <= 96) {
        B2_TRY(set_smem(attention_bwd_rows_kernel<3>, smem1));
        attention_bwd_rows_kernel<3><<<dim3(a.nh, a.B), ATT_WARPS * 32, smem1, s>>>(g, delta);
    } else if (n <= 192) {
        B2_TRY(set_smem(attention_bwd_rows_kernel<6>, smem1));
        attention_bwd_rows_kernel<6><<<dim3(a.nh, a.B), ATT_WARPS * 32, smem1, s>>>(g, delta);
    } else {
        B2_TRY(set_smem(attention_bwd_rows_kernel<MAX_KEY_ITERS>, smem1));
        attention_bwd_rows_kernel<MAX_KEY_ITERS><<<dim3(a.nh, a.B), ATT_WARPS * 32, smem1, s>>>(g, delta);
    }
    B2_LAUNCHED();
    const size_t smem2 = (size_t)2 * a.Lq * KPAD * 2 + (size_t)(2 * ATT_WARPS + 2) * a.Lq * 4;
    B2_TRY(set_smem(attention_bwd_keys_kernel, smem2));
    attention_bwd_keys_kernel<<<dim3(a.nh, a.B), ATT_WARPS * 32, smem2, s>>>(g, delta);
    B2_LAUNCHED();
    return 0;
}

int launch_fusion_fwd(const float* ctx, const float* vh, bf16* out, float* lse, int B, int L, int Nv, int H, cudaStream_t s) {
    B2_CHECK(H % 128 == 0, "fusion: H %d must be a multiple of 128", H);
    const size_t smem = (size_t)(FROWS * H + FROWS * Nv) * 4;
    B2_TRY(set_smem(fusion_fwd_kernel, smem));
    fusion_fwd_kernel<<<dim3((L + FROWS - 1) / FROWS, B), FTHREADS, smem, s>>>(ctx, vh, out, lse, L, Nv, H);
    B2_LAUNCHED();
    return 0;
}

int launch_fusion_bwd_rows(const float* dg, const float* ctx, const float* vh, const float* lse, float* pbuf, float* dsbuf,
                           float* dctx, int B, int L, int Nv, int H, cudaStream_t s) {
    B2_CHECK(H % 128 == 0, "fusion: H %d must be a multiple of 128", H);
    const size_t smem = (size_t)(FROWS * H + 2 * FROWS * Nv) * 4;
    B2_TRY(set_smem(fusion_bwd_rows_kernel, smem));
    fusion_bwd_rows_kernel<<<dim3((L + FROWS - 1) / FROWS, B), FTHREADS, smem, s>>>(dg, ctx, vh, lse, pbuf, dsbuf, dctx, L, Nv, H);
    B2_LAUNCHED();
    return 0;
}

int launch_fusion_bwd_keys(const float* dg, const float* ctx, const float* pbuf, const float* dsbuf, const float* dvh_in,
                           float* dvh32, bf16* dvh16, int B, int L, int Nv, int H, cudaStream_t s) {
    const size_t smem = (size_t)2 * L * FROWS * 4;
    B2_TRY(set_smem(fusion_bwd_keys_kernel, smem));
    fusion_bwd_keys_kernel<<<dim3((Nv + FROWS - 1) / FROWS, B), FTHREADS, smem, s>>>(dg, ctx, pbuf, dsbuf, dvh_in, dvh32, dvh16, L, Nv, H);
    B2_LAUNCHED();
    return 0;
}
