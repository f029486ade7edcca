// Persistent tcgen05 GEMM kernel for sm_100a (device code; instantiated per N-tile in gemm_inst*.cu).
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> shared-memory ring -> tcgen05.mma (one elected thread, fp32
//   accumulators in TMEM, DOUBLE-BUFFERED so the epilogue of tile i overlaps the main loop of tile i+1) ->
//   tcgen05.ld epilogue with a compile-time-specialised fused epilogue.
//   One CTA per SM loops over output tiles (static round-robin scheduler); three pipelines: smem full/empty
//   (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue) and the tile loop.
//
// Replaces, for the MarT hot path, every nn.Linear product of the reference model
// (MarT/models/modeling_unimo.py:212-287 CLIP attention/MLP, :317-478 BERT attention/FFN,
// :956-976 MLM head) and their autograd transposes (dX = dY W, dW = dY^T X).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (warp w owns TMEM lanes 32*(w%4) .. +31, i.e. 32 rows of the 128-row tile).
// Epilogue data path: TMEM -> registers (lane = row, 64 columns at a time) -> per-warp shared-memory transpose
// -> lane = column PAIR, so that every global access of the epilogue (bias, saved pre-activation, residual,
// outputs, reductions) is a coalesced 256-byte (fp32) / 128-byte (bf16) row segment.
#pragma once
#include "gemm.cuh"

namespace gemmk {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int GEMM_THREADS = 192;
constexpr int PAIR_EPI_WARPS = 8;                       // CTA-pair kernel: two epilogue warps per TMEM lane quadrant
constexpr int PAIR_THREADS = 64 + 32 * PAIR_EPI_WARPS;   // 320
constexpr int EPI_COLS = 64;   // columns per epilogue pass
constexpr int EPI_LD = 64;     // row of the per-warp transpose buffer (fp32 words; XOR-swizzled chunks instead of padding)

// fused epilogues (compile-time).  Every kind except GENERIC requires even N and even leading dimensions.
enum EpiKind {
    EPI_GENERIC = 0,  // any combination of the GemmOp epilogue fields, scalar accesses
    EPI_BF16 = 1,     // out_bf16 = acc*alpha + bias (+bias2)
    EPI_ACT = 2,      // pre = acc*alpha + bias (+bias2); out_preact = pre (or act'(pre)); out_bf16 = act(pre)
    EPI_RES = 3,      // out_f32 = acc*alpha (+bias) + residual   [+ out_bf16 copy]
    EPI_F32 = 4,      // out_f32 = acc*alpha (+bias)
    EPI_DACT = 5,     // out_bf16 = acc*alpha * dact_src   (dact_src holds the saved activation derivative, dact = 3)
    EPI_ACC = 6,      // out_f32 += acc*alpha   (vector reductions; split-K safe)
};

struct GemmDev {
    int M, N;
    int kb1, kb_total;  // k-blocks served by pair 1 / total
    int mt, nt, nsplit, kb_per;   // tile grid, split-K factor, k-blocks per split
    float alpha;
    const float* bias;
    const float* bias2;
    bf16* out_preact; long ldp;
    int act, save_grad;
    const bf16* dact_src; long ldx; int dact;
    const float* residual; long ldr;
    float* out_f32; long ldo32;
    bf16* out_bf16; long ldo16;
    int accumulate;
    DropCfg drop;                            // dropout on (acc*alpha + bias) before the residual add (EPI_RES / GENERIC)
    float* colsum_out; float* colsum_out2;   // EPI_DACT: += column sums of the output (bias gradient of the producing layer)
    unsigned long long* prof;   // b200mart_gemm_timeline: 16 stamps per CTA (null = off), see GemmStamp
    int debug;   // B200MART_GEMM_DEBUG bitmask (bring-up only): 1 no epilogue global access, 2 no TMA, 4 no MMA
    int b_static;   // GemmOp::b_static
};


// Intra-kernel timeline (b200mart_gemm_timeline; profiling runs only): per CTA 16 x u64
enum GemmStamp {
    ST_GT_START = 0, ST_CLK_START, ST_CLK_PROLOGUE, ST_CLK_TMA0, ST_CLK_FULL0, ST_CLK_MMA_ISSUED0, ST_CLK_TFULL0, ST_CLK_EPI0,
    ST_CLK_EPI_LAST, ST_CLK_END, ST_GT_END, ST_TILES_SMID, ST_SHAPE_MN, ST_SHAPE_K_GRID, ST_KIND, ST_UNUSED
};
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t smid() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void stamp(const GemmDev& p, int slot) {
    if (p.prof) p.prof[(size_t)blockIdx.x * 16 + slot] = (unsigned long long)clock64();
}
__device__ __forceinline__ void stamp_begin(const GemmDev& p, int kind, int bn) {
    if (!p.prof) return;
    unsigned long long* r = p.prof + (size_t)blockIdx.x * 16;
    r[ST_GT_START] = gtimer();
    r[ST_CLK_START] = (unsigned long long)clock64();
    r[ST_SHAPE_MN] = ((unsigned long long)(uint32_t)p.N << 32) | (uint32_t)p.M;
    r[ST_SHAPE_K_GRID] = ((unsigned long long)gridDim.x << 32) | (uint32_t)p.kb_total;
    r[ST_KIND] = ((unsigned long long)(uint32_t)bn << 32) | (uint32_t)kind;
}
__device__ __forceinline__ void stamp_end(const GemmDev& p, int tiles) {
    if (!p.prof) return;
    unsigned long long* r = p.prof + (size_t)blockIdx.x * 16;
    r[ST_CLK_END] = (unsigned long long)clock64();
    r[ST_GT_END] = gtimer();
    r[ST_TILES_SMID] = ((unsigned long long)smid() << 32) | (uint32_t)tiles;
}

template <int BN>
struct SmemLayout {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // as deep as 227 KB allows: the ring has to cover the TMA round trip (~1 us) at one stage per 128-512 MMA cycles
    static constexpr int STAGES = (BN >= 256) ? 4 : ((BN >= 128) ? 6 : 8);
    static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int EPI_BYTES = 4 * 32 * EPI_LD * 4;
    static constexpr int BAR_OFFSET = EPI_OFFSET + EPI_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // barriers + alignment slack
    static constexpr int TMEM_COLS = 2 * BN;               // two accumulator buffers (power of two >= 32)
};

__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float act_fwd(int act, float v) {
    return act == 1 ? gelu_erf(v) : (act == 2 ? quick_gelu(v) : v);
}
__device__ __forceinline__ float act_grad(int act, float x) { return act == 1 ? gelu_erf_grad(x) : quick_gelu_grad(x); }

// Per-warp transpose buffer of one epilogue pass: 32 rows x 64 columns fp32, NO padding; the 16-byte chunk c of row r lives
// at chunk (c ^ (r & 15)).  Writers (lane = row, 16 chunks) and readers (16 lanes = the 16 chunks of one row) both touch
// eight different 4-bank groups per quarter warp, so every 128-bit shared-memory access is conflict-free.
// (EC = 32 columns per pass, the half-footprint kernels: rows of 8 chunks, chunk c of row r at c ^ (r & 7), same property.)
template <int EC = EPI_COLS>
__device__ __forceinline__ int stg_off(int row, int chunk) { return row * EC + ((chunk ^ (row & (EC / 4 - 1))) << 2); }

// Epilogue work split of a pass: lane = (row parity hr = lane >> 4, column quad c = lane & 15); iteration i handles row
// 2 i + hr, columns 4 c .. 4 c + 3, so each warp-wide access covers two full 256-byte (fp32) / 128-byte (bf16) row segments.
// Per-pass inputs that come from HBM (residual / saved activation derivative): 16 vector loads per lane, all issued back to
// back before the TMEM read + transpose of the pass (pass 0: before the accumulator-ready wait).
// EC columns per pass: LPR = EC / 4 lanes cover one row segment, RPI = 32 / LPR rows per warp-wide access, IT = 32 / RPI
// iterations per pass (EC = 64: 16 lanes / 2 rows / 16 iterations; EC = 32: 8 lanes / 4 rows / 8 iterations).
template <int EPI, int EC = EPI_COLS>
struct EpiIn {
    static constexpr bool kRes = (EPI == EPI_RES), kDact = (EPI == EPI_DACT);
    static constexpr int LPR = EC / 4, RPI = 32 / LPR, IT = 32 / RPI;
    float4 rs[kRes ? IT : 1];
    uint2 xs[kDact ? IT : 1];
    template <bool FULL>
    __device__ __forceinline__ void load_t(const GemmDev& p, int row0, int nrows, int col, int hr) {
        if (kRes) {
            const float* rp = p.residual + (long)(row0 + hr) * p.ldr + col;
            const int ld2 = RPI * (int)p.ldr;
#pragma unroll
            for (int i = 0; i < IT; ++i)
                if (FULL || RPI * i + hr < nrows) rs[i] = *reinterpret_cast<const float4*>(rp + i * ld2);
        }
        if (kDact) {
            const bf16* xp = p.dact_src + (long)(row0 + hr) * p.ldx + col;
            const int ld2 = RPI * (int)p.ldx;
#pragma unroll
            for (int i = 0; i < IT; ++i)
                if (FULL || RPI * i + hr < nrows) xs[i] = *reinterpret_cast<const uint2*>(xp + i * ld2);
        }
    }
    __device__ __forceinline__ void load(const GemmDev& p, int row0, int nrows, int col0, int lane) {
        const int col = col0 + 4 * (lane % LPR);
        if (!(kRes || kDact) || col >= p.N) return;
        if (nrows >= 32) load_t<true>(p, row0, nrows, col, lane / LPR);
        else load_t<false>(p, row0, nrows, col, lane / LPR);
    }
};

__device__ __forceinline__ float rcp_fast(float x) {      // MUFU.RCP (1 ulp), no denormal / special-case slow path
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// activation value and derivative from ONE exponential.
//   exact GELU: y = x Phi(x), y' = Phi(x) + x phi(x); Phi through erf(z) = 1 - P(t) exp(-z^2), t = 1 / (1 + 0.3275911 z)
//   (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7 - far inside the bf16 rounding of the outputs), with z = |x| / sqrt(2) so
//   that exp(-z^2) = exp(-x^2 / 2) is also the Gaussian of the derivative: one MUFU.EX2 + one MUFU.RCP per element.
template <int ACT> __device__ __forceinline__ void act_both(float x, float& y, float& dy) {
    if (ACT == 1) {
        const float ax = fabsf(x);
        const float t = rcp_fast(fmaf(0.3275911f * 0.70710678118654752f, ax, 1.0f));
        const float e = __expf(-0.5f * x * x);
        float q = fmaf(1.061405429f, t, -1.453152027f);
        q = fmaf(q, t, 1.421413741f);
        q = fmaf(q, t, -0.284496736f);
        q = fmaf(q, t, 0.254829592f);
        const float h = 0.5f * q * t * e;             // = 0.5 erfc(|x| / sqrt 2) = Phi(-|x|)
        const float cdf = (x >= 0.f) ? 1.0f - h : h;
        y = x * cdf;
        dy = fmaf(x * 0.39894228040143268f, e, cdf);
    } else {                 // quick GELU: y = x s, s = sigmoid(1.702 x), y' = s + 1.702 x s (1 - s)
        const float sg = rcp_fast(1.0f + __expf(-1.702f * x));
        y = x * sg;
        dy = fmaf(1.702f * y, 1.0f - sg, sg);
    }
}

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint2 pack_bf16x4(float4 v) { return make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w)); }

// One 32-row x 64-column block of the output from the transpose buffer.  FULL: all 32 rows valid -> no predicates, the
// 16 iterations are independent straight-line chains the scheduler interleaves (the epilogue warps have no other latency
// hiding: one warp per scheduler).  DROP / OUT16 / SAVEG: runtime flags hoisted to compile time.
template <int EPI, int ACT, bool FULL, bool DROP, bool OUT16, bool SAVEG, int EC = EPI_COLS>
__device__ __forceinline__ void epilogue_rows_t(const GemmDev& p, const float* stg, const EpiIn<EPI, EC>& in, int row0, int nrows,
                                                int col, float4 b, int lane) {
    constexpr bool kF32 = (EPI == EPI_RES || EPI == EPI_F32 || EPI == EPI_ACC);
    constexpr bool k16 = (EPI == EPI_BF16 || EPI == EPI_ACT || EPI == EPI_DACT || (EPI == EPI_RES && OUT16));
    constexpr int LPR = EC / 4, RPI = 32 / LPR, IT = 32 / RPI;
    const int hr = lane / LPR, c = lane % LPR;
    float* o32 = kF32 ? p.out_f32 + (long)(row0 + hr) * p.ldo32 + col : nullptr;
    bf16* o16 = k16 ? p.out_bf16 + (long)(row0 + hr) * p.ldo16 + col : nullptr;
    bf16* opre = (EPI == EPI_ACT) ? p.out_preact + (long)(row0 + hr) * p.ldp + col : nullptr;
    const int ld32 = RPI * (int)p.ldo32, ld16 = RPI * (int)p.ldo16, ldp = RPI * (int)p.ldp;
    const float alpha = p.alpha;
    const uint32_t e0 = (uint32_t)(row0 + hr) * (uint32_t)p.N + (uint32_t)col;
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int kUnroll = (EPI == EPI_ACT) ? 4 : IT;
#pragma unroll(kUnroll)
    for (int i = 0; i < IT; ++i) {
        const int rl = RPI * i + hr;
        if (!FULL && rl >= nrows) continue;
        const float4 a = *reinterpret_cast<const float4*>(stg + stg_off<EC>(rl, c));
        float4 v = make_float4(fmaf(a.x, alpha, b.x), fmaf(a.y, alpha, b.y), fmaf(a.z, alpha, b.z), fmaf(a.w, alpha, b.w));
        if (EPI == EPI_BF16) {
            *reinterpret_cast<uint2*>(o16 + i * ld16) = pack_bf16x4(v);
        } else if (EPI == EPI_ACT) {
            float4 y, d;
            act_both<ACT>(v.x, y.x, d.x);
            act_both<ACT>(v.y, y.y, d.y);
            act_both<ACT>(v.z, y.z, d.z);
            act_both<ACT>(v.w, y.w, d.w);
            // saved for backward: the pre-activation, or the activation derivative itself (save_grad)
            *reinterpret_cast<uint2*>(opre + i * ldp) = pack_bf16x4(SAVEG ? d : v);
            *reinterpret_cast<uint2*>(o16 + i * ld16) = pack_bf16x4(y);
        } else if (EPI == EPI_RES) {
            if (DROP) {   // BertSelfOutput :389 / BertOutput :476: dropout(dense(x)) + residual
                const uint32_t e = e0 + (uint32_t)(RPI * i) * (uint32_t)p.N;
                v.x *= drop_mult(p.drop, e); v.y *= drop_mult(p.drop, e + 1);
                v.z *= drop_mult(p.drop, e + 2); v.w *= drop_mult(p.drop, e + 3);
            }
            const float4 r = in.rs[EpiIn<EPI, EC>::kRes ? i : 0];
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
            *reinterpret_cast<float4*>(o32 + i * ld32) = v;
            if (OUT16) *reinterpret_cast<uint2*>(o16 + i * ld16) = pack_bf16x4(v);
        } else if (EPI == EPI_F32) {
            *reinterpret_cast<float4*>(o32 + i * ld32) = v;
        } else if (EPI == EPI_DACT) {
            const uint2 xs = in.xs[EpiIn<EPI, EC>::kDact ? i : 0];
            const float2 x0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xs.x));
            const float2 x1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xs.y));
            v.x *= x0.x; v.y *= x0.y; v.z *= x1.x; v.w *= x1.y;
            cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
            *reinterpret_cast<uint2*>(o16 + i * ld16) = pack_bf16x4(v);
        } else if (EPI == EPI_ACC) {
            red_add_v4(o32 + i * ld32, v);
        }
    }
    if (EPI == EPI_DACT) {     // column sums of this block: fold the row groups, then one vector reduction per column quad
#pragma unroll
        for (int o = 16; o >= LPR; o >>= 1) {
            cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
            cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
        }
        if (hr == 0) {
            if (p.colsum_out) red_add_v4(p.colsum_out + col, cs);
            if (p.colsum_out2) red_add_v4(p.colsum_out2 + col, cs);
        }
    }
}

template <int EPI, int ACT, bool FULL, int EC = EPI_COLS>
__device__ __forceinline__ void epilogue_rows(const GemmDev& p, const float* stg, const EpiIn<EPI, EC>& in, int row0, int nrows,
                                              int col, float4 b, int lane) {
    if (EPI == EPI_RES) {
        const bool drop = p.drop.thresh != 0, o16 = p.out_bf16 != nullptr;
        if (drop && o16) epilogue_rows_t<EPI, ACT, FULL, true, true, false, EC>(p, stg, in, row0, nrows, col, b, lane);
        else if (drop) epilogue_rows_t<EPI, ACT, FULL, true, false, false, EC>(p, stg, in, row0, nrows, col, b, lane);
        else if (o16) epilogue_rows_t<EPI, ACT, FULL, false, true, false, EC>(p, stg, in, row0, nrows, col, b, lane);
        else epilogue_rows_t<EPI, ACT, FULL, false, false, false, EC>(p, stg, in, row0, nrows, col, b, lane);
    } else if (EPI == EPI_ACT) {
        if (p.save_grad) epilogue_rows_t<EPI, ACT, FULL, false, false, true, EC>(p, stg, in, row0, nrows, col, b, lane);
        else epilogue_rows_t<EPI, ACT, FULL, false, false, false, EC>(p, stg, in, row0, nrows, col, b, lane);
    } else {
        epilogue_rows_t<EPI, ACT, FULL, false, false, false, EC>(p, stg, in, row0, nrows, col, b, lane);
    }
}

template <int EPI, int EC = EPI_COLS>
__device__ __forceinline__ void epilogue_block(const GemmDev& p, const float* stg, const EpiIn<EPI, EC>& in, int row0, int nrows,
                                               int col0, int lane) {
    if (EPI == EPI_GENERIC) {
        // scalar, any flags: lane handles columns col0+lane and col0+32+lane
#pragma unroll 1
        for (int half = 0; half < EC / 32; ++half) {
            const int cl = half * 32 + lane;
            const int col = col0 + cl;
            if (col >= p.N) continue;
            float badd = 0.f;
            if (p.bias) badd = __ldg(p.bias + col);
            if (p.bias2) badd += __ldg(p.bias2 + col);
#pragma unroll 1
            for (int rr = 0; rr < nrows; ++rr) {
                const long row = row0 + rr;
                float v = stg[stg_off<EC>(rr, cl >> 2) + (cl & 3)] * p.alpha + badd;
                if (p.out_preact) p.out_preact[row * p.ldp + col] = __float2bfloat16((p.save_grad && p.act) ? act_grad(p.act, v) : v);
                v = act_fwd(p.act, v);
                if (p.dact) {
                    const float x = __bfloat162float(p.dact_src[row * p.ldx + col]);
                    v *= (p.dact == 3) ? x : act_grad(p.dact, x);
                }
                if (p.drop.thresh) v *= drop_mult(p.drop, (uint32_t)row * (uint32_t)p.N + (uint32_t)col);
                if (p.residual) v += p.residual[row * p.ldr + col];
                if (p.out_f32) {
                    if (p.accumulate) atomicAdd(p.out_f32 + row * p.ldo32 + col, v);
                    else p.out_f32[row * p.ldo32 + col] = v;
                }
                if (p.out_bf16) p.out_bf16[row * p.ldo16 + col] = __float2bfloat16(v);
            }
        }
        return;
    }
    const int col = col0 + 4 * (lane % (EC / 4));       // N % 4 == 0 -> the quad is valid or invalid together
    if (col >= p.N) return;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (EPI == EPI_BF16 || EPI == EPI_ACT || EPI == EPI_RES || EPI == EPI_F32) {
        if (p.bias) b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        if ((EPI == EPI_BF16 || EPI == EPI_ACT) && p.bias2) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias2 + col));
            b.x += t.x; b.y += t.y; b.z += t.z; b.w += t.w;
        }
    }
    if (EPI == EPI_ACT) {
        if (nrows >= 32) {
            if (p.act == 1) epilogue_rows<EPI, 1, true, EC>(p, stg, in, row0, nrows, col, b, lane);
            else epilogue_rows<EPI, 2, true, EC>(p, stg, in, row0, nrows, col, b, lane);
        } else {
            if (p.act == 1) epilogue_rows<EPI, 1, false, EC>(p, stg, in, row0, nrows, col, b, lane);
            else epilogue_rows<EPI, 2, false, EC>(p, stg, in, row0, nrows, col, b, lane);
        }
    } else {
        if (nrows >= 32) epilogue_rows<EPI, 0, true, EC>(p, stg, in, row0, nrows, col, b, lane);
        else epilogue_rows<EPI, 0, false, EC>(p, stg, in, row0, nrows, col, b, lane);
    }
}

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                  const GemmDev p) {
    using L = SmemLayout<BN>;
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled tiles need a 1024-byte aligned base
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + L::BAR_OFFSET;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (L::STAGES + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * L::STAGES + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * L::STAGES + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * L::STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.mt * p.nt * p.nsplit;

    if (threadIdx.x == 0) stamp_begin(p, EPI | (A_MN ? 16 : 0) | (B_MN ? 32 : 0), BN);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < L::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 4);   // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, L::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (threadIdx.x == 0) stamp(p, ST_CLK_PROLOGUE);
    pdl_wait();                  // everything above overlapped the previous kernel of the stream; global memory from here on

    // tile -> (m block, n block, k range).  m fastest: concurrently running CTAs share few B (weight) tiles.
    auto tile_coords = [&](int t, int& m0, int& n0, int& kb_begin, int& kb_end) {
        const int mb = t % p.mt;
        const int rest = t / p.mt;
        const int nb = rest % p.nt;
        const int z = rest / p.nt;
        m0 = mb * BM;
        n0 = nb * BN;
        kb_begin = z * p.kb_per;
        kb_end = min(p.kb_total, kb_begin + p.kb_per);
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                int m0, n0, kb_begin, kb_end;
                tile_coords(t, m0, n0, kb_begin, kb_end);
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sA = smem_base + stage * L::STAGE_BYTES;
                    const uint32_t sB = sA + L::A_BYTES;
                    if (p.debug & 2) {
                        mbar_arrive(full_bar(stage));
                        if (++stage == L::STAGES) { stage = 0; phase ^= 1u; }
                        continue;
                    }
                    mbar_arrive_expect_tx(full_bar(stage), L::STAGE_BYTES);
                    const bool second = kb >= p.kb1;
                    const CUtensorMap* ma = second ? &tmA2 : &tmA;
                    const CUtensorMap* mb = second ? &tmB2 : &tmB;
                    const int k0 = (second ? kb - p.kb1 : kb) * BK;
                    if (!A_MN) {
                        tma_load_2d(sA, ma, full_bar(stage), k0, m0);  // box {64 k, 128 m}
                    } else {
#pragma unroll
                        for (int c = 0; c < BM / 64; ++c)  // box {64 m, 64 k}
                            tma_load_2d(sA + c * (BK * 128), ma, full_bar(stage), m0 + c * 64, k0);
                    }
                    if (!B_MN) {
                        tma_load_2d(sB, mb, full_bar(stage), k0, n0);  // box {64 k, BN n}
                    } else {
#pragma unroll
                        for (int c = 0; c < BN / 64; ++c)
                            tma_load_2d(sB + c * (BK * 128), mb, full_bar(stage), n0 + c * 64, k0);
                    }
                    if (t == (int)blockIdx.x && kb == kb_begin) stamp(p, ST_CLK_TMA0);
                    if (++stage == L::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
            pdl_launch_dependents();   // all operand loads issued: the next kernel's CTAs may start their prologue
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                int m0, n0, kb_begin, kb_end;
                tile_coords(t, m0, n0, kb_begin, kb_end);
                if (kb_end <= kb_begin) continue;
                const int ab = it & 1;
                const uint32_t aph = (uint32_t)(it >> 1) & 1u;
                ++it;
                mbar_wait(tempty_bar(ab), aph ^ 1u);   // epilogue has drained this accumulator buffer
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(ab * BN);
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    if (it == 1 && kb == kb_begin) stamp(p, ST_CLK_FULL0);
                    const uint32_t sA = smem_base + stage * L::STAGE_BYTES;
                    const uint32_t sB = sA + L::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        if (p.debug & 4) break;
                        // K-major: rows of 128 B, 8-row swizzle atoms 1024 B apart (SBO); a UMMA_K=16
                        //          slice is 32 B further along the row.
                        // MN-major: 64-element MN chunks BK*128 B apart (LBO), 8-k-row atoms 1024 B
                        //          apart (SBO); a UMMA_K=16 slice is 16 rows = 2048 B further.
                        const uint64_t da = A_MN ? umma_desc_sw128(sA + k * 2048, BK * 128, 1024)
                                                 : umma_desc_sw128(sA + k * 32, 16, 1024);
                        const uint64_t db = B_MN ? umma_desc_sw128(sB + k * 2048, BK * 128, 1024)
                                                 : umma_desc_sw128(sB + k * 32, 16, 1024);
                        umma_bf16(tacc, da, db, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
                    if (++stage == L::STAGES) { stage = 0; phase ^= 1u; }
                }
                umma_commit(tfull_bar(ab));  // accumulator complete
                if (it == 1) stamp(p, ST_CLK_MMA_ISSUED0);
            }
        }
    } else {
        // ===================== epilogue (4 warps) =====================
        const int quad = warp & 3;
        float* stg = reinterpret_cast<float*>(smem_gen + L::EPI_OFFSET) + (warp - 2) * 32 * EPI_LD;
        int it = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            int m0, n0, kb_begin, kb_end;
            tile_coords(t, m0, n0, kb_begin, kb_end);
            if (kb_end <= kb_begin) continue;
            const int ab = it & 1;
            const uint32_t aph = (uint32_t)(it >> 1) & 1u;
            ++it;
            const int row0 = m0 + quad * 32;
            const int nrows = min(32, p.M - row0);   // may be <= 0 for the ragged last tile
            const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ab * BN);
            EpiIn<EPI> cur;
            if (!(p.debug & 1)) cur.load(p, row0, nrows, n0, lane);   // pass 0: fetched while the main loop still runs
            mbar_wait(tfull_bar(ab), aph);
            tc_fence_after();
            if (it == 1 && warp == 2 && lane == 0) stamp(p, ST_CLK_TFULL0);
#pragma unroll 1
            for (int c = 0; c < BN / EPI_COLS; ++c) {
                // HBM inputs of this pass are requested before the TMEM read + transpose so that their latency overlaps
                if (c > 0 && !(p.debug & 1)) cur.load(p, row0, nrows, n0 + c * EPI_COLS, lane);
                uint32_t r[EPI_COLS];
                tmem_ld32(tacc + (uint32_t)(c * EPI_COLS), r);
                tmem_ld32(tacc + (uint32_t)(c * EPI_COLS + 32), r + 32);
                tmem_ld_wait();
                if (c == BN / EPI_COLS - 1) {   // all TMEM reads of this buffer are done: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(ab));
                }
                const int col0 = n0 + c * EPI_COLS;
                if (col0 >= p.N || nrows <= 0 || (p.debug & 1)) continue;   // warp-uniform
#pragma unroll
                for (int q = 0; q < EPI_COLS / 4; ++q)
                    *reinterpret_cast<float4*>(stg + stg_off(lane, q)) = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                                                   __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                __syncwarp();
                epilogue_block<EPI>(p, stg, cur, row0, nrows, col0, lane);
                __syncwarp();   // the transpose buffer is reused by the next pass
            }
            if (warp == 2 && lane == 0) stamp(p, it == 1 ? ST_CLK_EPI0 : ST_CLK_EPI_LAST);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) stamp_end(p, ((int)num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, L::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a 2-CTA cluster computes one 256 x 256 output tile per step of its tile loop.  Each CTA
// loads its own 128 rows of A and HALF (128 rows) of the B tile, so the operand traffic per FLOP is half that of the
// single-CTA 128 x 256 tile; the leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 (M = 256) that reads both
// CTAs' shared memory and writes both CTAs' tensor memory.  Same roles and epilogue as above; barriers:
//   full[s]   leader only: transaction bytes of all four loads of a stage (both CTAs signal the leader's barrier)
//   empty[s]  both CTAs:   multicast tcgen05.commit from the leader frees the stage in both
//   tfull[b]  both CTAs:   multicast commit when the accumulator of a tile is complete
//   tempty[b] leader only: 8 arrivals (4 epilogue warps x 2 CTAs) before the MMA thread may overwrite buffer b
// Two footprints of the same kernel (template parameter L):
//   PairWide  one CTA per SM: 5 stages, TWO accumulator buffers (512 TMEM columns), 8 epilogue warps, 64-column passes.
//             The epilogue paces this kernel at the model's shapes (K = 768: 3.5 us of MMA per 128 x 256 tile against 6-9 us
//             of epilogue with four warps, profiles/gemm_timeline_r1f.txt), hence the second set of four epilogue warps.
//   PairHalf  TWO CTAs per SM: 3 stages (96 KB), ONE accumulator buffer (256 TMEM columns), 4 epilogue warps, 32-column
//             passes (16 KB of transpose buffers), <= 168 registers.  The step is a stream of small dependent GEMMs on several
//             CUDA streams; with whole-SM CTAs every prologue, pipeline fill and epilogue of one launch keeps an SM's tensor core
//             idle and a waiting kernel's CTAs out (profiles/timeline_r1_s5.txt).  At half footprint the second slot of every SM
//             takes a CTA of another launch (the other chain, a weight-gradient product, or - with programmatic dependent launch -
//             the prologue of the next kernel of the same stream), whose main loop runs under this CTA's prologue / epilogue.
template <int STAGES_, int NACC_, int EPIW_, int EC_, int MINB_>
struct PairLayoutT {
    static constexpr int BN_CTA = 128;                  // B rows loaded per CTA
    static constexpr int BN_TILE = 256;                 // accumulator columns per CTA
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN_CTA * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = STAGES_;
    static constexpr int NACC = NACC_;                  // accumulator buffers in tensor memory
    static constexpr int EPI_WARPS = EPIW_;
    static constexpr int EC = EC_;                      // columns per epilogue pass
    static constexpr int MINB = MINB_;                  // CTAs per SM the footprint allows
    static constexpr int THREADS = 64 + 32 * EPIW_;
    static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int EPI_BYTES = EPI_WARPS * 32 * EC * 4;
    static constexpr int BAR_OFFSET = EPI_OFFSET + EPI_BYTES;
    // alignment slack: the dynamic shared-memory window starts 1024-byte aligned in practice; the half footprint has no room
    // for a kilobyte of slack (2 x 115,712 bytes per SM), so it keeps 64 bytes and the kernel traps if the base is misaligned
    static constexpr int TOTAL = BAR_OFFSET + 256 + (MINB_ > 1 ? 64 : 1024);
    static constexpr int TMEM_COLS = NACC * BN_TILE;
};
using PairLayout = PairLayoutT<5, 2, PAIR_EPI_WARPS, EPI_COLS, 1>;
using PairHalf = PairLayoutT<3, 1, 4, 32, 2>;
// 16 epilogue warps (four per scheduler) on 32-column passes: the fused epilogues are instruction-issue bound (ncu source page,
// profiles/gemm_epilogue_r2.txt: dropout hashes, GELU, packing - not memory), and two warps per scheduler leave half of the issue
// slots to dependency stalls; same shared-memory footprint as PairLayout (16 x 4 KB of transpose buffers), <= 113 registers.
using PairWide16 = PairLayoutT<5, 2, 16, 32, 1>;
static_assert(PairHalf::TOTAL <= 115712, "two CTAs of the half-footprint kernel must fit one SM (228 KB - 2 x 1 KB reserved)");

template <bool A_MN, bool B_MN, int EPI, class L = PairLayout>
__global__ void __launch_bounds__(L::THREADS, L::MINB)
gemm_bf16_tcgen05_pair(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                       const GemmDev p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    if (L::MINB > 1 && smem_base != smem_u32(smem_raw)) __trap();      // no slack in the half footprint (see PairLayoutT::TOTAL)
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + L::BAR_OFFSET;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (L::STAGES + s); };
    auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * L::STAGES + b); };
    auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * L::STAGES + 2 + b); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * L::STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();          // 0 = leader
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int num_tiles = p.mt * p.nt * p.nsplit;     // mt counts 256-row blocks here

    if (threadIdx.x == 0) stamp_begin(p, EPI | (A_MN ? 16 : 0) | (B_MN ? 32 : 0), 512);
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < L::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int b = 0; b < L::NACC; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 2 * L::EPI_WARPS);     // one arrive per epilogue warp of both CTAs
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc_2cta(tmem_slot, L::TMEM_COLS);
        tmem_relinquish_2cta();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // the peer's barriers are initialised before anything may signal them
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    if (threadIdx.x == 0) stamp(p, ST_CLK_PROLOGUE);
    // Everything above overlapped the previous kernel of the stream.  griddepcontrol.wait (pdl_wait) must precede the first access
    // to memory that kernel may have written: every role waits below - except that the producer first requests the B tiles
    // of its first pipeline stages when B is static (model weights), so that the pipeline fill runs under the predecessor's tail.
    const bool b_early = p.b_static != 0 && (warp == 0 && lane == 0);
    if (!b_early) pdl_wait();

    auto tile_coords = [&](int t, int& m0, int& n0, int& kb_begin, int& kb_end) {
        const int mb = t % p.mt;
        const int rest = t / p.mt;
        const int nb = rest % p.nt;
        const int z = rest / p.nt;
        m0 = mb * 2 * BM + (int)rank * BM;            // this CTA's 128 rows of the 256-row pair tile
        n0 = nb * L::BN_TILE;
        kb_begin = z * p.kb_per;
        kb_end = min(p.kb_total, kb_begin + p.kb_per);
    };

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            auto load_b = [&](int stage, int kb, int nb0) {
                const uint32_t sB = smem_base + stage * L::STAGE_BYTES + L::A_BYTES;
                const bool second = kb >= p.kb1;
                const CUtensorMap* mb = second ? &tmB2 : &tmB;
                const int k0 = (second ? kb - p.kb1 : kb) * BK;
                if (!B_MN) {
                    tma_load_2d_2cta(sB, mb, full_bar(stage), k0, nb0);
                } else {
#pragma unroll
                    for (int c = 0; c < L::BN_CTA / 64; ++c)
                        tma_load_2d_2cta(sB + c * (BK * 128), mb, full_bar(stage), nb0 + c * 64, k0);
                }
            };
            int early = 0;      // stages of the FIRST tile whose barrier is armed and whose B half is already in flight
            if (b_early) {
                if (pair < num_tiles) {
                    int m0, n0, kb_begin, kb_end;
                    tile_coords(pair, m0, n0, kb_begin, kb_end);
                    const int nb0 = n0 + (int)rank * L::BN_CTA;
                    for (int kb = kb_begin; kb < kb_end && early < L::STAGES; ++kb, ++early) {
                        if (rank == 0) mbar_arrive_expect_tx(full_bar(early), 2 * L::STAGE_BYTES);   // fresh barriers: no empty wait
                        load_b(early, kb, nb0);
                    }
                }
                pdl_wait();
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int t = pair; t < num_tiles; t += npairs) {
                int m0, n0, kb_begin, kb_end;
                tile_coords(t, m0, n0, kb_begin, kb_end);
                const int nb0 = n0 + (int)rank * L::BN_CTA;      // this CTA's half of the B tile
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    const bool pre = early > 0;                  // this stage was armed (and its B requested) before the wait
                    if (pre) --early;
                    else mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sA = smem_base + stage * L::STAGE_BYTES;
                    if (!pre && rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * L::STAGE_BYTES);
                    const bool second = kb >= p.kb1;
                    const CUtensorMap* ma = second ? &tmA2 : &tmA;
                    const int k0 = (second ? kb - p.kb1 : kb) * BK;
                    if (!A_MN) {
                        tma_load_2d_2cta(sA, ma, full_bar(stage), k0, m0);
                    } else {
#pragma unroll
                        for (int c = 0; c < BM / 64; ++c)
                            tma_load_2d_2cta(sA + c * (BK * 128), ma, full_bar(stage), m0 + c * 64, k0);
                    }
                    if (!pre) load_b(stage, kb, nb0);
                    if (t == pair && kb == kb_begin) stamp(p, ST_CLK_TMA0);
                    if (++stage == L::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
            pdl_launch_dependents();   // all operand loads issued: the next kernel's CTAs may start their prologue
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread of the LEADER CTA) =====================
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, L::BN_TILE, A_MN, B_MN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = pair; t < num_tiles; t += npairs) {
                int m0, n0, kb_begin, kb_end;
                tile_coords(t, m0, n0, kb_begin, kb_end);
                if (kb_end <= kb_begin) continue;
                const int ab = it % L::NACC;
                const uint32_t aph = (uint32_t)(it / L::NACC) & 1u;
                ++it;
                mbar_wait(tempty_bar(ab), aph ^ 1u);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(ab * L::BN_TILE);
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    if (it == 1 && kb == kb_begin) stamp(p, ST_CLK_FULL0);
                    const uint32_t sA = smem_base + stage * L::STAGE_BYTES;
                    const uint32_t sB = sA + L::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t da = A_MN ? umma_desc_sw128(sA + k * 2048, BK * 128, 1024)
                                                 : umma_desc_sw128(sA + k * 32, 16, 1024);
                        const uint64_t db = B_MN ? umma_desc_sw128(sB + k * 2048, BK * 128, 1024)
                                                 : umma_desc_sw128(sB + k * 32, 16, 1024);
                        umma_bf16_2cta(tacc, da, db, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
                    }
                    umma_commit_2cta_mc(empty_bar(stage));     // frees the stage in both CTAs
                    if (++stage == L::STAGES) { stage = 0; phase ^= 1u; }
                }
                umma_commit_2cta_mc(tfull_bar(ab));            // accumulator complete, in both CTAs
                if (it == 1) stamp(p, ST_CLK_MMA_ISSUED0);
            }
        }
    } else {
        // ===================== epilogue (8 or 4 warps in each CTA) =====================
        // warp w reads TMEM lanes 32 (w % 4) .. +31 (hardware rule); with 8 warps, warps 2-5 take the first half of the tile's
        // column passes, warps 6-9 the second half
        constexpr int EC = L::EC;
        const int quad = warp & 3;
        const int chalf = (warp - 2) >> 2;
        constexpr int PASSES = L::BN_TILE / EC / (L::EPI_WARPS / 4);       // column passes per warp
        const int c_first = chalf * PASSES;
        float* stg = reinterpret_cast<float*>(smem_gen + L::EPI_OFFSET) + (warp - 2) * 32 * EC;
        int it = 0;
        for (int t = pair; t < num_tiles; t += npairs) {
            int m0, n0, kb_begin, kb_end;
            tile_coords(t, m0, n0, kb_begin, kb_end);
            if (kb_end <= kb_begin) continue;
            const int ab = it % L::NACC;
            const uint32_t aph = (uint32_t)(it / L::NACC) & 1u;
            ++it;
            const int row0 = m0 + quad * 32;
            const int nrows = min(32, p.M - row0);
            const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ab * L::BN_TILE);
            EpiIn<EPI, EC> cur;
            cur.load(p, row0, nrows, n0 + c_first * EC, lane);
            mbar_wait(tfull_bar(ab), aph);
            tc_fence_after();
            if (it == 1 && warp == 2 && lane == 0) stamp(p, ST_CLK_TFULL0);
#pragma unroll 1
            for (int c = c_first; c < c_first + PASSES; ++c) {
                if (c > c_first) cur.load(p, row0, nrows, n0 + c * EC, lane);
                uint32_t r[EC];
                tmem_ld32(tacc + (uint32_t)(c * EC), r);
                if (EC > 32) tmem_ld32(tacc + (uint32_t)(c * EC + 32), r + 32);
                tmem_ld_wait();
                if (c == c_first + PASSES - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(tempty_bar(ab));
                }
                const int col0 = n0 + c * EC;
                if (col0 >= p.N || nrows <= 0) continue;
#pragma unroll
                for (int q = 0; q < EC / 4; ++q)
                    *reinterpret_cast<float4*>(stg + stg_off<EC>(lane, q)) = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                                                       __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                __syncwarp();
                epilogue_block<EPI, EC>(p, stg, cur, row0, nrows, col0, lane);
                __syncwarp();
            }
            if (warp == 2 && lane == 0) stamp(p, it == 1 ? ST_CLK_EPI0 : ST_CLK_EPI_LAST);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) stamp_end(p, (num_tiles - pair + npairs - 1) / npairs);
    cluster_sync_all();          // neither CTA may leave (or free tensor memory) while the pair's MMAs / signals are in flight
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, L::TMEM_COLS);
    }
}

typedef void (*GemmKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemmDev);

// the (layout, epilogue) combinations the model uses get a specialised kernel; everything else is GENERIC
template <int BN>
inline GemmKernel pick_kernel_bn(int a_mn, int b_mn, int epi) {
    if (!a_mn && !b_mn) {
        switch (epi) {
            case EPI_BF16: return gemm_bf16_tcgen05<BN, false, false, EPI_BF16>;
            case EPI_ACT: return gemm_bf16_tcgen05<BN, false, false, EPI_ACT>;
            case EPI_RES: return gemm_bf16_tcgen05<BN, false, false, EPI_RES>;
            case EPI_F32: return gemm_bf16_tcgen05<BN, false, false, EPI_F32>;
            default: return gemm_bf16_tcgen05<BN, false, false, EPI_GENERIC>;
        }
    }
    if (!a_mn && b_mn) {
        switch (epi) {
            case EPI_DACT: return gemm_bf16_tcgen05<BN, false, true, EPI_DACT>;
            case EPI_RES: return gemm_bf16_tcgen05<BN, false, true, EPI_RES>;
            case EPI_F32: return gemm_bf16_tcgen05<BN, false, true, EPI_F32>;
            case EPI_ACC: return gemm_bf16_tcgen05<BN, false, true, EPI_ACC>;
            default: return gemm_bf16_tcgen05<BN, false, true, EPI_GENERIC>;
        }
    }
    if (a_mn && b_mn) {
        switch (epi) {
            case EPI_ACC: return gemm_bf16_tcgen05<BN, true, true, EPI_ACC>;
            case EPI_F32: return gemm_bf16_tcgen05<BN, true, true, EPI_F32>;
            default: return gemm_bf16_tcgen05<BN, true, true, EPI_GENERIC>;
        }
    }
    return gemm_bf16_tcgen05<BN, true, false, EPI_GENERIC>;
}

// CTA-pair kernels: instantiated for the (layout, epilogue) combinations of the model's large products
inline GemmKernel pick_kernel_pair(int a_mn, int b_mn, int epi) {
    if (!a_mn && !b_mn) {
        switch (epi) {
            case EPI_BF16: return gemm_bf16_tcgen05_pair<false, false, EPI_BF16>;
            case EPI_ACT: return gemm_bf16_tcgen05_pair<false, false, EPI_ACT>;
            case EPI_RES: return gemm_bf16_tcgen05_pair<false, false, EPI_RES>;
            case EPI_F32: return gemm_bf16_tcgen05_pair<false, false, EPI_F32>;
            default: return nullptr;
        }
    }
    if (!a_mn && b_mn) {
        switch (epi) {
            case EPI_DACT: return gemm_bf16_tcgen05_pair<false, true, EPI_DACT>;
            case EPI_RES: return gemm_bf16_tcgen05_pair<false, true, EPI_RES>;
            case EPI_F32: return gemm_bf16_tcgen05_pair<false, true, EPI_F32>;
            default: return nullptr;
        }
    }
    if (a_mn && b_mn && epi == EPI_ACC) return gemm_bf16_tcgen05_pair<true, true, EPI_ACC>;
    return nullptr;
}

// half-footprint CTA-pair kernels (two CTAs per SM): same (layout, epilogue) set
inline GemmKernel pick_kernel_pair_half(int a_mn, int b_mn, int epi) {
    if (!a_mn && !b_mn) {
        switch (epi) {
            case EPI_BF16: return gemm_bf16_tcgen05_pair<false, false, EPI_BF16, PairHalf>;
            case EPI_ACT: return gemm_bf16_tcgen05_pair<false, false, EPI_ACT, PairHalf>;
            case EPI_RES: return gemm_bf16_tcgen05_pair<false, false, EPI_RES, PairHalf>;
            case EPI_F32: return gemm_bf16_tcgen05_pair<false, false, EPI_F32, PairHalf>;
            default: return nullptr;
        }
    }
    if (!a_mn && b_mn) {
        switch (epi) {
            case EPI_DACT: return gemm_bf16_tcgen05_pair<false, true, EPI_DACT, PairHalf>;
            case EPI_RES: return gemm_bf16_tcgen05_pair<false, true, EPI_RES, PairHalf>;
            case EPI_F32: return gemm_bf16_tcgen05_pair<false, true, EPI_F32, PairHalf>;
            default: return nullptr;
        }
    }
    if (a_mn && b_mn && epi == EPI_ACC) return gemm_bf16_tcgen05_pair<true, true, EPI_ACC, PairHalf>;
    return nullptr;
}

inline GemmKernel pick_kernel_pair16(int a_mn, int b_mn, int epi) {
    if (!a_mn && !b_mn) {
        switch (epi) {
            case EPI_BF16: return gemm_bf16_tcgen05_pair<false, false, EPI_BF16, PairWide16>;
            case EPI_ACT: return gemm_bf16_tcgen05_pair<false, false, EPI_ACT, PairWide16>;
            case EPI_RES: return gemm_bf16_tcgen05_pair<false, false, EPI_RES, PairWide16>;
            case EPI_F32: return gemm_bf16_tcgen05_pair<false, false, EPI_F32, PairWide16>;
            default: return nullptr;
        }
    }
    if (!a_mn && b_mn) {
        switch (epi) {
            case EPI_DACT: return gemm_bf16_tcgen05_pair<false, true, EPI_DACT, PairWide16>;
            case EPI_RES: return gemm_bf16_tcgen05_pair<false, true, EPI_RES, PairWide16>;
            case EPI_F32: return gemm_bf16_tcgen05_pair<false, true, EPI_F32, PairWide16>;
            default: return nullptr;
        }
    }
    if (a_mn && b_mn && epi == EPI_ACC) return gemm_bf16_tcgen05_pair<true, true, EPI_ACC, PairWide16>;
    return nullptr;
}

// true when pick_kernel_bn has a specialised kernel for this combination
inline bool has_specialised(int a_mn, int b_mn, int epi) {
    if (epi == EPI_GENERIC) return true;
    if (!a_mn && !b_mn) return epi == EPI_BF16 || epi == EPI_ACT || epi == EPI_RES || epi == EPI_F32;
    if (!a_mn && b_mn) return epi == EPI_DACT || epi == EPI_RES || epi == EPI_F32 || epi == EPI_ACC;
    if (a_mn && b_mn) return epi == EPI_ACC || epi == EPI_F32;
    return false;
}

}  // namespace gemmk

// defined in gemm_inst{64,128,256}.cu
gemmk::GemmKernel gemm_pick_kernel_64(int a_mn, int b_mn, int epi);
gemmk::GemmKernel gemm_pick_kernel_128(int a_mn, int b_mn, int epi);
gemmk::GemmKernel gemm_pick_kernel_256(int a_mn, int b_mn, int epi);
gemmk::GemmKernel gemm_pick_kernel_pair(int a_mn, int b_mn, int epi);   // gemm_inst_pair.cu
gemmk::GemmKernel gemm_pick_kernel_pair_half(int a_mn, int b_mn, int epi);   // gemm_inst_pair_half.cu
gemmk::GemmKernel gemm_pick_kernel_pair16(int a_mn, int b_mn, int epi);       // gemm_inst_pair16.cu
