// tcgen05 GEMM for sm_100a: TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring ->
// tcgen05.mma (single elected thread, fp32 accumulator in TMEM) -> tcgen05.ld epilogue with fused
// bias / GELU / quick-GELU / activation-derivative / residual / dual output.
//
// Replaces, for the MarT hot path, every nn.Linear product of the reference model
// (MarT/models/modeling_unimo.py:212-287 CLIP attention/MLP, :317-478 BERT attention/FFN,
// :956-976 MLM head) and their autograd transposes (dX = dY W, dW = dY^T X).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (warp w owns TMEM lanes 32*(w%4) .. +31, i.e. 32 rows of the 128-row tile).
#include "gemm.cuh"
#include <stdarg.h>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int GEMM_THREADS = 192;

struct GemmDev {
    int M, N;
    int kb1, kb_total;  // k-blocks served by pair 1 / total
    float alpha;
    const float* bias;
    const float* bias2;
    bf16* out_preact; long ldp;
    int act;
    const bf16* dact_src; long ldx; int dact;
    const float* residual; long ldr;
    float* out_f32; long ldo32;
    bf16* out_bf16; long ldo16;
    int accumulate;
};

template <int BN>
struct SmemLayout {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN >= 256) ? 4 : 6;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // barriers + alignment slack
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                  const GemmDev p) {
    using L = SmemLayout<BN>;
    extern __shared__ uint8_t smem_raw[];
    // 128B-swizzled tiles need a 1024-byte aligned base
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + L::BAR_OFFSET;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (L::STAGES + s); };
    const uint32_t accum_bar = bar_base + 8u * (2 * L::STAGES);
    const uint32_t tmem_slot = bar_base + 8u * (2 * L::STAGES + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;

    // split-K: this CTA contracts k-blocks [kb_begin, kb_end)
    const int nsplit = gridDim.z;
    const int kb_per = (p.kb_total + nsplit - 1) / nsplit;
    const int kb_begin = blockIdx.z * kb_per;
    const int kb_end = min(p.kb_total, kb_begin + kb_per);
    const int nkb = kb_end - kb_begin;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < L::STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(accum_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (nkb > 0) {
        if (warp == 0) {
            // ===================== TMA producer =====================
            if (lane == 0) {
                int stage = 0;
                uint32_t phase = 0;
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sA = smem_base + stage * L::STAGE_BYTES;
                    const uint32_t sB = sA + L::A_BYTES;
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
                    if (++stage == L::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        } else if (warp == 1) {
            // ===================== MMA issuer (one thread) =====================
            if (lane == 0) {
                constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
                int stage = 0;
                uint32_t phase = 0;
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sA = smem_base + stage * L::STAGE_BYTES;
                    const uint32_t sB = sA + L::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // K-major: rows of 128 B, 8-row swizzle atoms 1024 B apart (SBO); a UMMA_K=16
                        //          slice is 32 B further along the row.
                        // MN-major: 64-element MN chunks BK*128 B apart (LBO), 8-k-row atoms 1024 B
                        //          apart (SBO); a UMMA_K=16 slice is 16 rows = 2048 B further.
                        const uint64_t da = A_MN ? umma_desc_sw128(sA + k * 2048, BK * 128, 1024)
                                                 : umma_desc_sw128(sA + k * 32, 16, 1024);
                        const uint64_t db = B_MN ? umma_desc_sw128(sB + k * 2048, BK * 128, 1024)
                                                 : umma_desc_sw128(sB + k * 32, 16, 1024);
                        umma_bf16(tmem_base, da, db, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
                    if (++stage == L::STAGES) { stage = 0; phase ^= 1u; }
                }
                umma_commit(accum_bar);  // accumulator complete
            }
        } else {
            // ===================== epilogue =====================
            mbar_wait(accum_bar, 0);
            tc_fence_after();
            const int quad = warp & 3;
            const int row = m0 + quad * 32 + lane;
            const bool row_ok = row < p.M;
            const bool vec_ok = (n0 + BN <= p.N);
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(c * 32), r);
                tmem_ld_wait();
                const int col0 = n0 + c * 32;
                if (row_ok && col0 < p.N) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
                const int ncols = min(32, p.N - col0);
                if (p.bias) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < ncols) v[j] += __ldg(p.bias + col0 + j);
                }
                if (p.bias2) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < ncols) v[j] += __ldg(p.bias2 + col0 + j);
                }
                if (p.out_preact) {
                    bf16* dst = p.out_preact + (long)row * p.ldp + col0;
                    if (vec_ok && (p.ldp & 7) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 q = make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]),
                                                 pack_bf16x2(v[j + 4], v[j + 5]), pack_bf16x2(v[j + 6], v[j + 7]));
                            *reinterpret_cast<uint4*>(dst + j) = q;
                        }
                    } else {
                        for (int j = 0; j < ncols; ++j) dst[j] = __float2bfloat16(v[j]);
                    }
                }
                if (p.act == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
                } else if (p.act == 2) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
                }
                if (p.dact) {
                    const bf16* src = p.dact_src + (long)row * p.ldx + col0;
                    if (vec_ok && (p.ldx & 7) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 q = *reinterpret_cast<const uint4*>(src + j);
                            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                float2 f = __bfloat1622float2(h[t]);
                                v[j + 2 * t] *= (p.dact == 1) ? gelu_erf_grad(f.x) : quick_gelu_grad(f.x);
                                v[j + 2 * t + 1] *= (p.dact == 1) ? gelu_erf_grad(f.y) : quick_gelu_grad(f.y);
                            }
                        }
                    } else {
                        for (int j = 0; j < ncols; ++j) {
                            float x = __bfloat162float(src[j]);
                            v[j] *= (p.dact == 1) ? gelu_erf_grad(x) : quick_gelu_grad(x);
                        }
                    }
                }
                if (p.residual) {
                    const float* src = p.residual + (long)row * p.ldr + col0;
                    if (vec_ok && (p.ldr & 3) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4 f = *reinterpret_cast<const float4*>(src + j);
                            v[j] += f.x; v[j + 1] += f.y; v[j + 2] += f.z; v[j + 3] += f.w;
                        }
                    } else {
                        for (int j = 0; j < ncols; ++j) v[j] += src[j];
                    }
                }
                if (p.out_f32) {
                    float* dst = p.out_f32 + (long)row * p.ldo32 + col0;
                    if (p.accumulate) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < ncols) atomicAdd(dst + j, v[j]);
                    } else if (vec_ok && (p.ldo32 & 3) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
                        for (int j = 0; j < ncols; ++j) dst[j] = v[j];
                    }
                }
                if (p.out_bf16) {
                    bf16* dst = p.out_bf16 + (long)row * p.ldo16 + col0;
                    if (vec_ok && (p.ldo16 & 7) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 q = make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]),
                                                 pack_bf16x2(v[j + 4], v[j + 5]), pack_bf16x2(v[j + 6], v[j + 7]));
                            *reinterpret_cast<uint4*>(dst + j) = q;
                        }
                    } else {
                        for (int j = 0; j < ncols; ++j) dst[j] = __float2bfloat16(v[j]);
                    }
                }
                }  // row_ok
                __syncwarp();  // tcgen05.ld is warp-collective: reconverge before the next chunk
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN);
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
    return fn;
}

// 2D bf16 tensor map over a row-major [rows][inner] array with leading dimension ld (elements);
// box = {64 inner elements (128 B), box_rows}; 128B swizzle; out-of-bounds elements read as zero.
int encode_map(CUtensorMap* map, const bf16* base, long inner, long rows, long ld, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    B2_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
    B2_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA operand base %p not 16-byte aligned", (const void*)base);
    B2_CHECK((ld * 2) % 16 == 0, "TMA operand leading dimension %ld elements is not a multiple of 16 bytes", ld);
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B2_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (inner %ld rows %ld ld %ld)", (int)r,
             inner, rows, ld);
    return 0;
}

typedef void (*GemmKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemmDev);

template <int BN>
GemmKernel pick_kernel(int a_mn, int b_mn) {
    if (!a_mn && !b_mn) return gemm_bf16_tcgen05<BN, false, false>;
    if (!a_mn && b_mn) return gemm_bf16_tcgen05<BN, false, true>;
    if (a_mn && b_mn) return gemm_bf16_tcgen05<BN, true, true>;
    return gemm_bf16_tcgen05<BN, true, false>;
}

}  // namespace

struct GemmPlan {
    CUtensorMap tmA, tmB, tmA2, tmB2;
    GemmDev dev;
    dim3 grid;
    int smem;
    GemmKernel kernel;
    double flops;
};

int gemm_plan_create(const GemmOp& op, GemmPlan** out) {
    *out = nullptr;
    B2_CHECK(op.M > 0 && op.N > 0 && op.K1 > 0, "gemm: bad shape M=%d N=%d K1=%d", op.M, op.N, op.K1);
    B2_CHECK(op.K2 == 0 || op.K1 % BK == 0, "gemm: K1=%d must be a multiple of %d when a second operand pair is used", op.K1, BK);
    B2_CHECK(op.split_k >= 1, "gemm: split_k must be >= 1");
    B2_CHECK(op.split_k == 1 || (op.accumulate && op.out_f32 && !op.out_bf16 && !op.out_preact && !op.bias && !op.bias2 &&
                                 !op.residual && op.act == 0 && op.dact == 0),
             "gemm: split-K needs a pure fp32 accumulate epilogue");
    const int bn = op.bn ? op.bn : 128;
    B2_CHECK(bn == 64 || bn == 128 || bn == 256, "gemm: unsupported N tile %d", bn);
    GemmPlan* p = new GemmPlan();
    memset(p, 0, sizeof(*p));
    int rc = 0;
    // operand maps: K-major -> [rows][K] inner = K; MN-major -> [K][rows] inner = rows
    auto enc = [&](CUtensorMap* m, const bf16* base, int rows, int K, long ld, int mn, int tile_rows) {
        if (!mn) return encode_map(m, base, K, rows, ld, tile_rows);
        return encode_map(m, base, rows, K, ld, BK);
    };
    rc = enc(&p->tmA, op.A, op.M, op.K1, op.lda, op.a_mn, BM);
    if (!rc) rc = enc(&p->tmB, op.B, op.N, op.K1, op.ldb, op.b_mn, bn);
    if (!rc && op.K2 > 0) {
        rc = enc(&p->tmA2, op.A2, op.M, op.K2, op.lda2, op.a_mn, BM);
        if (!rc) rc = enc(&p->tmB2, op.B2, op.N, op.K2, op.ldb2, op.b_mn, bn);
    } else if (!rc) {
        p->tmA2 = p->tmA;
        p->tmB2 = p->tmB;
    }
    if (rc) { delete p; return rc; }
    GemmDev& d = p->dev;
    d.M = op.M; d.N = op.N;
    d.kb1 = (op.K1 + BK - 1) / BK;
    d.kb_total = d.kb1 + (op.K2 + BK - 1) / BK;
    d.alpha = op.alpha; d.bias = op.bias; d.bias2 = op.bias2;
    d.out_preact = op.out_preact; d.ldp = op.ldp;
    d.act = op.act; d.dact_src = op.dact_src; d.ldx = op.ldx; d.dact = op.dact_src ? op.dact : 0;
    d.residual = op.residual; d.ldr = op.ldr;
    d.out_f32 = op.out_f32; d.ldo32 = op.ldo32; d.out_bf16 = op.out_bf16; d.ldo16 = op.ldo16;
    d.accumulate = op.accumulate;
    int split = op.split_k;
    if (split > d.kb_total) split = d.kb_total;
    p->grid = dim3((op.N + bn - 1) / bn, (op.M + BM - 1) / BM, split);
    if (bn == 64) { p->kernel = pick_kernel<64>(op.a_mn, op.b_mn); p->smem = SmemLayout<64>::TOTAL; }
    else if (bn == 128) { p->kernel = pick_kernel<128>(op.a_mn, op.b_mn); p->smem = SmemLayout<128>::TOTAL; }
    else { p->kernel = pick_kernel<256>(op.a_mn, op.b_mn); p->smem = SmemLayout<256>::TOTAL; }
    cudaError_t e = cudaFuncSetAttribute((const void*)p->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem);
    if (e != cudaSuccess) {
        const int smem = p->smem;
        delete p;
        return b200mart_set_err(3, "cudaFuncSetAttribute(smem=%d) failed: %s", smem, cudaGetErrorString(e));
    }
    p->flops = 2.0 * op.M * (double)op.N * (double)(op.K1 + op.K2);
    *out = p;
    return 0;
}

void gemm_plan_destroy(GemmPlan* p) { delete p; }

double gemm_plan_flops(const GemmPlan* p) { return p->flops; }

int gemm_plan_launch(const GemmPlan* p, cudaStream_t stream) {
    p->kernel<<<p->grid, GEMM_THREADS, p->smem, stream>>>(p->tmA, p->tmB, p->tmA2, p->tmB2, p->dev);
    B2_LAUNCHED();
    return 0;
}
