// Host side of the persistent tcgen05 GEMM (device code: gemm_kernel.cuh, instantiated in gemm_inst*.cu):
// tensor-map encoding, epilogue-kind / tile selection, plan objects.
#include "gemm_kernel.cuh"
#include <stdarg.h>
#include <stdlib.h>

using namespace gemmk;

namespace {

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

// persistent-grid cap (B200MART_GEMM_MAXCTAS, bring-up): fewer CTAs per GEMM leave SMs to the kernels of the other chains
int grid_cap() {
    static int cap = -1;
    if (cap < 0) { const char* e = getenv("B200MART_GEMM_MAXCTAS"); cap = e ? atoi(e) : 0; }
    return cap;
}

int num_sms() {
    static int n = 0;
    if (n) return n;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
        n = 148;
    return n;
}

}  // namespace

struct GemmPlan {
    CUtensorMap tmA, tmB, tmA2, tmB2;
    GemmDev dev;
    dim3 grid;
    int smem;
    GemmKernel kernel;
    double flops;
    bool pair;      // CTA-pair kernel: launched as 2-CTA clusters
    int threads;    // CTA size of the pair kernel (wide: 320, half footprint: 192)
};

int gemm_plan_create(const GemmOp& op, GemmPlan** out) {
    *out = nullptr;
    B2_CHECK(op.M > 0 && op.N > 0 && op.K1 > 0, "gemm: bad shape M=%d N=%d K1=%d", op.M, op.N, op.K1);
    B2_CHECK(op.K2 == 0 || op.K1 % BK == 0, "gemm: K1=%d must be a multiple of %d when a second operand pair is used", op.K1, BK);
    B2_CHECK(op.split_k >= 1, "gemm: split_k must be >= 1");
    B2_CHECK(op.split_k == 1 || (op.accumulate && op.out_f32 && !op.out_bf16 && !op.out_preact && !op.bias && !op.bias2 &&
                                 !op.residual && op.act == 0 && op.dact == 0),
             "gemm: split-K needs a pure fp32 accumulate epilogue");
    // ---- fused-epilogue kind
    int epi = EPI_GENERIC;
    // the specialised epilogues move 4 columns per lane: 16-byte fp32 / 8-byte bf16 vector accesses
    auto al = [](const void* q, uintptr_t a) { return (reinterpret_cast<uintptr_t>(q) & (a - 1)) == 0; };
    const bool even = (op.N % 4 == 0) && (!op.out_preact || (op.ldp % 4 == 0 && al(op.out_preact, 8))) &&
                      (!op.dact_src || (op.ldx % 4 == 0 && al(op.dact_src, 8))) &&
                      (!op.residual || (op.ldr % 4 == 0 && al(op.residual, 16))) &&
                      (!op.out_f32 || (op.ldo32 % 4 == 0 && al(op.out_f32, 16))) &&
                      (!op.out_bf16 || (op.ldo16 % 4 == 0 && al(op.out_bf16, 8))) && (!op.bias || al(op.bias, 16)) &&
                      (!op.bias2 || al(op.bias2, 16)) && (!op.colsum_out || al(op.colsum_out, 16)) &&
                      (!op.colsum_out2 || al(op.colsum_out2, 16));
    const bool has_dact = op.dact_src && op.dact;
    if (even) {
        if (op.accumulate && op.out_f32 && !op.out_bf16 && !op.out_preact && !op.bias && !op.bias2 && !op.residual &&
            op.act == 0 && !has_dact)
            epi = EPI_ACC;
        else if (!op.accumulate && has_dact && op.dact == 3 && op.out_bf16 && !op.out_f32 && !op.out_preact && !op.bias && !op.bias2 &&
                 !op.residual && op.act == 0)
            epi = EPI_DACT;
        else if (!op.accumulate && !has_dact && op.out_preact && op.out_bf16 && !op.out_f32 && !op.residual && op.bias &&
                 (op.act == 1 || op.act == 2))
            epi = EPI_ACT;
        else if (!op.accumulate && !has_dact && !op.out_preact && op.act == 0 && !op.bias2 && op.residual && op.out_f32)
            epi = EPI_RES;
        else if (!op.accumulate && !has_dact && !op.out_preact && op.act == 0 && !op.bias2 && !op.residual && op.out_f32 &&
                 !op.out_bf16 && !op.drop.thresh)
            epi = EPI_F32;
        else if (!op.accumulate && !has_dact && !op.out_preact && op.act == 0 && !op.residual && op.out_bf16 && !op.out_f32 &&
                 !op.drop.thresh)
            epi = EPI_BF16;
    }
    if (!has_specialised(op.a_mn, op.b_mn, epi)) epi = EPI_GENERIC;
    { const char* e = getenv("B200MART_GEMM_GENERIC"); if (e && atoi(e)) epi = EPI_GENERIC; }
    B2_CHECK(!(op.colsum_out || op.colsum_out2) || epi == EPI_DACT,
             "gemm: column sums are only produced by the saved-derivative (dact = 3) bf16 epilogue of a K-major A / MN-major B "
             "product with N and the leading dimensions multiples of 4 and 16-byte aligned column-sum vectors");
    // ---- N tile: rounds of the persistent grid x per-tile cost (main loop ~ bn, plus a fixed part); ties -> wider tile
    int bn = op.bn;
    if (bn == 0) {
        long best = -1;
        for (int cand : {64, 128, 256}) {
            const long tiles = (long)((op.M + BM - 1) / BM) * ((op.N + cand - 1) / cand) * op.split_k;
            const long rounds = (tiles + num_sms() - 1) / num_sms();
            const long cost = rounds * (cand + 64);
            if (best < 0 || cost <= best) { best = cost; bn = cand; }
        }
    }
    {   // automatic choice prefers the CTA-pair (256 x 256) kernel wherever one exists and the problem has >= 2 x 2 CTA
        // tiles: measured +4.5 % on the whole training step (half the CTAs per GEMM, half the operand traffic per FLOP).
        // B200MART_GEMM_PAIR=0 switches it off (A/B runs).
        static int pref = -1;
        if (pref < 0) { const char* e = getenv("B200MART_GEMM_PAIR"); pref = e ? atoi(e) : 1; }
        // B200MART_GEMM_SMALLN=n (default 0 = off, A/B switch for the next tuning step): products with N <= n keep the single-CTA
        // kernel and the tile the cost model picked (128 x 128 at N = 768: 96-150 CTAs instead of 48-78).  Their epilogue is bound by
        // the memory system (profiles/gemm_timeline_r1g.txt: 5.8-7.4 us for a 128 x 256 RES tile per CTA at ~56 GB/s per SM), so
        // spreading it over twice the SMs may beat the pair kernel's lower operand traffic.
        static int smalln = -1;
        if (smalln < 0) { const char* e = getenv("B200MART_GEMM_SMALLN"); smalln = e ? atoi(e) : 0; }
        const bool keep_single = smalln > 0 && op.N <= smalln && op.split_k == 1;
        if (pref && !keep_single && op.bn == 0 && op.N >= 512 && op.M >= 512 && gemm_pick_kernel_pair(op.a_mn, op.b_mn, epi) != nullptr) bn = 512;
    }
    B2_CHECK(bn == 64 || bn == 128 || bn == 256 || bn == 512, "gemm: unsupported N tile %d (64, 128, 256, or 512 = 256x256 CTA pair)", bn);
    const bool pair = (bn == 512);
    if (pair) {
        B2_CHECK(gemm_pick_kernel_pair(op.a_mn, op.b_mn, epi) != nullptr,
                 "gemm: no CTA-pair kernel for layout (%d,%d) / epilogue kind %d", op.a_mn, op.b_mn, epi);
        bn = 128;    // rows of B each CTA of the pair loads
    }
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
    d.act = op.act; d.save_grad = op.save_grad; d.dact_src = op.dact_src; d.ldx = op.ldx; d.dact = op.dact_src ? op.dact : 0;
    d.residual = op.residual; d.ldr = op.ldr;
    d.out_f32 = op.out_f32; d.ldo32 = op.ldo32; d.out_bf16 = op.out_bf16; d.ldo16 = op.ldo16;
    d.accumulate = op.accumulate;
    d.colsum_out = op.colsum_out; d.colsum_out2 = op.colsum_out2;
    d.drop = op.drop;
    { const char* e = getenv("B200MART_GEMM_DEBUG"); d.debug = e ? atoi(e) : 0; }
    { static int be = -1; if (be < 0) { const char* e = getenv("B200MART_GEMM_BEARLY"); be = e ? atoi(e) : 1; } d.b_static = be ? op.b_static : 0; }
    int split = op.split_k;
    if (split > d.kb_total) split = d.kb_total;
    d.kb_per = (d.kb_total + split - 1) / split;
    split = (d.kb_total + d.kb_per - 1) / d.kb_per;   // no empty k-ranges
    d.nsplit = split;
    d.mt = (op.M + BM - 1) / BM;
    d.nt = (op.N + bn - 1) / bn;
    const long tiles = (long)d.mt * d.nt * split;
    long cap = num_sms();
    if (grid_cap() > 0 && grid_cap() < cap) cap = grid_cap();
    if (op.max_ctas > 0 && op.max_ctas < cap) cap = op.max_ctas;
    p->grid = dim3((unsigned)(tiles < cap ? tiles : cap), 1, 1);
    p->pair = pair;
    if (pair) {
        // B200MART_GEMM_HALF=1: the half-footprint variant (two CTAs per SM; see PairLayoutT in gemm_kernel.cuh).
        // B200MART_GEMM_HALF_SLOTS = CTAs per SM ONE launch may occupy (default 2: a launch with more tiles than SM pairs runs them
        // side by side, the main loop of one under the epilogue of the other; 1 leaves the second slot to other launches).
        static int half = -1, slots = -1;
        if (half < 0) { const char* e = getenv("B200MART_GEMM_HALF"); half = e ? atoi(e) : 0; if (half == 1) half = 0xffff; }   // 1 = every kind, else bitmask by epilogue kind
        if (slots < 0) { const char* e = getenv("B200MART_GEMM_HALF_SLOTS"); slots = e ? atoi(e) : 2; if (slots < 1) slots = 1; if (slots > 2) slots = 2; }
        const bool use_half = ((half >> epi) & 1) && gemm_pick_kernel_pair_half(op.a_mn, op.b_mn, epi) != nullptr;
        d.mt = (op.M + 2 * BM - 1) / (2 * BM);
        d.nt = (op.N + PairLayout::BN_TILE - 1) / PairLayout::BN_TILE;
        const long ptiles = (long)d.mt * d.nt * split;
        long pairs = cap / 2;
        if (use_half && !(op.max_ctas > 0 && op.max_ctas < num_sms()) && !(grid_cap() > 0)) pairs *= slots;
        p->grid = dim3((unsigned)(2 * (ptiles < pairs ? ptiles : pairs)), 1, 1);
        static int epi16 = -1;      // B200MART_GEMM_EPI16: 16 epilogue warps (PairWide16); bitmask by epilogue kind, -1 = all
        if (epi16 == -1) { const char* e = getenv("B200MART_GEMM_EPI16"); epi16 = e ? atoi(e) : 0; if (epi16 < 0) epi16 = 0xffff; }
        const bool use16 = !use_half && ((epi16 >> epi) & 1) && gemm_pick_kernel_pair16(op.a_mn, op.b_mn, epi) != nullptr;
        p->kernel = use_half ? gemm_pick_kernel_pair_half(op.a_mn, op.b_mn, epi)
                             : (use16 ? gemm_pick_kernel_pair16(op.a_mn, op.b_mn, epi) : gemm_pick_kernel_pair(op.a_mn, op.b_mn, epi));
        p->smem = use_half ? PairHalf::TOTAL : (use16 ? PairWide16::TOTAL : PairLayout::TOTAL);
        p->threads = use_half ? PairHalf::THREADS : (use16 ? PairWide16::THREADS : PairLayout::THREADS);
    } else if (bn == 64) { p->kernel = gemm_pick_kernel_64(op.a_mn, op.b_mn, epi); p->smem = SmemLayout<64>::TOTAL; }
    else if (bn == 128) { p->kernel = gemm_pick_kernel_128(op.a_mn, op.b_mn, epi); p->smem = SmemLayout<128>::TOTAL; }
    else { p->kernel = gemm_pick_kernel_256(op.a_mn, op.b_mn, epi); p->smem = SmemLayout<256>::TOTAL; }
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

// b200mart_gemm_timeline: device buffer of `records` x 148 x 16 u64; launch i writes its CTAs' stamps into record i
static unsigned long long* g_timeline = nullptr;
static long g_timeline_records = 0, g_timeline_next = 0;

extern "C" int64_t b200mart_gemm_timeline(void* device_buffer, int64_t records) {
    const long used = g_timeline_next;
    g_timeline = static_cast<unsigned long long*>(device_buffer);
    g_timeline_records = device_buffer ? records : 0;
    g_timeline_next = 0;
    return used;
}

static unsigned long long* timeline_slot() {
    if (!g_timeline || g_timeline_next >= g_timeline_records) return nullptr;
    return g_timeline + (size_t)(g_timeline_next++) * 148 * 16;
}

int gemm_plan_launch(const GemmPlan* p, cudaStream_t stream, const DropCfg* drop) {
    if (g_timeline && !p->pair) {
        GemmDev dev = p->dev;
        if (drop) dev.drop = *drop;
        dev.prof = timeline_slot();
        CUDA_TRY(launch_pdl(p->kernel, p->grid, dim3(GEMM_THREADS), (size_t)p->smem, stream, p->tmA, p->tmB, p->tmA2, p->tmB2, dev));
        B2_LAUNCHED();
        return 0;
    }
    if (p->pair) {
        GemmDev dev = p->dev;
        if (drop) dev.drop = *drop;
        if (g_timeline) dev.prof = timeline_slot();
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = p->grid;
        cfg.blockDim = dim3(p->threads, 1, 1);
        cfg.dynamicSmemBytes = (size_t)p->smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = b200mart_pdl_enabled();
        cfg.attrs = attr;
        cfg.numAttrs = 2;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, p->kernel, p->tmA, p->tmB, p->tmA2, p->tmB2, dev));
        ++g_b200mart_launches;
        return 0;
    }
    if (drop) {   // per-step dropout seed: the plan (tensor maps, geometry) is reused, only the mask seed changes
        GemmDev dev = p->dev;
        dev.drop = *drop;
        CUDA_TRY(launch_pdl(p->kernel, p->grid, dim3(GEMM_THREADS), (size_t)p->smem, stream, p->tmA, p->tmB, p->tmA2, p->tmB2, dev));
    } else {
        CUDA_TRY(launch_pdl(p->kernel, p->grid, dim3(GEMM_THREADS), (size_t)p->smem, stream, p->tmA, p->tmB, p->tmA2, p->tmB2, p->dev));
    }
    B2_LAUNCHED();
    return 0;
}
