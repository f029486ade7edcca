// BertFusion (MarT/models/modeling_unimo.py:400-414): out = softmax(ctx vh^T) vh over the vision tokens, no scale,
// no mask.  fp32 throughout on CUDA cores: the scores are raw 768-dim dot products (O(sqrt(H)) magnitude, near
// one-hot softmax), where bf16 operand rounding would move probabilities by several percent (SURVEY.md §7 hard
// part 4); the op is 0.3 % of the step's FLOPs.  (Self-attention lives in attention_mma.cu.)
#include "kernels.cuh"

namespace {

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

template <typename K>
int set_smem(K kernel, size_t bytes) {
    B2_CHECK(bytes <= 227 * 1024, "attention: needs %zu bytes of shared memory (> 227 KB)", bytes);
    CUDA_TRY(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

}  // namespace

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
