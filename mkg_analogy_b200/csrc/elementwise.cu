// HBM-bound kernels of the MKGformer step: casts, im2col, embedding gathers + LayerNorm, LayerNorm
// forward/backward, column sums (bias gradients), embedding scatter-adds, AdamW.
// Layout rule: one warp per row for everything LayerNorm-shaped (H % 128 == 0, H <= 1024), lanes own
// float4 columns {lane*4 + 128*t}; all global accesses are 16-byte vectors.
#include "kernels.cuh"
#include <stdlib.h>

namespace {

constexpr int MAXV = 8;  // float4 per lane -> H <= 1024

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4_bf16(bf16* p, float4 v) {
    uint2 q = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    *reinterpret_cast<uint2*>(p) = q;
}

// ---------------------------------------------------------------- casts / fills
__global__ void cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long n) {
    long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    const long stride = (long)gridDim.x * blockDim.x * 8;
    for (; i + 8 <= n; i += stride) {
        float4 a = ld4(src + i), b = ld4(src + i + 4);
        uint4 q = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
        *reinterpret_cast<uint4*>(dst + i) = q;
    }
    if (i < n) for (long j = i; j < n && j < i + 8; ++j) dst[j] = __float2bfloat16(src[j]);
}
__global__ void fill_kernel(float* dst, float v, long n) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long stride = (long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] = v;
}

__global__ void axpy_kernel(const float* __restrict__ src, float* dst, long n) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long stride = (long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) dst[i] += src[i];
}
// row-wise fp32 -> bf16 cast into a padded destination (columns [cols, ld_dst) are zeroed)
__global__ void cast_pad_bf16_kernel(const float* __restrict__ src, long ld_src, bf16* __restrict__ dst, long ld_dst, int rows,
                                     int cols) {
    const long total = (long)rows * ld_dst;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / ld_dst;
        const int c = (int)(i % ld_dst);
        dst[i] = __float2bfloat16(c < cols ? src[r * ld_src + c] : 0.f);
    }
}

// ---------------------------------------------------------------- im2col (patch gather)
// one thread per VW consecutive pixels of a patch row: coalesced vector loads along the image row.
// out rows have leading dimension ldo (>= 3*p*p; padding columns are never read by the GEMM).
template <int VW>
__global__ void im2col_kernel(const float* __restrict__ px, bf16* __restrict__ out, int B, int S, int p, long ldo) {
    const int g = S / p, P = g * g;
    const int SV = S / VW;
    const long total = (long)B * 2 * 3 * S * SV;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const int xv = (int)(t % SV);
        long r = t / SV;
        const int y = (int)(r % S); r /= S;
        const int c = (int)(r % 3); r /= 3;
        const int img = (int)(r % 2);
        const int b = (int)(r / 2);
        const float* src = px + ((((long)b * 2 + img) * 3 + c) * S + y) * S + xv * VW;
        const int x = xv * VW;
        const int py = y / p, iy = y % p, pxx = x / p, ix = x % p;  // p % VW == 0 checked on the host
        const long row = ((long)b * 2 + img) * P + py * g + pxx;
        bf16* dst = out + row * ldo + (long)c * p * p + iy * p + ix;
        if (VW == 4) {
            st4_bf16(dst, ld4(src));
        } else if (VW == 2) {
            const float2 v = *reinterpret_cast<const float2*>(src);
            *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(v.x, v.y);
        } else {
            *dst = __float2bfloat16(*src);
        }
    }
}

// ---------------------------------------------------------------- LayerNorm core (warp per row)
template <bool WRITE_STATS>
__device__ __forceinline__ void ln_row(float4 (&v)[MAXV], int nv, int H, const float* gamma, const float* beta, float eps,
                                       float* y32, bf16* y16, float* mean_out, float* rstd_out, int lane,
                                       DropCfg drop = DropCfg{0u, 0u, 1.f}, uint32_t row_idx0 = 0u) {
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < MAXV; ++t)
        if (t < nv) s += v[t].x + v[t].y + v[t].z + v[t].w;
    const float mean = warp_sum(s) / H;
    float q = 0.f;
#pragma unroll
    for (int t = 0; t < MAXV; ++t)
        if (t < nv) {
            float a = v[t].x - mean, b = v[t].y - mean, c = v[t].z - mean, d = v[t].w - mean;
            q += a * a + b * b + c * c + d * d;
        }
    const float rstd = rsqrtf(warp_sum(q) / H + eps);
    if (WRITE_STATS && lane == 0) { *mean_out = mean; *rstd_out = rstd; }
#pragma unroll
    for (int t = 0; t < MAXV; ++t)
        if (t < nv) {
            const int col = lane * 4 + t * 128;
            const float4 g = ld4(gamma + col), b = ld4(beta + col);
            float4 o;
            o.x = (v[t].x - mean) * rstd * g.x + b.x;
            o.y = (v[t].y - mean) * rstd * g.y + b.y;
            o.z = (v[t].z - mean) * rstd * g.z + b.z;
            o.w = (v[t].w - mean) * rstd * g.w + b.w;
            if (drop.thresh) {   // dropout on the normalised output (BertEmbeddings :185); element index = row * H + col
                o.x *= drop_mult(drop, row_idx0 + col); o.y *= drop_mult(drop, row_idx0 + col + 1);
                o.z *= drop_mult(drop, row_idx0 + col + 2); o.w *= drop_mult(drop, row_idx0 + col + 3);
            }
            if (y32) st4(y32 + col, o);
            if (y16) st4_bf16(y16 + col, o);
        }
}

__global__ void layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float eps, float* y32, bf16* y16, float* mean,
                                     float* rstd, int M, int H) {
    pdl_wait();
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int nv = H / 128;
    float4 v[MAXV];
#pragma unroll
    for (int t = 0; t < MAXV; ++t)
        if (t < nv) v[t] = ld4(x + (long)row * H + lane * 4 + t * 128);
    ln_row<true>(v, nv, H, gamma, beta, eps, y32 ? y32 + (long)row * H : nullptr, y16 ? y16 + (long)row * H : nullptr,
                 mean + row, rstd + row, lane);
}

__global__ void vision_embed_ln_kernel(const float* __restrict__ patch_out, const float* __restrict__ cls,
                                       const float* __restrict__ pos, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float eps, float* x0, float* out, float* mean,
                                       float* rstd, int B, int P, int H) {
    const int Nv = 1 + 2 * P;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B * Nv) return;
    const int b = row / Nv, t = row % Nv;
    // token 0 = class embedding + pos[0]; token 1..P = image 0 patches + pos[1..P]; P+1..2P = image 1 + pos[1..P]
    const float* src = (t == 0) ? cls : patch_out + ((long)b * 2 * P + (t - 1)) * H;
    const int pidx = (t == 0) ? 0 : 1 + (t - 1) % P;
    const int nv = H / 128;
    float4 v[MAXV];
#pragma unroll
    for (int k = 0; k < MAXV; ++k)
        if (k < nv) {
            const int col = lane * 4 + k * 128;
            float4 a = ld4(src + col), p = ld4(pos + (long)pidx * H + col);
            v[k] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
            st4(x0 + (long)row * H + col, v[k]);
        }
    ln_row<true>(v, nv, H, gamma, beta, eps, out + (long)row * H, nullptr, mean + row, rstd + row, lane);
}

__global__ void text_embed_ln_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ tt,
                                     const float* __restrict__ E, const float* __restrict__ pos,
                                     const float* __restrict__ type, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float eps, float* pre, float* out32, bf16* out16,
                                     float* mean, float* rstd, int B, int L, int H, int V, int TV, int* status, DropCfg drop) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= B * L) return;
    const int l = row % L;
    long id = ids[row];
    long ty = tt ? tt[row] : 0;
    if (id < 0 || id >= V) { if (lane == 0) atomicOr(status, 1); id = 0; }  // reference would raise IndexError
    if (ty < 0 || ty >= TV) { if (lane == 0) atomicOr(status, 2); ty = 0; }
    const int nv = H / 128;
    float4 v[MAXV];
#pragma unroll
    for (int k = 0; k < MAXV; ++k)
        if (k < nv) {
            const int col = lane * 4 + k * 128;
            float4 a = ld4(E + id * H + col), p = ld4(pos + (long)l * H + col), t = ld4(type + ty * H + col);
            v[k] = make_float4(a.x + t.x + p.x, a.y + t.y + p.y, a.z + t.z + p.z, a.w + t.w + p.w);
            st4(pre + (long)row * H + col, v[k]);
        }
    ln_row<true>(v, nv, H, gamma, beta, eps, out32 + (long)row * H, out16 + (long)row * H, mean + row, rstd + row, lane, drop,
                 (uint32_t)row * (uint32_t)H);
}

// LayerNorm backward. Each warp walks rows r = warp_id, warp_id + nwarps, ...; dgamma/dbeta (and, optionally, the column
// sums of the produced dx = bias gradient of the linear layer that fed this LayerNorm's input) stay in registers (same
// lane->column mapping for every row) and are reduced through shared memory + one atomic per column per block.
template <bool COLSUM, int NV>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ add,
                     const bf16* __restrict__ dact_src, int dact, float* dx32, bf16* dx16, float* dgamma, float* dbeta,
                     float* dxsum, int M, int H, DropCfg dy_drop, DropCfg dx16_drop) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ float red[];  // [warps][H] x (2 or 3)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int nv = H / 128;
    float4 g[NV], ag[NV], ab[NV], ac[COLSUM ? NV : 1];
#pragma unroll
    for (int t = 0; t < NV; ++t) {
        ag[t] = make_float4(0, 0, 0, 0);
        ab[t] = make_float4(0, 0, 0, 0);
        if (COLSUM) ac[t] = make_float4(0, 0, 0, 0);
        if (t < nv) g[t] = ld4(gamma + lane * 4 + t * 128);
    }
    for (int row = blockIdx.x * nw + warp; row < M; row += gridDim.x * nw) {
        const float mu = mean[row], rs = rstd[row];
        float4 d[NV], xh[NV], av[NV];
        uint2 qv[NV];
        float s1 = 0.f, s2 = 0.f;
        // every HBM input of the row is requested up front (one round trip per row instead of two)
#pragma unroll
        for (int t = 0; t < NV; ++t)
            if (t < nv) {
                const long off = (long)row * H + lane * 4 + t * 128;
                if (add) av[t] = ld4(add + off);
                if (dact) qv[t] = *reinterpret_cast<const uint2*>(dact_src + off);
            }
#pragma unroll
        for (int t = 0; t < NV; ++t)
            if (t < nv) {
                const long off = (long)row * H + lane * 4 + t * 128;
                d[t] = ld4(dy + off);
                if (dy_drop.thresh) {   // this LayerNorm's output went through dropout: dy arrives wrt the dropped output
                    const uint32_t e = (uint32_t)off;
                    d[t].x *= drop_mult(dy_drop, e); d[t].y *= drop_mult(dy_drop, e + 1);
                    d[t].z *= drop_mult(dy_drop, e + 2); d[t].w *= drop_mult(dy_drop, e + 3);
                }
                float4 xv = ld4(x + off);
                xh[t] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
                ab[t].x += d[t].x; ab[t].y += d[t].y; ab[t].z += d[t].z; ab[t].w += d[t].w;
                ag[t].x += d[t].x * xh[t].x; ag[t].y += d[t].y * xh[t].y; ag[t].z += d[t].z * xh[t].z; ag[t].w += d[t].w * xh[t].w;
                d[t].x *= g[t].x; d[t].y *= g[t].y; d[t].z *= g[t].z; d[t].w *= g[t].w;
                s1 += d[t].x + d[t].y + d[t].z + d[t].w;
                s2 += d[t].x * xh[t].x + d[t].y * xh[t].y + d[t].z * xh[t].z + d[t].w * xh[t].w;
            }
        s1 = warp_sum(s1) / H;
        s2 = warp_sum(s2) / H;
#pragma unroll
        for (int t = 0; t < NV; ++t)
            if (t < nv) {
                const long off = (long)row * H + lane * 4 + t * 128;
                float4 o;
                o.x = rs * (d[t].x - s1 - xh[t].x * s2);
                o.y = rs * (d[t].y - s1 - xh[t].y * s2);
                o.z = rs * (d[t].z - s1 - xh[t].z * s2);
                o.w = rs * (d[t].w - s1 - xh[t].w * s2);
                if (dact) {
                    const uint2 q = qv[t];
                    const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.x));
                    const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.y));
                    if (dact == 1) {
                        o.x *= gelu_erf_grad(f0.x); o.y *= gelu_erf_grad(f0.y); o.z *= gelu_erf_grad(f1.x); o.w *= gelu_erf_grad(f1.y);
                    } else {
                        o.x *= quick_gelu_grad(f0.x); o.y *= quick_gelu_grad(f0.y); o.z *= quick_gelu_grad(f1.x); o.w *= quick_gelu_grad(f1.y);
                    }
                }
                if (add) {
                    const float4 a = av[t];
                    o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
                }
                if (dx32) st4(dx32 + off, o);      // full gradient: the residual path
                if (dx16_drop.thresh) {            // the dense branch fed this LayerNorm through dropout: mask its gradient
                    const uint32_t e = (uint32_t)off;
                    o.x *= drop_mult(dx16_drop, e); o.y *= drop_mult(dx16_drop, e + 1);
                    o.z *= drop_mult(dx16_drop, e + 2); o.w *= drop_mult(dx16_drop, e + 3);
                }
                if (COLSUM) { ac[t].x += o.x; ac[t].y += o.y; ac[t].z += o.z; ac[t].w += o.w; }
                if (dx16) st4_bf16(dx16 + off, o);
            }
    }
    float* rg = red;
    float* rb = red + nw * H;
    float* rc = red + 2 * nw * H;
#pragma unroll
    for (int t = 0; t < NV; ++t)
        if (t < nv) {
            st4(rg + warp * H + lane * 4 + t * 128, ag[t]);
            st4(rb + warp * H + lane * 4 + t * 128, ab[t]);
            if (COLSUM) st4(rc + warp * H + lane * 4 + t * 128, ac[t]);
        }
    __syncthreads();
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
        float sg = 0.f, sb = 0.f, sc = 0.f;
        for (int w = 0; w < nw; ++w) {
            sg += rg[w * H + c]; sb += rb[w * H + c];
            if (COLSUM) sc += rc[w * H + c];
        }
        atomicAdd(dgamma + c, sg);
        atomicAdd(dbeta + c, sb);
        if (COLSUM) atomicAdd(dxsum + c, sc);
    }
}

// Second design of the LayerNorm backward (default; B200MART_LNB=1 selects the one above).  The first keeps the per-column
// accumulators of dgamma / dbeta / column sums in registers (72 of its 226): one 8-warp block per SM, 12 % of the warp slots, an
// issue rate of 13-18 % and a third of the HBM rate (profiles/ln_bwd_r2.txt) - pure latency, no memory-level parallelism to hide it.
// Here the accumulators live in per-warp shared-memory rows (read-modify-write of the lane's own 16-byte slots: conflict-free),
// gamma is staged in shared memory, the optional `add` / saved-derivative inputs are loaded after the row reduction: <= 128
// registers, two blocks (16 warps) per SM; the per-block column totals leave as 4-wide vector reductions (a quarter of the
// L2 atomic operations per block).
__device__ __forceinline__ void red_add4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <bool COLSUM, int NV>
__global__ void __launch_bounds__(256, 2)
layernorm_bwd2_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                      const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ add,
                      const bf16* __restrict__ dact_src, int dact, float* dx32, bf16* dx16, float* dgamma, float* dbeta,
                      float* dxsum, int M, int H, DropCfg dy_drop, DropCfg dx16_drop) {
    extern __shared__ float red[];  // [2 or 3][warps][H] accumulators, then gamma [H]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int nv = H / 128;
    float* rg = red + warp * H;
    float* rb = red + (nw + warp) * H;
    float* rc = red + (2 * nw + warp) * H;
    float* gs = red + (COLSUM ? 3 : 2) * nw * H;
    for (int c = threadIdx.x * 4; c < H; c += blockDim.x * 4) st4(gs + c, ld4(gamma + c));     // parameters: not written by the predecessor
#pragma unroll
    for (int t = 0; t < NV; ++t)
        if (t < nv) {
            const int c = lane * 4 + t * 128;
            st4(rg + c, make_float4(0, 0, 0, 0));
            st4(rb + c, make_float4(0, 0, 0, 0));
            if (COLSUM) st4(rc + c, make_float4(0, 0, 0, 0));
        }
    __syncthreads();
    pdl_wait();
    pdl_launch_dependents();
    for (int row = blockIdx.x * nw + warp; row < M; row += gridDim.x * nw) {
        const float mu = mean[row], rs = rstd[row];
        float4 d[NV], xh[NV];
#pragma unroll
        for (int t = 0; t < NV; ++t)
            if (t < nv) {
                const long off = (long)row * H + lane * 4 + t * 128;
                d[t] = ld4(dy + off);
                xh[t] = ld4(x + off);
            }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int t = 0; t < NV; ++t)
            if (t < nv) {
                const int c = lane * 4 + t * 128;
                if (dy_drop.thresh) {   // this LayerNorm's output went through dropout: dy arrives wrt the dropped output
                    const uint32_t e = (uint32_t)((long)row * H + c);
                    d[t].x *= drop_mult(dy_drop, e); d[t].y *= drop_mult(dy_drop, e + 1);
                    d[t].z *= drop_mult(dy_drop, e + 2); d[t].w *= drop_mult(dy_drop, e + 3);
                }
                xh[t] = make_float4((xh[t].x - mu) * rs, (xh[t].y - mu) * rs, (xh[t].z - mu) * rs, (xh[t].w - mu) * rs);
                float4 ab = ld4(rb + c), ag = ld4(rg + c);
                ab.x += d[t].x; ab.y += d[t].y; ab.z += d[t].z; ab.w += d[t].w;
                ag.x += d[t].x * xh[t].x; ag.y += d[t].y * xh[t].y; ag.z += d[t].z * xh[t].z; ag.w += d[t].w * xh[t].w;
                st4(rb + c, ab);
                st4(rg + c, ag);
                const float4 g = ld4(gs + c);
                d[t].x *= g.x; d[t].y *= g.y; d[t].z *= g.z; d[t].w *= g.w;
                s1 += d[t].x + d[t].y + d[t].z + d[t].w;
                s2 += d[t].x * xh[t].x + d[t].y * xh[t].y + d[t].z * xh[t].z + d[t].w * xh[t].w;
            }
        s1 = warp_sum(s1) / H;
        s2 = warp_sum(s2) / H;
#pragma unroll
        for (int t = 0; t < NV; ++t)
            if (t < nv) {
                const int c = lane * 4 + t * 128;
                const long off = (long)row * H + c;
                float4 o;
                o.x = rs * (d[t].x - s1 - xh[t].x * s2);
                o.y = rs * (d[t].y - s1 - xh[t].y * s2);
                o.z = rs * (d[t].z - s1 - xh[t].z * s2);
                o.w = rs * (d[t].w - s1 - xh[t].w * s2);
                if (dact) {
                    const uint2 q = *reinterpret_cast<const uint2*>(dact_src + off);
                    const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.x));
                    const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.y));
                    if (dact == 1) {
                        o.x *= gelu_erf_grad(f0.x); o.y *= gelu_erf_grad(f0.y); o.z *= gelu_erf_grad(f1.x); o.w *= gelu_erf_grad(f1.y);
                    } else {
                        o.x *= quick_gelu_grad(f0.x); o.y *= quick_gelu_grad(f0.y); o.z *= quick_gelu_grad(f1.x); o.w *= quick_gelu_grad(f1.y);
                    }
                }
                if (add) {
                    const float4 a = ld4(add + off);
                    o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
                }
                if (dx32) st4(dx32 + off, o);      // full gradient: the residual path
                if (dx16_drop.thresh) {            // the dense branch fed this LayerNorm through dropout: mask its gradient
                    const uint32_t e = (uint32_t)off;
                    o.x *= drop_mult(dx16_drop, e); o.y *= drop_mult(dx16_drop, e + 1);
                    o.z *= drop_mult(dx16_drop, e + 2); o.w *= drop_mult(dx16_drop, e + 3);
                }
                if (COLSUM) {
                    float4 ac = ld4(rc + c);
                    ac.x += o.x; ac.y += o.y; ac.z += o.z; ac.w += o.w;
                    st4(rc + c, ac);
                }
                if (dx16) st4_bf16(dx16 + off, o);
            }
    }
    __syncthreads();
    for (int c = threadIdx.x * 4; c < H; c += blockDim.x * 4) {
        float4 sg = make_float4(0, 0, 0, 0), sb = sg, sc = sg;
        for (int w = 0; w < nw; ++w) {
            const float4 a = ld4(red + w * H + c), b = ld4(red + (nw + w) * H + c);
            sg.x += a.x; sg.y += a.y; sg.z += a.z; sg.w += a.w;
            sb.x += b.x; sb.y += b.y; sb.z += b.z; sb.w += b.w;
            if (COLSUM) { const float4 q = ld4(red + (2 * nw + w) * H + c); sc.x += q.x; sc.y += q.y; sc.z += q.z; sc.w += q.w; }
        }
        red_add4(dgamma + c, sg);
        red_add4(dbeta + c, sb);
        if (COLSUM) red_add4(dxsum + c, sc);
    }
}

// ---------------------------------------------------------------- column sums (bias gradients)
// block = 32 x 8 threads over a 32-column x ROWS strip; coalesced row-major reads, one atomic per column per block.
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, long ld, float* out, int M, int N, int rows_per_block) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float part[8][33];
    const int col = blockIdx.x * 32 + threadIdx.x;
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(M, r0 + rows_per_block);
    float s = 0.f;
    if (col < N)
        for (int r = r0 + threadIdx.y; r < r1; r += 8) s += (float)x[(long)r * ld + col];
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && col < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
        atomicAdd(out + col, t);
    }
}

// bf16 column sums, vectorised: a thread owns 8 consecutive columns (one 16-byte load per row), a warp reads 512 contiguous bytes
// of a row, the 8 row phases of a block are folded through shared memory, one 4-wide vector reduction per 4 columns per block.
// Needs N % 8 == 0, ld % 8 == 0 and 16-byte aligned x / out (the bias-gradient sums of the model: N = 768 / 2304 / 3072).
__global__ void __launch_bounds__(256)
colsum_bf16x8_kernel(const bf16* __restrict__ x, long ld, float* out, int M, int N, int rows_per_block) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float part[8][32][9];
    const int cg = blockIdx.x * 32 + threadIdx.x;          // column group of 8
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(M, r0 + rows_per_block);
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (cg * 8 < N)
        for (int r = r0 + threadIdx.y; r < r1; r += 8) {
            const uint4 v = *reinterpret_cast<const uint4*>(x + (long)r * ld + cg * 8);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
                s[2 * k] += f.x;
                s[2 * k + 1] += f.y;
            }
        }
#pragma unroll
    for (int k = 0; k < 8; ++k) part[threadIdx.y][threadIdx.x][k] = s[k];
    __syncthreads();
    // 256 threads fold 32 groups x 8 columns: thread (y, x) -> group x, columns 4 (y & 1) .. +3 for y < 2
    if (threadIdx.y < 2 && cg * 8 < N) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        const int k0 = 4 * threadIdx.y;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            t.x += part[i][threadIdx.x][k0]; t.y += part[i][threadIdx.x][k0 + 1];
            t.z += part[i][threadIdx.x][k0 + 2]; t.w += part[i][threadIdx.x][k0 + 3];
        }
        red_add4(out + cg * 8 + k0, t);
    }
}

// dst16 = bf16(src * dropout_mult), colsum[c] += sum_rows of the same values: turns an incoming fp32 gradient into the masked
// bf16 operand of the dW / dX products of a dense layer whose output went through dropout, plus that layer's bias gradient.
__global__ void mask_cast_colsum_kernel(const float* __restrict__ src, bf16* __restrict__ dst, float* colsum, int M, int N,
                                        int rows_per_block, DropCfg drop) {
    __shared__ float part[8][33];
    const int col = blockIdx.x * 32 + threadIdx.x;
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(M, r0 + rows_per_block);
    float s = 0.f;
    if (col < N)
        for (int r = r0 + threadIdx.y; r < r1; r += 8) {
            float v = src[(long)r * N + col];
            if (drop.thresh) v *= drop_mult(drop, (uint32_t)r * (uint32_t)N + (uint32_t)col);
            dst[(long)r * N + col] = __float2bfloat16(v);
            s += v;
        }
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && col < N && colsum) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
        atomicAdd(colsum + col, t);
    }
}

// ---------------------------------------------------------------- embedding backward
__global__ void text_embed_bwd_kernel(const float* __restrict__ d, const int64_t* __restrict__ ids,
                                      const int64_t* __restrict__ tt, float* dE, float* dpos, float* dtype, int B, int L,
                                      int H, int V, int TV) {
    const int row = blockIdx.x;
    const int l = row % L;
    long id = ids[row];
    long ty = tt ? tt[row] : 0;
    // same clamp as the forward kernel (which raised the status bit): an out-of-range id must never index past the tables
    if (id < 0 || id >= V) id = 0;
    if (ty < 0 || ty >= TV) ty = 0;
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
        const float v = d[(long)row * H + c];
        atomicAdd(dE + id * H + c, v);
        atomicAdd(dpos + (long)l * H + c, v);
        atomicAdd(dtype + ty * H + c, v);
    }
}

__global__ void vision_embed_bwd_kernel(const float* __restrict__ dx0, float* dcls, float* dpos, bf16* dpatch, int B,
                                        int P, int H) {
    const int Nv = 1 + 2 * P;
    const int row = blockIdx.x;
    const int b = row / Nv, t = row % Nv;
    const int pidx = (t == 0) ? 0 : 1 + (t - 1) % P;
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
        const float v = dx0[(long)row * H + c];
        atomicAdd(dpos + (long)pidx * H + c, v);
        if (t == 0) atomicAdd(dcls + c, v);
        else dpatch[((long)b * 2 * P + (t - 1)) * H + c] = __float2bfloat16(v);
    }
}

// ---------------------------------------------------------------- gathers / scatters
// Row / element gathers and scatters by a caller-supplied index list (the analogy-entity columns of the score head): an index
// outside [0, V) raises status bit 0 and reads row 0 (gathers) / is skipped (scatters) - never an out-of-bounds access.
__global__ void gather_rows_bf16_kernel(const bf16* __restrict__ src, const int64_t* __restrict__ idx, bf16* out, int R, int H, long V,
                                        int* status) {
    const int r = blockIdx.x;
    long i = idx[r];
    if (i < 0 || i >= V) { if (threadIdx.x == 0 && status) atomicOr(status, 1); i = 0; }
    const uint4* s = reinterpret_cast<const uint4*>(src + i * H);
    uint4* o = reinterpret_cast<uint4*>(out + (long)r * H);
    for (int c = threadIdx.x; c < H / 8; c += blockDim.x) o[c] = s[c];
}
__global__ void gather_f32_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, float* out, int R, long V) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R) { const long i = idx[r]; out[r] = (i < 0 || i >= V) ? 0.f : src[i]; }
}
__global__ void scatter_add_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, float* dst, int R, int H, long V) {
    const int r = blockIdx.x;
    if (idx[r] < 0 || idx[r] >= V) return;
    const long o = idx[r] * H;
    for (int c = threadIdx.x; c < H; c += blockDim.x) atomicAdd(dst + o + c, src[(long)r * H + c]);
}
__global__ void scatter_add_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, float* dst, int R, long V) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < R && idx[r] >= 0 && idx[r] < V) atomicAdd(dst + idx[r], src[r]);
}

// ---------------------------------------------------------------- AdamW (torch.optim.AdamW semantics)
__global__ void adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, bf16* __restrict__ p16, const AdamSeg* __restrict__ segs, int nseg,
                             long begin, long n, int zero_grads, float lr, float b1, float b2, float eps, float wd, float bc1,
                             float bc2, const float* __restrict__ dyn) {
    // dyn: device block {salt, lr, 1 - beta1^t, 1 - beta2^t} of the current step (b200mart_set_step_block): the per-step scalars of a
    // captured CUDA graph, whose kernel arguments are frozen
    if (dyn) { lr = __ldg(dyn + 1); bc1 = __ldg(dyn + 2); bc2 = __ldg(dyn + 3); }
    // one block per 4096-element chunk; the owning segment is found by binary search on the chunk start
    // (segments are aligned to 8 elements and chunks never straddle an unaligned boundary element-wise: each
    // thread re-checks its own element range).
  // grid-stride over 4096-element chunks: with a bounded grid (B200MART_ADAMW_CTAS) at most one of these 30-register blocks
  // sits on an SM, beside a GEMM CTA (53,760 of 65,536 registers) instead of in front of it
  for (long chunk = blockIdx.x; begin + chunk * 4096 < n; chunk += gridDim.x) {
    const long base = begin + chunk * 4096;     // processes the flat range [begin, n)
    for (int it = 0; it < 4; ++it) {
        const long i = base + ((long)it * blockDim.x + threadIdx.x) * 4;
        if (i >= n) break;
        int lo = 0, hi = nseg - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (segs[mid].end <= i) lo = mid + 1; else hi = mid;
        }
        const AdamSeg sg = segs[lo];
        float4 pv = ld4(p + i);
        if (!sg.skip && i >= sg.begin) {
            const float4 gv = ld4(g + i);
            float4 mv = ld4(m + i), vv = ld4(v + i);
            const float decay = sg.decay ? (1.0f - lr * wd) : 1.0f;
            const float step = lr / bc1;
            const float inv_sqrt_bc2 = rsqrtf(bc2);
#define ADAM1(P, G, Mm, Vv)                                                   \
    P *= decay; Mm = b1 * Mm + (1.0f - b1) * G; Vv = b2 * Vv + (1.0f - b2) * G * G; \
    P -= step * Mm / (sqrtf(Vv) * inv_sqrt_bc2 + eps);
            ADAM1(pv.x, gv.x, mv.x, vv.x)
            ADAM1(pv.y, gv.y, mv.y, vv.y)
            ADAM1(pv.z, gv.z, mv.z, vv.z)
            ADAM1(pv.w, gv.w, mv.w, vv.w)
#undef ADAM1
            st4(p + i, pv); st4(m + i, mv); st4(v + i, vv);
        }
        if (zero_grads) st4(g + i, make_float4(0.f, 0.f, 0.f, 0.f));   // the next step accumulates into a clean buffer
        if (p16) st4_bf16(p16 + i, pv);
    }
  }
}

// AdamW over this rank's shard of a flat range with the data-parallel exchange fused in, over NVSwitch multicast:
//   gradient  = mean over the ranks, pulled with multimem.ld_reduce.add from the multicast address of the (symmetric)
//               gradient buffers - the switch reads the element from every replica and returns the sum: the reduce-scatter;
//   update    = torch.optim.AdamW on the rank's own fp32 master / exp_avg / exp_avg_sq (1/N of the optimizer's HBM traffic);
//   broadcast = the new fp32 parameter and its bf16 mirror leave through multimem.st on the multicast addresses of the
//               (symmetric) parameter buffers - one send per rank, replicated by the switch: the all-gather.
// Every element is reduced once, by one rank, and the same bits are stored in every replica: replicas stay identical.
// The caller brackets the launch with two cross-rank barriers (gradients complete before / all reads done after).
__device__ __forceinline__ float4 mc_ld_reduce4(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st4(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mc_st2(void* mc, float a, float b) {     // 8 bytes of payload (four packed bf16)
    asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(mc), "f"(a), "f"(b) : "memory");
}
constexpr int MCQ = 8;     // 16-byte multimem.ld_reduce requests in flight per thread
constexpr int MCT = 1024;  // threads per CTA: 128 KB of gradient requests in flight per SM the kernel occupies
__global__ void __launch_bounds__(MCT, 1) adamw_mc_kernel(const float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, float* g_mc,
                                float* p_mc, bf16* p16_mc, const AdamSeg* __restrict__ segs, int nseg, long begin, long n,
                                float inv_world, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2) {
  // bounded persistent grid (the exchange is NVLink-latency bound, not HBM bound: a few CTAs with deep request queues keep the
  // link busy and leave the SMs to the backward pass, the way a collective library bounds its channels)
  for (long chunk = blockIdx.x; begin + chunk * (4L * MCT * MCQ) < n; chunk += gridDim.x) {
    const long base = begin + chunk * (4L * MCT * MCQ);
    // all MCQ gradient quads of the thread are requested before any is used: the NVLink round trip (~4 us) is paid once
    float4 gq[MCQ];
    bool live[MCQ];
#pragma unroll
    for (int it = 0; it < MCQ; ++it) {
        const long i = base + ((long)it * blockDim.x + threadIdx.x) * 4;
        live[it] = i < n;
        if (live[it]) gq[it] = mc_ld_reduce4(g_mc + i);
    }
#pragma unroll
    for (int it = 0; it < MCQ; ++it) {
        const long i = base + ((long)it * blockDim.x + threadIdx.x) * 4;
        if (!live[it]) continue;
        float4 gv = gq[it];
        // (clearing the consumed gradients through multimem.st as well was measured and rejected: another 4 bytes per parameter
        //  on NVLink; every rank clears its own replica with a local memset after the closing barrier of the extent instead)
        int lo = 0, hi = nseg - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (segs[mid].end <= i) lo = mid + 1; else hi = mid;
        }
        const AdamSeg sg = segs[lo];
        if (sg.skip || i < sg.begin) continue;            // never-updated parameters: identical on every replica already
        gv.x *= inv_world; gv.y *= inv_world; gv.z *= inv_world; gv.w *= inv_world;
        float4 pv = ld4(p + i), mv = ld4(m + i), vv = ld4(v + i);
        const float decay = sg.decay ? (1.0f - lr * wd) : 1.0f;
        const float step = lr / bc1;
        const float inv_sqrt_bc2 = rsqrtf(bc2);
#define ADAM1(P, G, Mm, Vv)                                                   \
    P *= decay; Mm = b1 * Mm + (1.0f - b1) * G; Vv = b2 * Vv + (1.0f - b2) * G * G; \
    P -= step * Mm / (sqrtf(Vv) * inv_sqrt_bc2 + eps);
        ADAM1(pv.x, gv.x, mv.x, vv.x)
        ADAM1(pv.y, gv.y, mv.y, vv.y)
        ADAM1(pv.z, gv.z, mv.z, vv.z)
        ADAM1(pv.w, gv.w, mv.w, vv.w)
#undef ADAM1
        st4(m + i, mv); st4(v + i, vv);
        mc_st4(p_mc + i, pv);
        mc_st2(p16_mc + i, __uint_as_float(pack_bf16x2(pv.x, pv.y)), __uint_as_float(pack_bf16x2(pv.z, pv.w)));
    }
  }
}

// ---------------------------------------------------------------- entity-image table gather (collator replacement)
// out[i] = table[idx[i]] (one preprocessed image per entity, fp32 or bf16 storage), zeros for idx < 0 ("[PAD] image",
// MarT/data/data_module.py:141-142).  One CTA row per image, 16-byte accesses; an index >= n_entities sets status bit 16
// and yields zeros instead of faulting.
template <typename T>
__global__ void gather_images_kernel(const T* __restrict__ table, const int64_t* __restrict__ idx, float* __restrict__ out,
                                     long image_elems, long n_entities, int* status) {
    const int img = blockIdx.y;
    const long e = idx[img];
    const bool ok = e >= 0 && e < n_entities;
    if (e >= n_entities && status && threadIdx.x == 0 && blockIdx.x == 0) atomicOr(status, 16);
    const T* src = table + (ok ? e : 0) * image_elems;
    float* dst = out + (long)img * image_elems;
    for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < image_elems; i += (long)gridDim.x * blockDim.x * 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) {
            if (sizeof(T) == 4) {
                v = ld4(reinterpret_cast<const float*>(src) + i);
            } else {
                const uint2 q = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(src) + i);
                const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.x));
                const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.y));
                v = make_float4(a.x, a.y, b.x, b.y);
            }
        }
        st4(dst + i, v);
    }
}

inline int blocks_for(long n, int per_block, int cap = 148 * 16) {
    long b = (n + per_block - 1) / per_block;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

int launch_cast_bf16(const float* src, bf16* dst, long n, cudaStream_t s) {
    cast_bf16_kernel<<<blocks_for(n, 256 * 8), 256, 0, s>>>(src, dst, n);
    B2_LAUNCHED();
    return 0;
}
// every replica of the multicast buffer: dst[i] += scale * src[i]  (reduced in the NVSwitch)
__global__ void multimem_add_kernel(const float* __restrict__ src, float* mc_dst, long n4, float scale) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long stride = (long)gridDim.x * blockDim.x;
    for (; i < n4; i += stride) {
        const float4 v = ld4(src + 4 * i);
        asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + 4 * i), "f"(v.x * scale),
                     "f"(v.y * scale), "f"(v.z * scale), "f"(v.w * scale)
                     : "memory");
    }
}
int launch_multimem_add(const float* src, float* mc_dst, long n, float scale, cudaStream_t s) {
    B2_CHECK(n % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(mc_dst) & 15) == 0,
             "multimem_add: n %ld must be a multiple of 4 and the buffers 16-byte aligned", n);
    if (n == 0) return 0;
    multimem_add_kernel<<<blocks_for(n / 4, 1024), 256, 0, s>>>(src, mc_dst, n / 4, scale);
    B2_LAUNCHED();
    return 0;
}
int launch_fill_f32(float* dst, float v, long n, cudaStream_t s) {
    fill_kernel<<<blocks_for(n, 1024), 256, 0, s>>>(dst, v, n);
    B2_LAUNCHED();
    return 0;
}
int launch_axpy_f32(const float* src, float* dst, long n, cudaStream_t s) {
    axpy_kernel<<<blocks_for(n, 1024), 256, 0, s>>>(src, dst, n);
    B2_LAUNCHED();
    return 0;
}
int launch_cast_pad_bf16(const float* src, long ld_src, bf16* dst, long ld_dst, int rows, int cols, cudaStream_t s) {
    cast_pad_bf16_kernel<<<blocks_for((long)rows * ld_dst, 1024), 256, 0, s>>>(src, ld_src, dst, ld_dst, rows, cols);
    B2_LAUNCHED();
    return 0;
}
int launch_im2col(const float* px, bf16* out, int B, int S, int p, long ldo, cudaStream_t s) {
    B2_CHECK(S % p == 0 && ldo >= 3L * p * p, "im2col: patch %d must divide image size %d", p, S);
    const int vw = (p % 4 == 0 && ldo % 4 == 0) ? 4 : ((p % 2 == 0 && ldo % 2 == 0) ? 2 : 1);
    const long total = (long)B * 2 * 3 * S * (S / vw);
    if (vw == 4) im2col_kernel<4><<<blocks_for(total, 256), 256, 0, s>>>(px, out, B, S, p, ldo);
    else if (vw == 2) im2col_kernel<2><<<blocks_for(total, 256), 256, 0, s>>>(px, out, B, S, p, ldo);
    else im2col_kernel<1><<<blocks_for(total, 256), 256, 0, s>>>(px, out, B, S, p, ldo);
    B2_LAUNCHED();
    return 0;
}
static int check_h(int H) {
    B2_CHECK(H % 128 == 0 && H <= 128 * MAXV, "hidden size %d must be a multiple of 128 and <= %d", H, 128 * MAXV);
    return 0;
}
int launch_vision_embed_ln(const float* patch_out, const float* cls, const float* pos, const float* gamma,
                           const float* beta, float eps, float* x0, float* out_f32, float* mean, float* rstd, int B,
                           int P, int H, cudaStream_t s) {
    B2_TRY(check_h(H));
    const int rows = B * (1 + 2 * P);
    vision_embed_ln_kernel<<<(rows + 3) / 4, 128, 0, s>>>(patch_out, cls, pos, gamma, beta, eps, x0, out_f32, mean, rstd, B, P, H);
    B2_LAUNCHED();
    return 0;
}
int launch_text_embed_ln(const int64_t* ids, const int64_t* tt, const float* E, const float* pos, const float* type,
                         const float* gamma, const float* beta, float eps, float* pre, float* out_f32, bf16* out_bf16,
                         float* mean, float* rstd, int B, int L, int H, int V, int TV, int* status, DropCfg drop, cudaStream_t s) {
    B2_TRY(check_h(H));
    text_embed_ln_kernel<<<(B * L + 3) / 4, 128, 0, s>>>(ids, tt, E, pos, type, gamma, beta, eps, pre, out_f32, out_bf16,
                                                        mean, rstd, B, L, H, V, TV, status, drop);
    B2_LAUNCHED();
    return 0;
}
int launch_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, float* y32, bf16* y16,
                         float* mean, float* rstd, int M, int H, cudaStream_t s) {
    B2_TRY(check_h(H));
    CUDA_TRY(launch_pdl(layernorm_fwd_kernel, dim3((M + 3) / 4), dim3(128), 0, s, x, gamma, beta, eps, y32, y16, mean, rstd, M, H));
    B2_LAUNCHED();
    return 0;
}
int launch_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                         const float* add, const bf16* dact_src, int dact, float* dx32, bf16* dx16, float* dgamma,
                         float* dbeta, float* dxsum, int M, int H, cudaStream_t s, DropCfg dy_drop, DropCfg dx16_drop) {
    B2_TRY(check_h(H));
    const int nw = 8;
    int blocks = (M + nw - 1) / nw;
    // rows per warp vs per-column atomics (2-3 x H per block): B200MART_LNB_BLOCKS = blocks per SM; 1: the register arrays allow one resident block
    static int per_sm = -1;
    if (per_sm < 0) { const char* e = getenv("B200MART_LNB_BLOCKS"); per_sm = e ? atoi(e) : 1; if (per_sm < 1) per_sm = 1; }
    if (blocks > 148 * per_sm) blocks = 148 * per_sm;
    const size_t smem = (size_t)(dxsum ? 3 : 2) * nw * H * sizeof(float);
    auto go = [&](auto kernel) -> int {
        CUDA_TRY(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(launch_pdl(kernel, dim3(blocks), dim3(nw * 32), smem, s, dy, x, mean, rstd, gamma, add, dact_src, dact_src ? dact : 0, dx32,
                            dx16, dgamma, dbeta, dxsum, M, H, dy_drop, dx16_drop));
        return 0;
    };
    const bool small = H <= 768;       // register arrays sized for the hidden size: 6 (base) or 8 (large) float4 per lane
    static int design = -1;
    if (design < 0) { const char* e = getenv("B200MART_LNB"); design = e ? atoi(e) : 2; }
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (design == 2 && al16(dgamma) && al16(dbeta) && (!dxsum || al16(dxsum)) && al16(gamma)) {
        blocks = (M + nw - 1) / nw;
        if (blocks > 148 * 2) blocks = 148 * 2;
        const size_t smem2 = ((size_t)(dxsum ? 3 : 2) * nw + 1) * H * sizeof(float);
        auto go2 = [&](auto kernel) -> int {
            CUDA_TRY(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            CUDA_TRY(launch_pdl(kernel, dim3(blocks), dim3(nw * 32), smem2, s, dy, x, mean, rstd, gamma, add, dact_src, dact_src ? dact : 0,
                                dx32, dx16, dgamma, dbeta, dxsum, M, H, dy_drop, dx16_drop));
            return 0;
        };
        if (dxsum) { B2_TRY(small ? go2(layernorm_bwd2_kernel<true, 6>) : go2(layernorm_bwd2_kernel<true, 8>)); }
        else { B2_TRY(small ? go2(layernorm_bwd2_kernel<false, 6>) : go2(layernorm_bwd2_kernel<false, 8>)); }
        B2_LAUNCHED();
        return 0;
    }
    if (dxsum) { B2_TRY(small ? go(layernorm_bwd_kernel<true, 6>) : go(layernorm_bwd_kernel<true, 8>)); }
    else { B2_TRY(small ? go(layernorm_bwd_kernel<false, 6>) : go(layernorm_bwd_kernel<false, 8>)); }
    B2_LAUNCHED();
    return 0;
}
int launch_colsum_bf16(const bf16* x, long ld, float* out, int M, int N, cudaStream_t s) {
    if (N % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        const int rpb8 = 64;
        CUDA_TRY(launch_pdl(colsum_bf16x8_kernel, dim3((N / 8 + 31) / 32, (M + rpb8 - 1) / rpb8), dim3(32, 8), 0, s, x, ld, out, M, N, rpb8));
        B2_LAUNCHED();
        return 0;
    }
    const int rpb = 128;
    CUDA_TRY(launch_pdl(colsum_kernel<bf16>, dim3((N + 31) / 32, (M + rpb - 1) / rpb), dim3(32, 8), 0, s, x, ld, out, M, N, rpb));
    B2_LAUNCHED();
    return 0;
}
int launch_colsum_f32(const float* x, long ld, float* out, int M, int N, cudaStream_t s) {
    const int rpb = 128;
    colsum_kernel<float><<<dim3((N + 31) / 32, (M + rpb - 1) / rpb), dim3(32, 8), 0, s>>>(x, ld, out, M, N, rpb);
    B2_LAUNCHED();
    return 0;
}
int launch_mask_cast_colsum(const float* src, bf16* dst, float* colsum, int M, int N, DropCfg drop, cudaStream_t s) {
    const int rpb = 64;
    mask_cast_colsum_kernel<<<dim3((N + 31) / 32, (M + rpb - 1) / rpb), dim3(32, 8), 0, s>>>(src, dst, colsum, M, N, rpb, drop);
    B2_LAUNCHED();
    return 0;
}
int launch_text_embed_bwd(const float* d, const int64_t* ids, const int64_t* tt, float* dE, float* dpos, float* dtype,
                          int B, int L, int H, int V, int TV, cudaStream_t s) {
    text_embed_bwd_kernel<<<B * L, 256, 0, s>>>(d, ids, tt, dE, dpos, dtype, B, L, H, V, TV);
    B2_LAUNCHED();
    return 0;
}
int launch_vision_embed_bwd(const float* dx0, float* dcls, float* dpos, bf16* dpatch, int B, int P, int H,
                            cudaStream_t s) {
    vision_embed_bwd_kernel<<<B * (1 + 2 * P), 256, 0, s>>>(dx0, dcls, dpos, dpatch, B, P, H);
    B2_LAUNCHED();
    return 0;
}
int launch_gather_rows_bf16(const bf16* src, const int64_t* idx, bf16* out, int R, int H, long V, int* status, cudaStream_t s) {
    B2_CHECK(H % 8 == 0, "gather_rows: H %d must be a multiple of 8", H);
    gather_rows_bf16_kernel<<<R, 96, 0, s>>>(src, idx, out, R, H, V, status);
    B2_LAUNCHED();
    return 0;
}
int launch_gather_f32(const float* src, const int64_t* idx, float* out, int R, long V, cudaStream_t s) {
    gather_f32_kernel<<<(R + 255) / 256, 256, 0, s>>>(src, idx, out, R, V);
    B2_LAUNCHED();
    return 0;
}
int launch_scatter_add_rows_f32(const float* src, const int64_t* idx, float* dst, int R, int H, long V, cudaStream_t s) {
    scatter_add_rows_kernel<<<R, 256, 0, s>>>(src, idx, dst, R, H, V);
    B2_LAUNCHED();
    return 0;
}
int launch_scatter_add_f32(const float* src, const int64_t* idx, float* dst, int R, long V, cudaStream_t s) {
    scatter_add_kernel<<<(R + 255) / 256, 256, 0, s>>>(src, idx, dst, R, V);
    B2_LAUNCHED();
    return 0;
}
int launch_adamw_mc(const float* p, float* m, float* v, float* g_mc, float* p_mc, bf16* p16_mc, const AdamSeg* segs_dev,
                    int nseg, long begin, long end, float inv_world, float lr, float b1, float b2, float eps, float wd, float bc1,
                    float bc2, cudaStream_t s, int max_ctas) {
    B2_CHECK(begin % 8 == 0 && end % 8 == 0 && begin < end, "adamw_mc: range [%ld, %ld) must be non-empty and 8-aligned", begin, end);
    long blocks = (end - begin + 4L * MCT * MCQ - 1) / (4L * MCT * MCQ);
    static int env_cap = -1;   // B200MART_MC_CTAS: bound of the grid (0 = one block per 32768 elements)
    if (env_cap < 0) { const char* e = getenv("B200MART_MC_CTAS"); env_cap = e ? atoi(e) : 32; if (env_cap < 0) env_cap = 0; }
    const int cap = max_ctas > 0 ? max_ctas : env_cap;     // the caller widens the grid for extents it exchanges after backward
    if (cap > 0 && blocks > cap) blocks = cap;
    adamw_mc_kernel<<<(unsigned)blocks, MCT, 0, s>>>(p, m, v, g_mc, p_mc, p16_mc, segs_dev, nseg, begin, end, inv_world, lr, b1, b2,
                                                    eps, wd, bc1, bc2);
    B2_LAUNCHED();
    return 0;
}

int launch_adamw(float* p, float* g, float* m, float* v, bf16* p16, const AdamSeg* segs_dev, int nseg, long begin, long end,
                 int zero_grads, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2, cudaStream_t s, const float* dyn) {
    B2_CHECK(begin % 8 == 0 && end % 8 == 0 && begin < end, "adamw: range [%ld, %ld) must be non-empty and 8-aligned", begin, end);
    long blocks = (end - begin + 4095) / 4096;
    static int cap = -1;       // B200MART_ADAMW_CTAS: bound of the grid (0 = one block per 4096 elements)
    if (cap < 0) { const char* e = getenv("B200MART_ADAMW_CTAS"); cap = e ? atoi(e) : 0; if (cap < 0) cap = 0; }
    if (cap > 0 && blocks > cap) blocks = cap;
    adamw_kernel<<<(unsigned)blocks, 256, 0, s>>>(p, g, m, v, p16, segs_dev, nseg, begin, end, zero_grads, lr, b1, b2, eps, wd,
                                                  bc1, bc2, dyn);
    B2_LAUNCHED();
    return 0;
}
int launch_gather_images(const void* table, int table_is_bf16, const int64_t* idx, float* out, int n_images, long image_elems,
                         long n_entities, int* status, cudaStream_t s) {
    B2_CHECK(table && idx && out && n_images > 0 && image_elems > 0 && n_entities > 0, "gather_images: bad arguments");
    B2_CHECK(image_elems % 4 == 0, "gather_images: image size %ld must be a multiple of 4 elements", image_elems);
    B2_CHECK(((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "gather_images: table / out must be 16-byte aligned");
    long per = (image_elems / 4 + 255) / 256;
    if (per > 64) per = 64;
    const dim3 grid((unsigned)per, (unsigned)n_images);
    if (table_is_bf16) gather_images_kernel<bf16><<<grid, 256, 0, s>>>(static_cast<const bf16*>(table), idx, out, image_elems, n_entities, status);
    else gather_images_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(table), idx, out, image_elems, n_entities, status);
    B2_LAUNCHED();
    return 0;
}
