// Score-head side kernels: [MASK] row location, label-smoothed cross-entropy with its softmax-backward fused
// (LabelSmoothSoftmaxCEV1, MarT/lit_models/utils.py:42-66; plain CE when eps = 0, transformer.py:25), the
// relaxation (cosine) loss of MarT/lit_models/transformer.py:103-108 with its gradient, and the rank of the gold
// entity (transformer.py:162-164, as a strict-greater count instead of two sorts).
#include "kernels.cuh"

namespace {

__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += red[w];
    return t;
}
__device__ __forceinline__ float block_reduce_max(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = -INFINITY;
    for (int w = 0; w < nw; ++w) t = fmaxf(t, red[w]);
    return t;
}

// one warp per sample; status bit 4 = a row without exactly one [MASK] (the reference asserts, transformer.py:74)
__global__ void find_mask_kernel(const int64_t* __restrict__ ids, int* mask_pos, int B, int L, int mask_id, int* status) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    int cnt = 0, pos = 0;
    for (int l = lane; l < L; l += 32)
        if (ids[(long)b * L + l] == mask_id) { ++cnt; pos = l; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        pos = max(pos, __shfl_xor_sync(0xffffffffu, pos, o));
    }
    if (lane == 0) {
        mask_pos[b] = pos;
        if (cnt != 1) atomicOr(status, 4);
    }
}

__global__ void gather_mask_rows_kernel(const bf16* __restrict__ z16, const int* __restrict__ mask_pos, bf16* zm, int L, int H) {
    const int b = blockIdx.x;
    const uint4* s = reinterpret_cast<const uint4*>(z16 + ((long)b * L + mask_pos[b]) * H);
    uint4* o = reinterpret_cast<uint4*>(zm + (long)b * H);
    for (int c = threadIdx.x; c < H / 8; c += blockDim.x) o[c] = s[c];
}

__device__ __forceinline__ bool row_selected(int select, const int64_t* pre_type, int b) {
    if (select == 0) return true;
    const bool rel = pre_type[b] == 2;
    return select == 2 ? rel : !rel;
}

// one CTA per sample.  target q: q[label] = 1-eps, q[c != label] = eps/C  (sum q = 1 - eps + (C-1) eps / C);
// loss_b = (sum q) * logsumexp(x) - sum_c q_c x_c ;  dx = ((sum q) softmax(x) - q) / n_selected.
__global__ void __launch_bounds__(256)
ls_ce_kernel(const LossGroup g, const int64_t* __restrict__ label, const int64_t* __restrict__ pre_type, float eps, float* loss,
             int B, int* status, const float* __restrict__ grad_scale) {
    __shared__ float red[8];
    __shared__ int nsel_s;
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        int c = 0;
        for (int i = 0; i < B; ++i) c += row_selected(g.select, pre_type, i) ? 1 : 0;
        nsel_s = c;
    }
    __syncthreads();
    const bool sel = row_selected(g.select, pre_type, b);
    const int C = g.C;
    if (!sel) {
        for (int c = tid; c < g.ldd; c += blockDim.x) {
            if (g.dlogits16) g.dlogits16[(long)b * g.ldd + c] = __float2bfloat16(0.f);
            if (g.dlogits32) g.dlogits32[(long)b * g.ldd + c] = 0.f;
        }
        return;
    }
    const float* x = g.logits + (long)b * g.ld;
    long yl = label[b];
    if (yl < 0 || yl >= C) { if (tid == 0 && status) atomicOr(status, 8); yl = 0; }   // the reference would raise IndexError
    const int y = (int)yl;
    float mx = -INFINITY, sx = 0.f;
    for (int c = tid; c < C; c += blockDim.x) {
        const float v = x[c];
        mx = fmaxf(mx, v);
        sx += v;
    }
    mx = block_reduce_max(mx, red);
    sx = block_reduce_sum(sx, red);
    float se = 0.f;
    for (int c = tid; c < C; c += blockDim.x) se += __expf(x[c] - mx);
    se = block_reduce_sum(se, red);
    const float lse = mx + __logf(se);
    const float q_other = eps / C, q_lab = 1.0f - eps;
    const float qsum = q_lab + (C - 1) * q_other;
    const float inv_n = 1.0f / nsel_s;
    if (tid == 0) {
        const float xl = x[y];
        const float l = qsum * lse - (q_lab * xl + q_other * (sx - xl));
        atomicAdd(loss, l * inv_n);
    }
    const float gs = inv_n * (grad_scale ? *grad_scale : 1.0f);      // incoming d(loss), applied at the seed of the backward pass
    for (int c = tid; c < g.ldd; c += blockDim.x) {
        float d = 0.f;
        if (c < C) d = (qsum * __expf(x[c] - lse) - (c == y ? q_lab : q_other)) * gs;
        if (g.dlogits16) g.dlogits16[(long)b * g.ldd + c] = __float2bfloat16(d);
        if (g.dlogits32) g.dlogits32[(long)b * g.ldd + c] = d;
    }
}

// one CTA (128 threads) per sample: relu(cos(z[qh], z[ah])) + 1 - cos(z[r0], z[r1]); torch cosine_similarity eps 1e-8
__global__ void __launch_bounds__(128)
relaxation_kernel(const float* __restrict__ z, const int64_t* __restrict__ rel_idx, const int64_t* __restrict__ q_head,
                  const int64_t* __restrict__ a_head, float alpha, float* loss, float* dz, int B, int L, int H, int* status,
                  const float* __restrict__ grad_scale) {
    __shared__ float red[4];
    const int b = blockIdx.x, tid = threadIdx.x;
    long pq = q_head[b], pa = a_head[b], p0 = rel_idx[2 * b], p1 = rel_idx[2 * b + 1];
    if (pq < 0 || pq >= L || pa < 0 || pa >= L || p0 < 0 || p0 >= L || p1 < 0 || p1 >= L) {
        if (tid == 0 && status) atomicOr(status, 8);
        pq = min(max(pq, 0L), (long)L - 1); pa = min(max(pa, 0L), (long)L - 1);
        p0 = min(max(p0, 0L), (long)L - 1); p1 = min(max(p1, 0L), (long)L - 1);
    }
    const float* zq = z + ((long)b * L + pq) * H;
    const float* za = z + ((long)b * L + pa) * H;
    const float* z0 = z + ((long)b * L + p0) * H;
    const float* z1 = z + ((long)b * L + p1) * H;
    float qa = 0, qq = 0, aa = 0, r01 = 0, r00 = 0, r11 = 0;
    for (int c = tid; c < H; c += blockDim.x) {
        const float q = zq[c], a = za[c], u = z0[c], v = z1[c];
        qa = fmaf(q, a, qa); qq = fmaf(q, q, qq); aa = fmaf(a, a, aa);
        r01 = fmaf(u, v, r01); r00 = fmaf(u, u, r00); r11 = fmaf(v, v, r11);
    }
    qa = block_reduce_sum(qa, red); qq = block_reduce_sum(qq, red); aa = block_reduce_sum(aa, red);
    r01 = block_reduce_sum(r01, red); r00 = block_reduce_sum(r00, red); r11 = block_reduce_sum(r11, red);
    const float nq = fmaxf(sqrtf(qq), 1e-8f), na = fmaxf(sqrtf(aa), 1e-8f);
    const float n0 = fmaxf(sqrtf(r00), 1e-8f), n1 = fmaxf(sqrtf(r11), 1e-8f);
    const float cos_h = qa / (nq * na), cos_r = r01 / (n0 * n1);
    const float w = alpha / B;
    if (tid == 0) atomicAdd(loss, w * (fmaxf(cos_h, 0.f) + 1.0f - cos_r));
    if (!dz) return;
    const float ws = w * (grad_scale ? *grad_scale : 1.0f);
    const float gh = cos_h > 0.f ? ws : 0.f;   // d relu
    const float gr = -ws;
    float* dq = dz + ((long)b * L + pq) * H;
    float* da = dz + ((long)b * L + pa) * H;
    float* d0 = dz + ((long)b * L + p0) * H;
    float* d1 = dz + ((long)b * L + p1) * H;
    for (int c = tid; c < H; c += blockDim.x) {
        const float q = zq[c], a = za[c], u = z0[c], v = z1[c];
        // d cos(x,y)/dx = y/(|x||y|) - cos * x/|x|^2   (rows may coincide -> atomics)
        atomicAdd(dq + c, gh * (a / (nq * na) - cos_h * q / (nq * nq)));
        atomicAdd(da + c, gh * (q / (nq * na) - cos_h * a / (na * na)));
        atomicAdd(d0 + c, gr * (v / (n0 * n1) - cos_r * u / (n0 * n0)));
        atomicAdd(d1 + c, gr * (u / (n0 * n1) - cos_r * v / (n1 * n1)));
    }
}

__global__ void scatter_mask_rows_kernel(const float* __restrict__ dzm, const int* __restrict__ mask_pos, float* dz, int L, int H) {
    const int b = blockIdx.x;
    float* d = dz + ((long)b * L + mask_pos[b]) * H;
    for (int c = threadIdx.x; c < H; c += blockDim.x) d[c] += dzm[(long)b * H + c];
}

__global__ void __launch_bounds__(256)
ranks_kernel(const float* __restrict__ logits, long ld, int C, const int64_t* __restrict__ label,
             const int64_t* __restrict__ pre_type, int select, int64_t* ranks, int B) {
    __shared__ float red[8];
    const int b = blockIdx.x;
    if (!row_selected(select, pre_type, b)) return;   // the row belongs to the other column group
    const float* x = logits + (long)b * ld;
    const long yl = label[b];
    const float gold = x[(yl < 0 || yl >= C) ? 0 : yl];
    float cnt = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) cnt += (x[c] > gold) ? 1.f : 0.f;
    cnt = block_reduce_sum(cnt, red);
    if (threadIdx.x == 0) ranks[b] = (int64_t)(cnt + 0.5f) + 1;
}

}  // namespace

int launch_find_mask(const int64_t* ids, int* mask_pos, int B, int L, int mask_id, int* status, cudaStream_t s) {
    find_mask_kernel<<<(B + 3) / 4, 128, 0, s>>>(ids, mask_pos, B, L, mask_id, status);
    B2_LAUNCHED();
    return 0;
}
int launch_gather_mask_rows(const bf16* z16, const int* mask_pos, bf16* zm, int B, int L, int H, cudaStream_t s) {
    B2_CHECK(H % 8 == 0, "gather_mask_rows: H %d must be a multiple of 8", H);
    gather_mask_rows_kernel<<<B, 96, 0, s>>>(z16, mask_pos, zm, L, H);
    B2_LAUNCHED();
    return 0;
}
int launch_ls_ce(const LossGroup& g, const int64_t* label, const int64_t* pre_type, float eps, float* loss, int B,
                 int* status, cudaStream_t s, const float* grad_scale) {
    B2_CHECK(g.C > 0 && g.ldd >= g.C, "ls_ce: bad class count %d / gradient leading dimension %ld", g.C, g.ldd);
    B2_CHECK(g.select == 0 || pre_type != nullptr, "ls_ce: row selection needs pre_type");
    ls_ce_kernel<<<B, 256, 0, s>>>(g, label, pre_type, eps, loss, B, status, grad_scale);
    B2_LAUNCHED();
    return 0;
}
int launch_relaxation(const float* z, const int64_t* rel_idx, const int64_t* q_head, const int64_t* a_head, float alpha,
                      float* loss, float* dz, int B, int L, int H, int* status, cudaStream_t s, const float* grad_scale) {
    relaxation_kernel<<<B, 128, 0, s>>>(z, rel_idx, q_head, a_head, alpha, loss, dz, B, L, H, status, grad_scale);
    B2_LAUNCHED();
    return 0;
}
int launch_scatter_mask_rows(const float* dzm, const int* mask_pos, float* dz, int B, int L, int H, cudaStream_t s) {
    scatter_mask_rows_kernel<<<B, 256, 0, s>>>(dzm, mask_pos, dz, L, H);
    B2_LAUNCHED();
    return 0;
}
int launch_ranks(const float* logits, long ld, int C, const int64_t* label, const int64_t* pre_type, int select,
                 int64_t* ranks, int B, cudaStream_t s) {
    ranks_kernel<<<B, 256, 0, s>>>(logits, ld, C, label, pre_type, select, ranks, B);
    B2_LAUNCHED();
    return 0;
}
