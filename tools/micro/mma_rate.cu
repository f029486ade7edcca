// Legacy mma.sync issue rate on sm_100a: bf16 m16n8k16 and tf32 m16n8k8, cycles per instruction per SM sub-core.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int KIND, int ILP>
__global__ void k(float* out, int iters, long long* cyc) {
    float c[ILP][4];
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = 2, a2 = 3, a3 = 4, b0 = 5, b1 = 6;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int KIND, int ILP>
void run(const char* name, int warps) {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<KIND, ILP><<<148, warps * 32>>>(out, iters, cyc);
    k<KIND, ILP><<<148, warps * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_sub = (double)h / ((double)iters * ILP * warps / 4.0);
    printf("%s warps/SM=%d ILP=%d: %.2f cycles per mma per sub-core (%.1f dense TFLOP/s at 1.9 GHz x 148 SMs)\n", name, warps, ILP, per_sub,
           (KIND == 0 ? 4096.0 : 1024.0) * 2 * 4 / per_sub * 1.9e9 * 148 / 1e12);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0, 1>("bf16 m16n8k16", 4); run<0, 4>("bf16 m16n8k16", 4); run<0, 8>("bf16 m16n8k16", 4); run<0, 8>("bf16 m16n8k16", 8); run<0, 8>("bf16 m16n8k16", 16);
    run<1, 1>("tf32 m16n8k8 ", 4); run<1, 4>("tf32 m16n8k8 ", 4); run<1, 8>("tf32 m16n8k8 ", 4); run<1, 8>("tf32 m16n8k8 ", 8); run<1, 8>("tf32 m16n8k8 ", 16);
    return 0;
}
