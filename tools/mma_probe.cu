// Issue-rate probe of the legacy warp-level tensor path on sm_100a: mma.sync.m16n8k8 tf32 (and m16n8k16 bf16 for
// comparison), 8 independent accumulator sets per warp, W warps per SM.  Prints MAC/clk/SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void probe(float* out, int iters, long long* cycles) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 11, b0 = 5, b1 = 9;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int kind = 0; kind < 2; ++kind)
        for (int warps : {4, 8, 16, 32}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (kind == 0) probe<0><<<148, warps * 32>>>(out, iters, cyc); else probe<1><<<148, warps * 32>>>(out, iters, cyc);
                cudaDeviceSynchronize();
            }
            long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            const double macs = (double)iters * 8 * warps * 16 * 8 * (kind ? 16 : 8);
            printf("%s warps/SM %2d: %.1f MAC/clk/SM (%lld cycles) %s\n", kind ? "bf16 m16n8k16" : "tf32 m16n8k8 ", warps, macs / h, h,
                   cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
