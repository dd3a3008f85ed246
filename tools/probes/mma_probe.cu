// Throughput of legacy mma.sync m16n8k8 tf32 on this GPU (per SM), to decide whether the rank-J
// projection of the row filter is worth moving to the tensor path.  nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void mma_loop(float* out, int iters) {
    float c[4][4] = {};
    unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = 7, a3 = 9, b0 = threadIdx.x ^ 5, b1 = 11;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            asm volatile(
                "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                : "+f"(c[k][0]), "+f"(c[k][1]), "+f"(c[k][2]), "+f"(c[k][3])
                : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0;
    for (int k = 0; k < 4; ++k) s += c[k][0] + c[k][1] + c[k][2] + c[k][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float* d;
    cudaMalloc(&d, 148 * 1024 * 4);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    for (int threads : {128, 256, 512, 1024}) {
        const int iters = 20000;
        mma_loop<<<p.multiProcessorCount, threads>>>(d, 10);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        mma_loop<<<p.multiProcessorCount, threads>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double mmas_per_sm = (double)iters * 4 * (threads / 32);
        const double clk = (double)p.clockRate * 1e3;  // Hz (base value reported by the runtime)
        printf("threads/SM %4d: %.3f ms, %.2f ns per MMA per SM, %.1f TFLOP/s tf32 dense (chip), err=%s\n", threads, ms,
               ms * 1e6 / mmas_per_sm, mmas_per_sm * p.multiProcessorCount * 2.0 * 16 * 8 * 8 / (ms * 1e-3) / 1e12,
               cudaGetErrorString(cudaGetLastError()));
        (void)clk;
    }
    return 0;
}
