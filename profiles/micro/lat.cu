// Latency microbenchmarks for the fp64 / shared-memory / barrier primitives the tile kernels use.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* clk, int nthreads_active) {
    extern __shared__ double sm[];
    double x = out[threadIdx.x & 31] + 1.0, y = 1.0000001;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
    __syncthreads();
    long long t0, t1;
    const int N = 256;
    // dependent DMUL chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x * y;
    t1 = clock64();
    if (threadIdx.x == 0) clk[0] = (t1 - t0);
    // dependent DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x + y;
    t1 = clock64();
    if (threadIdx.x == 0) clk[1] = (t1 - t0);
    // dependent division chain
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) x = 1.0 / x + 0.5;
    t1 = clock64();
    if (threadIdx.x == 0) clk[2] = (t1 - t0);
    // dependent shuffle chain (double)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __shfl_down_sync(0xffffffffu, x, 8);
    t1 = clock64();
    if (threadIdx.x == 0) clk[3] = (t1 - t0);
    // dependent shared-memory load chain (pointer chase through an index array)
    int* ism = (int*)(sm + 2048);
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) ism[i] = (i * 37 + 11) & 2047;
    __syncthreads();
    int p = threadIdx.x & 2047;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) p = ism[p];
    t1 = clock64();
    if (threadIdx.x == 0) clk[4] = (t1 - t0);
    // barrier
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) clk[5] = (t1 - t0);
    // barrier with one warp doing a smem store+load between
    t0 = clock64();
    for (int i = 0; i < N; i++) {
        if (threadIdx.x < 32) sm[threadIdx.x] = x + i;
        __syncthreads();
        x += sm[(threadIdx.x + 1) & 31];
        __syncthreads();
    }
    t1 = clock64();
    if (threadIdx.x == 0) clk[6] = (t1 - t0);
    // fp64 throughput: 8 independent chains per thread
    double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) {
        a0 = a0 * y + 1.0; a1 = a1 * y + 1.0; a2 = a2 * y + 1.0; a3 = a3 * y + 1.0;
        a4 = a4 * y + 1.0; a5 = a5 * y + 1.0; a6 = a6 * y + 1.0; a7 = a7 * y + 1.0;
    }
    t1 = clock64();
    if (threadIdx.x == 0) clk[7] = (t1 - t0);
    out[threadIdx.x] = x + p + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
    double* out; long long* clk;
    cudaMalloc(&out, 1024 * 8); cudaMemset(out, 0, 1024 * 8);
    cudaMalloc(&clk, 64);
    const char* names[] = {"DMUL dep", "DADD dep", "1/x+0.5 dep", "shfl.f64 dep", "LDS dep", "barrier", "store+bar+load+bar", "8xDFMA indep (per 8)"};
    for (int nt : {32, 256, 1024}) {
        k<<<1, nt, 48 * 1024>>>(out, clk, nt);
        cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, clk, 64, cudaMemcpyDeviceToHost);
        printf("threads=%d:", nt);
        for (int i = 0; i < 8; i++) printf("  %s %.1f", names[i], h[i] / 256.0);
        printf("\n");
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
