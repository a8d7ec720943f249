// Latency probes on B200 (sm_100a): dependent DADD chain, ordered sum from shared memory, L2 pointer chase,
// software grid barrier.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false lat.cu -o lat
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
__global__ void k_dadd(double* out, int n, long long* clk) {
    double s = out[0];
    const double a = out[1];
    long long c0 = clock64();
    for (int i = 0; i < n; ++i) s += a;
    long long c1 = clock64();
    out[2] = s;
    if (threadIdx.x == 0) clk[0] = c1 - c0;
}
__global__ void k_smemsum(double* out, int n, long long* clk) {
    __shared__ double r[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) r[i] = out[1] + i;
    __syncthreads();
    double s = 0;
    long long c0 = clock64();
#pragma unroll 8
    for (int i = 0; i < n; ++i) s += r[i & 1023];
    long long c1 = clock64();
    out[3] = s;
    if (threadIdx.x == 0) clk[1] = c1 - c0;
}
__global__ void k_chase(const int* p, int n, int* out, long long* clk) {
    int j = 0;
    long long c0 = clock64();
    for (int i = 0; i < n; ++i) j = __ldcg(p + j);
    long long c1 = clock64();
    out[0] = j;
    clk[2] = c1 - c0;
}
__device__ __forceinline__ void grid_sync(unsigned* counter, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        while (*((volatile unsigned*)counter) < target) {}
        __threadfence();
    }
    __syncthreads();
}
__global__ void k_bar(unsigned* counter, int n, long long* clk) {
    unsigned target = 0;
    long long c0 = clock64();
    for (int i = 0; i < n; ++i) grid_sync(counter, target);
    long long c1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) clk[3] = c1 - c0;
}
int main() {
    double* out; long long* clk; int* p; int* o; unsigned* ctr;
    cudaMalloc(&out, 64); cudaMalloc(&clk, 64); cudaMalloc(&o, 64); cudaMalloc(&ctr, 4);
    double h[2] = {1.0, 1e-3}; cudaMemcpy(out, h, 16, cudaMemcpyHostToDevice);
    const int N = 1 << 22;  // 16 MB of ints: L2 resident
    int* hp = new int[N];
    for (int i = 0; i < N; ++i) hp[i] = (int)(((long long)i * 1000003LL + 12345) % N);
    cudaMalloc(&p, N * 4); cudaMemcpy(p, hp, N * 4, cudaMemcpyHostToDevice);
    long long hc[8];
    for (int rep = 0; rep < 2; ++rep) {
        k_dadd<<<1, 32>>>(out, 4096, clk);
        k_smemsum<<<1, 32>>>(out, 4096, clk);
        k_chase<<<1, 1>>>(p, 2000, o, clk);
        cudaMemset(ctr, 0, 4);
        int n = 200; void* args[] = {&ctr, &n, &clk};
        cudaLaunchCooperativeKernel((void*)k_bar, dim3(148), dim3(1024), args, 0, 0);
        cudaDeviceSynchronize();
        cudaMemcpy(hc, clk, 64, cudaMemcpyDeviceToHost);
        printf("dadd dep %.2f cyc/op | smem ordered sum %.2f cyc/add | L2 chase %.1f cyc/load | grid barrier (148x1024) %.0f cyc  [%s]\n",
               hc[0] / 4096.0, hc[1] / 4096.0, hc[2] / 2000.0, hc[3] / 200.0, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
