// micro-benchmark: cost of a grid-wide barrier on B200 (cooperative launch, 1 CTA/SM)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sync_bench sync_bench.cu && ./sync_bench
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void k_cg(int n, float *sink)
{
    cg::grid_group grid = cg::this_grid();
    float a = threadIdx.x;
    for (int i = 0; i < n; ++i) {
        a = a * 1.0001f + 1.0f;
        grid.sync();
    }
    if (a == 12345.f) *sink = a;
}

// custom sense-free barrier: monotonically increasing counter, one arrival per CTA
__device__ __forceinline__ void grid_bar(unsigned int *ctr, unsigned int target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

__global__ void k_custom(int n, float *sink, unsigned int *ctr)
{
    float a = threadIdx.x;
    for (int i = 0; i < n; ++i) {
        a = a * 1.0001f + 1.0f;
        grid_bar(ctr, (unsigned int)(i + 1) * gridDim.x);
    }
    if (a == 12345.f) *sink = a;
}

// same, but every thread also has a store in flight before the barrier (the release has to cover it)
__global__ void k_custom_st(int n, float *buf, unsigned int *ctr)
{
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    float a = threadIdx.x;
    for (int i = 0; i < n; ++i) {
        a = a * 1.0001f + 1.0f;
        buf[tid + (size_t)(i & 1) * gridDim.x * blockDim.x] = a;
        grid_bar(ctr, (unsigned int)(i + 1) * gridDim.x);
    }
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *sink;
    unsigned int *ctr;
    cudaMalloc(&sink, 2 * sizeof(float) * 1024 * 256);
    cudaMalloc(&ctr, 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int threads : {256, 1024}) {
        for (int grid : {sms / 2, sms}) {
            for (int n : {0, 100}) {
                float ms[3];
                for (int variant = 0; variant < 3; ++variant) {
                    void *a0[] = {&n, &sink};
                    void *a1[] = {&n, &sink, &ctr};
                    for (int rep = 0; rep < 3; ++rep) {
                        cudaMemset(ctr, 0, 256);
                        cudaEventRecord(e0);
                        if (variant == 0) cudaLaunchCooperativeKernel((void *)k_cg, dim3(grid), dim3(threads), a0, 0, 0);
                        if (variant == 1) cudaLaunchCooperativeKernel((void *)k_custom, dim3(grid), dim3(threads), a1, 0, 0);
                        if (variant == 2) cudaLaunchCooperativeKernel((void *)k_custom_st, dim3(grid), dim3(threads), a1, 0, 0);
                        cudaEventRecord(e1);
                        cudaEventSynchronize(e1);
                        cudaEventElapsedTime(&ms[variant], e0, e1);
                    }
                }
                printf("threads %4d grid %3d syncs %3d : cg %.2f us  custom %.2f us  custom+store %.2f us  (%s)\n", threads, grid,
                       n, ms[0] * 1e3, ms[1] * 1e3, ms[2] * 1e3, cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    return 0;
}
