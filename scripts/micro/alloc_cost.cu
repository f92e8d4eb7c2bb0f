// Microbenchmark (profiling aid): host-side cost of the CUDA resource calls a small plan makes.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <vector>
using clk = std::chrono::steady_clock;
static double ms(clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); }
int main()
{
    cudaFree(0);
    for (int round = 0; round < 4; ++round) {
        auto t0 = clk::now();
        cudaDeviceProp prop;
        cudaGetDeviceProperties(&prop, 0);
        auto t1 = clk::now();
        int sm = 0;
        cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
        auto t2 = clk::now();
        cudaStream_t st;
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        auto t3 = clk::now();
        std::vector<void *> p(8);
        for (auto &q : p) cudaMalloc(&q, 200000);
        auto t4 = clk::now();
        for (auto &q : p) cudaFree(q);
        auto t5 = clk::now();
        for (auto &q : p) cudaMallocAsync(&q, 200000, st);
        auto t6 = clk::now();
        for (auto &q : p) cudaFreeAsync(q, st);
        cudaStreamSynchronize(st);
        auto t7 = clk::now();
        void *big;
        cudaMalloc(&big, 64u << 20);
        auto t8 = clk::now();
        cudaFree(big);
        auto t9 = clk::now();
        cudaStreamDestroy(st);
        auto t10 = clk::now();
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventDestroy(e);
        auto t11 = clk::now();
        printf("round %d: getProps %.3f  getAttr %.3f  streamCreate %.3f  8xMalloc(200K) %.3f  8xFree %.3f  8xMallocAsync %.3f  8xFreeAsync+sync %.3f  "
               "malloc64M %.3f  free64M %.3f  streamDestroy %.3f  event %.3f ms\n",
               round, ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), ms(t4, t5), ms(t5, t6), ms(t6, t7), ms(t7, t8), ms(t8, t9), ms(t9, t10), ms(t10, t11));
    }
    return 0;
}
