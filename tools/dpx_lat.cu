// Latency / ILP probe for DPX chains at low occupancy (2 warps per scheduler = 8 warps/SM).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
constexpr int ITERS = 8192;
template<int CH, int MODE>
__global__ void kern(unsigned* out, unsigned seed, unsigned b, unsigned c)
{
    unsigned x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = seed + threadIdx.x * 7 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) x[i] = __viaddmax_s16x2(x[i], b, c);
            if (MODE == 1) { x[i] = __vimax3_s16x2(x[i], b, c); }
            if (MODE == 2) { asm volatile("prmt.b32 %0, %0, %1, 0x80c4;" : "+r"(x[i]) : "r"(b)); }
            if (MODE == 3) { x[i] = __viaddmax_s16x2(x[i], b, c); x[i] = x[i] - b; }   // DPX -> IADD -> DPX chain
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int CH, int MODE>
void run(const char* name, int warps_per_sm, int ops, unsigned* d, int nsm, double clk)
{
    int threads = warps_per_sm * 32;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) kern<CH, MODE><<<nsm, threads>>>(d, 1234u, 0x00010001u, 0x00050003u);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); kern<CH, MODE><<<nsm, threads>>>(d, 1234u, 0x00010001u, 0x00050003u);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double cycles = best * 1e-3 * clk;
    double per_sched_ops = (double)ITERS * CH * ops * warps_per_sm / 4.0;   // warp-instr per scheduler
    printf("%-22s warps/SM=%2d chains=%d  cycles/warp-instr/scheduler=%.2f  (lat est %.1f cyc if 1 chain 1 warp)\n", name, warps_per_sm, CH,
           cycles / per_sched_ops, cycles / ((double)ITERS * ops));
}
int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    unsigned* d; CK(cudaMalloc(&d, (size_t)p.multiProcessorCount * 1024 * 4));
    int n = p.multiProcessorCount; double clk = p.clockRate * 1e3;
    run<1, 0>("viaddmax", 4, 1, d, n, clk);
    run<2, 0>("viaddmax", 4, 1, d, n, clk);
    run<4, 0>("viaddmax", 4, 1, d, n, clk);
    run<1, 0>("viaddmax", 8, 1, d, n, clk);
    run<2, 0>("viaddmax", 8, 1, d, n, clk);
    run<3, 0>("viaddmax", 8, 1, d, n, clk);
    run<4, 0>("viaddmax", 8, 1, d, n, clk);
    run<8, 0>("viaddmax", 8, 1, d, n, clk);
    run<1, 1>("vimax3", 4, 1, d, n, clk);
    run<1, 2>("prmt", 4, 1, d, n, clk);
    run<1, 3>("viaddmax+iadd chain", 4, 2, d, n, clk);
    run<2, 3>("viaddmax+iadd chain", 8, 2, d, n, clk);
    run<4, 3>("viaddmax+iadd chain", 8, 2, d, n, clk);
    return 0;
}
