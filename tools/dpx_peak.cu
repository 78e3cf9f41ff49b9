// Integer / DPX issue-rate micro-benchmark for sm_100a.
// Measures lane-ops/s of dependent-free chains of DPX and plain integer instructions on all SMs.
// Used to obtain the roofline denominator for the Smith-Waterman extension kernels.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

constexpr int CH = 8;      // independent chains per thread
constexpr int ITERS = 4096;

template<int MODE>
__global__ void __launch_bounds__(256) kern(unsigned* out, unsigned seed, unsigned b, unsigned c)
{
    unsigned x[CH];
    unsigned y[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { x[i] = seed + threadIdx.x * 7 + i; y[i] = seed * 3 + i + threadIdx.x * 5; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) x[i] = __viaddmax_s16x2(x[i], b, c);
            if (MODE == 1) x[i] = __vimax3_s16x2(x[i], b, c);
            if (MODE == 2) x[i] = __viaddmax_s16x2_relu(x[i], b, c);
            if (MODE == 3) x[i] = (unsigned)__viaddmax_s32((int)x[i], (int)b, (int)c);
            if (MODE == 4) x[i] = __byte_perm(x[i], b, c);
            if (MODE == 5) { x[i] = __viaddmax_s16x2(x[i], b, c); asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(b)); } // DPX + IADD
            if (MODE == 6) { x[i] = __viaddmax_s16x2(x[i], b, c); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(b), "r"(c)); } // DPX + real IMAD
            if (MODE == 7) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(b), "r"(c)); } // IMAD only
            if (MODE == 8) x[i] = __vmaxs2(x[i], b);
            if (MODE == 9) { asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(b)); }  // IADD
            if (MODE == 10) { x[i] = __viaddmax_s16x2(x[i], b, c); y[i] = __byte_perm(y[i], b, c); } // DPX + PRMT (same pipe?)
            if (MODE == 11) { asm volatile("lop3.b32 %0, %0, %1, %2, 0xE4;" : "+r"(x[i]) : "r"(b), "r"(c)); } // LOP3
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template<int MODE>
void run(const char* name, int opsPerIter, unsigned* d, int nsm)
{
    int blocks = nsm * 8;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) kern<MODE><<<blocks, 256>>>(d, 1234u, 0x00010001u, 0x00050003u);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        kern<MODE><<<blocks, 256>>>(d, 1234u, 0x00010001u, 0x00050003u);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double laneops = (double)blocks * 256 * ITERS * CH * opsPerIter;
    printf("{\"bench\":\"%s\",\"ms\":%.4f,\"tera_lane_ops_per_s\":%.3f,\"lane_ops_per_clk_per_sm_at_1965MHz\":%.2f}\n",
           name, best, laneops / best / 1e9, laneops / (best * 1e-3) / 1.965e9 / nsm);
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", p.name, p.multiProcessorCount, p.clockRate);
    unsigned* d; CK(cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 4));
    int n = p.multiProcessorCount;
    run<0>("viaddmax_s16x2", 1, d, n);
    run<1>("vimax3_s16x2", 1, d, n);
    run<2>("viaddmax_s16x2_relu", 1, d, n);
    run<3>("viaddmax_s32", 1, d, n);
    run<4>("prmt", 1, d, n);
    run<8>("vimax_s16x2", 1, d, n);
    run<9>("iadd", 1, d, n);
    run<11>("lop3", 1, d, n);
    run<7>("imad", 1, d, n);
    run<5>("viaddmax_s16x2+iadd(2 ops)", 2, d, n);
    run<6>("viaddmax_s16x2+imad(2 ops)", 2, d, n);
    run<10>("viaddmax_s16x2+prmt(2 ops)", 2, d, n);
    return 0;
}
