// fp64_latency.cu — how much instruction-level parallelism the FP64 pipe of a B200 SM needs at a given occupancy.
// Each thread runs ILP independent dependent-DFMA chains; blocks of 128 threads, WPS warps per scheduler resident.
// Prints the fraction of the saturated DFMA rate for ILP = 1..6 and 1..6 warps per scheduler.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chains(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = threadIdx.x * 1e-9 + k;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int k = 0; k < ILP; k++) x[k] = fma(x[k], a, b);
        }
    }
    double s = 0.;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += x[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
double run(int warps_per_scheduler, int sms, double* out) {
    // one block = 4 warps = one warp per scheduler; `warps_per_scheduler` blocks per SM (dynamic smem pins the residency)
    const int blocks = sms * warps_per_scheduler;
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    size_t smem = (220 * 1024) / warps_per_scheduler - 2048;
    cudaFuncSetAttribute(chains<ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    chains<ILP><<<blocks, 128, smem>>>(out, 16, 1.0000001, 1e-9);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        chains<ILP><<<blocks, 128, smem>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double flops = 2.0 * ILP * 16.0 * iters * 128.0 * blocks;
    return flops / (best * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double* out;
    cudaMalloc(&out, sizeof(double) * 128 * p.multiProcessorCount * 16);
    printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    printf("TFLOP/s (FMA = 2) by warps per scheduler (rows) and independent DFMA chains per thread (columns 1..6)\n");
    for (int w = 1; w <= 6; w++) {
        printf("w=%d:", w);
        printf(" %6.2f", run<1>(w, p.multiProcessorCount, out));
        printf(" %6.2f", run<2>(w, p.multiProcessorCount, out));
        printf(" %6.2f", run<3>(w, p.multiProcessorCount, out));
        printf(" %6.2f", run<4>(w, p.multiProcessorCount, out));
        printf(" %6.2f", run<5>(w, p.multiProcessorCount, out));
        printf(" %6.2f", run<6>(w, p.multiProcessorCount, out));
        printf("\n");
    }
    return 0;
}
