// fp64_nan.cu — does the FP64 pipe of a B200 SM slow down when some lanes of a warp carry NaN (or Inf) operands?
// Why it matters here: the idle lanes of the step kernel (host slot, padding) run the planet code on dummies; without the
// dummies the Kepler drift of the host slot is NaN throughout, and the whole kernel ran 21 % slower (profiles/r1_variants.md).
// Answer (profiles/r1_fp64_nan.txt): no — 33.5 TFLOP/s in every case; the slowdown was the Stumpff range-reduction loop
// running hundreds of trips on the host slot's garbage (ncu: 1 313 instead of 492 FP64 instructions per warp-step in the drifts).
// Dependent-DFMA chains, 4 per thread, 4 warps per scheduler; the lanes selected by `mask` start from the special value.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_nan fp64_nan.cu
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void chains(double* out, int iters, double a, double b, unsigned mask, double special) {
    double x[4];
    const bool sp = ((1u << (threadIdx.x & 7)) & mask) != 0;
#pragma unroll
    for (int k = 0; k < 4; k++) x[k] = sp ? special : threadIdx.x * 1e-9 + k;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int k = 0; k < 4; k++) x[k] = fma(x[k], a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = (x[0] + x[1]) + (x[2] + x[3]);
}

static double run(int sms, double* out, unsigned mask, double special) {
    const int blocks = sms * 4, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    chains<<<blocks, 128>>>(out, 16, 1.0000001, 1e-9, mask, special);
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        chains<<<blocks, 128>>>(out, iters, 1.0000001, 1e-9, mask, special);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return 2.0 * 4 * 16.0 * iters * 128.0 * blocks / (best * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double* out;
    cudaMalloc(&out, sizeof(double) * 128 * p.multiProcessorCount * 4);
    printf("%s: TFLOP/s of dependent DFMA chains (4 per thread, 4 warps per scheduler)\n", p.name);
    printf("all lanes finite            %6.2f\n", run(p.multiProcessorCount, out, 0u, 0.));
    printf("lane 0 of every 8 NaN       %6.2f\n", run(p.multiProcessorCount, out, 1u, nan("")));
    printf("all lanes NaN               %6.2f\n", run(p.multiProcessorCount, out, 0xffu, nan("")));
    printf("lane 0 of every 8 +Inf      %6.2f\n", run(p.multiProcessorCount, out, 1u, INFINITY));
    printf("lane 0 of every 8 subnormal %6.2f\n", run(p.multiProcessorCount, out, 1u, 1e-310));
    return 0;
}
