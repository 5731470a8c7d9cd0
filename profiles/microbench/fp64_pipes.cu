// FP64 microbenchmark: DFMA pipe vs DMMA (mma.sync m8n8k4 f64) throughput and power on B200.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = fma(x[j], a, b);
    double s = 0;
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a, double b) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main(int argc, char** argv) {
    const double secs = argc > 1 ? atof(argv[1]) : 2.0;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256, iters = 1 << 14;
    double* buf;
    cudaMalloc(&buf, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 2; ++which) {
        double spent = 0, best = 0, last = 0;
        int reps = 0;
        while (spent < secs) {
            cudaEventRecord(e0);
            if (which == 0) k_dfma<<<blocks, threads>>>(buf, iters, 0.999999, 1e-9);
            else k_dmma<<<blocks, threads>>>(buf, iters, 1e-3, 1e-3);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            spent += ms * 1e-3; ++reps;
            const double flops = which == 0 ? 2.0 * 8 * iters * (double)blocks * threads
                                            : 2.0 * 8 * 8 * 4 * 8.0 * iters * (double)blocks * (threads / 32);
            last = flops / (ms * 1e-3) / 1e12;
            if (last > best) best = last;
        }
        printf("%s: best %.2f TFLOP/s, sustained (last) %.2f TFLOP/s after %.1f s, %d launches\n", which ? "DMMA m8n8k4" : "DFMA", best, last, spent, reps);
        fflush(stdout);
        system("nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader");
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
