// Microbenchmark: per-SM throughput of the FP64-related instructions the temporal stage uses (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cvt_rate cvt_rate.cu ; run: ./cvt_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int seed, double alpha) {
    double d[8];
    float f[8];
    int n[8];
    for (int i = 0; i < 8; i++) { d[i] = seed * 0.37 + threadIdx.x * 0.001 + i; f[i] = (float)d[i]; n[i] = seed + i + threadIdx.x; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { d[i] = __int2double_rn(n[i]); n[i] = __double2hiint(d[i]) + it; }          // I2F.F64 (+ IADD)
            else if (MODE == 1) { f[i] = __double2float_rn(d[i]); d[i] = __hiloint2double(__double2hiint(d[i]), __float_as_int(f[i])); }   // F2F.F32.F64
            else if (MODE == 2) { d[i] = __fma_rn(d[i], alpha, 1.0); }                                  // DFMA
            else if (MODE == 3) { d[i] = __dadd_rn(d[i], alpha); }                                      // DADD
            else if (MODE == 4) { d[i] = (double)f[i]; f[i] = __int_as_float(__double2loint(d[i]) + it); }   // F2F.F64.F32
            else if (MODE == 5) { n[i] = __double2int_rn(d[i]); d[i] = __hiloint2double(__double2hiint(d[i]), n[i]); }   // F2I.F64
            else if (MODE == 6) { f[i] = __fadd_rn(f[i], 1.5f); }                                       // FADD (reference)
            else if (MODE == 7) { n[i] = __float2int_rn(f[i]); f[i] = __int_as_float(n[i] + it); }      // F2I.F32
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += d[i] + f[i] + n[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char *name) {
    double *d; cudaMalloc(&d, 148 * 8 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, 1, 0.9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, 2, 0.9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lanes = 148.0 * 8 * 256 * ITERS * 8;
    printf("%-14s %.3f ms  %.1f lanes/clk/SM @1.965GHz\n", name, ms, lanes / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(d);
}
int main() {
    run<0>("I2F.F64.S32"); run<1>("F2F.F32.F64"); run<2>("DFMA"); run<3>("DADD"); run<4>("F2F.F64.F32");
    run<5>("F2I.S32.F64"); run<6>("FADD"); run<7>("F2I.S32.F32");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
