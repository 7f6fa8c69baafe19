// Microbenchmark: per-SM issue rate of the integer instructions K1 is made of (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_rate alu_rate.cu ; run: ./alu_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed) {
    uint32_t v[8];
    for (int i = 0; i < 8; i++) v[i] = seed * 7 + threadIdx.x + i;
    const uint32_t c = seed * 3 + 1, d = seed + 5;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) v[i] = v[i] + c + d;                                  // IADD3
            else if (MODE == 1) v[i] = (v[i] & c) ^ d;                           // LOP3
            else if (MODE == 2) v[i] = __byte_perm(v[i], c, 0x2103 + (d & 0));   // PRMT
            else if (MODE == 3) v[i] = v[i] * c + d;                             // IMAD
            else if (MODE == 4) v[i] = __funnelshift_r(v[i], c, 7);              // SHF
            else if (MODE == 5) v[i] = v[i] > c ? d : v[i] + 1;                  // ISETP + SEL-ish
            else if (MODE == 6) v[i] = __dp2a_lo(c, v[i], d);                    // IDP.2A
            else if (MODE == 7) { v[i] = v[i] + c + d; v[(i + 1) & 7] = v[(i + 1) & 7] * c + d; }   // IADD3 + IMAD mix
        }
    }
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char *name, int per) {
    uint32_t *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, 2);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lanes = 148.0 * 8 * 256 * ITERS * 8 * per;
    printf("%-14s %.3f ms  %.1f lanes/clk/SM @1.965GHz\n", name, ms, lanes / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(d);
}
int main() {
    run<0>("IADD3", 1); run<1>("LOP3", 1); run<2>("PRMT", 1); run<3>("IMAD", 1); run<4>("SHF", 1); run<5>("ISETP+SEL", 1);
    run<6>("IDP.2A", 1); run<7>("IADD3+IMAD", 2);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
