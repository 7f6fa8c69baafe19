// Microbenchmark: issue rate of legacy warp-level MMA (IMMA.16832.U8, HMMA.16816.F32), IDP.4A and LDSM on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu ; run: ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(int *out, uint32_t seed) {
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    int c[4][4];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
    float f[4][4];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) f[i][j] = 0.f;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (MODE == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(f[i][0]), "+f"(f[i][1]), "+f"(f[i][2]), "+f"(f[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else {
#pragma unroll
                for (int j = 0; j < 4; j++) c[i][j] = (int)__dp4a(a0 + j, b0 + i, (unsigned)c[i][j]);
            }
        }
    }
    int s = 0;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += c[i][j] + (int)f[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char *name, double ops_per_inst, int inst_per_iter) {
    int *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, 2);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winst = 148.0 * 8 * 8 * ITERS * inst_per_iter;     // warp instructions
    printf("%s: %.3f ms, %.1f warp-inst/us total, %.3f warp-inst/clk/SM @1.965GHz, %.1f Tops\n", name, ms,
           winst / ms / 1e3, winst / (ms * 1e-3) / 148 / 1.965e9, winst * ops_per_inst / (ms * 1e-3) / 1e12);
    cudaFree(d);
}
int main() {
    run<0>("IMMA.16832.U8", 2.0 * 16 * 8 * 32, 4);
    run<1>("HMMA.16816.F32", 2.0 * 16 * 8 * 16, 4);
    run<2>("IDP.4A", 2.0 * 4 * 32, 16);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
