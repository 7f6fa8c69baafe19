// Microbenchmark: tensor-memory read throughput per SM (tcgen05.ld -> registers), the quantity that bounds a tcgen05
// pipeline whose accumulators have to come back to the ALUs once per output element (the Gaussian + temporal kernel of
// k_umma.cu reads 480 accumulator columns per 128 x 128 pixel tile and frame).
// Variants: 32x32b.x8 / .x16 / .x32, with and without .pack::16b (two 16-bit columns per register), 4 .. 32 warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_rate ldtm_rate.cu ; run: ./ldtm_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__device__ __forceinline__ uint32_t ld_once(uint32_t taddr) {
    uint32_t r[32];
    if (MODE == 0) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        return r[0] ^ r[1] ^ r[2] ^ r[3] ^ r[4] ^ r[5] ^ r[6] ^ r[7];
    } else if (MODE == 1) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) x ^= r[i];
        return x;
    } else if (MODE == 2) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < 32; i++) x ^= r[i];
        return x;
    } else if (MODE == 3) {      // 16 columns -> 8 registers
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        return r[0] ^ r[1] ^ r[2] ^ r[3] ^ r[4] ^ r[5] ^ r[6] ^ r[7];
    } else {                     // 32 columns -> 16 registers
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t x = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) x ^= r[i];
        return x;
    }
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_ldtm(uint32_t *out, long long *cycles, int iters, int pattern_check) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const uint32_t lanebase = (uint32_t)(32 * (warp & 3)) << 16;
    if (pattern_check && warp < 4) {
        // write column c of lane l = (l << 16) | (c + 0x100 * (c & 1)) so that the packed halves can be told apart
        for (int c = 0; c < 64; c++) {
            const uint32_t v = ((uint32_t)(32 * warp + lane) << 16) | (uint32_t)(c * 3 + 1);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + lanebase + c), "r"(v) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) acc ^= ld_once<MODE>(tmem + lanebase + 32 * ((warp >> 2) & 7) + ((it & 1) << 8));
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (pattern_check && blockIdx.x == 0 && warp == 0) {          // .sync.aligned: the whole warp executes the load
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tmem));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (lane == 1)
            for (int i = 0; i < 8; i++) out[4096 + i] = r[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int MODE>
static void run(const char *name, int cols, int bytes_per_lane, int threads) {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4 + 65536); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2048;
    k_ldtm<MODE><<<148, threads>>>(out, cyc, 64, 0);
    cudaDeviceSynchronize();
    k_ldtm<MODE><<<148, threads>>>(out, cyc, iters, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; i++) c += (double)h[i];
    c /= 148;
    const double warps = threads / 32.0;
    printf("%-28s %2d warps: %7.1f cyc/ld/warp-slot, %6.1f columns*lanes*4B per cycle per SM (TMEM bytes), %6.1f register bytes per cycle\n", name,
           threads / 32, c / iters, warps * 32 * cols * 4 * iters / c, warps * 32 * bytes_per_lane * iters / c);
    fflush(stdout);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {128, 256, 512, 1024}) {
        run<0>("32x32b.x8", 8, 32, threads);
        run<1>("32x32b.x16", 16, 64, threads);
        run<2>("32x32b.x32", 32, 128, threads);
        run<3>("32x32b.x8.pack::16b", 16, 32, threads);
        run<4>("32x32b.x16.pack::16b", 32, 64, threads);
    }
    // which half is which with pack::16b: lane 1, columns 0..15 hold 3c+1
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4 + 65536); cudaMalloc(&cyc, 148 * 8);
    k_ldtm<0><<<1, 128>>>(out, cyc, 1, 1);
    cudaDeviceSynchronize();
    uint32_t r[8]; cudaMemcpy(r, out + 4096, sizeof(r), cudaMemcpyDeviceToHost);
    printf("pack::16b registers of lane 1 over columns holding 3c+1 (c = 0..15):");
    for (int i = 0; i < 8; i++) printf(" %08x", r[i]);
    printf("\n");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
