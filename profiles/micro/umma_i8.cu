// Microbenchmark + known-answer test: the separable Gaussian pass as a banded (Toeplitz) u8 x u8 -> s32 product on the
// 5th-generation tensor cores: tcgen05.mma.cta_group::1.kind::i8 with the accumulator in TMEM, operands in shared memory
// in the canonical K-major no-swizzle layout (8 rows x 16 bytes core matrices), read back with tcgen05.ld.
// One CTA computes  D[y][x] = sum_i c[i] * G[y][x + i]   (y < 128, x < 128, i < 97: the k = 97 blur of 1080p full-res mode)
// as 7 MMAs M128 x N128 x K32: A_j = G[:, 32j .. 32j+32), B_j[n][k] = c[32j + k - n] = rows (n - 32j) of ONE master band
// matrix, so B_j is just a different start address in the same shared array.
// Checks the result against the CPU, then times a long chain of the same MMAs (and of M128 x N256 x K32) per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o umma_i8 umma_i8.cu ; run: ./umma_i8
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define ROWS 128
#define GCOLS 224           // 128 outputs + 96 halo columns
#define GQ (GCOLS / 16)     // 16-byte K chunks per row
#define TROWS 480           // master band rows u in [-224, 256): 128 (256 for the N = 256 timing run) rows are read from a multiple of 32
#define TOFF 224

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: [row group of 8][16-byte K chunk][8 rows][16 bytes]
__host__ __device__ inline int g_off(int row, int col) { return ((row >> 3) * GQ + (col >> 4)) * 128 + (row & 7) * 16 + (col & 15); }
__host__ __device__ inline int t_off(int urow, int k) { return ((urow >> 3) * 2 + (k >> 4)) * 128 + (urow & 7) * 16 + (k & 15); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);          // version 1 (Blackwell), SWIZZLE_NONE
}

__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// mode 0: one Toeplitz product, result to `out` [128][128] int32.  mode 1: `iters` chains of 7 MMAs (N = ncols), timing only.
__global__ void __launch_bounds__(128, 1) k_umma(const uint8_t *__restrict__ gimg, const uint8_t *__restrict__ band, int *__restrict__ out,
                                                 int mode, int iters, int ncols) {
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned char *sG = sm;                                  // ROWS x GCOLS bytes, blocked
    unsigned char *sT = sm + ROWS * GCOLS;                   // TROWS x 32 bytes, blocked
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < ROWS * GCOLS; i += 128) sG[g_off(i / GCOLS, i % GCOLS)] = gimg[i];
    for (int i = tid; i < TROWS * 32; i += 128) sT[t_off(i / 32, i % 32)] = band[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    // instruction descriptor: D = S32, A = B = unsigned 8 bit, both K-major, N, M = 128
    const uint32_t idesc = (2u << 4) | ((uint32_t)(ncols >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t gA = smem_u32(sG), tB = smem_u32(sT);
    uint32_t parity = 0;
    const int reps = mode == 0 ? 1 : iters;
    for (int it = 0; it < reps; it++) {
        if (tid == 0) {
#pragma unroll
            for (int j = 0; j < 7; j++) {
                const uint64_t ad = make_desc(gA + 2 * j * 128, 128, GQ * 128);
                const uint64_t bd = make_desc(tB + (TOFF - 32 * j) * 32, 128, 256);     // rows (n - 32 j) of the master band
                umma_i8(tmem, ad, bd, idesc, j > 0);
            }
            umma_commit(&bar);
        }
        mbar_wait(&bar, parity);
        parity ^= 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (mode == 0) {
        // warp w reads TMEM lanes 32 w .. 32 w + 31 (= rows y), 128 columns in 4 loads of 32
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int y = 32 * warp + lane;
#pragma unroll
            for (int i = 0; i < 32; i++) out[(blockIdx.x * 128 + y) * 128 + c0 + i] = (int)r[i];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256u) : "memory");
}

static std::vector<int> gauss_taps(int k) {          // cv2.getGaussianKernel(k, 0) in 8.8 fixed point with error diffusion
    std::vector<double> kern(k);
    double sigma = 0.3 * ((k - 1) * 0.5 - 1) + 0.8, s2 = -0.5 / (sigma * sigma), sum = 0;
    for (int i = 0; i < k; i++) { double x = i - (k - 1) * 0.5; kern[i] = exp(s2 * x * x); sum += kern[i]; }
    std::vector<int> c(k, 0);
    double err = 0; int tot = 0;
    for (int i = 0; i < k / 2; i++) { double adj = kern[i] / sum * 256.0 + err; int v = (int)nearbyint(adj); err = adj - v; c[i] = c[k - 1 - i] = v; tot += 2 * v; }
    c[k / 2] = 256 - tot;
    return c;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main() {
    const int K = 97;
    std::vector<int> c = gauss_taps(K);
    std::vector<uint8_t> g(ROWS * GCOLS), band(TROWS * 32, 0);
    srand(7);
    for (auto &v : g) v = (uint8_t)(rand() & 255);
    for (int ur = 0; ur < TROWS; ur++)
        for (int k = 0; k < 32; k++) { int i = k - (ur - TOFF); if (i >= 0 && i < K) band[ur * 32 + k] = (uint8_t)c[i]; }
    uint8_t *dg, *db; int *dout;
    int nsm = 0; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaMalloc(&dg, g.size())); CK(cudaMalloc(&db, band.size())); CK(cudaMalloc(&dout, (size_t)nsm * 128 * 128 * 4));
    CK(cudaMemcpy(dg, g.data(), g.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, band.data(), band.size(), cudaMemcpyHostToDevice));
    const size_t smem = ROWS * GCOLS + TROWS * 32;
    CK(cudaFuncSetAttribute(k_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_umma<<<1, 128, smem>>>(dg, db, dout, 0, 1, 128);
    CK(cudaDeviceSynchronize());
    std::vector<int> out(128 * 128);
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int y = 0; y < 128; y++)
        for (int x = 0; x < 128; x++) {
            int ref = 0;
            for (int i = 0; i < K; i++) ref += c[i] * g[y * GCOLS + x + i];
            if (ref != out[y * 128 + x]) { if (bad < 5) printf("mismatch y=%d x=%d got %d want %d\n", y, x, out[y * 128 + x], ref); bad++; }
        }
    printf("toeplitz k=97 via tcgen05.mma.kind::i8 (7 x M128 N128 K32): %s (%ld mismatches of 16384)\n", bad ? "FAILED" : "exact", bad);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int ncols : {128, 256}) {
        const int iters = 4096;
        k_umma<<<nsm, 128, smem>>>(dg, db, dout, 1, 16, ncols);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        k_umma<<<nsm, 128, smem>>>(dg, db, dout, 1, iters, ncols);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double mmas = (double)nsm * iters * 7, macs = mmas * 128.0 * ncols * 32;
        printf("tcgen05.mma.kind::i8 M128 N%d K32 (chains of 7 + commit + wait, 1 CTA/SM): %.3f ms, %.1f clk/MMA/SM @1.965GHz, %.1f Tops (u8 MAC = 2 ops)\n",
               ncols, ms, ms * 1e-3 * 1.965e9 / (iters * 7.0), 2 * macs / (ms * 1e-3) / 1e12);
    }
    printf("reference points (profiles/micro/mma_rate.log): mma.sync IMMA.16832.U8 1132.6 Tops, IDP.4A 147.4 Tops\n");
    return bad ? 2 : 0;
}
