// Gaussian blur (3 <= k <= 97) + temporal stage on the 5th-generation tensor cores: both passes of the separable
// 8.8 fixed-point Gaussian are banded (Toeplitz) u8 x u8 -> s32 products issued as tcgen05.mma.kind::i8 with the
// accumulators in tensor memory (SASS: UTCIMMA / LDTM), the gray tile staged by TMA (UTMALDG, 128-byte swizzle).
// Replaces cv2.GaussianBlur + mask_off_areas + VideoFrame.diff/threshold + accumulateWeighted of the reference
// (find_motion/find_motion.py:494, 619-635, 246-257, 651-659) in ONE pass over the image: the
// 16-bit horizontal sums never leave the SM (the two-kernel mma.sync path of k_wide.cu sends them through HBM as two
// byte planes, 2.03x the algorithmic traffic).  Integer arithmetic throughout: bit-exact.
//
// Input: the gray plane with a BORDER_REFLECT_101 apron of 48 pixels materialised (k_pad_gray below: from the BGR
// frames in full-resolution mode, from the resized gray plane otherwise), so no tile needs border logic.
// CTA = 128 x 128 output pixels of one stream, 1024 threads, one CTA per SM, walking the T frames of the call:
//   TMA: gray rows Y0-48 .. Y0+175 x columns X0-48 .. X0+207 -> shared, K-major SWIZZLE_128B (two 128-byte panels)
//   MMA1 (7 x M128 N224 K32): D1[x][y] = sum_q band[x - 32j][q] * gray[y][q]      horizontal pass, TRANSPOSED result:
//        tensor-memory lane = output column, tensor-memory column = row, so a thread reading its lane gets 16
//        consecutive ROWS of one column = one 16-byte chunk of the K-major operand of the vertical pass
//   split: LDTM -> low / high byte planes of the 16-bit sums -> shared (K-major, no swizzle)
//   MMA2 (2 x 7 x M128 N128 K32): D2lo / D2hi[x][y] = sum_q hor_lo/hi[x][q] * band[y - 32j][q]       vertical pass
//   epilogue: LDTM -> blur = (256 hi + lo + 32768) >> 16 -> mask -> bg8, threshold bit (warp ballot = one 32-pixel
//        word of the bit plane), float64 background update; the background of the thread's 16 pixels stays in
//        registers across the T frames.  MMA1 of frame t+1 is issued before the epilogue of frame t and runs under it.
// The band operand is ONE constant matrix: band[u][q] = c[q - u] (c zero-padded to 97 taps), of which every MMA reads
// 128 rows starting at a multiple of 32.
#include <cuda.h>

#include <vector>

#include "fm_common.cuh"

#define UB_PAD 48                     // apron of the padded gray plane = largest kernel radius
#define UB_T 128                      // output tile edge
#define UB_IN (UB_T + 2 * UB_PAD)     // 224 input rows / columns per tile
#define UB_THREADS 1024
#define UB_GSTAGE (2 * UB_IN * 128)   // bytes of one gray stage: two 128-byte-wide panels of 224 rows
#define UB_YQ (UB_IN / 16)            // 16-row chunks of the horizontal sums
#define UB_PLANE (UB_T * UB_IN)       // bytes of one byte plane of the horizontal sums
#define UB_BROWS 352                  // band rows u in [-224, 128)
#define UB_BOFF 224
#define UB_SMEM (2 * UB_GSTAGE + 2 * UB_PLANE + UB_BROWS * 32 + 1024 /* alignment */ + 64)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled fm_tma_encoder();     // k_fused.cu

__device__ __forceinline__ uint32_t usmem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ub_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(usmem(bar)), "r"(count));
}
__device__ __forceinline__ void ub_mbar_expect(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(usmem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ub_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nUB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra UB_DONE;\nbra UB_WAIT;\nUB_DONE:\n}\n"
        ::"r"(usmem(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void ub_tma_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(usmem(dst)), "l"(map), "r"(usmem(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// shared-memory matrix descriptors (sm_100 version bit set): K-major, no swizzle (LBO = stride of the two 16-byte K
// chunks, SBO = stride of 8-row groups) and K-major SWIZZLE_128B (rows of 128 bytes, SBO = 1024)
__device__ __forceinline__ uint64_t ub_desc_plain(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ull << 46);
}
__device__ __forceinline__ uint64_t ub_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void ub_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void ub_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(usmem(bar)) : "memory");
}
__device__ __forceinline__ void ub_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void ub_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
#define UB_FENCE_BEFORE() asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory")
#define UB_FENCE_AFTER() asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory")

struct UmmaParams {
    const uint8_t *band;        // [UB_BROWS * 32] bytes in the blocked K-major order of the shared copy
    double *bg;                 // [S][tiles][32 warps][8][32 lanes] double2
    const uint32_t *maskbits;   // [S][h][wpr]
    uint32_t *tbits;            // [S][T][flatwords] raw threshold bits
    size_t flatwords;
    const StreamState *state;
    const int *nvalid;
    int *rawrange;              // [S][T][2]
    uint8_t *blur_out;          // [S][T][h][w] parity tap or null
    int T, w, h, wpr, tilesX, tilesY, threshold;
    double alpha, beta;
};

template <bool SAFE>
__global__ void __launch_bounds__(UB_THREADS, 1) k_umma_blur(const __grid_constant__ CUtensorMap tmap, UmmaParams p) {
    extern __shared__ unsigned char ub_raw[];
    unsigned char *sm = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(ub_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *sGray = sm;                                     // [2 stages][2 panels][224][128], written by TMA
    unsigned char *sLo = sm + 2 * UB_GSTAGE;                       // [16 column groups][14 row chunks][8][16]
    unsigned char *sHi = sLo + UB_PLANE;
    unsigned char *sBand = sHi + UB_PLANE;                         // [44 row groups][2][8][16]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sBand + UB_BROWS * 32);      // tma[2], mma1, mma2
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = blockIdx.y, tile = blockIdx.x;
    const int ty = tile / p.tilesX, tx = tile - ty * p.tilesX;
    const int X0 = tx * UB_T, Y0 = ty * UB_T;
    const int Ts = min(p.T, __ldg(p.nvalid + s));
    if (Ts <= 0) return;
    const bool has_bg = p.state[s].has_bg != 0;

    for (int i = tid; i < UB_BROWS * 32 / 16; i += UB_THREADS)
        reinterpret_cast<uint4 *>(sBand)[i] = __ldg(reinterpret_cast<const uint4 *>(p.band) + i);
    if (tid == 0) {
        ub_mbar_init(&bars[0], 1);
        ub_mbar_init(&bars[1], 1);
        ub_mbar_init(&bars[2], 1);
        ub_mbar_init(&bars[3], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(usmem(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the band (generic-proxy writes) is read by the MMAs
    UB_FENCE_BEFORE();
    __syncthreads();
    UB_FENCE_AFTER();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tD1 = tmem, tLo = tmem + 256, tHi = tmem + 384;

    // this thread's pixels: column X0 + 32 (warp & 3) + lane, rows Y0 + 16 (warp >> 2) .. + 15
    const int lq = warp & 3, rg = warp >> 2;
    const int px = X0 + 32 * lq + lane, py = Y0 + 16 * rg;
    const bool okx = px < p.w;
    double2 *bgt = reinterpret_cast<double2 *>(p.bg) + ((((size_t)s * p.tilesX * p.tilesY + tile) * 32 + warp) * 8) * 32 + lane;
    double bg[16];
    if (has_bg) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const double2 v = bgt[i * 32];
            bg[2 * i] = v.x;
            bg[2 * i + 1] = v.y;
        }
    }
    uint32_t M = 0;                   // polygon mask bits of the 16 pixels (bit i = row py + i)
    if (okx) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int y = py + i;
            if (y < p.h) M |= ((__ldg(p.maskbits + ((size_t)s * p.h + y) * p.wpr + (px >> 5)) >> (px & 31)) & 1u) << i;
        }
    }
    uint32_t valid = 0;               // pixels inside the image
    if (okx) valid = py + 16 <= p.h ? 0xFFFFu : (py < p.h ? (1u << (p.h - py)) - 1u : 0u);

    const uint32_t idesc1 = (2u << 4) | ((uint32_t)(UB_IN >> 3) << 17) | ((128u >> 4) << 24);      // S32 += U8 x U8, M128, N224
    const uint32_t idesc2 = (2u << 4) | ((uint32_t)(UB_T >> 3) << 17) | ((128u >> 4) << 24);       // N128
    const uint32_t aBand = usmem(sBand), aGray = usmem(sGray), aLo = usmem(sLo), aHi = usmem(sHi);
    auto issue_tma = [&](int t) {                 // thread 0
        uint64_t *bar = &bars[t & 1];
        unsigned char *dst = sGray + (t & 1) * UB_GSTAGE;
        ub_mbar_expect(bar, UB_GSTAGE);
        ub_tma_4d(dst, &tmap, bar, X0, Y0, t, s);
        ub_tma_4d(dst + UB_IN * 128, &tmap, bar, X0 + 128, Y0, t, s);
    };
    auto issue_mma1 = [&](int t) {                // thread 0: horizontal pass of frame t into D1
        ub_mbar_wait(&bars[t & 1], (t >> 1) & 1);
        UB_FENCE_AFTER();
        const uint32_t g = aGray + (t & 1) * UB_GSTAGE;
#pragma unroll
        for (int j = 0; j < UB_IN / 32; j++) {
            const uint64_t ad = ub_desc_plain(aBand + (UB_BOFF - 32 * j) * 32, 128, 256);
            const uint64_t bd = ub_desc_sw128(g + (j >> 2) * (UB_IN * 128) + (j & 3) * 32);
            ub_mma(tD1, ad, bd, idesc1, j > 0);
        }
        ub_commit(&bars[2]);
    };
    if (tid == 0) {
        issue_tma(0);
        issue_mma1(0);
    }
    const int qoff = 0x4B400000 - p.threshold;
    const uint32_t thr2 = 2u * (uint32_t)p.threshold;
    const double nC = -(4503599627370496.0 * p.alpha);
    uint32_t *tw = p.tbits + (size_t)s * p.T * p.flatwords;
    const int wcol = (X0 >> 5) + lq;

    for (int t = 0; t < Ts; t++) {
        ub_mbar_wait(&bars[2], t & 1);            // D1 of frame t is complete (and gray stage t & 1 has been read)
        UB_FENCE_AFTER();
        if (tid == 0 && t + 1 < Ts) issue_tma(t + 1);
        // ---- split: 16-bit horizontal sums -> low / high byte planes, K-major for the vertical pass ----
        for (int yh = rg; yh < 2 * UB_YQ; yh += 8) {          // half chunks of 8 rows (register budget: 64 per thread)
            uint32_t r[8];
            ub_ld8(tD1 + ((uint32_t)(32 * lq) << 16) + 8 * yh, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            uint2 lo, hi;
            uint32_t a, b;
            a = __byte_perm(r[0], r[1], 0x5410); b = __byte_perm(r[2], r[3], 0x5410);
            lo.x = __byte_perm(a, b, 0x6420); hi.x = __byte_perm(a, b, 0x7531);
            a = __byte_perm(r[4], r[5], 0x5410); b = __byte_perm(r[6], r[7], 0x5410);
            lo.y = __byte_perm(a, b, 0x6420); hi.y = __byte_perm(a, b, 0x7531);
            const int xl = 32 * lq + lane;
            const int off = ((xl >> 3) * UB_YQ + (yh >> 1)) * 128 + (xl & 7) * 16 + 8 * (yh & 1);
            *reinterpret_cast<uint2 *>(sLo + off) = lo;
            *reinterpret_cast<uint2 *>(sHi + off) = hi;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        UB_FENCE_BEFORE();
        __syncthreads();
        if (tid == 0) {
            UB_FENCE_AFTER();
#pragma unroll
            for (int j = 0; j < UB_IN / 32; j++) {
                const uint64_t bd = ub_desc_plain(aBand + (UB_BOFF - 32 * j) * 32, 128, 256);
                ub_mma(tLo, ub_desc_plain(aLo + 2 * j * 128, 128, UB_YQ * 128), bd, idesc2, j > 0);
                ub_mma(tHi, ub_desc_plain(aHi + 2 * j * 128, 128, UB_YQ * 128), bd, idesc2, j > 0);
            }
            ub_commit(&bars[3]);
            if (t + 1 < Ts) issue_mma1(t + 1);    // runs on the tensor pipe under the epilogue below
        }
        ub_mbar_wait(&bars[3], t & 1);
        UB_FENCE_AFTER();
        // ---- epilogue: blur -> mask -> threshold bit -> background update, 16 rows of one column per thread ----
        const bool init = t == 0 && !has_bg;
        uint32_t anyw = 0, myword = 0;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            uint32_t lo[8], hi[8];
            ub_ld8(tLo + ((uint32_t)(32 * lq) << 16) + 16 * rg + 8 * half, lo);
            ub_ld8(tHi + ((uint32_t)(32 * lq) << 16) + 16 * rg + 8 * half, hi);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int i = 8 * half + j;
                uint32_t sv = ((hi[j] * 256u + lo[j] + 32768u) >> 16) & 0xFFu;
                if (M & (1u << i)) sv = 0;                                       // mask_off_areas paints BLACK into blur
                if (p.blur_out && (valid & (1u << i)))
                    p.blur_out[(((size_t)s * p.T + t) * p.h + py + i) * p.w + px] = (uint8_t)sv;
                bool bit;
                if (SAFE) {
                    const double X = __hiloint2double(0x43300000, (int)sv);      // 2^52 + blur
                    if (init) bg[i] = X - 4503599627370496.0;                    // ref_frame = blur.astype(float)
                    const int q = __float_as_int(__fadd_rn(__double2float_rn(bg[i]), 12582912.0f));
                    bit = (uint32_t)(q - qoff - (int)sv) > thr2;                 // |bg8 - blur| > threshold
                    bg[i] = __fma_rn(bg[i], p.beta, __fma_rn(X, p.alpha, nC));   // fma(bg, 1 - a, rn(blur * a))
                } else {
                    const double sd = __hiloint2double(0x43300000, (int)sv) - 4503599627370496.0;
                    if (init) bg[i] = sd;
                    const float f = fminf(fabsf(__double2float_rn(bg[i])), 255.0f);
                    const int b8 = __float_as_int(__fadd_rn(f, 12582912.0f)) & 0x1FF;
                    const int d = (int)sv - b8;
                    bit = (d < 0 ? -d : d) > p.threshold;
                    bg[i] = __fma_rn(bg[i], p.beta, __dmul_rn(sd, p.alpha));
                }
                const uint32_t word = __ballot_sync(0xffffffffu, bit && (valid & (1u << i)));    // 32 pixels of row py + i
                anyw |= word;
                if (lane == i) myword = word;
            }
        }
        if (lane < 16 && py + lane < p.h && wcol < p.wpr) tw[(size_t)(py + lane) * p.wpr + wcol] = myword;
        if (anyw && lane == 0) {                          // this warp's 16 rows hold something
            int *rr = p.rawrange + 2 * ((size_t)s * p.T + t);
            atomicMax(rr, min(py + 15, p.h - 1));
            atomicMax(rr + 1, p.h - 1 - py);
        }
        tw += p.flatwords;
        UB_FENCE_BEFORE();
        __syncthreads();          // D2 and the byte planes are free for frame t + 1
        UB_FENCE_AFTER();
    }
#pragma unroll
    for (int i = 0; i < 8; i++) bgt[i * 32] = make_double2(bg[2 * i], bg[2 * i + 1]);
    UB_FENCE_BEFORE();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// ---------------------------------------------------------------------------------------------
// padded gray plane: gpad[f][yp][xp] = gray(reflect101(yp - 48), reflect101(xp - 48)); BGR: convert on the way
// (find_motion.py:493).  One thread = 4 padded pixels.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ub_gray1(const uint8_t *q) { return (3735u * q[0] + 19235u * q[1] + 9798u * q[2] + 16384u) >> 15; }
__device__ __forceinline__ uint32_t ub_gray4(uint32_t w0, uint32_t w1, uint32_t w2) {
    const uint32_t C_BG = 7470u | (38470u << 16), C_R = 19596u;
    const uint32_t C_xB = 7470u << 16, C_GR = 38470u | (19596u << 16);
    uint32_t t0 = __dp2a_hi(C_R, w0, __dp2a_lo(C_BG, w0, 32768u));
    uint32_t t1 = __dp2a_hi(C_xB, w0, __dp2a_lo(C_GR, w1, 32768u));
    uint32_t t2 = __dp2a_hi(C_BG, w1, __dp2a_lo(C_R, w2, 32768u));
    uint32_t t3 = __dp2a_hi(C_GR, w2, __dp2a_lo(C_xB, w2, 32768u));
    return __byte_perm(__byte_perm(t0, t1, 0x0062), __byte_perm(t2, t3, 0x0062), 0x5410);
}

template <bool BGR>
__global__ void __launch_bounds__(256) k_pad_gray(const uint8_t *__restrict__ src, size_t sstride, size_t fstride, int T,
                                                  uint8_t *__restrict__ gpad, int w, int h, int Wp, int Hp,
                                                  const int *__restrict__ nvalid, int aligned4) {
    const int f = blockIdx.z, s = f / T, t = f - s * T;
    if (t >= __ldg(nvalid + s)) return;
    const int yp = blockIdx.y, xp = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (xp >= Wp) return;
    uint32_t out = 0;
    if (yp < h + 2 * UB_PAD && xp < w + 2 * UB_PAD) {
        const int y = fm_reflect101(yp - UB_PAD, h);
        const uint8_t *row = BGR ? src + (size_t)s * sstride + (size_t)t * fstride + (size_t)y * w * 3
                                 : src + ((size_t)f * h + y) * w;
        const int x = xp - UB_PAD;
        if (x >= 0 && x + 3 < w && aligned4) {
            if (BGR) {
                const uint32_t *q = reinterpret_cast<const uint32_t *>(row + 3 * x);
                out = ub_gray4(__ldg(q), __ldg(q + 1), __ldg(q + 2));
            } else {
                out = __ldg(reinterpret_cast<const uint32_t *>(row + x));
            }
        } else {
#pragma unroll
            for (int b = 0; b < 4; b++) {
                if (xp + b < w + 2 * UB_PAD) {
                    const int xx = fm_reflect101(x + b, w);
                    out |= (BGR ? ub_gray1(row + 3 * xx) : (uint32_t)row[xx]) << (8 * b);
                }
            }
        }
    }
    *reinterpret_cast<uint32_t *>(gpad + ((size_t)f * Hp + yp) * Wp + xp) = out;
}

// tiled background -> row-major float64 plane
__global__ void k_bg_export_umma(const double *__restrict__ bg, double *__restrict__ dst, int w, int h, int tilesX, int tilesY, int s) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const int tx = x / UB_T, ty = y / UB_T, lx = x % UB_T, ly = y % UB_T;
    const int warp = (lx >> 5) + 4 * (ly >> 4), lane = lx & 31, i = ly & 15;
    const size_t base = ((((size_t)s * tilesX * tilesY + (size_t)ty * tilesX + tx) * 32 + warp) * 8) * 32;
    dst[(size_t)y * w + x] = bg[(base + (size_t)(i >> 1) * 32 + lane) * 2 + (i & 1)];
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// bit planes are written a 32-pixel word per warp ballot: needs the flat bit order to equal the row-padded one (w % 32 == 0)
bool fm_umma_supported(const fm_ctx *c) { return c->k <= 2 * UB_PAD + 1 && (c->w % 32) == 0; }

static void umma_geom(const fm_ctx *c, int *tilesX, int *tilesY, int *Wp, int *Hp) {
    *tilesX = (c->w + UB_T - 1) / UB_T;
    *tilesY = (c->h + UB_T - 1) / UB_T;
    *Wp = *tilesX * UB_T + 128;             // the second panel of the last tile stays inside the plane
    *Hp = *tilesY * UB_T + 2 * UB_PAD;
}

size_t fm_umma_pad_bytes(const fm_ctx *c) {
    int tx, ty, Wp, Hp;
    umma_geom(c, &tx, &ty, &Wp, &Hp);
    return (size_t)c->S * c->Tmax * Wp * Hp;
}

size_t fm_umma_bg_doubles(const fm_ctx *c) {
    int tx, ty, Wp, Hp;
    umma_geom(c, &tx, &ty, &Wp, &Hp);
    return (size_t)c->S * tx * ty * UB_T * UB_T;
}

// band[u][q] = c48[q - u], u in [-224, 128), q in [0, 32), c48 = the taps centred in a 97-tap window; stored in the
// blocked K-major order of the shared copy: [row group of 8][16-byte half][8 rows][16 bytes]
int fm_umma_init(fm_ctx *c, const int *taps) {
    std::vector<uint8_t> band((size_t)UB_BROWS * 32, 0);
    const int shift = UB_PAD - (c->k >> 1);
    for (int ur = 0; ur < UB_BROWS; ur++)
        for (int q = 0; q < 32; q++) {
            const int i = q - (ur - UB_BOFF) - shift;
            if (i >= 0 && i < c->k) band[((ur >> 3) * 2 + (q >> 4)) * 128 + (ur & 7) * 16 + (q & 15)] = (uint8_t)taps[i];
        }
    FM_CUDA(cudaMalloc((void **)&c->uband, band.size()));
    FM_CUDA(cudaMemcpy(c->uband, band.data(), band.size(), cudaMemcpyHostToDevice));
    FM_CUDA(cudaMalloc((void **)&c->gpad, fm_umma_pad_bytes(c)));
    int rc;
    if ((rc = fm_ensure_smem((const void *)k_umma_blur<true>, UB_SMEM, c->cfg.device))) return rc;
    if ((rc = fm_ensure_smem((const void *)k_umma_blur<false>, UB_SMEM, c->cfg.device))) return rc;
    if (!fm_tma_encoder()) { fm_set_error("cuTensorMapEncodeTiled not available"); return FM_ECUDA; }
    return FM_OK;
}

int fm_launch_gray_plane(fm_ctx *c, const uint8_t *frames, size_t sstride, size_t fstride, int T, cudaStream_t st);   // k_frontend.cu

// frames != nullptr: full-resolution mode (BGR -> padded gray in one kernel); else the resized gray plane c->gray
int fm_launch_umma_blur(fm_ctx *c, const uint8_t *frames, size_t sstride, size_t fstride, int T, cudaStream_t st) {
    int tilesX, tilesY, Wp, Hp;
    umma_geom(c, &tilesX, &tilesY, &Wp, &Hp);
    const int F = c->S * T;
    dim3 pgrid((Wp / 4 + 255) / 256, Hp, F);
    if (frames) {
        if (c->cfg.flags & FM_FLAG_KEEP_PLANES) {          // parity tap of the gray conversion
            int rc = fm_launch_gray_plane(c, frames, sstride, fstride, T, st);
            if (rc) return rc;
        }
        const int al = ((((uintptr_t)frames) | sstride | fstride | ((size_t)c->w * 3)) & 3) == 0;
        k_pad_gray<true><<<pgrid, 256, 0, st>>>(frames, sstride, fstride, T, c->gpad, c->w, c->h, Wp, Hp, c->nvalid, al);
    } else {
        const int al = (c->w & 3) == 0;
        k_pad_gray<false><<<pgrid, 256, 0, st>>>(c->gray, 0, 0, T, c->gpad, c->w, c->h, Wp, Hp, c->nvalid, al);
    }
    FM_LAUNCH_CHECK();
    CUtensorMap tmap;
    cuuint64_t dims[4] = {(cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)T, (cuuint64_t)c->S};
    cuuint64_t strides[3] = {(cuuint64_t)Wp, (cuuint64_t)Wp * Hp, (cuuint64_t)Wp * Hp * T};
    cuuint32_t box[4] = {128, UB_IN, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fm_tma_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void *)c->gpad, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fm_set_error("cuTensorMapEncodeTiled (padded gray plane) failed (%d)", (int)r); return FM_ECUDA; }
    UmmaParams p;
    p.band = c->uband; p.bg = c->bg; p.maskbits = c->maskbits; p.tbits = c->tflat;
    p.flatwords = (size_t)c->ntiles * FM_TILE_WORDS;
    p.state = c->state; p.nvalid = c->nvalid; p.rawrange = c->rawrange;
    p.blur_out = (c->cfg.flags & FM_FLAG_KEEP_PLANES) ? c->blur : nullptr;
    p.T = T; p.w = c->w; p.h = c->h; p.wpr = c->wpr; p.tilesX = tilesX; p.tilesY = tilesY; p.threshold = c->cfg.threshold;
    p.alpha = c->cfg.avg; p.beta = 1.0 - p.alpha;
    const bool safe = p.alpha >= 0.0 && p.alpha <= 1.0 && p.threshold >= 0;
    dim3 grid(tilesX * tilesY, c->S);
    if (safe) k_umma_blur<true><<<grid, UB_THREADS, UB_SMEM, st>>>(tmap, p);
    else k_umma_blur<false><<<grid, UB_THREADS, UB_SMEM, st>>>(tmap, p);
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int fm_launch_bg_export_umma(fm_ctx *c, int stream, double *dst_dev, cudaStream_t st) {
    int tilesX, tilesY, Wp, Hp;
    umma_geom(c, &tilesX, &tilesY, &Wp, &Hp);
    dim3 grid((c->w + 127) / 128, c->h);
    k_bg_export_umma<<<grid, 128, 0, st>>>(c->bg, dst_dev, c->w, c->h, tilesX, tilesY, stream);
    FM_LAUNCH_CHECK();
    return FM_OK;
}
