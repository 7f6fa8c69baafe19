// K1: fused, time-blocked stencil + background kernel (full-resolution mode, Gaussian k <= 5).
//
// One CTA owns a 128 x FT_H pixel tile of one stream and walks the T frames of the call in order:
//   BGR tile + 2 px halo --(TMA, double buffered)--> gray bytes in shared memory
//   --> horizontal 8.8 fixed-point Gaussian pass as a banded (Toeplitz) u8 x u8 matrix product on the
//       tensor cores (IMMA.16832.U8: A = 16 gray rows x 32 columns via LDSM, B = the taps), sums packed
//       2x16-bit into shared memory
//   --> vertical pass on the packed lanes in registers (sliding 5-row window)
//   --> polygon mask --> bg8 = rne(f32(bg)), |blur - bg8| > threshold --> bit-packed mask out
//   --> bg = fma(bg, 1-alpha, rn(blur*alpha))
// Each thread keeps the float64 background of its 16 pixels (one row, two 8-pixel segments) in registers across all T frames,
// so the background costs one 16 B/px HBM round trip per call instead of per frame.
// Replaces blur_frame + mask_off_areas + find_diff's diff/threshold/accumulateWeighted
// (find_motion/find_motion.py:487-494, 619-635, 246-257, 651-659).
//
// Packed arithmetic: for k in {1,3,5} the taps are g*[b0,b1,b2,b1,b0] with g >= 16, so the
// horizontal sums (<= 255*256/g) and the vertical sums (<= 255*(256/g)^2 <= 65280) fit 16 bits
// and two pixels share one 32-bit IMAD; (ver + 32768) >> 16 == (v' + round) >> shift exactly.
#include <cuda.h>

#include "fm_common.cuh"

#define FT_W 128
#ifndef FT_H
#define FT_H 32            // tile height; a thread owns one row x 16 pixels, a warp 4 rows x 128 columns
#endif
#define FT_PX 16           // pixels (and float64 background values) per thread
#define FUSED_WARPS (FT_H / 4)
#define FG_WORDS 36        // gray words per shared row: cols x0-4 .. x0+139 (18 units of 8 pixels)
#define FG_ROWS (FT_H + 4) // rows y0-2 .. y0+FT_H+1
#define FH_MB ((FG_ROWS + 15) / 16)     // 16-row blocks of the tensor-core pass
#define FH_ROWS (16 * FH_MB)            // rows of the shared planes (the rows past FG_ROWS are scratch: no row guards)
#ifndef FUSED_MIN_CTAS
#define FUSED_MIN_CTAS (1024 / (8 * FT_H))      // 64 registers per thread: 1024 resident threads per SM
#endif
#define FH_WORDS 68         // packed horizontal sums per shared row: 64 pixel pairs + 4 (bank spread for the D-fragment stores)
#define FUSED_THREADS (8 * FT_H)
#define RAW_PITCH 432      // bytes per staged BGR row: bytes x0*3-16 .. x0*3+415 (TMA box of 108 u32)
#define RAW_BYTES (RAW_PITCH * FG_ROWS)                 // bytes one TMA box delivers
#define RAW_STAGE ((RAW_BYTES + 127) / 128 * 128)        // stage stride (TMA destinations are 128 B aligned)

// ---- TMA / mbarrier primitives (sm_100a PTX) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

struct FusedParams {
    int T, w, h, wpr;           // T = frames per stream of the call
    const int *nvalid;          // [S] real frames of each stream in this call (ragged batches), <= T
    int tilesX, tilesY;
    double *bg;                 // [S][tiles][8 warps][8 pairs][32 lanes] double2 (thread-private, coalesced)
    const uint32_t *maskbits;   // [S][h][wpr]
    uint32_t *tbits;            // [S][T][flatwords]  (row-padded == flat because w % 32 == 0)
    size_t flatwords;
    const StreamState *state;
    int b0, b1, b2, shift, rnd; // taps / g, and the rounding of the final shift (packed in both lanes)
    int threshold;
    double alpha, beta;
    uint8_t *gray_out, *blur_out;   // [S][T][h][w], only with KEEP
    int *rawrange;                  // [S][T][2] (max y, max h-1-y) of rows with pixels above threshold
    uint4 bfr[32];                  // per lane: B fragments of the horizontal pass (fused_bfr)
};

// B fragments of the horizontal pass (constant per call): the banded tap matrix for the two 8-column output blocks of a
// 32-byte window.  Output column n of block nb is gray byte 4 + 16 j + 8 nb + n of its row, window byte k is gray byte
// 16 j + k, so the tap index is k - n - 8 nb - 2 (taps b0 b1 b2 b1 b0).  Lane = 4 g + tq holds bytes k = 16 r + 4 tq + i of
// column n = g: words (nb, r) = (0,0) (0,1) (1,0) (1,1).
static void fused_bfr(FusedParams &p) {
    for (int lane = 0; lane < 32; lane++) {
        const int g = lane >> 2, tq = lane & 3;
        uint32_t v[4];
        for (int nb = 0; nb < 2; nb++)
            for (int r = 0; r < 2; r++) {
                uint32_t x = 0;
                for (int i = 0; i < 4; i++) {
                    const int idx = 16 * r + 4 * tq + i - g - 8 * nb - 2;
                    const int tap = (idx == 2) ? p.b2 : (idx == 1 || idx == 3) ? p.b1 : (idx == 0 || idx == 4) ? p.b0 : 0;
                    x |= (uint32_t)tap << (8 * i);
                }
                v[2 * nb + r] = x;
            }
        p.bfr[lane] = make_uint4(v[0], v[1], v[2], v[3]);
    }
}

__device__ __forceinline__ uint32_t gray4(uint32_t w0, uint32_t w1, uint32_t w2) {
    // 4 BGR pixels in 3 words -> 4 gray bytes.  Y = (3735 B + 19235 G + 9798 R + 16384) >> 15 computed as
    // (7470 B + 38470 G + 19596 R + 32768) >> 16 with two 16-bit x 8-bit dot products per pixel (IDP.2A);
    // the lo/hi byte-pair selection of IDP.2A picks each pixel's bytes straight out of the three words.
    // The result is byte 2 of the accumulator.
    const uint32_t C_BG = 7470u | (38470u << 16), C_R = 19596u;              // byte pairs [B,G], [R,x]
    const uint32_t C_xB = 7470u << 16, C_GR = 38470u | (19596u << 16);       // byte pairs [x,B], [G,R]
    uint32_t t0 = __dp2a_hi(C_R, w0, __dp2a_lo(C_BG, w0, 32768u));           // w0 = [B0,G0,R0,B1]
    uint32_t t1 = __dp2a_hi(C_xB, w0, __dp2a_lo(C_GR, w1, 32768u));          // w1 = [G1,R1,B2,G2]
    uint32_t t2 = __dp2a_hi(C_BG, w1, __dp2a_lo(C_R, w2, 32768u));           // w2 = [R2,B3,G3,R3]
    uint32_t t3 = __dp2a_hi(C_GR, w2, __dp2a_lo(C_xB, w2, 32768u));
    return __byte_perm(__byte_perm(t0, t1, 0x0062), __byte_perm(t2, t3, 0x0062), 0x5410);
}

// ---- tensor-core primitives for the horizontal pass ----
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void imma_u8(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                        uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(0));
}

__device__ __forceinline__ double u8_to_f64(uint32_t v) {
    return __int2double_rn((int)v);          // exact for 0..255; one XU op instead of MOV + DADD
}

template <bool SAFE>
__device__ __forceinline__ int bg8_magic(double b) {
    // 0x4B400000 + rne(|float32(b)|) saturated to 255
    float f = __double2float_rn(b);
    if (!SAFE) f = fminf(fabsf(f), 255.0f);
    return __float_as_int(__fadd_rn(f, 12582912.0f));
}

// bits = 2 * bits + (d > thr2) in two instructions: d + ~thr2 carries out exactly when d > thr2 (unsigned), and the
// carry is shifted into the accumulator by an add-with-carry.
__device__ __forceinline__ void push_gt(uint32_t &bits, uint32_t d, uint32_t nthr2) {
    asm("{\n .reg .u32 t;\n add.cc.u32 t, %1, %2;\n addc.u32 %0, %0, %0;\n}" : "+r"(bits) : "r"(d), "r"(nthr2));
}

// 8x8 transpose of 4-bit elements across the 8 lanes of a lane octet: in: lane i holds e[r] (nibble r)
// = bits of row r; out: lane r holds nibble i = bits of lane i  -> a 32-pixel row word.
__device__ __forceinline__ uint32_t nibble_transpose8(uint32_t x, int lane) {
    uint32_t o = __shfl_xor_sync(0xffffffffu, x, 4);
    x = (lane & 4) ? ((o >> 16) | (x & 0xFFFF0000u)) : ((x & 0x0000FFFFu) | (o << 16));
    o = __shfl_xor_sync(0xffffffffu, x, 2);
    x = (lane & 2) ? (((o >> 8) & 0x00FF00FFu) | (x & 0xFF00FF00u)) : ((x & 0x00FF00FFu) | ((o & 0x00FF00FFu) << 8));
    o = __shfl_xor_sync(0xffffffffu, x, 1);
    x = (lane & 1) ? (((o >> 4) & 0x0F0F0F0Fu) | (x & 0xF0F0F0F0u)) : ((x & 0x0F0F0F0Fu) | ((o & 0x0F0F0F0Fu) << 4));
    return x;
}

// One thread: 5 rows of packed horizontal sums in -> ONE output row x 16 pixels (two segments of 8 consecutive pixels, 64
// pixels apart, so that the 128-bit shared loads of a quarter-warp are contiguous): vertical pass, mask, threshold bits,
// background update.  The 16 threshold bits of a thread are two bytes of the bit plane: no transposition across lanes.
// INIT: first frame of a stream (ref_frame = blur.astype(float));  MASKED: the warp has masked pixels (M bit j = pixel j);
// SH8: k = 5 (taps 1,4,6,4,1 compiled in, final shift 8 done by byte selection).
template <bool KEEP, bool SAFE, bool INIT, bool MASKED, bool SH8>
__device__ __forceinline__ uint32_t fused_rows(const uint32_t *shw, double (&bg)[FT_PX], uint32_t M, const FusedParams &p,
                                               uint8_t *blur_out, uint32_t vmask) {
    const uint32_t b0 = p.b0, b1 = p.b1, b2 = p.b2;
    const int qoff = 0x4B400000 - p.threshold;
    const unsigned nthr2 = ~(2u * (unsigned)p.threshold);
    const double nC = -(4503599627370496.0 * p.alpha);
    uint32_t bits = 0;
#pragma unroll
    for (int seg = 0; seg < 2; seg++) {
        // rows y-2 .. y+2 of the segment, pixels (0,1) (2,3) (4,5) (6,7) per 128-bit load, 16 bits each; accumulated as they
        // arrive (outer rows, then the +-1 rows, then the centre) to keep the live registers low
        const uint32_t *sp = shw + 32 * seg;
        uint32_t v[4];             // the two pixels of a pair in bits [0,8) and [16,24) (SH8: in bytes 1 and 3)
        {
            const uint4 r0 = *reinterpret_cast<const uint4 *>(sp), r4 = *reinterpret_cast<const uint4 *>(sp + 4 * FH_WORDS);
            const uint32_t rn = SH8 ? 0x00800080u : (uint32_t)p.rnd;
            if (SH8) { v[0] = r0.x + r4.x + rn; v[1] = r0.y + r4.y + rn; v[2] = r0.z + r4.z + rn; v[3] = r0.w + r4.w + rn; }
            else { v[0] = b0 * (r0.x + r4.x) + rn; v[1] = b0 * (r0.y + r4.y) + rn; v[2] = b0 * (r0.z + r4.z) + rn; v[3] = b0 * (r0.w + r4.w) + rn; }
        }
        {
            const uint4 r1 = *reinterpret_cast<const uint4 *>(sp + FH_WORDS), r3 = *reinterpret_cast<const uint4 *>(sp + 3 * FH_WORDS);
            const uint32_t m1 = SH8 ? 4u : (uint32_t)b1;
            v[0] += m1 * (r1.x + r3.x); v[1] += m1 * (r1.y + r3.y); v[2] += m1 * (r1.z + r3.z); v[3] += m1 * (r1.w + r3.w);
        }
        {
            const uint4 r2 = *reinterpret_cast<const uint4 *>(sp + 2 * FH_WORDS);
            const uint32_t m2 = SH8 ? 6u : (uint32_t)b2;
            v[0] += m2 * r2.x; v[1] += m2 * r2.y; v[2] += m2 * r2.z; v[3] += m2 * r2.w;
        }
        if (!SH8) {
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = (v[j] >> p.shift) & 0x00FF00FFu;
        }
        if (KEEP && (vmask & (1u << seg))) {
            uint2 o;
            if (SH8) { o.x = __byte_perm(v[0], v[1], 0x7531); o.y = __byte_perm(v[2], v[3], 0x7531); }
            else { o.x = __byte_perm(v[0], v[1], 0x6420); o.y = __byte_perm(v[2], v[3], 0x6420); }
            const uint32_t mk = (M >> (8 * seg)) & 0xFFu;
            o.x &= ~((((mk & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu);
            o.y &= ~(((((mk >> 4) & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu);
            *reinterpret_cast<uint2 *>(blur_out + 64 * seg) = o;
        }
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int idx = 8 * seg + c;
            uint32_t sv = SH8 ? __byte_perm(v[c >> 1], 0, (c & 1) ? 0x4443 : 0x4441)
                              : ((c & 1) ? (v[c >> 1] >> 16) : (v[c >> 1] & 0xFFFFu));
            if (MASKED && (M & (1u << idx))) sv = 0;          // mask_off_areas paints BLACK into blur
            if (SAFE) {
                // rn(blur * alpha) without an int -> double conversion (the XU pipe converts 16 lanes/clk/SM):
                // X = 2^52 + blur is assembled from its bit pattern, and X * alpha - 2^52 * alpha is exactly
                // blur * alpha before the single rounding of the FMA (2^52 * alpha is exact).
                const double X = __hiloint2double(0x43300000, (int)sv);
                if (INIT) bg[idx] = X - 4503599627370496.0;   // ref_frame = blur.astype(float)
                int q = bg8_magic<SAFE>(bg[idx]);
                push_gt(bits, (unsigned)(q - qoff - (int)sv), nthr2);           // |bg8 - blur| > threshold
                bg[idx] = __fma_rn(bg[idx], p.beta, __fma_rn(X, p.alpha, nC));
            } else {
                const double sd = u8_to_f64(sv);
                if (INIT) bg[idx] = sd;
                int q = bg8_magic<SAFE>(bg[idx]);
                push_gt(bits, (unsigned)(q - qoff - (int)sv), nthr2);
                bg[idx] = __fma_rn(bg[idx], p.beta, __dmul_rn(sd, p.alpha));
            }
        }
    }
    return __brev(bits) >> (32 - FT_PX);      // the first pixel was pushed first: bit idx = pixel idx
}

template <bool KEEP, bool SAFE>
__global__ void __launch_bounds__(FUSED_THREADS, FUSED_MIN_CTAS) k_fused(const __grid_constant__ CUtensorMap tmap, FusedParams p) {
    extern __shared__ __align__(128) unsigned char fsm[];
    unsigned char *raw = fsm;                                                  // [2][RAW_STAGE] staged BGR rows (TMA)
    uint32_t *sg = reinterpret_cast<uint32_t *>(fsm + 2 * RAW_STAGE);        // [FH_ROWS][FG_WORDS] gray bytes (FG_ROWS used)
    uint32_t *sh = sg + FH_ROWS * FG_WORDS;                                  // [FH_ROWS][FH_WORDS] packed horizontal sums
    uint64_t *bars = reinterpret_cast<uint64_t *>(sh + FH_ROWS * FH_WORDS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = blockIdx.z;
    const int tx = blockIdx.x, ty = blockIdx.y, tile = ty * p.tilesX + tx;
    const int x0 = tx * FT_W, y0 = ty * FT_H;
    const int w = p.w, h = p.h;
    const bool border = (x0 == 0) || (x0 + FT_W >= w) || (y0 == 0) || (y0 + FT_H >= h);
    const bool has_bg = p.state[s].has_bg != 0;
    const int Ts = min(p.T, __ldg(p.nvalid + s));            // frames of this stream that are real
    if (Ts <= 0) return;

    // this thread's pixels: row y0 + tid / 8, columns x0 + 8 (tid % 8) .. + 7 and the same 64 pixels further right
    const int trow = tid >> 3, xg = tid & 7;
    const int px = x0 + 8 * xg, py = y0 + trow;
    double2 *bgt = reinterpret_cast<double2 *>(p.bg) +
                   ((((size_t)s * p.tilesX * p.tilesY + tile) * FUSED_WARPS + warp) * (FT_PX / 2)) * 32 + lane;
    double bg[FT_PX];
    if (has_bg) {
#pragma unroll
        for (int i = 0; i < FT_PX / 2; i++) {
            double2 v = bgt[i * 32];
            bg[2 * i] = v.x;
            bg[2 * i + 1] = v.y;
        }
    }
    // polygon mask bits of the 16 pixels (bit j = pixel j of the first segment, bit 8 + j of the second), loaded once per call
    uint32_t M = 0;
    const bool okA = py < h && px < w, okB = py < h && px + 64 < w;
    const uint32_t vmask = (okA ? 1u : 0u) | (okB ? 2u : 0u);
    if (okA) M = (__ldg(p.maskbits + ((size_t)s * h + py) * p.wpr + (px >> 5)) >> (px & 31)) & 0xFFu;
    if (okB) M |= ((__ldg(p.maskbits + ((size_t)s * h + py) * p.wpr + ((px + 64) >> 5)) >> (px & 31)) & 0xFFu) << 8;
    const bool wmasked = __any_sync(0xffffffffu, M != 0);       // warp-uniform choice of the masked variant
    // TMA pipeline: frame t+1 lands in the other stage while frame t is being processed
    const int cx = (x0 * 3) / 4 - 4, cy = y0 - 2;        // box origin in (u32 column, row); OOB is zero-filled
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bars[0], RAW_BYTES);
        tma_load_4d(raw, &tmap, &bars[0], cx, cy, 0, s);
    }
    // the thread's two bytes of the bit plane (row py, pixels px .. px+7 and px+64 .. px+71)
    uint8_t *tb = reinterpret_cast<uint8_t *>(p.tbits + ((size_t)s * p.T) * p.flatwords + (size_t)py * p.wpr) + (px >> 3);
    // B fragments of the horizontal pass (constant, prepared by the host: fused_bfr)
    uint32_t bfr[2][2];
    {
        const uint4 b = p.bfr[lane];
        bfr[0][0] = b.x; bfr[0][1] = b.y; bfr[1][0] = b.z; bfr[1][1] = b.w;
        // opaque: otherwise the compiler re-reads the (lane-indexed, hence serialised) constant bank in every frame
        asm volatile("" : "+r"(bfr[0][0]), "+r"(bfr[0][1]), "+r"(bfr[1][0]), "+r"(bfr[1][1]));
    }
    // A-fragment row addresses of LDSM.x4 (lane L supplies row L&7 of matrix L>>3; matrices: rows 0-7 / 8-15 of
    // bytes 0-15, then of bytes 16-31) and D-fragment store positions
    const int hwin = warp & 7;                                  // the warp's 16-column window of the horizontal pass
    const uint32_t sg_lane = smem_u32(sg) + ((lane & 7) + 8 * ((lane >> 3) & 1)) * (FG_WORDS * 4) + 16 * hwin + 16 * (lane >> 4);
    uint32_t *sh_lane = sh + (lane >> 2) * FH_WORDS + 8 * hwin + (lane & 3);

    for (int t = 0; t < Ts; t++) {
        // No barrier here: every thread that gets this far has passed the second barrier of frame t-1, i.e. all
        // conversions out of raw stage (t+1)&1 (frame t-1) and all reads of the gray plane are complete.
        if (tid == 0 && t + 1 < Ts) {
            mbar_expect_tx(&bars[(t + 1) & 1], RAW_BYTES);
            tma_load_4d(raw + ((t + 1) & 1) * RAW_STAGE, &tmap, &bars[(t + 1) & 1], cx, cy, t + 1, s);
        }
        mbar_wait(&bars[t & 1], (t >> 1) & 1);
        // ---- staged BGR -> gray bytes in shared memory (tile + halo) ----
        {
            // the staged rows are dense (pitch 432 B = 36 groups of 4 pixels = 12 B, the first at byte 4 because the
            // TMA box must start 16 B aligned in global memory -- a box at x0 * 3 - 12 raises "illegal instruction") and
            // so are the gray rows (36 words): group u is raw + 4 + 12 u -> sg[u].
            const unsigned char *rs = raw + (t & 1) * RAW_STAGE + 4;
            // two units (8 pixels, 24 bytes -> 2 gray words) per step: half the address arithmetic and loop control
#pragma unroll
            for (int i = 0; i < (FG_ROWS * FG_WORDS / 2 + FUSED_THREADS - 1) / FUSED_THREADS; i++) {
                int u = tid + i * FUSED_THREADS;
                if (u < FG_ROWS * FG_WORDS / 2) {
                    // 24 bytes at byte 4 (mod 8) of the stage: 4 + 8 + 8 + 4
                    const unsigned char *q = rs + u * 24;
                    const uint32_t a0 = *reinterpret_cast<const uint32_t *>(q), a5 = *reinterpret_cast<const uint32_t *>(q + 20);
                    const uint2 a12 = *reinterpret_cast<const uint2 *>(q + 4), a34 = *reinterpret_cast<const uint2 *>(q + 12);
                    const uint32_t g0 = gray4(a0, a12.x, a12.y), g1 = gray4(a34.x, a34.y, a5);
                    *reinterpret_cast<uint2 *>(sg + 2 * u) = make_uint2(g0, g1);
                }
            }
        }
        __syncthreads();
        if (border) {               // BORDER_REFLECT_101 for the 2-pixel ring outside the image
            unsigned char *sb = reinterpret_cast<unsigned char *>(sg);
            for (int ry = tid; ry < FG_ROWS; ry += FUSED_THREADS) {
                unsigned char *row = sb + ry * (FG_WORDS * 4) + 4;       // row[c] = column x0 + c
                if (x0 == 0) { row[-1] = row[1]; row[-2] = row[2]; }
                if (x0 + FT_W >= w) { int e = w - x0; row[e] = row[e - 2]; row[e + 1] = row[e - 3]; }
            }
            __syncthreads();
            for (int i = tid; i < FG_WORDS; i += FUSED_THREADS) {
                if (y0 == 0) { sg[1 * FG_WORDS + i] = sg[3 * FG_WORDS + i]; sg[0 * FG_WORDS + i] = sg[4 * FG_WORDS + i]; }
                if (y0 + FT_H >= h) {
                    int e = h - y0 + 2;       // shared row of image row h
                    sg[e * FG_WORDS + i] = sg[(e - 2) * FG_WORDS + i];
                    sg[(e + 1) * FG_WORDS + i] = sg[(e - 3) * FG_WORDS + i];
                }
            }
            __syncthreads();
        }
        if (KEEP) {
            for (int u = tid; u < FT_H * (FT_W / 4); u += FUSED_THREADS) {
                int ry = u / (FT_W / 4), ux = u - ry * (FT_W / 4);
                int gy = y0 + ry, gx = x0 + 4 * ux;
                if (gy < h && gx < w)
                    *reinterpret_cast<uint32_t *>(p.gray_out + (((size_t)s * p.T + t) * h + gy) * w + gx) =
                        sg[(ry + 2) * FG_WORDS + ux + 1];
            }
        }
        // ---- horizontal pass on the tensor cores: warp = one 16-column window, every (FUSED_WARPS / 8)-th block of 16 rows ----
        {
            constexpr int MBS = FUSED_WARPS / 8;                 // warps per window
            constexpr int NMB = (FH_MB + MBS - 1) / MBS;         // blocks per warp (the last may not exist for every warp)
            const int mb0 = warp >> 3;
            uint32_t a[NMB][4];
#pragma unroll
            for (int i = 0; i < NMB; i++)
                if (MBS == 1 || mb0 + i * MBS < FH_MB) ldsm_x4(sg_lane + (mb0 + i * MBS) * (16 * FG_WORDS * 4), a[i][0], a[i][1], a[i][2], a[i][3]);
#pragma unroll
            for (int i = 0; i < NMB; i++) {
                if (MBS == 1 || mb0 + i * MBS < FH_MB) {
                    int d0[4], d1[4];
                    imma_u8(d0, a[i][0], a[i][1], a[i][2], a[i][3], bfr[0][0], bfr[0][1]);
                    imma_u8(d1, a[i][0], a[i][1], a[i][2], a[i][3], bfr[1][0], bfr[1][1]);
                    uint32_t *o = sh_lane + (mb0 + i * MBS) * (16 * FH_WORDS);
                    o[0] = __byte_perm(d0[0], d0[1], 0x5410);
                    o[4] = __byte_perm(d1[0], d1[1], 0x5410);
                    o[8 * FH_WORDS] = __byte_perm(d0[2], d0[3], 0x5410);
                    o[8 * FH_WORDS + 4] = __byte_perm(d1[2], d1[3], 0x5410);
                }
            }
        }
        __syncthreads();
        // ---- vertical pass on packed pairs (5 rows x 16 pixels per thread), then the temporal update ----
        const uint32_t *sgw = sh + trow * FH_WORDS + 4 * xg;
        asm volatile("" : "+r"(M));        // opaque per frame: keeps the 16 single-bit tests of M from being hoisted (and spilled)
        uint8_t *bo = KEEP ? p.blur_out + (((size_t)s * p.T + t) * h + py) * w + px : nullptr;
        uint32_t bits;
        if (t == 0 && !has_bg) bits = fused_rows<KEEP, SAFE, true, true, false>(sgw, bg, M, p, bo, vmask);
        else if (wmasked) bits = p.shift == 8 ? fused_rows<KEEP, SAFE, false, true, true>(sgw, bg, M, p, bo, vmask)
                                              : fused_rows<KEEP, SAFE, false, true, false>(sgw, bg, M, p, bo, vmask);
        else if (p.shift == 8) bits = fused_rows<KEEP, SAFE, false, false, true>(sgw, bg, M, p, bo, vmask);
        else bits = fused_rows<KEEP, SAFE, false, false, false>(sgw, bg, M, p, bo, vmask);
        // ---- two bytes of the bit plane per thread ----
        if (!okA) bits = 0;
        if (!okB) bits &= 0xFFu;
        if (okA) tb[0] = (uint8_t)bits;
        if (okB) tb[8] = (uint8_t)(bits >> 8);
        if (__any_sync(0xffffffffu, bits != 0) && lane == 0) {       // this warp's 4 rows hold something
            int *rr = p.rawrange + 2 * ((size_t)s * p.T + t);
            atomicMax(rr, min(y0 + 4 * warp + 3, h - 1));
            atomicMax(rr + 1, h - 1 - (y0 + 4 * warp));
        }
        tb += 4 * p.flatwords;
    }
#pragma unroll
    for (int i = 0; i < FT_PX / 2; i++) bgt[i * 32] = make_double2(bg[2 * i], bg[2 * i + 1]);
}

// tiled background -> row-major float64 plane
__global__ void k_bg_export_fused(const double *__restrict__ bg, double *__restrict__ dst, int w, int h, int tilesX,
                                  int tilesY, int s) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    int tx = x / FT_W, ty = y / FT_H, tile = ty * tilesX + tx;
    int lx = x - tx * FT_W, ly = y - ty * FT_H;
    int tid = ly * 8 + ((lx & 63) >> 3);           // thread = row, two 8-pixel segments 64 pixels apart
    int warp = tid >> 5, lane = tid & 31;
    int idx = 8 * (lx >> 6) + (lx & 7);            // pixel index inside the thread
    size_t base = ((((size_t)s * tilesX * tilesY + tile) * FUSED_WARPS + warp) * (FT_PX / 2)) * 32;
    dst[(size_t)y * w + x] = bg[(base + (size_t)(idx >> 1) * 32 + lane) * 2 + (idx & 1)];
}

bool fm_fused_supported(const fm_ctx *c) {
    return c->resize_mode == 0 && c->k <= 5 && (c->w % 32) == 0 && c->w >= 4 && c->h >= 4 &&
           ((size_t)c->W * c->H * 3) % 16 == 0;
}

size_t fm_fused_bg_doubles(const fm_ctx *c) {
    size_t tilesX = (c->w + FT_W - 1) / FT_W, tilesY = (c->h + FT_H - 1) / FT_H;
    return (size_t)c->S * tilesX * tilesY * FT_W * FT_H;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled fm_tma_encoder() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)sym;
    }
    return fn;
}

#define FUSED_SMEM (2 * RAW_STAGE + FH_ROWS * FG_WORDS * 4 + FH_ROWS * FH_WORDS * 4 + 16)

int fm_launch_fused(fm_ctx *c, const uint8_t *frames, size_t sstride, size_t fstride, int T, cudaStream_t st) {
    if ((((uintptr_t)frames) & 15) || (sstride & 15) || (fstride & 15)) {
        fm_set_error("fused front end needs 16-byte aligned frames and strides (TMA)");
        return FM_EINVAL;
    }
    PFN_encodeTiled enc = fm_tma_encoder();
    if (!enc) { fm_set_error("cuTensorMapEncodeTiled not available"); return FM_ECUDA; }
    // the call's frames as a 4-D u32 tensor: (W*3/4 words, H rows, T frames, S streams)
    CUtensorMap tmap;
    cuuint64_t dims[4] = {(cuuint64_t)c->W * 3 / 4, (cuuint64_t)c->H, (cuuint64_t)T, (cuuint64_t)c->S};
    cuuint64_t strides[3] = {(cuuint64_t)c->W * 3, (cuuint64_t)fstride, (cuuint64_t)(c->S > 1 ? sstride : fstride * T)};
    cuuint32_t box[4] = {RAW_PITCH / 4, FG_ROWS, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, (void *)frames, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fm_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return FM_ECUDA; }
    FusedParams p;
    p.T = T; p.nvalid = c->nvalid;
    p.w = c->w; p.h = c->h; p.wpr = c->wpr;
    p.tilesX = (c->w + FT_W - 1) / FT_W; p.tilesY = (c->h + FT_H - 1) / FT_H;
    p.bg = c->bg; p.maskbits = c->maskbits; p.tbits = c->tflat;
    p.flatwords = (size_t)c->ntiles * FM_TILE_WORDS;
    p.state = c->state;
    // taps / g for k in {1,3,5}: [0,0,1,0,0] g=256, [0,1,2,1,0] g=64, [1,4,6,4,1] g=16
    int lg;
    if (c->k == 1) { p.b0 = 0; p.b1 = 0; p.b2 = 1; lg = 8; }
    else if (c->k == 3) { p.b0 = 0; p.b1 = 1; p.b2 = 2; lg = 6; }
    else { p.b0 = 1; p.b1 = 4; p.b2 = 6; lg = 4; }
    p.shift = 16 - 2 * lg;
    int r1 = p.shift ? (1 << (p.shift - 1)) : 0;
    p.rnd = r1 | (r1 << 16);
    p.threshold = c->cfg.threshold;
    p.alpha = c->cfg.avg; p.beta = 1.0 - p.alpha;
    p.gray_out = c->gray; p.blur_out = c->blur;
    p.rawrange = c->rawrange;
    const bool keep = (c->cfg.flags & FM_FLAG_KEEP_PLANES) != 0;
    const bool safe = p.alpha >= 0.0 && p.alpha <= 1.0 && p.threshold >= 0;
    fused_bfr(p);
    dim3 grid(p.tilesX, p.tilesY, c->S);
    int rc;
    if ((rc = fm_ensure_smem((const void *)k_fused<true, true>, FUSED_SMEM, c->cfg.device))) return rc;
    if ((rc = fm_ensure_smem((const void *)k_fused<true, false>, FUSED_SMEM, c->cfg.device))) return rc;
    if ((rc = fm_ensure_smem((const void *)k_fused<false, true>, FUSED_SMEM, c->cfg.device))) return rc;
    if ((rc = fm_ensure_smem((const void *)k_fused<false, false>, FUSED_SMEM, c->cfg.device))) return rc;
    if (keep) {
        if (safe) k_fused<true, true><<<grid, FUSED_THREADS, FUSED_SMEM, st>>>(tmap, p);
        else k_fused<true, false><<<grid, FUSED_THREADS, FUSED_SMEM, st>>>(tmap, p);
    } else {
        if (safe) k_fused<false, true><<<grid, FUSED_THREADS, FUSED_SMEM, st>>>(tmap, p);
        else k_fused<false, false><<<grid, FUSED_THREADS, FUSED_SMEM, st>>>(tmap, p);
    }
    FM_LAUNCH_CHECK();
    return FM_OK;
}

int fm_launch_bg_export_fused(fm_ctx *c, int stream, double *dst_dev, cudaStream_t st) {
    dim3 grid((c->w + 127) / 128, c->h);
    k_bg_export_fused<<<grid, 128, 0, st>>>(c->bg, dst_dev, c->w, c->h, (c->w + FT_W - 1) / FT_W,
                                            (c->h + FT_H - 1) / FT_H, stream);
    FM_LAUNCH_CHECK();
    return FM_OK;
}
