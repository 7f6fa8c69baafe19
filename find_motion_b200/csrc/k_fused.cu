// K1: fused, time-blocked stencil + background kernel (full-resolution mode, Gaussian k <= 5).
//
// One CTA owns a 128x64 pixel tile of one stream and walks the T frames of the call in order:
//   BGR tile + 2 px halo --(128-bit/32-bit coalesced loads)--> gray bytes in shared memory
//   --> horizontal then vertical 8.8 fixed-point Gaussian on packed 2x16-bit lanes (registers)
//   --> polygon mask --> bg8 = rne(f32(bg)), |blur - bg8| > threshold --> bit-packed mask out
//   --> bg = fma(bg, 1-alpha, rn(blur*alpha))
// Each thread keeps the float64 background of its 4x8 pixels in registers across all T frames,
// so the background costs one 16 B/px HBM round trip per call instead of per frame.
// Replaces blur_frame + mask_off_areas + find_diff's diff/threshold/accumulateWeighted
// (find_motion/find_motion.py:487-494, 619-635, 246-257, 651-659).
//
// Packed arithmetic: for k in {1,3,5} the taps are g*[b0,b1,b2,b1,b0] with g >= 16, so the
// horizontal sums (<= 255*256/g) and the vertical sums (<= 255*(256/g)^2 <= 65280) fit 16 bits
// and two pixels share one 32-bit IMAD; (ver + 32768) >> 16 == (v' + round) >> shift exactly.
#include "fm_common.cuh"

#define FT_W 128
#define FT_H 64
#define FG_WORDS 34        // gray words per shared row: cols x0-4 .. x0+131
#define FG_ROWS 68         // rows y0-2 .. y0+65
#define FUSED_THREADS 256

struct FusedParams {
    const uint8_t *frames;
    size_t sstride, fstride;
    int T, w, h, wpr;
    int tilesX, tilesY;
    double *bg;                 // [S][tiles][8 warps][8 rows][2 pairs][32 lanes] double2
    const uint32_t *maskbits;   // [S][h][wpr]
    uint32_t *tbits;            // [S][T][flatwords]  (row-padded == flat because w % 32 == 0)
    size_t flatwords;
    const StreamState *state;
    int b0, b1, b2, shift, rnd; // taps / g, and the rounding of the final shift (packed in both lanes)
    int threshold;
    double alpha, beta;
    uint8_t *gray_out, *blur_out;   // [S][T][h][w], only with KEEP
};

__device__ __forceinline__ uint32_t gray4(uint32_t w0, uint32_t w1, uint32_t w2) {
    // 4 BGR pixels in 3 words -> 4 gray bytes.  Y = (3735 B + 19235 G + 9798 R + 16384) >> 15 with the
    // coefficients split in bytes: c = 256*hi + lo, two dp4a per pixel.
    const uint32_t LO_A = 151u | (35u << 8) | (70u << 16);        // bytes [B,G,R,x]
    const uint32_t HI_A = 14u | (75u << 8) | (38u << 16);
    const uint32_t LO_D = (151u << 8) | (35u << 16) | (70u << 24);   // bytes [x,B,G,R]
    const uint32_t HI_D = (14u << 8) | (75u << 16) | (38u << 24);
    uint32_t p1 = __funnelshift_r(w0, w1, 24);                       // [B1,G1,R1,B2]
    uint32_t p2 = __funnelshift_r(w1, w2, 16);                       // [B2,G2,R2,B3]
    uint32_t y0 = (__dp4a(w0, HI_A, 0u) * 256u + __dp4a(w0, LO_A, 16384u)) >> 15;
    uint32_t y1 = (__dp4a(p1, HI_A, 0u) * 256u + __dp4a(p1, LO_A, 16384u)) >> 15;
    uint32_t y2 = (__dp4a(p2, HI_A, 0u) * 256u + __dp4a(p2, LO_A, 16384u)) >> 15;
    uint32_t y3 = (__dp4a(w2, HI_D, 0u) * 256u + __dp4a(w2, LO_D, 16384u)) >> 15;
    return y0 | (y1 << 8) | (y2 << 16) | (y3 << 24);
}

__device__ __forceinline__ double u8_to_f64(uint32_t v) {
    return __dadd_rn(__hiloint2double(0x43300000, (int)v), -4503599627370496.0);
}

template <bool SAFE>
__device__ __forceinline__ int bg8_magic(double b) {
    // 0x4B400000 + rne(|float32(b)|) saturated to 255
    float f = __double2float_rn(b);
    if (!SAFE) f = fminf(fabsf(f), 255.0f);
    return __float_as_int(__fadd_rn(f, 12582912.0f));
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

template <bool KEEP, bool SAFE>
__global__ void __launch_bounds__(FUSED_THREADS, 2) k_fused(FusedParams p) {
    __shared__ uint32_t sg[FG_ROWS * FG_WORDS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = blockIdx.y;
    const int tile = blockIdx.x;
    const int ty = tile / p.tilesX, tx = tile - ty * p.tilesX;
    const int x0 = tx * FT_W, y0 = ty * FT_H;
    const int w = p.w, h = p.h;
    const bool border = (x0 == 0) || (x0 + FT_W >= w) || (y0 == 0) || (y0 + FT_H >= h);
    const bool has_bg = p.state[s].has_bg != 0;

    // this thread's pixels: columns x0 + 4*lane .. +3, rows y0 + 8*warp .. +7
    const int px = x0 + 4 * lane, py = y0 + 8 * warp;
    double2 *bgt = reinterpret_cast<double2 *>(p.bg) +
                   ((((size_t)s * p.tilesX * p.tilesY + tile) * 8 + warp) * 16) * 32 + lane;
    double bg[32];
    if (has_bg) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            double2 v = bgt[i * 32];
            bg[2 * i] = v.x;
            bg[2 * i + 1] = v.y;
        }
    }
    // polygon mask bits of the 32 pixels (bit 4r+c), loaded once per call
    uint32_t M = 0;
    if (px < w) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            int y = py + r;
            if (y < h) {
                uint32_t mw = __ldg(p.maskbits + ((size_t)s * h + y) * p.wpr + (px >> 5));
                M |= ((mw >> (px & 31)) & 0xFu) << (4 * r);
            }
        }
    }
    const uint8_t *src = p.frames + (size_t)s * p.sstride;
    uint32_t *tw = p.tbits + (size_t)s * p.T * p.flatwords;
    const int b0 = p.b0, b1 = p.b1, b2 = p.b2;

    for (int t = 0; t < p.T; t++) {
        __syncthreads();            // previous frame's readers are done with sg
        // ---- BGR -> gray into shared memory (tile + halo), 4 pixels per unit ----
        for (int u = tid; u < FG_ROWS * FG_WORDS; u += FUSED_THREADS) {
            int ry = u / FG_WORDS, ux = u - ry * FG_WORDS;
            int gy = y0 - 2 + ry, gx = x0 - 4 + 4 * ux;
            uint32_t g = 0;
            if ((unsigned)gy < (unsigned)h && (unsigned)gx < (unsigned)w) {
                const uint32_t *q = reinterpret_cast<const uint32_t *>(src + ((size_t)gy * w + gx) * 3);
                g = gray4(__ldg(q), __ldg(q + 1), __ldg(q + 2));
            }
            sg[u] = g;
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
        // ---- separable blur on packed pairs, sliding 5-row window, then the temporal update ----
        const uint32_t *sgw = sg + (8 * warp) * FG_WORDS + lane;
        uint32_t win[5][2];
        uint32_t bits = 0;
#pragma unroll
        for (int rr = 0; rr < 12; rr++) {
            uint32_t W0 = sgw[rr * FG_WORDS], W1 = sgw[rr * FG_WORDS + 1], W2 = sgw[rr * FG_WORDS + 2];
            uint32_t Ea = __byte_perm(W0, 0, 0x4342), Eb = __byte_perm(W1, 0, 0x4140);
            uint32_t Ec = __byte_perm(W1, 0, 0x4342), Ed = __byte_perm(W2, 0, 0x4140);
            uint32_t Oa = __funnelshift_r(Ea, Eb, 16), Ob = __funnelshift_r(Eb, Ec, 16), Oc = __funnelshift_r(Ec, Ed, 16);
            uint32_t h0 = b0 * (Ea + Ec) + b1 * (Oa + Ob) + b2 * Eb;
            uint32_t h1 = b0 * (Eb + Ed) + b1 * (Ob + Oc) + b2 * Ec;
#pragma unroll
            for (int i = 0; i < 4; i++) { win[i][0] = win[i + 1][0]; win[i][1] = win[i + 1][1]; }
            win[4][0] = h0;
            win[4][1] = h1;
            if (rr >= 4) {
                const int r = rr - 4;          // output row of this thread
                uint32_t v[2];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    uint32_t a = b0 * (win[0][j] + win[4][j]) + b1 * (win[1][j] + win[3][j]) + b2 * win[2][j] + p.rnd;
                    v[j] = (a >> p.shift) & 0x00FF00FFu;
                }
                if (KEEP) {
                    int y = py + r;
                    if (y < h && px < w) {
                        uint32_t o = (v[0] & 0xFF) | ((v[0] >> 16) << 8) | ((v[1] & 0xFF) << 16) | ((v[1] >> 16) << 24);
                        uint32_t mk = (M >> (4 * r)) & 0xFu;
                        uint32_t keep = ((mk & 1) ? 0u : 0xFFu) | ((mk & 2) ? 0u : 0xFF00u) | ((mk & 4) ? 0u : 0xFF0000u) |
                                        ((mk & 8) ? 0u : 0xFF000000u);
                        *reinterpret_cast<uint32_t *>(p.blur_out + (((size_t)s * p.T + t) * h + y) * w + px) = o & keep;
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int idx = 4 * r + c;
                    uint32_t sv = (c & 1) ? (v[c >> 1] >> 16) : (v[c >> 1] & 0xFFFFu);
                    if (M & (1u << idx)) sv = 0;                    // mask_off_areas paints BLACK into blur
                    if (t == 0 && !has_bg) bg[idx] = u8_to_f64(sv); // ref_frame = blur.astype(float)
                    int q = bg8_magic<SAFE>(bg[idx]);
                    int d = q - (0x4B400000 + (int)sv);
                    if ((unsigned)(d + p.threshold) > (unsigned)(2 * p.threshold)) bits |= 1u << idx;
                    bg[idx] = __fma_rn(bg[idx], p.beta, __dmul_rn(u8_to_f64(sv), p.alpha));
                }
            }
        }
        // ---- 8 lanes x 8 rows of nibbles -> one 32-pixel word per lane, coalesced store ----
        uint32_t word = nibble_transpose8(bits, lane);
        {
            int y = py + (lane & 7);
            int xw = (x0 >> 5) + (lane >> 3);
            if (y < h && xw < p.wpr) tw[(size_t)y * p.wpr + xw] = word;
        }
        tw += p.flatwords;
        src += p.fstride;
    }
#pragma unroll
    for (int i = 0; i < 16; i++) bgt[i * 32] = make_double2(bg[2 * i], bg[2 * i + 1]);
}

// tiled background -> row-major float64 plane
__global__ void k_bg_export_fused(const double *__restrict__ bg, double *__restrict__ dst, int w, int h, int tilesX,
                                  int tilesY, int s) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    int tx = x / FT_W, ty = y / FT_H, tile = ty * tilesX + tx;
    int lx = x - tx * FT_W, ly = y - ty * FT_H;
    int warp = ly >> 3, r = ly & 7, lane = lx >> 2, c = lx & 3;
    int idx = 4 * r + c;           // pixel index inside the thread
    size_t base = ((((size_t)s * tilesX * tilesY + tile) * 8 + warp) * 16) * 32;
    dst[(size_t)y * w + x] = bg[(base + (size_t)(idx >> 1) * 32 + lane) * 2 + (idx & 1)];
}

bool fm_fused_supported(const fm_ctx *c) {
    return c->resize_mode == 0 && c->k <= 5 && (c->w % 32) == 0 && c->w >= 4 && c->h >= 4 &&
           ((size_t)c->W * c->H * 3) % 4 == 0;
}

size_t fm_fused_bg_doubles(const fm_ctx *c) {
    size_t tilesX = (c->w + FT_W - 1) / FT_W, tilesY = (c->h + FT_H - 1) / FT_H;
    return (size_t)c->S * tilesX * tilesY * FT_W * FT_H;
}

int fm_launch_fused(fm_ctx *c, const uint8_t *frames, size_t sstride, size_t fstride, int T, cudaStream_t st) {
    if ((((uintptr_t)frames) & 3) || (sstride & 3) || (fstride & 3)) {
        fm_set_error("fused front end needs 4-byte aligned frames and strides");
        return FM_EINVAL;
    }
    FusedParams p;
    p.frames = frames; p.sstride = sstride; p.fstride = fstride;
    p.T = T; p.w = c->w; p.h = c->h; p.wpr = c->wpr;
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
    const bool keep = (c->cfg.flags & FM_FLAG_KEEP_PLANES) != 0;
    const bool safe = p.alpha >= 0.0 && p.alpha <= 1.0 && p.threshold >= 0;
    dim3 grid(p.tilesX * p.tilesY, c->S);
    if (keep) {
        if (safe) k_fused<true, true><<<grid, FUSED_THREADS, 0, st>>>(p);
        else k_fused<true, false><<<grid, FUSED_THREADS, 0, st>>>(p);
    } else {
        if (safe) k_fused<false, true><<<grid, FUSED_THREADS, 0, st>>>(p);
        else k_fused<false, false><<<grid, FUSED_THREADS, 0, st>>>(p);
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
