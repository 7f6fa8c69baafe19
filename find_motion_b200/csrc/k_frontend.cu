// Generic front end: INTER_AREA resize + gray (K0), gray, separable 8.8 fixed-point Gaussian,
// mask application.  Produces the masked blur plane the temporal kernel consumes.
// Replaces VideoMotion.blur_frame + mask_off_areas (find_motion/find_motion.py:487-494, 619-635).
#include "fm_common.cuh"

// ---------------------------------------------------------------------------------------------
// gray: Y = (3735 B + 19235 G + 9798 R + 16384) >> 15
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fm_gray(uint32_t b, uint32_t g, uint32_t r) {
    return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15;
}

// identity-resize mode: one thread per 4 pixels of a frame (12 bytes in, 4 bytes out)
__global__ void __launch_bounds__(256) k_gray(const uint8_t *__restrict__ frames, size_t sstride,
                                              size_t fstride, int T, int N, uint8_t *__restrict__ gray) {
    int f = blockIdx.y;   // s*T + t
    int s = f / T, t = f - s * T;
    const uint8_t *src = frames + (size_t)s * sstride + (size_t)t * fstride;
    uint8_t *dst = gray + (size_t)f * N;
    int i4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= N) return;
    if (i4 + 4 <= N && ((((uintptr_t)src) & 3) == 0) && ((((uintptr_t)dst) & 3) == 0)) {
        const uint32_t *p = reinterpret_cast<const uint32_t *>(src + (size_t)i4 * 3);
        uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        uint32_t y0 = fm_gray(w0 & 255, (w0 >> 8) & 255, (w0 >> 16) & 255);
        uint32_t y1 = fm_gray(w0 >> 24, w1 & 255, (w1 >> 8) & 255);
        uint32_t y2 = fm_gray((w1 >> 16) & 255, w1 >> 24, w2 & 255);
        uint32_t y3 = fm_gray((w2 >> 8) & 255, (w2 >> 16) & 255, w2 >> 24);
        *reinterpret_cast<uint32_t *>(dst + i4) = y0 | (y1 << 8) | (y2 << 16) | (y3 << 24);
    } else {
        for (int i = i4; i < min(i4 + 4, N); i++) {
            const uint8_t *p = src + (size_t)i * 3;
            dst[i] = (uint8_t)fm_gray(p[0], p[1], p[2]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K0: INTER_AREA resize (float32, unfused, strictly ordered) + gray.  SURVEY.md A.1.
// One CTA per (frame, destination row).  Source rows are staged through shared memory with
// coalesced loads; horizontal chains run per (source row, dx) and the vertical chain per (dx, c).
// ---------------------------------------------------------------------------------------------
struct K0Params {
    const uint8_t *frames;
    size_t sstride, fstride;
    int T, W, H, w, h;
    int mode, fx, fy;            // 1 = tables, 2 = integer ratio
    const int *xstart, *xidx;
    const float *xwt;
    const int *ystart, *yidx;
    const float *ywt;
    int max_ytaps;
    uint8_t *gray;
};

#define K0_THREADS 512
#define K0_GROUPS 4       // source rows in flight (K0_THREADS/K0_GROUPS threads each)

__global__ void __launch_bounds__(K0_THREADS) k_resize_gray(K0Params p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int f = blockIdx.y;
    const int s = f / p.T, t = f - s * p.T;
    const int dy = blockIdx.x;
    const uint8_t *src = p.frames + (size_t)s * p.sstride + (size_t)t * p.fstride;
    const int rowbytes = p.W * 3;
    const int rowpitch = (rowbytes + 16 + 15) & ~15;          // room for a misaligned start
    unsigned char *rowbuf = smem;                               // [K0_GROUPS][rowpitch]
    float *bufs = reinterpret_cast<float *>(smem + (size_t)K0_GROUPS * rowpitch);   // [ny][3w]
    const int w3 = p.w * 3;
    const int tid = threadIdx.x;
    const int gsz = K0_THREADS / K0_GROUPS;
    const int grp = tid / gsz, gt = tid - grp * gsz;

    int y0, ny;
    if (p.mode == 1) {
        y0 = p.ystart[dy];
        ny = p.ystart[dy + 1] - y0;
    } else {
        y0 = 0;
        ny = p.fy;
    }

    for (int jb = 0; jb < ny; jb += K0_GROUPS) {
        int j = jb + grp;
        int mis = 0;
        if (j < ny) {
            int sy = (p.mode == 1) ? p.yidx[y0 + j] : dy * p.fy + j;
            const uint8_t *row = src + (size_t)sy * rowbytes;
            mis = (int)((uintptr_t)row & 15);
            const uint4 *row16 = reinterpret_cast<const uint4 *>(row - mis);
            int n16 = (mis + rowbytes + 15) >> 4;
            uint4 *dst16 = reinterpret_cast<uint4 *>(rowbuf + (size_t)grp * rowpitch);
            // partial first/last 16-byte blocks are loaded bytewise so that nothing outside the
            // row (possibly outside the caller's buffer) is touched
            for (int i = gt; i < n16; i += gsz) {
                if ((i == 0 && mis) || (i == n16 - 1 && ((mis + rowbytes) & 15))) {
                    unsigned char *d = reinterpret_cast<unsigned char *>(dst16 + i);
                    for (int b = 0; b < 16; b++) {
                        int o = i * 16 + b - mis;
                        d[b] = (o >= 0 && o < rowbytes) ? row[o] : 0;
                    }
                } else {
                    dst16[i] = __ldg(row16 + i);
                }
            }
        }
        __syncthreads();
        if (j < ny) {
            const unsigned char *rb = rowbuf + (size_t)grp * rowpitch + mis;
            float *out = bufs + (size_t)j * w3;
            if (p.mode == 1) {
                for (int dx = gt; dx < p.w; dx += gsz) {
                    int a = p.xstart[dx], b = p.xstart[dx + 1];
                    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
                    for (int q = a; q < b; q++) {
                        const unsigned char *px = rb + p.xidx[q] * 3;
                        float al = p.xwt[q];
                        b0 = __fadd_rn(b0, __fmul_rn((float)px[0], al));
                        b1 = __fadd_rn(b1, __fmul_rn((float)px[1], al));
                        b2 = __fadd_rn(b2, __fmul_rn((float)px[2], al));
                    }
                    out[dx * 3 + 0] = b0;
                    out[dx * 3 + 1] = b1;
                    out[dx * 3 + 2] = b2;
                }
            } else {   // integer ratio: exact integer row sums, kept as int bit patterns
                int *outi = reinterpret_cast<int *>(out);
                for (int dx = gt; dx < p.w; dx += gsz) {
                    int s0 = 0, s1 = 0, s2 = 0;
                    const unsigned char *px = rb + dx * p.fx * 3;
                    for (int q = 0; q < p.fx; q++, px += 3) {
                        s0 += px[0];
                        s1 += px[1];
                        s2 += px[2];
                    }
                    outi[dx * 3 + 0] = s0;
                    outi[dx * 3 + 1] = s1;
                    outi[dx * 3 + 2] = s2;
                }
            }
        }
        __syncthreads();
    }
    // vertical chains + rounding + gray
    unsigned char *small = rowbuf;        // reuse: [3w] resized BGR bytes of this row
    for (int i = tid; i < w3; i += K0_THREADS) {
        int v;
        if (p.mode == 1) {
            float sum = __fmul_rn(p.ywt[y0], bufs[i]);
            for (int j = 1; j < ny; j++) sum = __fadd_rn(sum, __fmul_rn(p.ywt[y0 + j], bufs[(size_t)j * w3 + i]));
            v = __float2int_rn(sum);
        } else {
            const int *bi = reinterpret_cast<const int *>(bufs);
            int sum = 0;
            for (int j = 0; j < ny; j++) sum += bi[(size_t)j * w3 + i];
            if (p.fx == 2 && p.fy == 2) {
                v = (sum + 2) >> 2;
            } else {
                float sc = __fdiv_rn(1.f, (float)(p.fx * p.fy));
                v = __float2int_rn(__fmul_rn((float)sum, sc));
            }
        }
        small[i] = (unsigned char)min(max(v, 0), 255);
    }
    __syncthreads();
    uint8_t *g = p.gray + ((size_t)f * p.h + dy) * p.w;
    for (int dx = tid; dx < p.w; dx += K0_THREADS)
        g[dx] = (uint8_t)fm_gray(small[dx * 3], small[dx * 3 + 1], small[dx * 3 + 2]);
}

// ---------------------------------------------------------------------------------------------
// generic separable Gaussian, two passes through a u16 plane.
// This is the fallback for wide kernels; the fused kernel (k_fused.cu) covers small k.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_hblur(const uint8_t *__restrict__ gray, uint16_t *__restrict__ hor,
                                               const int *__restrict__ coef, int k, int w, int h) {
    extern __shared__ int sc[];
    for (int i = threadIdx.x; i < k; i += blockDim.x) sc[i] = coef[i];
    __syncthreads();
    int f = blockIdx.z, y = blockIdx.y;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    const uint8_t *row = gray + ((size_t)f * h + y) * w;
    int r = k >> 1;
    int acc = 0;
    if (x - r >= 0 && x + r < w) {
        const uint8_t *q = row + x - r;
        for (int i = 0; i < k; i++) acc += sc[i] * q[i];
    } else {
        for (int i = 0; i < k; i++) acc += sc[i] * row[fm_reflect101(x + i - r, w)];
    }
    hor[((size_t)f * h + y) * w + x] = (uint16_t)acc;
}

__global__ void __launch_bounds__(256) k_vblur(const uint16_t *__restrict__ hor, uint8_t *__restrict__ blur,
                                               const int *__restrict__ coef, int k, int w, int h, int wpr,
                                               int T, const uint32_t *__restrict__ maskbits) {
    extern __shared__ int sc[];
    for (int i = threadIdx.x; i < k; i += blockDim.x) sc[i] = coef[i];
    __syncthreads();
    int f = blockIdx.z, y = blockIdx.y;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    const uint16_t *pl = hor + (size_t)f * h * w;
    int r = k >> 1;
    int acc = 0;
    if (y - r >= 0 && y + r < h) {
        const uint16_t *q = pl + (size_t)(y - r) * w + x;
        for (int j = 0; j < k; j++) acc += sc[j] * q[(size_t)j * w];
    } else {
        for (int j = 0; j < k; j++) acc += sc[j] * pl[(size_t)fm_reflect101(y + j - r, h) * w + x];
    }
    int v = (acc + 32768) >> 16;
    int s = f / T;
    uint32_t m = maskbits[((size_t)s * h + y) * wpr + (x >> 5)];
    if ((m >> (x & 31)) & 1) v = 0;
    blur[((size_t)f * h + y) * w + x] = (uint8_t)v;
}

int fm_launch_frontend(fm_ctx *c, const uint8_t *frames, size_t sstride, size_t fstride, int T,
                       cudaStream_t st) {
    const int F = c->S * T;
    if (c->resize_mode == 0) {
        dim3 grid(((c->N + 3) / 4 + 255) / 256, F);
        k_gray<<<grid, 256, 0, st>>>(frames, sstride, fstride, T, c->N, c->gray);
        FM_LAUNCH_CHECK();
    } else {
        K0Params p;
        p.frames = frames; p.sstride = sstride; p.fstride = fstride;
        p.T = T; p.W = c->W; p.H = c->H; p.w = c->w; p.h = c->h;
        p.mode = c->resize_mode; p.fx = c->fx; p.fy = c->fy;
        p.xstart = c->xtab.start; p.xidx = c->xtab.idx; p.xwt = c->xtab.wt;
        p.ystart = c->ytab.start; p.yidx = c->ytab.idx; p.ywt = c->ytab.wt;
        p.max_ytaps = c->resize_mode == 1 ? c->ytab.max_taps : c->fy;
        p.gray = c->gray;
        int rowpitch = (c->W * 3 + 16 + 15) & ~15;
        size_t smem = (size_t)K0_GROUPS * rowpitch + (size_t)p.max_ytaps * c->w * 3 * sizeof(float);
        if (smem > 220 * 1024) {
            fm_set_error("resize front end needs %zu bytes of shared memory (frame too wide / ratio too large)", smem);
            return FM_ERANGE;
        }
        static size_t configured = 0;
        if (smem > configured) {
            FM_CUDA(cudaFuncSetAttribute(k_resize_gray, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = smem;
        }
        dim3 grid(c->h, F);
        k_resize_gray<<<grid, K0_THREADS, smem, st>>>(p);
        FM_LAUNCH_CHECK();
    }
    // separable blur
    dim3 bgrid((c->w + 255) / 256, c->h, F);
    size_t sm = (size_t)c->k * sizeof(int);
    k_hblur<<<bgrid, 256, sm, st>>>(c->gray, c->hor, c->coef, c->k, c->w, c->h);
    FM_LAUNCH_CHECK();
    k_vblur<<<bgrid, 256, sm, st>>>(c->hor, c->blur, c->coef, c->k, c->w, c->h, c->wpr, T, c->maskbits);
    FM_LAUNCH_CHECK();
    return FM_OK;
}
