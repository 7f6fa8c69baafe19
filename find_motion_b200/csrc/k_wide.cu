// Wide-kernel Gaussian blur (any k that the fused kernel does not take: 7 ... 97 ... 193 ... 385 ...) as two
// banded (Toeplitz) u8 x u8 -> s32 matrix products on the tensor cores (IMMA.16832.U8.U8, 1.13 Pop/s measured
// on B200 against 0.15 Pop/s for IDP.4A: profiles/micro/mma_rate.cu).
// Replaces cv2.GaussianBlur(gray, (k,k), 0) of blur_frame (find_motion/find_motion.py:494) and the
// polygon masks painted into the blur (fm.py:619-635).  Integer arithmetic throughout: bit-exact.
//
// Pass 1 (k_wide_h), horizontal: D[row][col] = sum_j gray[row][col + j - r] * c[j].
//   A = 16 gray rows x 32 columns (ldmatrix from a shared tile with the BORDER_REFLECT_101 columns and rows
//   materialised), B = the taps as a banded 32 x 8 block that depends only on (window step, column block).
//   The 16-bit sums are split into a low-byte and a high-byte plane, laid out for pass 2:
//     plane[frame][group of 32 padded rows][half][column][4 words],   word (4*half + j) byte i = row j' + 8 i of
//   the group (j' = 4*half + j), i.e. the 16 bytes of a (column, half) are one ldmatrix row.
//   Padded row p holds image row reflect101(p - r), so pass 2 has no border logic.
// Pass 2 (k_wide_v), vertical: out[y][x] = (256 * sum_j hi[y + j][x] c[j] + sum_j lo[y + j][x] c[j] + 32768) >> 16.
//   A = 16 columns x 32 padded rows of a byte plane (one ldmatrix.x4 per plane), B = the taps with the k index
//   permuted to the row order of the plane words (k = 4t+i <-> row t + 8i, k = 16+4t+i <-> row 4 + t + 8i).
//   Epilogue: transpose through shared memory, apply the polygon mask bits, store 32 bytes per lane.
#include "fm_common.cuh"

#define WH_COLS 128        // pass 1 CTA tile: 32 padded rows (one group) x 128 columns, warp = 32 x 32
#define WV_ROWS 128        // pass 2 CTA tile: 128 output rows x 32 columns, warp = 32 x 32
#define WV_COLS 32

__device__ __forceinline__ uint32_t wsmem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void wimma(int (&d)[4], const uint32_t (&a)[4], uint2 b) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

struct WideGeom {
    int k, r;
    int R16;        // left pad of the pass-1 window, r rounded up to 16
    int Sh;         // 32-column steps of a pass-1 warp window
    int Sv;         // 32-row groups of a pass-2 warp window
    int pitch;      // bytes per row of the pass-1 shared tile
    int NGa;        // allocated row groups per frame
};

static WideGeom wide_geom(const fm_ctx *c) {
    WideGeom g;
    g.k = c->k; g.r = c->k >> 1;
    g.R16 = (g.r + 15) / 16 * 16;
    g.Sh = (g.R16 + 32 + g.r + 31) / 32;
    g.Sv = 1 + (2 * g.r + 31) / 32;
    int L = WH_COLS - 32 + 32 * g.Sh;                                  // the last warp's last step ends here
    while ((L % 128) != 16 && (L % 128) != 48 && (L % 128) != 80 && (L % 128) != 112) L += 16;   // ldmatrix: 8 rows -> 8 bank groups
    g.pitch = L;
    g.NGa = (c->h + 31) / 32 + g.Sv - 1;
    return g;
}

size_t fm_wide_plane_bytes(const fm_ctx *c) {
    WideGeom g = wide_geom(c);
    return (size_t)c->S * c->Tmax * g.NGa * 2 * c->w * 16;          // one byte plane
}

// Tap tables: entry d = 4 s - nb + 3 (s = window step, nb = 8-wide output block), lane (g = lane >> 2, t = lane & 3),
// two registers of 4 taps.  Pass 1: tap index 8 (d - 3) + kk - g + r - R16, kk = 16 reg + 4 t + i.
// Pass 2: tap index 8 (d - 3) + row(kk) - g with the plane's row order.
int fm_wide_init(fm_ctx *c, const int *taps) {
    WideGeom g = wide_geom(c);
    const int nh = 4 * g.Sh + 3, nv = 4 * g.Sv + 3;
    uint2 *hh = (uint2 *)malloc((size_t)(nh + nv) * 32 * sizeof(uint2));
    if (!hh) { fm_set_error("out of host memory"); return FM_ENOMEM; }
    uint2 *hv = hh + (size_t)nh * 32;
    for (int pass = 0; pass < 2; pass++) {
        uint2 *tab = pass ? hv : hh;
        const int n = pass ? nv : nh;
        for (int d = 0; d < n; d++)
            for (int lane = 0; lane < 32; lane++) {
                const int gq = lane >> 2, t = lane & 3;
                uint32_t reg[2] = {0, 0};
                for (int rg = 0; rg < 2; rg++)
                    for (int i = 0; i < 4; i++) {
                        int j;
                        if (pass == 0) j = 8 * (d - 3) + (16 * rg + 4 * t + i) - gq + g.r - g.R16;
                        else j = 8 * (d - 3) + (4 * rg + t + 8 * i) - gq;
                        if (j >= 0 && j < g.k) reg[rg] |= (uint32_t)(taps[j] & 255) << (8 * i);
                    }
                tab[(size_t)d * 32 + lane] = make_uint2(reg[0], reg[1]);
            }
    }
    cudaError_t e = cudaMalloc((void **)&c->wtab, (size_t)(nh + nv) * 32 * sizeof(uint2));
    if (e == cudaSuccess) e = cudaMemcpy(c->wtab, hh, (size_t)(nh + nv) * 32 * sizeof(uint2), cudaMemcpyHostToDevice);
    free(hh);
    if (e != cudaSuccess) { fm_set_error("wide-blur tap tables: %s", cudaGetErrorString(e)); return FM_ECUDA; }
    return FM_OK;
}

// grid: (ceil(w / 128), NGa, F), 128 threads.  dynamic smem: 32 * pitch + (4 Sh + 3) * 256
__global__ void __launch_bounds__(128) k_wide_h(const uint8_t *__restrict__ gray, uint32_t *__restrict__ plo,
                                                uint32_t *__restrict__ phi, const uint2 *__restrict__ tabg, int w, int h,
                                                int r, int R16, int Sh, int pitch, int NGa) {
    extern __shared__ __align__(16) unsigned char wsm[];
    unsigned char *tile = wsm;                                              // [32][pitch] gray bytes
    uint2 *tab = reinterpret_cast<uint2 *>(wsm + 32 * pitch);             // [4 Sh + 3][32]
    const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
    const int f = blockIdx.z, G = blockIdx.y, X0 = blockIdx.x * WH_COLS;
    for (int i = tid; i < (4 * Sh + 3) * 32; i += 128) tab[i] = __ldg(tabg + i);
    // stage the window: shared column cc <-> image column X0 - R16 + cc (reflected), row rr <-> padded row 32 G + rr
    {
        const int words = pitch >> 2;
        const uint8_t *fr = gray + (size_t)f * h * w;
        for (int i = tid; i < 32 * words; i += 128) {
            const int rr = i / words, cw = i - rr * words;
            const uint8_t *row = fr + (size_t)fm_reflect101(32 * G + rr - r, h) * w;
            const int x = X0 - R16 + 4 * cw;
            uint32_t v;
            if (x >= 0 && x + 3 < w) v = __ldg(reinterpret_cast<const uint32_t *>(row + x));
            else {
                v = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) v |= (uint32_t)row[fm_reflect101(x + b, w)] << (8 * b);
            }
            *reinterpret_cast<uint32_t *>(tile + rr * pitch + 4 * cw) = v;
        }
    }
    __syncthreads();
    int acc[2][4][4];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
            for (int d = 0; d < 4; d++) acc[a][b][d] = 0;
    const uint32_t abase = wsmem_u32(tile) + ((lane & 7) + 8 * ((lane >> 3) & 1)) * pitch + 32 * wq + 16 * (lane >> 4);
    for (int s = 0; s < Sh; s++) {
        uint32_t a0[4], a1[4];
        wldsm_x4(abase + 32 * s, a0);
        wldsm_x4(abase + 32 * s + 16 * pitch, a1);
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            const uint2 b = tab[(4 * s - nb + 3) * 32 + lane];
            wimma(acc[0][nb], a0, b);
            wimma(acc[1][nb], a1, b);
        }
    }
    // thread (g, t): rows g, g+8 (tile 0), g+16, g+24 (tile 1) of columns 2t, 2t+1 of each block = word g of the group
    const int g = lane >> 2, t = lane & 3;
    const size_t gbase = (((size_t)f * NGa + G) * 2 + (g >> 2)) * w;
#pragma unroll
    for (int nb = 0; nb < 4; nb++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int x = X0 + 32 * wq + 8 * nb + 2 * t + e;
            if (x < w) {
                const uint32_t p = __byte_perm(acc[0][nb][e], acc[0][nb][2 + e], 0x5140);     // [r0.b0, r8.b0, r0.b1, r8.b1]
                const uint32_t q = __byte_perm(acc[1][nb][e], acc[1][nb][2 + e], 0x5140);
                const size_t o = (gbase + x) * 4 + (g & 3);
                plo[o] = __byte_perm(p, q, 0x5410);
                phi[o] = __byte_perm(p, q, 0x7632);
            }
        }
}

// grid: (ceil(w / 32), ceil(h / 128), F), 128 threads.
// dynamic smem: 2 planes * (4 + Sv - 1) groups * 2 halves * 32 columns * 16 B  +  (4 Sv + 3) * 256  +  4 * 1024
__global__ void __launch_bounds__(128) k_wide_v(const uint4 *__restrict__ plo, const uint4 *__restrict__ phi,
                                                uint8_t *__restrict__ blur, const uint2 *__restrict__ tabg, int w, int h,
                                                int Sv, int NGa, int wpr, int T, const uint32_t *__restrict__ maskbits) {
    extern __shared__ __align__(16) unsigned char wsm[];
    const int NGt = 4 + Sv - 1;
    uint4 *sA = reinterpret_cast<uint4 *>(wsm);                              // [plane][NGt][half][32 cols]
    uint2 *tab = reinterpret_cast<uint2 *>(wsm + (size_t)2 * NGt * 2 * WV_COLS * 16);
    unsigned char *outb = reinterpret_cast<unsigned char *>(tab + (4 * Sv + 3) * 32);      // [4 warps][32 rows][32 cols]
    const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;
    const int f = blockIdx.z, Y0 = blockIdx.y * WV_ROWS, X0 = blockIdx.x * WV_COLS, G0 = Y0 >> 5;
    for (int i = tid; i < (4 * Sv + 3) * 32; i += 128) tab[i] = __ldg(tabg + i);
    {
        const int per = NGt * 2 * WV_COLS;            // chunks per plane
        for (int i = tid; i < 2 * per; i += 128) {
            const int pl = i >= per, j = pl ? i - per : i;
            const int col = j & (WV_COLS - 1), gh = j >> 5;          // gh = 2 * group + half
            const int G = G0 + (gh >> 1);
            uint4 v = make_uint4(0, 0, 0, 0);
            if (X0 + col < w && G < NGa) v = __ldg((pl ? phi : plo) + (((size_t)f * NGa + G) * 2 + (gh & 1)) * w + X0 + col);
            sA[i] = v;
        }
    }
    __syncthreads();
    int acc[2][2][4][4];       // [plane][column tile][row block][4]
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
            for (int c2 = 0; c2 < 4; c2++)
#pragma unroll
                for (int d = 0; d < 4; d++) acc[a][b][c2][d] = 0;
    // ldmatrix rows: matrices (cols 0-7, half 0), (cols 8-15, half 0), (cols 0-7, half 1), (cols 8-15, half 1)
    const uint32_t abase = wsmem_u32(sA) + (((lane >> 4) * WV_COLS) + (lane & 7) + 8 * ((lane >> 3) & 1)) * 16;
    const uint32_t plane_stride = NGt * 2 * WV_COLS * 16;
    for (int s = 0; s < Sv; s++) {
        uint32_t a[2][2][4];
#pragma unroll
        for (int pl = 0; pl < 2; pl++)
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
                wldsm_x4(abase + pl * plane_stride + (wq + s) * (2 * WV_COLS * 16) + mt * 256, a[pl][mt]);
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            const uint2 b = tab[(4 * s - nb + 3) * 32 + lane];
#pragma unroll
            for (int pl = 0; pl < 2; pl++)
#pragma unroll
                for (int mt = 0; mt < 2; mt++) wimma(acc[pl][mt][nb], a[pl][mt], b);
        }
    }
    // epilogue: thread (g, t) holds columns 16 mt + g (+8), output rows 8 nb + 2t (+1) of the warp's 32 x 32 block
    const int g = lane >> 2, t = lane & 3;
    unsigned char *ob = outb + wq * 1024;
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nb = 0; nb < 4; nb++)
#pragma unroll
            for (int d = 0; d < 4; d++) {
                const int v = ((acc[1][mt][nb][d] << 8) + acc[0][mt][nb][d] + 32768) >> 16;
                ob[(8 * nb + 2 * t + (d & 1)) * 32 + 16 * mt + g + 8 * (d >> 1)] = (unsigned char)v;
            }
    __syncwarp();
    const int y = Y0 + 32 * wq + lane;
    if (y < h) {
        const int sidx = f / T;
        const uint32_t m = maskbits[((size_t)sidx * h + y) * wpr + (X0 >> 5)];
        const uint32_t *src = reinterpret_cast<const uint32_t *>(ob + lane * 32);
        uint8_t *dst = blur + ((size_t)f * h + y) * w + X0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (X0 + 4 * i < w) {
                const uint32_t mk = (m >> (4 * i)) & 0xFu;
                const uint32_t zero = (((mk * 0x00204081u) & 0x01010101u) * 0xFFu);       // masked pixels -> 0xFF bytes
                *reinterpret_cast<uint32_t *>(dst + 4 * i) = src[i] & ~zero;
            }
        }
    }
}

int fm_launch_wide_blur(fm_ctx *c, int T, cudaStream_t st) {
    const WideGeom g = wide_geom(c);
    const int F = c->S * T;
    uint32_t *plo = reinterpret_cast<uint32_t *>(c->hor);
    uint32_t *phi = plo + fm_wide_plane_bytes(c) / 4;
    const uint2 *tabh = c->wtab, *tabv = c->wtab + (size_t)(4 * g.Sh + 3) * 32;
    const size_t smh = (size_t)32 * g.pitch + (size_t)(4 * g.Sh + 3) * 256;
    const size_t smv = (size_t)2 * (4 + g.Sv - 1) * 2 * WV_COLS * 16 + (size_t)(4 * g.Sv + 3) * 256 + 4 * 1024;
    if (smh > 200 * 1024 || smv > 200 * 1024) {
        fm_set_error("Gaussian kernel %d too wide for the tensor-core blur (%zu / %zu bytes of shared memory)", c->k, smh, smv);
        return FM_ERANGE;
    }
    static size_t conf_h[FM_MAX_DEVICES] = {0}, conf_v[FM_MAX_DEVICES] = {0};
    size_t &ch = conf_h[c->cfg.device % FM_MAX_DEVICES], &cv = conf_v[c->cfg.device % FM_MAX_DEVICES];
    if (smh > ch) { FM_CUDA(cudaFuncSetAttribute(k_wide_h, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smh)); ch = smh; }
    if (smv > cv) { FM_CUDA(cudaFuncSetAttribute(k_wide_v, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smv)); cv = smv; }
    dim3 hgrid((c->w + WH_COLS - 1) / WH_COLS, g.NGa, F);
    k_wide_h<<<hgrid, 128, smh, st>>>(c->gray, plo, phi, tabh, c->w, c->h, g.r, g.R16, g.Sh, g.pitch, g.NGa);
    FM_LAUNCH_CHECK();
    dim3 vgrid((c->w + WV_COLS - 1) / WV_COLS, (c->h + WV_ROWS - 1) / WV_ROWS, F);
    k_wide_v<<<vgrid, 128, smv, st>>>(reinterpret_cast<const uint4 *>(plo), reinterpret_cast<const uint4 *>(phi), c->blur,
                                      tabv, c->w, c->h, g.Sv, g.NGa, c->wpr, T, c->maskbits);
    FM_LAUNCH_CHECK();
    return FM_OK;
}
