// Mask rasterisation kernel: paints the polygons that mask_off_areas draws in BLACK on every
// blurred frame (find_motion/find_motion.py:619-635) ONCE into a bit plane, with the exact
// footprint of cv2.rectangle(FILLED) and cv2.fillConvexPoly: 8-connected
// Bresenham outline (after clipLine, drawn left to right) united with the 16.16 fixed-point
// scanline fill.  One CTA per polygon; edges are drawn by one thread each, the sequential
// edge walker runs on thread 0 and the row spans it produces are filled by the whole CTA.
#include "fm_common.cuh"

__device__ __forceinline__ void set_bit(uint32_t *plane, int wpr, int x, int y) {
    atomicOr(plane + (size_t)y * wpr + (x >> 5), 1u << (x & 31));
}

__device__ void set_span(uint32_t *plane, int wpr, int y, int x0, int x1) {
    for (int j = x0 >> 5; j <= (x1 >> 5); j++) {
        int lo = max(x0 - 32 * j, 0), hi = min(x1 - 32 * j, 31);
        uint32_t m = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
        atomicOr(plane + (size_t)y * wpr + j, m);
    }
}

// cv2.clipLine on [0,w) x [0,h) with 64-bit intermediates
__device__ bool clip_line(int w, int h, long long &x1, long long &y1, long long &x2, long long &y2) {
    if (w <= 0 || h <= 0) return false;
    long long right = w - 1, bottom = h - 1;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (a - y1) * (x2 - x1) / (y2 - y1);
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (a - y2) * (x2 - x1) / (y2 - y1);
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (a - x1) * (y2 - y1) / (x2 - x1);
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (a - x2) * (y2 - y1) / (x2 - x1);
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

__device__ void draw_line8(uint32_t *plane, int wpr, int w, int h, int ax, int ay, int bx, int by) {
    long long x1 = ax, y1 = ay, x2 = bx, y2 = by;
    if (!clip_line(w, h, x1, y1, x2, y2)) return;
    int dx = (int)(x2 - x1), dy = (int)(y2 - y1);
    int x = (int)x1, y = (int)y1;
    if (dx < 0) {            // leftToRight: start from the smaller x
        x = (int)x2; y = (int)y2;
        dx = -dx; dy = -dy;
    }
    int sy = dy >= 0 ? 1 : -1;
    int ady = dy < 0 ? -dy : dy;
    if (dx >= ady) {
        int err = dx - 2 * ady;
        for (int i = 0; i <= dx; i++) {
            set_bit(plane, wpr, x, y);
            if (err < 0) { y += sy; err += 2 * dx - 2 * ady; } else err -= 2 * ady;
            x++;
        }
    } else {
        int err = ady - 2 * dx;
        for (int i = 0; i <= ady; i++) {
            set_bit(plane, wpr, x, y);
            if (err < 0) { x++; err += 2 * ady - 2 * dx; } else err -= 2 * dx;
            y += sy;
        }
    }
}

__global__ void __launch_bounds__(128) k_mask_raster(uint32_t *__restrict__ plane, int w, int h, int wpr,
                                                      const int *__restrict__ offs, const int *__restrict__ pts,
                                                      int *__restrict__ spans /* [npoly][h][2] */) {
    const int poly = blockIdx.x;
    const int p0 = offs[poly], n = offs[poly + 1] - p0;
    const int *v = pts + 2 * p0;
    const int tid = threadIdx.x;
    if (n == 2) {   // cv2.rectangle(..., FILLED): inclusive corners, clipped to the image
        int xa = min(v[0], v[2]), xb = max(v[0], v[2]), ya = min(v[1], v[3]), yb = max(v[1], v[3]);
        xa = max(xa, 0); ya = max(ya, 0); xb = min(xb, w - 1); yb = min(yb, h - 1);
        if (xb < xa || yb < ya) return;
        for (int y = ya + tid; y <= yb; y += blockDim.x) set_span(plane, wpr, y, xa, xb);
        return;
    }
    if (n < 2) {
        if (n == 1 && tid == 0) draw_line8(plane, wpr, w, h, v[0], v[1], v[0], v[1]);
        return;
    }
    // outline
    for (int e = tid; e < n; e += blockDim.x) {
        int a = e == 0 ? n - 1 : e - 1;
        draw_line8(plane, wpr, w, h, v[2 * a], v[2 * a + 1], v[2 * e], v[2 * e + 1]);
    }
    int *sp = spans + (size_t)poly * h * 2;
    for (int y = tid; y < h; y += blockDim.x) { sp[2 * y] = 1; sp[2 * y + 1] = 0; }   // empty
    __syncthreads();
    if (tid == 0 && n >= 3) {
        const long long XY_ONE = 1LL << 16;
        int ymin = v[1], ymax = v[1], xmin = v[0], xmax = v[0], imin = 0;
        for (int i = 0; i < n; i++) {
            int px = v[2 * i], py = v[2 * i + 1];
            if (py < ymin) { ymin = py; imin = i; }
            ymax = max(ymax, py); xmax = max(xmax, px); xmin = min(xmin, px);
        }
        if (!(xmax < 0 || ymax < 0 || xmin >= w || ymin >= h)) {
            ymax = min(ymax, h - 1);
            int eidx[2] = {imin, imin}, edi[2] = {1, n - 1}, eye[2] = {ymin, ymin};
            long long ex[2] = {-XY_ONE, -XY_ONE}, edx[2] = {0, 0};
            int edges = n;
            int y = ymin;
            do {
                for (int i = 0; i < 2; i++) {
                    if (y >= eye[i]) {
                        int idx0 = eidx[i], di = edi[i];
                        int idx = idx0 + di;
                        if (idx >= n) idx -= n;
                        for (; edges-- > 0;) {
                            int ty = v[2 * idx + 1];
                            if (ty > y) {
                                long long xs = (long long)v[2 * idx0] << 16, xe = (long long)v[2 * idx] << 16;
                                eye[i] = ty;
                                edx[i] = ((xe - xs) * 2 + ((long long)ty - y)) / (2 * ((long long)ty - y));
                                ex[i] = xs;
                                eidx[i] = idx;
                                break;
                            }
                            idx0 = idx;
                            idx += di;
                            if (idx >= n) idx -= n;
                        }
                    }
                }
                if (edges < 0) break;
                if (y >= 0) {
                    int l = 0, r = 1;
                    if (ex[0] > ex[1]) { l = 1; r = 0; }
                    int xx1 = (int)((ex[l] + (XY_ONE >> 1)) >> 16);
                    int xx2 = (int)((ex[r] + (XY_ONE >> 1)) >> 16);
                    if (xx2 >= 0 && xx1 < w) {
                        if (xx1 < 0) xx1 = 0;
                        if (xx2 >= w) xx2 = w - 1;
                        sp[2 * y] = xx1;
                        sp[2 * y + 1] = xx2;
                    }
                }
                ex[0] += edx[0];
                ex[1] += edx[1];
            } while (++y <= ymax);
        }
    }
    __syncthreads();
    for (int y = tid; y < h; y += blockDim.x) {
        int a = sp[2 * y], b = sp[2 * y + 1];
        if (b >= a) set_span(plane, wpr, y, a, b);
    }
}

// row-padded mask bit plane -> flat bit order (for kernels that address pixels by flat index)
__global__ void k_mask_flatten(const uint32_t *__restrict__ plane, uint32_t *__restrict__ flat, int w, int h,
                               int wpr, int nwords) {
    int wi = blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= nwords) return;
    uint32_t v = 0;
    long long N = (long long)w * h;
    for (int b = 0; b < 32; b++) {
        long long i = (long long)wi * 32 + b;
        if (i >= N) break;
        int y = (int)(i / w), x = (int)(i - (long long)y * w);
        v |= ((plane[(size_t)y * wpr + (x >> 5)] >> (x & 31)) & 1u) << b;
    }
    flat[wi] = v;
}

int fm_launch_masks(fm_ctx *c, int stream, int n_polys, const int *offs, const int *pts_scaled, int npts,
                    cudaStream_t st) {
    uint32_t *plane = c->maskbits + (size_t)stream * c->h * c->wpr;
    uint32_t *flat = c->maskflat + (size_t)stream * c->ntiles * FM_TILE_WORDS;
    FM_CUDA(cudaMemsetAsync(plane, 0, (size_t)c->h * c->wpr * 4, st));
    if (n_polys > 0) {
        int *d_offs = nullptr, *d_pts = nullptr, *d_spans = nullptr;
        FM_CUDA(cudaMalloc(&d_offs, (size_t)(n_polys + 1) * sizeof(int)));
        FM_CUDA(cudaMalloc(&d_pts, (size_t)npts * 2 * sizeof(int)));
        FM_CUDA(cudaMalloc(&d_spans, (size_t)n_polys * c->h * 2 * sizeof(int)));
        FM_CUDA(cudaMemcpyAsync(d_offs, offs, (size_t)(n_polys + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
        FM_CUDA(cudaMemcpyAsync(d_pts, pts_scaled, (size_t)npts * 2 * sizeof(int), cudaMemcpyHostToDevice, st));
        k_mask_raster<<<n_polys, 128, 0, st>>>(plane, c->w, c->h, c->wpr, d_offs, d_pts, d_spans);
        FM_LAUNCH_CHECK();
        FM_CUDA(cudaStreamSynchronize(st));
        cudaFree(d_offs); cudaFree(d_pts); cudaFree(d_spans);
    }
    int nwords = c->ntiles * FM_TILE_WORDS;
    k_mask_flatten<<<(nwords + 127) / 128, 128, 0, st>>>(plane, flat, c->w, c->h, c->wpr, nwords);
    FM_LAUNCH_CHECK();
    return FM_OK;
}
