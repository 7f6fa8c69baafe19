// Shared declarations of libfmgpu.so (sm_100a only).  See include/fm_gpu.h for the C ABI and
// DESIGN.md for the data layout.  Arithmetic follows SURVEY.md Appendix A (verified against the
// reference's cv2 call chain, find_motion/find_motion.py:487-494, 619-700, 549-589).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/fm_gpu.h"

#define FM_MAX_K 1024          // widest supported Gaussian (taps)
#define FM_TILE_PX 512         // pixels per background tile: 32 lanes x 16 px
#define FM_TILE_WORDS 16       // 32-bit threshold words per tile
#define FM_TIMING_RING 16384
#define FM_MAX_DEVICES 64      // per-device caches of kernel attributes (cudaFuncSetAttribute is per device)

struct StreamState {           // per stream, device resident (find_motion.py:362-371)
    int has_bg;                // ref_frame is not None
    int counter;               // movement_counter
    int decay;                 // movement_decay
    int cache_len;             // len(frame_cache)
};

struct ResizeTab {             // INTER_AREA decimation tables of one axis
    int *start;                // [dst+1] prefix offsets into idx/wt
    int *idx;                  // source index per tap
    float *wt;                 // float32 weight per tap
    int max_taps;
};

struct CclScratch {            // run-based labelling scratch for `frames` frames at a time
    int frames;                // sub-batch capacity
    int cap;                   // run slots per row
    size_t slots;              // per frame: h * cap
    uint16_t *xs, *xe;         // [frames][slots]
    int *rowcnt;               // [frames][h]
    int *parent;               // [frames][1 + slots]   (id 0 = the outside of the image)
    int *area2;                // [frames][slots]
    int *bbox;                 // [frames][slots][4]  xmin, ymin, xmax, ymax
};

struct fm_ctx {
    fm_config cfg;
    fm_info info;
    int S, Tmax, W, H, w, h, k, wpr;
    int N;                     // w*h
    int ntiles;                // ceil(N / FM_TILE_PX)
    int resize_mode;           // 0 identity, 1 general tables, 2 integer ratio
    bool fused;                // K1 fused stencil+background kernel drives the front end
    bool wide_fused;           // wide Gaussian: the vertical pass runs the temporal stage too (k_wide_vt)
    bool umma;                 // blur + temporal stage on tcgen05 tensor cores (k_umma.cu), k <= 97
    uint8_t *uband;            // band (Toeplitz) operand of the tcgen05 blur (apron-plane kernel)
    uint8_t *uband_f;          // ... of the kernel that reads the BGR frames directly
    bool umma_direct;          // full-resolution mode with TMA-compatible rows: no gray / apron plane at all
    int umma_ra;               // apron class of the direct kernel (16: k <= 33, 48: k <= 97)
    uint8_t *gpad;             // [S][Tmax][Hp][Wp] gray plane with the BORDER_REFLECT_101 apron materialised
    int fx, fy;                // integer ratios (mode 2)
    int maxc;
    // tables
    int *coef;                 // [k] 8.8 fixed-point Gaussian taps
    uint32_t *wtab;            // tap tables of the tensor-core blur (k_wide.cu): pass 1 (uint2 per lane) then pass 2 (uint4 per lane)
    ResizeTab xtab, ytab;
    int *g4start, *g4n, *g4off;   // x taps regrouped in 4-pixel groups with zero-weight padding (k_resize_gray_g4)
    float4 *g4w;
    int g4max;                    // most tap groups of any destination column
    struct RowsPlan *rows;        // plan of the row-per-lane resize kernel (k_resize_rows.cu); null: not applicable
    // planes
    uint8_t *gray;             // [S][Tmax][h][w]
    uint16_t *hor;             // horizontal pass: u16 [S][Tmax][h][w] (naive) or the low/high byte planes of k_wide.cu
    uint8_t *blur;             // [S][Tmax][h][w]  masked blur
    double *bg;                // [S][ntiles][8][32][2]  float64 background, tiled
    uint32_t *maskbits;        // [S][h][wpr]  1 = zero the blur here
    uint32_t *maskflat;        // [S][ntiles*16] same mask, flat bit order (fused kernels)
    uint32_t *tflat;           // [S][Tmax][ntiles*16]  raw threshold, flat bit order
    uint32_t *dil;             // [S][Tmax][h][wpr]  dilated threshold, row-padded bit plane
    uint32_t *fill;            // [S][Tmax][h][wpr]  dilated threshold with holes filled
    int *any;                  // [S][Tmax][4] range of set pixels: (max y, max h-1-y, max word column j, max wpr-1-j), -1 = none
    int *rawrange;             // [S][Tmax][2] row range of the RAW threshold (written by the temporal kernels)
    int *heavy;                // [S][Tmax] frame needs the global-memory labelling kernel
    int *ncomp;                // [S][Tmax]
    int *ncounted;             // [S][Tmax]
    fm_component *comps;       // [S][Tmax][maxc]
    fm_frame_stats *stats;     // [S][Tmax]
    StreamState *state;        // [S]
    int *errflag;              // device error word (capacity overflow), sticky until fm_ctx_check / fm_ctx_reset
    int *nvalid;               // [S] frames of each stream that are real in the current call (ragged batches)
    int *nvalid_host;          // [S] last uploaded copy (uploaded only when it changes)
    CclScratch ccl;
    int *spans;                // mask raster scratch [h][2]
    int last_T;                // frames per stream of the last call
    bool planes_valid;
    // host entry points (fm_process_host, fm_submit_host / fm_wait): two slots, so that the host->device copy of
    // batch i+1 runs on the copy stream while the kernels of batch i run on the compute stream
    uint8_t *stage_dev[2];
    size_t stage_bytes[2];
    fm_frame_stats *stats_pinned[2];   // [S][Tmax] each
    int slot_T[2];             // frames per stream of the batch in flight in the slot (0 = idle)
    int *err_pinned;           // [2] error word of the slot's batch
    cudaStream_t own_stream;   // compute
    cudaStream_t copy_stream;  // host -> device frame copies
    cudaEvent_t ev_copied[2], ev_done[2];
    // timing
    bool timing;               // bracket the kernel groups with events (no sync inside fm_process)
    cudaEvent_t *evs;          // [FM_TIMING_RING][4]
    int ev_pending;
    double t_ms[3];
    int64_t t_calls;
};

extern std::atomic<unsigned long long> g_launches;
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device and per function: one mutex-protected cache for
// all kernels, so that host threads driving contexts on different (or the same) devices do not race
int fm_ensure_smem(const void *func, size_t bytes, int device);
void fm_set_error(const char *fmt, ...);

#define FM_CUDA(call)                                                                     \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            fm_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                         __LINE__);                                                       \
            return FM_ECUDA;                                                              \
        }                                                                                 \
    } while (0)

#define FM_LAUNCH_CHECK()                                                         \
    do {                                                                          \
        g_launches.fetch_add(1, std::memory_order_relaxed);                       \
        cudaError_t e_ = cudaGetLastError();                                      \
        if (e_ != cudaSuccess) {                                                  \
            fm_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), \
                         __FILE__, __LINE__);                                     \
            return FM_ECUDA;                                                      \
        }                                                                         \
    } while (0)

// stage launchers (each enqueues on `st`, returns FM_OK / error)
int fm_launch_frontend(fm_ctx *c, const uint8_t *frames, size_t sstride, size_t fstride, int T,
                       cudaStream_t st);
int fm_launch_temporal(fm_ctx *c, int T, cudaStream_t st);
int fm_launch_fused(fm_ctx *c, const uint8_t *frames, size_t sstride, size_t fstride, int T, cudaStream_t st);
int fm_ccl_configure(fm_ctx *c);
int fm_launch_morph_ccl(fm_ctx *c, int T, cudaStream_t st, fm_frame_stats *stats_out);
int fm_launch_masks(fm_ctx *c, int stream, int n_polys, const int *offs, const int *pts_scaled,
                    int npts, cudaStream_t st);
int fm_launch_bg_export(fm_ctx *c, int stream, double *dst_dev, cudaStream_t st);
int fm_launch_thresh_export(fm_ctx *c, int stream, int t, uint8_t *dst_dev, cudaStream_t st);
int fm_launch_mask_export(fm_ctx *c, int stream, uint8_t *dst_dev, cudaStream_t st);
bool fm_fused_supported(const fm_ctx *c);
size_t fm_wide_plane_bytes(const fm_ctx *c);
int fm_wide_init(fm_ctx *c, const int *taps);
bool fm_wide_fused_supported(const fm_ctx *c);
size_t fm_wide_bg_doubles(const fm_ctx *c);
int fm_launch_bg_export_wide(fm_ctx *c, int stream, double *dst_dev, cudaStream_t st);
int fm_launch_wide_blur(fm_ctx *c, const uint8_t *frames, size_t sstride, size_t fstride, int T, cudaStream_t st);
size_t fm_fused_bg_doubles(const fm_ctx *c);
int fm_launch_bg_export_fused(fm_ctx *c, int stream, double *dst_dev, cudaStream_t st);
int fm_launch_resize_bgr(int device, const uint8_t *src_dev, int W, int H, int w, int h, int mode, int fx, int fy,
                         const ResizeTab &xt, const ResizeTab &yt, uint8_t *dst_dev, cudaStream_t st);
int fm_rows_plan(fm_ctx *c, const int *xstart, const int *xidx, const float *xwt, const int *ystart, const int *yidx);
void fm_rows_free(fm_ctx *c);
int fm_rows_plan_describe(int W, int w, int h, const int *xstart, const int *xidx, const float *xwt, const int *ystart,
                          const int *yidx, fm_rows_plan_info *out);
bool fm_rows_usable(const fm_ctx *c, const uint8_t *frames, size_t sstride, size_t fstride);
int fm_launch_resize_rows(fm_ctx *c, const uint8_t *frames, size_t sstride, size_t fstride, int T, cudaStream_t st);
bool fm_umma_supported(const fm_ctx *c);
bool fm_umma_preferred(const fm_ctx *c);
int fm_umma_init(fm_ctx *c, const int *taps);
size_t fm_umma_bg_doubles(const fm_ctx *c);
int fm_launch_umma_blur(fm_ctx *c, const uint8_t *frames, size_t sstride, size_t fstride, int T, cudaStream_t st);
int fm_launch_bg_export_umma(fm_ctx *c, int stream, double *dst_dev, cudaStream_t st);
int fm_ccl_alloc(CclScratch *s, int frames, int h, int cap);
void fm_ccl_free(CclScratch *s);
int fm_ccl_plane(int device, const uint8_t *plane_host, int w, int h, int max_n, fm_component *out,
                 int *n);

__device__ __forceinline__ int fm_reflect101(int i, int n) {
    // BORDER_REFLECT_101 for any offset
    if ((unsigned)i < (unsigned)n) return i;
    if (n == 1) return 0;
    int p = 2 * (n - 1);
    i %= p;
    if (i < 0) i += p;
    return i >= n ? p - i : i;
}
