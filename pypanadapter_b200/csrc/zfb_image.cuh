// Waterfall image on the device: the reference keeps a float64
// img_array of (w//4, w) on the host, rolls the WHOLE image by one row per
// update (np.roll, pypanadapter_spectrum.py:1651-1652), redraws grid and tick
// marks (S:1647-1648, S:1655-1662) and hands it to pyqtgraph, which maps
// [minlev, maxlev] onto a 256-entry colour table (S:1592-1594, S:1612-1623;
// pyqtgraph ImageItem: index = clip(trunc((v - min) * 256 / (max - min)), 0,
// 255)).  Here the rows stay in the device ring; a pixel of the image the
// reference would hold after the same sequence of updates is a pure function
// of (y, x), the ring and three counters (wf_pixel), so the image -- as
// float32 img_array, as 8-bit colour indices or as RGBA through the table --
// is produced by one streaming pass only when it is displayed, and only the
// 8-bit image crosses PCIe.  Waterfall.autolevel's np.percentile(img_array[
// img_array < 0], [2, 98]) (S:1676) is an exact order-statistic selection by
// three 11/11/10-bit radix histograms over the same pixel function.
#pragma once
#include "zfb_common.cuh"

namespace zfb {

constexpr int IMG_F32 = 0, IMG_U8 = 1, IMG_RGBA = 2;

struct ImageParams {
    const float *ring;        // [ring_rows][W] finished rows
    int    ring_rows;
    long long written;        // rows written to the ring so far (monotone)
    int    W, H;              // image is [H][W], H = W//4 in the reference
    int    have;              // rows of the image that come from the ring
    int    nseen;             // image_update calls since init_image, saturated at 1<<20
    int    scroll;            // AppState.scroll: > 0 newest row near the bottom, else at the top
    int    tick_step;         // W // 10
    double minlev, scale;     // level mapping: (v - minlev) * scale, scale = 256 / (maxlev - minlev)
    const unsigned int *lut;  // 256 packed RGBA entries (IMG_RGBA)
    void  *out;
};

// tick marks are redrawn on every update and scroll with the image: every row
// that has passed through the marked band keeps them (S:1655-1662)
__device__ __forceinline__ bool wf_tick_row(const ImageParams &p, int y) {
    if (p.nseen <= 0) return false;
    if (p.scroll > 0) {
        if (y >= 5 && y < 15) return true;
        if (y < 5) return p.nseen > 5 - y;          // moved up out of the band
        if (y == p.H - 1) return p.nseen > 6;       // wrapped around from row 0
        return false;
    }
    if (y >= p.H - 10 && y < p.H - 2) return true;
    if (y >= p.H - 2 && y < p.H) return p.nseen > y - (p.H - 3);
    return false;
}

__device__ __forceinline__ bool wf_tick_col(const ImageParams &p, int x) {
    if (x >= p.W - 1 || p.tick_step <= 0) return false;
    const int i = x / p.tick_step;
    return (x - i * p.tick_step == 0) && i != 5 && i != 10;
}

// ring row shown at image row y, or -1 for the -500 fill (S:1631)
__device__ __forceinline__ const float *wf_row(const ImageParams &p, int y) {
    int j;                                           // age of the row, 0 = newest
    if (p.scroll > 0) {
        j = p.H - 2 - y;
        if (j < 0) j += p.H;
    } else {
        j = y;
    }
    if (j >= p.have) return nullptr;
    long long slot = (p.written - 1 - j) % p.ring_rows;
    return p.ring + (size_t)slot * p.W;
}

__device__ __forceinline__ float wf_fix(const ImageParams &p, bool from_ring, bool tick_row, int x, float v) {
    const bool edge = (x == 0) || (x == p.W - 1);
    if (from_ring) {
        if (edge || x == p.W / 2) v = 0.f;           // grid bins zeroed on the row (S:1647-1648)
    } else {
        v = edge ? 0.f : -500.f;                     // init_image (S:1631-1635)
    }
    if (tick_row && wf_tick_col(p, x)) v = 0.f;
    return v;
}

__device__ __forceinline__ unsigned int wf_level(const ImageParams &p, float v) {
    double d = ((double)v - p.minlev) * p.scale;
    if (!(d > 0.0)) d = 0.0;                         // also NaN
    if (d > 255.0) d = 255.0;
    return (unsigned int)d;                          // truncation, like astype after clip
}

template <int KIND>
__global__ void __launch_bounds__(256) wf_image_kernel(const ImageParams p) {
    const int y = blockIdx.y;
    const float *row = wf_row(p, y);
    const bool tick_row = wf_tick_row(p, y);
    const bool vec = (p.W & 3) == 0;
    if (vec) {
        const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
        if (x0 >= p.W) return;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row) v = __ldg((const float4 *)(row + x0));
        float q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) q[e] = wf_fix(p, row != nullptr, tick_row, x0 + e, q[e]);
        const size_t at = (size_t)y * p.W + x0;
        if (KIND == IMG_F32) {
            *(float4 *)((float *)p.out + at) = make_float4(q[0], q[1], q[2], q[3]);
        } else if (KIND == IMG_U8) {
            unsigned int pk = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) pk |= wf_level(p, q[e]) << (8 * e);
            *(unsigned int *)((unsigned char *)p.out + at) = pk;
        } else {
            uint4 c;
            c.x = p.lut[wf_level(p, q[0])];
            c.y = p.lut[wf_level(p, q[1])];
            c.z = p.lut[wf_level(p, q[2])];
            c.w = p.lut[wf_level(p, q[3])];
            *(uint4 *)((unsigned int *)p.out + at) = c;
        }
    } else {
        const int xb = (int)blockIdx.x * 1024;
        const int xe = min(p.W, xb + 1024);
        for (int x = xb + (int)threadIdx.x; x < xe; x += 256) {
            const float v = wf_fix(p, row != nullptr, tick_row, x, row ? row[x] : 0.f);
            const size_t at = (size_t)y * p.W + x;
            if (KIND == IMG_F32) ((float *)p.out)[at] = v;
            else if (KIND == IMG_U8) ((unsigned char *)p.out)[at] = (unsigned char)wf_level(p, v);
            else ((unsigned int *)p.out)[at] = p.lut[wf_level(p, v)];
        }
    }
}

// ---- exact order statistics of {img[y][x] : img[y][x] < 0} ---------------
constexpr int SEL_TARGETS = 4;            // ranks resolved per sweep
constexpr int SEL_BINS = 2048;

struct SelectParams {
    ImageParams img;
    int pass;                             // 0: key bits 31..21, 1: bits 20..10, 2: bits 9..0
    int ntargets;
    unsigned int prefix[SEL_TARGETS];     // pass 1: key >> 21, pass 2: key >> 10 of each target
    unsigned int *hist;                   // [SEL_TARGETS][SEL_BINS] (pass 0 uses histogram 0 only)
    int rows_per_cta;
};

// monotone map of the floats below zero onto unsigned keys (ascending)
__device__ __forceinline__ unsigned int wf_key(float v) { return ~__float_as_uint(v); }
__host__ inline float wf_unkey(unsigned int k) {
    const unsigned int u = ~k;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

__global__ void __launch_bounds__(256) wf_select_kernel(const SelectParams s) {
    __shared__ unsigned int h[SEL_TARGETS * SEL_BINS];
    const ImageParams &p = s.img;
    const int nh = (s.pass == 0) ? 1 : s.ntargets;
    for (int i = threadIdx.x; i < nh * SEL_BINS; i += blockDim.x) h[i] = 0u;
    __syncthreads();
    const int y0 = blockIdx.x * s.rows_per_cta;
    const int y1 = min(p.H, y0 + s.rows_per_cta);
    // runs of equal bins (the -500 fill, flat floors) cost one atomic per run and thread
    int run_at = -1;
    unsigned int run_n = 0;
    for (int y = y0; y < y1; ++y) {
        const float *row = wf_row(p, y);
        const bool tick_row = wf_tick_row(p, y);
        for (int x = threadIdx.x; x < p.W; x += blockDim.x) {
            const float v = wf_fix(p, row != nullptr, tick_row, x, row ? __ldg(row + x) : 0.f);
            if (!(v < 0.f)) continue;
            const unsigned int k = wf_key(v);
            if (s.pass == 0) {
                const int at = (int)(k >> 21);
                if (at == run_at) { ++run_n; } else {
                    if (run_n) atomicAdd(&h[run_at], run_n);
                    run_at = at;
                    run_n = 1;
                }
            } else {
                for (int t = 0; t < s.ntargets; ++t) {
                    const bool hit = (s.pass == 1) ? ((k >> 21) == s.prefix[t]) : ((k >> 10) == s.prefix[t]);
                    if (!hit) continue;
                    const int at = t * SEL_BINS + (int)((s.pass == 1) ? ((k >> 10) & 0x7FFu) : (k & 0x3FFu));
                    if (at == run_at) { ++run_n; } else {
                        if (run_n) atomicAdd(&h[run_at], run_n);
                        run_at = at;
                        run_n = 1;
                    }
                }
            }
        }
    }
    if (run_n) atomicAdd(&h[run_at], run_n);
    __syncthreads();
    for (int i = threadIdx.x; i < nh * SEL_BINS; i += blockDim.x)
        if (h[i]) atomicAdd(&s.hist[i], h[i]);
}

}  // namespace zfb
