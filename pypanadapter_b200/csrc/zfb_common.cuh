// Shared device helpers and constants for the zoom-FFT PSD kernels (sm_100a).
#pragma once
#include <stdint.h>

#include "zfb_platform.h"

namespace zfb {

// ---- decimator geometry -------------------------------------------------
// One CTA filters a REGION of one frame that lives in shared memory; each
// thread owns a BLK-sample run of it.  Outputs are only taken from the middle
// T <= TMAX samples; WARM samples on either side absorb the start-up
// transient of regions that do not begin/end at a true chunk edge (the slow
// pole pair of cheby1(8,.05,.4) has radius 0.935: 0.935^320 = 5e-10 in state,
// below fp32 resolution in the outputs).
constexpr int BLK      = 64;                 // samples per thread
constexpr int WARM     = 320;                // warm-up halo each side
// threads per CTA is a kernel template parameter NT (256 or 128):
//   REGION = BLK*NT samples in smem, at most REGION - 2*WARM outputs per tile
constexpr int NTHR_BIG   = 256;              // 16384-sample region, 1 CTA / SM
constexpr int NTHR_SMALL = 128;              //  8192-sample region, 3 CTAs / SM
__host__ __device__ constexpr int region_of(int nt) { return BLK * nt; }
__host__ __device__ constexpr int tmax_of(int nt) { return BLK * nt - 2 * WARM; }
constexpr int BLK_PAD  = BLK + 1;            // +1 complex: conflict-free LDS.64
constexpr int PADLEN   = 27;                 // scipy sosfiltfilt odd extension
constexpr int NSEC     = 4;                  // biquads in cheby1 order 8
constexpr int NSTATE   = 2 * NSEC;
constexpr int JTERMS   = 5;                  // blocks of history in hand-off (64-sample runs)
constexpr int JTERMS32 = 10;                 // same for the 32-sample runs of the strip kernel

// input kinds of a decimation stage / of the Welch kernel
constexpr int KIND_C64_RAW = 0;  // complex64 chunk: flip + LO mix on load
constexpr int KIND_U8_RAW  = 1;  // uint8 IQ chunk: convert + flip + LO mix
constexpr int KIND_C64_MID = 2;  // complex64 intermediate (already mixed)
constexpr int KIND_CS16_RAW = 3; // int16 IQ chunk (SoapySDR CS16): fir_run_kernel only (128 threads, no channel batch)

struct DecimConst {
    float na1[NSEC], na2[NSEC];          // -a1, -a2 of each section
    float g;                             // (b0 of section 0)^2: gain of the forward + backward pass
    float zi[NSEC];                      // steady state of the all-pole cascade per unit scaled input
    float Mp[JTERMS][NSTATE][NSTATE];    // Mp[j] = (state transition over BLK)^j
    float Mp32[JTERMS32][NSTATE][NSTATE];// same over 32 samples (|M32^10| = 3e-9)
    float bc[9];                         // numerator of the causal half G(z) of H(z)H(1/z) (zfb_iirstream.cuh)
};

// per-channel software-LO tables of the channel-batched launches
// (zfb_process_channels_*: many zoom centres over the same frames in one grid)
struct ChannelLo {
    unsigned long long phase_inc;   // frac(f_demod/fs) * 2^64
    float2 run[32];                 // fir_run_kernel: sqrt(2) * exp(-2 pi i f/fs j)
    float2 dec_small[8];            // exact stage 0 / strips: sqrt(2) g^2 exp(-2 pi i f/fs v)
    float2 dec_big[32];             // exp(-2 pi i f/fs * it*STRIP_NT*VEC)  (strip kernel, 128 threads)
    float2 dec_big64[32];           // the same for the 64-thread strip regions
    int    late;                    // fir_run_kernel: mix at the chain's output (FirRunParams::late)
    int    pad_;
    float2 out[16];                 // sqrt(2) * exp(-2 pi i f/fs * (j << NS))
};

// packed fp32x2 arithmetic (Blackwell FFMA2/FADD2/FMUL2): re and im of a
// sample share every real filter coefficient, so one issue slot does both.
__device__ __forceinline__ float2 pk_fma(float a, float2 x, float2 y) {
    return __ffma2_rn(make_float2(a, a), x, y);
}
__device__ __forceinline__ float2 pk_add(float2 x, float2 y) {
    return __fadd2_rn(x, y);
}
__device__ __forceinline__ float2 pk_mul(float a, float2 x) {
    return __fmul2_rn(make_float2(a, a), x);
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// exp(-2*pi*i * frac) for frac = (pos * inc mod 2^64) / 2^64
__device__ __forceinline__ float2 lo_phasor(long long pos, unsigned long long inc) {
    unsigned long long ph = (unsigned long long)pos * inc;
    int p32 = (int)(ph >> 32);
    float s, c;
    sincospif((float)p32 * 4.656612873077393e-10f, &s, &c);   // 2^-31
    return make_float2(c, -s);
}

// pyrtlsdr: u / 127.5 - 1
__device__ __forceinline__ float u8_to_f(unsigned int u) {
    return fmaf((float)u, 1.0f / 127.5f, -1.0f);
}

}  // namespace zfb
