// Exact decimate-by-2 stage: one zero-phase cheby1(8, 0.05, 0.4) pass with
// scipy.signal.decimate's per-chunk semantics, as a block-parallel IIR.
//
// Replaces scipy.signal.decimate(x_mix, 2) as called by the reference at
// pypanadapter_spectrum.py:2098 / pypanadapter_thread.py:1534, i.e.
// scipy/signal/_signaltools.py:5317-5369 -> sosfiltfilt (:5091-5204):
//   ext = odd_ext(x, 27); forward cascade from zi*ext[0]; backward cascade
//   from zi*y_fwd[-1]; drop the pads; keep every 2nd sample.
// The first stage also fuses the sample conversion (pyrtlsdr u8 -> complex,
// S:543), the np.flip (S:460,543; T:460) and the LO mix (S:2090-2094).
//
// Parallel scheme (one CTA = one REGION of one frame, resident in smem):
//   * the pass H(z) = g (1+z^-1)^8 / prod A_k(z) is split: the sweeps run the
//     all-pole cascade only, the binomial numerator is one FIR pass per
//     direction afterwards, g^2 is folded into the load (see `pole` below);
//   * thread t owns samples [64t, 64t+64) of the region;
//   * sweep 1: every thread runs the all-pole cascade over its run from a zero
//     state and publishes the 8-value final state z_t;
//   * hand-off: the true incoming state is s_t = sum_j M^(j-1) z_(t-j) (M = the
//     cascade's state transition over 64 samples; |M^5| = 2e-9, so 5 terms);
//     the run that contains the chunk's first extended sample starts instead
//     from the exact steady state zi*ext[0] and ends the sum ("anchor");
//   * sweep 2: rerun from s_t, now storing, software-pipelined across the four
//     sections.  Same again backwards.
//   A region that starts/ends inside the chunk anchors on a steady-state guess
//   WARM=320 samples outside its outputs; true chunk edges are exact.
// The same stage body (exact_stage_inplace) serves the tiled kernel
// (decim2_exact_kernel: mode exact, and the last stage of mode fast) and the
// fused edge-strip cascade of mode fast (strip_cascade_kernel).
#pragma once
#include "zfb_common.cuh"

#ifndef ZFB_SWEEP_UNROLL
#define ZFB_SWEEP_UNROLL 12
#endif

namespace zfb {

constexpr int SWEEP_UNROLL = ZFB_SWEEP_UNROLL;     // steps of the pipelined sweep per loop trip

__constant__ DecimConst c_dec;

struct StageParams {
    const void *in;            // [frames][in_stride] samples of the stage's kind
    float2     *out;           // [frames][out_stride] complex64
    long long   in_stride;     // in samples
    long long   out_stride;
    int         L;             // stage input length per frame
    int         T;             // outputs span per tile (input-index units, %16==0)
    int         flip;
    unsigned long long phase_inc;   // frac(f_demod/fs) * 2^64
    // strip mode (ZFB_MODE_FAST edge strips): blockIdx.y = 2*frame + side; each
    // side is a short chunk of its own (L samples) cut from the frame's ends
    int         strips;
    long long   side_in_off;   // element offset of side 1 inside a frame's input
    long long   side_out_off;  // element offset of side 1 inside a frame's output
    int         pos_off;       // raw kinds: absolute position of side 1's sample 0 (LO phase, flip)
    int         Lfull;         // raw kinds: length of the whole frame (flip index); L otherwise
    int         w_lo[2], w_hi[2];   // outputs [w_lo, w_hi) of a side are written
    const ChannelLo *chan;     // channel-batched launches: LO tables per channel, else null
    int         chan_frames;   // frames per channel in this launch (batch b = ch*chan_frames + frame)
    float2      lo_small[8];   // sqrt(2)*g*exp(-2pi i f/fs v), v = 0..7
    float2      lo_big[32];    // exp(-2pi i f/fs * it*NT*VEC)
};

struct Sec4 {
    float2 w1[NSEC], w2[NSEC];
};

__device__ __forceinline__ void sec_zero(Sec4 &s) {
#pragma unroll
    for (int k = 0; k < NSEC; ++k) {
        s.w1[k] = make_float2(0.f, 0.f);
        s.w2[k] = make_float2(0.f, 0.f);
    }
}

// steady state of the all-pole cascade for the constant (already scaled) input x0
__device__ __forceinline__ void sec_steady(Sec4 &s, float2 x0) {
#pragma unroll
    for (int k = 0; k < NSEC; ++k) {
        float2 w = pk_mul(c_dec.zi[k], x0);
        s.w1[k] = w;
        s.w2[k] = w;
    }
}

// The filter of one pass is H(z) = g (1+z^-1)^8 / prod_k A_k(z).  All of it is
// LTI, so the numerator is pulled out of the recursion: the sweeps run only
// the all-pole cascade 1/prod A_k (2 packed FMAs per section and sample,
//   w = v - a1 w1 - a2 w2, section output = w),
// and the binomial FIR (1+z^-1)^8 is applied once per pass afterwards
// (fir_causal / fir_anticausal_even below).  With the constant-past /
// constant-future boundary model of sosfiltfilt this is exact, not an
// approximation.
__device__ __forceinline__ float2 pole(float2 v, float2 &w1, float2 &w2, float na1, float na2) {
    float2 t = pk_fma(na2, w2, v);
    float2 w = pk_fma(na1, w1, t);
    w2 = w1;
    w1 = w;
    return w;
}

// Full BLK-sample run, software-pipelined across the cascade: in step i
// section k works on sample i-k, so the four recurrences are independent
// instruction streams (the serial form leaves the FMA pipe idle ~60 %).
template <bool BWD, bool STORE, int B = BLK>
__device__ __forceinline__ void sweep_full(float2 *blk, Sec4 &s, const float (&na1)[NSEC],
                                           const float (&na2)[NSEC]) {
#define ZFB_POS(i) (BWD ? (B - 1 - (i)) : (i))
    float2 p0, p1, p2;
    p0 = pole(blk[ZFB_POS(0)], s.w1[0], s.w2[0], na1[0], na2[0]);
    {
        float2 n1 = pole(p0, s.w1[1], s.w2[1], na1[1], na2[1]);
        p0 = pole(blk[ZFB_POS(1)], s.w1[0], s.w2[0], na1[0], na2[0]);
        p1 = n1;
    }
    {
        float2 n2 = pole(p1, s.w1[2], s.w2[2], na1[2], na2[2]);
        float2 n1 = pole(p0, s.w1[1], s.w2[1], na1[1], na2[1]);
        p0 = pole(blk[ZFB_POS(2)], s.w1[0], s.w2[0], na1[0], na2[0]);
        p2 = n2;
        p1 = n1;
    }
    // steady part: B - 3 steps (61 = 1 + 60 for B = 64), unrolled by ZFB_SWEEP_UNROLL
#define ZFB_SWEEP_STEP(i)                                                         \
    {                                                                             \
        float2 y = pole(p2, s.w1[3], s.w2[3], na1[3], na2[3]);                    \
        float2 n2 = pole(p1, s.w1[2], s.w2[2], na1[2], na2[2]);                   \
        float2 n1 = pole(p0, s.w1[1], s.w2[1], na1[1], na2[1]);                   \
        float2 n0 = pole(blk[ZFB_POS(i)], s.w1[0], s.w2[0], na1[0], na2[0]);      \
        if (STORE) blk[ZFB_POS((i) - 3)] = y;                                     \
        p2 = n2;                                                                  \
        p1 = n1;                                                                  \
        p0 = n0;                                                                  \
    }
    ZFB_SWEEP_STEP(3)
#pragma unroll SWEEP_UNROLL
    for (int i = 4; i < B; ++i) ZFB_SWEEP_STEP(i)
#undef ZFB_SWEEP_STEP
    {
        float2 y = pole(p2, s.w1[3], s.w2[3], na1[3], na2[3]);
        float2 n2 = pole(p1, s.w1[2], s.w2[2], na1[2], na2[2]);
        float2 n1 = pole(p0, s.w1[1], s.w2[1], na1[1], na2[1]);
        if (STORE) blk[ZFB_POS(B - 3)] = y;
        p2 = n2;
        p1 = n1;
    }
    {
        float2 y = pole(p2, s.w1[3], s.w2[3], na1[3], na2[3]);
        float2 n2 = pole(p1, s.w1[2], s.w2[2], na1[2], na2[2]);
        if (STORE) blk[ZFB_POS(B - 2)] = y;
        p2 = n2;
    }
    {
        float2 y = pole(p2, s.w1[3], s.w2[3], na1[3], na2[3]);
        if (STORE) blk[ZFB_POS(B - 1)] = y;
    }
#undef ZFB_POS
}

template <bool BWD, bool STORE, int B = BLK>
__device__ __forceinline__ void sweep(float2 *blk, int lo, int hi, Sec4 &s,
                                      const float (&na1)[NSEC], const float (&na2)[NSEC]) {
    if (lo == 0 && hi == B) {
        sweep_full<BWD, STORE, B>(blk, s, na1, na2);
    } else {
        for (int i = lo; i < hi; ++i) {
            const int q = BWD ? (hi - 1 - (i - lo)) : i;
            float2 v = blk[q];
#pragma unroll
            for (int k = 0; k < NSEC; ++k) v = pole(v, s.w1[k], s.w2[k], na1[k], na2[k]);
            if (STORE) blk[q] = v;
        }
    }
}

// 9-tap binomial (1+z^-1)^8 over a 9-sample window w[0..8]
__device__ __forceinline__ float2 binom9(const float2 *w) {
    float2 a = pk_add(w[0], w[8]);
    float2 b = pk_add(w[1], w[7]);
    float2 c = pk_add(w[2], w[6]);
    float2 d = pk_add(w[3], w[5]);
    float2 r = pk_mul(70.0f, w[4]);
    r = pk_fma(56.0f, d, r);
    r = pk_fma(28.0f, c, r);
    r = pk_fma(8.0f, b, r);
    return pk_add(a, r);
}

// y[n] = sum_j C(8,j) x[n-j] over one thread's run, in place; hist = x[-8..-1]
template <int B = BLK>
__device__ __forceinline__ void fir_causal(float2 *blk, const float2 (&hist)[8]) {
    float2 win[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) win[j] = hist[j];
#pragma unroll 2
    for (int b = 0; b < B; b += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) win[8 + j] = blk[b + j];
#pragma unroll
        for (int j = 0; j < 8; ++j) blk[b + j] = binom9(win + j);
#pragma unroll
        for (int j = 0; j < 8; ++j) win[j] = win[8 + j];
    }
}

// y[n] = sum_j C(8,j) x[n+j] at the even n of one thread's run, in place
// (odd positions keep x); ahead = x[BLK .. BLK+7]
template <int B = BLK>
__device__ __forceinline__ void fir_anticausal_even(float2 *blk, const float2 (&ahead)[8]) {
    float2 win[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) win[j] = blk[j];
#pragma unroll 2
    for (int b = 0; b < B; b += 8) {
        if (b + 8 < B) {
#pragma unroll
            for (int j = 0; j < 8; ++j) win[8 + j] = blk[b + 8 + j];
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) win[8 + j] = ahead[j];
        }
#pragma unroll
        for (int j = 0; j < 8; j += 2) blk[b + j] = binom9(win + j);
#pragma unroll
        for (int j = 0; j < 8; ++j) win[j] = win[8 + j];
    }
}

template <int NT>
__device__ __forceinline__ void publish(float2 *zbuf, int t, const Sec4 &s) {
#pragma unroll
    for (int k = 0; k < NSEC; ++k) {
        zbuf[(2 * k) * NT + t]     = s.w1[k];
        zbuf[(2 * k + 1) * NT + t] = s.w2[k];
    }
}

// s = sum_{j=1..jmax} M^(j-1) * z_{t -/+ j}, M = state transition over one B-sample run
template <bool BWD, int NT, int B = BLK>
__device__ __forceinline__ void handoff(const float2 *zbuf, int t, int jmax, Sec4 &s) {
    float2 acc[NSTATE];
#pragma unroll
    for (int r = 0; r < NSTATE; ++r) acc[r] = make_float2(0.f, 0.f);
    for (int j = 1; j <= jmax; ++j) {
        const int tt = BWD ? t + j : t - j;
        float2 z[NSTATE];
#pragma unroll
        for (int r = 0; r < NSTATE; ++r) z[r] = zbuf[r * NT + tt];
        if (j == 1) {
#pragma unroll
            for (int r = 0; r < NSTATE; ++r) acc[r] = z[r];
        } else {
#pragma unroll
            for (int r = 0; r < NSTATE; ++r) {
#pragma unroll
                for (int c = 0; c < 2 * (r / 2 + 1); ++c)    // block lower triangular
                    acc[r] = pk_fma(B == BLK ? c_dec.Mp[j - 1][r][c] : c_dec.Mp32[j - 1][r][c], z[c], acc[r]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NSEC; ++k) {
        s.w1[k] = acc[2 * k];
        s.w2[k] = acc[2 * k + 1];
    }
}

// run-padded layout: one spare complex after every B-sample run
template <int B = BLK>
__device__ __forceinline__ int sidx(int q) {
    return q + (q >> (B == 64 ? 6 : 5));
}
static_assert(BLK == 64, "sidx assumes runs of 64 (default) or 32 samples");

// ---- region load: convert / flip / mix / gain, zero outside [0, L) -------
// 128-bit coalesced loads, CH of them in flight per thread before the first
// use; the scalar path only serves ragged frame ends and unaligned frames.
// uint8 -> float without the slow I2F pipe: the byte is dropped into the
// mantissa of 2^23 (PRMT), minus 2^23 is exact, then u/127.5 - 1 (pyrtlsdr).
__device__ __forceinline__ float2 u8pair_to_iq(unsigned int word, int hi) {
    const unsigned int fi = __byte_perm(word, 0x4B000000u, hi ? 0x7442 : 0x7440);
    const unsigned int fq = __byte_perm(word, 0x4B000000u, hi ? 0x7443 : 0x7441);
    float2 f = make_float2(__uint_as_float(fi), __uint_as_float(fq));
    f = pk_add(f, make_float2(-8388608.0f, -8388608.0f));
    return __ffma2_rn(f, make_float2(1.0f / 127.5f, 1.0f / 127.5f), make_float2(-1.0f, -1.0f));
}

// the same bytes as exact integers 0..255 (conversion folded into a later linear stage)
__device__ __forceinline__ float2 u8pair_to_raw(unsigned int word, int hi) {
    const unsigned int fi = __byte_perm(word, 0x4B000000u, hi ? 0x7442 : 0x7440);
    const unsigned int fq = __byte_perm(word, 0x4B000000u, hi ? 0x7443 : 0x7441);
    return pk_add(make_float2(__uint_as_float(fi), __uint_as_float(fq)), make_float2(-8388608.0f, -8388608.0f));
}

// int16 IQ pair (I in the low half) as exact integers -32768..32767: the sign bit is flipped, the
// 16 bits are dropped into the mantissa of 2^23 (PRMT) and 2^23 + 32768 is taken off again -- no I2F
__device__ __forceinline__ float2 cs16pair_to_raw(unsigned int word) {
    const unsigned int w = word ^ 0x80008000u;
    const unsigned int fi = __byte_perm(w, 0x4B000000u, 0x7410);
    const unsigned int fq = __byte_perm(w, 0x4B000000u, 0x7432);
    return pk_add(make_float2(__uint_as_float(fi), __uint_as_float(fq)), make_float2(-8421376.0f, -8421376.0f));
}

template <int KIND, int NT, bool CHAN = false, int B = BLK>
__device__ __forceinline__ void load_region(float2 *buf, const StageParams &p,
                                            const char *frame_in, int rs, int tid, int posoff, int ch = 0) {
    constexpr int VEC = (KIND == KIND_U8_RAW) ? 8 : 2;
    constexpr int ITERS = B / VEC;
    constexpr int CH = ITERS < 8 ? ITERS : 8;     // loads in flight per thread
    const int L = p.L;
    const bool fl = (KIND != KIND_C64_MID) && p.flip;
    float2 b0 = make_float2(1.f, 0.f);
    const ChannelLo *cl = CHAN ? p.chan + ch : nullptr;
    if (KIND != KIND_C64_MID)
        b0 = lo_phasor((long long)rs + (long long)posoff + (long long)tid * VEC,
                       CHAN ? cl->phase_inc : p.phase_inc);
    const float g = c_dec.g;
    const size_t esz = (KIND == KIND_U8_RAW) ? 2 : 8;
    const int Lf = p.Lfull;               // raw kinds index the whole frame: element = posoff + pos

#pragma unroll 1
    for (int it0 = 0; it0 < ITERS; it0 += CH) {
        uint4 raw[CH];
        bool vec_ok[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int pos = rs + ((it0 + c) * NT + tid) * VEC;
            // sample index of element e: flip ? Lf-1-(posoff+pos+e) : posoff+pos+e
            const long long i0 = fl ? (long long)Lf - VEC - pos - posoff : (long long)pos + posoff;
            const char *a = frame_in + (size_t)i0 * esz;
            vec_ok[c] = (pos >= 0) && (pos + VEC <= L) && ((((uintptr_t)a) & 15) == 0);
            raw[c] = vec_ok[c] ? __ldg((const uint4 *)a) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            const int it = it0 + c;
            const int q = (it * NT + tid) * VEC;
            const int pos = rs + q;
            float2 v[VEC];
            if (vec_ok[c]) {
                if (KIND == KIND_U8_RAW) {
                    const unsigned int wds[4] = {raw[c].x, raw[c].y, raw[c].z, raw[c].w};
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        const int ee = fl ? (VEC - 1 - e) : e;
                        v[e] = u8pair_to_iq(wds[ee >> 1], ee & 1);
                    }
                } else {
                    const float2 lo2 = make_float2(__uint_as_float(raw[c].x), __uint_as_float(raw[c].y));
                    const float2 hi2 = make_float2(__uint_as_float(raw[c].z), __uint_as_float(raw[c].w));
                    v[0] = fl ? hi2 : lo2;
                    v[VEC - 1] = fl ? lo2 : hi2;
                }
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const int pe = pos + e;
                    v[e] = make_float2(0.f, 0.f);
                    if (pe >= 0 && pe < L) {
                        const long long ie = fl ? (long long)Lf - 1 - pe - posoff : (long long)pe + posoff;
                        if (KIND == KIND_U8_RAW) {
                            const unsigned char *src = (const unsigned char *)frame_in;
                            v[e] = make_float2(u8_to_f(src[2 * ie]), u8_to_f(src[2 * ie + 1]));
                        } else {
                            v[e] = __ldg((const float2 *)frame_in + ie);
                        }
                    }
                }
            }
            if (KIND == KIND_C64_MID) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) v[e] = pk_mul(g, v[e]);
            } else {
                const float2 bi = cmul(b0, CHAN ? (NT == 64 ? cl->dec_big64[it] : cl->dec_big[it]) : p.lo_big[it]);
#pragma unroll
                for (int e = 0; e < VEC; ++e)
                    v[e] = cmul(v[e], cmul(bi, CHAN ? cl->dec_small[e] : p.lo_small[e]));
            }
            const int s0 = sidx<B>(q);        // VEC consecutive samples share a run
#pragma unroll
            for (int e = 0; e < VEC; ++e) buf[s0 + e] = v[e];
        }
    }
}

// One exact zero-phase stage on a region that already sits in shared memory
// (stage input of length L at region coordinate q = pos - rs, g^2-scaled).
// On return (after a barrier) the filtered samples are at the EVEN positions:
// output m of the stage = buf[sidx(2m - rs)].
template <int NT, int B = BLK>
__device__ __forceinline__ void exact_stage_inplace(float2 *buf, float2 *zbuf, int L, int rs, int tid) {
    constexpr int REGION = B * NT;
    constexpr int LB = (B == 64) ? 6 : 5;
    constexpr int JT = (B == BLK) ? JTERMS : JTERMS32;
    // odd extension (scipy odd_ext, 27 samples each side) where it falls in the region
    // (slots 0..26: left pad, 32..58: right pad; one-warp CTAs take two turns)
    for (int slot = tid; slot < 32 + PADLEN; slot += NT) {
        if (slot < PADLEN) {
            const int pos = -1 - slot;                // 2*x[0] - x[-pos]
            const int q = pos - rs;
            if (q >= 0) {
                float2 x0 = buf[sidx<B>(-rs)], xm = buf[sidx<B>(-pos - rs)];
                buf[sidx<B>(q)] = make_float2(2.f * x0.x - xm.x, 2.f * x0.y - xm.y);
            }
        } else if (slot >= 32) {
            const int pos = L + (slot - 32);          // 2*x[L-1] - x[2(L-1)-pos]
            const int q = pos - rs;
            if (q < REGION) {
                float2 x1 = buf[sidx<B>(L - 1 - rs)], xm = buf[sidx<B>(2 * (L - 1) - pos - rs)];
                buf[sidx<B>(q)] = make_float2(2.f * x1.x - xm.x, 2.f * x1.y - xm.y);
            }
        }
    }
    __syncthreads();

    // valid part of the region in region coordinates
    const int vlo = max(0, -PADLEN - rs);
    const int vhi = min(REGION, L + PADLEN - rs);
    const int a_f = vlo >> LB;                     // run holding the first valid sample
    const int a_b = (vhi - 1) >> LB;               // run holding the last valid sample
    const bool active = (tid >= a_f) && (tid <= a_b);
    const int lo = max(vlo - tid * B, 0);
    const int hi = min(vhi - tid * B, B);
    float2 *blk = buf + tid * (B + 1);

    float na1[NSEC], na2[NSEC];
#pragma unroll
    for (int k = 0; k < NSEC; ++k) {
        na1[k] = c_dec.na1[k];
        na2[k] = c_dec.na2[k];
    }

    Sec4 s;
    // A run that holds a chunk edge is only partly valid.  Its lane would crawl through the
    // un-pipelined loop (four dependent sections per sample) while the rest of the CTA waits at
    // the next barrier: the run is made whole instead.  The samples before the first valid one
    // are set to ITS value -- the start-up state is the steady state for exactly that constant,
    // so sweeping over them from it changes nothing -- and, before the backward pass, the
    // samples after the last valid one to the forward output there (sosfiltfilt's constant
    // future).  Every active lane then runs the full pipelined sweep.
    if (active && tid == a_f) {
        const float2 c = blk[lo];
        for (int i = 0; i < lo; ++i) blk[i] = c;
    }
    // ---------------- forward ----------------
    if (active) {
        if (tid == a_f) sec_steady(s, blk[0]); else sec_zero(s);
        sweep_full<false, false, B>(blk, s, na1, na2);
        publish<NT>(zbuf, tid, s);
    }
    __syncthreads();
    if (active) {
        if (tid == a_f) sec_steady(s, blk[0]);
        else handoff<false, NT, B>(zbuf, tid, min(JT, tid - a_f), s);
        sweep_full<false, true, B>(blk, s, na1, na2);
    }
    __syncthreads();
    // numerator of the forward pass: (1+z^-1)^8 over the whole region.  The 8
    // samples before a run belong to the neighbour, who rewrites them in place.
    {
        float2 hist[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            hist[j] = (tid > 0) ? buf[sidx<B>(tid * B - 8 + j)] : make_float2(0.f, 0.f);
        __syncthreads();
        if (active) fir_causal<B>(blk, hist);        // runs without valid samples are never read
    }
    __syncthreads();
    // ---------------- backward ----------------
    if (active && tid == a_b) {
        const float2 c = blk[hi - 1];
        for (int i = hi; i < B; ++i) blk[i] = c;
    }
    if (active) {
        if (tid == a_b) sec_steady(s, blk[B - 1]); else sec_zero(s);
        sweep_full<true, false, B>(blk, s, na1, na2);
        publish<NT>(zbuf, tid, s);
    }
    __syncthreads();
    if (active) {
        if (tid == a_b) sec_steady(s, blk[B - 1]);
        else handoff<true, NT, B>(zbuf, tid, min(JT, a_b - tid), s);
        sweep_full<true, true, B>(blk, s, na1, na2);
    }
    __syncthreads();
    // numerator of the backward pass, (1+z)^8, only where a sample is kept
    {
        float2 ahead[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            ahead[j] = (tid < NT - 1) ? buf[sidx<B>((tid + 1) * B + j)] : make_float2(0.f, 0.f);
        __syncthreads();
        if (active) fir_anticausal_even<B>(blk, ahead);
    }
    __syncthreads();
}

template <int KIND, int NT>
__global__ void __launch_bounds__(NT, (NT == NTHR_BIG ? 1 : 3)) decim2_exact_kernel(const StageParams p) {
    ZFB_DYN_SMEM(smem_raw);
    float2 *buf  = reinterpret_cast<float2 *>(smem_raw);          // NT*BLK_PAD
    float2 *zbuf = buf + NT * BLK_PAD;                            // NSTATE*NT

    const int tid   = threadIdx.x;
    const int tile  = blockIdx.x;
    int frame = blockIdx.y, side = 0;
    if (p.strips) {
        side = frame & 1;
        frame >>= 1;
    }
    const int L     = p.L;
    const int p0    = tile * p.T;                 // first output position (even)
    const int rs    = p0 - WARM;                  // region start, ext coordinates

    const size_t esz = (KIND == KIND_U8_RAW) ? 2 : 8;
    const char *frame_in = (const char *)p.in +
                           ((size_t)frame * (size_t)p.in_stride + (size_t)side * (size_t)p.side_in_off) * esz;

    load_region<KIND, NT>(buf, p, frame_in, rs, tid, (KIND != KIND_C64_MID && side) ? p.pos_off : 0);
    __syncthreads();

    exact_stage_inplace<NT>(buf, zbuf, L, rs, tid);

    // ---------------- keep every 2nd sample of [p0, min(p0+T, L)) ----------------
    const int span = min(p.T, L - p0);
    const int nout = (span + 1) >> 1;
    float2 *out = p.out + (size_t)frame * (size_t)p.out_stride + (size_t)side * (size_t)p.side_out_off + (p0 >> 1);
    const int wlo = p.w_lo[side] - (p0 >> 1), whi = p.w_hi[side] - (p0 >> 1);
    for (int i = tid; i < nout; i += NT)
        if (i >= wlo && i < whi) out[i] = buf[sidx(WARM + 2 * i)];
}

// ZFB_MODE_FAST edge strips: the exact cascade on one short chunk cut from a frame's left
// (side 0) or right (side 1) end.  A launch runs `nstages` consecutive stages of every strip
// in ONE CTA (stage s leaves its outputs in shared memory, the part the next stage needs is
// compacted to the front of the region) with NT = region / 64 threads; the strip halves in
// length from stage to stage, so the cascade is cut into launches whose regions fit the
// strip -- 8192 samples (128 threads) for the first stage of R = 16, 2048 (one warp) for the
// last two -- instead of dragging 128 threads, of which 9 to 23 hold samples, and 75 KB of
// shared memory through the late stages (round 1: 26 % of the cfg2 step).  Between launches
// the strips live in a small global buffer [frame][side][cap]; the last launch writes the
// `keep` outputs at the chunk's end.
// (measured in round 1: 256 threads x 32-sample runs -- half the serial chain per thread, but
// twice the hand-off terms -- is 36 % SLOWER than 64-sample runs; the run length stays a
// template parameter of the stage body)
constexpr int STRIP_NT = 128;          // widest region (stage 0 of the fused cascades; LO tables are built for it)
constexpr int STRIP_BLK = 64;
constexpr int STRIP_LEAD = 32;         // region position 0 sits this far before the strip (>= PADLEN, even)
constexpr size_t strip_smem(int nt = STRIP_NT) { return (size_t)(nt * (STRIP_BLK + 1) + NSTATE * nt) * sizeof(float2); }
// threads a stage of `len` input samples needs: lead-in, the samples, the odd extension
__host__ __device__ constexpr int strip_threads_for(int len) {
    return (STRIP_LEAD + len + PADLEN + STRIP_BLK) / STRIP_BLK <= 32 ? 32 : ((STRIP_LEAD + len + PADLEN + STRIP_BLK) / STRIP_BLK <= 64 ? 64 : 128);
}

struct StripParams {
    StageParams st;            // first stage's load: in, in_stride, side_in_off, L = len[0], flip, Lfull, pos_off, LO tables
    int   nstages;             // stages fused in this launch
    int   len[16];             // strip length at the input of stage s of this launch
    int   last;                // 1: the launch ends the cascade and writes `keep` outputs to out
    int   keep;                // outputs of the last stage that are written (K)
    float2 *out;               // decimated chunks [frames][out_stride]
    long long out_stride;
    int   ndec;                // decimated chunk length
    float2 *mid_out;           // !last: [frames][2][mid_cap] first / last next_len outputs of the launch's last stage
    long long mid_cap;
    int   next_len;
};

template <int KIND, bool CHAN = false, int NT = STRIP_NT>
__global__ void __launch_bounds__(NT, (NT == 128 ? 3 : (NT == 64 ? 6 : 12))) strip_cascade_kernel(const StripParams sp) {
    constexpr int B = STRIP_BLK;
    ZFB_DYN_SMEM(smem_raw);
    float2 *buf  = reinterpret_cast<float2 *>(smem_raw);
    float2 *zbuf = buf + NT * (B + 1);
    const int tid = threadIdx.x;
    const int side = blockIdx.x;
    const int frame = blockIdx.y;                      // output (batch) index
    const StageParams &p = sp.st;
    const int ch = CHAN ? frame / p.chan_frames : 0;
    const int in_frame = CHAN ? frame % p.chan_frames : frame;
    const size_t esz = (KIND == KIND_U8_RAW) ? 2 : 8;
    const char *frame_in = (const char *)p.in +
                           ((size_t)in_frame * (size_t)p.in_stride + (size_t)side * (size_t)p.side_in_off) * esz;
    const int rs = -STRIP_LEAD;

    load_region<KIND, NT, CHAN, B>(buf, p, frame_in, rs, tid, (KIND != KIND_C64_MID && side) ? p.pos_off : 0, ch);
    __syncthreads();
    const float g = c_dec.g;
    for (int s = 0; s < sp.nstages; ++s) {
        const int L = sp.len[s];
        exact_stage_inplace<NT, B>(buf, zbuf, L, rs, tid);
        const int nout = (L + 1) >> 1;
        if (s == sp.nstages - 1) {
            if (sp.last) {
                float2 *out = sp.out + (size_t)frame * (size_t)sp.out_stride;
                const int m0 = side ? nout - sp.keep : 0;
                const int d0 = side ? sp.ndec - sp.keep : 0;
                for (int i = tid; i < sp.keep; i += NT) out[d0 + i] = buf[sidx<B>(STRIP_LEAD + 2 * (m0 + i))];
            } else {
                // hand the next launch its input: the first (side 0) / last (side 1) next_len outputs
                float2 *out = sp.mid_out + ((size_t)frame * 2 + (size_t)side) * (size_t)sp.mid_cap;
                const int off = side ? nout - sp.next_len : 0;
                for (int j = tid; j < sp.next_len; j += NT) out[j] = buf[sidx<B>(STRIP_LEAD + 2 * (off + j))];
            }
        } else {
            // the next stage works on the first (side 0) / last (side 1) len[s+1] outputs:
            // move them, gain-scaled, to region positions STRIP_LEAD + j.  dst <= src, so
            // ascending chunks are safe once a chunk's reads precede its writes.
            const int Ln = sp.len[s + 1];
            const int off = side ? nout - Ln : 0;
            constexpr int CH = 8;
            for (int j0 = 0; j0 < Ln; j0 += CH * NT) {
                float2 v[CH];
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int j = j0 + c * NT + tid;
                    v[c] = (j < Ln) ? buf[sidx<B>(STRIP_LEAD + 2 * (off + j))] : make_float2(0.f, 0.f);
                }
                __syncthreads();
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int j = j0 + c * NT + tid;
                    if (j < Ln) buf[sidx<B>(STRIP_LEAD + j)] = pk_mul(g, v[c]);
                }
                __syncthreads();
            }
        }
    }
}

}  // namespace zfb
