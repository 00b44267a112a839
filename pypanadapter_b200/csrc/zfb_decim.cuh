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
//   * thread t owns samples [64t, 64t+64) of the region;
//   * sweep 1: every thread runs the 4-biquad cascade over its run from a zero
//     state and publishes the 8-value final state z_t;
//   * hand-off: the true incoming state is s_t = sum_j M^(j-1) z_(t-j) (M = the
//     cascade's state transition over 64 samples; |M^5| < 4e-8 so 5 terms);
//     the run that contains the chunk's first extended sample starts instead
//     from the exact steady state zi*ext[0] and ends the sum ("anchor");
//   * sweep 2: rerun from s_t, now storing.  Same again backwards.
//   A region that starts/ends inside the chunk anchors on a steady-state guess
//   WARM=320 samples outside its outputs; true chunk edges are exact.
#pragma once
#include "zfb_common.cuh"

namespace zfb {

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
    float2      lo_small[8];   // sqrt(2)*g*exp(-2pi i f/fs v), v = 0..7
    float2      lo_big[32];    // exp(-2pi i f/fs * it*NTHR*VEC)
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

// steady state for the constant (already g-scaled) input x0
__device__ __forceinline__ void sec_steady(Sec4 &s, float2 x0) {
#pragma unroll
    for (int k = 0; k < NSEC; ++k) {
        float2 w = pk_mul(c_dec.zi[k], x0);
        s.w1[k] = w;
        s.w2[k] = w;
    }
}

// one sample through the cascade, direct form II with numerator (1+z^-1)^2:
//   w = v - a1 w1 - a2 w2 ;  y = w + 2 w1 + w2
__device__ __forceinline__ float2 cascade(float2 v, Sec4 &s, const float (&na1)[NSEC],
                                          const float (&na2)[NSEC]) {
#pragma unroll
    for (int k = 0; k < NSEC; ++k) {
        float2 t = pk_fma(na2[k], s.w2[k], v);
        float2 w = pk_fma(na1[k], s.w1[k], t);
        float2 u = pk_add(w, s.w2[k]);
        v = pk_fma(2.0f, s.w1[k], u);
        s.w2[k] = s.w1[k];
        s.w1[k] = w;
    }
    return v;
}

template <bool BWD, bool STORE>
__device__ __forceinline__ void sweep(float2 *blk, int lo, int hi, Sec4 &s,
                                      const float (&na1)[NSEC], const float (&na2)[NSEC]) {
    if (lo == 0 && hi == BLK) {
#pragma unroll 8
        for (int i = 0; i < BLK; ++i) {
            const int q = BWD ? (BLK - 1 - i) : i;
            float2 v = cascade(blk[q], s, na1, na2);
            if (STORE) blk[q] = v;
        }
    } else {
        for (int i = lo; i < hi; ++i) {
            const int q = BWD ? (hi - 1 - (i - lo)) : i;
            float2 v = cascade(blk[q], s, na1, na2);
            if (STORE) blk[q] = v;
        }
    }
}

__device__ __forceinline__ void publish(float2 *zbuf, int t, const Sec4 &s) {
#pragma unroll
    for (int k = 0; k < NSEC; ++k) {
        zbuf[(2 * k) * NTHR + t]     = s.w1[k];
        zbuf[(2 * k + 1) * NTHR + t] = s.w2[k];
    }
}

// s = sum_{j=1..jmax} Mp[j-1] * z_{t -/+ j}
template <bool BWD>
__device__ __forceinline__ void handoff(const float2 *zbuf, int t, int jmax, Sec4 &s) {
    float2 acc[NSTATE];
#pragma unroll
    for (int r = 0; r < NSTATE; ++r) acc[r] = make_float2(0.f, 0.f);
    for (int j = 1; j <= jmax; ++j) {
        const int tt = BWD ? t + j : t - j;
        float2 z[NSTATE];
#pragma unroll
        for (int r = 0; r < NSTATE; ++r) z[r] = zbuf[r * NTHR + tt];
        if (j == 1) {
#pragma unroll
            for (int r = 0; r < NSTATE; ++r) acc[r] = z[r];
        } else {
#pragma unroll
            for (int r = 0; r < NSTATE; ++r) {
#pragma unroll
                for (int c = 0; c < 2 * (r / 2 + 1); ++c)    // block lower triangular
                    acc[r] = pk_fma(c_dec.Mp[j - 1][r][c], z[c], acc[r]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NSEC; ++k) {
        s.w1[k] = acc[2 * k];
        s.w2[k] = acc[2 * k + 1];
    }
}

__device__ __forceinline__ int sidx(int q) { return q + (q >> 6); }   // BLK_PAD layout

// ---- region load: convert / flip / mix / gain, zero outside [0, L) -------
template <int KIND>
__device__ __forceinline__ void load_region(float2 *buf, const StageParams &p,
                                            const char *frame_in, int rs, int tid) {
    constexpr int VEC = (KIND == KIND_U8_RAW) ? 8 : 2;
    constexpr int ITERS = REGION / (NTHR * VEC);
    const int L = p.L;
    float2 b0 = make_float2(1.f, 0.f);
    if (KIND != KIND_C64_MID) b0 = lo_phasor((long long)rs + (long long)tid * VEC, p.phase_inc);
    const float g = c_dec.g;

#pragma unroll 2
    for (int it = 0; it < ITERS; ++it) {
        const int q = (it * NTHR + tid) * VEC;
        const int pos = rs + q;
        float2 v[VEC];
        if (pos >= L || pos + VEC <= 0) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) v[e] = make_float2(0.f, 0.f);
        } else {
            const bool full = (pos >= 0) && (pos + VEC <= L);
            if (KIND == KIND_U8_RAW) {
                const unsigned char *src = (const unsigned char *)frame_in;
                // sample index of element e: flip ? L-1-(pos+e) : pos+e
                const long long i0 = p.flip ? (long long)L - VEC - pos : (long long)pos;
                const unsigned char *a = src + 2 * i0;
                if (full && ((((uintptr_t)a) & 15) == 0)) {
                    uint4 raw = __ldg((const uint4 *)a);
                    unsigned int wds[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        const int ee = p.flip ? (VEC - 1 - e) : e;
                        unsigned int h = (wds[ee >> 1] >> ((ee & 1) * 16)) & 0xffffu;
                        v[e] = make_float2(u8_to_f(h & 0xffu), u8_to_f(h >> 8));
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        const int pe = pos + e;
                        if (pe >= 0 && pe < L) {
                            const long long ie = p.flip ? (long long)L - 1 - pe : (long long)pe;
                            v[e] = make_float2(u8_to_f(src[2 * ie]), u8_to_f(src[2 * ie + 1]));
                        } else {
                            v[e] = make_float2(0.f, 0.f);
                        }
                    }
                }
            } else {
                const float2 *src = (const float2 *)frame_in;
                const bool fl = (KIND == KIND_C64_RAW) && p.flip;
                const long long i0 = fl ? (long long)L - VEC - pos : (long long)pos;
                const float2 *a = src + i0;
                if (full && ((((uintptr_t)a) & 15) == 0)) {
                    float4 raw = __ldg((const float4 *)a);
                    if (fl) {
                        v[0] = make_float2(raw.z, raw.w);
                        v[1] = make_float2(raw.x, raw.y);
                    } else {
                        v[0] = make_float2(raw.x, raw.y);
                        v[1] = make_float2(raw.z, raw.w);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        const int pe = pos + e;
                        if (pe >= 0 && pe < L) {
                            const long long ie = fl ? (long long)L - 1 - pe : (long long)pe;
                            v[e] = __ldg(src + ie);
                        } else {
                            v[e] = make_float2(0.f, 0.f);
                        }
                    }
                }
            }
            if (KIND == KIND_C64_MID) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) v[e] = pk_mul(g, v[e]);
            } else {
                const float2 bi = cmul(b0, p.lo_big[it]);
#pragma unroll
                for (int e = 0; e < VEC; ++e) v[e] = cmul(v[e], cmul(bi, p.lo_small[e]));
            }
        }
        const int s0 = sidx(q);           // VEC consecutive samples share a block
#pragma unroll
        for (int e = 0; e < VEC; ++e) buf[s0 + e] = v[e];
    }
}

template <int KIND>
__global__ void __launch_bounds__(NTHR, 1) decim2_exact_kernel(const StageParams p) {
    ZFB_DYN_SMEM(smem_raw);
    float2 *buf  = reinterpret_cast<float2 *>(smem_raw);          // NTHR*BLK_PAD
    float2 *zbuf = buf + NTHR * BLK_PAD;                          // NSTATE*NTHR

    const int tid   = threadIdx.x;
    const int tile  = blockIdx.x;
    const int frame = blockIdx.y;
    const int L     = p.L;
    const int p0    = tile * p.T;                 // first output position (even)
    const int rs    = p0 - WARM;                  // region start, ext coordinates

    const size_t esz = (KIND == KIND_U8_RAW) ? 2 : 8;
    const char *frame_in = (const char *)p.in + (size_t)frame * (size_t)p.in_stride * esz;

    load_region<KIND>(buf, p, frame_in, rs, tid);
    __syncthreads();

    // odd extension (scipy odd_ext, 27 samples each side) where it falls in the region
    if (tid < PADLEN) {
        const int pos = -1 - tid;                 // 2*x[0] - x[-pos]
        const int q = pos - rs;
        if (q >= 0) {
            float2 x0 = buf[sidx(-rs)], xm = buf[sidx(-pos - rs)];
            buf[sidx(q)] = make_float2(2.f * x0.x - xm.x, 2.f * x0.y - xm.y);
        }
    } else if (tid >= 32 && tid < 32 + PADLEN) {
        const int pos = L + (tid - 32);           // 2*x[L-1] - x[2(L-1)-pos]
        const int q = pos - rs;
        if (q < REGION) {
            float2 x1 = buf[sidx(L - 1 - rs)], xm = buf[sidx(2 * (L - 1) - pos - rs)];
            buf[sidx(q)] = make_float2(2.f * x1.x - xm.x, 2.f * x1.y - xm.y);
        }
    }
    __syncthreads();

    // valid part of the region in region coordinates
    const int vlo = max(0, -PADLEN - rs);
    const int vhi = min(REGION, L + PADLEN - rs);
    const int a_f = vlo >> 6;                     // run holding the first valid sample
    const int a_b = (vhi - 1) >> 6;               // run holding the last valid sample
    const bool active = (tid >= a_f) && (tid <= a_b);
    const int lo = max(vlo - tid * BLK, 0);
    const int hi = min(vhi - tid * BLK, BLK);
    float2 *blk = buf + tid * BLK_PAD;

    float na1[NSEC], na2[NSEC];
#pragma unroll
    for (int k = 0; k < NSEC; ++k) {
        na1[k] = c_dec.na1[k];
        na2[k] = c_dec.na2[k];
    }

    Sec4 s;
    // ---------------- forward ----------------
    if (active) {
        if (tid == a_f) sec_steady(s, blk[lo]); else sec_zero(s);
        sweep<false, false>(blk, lo, hi, s, na1, na2);
        publish(zbuf, tid, s);
    }
    __syncthreads();
    if (active) {
        if (tid == a_f) sec_steady(s, blk[lo]);
        else handoff<false>(zbuf, tid, min(JTERMS, tid - a_f), s);
        sweep<false, true>(blk, lo, hi, s, na1, na2);
    }
    __syncthreads();
    // ---------------- backward ----------------
    if (active) {
        if (tid == a_b) sec_steady(s, blk[hi - 1]); else sec_zero(s);
        sweep<true, false>(blk, lo, hi, s, na1, na2);
        publish(zbuf, tid, s);
    }
    __syncthreads();
    if (active) {
        if (tid == a_b) sec_steady(s, blk[hi - 1]);
        else handoff<true>(zbuf, tid, min(JTERMS, a_b - tid), s);
        sweep<true, true>(blk, lo, hi, s, na1, na2);
    }
    __syncthreads();

    // ---------------- keep every 2nd sample of [p0, min(p0+T, L)) ----------------
    const int span = min(p.T, L - p0);
    const int nout = (span + 1) >> 1;
    float2 *out = p.out + (size_t)frame * (size_t)p.out_stride + (p0 >> 1);
    for (int i = tid; i < nout; i += NTHR) out[i] = buf[sidx(WARM + 2 * i)];
}

}  // namespace zfb
