// ZFB_MODE_FAST interior: fused NCO mix + multistage polyphase FIR decimator.
//
// Replaces, for the INTERIOR of a chunk, all but the last scipy.signal.
// decimate(x, 2) call of the reference's zoom loop (pypanadapter_spectrum.py:
// 2096-2098, pypanadapter_thread.py:1532-1534).  Every stage of the reference
// is a zero-phase LTI filter |H(f)|^2 (cheby1 order 8, applied forwards and
// backwards) followed by [::2].  Only the band |f| <= B = 0.7 fs/R of the
// early stages' output can reach the row: the last stage -- which stays the
// exact IIR kernel -- rejects everything beyond B by >= 178 dB.  So an early
// stage only has to (a) keep aliases out of |f| <= B and (b) have a smooth,
// known gain there.  Short symmetric FIRs (7..27 taps, designed on the host
// in fp64, pypanadapter_b200/fastdesign.py) do (a) to >= 140 dB, and one
// 33-tap compensator at the last stage's input rate restores the product of
// the reference's |H_s|^2 over |f| <= B to 1e-5.  The per-chunk edge semantics
// of the reference (odd extension, steady-state initial conditions at every
// stage) are NOT LTI; the first and last samples of the decimated chunk are
// therefore recomputed by the exact kernels on two short strips and overwrite
// what this kernel produced there (zfb_engine.cu).
//
// One CTA = one tile of TO outputs.  Levels: 0 = mixed input, s = after s
// decimating stages; each level lives in shared memory in polyphase (even/odd)
// layout so that the stride-2 reads of a decimator are unit-stride across
// threads.  Values outside a level's [0, L_level) are zero (zero extension),
// independent of the tiling.
#pragma once
#include "zfb_decim.cuh"

namespace zfb {

constexpr int FIR_MAX_STAGES = 3;     // decimating stages per launch
constexpr int FIR_MAX_HALF = 20;      // half length M of a stage (2M+1 taps)
constexpr int FIR_COMP_MAX_HALF = 24;   // (dispatch tables below enumerate 1..20 / 1..24)
constexpr int FIR_NT = 256;

struct FirChainParams {
    const void *in;            // [frames][in_stride] samples of the chain's input kind
    long long   in_stride;
    int         L;             // input length per frame
    int         flip;
    unsigned long long phase_inc;
    float2      lo_small[8];   // amp * exp(-2 pi i f/fs v), v = 0..7
    float2      lo_big[32];    // exp(-2 pi i f/fs * it*FIR_NT*VEC)
    int         ns;            // decimating stages, 1..FIR_MAX_STAGES
    int         M[FIR_MAX_STAGES];
    float       h[FIR_MAX_STAGES][FIR_MAX_HALF + 1];   // h[s][j], j = 0..M[s] (centre first)
    int         Mc;            // compensator half length, -1: none
    float       hc[FIR_COMP_MAX_HALF + 1];
    float2     *out;           // [frames][out_stride]
    long long   out_stride;
    int         TO;            // outputs per tile
    int         n[FIR_MAX_STAGES + 1];   // samples held per level (level 0 rounded to the load vector)
};

// geometry of one tile, shared by host (smem sizing) and device
struct FirTile {
    int lo[FIR_MAX_STAGES + 1];   // first position held per level (level 0 aligned down to 8)
    int n[FIR_MAX_STAGES + 1];    // count per level
    int c[FIR_MAX_STAGES + 1];    // c[l]: index in level l-1 of the centre tap of level-l output 0 (l >= 1)
};

__host__ __device__ inline void fir_tile_geometry(const FirChainParams &p, int o0, FirTile &t) {
    const int mc = p.Mc < 0 ? 0 : p.Mc;
    t.lo[p.ns] = o0 - mc;
    t.n[p.ns] = p.TO + 2 * mc;
    for (int l = p.ns; l >= 1; --l) {
        const int M = p.M[l - 1];
        int lo = 2 * t.lo[l] - M;
        int n = 2 * (t.n[l] - 1) + 2 * M + 1;
        if (l == 1) {                      // level 0: align to the 8-sample load vector
            const int lo_al = lo - (((lo % 8) + 8) % 8);
            n += lo - lo_al;
            n = (n + 7) / 8 * 8;
            t.c[1] = M + (lo - lo_al);
            lo = lo_al;
        } else {
            t.c[l] = M;
        }
        t.lo[l - 1] = lo;
        t.n[l - 1] = n;
    }
}

// even/odd arrays of a level with n samples: E at base, O at base + eo_half(n)
__host__ __device__ constexpr int eo_half(int n) { return ((n + 1) / 2 + 3) | 1; }

struct StageIO {
    const float2 *E, *O;     // polyphase input level
    int c;                   // input index of the centre tap of output 0
    int n, i_lo, i_hi;       // outputs held / first and one-past-last non-zero output
    int mode;                // 0: polyphase smem, 1: linear smem, 2: global
    float2 *dstE, *dstO;
};

__device__ __forceinline__ void stage_store(const StageIO &io, int i, float2 v) {
    if (io.mode == 0) ((i & 1) ? io.dstO : io.dstE)[i >> 1] = v;
    else io.dstE[i] = v;
}

// One decimate-by-2 FIR stage with compile-time half length M: the taps sit in
// registers and every tap is an LDS with an immediate offset from two base
// pointers (A: the polyphase array holding the centre tap, B: the other one).
// Output i: centre = 2i + c; even tap offsets stay in A, odd ones are in B.
template <int M>
__device__ __forceinline__ void fir_stage(const float *hp, const StageIO &io, int tid) {
    float h[M + 1];
#pragma unroll
    for (int j = 0; j <= M; ++j) h[j] = hp[j];
    const bool c_odd = io.c & 1;
    const float2 *A = (c_odd ? io.O : io.E) + (io.c >> 1);
    const float2 *B = (c_odd ? io.E + 1 : io.O) + (io.c >> 1);
    for (int i = tid; i < io.n; i += FIR_NT) {
        float2 acc = make_float2(0.f, 0.f);
        if (i >= io.i_lo && i < io.i_hi) {
            const float2 *a = A + i, *b = B + i;
            acc = pk_mul(h[0], a[0]);
#pragma unroll
            for (int j = 1; j <= M; ++j) {
                float2 xs;
                if (j & 1) xs = pk_add(b[(j - 1) / 2], b[-(j + 1) / 2]);     // offsets +-j, j odd
                else xs = pk_add(a[j / 2], a[-j / 2]);
                acc = pk_fma(h[j], xs, acc);
            }
        }
        if (io.mode != 2 || (i >= io.i_lo && i < io.i_hi)) stage_store(io, i, acc);
    }
}

__device__ __forceinline__ void fir_stage_dispatch(int M, const float *hp, const StageIO &io, int tid) {
    // exact half length only: a padded instantiation would read past the halo
#define ZFB_CASE(m) case m: fir_stage<m>(hp, io, tid); break;
    switch (M) {
        ZFB_CASE(1) ZFB_CASE(2) ZFB_CASE(3) ZFB_CASE(4) ZFB_CASE(5) ZFB_CASE(6) ZFB_CASE(7) ZFB_CASE(8)
        ZFB_CASE(9) ZFB_CASE(10) ZFB_CASE(11) ZFB_CASE(12) ZFB_CASE(13) ZFB_CASE(14) ZFB_CASE(15)
        ZFB_CASE(16) ZFB_CASE(17) ZFB_CASE(18) ZFB_CASE(19) ZFB_CASE(20)
        default: break;
    }
#undef ZFB_CASE
}

// symmetric compensator (no decimation): out[i] = sum_j hc[|j|] x[i + j], x linear in smem
template <int MC>
__device__ __forceinline__ void comp_stage(const float *hp, const float2 *x, float2 *out, int n_out, int tid) {
    float h[MC + 1];
#pragma unroll
    for (int j = 0; j <= MC; ++j) h[j] = hp[j];
    for (int i = tid; i < n_out; i += FIR_NT) {
        const float2 *a = x + i;
        float2 acc = pk_mul(h[0], a[0]);
#pragma unroll
        for (int j = 1; j <= MC; ++j) acc = pk_fma(h[j], pk_add(a[j], a[-j]), acc);
        out[i] = acc;
    }
}

__device__ __forceinline__ void comp_dispatch(int Mc, const float *hp, const float2 *x, float2 *out, int n_out,
                                              int tid) {
#define ZFB_CASE(m) case m: comp_stage<m>(hp, x, out, n_out, tid); break;
    switch (Mc) {
        ZFB_CASE(1) ZFB_CASE(2) ZFB_CASE(3) ZFB_CASE(4) ZFB_CASE(5) ZFB_CASE(6) ZFB_CASE(7) ZFB_CASE(8)
        ZFB_CASE(9) ZFB_CASE(10) ZFB_CASE(11) ZFB_CASE(12) ZFB_CASE(13) ZFB_CASE(14) ZFB_CASE(15)
        ZFB_CASE(16) ZFB_CASE(17) ZFB_CASE(18) ZFB_CASE(19) ZFB_CASE(20) ZFB_CASE(21) ZFB_CASE(22)
        ZFB_CASE(23) ZFB_CASE(24)
        default: break;
    }
#undef ZFB_CASE
}

template <int KIND>
__global__ void __launch_bounds__(FIR_NT) fir_chain_kernel(const FirChainParams p) {
    constexpr int VEC = (KIND == KIND_U8_RAW) ? 8 : 2;
    ZFB_DYN_SMEM(smem_raw);
    float2 *sm = reinterpret_cast<float2 *>(smem_raw);
    const int tid = threadIdx.x;
    const int frame = blockIdx.y;
    const int o0 = blockIdx.x * p.TO;

    FirTile t;
    fir_tile_geometry(p, o0, t);
    int Llev[FIR_MAX_STAGES + 1];
    Llev[0] = p.L;
    for (int l = 1; l <= p.ns; ++l) Llev[l] = (Llev[l - 1] + 1) >> 1;

    // buffers: level l (polyphase) alternates between bufA and bufB
    const int sizeA = 2 * eo_half(t.n[0]);
    float2 *bufA = sm;
    float2 *bufB = sm + sizeA;

    // ---------------- level 0: load, convert, flip, LO mix ----------------
    {
        const size_t esz = (KIND == KIND_U8_RAW) ? 2 : 8;
        const char *frame_in = (const char *)p.in + (size_t)frame * (size_t)p.in_stride * esz;
        const int L = p.L;
        const bool fl = (KIND != KIND_C64_MID) && p.flip;
        const int half0 = eo_half(t.n[0]);
        float2 *E = bufA, *O = bufA + half0;
        const int nvec = t.n[0] / VEC;
        float2 b0 = make_float2(1.f, 0.f);
        if (KIND != KIND_C64_MID)
            b0 = lo_phasor((long long)t.lo[0] + (long long)tid * VEC, p.phase_inc);
        for (int it = 0; it * FIR_NT < nvec; ++it) {
            const int v = it * FIR_NT + tid;
            if (v >= nvec) break;
            const int r0 = v * VEC;                 // index within the level (even)
            const int pos = t.lo[0] + r0;
            float2 s[VEC];
            const long long i0 = fl ? (long long)L - VEC - pos : (long long)pos;
            const char *a = frame_in + (size_t)i0 * esz;
            if (pos >= 0 && pos + VEC <= L && ((((uintptr_t)a) & 15) == 0)) {
                const uint4 raw = __ldg((const uint4 *)a);
                if (KIND == KIND_U8_RAW) {
                    const unsigned int wds[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        const int ee = fl ? (VEC - 1 - e) : e;
                        s[e] = u8pair_to_iq(wds[ee >> 1], ee & 1);
                    }
                } else {
                    const float2 lo2 = make_float2(__uint_as_float(raw.x), __uint_as_float(raw.y));
                    const float2 hi2 = make_float2(__uint_as_float(raw.z), __uint_as_float(raw.w));
                    s[0] = fl ? hi2 : lo2;
                    s[VEC - 1] = fl ? lo2 : hi2;
                }
            } else {
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const int pe = pos + e;
                    s[e] = make_float2(0.f, 0.f);
                    if (pe >= 0 && pe < L) {
                        const long long ie = fl ? (long long)L - 1 - pe : (long long)pe;
                        if (KIND == KIND_U8_RAW) {
                            const unsigned char *src = (const unsigned char *)frame_in;
                            s[e] = make_float2(u8_to_f(src[2 * ie]), u8_to_f(src[2 * ie + 1]));
                        } else {
                            s[e] = __ldg((const float2 *)frame_in + ie);
                        }
                    }
                }
            }
            if (KIND != KIND_C64_MID) {
                const float2 bi = cmul(b0, p.lo_big[it]);
#pragma unroll
                for (int e = 0; e < VEC; ++e) s[e] = cmul(s[e], cmul(bi, p.lo_small[e]));
            }
#pragma unroll
            for (int e = 0; e < VEC; e += 2) {
                E[(r0 + e) >> 1] = s[e];
                O[(r0 + e) >> 1] = s[e + 1];
            }
        }
    }
    __syncthreads();

    // ---------------- decimating stages ----------------
    float2 *src = bufA, *dst = bufB;
    float2 *frame_out = p.out + (size_t)frame * (size_t)p.out_stride;
    for (int l = 1; l <= p.ns; ++l) {
        const bool last = (l == p.ns);
        StageIO io;
        io.E = src;
        io.O = src + eo_half(t.n[l - 1]);
        io.c = t.c[l];
        io.n = t.n[l];
        // outputs whose position lies outside the level are zero (zero extension)
        io.i_lo = max(0, -t.lo[l]);
        io.i_hi = min(t.n[l], Llev[l] - t.lo[l]);
        if (last && p.Mc < 0) {
            io.mode = 2;                                  // straight to global
            io.dstE = frame_out + t.lo[l];
            io.dstO = nullptr;
        } else if (last) {
            io.mode = 1;                                  // linear, for the compensator
            io.dstE = dst;
            io.dstO = nullptr;
        } else {
            io.mode = 0;                                  // polyphase, for the next decimator
            io.dstE = dst;
            io.dstO = dst + eo_half(t.n[l]);
        }
        fir_stage_dispatch(p.M[l - 1], p.h[l - 1], io, tid);
        __syncthreads();
        float2 *tmp = src; src = dst; dst = tmp;
    }

    // ---------------- compensator at the output rate ----------------
    if (p.Mc >= 0) {
        const int Lout = Llev[p.ns];
        const int n_out = min(p.TO, Lout - o0);
        comp_dispatch(p.Mc, p.hc, src + p.Mc, frame_out + o0, n_out, tid);
    }
}

// dynamic shared memory of a launch: level 0 in bufA, the largest later level in bufB,
// and (ping-pong) level 2 back in bufA, which always fits
inline size_t fir_chain_smem(const FirChainParams &p) {
    FirTile t;
    fir_tile_geometry(p, 0, t);
    const size_t a = 2 * (size_t)eo_half(t.n[0]);
    size_t b = 0;
    for (int l = 1; l <= p.ns; ++l) {
        const size_t need = (l == p.ns && p.Mc >= 0) ? (size_t)t.n[l] + 8 : 2 * (size_t)eo_half(t.n[l]);
        if (need > b) b = need;
    }
    return (a + b) * sizeof(float2);
}

}  // namespace zfb

// ===========================================================================
// Register-blocked variant of the FIR chain for the tap sets the design
// actually produces for R = 4, 8, >= 16 (last chain): every thread owns a run
// of RUN0 = 32 consecutive level-0 samples, keeps its data in registers through
// all stages (sliding windows with static indexing) and goes to shared memory
// only for the M samples it needs from its neighbours' runs.  ~3x less shared
// memory traffic than fir_chain_kernel (which stays as the general fallback:
// ncu showed it at 83 % of the LSU/shared-memory wavefront peak).
// ===========================================================================
#ifndef ZFB_FIR_MINB
#define ZFB_FIR_MINB 2
#endif

namespace zfb {

constexpr int RUN0 = 32;

struct FirRunParams {
    const void *in;
    long long   in_stride;
    int         L;
    int         flip;
    unsigned long long phase_inc;
    float2      lo_run[RUN0];      // amp * exp(-2 pi i f/fs j), j = 0..31
    float       h0[FIR_MAX_HALF + 1], h1[FIR_MAX_HALF + 1], h2[FIR_MAX_HALF + 1];
    float       hc[FIR_COMP_MAX_HALF + 1];
    // uint8 input with the late mix: the first stage works on the RAW byte values u (the
    // conversion's u/127.5 - 1 is folded into it: taps h0/127.5, and the unit DC gain turns the
    // -1 into one constant per output, `bias0` = 127.5 * sum(h0s) as the fp32 taps give it)
    float       h0s[FIR_MAX_HALF + 1];
    float       bias0;
    float2     *out;
    long long   out_stride;
    int         ht;                // halo threads per side
    const ChannelLo *chan;         // channel-batched launches (template CH), else null
    int         chan_frames;       // frames per channel: batch b = ch*chan_frames + frame
    // Late mix.  Mixing then filtering with h equals filtering with h[n] e^{+i theta n} and
    // mixing afterwards; when the LO offset is negligible against every filter bandwidth of the
    // chain (|f_demod| R / fs <= 1e-3: the chain's gain, smooth over 0.7 fs/R, moves by < 1e-3 dB
    // -- the reference's own LO sits at 1 Hz, S:2090) the modulation of the taps is dropped and
    // the LO is applied to the chain's OUTPUT, 2^NS times fewer complex products.
    int         late;
    float2      lo_out[RUN0 / 2];  // amp * exp(-2 pi i f/fs * (j << NS)), j = 0 .. (RUN0 >> NS) - 1
};

// neighbour exchange of one level: a thread publishes its run (or only the M
// samples at either end when that is all a neighbour can need) and reads the
// M samples left / right of its run with compile-time offsets
template <int RUN, int M, int NT = FIR_NT>
struct LevelStore {
    static constexpr bool EDGES = (2 * M <= RUN);
    static constexpr int STRIDE = (EDGES ? 2 * M : RUN) | 1;       // odd: conflict-free LDS.64
    static constexpr int SIZE = NT * STRIDE;                   // float2 elements

    __device__ static __forceinline__ void publish(float2 *sm, int t, const float2 (&x)[RUN]) {
        float2 *p = sm + t * STRIDE;
        if (EDGES) {
#pragma unroll
            for (int j = 0; j < M; ++j) {
                p[j] = x[j];
                p[M + j] = x[RUN - M + j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < RUN; ++j) p[j] = x[j];
        }
    }
    // win[0..M) = samples at run positions -M..-1, win[M+RUN .. M+RUN+M) = positions RUN..RUN+M-1
    template <int W>
    __device__ static __forceinline__ void halo(const float2 *sm, int t, float2 (&win)[W]) {
        static_assert(W == RUN + 2 * M, "window size");
#pragma unroll
        for (int d = 1; d <= M; ++d) {              // position -d
            int idx;
            if (EDGES) {
                idx = max(t - 1, 0) * STRIDE + 2 * M - d;
            } else {
                const int c = (d + RUN - 1) / RUN;
                idx = max(t - c, 0) * STRIDE + (RUN * c - d);
            }
            win[M - d] = sm[idx];
        }
#pragma unroll
        for (int d = 0; d < M; ++d) {               // position RUN + d
            int idx;
            if (EDGES) {
                idx = min(t + 1, NT - 1) * STRIDE + d;
            } else {
                const int c = d / RUN;
                idx = min(t + 1 + c, NT - 1) * STRIDE + (d - RUN * c);
            }
            win[M + RUN + d] = sm[idx];
        }
    }
};

// RUN/2 outputs of a decimate-by-2 symmetric FIR from a window win[-M .. RUN-1+M]
template <int RUN, int M>
__device__ __forceinline__ void fir_decim_regs(const float2 (&win)[RUN + 2 * M], const float *hp,
                                               float2 (&y)[RUN / 2], float bias = 0.f) {
    float h[M + 1];
#pragma unroll
    for (int j = 0; j <= M; ++j) h[j] = hp[j];
    const float2 nb = make_float2(-bias, -bias);
#pragma unroll
    for (int m = 0; m < RUN / 2; ++m) {
        float2 acc = pk_fma(h[0], win[M + 2 * m], nb);
#pragma unroll
        for (int j = 1; j <= M; ++j) acc = pk_fma(h[j], pk_add(win[M + 2 * m - j], win[M + 2 * m + j]), acc);
        y[m] = acc;
    }
}

template <int RUN, int M>
__device__ __forceinline__ void fir_same_regs(const float2 (&win)[RUN + 2 * M], const float *hp,
                                              float2 (&y)[RUN]) {
    float h[M + 1];
#pragma unroll
    for (int j = 0; j <= M; ++j) h[j] = hp[j];
#pragma unroll
    for (int m = 0; m < RUN; ++m) {
        float2 acc = pk_mul(h[0], win[M + m]);
#pragma unroll
        for (int j = 1; j <= M; ++j) acc = pk_fma(h[j], pk_add(win[M + m - j], win[M + m + j]), acc);
        y[m] = acc;
    }
}

// zero the outputs whose level position lies outside [0, Llev)
template <int RUN>
__device__ __forceinline__ void mask_level(float2 (&y)[RUN], int pos0, int Llev) {
#pragma unroll
    for (int j = 0; j < RUN; ++j)
        if (pos0 + j < 0 || pos0 + j >= Llev) y[j] = make_float2(0.f, 0.f);
}

// one stage: publish the input runs, barrier, gather the halos, filter
// `edge` (CTA-uniform): the tile reaches outside [0, L) at this level; interior tiles (all but the
// first and last of a frame) skip the per-output range tests (ncu: 64 ISETP + 60 FSEL of 1 025
// instructions per run)
template <int RUN, int M, int NT = FIR_NT>
__device__ __forceinline__ void run_stage(float2 *sm, int t, const float2 (&x)[RUN], const float *hp,
                                          float2 (&y)[RUN / 2], int pos0_out, int L_out, bool edge, float bias = 0.f) {
    LevelStore<RUN, M, NT>::publish(sm, t, x);
    __syncthreads();
    float2 win[RUN + 2 * M];
#pragma unroll
    for (int j = 0; j < RUN; ++j) win[M + j] = x[j];
    LevelStore<RUN, M, NT>::template halo<RUN + 2 * M>(sm, t, win);
    fir_decim_regs<RUN, M>(win, hp, y, bias);
    if (edge) mask_level<RUN / 2>(y, pos0_out, L_out);
}

template <int NS, int M0, int M1, int M2, int MC, int NT = FIR_NT>
struct FirRunShape {
    static constexpr int RUN_OUT = RUN0 >> NS;
    static constexpr int S0 = LevelStore<RUN0, M0, NT>::SIZE;
    static constexpr int S1 = NS >= 2 ? LevelStore<RUN0 / 2, (M1 > 0 ? M1 : 1), NT>::SIZE : 0;
    static constexpr int S2 = NS >= 3 ? LevelStore<RUN0 / 4, (M2 > 0 ? M2 : 1), NT>::SIZE : 0;
    static constexpr int SC = MC >= 0 ? LevelStore<RUN_OUT, (MC > 0 ? MC : 1), NT>::SIZE : 0;   // MC < 0: no compensator
    static constexpr size_t SMEM = (size_t)(S0 + S1 + S2 + SC) * sizeof(float2);
    // dependency cone of a final output, in level-0 samples
    static constexpr int CONE = M0 + (NS >= 2 ? 2 * M1 : 0) + (NS >= 3 ? 4 * M2 : 0) + (MC > 0 ? (1 << NS) * MC : 0);
    static constexpr int HT = (CONE + RUN0 - 1) / RUN0;
    static constexpr int SPAN = (NT - 2 * HT) * RUN0;          // level-0 samples a tile finishes
};

template <int KIND, int NS, int M0, int M1, int M2, int MC, bool CH = false, int NT = FIR_NT>
__global__ void __launch_bounds__(NT, (ZFB_FIR_MINB * FIR_NT / NT)) fir_run_kernel(const FirRunParams p) {
    using SH = FirRunShape<NS, M0, M1, M2, MC, NT>;
    constexpr int VEC = (KIND == KIND_U8_RAW) ? 8 : (KIND == KIND_CS16_RAW) ? 4 : 2;    // samples per 16 bytes
    ZFB_DYN_SMEM(smem_raw);
    float2 *sm0 = reinterpret_cast<float2 *>(smem_raw);
    float2 *sm1 = sm0 + SH::S0;
    float2 *sm2 = sm1 + SH::S1;
    float2 *smc = sm2 + SH::S2;
    const int t = threadIdx.x;
    const int frame = blockIdx.y;                                   // output (batch) index
    const int ch = CH ? frame / p.chan_frames : 0;
    const int in_frame = CH ? frame % p.chan_frames : frame;
    const int lo0 = blockIdx.x * SH::SPAN - SH::HT * RUN0;          // level-0 position of thread 0, sample 0
    const int pos0 = lo0 + t * RUN0;
    const int L = p.L;

    // ---------------- level 0: this thread's 32 samples ----------------
    float2 x0[RUN0];
    const bool late_mix = KIND != KIND_C64_MID && (CH ? p.chan[ch].late : p.late);
    // integer wire formats with the late mix: the raw sample values go through the first stage
    // (its taps carry 1/127.5 or 1/32768, `bias0` the uint8 offset)
    const bool fold = (KIND == KIND_U8_RAW || KIND == KIND_CS16_RAW) && late_mix;
    constexpr float kCs16 = 1.0f / 32768.0f;
    {
        const size_t esz = (KIND == KIND_U8_RAW) ? 2 : (KIND == KIND_CS16_RAW) ? 4 : 8;
        const char *frame_in = (const char *)p.in + (size_t)in_frame * (size_t)p.in_stride * esz;
        const bool fl = (KIND != KIND_C64_MID) && p.flip;
        // element e of the run <-> sample index fl ? L-1-(pos0+e) : pos0+e
        const long long i_first = fl ? (long long)L - RUN0 - pos0 : (long long)pos0;
        const char *a = frame_in + (size_t)i_first * esz;
        if (pos0 >= 0 && pos0 + RUN0 <= L && ((((uintptr_t)a) & 15) == 0)) {
            constexpr int NV = RUN0 / VEC;
            uint4 raw[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) raw[v] = __ldg((const uint4 *)a + v);
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                if (KIND == KIND_U8_RAW) {
                    const unsigned int wds[4] = {raw[v].x, raw[v].y, raw[v].z, raw[v].w};
                    if (fold) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int idx = v * 8 + e;
                            x0[fl ? RUN0 - 1 - idx : idx] = u8pair_to_raw(wds[e >> 1], e & 1);
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int idx = v * 8 + e;             // memory order
                            x0[fl ? RUN0 - 1 - idx : idx] = u8pair_to_iq(wds[e >> 1], e & 1);
                        }
                    }
                } else if (KIND == KIND_CS16_RAW) {
                    const unsigned int wds[4] = {raw[v].x, raw[v].y, raw[v].z, raw[v].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int idx = v * 4 + e;
                        const float2 r = cs16pair_to_raw(wds[e]);
                        x0[fl ? RUN0 - 1 - idx : idx] = fold ? r : pk_mul(kCs16, r);
                    }
                } else {
                    const int i0 = v * 2, i1 = v * 2 + 1;
                    x0[fl ? RUN0 - 1 - i0 : i0] = make_float2(__uint_as_float(raw[v].x), __uint_as_float(raw[v].y));
                    x0[fl ? RUN0 - 1 - i1 : i1] = make_float2(__uint_as_float(raw[v].z), __uint_as_float(raw[v].w));
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < RUN0; ++e) {
                const int pe = pos0 + e;
                x0[e] = (fold && KIND == KIND_U8_RAW) ? make_float2(127.5f, 127.5f) : make_float2(0.f, 0.f);   // zero signal
                if (pe >= 0 && pe < L) {
                    const long long ie = fl ? (long long)L - 1 - pe : (long long)pe;
                    if (KIND == KIND_U8_RAW) {
                        const unsigned char *src = (const unsigned char *)frame_in;
                        x0[e] = fold ? make_float2((float)src[2 * ie], (float)src[2 * ie + 1])
                                     : make_float2(u8_to_f(src[2 * ie]), u8_to_f(src[2 * ie + 1]));
                    } else if (KIND == KIND_CS16_RAW) {
                        const short *src = (const short *)frame_in;
                        const float2 r = make_float2((float)src[2 * ie], (float)src[2 * ie + 1]);
                        x0[e] = fold ? r : pk_mul(kCs16, r);
                    } else {
                        x0[e] = __ldg((const float2 *)frame_in + ie);
                    }
                }
            }
        }
        if (KIND != KIND_C64_MID && !late_mix) {
            const ChannelLo *cl = CH ? p.chan + ch : nullptr;
            const float2 b0 = lo_phasor((long long)pos0, CH ? cl->phase_inc : p.phase_inc);
#pragma unroll
            for (int e = 0; e < RUN0; ++e) x0[e] = cmul(x0[e], cmul(b0, CH ? cl->run[e] : p.lo_run[e]));
        }
    }

    int Llev[4];
    Llev[0] = L;
    Llev[1] = (L + 1) >> 1;
    Llev[2] = (Llev[1] + 1) >> 1;
    Llev[3] = (Llev[2] + 1) >> 1;
    // does any thread of this tile hold a position outside the frame (at level 0, hence at any level)?
    const bool edge = lo0 < 0 || lo0 + NT * RUN0 > ((L >> NS) << NS);

    // ---------------- decimating stages, all in registers ----------------
    float2 yf[SH::RUN_OUT];                 // level NS run of this thread
    {
        float2 y1[RUN0 / 2];
        run_stage<RUN0, M0, NT>(sm0, t, x0, fold ? p.h0s : p.h0, y1, pos0 >> 1, Llev[1], edge, fold ? p.bias0 : 0.f);
        if constexpr (NS == 1) {
#pragma unroll
            for (int j = 0; j < RUN0 / 2; ++j) yf[j] = y1[j];
        } else {
            float2 y2[RUN0 / 4];
            run_stage<RUN0 / 2, M1, NT>(sm1, t, y1, p.h1, y2, pos0 >> 2, Llev[2], edge);
            if constexpr (NS == 2) {
#pragma unroll
                for (int j = 0; j < RUN0 / 4; ++j) yf[j] = y2[j];
            } else {
                run_stage<RUN0 / 4, M2, NT>(sm2, t, y2, p.h2, yf, pos0 >> 3, Llev[3], edge);
            }
        }
    }

    // ---------------- compensator (last chain only) ----------------
    constexpr int RO = SH::RUN_OUT;
    float2 out[RO];
    if constexpr (MC > 0) {
        LevelStore<RO, MC, NT>::publish(smc, t, yf);
        __syncthreads();
        float2 win[RO + 2 * MC];
#pragma unroll
        for (int j = 0; j < RO; ++j) win[MC + j] = yf[j];
        LevelStore<RO, MC, NT>::template halo<RO + 2 * MC>(smc, t, win);
        fir_same_regs<RO, MC>(win, p.hc, out);
    } else {
        (void)smc;
#pragma unroll
        for (int j = 0; j < RO; ++j) out[j] = yf[j];
    }

    if (late_mix) {
        // output j of this thread sits at input position (pos0 >> NS << NS) + (j << NS)
        const ChannelLo *cl = CH ? p.chan + ch : nullptr;
        const float2 b0 = lo_phasor((long long)(pos0 >> NS) << NS, CH ? cl->phase_inc : p.phase_inc);
#pragma unroll
        for (int j = 0; j < RO; ++j) out[j] = cmul(out[j], cmul(b0, CH ? cl->out[j] : p.lo_out[j]));
    }

    if (t >= SH::HT && t < NT - SH::HT) {
        const int po = pos0 >> NS;
        float2 *frame_out = p.out + (size_t)frame * (size_t)p.out_stride;
#pragma unroll
        for (int j = 0; j < RO; ++j)
            if (po + j >= 0 && po + j < Llev[NS]) frame_out[po + j] = out[j];
    }
}

}  // namespace zfb
