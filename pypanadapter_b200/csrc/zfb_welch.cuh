// Welch PSD of (decimated) chunks, fused with detrend, window, shared-memory
// Stockham FFT, |X|^2, segment mean, density scale, fftshift and centre crop;
// plus the row finalisation (EMA on linear power, 20*log10).
//
// Replaces scipy.signal.welch(chunk, fs, window=taper, nperseg=N, nfft=N)
// (pypanadapter_spectrum.py:2111, pypanadapter_thread.py:1536/1538 ->
// scipy/signal/_spectral_py.py:515-972), np.fft.fftshift(spec)[N//2-W//2:
// N//2+W//2] (S:2114, T:1543) and 20*np.log10(abs(spec)) (S:2117-2119, T:1548).
//
// FFT: N = 2^m points, N/PPT threads, PPT (8 or 16) points per thread.
// Radix-8 passes (+ radix-4 passes when m % 3 != 0); butterflies are done in
// registers, passes exchange data through shared memory in Stockham autosort
// order with a +1/16 padding.  The first pass is fed straight from global
// memory and the last pass accumulates |X|^2 in registers, so a segment
// touches smem only between passes.  With 50 % overlap the second half of a
// segment is the first half of the next one *in the same thread's registers*
// (element tid + m*N/PPT moves from m to m - PPT/2), so every sample is read
// from global memory exactly once.  Twiddles come from an fp64-generated
// table exp(-2 pi i k / N).
#pragma once
#include "zfb_common.cuh"

namespace zfb {

constexpr int WF_SPARSE_MAX = 12;

struct WelchParams {
    const void  *in;          // [frames][in_stride], kind per template
    long long    in_stride;
    int          len;         // samples per frame available to Welch
    int          flip;        // only for raw kinds (R == 1 path)
    int          nperseg;     // window length (== N unless the chunk is short)
    int          hop;
    int          nseg;
    int          seg_per_split;   // segments handled by one CTA
    int          nsplit;          // CTAs per frame (gridDim.x)
    int          reuse;           // hop == N/2 && nperseg == N: register reuse
    int          prepared;        // 1: blocks are already detrended + windowed (radix-16 front pass)
    const float *window;      // nperseg taps
    const float2*twiddle;     // N entries exp(-2 pi i k / N)
    int          W;           // row width
    float       *pow_out;     // [frames][nsplit][W] sum_s |X|^2, fftshifted + cropped
    // prepared blocks of the radix-16 front pass (frame index = 16 * frame + residue r): the
    // segment mean is removed after the FFT, X -= mean[frame][s] * wf16[r][k]; null: nothing to remove
    const float2*seg_mean;    // [frames / 16][nseg]
    const float2*wf16;        // [16][N]: FFT(window)[16 k + r]
    // cosine-sum windows (boxcar, hann, hamming, blackman, nuttall, flattop ...) have a handful of
    // non-zero FFT bins: then only those are corrected (wf_n > 0) and wf16 is not read at all
    int          wf_n;        // 0: dense table
    const int   *wf_bin;      // [wf_n] bin index 16 k + r of the big FFT (device memory: indexed in a loop)
    const float2*wf_val;      // [wf_n]
};

__device__ __forceinline__ int fpad(int a) { return a + (a >> 4); }

// complex add / subtract as ONE packed instruction (FADD2; the negation is an operand modifier):
// same IEEE result per component as two FADDs, half the issue slots -- the butterflies are 56 %
// additions (welch_kernel<12,16>: 480 FADD of 1640 SASS instructions became 208 FADD2 + 64 FADD)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// multiply by -i (forward FFT quarter turn)
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

__device__ __forceinline__ void bfly2(float2 &a, float2 &b) {
    float2 t = csub(a, b);
    a = cadd(a, b);
    b = t;
}

// in-place 4-point DFT, natural order output
__device__ __forceinline__ void dft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
    bfly2(a0, a2);
    bfly2(a1, a3);
    a3 = mul_mi(a3);
    bfly2(a0, a1);
    bfly2(a2, a3);
    // now a0=X0, a1=X2, a2=X1, a3=X3
    float2 t = a1; a1 = a2; a2 = t;
}

// in-place 8-point DFT, natural order output
__device__ __forceinline__ void dft8(float2 *v) {
    const float h = 0.70710678118654752440f;
    bfly2(v[0], v[4]); bfly2(v[1], v[5]); bfly2(v[2], v[6]); bfly2(v[3], v[7]);
    v[5] = make_float2((v[5].x + v[5].y) * h, (v[5].y - v[5].x) * h);    // * exp(-i pi/4)
    v[6] = mul_mi(v[6]);                                                  // * exp(-i pi/2)
    v[7] = make_float2((v[7].y - v[7].x) * h, -(v[7].x + v[7].y) * h);   // * exp(-3i pi/4)
    bfly2(v[0], v[2]); bfly2(v[1], v[3]); bfly2(v[4], v[6]); bfly2(v[5], v[7]);
    v[3] = mul_mi(v[3]);
    v[7] = mul_mi(v[7]);
    bfly2(v[0], v[1]); bfly2(v[2], v[3]); bfly2(v[4], v[5]); bfly2(v[6], v[7]);
    // bit-reversed -> natural: (0,4,2,6,1,5,3,7)
    float2 t;
    t = v[1]; v[1] = v[4]; v[4] = t;
    t = v[3]; v[3] = v[6]; v[6] = t;
}

// 16-point forward DFT, natural order in and out: X[r] = sum_q v[q] W16^(q r)
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
    // q = 4 q1 + q0, r = r1 + 4 r0:  4-point DFTs over q1, twiddle W16^(q0 r1), 4-point DFTs over q0
    const float c1 = 0.92387953251128673848f, s1 = 0.38268343236508978178f;   // cos, sin(pi/8)
    const float h = 0.70710678118654752440f;
    float2 t[4][4];                       // t[q0][r1]
#pragma unroll
    for (int q0 = 0; q0 < 4; ++q0) {
        float2 a0 = v[q0], a1 = v[4 + q0], a2 = v[8 + q0], a3 = v[12 + q0];
        dft4(a0, a1, a2, a3);
        t[q0][0] = a0; t[q0][1] = a1; t[q0][2] = a2; t[q0][3] = a3;
    }
    // W16^m = exp(-2 pi i m / 16), m = q0 * r1
    const float2 w1 = make_float2(c1, -s1), w2 = make_float2(h, -h), w3 = make_float2(s1, -c1);
    const float2 w6 = make_float2(-h, -h), w9 = make_float2(-c1, s1);
    t[1][1] = cmul(t[1][1], w1); t[1][2] = cmul(t[1][2], w2); t[1][3] = cmul(t[1][3], w3);
    t[2][1] = cmul(t[2][1], w2); t[2][2] = mul_mi(t[2][2]);   t[2][3] = cmul(t[2][3], w6);
    t[3][1] = cmul(t[3][1], w3); t[3][2] = cmul(t[3][2], w6); t[3][3] = cmul(t[3][3], w9);
#pragma unroll
    for (int r1 = 0; r1 < 4; ++r1) {
        float2 a0 = t[0][r1], a1 = t[1][r1], a2 = t[2][r1], a3 = t[3][r1];
        dft4(a0, a1, a2, a3);
        v[r1] = a0; v[r1 + 4] = a1; v[r1 + 8] = a2; v[r1 + 12] = a3;
    }
}

template <int RADIX> struct Log2R;
template <> struct Log2R<2> { static constexpr int v = 1; };
template <> struct Log2R<4> { static constexpr int v = 2; };
template <> struct Log2R<8> { static constexpr int v = 3; };
template <> struct Log2R<16> { static constexpr int v = 4; };

// One Stockham pass of radix RADIX over the PPT values a thread holds.
// Thread `tid` of nt = N/PPT holds v[m] = data[tid + m*nt].  A thread does
// NB = PPT/RADIX butterflies j = tid + c*nt whose inputs j + q*N/RADIX are
// exactly its own v[c + NB*q].  Ns = product of the radices done so far.
// Threads with tid >= nt (only when N/PPT < 32) idle but keep the barriers.
// `tw` is the table exp(-2 pi i k / 2^log2tw) (log2tw may exceed log2 of this
// FFT's length: sub-FFTs of the four-step path share the big table).
template <int RADIX, int PPT>
__device__ __forceinline__ void fft_pass(float2 (&v)[PPT], int tid, int nt, int log2tw, int log2Ns,
                                         const float2 *__restrict__ tw, float2 *sm, bool last,
                                         bool active) {
    constexpr int NB = PPT / RADIX;               // butterflies per thread
    constexpr int LR = Log2R<RADIX>::v;
    const int Ns = 1 << log2Ns;
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        const int j = tid + c * nt;
        const int k = j & (Ns - 1);
        float2 a[RADIX];
#pragma unroll
        for (int r = 0; r < RADIX; ++r) a[r] = v[c + NB * r];
        if (log2Ns > 0 && active) {
            // twiddle w^r, w = exp(-2 pi i k / (Ns*RADIX)) = tw[k * 2^log2tw/(Ns*RADIX)].
            // The table lookups of a warp are scattered (k differs per lane): ncu showed
            // the LSU/L1 pipe at 74 % with one lookup per r, so only w, w^2, w^4 are
            // loaded and the other powers are one complex product each.
            const int step = k << (log2tw - log2Ns - LR);
            const float2 w1 = __ldg(tw + step);
            a[1] = cmul(a[1], w1);
            if constexpr (RADIX >= 4) {
                const float2 w2 = __ldg(tw + 2 * step);
                const float2 w3 = cmul(w1, w2);
                a[2] = cmul(a[2], w2);
                a[3] = cmul(a[3], w3);
                if constexpr (RADIX >= 8) {
                    const float2 w4 = __ldg(tw + 4 * step);
                    const float2 w5 = cmul(w1, w4), w6 = cmul(w2, w4), w7 = cmul(w3, w4);
                    a[4] = cmul(a[4], w4);
                    a[5] = cmul(a[5], w5);
                    a[6] = cmul(a[6], w6);
                    a[7] = cmul(a[7], w7);
                    if constexpr (RADIX == 16) {
                        const float2 w8 = __ldg(tw + 8 * step);
                        a[8] = cmul(a[8], w8);
                        a[9] = cmul(a[9], cmul(w1, w8));
                        a[10] = cmul(a[10], cmul(w2, w8));
                        a[11] = cmul(a[11], cmul(w3, w8));
                        a[12] = cmul(a[12], cmul(w4, w8));
                        a[13] = cmul(a[13], cmul(w5, w8));
                        a[14] = cmul(a[14], cmul(w6, w8));
                        a[15] = cmul(a[15], cmul(w7, w8));
                    }
                }
            }
        }
        if constexpr (RADIX == 16) dft16(a);
        else if constexpr (RADIX == 8) dft8(a);
        else if constexpr (RADIX == 4) dft4(a[0], a[1], a[2], a[3]);
        else bfly2(a[0], a[1]);
        if (last) {
#pragma unroll
            for (int r = 0; r < RADIX; ++r) v[c + NB * r] = a[r];
        } else if (active) {
            const int j0 = ((j >> log2Ns) << (log2Ns + LR)) + k;
#pragma unroll
            for (int r = 0; r < RADIX; ++r) sm[fpad(j0 + r * Ns)] = a[r];
        }
    }
    if (!last) {
        __syncthreads();
        if (active) {
#pragma unroll
            for (int m = 0; m < PPT; ++m) v[m] = sm[fpad(tid + m * nt)];
        }
        __syncthreads();
    }
}

// full N-point forward FFT of the thread-distributed array; on return
// v[m] = X[tid + m*N/PPT] (natural order)
template <int LOG2N, int PPT>
__device__ __forceinline__ void fft_block(float2 (&v)[PPT], int tid, const float2 *tw, float2 *sm) {
    constexpr int NT = (1 << LOG2N) / PPT;
    const bool active = tid < NT;
    int log2Ns = 0;
#ifndef ZFB_WELCH_NO_R16
    if constexpr (PPT == 16) {
        // 16 points per thread = one radix-16 butterfly per pass: 2048 = 16.16.8,
        // 4096 = 16.16.16, 8192 = 16.16.8.4 -- one exchange through shared memory
        // (and two barriers) fewer than the radix-8 schedule, 4 table lookups per
        // thread and pass instead of 6
        static_assert(LOG2N >= 11 && LOG2N <= 13, "radix-16 schedule covers 2048..8192");
        fft_pass<16, PPT>(v, tid, NT, LOG2N, 0, tw, sm, false, active);
        fft_pass<16, PPT>(v, tid, NT, LOG2N, 4, tw, sm, false, active);
        if constexpr (LOG2N == 11) {
            fft_pass<8, PPT>(v, tid, NT, LOG2N, 8, tw, sm, true, active);
        } else if constexpr (LOG2N == 12) {
            fft_pass<16, PPT>(v, tid, NT, LOG2N, 8, tw, sm, true, active);
        } else {
            fft_pass<8, PPT>(v, tid, NT, LOG2N, 8, tw, sm, false, active);
            fft_pass<4, PPT>(v, tid, NT, LOG2N, 11, tw, sm, true, active);
        }
        return;
    }
#endif
    constexpr int R8 = (LOG2N % 3 == 1) ? (LOG2N / 3 - 1) : (LOG2N / 3);   // radix-8 passes
    constexpr int REM = LOG2N - 3 * R8;                                   // 0, 2 or 4 bits
#pragma unroll
    for (int p = 0; p < R8; ++p) {
        const bool last = (REM == 0) && (p == R8 - 1);
        fft_pass<8, PPT>(v, tid, NT, LOG2N, log2Ns, tw, sm, last, active);
        log2Ns += 3;
    }
    if (REM == 2) {
        fft_pass<4, PPT>(v, tid, NT, LOG2N, log2Ns, tw, sm, true, active);
    } else if (REM == 4) {
        fft_pass<4, PPT>(v, tid, NT, LOG2N, log2Ns, tw, sm, false, active);
        log2Ns += 2;
        fft_pass<4, PPT>(v, tid, NT, LOG2N, log2Ns, tw, sm, true, active);
    }
}

template <int KIND>
__device__ __forceinline__ float2 welch_fetch(const void *frame_in, int idx, int len, int flip) {
    if (KIND == KIND_U8_RAW) {
        const unsigned char *s = (const unsigned char *)frame_in;
        const int i = flip ? (len - 1 - idx) : idx;
        const unsigned short h = __ldg((const unsigned short *)(s + 2 * (size_t)i));
        return make_float2(u8_to_f(h & 0xffu), u8_to_f(h >> 8));
    } else {
        const float2 *s = (const float2 *)frame_in;
        const int i = (KIND == KIND_C64_RAW && flip) ? (len - 1 - idx) : idx;
        return __ldg(s + i);
    }
}

// sum a complex value over the CTA (all threads get the result)
template <int NTHREADS>
__device__ __forceinline__ float2 block_sum(float2 sum, int tid, float2 *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum.x += __shfl_xor_sync(0xffffffffu, sum.x, o);
        sum.y += __shfl_xor_sync(0xffffffffu, sum.y, o);
    }
    if (NTHREADS > 32) {
        if ((tid & 31) == 0) red[tid >> 5] = sum;
        __syncthreads();
        if (tid < 32) {
            float2 t = (tid < NTHREADS / 32) ? red[tid] : make_float2(0.f, 0.f);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                t.x += __shfl_xor_sync(0xffffffffu, t.x, o);
                t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
            }
            if (tid == 0) red[32] = t;
        }
        __syncthreads();
        sum = red[32];
    }
    return sum;
}

template <int LOG2N, int PPT>
struct WelchShape {
    static constexpr int N = 1 << LOG2N;
    static constexpr int NT = N / PPT;                       // working threads
    static constexpr int NTHREADS = NT < 32 ? 32 : NT;       // launched threads
#ifndef ZFB_WELCH_MINB_BIG
#define ZFB_WELCH_MINB_BIG 1
#endif
    static constexpr int MINB = NTHREADS <= 256 ? 2 : ZFB_WELCH_MINB_BIG;     // >= 2 CTAs/SM: the passes are barrier-bound
    static constexpr size_t SMEM = (size_t)(N + (N >> 4) + 1) * sizeof(float2);
};

// KEEP: how many of a thread's PPT bins (k = tid + m*NT) at EACH end of the spectrum can fall inside
// the kept W columns: m < KEEP or m >= PPT - KEEP.  The reference crops the decimated spectrum to its
// centre N/R bins (S:2114, T:1543), so for R >= 2 most of the last pass's outputs are never looked
// at: with KEEP < PPT/2 only the kept ones are accumulated (2*KEEP instead of PPT registers) and the
// compiler prunes the last butterfly to the outputs that are used.  KEEP = PPT/2: everything.
// DENSE (needs KEEP <= 2, <= 256 threads): 768 threads per SM (3 CTAs of 256, 6 of 128) at 80 registers.  What makes room is the pruned
// accumulator and NOT carrying the overlapping half of a segment in registers (16 of them): it is
// fetched again (an L1/L2 hit: the same CTA read it one segment ago).
template <int LOG2N, int PPT, int KIND, int KEEP = PPT / 2, bool DENSE = false>
__global__ void __launch_bounds__((WelchShape<LOG2N, PPT>::NTHREADS), (DENSE ? 768 / WelchShape<LOG2N, PPT>::NTHREADS : WelchShape<LOG2N, PPT>::MINB))
welch_kernel(const WelchParams p) {
    using S = WelchShape<LOG2N, PPT>;
    constexpr int N = S::N;
    constexpr int NT = S::NT;
    constexpr int H = PPT / 2;
    static_assert(KEEP >= 1 && KEEP <= H, "KEEP counts bins per spectrum end");
    static_assert(!DENSE || (KEEP <= 2 && S::NTHREADS <= 256), "DENSE: pruned accumulator, <= 256 threads");
#define ZFB_KEPT(m) ((m) < KEEP || (m) >= PPT - KEEP)
    ZFB_DYN_SMEM(smem_raw);
    float2 *sm = reinterpret_cast<float2 *>(smem_raw);     // fpad(N) complex
    __shared__ float2 red[33];

    const int tid = threadIdx.x;
    const bool active = (NT == S::NTHREADS) ? true : (tid < NT);
    const int split = blockIdx.x;
    const int frame = blockIdx.y;
    const size_t esz = (KIND == KIND_U8_RAW) ? 2 : 8;
    const bool full = p.nperseg == N;       // the usual case: no bounds test per element
    const char *frame_in = (const char *)p.in + (size_t)frame * (size_t)p.in_stride * esz;

    float acc[PPT];
#pragma unroll
    for (int m = 0; m < PPT; ++m) acc[m] = 0.f;
    const float inv_n = 1.0f / (float)p.nperseg;

    const int s_begin = split * p.seg_per_split;
    const int s_end = min(p.nseg, s_begin + p.seg_per_split);
    float2 raw[PPT];

    // p.reuse (hop == N/2, nperseg == N; every BASELINE configuration): the second half of a segment
    // is the first half of the next in the same thread's registers, and no element needs a bounds
    // test -- a uniform branch picks that load sequence (H shifts + H unconditional loads)
    // sparse post-FFT mean removal (prepared blocks): the listed bins have distinct residues (host
    // guarantee), so a CTA holds at most one of them and one thread owns it for all segments
    int sparse_m = -1;                 // sparse_m + PPT * (index in the list), or -1
    if (p.prepared && p.seg_mean != nullptr && p.wf_n > 0) {
        for (int j = 0; j < p.wf_n; ++j) {
            const int b = __ldg(p.wf_bin + j);
            if ((b & 15) == (frame & 15) && ((b >> 4) & (NT - 1)) == tid) sparse_m = (b >> 4) / NT + PPT * j;
        }
    }
    const bool fastload = p.reuse && !p.prepared;
    constexpr bool CARRY = !DENSE;
    if (CARRY && fastload && s_begin < s_end) {           // first half of the first segment, parked in the upper half
#pragma unroll
        for (int m = 0; m < H; ++m)
            raw[m + H] = active ? welch_fetch<KIND>(frame_in, s_begin * p.hop + tid + m * NT, p.len, p.flip)
                                : make_float2(0.f, 0.f);
    }
    for (int s = s_begin; s < s_end; ++s) {
        const int base = s * p.hop;
        if (fastload && !CARRY) {
#pragma unroll
            for (int m = 0; m < PPT; ++m)
                raw[m] = active ? welch_fetch<KIND>(frame_in, base + tid + m * NT, p.len, p.flip) : make_float2(0.f, 0.f);
        } else if (fastload) {
#pragma unroll
            for (int m = 0; m < H; ++m) raw[m] = raw[m + H];
#pragma unroll
            for (int m = 0; m < H; ++m)
                raw[m + H] = active ? welch_fetch<KIND>(frame_in, base + tid + (m + H) * NT, p.len, p.flip)
                                    : make_float2(0.f, 0.f);
        } else {
#pragma unroll
            for (int m = 0; m < PPT; ++m) {
                const int idx = tid + m * NT;
                raw[m] = (active && idx < p.nperseg) ? welch_fetch<KIND>(frame_in, base + idx, p.len, p.flip)
                                                     : make_float2(0.f, 0.f);
            }
        }
        float2 v[PPT];
        if (p.prepared) {
#pragma unroll
            for (int m = 0; m < PPT; ++m) v[m] = raw[m];
        } else {
            float2 sum = make_float2(0.f, 0.f);
#pragma unroll
            for (int m = 0; m < PPT; ++m) sum = cadd(sum, raw[m]);
            // detrend='constant': subtract the segment's complex mean
            sum = block_sum<S::NTHREADS>(sum, tid, red);
            const float2 mean = make_float2(sum.x * inv_n, sum.y * inv_n);
            float w[PPT];
            if (full) {
#pragma unroll
                for (int m = 0; m < PPT; ++m) w[m] = active ? __ldg(p.window + tid + m * NT) : 0.f;
            } else {
#pragma unroll
                for (int m = 0; m < PPT; ++m) {
                    const int idx = tid + m * NT;
                    w[m] = (active && idx < p.nperseg) ? __ldg(p.window + idx) : 0.f;
                }
            }
#pragma unroll
            for (int m = 0; m < PPT; ++m) {
                v[m].x = (raw[m].x - mean.x) * w[m];
                v[m].y = (raw[m].y - mean.y) * w[m];
            }
        }
        fft_block<LOG2N, PPT>(v, tid, p.twiddle, sm);
        if (p.prepared && p.seg_mean != nullptr && active) {
            if (p.wf_n > 0) {
                if (sparse_m >= 0) {                       // the one thread of the CTA that owns a listed bin
                    const float2 mean = __ldg(p.seg_mean + (size_t)(frame >> 4) * p.nseg + s);
                    const float2 c = cmul(make_float2(-mean.x, -mean.y), __ldg(p.wf_val + sparse_m / PPT));
                    const int own = sparse_m % PPT;
#pragma unroll
                    for (int m = 0; m < PPT; ++m) {       // blend, not v[own]: the array stays in registers
                        const float f = (m == own) ? 1.f : 0.f;
                        v[m].x = fmaf(f, c.x, v[m].x);
                        v[m].y = fmaf(f, c.y, v[m].y);
                    }
                }
            } else {
                const float2 mean = __ldg(p.seg_mean + (size_t)(frame >> 4) * p.nseg + s);
                const float2 nm = make_float2(-mean.x, -mean.y);
                const float2 *wf = p.wf16 + (size_t)(frame & 15) * N + tid;
#pragma unroll
                for (int m = 0; m < PPT; ++m) v[m] = cadd(v[m], cmul(nm, __ldg(wf + m * NT)));
            }
        }
#pragma unroll
        for (int m = 0; m < PPT; ++m)
            if (ZFB_KEPT(m)) acc[m] = fmaf(v[m].x, v[m].x, fmaf(v[m].y, v[m].y, acc[m]));
    }

    // fftshift + centre crop: natural bin k sits at column (k + N/2) mod N
    if (!active) return;
    const int c0 = N / 2 - p.W / 2;
    float *row = p.pow_out + ((size_t)frame * p.nsplit + split) * p.W;
#pragma unroll
    for (int m = 0; m < PPT; ++m) {
        if (!ZFB_KEPT(m)) continue;
        const int k = tid + m * NT;
        const int col = ((k + N / 2) & (N - 1)) - c0;
        if (col >= 0 && col < p.W) row[col] = acc[m];
    }
#undef ZFB_KEPT
}

// rows: sum the per-split partial sums in fixed order, scale, optional EMA
// across consecutive frames on linear power, then dB20.
//   a_0 = p_0 (when the state is empty), a_i = alpha p_i + (1-alpha) a_(i-1)
struct FinalizeParams {
    float *pow_io;            // [nframes][nsplit][Wp] partial sums from the Welch kernels
    int    nframes, nsplit, W;   // W: row width
    int    Wp;                // width of a pow row (== W except one-sided rows)
    // one-sided rows (real input, R = 1): row column c is bin k = (c + os_lo + (M+1)/2) mod M
    // of the M = N/2+1 one-sided bins (np.fft.fftshift of an odd-length array), doubled
    // except DC and Nyquist; pow holds the full two-sided, fftshifted N-bin row
    int    onesided, os_lo, os_N;
    float  scale;             // 1 / (fs * sum w^2) / nseg
    float  alpha;             // < 0: EMA off
    int    linear;            // 1: emit linear power instead of dB20
    float *ema_state;         // [W]
    int    ema_have;          // the state holds a row (host-side knowledge, launch order)
    float *rows;              // [nframes][W] or null
    float *ring;              // [ring_rows][W] or null
    long long ring_pos;       // ring slot of frame 0
    int    ring_rows;
    // channel-batched launches: frame index b = ch*chan_frames + f lands at
    // rows[ch*chan_row_stride + f*W]; chan_frames == 0: rows[b*W]
    int    chan_frames;
    long long chan_row_stride;
};

__device__ __forceinline__ void emit_row_value(const FinalizeParams &p, int f, int col, float v) {
    const float out = p.linear ? v : 20.0f * log10f(fabsf(v));
    if (p.rows) {
        const size_t at = p.chan_frames ? (size_t)(f / p.chan_frames) * (size_t)p.chan_row_stride +
                                              (size_t)(f % p.chan_frames) * p.W
                                        : (size_t)f * p.W;
        p.rows[at + col] = out;
    }
    // only the newest ring_rows frames of a launch enter the ring: older ones would share
    // a slot with them (unordered writes from different threads)
    if (p.ring && f >= p.nframes - p.ring_rows)
        p.ring[(size_t)((p.ring_pos + f) % p.ring_rows) * p.W + col] = out;
}

// column of the pow row that feeds row column `col`, and the one-sided factor
__device__ __forceinline__ int pow_col(const FinalizeParams &p, int col, float &factor) {
    factor = 1.f;
    if (!p.onesided) return col;
    const int M = p.os_N / 2 + 1;
    int k = col + p.os_lo + (M + 1) / 2;
    if (k >= M) k -= M;
    factor = (k == 0 || k == p.os_N / 2) ? 1.f : 2.f;
    return (k + p.os_N / 2) & (p.os_N - 1);
}

// without EMA: one thread per (frame, column)
__global__ void reduce_rows_kernel(const FinalizeParams p) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)p.nframes * p.W) return;
    const int f = (int)(idx / p.W), col = (int)(idx % p.W);
    float factor;
    const float *src = p.pow_io + (size_t)f * p.nsplit * p.Wp + pow_col(p, col, factor);
    float pw = 0.f;
    for (int s = 0; s < p.nsplit; ++s) pw += src[(size_t)s * p.Wp];
    pw *= p.scale * factor;
    emit_row_value(p, f, col, pw);
}

// EMA is order dependent (a_i needs a_(i-1)) but only along the frames: one CTA
// owns EMA_COLS columns.  All warps reduce the per-split sums of EMA_FR frames
// into a shared-memory tile (32-byte sectors, EMA_CH independent loads in
// flight per thread), the first EMA_COLS threads walk the recurrence over the
// tile (nothing but LDS + 2 dependent ops per frame), all warps turn the
// averaged power into rows.  Same operations in the same order per (frame,
// column) as a serial walk: a row does not depend on the batch it is in.
constexpr int EMA_COLS = 4;          // (8 -> 4: twice the CTAs and one 512-frame tile; the kernel is latency-, not traffic-bound)
constexpr int EMA_NT = 512;
constexpr int EMA_CH = 4;
constexpr int EMA_FPP = EMA_NT / EMA_COLS;          // frames per pass of the CTA
constexpr int EMA_FR = EMA_FPP * EMA_CH;            // frames per tile

__global__ void __launch_bounds__(EMA_NT) ema_rows_kernel(const FinalizeParams p) {
    __shared__ float tile[EMA_FR * EMA_COLS];
    const int c = threadIdx.x % EMA_COLS, fs = threadIdx.x / EMA_COLS;
    const int col = blockIdx.x * EMA_COLS + c;
    const bool ok = col < p.W;
    const bool walker = threadIdx.x < EMA_COLS;
    bool have = p.ema_have != 0;
    float a = (walker && have && ok) ? p.ema_state[col] : 0.f;
    const size_t stride = (size_t)p.nsplit * p.Wp;
    float factor = 1.f;
    const int scol = ok ? pow_col(p, col, factor) : 0;
    const float scale = p.scale * factor;
    for (int f0 = 0; f0 < p.nframes; f0 += EMA_FR) {
        const int n = min(EMA_FR, p.nframes - f0);
        {
            float pw[EMA_CH];
#pragma unroll
            for (int k = 0; k < EMA_CH; ++k) pw[k] = 0.f;
            for (int s = 0; s < p.nsplit; ++s) {
#pragma unroll
                for (int k = 0; k < EMA_CH; ++k) {
                    const int i = fs + k * EMA_FPP;
                    if (ok && i < n) pw[k] += p.pow_io[(size_t)(f0 + i) * stride + (size_t)s * p.Wp + scol];
                }
            }
#pragma unroll
            for (int k = 0; k < EMA_CH; ++k) tile[(fs + k * EMA_FPP) * EMA_COLS + c] = pw[k] * scale;
        }
        __syncthreads();
        if (walker) {
#pragma unroll 8
            for (int i = 0; i < n; ++i) {
                const float pw = tile[i * EMA_COLS + c];
                a = have ? fmaf(p.alpha, pw - a, a) : pw;
                have = true;
                tile[i * EMA_COLS + c] = a;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < EMA_CH; ++k) {
            const int i = fs + k * EMA_FPP;
            if (ok && i < n) emit_row_value(p, f0 + i, col, tile[i * EMA_COLS + c]);
        }
        __syncthreads();
    }
    if (walker && ok && p.nframes > 0) p.ema_state[col] = a;
}

}  // namespace zfb
