// Streaming form of one zero-phase decimate-by-2 stage (interior / LTI part).
//
// Replaces, for samples far enough from the chunk ends, what
// scipy.signal.decimate(x, 2) computes at pypanadapter_spectrum.py:2098 /
// pypanadapter_thread.py:1534 (scipy/signal/_signaltools.py:5317-5369:
// cheby1(8, 0.05, 0.4) through sosfiltfilt, then [::2]).  In the interior that
// is the LTI zero-phase filter Z(z) = H(z) H(1/z).  Z is split into partial
// fractions, Z(z) = G(z) + G(1/z), with G causal and of order 8 with H's poles:
//     G(z) = Bc(z) / prod_k A_k(z),  Bc of degree 8 (host, fp64: build_decim_const).
// So y[n] = (G x)[n] + (G~ x)[n]: one causal and one anti-causal recursion, both
// fed by the INPUT -- no forward-filtered intermediate signal, and each needs
// its 9-tap numerator only at the kept (even) n:
//     8 (all-pole) + 4.5 (numerator) packed FMAs per sample and direction,
// against 44.5 + hand-off for the two-sweep shared-memory kernel (zfb_decim.cuh).
//
// Parallel scheme: a LANE owns a contiguous stream of S samples of one frame;
// it warms its recursion up over the Wm samples before (after) the stream
// (0.935^256 = 3e-8 of the state) and then runs through the stream with its
// whole state in registers, software-pipelined across the four sections (the
// four recurrences are independent instruction streams; the numerator is in
// transposed form, four partial sums per lane).  No barriers, no hand-off, no
// idle lanes; a warp of 32 streams keeps the FMA pipe busy on its own.
//
// Data movement is all 1-D TMA (cp.async.bulk, zfb_tma.cuh): every lane pulls
// 256-byte pieces of ITS stream into a padded shared-memory row (row pitch 272 B:
// the lanes' LDS.128 hit distinct banks), NS pieces in flight behind per-lane
// mbarriers, and pushes 128-byte pieces of results back the same way.  The
// forward pass writes its half of the sum to the output; the backward pass of
// the same lane pulls it back (L2), adds its own half and stores the result.
// A per-lane LDG/STG.128 here would touch 32 cache lines per warp instruction
// (32 L1 wavefronts); the bulk copies bypass that path.
//
// Outside [0, L) the input is held constant (x[0], x[L-1]); the reference's own
// edge rules (odd extension, steady-state zi at every pass) differ from that
// only within ~256 samples of a chunk end, which mode FAST recomputes exactly on
// its edge strips (zfb_engine.cu).
#pragma once
#include "zfb_decim.cuh"
#include "zfb_tma.cuh"

namespace zfb {

constexpr int IS_CH = 32;                          // samples per piece and lane
constexpr int IS_XPITCH = IS_CH * 8 + 16;          // 272 B rows: conflict-free LDS.128
constexpr int IS_OPITCH = (IS_CH / 2) * 8 + 16;    // 144 B rows of 16 outputs
constexpr int IS_LAG = 6;                          // samples between a section-0 input and the cascade's output

struct IirStreamParams {
    const float2 *in;          // [frames][in_stride] complex64, 16-byte aligned rows
    long long     in_stride;
    int           L;           // stage input length per frame
    float2       *out;         // [frames][out_stride]
    long long     out_stride;
    int           S;           // samples per stream (multiple of IS_CH)
    int           Wm;          // warm-up samples (multiple of IS_CH)
    int           nspf;        // streams per frame
    int           n0;          // input position of stream 0's first sample (multiple of 4)
    int           m_lo, m_hi;  // outputs [m_lo, m_hi) of a frame are written
    long long     nstreams;    // frames * nspf
};

template <int NS, int NO>
struct IirStreamShape {
    static constexpr int XBYTES = NS * 32 * IS_XPITCH;
    static constexpr int OBYTES = NO * 32 * IS_OPITCH;
    static constexpr int BARS = (NS + NO) * 32;
    static constexpr size_t SMEM = (size_t)XBYTES + OBYTES + BARS * sizeof(uint64_t);
};

struct IirState {
    Sec4   s;
    float2 q1a, q1b, q2a, q2b, q3a, q3b;     // section outputs waiting for the next section
    float2 A1, A2, A3, A4;                   // partial sums of the next four kept outputs
};

__device__ __forceinline__ void iir_init(IirState &st, float2 c) {
    sec_steady(st.s, c);
    st.q1a = st.q1b = st.s.w1[0];
    st.q2a = st.q2b = st.s.w1[1];
    st.q3a = st.q3b = st.s.w1[2];
    st.A1 = st.A2 = st.A3 = st.A4 = make_float2(0.f, 0.f);
}

// one pair of samples, xa first in processing order; with FIR, `o` is the kept output that completes
template <bool BWD, bool FIR>
__device__ __forceinline__ void iir_step(IirState &st, float2 xa, float2 xb, const float (&na1)[NSEC],
                                         const float (&na2)[NSEC], const float (&bc)[9], float2 &o) {
    const float2 va = pole(st.q3a, st.s.w1[3], st.s.w2[3], na1[3], na2[3]);
    const float2 vb = pole(st.q3b, st.s.w1[3], st.s.w2[3], na1[3], na2[3]);
    st.q3a = pole(st.q2a, st.s.w1[2], st.s.w2[2], na1[2], na2[2]);
    st.q3b = pole(st.q2b, st.s.w1[2], st.s.w2[2], na1[2], na2[2]);
    st.q2a = pole(st.q1a, st.s.w1[1], st.s.w2[1], na1[1], na2[1]);
    st.q2b = pole(st.q1b, st.s.w1[1], st.s.w2[1], na1[1], na2[1]);
    st.q1a = pole(xa, st.s.w1[0], st.s.w2[0], na1[0], na2[0]);
    st.q1b = pole(xb, st.s.w1[0], st.s.w2[0], na1[0], na2[0]);
    if (FIR) {
        if (!BWD) {     // va = v[p] (p even, kept), vb = v[p+1]:  y[p] = sum_j bc[j] v[p-j]
            o = pk_fma(bc[0], va, st.A1);
            st.A1 = pk_fma(bc[1], vb, pk_fma(bc[2], va, st.A2));
            st.A2 = pk_fma(bc[3], vb, pk_fma(bc[4], va, st.A3));
            st.A3 = pk_fma(bc[5], vb, pk_fma(bc[6], va, st.A4));
            st.A4 = pk_fma(bc[7], vb, pk_mul(bc[8], va));
        } else {        // va = v[p+1], vb = v[p] (p even, kept):  y[p] = sum_j bc[j] v[p+j]
            o = pk_fma(bc[0], vb, pk_fma(bc[1], va, st.A1));
            st.A1 = pk_fma(bc[2], vb, pk_fma(bc[3], va, st.A2));
            st.A2 = pk_fma(bc[4], vb, pk_fma(bc[5], va, st.A3));
            st.A3 = pk_fma(bc[6], vb, pk_fma(bc[7], va, st.A4));
            st.A4 = pk_mul(bc[8], vb);
        }
    }
}

// 32 samples of one lane from its shared-memory row.  MODE 0: warm-up (recursion only),
// 1: numerator running, nothing kept, 2: the 16 kept outputs go to `orow` (forward: stored;
// backward: added to what the row holds -- the forward half, or zeros).
template <bool BWD, int MODE>
__device__ __forceinline__ void iir_piece(const unsigned char *xrow, unsigned char *orow, IirState &st,
                                          const float (&na1)[NSEC], const float (&na2)[NSEC],
                                          const float (&bc)[9]) {
    float2 held = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < IS_CH / 2; ++t) {
        const int xi = BWD ? (IS_CH / 2 - 1 - t) : t;
        const float4 x = *reinterpret_cast<const float4 *>(xrow + xi * 16);
        const float2 lo = make_float2(x.x, x.y), hi = make_float2(x.z, x.w);
        float2 o;
        iir_step<BWD, (MODE >= 1)>(st, BWD ? hi : lo, BWD ? lo : hi, na1, na2, bc, o);
        if (MODE == 2) {
            // output index within the piece: forward t, backward 15 - t; two share a 16-byte slot
            if ((t & 1) == 0) {
                held = o;
            } else {
                const int slot = BWD ? (IS_CH / 2 - 1 - t) / 2 : t / 2;
                float4 *dst = reinterpret_cast<float4 *>(orow + slot * 16);
                if (!BWD) {
                    *dst = make_float4(held.x, held.y, o.x, o.y);
                } else {
                    const float4 p = *dst;
                    *dst = make_float4(p.x + o.x, p.y + o.y, p.z + held.x, p.w + held.y);
                }
            }
        }
    }
}

template <int NS, int NO>
__global__ void __launch_bounds__(32) iir_stream_kernel(const IirStreamParams p) {
    using SH = IirStreamShape<NS, NO>;
    ZFB_DYN_SMEM(smem_raw);
    const int lane = threadIdx.x;
    unsigned char *xbuf = smem_raw;                               // [NS][32][IS_XPITCH]
    unsigned char *obuf = smem_raw + SH::XBYTES;                  // [NO][32][IS_OPITCH]
    uint64_t *xbar = reinterpret_cast<uint64_t *>(smem_raw + SH::XBYTES + SH::OBYTES);   // [NS][32]
    uint64_t *pbar = xbar + NS * 32;                              // [NO][32]

#pragma unroll
    for (int i = 0; i < NS; ++i) mbar_init(xbar + i * 32 + lane, 1);
#pragma unroll
    for (int i = 0; i < NO; ++i) mbar_init(pbar + i * 32 + lane, 1);
    mbar_fence_init();
    __syncwarp();

    long long g = (long long)blockIdx.x * 32 + lane;
    const bool live = g < p.nstreams;
    if (!live) g = p.nstreams - 1;                 // same work, nothing stored
    const int frame = (int)(g / p.nspf);
    const int a = p.n0 + (int)(g % p.nspf) * p.S;  // first sample of the stream
    const float2 *in_row = p.in + (size_t)frame * (size_t)p.in_stride;
    float2 *out_row = p.out + (size_t)frame * (size_t)p.out_stride;
    const int L = p.L;
    const int m_lo = live ? p.m_lo : 0, m_hi = live ? p.m_hi : 0;
    const int jw = p.Wm / IS_CH;                   // warm-up pieces
    const int npieces = jw + p.S / IS_CH;

    float na1[NSEC], na2[NSEC], bc[9];
#pragma unroll
    for (int k = 0; k < NSEC; ++k) {
        na1[k] = c_dec.na1[k];
        na2[k] = c_dec.na2[k];
    }
#pragma unroll
    for (int j = 0; j < 9; ++j) bc[j] = c_dec.bc[j];

    long long gx = 0;        // x pieces consumed so far (both directions): slot gx % NS, parity (gx / NS) & 1
    long long gp = 0;        // forward halves pulled back so far (backward direction only)

    // input piece j of a direction: positions [cin, cin + 32) of the frame into slot `slot`
    auto issue_x = [&](int cin, int slot) {
        unsigned char *row = xbuf + ((size_t)slot * 32 + lane) * IS_XPITCH;
        uint64_t *bar = xbar + slot * 32 + lane;
        const float2 *src = in_row + cin;
        if (cin >= 0 && cin + IS_CH <= L && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
            mbar_arrive_expect_tx(bar, IS_CH * 8);
            bulk_g2s(row, src, IS_CH * 8, bar);
        } else {
            float2 *r = reinterpret_cast<float2 *>(row);
            for (int i = 0; i < IS_CH; ++i) {
                int q = cin + i;
                q = q < 0 ? 0 : (q >= L ? L - 1 : q);
                r[i] = in_row[q];
            }
            mbar_arrive(bar);
        }
    };
    auto out_bulk_ok = [&](int mc) {
        return mc >= m_lo && mc + IS_CH / 2 <= m_hi && ((reinterpret_cast<uintptr_t>(out_row + mc) & 15) == 0);
    };

#pragma unroll 1
    for (int dir = 0; dir < 2; ++dir) {
        const bool bwd = dir == 1;
        // output-side position of piece j: forward a - Wm + 32 j, backward a + S + Wm - 32 (j + 1);
        // the input piece runs IS_LAG samples ahead of it in processing order
        const int top = a + p.S + p.Wm;
        auto piece_po = [&](int j) { return bwd ? top - IS_CH * (j + 1) : a - p.Wm + IS_CH * j; };
        auto piece_cin = [&](int j) { return bwd ? piece_po(j) - IS_LAG : piece_po(j) + IS_LAG; };

        // initial state: steady state for the constant the first warm-up sample would hold for ever
        IirState st;
        {
            int q0 = bwd ? piece_cin(0) + IS_CH - 1 : piece_cin(0);
            q0 = q0 < 0 ? 0 : (q0 >= L ? L - 1 : q0);
            iir_init(st, in_row[q0]);
        }
        for (int j = 0; j < NS - 1 && j < npieces; ++j) issue_x(piece_cin(j), (int)((gx + j) % NS));

#pragma unroll 1
        for (int j = 0; j < npieces; ++j) {
            if (j + NS - 1 < npieces) issue_x(piece_cin(j + NS - 1), (int)((gx + NS - 1) % NS));
            const int e = j - jw;                             // kept piece number (>= 0: outputs are kept)
            const int mc = piece_po(j) >> 1;                  // first output of the piece
            if (bwd) {
                // pull the forward half of the NEXT kept piece into its row (or zeros where the
                // piece is not written whole); its row was last read by the store of piece e+1-NO
                const int en = e + 1;
                if (en >= 0 && en < p.S / IS_CH) {
                    bulk_wait_read<(NO >= 2 ? NO - 2 : 0)>();
                    const int slot = (int)((gp + en) % NO);
                    unsigned char *row = obuf + ((size_t)slot * 32 + lane) * IS_OPITCH;
                    uint64_t *bar = pbar + slot * 32 + lane;
                    const int mcn = piece_po(j + 1) >> 1;
                    if (out_bulk_ok(mcn)) {
                        mbar_arrive_expect_tx(bar, IS_CH * 4);
                        bulk_g2s(row, out_row + mcn, IS_CH * 4, bar);
                    } else {
                        float4 *r = reinterpret_cast<float4 *>(row);
#pragma unroll
                        for (int i = 0; i < IS_CH / 4; ++i) r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        mbar_arrive(bar);
                    }
                }
            } else if (e >= 0) {
                bulk_wait_read<NO - 1>();                     // the row's previous store has left it
            }
            const int xs = (int)(gx % NS);
            mbar_wait(xbar + xs * 32 + lane, (uint32_t)((gx / NS) & 1));
            const unsigned char *xrow = xbuf + ((size_t)xs * 32 + lane) * IS_XPITCH;
            if (e < -1) {
                if (bwd) iir_piece<true, 0>(xrow, nullptr, st, na1, na2, bc);
                else iir_piece<false, 0>(xrow, nullptr, st, na1, na2, bc);
            } else if (e == -1) {
                if (bwd) iir_piece<true, 1>(xrow, nullptr, st, na1, na2, bc);
                else iir_piece<false, 1>(xrow, nullptr, st, na1, na2, bc);
            } else {
                const int os = (int)(((bwd ? gp : 0) + e) % NO);
                unsigned char *orow = obuf + ((size_t)os * 32 + lane) * IS_OPITCH;
                if (bwd) {
                    mbar_wait(pbar + os * 32 + lane, (uint32_t)((((gp + e) / NO)) & 1));
                    iir_piece<true, 2>(xrow, orow, st, na1, na2, bc);
                } else {
                    iir_piece<false, 2>(xrow, orow, st, na1, na2, bc);
                }
                if (out_bulk_ok(mc)) {
                    fence_proxy_async();
                    bulk_s2g(out_row + mc, orow, IS_CH * 4);
                } else {
                    const float2 *r = reinterpret_cast<const float2 *>(orow);
                    for (int i = 0; i < IS_CH / 2; ++i) {
                        const int m = mc + i;
                        if (m >= m_lo && m < m_hi) {
                            float2 v = r[i];
                            if (bwd) {
                                const float2 f = out_row[m];
                                v = make_float2(v.x + f.x, v.y + f.y);
                            }
                            out_row[m] = v;
                        }
                    }
                }
                bulk_commit();
            }
            gx += 1;
        }
        if (bwd) gp += p.S / IS_CH;
        bulk_wait<0>();          // forward halves are in place before the backward pass pulls them back
    }
}

}  // namespace zfb
