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
// Data movement is tiled TMA (cp.async.bulk.tensor, zfb_tma.cuh): the frames are
// described to the copy engine as a 4-D tensor [frame][stream][16-sample block][16
// samples], so ONE request brings the next 16 samples of all 32 streams of a warp
// (a [32 rows][128 B] box, SWIZZLE_128B: the lanes' LDS.128 of their own rows hit
// distinct banks) and one request takes 8 results of every stream back ([32][64 B],
// SWIZZLE_64B).  A warm-up block that lies before (after) a stream is the tail
// (head) of the neighbouring stream -- the same box one stream over; rows that
// fall outside the frame are zero-filled / dropped by the engine itself, which is
// the zero extension this kernel assumes: there is not one bounds check, fallback
// path or per-lane address in it.  (The first version issued one 256-byte
// cp.async.bulk per LANE; UBLKCP is a uniform-datapath instruction, the compiler
// serialises it over the lanes and the copy engine needs ~46 cycles per request:
// 4000 cycles per 32 samples, 326 us for the cfg2 last stage, profiles/r02a_*.)
// The forward pass writes its half of the sum to the output; the backward pass of
// the same warp pulls it back (L2), adds its own half and stores the result.
//
// Outside [0, L) the input is zero; the reference's own edge rules (odd
// extension, steady-state zi at every pass) differ from that only within ~256
// samples of a chunk end, which mode FAST recomputes exactly on its edge strips
// and patches in afterwards (zfb_engine.cu).
#pragma once
#include "zfb_decim.cuh"
#include "zfb_tma.cuh"

namespace zfb {

constexpr int IS_BLK = 16;                         // samples per block (one 128-byte row of a tile)
constexpr int IS_LAG = 6;                          // samples between a section-0 input and the cascade's output
constexpr int IS_XTILE = 32 * IS_BLK * 8;          // 4096 B: [32 streams][16 samples]
constexpr int IS_OTILE = 32 * (IS_BLK / 2) * 8;    // 2048 B: [32 streams][8 results]

struct IirStreamParams {
    int S16;              // blocks per stream (stream length S = 16 * S16)
    int Wb;               // warm-up blocks (Wb + 1 <= S16)
    int nspf;             // streams per frame (rows beyond it are out of bounds for the copy engine)
    int groups;           // warps per frame: ceil(nspf / 32)
    // L2 residency hints (the launch moves more bytes than L2 holds: every input block is read by
    // both passes, every result tile is written, pulled back and written again).  The backward
    // pass starts where the forward pass ended, so the LAST blocks of a stream are the first to be
    // wanted again: blocks >= keep_from are loaded / stored evict_last by the forward pass, the
    // others and everything the backward pass reads evict_first.  keep_from >= S16: no hints.
    int keep_from;
};

template <int NS, int NO>
struct IirStreamShape {
    static constexpr int XBYTES = NS * IS_XTILE;
    static constexpr int OBYTES = NO * IS_OTILE;
    static constexpr size_t SMEM = 1024 + (size_t)XBYTES + OBYTES + (NS + NO) * sizeof(uint64_t);
};

struct IirState {
    Sec4   s;
    float2 q1a, q1b, q2a, q2b, q3a, q3b;     // section outputs waiting for the next section
    float2 A1, A2, A3, A4;                   // partial sums of the next four kept outputs
};

__device__ __forceinline__ void iir_zero(IirState &st) {
    sec_zero(st.s);
    st.q1a = st.q1b = st.q2a = st.q2b = st.q3a = st.q3b = make_float2(0.f, 0.f);
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

// One block (16 cascade outputs) of one lane.  The inputs lead the outputs by IS_LAG = 6
// samples in processing order: 10 come from the lane's row of tile `ta`, 6 from tile `tb`
// (the next one in processing order).  `xsw` / `osw` are the lane's swizzle masks.
// MODE 0: warm-up (recursion only); 1: numerator running, nothing kept; 2: the 8 kept outputs
// go to the lane's row of `orow` (forward: stored; backward: added to what the row holds).
template <bool BWD, int MODE>
__device__ __forceinline__ void iir_block(const unsigned char *ta, const unsigned char *tb, unsigned xsw,
                                          unsigned char *orow, unsigned osw, IirState &st,
                                          const float (&na1)[NSEC], const float (&na2)[NSEC],
                                          const float (&bc)[9]) {
    float2 held = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < IS_BLK / 2; ++t) {
        // 16-byte chunk (two samples) of this step
        const unsigned char *tile = (t < 5) ? ta : tb;
        const int chunk = BWD ? ((t < 5) ? 4 - t : 12 - t) : ((t < 5) ? 3 + t : t - 5);
        const float4 x = *reinterpret_cast<const float4 *>(tile + (((unsigned)chunk << 4) ^ xsw));
        const float2 lo = make_float2(x.x, x.y), hi = make_float2(x.z, x.w);
        float2 o;
        iir_step<BWD, (MODE >= 1)>(st, BWD ? hi : lo, BWD ? lo : hi, na1, na2, bc, o);
        if (MODE == 2) {
            if ((t & 1) == 0) {
                held = o;
            } else {
                const int oc = BWD ? (IS_BLK / 2 - 1 - t) / 2 : t / 2;       // result chunk (two outputs)
                float4 *dst = reinterpret_cast<float4 *>(orow + (((unsigned)oc << 4) ^ osw));
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
__global__ void __launch_bounds__(32) iir_stream_kernel(ZFB_TMAP_PARAM tm_in, ZFB_TMAP_PARAM tm_out,
                                                        const IirStreamParams p) {
    using SH = IirStreamShape<NS, NO>;
    ZFB_DYN_SMEM(smem_raw);
    const int lane = threadIdx.x;
    // tiles want 1024-byte alignment (the swizzle pattern is a function of the address)
    unsigned char *base = smem_raw + smem_align_pad(smem_raw, 1024);
    unsigned char *xbuf = base;                                   // [NS] tiles of [32][128 B]
    unsigned char *obuf = base + SH::XBYTES;                      // [NO] tiles of [32][64 B]
    uint64_t *xbar = reinterpret_cast<uint64_t *>(base + SH::XBYTES + SH::OBYTES);   // [NS]
    uint64_t *pbar = xbar + NS;                                   // [NO]
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) mbar_init(xbar + i, 1);
#pragma unroll
        for (int i = 0; i < NO; ++i) mbar_init(pbar + i, 1);
        mbar_fence_init();
    }
    __syncwarp();

    const int frame = blockIdx.x / p.groups;
    const int s0 = (blockIdx.x % p.groups) * 32;       // first stream of this warp
    const int S16 = p.S16, Wb = p.Wb;
    const unsigned xrow = (unsigned)lane * (IS_BLK * 8), xsw = (unsigned)(lane & 7) << 4;
    const unsigned orow_off = (unsigned)lane * (IS_BLK * 4), osw = (unsigned)((lane >> 1) & 3) << 4;

    float na1[NSEC], na2[NSEC], bc[9];
#pragma unroll
    for (int k = 0; k < NSEC; ++k) {
        na1[k] = c_dec.na1[k];
        na2[k] = c_dec.na2[k];
    }
#pragma unroll
    for (int j = 0; j < 9; ++j) bc[j] = c_dec.bc[j];

    unsigned gx = 0;          // x tiles consumed so far (both directions): slot gx % NS, parity (gx / NS) & 1
    const int npieces = Wb + S16;
    const bool hints = p.keep_from < S16;
    const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();

#pragma unroll 1
    for (int dir = 0; dir < 2; ++dir) {
        const bool bwd = dir == 1;
        // piece k works on output block jb(k) of the stream; x tile q is block jx(q)
        //   forward : jb = k - Wb,            jx(q) = q - Wb            (tiles k, k+1)
        //   backward: jb = S16 + Wb - 1 - k,  jx(q) = S16 + Wb - 1 - q  (tiles k, k+1)
        auto issue_x = [&](int q) {
            const int jx = bwd ? S16 + Wb - 1 - q : q - Wb;
            int c1 = jx, c2 = s0;
            if (jx < 0) { c1 += S16; c2 -= 1; }
            else if (jx >= S16) { c1 -= S16; c2 += 1; }
            const unsigned slot = (gx + (unsigned)q) % NS;
            mbar_arrive_expect_tx(xbar + slot, IS_XTILE);
            if (hints)
                tma_load_4d_hint(xbuf + slot * IS_XTILE, &tm_in, 0, c1, c2, frame, xbar + slot,
                                 (!bwd && jx >= p.keep_from) ? pol_keep : pol_drop);
            else
                tma_load_4d(xbuf + slot * IS_XTILE, &tm_in, 0, c1, c2, frame, xbar + slot);
        };
        IirState st;
        iir_zero(st);
        if (lane == 0)
            for (int q = 0; q < NS - 1 && q <= npieces; ++q) issue_x(q);
        mbar_wait(xbar + gx % NS, (gx / NS) & 1);                 // tile 0

#pragma unroll 1
        for (int k = 0; k < npieces; ++k) {
            const int jb = bwd ? S16 + Wb - 1 - k : k - Wb;
            const int e = bwd ? (jb < S16 ? S16 - 1 - jb : -(jb - S16 + 1)) : jb;   // kept piece number, < 0: warm-up
            // everybody has left tile k-1: its slot takes tile k + NS - 1
            __syncwarp();
            if (lane == 0) {
                if (k + NS - 1 <= npieces) issue_x(k + NS - 1);
                if (bwd) {
                    // forward half of the kept piece NO - 2 ahead into its row tile (last read by the
                    // store of kept piece en - NO: at most one later store may still be reading)
                    static_assert(NO >= 3, "row tiles: one being filled, NO - 2 on their way, one draining");
                    const int en = e + (NO - 2);
                    if (en >= 0 && en < S16) {
                        bulk_wait_read<1>();
                        const unsigned slot = (unsigned)en % NO;
                        mbar_arrive_expect_tx(pbar + slot, IS_OTILE);
                        if (hints)
                            tma_load_4d_hint(obuf + slot * IS_OTILE, &tm_out, 0, S16 - 1 - en, s0, frame, pbar + slot,
                                             pol_drop);
                        else
                            tma_load_4d(obuf + slot * IS_OTILE, &tm_out, 0, S16 - 1 - en, s0, frame, pbar + slot);
                    }
                } else if (e >= 0) {
                    bulk_wait_read<NO - 1>();                     // the row tile's previous store has left it
                }
            }
            __syncwarp();
            const unsigned sa = (gx + (unsigned)k) % NS, sb = (gx + (unsigned)k + 1) % NS;
            mbar_wait(xbar + sb, ((gx + (unsigned)k + 1) / NS) & 1);                  // tile k + 1
            const unsigned char *ta = xbuf + sa * IS_XTILE + xrow, *tb = xbuf + sb * IS_XTILE + xrow;
            if (e < -1) {
                if (bwd) iir_block<true, 0>(ta, tb, xsw, nullptr, 0, st, na1, na2, bc);
                else iir_block<false, 0>(ta, tb, xsw, nullptr, 0, st, na1, na2, bc);
            } else if (e == -1) {
                if (bwd) iir_block<true, 1>(ta, tb, xsw, nullptr, 0, st, na1, na2, bc);
                else iir_block<false, 1>(ta, tb, xsw, nullptr, 0, st, na1, na2, bc);
            } else {
                const unsigned os = (unsigned)e % NO;
                unsigned char *orow = obuf + os * IS_OTILE + orow_off;
                if (bwd) {
                    mbar_wait(pbar + os, ((unsigned)e / NO) & 1);
                    iir_block<true, 2>(ta, tb, xsw, orow, osw, st, na1, na2, bc);
                } else {
                    iir_block<false, 2>(ta, tb, xsw, orow, osw, st, na1, na2, bc);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    if (hints && !bwd)
                        tma_store_4d_hint(&tm_out, 0, jb, s0, frame, obuf + os * IS_OTILE,
                                          jb >= p.keep_from ? pol_keep : pol_drop);
                    else
                        tma_store_4d(&tm_out, 0, jb, s0, frame, obuf + os * IS_OTILE);
                    bulk_commit();
                }
            }
        }
        gx += (unsigned)npieces + 1;
        if (lane == 0) bulk_wait<0>();   // forward halves are in place before the backward pass pulls them back
        __syncwarp();
    }
}

}  // namespace zfb
