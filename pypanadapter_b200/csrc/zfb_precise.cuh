// fp64 path for rows with few Welch segments.
//
// north_star asks for 0.01 dB on every bin above -100 dBFS.  The fp32 chain's
// rounding noise sits 110-120 dB under the strongest signal of a chunk; averaged
// over many segments that is far below the bar, but in a row of one to six
// segments an isolated bin 75-85 dB under a strong tone can be off by a few
// hundredths (round 1: 25 of 1594 random configurations, all of this kind).
// Such rows are tiny jobs -- a handful of FFTs -- so they take this path: the
// reference's own arithmetic in double precision, parallel enough to finish in
// well under a millisecond, with no attempt at the fp32 pipe's throughput.
//
//   px_load_kernel   S:543 (u/127.5 - 1), np.flip (S:460), LO (S:2090-2094)
//   px_iir_kernel    scipy.signal.decimate(x, 2): sosfiltfilt's odd extension,
//                    sosfilt_zi start-up at both passes, direct form II
//                    transposed sections as scipy's _sosfilt evaluates them
//                    (scipy/signal/_signaltools.py:5091-5204, 5317-5369); one
//                    thread per stream of PX_STREAM samples, warmed up over
//                    PX_WARM samples (0.935^576 = 2e-17) unless the stream starts
//                    at the chunk's own (extended) end, where the start-up is exact
//   px_welch_kernel  scipy.signal.welch per segment: detrend, window, radix-2
//                    Stockham FFT in global memory, |X|^2 (S:2111)
//   px_rows_kernel   mean over segments, density scale, fftshift + crop, EMA, dB20
#pragma once
#include "zfb_common.cuh"

namespace zfb {

constexpr int PX_STREAM = 1024;       // (2048 left a batch of 64 cfg5 frames with 17 CTAs per launch)
constexpr int PX_WARM = 576;

struct PreciseConst {
    double sos[NSEC][6];     // b0 b1 b2 1 a1 a2
    double zi[NSEC][2];      // sosfilt_zi: DF2T state per unit input
};

struct PxLoadParams {
    const void *in;          // [frames][in_stride] wire samples (uint8 IQ or complex64)
    long long   in_stride;
    int         n;           // samples per frame
    int         kind;        // KIND_U8_RAW / KIND_C64_RAW
    int         flip;
    int         mix;         // 0: no LO (fft_ratio 1 or ZFB_FLAG_NO_LO)
    unsigned long long phase_inc;   // frac(f_demod / fs) * 2^64
    double      amp;         // sqrt(2) (S:2093)
    double2    *out;         // [frames][out_stride]
    long long   out_stride;
    int         frames;
};

__global__ void __launch_bounds__(256) px_load_kernel(const PxLoadParams p) {
    const long long total = (long long)p.frames * p.n;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(g / p.n), i = (int)(g % p.n);
        const int src = p.flip ? p.n - 1 - i : i;
        double re, im;
        if (p.kind == KIND_U8_RAW) {
            const unsigned char *s = (const unsigned char *)p.in + ((size_t)f * (size_t)p.in_stride + (size_t)src) * 2;
            re = (double)s[0] / 127.5 - 1.0;          // pyrtlsdr: iq /= 127.5; iq -= (1 + 1j)
            im = (double)s[1] / 127.5 - 1.0;
        } else {
            const float2 v = ((const float2 *)p.in)[(size_t)f * (size_t)p.in_stride + (size_t)src];
            re = (double)v.x;
            im = (double)v.y;
        }
        if (p.mix) {
            const unsigned long long ph = (unsigned long long)i * p.phase_inc;     // mod 2^64
            double s, c;
            sincospi(-2.0 * ((double)(ph >> 11) * (1.0 / 9007199254740992.0)), &s, &c);
            const double lr = p.amp * c, li = p.amp * s;
            const double r2 = re * lr - im * li, i2 = re * li + im * lr;
            re = r2;
            im = i2;
        }
        p.out[(size_t)f * (size_t)p.out_stride + (size_t)i] = make_double2(re, im);
    }
}

struct PxIirParams {
    const double2 *x;        // forward: stage input [frames][x_stride]; backward: forward output [frames][x_stride]
    long long      x_stride;
    double2       *y;        // forward: [frames][y_stride] over ext positions -27 .. L+26; backward: outputs [frames][y_stride]
    long long      y_stride;
    int            L;        // stage input length
    int            nstreams; // per frame
    int            frames;
    int            backward;
};

// ext position e in [0, L + 2*PADLEN): odd extension of x (scipy odd_ext)
__device__ __forceinline__ double2 px_ext(const double2 *x, int L, int e) {
    const int n = e - PADLEN;
    if (n < 0) {
        const double2 a = x[0], b = x[-n];
        return make_double2(2.0 * a.x - b.x, 2.0 * a.y - b.y);
    }
    if (n >= L) {
        const double2 a = x[L - 1], b = x[2 * (L - 1) - n];
        return make_double2(2.0 * a.x - b.x, 2.0 * a.y - b.y);
    }
    return x[n];
}

__constant__ PreciseConst c_px;

__global__ void __launch_bounds__(128) px_iir_kernel(const PxIirParams p) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.frames * p.nstreams) return;
    const int f = g / p.nstreams, j = g % p.nstreams;
    const int E = p.L + 2 * PADLEN;                 // extended length
    const int a = j * PX_STREAM, b = min(E, a + PX_STREAM);
    double z1r[NSEC], z1i[NSEC], z2r[NSEC], z2i[NSEC];
    // processing order index t runs over ext positions: forward e = t, backward e = E - 1 - t
    const int start = max(0, a - PX_WARM);
    const double2 *x = p.x + (size_t)f * (size_t)p.x_stride;
    auto sample = [&](int t) -> double2 {
        if (!p.backward) return px_ext(x, p.L, t);
        return x[E - 1 - t];                        // forward output, stored by ext position
    };
    {
        const double2 x0 = sample(start);           // sosfilt_zi * first sample (exact at t = 0)
        double gr = x0.x, gi = x0.y;
#pragma unroll
        for (int k = 0; k < NSEC; ++k) {
            z1r[k] = c_px.zi[k][0] * gr;
            z1i[k] = c_px.zi[k][0] * gi;
            z2r[k] = c_px.zi[k][1] * gr;
            z2i[k] = c_px.zi[k][1] * gi;
            // the next section sees this section's steady output: input * sum(b) / sum(a)
            const double dc = (c_px.sos[k][0] + c_px.sos[k][1] + c_px.sos[k][2]) /
                              (1.0 + c_px.sos[k][4] + c_px.sos[k][5]);
            gr *= dc;
            gi *= dc;
        }
    }
    double2 *y = p.y + (size_t)f * (size_t)p.y_stride;
    for (int t = start; t < b; ++t) {
        const double2 v = sample(t);
        double vr = v.x, vi = v.y;
#pragma unroll
        for (int k = 0; k < NSEC; ++k) {            // direct form II transposed (scipy _sosfilt)
            const double b0 = c_px.sos[k][0], b1 = c_px.sos[k][1], b2 = c_px.sos[k][2];
            const double a1 = c_px.sos[k][4], a2 = c_px.sos[k][5];
            const double yr = b0 * vr + z1r[k], yi = b0 * vi + z1i[k];
            z1r[k] = b1 * vr - a1 * yr + z2r[k];
            z1i[k] = b1 * vi - a1 * yi + z2i[k];
            z2r[k] = b2 * vr - a2 * yr;
            z2i[k] = b2 * vi - a2 * yi;
            vr = yr;
            vi = yi;
        }
        if (t >= a) {
            if (!p.backward) {
                y[t] = make_double2(vr, vi);
            } else {
                const int n = (E - 1 - t) - PADLEN;          // chunk position of this output
                if (n >= 0 && n < p.L && (n & 1) == 0) y[n >> 1] = make_double2(vr, vi);
            }
        }
    }
}

struct PxWelchParams {
    const void   *x;          // decimated chunks [frames][x_stride]: double2, or wire samples when from_wire
    long long     x_stride;
    int           from_wire;  // 1: fft_ratio 1 -- read the caller's samples (kind / flip), no copy in between
    int           kind, flip, len;
    int           nperseg, hop, nseg, log2N;
    const double *window;     // nperseg taps
    const double2*twiddle;    // [N/2] exp(-i pi m / (N/2)) (host, fp64): pass s uses entry k * (N/2 >> s)
    double2      *work;       // [frames][nseg][2][N] ping-pong
    double       *pow;        // [frames][nseg][N] |X|^2, natural order
};

__device__ __forceinline__ double2 px_fetch(const PxWelchParams &p, int f, int i) {
    if (!p.from_wire) return ((const double2 *)p.x)[(size_t)f * (size_t)p.x_stride + (size_t)i];
    const int src = p.flip ? p.len - 1 - i : i;
    if (p.kind == KIND_U8_RAW) {
        const unsigned char *s = (const unsigned char *)p.x + ((size_t)f * (size_t)p.x_stride + (size_t)src) * 2;
        return make_double2((double)s[0] / 127.5 - 1.0, (double)s[1] / 127.5 - 1.0);
    }
    const float2 v = ((const float2 *)p.x)[(size_t)f * (size_t)p.x_stride + (size_t)src];
    return make_double2((double)v.x, (double)v.y);
}

constexpr int PX_WELCH_NT = 512;

__global__ void __launch_bounds__(PX_WELCH_NT) px_welch_kernel(const PxWelchParams p) {
    __shared__ double red[2][PX_WELCH_NT];
    const int seg = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const int N = 1 << p.log2N;
    const int base = seg * p.hop;
    // detrend='constant': the segment's complex mean
    double sr = 0.0, si = 0.0;
    for (int i = tid; i < p.nperseg; i += PX_WELCH_NT) {
        const double2 v = px_fetch(p, f, base + i);
        sr += v.x;
        si += v.y;
    }
    red[0][tid] = sr;
    red[1][tid] = si;
    __syncthreads();
    for (int o = PX_WELCH_NT / 2; o > 0; o >>= 1) {
        if (tid < o) {
            red[0][tid] += red[0][tid + o];
            red[1][tid] += red[1][tid + o];
        }
        __syncthreads();
    }
    const double mr = red[0][0] / (double)p.nperseg, mi = red[1][0] / (double)p.nperseg;
    double2 *A = p.work + (((size_t)f * p.nseg + seg) * 2) * (size_t)N;
    double2 *B = A + N;
    for (int i = tid; i < N; i += PX_WELCH_NT) {
        double2 v = make_double2(0.0, 0.0);
        if (i < p.nperseg) {
            const double2 s = px_fetch(p, f, base + i);
            const double w = p.window[i];
            v = make_double2((s.x - mr) * w, (s.y - mi) * w);
        }
        A[i] = v;
    }
    __syncthreads();
    // radix-2 Stockham autosort, forward transform
    for (int s = 0; s < p.log2N; ++s) {
        const int Ns = 1 << s;
        for (int j = tid; j < N / 2; j += PX_WELCH_NT) {
            const int k = j & (Ns - 1);
            const double2 u = A[j], v = A[j + N / 2];
            const double2 w = p.twiddle[(size_t)k << (p.log2N - 1 - s)];   // exp(-2 pi i k / (2 Ns))
            const double cs = w.x, sn = w.y;
            const double tr = v.x * cs - v.y * sn, ti = v.x * sn + v.y * cs;
            const int j0 = ((j >> s) << (s + 1)) + k;
            B[j0] = make_double2(u.x + tr, u.y + ti);
            B[j0 + Ns] = make_double2(u.x - tr, u.y - ti);
        }
        __syncthreads();
        double2 *t = A; A = B; B = t;
    }
    double *pw = p.pow + ((size_t)f * p.nseg + seg) * (size_t)N;
    for (int i = tid; i < N; i += PX_WELCH_NT) pw[i] = A[i].x * A[i].x + A[i].y * A[i].y;
}

struct PxRowsParams {
    const double *pow;        // [frames][nseg][N]
    int    frames, nseg, log2N;
    int    W;                 // row width
    int    onesided, os_lo;   // one-sided rows (real input): see FinalizeParams
    double scale;             // 1 / (fs * sum w^2) / nseg
    double alpha;             // < 0: EMA off
    int    linear;
    float *ema_state;
    int    ema_have;
    float *rows;              // [frames][W] or null
    float *ring;
    long long ring_pos;
    int    ring_rows;
};

// one thread per column, frames in order (the EMA recurrence runs along them); without the EMA
// the frames are independent and gridDim.y spreads them (frame f = blockIdx.y, blockIdx.y + gridDim.y, ...)
__global__ void __launch_bounds__(256) px_rows_kernel(const PxRowsParams p) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= p.W) return;
    const int N = 1 << p.log2N;
    int k;                    // natural-order bin feeding this column
    double factor = 1.0;
    if (p.onesided) {
        const int M = N / 2 + 1;
        int kk = col + p.os_lo + (M + 1) / 2;
        if (kk >= M) kk -= M;
        factor = (kk == 0 || kk == N / 2) ? 1.0 : 2.0;
        k = kk;
    } else {
        const int c0 = N / 2 - p.W / 2;
        k = (col + c0 + N / 2) & (N - 1);      // fftshift: column c <- bin (c + N/2) mod N
    }
    bool have = p.ema_have != 0;
    double a = (have && p.alpha >= 0.0) ? (double)p.ema_state[col] : 0.0;
    for (int f = (int)blockIdx.y; f < p.frames; f += (int)gridDim.y) {
        double pw = 0.0;
        for (int s = 0; s < p.nseg; ++s) pw += p.pow[((size_t)f * p.nseg + s) * (size_t)N + k];
        pw *= p.scale * factor;
        if (p.alpha >= 0.0) {
            // the state between launches is fp32 (shared with the fp32 path): round it after every
            // frame, so that a row does not depend on where a launch group or a call ends
            a = (double)(float)(have ? a + p.alpha * (pw - a) : pw);
            have = true;
            pw = a;
        }
        const float out = p.linear ? (float)pw : (float)(20.0 * log10(fabs(pw)));
        if (p.rows) p.rows[(size_t)f * p.W + col] = out;
        if (p.ring && f >= p.frames - p.ring_rows)
            p.ring[(size_t)((p.ring_pos + f) % p.ring_rows) * p.W + col] = out;
    }
    if (p.alpha >= 0.0 && p.frames > 0 && blockIdx.y == 0) p.ema_state[col] = (float)a;
}

// decimated chunk of frame 0 as complex64 (zfb_debug_read_decimated)
__global__ void __launch_bounds__(256) px_to_c64_kernel(const double2 *in, float2 *out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float2((float)in[i].x, (float)in[i].y);
}

}  // namespace zfb
