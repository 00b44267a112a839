// Taper (window) design and its preview spectrum on the device.
//
// Replaces, for the closed-form families of the reference's taper dialog
// (FFTTaperingControl.taper_list, pypanadapter_spectrum.py:1222-1243), the
// scipy.signal.get_window(AppState.fft_tapering, 51) call of ShowCurve
// (S:1366) and its preview
//     fft = np.fft.fft(taperdata, 2048) / (len(taperdata) / 2.0)
//     taperfft = 20 * np.log10(np.abs(fft / np.max(np.abs(fft))))      (S:1374-1376)
// Formulas follow scipy/signal/windows/_windows.py (1.18.1) evaluated in fp64:
// a periodic window (fftbins=True, what get_window builds) is the symmetric one
// of n + 1 points without its last sample.  chebwin / dpss / slepian need a
// polynomial design or an eigenproblem and stay on the host (scipy).
#pragma once
#include "zfb_common.cuh"

namespace zfb {

enum TaperKind {
    TAPER_BOXCAR = 0, TAPER_TRIANG, TAPER_BARTLETT, TAPER_HANN, TAPER_HAMMING, TAPER_BLACKMAN, TAPER_NUTTALL,
    TAPER_BLACKMANHARRIS, TAPER_FLATTOP, TAPER_BOHMAN, TAPER_BARTHANN, TAPER_PARZEN, TAPER_KAISER, TAPER_GAUSSIAN,
    TAPER_GENERAL_GAUSSIAN, TAPER_EXPONENTIAL, TAPER_TUKEY, TAPER_COUNT
};

struct TaperParams {
    int    kind;
    int    n;          // taps wanted
    int    M;          // length of the symmetric design: n + 1 when periodic
    double p0, p1;
    double *out;       // n doubles
};

__device__ __forceinline__ double taper_cos_sum(const double *a, int na, int i, int M) {
    // general_cosine: fac = linspace(-pi, pi, M); w = sum_k a_k cos(k fac)
    const double fac = -1.0 + 2.0 * (double)i / (double)(M - 1);          // in units of pi
    double w = 0.0;
    for (int k = 0; k < na; ++k) w += a[k] * cospi((double)k * fac);
    return w;
}

__global__ void __launch_bounds__(256) taper_kernel(const TaperParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const int M = p.M;
    const double n = (double)i;
    double w = 1.0;
    if (p.n == 1) {            // scipy's _len_guards: a one-point window is 1
        p.out[i] = 1.0;
        return;
    }
    switch (p.kind) {
        case TAPER_BOXCAR: w = 1.0; break;
        case TAPER_TRIANG: {
            // n' = 1 .. (M+1)//2 mirrored; even M: (2n'-1)/M, odd M: 2n'/(M+1)
            const int half = (M + 1) / 2;
            const int j = i < half ? i + 1 : M - i;
            w = (M % 2 == 0) ? (2.0 * j - 1.0) / (double)M : 2.0 * j / ((double)M + 1.0);
            break;
        }
        case TAPER_BARTLETT:
            w = (n <= (M - 1) / 2.0) ? 2.0 * n / (M - 1) : 2.0 - 2.0 * n / (M - 1);
            break;
        case TAPER_HANN: { const double a[2] = {0.5, 0.5}; w = taper_cos_sum(a, 2, i, M); break; }
        case TAPER_HAMMING: { const double a[2] = {0.54, 0.46}; w = taper_cos_sum(a, 2, i, M); break; }
        case TAPER_BLACKMAN: { const double a[3] = {0.42, 0.50, 0.08}; w = taper_cos_sum(a, 3, i, M); break; }
        case TAPER_NUTTALL: {
            const double a[4] = {0.3635819, 0.4891775, 0.1365995, 0.0106411};
            w = taper_cos_sum(a, 4, i, M);
            break;
        }
        case TAPER_BLACKMANHARRIS: {
            const double a[4] = {0.35875, 0.48829, 0.14128, 0.01168};
            w = taper_cos_sum(a, 4, i, M);
            break;
        }
        case TAPER_FLATTOP: {
            const double a[5] = {0.21557895, 0.41663158, 0.277263158, 0.083578947, 0.006947368};
            w = taper_cos_sum(a, 5, i, M);
            break;
        }
        case TAPER_BOHMAN: {
            if (i == 0 || i == M - 1) { w = 0.0; break; }
            const double fac = fabs(-1.0 + 2.0 * n / (M - 1));
            w = (1.0 - fac) * cospi(fac) + sinpi(fac) / 3.14159265358979323846;
            break;
        }
        case TAPER_BARTHANN: {
            const double fac = fabs(n / (M - 1) - 0.5);
            w = 0.62 - 0.48 * fac + 0.38 * cospi(2.0 * fac);
            break;
        }
        case TAPER_PARZEN: {
            const double x = fabs(n - (M - 1) / 2.0);
            const double r = x / (M / 2.0);
            w = (x <= (M - 1) / 4.0) ? 1.0 - 6.0 * r * r + 6.0 * r * r * r : 2.0 * (1.0 - r) * (1.0 - r) * (1.0 - r);
            break;
        }
        case TAPER_KAISER: {
            const double alpha = (M - 1) / 2.0;
            const double t = (n - alpha) / alpha;
            w = cyl_bessel_i0(p.p0 * sqrt(fmax(0.0, 1.0 - t * t))) / cyl_bessel_i0(p.p0);
            break;
        }
        case TAPER_GAUSSIAN: {
            const double x = n - (M - 1) / 2.0;
            w = exp(-x * x / (2.0 * p.p0 * p.p0));
            break;
        }
        case TAPER_GENERAL_GAUSSIAN: {
            const double x = n - (M - 1) / 2.0;
            w = exp(-0.5 * pow(fabs(x / p.p1), 2.0 * p.p0));
            break;
        }
        case TAPER_EXPONENTIAL:          // p0 = center, p1 = tau
            w = exp(-fabs(n - p.p0) / p.p1);
            break;
        case TAPER_TUKEY: {
            const double al = p.p0;
            if (al <= 0.0) { w = 1.0; break; }
            if (al >= 1.0) { const double a[2] = {0.5, 0.5}; w = taper_cos_sum(a, 2, i, M); break; }
            const int width = (int)floor(al * (M - 1) / 2.0);
            if (i <= width) w = 0.5 * (1.0 + cospi(-1.0 + 2.0 * n / al / (M - 1)));
            else if (i < M - width - 1) w = 1.0;
            else w = 0.5 * (1.0 + cospi(-2.0 / al + 1.0 + 2.0 * n / al / (M - 1)));
            break;
        }
        default: w = 1.0;
    }
    p.out[i] = w;
}

// |DFT(taper zero-padded to nfft)| in fp64: one thread per bin (ntaps is tiny: 51 in the dialog)
__global__ void __launch_bounds__(256) taper_spectrum_kernel(const double *taper, int ntaps, int nfft, double *mag) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nfft) return;
    double re = 0.0, im = 0.0;
    for (int n = 0; n < ntaps; ++n) {
        // phase -2 pi k n / nfft, reduced exactly in integers
        const long long r = ((long long)k * n) % nfft;
        double s, c;
        sincospi(-2.0 * (double)r / (double)nfft, &s, &c);
        re += taper[n] * c;
        im += taper[n] * s;
    }
    mag[k] = sqrt(re * re + im * im);
}

// 20*log10(mag / max(mag)) (the reference divides by len/2 first: it cancels in the ratio)
__global__ void __launch_bounds__(256) taper_db_kernel(const double *mag, int nfft, float *out_db) {
    __shared__ double red[256];
    double m = 0.0;
    for (int k = threadIdx.x; k < nfft; k += blockDim.x) m = fmax(m, mag[k]);
    red[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    const double mx = red[0];
    for (int k = threadIdx.x; k < nfft; k += blockDim.x) out_db[k] = (float)(20.0 * log10(mag[k] / mx));
}

}  // namespace zfb
