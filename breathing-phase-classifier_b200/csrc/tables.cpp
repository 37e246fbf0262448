// Host-side constant tables.  See tables.hpp.  float64 construction, rounded once to the dtype the reference's
// library stores (float32 banks, complex64 CQT bases), so device arithmetic starts from the same constants.
#include "tables.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

namespace bpc {

static const double kPi = 3.141592653589793238462643383279502884;

std::vector<double> linspace(double start, double stop, int num) {
    std::vector<double> y(num);
    if (num == 1) { y[0] = start; return y; }
    const double step = (stop - start) / double(num - 1);
    for (int i = 0; i < num; ++i) y[i] = double(i) * step + start;
    y[num - 1] = stop;
    return y;
}

std::vector<double> hann_periodic(int n) {
    // scipy general_cosine(n+1, [0.5, 0.5], sym=True)[:-1]: fac = linspace(-pi, pi, n+1); w = 0.5 + 0.5*cos(fac)
    std::vector<double> fac = linspace(-kPi, kPi, n + 1);
    std::vector<double> w(n);
    for (int i = 0; i < n; ++i) w[i] = 0.5 + 0.5 * std::cos(fac[i]);
    return w;
}

std::vector<double> hamming_sym(int n) {
    // numpy.hamming: n_ = arange(1-M, M, 2); 0.54 + 0.46*cos(pi*n_/(M-1))
    std::vector<double> w(n);
    for (int i = 0; i < n; ++i) w[i] = 0.54 + 0.46 * std::cos(kPi * double(1 - n + 2 * i) / double(n - 1));
    return w;
}

// ---- Slaney mel scale (librosa.hz_to_mel / mel_to_hz, htk=False)
static double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

SparseBank mel_bank(int sr, int n_fft, int n_mels, double fmin, double fmax) {
    SparseBank b;
    b.rows = n_mels;
    b.cols = 1 + n_fft / 2;
    b.dense.assign(size_t(b.rows) * b.cols, 0.f);
    const double val = 1.0 / (double(n_fft) * (1.0 / double(sr)));        // numpy.fft.rfftfreq
    std::vector<double> mels = linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2);
    std::vector<double> mel_f(n_mels + 2);
    for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz(mels[i]);
    for (int i = 0; i < n_mels; ++i) {
        const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        for (int k = 0; k < b.cols; ++k) {
            const double fk = double(k) * val;
            const double lower = -(mel_f[i] - fk) / fd0;
            const double upper = (mel_f[i + 2] - fk) / fd1;
            const float w32 = float(std::max(0.0, std::min(lower, upper)));   // weights[i] = ... (float32 store)
            b.dense[size_t(i) * b.cols + k] = float(double(w32) * enorm);      // weights *= enorm (in-place, f64 math)
        }
    }
    b.start.assign(b.rows, 0);
    b.count.assign(b.rows, 0);
    for (int i = 0; i < b.rows; ++i) {
        int lo = -1, hi = -1;
        for (int k = 0; k < b.cols; ++k)
            if (b.dense[size_t(i) * b.cols + k] != 0.f) { if (lo < 0) lo = k; hi = k; }
        if (lo >= 0) { b.start[i] = lo; b.count[i] = hi - lo + 1; }
        b.width = std::max(b.width, b.count[i]);
    }
    b.w.assign(size_t(b.rows) * std::max(1, b.width), 0.f);
    for (int i = 0; i < b.rows; ++i)
        for (int j = 0; j < b.count[i]; ++j) b.w[size_t(i) * b.width + j] = b.dense[size_t(i) * b.cols + b.start[i] + j];
    return b;
}

std::vector<float> dct2_ortho(int n_out, int n_in) {
    // scipy.fftpack.dct(type=2, norm='ortho'): X[k] = f_k * 2 * sum_n x[n] cos(pi k (2n+1) / (2N))
    std::vector<float> d(size_t(n_out) * n_in);
    for (int k = 0; k < n_out; ++k) {
        const double f = (k == 0) ? std::sqrt(1.0 / (4.0 * n_in)) : std::sqrt(1.0 / (2.0 * n_in));
        for (int n = 0; n < n_in; ++n)
            d[size_t(k) * n_in + n] = float(2.0 * f * std::cos(kPi * double(k) * double(2 * n + 1) / double(2 * n_in)));
    }
    return d;
}

std::vector<double> tuning_edges() { return linspace(-0.5, 0.5, kNumTunings + 1); }

std::vector<float> chroma_bank(int sr, int n_fft, double tuning) {
    const int n_chroma = 12;
    const int nb = n_fft;                                 // columns before the final slice
    // frequencies = linspace(0, sr, n_fft, endpoint=False)[1:]
    std::vector<double> frqbins(nb);
    const double a440 = 440.0 * std::pow(2.0, tuning / double(n_chroma));
    const double step = double(sr) / double(n_fft);
    for (int k = 1; k < nb; ++k) frqbins[k] = double(n_chroma) * std::log2((double(k) * step) / (a440 / 16.0));
    frqbins[0] = frqbins[1] - 1.5 * n_chroma;
    std::vector<double> bw(nb);
    for (int k = 0; k + 1 < nb; ++k) bw[k] = std::max(frqbins[k + 1] - frqbins[k], 1.0);
    bw[nb - 1] = 1.0;
    const double half = std::round(double(n_chroma) / 2.0);
    std::vector<double> wts(size_t(n_chroma) * nb);
    for (int k = 0; k < nb; ++k) {
        double col2 = 0.0;
        for (int c = 0; c < n_chroma; ++c) {
            double D = frqbins[k] - double(c);
            D = D + half + 10.0 * n_chroma;
            D = D - std::floor(D / n_chroma) * n_chroma;           // numpy.remainder for a positive divisor
            D -= half;
            const double v = std::exp(-0.5 * std::pow(2.0 * D / bw[k], 2.0));
            wts[size_t(c) * nb + k] = v;
            col2 += v * v;
        }
        double len = std::sqrt(col2);
        if (len < 2.2250738585072014e-308) len = 1.0;              // util.normalize threshold = tiny(float64)
        const double oct = std::exp(-0.5 * std::pow((frqbins[k] / n_chroma - 5.0) / 2.0, 2.0));
        for (int c = 0; c < n_chroma; ++c) wts[size_t(c) * nb + k] = wts[size_t(c) * nb + k] / len * oct;
    }
    const int cols = 1 + n_fft / 2;
    std::vector<float> out(size_t(n_chroma) * cols);
    for (int c = 0; c < n_chroma; ++c) {                           // np.roll(wts, -3, axis=0): out[c] = wts[(c+3)%12]
        const int src = (c + 3) % n_chroma;
        for (int k = 0; k < cols; ++k) out[size_t(c) * cols + k] = float(wts[size_t(src) * nb + k]);
    }
    return out;
}

// ---- Kaiser half-band (oracle/librosa_shim/librosa/_core.py::default_halfband; scipy.signal.firwin semantics)
static double bessel_i0(double x) { return std::cyl_bessel_i(0.0, x); }

std::vector<double> halfband_taps(int numtaps, double atten_db) {
    double beta;
    if (atten_db > 50) beta = 0.1102 * (atten_db - 8.7);
    else if (atten_db > 21) beta = 0.5842 * std::pow(atten_db - 21, 0.4) + 0.07886 * (atten_db - 21);
    else beta = 0.0;
    const double alpha = 0.5 * (numtaps - 1);
    std::vector<double> h(numtaps);
    double s = 0.0;
    for (int i = 0; i < numtaps; ++i) {
        const double m = double(i) - alpha;
        const double x = 0.5 * m;                                   // right * m with right = 0.5, left = 0
        const double sinc = (x == 0.0) ? 1.0 : std::sin(kPi * x) / (kPi * x);
        const double r = (double(i) - alpha) / alpha;
        const double win = bessel_i0(beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / bessel_i0(beta);
        h[i] = 0.5 * sinc * win;
        s += h[i];
    }
    for (auto& v : h) v /= s;                                       // firwin scale (pass_zero) == unit DC gain
    double s2 = 0.0;
    for (auto v : h) s2 += v;
    for (auto& v : h) v /= s2;                                      // default_halfband: taps / sum(taps)
    return h;
}

// ---- sample-rate converter (oracle/resample.py): Kaiser-windowed sinc, 150 dB, transition band [0.9125, 1] of the
// lower of the two Nyquist rates (soxr HQ's pass band), as a polyphase table.
ResampleFilter resample_filter(int sr_in, int sr_out) {
    ResampleFilter f;
    if (sr_in <= 0 || sr_out <= 0) return f;
    int a = sr_in, b = sr_out;
    while (b) { const int t = a % b; a = b; b = t; }
    f.p = sr_out / a;
    f.q = sr_in / a;
    const double r = std::min(1.0, double(f.p) / double(f.q));
    const double fc = 0.5 * 0.95625 * r;                            // cut-off in cycles per input sample
    const double delta = 0.5 * 0.0875 * r;                          // transition width
    const double atten = 150.0, beta = 0.1102 * (atten - 8.7);
    const int ntaps = (int)std::ceil((atten - 7.95) / (14.36 * delta));
    f.half = (ntaps + 1) / 2;
    const double i0b = bessel_i0(beta);
    f.tab.assign((size_t)f.p * 2 * f.half, 0.0);
    for (int ph = 0; ph < f.p; ++ph) {
        double* row = f.tab.data() + (size_t)ph * 2 * f.half;
        double s = 0.0;
        for (int j = 0; j < 2 * f.half; ++j) {
            const double u = double(ph) / double(f.p) + double(f.half - 1 - j);
            const double xr = u / double(f.half);
            double w = 0.0;
            if (std::fabs(xr) <= 1.0) w = bessel_i0(beta * std::sqrt(std::max(0.0, 1.0 - xr * xr))) / i0b;
            const double z = 2.0 * fc * u;
            const double sinc = (z == 0.0) ? 1.0 : std::sin(kPi * z) / (kPi * z);
            row[j] = 2.0 * fc * sinc * w;
            s += row[j];
        }
        for (int j = 0; j < 2 * f.half; ++j) row[j] /= s;
    }
    return f;
}

void fft_inplace(std::vector<std::complex<double>>& a) {
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const double ang = -2.0 * kPi * double(k) / double(len);
                const std::complex<double> w(std::cos(ang), std::sin(ang));
                const std::complex<double> u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
}

CqtBasisEll cqt_basis(int sr, double tuning, std::vector<std::complex<float>>* dense_out) {
    // librosa.vqt(fmin=C1, n_bins=252, bins_per_octave=36, tuning, filter_scale=1, norm=1, sparsity=0.01, 'hann')
    const int bpo = kCqtBinsPerOct, n_oct = kCqtOctaves, n_bins = kCqtBins, n_fft = 512;
    const double c1 = 440.0 * std::pow(2.0, (24.0 - 69.0) / 12.0);
    const double fmin = c1 * std::pow(2.0, tuning / double(bpo));
    std::vector<double> ratios;
    for (int o = 0; o < n_oct; ++o)
        for (int r = 0; r < bpo; ++r) ratios.push_back(std::pow(2.0, double(o)) * std::pow(2.0, double(r) / bpo));
    std::sort(ratios.begin(), ratios.end());
    std::vector<double> freqs(n_bins), logf(n_bins), alpha(n_bins), lengths(n_bins);
    for (int k = 0; k < n_bins; ++k) { freqs[k] = ratios[k] * fmin; logf[k] = std::log2(freqs[k]); }
    for (int k = 0; k < n_bins; ++k) {                               // filters.relative_bandwidth
        double b;
        if (k == 0) b = 1.0 / (logf[1] - logf[0]);
        else if (k == n_bins - 1) b = 1.0 / (logf[k] - logf[k - 1]);
        else b = 2.0 / (logf[k + 1] - logf[k - 1]);
        const double p = std::pow(2.0, 2.0 / b);
        alpha[k] = (p - 1.0) / (p + 1.0);
        lengths[k] = (1.0 / alpha[k]) * double(sr) / freqs[k];       // Q * sr / freqs (gamma = 0)
    }
    CqtBasisEll out;
    out.sqrt_len.resize(n_bins);
    for (int k = 0; k < n_bins; ++k) out.sqrt_len[k] = std::sqrt(lengths[k]);
    out.col.assign(size_t(bpo) * kCqtEllWidth, int16_t(-1));
    out.re.assign(size_t(bpo) * kCqtEllWidth, 0.f);
    out.im.assign(size_t(bpo) * kCqtEllWidth, 0.f);
    if (dense_out) dense_out->assign(size_t(bpo) * (n_fft / 2 + 1), std::complex<float>(0.f, 0.f));

    for (int r = 0; r < bpo; ++r) {                                   // top octave at the full rate
        const int k = n_bins - bpo + r;
        const double ilen = lengths[k], freq = freqs[k];
        const double lo = std::floor(-ilen / 2.0), hi = std::floor(ilen / 2.0);
        const int n = int(std::ceil(hi - lo));
        std::vector<std::complex<double>> sig(n);
        std::vector<double> win = hann_periodic(n);
        double l1 = 0.0;
        for (int i = 0; i < n; ++i) {
            const double ang = (lo + double(i)) * 2 * kPi * freq / double(sr);
            sig[i] = std::complex<double>(std::cos(ang), std::sin(ang)) * win[i];
            l1 += std::abs(sig[i]);
        }
        if (l1 < 2.2250738585072014e-308) l1 = 1.0;
        std::vector<std::complex<double>> padded(n_fft, std::complex<double>(0.0, 0.0));
        const int lpad = (n_fft - n) / 2;
        const double scale = ilen / double(n_fft);
        for (int i = 0; i < n; ++i) {
            const std::complex<double> v = sig[i] / l1;
            const std::complex<float> c64(float(v.real()), float(v.imag()));            // np.asarray(..., complex64)
            const std::complex<double> s(double(c64.real()) * scale, double(c64.imag()) * scale);
            const std::complex<float> c64b(float(s.real()), float(s.imag()));           // basis *= lengths / n_fft
            padded[lpad + i] = std::complex<double>(c64b.real(), c64b.imag());
        }
        fft_inplace(padded);
        const int nb = n_fft / 2 + 1;
        std::vector<double> mags(nb), sorted(nb);
        double norm = 0.0;
        for (int j = 0; j < nb; ++j) { mags[j] = std::abs(padded[j]); norm += mags[j]; }
        sorted = mags;
        std::sort(sorted.begin(), sorted.end());
        double cum = 0.0, thr = sorted[nb - 1];
        for (int j = 0; j < nb; ++j) {                                 // util.sparsify_rows(quantile=0.01)
            cum += sorted[j] / norm;
            if (!(cum < 0.01)) { thr = sorted[j]; break; }
        }
        int w = 0;
        for (int j = 0; j < nb; ++j) {
            if (mags[j] >= thr) {
                const std::complex<float> v(float(padded[j].real()), float(padded[j].imag()));
                if (dense_out) (*dense_out)[size_t(r) * nb + j] = v;
                if (w < kCqtEllWidth) {
                    out.col[size_t(r) * kCqtEllWidth + w] = int16_t(j);
                    out.re[size_t(r) * kCqtEllWidth + w] = v.real();
                    out.im[size_t(r) * kCqtEllWidth + w] = v.imag();
                }
                ++w;
            }
        }
        if (w > kCqtEllWidth) { out.col.clear(); return out; }         // caller treats as fatal
    }
    return out;
}

}  // namespace bpc
