// Host-side constant tables of the precompute path (filterbanks, windows, DCT matrices, chroma banks, CQT bases).
// Everything here restates the *published* construction used by librosa 0.10.2 (the library the reference calls
// at src/precompute/process.py:32,43,52,53,74 and methods.py:137); nothing is read from Python at run time.
#pragma once
#include <cstdint>
#include <vector>
#include <complex>
#include "../../include/bpc.h"

namespace bpc {

constexpr int kNumTunings = 100;     // pitch_tuning histogram bins (resolution 0.01)
constexpr int kCqtBinsPerOct = 36;
constexpr int kCqtOctaves = 7;
constexpr int kCqtBins = kCqtBinsPerOct * kCqtOctaves;   // 252
constexpr int kCqtEllWidth = 20;     // max non-zeros per sparsified basis row (measured: <= 16)
constexpr int kHalfbandTaps = 127;

struct SparseBank {                  // triangular mel bank in band form: row m covers bins [start, start+count)
    int rows = 0, cols = 0, width = 0;           // width = max count
    std::vector<float> dense;                    // [rows, cols]
    std::vector<int32_t> start, count;           // [rows]
    std::vector<float> w;                        // [rows, width], zero padded
};

struct CqtBasisEll {                 // one tuning: 36 rows, ELL format over rfft bins of an n_fft=512 frame
    std::vector<int16_t> col;        // [36, kCqtEllWidth], -1 = empty
    std::vector<float> re, im;       // [36, kCqtEllWidth]
    std::vector<double> sqrt_len;    // [252] sqrt(lengths) at the full rate (vqt `V /= sqrt(lengths)`)
};

// numpy.linspace(start, stop, num) semantics (endpoint=True)
std::vector<double> linspace(double start, double stop, int num);
std::vector<double> hann_periodic(int n);                    // scipy.signal.get_window('hann', n, fftbins=True)
std::vector<double> hamming_sym(int n);                      // numpy.hamming(n)
SparseBank mel_bank(int sr, int n_fft, int n_mels, double fmin, double fmax);   // librosa.filters.mel (slaney)
std::vector<float> dct2_ortho(int n_out, int n_in);          // rows k < n_out of the ortho DCT-II of length n_in
std::vector<float> chroma_bank(int sr, int n_fft, double tuning);               // librosa.filters.chroma [12, 1+n_fft/2]
std::vector<double> tuning_edges();                          // numpy.linspace(-0.5, 0.5, 101)
std::vector<double> halfband_taps(int numtaps = kHalfbandTaps, double atten_db = 150.0);
CqtBasisEll cqt_basis(int sr, double tuning, std::vector<std::complex<float>>* dense_out = nullptr);
void fft_inplace(std::vector<std::complex<double>>& a);      // radix-2, power-of-two length, forward

// Polyphase table of the sample-rate converter (oracle/resample.py::polyphase_table): p phases x 2 half taps.
struct ResampleFilter {
    int p = 0, q = 0, half = 0;      // sr_out / sr_in = p / q in lowest terms; taps span (-half, half] input samples
    std::vector<double> tab;         // [p, 2 half]: tab[f][j] = h(f / p + half - 1 - j), every phase scaled to unit DC gain
};
ResampleFilter resample_filter(int sr_in, int sr_out);

}  // namespace bpc
