"""The reference's own float32 noise on the five near-cancelling scalars (indices 10, 29, 30, 33, 34).

methods.py:65 scipy.stats.skew(centroid) runs on a float64 array, methods.py:99-100 skew / kurtosis run on the float32
waveform IN float32 (scipy keeps the input dtype: means and central moments are float32 reductions over 16000 samples),
and methods.py:105-110 np.correlate(y, y) accumulates float32 products in float32.  This script evaluates each of them
the reference's way and in float64 on the same segments (8 golden + N synthetic) and prints the largest absolute and
relative differences: the floor below which "rtol 1e-4 against the reference" asks for more digits than the reference
itself has.  CPU only; DESIGN.md section 2 quotes its output.
"""
import os
import sys

import numpy as np
import scipy.stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pipeline as P  # noqa: E402


def main(n_synth=24):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden_segments.npz"))
    ys = [q.astype(np.float32) / np.float32(32768.0) for q in gold["pcm16"]]
    ys += [P.synth_segment(1000 + i) for i in range(n_synth)]
    names = {10: "skew(centroid)", 29: "skew(y)", 30: "kurtosis(y)", 33: "ac[160]/ac[0]", 34: "ac[320]/ac[0]"}
    worst_abs = {k: 0.0 for k in names}
    worst_rel = {k: 0.0 for k in names}
    worst_excess = {k: -1.0 for k in names}
    for y in ys:
        d = {}
        sc = P.scalar_features(y, debug=d)
        y64 = y.astype(np.float64)
        ref = {
            # the centroid array is float64 already; its float32 ingredient is |X| (complex64 -> float32 magnitudes)
            10: float(scipy.stats.skew(d["centroid"].astype(np.float64))),
            29: float(scipy.stats.skew(y64)),
            30: float(scipy.stats.kurtosis(y64)),
        }
        ac0 = float(np.dot(y64, y64))
        ref[33] = float(np.dot(y64[:-160], y64[160:]) / ac0)
        ref[34] = float(np.dot(y64[:-320], y64[320:]) / ac0)
        for k in names:
            a = abs(float(sc[k]) - ref[k])
            worst_abs[k] = max(worst_abs[k], a)
            worst_rel[k] = max(worst_rel[k], a / max(abs(ref[k]), 1e-30))
            worst_excess[k] = max(worst_excess[k], a - (1e-4 * abs(ref[k]) + 2e-6))
    print(f"{len(ys)} segments; reference arithmetic (float32 where the reference is) vs float64")
    for k, nm in names.items():
        print(f"  [{k:2d}] {nm:16s} max |f32 - f64| = {worst_abs[k]:.2e}   max rel = {worst_rel[k]:.2e}   "
              f"max(|d| - (1e-4 |ref| + 2e-6)) = {worst_excess[k]:+.2e}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 24)
