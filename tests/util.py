"""shared helpers for the parity tests"""
import numpy as np


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def max_abs(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))


# stated tolerances, against the reference's ggml CPU backend on the same model file + audio
TOL = {
    "mel":  dict(max_abs=2e-4, rel_l2=1e-5),
    "f16":  dict(max_abs=1e-2, rel_l2=2e-3),
    "q8_0": dict(max_abs=0.15, rel_l2=3e-2),   # floor set by ggml's Q8_0 quantisation of activations (not reproduced)
    "q4_0": dict(max_abs=0.15, rel_l2=3e-2),
    "quant_vs_f32_restatement": dict(max_abs=1e-2, rel_l2=2e-3),   # proves block decode + GEMM are exact
}
