"""The arithmetic behind precise mode (lb2_set_option("precise", 1)), checked on the CPU with numpy: an fp32 value
split as hi = fp16(x), lo = fp16(x - hi) keeps ~22 bits, and the three-term product hi*Wh + hi*Wl + lo*Wh accumulated
in fp32 reproduces an fp32 3x3 convolution layer to ~1e-6 where plain fp16 operands give ~1e-3 — the factor the GPU
path shows end to end (tests/test_gpu_parity.py::test_precise_mode_correctness_set_within_2e_4)."""
import numpy as np


def split(x):
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def test_split_keeps_22_bits():
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(100000) * np.exp(rng.uniform(-6, 2, 100000))).astype(np.float32)
    hi, lo = split(x)
    rel = np.abs((hi + lo) - x) / np.maximum(np.abs(x), 1e-30)
    big = np.abs(x) > 2.0 ** -3          # lo stays a normal fp16 number (|lo| >= 2^-14) down to about here
    assert rel[big].max() < 2.0 ** -21
    # everywhere: half an fp16 subnormal quantum (2^-25) or 2^-22 relative, whichever is larger
    assert np.all(np.abs((hi + lo) - x) <= np.maximum(2.0 ** -25, 2.0 ** -22 * np.abs(x)))


def test_three_term_product_matches_fp32_layer():
    rng = np.random.default_rng(2)
    c_in, c_out, rows = 128, 64, 400
    a = rng.standard_normal((rows, 9 * c_in)).astype(np.float32)                      # im2col rows of ELU-like activations
    a = np.where(a > 0, a, np.expm1(a)).astype(np.float32)
    w = rng.uniform(-1, 1, (9 * c_in, c_out)).astype(np.float32) * np.float32(np.sqrt(6.0 / (9 * c_in)))
    want = a.astype(np.float64) @ w.astype(np.float64)
    ah, al = split(a)
    wh, wl = split(w)
    plain = ah @ wh                                                                    # fp16 operands, fp32 accumulation
    precise = ah @ wh + ah @ wl + al @ wh
    e_plain, e_precise = np.abs(plain - want).max(), np.abs(precise - want).max()
    assert 1e-4 < e_plain < 1e-2
    assert e_precise < 5e-6 and e_precise < e_plain / 200
    # the first layer's inputs are 0/1: no residual, two terms suffice
    b = (rng.random((rows, 25 * 32)) < 0.2).astype(np.float32)
    w1 = rng.uniform(-1, 1, (25 * 32, c_out)).astype(np.float32) * np.float32(np.sqrt(6.0 / 800))
    w1h, w1l = split(w1)
    assert split(b)[1].max() == 0.0
    assert np.abs((b @ w1h + b @ w1l) - b.astype(np.float64) @ w1.astype(np.float64)).max() < 5e-6
