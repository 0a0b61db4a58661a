/* Deterministic synthetic network weights — TEST INFRASTRUCTURE (oracle side).
 *
 * The reference's weight translation units NN128.cpp / NNValue.cpp are missing from the
 * snapshot (/root/reference/.MISSING_LARGE_BLOBS:1-3); only their extern declarations exist
 * (Network.cpp:82-137). Parity and throughput therefore run on seeded synthetic weights with
 * exactly those shapes. This header is the single statement of the generator; it is mirrored
 * bit-for-bit in numpy by leela_b200/synth.py (tests/test_synth.py checks the two agree).
 *
 *   array id      : policy conv i (1..13): w = 2(i-1), b = 2(i-1)+1
 *                   value  conv j (1..12): w = 32+2(j-1), b = 33+2(j-1)
 *                   val_ip13: w 56, b 57;  val_ip14: w 58, b 59
 *   element idx   : z = splitmix64_finalize(seed*0x9E3779B97F4A7C15 + (id+1)*0xD1B54A32D192ED03 + idx)
 *                   u = z >> 40 (24 bits); t = 2*(u * 2^-24) - 1   (exact in fp32)
 *   weights       : t * (float(sqrt(6/fan_in)) * gain)            (fp32 multiplies)
 *   biases        : t * 0.1f
 *   gain          : 1 everywhere except the last policy conv (conv13), which uses policy_gain
 */
#ifndef LB2_SYNTH_WEIGHTS_H
#define LB2_SYNTH_WEIGHTS_H
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#define LB2_SYNTH_DEFAULT_SEED 20260001ull
#define LB2_SYNTH_DEFAULT_POLICY_GAIN 2.0f

static inline uint64_t lb2_synth_mix(uint64_t seed, uint32_t id, uint64_t idx) {
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + (uint64_t)(id + 1) * 0xD1B54A32D192ED03ull + idx;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

static inline float lb2_synth_unit(uint64_t seed, uint32_t id, uint64_t idx) {
    float u = (float)(uint32_t)(lb2_synth_mix(seed, id, idx) >> 40) * (1.0f / 16777216.0f);
    return 2.0f * u - 1.0f;
}

/* fill weights: fan_in = c_in*k*k (conv) or n_in (inner product) */
static inline void lb2_synth_fill_weights(float* dst, size_t n, uint64_t seed, uint32_t id,
                                          int fan_in, float gain) {
    float scale = (float)sqrt(6.0 / (double)fan_in) * gain;
    for (size_t i = 0; i < n; i++) dst[i] = lb2_synth_unit(seed, id, i) * scale;
}

static inline void lb2_synth_fill_biases(float* dst, size_t n, uint64_t seed, uint32_t id) {
    for (size_t i = 0; i < n; i++) dst[i] = lb2_synth_unit(seed, id, i) * 0.1f;
}

#endif
