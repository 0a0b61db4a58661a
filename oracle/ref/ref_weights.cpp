/* Definitions of the 54 weight arrays the reference declares extern (Network.cpp:82-137, the
 * USE_BLAS / 128-wide policy shapes and the value net), filled at static-initialisation time
 * by the shared synthetic generator (oracle/synth_weights.h). Stands in for the missing
 * NN128.cpp / NNValue.cpp (.MISSING_LARGE_BLOBS). TEST INFRASTRUCTURE ONLY.
 * Seed / final-conv gain can be overridden with LB2_WEIGHT_SEED / LB2_POLICY_GAIN. */
#include <array>
#include <cstdlib>
#include "../synth_weights.h"

namespace {
uint64_t seed() {
    const char* s = std::getenv("LB2_WEIGHT_SEED");
    return s ? std::strtoull(s, nullptr, 10) : LB2_SYNTH_DEFAULT_SEED;
}
float policy_gain() {
    const char* s = std::getenv("LB2_POLICY_GAIN");
    return s ? std::strtof(s, nullptr) : LB2_SYNTH_DEFAULT_POLICY_GAIN;
}
template <size_t N> std::array<float, N> W(uint32_t id, int fan_in, float gain = 1.0f) {
    std::array<float, N> a;
    lb2_synth_fill_weights(a.data(), N, seed(), id, fan_in, gain);
    return a;
}
template <size_t N> std::array<float, N> B(uint32_t id) {
    std::array<float, N> a;
    lb2_synth_fill_biases(a.data(), N, seed(), id);
    return a;
}
}  // namespace

extern const std::array<float, 76800> conv1_w = W<76800>(0, 800);
extern const std::array<float, 96> conv1_b = B<96>(1);
extern const std::array<float, 110592> conv2_w = W<110592>(2, 864);
extern const std::array<float, 128> conv2_b = B<128>(3);
extern const std::array<float, 147456> conv3_w = W<147456>(4, 1152);
extern const std::array<float, 128> conv3_b = B<128>(5);
extern const std::array<float, 147456> conv4_w = W<147456>(6, 1152);
extern const std::array<float, 128> conv4_b = B<128>(7);
extern const std::array<float, 147456> conv5_w = W<147456>(8, 1152);
extern const std::array<float, 128> conv5_b = B<128>(9);
extern const std::array<float, 147456> conv6_w = W<147456>(10, 1152);
extern const std::array<float, 128> conv6_b = B<128>(11);
extern const std::array<float, 147456> conv7_w = W<147456>(12, 1152);
extern const std::array<float, 128> conv7_b = B<128>(13);
extern const std::array<float, 147456> conv8_w = W<147456>(14, 1152);
extern const std::array<float, 128> conv8_b = B<128>(15);
extern const std::array<float, 147456> conv9_w = W<147456>(16, 1152);
extern const std::array<float, 128> conv9_b = B<128>(17);
extern const std::array<float, 147456> conv10_w = W<147456>(18, 1152);
extern const std::array<float, 128> conv10_b = B<128>(19);
extern const std::array<float, 147456> conv11_w = W<147456>(20, 1152);
extern const std::array<float, 128> conv11_b = B<128>(21);
extern const std::array<float, 147456> conv12_w = W<147456>(22, 1152);
extern const std::array<float, 128> conv12_b = B<128>(23);
extern const std::array<float, 1152> conv13_w = W<1152>(24, 1152, policy_gain());
extern const std::array<float, 1> conv13_b = B<1>(25);
extern const std::array<float, 51200> val_conv1_w = W<51200>(32, 800);
extern const std::array<float, 64> val_conv1_b = B<64>(33);
extern const std::array<float, 36864> val_conv2_w = W<36864>(34, 576);
extern const std::array<float, 64> val_conv2_b = B<64>(35);
extern const std::array<float, 36864> val_conv3_w = W<36864>(36, 576);
extern const std::array<float, 64> val_conv3_b = B<64>(37);
extern const std::array<float, 36864> val_conv4_w = W<36864>(38, 576);
extern const std::array<float, 64> val_conv4_b = B<64>(39);
extern const std::array<float, 36864> val_conv5_w = W<36864>(40, 576);
extern const std::array<float, 64> val_conv5_b = B<64>(41);
extern const std::array<float, 36864> val_conv6_w = W<36864>(42, 576);
extern const std::array<float, 64> val_conv6_b = B<64>(43);
extern const std::array<float, 36864> val_conv7_w = W<36864>(44, 576);
extern const std::array<float, 64> val_conv7_b = B<64>(45);
extern const std::array<float, 36864> val_conv8_w = W<36864>(46, 576);
extern const std::array<float, 64> val_conv8_b = B<64>(47);
extern const std::array<float, 36864> val_conv9_w = W<36864>(48, 576);
extern const std::array<float, 64> val_conv9_b = B<64>(49);
extern const std::array<float, 36864> val_conv10_w = W<36864>(50, 576);
extern const std::array<float, 64> val_conv10_b = B<64>(51);
extern const std::array<float, 36864> val_conv11_w = W<36864>(52, 576);
extern const std::array<float, 64> val_conv11_b = B<64>(53);
extern const std::array<float, 576> val_conv12_w = W<576>(54, 576);
extern const std::array<float, 1> val_conv12_b = B<1>(55);
extern const std::array<float, 92416> val_ip13_w = W<92416>(56, 361);
extern const std::array<float, 256> val_ip13_b = B<256>(57);
extern const std::array<float, 256> val_ip14_w = W<256>(58, 256);
extern const std::array<float, 1> val_ip14_b = B<1>(59);
