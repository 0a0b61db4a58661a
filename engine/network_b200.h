// network_b200.h — what engine/main.cpp needs from engine/network_b200.cpp besides the
// reference's own Network.h interface.
#pragma once
#include <cstdint>
#include <string>

#include "GameState.h"
#include "Network.h"
#include "leela_b200.h"

namespace leela_b200 {

// The 32 feature planes of gather_features_policy (value_net = false, Network.cpp:883-1046) or
// gather_features_value (true, Network.cpp:1048-1201), packed one uint32 per board point
// (idx = y*19 + x, bit c = plane c) — the input format of the C ABI.
void pack_features(FastState* state, bool value_net, uint32_t* packed, Network::BoardPlane* ladder);

// 0 (default): planes through the reference's FastBoard queries; 1: through lb2_planes_from_position (the
// library's own board and ladder reader); 2: both, aborting on the first difference
void set_planes_mode(int mode);
long planes_checked();
void set_weights_path(const std::string& path);   // "LB2WGT01" file, see leela_b200/fileio.py
void set_max_outstanding(int n);                   // async policy requests a search thread may have in flight
lb2_ctx* context();

// start-up known-answer test against the answers shipped with the weights file (GTP.cpp:105-125 for this evaluator)
bool self_test(GameState& state);

}  // namespace leela_b200
