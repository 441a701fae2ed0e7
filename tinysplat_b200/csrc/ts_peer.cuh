// Pointer tables of the peer-memory gradient exchange (peer.cu, project.cu): one pointer per
// data-parallel rank, passed to the kernels by value.
#pragma once
#include "ts_common.cuh"

namespace ts {

constexpr int kMaxPeers = 8;          // one NVSwitch domain of a B200 box
constexpr int kBarrierSlots = 16;     // one per pushed chunk of rows + one for the stored shard gradients
constexpr int kCamRowFloats = 32;     // 3x4 view | 4x4 full projection | fx fy | pad

struct PeerPtrs {
    void* p[kMaxPeers];
};

}  // namespace ts
