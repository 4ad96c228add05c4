// Kernel argument blocks and launchers of the HCA kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "formats.h"
#include "hca_tables_dev.h"

namespace cri {

struct HcaDecodeArgs {
    const uint8_t* in;
    uint8_t* out;
    const HcaStreamDev* streams;
    const HcaUnit* units;
    const HcaLane* lanes;
    const uint8_t* cipher;      // [n][256]
    const uint8_t* ath;         // [n][128]
    uint4* quant;               // [group][channel][8][16][32] x 8 int16
    float4* gain;               // [group][channel][32][32] x 4 fp32
    uint32_t* inten;            // [group][channel][32]
    int32_t* status;
    uint64_t total_groups;      // unit blocks x steps
    uint32_t steps;             // frames per unit + 1 (step 0 = look-back frame)
    uint32_t max_channels;
};

// `mid` (optional) is recorded between the unpack and the transform kernel.
void launch_hca_decode(const HcaDecodeArgs& a, uint32_t n_lanes, cudaStream_t s, uint64_t* launches, cudaEvent_t mid);
uint32_t hca_imdct_lane_granule();

}  // namespace cri
