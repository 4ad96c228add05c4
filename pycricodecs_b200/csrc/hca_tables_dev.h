// Device-side launch tables of the HCA kernels (decode, crypt, encode).
//
// Vocabulary: a "unit" is a run of consecutive frames of one stream that one
// lane walks (decode: plus one look-back frame, because the IMDCT overlap needs
// the previous subframe's DCT output; frames are otherwise independent,
// CriCodecs/hca.cpp:1990-1991). 32 units form a "unit block"; step t of a unit
// block is the t-th frame of each of its 32 units, and the intermediate arrays
// are laid out [unit block][step][...][lane] so that the unpack kernel (lane =
// unit) writes and the transform kernel (lane = unit x channel) reads them as
// coalesced 512-byte rows.
#pragma once
#include <cstdint>
#include <vector>

namespace cri {

constexpr int kHcaMaxChannels = 16;

struct HcaStreamDev {
    uint64_t in_off;        // byte offset of frame 0 in the input blob
    uint64_t out_off;       // decode: first PCM sample in the output blob; crypt/encode: frame 0 in the output blob
    uint32_t frame_size;
    uint32_t frame_count;
    uint32_t out_samples;   // decode: samples per channel in the WAV image
    uint32_t delay;         // decode: samples dropped at the start (encoder delay)
    uint32_t cipher;        // index into the cipher-table array (0 = identity)
    uint32_t ath;           // index into the ATH-curve array (0 = all zero)
    uint32_t frame_base;    // decode: index of the stream's frame 0 in the job-wide per-frame arrays (noise generator)
    uint32_t unit_base;     // decode, general kernels: index of the stream's first unit (runs of frames, in order)
    uint8_t channels, total_bands, base_bands, stereo_bands;
    uint8_t bands_per_hfr, hfr_groups, min_res, max_res;
    uint8_t type[kHcaMaxChannels];    // 0 discrete, 1 stereo primary, 2 stereo secondary
    uint8_t coded[kHcaMaxChannels];   // coded band count per channel
    uint8_t joint;          // 1 if any HFR / intensity reconstruction is needed
    uint8_t v3;             // version 3.0 bitstream: extra scalefactors for the HFR scales, delta-coded intensities
    uint8_t noise;          // v3.0 with min_res == 0: resolution-0 bands are rebuilt by the noise generator
    uint8_t pad;
};

struct HcaUnit {
    uint32_t stream;
    uint32_t first;   // first frame of the run
    uint32_t count;   // frames in the run (0 = idle lane)
};

struct HcaUnitBlock {
    uint64_t q_off;    // int16 quantised spectra: [step][channel][subframe][16 chunks][32 lanes][8]   (uint4 units)
    uint64_t g_off;    // fp32 gains / HFR multipliers: [step][channel][128][32 lanes]
    uint64_t i_off;    // intensity nibbles: [step][channel][32 lanes] (uint32)
    uint32_t channels; // max channels of the block's units (strides)
    uint32_t steps;    // max run length + 1 (step 0 = look-back frame)
};

// one transform-kernel lane: a (unit, channel) pair
struct HcaLane {
    uint32_t unit;     // global unit index (block = unit / 32, lane-in-block = unit % 32); 0xFFFFFFFF = idle
    uint32_t channel;
};

struct HcaJob {
    std::vector<HcaStreamDev> streams;
    std::vector<HcaUnit> units;            // padded to a multiple of 32
    std::vector<HcaUnitBlock> blocks;
    std::vector<HcaLane> lanes;            // padded to a multiple of 32
    std::vector<uint8_t> cipher_tables;    // 256 bytes each, [0] identity
    std::vector<uint8_t> ath_tables;       // 128 bytes each, [0] zero
    uint64_t q_bytes = 0, g_bytes = 0, i_bytes = 0, s_bytes = 0;
    uint64_t i_count = 0;                  // entries of the intensity array (the "kept" bytes follow them)
    uint64_t n_bytes = 0;                  // noise generator side arrays (v3.0 streams with min_res == 0), see run_hca
    uint64_t noise_frames = 0;             // frames of the job-wide per-frame arrays
    uint8_t* d_n = nullptr;
    uint32_t scratch_words = 0;
    std::vector<uint64_t> frame_prefix;    // crypt: exclusive prefix of frame counts per stream
    uint64_t* d_frame_prefix = nullptr;
    std::vector<uint64_t> group_prefix;    // crypt, staged kernel: exclusive prefix of frame groups per stream
    uint64_t* d_group_prefix = nullptr;
    std::vector<uint2> group_table;        // crypt, LUT kernel: (stream, first frame) of every group
    uint2* d_group_table = nullptr;
    uint32_t frames_per_group = 0, group_bytes = 0, lut_table = 0, min_frame = 0;
    uint32_t enc_frame_words = 0;
    std::vector<uint16_t> crc_mul;         // encode: [stream][32] CRC chunk multipliers
    uint16_t* d_crc_mul = nullptr;
    uint32_t uniform = 0;                  // decode: see HcaDecodeArgs::uniform
    // decode fast path (hca_fast_kernels.cu)
    std::vector<uint32_t> dec_prefix;      // [n + 1] exclusive prefix of the frames decoded per stream
    uint32_t* d_dec_prefix = nullptr;
    uint32_t run_len = 0, n_runs = 0;      // n_runs == 0: general path
    bool any_joint = false;                // some stream has an intensity-stereo pair or HFR bands
    bool any_pair = false;                 // some stream has an intensity-stereo pair
    uint64_t total_frames = 0, spec_bytes = 0;
    uint8_t* d_spec = nullptr;
    uint8_t* d_s = nullptr;
    uint32_t max_channels = 1, max_steps = 0;

    HcaStreamDev* d_streams = nullptr;
    HcaUnit* d_units = nullptr;
    HcaUnitBlock* d_blocks = nullptr;
    HcaLane* d_lanes = nullptr;
    uint8_t* d_cipher = nullptr;
    uint8_t* d_ath = nullptr;
    uint8_t* d_q = nullptr;
    uint8_t* d_g = nullptr;
    uint8_t* d_i = nullptr;
    uint64_t* d_block_step_prefix = nullptr;  // exclusive prefix of steps per unit block (maps a flat group id to (block, step))
    std::vector<uint64_t> block_step_prefix;
    uint64_t total_groups = 0;
};

}  // namespace cri
