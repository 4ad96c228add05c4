// Kernel argument blocks and launchers of the HCA kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "formats.h"
#include "hca_tables_dev.h"

namespace cri {

// Intermediate arrays are indexed by frame slot = unit * steps + step (step 0 = look-back frame of the unit).
struct HcaDecodeArgs {
    const uint8_t* in;
    uint8_t* out;
    const HcaStreamDev* streams;
    const HcaUnit* units;       // padded to a multiple of 32
    const uint8_t* cipher;      // [n][256]
    const uint8_t* ath;         // [n][128]
    uint32_t* scratch;          // [slot][scratch_words] aligned, deciphered, big-endian frame words
    uint4* quant;               // [slot][channel][8 subframes][16] x 8 int16 quantised spectra
    float* gain;                // [slot][channel][128] gain per coded band, HFR multiplier per reconstructed band
    uint32_t* inten;            // [slot][channel] 8 intensity nibbles
    uint8_t* carry;             // [slot][channel] k > 0: nibbles k..7 were not coded and keep the previous frame's values
                                // (hca.cpp:1368-1372, 1410-1412 + its caller :1185); resolved by the intensity scan
    uint32_t carry_scan;        // some stream has an intensity pair: run the scan between unpack and transform
    int32_t* status;
    uint64_t total_groups;      // unit blocks (32 units) x steps: one unpack warp each
    uint32_t n_units;
    uint32_t steps;             // frames per unit + 1
    uint32_t max_channels;
    uint32_t scratch_words;     // multiple of 4
    uint32_t uniform;           // 0 mixed batch; 1 every stream is mono, 2 stereo: all bands coded, no HFR / intensity
    // fast path (hca_fast_kernels.cu): frames flattened over the streams, cut into runs of run_len frames;
    // scratch is then indexed by flattened frame and has one spare row
    const uint32_t* dec_prefix; // [n_streams + 1] exclusive prefix of the frames decoded per stream
    float4* spec;               // [transform warp][frame in run][subframe][32 chunks][32 lanes] dequantised spectra
    uint64_t total_frames;
    uint32_t n_streams;
    uint32_t run_len;
    uint32_t n_runs;            // 0 = not on the fast path
    uint32_t joint;             // fast path: some stream has an intensity-stereo pair or HFR bands
    uint32_t run_base;          // fast path, pipelined launches: this launch covers runs [run_base, run_base + run_count)
    uint32_t run_count;
    // v3.0 noise generator (null when no stream needs it): the unpack kernel leaves, per frame slot and channel, the
    // band classes and the number of generator draws per subframe; a scan turns the per-frame totals into the
    // generator state at the start of every frame (the state runs through the whole stream, hca.cpp:1602-1635)
    uint8_t* sfres;             // [slot][channel][128] scalefactor | 0x40 noise band | 0x80 valid band
    uint32_t* draws;            // [slot][channel] draws per subframe
    uint32_t* frame_draws;      // [job frame] draws per subframe, all channels
    uint32_t* frame_state;      // [job frame] generator state at the start of the frame
    uint64_t one2;              // {1.0f, 1.0f}: the multiplier of the transform kernel's two-wide sums (hca_sum2)
    uint32_t force_careful;     // tests: take the end-of-frame-checked reader variants everywhere (CRI_HCA_CAREFUL=1)
};

// `mid` (optional) is recorded between the unpack and the transform kernel.
void launch_hca_decode(const HcaDecodeArgs& a, cudaStream_t s, uint64_t* launches, cudaEvent_t mid);
void launch_hca_noise_scan(const HcaDecodeArgs& a, cudaStream_t s, uint64_t* launches);   // between the two, when a.sfres
void launch_hca_intensity_scan(const HcaDecodeArgs& a, cudaStream_t s, uint64_t* launches);   // between the two, when a.carry_scan
void launch_hca_imdct(const HcaDecodeArgs& a, cudaStream_t s, uint64_t* launches);   // second half of launch_hca_decode
// Streams and events of the pipelined form of the fast path (owned by the context): the job's runs are cut into `chunks`
// pieces, the unpack kernels run back to back on the job's stream and every transform kernel follows its own unpack kernel
// on the high-priority `side` stream, so the transform of chunk k shares the SMs with the unpack of chunk k + 1.
struct HcaFastPipe {
    cudaStream_t side = nullptr;
    cudaEvent_t start = nullptr, done = nullptr;
    cudaEvent_t unpacked[16] = {};
    uint32_t chunks = 1;
};
void launch_hca_decode_fast(const HcaDecodeArgs& a, cudaStream_t s, uint64_t* launches, cudaEvent_t mid, const HcaFastPipe* pipe);
uint32_t hca_fast_chunks();              // CRI_HCA_CHUNKS (default 1 = two launches, no overlap)
uint32_t hca_fast_threads_per_cta();     // columns per transform CTA
uint32_t hca_fast_ctas_per_sm();

struct HcaCryptArgs {
    const uint8_t* in;
    uint8_t* out;                  // same byte offsets as `in`
    const HcaStreamDev* streams;   // in_off = frame 0, cipher = table index
    const uint64_t* frame_prefix;  // [n_streams + 1] exclusive prefix of frame counts
    const uint8_t* tables;         // [n_tables][256] (already inverted when encrypting)
    uint64_t n_frames;
    uint32_t n_streams;
    uint32_t n_tables;
    // staged kernel: groups of up to frames_per_group consecutive frames of one stream, one warp each
    const uint64_t* group_prefix;  // [n_streams + 1] exclusive prefix of groups per stream
    const uint2* group_table;      // [n_groups] (stream, first frame of the group): what the LUT kernel reads instead of searching
    uint64_t n_groups;             // 0 = lane-per-frame kernel (frames too large to stage)
    uint32_t frames_per_group;
    uint32_t group_bytes;          // shared-memory bytes per group (multiple of 16, >= frames_per_group * frame_size + 32)
    uint32_t lut_table;            // the cipher table most streams use: the LUT kernel keeps a bank-replicated copy of it
    uint32_t min_frame;            // smallest frame size of the job (the LUT kernel wants >= 128 bytes)
};
void launch_hca_crypt(const HcaCryptArgs& a, cudaStream_t s, uint64_t* launches);

struct HcaEncodeArgs {
    const uint8_t* in;             // WAV images
    uint8_t* out;                  // HCA images (headers are host-built patches)
    const HcaStreamDev* streams;   // in_off = first PCM sample, out_off = frame 0, out_samples = samples per channel
    const uint64_t* frame_prefix;  // [n_streams + 1]
    const uint16_t* crc_mul;       // [n_streams][32]: x^(8 * bytes behind lane l's CRC chunk) mod P (hca_engine.cu)
    int32_t* status;
    uint64_t n_frames;
    uint32_t n_streams;
    uint32_t max_channels;
    uint32_t frame_words;          // 32-bit words of the shared frame buffer (>= frame_size / 4 + 2)
    uint32_t smem_per_warp;        // filled in by the launcher
};
int launch_hca_encode(HcaEncodeArgs a, cudaStream_t s, uint64_t* launches);   // -1: frame too large for shared memory

}  // namespace cri
