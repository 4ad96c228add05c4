// Host-side container/header logic of the ADX / HCA hot path: WAV ingest, ADX
// and HCA header parsing, output-size planning, cipher-table derivation, CRC16.
// Everything here runs once per stream on the host; the per-sample / per-frame
// work lives in the CUDA kernels (adx_kernels.cu, hca_*_kernels.cu).
//
// Reference behaviour mirrored (file:line into Youjose/PyCriCodecs CriCodecs/):
//   WAV ingest           pcm.cpp:291-342, 411-444 (16-bit PCM only)
//   WAV image writer     pcm.cpp:350-375, 547-556
//   ADX header           adx.cpp:145-183, 298-358 ; coefficients adx.cpp:58-64
//   ADX encode planning  adx.cpp:416-486
//   HCA header           hca.cpp:628-984 ; cipher tables hca.cpp:499-617
//   HCA encode planning  hca.cpp:2206-2462 ; header writer hca.cpp:3109-3164
//   CryptHeader          hca.cpp:3166-3250
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <thread>
#include <vector>

namespace cri {

// Status codes shared by the C-ABI (include/cricodecs_b200.h documents them).
enum : int {
    OK = 0,
    // ADX: reference AdxErrorCode values -1..-18 are passed through unchanged.
    // WAV ingest errors are reported as ERR_WAV_BASE + pcm.cpp code (-1..-10).
    ERR_WAV_BASE = -100,
    ERR_HCA_HEADER = -201,   // py_codec_err(-1)
    ERR_HCA_DECODE = -202,   // py_codec_err(-2)
    ERR_HCA_CHANNELS = -203, // py_codec_err(-3)
    ERR_HCA_ENCODE = -204,   // py_codec_err(-4)
    ERR_UNSUPPORTED = -300,  // valid input that this build does not handle yet
    ERR_BUFFER = -301,       // caller's output buffer too small / truncated input
    ERR_CUDA = -400,
};

inline uint32_t be16(const uint8_t* p) { return ((uint32_t)p[0] << 8) | p[1]; }
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline uint32_t le16(const uint8_t* p) { return p[0] | ((uint32_t)p[1] << 8); }
inline uint32_t le32(const uint8_t* p) { return p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline void put_be16(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)v; }
inline void put_be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
inline void put_le16(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
inline void put_le32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

uint16_t crc16(const uint8_t* p, size_t n);

// Header parsing is independent per stream and, for a batch of thousands of streams scattered over a blob, mostly cache
// misses: large batches are parsed by a handful of host threads (fn(i) must only touch stream i's slots).
template <class F>
inline void parallel_for(uint32_t n, F fn) {
    const unsigned hw = std::thread::hardware_concurrency();
    const uint32_t workers = n < 512 ? 1u : std::min<uint32_t>(std::min<uint32_t>(hw ? hw : 1u, 8u), n / 256);
    if (workers <= 1) {
        for (uint32_t i = 0; i < n; i++) fn(i);
        return;
    }
    std::vector<std::thread> pool;
    for (uint32_t w = 0; w < workers; w++)
        pool.emplace_back([=] {
            const uint32_t lo = (uint32_t)((uint64_t)n * w / workers), hi = (uint32_t)((uint64_t)n * (w + 1) / workers);
            for (uint32_t i = lo; i < hi; i++) fn(i);
        });
    for (auto& t : pool) t.join();
}

// ---------------------------------------------------------------- WAV
// Sample encodings the reference converts to PCM16 before encoding (PCM::load_WAVE / Get_PCM16, pcm.cpp:291-327, 455-545)
enum WavSampleFormat : uint8_t {
    WAV_S16 = 0,   // 9..16 valid bits in 2 bytes: used as they are
    WAV_U8 = 1,    // <= 8 valid bits in 1 byte:   (u8 - (1 << (bits - 1))) << 8
    WAV_S24 = 2,   // 17..24 valid bits in 3 bytes: sign-extended >> (bits - 16)
    WAV_S32 = 3,   // 32 bits in 4 bytes:          >> 16
    WAV_F32 = 4,   // IEEE float:  clamp((int)(x * 32767))
    WAV_F64 = 5,   // IEEE double: clamp((int)(x * 32767))
};
struct WavInfo {
    int channels = 0, rate = 0, looping = 0;
    uint32_t loop_start = 0, loop_end = 0;
    uint32_t loop_count = 0;     // NumberofSampleLoops of the smpl chunk (loop_start / loop_end are loop 0)
    size_t data_offset = 0;      // byte offset of the first PCM sample in the image
    uint32_t total_samples = 0;  // over all channels (the reference's ColumnSize)
    uint8_t format = WAV_S16;    // WavSampleFormat
    uint8_t sample_bytes = 2;
    uint8_t shift = 0;           // WAV_U8: valid bits; WAV_S24: bits - 16
};
// the reference's conversion of sample `index` (over all channels) to PCM16
int16_t wav_sample_s16(const WavInfo& w, const uint8_t* data, size_t index);
int parse_wav(const uint8_t* d, size_t n, WavInfo* w);  // 0 or pcm.cpp code (-1..-8)
size_t wav_header_size(bool looping);
void write_wav_header(uint8_t* out, uint32_t samples_per_channel, int channels, int rate, bool looping,
                      uint32_t loop_start, uint32_t loop_end);

// ---------------------------------------------------------------- ADX
struct AdxInfo {
    int mode = 0, block_size = 0, bit_depth = 0, channels = 0, version = 0, data_offset = 0, looping = 0;
    uint32_t rate = 0, samples = 0, highpass = 0, loop_start = 0, loop_end = 0;
    int16_t history[256][2] = {};
    int coef[2] = {0, 0};
    uint32_t blocks = 0;          // blocks per channel the decoder walks (adx.cpp:386)
    uint32_t samples_per_block = 0;
};
void adx_coefficients(unsigned highpass, unsigned rate, int coef[2]);
int parse_adx(const uint8_t* d, size_t n, AdxInfo* a);   // 0 or AdxErrorCode (-1..-9)

struct AdxEncPlan {
    int channels = 0, mode = 3, block_size = 18, bit_depth = 4, version = 4, filter = 0;
    unsigned highpass = 500;
    uint32_t rate = 0, samples = 0;  // samples per channel as written to the header
    uint32_t frames = 0;             // blocks per channel that are coded
    uint32_t samples_per_block = 0;
    int header_size = 0;
    size_t out_size = 0;
    int coef[2] = {0, 0};
    int looping = 0;                 // the WAV carries a sampler loop: the header gets a loop table (adx.cpp:94-143)
    uint32_t loop_start = 0, loop_end = 0;
};
// Validation order and codes as ADX::Encode (adx.cpp:424-442): -10..-18.
int plan_adx_encode(const WavInfo& w, unsigned bit_depth, unsigned block_size, unsigned mode, unsigned highpass,
                    unsigned filter, unsigned version, AdxEncPlan* p, bool looping = false);
// Writes the header (incl. initial history, "(c)CRI") and the EOF block; block payload is the kernel's job.
void write_adx_frame(uint8_t* out, const AdxEncPlan& p, const int16_t* first_samples);

// ---------------------------------------------------------------- HCA
struct HcaInfo {
    unsigned version = 0, header_size = 0, channels = 0, rate = 0, frame_count = 0, delay = 0, padding = 0;
    unsigned frame_size = 0, min_res = 0, max_res = 0, tracks = 0, channel_config = 0;
    unsigned total_bands = 0, base_bands = 0, stereo_bands = 0, bands_per_hfr = 0, ms_stereo = 0, hfr_groups = 0;
    unsigned ath_type = 0, ciph_type = 0;
    unsigned loop_flag = 0, loop_start_frame = 0, loop_end_frame = 0, loop_start_delay = 0, loop_end_padding = 0;
    uint8_t ath[128] = {};
    uint8_t type[16] = {};       // 0 discrete, 1 stereo primary, 2 stereo secondary
    unsigned coded[16] = {};
};
int parse_hca(const uint8_t* d, size_t n, HcaInfo* h);   // 0 or negative (any error maps to ERR_HCA_HEADER)
uint64_t mix_subkey(uint64_t key, unsigned subkey);
int cipher_table(int type, uint64_t key, uint8_t table[256]);
// Header half of HcaCrypt: toggles the signature masks, sets the ciph type, rewrites the header CRC.
void crypt_header(uint8_t* header, unsigned header_size, unsigned new_type);

struct HcaEncPlan {
    unsigned channels = 0, rate = 0, samples = 0, frame_size = 0, frame_count = 0, delay = 128, padding = 0, header_size = 96;
    unsigned total_bands = 0, base_bands = 0, stereo_bands = 0, hfr_groups = 0, bands_per_hfr = 0, hfr_band_count = 0;
    unsigned channel_config = 0;
    uint8_t type[16] = {};
    unsigned coded[16] = {};
    // looping input (initHCAEncode / CalculateLoopInfo / CalculateHeaderSize, hca.cpp:2292-2321, 2440-2449): the encoder
    // is fed  pre_zero x silence, pre_first x the first sample frame, main_samples input frames, post_samples frames
    // from the loop start, then silence
    unsigned loop_flag = 0, loop_start_frame = 0, loop_end_frame = 0, loop_start_delay = 0, loop_end_padding = 0;
    unsigned loop_start = 0, pre_zero = 0, pre_first = 0, main_samples = 0, post_samples = 0;
};
int plan_hca_encode(unsigned channels, unsigned rate, unsigned samples_per_channel, unsigned quality, HcaEncPlan* p);
// the same for a WAV with a sampler loop; -300 where the reference itself reads outside its buffers
int plan_hca_encode_loop(const WavInfo& w, unsigned quality, HcaEncPlan* p);
void write_hca_header(uint8_t* out, const HcaEncPlan& p);

}  // namespace cri
