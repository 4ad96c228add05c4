// Device-side launch tables and kernel launchers shared by engine.cu and the
// *_kernels.cu translation units. Names follow the codec domain: an ADX "chain"
// is one channel of one stream (a strictly serial predictor recurrence), an HCA
// "unit" is a run of consecutive frames of one stream.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace cri {

// ------------------------------------------------------------------ ADX
struct AdxChain {
    uint64_t in_off;      // decode: first block of this channel; encode: first PCM sample of this channel (bytes into the in blob)
    uint64_t out_off;     // decode: first PCM sample of this channel; encode: first block of this channel (bytes into the out blob)
    uint64_t eof_off;     // decode: channel-0 block of frame 0 (EOF marker probe); unused for encode
    uint32_t blocks;      // blocks to walk
    uint32_t samples;     // valid samples per channel (decode: clip writes; encode: pad reads with zeros)
    uint32_t in_stride;   // decode: bytes between this channel's blocks; encode: int16 elements between samples
    uint32_t out_stride;  // decode: int16 elements between samples; encode: bytes between blocks
    int32_t coef0, coef1;
    int16_t hist1, hist2;
    uint8_t mode, bit_depth, block_size, filter;  // filter: encode mode 2 predictor index
    uint8_t channels, channel;                    // of the stream / this chain's index; fast lists: uniform per warp
    uint8_t pad[2];
    uint32_t stream;
};

void launch_adx_decode(const uint8_t* d_in, uint8_t* d_out, const AdxChain* d_chains, uint32_t n_fast, uint32_t n_generic,
                       cudaStream_t s, uint64_t* launches);
void launch_adx_encode(const uint8_t* d_in, uint8_t* d_out, const AdxChain* d_chains, uint32_t n_fast, uint32_t n_generic,
                       cudaStream_t s, uint64_t* launches);

// WAV ingest: samples of other encodings are converted to PCM16 into a spare region of the input blob before the encode
// kernels run (the reference's PCM::Get_PCM16, pcm.cpp:530-545).
struct PcmConv {
    uint64_t src_off;     // first sample in the input blob
    uint64_t dst_off;     // first int16 of the converted copy (2-byte aligned, inside the blob's conversion region)
    uint32_t count;       // samples over all channels
    uint8_t format;       // WavSampleFormat
    uint8_t shift;
    uint8_t kind;         // 0: convert samples src[i]; 1: repeat the first sample frame (src[i % channels]); 2: silence
    uint8_t channels;
};
void launch_pcm_convert(uint8_t* d_blob, const PcmConv* d_conv, uint32_t n, uint32_t max_count, cudaStream_t s, uint64_t* launches);

// Scatter host-built header/trailer bytes into the output blob (one patch per stream piece).
struct Patch {
    uint64_t dst_off;
    uint32_t src_off;
    uint32_t bytes;
};
void launch_scatter_patches(uint8_t* d_out, const uint8_t* d_patch_bytes, const Patch* d_patches, uint32_t n,
                            cudaStream_t s, uint64_t* launches);

// Gather byte ranges of a device blob into a packed device buffer (the device-pointer entry points fetch the
// stream headers this way: the host plans from headers only).
struct Segment {
    uint64_t src_off;
    uint64_t dst_off;
    uint32_t bytes;
    uint32_t pad;
};
void launch_gather_segments(uint8_t* d_dst, const uint8_t* d_src, const Segment* d_segs, uint32_t n, cudaStream_t s,
                            uint64_t* launches);

}  // namespace cri
