// ADX (CRI ADPCM) decode / encode kernels for sm_100a.
//
// What is computed (reference: CriCodecs/adx.cpp): per channel a 2-tap
// fixed-point predictor  s = q*scale + (c0*h1 >> 12) + (c1*h2 >> 12), clamped to
// int16, q a signed n-bit code (ChannelFrame::Decode, adx.cpp:189-214); the
// encoder searches a per-block scale from the residual range against RAW history
// and then quantises against the SIMULATED decoder history (ChannelFrame::Encode,
// adx.cpp:215-273). The clamp and the two floor shifts make the recurrence
// non-associative, so one channel of one stream ("chain") is strictly serial and
// the parallelism is chains: one lane per chain, 32 chains per warp.
//
// Memory plan (fast path: 4-bit codes, 18-byte blocks -- the configuration
// every CRI tool emits): a warp walks its 32 chains in tiles of kTile blocks.
// Global traffic is staged through shared memory so that HBM sees coalesced
// runs: tile-in is loaded cooperatively (all lanes read consecutive bytes /
// samples of ONE chain at a time), each lane then runs its own recurrence out of
// shared memory, and tile-out is stored cooperatively the same way. The serial
// recurrence is latency-bound (~20 dependent cycles per sample), HBM traffic is
// 82 B per block (18 B code + 64 B PCM).
//
// Any other bit depth / block size / more than 32 samples per block takes the
// generic kernel: same arithmetic, one thread per chain straight on global memory.
#include <cstdint>

#include "kernels.h"

namespace cri {
namespace {

constexpr int kTile = 8;               // blocks per tile
constexpr int kSpb = 32;               // samples per block on the fast path
constexpr int kBlk = 18;               // bytes per block on the fast path
constexpr int kCodeRow = kTile * kBlk + 4;   // 148 B = 37 words: odd word stride, no bank conflicts
constexpr int kPcmRow = kTile * kSpb + 2;    // 258 int16 = 129 words
constexpr int kWarps = 2;               // 2 warps x 21 KB staging stays under the 48 KB static limit

__device__ __forceinline__ int clamp16(int v) { return min(max(v, -32768), 32767); }

__device__ __forceinline__ int decode_scale(int raw, int mode, int& c0, int& c1) {
    if (mode == 3) return raw + 1;
    if (mode == 4) return (int)(1u << ((12 - raw) & 31));
    // mode 2: predictor index in the top 3 bits (4 pairs exist), adx.cpp:196-201
    const int pred = (raw >> 13) & 3;
    c0 = pred == 0 ? 0 : pred == 1 ? 0x0F00 : pred == 2 ? 0x1CC0 : 0x1880;
    c1 = pred == 0 ? 0 : pred == 1 ? 0 : pred == 2 ? -0x0D00 : -0x0DC0;
    return (raw & 0x1FFF) + 1;
}

// ------------------------------------------------------------ decode, fast
__global__ void __launch_bounds__(kWarps * 32)
adx_decode_fast_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const AdxChain* __restrict__ chains,
                       uint32_t n_chains) {
    __shared__ __align__(16) uint8_t s_code[kWarps][32][kCodeRow];
    __shared__ __align__(16) int16_t s_pcm[kWarps][32][kPcmRow];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t first = (blockIdx.x * kWarps + warp) * 32u;
    if (first >= n_chains) return;
    const uint32_t mine = first + lane;
    const bool active = mine < n_chains;
    AdxChain ch = chains[active ? mine : first];
    if (!active) ch.blocks = 0;
    uint32_t warp_blocks = ch.blocks;
#pragma unroll
    for (int o = 16; o; o >>= 1) warp_blocks = max(warp_blocks, __shfl_xor_sync(0xFFFFFFFFu, warp_blocks, o));
    const uint32_t in_warp = min(32u, n_chains - first);

    int h1 = ch.hist1, h2 = ch.hist2, c0 = ch.coef0, c1 = ch.coef1;
    bool ended = false;  // EOF block seen (adx.cpp:405-406): the rest of the stream stays silent

    for (uint32_t b0 = 0; b0 < warp_blocks; b0 += kTile) {
        // ---- stage in: chain j's next kTile blocks, all lanes on one chain at a time
        for (uint32_t j = 0; j < in_warp; j++) {
            const AdxChain& cj = chains[first + j];
            const uint32_t nb = cj.blocks > b0 ? min((uint32_t)kTile, cj.blocks - b0) : 0u;
            for (uint32_t idx = lane; idx < nb * kBlk; idx += 32) {
                const uint32_t t = idx / kBlk, k = idx - t * kBlk;
                s_code[warp][j][idx] = in[cj.in_off + (uint64_t)(b0 + t) * cj.in_stride + k];
            }
        }
        __syncwarp();
        // ---- each lane decodes its own blocks out of shared memory
        const uint32_t nb = ch.blocks > b0 ? min((uint32_t)kTile, ch.blocks - b0) : 0u;
        for (uint32_t t = 0; t < nb; t++) {
            int16_t* dst = &s_pcm[warp][lane][t * kSpb];
            if (!ended) {
                const uint64_t probe = ch.eof_off + (uint64_t)(b0 + t) * ch.in_stride;
                ended = in[probe] == 0x80 && in[probe + 1] == 0x01;
            }
            if (ended) {
#pragma unroll
                for (int i = 0; i < kSpb; i++) dst[i] = 0;
                continue;
            }
            const uint8_t* src = &s_code[warp][lane][t * kBlk];
            const int scale = decode_scale((src[0] << 8) | src[1], ch.mode, c0, c1);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int byte = src[2 + k];
                const int q_hi = ((int)(byte << 24)) >> 28, q_lo = ((int)(byte << 28)) >> 28;
                int s = q_hi * scale + ((c0 * h1) >> 12) + ((c1 * h2) >> 12);
                s = clamp16(s);
                h2 = h1; h1 = s;
                dst[2 * k] = (int16_t)s;
                s = q_lo * scale + ((c0 * h1) >> 12) + ((c1 * h2) >> 12);
                s = clamp16(s);
                h2 = h1; h1 = s;
                dst[2 * k + 1] = (int16_t)s;
            }
        }
        __syncwarp();
        // ---- stage out: chain j's samples, lanes on consecutive samples (stride = channels)
        for (uint32_t j = 0; j < in_warp; j++) {
            const AdxChain& cj = chains[first + j];
            const uint32_t nbj = cj.blocks > b0 ? min((uint32_t)kTile, cj.blocks - b0) : 0u;
            const uint32_t base = b0 * kSpb;
            for (uint32_t idx = lane; idx < nbj * kSpb; idx += 32) {
                if (base + idx < cj.samples)
                    *reinterpret_cast<int16_t*>(out + cj.out_off + (uint64_t)(base + idx) * cj.out_stride * 2) =
                        s_pcm[warp][j][idx];
            }
        }
        __syncwarp();
    }
}

// --------------------------------------------------------- decode, generic
__device__ __forceinline__ uint32_t read_bits_be(const uint8_t* p, uint32_t bitpos, int count) {
    uint32_t v = 0;
    for (int i = 0; i < count; i++, bitpos++) v = (v << 1) | ((p[bitpos >> 3] >> (7 - (bitpos & 7))) & 1u);
    return v;
}

__global__ void adx_decode_generic_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                          const AdxChain* __restrict__ chains, uint32_t n_chains) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chains) return;
    const AdxChain ch = chains[i];
    const int spb = (ch.block_size - 2) * 8 / ch.bit_depth;
    int h1 = ch.hist1, h2 = ch.hist2, c0 = ch.coef0, c1 = ch.coef1;
    bool ended = false;
    for (uint32_t b = 0; b < ch.blocks; b++) {
        const uint64_t probe = ch.eof_off + (uint64_t)b * ch.in_stride;
        if (!ended) ended = in[probe] == 0x80 && in[probe + 1] == 0x01;
        const uint8_t* src = in + ch.in_off + (uint64_t)b * ch.in_stride;
        int scale = 0;
        if (!ended) scale = decode_scale((src[0] << 8) | src[1], ch.mode, c0, c1);
        for (int k = 0; k < spb; k++) {
            const uint32_t idx = b * spb + k;
            int s = 0;
            if (!ended) {
                const int sh = 32 - ch.bit_depth;
                const int q = ((int)(read_bits_be(src + 2, k * ch.bit_depth, ch.bit_depth) << sh)) >> sh;
                s = clamp16(q * scale + ((c0 * h1) >> 12) + ((c1 * h2) >> 12));
                h2 = h1; h1 = s;
            }
            if (idx < ch.samples) {  // byte stores: this path also serves odd output addresses
                uint8_t* o = out + ch.out_off + (uint64_t)idx * ch.out_stride * 2;
                o[0] = (uint8_t)s;
                o[1] = (uint8_t)(s >> 8);
            }
        }
    }
}

// ------------------------------------------------------------ encode core
// One block of one chain. `smp` holds the block's samples; writes block_size
// bytes (zero-initialised by the caller) through put(byte_index, value).
struct ScaleChoice {
    int scale;      // value used by the quantiser / simulated decoder
    int word;       // 16-bit word stored in the block
};

__device__ __forceinline__ ScaleChoice choose_scale(int mn, int mx, int limit, int mode, int filter) {
    // C division truncates toward zero; ~limit == -(limit+1)  (adx.cpp:236-238)
    const int a = mx / limit, b = mn / ~limit;
    int scale = (a > b ? a : b) & 0xFFFF;
    if (scale > 0x1000) scale = 0x1000;
    ScaleChoice r;
    if (mode == 4) {
        const int power = scale == 0 ? 0 : (31 - __clz(scale)) + 1;
        r.scale = (1 << power) & 0xFFFF;
        r.word = (12 - power) & 0xFFFF;
    } else if (mode == 2) {
        r.scale = scale;
        r.word = ((filter << 13) | (scale & 0x1FFF)) & 0xFFFF;
    } else {
        r.scale = scale;
        r.word = scale;
    }
    return r;
}

// Exact C-style truncating division of v (|v| < 2^20) by d (1..8192) after the
// reference's round-half-away bias; only quotients in [-9, 8] need to be exact
// because the result is clamped to n-bit range right after (adx.cpp:258-260).
__device__ __forceinline__ int div_trunc(int v, int d) { return v / d; }

// ------------------------------------------------------------ encode, fast
__global__ void __launch_bounds__(kWarps * 32)
adx_encode_fast_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const AdxChain* __restrict__ chains,
                       uint32_t n_chains) {
    __shared__ __align__(16) int16_t s_pcm[kWarps][32][kPcmRow];
    __shared__ __align__(16) uint8_t s_code[kWarps][32][kCodeRow];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t first = (blockIdx.x * kWarps + warp) * 32u;
    if (first >= n_chains) return;
    const uint32_t mine = first + lane;
    const bool active = mine < n_chains;
    AdxChain ch = chains[active ? mine : first];
    if (!active) ch.blocks = 0;
    uint32_t warp_blocks = ch.blocks;
#pragma unroll
    for (int o = 16; o; o >>= 1) warp_blocks = max(warp_blocks, __shfl_xor_sync(0xFFFFFFFFu, warp_blocks, o));
    const uint32_t in_warp = min(32u, n_chains - first);

    int h1 = ch.hist1, h2 = ch.hist2;
    const int c0 = ch.coef0, c1 = ch.coef1;
    const int limit = 7;

    for (uint32_t b0 = 0; b0 < warp_blocks; b0 += kTile) {
        for (uint32_t j = 0; j < in_warp; j++) {  // stage in: lanes on consecutive samples of chain j
            const AdxChain& cj = chains[first + j];
            const uint32_t nb = cj.blocks > b0 ? min((uint32_t)kTile, cj.blocks - b0) : 0u;
            const uint32_t base = b0 * kSpb;
            for (uint32_t idx = lane; idx < nb * kSpb; idx += 32) {
                int16_t v = 0;
                if (base + idx < cj.samples)
                    v = *reinterpret_cast<const int16_t*>(in + cj.in_off + (uint64_t)(base + idx) * cj.in_stride * 2);
                s_pcm[warp][j][idx] = v;
            }
        }
        __syncwarp();
        const uint32_t nb = ch.blocks > b0 ? min((uint32_t)kTile, ch.blocks - b0) : 0u;
        for (uint32_t t = 0; t < nb; t++) {
            const int16_t* smp = &s_pcm[warp][lane][t * kSpb];
            uint8_t* dst = &s_code[warp][lane][t * kBlk];
            // pass 1: residual range against RAW history (adx.cpp:221-230)
            const int o1 = h1, o2 = h2;
            int mn = 0, mx = 0;
#pragma unroll
            for (int i = 0; i < kSpb; i++) {
                const int s = smp[i];
                const int r = (s * 4096 - c0 * h1 - c1 * h2) >> 12;
                mn = min(mn, r); mx = max(mx, r);
                h2 = h1; h1 = s;
            }
            if (mn == 0 && mx == 0) {  // silent residual: all-zero block, history stays raw (adx.cpp:231-234)
#pragma unroll
                for (int k = 0; k < kBlk; k++) dst[k] = 0;
                continue;
            }
            const ScaleChoice sc = choose_scale(mn, mx, limit, ch.mode, ch.filter);
            dst[0] = (uint8_t)(sc.word >> 8);
            dst[1] = (uint8_t)sc.word;
            const int scale = sc.scale ? sc.scale : 1;  // adx.cpp:256-257
            const int half = scale >> 1;
            h1 = o1; h2 = o2;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                int byte = 0;
#pragma unroll
                for (int n = 0; n < 2; n++) {
                    const int s = smp[2 * k + n];
                    const int pred = c0 * h1 + c1 * h2;
                    int d = (s * 4096 - pred) >> 12;
                    d = d > 0 ? d + half : d - half;
                    d = div_trunc(d, scale);
                    d = min(max(d, -8), 7);
                    const int sim = clamp16((d * 4096 * scale + pred) >> 12);
                    h2 = h1; h1 = sim;
                    byte = (byte << 4) | (d & 0xF);
                }
                dst[2 + k] = (uint8_t)byte;
            }
        }
        __syncwarp();
        for (uint32_t j = 0; j < in_warp; j++) {  // stage out: lanes on consecutive bytes of chain j's blocks
            const AdxChain& cj = chains[first + j];
            const uint32_t nbj = cj.blocks > b0 ? min((uint32_t)kTile, cj.blocks - b0) : 0u;
            for (uint32_t idx = lane; idx < nbj * kBlk; idx += 32) {
                const uint32_t t = idx / kBlk, k = idx - t * kBlk;
                out[cj.out_off + (uint64_t)(b0 + t) * cj.out_stride + k] = s_code[warp][j][idx];
            }
        }
        __syncwarp();
    }
}

// --------------------------------------------------------- encode, generic
__global__ void adx_encode_generic_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                          const AdxChain* __restrict__ chains, uint32_t n_chains) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chains) return;
    const AdxChain ch = chains[i];
    const int depth = ch.bit_depth;
    const int spb = (ch.block_size - 2) * 8 / depth;
    const int limit = (1 << (depth - 1)) - 1;
    int h1 = ch.hist1, h2 = ch.hist2;
    const int c0 = ch.coef0, c1 = ch.coef1;
    auto sample = [&](uint32_t idx) -> int {
        if (idx >= ch.samples) return 0;
        const uint8_t* q = in + ch.in_off + (uint64_t)idx * ch.in_stride * 2;  // byte loads: any alignment
        return (int)(int16_t)(q[0] | (q[1] << 8));
    };
    for (uint32_t b = 0; b < ch.blocks; b++) {
        uint8_t* dst = out + ch.out_off + (uint64_t)b * ch.out_stride;
        for (int k = 0; k < ch.block_size; k++) dst[k] = 0;
        const int o1 = h1, o2 = h2;
        int mn = 0, mx = 0;
        for (int k = 0; k < spb; k++) {
            const int s = sample(b * spb + k);
            const int r = (s * 4096 - c0 * h1 - c1 * h2) >> 12;
            mn = min(mn, r); mx = max(mx, r);
            h2 = h1; h1 = s;
        }
        if (mn == 0 && mx == 0) continue;
        const ScaleChoice sc = choose_scale(mn, mx, limit, ch.mode, ch.filter);
        dst[0] = (uint8_t)(sc.word >> 8);
        dst[1] = (uint8_t)sc.word;
        const int scale = sc.scale ? sc.scale : 1;
        const int half = scale >> 1;
        h1 = o1; h2 = o2;
        uint32_t bitpos = 16;
        for (int k = 0; k < spb; k++) {
            const int s = sample(b * spb + k);
            const int pred = c0 * h1 + c1 * h2;
            int d = (s * 4096 - pred) >> 12;
            d = d > 0 ? d + half : d - half;
            d = d / scale;
            d = min(max(d, ~limit), limit);
            const int sim = clamp16((d * 4096 * scale + pred) >> 12);
            h2 = h1; h1 = sim;
            for (int bit = depth - 1; bit >= 0; bit--, bitpos++)
                if ((d >> bit) & 1) dst[bitpos >> 3] |= (uint8_t)(0x80u >> (bitpos & 7));
        }
    }
}

__global__ void scatter_patches_kernel(uint8_t* __restrict__ out, const uint8_t* __restrict__ bytes,
                                       const Patch* __restrict__ patches, uint32_t n) {
    const uint32_t p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= n) return;
    const Patch pt = patches[p];
    for (uint32_t i = threadIdx.x & 31; i < pt.bytes; i += 32) out[pt.dst_off + i] = bytes[pt.src_off + i];
}

}  // namespace

void launch_adx_decode(const uint8_t* d_in, uint8_t* d_out, const AdxChain* d_chains, uint32_t n_fast, uint32_t n_generic,
                       cudaStream_t s, uint64_t* launches) {
    if (n_fast) {
        const uint32_t per_cta = kWarps * 32;
        adx_decode_fast_kernel<<<(n_fast + per_cta - 1) / per_cta, per_cta, 0, s>>>(d_in, d_out, d_chains, n_fast);
        ++*launches;
    }
    if (n_generic) {
        adx_decode_generic_kernel<<<(n_generic + 63) / 64, 64, 0, s>>>(d_in, d_out, d_chains + n_fast, n_generic);
        ++*launches;
    }
}

void launch_adx_encode(const uint8_t* d_in, uint8_t* d_out, const AdxChain* d_chains, uint32_t n_fast, uint32_t n_generic,
                       cudaStream_t s, uint64_t* launches) {
    if (n_fast) {
        const uint32_t per_cta = kWarps * 32;
        adx_encode_fast_kernel<<<(n_fast + per_cta - 1) / per_cta, per_cta, 0, s>>>(d_in, d_out, d_chains, n_fast);
        ++*launches;
    }
    if (n_generic) {
        adx_encode_generic_kernel<<<(n_generic + 63) / 64, 64, 0, s>>>(d_in, d_out, d_chains + n_fast, n_generic);
        ++*launches;
    }
}

void launch_scatter_patches(uint8_t* d_out, const uint8_t* d_bytes, const Patch* d_patches, uint32_t n, cudaStream_t s,
                            uint64_t* launches) {
    if (!n) return;
    scatter_patches_kernel<<<(n + 3) / 4, 128, 0, s>>>(d_out, d_bytes, d_patches, n);
    ++*launches;
}

}  // namespace cri
