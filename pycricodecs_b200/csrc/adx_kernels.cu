// ADX (CRI ADPCM) decode / encode kernels for sm_100a.
//
// What is computed (reference: CriCodecs/adx.cpp): per channel a 2-tap fixed-point predictor
//   s = q*scale + (c0*h1 >> 12) + (c1*h2 >> 12), clamped to int16, q a signed n-bit code
// (ChannelFrame::Decode, adx.cpp:189-214); the encoder picks a per-block scale from the residual range against
// RAW history and then quantises against the SIMULATED decoder history (ChannelFrame::Encode, adx.cpp:215-273).
// The clamp and the floor shifts make the recurrence non-associative, so one channel of one stream (a "chain") is
// strictly serial and the parallelism is chains: one lane per chain, 32 chains (32/channels streams) per warp.
//
// Fast path (4-bit codes, 18-byte blocks, 1/2/4/8/16/32 channels -- what every CRI tool emits): the warp walks its
// streams in tiles of kTile blocks. A stream's tile is CONTIGUOUS in both blobs (blocks of all channels are
// interleaved per frame, PCM is interleaved per sample), so the warp copies the aligned 32-bit words that cover
// each stream's tile with cp.async into a double-buffered shared stage (next tile in flight while this one is
// decoded; no alignment requirement on the stream itself), every lane runs its own recurrence out of shared
// memory, and results leave as coalesced stores per stream. The recurrence is latency bound (~20 dependent
// cycles per decoded sample, ~70 per encoded sample); HBM traffic is the compulsory 82 B per block.
//
// Anything else (other bit depths / block sizes / channel counts, odd PCM alignment) takes the generic kernels:
// same arithmetic, one thread per chain straight on global memory.
#include <algorithm>
#include <cstdint>
#include <type_traits>
#ifdef CRI_ADX_TIMING
#include <cstdio>
#endif

#include "kernels.h"

namespace cri {
namespace {

constexpr int kTile = 8;                 // blocks per chain per tile
constexpr int kSpb = 32;                 // samples per block on the fast path
constexpr int kBlk = 18;                 // bytes per block on the fast path
// input stages: per stream one row of whole 16-byte chunks covering its tile (+ one chunk for the misalignment)
constexpr int kCodeWords = 32 * ((kTile * kBlk + 15) / 16 + 1) * 4;        // 32 chains x kTile blocks
constexpr int kPcmWords = 32 * ((kTile * kSpb * 2 + 15) / 16 + 1) * 4;     // 32 chains x kTile x 32 samples
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kMovers = 3;               // mover warps per worker warp
// A CTA holds kGroups independent worker + movers groups: warp w belongs to group w % kGroups with role w / kGroups
// (0 = worker). Warps map to the SM's four schedulers by warp index, so the four workers of a CTA sit on four
// different schedulers -- with one 4-warp CTA per group, every resident CTA's worker landed on scheduler 0 and three
// to four serial recurrences shared one issue port while the other three schedulers idled with the movers.
constexpr int kGroups = 4;
constexpr int kGroupThreads = 32 * (1 + kMovers);
__device__ __forceinline__ void group_sync(int group) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(kGroupThreads) : "memory");
}

__device__ __forceinline__ int clamp16(int v) { return min(max(v, -32768), 32767); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ int decode_scale(int raw, int mode, int& c0, int& c1) {
    if (mode == 3) return raw + 1;
    if (mode == 4) return (int)(1u << ((12 - raw) & 31));
    // mode 2: predictor index in the top 3 bits (4 pairs exist), adx.cpp:196-201
    const int pred = (raw >> 13) & 3;
    c0 = pred == 0 ? 0 : pred == 1 ? 0x0F00 : pred == 2 ? 0x1CC0 : 0x1880;
    c1 = pred == 0 ? 0 : pred == 1 ? 0 : pred == 2 ? -0x0D00 : -0x0DC0;
    return (raw & 0x1FFF) + 1;
}

// Warp specialisation: a CTA is two warps working on the same 32 chains. Warp 0 ("worker") runs the 32 serial
// recurrences out of shared memory; warp 1 ("mover") keeps the next input tile arriving (cp.async) and drains the
// previous output tile to HBM, so the recurrence never waits for memory. One group barrier per tile hands the
// double-buffered tiles over.
struct StreamInfo {
    uint64_t in_base;    // first byte of the stream's payload in the input blob (decode: frame 0, encode: sample 0)
    uint64_t out_base;   // first byte of the stream's payload in the output blob
    uint32_t blocks;     // blocks per channel to walk
    uint32_t samples;    // valid samples per channel
};

// mover: request the aligned 16-byte chunks that cover bytes [in_base + lo, in_base + hi) of every stream into its stage
// row (the blob starts 256-byte aligned and has slack behind it, so the chunks may overhang the stream on both sides)
__device__ __forceinline__ void mover_request(uint32_t* stage, int row_words, const uint8_t* blob, const StreamInfo* info,
                                              int nstreams, uint32_t b0, int unit_bytes, bool clip_samples, int nch, int lane,
                                              int mover) {
    for (int s = mover; s < nstreams; s += kMovers) {
        const StreamInfo si = info[s];
        const uint32_t nb = si.blocks > b0 ? min((uint32_t)kTile, si.blocks - b0) : 0u;
        uint64_t lo = (uint64_t)b0 * unit_bytes, hi = lo + (uint64_t)nb * unit_bytes;
        if (clip_samples) hi = min(hi, (uint64_t)si.samples * nch * 2);   // encode: PCM ends with the stream, the rest is padding
        if (hi <= lo) continue;
        const uint64_t first = (si.in_base + lo) & ~(uint64_t)15;
        const int chunks = (int)(((si.in_base + hi + 15) & ~(uint64_t)15) - first) >> 4;
        for (int w = lane; w < chunks; w += 32) cp_async16(stage + s * row_words + 4 * w, blob + first + 16ull * w);
    }
}

// ------------------------------------------------------------ decode, fast
// Decode groups are one worker and ONE mover warp: the mover does not copy, it issues bulk copies (cp.async.bulk: the
// copy engine moves a stream's tile between HBM and shared memory by itself and reports on an mbarrier / bulk group),
// one lane per stream. With cp.async requests and load / store loops the movers issued 40 % of the kernel's
// instructions on the worker's own scheduler, and the lone recurrence warp ran a third slower for it.
constexpr int kDecGroupThreads = 64;
__device__ __forceinline__ void dec_group_sync(int group) {
    asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(kDecGroupThreads) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
// PCM tile: per stream one row of kTile x 32 x channels samples, as in the WAV, placed p = (address of the row's first
// byte in the output blob) mod 16 bytes into a 16-byte aligned row of 16 more bytes: shared and global addresses of
// every sample then agree modulo 16, so all whole 16-byte units of the row leave in one bulk store (both ends of a bulk
// copy must be 16-byte aligned) and only the < 16 bytes at either end are stored by the lane itself.
constexpr int kPcmTileBytes = 32 * kTile * kSpb * 2 + 32 * 16;
struct alignas(16) DecodeStage {
    uint32_t code[3][kCodeWords];             // tile t in stage t % 3: two tiles are in flight
    alignas(16) uint8_t pcm[2][kPcmTileBytes];
    StreamInfo info[32];
    alignas(8) uint64_t full[3];              // mbarriers: the bulk loads of a stage have landed
};

__global__ void __launch_bounds__(kGroups * kDecGroupThreads, 1)
adx_decode_fast_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const AdxChain* __restrict__ chains,
                       uint32_t n_chains) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    const int group = (threadIdx.x >> 5) % kGroups, role = (threadIdx.x >> 5) / kGroups, lane = threadIdx.x & 31;
    DecodeStage& stage = reinterpret_cast<DecodeStage*>(s_dyn)[group];
    auto& s_code = stage.code;
    StreamInfo* s_info = stage.info;
    const uint32_t first = (blockIdx.x * kGroups + group) * 32u;
    if (first >= n_chains) return;
    const AdxChain ch = chains[first + lane];               // the list is padded to whole warps (idle: blocks == 0)
    const int nch = __shfl_sync(kFull, (int)ch.channels, 0);
    const int nstreams = 32 / nch;
    const int frame_bytes = nch * kBlk;
    const int row_words = ((kTile * frame_bytes + 15) / 16 + 1) * 4;   // whole 16-byte chunks
    const uint32_t tile_bytes = (uint32_t)(kTile * kSpb * 2 * nch);   // PCM of one stream's tile
    const uint32_t out_row = tile_bytes + 16;               // bytes per stream in the PCM tile
    uint32_t warp_blocks = ch.blocks;
#pragma unroll
    for (int o = 16; o; o >>= 1) warp_blocks = max(warp_blocks, __shfl_xor_sync(kFull, warp_blocks, o));
    const int slot = lane / nch;                            // this lane's stream within the CTA
    if (role == 0 && ch.channel == 0) s_info[slot] = StreamInfo{ch.eof_off, ch.out_off, ch.blocks, ch.samples};
    if (role == 1 && lane == 0) {
        for (int k = 0; k < 3; k++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&stage.full[k])), "r"(nstreams));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    dec_group_sync(group);
    const uint32_t ntiles = (warp_blocks + kTile - 1) / kTile;

    // mover, lane s = stream s: one bulk load brings the aligned 16-byte units that cover the stream's tile into its
    // stage row (the blob starts 256-byte aligned and has slack behind it); every stream lane arrives once per tile
    auto request = [&](uint32_t t) {
        if (lane >= nstreams) return;
        const StreamInfo si = s_info[lane];
        const uint32_t b0 = t * kTile;
        const uint32_t nb = si.blocks > b0 ? min((uint32_t)kTile, si.blocks - b0) : 0u;
        const uint32_t bar = smem_u32(&stage.full[t % 3]);
        if (nb) {
            const uint64_t lo = si.in_base + (uint64_t)b0 * frame_bytes, hi = lo + (uint64_t)nb * frame_bytes;
            const uint64_t from = lo & ~(uint64_t)15;
            const uint32_t bytes = (uint32_t)(((hi + 15) & ~(uint64_t)15) - from);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(&s_code[t % 3][lane * row_words])), "l"(in + from), "r"(bytes), "r"(bar) : "memory");
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
        }
    };
    // mover, lane s = stream s: tile t of the stream is one contiguous run of the WAV. Whole 16-byte units go out in one
    // bulk store, the < 16 bytes in front of and behind them sample by sample; returns once the copy engine has read the
    // tile (the worker overwrites it two tiles later).
    auto store_tile = [&](uint32_t t) {
#ifdef CRI_ADX_NOSTORE                                      // development: the worker's pace without the stores
        if (t != 0xFFFFFFFFu) return;
#endif
        if (lane < nstreams) {
            const StreamInfo si = s_info[lane];
            const uint32_t b0 = t * kTile, s0 = b0 * kSpb;
            const uint32_t nb = si.blocks > b0 ? min((uint32_t)kTile, si.blocks - b0) : 0u;
            const uint32_t count = s0 < si.samples ? min(nb * kSpb, si.samples - s0) : 0u;
            const uint32_t bytes = count * nch * 2;
            uint8_t* dst = out + si.out_base + (size_t)t * tile_bytes;
            const uint32_t ph = (uint32_t)(si.out_base & 15);             // tile_bytes is a multiple of 16
            const uint8_t* src = &stage.pcm[t & 1][lane * out_row + ph];
            const uint32_t head = min((16u - ph) & 15u, bytes), mid = (bytes - head) & ~15u, tail = bytes - head - mid;
            if (mid)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(dst + head), "r"(smem_u32(src + head)), "r"(mid) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#pragma unroll
            for (uint32_t k = 0; k < 14; k += 2) {          // ph is even: at most 7 samples on either side
                if (k < head) *reinterpret_cast<int16_t*>(dst + k) = *reinterpret_cast<const int16_t*>(src + k);
                if (k < tail) *reinterpret_cast<int16_t*>(dst + head + mid + k) = *reinterpret_cast<const int16_t*>(src + head + mid + k);
            }
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    };

    if (role == 1) {
        request(0);
        if (ntiles > 1) request(1);
        if (ntiles) mbar_wait(smem_u32(&stage.full[0]), 0);  // tile 0 has landed, tile 1 is in flight
    }
    dec_group_sync(group);

    int h1 = ch.hist1, h2 = ch.hist2, c0 = ch.coef0, c1 = ch.coef1;
    const int zero = ch.pad[0];
    bool ended = false;   // EOF block seen (adx.cpp:405-406): the rest of the stream stays silent
#ifdef CRI_ADX_TIMING     // development: where a tile's time goes (tools/build_variant.sh, UNIT=adx_kernels.cu)
    long long t_work = 0, t_wait = 0;
#endif
    for (uint32_t t = 0; t < ntiles; t++) {
        const int buf = t & 1;
        const uint32_t b0 = t * kTile;
#ifdef CRI_ADX_TIMING
        const long long t_begin = clock64();
#endif
        if (role == 1) {
#ifndef CRI_ADX_NOMOVE                                      // development: the worker's pace with an idle mover (stale data)
            if (t + 2 < ntiles) request(t + 2);
            if (t >= 1) store_tile(t - 1);
            if (t + 1 < ntiles) mbar_wait(smem_u32(&stage.full[(t + 1) % 3]), ((t + 1) / 3) & 1);   // tile t + 1 has landed
#endif
        } else {
            const uint32_t nb = ch.blocks > b0 ? min((uint32_t)kTile, ch.blocks - b0) : 0u;
            const int nch_rt = nch;
#ifndef CRI_ADX_NOMOVE
            mbar_wait(smem_u32(&stage.full[t % 3]), (t / 3) & 1);   // (the mover saw it before the barrier: passes at once)
#endif
            // byte 0 of this lane's first frame inside its stream's row
            const uint8_t* row = reinterpret_cast<const uint8_t*>(&s_code[t % 3][slot * row_words]) +
                                 (int)((ch.eof_off + (uint64_t)b0 * frame_bytes) & 15);
            int16_t* my_pcm = reinterpret_cast<int16_t*>(&stage.pcm[buf][slot * out_row + ((ch.out_off - 2u * ch.channel) & 15)]) + ch.channel;
            // A block reaches the registers as five words (six aligned shared-memory loads and a funnel shift: blocks start
            // on any byte), one block AHEAD of the one being decoded, so that its latency -- and that of the EOF probe,
            // which needs channel 0's scale word -- hides under 32 samples of recurrence instead of standing in front of
            // them. The loop runs to the warp's block count so that the shuffle below is warp-wide.
            const uint32_t nb_warp = warp_blocks > b0 ? min((uint32_t)kTile, warp_blocks - b0) : 0u;
            auto fetch = [&](uint32_t tb, uint32_t (&w)[5]) {
                const uint32_t at = (uint32_t)__cvta_generic_to_shared(row + tb * frame_bytes + ch.channel * kBlk);
                uint32_t r[6];
#pragma unroll
                for (int i = 0; i < 6; i++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r[i]) : "r"((at & ~3u) + 4u * i));
#pragma unroll
                for (int i = 0; i < 5; i++) w[i] = __funnelshift_r(r[i], r[i + 1], (at & 3u) * 8u);
            };
            // the sample stride in the PCM tile as a compile-time constant for mono and stereo (store offsets become
            // immediates; with a run-time stride every store waits for its own uniform-register add)
            auto decode_tile = [&](auto stride_tag) {
            constexpr int kStride = decltype(stride_tag)::value;
            const int nch = kStride ? kStride : nch_rt;
            uint32_t cur[5], nxt[5];
            fetch(0, cur);
            // A lone warp pays for every branch in full (nothing else issues while it resolves), so the block loop has
            // one: blocks past a lane's count decode stale bytes into samples the movers never copy; an EOF block
            // (adx.cpp:405-406) zeroes scale and coefficients, which makes the recurrence emit the silence the reference
            // writes; the scale modes are selects; the exact, clamped recurrence is one rarely taken branch at the end.
            for (uint32_t tb = 0; tb < nb_warp; tb++) {
                fetch(tb + 1, nxt);                           // past the tile: stale bytes of the stage, never used
                int16_t* dst = my_pcm + tb * kSpb * nch;      // sample i of this block -> dst[i * nch]
                const int head = (int)__byte_perm(cur[0], 0, 0x4401);            // the block's scale word (big-endian)
                const int head0 = __shfl_sync(kFull, head, lane - ch.channel);   // channel 0's scale word of this frame
                uint32_t blkw[5];
#pragma unroll
                for (int i = 0; i < 5; i++) { blkw[i] = cur[i]; cur[i] = nxt[i]; }
                ended = ended || head0 == 0x8001;
                // decode_scale, as selects
                const int pred = (head >> 13) & 3;
                const bool m3 = ch.mode == 3, m4 = ch.mode == 4;
                if (!m3 && !m4) {
                    c0 = pred == 0 ? 0 : pred == 1 ? 0x0F00 : pred == 2 ? 0x1CC0 : 0x1880;
                    c1 = pred == 0 ? 0 : pred == 1 ? 0 : pred == 2 ? -0x0D00 : -0x0DC0;
                }
                const int scale_raw = m3 ? head + 1 : m4 ? (int)(1u << ((12 - head) & 31)) : (head & 0x1FFF) + 1;
                const bool silent = ended || tb >= nb;        // (past the lane's blocks: keeps stale bytes from forcing the redo)
                const int scale = silent ? 0 : scale_raw, k0 = silent ? 0 : c0, k1 = silent ? 0 : c1;
                auto mul = [](int a, int b2) -> int { return (int)((uint32_t)a * (uint32_t)b2); };   // wraps (a discarded pass may overflow)
                // Speculative pass without the int16 clamp (adx.cpp:209): as long as no sample leaves the int16 range
                // the clamp is the identity. Every product stays below 2^31 for scales up to 0x2000 and |c| <= 0x2000;
                // otherwise, or when a sample leaves the range, the block is redone with the reference's clamped
                // recurrence from the saved history.
                int prod[kSpb];                               // the products do not depend on the history: all 32 first
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const uint32_t wd = blkw[(2 + k) >> 2];
                    const int pos = 8 * ((2 + k) & 3);        // the byte's high nibble is the first sample
                    // `zero` (a padding byte of the chain record) occupies the multiply-add's addend: the assembler
                    // would otherwise fold the product into the addition below, where it would wait for the shift
                    prod[2 * k] = mul(((int)(wd << (24 - pos))) >> 28, scale) + zero;
                    prod[2 * k + 1] = mul(((int)(wd << (28 - pos))) >> 28, scale) + zero;
                }
                // s_n = (c0*s_{n-1} >> 12) + x_n with x_n = prod_n + (c1*s_{n-2} >> 12) formed one sample ahead: two
                // dependent instructions per sample (multiply, shift-add) sit on the serial path, the rest fills the
                // gaps between them
                int s1 = h1, s2 = h2;
                int x = prod[0] + (mul(k1, s2) >> 12);
                int lo = 0, hi = 0;
#pragma unroll
                for (int n = 0; n < kSpb; n++) {
                    const int sn = (mul(k0, s1) >> 12) + x;
                    if (n + 1 < kSpb) x = prod[n + 1] + (mul(k1, s1) >> 12);
                    dst[n * nch] = (int16_t)sn;
                    lo = min(lo, sn); hi = max(hi, sn);
                    s2 = s1; s1 = sn;
                }
                const bool redo = lo < -32768 || hi > 32767 || scale > 0x2000 || abs(k0) > 0x2000 || abs(k1) > 0x2000;
                if (!redo) { h1 = s1; h2 = s2; }
                if (redo) {
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        const int byte = (int)((blkw[(2 + k) >> 2] >> (8 * ((2 + k) & 3))) & 0xFFu);
                        const int q_hi = ((int)(byte << 24)) >> 28, q_lo = ((int)(byte << 28)) >> 28;
                        int s = (q_hi * scale + ((k1 * h2) >> 12)) + ((k0 * h1) >> 12);
                        s = clamp16(s);
                        h2 = h1; h1 = s;
                        dst[(2 * k) * nch] = (int16_t)s;
                        s = (q_lo * scale + ((k1 * h2) >> 12)) + ((k0 * h1) >> 12);
                        s = clamp16(s);
                        h2 = h1; h1 = s;
                        dst[(2 * k + 1) * nch] = (int16_t)s;
                    }
                }
            }
            };
            if (nch_rt == 2) decode_tile(std::integral_constant<int, 2>{});
            else if (nch_rt == 1) decode_tile(std::integral_constant<int, 1>{});
            else decode_tile(std::integral_constant<int, 0>{});
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tile's samples, before the bulk store reads them
        }
#ifdef CRI_ADX_TIMING
        const long long t_mid = clock64();
        dec_group_sync(group);
        t_work += t_mid - t_begin;
        t_wait += clock64() - t_mid;
#else
        dec_group_sync(group);
#endif
    }
    if (role == 1 && ntiles) store_tile(ntiles - 1);
#ifdef CRI_ADX_TIMING
    if ((blockIdx.x == 0 || blockIdx.x == 77) && lane == 0 && group < 2)
        printf("adx decode cta %d group %d role %d: tiles %u work %lld wait %lld cycles\n", blockIdx.x, group, role, ntiles, t_work, t_wait);
#endif
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

// Exact C-style truncating division by a block-invariant divisor d (1..8192) for |v| < 2^19:
// q = (|v| * ceil(2^32 / d)) >> 32 is exact while |v| * d < 2^32 (round-up magic number, error < d / 2^32 per unit).
__device__ __forceinline__ uint32_t div_magic(int d) { return 0xFFFFFFFFu / (uint32_t)d + 1u; }   // ceil(2^32 / d) for d >= 2
__device__ __forceinline__ int div_trunc(int v, uint32_t magic, int d) {
    if (d == 1) return v;
    const int q = (int)__umulhi((uint32_t)abs(v), magic);
    return v < 0 ? -q : q;
}

// ------------------------------------------------------------ encode, fast
// Per (chain, block) result of the movers' share of ChannelFrame::Encode's first pass (adx.cpp:221-230). The residuals of a
// block's samples 2 .. 31 are taken against RAW samples of the same block; only the first two see the history the previous
// block left behind (its SIMULATED samples), so those two stay with the worker.
struct BlockHead {
    int32_t mn, mx;     // range of the residuals of samples 2 .. 31 of the block
};
struct alignas(16) EncodeStage {
    uint32_t pcm[2][kPcmWords];
    uint8_t code[2][32 * kTile * kBlk + 128];   // per stream: blocks in file order
    BlockHead head[2][32 * kTile];              // [tile parity][block in tile][chain]
    StreamInfo info[32];
};

__global__ void __launch_bounds__(kGroups * kGroupThreads, 1)
adx_encode_fast_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const AdxChain* __restrict__ chains,
                       uint32_t n_chains) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    const int group = (threadIdx.x >> 5) % kGroups, role = (threadIdx.x >> 5) / kGroups, lane = threadIdx.x & 31;
    EncodeStage& stage = reinterpret_cast<EncodeStage*>(s_dyn)[group];
    // reciprocals of every scale the quantiser can meet on this path: a table look-up instead of a division per block
    uint32_t* s_magic = reinterpret_cast<uint32_t*>(s_dyn + kGroups * sizeof(EncodeStage));
    for (int i = threadIdx.x; i <= 0x1000; i += blockDim.x) s_magic[i] = i >= 2 ? 0xFFFFFFFFu / (uint32_t)i + 1u : 0xFFFFFFFFu;   // ceil(2^32 / i); i = 1: see below
    __syncthreads();
    auto& s_pcm = stage.pcm;
    auto& s_code = stage.code;
    StreamInfo* s_info = stage.info;
    const uint32_t first = (blockIdx.x * kGroups + group) * 32u;
    if (first >= n_chains) return;
    const AdxChain ch = chains[first + lane];
    const int nch = __shfl_sync(kFull, (int)ch.channels, 0);
    const int nstreams = 32 / nch;
    const int frame_bytes = nch * kSpb * 2;                  // PCM bytes of one block of every channel
    const int row_words = ((kTile * frame_bytes + 15) / 16 + 1) * 4;   // whole 16-byte chunks
    const int code_row = kTile * nch * kBlk + 4;             // bytes per stream in the block tile (odd word stride)
    uint32_t warp_blocks = ch.blocks;
#pragma unroll
    for (int o = 16; o; o >>= 1) warp_blocks = max(warp_blocks, __shfl_xor_sync(kFull, warp_blocks, o));
    const int slot = lane / nch;
    const uint64_t stream_base = ch.in_off - 2ull * ch.channel;   // first PCM sample of the stream
    if (role == 0 && ch.channel == 0) s_info[slot] = StreamInfo{stream_base, ch.out_off, ch.blocks, ch.samples};
    group_sync(group);
    const uint32_t ntiles = (warp_blocks + kTile - 1) / kTile;

    auto store_tile = [&](uint32_t t) {                     // mover: tile t of every stream is contiguous in the ADX image
        const uint32_t b0 = t * kTile;
        for (int s = role - 1; s < nstreams; s += kMovers) {
            const StreamInfo si = s_info[s];
            const uint32_t nb = si.blocks > b0 ? min((uint32_t)kTile, si.blocks - b0) : 0u;
            uint8_t* dst = out + si.out_base + (uint64_t)b0 * nch * kBlk;
            const uint32_t bytes = nb * nch * kBlk;                 // even
            const uint8_t* srow = &s_code[t & 1][s * code_row];
            if ((reinterpret_cast<uintptr_t>(dst) & 1) == 0) {
                for (uint32_t e = lane; e < bytes / 2; e += 32)
                    reinterpret_cast<uint16_t*>(dst)[e] = reinterpret_cast<const uint16_t*>(srow)[e];
            } else {
                for (uint32_t e = lane; e < bytes; e += 32) dst[e] = srow[e];
            }
        }
    };

    const int c0 = ch.coef0, c1 = ch.coef1;
    const int limit = 7;
    // movers, first pass of tile t (lane = chain, the tile's blocks shared out over the three movers): range of the residuals
    // that depend on raw samples only. Samples past the end of the stream are silence (adx.cpp:450-460): zeroed in the
    // tile, so the worker reads them without a bounds check.
    auto first_pass = [&](uint32_t t) {
        const int buf = t & 1;
        const uint32_t b0 = t * kTile;
        const uint32_t nb = ch.blocks > b0 ? min((uint32_t)kTile, ch.blocks - b0) : 0u;
        int16_t* row = reinterpret_cast<int16_t*>(reinterpret_cast<uint8_t*>(&s_pcm[buf][slot * row_words]) +
                                                  (int)((stream_base + (uint64_t)b0 * frame_bytes) & 15)) + ch.channel;
        for (uint32_t tb = role - 1; tb < nb; tb += kMovers) {
            const uint32_t sbase = (b0 + tb) * kSpb;
            int16_t* q = row + (size_t)tb * kSpb * nch;
            if (sbase + kSpb > ch.samples) {                               // the stream ends inside this block: silence behind it
                for (int i = 0; i < kSpb; i++)
                    if (sbase + i >= ch.samples) q[i * nch] = 0;
            }
            // the arithmetic shift is monotone, so the range is taken over the unshifted values (two samples per
            // three-input min / max) and shifted once
            int h2 = q[0], h1 = q[nch];
            int mn = 0, mx = 0;
            q += 2 * nch;
            const int nc0 = -c0, nc1 = -c1;
#pragma unroll 5
            for (int i = 2; i < kSpb; i += 2, q += 2 * nch) {
                const int sa = q[0], sb = q[nch];
                const int va = sa * 4096 + (nc0 * h1 + nc1 * h2);
                const int vb = sb * 4096 + (nc0 * sa + nc1 * h1);
                mn = __vimin3_s32(mn, va, vb); mx = __vimax3_s32(mx, va, vb);
                h2 = sa; h1 = sb;
            }
            stage.head[buf][tb * 32 + lane] = BlockHead{mn >> 12, mx >> 12};
        }
    };

    if (role >= 1) {
        mover_request(&s_pcm[0][0], row_words, in, s_info, nstreams, 0, frame_bytes, true, nch, lane, role - 1);
        cp_commit();
        cp_wait_all();
    }
    group_sync(group);
    if (role >= 1 && ntiles) first_pass(0);
    group_sync(group);

    int h1 = ch.hist1, h2 = ch.hist2;
#ifdef CRI_ADX_TIMING
    long long t_work = 0, t_wait = 0;
#endif
    for (uint32_t t = 0; t < ntiles; t++) {
        const int buf = t & 1;
        const uint32_t b0 = t * kTile;
#ifdef CRI_ADX_TIMING
        const long long t_begin = clock64();
#endif
        if (role >= 1) {
            if (t + 1 < ntiles) mover_request(&s_pcm[buf ^ 1][0], row_words, in, s_info, nstreams, b0 + kTile, frame_bytes, true, nch, lane, role - 1);
            cp_commit();
            if (t >= 1) store_tile(t - 1);
            cp_wait_all();
            asm volatile("bar.sync %0, %1;" ::"r"(5 + group), "n"(32 * kMovers) : "memory");   // every mover's part of tile t + 1 has landed
            if (t + 1 < ntiles) first_pass(t + 1);
        } else {
            const uint32_t nb = ch.blocks > b0 ? min((uint32_t)kTile, ch.blocks - b0) : 0u;
            const uint8_t* row = reinterpret_cast<const uint8_t*>(&s_pcm[buf][slot * row_words]) +
                                 (int)((stream_base + (uint64_t)b0 * frame_bytes) & 15);
            for (uint32_t tb = 0; tb < nb; tb++) {
                uint8_t* dst = &s_code[buf][slot * code_row + (tb * nch + ch.channel) * kBlk];
                const BlockHead hd = stage.head[buf][tb * 32 + lane];
                // this block's samples (the movers have zeroed what lies past the end of the stream)
                int smp[kSpb];
#pragma unroll
                for (int i = 0; i < kSpb; i++)
                    smp[i] = *(reinterpret_cast<const int16_t*>(row) + (size_t)(tb * kSpb + i) * nch + ch.channel);
                // pass 1 (adx.cpp:221-230): the first two residuals see the history the previous block left, the others came
                // from the movers
                const int r0 = (smp[0] * 4096 - c0 * h1 - c1 * h2) >> 12;
                const int r1 = (smp[1] * 4096 - c0 * smp[0] - c1 * h1) >> 12;
                const int mn = min(min(hd.mn, r0), r1), mx = max(max(hd.mx, r0), r1);
                if (mn == 0 && mx == 0) {  // silent residual: all-zero block, history stays raw (adx.cpp:231-234)
#pragma unroll
                    for (int k = 0; k < kBlk; k++) dst[k] = 0;
                    h1 = smp[kSpb - 1]; h2 = smp[kSpb - 2];
                    continue;
                }
                const ScaleChoice sc = choose_scale(mn, mx, limit, ch.mode, ch.filter);
                dst[0] = (uint8_t)(sc.word >> 8);
                dst[1] = (uint8_t)sc.word;
                const int scale = sc.scale ? sc.scale : 1;  // adx.cpp:256-257
                const int half = scale >> 1;
                // pass 2 (adx.cpp:254-266), speculative form. With t = (s << 12) - c0*h1 - c1*h2 the reference computes
                //   delta = trunc((t >> 12 -+ half) / scale) clamped to [-8, 7],  h1' = clamp16(((delta << 12) * scale + pred) >> 12).
                // (delta << 12) * scale is a multiple of 4096, so h1' = clamp16(delta * scale + (pred >> 12)), and
                // pred >> 12 = s - ceil(t / 4096). The int16 clamp does not fire on ordinary audio: the block is first run
                // without it (9 dependent operations per sample instead of 14; `range_s` collects the biased values), and
                // redone with the reference's own sequence from the saved history if a value left the int16 range. (The
                // delta clamp stays: the scale comes from residuals against RAW history, so it fires in ~1 block of 8.)
                bool exact = abs(c0) > 0x2000 || abs(c1) > 0x2000 || scale > 0x1000;
                if (!exact) {
                    // q = floor((|r| + half) / scale) as the upper word of (|r| + half) * M, M = ceil(2^32 / scale): exact
                    // while (|r| + half) * scale < 2^32. The addend half * M is a block constant, so the division is ONE
                    // wide multiply-add on the path (|r| * M + c64). scale = 1 (M would be 2^32): M = c64 = 2^32 - 1 gives
                    // upper(|r| * 2^32 + (2^32 - 1 - |r|)) = |r|.
                    const uint32_t magic = s_magic[scale];                            // 1 <= scale <= 0x1000
                    const unsigned long long c64 = (unsigned long long)(uint32_t)half * magic + (scale == 1 ? 0xFFFFFFFFull : 0ull);
                    // Serial path per sample: t -> r -> |r| -> q -> min(q, limit) -> t' (five operations). The next t is
                    //   s'*4096 - c1*sim_{n-1} - c0*sim_n  with  sim_n = delta*scale + P  =  B - (c0*scale) * delta,
                    // where B = s'*4096 - c1*sim_{n-1} - c0*P and the sign of delta (known from r, early) is folded into
                    // the multiplier, so neither the sign nor the simulated sample itself sits on the path.
                    const int a_pos = c0 * scale;
                    int range_s = 0;
                    int g1 = h1, g2 = h2;                                             // sim_{n-1}, sim_{n-2}
                    int t = (smp[0] * 4096 - c1 * g2) - c0 * g1;
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        int byte = 0;
#pragma unroll
                        for (int n = 0; n < 2; n++) {
                            const int i = 2 * k + n;
                            const int sm = smp[i];
                            const int r = t >> 12;
                            const bool neg = r < 0;
                            const int q = (int)(((unsigned long long)(uint32_t)abs(r) * magic + c64) >> 32);
                            const int qc = min(q, neg ? 8 : 7);                       // clamp of delta to [-8, 7], on the magnitude
                            const int pshift = sm - ((t + 4095) >> 12);               // pred >> 12
                            const int d = neg ? -qc : qc;
                            const int sim = d * scale + pshift;
                            range_s |= sim + 32768;
                            if (i < kSpb - 1) {         // one multiply-add behind qc: everything else of the next t is ready by then
                                const int ready = smp[i + 1] * 4096 - c1 * g1 - c0 * pshift, step = neg ? a_pos : -a_pos;
                                asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(t) : "r"(qc), "r"(step), "r"(ready));
                            }
                            g2 = g1; g1 = sim;
                            byte = (byte << 4) | (d & 0xF);
                        }
                        dst[2 + k] = (uint8_t)byte;
                    }
                    if ((unsigned)range_s > 0xFFFFu) exact = true;
                    else { h1 = g1; h2 = g2; }
                }
                if (exact) {
                    const uint32_t magic = div_magic(scale);
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        int byte = 0;
#pragma unroll
                        for (int n = 0; n < 2; n++) {
                            const int pred = c0 * h1 + c1 * h2;
                            int d = (smp[2 * k + n] * 4096 - pred) >> 12;
                            d = d > 0 ? d + half : d - half;
                            d = div_trunc(d, magic, scale);
                            d = min(max(d, -8), 7);
                            const int sim = clamp16((d * 4096 * scale + pred) >> 12);
                            h2 = h1; h1 = sim;
                            byte = (byte << 4) | (d & 0xF);
                        }
                        dst[2 + k] = (uint8_t)byte;
                    }
                }
            }
        }
#ifdef CRI_ADX_TIMING
        const long long t_mid = clock64();
        group_sync(group);
        t_work += t_mid - t_begin;
        t_wait += clock64() - t_mid;
#else
        group_sync(group);
#endif
    }
    if (role >= 1 && ntiles) store_tile(ntiles - 1);
#ifdef CRI_ADX_TIMING
    if (blockIdx.x == 0 && lane == 0 && group == 0)
        printf("adx encode cta %d group %d role %d: tiles %u work %lld wait %lld cycles\n", blockIdx.x, group, role, ntiles, t_work, t_wait);
#endif
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

__global__ void gather_segments_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src,
                                       const Segment* __restrict__ segs, uint32_t n) {
    const uint32_t p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= n) return;
    const Segment sg = segs[p];
    for (uint32_t i = threadIdx.x & 31; i < sg.bytes; i += 32) dst[sg.dst_off + i] = src[sg.src_off + i];
}

// ------------------------------------------------------------ WAV ingest
// One CTA per (stream, chunk of kConvChunk samples); conversion rules of the reference: PCM8_to_PCM16, PCM_to_PCM16,
// Float_to_PCM (pcm.cpp:455-527), with the x86 (int) cast of an out-of-range float (INT_MIN) restated explicitly.
constexpr uint32_t kConvChunk = 8192;
__global__ void pcm_convert_kernel(uint8_t* blob, const PcmConv* __restrict__ conv, uint32_t n) {
    const uint32_t s = blockIdx.y;
    if (s >= n) return;
    const PcmConv c = conv[s];
    const uint32_t begin = blockIdx.x * kConvChunk;
    if (begin >= c.count) return;
    const uint32_t end = min(begin + kConvChunk, c.count);
    const uint8_t* src = blob + c.src_off;
    int16_t* dst = reinterpret_cast<int16_t*>(blob + c.dst_off);
    for (uint32_t j = begin + threadIdx.x; j < end; j += blockDim.x) {
        if (c.kind == 2) { dst[j] = 0; continue; }
        const uint32_t i = c.kind == 1 ? j % c.channels : j;
        int v;
        if (c.format == 0) {
            v = (int)(int16_t)((uint32_t)src[2ull * i] | ((uint32_t)src[2ull * i + 1] << 8));
        } else if (c.format == 1) {
            v = ((int)src[i] - (1 << (c.shift - 1))) << 8;
        } else if (c.format == 2) {
            const uint8_t* p = src + 3ull * i;
            int w = (int)p[0] | ((int)p[1] << 8) | ((int)p[2] << 16);
            if (w & 0x800000) w |= ~0xFFFFFF;
            v = w >> c.shift;
        } else if (c.format == 3) {
            const uint8_t* p = src + 4ull * i;
            const int w = (int)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
            v = w >> 16;
        } else if (c.format == 4) {
            const uint8_t* p = src + 4ull * i;
            const float f = __uint_as_float((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
            const float x = __fmul_rn(f, 32767.0f);
            v = fabsf(x) < 2147483648.0f ? __float2int_rz(x) : INT_MIN;
            v = clamp16(v);
        } else {
            const uint8_t* p = src + 8ull * i;
            unsigned long long bits = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) bits |= (unsigned long long)p[k] << (8 * k);
            const double x = __dmul_rn(__longlong_as_double((long long)bits), 32767.0);
            v = fabs(x) < 2147483648.0 ? __double2int_rz(x) : INT_MIN;
            v = clamp16(v);
        }
        dst[j] = (int16_t)v;
    }
}

}  // namespace

void launch_pcm_convert(uint8_t* d_blob, const PcmConv* d_conv, uint32_t n, uint32_t max_count, cudaStream_t s, uint64_t* launches) {
    if (!n) return;
    // grid.x covers the longest stream; CTAs past the end of a shorter one return at once
    const uint32_t chunks = std::max<uint32_t>(1, (max_count + kConvChunk - 1) / kConvChunk);
    for (uint32_t first = 0; first < n; first += 65535) {
        const uint32_t cnt = std::min<uint32_t>(65535, n - first);
        pcm_convert_kernel<<<dim3(chunks, cnt), 256, 0, s>>>(d_blob, d_conv + first, cnt);
    }
    ++*launches;
}

void launch_adx_decode(const uint8_t* d_in, uint8_t* d_out, const AdxChain* d_chains, uint32_t n_fast, uint32_t n_generic,
                       cudaStream_t s, uint64_t* launches) {
    if (n_fast) {
        const unsigned groups = (n_fast + 31) / 32;
        cudaFuncSetAttribute(adx_decode_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kGroups * sizeof(DecodeStage)));
        adx_decode_fast_kernel<<<(groups + kGroups - 1) / kGroups, kGroups * kDecGroupThreads, kGroups * sizeof(DecodeStage), s>>>(d_in, d_out, d_chains, n_fast);
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
        const unsigned groups = (n_fast + 31) / 32;
        const size_t smem_e = kGroups * sizeof(EncodeStage) + (0x1000 + 4) * sizeof(uint32_t);   // stages + the reciprocal table
        cudaFuncSetAttribute(adx_encode_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_e);
        adx_encode_fast_kernel<<<(groups + kGroups - 1) / kGroups, kGroups * kGroupThreads, smem_e, s>>>(d_in, d_out, d_chains, n_fast);
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

void launch_gather_segments(uint8_t* d_dst, const uint8_t* d_src, const Segment* d_segs, uint32_t n, cudaStream_t s,
                            uint64_t* launches) {
    if (!n) return;
    gather_segments_kernel<<<(n + 3) / 4, 128, 0, s>>>(d_dst, d_src, d_segs, n);
    ++*launches;
}

}  // namespace cri
