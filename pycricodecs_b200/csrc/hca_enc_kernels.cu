// HCA v2.0 encode kernel for sm_100a: one WARP per frame.
//
// Reference: EncodeFrame (CriCodecs/hca.cpp:2965-2988, a VGAudio port): PCM ->
// float, windowed MDCT per subframe (:2529-2553, DCT4 :2481-2527), intensity
// stereo (:2561-2609), scalefactors (:2611-2637), scaled spectra (:2639-2654),
// HFR group scales (:2656-2706), header length (:2708-2750), noise-level and
// evaluation-boundary binary searches over the frame's bit budget
// (CalculateUsedBits :2763-2790, searches :2792-2866), quantisation (:2868-2892),
// bit packing + CRC16 (:2894-2963).
//
// Frames are independent: the only state that crosses a frame boundary is the
// previous 128 raw input samples (:2552), which this kernel re-reads. So the
// batch is a flat list of frames, one warp each. Inside the warp:
//   * the MDCT keeps the 64 complex points of the reference's DCT4 two per lane
//     (tools/gen_dct.py): one in-lane pass + five __shfl_xor passes, every
//     product and sum rounded separately (no FMA: bit parity with the reference);
//   * everything per band (scalefactors, scaled spectra, bit costs, quantised
//     codes) is owned by lane = band mod 32, so shared-memory rows are read
//     conflict-free and integer bit counts are warp reductions;
//   * the few fp32 running sums whose rounding depends on order (stereo energies,
//     HFR group averages) are accumulated by one lane per subframe / group in the
//     reference's order; the fp64 corners (sqrt(2), 1.0/avg) are evaluated in fp64;
//   * the bitstream is assembled with warp prefix sums of code lengths and
//     shared-memory atomicOr (fields of known length at computed offsets).
#include <algorithm>
#include <cstdint>
#include <type_traits>

#include "cri_tables.h"
#include "hca_kernels.h"

namespace cri {
namespace {

__constant__ uint32_t e_scaling[64] = CRI_TBL_DEC_SCALING;
__constant__ uint32_t e_qscaling[64] = CRI_TBL_ENC_Q_SCALING;
__constant__ uint8_t e_curve[59] = CRI_TBL_ENC_RES_CURVE;
__constant__ uint8_t e_qbits[128] = CRI_TBL_ENC_Q_BITS;
__constant__ uint8_t e_qcode[128] = CRI_TBL_ENC_Q_CODE;
__constant__ uint32_t e_inv_step[16] = CRI_TBL_ENC_INV_STEP;
__constant__ uint32_t e_dead_zone[16] = CRI_TBL_ENC_DEAD_ZONE;
__constant__ uint32_t e_ratio_bounds[14] = CRI_TBL_ENC_RATIO_BOUNDS;
__constant__ uint8_t e_max_bits[16] = CRI_TBL_MAX_BITS;
__constant__ uint32_t e_cost_rows[64] = CRI_TBL_ENC_COST_ROWS;
__constant__ uint32_t e_rank_keys[2 * 49 * 2] = CRI_TBL_ENC_RANK_KEYS;
__constant__ uint32_t e_rank_rows[16] = CRI_TBL_ENC_RANK_ROWS;
constexpr uint32_t kRankShift = 21, kRankBase = 459, kRankBuckets = 49;   // tools/gen_tables.py: enc_cost_ranks()

#include "hca_dct_gen.inc"

#ifndef HCA_ENC_WARPS
#define HCA_ENC_WARPS 10          // x 9.6 KB of shared memory per stereo frame: two CTAs per SM
#endif
constexpr int kEncWarps = HCA_ENC_WARPS;
#ifndef HCA_ENC_MIN_CTAS
#define HCA_ENC_MIN_CTAS (HCA_ENC_WARPS >= 20 ? 1 : 20 / HCA_ENC_WARPS)
#endif
// CONVOY: the warps of a CTA meet at every phase boundary, so that they fetch the same instruction-cache lines at about
// the same time. It paid while the kernel was ~120 KB of SASS with a per-coefficient bit-cost loop (one CTA of 20 warps:
// 23.5 -> 19.2 ms per 8192 streams); with the counted bit costs the barrier stalls cost more than the shared fetches
// save. Measured per 8192 streams: 20 warps with convoy 15.9 ms, without 14.7; 2 x 10 warps without 14.4; 3 x 7 warps
// (80 registers) 15.9; 4 x 5 warps 16.4. A warp behind the last frame redoes it and stores nothing, so the barrier stays
// legal if it is switched back on (-DHCA_ENC_CONVOY=1). What still matters is that the warps of a CTA run the same phase
// at about the same time: see the round experiments in launch_hca_encode.
#ifndef HCA_ENC_CONVOY
#define HCA_ENC_CONVOY 0
#endif
#ifndef HCA_ENC_ROUND_SYNC
#define HCA_ENC_ROUND_SYNC 0
#endif
// HCA_ENC_CONVOY is a bit mask over the nine phase boundaries (0: after the MDCT, 1: intensity, 2: scalefactors, 3: scaled
// spectra, 4: noise level, 5: boundary, 6: header, 7: quantised, 8: packed); a boundary without the CTA barrier still orders
// the warp's own shared memory.
#define CONVOY(k) do { if ((HCA_ENC_CONVOY >> (k)) & 1) __syncthreads(); else __syncwarp(); } while (0)
#ifndef HCA_ENC_RANK_UNROLL
#define HCA_ENC_RANK_UNROLL 8        // coefficients per unrolled body of the rank loop (code size against loop overhead)
#endif
constexpr int kRankUnroll = HCA_ENC_RANK_UNROLL;
constexpr int kSpecRow = 128;
constexpr unsigned kFull = 0xFFFFFFFFu;

struct EncTables {              // per-CTA shared copies (per-lane indices diverge)
    uint4 cost[16];             // per resolution: bits(-N), bits(P), 8 * full, overfull (tools/gen_tables.py: enc_cost_rows)
    uint2 rank_key[2 * 49];     // counted bit costs: [x < 0][bucket] = (interval end inside the bucket or 0, ends in higher buckets)
    uint2 rank_mask[16];        // [rank] = one in every nibble whose resolution holds a coefficient of that rank inside its interval
    uint32_t rank_row[16];      // per resolution: nibble shift | word B << 7 | 8 * full << 8 | overfull << 16
    float scaling[64];
    float qscaling[64];
    float inv_step[16];
    float dead_zone[16];
    uint8_t curve[60];
    uint16_t crc[4][256];       // CRC-16 (poly 0x8005, MSB first) of byte v followed by k zero bytes
    uint32_t qpack[128];        // [resolution * 16 + value + 8] = code length << 16 | code (prefix codebooks)
    uint8_t max_bits[16];
};

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}
__device__ __forceinline__ int warp_excl_scan(int v, int lane, int* total) {
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(kFull, x, o);
        if (lane >= o) x += y;
    }
    *total = __shfl_sync(kFull, x, 31);
    return x - v;
}

__device__ __forceinline__ int find_scalefactor(const float* table, float v) {   // hca.cpp:2611-2623
    unsigned lo = 0, hi = 63;
    while (lo < hi) {
        const unsigned mid = (lo + hi) >> 1;
        if (table[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return (int)lo;
}

__device__ __forceinline__ int enc_resolution(const EncTables& tb, int sf, int noise) {   // hca.cpp:2752-2761
    if (sf == 0) return 0;
    int pos = noise - 5 * sf / 2 + 2;
    pos = min(max(pos, 0), 58);
    return tb.curve[pos];
}

// Per-warp shared-memory view of one frame being encoded.
struct FrameSmem {
    float* spec;        // [nch][8][128] spectra, later scaled spectra
    int16_t* pcm;       // [nch][1152] previous 128 + current 1024 samples
    uint8_t* sf;        // [nch][128]
    uint8_t* res;       // [nch][128] during the bit allocation: coefficients of the band that sit on the positive clamp; then the final resolutions
    uint32_t* bits;     // packed frame, MSB-first 32-bit words
    float* hfr_avg;     // [nch][8]
    int* hfr_scale;     // [nch][8]
    uint8_t* inten;     // [nch][8]
    int* header_bits;   // [nch]
    int* delta_bits;    // [nch]
};

// Bits of one band's eight coefficients at resolution r (the per-coefficient terms of CalculateUsedBits, hca.cpp:2772-2787):
// a coefficient costs full[r] bits, one less inside (-N[r], P[r]) -- the dead zone of the sign-magnitude codes, the short
// codes of the prefix codebooks -- and nothing at all where the clamp value quantises out of the codebook (`clamped`
// = how many of the band's coefficients sit on the clamp). tests/test_tables.py checks the rows against the reference's
// quantise-and-look-up arithmetic.
__device__ __forceinline__ int band_cost(const EncTables& tb, const float* sp_band, int r, int clamped) {
    const uint4 row = tb.cost[r];
    const float neg_n = __uint_as_float(row.x), p = __uint_as_float(row.y);
    int inside = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float x = sp_band[j * kSpecRow];
        inside += (x > neg_n && x < p) ? 1 : 0;
    }
    return (int)row.z - inside - (int)row.w * clamped;
}

// Total frame bits for (noise level, evaluation boundary): CalculateUsedBits, hca.cpp:2763-2790.
__device__ __forceinline__ int used_bits(const EncTables& tb, const FrameSmem& fs, const HcaStreamDev& S, int lane, int noise_level,
                                         int boundary) {
    int len = 0;
    const int nch = S.channels;
    for (int c = 0; c < nch; c++) {
        const int coded = S.coded[c];
        const float* sp = fs.spec + (size_t)c * 8 * kSpecRow;
        for (int b = lane; b < coded; b += 32) {
            const int noise = b < boundary ? noise_level - 1 : noise_level;
            len += band_cost(tb, sp + b, enc_resolution(tb, fs.sf[c * 128 + b], noise), fs.res[c * 128 + b]);
        }
    }
    len = warp_sum(len);
    int hdr = 48;
    for (int c = 0; c < nch; c++) hdr += fs.header_bits[c];
    return len + hdr;
}

// The boundary search (BinarySearchBoundary, hca.cpp:2834-2850) only ever compares two noise levels per band: bands
// below the boundary use noise_level - 1, the others noise_level. With hi[b] / lo[b] = bits of band b (all channels)
// at noise_level / noise_level - 1,  used_bits(noise_level, m) = header + sum(hi) + sum over b < m of (lo[b] - hi[b])
// -- the same integers added in another order. One pass fills pre[m] = that value for m = 0..128; every probe of the
// search is then a lookup instead of a pass over the 2048 coefficients.
__device__ __forceinline__ void boundary_table(const EncTables& tb, const FrameSmem& fs, const HcaStreamDev& S, int lane, int noise_level,
                                               int* pre) {
    const int nch = S.channels;
    int diff[4] = {0, 0, 0, 0};
    int tot = 0;
    for (int c = 0; c < nch; c++) {
        const int coded = S.coded[c];
        const float* sp = fs.spec + (size_t)c * 8 * kSpecRow;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int b = lane + 32 * k;
            if (b < coded) {
                const int sfv = fs.sf[c * 128 + b];
                const int r_hi = enc_resolution(tb, sfv, noise_level), r_lo = enc_resolution(tb, sfv, noise_level - 1);
                const int clamped = fs.res[c * 128 + b];
                const int c_hi = band_cost(tb, sp + b, r_hi, clamped);
                const int c_lo = r_lo == r_hi ? c_hi : band_cost(tb, sp + b, r_lo, clamped);
                tot += c_hi;
                diff[k] += c_lo - c_hi;
            }
        }
    }
    int base = warp_sum(tot) + 48;
    for (int c = 0; c < nch; c++) base += fs.header_bits[c];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int total;
        const int excl = warp_excl_scan(diff[k], lane, &total);
        pre[lane + 32 * k] = base + excl;              // boundary m = lane + 32k: bands below m switched to the lower level
        base += total;
    }
    if (lane == 0) pre[128] = base;
    __syncwarp();
}

// CalculateOptimalDeltaLength + CalculateFrameHeaderLength, hca.cpp:2708-2750 (warp-collective).
__device__ __forceinline__ void header_lengths(const FrameSmem& fs, const HcaStreamDev& S, int lane) {
    const int nch = S.channels;
    for (int c = 0; c < nch; c++) {
        const int coded = S.coded[c];
        const uint8_t* sf = fs.sf + c * 128;
        // a delta width db costs db bits per band behind the first, 6 more where |delta| > 2^(db-1) - 1, i.e. where delta
        // has more than db - 1 significant bits: the five escape counts (at most 127 each) travel through the warp sums as
        // 8-bit fields, four in one word
        int any = 0;
        uint32_t esc_lo = 0, esc_5 = 0;
        for (int b = lane; b < coded; b += 32) {
            any |= sf[b] != 0;
            if (b >= 1) {
                const uint32_t delta = (uint32_t)abs((int)sf[b] - (int)sf[b - 1]);
                esc_lo += (delta > 0u ? 1u : 0u) + (delta > 1u ? 1u << 8 : 0u) + (delta > 3u ? 1u << 16 : 0u) + (delta > 7u ? 1u << 24 : 0u);
                esc_5 += delta > 15u ? 1u : 0u;
            }
        }
        any = __any_sync(kFull, any);
        esc_lo = (uint32_t)warp_sum((int)esc_lo);
        esc_5 = (uint32_t)warp_sum((int)esc_5);
        int best_bits = 6, best_len = 3 + 6 * coded;
#pragma unroll
        for (int db = 1; db < 6; db++) {
            const int escapes = db < 5 ? (int)((esc_lo >> (8 * (db - 1))) & 0xFFu) : (int)esc_5;
            const int len = 3 + 6 + db * max(coded - 1, 0) + 6 * escapes;
            if (len < best_len) { best_len = len; best_bits = db; }
        }
        if (!any) { best_len = 3; best_bits = 0; }
        if (S.type[c] == 2) best_len += 32;
        else if (S.hfr_groups > 0) best_len += 6 * S.hfr_groups;
        if (lane == 0) { fs.header_bits[c] = best_len; fs.delta_bits[c] = best_bits; }
    }
    __syncwarp();
}

// The last s with prefix[s] <= f, searched by the whole warp: 32 probes spread over the open interval per step (three
// dependent loads for 8192 streams where a bisection takes thirteen).
__device__ __forceinline__ uint32_t find_stream_warp(const uint64_t* __restrict__ prefix, uint32_t n, uint64_t f, int lane) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const uint32_t probe = lo + (uint32_t)(((uint64_t)(hi - lo) * (uint32_t)(lane + 1)) / 33u);      // lo <= probe < hi, non-decreasing in lane
        const int k = __popc(__ballot_sync(kFull, prefix[probe] <= f));                                  // probes 0 .. k - 1 hold
        const uint32_t below = __shfl_sync(kFull, probe, max(k - 1, 0)), above = __shfl_sync(kFull, probe, min(k, 31));
        lo = k ? below : lo;
        hi = k < 32 ? above : hi;
    }
    return lo;
}

// COUNTED (mono / stereo batches): the bit-allocation search does not go back to the 2048 scaled coefficients for every
// probe. Each coefficient is classified ONCE, when it is scaled: its rank r* (tools/gen_tables.py: enc_cost_ranks) says
// in which resolutions' short-code intervals it lies, and a band keeps, per resolution, how many of its eight
// coefficients do -- fifteen 4-bit counts in two registers per band (word A: resolutions 1..7 in sorted order + the
// number of coefficients on the positive clamp in nibble 7; word B: resolutions 8..15). A probe of CalculateUsedBits
// (hca.cpp:2763-2790) is then, per band, one table row and one nibble: the same integers as the per-coefficient count.
template <bool COUNTED>
__global__ void __launch_bounds__(kEncWarps * 32, HCA_ENC_MIN_CTAS)
hca_encode_kernel(HcaEncodeArgs a) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ EncTables tb;
    __shared__ uint32_t s_first_stream;
    if (threadIdx.x < 32) {                                   // stream of the CTA's first frame (round 0)
        const uint64_t f0 = min((uint64_t)blockIdx.x * (blockDim.x >> 5), a.n_frames - 1);
        const uint32_t s0 = find_stream_warp(a.frame_prefix, a.n_streams, f0, (int)threadIdx.x);
        if (threadIdx.x == 0) s_first_stream = s0;
    }
    if (COUNTED) {
        for (int i = threadIdx.x; i < 2 * (int)kRankBuckets; i += blockDim.x) tb.rank_key[i] = make_uint2(e_rank_keys[2 * i], e_rank_keys[2 * i + 1]);
        for (int i = threadIdx.x; i < 16; i += blockDim.x) {
            const int na = min(i, 7), nb = max(i - 7, 0);
            tb.rank_mask[i] = make_uint2(na ? 0x01111111u >> (28 - 4 * na) : 0u, nb ? 0x11111111u >> (32 - 4 * nb) : 0u);
            const uint32_t row = e_rank_rows[i], pos = (row & 63u) >> 2;       // position in the sorted order; 15 = resolution 0
            const uint32_t where = pos < 7 ? 4 * pos : pos < 15 ? (4 * (pos - 7)) | 0x80u : 28u;
            tb.rank_row[i] = (row & ~63u) | where;
        }
    }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) {
        tb.scaling[i] = __uint_as_float(e_scaling[i]);
        tb.qscaling[i] = __uint_as_float(e_qscaling[i]);
    }
    for (int i = threadIdx.x; i < 16; i += blockDim.x) {
        tb.cost[i] = make_uint4(e_cost_rows[4 * i], e_cost_rows[4 * i + 1], e_cost_rows[4 * i + 2], e_cost_rows[4 * i + 3]);
        tb.inv_step[i] = __uint_as_float(e_inv_step[i]);
        tb.dead_zone[i] = __uint_as_float(e_dead_zone[i]);
        tb.max_bits[i] = e_max_bits[i];
    }
    for (int i = threadIdx.x; i < 59; i += blockDim.x) tb.curve[i] = e_curve[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        auto step = [](uint32_t c, uint32_t byte) {             // table-free byte step
            const uint32_t v = ((c >> 8) ^ byte) & 0xFF;
            return ((c << 8) ^ (v << 1) ^ (v << 2) ^ ((__popc(v) & 1) ? 0x8003u : 0u)) & 0xFFFFu;
        };
        uint32_t c = step(0, (uint32_t)i);
#pragma unroll
        for (int k = 0; k < 4; k++) { tb.crc[k][i] = (uint16_t)c; c = step(c, 0); }
    }
    for (int i = threadIdx.x; i < 128; i += blockDim.x) {
        tb.qpack[i] = ((uint32_t)e_qbits[i] << 16) | ((uint32_t)e_qcode[i] & ((1u << e_qbits[i]) - 1u));
    }
    __syncthreads();

    // Every warp walks the frame list with the stride of the whole grid (one round with the default launch: a CTA per
    // kEncWarps frames; see launch_hca_encode). With the convoy barrier all warps make the same number of rounds: a
    // warp behind the last frame redoes it and stores nothing.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#ifdef HCA_ENC_STAGGER_NS           // experiment: resident CTAs whose warps start one after the other, spread over one frame time
    __nanosleep((unsigned)(((blockIdx.x / 148u) * (blockDim.x >> 5) + warp) % 20u) * HCA_ENC_STAGGER_NS);
#endif
    const uint64_t grid_warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const uint32_t rounds = grid_warps >= a.n_frames ? 1u : (uint32_t)((a.n_frames + grid_warps - 1) / grid_warps);   // (64-bit division: rare path)
    for (uint32_t round = 0; round < rounds; round++) {
    const uint64_t f_own = round * grid_warps + (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    const bool surplus = f_own >= a.n_frames;
#if HCA_ENC_ROUND_SYNC
    __syncthreads();                                          // the CTA's warps start every round together (shared instruction fetch)
#else
    if (surplus && !HCA_ENC_CONVOY) break;
#endif
    const uint64_t f = surplus ? a.n_frames - 1 : f_own;
    // the frame's stream = the last s with frame_prefix[s] <= f. In the first round the CTA's frames are neighbours: warp 0
    // has looked the first one up while the tables were built, the others walk on from there (usually zero or one step).
    uint32_t lo = round == 0 ? s_first_stream : find_stream_warp(a.frame_prefix, a.n_streams, f, lane);
    while (lo + 1 < a.n_streams && a.frame_prefix[lo + 1] <= f) lo++;
    const uint32_t stream = lo;
    const uint32_t mul = a.crc_mul[(size_t)stream * 32 + lane];      // this lane's CRC chunk multiplier (needed last, requested first)
    const HcaStreamDev& S = a.streams[stream];
    const uint32_t frame = (uint32_t)(f - a.frame_prefix[stream]);
    const int nch = S.channels;
    const int MC = (int)a.max_channels;

    // carve the warp's shared memory
    FrameSmem fs;
    {
        uint8_t* p = s_dyn + (size_t)warp * a.smem_per_warp;
        fs.spec = reinterpret_cast<float*>(p); p += (size_t)MC * 8 * kSpecRow * 4;
        fs.hfr_avg = reinterpret_cast<float*>(p); p += (size_t)MC * 8 * 4;
        fs.hfr_scale = reinterpret_cast<int*>(p); p += (size_t)MC * 8 * 4;
        fs.header_bits = reinterpret_cast<int*>(p); p += (size_t)MC * 4;
        fs.delta_bits = reinterpret_cast<int*>(p); p += (size_t)MC * 4;
        // scratch: the boundary table (129 words) during the bit allocation, then the packed frame
        fs.pcm = reinterpret_cast<int16_t*>(p);
        fs.bits = reinterpret_cast<uint32_t*>(p);
        {
            const size_t tab_bytes = 129 * 4 + 12, bit_bytes = (size_t)a.frame_words * 4;
            p += (tab_bytes > bit_bytes ? tab_bytes : bit_bytes + 15) / 16 * 16;
        }
        fs.sf = p; p += (size_t)MC * 128;
        fs.res = p; p += (size_t)MC * 128;
        fs.inten = p;
    }
    const int frame_size = (int)S.frame_size;

    // ---- PCM: the MDCT of subframe `sub` reads sample frames [1024 * frame - 128 + 128 * sub, + 256) straight from the
    // blob (previous 128 + current 128; 2-byte loads that L1 serves after the first touch of a line); outside the
    // stream there is silence (hca.cpp:3035-3053)
    const int16_t* pcm_base = reinterpret_cast<const int16_t*>(a.in + S.in_off);
    const long long pcm_n0 = (long long)frame * 1024 - 128;             // stream index of the frame's first wanted sample frame
    const long long pcm_end = (long long)S.out_samples;
    // a frame whose whole window lies inside the stream (all but the first and the last one) loads without bounds checks
    const bool interior = pcm_n0 >= 0 && pcm_n0 + 1152 <= pcm_end;
    if (interior) {                                           // pull the window towards L2 while the first block is on its way
        const uint8_t* w0p = reinterpret_cast<const uint8_t*>(pcm_base + pcm_n0 * nch);
        for (int k = lane * 128; k < 1152 * 2 * nch; k += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(w0p + k));
    }
    auto sample = [&](int idx /* sample frame inside the 1152-frame window */, int c) -> float {
        const long long n = pcm_n0 + idx;
        const int v = (interior || (n >= 0 && n < pcm_end)) ? (int)__ldg(pcm_base + n * nch + c) : 0;
        return (float)v;
    };
    // ---- MDCT: 8 subframes x channels (hca.cpp:2470-2559)
    {
        // PcmToFloat's factor 1/32768 (hca.cpp:2470-2479) is folded into the window: w * (s * 2^-15) and (w * 2^-15) * s
        // are the same real number rounded once (both scalings are exact), so the products are bit-identical
        const float k = 1.0f / 32768.0f;
        const float w0 = __fmul_rn(__uint_as_float(kMdctWin[4 * lane]), k), w1 = __fmul_rn(__uint_as_float(kMdctWin[4 * lane + 1]), k);
        const float w2 = __fmul_rn(__uint_as_float(kMdctWin[4 * lane + 2]), k), w3 = __fmul_rn(__uint_as_float(kMdctWin[4 * lane + 3]), k);
        const float pc0 = __uint_as_float(kMdctPreCos[lane]), ps0 = __uint_as_float(kMdctPreSin[lane]);
        const float pc1 = __uint_as_float(kMdctPreCos[lane + 32]), ps1 = __uint_as_float(kMdctPreSin[lane + 32]);
        // Passes 1..5 exchange with lane ^ (32 >> s). The lane with that bit set ("back") takes other - mine and rotates
        // it; the front lane takes mine + other as it is. Both are ONE instruction stream: the sum is fma(mine, +-1, other)
        // (the product is exact, so this is the add or the subtract, rounded once), and a front lane rotates by
        // (cos, sin, "-cos") = (1, 0, 1): x * 1 + y * 0 and x * 0 + y * 1 return x and y unchanged up to the sign of a
        // zero, which nothing downstream can tell apart (magnitudes, comparisons against non-zero bounds, products).
        float tc[6], ts[6], tn[6];
#pragma unroll
        for (int s = 0; s < 6; s++) {
            const bool front = s > 0 && !(lane & (32 >> s));
            tc[s] = front ? 1.0f : __uint_as_float(kMdctCos[s * 32 + lane]);
            ts[s] = front ? 0.0f : __uint_as_float(kMdctSin[s * 32 + lane]);
            tn[s] = front ? 1.0f : -tc[s];
        }
        const int d0 = kMdctDest[4 * lane], d1 = kMdctDest[4 * lane + 1], d2 = kMdctDest[4 * lane + 2], d3 = kMdctDest[4 * lane + 3];
        const int i0 = 2 * lane, i1 = 63 - 2 * lane, i2 = 64 + 2 * lane, i3 = 127 - 2 * lane;
        float sg[6];
#pragma unroll
        for (int s = 1; s < 6; s++) sg[s] = (lane & (32 >> s)) ? -1.0f : 1.0f;
        // INTERIOR (all but the first and the last frame of a stream: the whole window lies inside the stream): a sample
        // is one address and one 2-byte load; the other frames check every index against the stream's ends
        auto mdct_channels = [&](auto interior_tag) {
            constexpr bool INTERIOR = decltype(interior_tag)::value;
            const int e0 = i0 * nch, e1 = i1 * nch, e2 = i2 * nch, e3 = i3 * nch;
            for (int c = 0; c < nch; c++) {
                const int16_t* qc = pcm_base + pcm_n0 * nch + c;          // window index 0 (in front of the blob for frame 0: INTERIOR only)
                auto block = [&](int at, float& v0, float& v1, float& v2, float& v3) {     // the four samples of this lane, window position `at`
                    if (INTERIOR) {
                        const int16_t* q = qc + at * nch;
                        v0 = (float)(int)__ldg(q + e0); v1 = (float)(int)__ldg(q + e1);
                        v2 = (float)(int)__ldg(q + e2); v3 = (float)(int)__ldg(q + e3);
                    } else {
                        v0 = sample(at + i0, c); v1 = sample(at + i1, c); v2 = sample(at + i2, c); v3 = sample(at + i3, c);
                    }
                };
                // the block in front of a subframe is the previous subframe's own block: four loads per subframe, not
                // eight; the next subframe's block is requested before this one's butterflies
                float p0, p1, p2, p3, n0, n1, n2, n3;
                block(0, p0, p1, p2, p3);
                block(128, n0, n1, n2, n3);
                for (int sub = 0; sub < 8; sub++) {
                    const float c0 = n0, c1 = n1, c2 = n2, c3 = n3;
                    if (sub < 7) block(256 + sub * 128, n0, n1, n2, n3);
                    // windowing (hca.cpp:2537-2546): in[i] = W[63-i]*(-cur[64+i]) - (-W[64+i])*cur[63-i],
                    //                               in[64+i] = W[i]*prv[i] - (-W[127-i])*prv[127-i]
                    const float in_a = __fsub_rn(__fmul_rn(w1, -c2), __fmul_rn(-w2, c1));   // in[2l]
                    const float in_b = __fsub_rn(__fmul_rn(w1, p1), __fmul_rn(-w2, p2));    // in[127-2l]
                    const float in_c = __fsub_rn(__fmul_rn(w0, p0), __fmul_rn(-w3, p3));    // in[64+2l]
                    const float in_d = __fsub_rn(__fmul_rn(w0, -c3), __fmul_rn(-w3, c0));   // in[63-2l]
                    // pre-rotation (hca.cpp:2490-2498): z[k] from in[2k], in[127-2k]
                    float re0 = __fadd_rn(__fmul_rn(in_a, pc0), __fmul_rn(in_b, ps0));
                    float im0 = __fsub_rn(__fmul_rn(in_a, ps0), __fmul_rn(in_b, pc0));
                    float re1 = __fadd_rn(__fmul_rn(in_c, pc1), __fmul_rn(in_d, ps1));
                    float im1 = __fsub_rn(__fmul_rn(in_c, ps1), __fmul_rn(in_d, pc1));
                    // pass 0: z[l] with z[l+32], in lane
                    {
                        const float ar = __fsub_rn(re0, re1), ai = __fsub_rn(im0, im1);
                        re0 = __fadd_rn(re0, re1); im0 = __fadd_rn(im0, im1);
                        re1 = __fadd_rn(__fmul_rn(ar, tc[0]), __fmul_rn(ai, ts[0]));
                        im1 = __fsub_rn(__fmul_rn(ar, ts[0]), __fmul_rn(ai, tc[0]));
                    }
                    // passes 1..5: partner lane = lane ^ (32 >> s); the lane with that bit set holds the "back" element
#pragma unroll
                    for (int s = 1; s < 6; s++) {
                        const int d = 32 >> s;
                        const float o_re0 = __shfl_xor_sync(kFull, re0, d), o_im0 = __shfl_xor_sync(kFull, im0, d);
                        const float o_re1 = __shfl_xor_sync(kFull, re1, d), o_im1 = __shfl_xor_sync(kFull, im1, d);
                        // front: mine + other; back: other - mine
                        const float a_re0 = __fmaf_rn(re0, sg[s], o_re0), a_im0 = __fmaf_rn(im0, sg[s], o_im0);
                        const float a_re1 = __fmaf_rn(re1, sg[s], o_re1), a_im1 = __fmaf_rn(im1, sg[s], o_im1);
                        re0 = __fadd_rn(__fmul_rn(a_re0, tc[s]), __fmul_rn(a_im0, ts[s]));
                        im0 = __fadd_rn(__fmul_rn(a_re0, ts[s]), __fmul_rn(a_im0, tn[s]));
                        re1 = __fadd_rn(__fmul_rn(a_re1, tc[s]), __fmul_rn(a_im1, ts[s]));
                        im1 = __fadd_rn(__fmul_rn(a_re1, ts[s]), __fmul_rn(a_im1, tn[s]));
                    }
                    float* sp = fs.spec + ((size_t)c * 8 + sub) * kSpecRow;
                    sp[d0] = __fmul_rn(re0, 0.125f); sp[d1] = __fmul_rn(im0, 0.125f);
                    sp[d2] = __fmul_rn(re1, 0.125f); sp[d3] = __fmul_rn(im1, 0.125f);
                    p0 = c0; p1 = c1; p2 = c2; p3 = c3;
                }
            }
        };
        if (interior) mdct_channels(std::true_type{}); else mdct_channels(std::false_type{});
    }
    CONVOY(0);

    // ---- intensity stereo (hca.cpp:2561-2609): one lane per subframe accumulates the energies in band order
    if (S.stereo_bands > 0) {
        for (int c = 0; c + 1 < nch; c++) {
            if (S.type[c] != 1) continue;
            float ratio = 1.0f;
            if (lane < 8) {
                const float* l = fs.spec + ((size_t)c * 8 + lane) * kSpecRow;
                const float* r = fs.spec + ((size_t)(c + 1) * 8 + lane) * kSpecRow;
                float el = 0.f, er = 0.f, et = 0.f;
                for (int b = S.base_bands; b < S.total_bands; b++) {
                    el = __fadd_rn(el, fabsf(l[b]));
                    er = __fadd_rn(er, fabsf(r[b]));
                    et = __fadd_rn(et, fabsf(__fadd_rn(l[b], r[b])));
                }
                et = __fmul_rn(et, 2.0f);
                const float elr = __fadd_rn(er, el);
                const float stored = __fdiv_rn(__fmul_rn(2.0f, el), elr);
                ratio = __fdiv_rn(elr, et);
                const double half_sqrt2 = 1.4142135623730951 / 2;
                if ((double)ratio < 0.5) ratio = 0.5f;
                else if ((double)ratio > half_sqrt2) ratio = (float)half_sqrt2;
                int q = 1;
                if (er > 0.f || el > 0.f) {
                    while (q < 13 && __uint_as_float(e_ratio_bounds[q]) >= stored) q++;
                } else {
                    q = 0;
                    ratio = 1.0f;
                }
                fs.inten[(c + 1) * 8 + lane] = (uint8_t)q;
            }
            for (int sub = 0; sub < 8; sub++) {
                const float rt = __shfl_sync(kFull, ratio, sub);
                float* l = fs.spec + ((size_t)c * 8 + sub) * kSpecRow;
                float* r = fs.spec + ((size_t)(c + 1) * 8 + sub) * kSpecRow;
                for (int b = S.base_bands + lane; b < S.total_bands; b += 32) {
                    l[b] = __fmul_rn(__fadd_rn(l[b], r[b]), rt);
                    r[b] = 0.f;
                }
            }
        }
        __syncwarp();
    }

    CONVOY(1);
    // ---- scalefactors (hca.cpp:2625-2637)
    for (int c = 0; c < nch; c++) {
        const int coded = S.coded[c];
        const float* sp = fs.spec + (size_t)c * 8 * kSpecRow;
        // the lane's four bands are searched side by side: 63 table entries = always six halvings, so the four
        // bisections (find_scalefactor) run in lock step as independent chains of shared-memory loads
        float mx[4];
        unsigned lo4[4], hi4[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int b = lane + 32 * k;
            mx[k] = 0.f;
            if (b < coded) {
#pragma unroll
                for (int j = 0; j < 8; j++) { const float v = fabsf(sp[j * kSpecRow + b]); mx[k] = mx[k] < v ? v : mx[k]; }
            }
            lo4[k] = 0; hi4[k] = 63;
        }
#pragma unroll
        for (int step = 0; step < 6; step++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned mid = (lo4[k] + hi4[k]) >> 1;
                const bool le = tb.scaling[mid] <= mx[k];
                lo4[k] = le ? mid + 1 : lo4[k];
                hi4[k] = le ? hi4[k] : mid;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) fs.sf[c * 128 + lane + 32 * k] = lane + 32 * k < coded ? (uint8_t)lo4[k] : (uint8_t)0;
    }
    __syncwarp();

    // ---- HFR group averages on the raw spectra (hca.cpp:2656-2674), one lane per group, reference order
    const int hfr_start = S.stereo_bands + S.base_bands;
    const int hfr_band_count = S.total_bands - S.base_bands - S.stereo_bands;
    if (S.hfr_groups > 0) {
        for (int c = 0; c < nch; c++) {
            if (S.type[c] == 2) continue;
            if (lane < S.hfr_groups) {
                const float* sp = fs.spec + (size_t)c * 8 * kSpecRow;
                float sum = 0.f;
                int count = 0;
                int band = hfr_start + lane * S.bands_per_hfr;
                for (int i = 0; i < S.bands_per_hfr && band < 128; band++, i++) {
                    for (int j = 0; j < 8; j++) sum = __fadd_rn(sum, fabsf(sp[j * kSpecRow + band]));
                    count += 8;
                }
                fs.hfr_avg[c * 8 + lane] = __fdiv_rn(sum, (float)count);
            }
        }
        __syncwarp();
    }

    CONVOY(2);
    // ---- scaled spectra, in place (hca.cpp:2639-2654)
    uint32_t cnt_a[2][4] = {}, cnt_b[2][4] = {};    // COUNTED: per band (channel c, band lane + 32 k) the interval counts
    if constexpr (COUNTED) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int coded = c < nch ? (int)S.coded[c] : 0;
            float* sp = fs.spec + (size_t)c * 8 * kSpecRow;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int b = lane + 32 * k;
                uint32_t wa = 0, wb = 0;
                if (b < coded) {
                    const int sfv = fs.sf[c * 128 + b];
                    const float ks = sfv ? tb.qscaling[sfv] : 0.f;       // no scalefactor: every coefficient becomes (+-) 0
#pragma unroll kRankUnroll
                    for (int j = 0; j < 8; j++) {
                        const float v = fminf(fmaxf(__fmul_rn(sp[j * kSpecRow + b], ks), -0.9999999f), 0.9999999f);
                        sp[j * kSpecRow + b] = v;
                        const uint32_t u = __float_as_uint(v), au = u & 0x7FFFFFFFu;
                        const uint2 key = tb.rank_key[max(au >> kRankShift, kRankBase) - kRankBase + (u >> 31) * kRankBuckets];
                        const uint2 m = tb.rank_mask[key.y + (au < key.x ? 1u : 0u)];
                        wa += m.x + (v == 0.9999999f ? 1u << 28 : 0u);
                        wb += m.y;
                    }
                }
                cnt_a[c][k] = wa; cnt_b[c][k] = wb;
            }
        }
    }
    for (int c = 0; c < nch && !COUNTED; c++) {
        const int coded = S.coded[c];
        float* sp = fs.spec + (size_t)c * 8 * kSpecRow;
        for (int b = lane; b < coded; b += 32) {
            const int sfv = fs.sf[c * 128 + b];
            const float ks = tb.qscaling[sfv];
            int clamped = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                float v = __fmul_rn(sp[j * kSpecRow + b], ks);
                v = v > 0.9999999f ? 0.9999999f : v < -0.9999999f ? -0.9999999f : v;
                v = sfv == 0 ? 0.f : v;
                clamped += v == 0.9999999f ? 1 : 0;
                sp[j * kSpecRow + b] = v;
            }
            fs.res[c * 128 + b] = (uint8_t)clamped;
        }
        for (int b = coded + lane; b < 128; b += 32) fs.res[c * 128 + b] = 0;
    }
    __syncwarp();

    // ---- HFR scales (hca.cpp:2676-2706)
    if (S.hfr_groups > 0) {
        const int lim = min(hfr_band_count, (int)S.total_bands - hfr_band_count);
        for (int c = 0; c < nch; c++) {
            if (S.type[c] == 2) continue;
            if (lane < S.hfr_groups) {
                const float* sp = fs.spec + (size_t)c * 8 * kSpecRow;
                float sum = 0.f;
                int count = 0;
                int band = lane * S.bands_per_hfr;            // each earlier group consumed bands_per_hfr mirrored bands ...
                if (band > lim) band = lim;                   // ... unless it ran into the limit
                for (int i = 0; i < S.bands_per_hfr && band < lim; band++, i++) {
                    for (int j = 0; j < 8; j++) sum = __fadd_rn(sum, fabsf(sp[j * kSpecRow + (hfr_start - band - 1)]));
                    count += 8;
                }
                const float avg = __fdiv_rn(sum, (float)count);
                float g = fs.hfr_avg[c * 8 + lane];
                if (avg > 0.0f) {
                    double m = 1.0 / (double)avg;
                    const double s2 = 1.4142135623730951;
                    if (s2 < m) m = s2;
                    g = (float)((double)g * m);
                }
                fs.hfr_scale[c * 8 + lane] = find_scalefactor(tb.scaling, g);
            }
        }
        __syncwarp();
    }

    CONVOY(3);
    // ---- bit allocation (hca.cpp:2809-2866)
    header_lengths(fs, S, lane);
    const int avail = frame_size * 8;
    int noise_level, boundary = 0;
    bool failed = false;
    // COUNTED: what a probe needs of a band besides its counts -- the scalefactor's share of the resolution index
    // (hca.cpp:2752-2761) and the row stride (0 for bands without a scalefactor: row 0 costs nothing)
    int band_off[2][4] = {}, band_nz4[2][4] = {}, hdr_total = 48;
    bool any_clamped = false;
    auto load_bands = [&]() {
        if constexpr (COUNTED) {
            bool cl = false;
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int coded = c < nch ? (int)S.coded[c] : 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int b = lane + 32 * k;
                    const int sfv = b < coded ? (int)fs.sf[c * 128 + b] : 0;
                    band_off[c][k] = 2 - 5 * sfv / 2;
                    band_nz4[c][k] = sfv ? 4 : 0;
                    if (!sfv) { cnt_a[c][k] = 0; cnt_b[c][k] = 0; }     // also after the search below gave a band up
                    cl |= (cnt_a[c][k] >> 28) != 0;
                }
            }
            any_clamped = __any_sync(kFull, cl);
            hdr_total = 48;
            for (int c = 0; c < nch; c++) hdr_total += fs.header_bits[c];
        }
    };
    auto band_bits = [&](int c, int k, int noise, auto with_clamped) -> int {
        const int pos = min(max(noise + band_off[c][k], 0), 58);
        const uint32_t row = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(tb.rank_row) + (uint32_t)tb.curve[pos] * (uint32_t)band_nz4[c][k]);
        const uint32_t w = (row & 0x80u) ? cnt_b[c][k] : cnt_a[c][k];
        int bits = (int)((row >> 8) & 0xFFu) - (int)(__funnelshift_r(w, 0u, row) & 15u);       // shift = row & 31
        if (decltype(with_clamped)::value) bits -= (int)(row >> 16) * (int)(cnt_a[c][k] >> 28);
        return bits;
    };
    auto used_bits_counted = [&](int noise) -> int {
        int len = 0;
        if (any_clamped) {
#pragma unroll
            for (int c = 0; c < 2; c++)
#pragma unroll
                for (int k = 0; k < 4; k++) len += band_bits(c, k, noise, std::true_type{});
        } else {
#pragma unroll
            for (int c = 0; c < 2; c++)
#pragma unroll
                for (int k = 0; k < 4; k++) len += band_bits(c, k, noise, std::false_type{});
        }
        return warp_sum(len) + hdr_total;
    };
    load_bands();
    {
        int highest = (int)S.base_bands + (int)S.stereo_bands - 1;
        for (;;) {
            int lo_l = 0, hi_l = 255, mid_value = 0;          // BinarySearchLevel
            while (lo_l != hi_l) {
                const int mid = (lo_l + hi_l) / 2;
                if constexpr (COUNTED) mid_value = used_bits_counted(mid);
                else mid_value = used_bits(tb, fs, S, lane, mid, 0);
                if (mid_value > avail) lo_l = mid + 1; else hi_l = mid;
            }
            noise_level = (lo_l == 255 && mid_value > avail) ? -1 : lo_l;
            if (noise_level >= 0) break;
            highest -= 2;
            if (highest < 0) { failed = true; break; }
            if (lane == 0)
                for (int c = 0; c < nch; c++) { fs.sf[c * 128 + highest + 1] = 0; fs.sf[c * 128 + highest + 2] = 0; }
            __syncwarp();
            header_lengths(fs, S, lane);
            load_bands();
        }
    }
    CONVOY(4);
    if (!failed && noise_level != 0) {                        // BinarySearchBoundary
        int* pre = reinterpret_cast<int*>(fs.pcm);            // scratch shared with the frame buffer, which is filled later
        if constexpr (COUNTED) {                              // boundary_table() from the counts
            int diff[4] = {0, 0, 0, 0}, tot = 0;
#pragma unroll
            for (int c = 0; c < 2; c++)
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int c_hi = band_bits(c, k, noise_level, std::true_type{}), c_lo = band_bits(c, k, noise_level - 1, std::true_type{});
                    tot += c_hi;
                    diff[k] += c_lo - c_hi;
                }
            int base = warp_sum(tot) + hdr_total;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                int total;
                const int excl = warp_excl_scan(diff[k], lane, &total);
                pre[lane + 32 * k] = base + excl;
                base += total;
            }
            if (lane == 0) pre[128] = base;
            __syncwarp();
        } else
        boundary_table(tb, fs, S, lane, noise_level, pre);
        int lo_b = 0, hi_b = 127;
        while (abs(hi_b - lo_b) > 1) {
            const int mid = (lo_b + hi_b) / 2;
            const int v = pre[mid];
            if (avail < v) hi_b = mid - 1; else lo_b = mid;
        }
        if (lo_b == hi_b) boundary = lo_b < 127 ? lo_b : -1;
        else boundary = pre[hi_b] > avail ? lo_b : hi_b;
        if (boundary < 0) failed = true;
    }
    if (failed && lane == 0 && !surplus) a.status[stream] = ERR_HCA_ENCODE;   // EncodeFrame gives up (HcaErrorCode, hca.cpp:2976-2984);
    CONVOY(5);                                                                 //  the warp stays with its CTA, nothing of the frame is stored

    // ---- final resolutions (hca.cpp:2868-2876)
    if constexpr (COUNTED) {                                  // from the band state the search kept in registers
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int b = lane + 32 * k;
                const int pos = min(max((b < boundary ? noise_level - 1 : noise_level) + band_off[c][k], 0), 58);
                if (c < nch) fs.res[c * 128 + b] = band_nz4[c][k] ? tb.curve[pos] : (uint8_t)0;
            }
    } else {
        for (int c = 0; c < nch; c++) {
            const int coded = S.coded[c];
            for (int b = lane; b < 128; b += 32)
                fs.res[c * 128 + b] = b < coded ? (uint8_t)enc_resolution(tb, fs.sf[c * 128 + b], b < boundary ? noise_level - 1 : noise_level) : 0;
        }
    }
    __syncwarp();

    for (uint32_t i = lane; i < a.frame_words; i += 32) fs.bits[i] = 0;   // the PCM stage / boundary table are dead now
    __syncwarp();
    // ---- pack (hca.cpp:2894-2963): sync word, noise level, boundary, per-channel headers, spectra, CRC
    const int limit_bits = (frame_size - 2) * 8 + 16;         // writer buffer = frame_size - 2 bytes after the sync word
    // One piece (at most 24 bits) into the frame buffer, without a branch on whether it straddles a word: the code is
    // placed in a 64-bit window that starts at its first word, the second word is ORed only if something landed in it.
    auto put_bits = [&](uint32_t code, int len, int at) {                 // the reference's writer drops what does not fit (IO.cpp:131-134)
        if (len > 0 && at + len <= limit_bits) {
            const int w = at >> 5, bo = at & 31;
            const unsigned long long window = (unsigned long long)code << (64 - bo - len);     // bo + len <= 55
            const uint32_t second = (uint32_t)window;
            atomicOr(&fs.bits[w], (uint32_t)(window >> 32));
            if (second) atomicOr(&fs.bits[w + 1], second);
        }
    };
    // Header fields of known length need no prefix sum: sync word + noise level + boundary are word 0, the 3-bit delta
    // width, raw 6-bit scalefactors, intensities and HFR scales sit at computed offsets. Only delta-coded scalefactors have
    // data-dependent lengths: a lane takes bands 4 l .. 4 l + 3 (two pieces of at most 22 bits) and ONE warp prefix sum
    // per channel places them.
    if (lane == 0) fs.bits[0] = 0xFFFF0000u | ((uint32_t)noise_level << 7) | (uint32_t)boundary;
    __syncwarp();
    int cursor = 32;
    for (int c = 0; c < nch; c++) {
        const int coded = S.coded[c];
        const int db = fs.delta_bits[c];
        const uint8_t* sf = fs.sf + c * 128;
        if (lane == 0) put_bits((uint32_t)db, 3, cursor);
        cursor += 3;
        const int b4 = 4 * lane, nb = min(max(coded - b4, 0), 4);         // this lane's bands
        if (db == 6) {
            uint32_t code = 0;
            for (int h = 0; h < nb; h++) code = (code << 6) | sf[b4 + h];
            put_bits(code, 6 * nb, cursor + 24 * lane);
            cursor += 6 * coded;
        } else if (db != 0) {
            const int maxd = (1 << (db - 1)) - 1, esc = (1 << db) - 1;
            uint32_t piece[2] = {0, 0};
            int plen[2] = {0, 0};
#pragma unroll
            for (int h = 0; h < 4; h++) {
                const int b = b4 + h;
                uint32_t code = 0;
                int len = 0;
                if (b == 0 && coded > 0) { code = sf[0]; len = 6; }
                else if (b < coded) {
                    const int delta = (int)sf[b] - (int)sf[b - 1];
                    if (abs(delta) > maxd) { code = ((uint32_t)esc << 6) | sf[b]; len = db + 6; }
                    else { code = (uint32_t)(maxd + delta); len = db; }
                }
                piece[h >> 1] = (piece[h >> 1] << len) | code;
                plen[h >> 1] += len;
            }
            int total;
            const int at = cursor + warp_excl_scan(plen[0] + plen[1], lane, &total);
            put_bits(piece[0], plen[0], at);
            put_bits(piece[1], plen[1], at + plen[0]);
            cursor += total;
        }
        if (S.type[c] == 2) {
            if (lane < 8) put_bits(fs.inten[c * 8 + lane], 4, cursor + 4 * lane);
            cursor += 32;
        } else if (S.hfr_groups > 0) {
            if (lane < S.hfr_groups) put_bits((uint32_t)fs.hfr_scale[c * 8 + lane], 6, cursor + 6 * lane);
            cursor += 6 * (int)S.hfr_groups;
        }
    }
    CONVOY(6);
    // Spectra in two phases. (1) Every lane quantises its bands (4 l .. 4 l + 3 of every channel) for all eight subframes --
    // the band's constants are fetched once, not once per subframe -- and leaves (length << 16 | code) in place of the
    // scaled value (QuantizeSpectra + WriteSpectra, hca.cpp:2878-2936). (2) In bitstream order (subframe-major, channel-
    // minor) a lane picks its four words up with one 16-byte load, joins them into two codes of at most 24 bits, and one
    // warp prefix sum of the lengths places them: 16 prefix sums per stereo frame.
    for (int c = 0; c < nch; c++) {
        const int coded = S.coded[c];
        int mb[4], down[4], tbase[4];
        uint32_t mmask[4];
        bool prefix[4], looked_up[4];
        float inv[4], up[4];
#pragma unroll
        for (int h = 0; h < 4; h++) {
            const int b = 4 * lane + h;
            const int r = b < coded ? (int)fs.res[c * 128 + b] : 0;
            inv[h] = tb.inv_step[r];
            up[h] = __fadd_rn(inv[h], 1.0f);
            mb[h] = (int)tb.max_bits[r] - 1;
            down[h] = r < 8 ? r + 1 : (1 << mb[h]);                        // (int)(inv + 0.5)
            prefix[h] = r < 8;
            looked_up[h] = r >= 1 && r < 8;                                // resolution 0: entry 8 of row 0 (all zero), whatever the band holds
            tbase[h] = looked_up[h] ? r * 16 + 8 : 8;
            mmask[h] = (1u << (mb[h] & 31)) - 1u;
        }
        for (int sub = 0; sub < 8; sub++) {
            uint4* row = reinterpret_cast<uint4*>(fs.spec + ((size_t)c * 8 + sub) * kSpecRow + 4 * lane);
            const uint4 v = *row;
            const uint32_t x[4] = {v.x, v.y, v.z, v.w};
            uint32_t o[4];
#pragma unroll
            for (int h = 0; h < 4; h++) {
                const int q = __float2int_rz(__fadd_rn(__fmul_rn(__uint_as_float(x[h]), inv[h]), up[h])) - down[h];
                // prefix codebooks (r <= 7): length << 16 | code from the table; sign-magnitude (r >= 8): |q| then the sign
                // bit, and a zero gives the sign bit back (|q| <= 2^mb - 1, so neither form needs a mask)
                const uint32_t from_table = tb.qpack[tbase[h] + (looked_up[h] ? q : 0)];
                const uint32_t mag = (uint32_t)abs(q) & mmask[h];
                const uint32_t computed = ((uint32_t)(mb[h] + (q != 0 ? 1 : 0)) << 16) | (mag << 1) | ((uint32_t)q >> 31);
                o[h] = prefix[h] ? from_table : computed;
            }
            *row = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
    CONVOY(7);
    // rows in bitstream order (subframe-major, channel-minor), two per prefix sum: the lengths of a lane's codes in rows
    // i and i + 1 travel through the scan as two 16-bit fields (a row holds at most 128 x 12 bits)
    int row_sub = 0, row_c = 0;                               // (subframe, channel) of the next row: no division in the loop
    for (int i = 0; i < 8 * nch; i += 2) {
        uint32_t code[2][2];
        int len[2][2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int sub = row_sub, c = row_c;
            if (++row_c == nch) { row_c = 0; row_sub++; }
            const uint4 e = *reinterpret_cast<const uint4*>(fs.spec + ((size_t)c * 8 + sub) * kSpecRow + 4 * lane);
            const int l1 = (int)(e.y >> 16), l3 = (int)(e.w >> 16);
            code[h][0] = ((e.x & 0xFFFFu) << l1) | (e.y & 0xFFFFu); code[h][1] = ((e.z & 0xFFFFu) << l3) | (e.w & 0xFFFFu);
            len[h][0] = (int)(e.x >> 16) + l1; len[h][1] = (int)(e.z >> 16) + l3;
        }
        int total;
        const int excl = warp_excl_scan((len[0][0] + len[0][1]) | ((len[1][0] + len[1][1]) << 16), lane, &total);
        const int at0 = cursor + (excl & 0xFFFF), at1 = cursor + (total & 0xFFFF) + (excl >> 16);
        cursor += (total & 0xFFFF) + (total >> 16);
        put_bits(code[0][0], len[0][0], at0);
        put_bits(code[0][1], len[0][1], at0 + len[0][0]);
        put_bits(code[1][0], len[1][0], at1);
        put_bits(code[1][1], len[1][1], at1 + len[1][0]);
    }
    CONVOY(8);

    // ---- CRC16 over the first frame_size - 2 bytes. The CRC (init 0, no final xor) is linear: every lane takes a
    // contiguous chunk of whole words (four bytes per step: slice-by-4 tables), multiplies its partial CRC by
    // x^(8 * bytes behind the chunk) mod P (host-made per-stream constants) and the 32 products are XORed together.
    uint8_t* dst = a.out + S.out_off + (uint64_t)frame * frame_size;
    uint32_t crc = 0;
    {
        const int body = frame_size - 2, chunk = ((((body + 3) >> 2) + 31) >> 5) << 2;     // bytes per lane, a multiple of 4
        const int b0 = min(lane * chunk, body), b1 = min(b0 + chunk, body);
        uint32_t part = 0;
        int i = b0;
        for (; i + 4 <= b1; i += 4) {                        // fs.bits words hold the frame's bytes first-byte-on-top
            const uint32_t x = fs.bits[i >> 2] ^ (part << 16);
            part = (uint32_t)tb.crc[3][x >> 24] ^ (uint32_t)tb.crc[2][(x >> 16) & 0xFF] ^ (uint32_t)tb.crc[1][(x >> 8) & 0xFF] ^ (uint32_t)tb.crc[0][x & 0xFF];
        }
        for (; i < b1; i++) {
            const uint32_t byte = (fs.bits[i >> 2] >> (24 - 8 * (i & 3))) & 0xFF;
            part = ((part << 8) & 0xFFFF) ^ (uint32_t)tb.crc[0][((part >> 8) ^ byte) & 0xFF];
        }
        uint32_t prod = 0;                                   // carry-less part * mul mod x^16 + x^15 + x^2 + 1
#pragma unroll
        for (int bit = 15; bit >= 0; bit--) {
            prod = ((prod << 1) ^ ((prod & 0x8000u) ? 0x18005u : 0u)) & 0xFFFFu;
            if ((part >> bit) & 1u) prod ^= mul;
        }
        crc = prod;
#pragma unroll
        for (int o = 16; o; o >>= 1) crc ^= __shfl_xor_sync(kFull, crc, o);
    }
    // ---- the frame leaves as 32-bit words wherever it lands in the blob: bytes up to the first aligned address and
    // behind the last whole word go one by one, word k in between is the frame's bytes nh + 4k .. nh + 4k + 3, i.e.
    // the packed words k and k + 1 funnel-shifted by nh bytes and turned into memory order. The CRC takes the place of
    // the last two bytes first.
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int i = frame_size - 2 + k, sh = 24 - 8 * (i & 3);
            fs.bits[i >> 2] = (fs.bits[i >> 2] & ~(0xFFu << sh)) | (((k == 0 ? crc >> 8 : crc) & 0xFFu) << sh);
        }
    }
    __syncwarp();
    if (!failed && !surplus) {
        auto byte_at = [&](int i) -> uint8_t { return (uint8_t)(fs.bits[i >> 2] >> (24 - 8 * (i & 3))); };
        const int nh = min((int)((4 - (reinterpret_cast<uintptr_t>(dst) & 3)) & 3), frame_size);
        const int nw = (frame_size - nh) >> 2, nt = frame_size - nh - 4 * nw;
        if (lane < nh) dst[lane] = byte_at(lane);
        if (lane >= 8 && lane < 8 + nt) dst[nh + 4 * nw + lane - 8] = byte_at(nh + 4 * nw + lane - 8);
        uint32_t* dw = reinterpret_cast<uint32_t*>(dst + nh);
        for (int k = lane; k < nw; k += 32)
            dw[k] = __byte_perm(__funnelshift_l(fs.bits[k + 1], fs.bits[k], 8 * nh), 0, 0x0123);
    }
    __syncwarp();                                             // the next round reuses the warp's shared memory
    }
}

}  // namespace

size_t hca_encode_smem_per_warp(uint32_t max_channels, uint32_t frame_words) {
    size_t n = (size_t)max_channels * 8 * kSpecRow * 4 + (size_t)max_channels * 8 * 4 * 2 + (size_t)max_channels * 4 * 2 +
               (std::max((size_t)129 * 4 + 12, (size_t)frame_words * 4) + 15) / 16 * 16 + (size_t)max_channels * 128 * 2 +
               (size_t)max_channels * 8;
    return (n + 15) / 16 * 16;
}

int launch_hca_encode(HcaEncodeArgs a, cudaStream_t s, uint64_t* launches) {
    if (!a.n_frames) return 0;
    a.smem_per_warp = (uint32_t)hca_encode_smem_per_warp(a.max_channels, a.frame_words);
    // wide streams (up to 8 channels x 6.6 KB) get fewer warps per CTA so the CTA still fits in shared memory
    int warps = kEncWarps;
    while (warps > 1 && (size_t)a.smem_per_warp * warps > 200 * 1024) warps >>= 1;
    const size_t smem = (size_t)a.smem_per_warp * warps;
    if (smem > 200 * 1024) return -1;
    auto kernel = a.max_channels <= 2 ? hca_encode_kernel<true> : hca_encode_kernel<false>;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // One CTA per 2 x `warps` frames: every warp makes two rounds over the frame list. More rounds per CTA amortise the
    // table set-up further, but the warps of a CTA drift apart from round to round and stop sharing instruction fetches (the
    // kernel is ~100 KB of SASS). Measured per 8192 streams: 1 / 2 / 3 / 4 / 8 rounds -> 12.83 / 12.42 / 14.55 / 15.23 /
    // 16.30 ms, resident CTAs only (HCA_ENC_GRID = 2) 16.1 ms; with a CTA barrier at the top of every round
    // (HCA_ENC_ROUND_SYNC) 2 / 4 / 8 rounds / resident -> 12.78 / 12.87 / 13.10 / 13.72 ms; resident CTAs whose warps are
    // deliberately spread over a frame time (HCA_ENC_STAGGER_NS) 17.1 ms -- twenty phases per SM are the worst case.
    const uint64_t want = (a.n_frames + warps - 1) / warps;
#ifndef HCA_ENC_ROUNDS
#define HCA_ENC_ROUNDS 2
#endif
    uint64_t grid = (want + HCA_ENC_ROUNDS - 1) / HCA_ENC_ROUNDS;
#ifdef HCA_ENC_GRID
    if (HCA_ENC_GRID == 2) {
        int dev = 0, sm_count = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        grid = std::min<uint64_t>(want, 2ull * (uint64_t)std::max(sm_count, 1));
    }
#endif
    kernel<<<(unsigned)grid, warps * 32, smem, s>>>(a);
    ++*launches;
    return 0;
}

}  // namespace cri
