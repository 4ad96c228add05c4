// HCA decode, first kernel: frame check + bitstream unpack (clHCA_DecodeBlock_unpack, CriCodecs/hca.cpp:1149-1205).
//
// Reference pipeline per frame: sync word, CRC16 over the whole frame (:1166), cipher LUT (:1169), noise level
// + evaluation boundary, per channel scalefactors (:1290-1358), intensity / HFR scales (:1361-1441), resolution
// and gain per band (:1444-1507), then 8 x channels runs of variable-length codes (:1540-1571).
//
// One LANE per frame. Code lengths are data dependent, so a frame's 2048 codes are one serial chain and the
// batch supplies the parallelism (770 k frames at the headline config).
//   Phase 1 walks the frame once with 16-byte loads: CRC16 (table-free byte step), cipher LUT, byte swap, into
//   an aligned scratch row in HBM/L2.
//   Phase 2 parses from a 128-bit register window that is topped up 64 bits at a time, one 8-byte load in flight
//   ahead of use (the row comes back from L2/HBM, not L1), and checked once per 4 codes. The per-code path is
//   branch-free: lanes of a warp sit on different resolutions, so both code families are evaluated and
//   selected; the end-of-frame rule of the reference's reader (a read that would cross the end yields 0,
//   hca.cpp:232-233) is decided once per 128-code run, with an exact per-code path for runs that could cross.
//   Quantised coefficients leave through a 128-byte-per-lane shared tile so that HBM sees full 128-byte rows.
// The second kernel (hca_imdct_kernels.cu) turns spectra into PCM.
#include <cstdint>
#include <type_traits>

#include "cri_tables.h"
#include "hca_kernels.h"

namespace cri {
namespace {

__constant__ uint8_t c_invert[66] = CRI_TBL_INVERT;
__constant__ uint32_t c_scaling[64] = CRI_TBL_DEC_SCALING;
__constant__ uint32_t c_range[16] = CRI_TBL_DEC_RANGE;
__constant__ uint32_t c_conv[128] = CRI_TBL_SCALE_CONV;
__constant__ uint8_t c_read_bits[128] = CRI_TBL_READ_BITS;
__constant__ int8_t c_read_vals[128] = CRI_TBL_READ_VALS;
__constant__ uint8_t c_max_bits[16] = CRI_TBL_MAX_BITS;

constexpr int kUnpackThreads = 128;
constexpr unsigned kFull = 0xFFFFFFFFu;

struct UnpackTables {           // per-CTA copies: per-lane indices diverge, shared memory does not serialise
    uint8_t invert[68];
    uint8_t code[128];          // resolutions 0..7: (value + 8) | bits << 4
    uint8_t max_bits[16];
    float scaling[64];
    float range[16];
    float conv[128];
};

// CRC-16 (poly 0x8005, MSB first) of one more byte without a table: the reference's
// table entry is (v<<1) ^ (v<<2) ^ (parity(v) ? 0x8003 : 0)  (tests/test_tables.py).
__device__ __forceinline__ uint32_t crc16_step(uint32_t crc, uint32_t byte) {
    const uint32_t v = ((crc >> 8) ^ byte) & 0xFF;
    const uint32_t t = (v << 1) ^ (v << 2) ^ ((__popc(v) & 1) ? 0x8003u : 0u);
    return ((crc << 8) ^ t) & 0xFFFF;
}

// MSB-first reader over big-endian 32-bit words: 128 bits in registers (w3 holds the next bits), 64 more
// prefetched. top_up() must run at least once per 48 consumed bits.
struct BitWindow {
    uint32_t w3, w2, w1, w0;
    uint2 ahead, ahead2;        // next 2 x 64 bits, already requested (the row comes back from L2 / HBM)
    const uint2* next_ptr;      // what to load after `ahead2`
    int have;                   // valid bits in the window
    int loaded;                 // bits taken from the row so far (window + consumed): position = loaded - have

    __device__ __forceinline__ void init(const uint32_t* row) {
        const uint4 a = *reinterpret_cast<const uint4*>(row);
        w3 = a.x; w2 = a.y; w1 = a.z; w0 = a.w;
        next_ptr = reinterpret_cast<const uint2*>(row + 4);
        ahead = *next_ptr++;
        ahead2 = *next_ptr++;
        have = 128; loaded = 128;
    }
    __device__ __forceinline__ int position() const { return loaded - have; }
    __device__ __forceinline__ uint32_t peek(int n) const { return __funnelshift_l(w3, 0u, n); }   // top n bits, n in 0..31
    __device__ __forceinline__ void skip(int n) {                                            // n in 0..31
        w3 = __funnelshift_l(w2, w3, n);
        w2 = __funnelshift_l(w1, w2, n);
        w1 = __funnelshift_l(w0, w1, n);
        w0 <<= n;
        have -= n;
    }
    __device__ __forceinline__ void top_up() {
        if (have <= 64) {       // w1:w0 hold no valid bits; append `ahead` behind the `have` valid bits of w3:w2
            const uint64_t n64 = ((uint64_t)ahead.x << 32) | ahead.y;
            const uint64_t hi = (((uint64_t)w3 << 32) | w2) | ((n64 >> 1) >> (have - 1));
            const uint64_t lo = n64 << (64 - have);
            w3 = (uint32_t)(hi >> 32); w2 = (uint32_t)hi; w1 = (uint32_t)(lo >> 32); w0 = (uint32_t)lo;
            have += 64; loaded += 64;
            ahead = ahead2;
            ahead2 = *next_ptr++;
        }
    }
    // generic read for the frame header (rare, not on the per-coefficient path)
    __device__ __forceinline__ uint32_t read(int n, int nbits) {
        const uint32_t v = position() + n <= nbits ? peek(n) : 0u;
        skip(n);
        top_up();
        return v;
    }
};

__global__ void __launch_bounds__(kUnpackThreads)
hca_unpack_kernel(HcaDecodeArgs a) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ UnpackTables tb;
    for (int i = threadIdx.x; i < 66; i += blockDim.x) tb.invert[i] = c_invert[i];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) {
        tb.code[i] = (uint8_t)((c_read_vals[i] + 8) | (c_read_bits[i] << 4));
        tb.conv[i] = __uint_as_float(c_conv[i]);
    }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) tb.scaling[i] = __uint_as_float(c_scaling[i]);
    for (int i = threadIdx.x; i < 16; i += blockDim.x) {
        tb.range[i] = __uint_as_float(c_range[i]);
        tb.max_bits[i] = c_max_bits[i];
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t group = (uint64_t)blockIdx.x * (kUnpackThreads / 32) + warp;
    if (group >= a.total_groups) return;                    // whole warp
    const uint32_t block = (uint32_t)(group / a.steps), step = (uint32_t)(group % a.steps);
    const uint32_t unit = block * 32 + lane;
    const HcaUnit u = a.units[unit];
    // step 0 is the look-back frame in front of the run (needed only for its last subframe)
    const bool active = u.count != 0 && !(step == 0 && u.first == 0) && step <= u.count;
    const uint32_t frame = u.first + step - 1;
    const HcaStreamDev& S = a.streams[u.stream];
    const int nch = active ? S.channels : 0;
    const uint64_t slot = (uint64_t)unit * a.steps + step;   // frame slot of the intermediate arrays

    // per-warp shared scratch: 4 KB that first holds the scalefactors of the channel being parsed ([band][lane]
    // bytes) and later the store tile ([lane][8 x 16 B], XOR-swizzled); then per channel 4 KB of
    // (resolution | max_bits << 4) per band ([band][lane] bytes)
    uint8_t* wbase = s_dyn + (size_t)warp * (4096 + 4096 * a.max_channels);
    uint8_t* s_sf = wbase + lane;                            // + band * 32
    uint8_t* s_rb = wbase + 4096 + lane;                     // + (channel * 128 + band) * 32
    uint4* s_stage = reinterpret_cast<uint4*>(wbase);

    bool bad = false;
    uint32_t* words = a.scratch + slot * a.scratch_words;    // this frame's aligned, deciphered, byte-swapped copy
    const int frame_size = active ? (int)S.frame_size : 0;
    const int nbits = frame_size * 8;

    // ---- phase 1: CRC over the raw frame, cipher LUT, byte swap -> scratch row (16-byte loads, one row ahead)
    if (active) {
        const uint8_t* src = a.in + S.in_off + (uint64_t)frame * S.frame_size;
        const uintptr_t addr = reinterpret_cast<uintptr_t>(src);
        const uint4* ap = reinterpret_cast<const uint4*>(addr & ~(uintptr_t)15);
        const int lead = (int)(addr & 15);                  // bytes of the first row that precede the frame
        const int wsel = lead >> 2, sh = (lead & 3) * 8;
        const uint8_t* cipher = S.cipher ? a.cipher + (size_t)S.cipher * 256 : nullptr;
        uint32_t crc = 0;
        const int nrows = (frame_size + 15) >> 4;
        uint4 cur = __ldg(ap), nxt = __ldg(ap + 1);
        for (int row = 0; row < nrows; row++) {
            const uint4 nn = __ldg(ap + row + 2);           // the input blob has 64 bytes of slack behind it
            uint32_t v[5];                                  // the 5 aligned words that cover this output row
            {
                const uint32_t t[8] = {cur.x, cur.y, cur.z, cur.w, nxt.x, nxt.y, nxt.z, nxt.w};
#pragma unroll
                for (int k = 0; k < 5; k++) v[k] = wsel == 0 ? t[k] : wsel == 1 ? t[k + 1] : wsel == 2 ? t[k + 2] : t[k + 3];
            }
            uint32_t o[4];
#pragma unroll
            for (int wq = 0; wq < 4; wq++) {
                const uint32_t raw = __funnelshift_r(v[wq], v[wq + 1], sh);  // 4 frame bytes, memory order
                uint32_t be = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    uint32_t bb = (raw >> (8 * k)) & 0xFF;
                    if (16 * row + 4 * wq + k < frame_size) crc = crc16_step(crc, bb); else bb = 0;
                    if (cipher) bb = __ldg(cipher + bb);
                    be = (be << 8) | bb;
                }
                o[wq] = be;
            }
            reinterpret_cast<uint4*>(words)[row] = make_uint4(o[0], o[1], o[2], o[3]);
            cur = nxt; nxt = nn;
        }
        // zero rows of slack so the prefetching reader never sees stale data
        reinterpret_cast<uint4*>(words)[nrows] = make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4*>(words)[nrows + 1] = make_uint4(0, 0, 0, 0);
        if (crc != 0) bad = true;                           // a valid frame's CRC over all its bytes is 0 (hca.cpp:1166)
    }

    // ---- phase 2: frame header
    BitWindow br;
    br.init(words);
    uint32_t packed = 0;
    if (active) {
        if (br.read(16, nbits) != 0xFFFF) bad = true;        // sync word (cipher tables keep 0xFF fixed)
        const uint32_t noise_level = br.read(9, nbits), boundary = br.read(7, nbits);
        packed = (noise_level << 8) - boundary;
    }
    const uint8_t* ath = a.ath + (size_t)(active ? S.ath : 0) * 128;
    int run_bits[kHcaMaxChannels];                           // worst-case bits of one 128-code run per channel
    uint32_t frame_draws = 0;                                // noise generator draws per subframe, all channels
    for (int c = 0; c < nch && !bad; c++) {
        const int coded = S.coded[c];
        const int type = S.type[c];
        // scalefactors (hca.cpp:1290-1358); v3.0 appends one per HFR group to every non-secondary channel
        const int extra = (S.v3 && type != 2) ? (int)S.hfr_groups : 0;
        const int n_sf = coded + extra;                      // <= 127 - extra (plan_hca_decode, v3_supported)
        const uint32_t delta_bits = br.read(3, nbits);
        if (delta_bits >= 6) {
            for (int i = 0; i < n_sf; i++) s_sf[i * 32] = (uint8_t)br.read(6, nbits);
        } else if (delta_bits > 0) {
            const uint32_t escape = (1u << delta_bits) - 1;
            uint32_t v = br.read(6, nbits);
            s_sf[0] = (uint8_t)v;
            for (int i = 1; i < n_sf; i++) {
                const uint32_t d = br.read((int)delta_bits, nbits);
                if (d == escape) {
                    v = br.read(6, nbits);
                } else {
                    const int test = (int)v + ((int)d - (int)(escape >> 1));
                    if (test < 0 || test >= 64) { bad = true; break; }
                    v = (v - (escape >> 1) + d) & 0x3F;
                }
                s_sf[i * 32] = (uint8_t)v;
            }
        } else {
            for (int i = 0; i < 128; i++) s_sf[i * 32] = 0;
        }
        if (bad) break;
        // v3.0: the HFR scales are the extra scalefactors, copied to the top of the table starting ONE entry past the
        // last one read (hca.cpp:1353-1355); that entry is never written in a supported stream, so it reads as zero
        if (extra > 0 && delta_bits > 0) {
            s_sf[127 * 32] = 0;
            for (int i = 1; i < extra; i++) s_sf[(127 - i) * 32] = s_sf[(n_sf - i) * 32];
        }
        // intensity (secondary channel) or HFR scales (others), hca.cpp:1361-1441
        if (type == 2) {
            const uint32_t v0 = br.position() + 4 <= nbits ? br.peek(4) : 0u;
            uint32_t inten = v0;
            uint32_t keep = 0;                               // first nibble that keeps the previous frame's value (0: none)
            if (!S.v3) {
                if (v0 < 15) {
                    br.skip(4);
                    br.top_up();
                    for (int i = 1; i < 8; i++) inten |= br.read(4, nbits) << (4 * i);
                } else {
                    keep = 1;                                // 15 is stored, not consumed, and nothing else is read (:1368-1372)
                }
            } else {                                         // hca.cpp:1382-1424
                br.skip(4);
                br.top_up();
                if (v0 < 15) {
                    const uint32_t db = br.read(2, nbits);
                    if (db == 3) {
                        for (int i = 1; i < 8; i++) inten |= br.read(4, nbits) << (4 * i);
                    } else {
                        const uint32_t escape = (2u << db) - 1;
                        uint32_t v = v0;
                        for (int i = 1; i < 8; i++) {
                            const uint32_t d = br.read((int)db + 1, nbits);
                            if (d == escape) {
                                v = br.read(4, nbits);
                            } else {
                                v = (v - (escape >> 1) + d) & 0xFF;
                                // out of range: unpack_intensity returns here (:1410-1412), its caller ignores that
                                // (:1185) and goes on with the bits that follow and the older intensities
                                if (v > 15) { keep = (uint32_t)i; break; }
                            }
                            inten |= v << (4 * i);
                        }
                    }
                } else {
                    inten = 0x77777777u;
                }
            }
            a.carry[slot * a.max_channels + c] = (uint8_t)keep;
            a.inten[slot * a.max_channels + c] = inten;
        } else if (!S.v3) {
            for (int g = 0; g < S.hfr_groups; g++) s_sf[(128 - S.hfr_groups + g) * 32] = (uint8_t)br.read(6, nbits);
        }
        if (bad) break;
        // resolution + gain per band (hca.cpp:1444-1507)
        float4* gdst = reinterpret_cast<float4*>(a.gain + (slot * a.max_channels + c) * 128);
        const bool noise = S.noise && a.sfres != nullptr;
        uint32_t* cls = reinterpret_cast<uint32_t*>(a.sfres + (noise ? (slot * a.max_channels + c) * 128 : 0));
        int n_noise = 0, n_valid = 0;
        int sum_bits = 0;
        for (int i0 = 0; i0 < 128; i0 += 4) {
            float g[4];
            uint32_t cls4 = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int i = i0 + k;
                uint32_t r = 0;
                g[k] = 0.f;
                if (i < coded) {
                    const uint32_t sf = s_sf[i * 32];
                    uint32_t tag = sf;
                    if (sf > 0) {
                        const int level = (int)ath[i] + (int)((packed + (uint32_t)i) >> 8);
                        const int cp = level + 1 - (int)((5 * sf) >> 1);
                        r = cp < 0 ? 15u : cp <= 65 ? (uint32_t)tb.invert[cp] : 0u;
                        if (r > S.max_res) r = S.max_res; else if (r < S.min_res) r = S.min_res;
                        // hca.cpp:1479-1487: bands with a scalefactor are "noise" (resolution 0) or "valid"
                        if (r < 1) { tag |= 0x40; n_noise++; } else { tag |= 0x80; n_valid++; }
                    }
                    cls4 |= tag << (8 * k);
                    g[k] = __fmul_rn(tb.scaling[sf], tb.range[r]);
                }
                const uint32_t mb = i < coded ? tb.max_bits[r] : 0u;
                sum_bits += (int)mb;
                s_rb[(c * 128 + i) * 32] = (uint8_t)(r | (mb << 4));     // bands past `coded`: r = 0, 0 bits
            }
            gdst[i0 >> 2] = make_float4(g[0], g[1], g[2], g[3]);
            if (noise) cls[i0 >> 2] = cls4;
        }
        run_bits[c] = sum_bits;
        if (noise) {
            // the generator runs only in channels that have both kinds of band (hca.cpp:1605-1606)
            const uint32_t d = (n_noise > 0 && n_valid > 0) ? (uint32_t)n_noise : 0u;
            a.draws[slot * a.max_channels + c] = d;
            frame_draws += d;
        }
        // HFR multipliers for the bands above the coded ones (hca.cpp:1638-1683, v2.0 rule)
        if (S.bands_per_hfr && type != 2) {
            const int start = S.base_bands + S.stereo_bands;
            int high = start, low = start - 1;
            float* gflat = a.gain + (slot * a.max_channels + c) * 128;
            // v3.0: the low band stops moving down after the first half of the groups (hca.cpp:1652-1661)
            const int moving = S.v3 ? (int)S.hfr_groups >> 1 : (int)S.hfr_groups;
            for (int g = 0; g < S.hfr_groups; g++)
                for (int i = 0; i < S.bands_per_hfr; i++) {
                    if (high >= S.total_bands || low < 0) break;
                    int k = (int)s_sf[(128 - S.hfr_groups + g) * 32] - (int)s_sf[low * 32] + 63;
                    k &= ~(k >> 31);
                    gflat[high] = tb.conv[k];
                    high++;
                    if (g < moving) low--;
                }
        }
    }

    if (active && S.noise && a.sfres != nullptr) {
        if (step != 0) a.frame_draws[S.frame_base + frame] = bad ? 0u : frame_draws;
        if (bad)                                             // a failed frame draws nothing (its stream is reported as failed)
            for (int c = 0; c < nch; c++) a.draws[slot * a.max_channels + c] = 0;
    }

    // ---- spectra: subframe-major, channel-minor runs of codes (hca.cpp:1540-1571). The loops are warp-uniform
    // (bounded by the widest stream of the warp) because the store tile is flushed cooperatively.
    int warp_nch = nch;
#pragma unroll
    for (int o = 16; o; o >>= 1) warp_nch = max(warp_nch, __shfl_xor_sync(kFull, warp_nch, o));
    __syncwarp();                                            // the scalefactor bytes are dead: their 4 KB become the store tile
    const bool lookback = step == 0;
    const int sw = lane & 7;                                 // swizzle of this lane's tile row
    for (int sub = 0; sub < 8; sub++) {
        for (int c = 0; c < warp_nch; c++) {
            const bool mine = c < nch && !bad;
            const int coded = mine ? (int)S.coded[c] : 0;
            // can this run cross the end of the frame? (only corrupt / wrongly keyed frames do)
            const bool careful = mine && br.position() + run_bits[c < nch ? c : 0] > nbits;
            const uint8_t* rbp = s_rb + (size_t)c * 128 * 32;
            const bool any_careful = __any_sync(kFull, careful);
            for (int halfband = 0; halfband < 2; halfband++) {
                auto decode_half = [&](auto careful_tag) {
                    constexpr bool kCareful = decltype(careful_tag)::value;
#pragma unroll 1
                    for (int i0 = halfband * 64; i0 < halfband * 64 + 64; i0 += 8) {
                        uint32_t pk[4] = {0, 0, 0, 0};
                        if (i0 < coded) {
#pragma unroll
                            for (int k = 0; k < 8; k++) {
                                const uint32_t rb = rbp[(i0 + k) * 32];
                                const int bits = (int)(rb >> 4);
                                uint32_t code = br.peek(bits);
                                if (kCareful && br.position() + bits > nbits) code = 0;
                                // sign-magnitude family (resolution >= 8): LSB is the sign, zero gives one bit back
                                const int mag = (int)(code >> 1);
                                const int v_hi = (code & 1) ? -mag : mag;
                                const int used_hi = code < 2 ? bits - 1 : bits;
                                // prefix-codebook family (resolution <= 7): one table byte = (value + 8) | bits << 4
                                const uint32_t e = tb.code[((rb & 7) << 4) | (code & 15)];
                                const int v_lo = (int)(e & 15) - 8;
                                const int used_lo = (int)(e >> 4);
                                const bool hi = rb & 8;
                                const int v = hi ? v_hi : v_lo;
                                br.skip(hi ? used_hi : used_lo);
                                if ((k & 3) == 3) br.top_up();
                                pk[k >> 1] |= ((uint32_t)v & 0xFFFFu) << (16 * (k & 1));
                            }
                        }
                        s_stage[lane * 8 + (((i0 >> 3) & 7) ^ sw)] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                };
                if (any_careful) decode_half(std::true_type{}); else decode_half(std::false_type{});
                __syncwarp();
                // flush: 64 coefficients (128 B) of each lane's frame as one full row; 4 frames per warp store
                if (!lookback || sub == 7) {
                    const uint64_t my_row = ((slot * a.max_channels + c) * 8 + sub) * 16 + halfband * 8;  // uint4 index
                    const bool my_ok = c < nch;
#pragma unroll
                    for (int it = 0; it < 8; it++) {
                        const int j = it * 4 + (lane >> 3);
                        const uint64_t row = __shfl_sync(kFull, my_row, j);
                        const bool ok = __shfl_sync(kFull, (int)my_ok, j);
                        if (ok) a.quant[row + (lane & 7)] = s_stage[j * 8 + ((lane & 7) ^ (j & 7))];
                    }
                }
                __syncwarp();
            }
        }
    }
    if (bad && active) a.status[u.stream] = ERR_HCA_DECODE;
}

}  // namespace

size_t hca_unpack_smem(uint32_t max_channels) { return (size_t)(kUnpackThreads / 32) * (4096 + 4096 * (size_t)max_channels); }

void launch_hca_decode(const HcaDecodeArgs& a, cudaStream_t s, uint64_t* launches, cudaEvent_t mid) {
    if (!a.total_groups) return;
    const size_t smem = hca_unpack_smem(a.max_channels);
    cudaFuncSetAttribute(hca_unpack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const uint64_t groups_per_cta = kUnpackThreads / 32;
    hca_unpack_kernel<<<(unsigned)((a.total_groups + groups_per_cta - 1) / groups_per_cta), kUnpackThreads, smem, s>>>(a);
    ++*launches;
    if (a.sfres) launch_hca_noise_scan(a, s, launches);
    if (a.carry_scan) launch_hca_intensity_scan(a, s, launches);
    if (mid) cudaEventRecord(mid, s);
    launch_hca_imdct(a, s, launches);
}

}  // namespace cri
