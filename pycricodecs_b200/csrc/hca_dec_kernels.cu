// HCA decode kernels for sm_100a: bitstream unpack and IMDCT transform.
//
// Reference pipeline per frame (CriCodecs/hca.cpp): sync + CRC16 + cipher LUT,
// frame header, per-channel scalefactors / intensity / HFR scales, resolution
// and gain per band, 8 x channels runs of variable-length codes
// (clHCA_DecodeBlock_unpack, :1149-1205), then per subframe HFR reconstruction,
// intensity stereo and a 128-point DCT-IV with window + overlap-add
// (clHCA_DecodeBlock_transform, :1207-1233), then float -> PCM16 (:339-360).
//
// Two kernels, split where the parallelism changes shape:
//
//  hca_unpack_kernel     one LANE per frame. The code lengths are data dependent,
//      so a frame's 2048 codes are one serial chain; the batch supplies the
//      parallelism (770 k frames at the headline config). Each lane streams its
//      frame with 16-byte loads, keeps a 64-bit bit window in registers and writes
//      quantised coefficients (int16) and per-band gains (fp32) to an intermediate
//      laid out [step][channel][...][lane], i.e. every warp store is one
//      contiguous 512-byte row.
//
//  hca_imdct_kernel      one THREAD per (run of frames, channel), the whole
//      128-point transform in registers (hca_dct_gen.inc, generated): the stage
//      shuffles of the network are register renamings, so the instruction stream
//      is the transform's ~4 k separately rounded fp32 operations (no FMA: the
//      reference build has none, and PCM parity is bit-exact) plus dequantisation
//      and PCM conversion. The overlap state (first half of the previous DCT
//      output) is carried in registers along the run. PCM leaves through a
//      per-warp shared-memory tile so that global stores are coalesced per stream.
//
// HBM traffic per stereo frame: frame_size + 4096 B compulsory, plus the
// intermediate (4 KB int16 + 1 KB gains written and read once).
#include <cstdint>

#include "cri_tables.h"
#include "hca_kernels.h"

namespace cri {
namespace {

__constant__ uint8_t c_invert[66] = CRI_TBL_INVERT;
__constant__ uint32_t c_scaling[64] = CRI_TBL_DEC_SCALING;
__constant__ uint32_t c_range[16] = CRI_TBL_DEC_RANGE;
__constant__ uint32_t c_conv[128] = CRI_TBL_SCALE_CONV;
__constant__ uint32_t c_intensity[16] = CRI_TBL_INTENSITY_RATIO;
__constant__ uint8_t c_read_bits[128] = CRI_TBL_READ_BITS;
__constant__ int8_t c_read_vals[128] = CRI_TBL_READ_VALS;

#include "hca_dct_gen.inc"

// ------------------------------------------------------------------ unpack
constexpr int kUnpackThreads = 128;

struct UnpackTables {           // small per-CTA copies: per-lane indices diverge, shared memory does not serialise
    uint8_t invert[68];
    uint8_t code[128];          // (value + 8) | bits << 4 for resolutions 1..7
    float scaling[64];
    float range[16];
    float conv[128];
};

// CRC-16 (poly 0x8005, MSB first) of one more byte without a table: the reference's
// table entry is (v<<1) ^ (v<<2) ^ (parity(v) ? 0x8003 : 0)  (checked in tests/test_tables.py).
__device__ __forceinline__ uint32_t crc16_step(uint32_t crc, uint32_t byte) {
    const uint32_t v = ((crc >> 8) ^ byte) & 0xFF;
    const uint32_t t = (v << 1) ^ (v << 2) ^ ((__popc(v) & 1) ? 0x8003u : 0u);
    return ((crc << 8) ^ t) & 0xFFFF;
}

struct FrameReader {
    const uint4* src;           // 16-byte aligned load cursor
    uint4 cur;
    int sub;                    // next 32-bit lane of `cur`
    int byte_pos;               // frame-relative index of the next byte that fetch() returns
    int frame_size;
    uint32_t crc;
    const uint8_t* cipher;      // nullptr = identity
    uint64_t win;               // bit window, MSB first
    int have;                   // valid bits in win
    int pos;                    // frame-relative bit position of win's MSB
    int nbits;

    __device__ __forceinline__ uint32_t fetch_word() {  // next 4 stream bytes, big endian, CRC'd and deciphered
        if (sub == 4) { cur = __ldg(src++); sub = 0; }
        uint32_t w = sub == 0 ? cur.x : sub == 1 ? cur.y : sub == 2 ? cur.z : cur.w;
        sub++;
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t b = (w >> (8 * k)) & 0xFF;   // memory order = little-endian lanes
            const int p = byte_pos + k;
            if (p >= 0 && p < frame_size) crc = crc16_step(crc, b);
            if (cipher) b = __ldg(cipher + b);
            out = (out << 8) | b;
        }
        byte_pos += 4;
        return out;
    }

    __device__ __forceinline__ void init(const uint8_t* frame, int size, const uint8_t* table) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(frame);
        src = reinterpret_cast<const uint4*>(a & ~(uintptr_t)15);
        const int lead = (int)(a & 15);            // bytes of the first 16-byte row that precede the frame
        sub = 4;
        byte_pos = -lead;
        frame_size = size;
        crc = 0;
        cipher = table;
        nbits = size * 8;
        pos = 0;
        for (int k = 0; k < (lead >> 2); k++) fetch_word();
        win = (uint64_t)fetch_word() << 32;
        have = 32 - 8 * (lead & 3);
        win <<= 8 * (lead & 3);
        win |= (uint64_t)fetch_word() << (32 - have);
        have += 32;
        if (have <= 32) { win |= (uint64_t)fetch_word() << (32 - have); have += 32; }
    }

    __device__ __forceinline__ uint32_t peek(int n) const {  // n in 0..31; 0 once the read would cross the frame end
        const uint32_t v = ((uint32_t)(win >> 32) >> 1) >> (31 - n);
        return pos + n <= nbits ? v : 0u;
    }
    __device__ __forceinline__ void skip(int n) {            // n in 0..16; keeps more than 32 valid bits in the window
        win <<= n;
        have -= n;
        pos += n;
        if (have <= 32) {
            win |= (uint64_t)fetch_word() << (32 - have);
            have += 32;
        }
    }
    __device__ __forceinline__ uint32_t read(int n) { const uint32_t v = peek(n); skip(n); return v; }

    __device__ __forceinline__ uint32_t finish_crc() {       // run the CRC to the end of the frame
        while (byte_pos < frame_size) fetch_word();
        return crc;
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
    for (int i = threadIdx.x; i < 16; i += blockDim.x) tb.range[i] = __uint_as_float(c_range[i]);
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const uint64_t group = (uint64_t)blockIdx.x * (kUnpackThreads / 32) + (threadIdx.x >> 5);
    if (group >= a.total_groups) return;
    const uint32_t block = (uint32_t)(group / a.steps), step = (uint32_t)(group % a.steps);
    const HcaUnit u = a.units[block * 32 + lane];
    // step 0 is the look-back frame in front of the run (needed only for its last subframe)
    if (u.count == 0 || (step == 0 && u.first == 0) || (step > u.count)) return;
    const uint32_t frame = u.first + step - 1;
    const HcaStreamDev& S = a.streams[u.stream];
    const int nch = S.channels;
    const bool lookback = step == 0;

    // per-thread shared scratch: scalefactors of the channel being parsed + resolutions of every channel
    const int T = blockDim.x;
    uint8_t* s_sf = s_dyn;                                   // [128][T] bytes
    uint32_t* s_res = reinterpret_cast<uint32_t*>(s_dyn + 128 * T);  // [channel][16][T] words of 8 nibbles
    const int tid = threadIdx.x;

    FrameReader br;
    br.init(a.in + S.in_off + (uint64_t)frame * S.frame_size, (int)S.frame_size,
            S.cipher ? a.cipher + (size_t)S.cipher * 256 : nullptr);
    bool bad = br.read(16) != 0xFFFF;                        // sync word (never enciphered: table[0xFF] = 0xFF)

    const uint32_t noise_level = br.read(9), boundary = br.read(7);
    const uint32_t packed = (noise_level << 8) - boundary;
    const uint8_t* ath = a.ath + (size_t)S.ath * 128;
    const uint64_t slot = ((uint64_t)block * a.steps + step) * a.max_channels;

    for (int c = 0; c < nch && !bad; c++) {
        const int coded = S.coded[c];
        const int type = S.type[c];
        // ---- scalefactors (hca.cpp:1290-1358, v2.0 and older)
        const uint32_t delta_bits = br.read(3);
        if (delta_bits >= 6) {
            for (int i = 0; i < coded; i++) s_sf[i * T + tid] = (uint8_t)br.read(6);
        } else if (delta_bits > 0) {
            const uint32_t escape = (1u << delta_bits) - 1;
            uint32_t v = br.read(6);
            s_sf[tid] = (uint8_t)v;
            for (int i = 1; i < coded; i++) {
                const uint32_t d = br.read((int)delta_bits);
                if (d == escape) {
                    v = br.read(6);
                } else {
                    const int test = (int)v + ((int)d - (int)(escape >> 1));
                    if (test < 0 || test >= 64) { bad = true; break; }
                    v = (v - (escape >> 1) + d) & 0x3F;
                }
                s_sf[i * T + tid] = (uint8_t)v;
            }
        } else {
            for (int i = 0; i < 128; i++) s_sf[i * T + tid] = 0;
        }
        if (bad) break;
        // ---- intensity (secondary) or HFR scales (others), hca.cpp:1361-1441
        uint32_t inten = 0;
        if (type == 2) {
            const uint32_t v0 = br.peek(4);
            inten = v0;
            if (v0 < 15) {
                br.skip(4);
                for (int i = 1; i < 8; i++) inten |= br.read(4) << (4 * i);
            }
            a.inten[(slot + c) * 32 + lane] = inten;
        } else {
            for (int g = 0; g < S.hfr_groups; g++) s_sf[(128 - S.hfr_groups + g) * T + tid] = (uint8_t)br.read(6);
        }
        // ---- resolution + gain per band (hca.cpp:1444-1507)
        float4* gdst = a.gain + ((slot + c) * 32) * 32 + lane;   // [32 chunks of 4 bands][32 lanes]
        for (int i0 = 0; i0 < 128; i0 += 8) {
            uint32_t resw = 0;
            float g[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int i = i0 + k;
                uint32_t r = 0;
                g[k] = 0.f;
                if (i < coded) {
                    const uint32_t sf = s_sf[i * T + tid];
                    if (sf > 0) {
                        const int level = (int)ath[i] + (int)((packed + (uint32_t)i) >> 8);
                        const int cp = level + 1 - (int)((5 * sf) >> 1);
                        r = cp < 0 ? 15u : cp <= 65 ? (uint32_t)tb.invert[cp] : 0u;
                        if (r > S.max_res) r = S.max_res; else if (r < S.min_res) r = S.min_res;
                    }
                    g[k] = __fmul_rn(tb.scaling[sf], tb.range[r]);
                }
                resw |= r << (4 * k);
            }
            s_res[(c * 16 + (i0 >> 3)) * T + tid] = resw;
            if (i0 < coded) {
                gdst[(i0 >> 2) * 32] = make_float4(g[0], g[1], g[2], g[3]);
                gdst[((i0 >> 2) + 1) * 32] = make_float4(g[4], g[5], g[6], g[7]);
            }
        }
        // ---- HFR multipliers for the bands above the coded ones (hca.cpp:1638-1683, v2.0 rule)
        if (S.bands_per_hfr && type != 2) {
            const int start = S.base_bands + S.stereo_bands;
            int high = start, low = start - 1;
            float* gflat = reinterpret_cast<float*>(a.gain + ((slot + c) * 32) * 32);
            for (int g = 0; g < S.hfr_groups; g++)
                for (int i = 0; i < S.bands_per_hfr; i++) {
                    if (high >= S.total_bands || low < 0) break;
                    int k = (int)s_sf[(128 - S.hfr_groups + g) * T + tid] - (int)s_sf[low * T + tid] + 63;
                    k &= ~(k >> 31);
                    gflat[((high >> 2) * 32 + lane) * 4 + (high & 3)] = tb.conv[k];
                    high++; low--;
                }
        }
    }

    // ---- spectra: subframe-major, channel-minor runs of codes (hca.cpp:1540-1571)
    const uint64_t max_bits_lo = 0x6544443320ull;            // resolutions 0..9: 0,2,3,3,4,4,4,4,5,6 (4 bits each, r0 lowest)
    for (int sub = 0; sub < 8 && !bad; sub++) {
        for (int c = 0; c < nch; c++) {
            const int coded = S.coded[c];
            uint4* qdst = a.quant + (((slot + c) * 8 + sub) * 16) * 32 + lane;
            const bool keep = !lookback || sub == 7;
            for (int i0 = 0; i0 < coded; i0 += 8) {
                const uint32_t resw = s_res[(c * 16 + (i0 >> 3)) * T + tid];
                int q[8];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    q[k] = 0;
                    if (i0 + k < coded) {
                        const uint32_t r = (resw >> (4 * k)) & 15;
                        const int bits = r < 10 ? (int)((max_bits_lo >> (4 * r)) & 15) : (int)r - 3;
                        const uint32_t code = br.peek(bits);
                        int used;
                        if (r > 7) {
                            const int mag = (int)(code >> 1);
                            q[k] = (code & 1) ? -mag : mag;
                            used = bits - (mag == 0);
                        } else {
                            const uint32_t e = tb.code[(r << 4) + code];
                            q[k] = (int)(e & 15) - 8;
                            used = (int)(e >> 4);
                        }
                        br.skip(used);
                    }
                }
                if (keep) {
                    uint4 v;
                    v.x = (uint32_t)(q[0] & 0xFFFF) | ((uint32_t)q[1] << 16);
                    v.y = (uint32_t)(q[2] & 0xFFFF) | ((uint32_t)q[3] << 16);
                    v.z = (uint32_t)(q[4] & 0xFFFF) | ((uint32_t)q[5] << 16);
                    v.w = (uint32_t)(q[6] & 0xFFFF) | ((uint32_t)q[7] << 16);
                    qdst[(i0 >> 3) * 32] = v;
                }
            }
        }
    }
    if (!bad) bad = br.finish_crc() != 0;                    // a valid frame's CRC over all bytes is 0 (hca.cpp:1166)
    if (bad) a.status[u.stream] = ERR_HCA_DECODE;
}

// --------------------------------------------------------------- transform
constexpr int kImdctThreads = 64;
constexpr int kTileRow = 130;   // int16 per lane: 128 samples + 2 pad -> 65-word rows, conflict-free

__device__ __forceinline__ int pcm16(float f) {              // hca.cpp:339-360; (int) of an out-of-range float is INT_MIN on x86
    const float v = __fmul_rn(f, 32768.0f);
    int s = __float2int_rz(v);
    if (!(fabsf(v) < 2147483648.0f)) s = INT_MIN;
    return max(-32768, min(32767, s));
}

__global__ void __launch_bounds__(kImdctThreads)
hca_imdct_kernel(HcaDecodeArgs a) {
    __shared__ __align__(16) int16_t s_tile[kImdctThreads / 32][32][kTileRow];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t li = blockIdx.x * kImdctThreads + threadIdx.x;   // lanes[] is padded to the grid
    const HcaLane me = a.lanes[li];
    const bool idle = me.unit == 0xFFFFFFFFu;
    HcaUnit u{0, 0, 0};
    if (!idle) u = a.units[me.unit];
    const HcaStreamDev& S = a.streams[u.stream];
    const int ch = (int)me.channel;
    const int nch = idle ? 1 : S.channels;
    const int coded = idle ? 0 : S.coded[ch];
    const int type = idle ? 0 : S.type[ch];
    const uint32_t block = idle ? 0 : me.unit >> 5, ulane = idle ? 0 : me.unit & 31;
    uint32_t steps_here = idle ? 0 : u.count + 1;
    uint32_t warp_steps = steps_here;
#pragma unroll
    for (int o = 16; o; o >>= 1) warp_steps = max(warp_steps, __shfl_xor_sync(0xFFFFFFFFu, warp_steps, o));
    const bool warp_joint = __any_sync(0xFFFFFFFFu, !idle && S.joint);

    float x[128];
    float dprev[64];
#pragma unroll
    for (int i = 0; i < 64; i++) dprev[i] = 0.f;

    for (uint32_t step = 0; step < warp_steps; step++) {
        const bool live = step < steps_here && !(step == 0 && u.first == 0);
        const uint32_t frame = u.first + step - 1;
        const uint64_t slot = ((uint64_t)block * a.steps + step) * a.max_channels + ch;
        uint32_t inten = 0;
        if (warp_joint && live && type == 2) inten = a.inten[slot * 32 + ulane];
        for (int sub = (step == 0 ? 7 : 0); sub < 8; sub++) {
            if (live) {
                // ---- dequantise: spectra = gain * q  (hca.cpp:1568)
                const uint4* qsrc = a.quant + ((slot * 8 + sub) * 16) * 32 + ulane;
                const float4* gsrc = a.gain + (slot * 32) * 32 + ulane;
#pragma unroll
                for (int c8 = 0; c8 < 16; c8++) {
                    if (c8 * 8 < coded) {
                        const uint4 q = qsrc[c8 * 32];
                        const float4 g0 = gsrc[(2 * c8) * 32], g1 = gsrc[(2 * c8 + 1) * 32];
                        const int q0 = (int)(short)(q.x & 0xFFFF), q1 = (int)q.x >> 16, q2 = (int)(short)(q.y & 0xFFFF), q3 = (int)q.y >> 16;
                        const int q4 = (int)(short)(q.z & 0xFFFF), q5 = (int)q.z >> 16, q6 = (int)(short)(q.w & 0xFFFF), q7 = (int)q.w >> 16;
                        x[c8 * 8 + 0] = c8 * 8 + 0 < coded ? __fmul_rn(g0.x, (float)q0) : 0.f;
                        x[c8 * 8 + 1] = c8 * 8 + 1 < coded ? __fmul_rn(g0.y, (float)q1) : 0.f;
                        x[c8 * 8 + 2] = c8 * 8 + 2 < coded ? __fmul_rn(g0.z, (float)q2) : 0.f;
                        x[c8 * 8 + 3] = c8 * 8 + 3 < coded ? __fmul_rn(g0.w, (float)q3) : 0.f;
                        x[c8 * 8 + 4] = c8 * 8 + 4 < coded ? __fmul_rn(g1.x, (float)q4) : 0.f;
                        x[c8 * 8 + 5] = c8 * 8 + 5 < coded ? __fmul_rn(g1.y, (float)q5) : 0.f;
                        x[c8 * 8 + 6] = c8 * 8 + 6 < coded ? __fmul_rn(g1.z, (float)q6) : 0.f;
                        x[c8 * 8 + 7] = c8 * 8 + 7 < coded ? __fmul_rn(g1.w, (float)q7) : 0.f;
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; k++) x[c8 * 8 + k] = 0.f;
                    }
                }
            }
            if (warp_joint) {
                // ---- HFR: copy mirrored low bands upward (hca.cpp:1638-1683)
                if (live && S.bands_per_hfr && type != 2) {
                    float tmp[128];
#pragma unroll
                    for (int i = 0; i < 128; i++) tmp[i] = x[i];
                    const int start = S.base_bands + S.stereo_bands;
                    const int room = min(min((int)S.total_bands - start, (int)S.hfr_groups * (int)S.bands_per_hfr), start);
                    const float* gflat = reinterpret_cast<const float*>(a.gain + (slot * 32) * 32);
                    for (int n = 0; n < room; n++) {
                        const int high = start + n, low = start - 1 - n;
                        tmp[high] = __fmul_rn(gflat[((high >> 2) * 32 + ulane) * 4 + (high & 3)], tmp[low]);
                    }
                    const int last = start + max(room, 0) - 1;
                    if (last >= 0) tmp[last] = 0.f;
#pragma unroll
                    for (int i = 0; i < 128; i++) x[i] = tmp[i];
                }
                // ---- intensity stereo: the secondary channel is rebuilt from the primary (hca.cpp:1696-1714)
                const float rl = __uint_as_float(c_intensity[(__shfl_down_sync(0xFFFFFFFFu, inten, 1) >> (4 * sub)) & 15]);
                const float rr_self = __fsub_rn(2.0f, __uint_as_float(c_intensity[(inten >> (4 * sub)) & 15]));
                const int lo = S.base_bands, hi = S.total_bands;
#pragma unroll
                for (int i = 0; i < 128; i++) {
                    const float left = __shfl_up_sync(0xFFFFFFFFu, x[i], 1);
                    if (live && i >= lo && i < hi) {
                        if (type == 2) x[i] = __fmul_rn(left, rr_self);
                        else if (type == 1) x[i] = __fmul_rn(x[i], rl);
                    }
                }
            }
            if (live) hca_dct4_dec(x);
            if (step == 0) {                       // look-back subframe: only its DCT output is needed
                if (live) hca_imdct_carry(x, dprev);
                continue;
            }
            // ---- window + overlap, PCM16 into the warp's tile
            if (live) hca_imdct_window(x, dprev, [&](int i, float w) { s_tile[warp][lane][i] = (int16_t)pcm16(w); });
            __syncwarp();
            // ---- coalesced store: all lanes write consecutive samples of ONE (unit, channel) at a time
            const long long n0 = (long long)frame * 1024 + sub * 128 - (long long)S.delay;  // stream sample index of tile[0]
            long long base = 0;
            int i_lo = 0, i_hi = 0, stride = 0;
            if (live) {
                base = (long long)S.out_off + (n0 * nch + ch) * 2;
                stride = nch * 2;
                i_lo = (int)max(0ll, -n0);
                i_hi = (int)min(128ll, (long long)S.out_samples - n0);
            }
            for (int j = 0; j < 32; j++) {
                const long long b = __shfl_sync(0xFFFFFFFFu, base, j);
                const int st = __shfl_sync(0xFFFFFFFFu, stride, j);
                const int lo = __shfl_sync(0xFFFFFFFFu, i_lo, j), hi = __shfl_sync(0xFFFFFFFFu, i_hi, j);
                if (lo >= hi) continue;
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const int i = lane + 32 * m;
                    if (i >= lo && i < hi) *reinterpret_cast<int16_t*>(a.out + b + (long long)i * st) = s_tile[warp][j][i];
                }
            }
            __syncwarp();
        }
    }
}

}  // namespace

void launch_hca_decode(const HcaDecodeArgs& a, uint32_t n_lanes, cudaStream_t s, uint64_t* launches, cudaEvent_t mid) {
    if (!a.total_groups) return;
    const size_t per_thread = 128 + 64 * (size_t)a.max_channels;
    const size_t smem = per_thread * kUnpackThreads;
    cudaFuncSetAttribute(hca_unpack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const uint64_t groups_per_cta = kUnpackThreads / 32;
    hca_unpack_kernel<<<(unsigned)((a.total_groups + groups_per_cta - 1) / groups_per_cta), kUnpackThreads, smem, s>>>(a);
    ++*launches;
    if (mid) cudaEventRecord(mid, s);
    hca_imdct_kernel<<<n_lanes / kImdctThreads, kImdctThreads, 0, s>>>(a);
    ++*launches;
}

uint32_t hca_imdct_lane_granule() { return kImdctThreads; }

}  // namespace cri
