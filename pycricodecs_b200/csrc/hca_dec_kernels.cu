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
//  hca_unpack_kernel   one LANE per frame. Code lengths are data dependent, so a
//      frame's 2048 codes are one serial chain and the batch supplies the
//      parallelism (770 k frames at the headline config). Phase 1 walks the frame
//      once: CRC16 (table-free byte step), cipher LUT, byte swap, into an aligned
//      scratch row. Phase 2 parses from a 64-bit register window refilled with
//      prefetched 32-bit words; the per-code path is branch-free (both code
//      families are evaluated and selected) because lanes of a warp sit on
//      different resolutions and would otherwise serialise. Quantised
//      coefficients go through a 128-byte-per-lane shared tile so that HBM
//      sees full 128-byte rows.
//
//  hca_imdct_kernel    one WARP per run of frames of one stream. Coefficient p of
//      a 128-point block lives in lane p>>2, register p&3 and never moves: the
//      reference's 7 sum/difference passes pair slots differing in bit 0..6 of
//      p, its 7 rotation passes pair bit 6..0, the window pairs bit 0 (see
//      tools/gen_dct.py), so every exchange is a register swap or one
//      __shfl_xor, every global access is a coalesced 256/512-byte row, and the
//      overlap state is two registers per lane. All products and sums are
//      separately rounded (__fmul_rn/__fadd_rn): the reference build has no FMA,
//      and PCM parity is bit-exact.
//
// HBM traffic per stereo frame: frame_size + 4096 B compulsory, plus the
// intermediate (4 KB int16 spectra + 1 KB gains + the scratch row, written and
// read once; the scratch row normally stays in L2).
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
__constant__ uint8_t c_max_bits[16] = CRI_TBL_MAX_BITS;

#include "hca_dct_gen.inc"

// ------------------------------------------------------------------ unpack
constexpr int kUnpackThreads = 128;
constexpr int kStageRow = 9;    // uint4 per lane in the store tile: 8 payload + 1 pad (36-word rows: conflict-free)

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

struct BitWindow {              // MSB-first reader over big-endian 32-bit words
    uint64_t win;
    const uint32_t* next_ptr;
    uint32_t next, next2;       // two prefetched words: the scratch row comes back from L2, one word ahead is too late
    int have;                   // valid bits in win (kept above 32)
    int pos;                    // bits consumed so far
    int nbits;

    __device__ __forceinline__ void init(const uint32_t* words, int frame_bits) {
        win = ((uint64_t)words[0] << 32) | words[1];
        next = words[2];
        next2 = words[3];
        next_ptr = words + 4;
        have = 64; pos = 0; nbits = frame_bits;
    }
    // n in 0..16. A read that would cross the end of the frame returns 0 (hca.cpp:232-233).
    __device__ __forceinline__ uint32_t peek(int n) const {
        const uint32_t v = ((uint32_t)(win >> 32) >> 1) >> (31 - n);
        return pos + n <= nbits ? v : 0u;
    }
    __device__ __forceinline__ void skip(int n) {
        win <<= n;
        have -= n;
        pos += n;
        if (have <= 32) {       // predicated, no divergence: shift in the prefetched word, fetch the one after
            win |= (uint64_t)next << (32 - have);
            have += 32;
            next = next2;
            next2 = *next_ptr++;
        }
    }
    __device__ __forceinline__ uint32_t read(int n) { const uint32_t v = peek(n); skip(n); return v; }
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

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    const int T = kUnpackThreads;
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

    // shared scratch: [0,128T) scalefactors of the channel being parsed; then per channel 128 bytes of
    // (resolution | max_bits << 4) per band; then the store tile.
    uint8_t* s_sf = s_dyn;
    uint8_t* s_rb = s_dyn + 128 * T;
    uint4* s_stage = reinterpret_cast<uint4*>(s_dyn + (size_t)(128 + 128 * a.max_channels) * T) + (size_t)warp * 32 * kStageRow;

    bool bad = false;
    uint32_t* words = a.scratch + slot * a.scratch_words;    // this frame's aligned, deciphered, byte-swapped copy
    const int frame_size = active ? (int)S.frame_size : 0;

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
            // the 5 aligned words that cover this output row, then a byte funnel
            uint32_t v[5];
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
    br.init(words, frame_size * 8);
    uint32_t packed = 0;
    if (active) {
        if (br.read(16) != 0xFFFF) bad = true;               // sync word (cipher tables keep 0xFF fixed)
        const uint32_t noise_level = br.read(9), boundary = br.read(7);
        packed = (noise_level << 8) - boundary;
    }
    const uint8_t* ath = a.ath + (size_t)(active ? S.ath : 0) * 128;
    for (int c = 0; c < nch && !bad; c++) {
        const int coded = S.coded[c];
        const int type = S.type[c];
        // scalefactors (hca.cpp:1290-1358, v2.0 and older)
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
        // intensity (secondary channel) or HFR scales (others), hca.cpp:1361-1441
        if (type == 2) {
            const uint32_t v0 = br.peek(4);
            uint32_t inten = v0;
            if (v0 < 15) {
                br.skip(4);
                for (int i = 1; i < 8; i++) inten |= br.read(4) << (4 * i);
            }
            a.inten[slot * a.max_channels + c] = inten;
        } else {
            for (int g = 0; g < S.hfr_groups; g++) s_sf[(128 - S.hfr_groups + g) * T + tid] = (uint8_t)br.read(6);
        }
        // resolution + gain per band (hca.cpp:1444-1507)
        float4* gdst = reinterpret_cast<float4*>(a.gain + (slot * a.max_channels + c) * 128);
        for (int i0 = 0; i0 < 128; i0 += 4) {
            float g[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
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
                s_rb[(c * 128 + i) * T + tid] = (uint8_t)(r | ((uint32_t)tb.max_bits[r] << 4));
            }
            gdst[i0 >> 2] = make_float4(g[0], g[1], g[2], g[3]);
        }
        // HFR multipliers for the bands above the coded ones (hca.cpp:1638-1683, v2.0 rule)
        if (S.bands_per_hfr && type != 2) {
            const int start = S.base_bands + S.stereo_bands;
            int high = start, low = start - 1;
            float* gflat = a.gain + (slot * a.max_channels + c) * 128;
            for (int g = 0; g < S.hfr_groups; g++)
                for (int i = 0; i < S.bands_per_hfr; i++) {
                    if (high >= S.total_bands || low < 0) break;
                    int k = (int)s_sf[(128 - S.hfr_groups + g) * T + tid] - (int)s_sf[low * T + tid] + 63;
                    k &= ~(k >> 31);
                    gflat[high] = tb.conv[k];
                    high++; low--;
                }
        }
    }

    // ---- spectra: subframe-major, channel-minor runs of codes (hca.cpp:1540-1571). The loops are warp-uniform
    // (bounded by the widest stream of the warp) because the store tile is flushed cooperatively.
    int warp_nch = nch;
#pragma unroll
    for (int o = 16; o; o >>= 1) warp_nch = max(warp_nch, __shfl_xor_sync(0xFFFFFFFFu, warp_nch, o));
    const bool lookback = step == 0;
    uint4* my_stage = s_stage + lane * kStageRow;
    for (int sub = 0; sub < 8; sub++) {
        for (int c = 0; c < warp_nch; c++) {
            const bool mine = c < nch && !bad;
            const int coded = mine ? (int)S.coded[c] : 0;
            for (int halfband = 0; halfband < 2; halfband++) {
#pragma unroll 1
                for (int i0 = halfband * 64; i0 < halfband * 64 + 64; i0 += 8) {
                    uint32_t pk[4] = {0, 0, 0, 0};
                    if (i0 < coded) {
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const int i = i0 + k;
                            const uint32_t rb = i < coded ? s_rb[(c * 128 + i) * T + tid] : 0u;
                            const uint32_t r = rb & 15;
                            const int bits = (int)(rb >> 4);
                            const uint32_t code = br.peek(bits);
                            // sign-magnitude family (resolution >= 8): LSB is the sign, zero gives one bit back
                            const int mag = (int)(code >> 1);
                            const int v_hi = (code & 1) ? -mag : mag;
                            const int used_hi = bits - (mag == 0);
                            // prefix-codebook family (resolution <= 7)
                            const uint32_t e = tb.code[((r & 7) << 4) | (code & 15)];
                            const int v_lo = (int)(e & 15) - 8;
                            const int used_lo = (int)(e >> 4);
                            const bool hi = r > 7;
                            const int v = hi ? v_hi : v_lo;
                            br.skip(hi ? used_hi : used_lo);
                            pk[k >> 1] |= ((uint32_t)v & 0xFFFFu) << (16 * (k & 1));
                        }
                    }
                    my_stage[(i0 >> 3) & 7] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
                __syncwarp();
                // flush: 64 coefficients (128 B) of each lane's frame as one full row; 4 frames per warp store
                if (!lookback || sub == 7) {
                    const uint64_t my_row = ((slot * a.max_channels + c) * 8 + sub) * 16 + halfband * 8;  // uint4 index
                    const bool my_ok = c < nch;
#pragma unroll
                    for (int it = 0; it < 8; it++) {
                        const int j = it * 4 + (lane >> 3);
                        const uint64_t row = __shfl_sync(0xFFFFFFFFu, my_row, j);
                        const bool ok = __shfl_sync(0xFFFFFFFFu, (int)my_ok, j);
                        if (ok) a.quant[row + (lane & 7)] = s_stage[j * kStageRow + (lane & 7)];
                    }
                }
                __syncwarp();
            }
        }
    }
    if (bad && active) a.status[u.stream] = ERR_HCA_DECODE;
}

// --------------------------------------------------------------- transform
constexpr int kImdctWarps = 4;

__device__ __forceinline__ int pcm16(float f) {   // hca.cpp:339-360; (int) of an out-of-range float is INT_MIN on x86
    const float v = __fmul_rn(f, 32768.0f);
    int s = __float2int_rz(v);
    if (!(fabsf(v) < 2147483648.0f)) s = INT_MIN;
    return max(-32768, min(32767, s));
}

__device__ __forceinline__ float flip(float v, uint32_t mask) { return __uint_as_float(__float_as_uint(v) ^ mask); }

struct Spectra4 {              // one lane's share of a 128-point block as it sits in memory
    uint2 q;                    // 4 x int16 quantised coefficients
    float4 g;                   // 4 gains (or HFR multipliers above the coded bands)
};

__global__ void __launch_bounds__(kImdctWarps * 32)
hca_imdct_kernel(HcaDecodeArgs a) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ __align__(16) float s_rot_s[7 * 128];
    __shared__ __align__(16) float s_rot_c[7 * 128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t unit = blockIdx.x * kImdctWarps + warp;
    const bool dead = unit >= a.n_units || a.units[unit < a.n_units ? unit : 0].count == 0;
    // rotation factors per slot (p = 4*lane + r): one shared copy per CTA, read as one LDS.128 per pass and table.
    // Keeping the 56 per-lane factors in registers instead costs half the occupancy.
    for (int i = threadIdx.x; i < 7 * 128; i += blockDim.x) {
        s_rot_s[i] = __uint_as_float(kRotS[i]);
        s_rot_c[i] = __uint_as_float(kRotC[i]);
    }
    __syncthreads();
    if (dead) return;
    const HcaUnit u = a.units[unit];
    const HcaStreamDev& S = a.streams[u.stream];
    const int nch = S.channels;
    const int MC = (int)a.max_channels;
    // per-warp shared: PCM tile [MC][128] int16, overlap carry [MC][32] float2, HFR scratch [128] float
    uint8_t* base = s_dyn + (size_t)warp * (MC * 512 + 512);
    int16_t* tile = reinterpret_cast<int16_t*>(base);
    float2* carry = reinterpret_cast<float2*>(base + MC * 256);
    float* xs = reinterpret_cast<float*>(base + MC * 512);

    const float wa0 = __uint_as_float(kWinA[2 * lane]), wa1 = __uint_as_float(kWinA[2 * lane + 1]);
    const float wb0 = __uint_as_float(kWinB[2 * lane]), wb1 = __uint_as_float(kWinB[2 * lane + 1]);
    const int pa0 = kWinPosA[2 * lane], pa1 = kWinPosA[2 * lane + 1], pb0 = kWinPosB[2 * lane], pb1 = kWinPosB[2 * lane + 1];
    uint32_t sgn[5];   // sum/difference passes across lanes: the slot with the pass bit set holds b of (a+b, a-b)
#pragma unroll
    for (int b = 0; b < 5; b++) sgn[b] = (lane >> b) & 1 ? 0x80000000u : 0u;
    const float4* rot_s = reinterpret_cast<const float4*>(s_rot_s) + lane;   // [pass * 32]
    const float4* rot_c = reinterpret_cast<const float4*>(s_rot_c) + lane;

    for (int c = 0; c < nch; c++) carry[c * 32 + lane] = make_float2(0.f, 0.f);
    const int total = S.total_bands, basebands = S.base_bands;
    const int start = S.base_bands + S.stereo_bands;
    const int room = min(min(total - start, (int)S.hfr_groups * (int)S.bands_per_hfr), start);
    const bool joint = S.joint;
    const uint64_t slot0 = (uint64_t)unit * a.steps;

    auto fetch = [&](uint32_t step, int sub, int c) {
        Spectra4 r;
        const uint64_t sc = (slot0 + step) * MC + c;
        r.q = reinterpret_cast<const uint2*>(a.quant + (sc * 8 + sub) * 16)[lane];
        r.g = reinterpret_cast<const float4*>(a.gain + sc * 128)[lane];
        return r;
    };

    const uint32_t first_step = u.first == 0 ? 1u : 0u;
    Spectra4 nxt = fetch(first_step, first_step == 0 ? 7 : 0, 0);
    for (uint32_t step = first_step; step <= u.count; step++) {
        const uint32_t frame = u.first + step - 1;
        const uint64_t slot = slot0 + step;
        for (int sub = (step == 0 ? 7 : 0); sub < 8; sub++) {
            float xl[4] = {0.f, 0.f, 0.f, 0.f};   // primary channel's spectra for intensity stereo
            for (int c = 0; c < nch; c++) {
                const int coded = S.coded[c];
                const int type = S.type[c];
                const Spectra4 cur = nxt;
                {   // software prefetch: the next block's row is requested before this block's arithmetic starts
                    int nc = c + 1, ns = sub;
                    uint32_t nstep = step;
                    if (nc == nch) { nc = 0; ns++; }
                    if (ns == 8) { ns = 0; nstep++; }
                    if (nstep <= u.count) nxt = fetch(nstep, ns, nc);
                }
                // ---- dequantise: spectra = gain * q  (hca.cpp:1568); bands past the coded count are zero
                float x[4];
                x[0] = 4 * lane + 0 < coded ? __fmul_rn(cur.g.x, (float)(int)(short)(cur.q.x & 0xFFFF)) : 0.f;
                x[1] = 4 * lane + 1 < coded ? __fmul_rn(cur.g.y, (float)((int)cur.q.x >> 16)) : 0.f;
                x[2] = 4 * lane + 2 < coded ? __fmul_rn(cur.g.z, (float)(int)(short)(cur.q.y & 0xFFFF)) : 0.f;
                x[3] = 4 * lane + 3 < coded ? __fmul_rn(cur.g.w, (float)((int)cur.q.y >> 16)) : 0.f;
                if (joint) {
                    // ---- HFR: mirrored low bands scaled into the high bands (hca.cpp:1638-1683)
                    if (S.bands_per_hfr && type != 2) {
                        __syncwarp();
                        reinterpret_cast<float4*>(xs)[lane] = make_float4(x[0], x[1], x[2], x[3]);
                        __syncwarp();
                        const float gg[4] = {cur.g.x, cur.g.y, cur.g.z, cur.g.w};
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            const int p = 4 * lane + r;
                            if (p >= start && p < start + room) x[r] = __fmul_rn(gg[r], xs[2 * start - 1 - p]);
                            if (p == start + room - 1) x[r] = 0.f;
                        }
                    }
                    // ---- intensity stereo: the secondary channel is rebuilt from the primary (hca.cpp:1696-1714)
                    if (type == 1) {
                        const uint32_t inten = a.inten[slot * MC + c + 1];
                        const float rl = __uint_as_float(c_intensity[(inten >> (4 * sub)) & 15]);
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            xl[r] = x[r];
                            const int p = 4 * lane + r;
                            if (p >= basebands && p < total) x[r] = __fmul_rn(x[r], rl);
                        }
                    } else if (type == 2) {
                        const uint32_t inten = a.inten[slot * MC + c];
                        const float rr = __fsub_rn(2.0f, __uint_as_float(c_intensity[(inten >> (4 * sub)) & 15]));
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            const int p = 4 * lane + r;
                            if (p >= basebands && p < total) x[r] = __fmul_rn(xl[r], rr);
                        }
                    }
                }
                // ---- 7 sum/difference passes: slot bit 0, 1 (registers), 2..6 (lanes)
                {
                    const float t0 = __fadd_rn(x[0], x[1]), t1 = __fsub_rn(x[0], x[1]), t2 = __fadd_rn(x[2], x[3]), t3 = __fsub_rn(x[2], x[3]);
                    x[0] = __fadd_rn(t0, t2); x[2] = __fsub_rn(t0, t2); x[1] = __fadd_rn(t1, t3); x[3] = __fsub_rn(t1, t3);
                }
#pragma unroll
                for (int b = 0; b < 5; b++) {
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const float other = __shfl_xor_sync(0xFFFFFFFFu, x[r], 1 << b);
                        x[r] = __fadd_rn(other, flip(x[r], sgn[b]));
                    }
                }
                // ---- 7 rotation passes: slot bit 6..2 (lanes), 1, 0 (registers):  v*S + partner*C
#pragma unroll
                for (int st = 0; st < 5; st++) {
                    const float4 s4 = rot_s[st * 32], c4 = rot_c[st * 32];
                    const float ss[4] = {s4.x, s4.y, s4.z, s4.w}, cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const float other = __shfl_xor_sync(0xFFFFFFFFu, x[r], 16 >> st);
                        x[r] = __fadd_rn(__fmul_rn(x[r], ss[r]), __fmul_rn(other, cc[r]));
                    }
                }
                {
                    const float4 s5 = rot_s[5 * 32], c5 = rot_c[5 * 32], s6 = rot_s[6 * 32], c6 = rot_c[6 * 32];
                    const float y0 = __fadd_rn(__fmul_rn(x[0], s5.x), __fmul_rn(x[2], c5.x));
                    const float y2 = __fadd_rn(__fmul_rn(x[2], s5.z), __fmul_rn(x[0], c5.z));
                    const float y1 = __fadd_rn(__fmul_rn(x[1], s5.y), __fmul_rn(x[3], c5.y));
                    const float y3 = __fadd_rn(__fmul_rn(x[3], s5.w), __fmul_rn(x[1], c5.w));
                    x[0] = __fadd_rn(__fmul_rn(y0, s6.x), __fmul_rn(y1, c6.x));
                    x[1] = __fadd_rn(__fmul_rn(y1, s6.y), __fmul_rn(y0, c6.y));
                    x[2] = __fadd_rn(__fmul_rn(y2, s6.z), __fmul_rn(y3, c6.z));
                    x[3] = __fadd_rn(__fmul_rn(y3, s6.w), __fmul_rn(y2, c6.w));
                }
                // ---- window + overlap (hca.cpp:1983-1992): odd slots hold dct[j>=64], even slots dct[127-j]
                const float2 prev = carry[c * 32 + lane];
                carry[c * 32 + lane] = make_float2(x[0], x[2]);
                if (step != 0) {
                    int16_t* t = tile + c * 128;
                    t[pa0] = (int16_t)pcm16(__fadd_rn(__fmul_rn(wa0, x[1]), __fmul_rn(wb0, prev.x)));
                    t[pb0] = (int16_t)pcm16(__fsub_rn(__fmul_rn(wb0, x[1]), __fmul_rn(wa0, prev.x)));
                    t[pa1] = (int16_t)pcm16(__fadd_rn(__fmul_rn(wa1, x[3]), __fmul_rn(wb1, prev.y)));
                    t[pb1] = (int16_t)pcm16(__fsub_rn(__fmul_rn(wb1, x[3]), __fmul_rn(wa1, prev.y)));
                }
            }
            if (step == 0) continue;
            __syncwarp();
            // ---- interleave the channels and store this subframe's samples (contiguous in the WAV image)
            const long long n0 = (long long)frame * 1024 + sub * 128 - (long long)S.delay;
            uint8_t* dst = a.out + S.out_off;
            if (nch == 2 && ((S.out_off & 3) == 0)) {
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const int i = lane + 32 * m;
                    const long long n = n0 + i;
                    if (n >= 0 && n < (long long)S.out_samples) {
                        const uint32_t w = (uint32_t)(uint16_t)tile[i] | ((uint32_t)(uint16_t)tile[128 + i] << 16);
                        *reinterpret_cast<uint32_t*>(dst + n * 4) = w;
                    }
                }
            } else {
                for (int e = lane; e < 128 * nch; e += 32) {
                    const int i = e / nch, c = e - i * nch;
                    const long long n = n0 + i;
                    if (n >= 0 && n < (long long)S.out_samples)
                        *reinterpret_cast<int16_t*>(dst + (n * nch + c) * 2) = tile[c * 128 + i];
                }
            }
            __syncwarp();
        }
    }
}

}  // namespace

size_t hca_unpack_smem(uint32_t max_channels) {
    return (size_t)(128 + 128 * max_channels) * kUnpackThreads + (size_t)(kUnpackThreads / 32) * 32 * kStageRow * sizeof(uint4);
}

void launch_hca_decode(const HcaDecodeArgs& a, cudaStream_t s, uint64_t* launches, cudaEvent_t mid) {
    if (!a.total_groups) return;
    const size_t smem = hca_unpack_smem(a.max_channels);
    cudaFuncSetAttribute(hca_unpack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const uint64_t groups_per_cta = kUnpackThreads / 32;
    hca_unpack_kernel<<<(unsigned)((a.total_groups + groups_per_cta - 1) / groups_per_cta), kUnpackThreads, smem, s>>>(a);
    ++*launches;
    if (mid) cudaEventRecord(mid, s);
    const size_t smem2 = (size_t)kImdctWarps * (a.max_channels * 512 + 512);
    cudaFuncSetAttribute(hca_imdct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    hca_imdct_kernel<<<(a.n_units + kImdctWarps - 1) / kImdctWarps, kImdctWarps * 32, smem2, s>>>(a);
    ++*launches;
}

}  // namespace cri
