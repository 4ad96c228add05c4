// HCA decode, fast path: every stream of the job is mono or stereo (discrete channels or one intensity-stereo pair,
// with or without high-frequency reconstruction), v <= 2.0 -- everything the reference encoder emits.
//
// Work is laid out over the FLATTENED frame list of the job (frame g = 0 .. G-1 in stream order, dec_prefix[s] =
// first g of stream s) cut into runs of `run_len` consecutive frames:
//
//   hca_unpack_fast_kernel   one LANE per frame (clHCA_DecodeBlock_unpack, CriCodecs/hca.cpp:1149-1205, plus the
//                            dequantisation spectra = gain * q, :1568). A frame's 2048 variable-length codes are
//                            one serial chain, so the batch supplies the parallelism. A warp takes the j-th frame
//                            of 32 consecutive runs, which makes every store of four dequantised coefficients a
//                            coalesced 256/512-byte row of the intermediate array.
//   hca_imdct_fast_kernel    one LANE per (run, channel) column: the whole 128-point DCT-IV lives in 128 registers
//                            as straight-line code with the rotation factors as immediates (tools/gen_dct.py; the
//                            reference network imdct_transform, hca.cpp:1898-1992, operand order and rounding points
//                            unchanged, no FMA), so a transform is its 3968 fp32 operations plus I/O: no shuffles,
//                            no shared-memory tables. The overlap state (dct[0..63] of the previous subframe) is a
//                            256-byte column of shared memory; PCM leaves through a per-warp shared tile as
//                            coalesced rows. A column starts by transforming the last subframe of the frame in front
//                            of its run (read from the neighbouring run's slot), which is all the look-back the
//                            overlap needs (hca.cpp:1990-1991).
//
// Intermediate array (HBM): spec[imdct warp W][frame in run j][subframe][chunk 0..31][lane 0..31] float4, where lane =
// channel * (32 / NCH) + run % (32 / NCH) and chunk i holds coefficients 4i .. 4i+3. Both kernels touch it with
// fully coalesced 512-byte rows. Algorithmic bytes per stereo frame: 8192 written + 8192 read, 4096 PCM written.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <type_traits>

#include "cri_tables.h"
#include "hca_kernels.h"

namespace cri {
namespace {

__constant__ uint8_t c_invert[66] = CRI_TBL_INVERT;
__constant__ uint32_t c_scaling[64] = CRI_TBL_DEC_SCALING;
__constant__ uint32_t c_range[16] = CRI_TBL_DEC_RANGE;
__constant__ int8_t c_read_vals[128] = CRI_TBL_READ_VALS;
__constant__ uint8_t c_max_bits[16] = CRI_TBL_MAX_BITS;
__constant__ uint32_t c_conv[128] = CRI_TBL_SCALE_CONV;          // HFR: scale_conversion_table (hca.cpp:1579-1598)
__constant__ uint32_t c_intensity[16] = CRI_TBL_INTENSITY_RATIO;  // intensity stereo ratios (hca.cpp:1689-1693)

// Two-wide fp32 helpers of the generated transform (sm_100a f32x2: one instruction, two separately rounded results).
// hca_bfly2: (a, b) <- (a + b, a - b) on two register pairs. hca_sum2: d = p + q, written as fma(q, one, p) with `one`
// = {1.0f, 1.0f} read from the kernel arguments: ptxas contracts a packed add of products into FFMA2 even under
// -fmad=false, which would drop the rounding of the products; it cannot contract through a multiplier it does not know.
#ifndef HCA_SCALAR_SUMS
#define HCA_SCALAR_SUMS 0          // experiments: 1 = every sum a scalar instruction (tools/build_variant.sh)
#endif
__device__ __forceinline__ void hca_bfly2(float& a0, float& a1, float& b0, float& b1) {
    if (HCA_SCALAR_SUMS) {
        const float s0 = __fadd_rn(a0, b0), s1 = __fadd_rn(a1, b1), d0 = __fsub_rn(a0, b0), d1 = __fsub_rn(a1, b1);
        a0 = s0; a1 = s1; b0 = d0; b1 = d1;
        return;
    }
    unsigned long long a, b, s, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(s) : "l"(a), "l"(b));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(s));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(b0), "=f"(b1) : "l"(d));
}
__device__ __forceinline__ void hca_sum2(unsigned long long one, float p0, float p1, float q0, float q1, float& d0, float& d1) {
    if (HCA_SCALAR_SUMS) { d0 = __fadd_rn(p0, q0); d1 = __fadd_rn(p1, q1); return; }
    unsigned long long p, q, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(p0), "f"(p1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(q0), "f"(q1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(q), "l"(one), "l"(p));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

#include "hca_dct_thread_gen.inc"
#ifdef CRI_HCA_PAIR_KERNEL        // experiment ("measured and dropped"): tools/build_variant.sh pair -DCRI_HCA_PAIR_KERNEL
#include "hca_dct_pair_gen.inc"
#endif

constexpr int kFastThreads = 128;             // unpack kernel
constexpr int kFastWarps = kFastThreads / 32;
constexpr unsigned kFull = 0xFFFFFFFFu;

// stream of flattened frame g: the s with prefix[s] <= g < prefix[s + 1] (streams without frames are skipped)
__device__ __forceinline__ uint32_t find_stream(const uint32_t* __restrict__ prefix, uint32_t n, uint32_t g) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------------ unpack
struct FastTables {
    float gain[1024];       // [scalefactor << 4 | resolution] = scaling[sf] * range[res]   (calculate_gain, hca.cpp:1498-1507)
    float code[128];        // resolutions 0..7, [4 peeked bits << 3 | res]: the value (read_val_table) as a float
    uint16_t crc[4][256];   // CRC-16 (poly 0x8005, MSB first) of byte v followed by k zero bytes
    uint8_t invert[68];
    uint8_t max_bits[16];
};

__device__ __forceinline__ uint32_t crc16_step(uint32_t crc, uint32_t byte) {   // poly 0x8005, MSB first, table-free
    const uint32_t v = ((crc >> 8) ^ byte) & 0xFF;
    const uint32_t t = (v << 1) ^ (v << 2) ^ ((__popc(v) & 1) ? 0x8003u : 0u);
    return ((crc << 8) ^ t) & 0xFFFF;
}

// CRC-16 over four more message bytes (w = bytes in memory order): the register is folded into the first two bytes,
// then the four table terms are independent (slice-by-4).
__device__ __forceinline__ uint32_t crc16_word(const uint16_t (&T)[4][256], uint32_t c, uint32_t w) {
    const uint32_t x = w ^ __byte_perm(c, 0, 0x4401);
    return (uint32_t)T[3][x & 0xFF] ^ (uint32_t)T[2][(x >> 8) & 0xFF] ^ (uint32_t)T[1][(x >> 16) & 0xFF] ^ (uint32_t)T[0][x >> 24];
}
__device__ __forceinline__ uint32_t crc16_byte(const uint16_t (&T0)[256], uint32_t c, uint32_t byte) {
    return ((c << 8) & 0xFFFF) ^ (uint32_t)T0[((c >> 8) ^ byte) & 0xFF];
}

// MSB-first reader over big-endian 32-bit words: 128 bits in registers (w3 holds the next bits). Refills come from a
// per-lane ring of 2 x 16 bytes in shared memory that cp.async keeps filled from the frame's scratch row: a refill is
// then a shared-memory load, and the copy that replaces a consumed slot is in flight for at least three top_up()
// calls before its data is needed (the other slot holds 128 bits; a call consumes at most 64), which is what
// `wait_group 2` in front of the refill checks. (A register prefetch does not work here: refills are per-lane
// events, but a register written by one lane's load is a scoreboard dependency for the whole warp, so the warp
// waited on L2 in nearly every iteration.) The reader runs past the frame by up to 48 bytes: the scratch row has
// slack. top_up() must run at least once per 48 consumed bits.
struct BitWindow {
    uint32_t w3, w2, w1, w0;
    uint32_t ring;              // shared-memory byte address of this lane's slot 0 (slot 1 is kRingSlotStride above)
    const uint4* next16;        // next 16-byte chunk of the row to request
    int rd;                     // next 8-byte half to consume: slot rd >> 1, half rd & 1
    int have;
    int loaded;
    static constexpr uint32_t kRingSlotStride = 32 * 16;

    __device__ __forceinline__ void request(uint32_t slot) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + slot * kRingSlotStride), "l"(next16));
        next16++;
    }
    __device__ __forceinline__ void init(const uint32_t* row, uint32_t ring_addr) {
        const uint4 a = *reinterpret_cast<const uint4*>(row);
        w3 = a.x; w2 = a.y; w1 = a.z; w0 = a.w;
        ring = ring_addr;
        next16 = reinterpret_cast<const uint4*>(row) + 1;
        request(0);
        request(1);
        asm volatile("cp.async.commit_group;");
        rd = 0;
        have = 128; loaded = 128;
    }
    __device__ __forceinline__ int position() const { return loaded - have; }
    __device__ __forceinline__ uint32_t peek(int n) const { return __funnelshift_l(w3, 0u, n); }   // n in 0..31
    __device__ __forceinline__ void skip(int n) {                                            // n in 0..31
        w3 = __funnelshift_l(w2, w3, n);
        w2 = __funnelshift_l(w1, w2, n);
        w1 = __funnelshift_l(w0, w1, n);
        w0 <<= n;
        have -= n;
    }
    // Branch-free: in nearly every call some lane of the warp needs its refill, so the warp would walk the refill
    // path anyway; predicated, it needs no reconvergence barrier and its loads and shifts interleave with the
    // value lookups of the codes around it.
    __device__ __forceinline__ void top_up() {
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        const bool need = have <= 64;
        const uint32_t slot = (uint32_t)rd >> 1;
        uint32_t vx = 0, vy = 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p ld.shared.v2.u32 {%0, %1}, [%2];\n\t}"
                     : "+r"(vx), "+r"(vy) : "r"(ring + slot * kRingSlotStride + ((uint32_t)rd & 1) * 8), "r"((uint32_t)need));
        const int h = need ? have : 64;                               // keeps the shift counts in range when nothing is loaded
        const uint64_t n64 = ((uint64_t)vx << 32) | vy;
        const uint64_t hi = (((uint64_t)w3 << 32) | w2) | ((n64 >> 1) >> (h - 1));
        const uint64_t lo = n64 << (64 - h);
        w3 = need ? (uint32_t)(hi >> 32) : w3;
        w2 = need ? (uint32_t)hi : w2;
        w1 = need ? (uint32_t)(lo >> 32) : w1;
        w0 = need ? (uint32_t)lo : w0;
        have += need ? 64 : 0;
        loaded += need ? 64 : 0;
        const bool refill = need && (rd & 1);                         // both halves of the slot are consumed: refill it
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                     ::"r"(ring + slot * kRingSlotStride), "l"(next16), "r"((uint32_t)refill) : "memory");
        next16 += refill ? 1 : 0;
        rd = need ? (rd + 1) & 3 : rd;
        asm volatile("cp.async.commit_group;");
    }
    __device__ __forceinline__ uint32_t read(int n, int nbits) {          // header fields (not the per-coefficient path)
        const uint32_t v = position() + n <= nbits ? peek(n) : 0u;
        skip(n);
        top_up();
        return v;
    }
};

// Per band and lane one 16-bit word: [5:2] resolution, [11:6] scalefactor, [15:12] code length (max_bit_table).
// word & 0xFFC is the byte offset of the band's gain in FastTables::gain, word & 0x3C = 4 * resolution.
//
// Code lengths: a code of resolution r occupies max_bits[r] bits, or one bit less when its peeked value is below
// kShortBelow[r] -- for the sign-magnitude family (r >= 8) that is the "zero gives the sign bit back" rule (code < 2),
// for the prefix codebooks (r <= 7) it restates read_bit_table (hca.cpp:1513-1526; tests/test_tables.py checks the
// restatement). So the bit position, the only serial dependency between codes, advances by a compare and a subtract;
// the value lookups hang off the side of the chain.
constexpr uint32_t kShortBelowLo = 0x26AE2620u;   // nibble r = threshold of resolution r, r = 0..7: 0,2,6,2,14,10,6,2
constexpr uint32_t kShortBelowHi = 0x22222222u;   // r = 8..15: 2
// JOINT: some stream of the batch has an intensity-stereo pair or HFR bands (their header fields and the HFR pass are
// compiled out otherwise)
template <int NCH, bool JOINT>
__global__ void __launch_bounds__(kFastThreads, 3)
hca_unpack_fast_kernel(HcaDecodeArgs a) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ FastTables tb;
    constexpr int RW = 32 / NCH;                              // runs per transform warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t R = a.run_len;
    const uint32_t wid = blockIdx.x * kFastWarps + warp;     // (block of 32 runs, frame in run)
    const uint32_t rb = wid / R, j = wid - rb * R;
    const uint32_t r = a.run_base + rb * 32 + lane;
    const uint64_t g64 = (uint64_t)r * R + j;
    const bool active = r < a.n_runs && r - a.run_base < a.run_count && g64 < a.total_frames;
    const uint32_t g = active ? (uint32_t)g64 : (uint32_t)a.total_frames;   // scratch row G is a dummy for idle lanes
    uint32_t stream = 0, frame = 0;
    if (active) {
        stream = find_stream(a.dec_prefix, a.n_streams, g);
        frame = g - __ldg(a.dec_prefix + stream);
    }
    if (active) {   // pull the frame towards L2 while the tables are built: each lane reads its own 682-byte frame
        const HcaStreamDev& S0 = a.streams[stream];
        const uint8_t* src = a.in + S0.in_off + (uint64_t)frame * S0.frame_size;
        for (uint32_t k = 0; k < S0.frame_size + 127; k += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + k));
    }
    for (int i = threadIdx.x; i < 1024; i += kFastThreads)
        tb.gain[i] = __fmul_rn(__uint_as_float(c_scaling[i >> 4]), __uint_as_float(c_range[i & 15]));
    for (int i = threadIdx.x; i < 128; i += kFastThreads)
        tb.code[((i & 15) << 3) | (i >> 4)] = (float)(int)c_read_vals[i];
    for (int i = threadIdx.x; i < 66; i += kFastThreads) tb.invert[i] = c_invert[i];
    for (int i = threadIdx.x; i < 16; i += kFastThreads) tb.max_bits[i] = c_max_bits[i];
    for (int i = threadIdx.x; i < 256; i += kFastThreads) {
        uint32_t c = crc16_step(0, (uint32_t)i);
#pragma unroll
        for (int k = 0; k < 4; k++) { tb.crc[k][i] = (uint16_t)c; c = crc16_step(c, 0); }
    }
    __syncthreads();

    if ((uint64_t)rb * 32 >= a.run_count || (uint64_t)a.run_base + rb * 32 >= a.n_runs) return;   // whole warp
    const HcaStreamDev& S = a.streams[stream];

    uint16_t* tab = reinterpret_cast<uint16_t*>(s_dyn) + (size_t)warp * (NCH * 128 * 32) + lane;   // + (c * 128 + band) * 32

    bool bad = false;
    uint32_t* words = a.scratch + (uint64_t)g * a.scratch_words;   // this frame's aligned, deciphered, big-endian copy
    const int frame_size = active ? (int)S.frame_size : 0;
    const int nbits = frame_size * 8;

    // ---- phase 1: CRC over the raw frame (four bytes per step, slice-by-4 tables), cipher LUT, byte swap -> scratch
    // row; 16-byte loads one row ahead
    if (active) {
        const uint8_t* src = a.in + S.in_off + (uint64_t)frame * S.frame_size;
        const uintptr_t addr = reinterpret_cast<uintptr_t>(src);
        const uint4* ap = reinterpret_cast<const uint4*>(addr & ~(uintptr_t)15);
        const int lead = (int)(addr & 15);
        const int wsel = lead >> 2, sh = (lead & 3) * 8;
        const uint8_t* cipher = S.cipher ? a.cipher + (size_t)S.cipher * 256 : nullptr;
        uint32_t crc = 0;
        const int nrows = (frame_size + 15) >> 4, full_rows = frame_size >> 4;
        // The frame's aligned 16-byte rows reach the lane through a ring of kSlots rows in shared memory (the warp's band
        // table region, not in use yet; [slot][lane] x 16 bytes, conflict-free), filled by cp.async in four groups: every
        // row of the ring is in flight at once, where loads into registers kept four rows -- and one L2 latency per
        // four rows -- on the lane's critical path (the phase was 7 % of the kernel's instructions and 21 % of its
        // stall samples).
        constexpr int kSlots = NCH * 16, kGroup = kSlots / 4;
        const uint32_t stage = (uint32_t)__cvta_generic_to_shared(s_dyn + (size_t)warp * (NCH * 128 * 32 * sizeof(uint16_t))) + lane * 16;
        int next_row = 1;                                        // aligned row 0 is loaded directly
        auto issue_group = [&]() {
#pragma unroll
            for (int k = 0; k < kGroup; k++) {
                const int ar = next_row + k;
                if (ar <= nrows)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stage + (uint32_t)((ar - 1) % kSlots) * 512u), "l"(ap + ar) : "memory");
            }
            next_row += kGroup;
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        uint4 cur = __ldg(ap);
#pragma unroll
        for (int k = 0; k < 4; k++) issue_group();
        auto fetch_row = [&](int row, uint32_t (&raw)[4]) {      // the row's four words, frame bytes in memory order
            if (row % kGroup == 0) asm volatile("cp.async.wait_group 3;" ::: "memory");   // the group with aligned rows row + 1 .. row + kGroup
            uint4 nxt;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(nxt.x), "=r"(nxt.y), "=r"(nxt.z), "=r"(nxt.w)
                         : "r"(stage + (uint32_t)(row % kSlots) * 512u) : "memory");
            const uint32_t t[8] = {cur.x, cur.y, cur.z, cur.w, nxt.x, nxt.y, nxt.z, nxt.w};
            uint32_t v[5];
#pragma unroll
            for (int k = 0; k < 5; k++) v[k] = wsel == 0 ? t[k] : wsel == 1 ? t[k + 1] : wsel == 2 ? t[k + 2] : t[k + 3];
#pragma unroll
            for (int wq = 0; wq < 4; wq++) raw[wq] = __funnelshift_r(v[wq], v[wq + 1], sh);
            cur = nxt;
            if (row % kGroup == kGroup - 1) issue_group();      // this group's slots are consumed: the group after the ring's takes them
        };
        auto big_endian = [&](uint32_t w) -> uint32_t {          // deciphered, first frame byte in the top bits
            if (!cipher) return __byte_perm(w, 0, 0x0123);
            return ((uint32_t)__ldg(cipher + (w & 0xFF)) << 24) | ((uint32_t)__ldg(cipher + ((w >> 8) & 0xFF)) << 16) |
                   ((uint32_t)__ldg(cipher + ((w >> 16) & 0xFF)) << 8) | (uint32_t)__ldg(cipher + (w >> 24));
        };
        for (int row = 0; row < full_rows; row++) {
            uint32_t raw[4], o[4];
            fetch_row(row, raw);
#pragma unroll
            for (int wq = 0; wq < 4; wq++) {
                crc = crc16_word(tb.crc, crc, raw[wq]);
                o[wq] = big_endian(raw[wq]);
            }
            reinterpret_cast<uint4*>(words)[row] = make_uint4(o[0], o[1], o[2], o[3]);
        }
        if (frame_size & 15) {                                   // last, partial row: bytes past the frame read as 0
            uint32_t raw[4], o[4];
            fetch_row(full_rows, raw);
            const int rem = frame_size & 15;
#pragma unroll
            for (int wq = 0; wq < 4; wq++) {
                const int nb = min(4, max(0, rem - 4 * wq));
                const uint32_t w = nb == 4 ? raw[wq] : nb == 0 ? 0u : raw[wq] & (0xFFFFFFFFu >> (32 - 8 * nb));
                if (nb == 4) crc = crc16_word(tb.crc, crc, w);
                else for (int k = 0; k < nb; k++) crc = crc16_byte(tb.crc[0], crc, (w >> (8 * k)) & 0xFF);
                o[wq] = big_endian(w);                           // every cipher table maps 0 to 0
            }
            reinterpret_cast<uint4*>(words)[full_rows] = make_uint4(o[0], o[1], o[2], o[3]);
        }
        reinterpret_cast<uint4*>(words)[nrows] = make_uint4(0, 0, 0, 0);     // slack for the prefetching reader
        reinterpret_cast<uint4*>(words)[nrows + 1] = make_uint4(0, 0, 0, 0);
        if (crc != 0) bad = true;                               // a valid frame's CRC over all its bytes is 0 (hca.cpp:1166)
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");        // the ring is the band tables' memory from here on
    __syncwarp();

    // ---- phase 2: frame header (hca.cpp:1162-1178), then per channel scalefactors (:1290-1358) and, as each one is
    // known, the band's resolution (:1444-1494). No HFR scales / intensity on this path.
    BitWindow br;
    br.init(words, (uint32_t)__cvta_generic_to_shared(s_dyn + (size_t)kFastWarps * (NCH * 128 * 32 * sizeof(uint16_t)) + (size_t)warp * 1024 + lane * 16));
    uint32_t packed = 0;
    if (active) {
        if (br.read(16, nbits) != 0xFFFF) bad = true;
        const uint32_t noise_level = br.read(9, nbits), boundary = br.read(7, nbits);
        packed = (noise_level << 8) - boundary;
    }
    const bool has_ath = active && S.ath != 0;                   // v2.0 streams have no ATH curve
    const uint8_t* ath = a.ath + (size_t)(has_ath ? S.ath : 0) * 128;
    const uint32_t min_res = S.min_res, max_res = S.max_res;
    const int hfr_groups = JOINT && active ? (int)S.hfr_groups : 0;
    uint64_t hfr_sf[NCH];                                        // 6-bit HFR scalefactors of the channel, group g at bit 6g
    int run_bits[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) hfr_sf[c] = 0;
    // A read that would cross the end of the frame yields 0 (hca.cpp:232-233). The per-read check is only compiled
    // into the variant used when some lane's frame is too short to be sure its header ends inside it.
    auto parse_channel = [&](int c, auto safe_tag) {
        constexpr bool kSafe = decltype(safe_tag)::value;
        uint16_t* tc = tab + c * 128 * 32;
        const int coded = active && !bad ? (int)S.coded[c] : 0;
        int sum_bits = 0;
        if (coded) {
            uint32_t db = br.peek(3);
            if (!kSafe && br.position() + 3 > nbits) db = 0;
            br.skip(3);
            // delta_bits 0: all zero; 1..5: 6 raw bits, then deltas with escape (1 << db) - 1 to 6 raw bits; >= 6: raw
            const int n_first = db == 0 ? 0 : 6;
            const int n_rest = db >= 6 ? 6 : (int)db;
            const bool delta = db >= 1 && db <= 5;
            const uint32_t escape = (1u << db) - 1, half = escape >> 1;
            uint32_t v = 0;
            for (int i = 0; i < coded; i++) {
                const int n1 = i == 0 ? n_first : n_rest;
                uint32_t d = br.peek(n1);
                uint32_t wide = br.peek(n1 + 6) & 63;            // the 6 raw bits behind an escape code
                if (!kSafe) {
                    const int pos = br.position();
                    if (pos + n1 > nbits) d = 0;
                    if (pos + n1 + 6 > nbits) wide = 0;
                }
                const bool rel = delta && i > 0;
                const bool esc = rel && d == escape;
                const int test = (int)v + (int)d - (int)half;
                if (rel && !esc && (test < 0 || test >= 64)) bad = true;    // HCA_ERROR_UNPACK (hca.cpp:1330-1333)
                v = esc ? wide : rel ? ((uint32_t)test & 63u) : d;
                br.skip(n1 + (esc ? 6 : 0));
                if ((i & 3) == 3) br.top_up();                   // at most 4 x 11 bits between refills
                uint32_t res = 0;
                if (v > 0) {
                    const int level = (has_ath ? (int)ath[i] : 0) + (int)((packed + (uint32_t)i) >> 8);
                    const int cp = level + 1 - (int)((5 * v) >> 1);
                    res = cp < 0 ? 15u : cp <= 65 ? (uint32_t)tb.invert[cp] : 0u;
                    res = min(max(res, min_res), max_res);
                }
                const uint32_t mb = tb.max_bits[res];
                sum_bits += (int)mb;
                tc[i * 32] = (uint16_t)((res << 2) | (v << 6) | (mb << 12));
            }
            br.top_up();
            // intensity indices of a secondary channel, HFR scales of the others (unpack_intensity, hca.cpp:1361-1441, v <= 2.0)
            if (!JOINT) {
            } else if (S.type[c] == 2) {
                uint32_t v0 = br.peek(4);
                if (!kSafe && br.position() + 4 > nbits) v0 = 0;
                uint32_t inten = v0;
                if (v0 < 15) {
                    br.skip(4);
                    for (int i = 1; i < 8; i++) {
                        uint32_t v = br.peek(4);
                        if (!kSafe && br.position() + 4 > nbits) v = 0;
                        br.skip(4);
                        inten |= v << (4 * i);
                    }
                    br.top_up();
                }
                a.inten[g] = inten;
                a.carry[g] = v0 < 15 ? 0 : 1;           // 15: the other seven keep the previous frame's values (:1368-1372)
            } else {
                uint64_t packed_sf = 0;
                for (int grp = 0; grp < hfr_groups; grp++) {
                    uint32_t v = br.peek(6);
                    if (!kSafe && br.position() + 6 > nbits) v = 0;
                    br.skip(6);
                    if ((grp & 3) == 3) br.top_up();
                    packed_sf |= (uint64_t)v << (6 * grp);
                }
                br.top_up();
                hfr_sf[c] = packed_sf;
            }
        }
        for (int i = coded; i < 128; i++) tc[i * 32] = 0;       // bands past the coded count: 0 from 0 bits
        run_bits[c] = sum_bits;
    };
    int hdr_bound = 32;                                          // most bits the header can take for this stream's geometry
#pragma unroll
    for (int c = 0; c < NCH; c++)
        hdr_bound += 3 + 6 + 11 * max((int)S.coded[c] - 1, 0) + (!JOINT ? 0 : S.type[c] == 2 ? 32 : 6 * hfr_groups);
    const bool hdr_safe = !a.force_careful && (!active || nbits >= hdr_bound);
    if (__all_sync(kFull, hdr_safe)) {
#pragma unroll
        for (int c = 0; c < NCH; c++) parse_channel(c, std::true_type{});
    } else {
#pragma unroll
        for (int c = 0; c < NCH; c++) parse_channel(c, std::false_type{});
    }
    if (bad) {                                               // nothing of a bad frame is decoded: every code is 0 bits
#pragma unroll
        for (int c = 0; c < NCH; c++)
            for (int i = 0; i < 128; i++) tab[(c * 128 + i) * 32] = 0;
    }
    __syncwarp();

    // ---- spectra: subframe-major, channel-minor runs of 128 codes (hca.cpp:1540-1571), dequantised on the way out
    const uint32_t W = r / RW, rr = r % RW;
    float4* dst_frame = a.spec + ((uint64_t)W * R + j) * (8 * 1024) + rr;   // + sub * 1024 + (c * RW) + chunk * 32
    const float* gain_tab = tb.gain;
    const float* code_tab = tb.code;
    // HFR geometry (clHCA_DecodeHeader, hca.cpp:872-874): bands [start, start + room) mirror [start - room, start)
    const int hfr_start = (int)S.base_bands + (int)S.stereo_bands;
    const bool hfr_on = JOINT && active && S.bands_per_hfr != 0;
    const int hfr_room = hfr_on ? max(0, min(min((int)S.total_bands - hfr_start, hfr_groups * (int)S.bands_per_hfr), hfr_start)) : 0;
    for (int sub = 0; sub < 8; sub++) {
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            // can this run cross the end of the frame? (only corrupt / wrongly keyed frames do)
            const bool careful = a.force_careful || br.position() + run_bits[c] > nbits;
            const bool any_careful = __any_sync(kFull, careful);
            const uint16_t* tp = tab + c * 128 * 32;
            float4* dst = dst_frame + sub * 1024 + c * RW;
            auto decode_run = [&](auto careful_tag) {
                constexpr bool kCareful = decltype(careful_tag)::value;
#pragma unroll 2
                for (int chunk = 0; chunk < 32; chunk++) {
                    float f[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t t = tp[(chunk * 4 + k) * 32];
                        const int bits = (int)(t >> 12);
                        const uint32_t res4 = t & 0x3C;
                        const uint32_t short_below = __funnelshift_rc(kShortBelowLo, kShortBelowHi, res4) & 15;
                        uint32_t code = br.peek(bits);
                        if (kCareful && br.position() + bits > nbits) code = 0;   // reader rule, hca.cpp:232-233
                        br.skip(bits - (code < short_below ? 1 : 0));
                        // value: sign-magnitude (resolution >= 8, LSB = sign) or prefix codebook (resolution <= 7)
                        const float q_hi = __uint_as_float(__float_as_uint((float)(code >> 1)) | (code << 31));
                        float q = q_hi;                                           // codes of the prefix family are < 16
                        if (!(t & 0x20)) q = *reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(code_tab) + (code << 5) + (res4 & 0x1C));
                        const float gain = *reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(gain_tab) + (t & 0xFFC));
                        f[k] = __fmul_rn(gain, q);                                // spectra = gain * q (hca.cpp:1568)
                    }
                    br.top_up();
                    if (active) __stcs(dst + chunk * 32, make_float4(f[0], f[1], f[2], f[3]));
                }
            };
            if (any_careful) decode_run(std::true_type{}); else decode_run(std::false_type{});
            // high-frequency reconstruction (hca.cpp:1638-1683): band start + n = conv[hfr scale - sf(low) + 63] * band
            // start - 1 - n, from this lane's own freshly stored row (intensity scaling happens later, in the transform
            // kernel, so the row still holds the unscaled spectra the reference mirrors)
            if (JOINT && __any_sync(kFull, hfr_on && S.type[c] != 2 && !bad)) {
                if (hfr_on && S.type[c] != 2 && !bad) {
                    float* row = reinterpret_cast<float*>(dst);
                    auto at = [&](int band) -> float* { return row + (size_t)(band >> 2) * 128 + (band & 3); };
                    const int bph = (int)S.bands_per_hfr;
                    int grp = 0, in_grp = 0;
                    for (int n = 0; n < hfr_room; n++) {
                        const int low = hfr_start - 1 - n, high = hfr_start + n;
                        const int sf_low = (int)((tp[low * 32] >> 6) & 63);
                        int k = (int)((hfr_sf[c] >> (6 * grp)) & 63) - sf_low + 63;
                        k = max(k, 0);
                        *at(high) = __fmul_rn(__uint_as_float(c_conv[k]), *at(low));
                        if (++in_grp == bph) { in_grp = 0; grp++; }
                    }
                    if (hfr_start + hfr_room >= 1) *at(hfr_start + hfr_room - 1) = 0.f;   // hca.cpp:1681 (band start - 1 if nothing was mirrored)
                }
            }
        }
    }
    if (bad && active) a.status[stream] = ERR_HCA_DECODE;
}

// ------------------------------------------------------------------------------------------------ transform
struct __align__(16) RowDesc {
    long long off;     // byte offset in the output blob of the row's sample 0 (may lie in front of the stream: delay)
    int lo, hi;        // valid samples [lo, hi) of the row's 128
};

__device__ __forceinline__ void load_spectra(float (&x)[128], const float4* __restrict__ src, bool ok) {
#pragma unroll
    for (int i = 0; i < 32; i++) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) v = __ldcs(src + i * 32);
        x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
    }
}

// PCM conversion (clHCA_ReadSamples16, hca.cpp:339-360): (int)(wave * 32768) clamped to int16. The window already
// carries the factor 32768. cvt.rzi.s16 truncates and saturates in one instruction; it differs from the reference's
// x86 (int) cast only for v >= 2^31 (the cast yields INT_MIN there), which cannot happen on this path: |gain * q| <
// scaling[63] = 11.32, the DCT-IV is orthonormal scaled (|dct| <= 0.125 * 128 * 11.32 = 181) and the window sums two
// products of factors below 1, so |v| < 2 * 181 * 32768 < 2^24.
__device__ __forceinline__ short pcm16_sat(float v) {
    short s;
    asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(s) : "f"(v));
    return s;
}

// One CTA of THREADS = 256 per SM (2 warps per scheduler; the 252 registers per thread leave room for no more).
// Measured on B200 (8192 stereo streams): with every sum a scalar instruction, 128 / 256 / 384 threads per SM all
// issued ~0.6 instructions per cycle and scheduler with "no instruction" the top stall -- the body is ~70 KB of
// straight-line code that every warp streams once per subframe, and instruction supply set the pace whatever the
// warp count. Issuing the paired sums two-wide (FADD2 / FFMA2, see hca_bfly2 / hca_sum2) cut the instruction count
// by a fifth and removed that stall; what remains is the fp32 pipe itself (two-wide forms take two pipe cycles:
// tools/ubench/fp32_rates.cu) and fixed-latency waits that two warps per scheduler cannot hide. The CTA-wide convoy
// barrier (warps fetch the same instruction-cache lines together) is worth ~5 %; barriers among the warps of one
// scheduler added nothing.
template <int NCH, int THREADS, int CONVOY, bool JOINT>
__global__ void __launch_bounds__(THREADS, 1)
hca_imdct_fast_kernel(HcaDecodeArgs a) {
    constexpr int RW = 32 / NCH;                 // runs (= tile rows) per warp
    constexpr int ROW_WORDS = 64 * NCH;          // 128 samples x NCH channels x int16
    constexpr int PITCH = ROW_WORDS + 1;         // odd word pitch: the column writes of the window are conflict-free
    constexpr int WARP_WORDS = RW * PITCH + RW * 4;
    extern __shared__ __align__(16) uint8_t s_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* carry = reinterpret_cast<float4*>(s_dyn) + threadIdx.x;                                  // [16][THREADS]
    uint32_t* tile = reinterpret_cast<uint32_t*>(s_dyn + 16 * THREADS * sizeof(float4)) + warp * WARP_WORDS;
    RowDesc* rows = reinterpret_cast<RowDesc*>(tile + RW * PITCH);

    const int rr = lane % RW, ch = lane / RW;
    const uint32_t W = a.run_base / RW + blockIdx.x * (THREADS / 32) + warp;
    const uint32_t R = a.run_len;
    const uint32_t r = W * RW + rr;
    const bool live = r < a.n_runs && r - a.run_base < a.run_count;
    const uint32_t G = (uint32_t)a.total_frames;
    uint32_t g = live ? r * R : G;
    uint32_t s = 0, f = 0, cnt = 0;
    if (live) {
        s = find_stream(a.dec_prefix, a.n_streams, g);
        const uint32_t p0 = __ldg(a.dec_prefix + s);
        f = g - p0;
        cnt = __ldg(a.dec_prefix + s + 1) - p0;
    }

    // CONVOY > 0: all warps of the CTA meet every CONVOY-th sync point of the generated code (one per 128 fp32
    // instructions), so they fetch the same instruction-cache lines at the same time
    auto convoy = [&](int k) {
        if (CONVOY > 0 && k % CONVOY == 0) __syncthreads();
    };
    // intensity stereo (apply_intensity_stereo, hca.cpp:1696-1714): bands [base, total) of both channels are the PRIMARY
    // channel's spectra times ratio / (2 - ratio); the primary sits NCH-pair lanes below (lane & (RW - 1)). HFR has
    // already been applied by the unpack kernel, on the unscaled spectra, as in the reference's order.
    auto intensity = [&](float (&x)[128], bool pair_joint, uint32_t inten, int sub, int base, int total) {
        if (!JOINT || NCH != 2 || !__any_sync(kFull, pair_joint)) return;
        const float rl = __uint_as_float(c_intensity[(inten >> (4 * sub)) & 15]);
        const float mine = ch == 0 ? rl : __fsub_rn(2.0f, rl);
#pragma unroll
        for (int i = 0; i < 128; i++) {
            const float left = __shfl_sync(kFull, x[i], lane & (RW - 1));
            if (pair_joint && i >= base && i < total) x[i] = __fmul_rn(left, mine);
        }
    };
    float x[128];
    {   // look-back: the DCT output of the last subframe in front of the run (zero at the start of a stream)
        const bool lb = live && f > 0;
        const uint32_t r1 = lb ? r - 1 : 0;
        const float4* src = a.spec + (((uint64_t)(r1 / RW) * R + (R - 1)) * 8 + 7) * 1024 + ch * RW + (r1 % RW);
        load_spectra(x, src, lb);
        {
            const HcaStreamDev& S = a.streams[s];
            const bool pj = JOINT && lb && S.type[ch] != 0;
            intensity(x, pj, pj ? __ldg(a.inten + g - 1) : 0u, 7, (int)S.base_bands, (int)S.total_bands);
        }
        hca_dct4_dec(x, a.one2, convoy);
        hca_carry_thread<THREADS>(x, carry);
    }

    const float4* src = a.spec + (uint64_t)W * R * (8 * 1024) + lane;     // block (j, sub) at + (j * 8 + sub) * 1024
    uint8_t* trow = reinterpret_cast<uint8_t*>(tile + rr * PITCH) + 2 * ch;
    load_spectra(x, src, live && g < G);
    for (uint32_t j = 0; j < R; j++, g++) {
        const bool ok = live && g < G;
        bool fresh = false;
        if (ok && f >= cnt) {                    // the run crosses into the next stream: the overlap state starts at zero
            do {
                s++;
                cnt = __ldg(a.dec_prefix + s + 1) - __ldg(a.dec_prefix + s);
            } while (cnt == 0);
            f = 0;
            fresh = true;
        }
        if (__any_sync(kFull, fresh)) {
            if (fresh) {
#pragma unroll
                for (int q = 0; q < 16; q++) carry[q * THREADS] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        long long out_off = 0;
        int out_samples = 0, delay = 0;
        bool pair_joint = false;
        uint32_t inten = 0;
        int jbase = 0, jtotal = 0;
        if (ok) {
            const HcaStreamDev& S = a.streams[s];
            out_off = (long long)S.out_off;
            out_samples = (int)S.out_samples;
            delay = (int)S.delay;
            if (JOINT && NCH == 2 && S.type[ch] != 0) {
                pair_joint = true;
                inten = __ldg(a.inten + g);
                jbase = (int)S.base_bands; jtotal = (int)S.total_bands;
            }
        }
        const bool ok_next_frame = live && j + 1 < R && g + 1 < G;
#pragma unroll 1
        for (int sub = 0; sub < 8; sub++) {
            intensity(x, pair_joint, inten, sub, jbase, jtotal);
            hca_dct4_dec(x, a.one2, convoy);
            if (ch == 0) {
                const long long n0 = (long long)f * 1024 + sub * 128 - delay;     // stream sample index of the row's sample 0
                RowDesc d;
                d.off = out_off + n0 * (2 * NCH);
                d.lo = ok ? (int)max(0ll, min(128ll, -n0)) : 0;
                d.hi = ok ? (int)max(0ll, min(128ll, (long long)out_samples - n0)) : 0;
                rows[rr] = d;
            }
            // the next block of spectra is loaded into the registers the window frees, 16 bytes at a time
            src += 1024;
            const bool ok_next = sub < 7 ? ok : ok_next_frame;
            // lanes without a next block (end of the run / of the job, idle lanes) read block 0 instead: their rows are
            // never stored, so the values do not matter, and the loads need no predicate and no zero fill
            const float4* nsrc = ok_next ? src : a.spec + lane;
            hca_window_thread<THREADS>(x, carry, a.one2,
                [](float v) { return pcm16_sat(v); },
                [&](int i, short v) { *reinterpret_cast<short*>(trow + i * (2 * NCH)) = v; },
                [&](int c) {
                    const float4 v = __ldcs(nsrc + c * 32);
                    x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
                }, convoy);
            __syncwarp();
            // ---- coalesced copy-out, one tile row (= 128 consecutive samples of one stream, all channels) at a time
#pragma unroll 4
            for (int row = 0; row < RW; row++) {
                const RowDesc d = rows[row];
                if (d.lo >= d.hi) continue;
                uint8_t* dst = a.out + d.off;
                const uint32_t* trw = tile + row * PITCH;
                const bool full = d.lo == 0 && d.hi == 128;
                if (NCH == 2) {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int w = lane + 32 * k;                      // word w = sample w (left | right << 16)
                        if (full || (w >= d.lo && w < d.hi)) reinterpret_cast<uint32_t*>(dst)[w] = trw[w];
                    }
                } else if (full && (d.off & 3) == 0) {
#pragma unroll
                    for (int k = 0; k < 2; k++) reinterpret_cast<uint32_t*>(dst)[lane + 32 * k] = trw[lane + 32 * k];
                } else {
                    const uint16_t* t16 = reinterpret_cast<const uint16_t*>(trw);
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int i = lane + 32 * k;
                        if (i >= d.lo && i < d.hi) reinterpret_cast<uint16_t*>(dst)[i] = t16[i];
                    }
                }
            }
            __syncwarp();
        }
        f++;
    }
}

#ifdef CRI_HCA_PAIR_KERNEL
// ------------------------------------------------------------------------------------------------ transform, warp pairs
// The same transform with every column split over a PAIR of warps: warp h of the pair keeps coefficients 64 h .. 64 h + 63
// of the pair's 32 columns (lane = column, as above) in 64 registers. Twelve of the fourteen passes, the window and the
// overlap state stay inside a half (tools/gen_dct.py: gen_imdct_half); for the two passes in the middle every lane
// publishes its 64 values in shared memory once, the pair meets at a named barrier and each lane recomputes the sums
// from its own and its partner's value. Half the registers per thread means twice the warps per SM for the same number
// of columns in flight -- the thread-resident kernel above sits at two warps per scheduler, which is what left it at
// about half of the fp32 rate -- and half the straight-line code per warp. The two warps of a pair run different
// instruction streams (all rotation factors are immediates), so the split is by warp, not by lane.
//
// Shared memory per pair: exchange area 2 x 16 x 32 float4 (16 KB), PCM tile (both warps' samples of a subframe meet
// there; rows leave coalesced, half of them by each warp), overlap state 8 float4 per thread.
template <int H, class Store, class Barrier, class Load, class Sync>
__device__ __forceinline__ void dct_half(float (&x)[64], const unsigned long long one, Store store, Barrier barrier, Load load, Sync sync) {
    if (H == 0) hca_dct4_dec_h0(x, one, store, barrier, load, sync);
    else hca_dct4_dec_h1(x, one, store, barrier, load, sync);
}

template <int NCH, int PAIRS, int CONVOY, bool JOINT>
struct PairXf {
    static constexpr int THREADS = PAIRS * 64;
    static constexpr int RW = 32 / NCH;                 // runs (= tile rows) per pair
    static constexpr int ROW_WORDS = 64 * NCH;          // 128 samples x NCH channels x int16
    static constexpr int PITCH = ROW_WORDS + 1;         // odd word pitch: the column writes of the window are conflict-free
    static constexpr int TILE_WORDS = RW * PITCH + RW * 4;
    static constexpr size_t kCarryBytes = (size_t)8 * THREADS * sizeof(float4);
    static constexpr size_t kXchgBytes = (size_t)2 * 16 * 32 * sizeof(float4);
    static constexpr size_t kPairBytes = kXchgBytes + (size_t)TILE_WORDS * sizeof(uint32_t);
    static constexpr size_t kSmem = kCarryBytes + PAIRS * kPairBytes;

    template <int H>
    static __device__ __forceinline__ void body(const HcaDecodeArgs& a, uint8_t* s_dyn) {
        const int lane = threadIdx.x & 31, pair = threadIdx.x >> 6;
        float4* carry = reinterpret_cast<float4*>(s_dyn) + threadIdx.x;                                  // [8][THREADS]
        uint8_t* pair_mem = s_dyn + kCarryBytes + (size_t)pair * kPairBytes;
        float4* xw = reinterpret_cast<float4*>(pair_mem) + H * 512 + lane;                                 // own half: [16][32]
        const float4* xr = reinterpret_cast<const float4*>(pair_mem) + (1 - H) * 512 + lane;               // the partner's
        uint32_t* tile = reinterpret_cast<uint32_t*>(pair_mem + kXchgBytes);
        RowDesc* rows = reinterpret_cast<RowDesc*>(tile + RW * PITCH);

        const int rr = lane % RW, ch = lane / RW;
        const uint32_t W = blockIdx.x * PAIRS + pair;
        const uint32_t R = a.run_len;
        const uint32_t r = W * RW + rr;
        const bool live = r < a.n_runs;
        const uint32_t G = (uint32_t)a.total_frames;
        uint32_t g = live ? r * R : G;
        uint32_t s = 0, f = 0, cnt = 0;
        if (live) {
            s = find_stream(a.dec_prefix, a.n_streams, g);
            const uint32_t p0 = __ldg(a.dec_prefix + s);
            f = g - p0;
            cnt = __ldg(a.dec_prefix + s + 1) - p0;
        }
        auto convoy = [&](int k) {
            if (CONVOY > 0 && k % CONVOY == 0) __syncthreads();
        };
        auto meet = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory"); };
        auto publish = [&](int q, float v0, float v1, float v2, float v3) { xw[q * 32] = make_float4(v0, v1, v2, v3); };
        auto partner = [&](int q) -> float4 { return xr[q * 32]; };
        // intensity stereo (see hca_imdct_fast_kernel): this half's bands only
        auto intensity = [&](float (&x)[64], bool pair_joint, uint32_t inten, int sub, int base, int total) {
            if (!JOINT || NCH != 2 || !__any_sync(kFull, pair_joint)) return;
            const float rl = __uint_as_float(c_intensity[(inten >> (4 * sub)) & 15]);
            const float mine = ch == 0 ? rl : __fsub_rn(2.0f, rl);
#pragma unroll
            for (int i = 0; i < 64; i++) {
                const float left = __shfl_sync(kFull, x[i], lane & (RW - 1));
                if (pair_joint && 64 * H + i >= base && 64 * H + i < total) x[i] = __fmul_rn(left, mine);
            }
        };
        auto load_half = [&](float (&x)[64], const float4* __restrict__ src, bool ok) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) v = __ldcs(src + (16 * H + i) * 32);
                x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
            }
        };
        float x[64];
        {   // look-back: the DCT output of the last subframe in front of the run (zero at the start of a stream)
            const bool lb = live && f > 0;
            const uint32_t r1 = lb ? r - 1 : 0;
            const float4* src = a.spec + (((uint64_t)(r1 / RW) * R + (R - 1)) * 8 + 7) * 1024 + ch * RW + (r1 % RW);
            load_half(x, src, lb);
            {
                const HcaStreamDev& S = a.streams[s];
                const bool pj = JOINT && lb && S.type[ch] != 0;
                intensity(x, pj, pj ? __ldg(a.inten + g - 1) : 0u, 7, (int)S.base_bands, (int)S.total_bands);
            }
            dct_half<H>(x, a.one2, publish, meet, partner, convoy);
            if (H == 0) hca_carry_h0<THREADS>(x, carry); else hca_carry_h1<THREADS>(x, carry);
            meet();                                   // the partner has read this half's published values
        }

        const float4* src = a.spec + (uint64_t)W * R * (8 * 1024) + lane;     // block (j, sub) at + (j * 8 + sub) * 1024
        uint8_t* trow = reinterpret_cast<uint8_t*>(tile + rr * PITCH) + 2 * ch;
        load_half(x, src, live && g < G);
        for (uint32_t j = 0; j < R; j++, g++) {
            const bool ok = live && g < G;
            bool fresh = false;
            if (ok && f >= cnt) {                    // the run crosses into the next stream: the overlap state starts at zero
                do {
                    s++;
                    cnt = __ldg(a.dec_prefix + s + 1) - __ldg(a.dec_prefix + s);
                } while (cnt == 0);
                f = 0;
                fresh = true;
            }
            if (__any_sync(kFull, fresh)) {
                if (fresh) {
#pragma unroll
                    for (int q = 0; q < 8; q++) carry[q * THREADS] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            long long out_off = 0;
            int out_samples = 0, delay = 0;
            bool pair_joint = false;
            uint32_t inten = 0;
            int jbase = 0, jtotal = 0;
            if (ok) {
                const HcaStreamDev& S = a.streams[s];
                out_off = (long long)S.out_off;
                out_samples = (int)S.out_samples;
                delay = (int)S.delay;
                if (JOINT && NCH == 2 && S.type[ch] != 0) {
                    pair_joint = true;
                    inten = __ldg(a.inten + g);
                    jbase = (int)S.base_bands; jtotal = (int)S.total_bands;
                }
            }
            const bool ok_next_frame = live && j + 1 < R && g + 1 < G;
#pragma unroll 1
            for (int sub = 0; sub < 8; sub++) {
                intensity(x, pair_joint, inten, sub, jbase, jtotal);
                dct_half<H>(x, a.one2, publish, meet, partner, convoy);
                if (ch == 0 && H == 0) {
                    const long long n0 = (long long)f * 1024 + sub * 128 - delay;     // stream sample index of the row's sample 0
                    RowDesc d;
                    d.off = out_off + n0 * (2 * NCH);
                    d.lo = ok ? (int)max(0ll, min(128ll, -n0)) : 0;
                    d.hi = ok ? (int)max(0ll, min(128ll, (long long)out_samples - n0)) : 0;
                    rows[rr] = d;
                }
                src += 1024;
                const bool ok_next = sub < 7 ? ok : ok_next_frame;
                const float4* nsrc = ok_next ? src : a.spec + lane;   // idle lanes read block 0: never stored, no predicate needed
                auto cvt = [](float v) { return pcm16_sat(v); };
                auto emit = [&](int i, short v) { *reinterpret_cast<short*>(trow + i * (2 * NCH)) = v; };
                auto refill = [&](int c) {
                    const float4 v = __ldcs(nsrc + (16 * H + c) * 32);
                    x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
                };
                if (H == 0) hca_window_h0<THREADS>(x, carry, a.one2, cvt, emit, refill);
                else hca_window_h1<THREADS>(x, carry, a.one2, cvt, emit, refill);
                meet();                              // both halves' samples are in the tile
                // ---- coalesced copy-out, one tile row (= 128 consecutive samples of one stream, all channels) at a time;
                // the rows are shared out between the two warps
#pragma unroll 4
                for (int row = H; row < RW; row += 2) {
                    const RowDesc d = rows[row];
                    if (d.lo >= d.hi) continue;
                    uint8_t* dst = a.out + d.off;
                    const uint32_t* trw = tile + row * PITCH;
                    const bool full = d.lo == 0 && d.hi == 128;
                    if (NCH == 2) {
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int w = lane + 32 * k;                      // word w = sample w (left | right << 16)
                            if (full || (w >= d.lo && w < d.hi)) reinterpret_cast<uint32_t*>(dst)[w] = trw[w];
                        }
                    } else if (full && (d.off & 3) == 0) {
#pragma unroll
                        for (int k = 0; k < 2; k++) reinterpret_cast<uint32_t*>(dst)[lane + 32 * k] = trw[lane + 32 * k];
                    } else {
                        const uint16_t* t16 = reinterpret_cast<const uint16_t*>(trw);
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int i = lane + 32 * k;
                            if (i >= d.lo && i < d.hi) reinterpret_cast<uint16_t*>(dst)[i] = t16[i];
                        }
                    }
                }
            }
            f++;
        }
    }
};

template <int NCH, int PAIRS, int CONVOY, bool JOINT>
__global__ void __launch_bounds__(PAIRS * 64, 1)
hca_imdct_pair_kernel(const __grid_constant__ HcaDecodeArgs a) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    if ((threadIdx.x >> 5) & 1) PairXf<NCH, PAIRS, CONVOY, JOINT>::template body<1>(a, s_dyn);
    else PairXf<NCH, PAIRS, CONVOY, JOINT>::template body<0>(a, s_dyn);
}

#ifndef HCA_PAIRS
#define HCA_PAIRS 6
#endif
constexpr int kPairs = HCA_PAIRS;             // warp pairs per CTA (one CTA per SM): 12 warps, 192 columns in flight

template <int NCH, int PAIRS, int CONVOY, bool JOINT>
void launch_xf_pair(const HcaDecodeArgs& a, cudaStream_t s) {
    using K = PairXf<NCH, PAIRS, CONVOY, JOINT>;
    cudaFuncSetAttribute(hca_imdct_pair_kernel<NCH, PAIRS, CONVOY, JOINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::kSmem);
    const uint32_t col_warps = (a.n_runs + K::RW - 1) / K::RW;
    hca_imdct_pair_kernel<NCH, PAIRS, CONVOY, JOINT><<<(col_warps + PAIRS - 1) / PAIRS, PAIRS * 64, K::kSmem, s>>>(a);
}

#endif  // CRI_HCA_PAIR_KERNEL

constexpr int kXfThreads = 256;          // one launch per kernel: one transform CTA owns an SM
constexpr int kXfThreadsPiped = 128;     // pipelined launches: half an SM's registers, the rest hosts unpack CTAs

template <int NCH, int THREADS, int CONVOY, bool JOINT>
void launch_xf(const HcaDecodeArgs& a, cudaStream_t s) {
    constexpr int RW = 32 / NCH;
    const size_t smem_t = 16 * THREADS * sizeof(float4) + (size_t)(THREADS / 32) * (RW * (64 * NCH + 1) + RW * 4) * sizeof(uint32_t);
    cudaFuncSetAttribute(hca_imdct_fast_kernel<NCH, THREADS, CONVOY, JOINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t);
    const uint32_t runs = std::min(a.run_count, a.n_runs - std::min(a.n_runs, a.run_base));
    const uint32_t warps = (runs + RW - 1) / RW, per_cta = THREADS / 32;
    if (!warps) return;
    hca_imdct_fast_kernel<NCH, THREADS, CONVOY, JOINT><<<(warps + per_cta - 1) / per_cta, THREADS, smem_t, s>>>(a);
}

template <int NCH, bool JOINT>
void launch_unpack(const HcaDecodeArgs& a, cudaStream_t s) {
    const size_t smem_u = (size_t)kFastWarps * (NCH * 128 * 32 * sizeof(uint16_t) + 1024);   // band tables + the bit readers' rings
    cudaFuncSetAttribute(hca_unpack_fast_kernel<NCH, JOINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_u);
    const uint32_t runs = std::min(a.run_count, a.n_runs - std::min(a.n_runs, a.run_base));
    const uint64_t unpack_warps = (uint64_t)((runs + 31) / 32) * a.run_len;
    if (!unpack_warps) return;
    hca_unpack_fast_kernel<NCH, JOINT><<<(unsigned)((unpack_warps + kFastWarps - 1) / kFastWarps), kFastThreads, smem_u, s>>>(a);
}

template <int NCH, bool JOINT>
void launch_fast(HcaDecodeArgs a, cudaStream_t s, uint64_t* launches, cudaEvent_t mid, const HcaFastPipe* pipe) {
#ifdef CRI_HCA_PIPELINED
    const uint32_t chunks = pipe && pipe->side ? std::min<uint32_t>(pipe->chunks, 16) : 1;
#else
    const uint32_t chunks = 1;
    (void)pipe;
#endif
    if (chunks <= 1) {
        a.run_base = 0;
        a.run_count = a.n_runs;
        launch_unpack<NCH, JOINT>(a, s);
        ++*launches;
        if (JOINT && a.carry_scan) launch_hca_intensity_scan(a, s, launches);
        if (mid) cudaEventRecord(mid, s);
#ifdef CRI_HCA_PAIR_KERNEL
        static const int xf = [] { const char* e = getenv("CRI_HCA_XF"); return e && *e ? atoi(e) : 0; }();
        if (xf) launch_xf_pair<NCH, kPairs, 0, JOINT>(a, s);
        else
#endif
        launch_xf<NCH, kXfThreads, 2, JOINT>(a, s);   // CONVOY = 2: the CTA meets every 256 fp32 instructions (measured: -5 % vs none)
        ++*launches;
        return;
    }
#ifdef CRI_HCA_PIPELINED
    // pipelined: chunk k's transform (side stream, high priority, 128-thread CTAs) runs beside chunk k + 1's unpack
    const uint32_t gran = 128;                                            // runs: whole unpack warps and whole transform CTAs
    const uint32_t per = ((a.n_runs + chunks - 1) / chunks + gran - 1) / gran * gran;
    cudaEventRecord(pipe->start, s);
    cudaStreamWaitEvent(pipe->side, pipe->start, 0);
    uint32_t k = 0;
    for (uint32_t base = 0; base < a.n_runs; base += per, k++) {
        a.run_base = base;
        a.run_count = std::min(per, a.n_runs - base);
        launch_unpack<NCH, JOINT>(a, s);
        ++*launches;
        cudaEventRecord(pipe->unpacked[k], s);
        cudaStreamWaitEvent(pipe->side, pipe->unpacked[k], 0);
        launch_xf<NCH, kXfThreadsPiped, 2, JOINT>(a, pipe->side);
        ++*launches;
    }
    if (mid) cudaEventRecord(mid, s);                                      // all unpack kernels are done here
    cudaEventRecord(pipe->done, pipe->side);
    cudaStreamWaitEvent(s, pipe->done, 0);
#endif
}

}  // namespace

uint32_t hca_fast_threads_per_cta() {          // columns per transform CTA (the host sizes run_len by it)
#ifdef CRI_HCA_PAIR_KERNEL
    const char* e = getenv("CRI_HCA_XF");
    if (e && *e && atoi(e) != 0) return kPairs * 32;
#endif
    return hca_fast_chunks() > 1 ? kXfThreadsPiped : kXfThreads;
}
uint32_t hca_fast_ctas_per_sm() { return 1; }

uint32_t hca_fast_chunks() {
#ifdef CRI_HCA_PIPELINED          // experiment ("measured and dropped"): build with -DCRI_HCA_PIPELINED, run with CRI_HCA_CHUNKS=k
    const char* e = getenv("CRI_HCA_CHUNKS");
    const int v = e && *e ? atoi(e) : 1;
    return (uint32_t)std::min(std::max(v, 1), 16);
#else
    return 1;
#endif
}

void launch_hca_decode_fast(const HcaDecodeArgs& a, cudaStream_t s, uint64_t* launches, cudaEvent_t mid, const HcaFastPipe* pipe) {
    if (!a.n_runs) return;
#ifdef CRI_DEV_ONE_VARIANT       // development builds: one template variant only (compile time)
    launch_fast<2, false>(a, s, launches, mid, pipe);
#else
    if (a.uniform == 2) {
        if (a.joint) launch_fast<2, true>(a, s, launches, mid, pipe); else launch_fast<2, false>(a, s, launches, mid, pipe);
    } else {
        if (a.joint) launch_fast<1, true>(a, s, launches, mid, pipe); else launch_fast<1, false>(a, s, launches, mid, pipe);
    }
#endif
}

}  // namespace cri
