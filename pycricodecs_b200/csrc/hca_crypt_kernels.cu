// HCA frame en/decryption kernel for sm_100a.
//
// Reference: HcaCrypt (CriCodecs/hca.cpp:3271-3337) maps every byte of every
// frame through a 256-entry substitution table derived from the key
// (cipher_decrypt, :491-497; the inverse table when encrypting, :3315-3320) and
// rewrites the frame's trailing CRC16 over the first frame_size-2 mapped bytes
// (:3326). The header half (signature masks, ciph type, header CRC) is done on
// the host (formats.cpp: crypt_header) and scattered as a patch.
//
// Two kernels. hca_crypt_staged_kernel (frames up to 12 KB, i.e. every real stream): one WARP per group of up to 32
// consecutive frames of one stream -- contiguous in the blob, so the group moves HBM -> shared memory -> HBM as
// coalesced 16-byte rows (cp.async in, uint4 out; input and output share their byte offsets) -- and one LANE per
// frame works on it in shared memory: substitution table (shared), CRC16 four bytes per step (slice-by-4 tables,
// shared). HBM traffic is the compulsory 2 x frame_size per frame. hca_crypt_kernel is the any-size path: one lane
// per frame straight on global memory.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <type_traits>

#include "hca_kernels.h"

namespace cri {
namespace {

constexpr int kCryptThreads = 128;
constexpr int kSmemTables = 16;

__device__ __forceinline__ uint32_t crc16_byte(uint32_t crc, uint32_t byte) {   // poly 0x8005, MSB first (hca.cpp:205-211)
    const uint32_t v = ((crc >> 8) ^ byte) & 0xFF;
    const uint32_t t = (v << 1) ^ (v << 2) ^ ((__popc(v) & 1) ? 0x8003u : 0u);
    return ((crc << 8) ^ t) & 0xFFFF;
}

__global__ void __launch_bounds__(kCryptThreads)
hca_crypt_kernel(HcaCryptArgs a) {
    __shared__ uint8_t s_tab[kSmemTables * 256];
    const uint32_t cached = min(a.n_tables, (uint32_t)kSmemTables);
    for (uint32_t i = threadIdx.x; i < cached * 256; i += blockDim.x) s_tab[i] = a.tables[i];
    __syncthreads();
    const uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.n_frames) return;
    // frame -> stream by binary search over the exclusive prefix of frame counts
    uint32_t lo = 0, hi = a.n_streams;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a.frame_prefix[mid] <= f) lo = mid; else hi = mid;
    }
    const HcaStreamDev& S = a.streams[lo];
    const uint32_t n = S.frame_size;
    const uint64_t off = S.in_off + (f - a.frame_prefix[lo]) * n;   // same offset in both blobs
    const uint8_t* src = a.in + off;
    uint8_t* dst = a.out + off;
    const uint8_t* tab = S.cipher < cached ? s_tab + S.cipher * 256 : a.tables + (size_t)S.cipher * 256;
    const uint32_t body = n - 2;                                    // bytes covered by the rewritten CRC
    uint32_t crc = 0;
    uint32_t i = 0;
    const uint32_t head = min(n, (uint32_t)((4 - (reinterpret_cast<uintptr_t>(src) & 3)) & 3));
    for (; i < head; i++) {
        const uint32_t b = tab[src[i]];
        if (i < body) crc = crc16_byte(crc, b);
        dst[i] = (uint8_t)b;
    }
    for (; i + 4 <= n; i += 4) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(src + i);
        uint32_t o = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t b = tab[(w >> (8 * k)) & 0xFF];
            if (i + k < body) crc = crc16_byte(crc, b);
            o |= b << (8 * k);
        }
        *reinterpret_cast<uint32_t*>(dst + i) = o;
    }
    for (; i < n; i++) {
        const uint32_t b = tab[src[i]];
        if (i < body) crc = crc16_byte(crc, b);
        dst[i] = (uint8_t)b;
    }
    dst[n - 2] = (uint8_t)(crc >> 8);
    dst[n - 1] = (uint8_t)crc;
}

// ---- staged kernel
constexpr int kStageWarps = 4;

__device__ __forceinline__ uint32_t crc16_word(const uint16_t (&T)[4][256], uint32_t c, uint32_t w) {   // w: 4 bytes, memory order
    const uint32_t x = w ^ __byte_perm(c, 0, 0x4401);
    return (uint32_t)T[3][x & 0xFF] ^ (uint32_t)T[2][(x >> 8) & 0xFF] ^ (uint32_t)T[1][(x >> 16) & 0xFF] ^ (uint32_t)T[0][x >> 24];
}
__device__ __forceinline__ uint32_t crc16_tab(const uint16_t (&T0)[256], uint32_t c, uint32_t byte) {
    return ((c << 8) & 0xFFFF) ^ (uint32_t)T0[((c >> 8) ^ byte) & 0xFF];
}

__global__ void __launch_bounds__(kStageWarps * 32)
hca_crypt_staged_kernel(HcaCryptArgs a) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ uint8_t s_tab[kSmemTables * 256];
    __shared__ uint16_t s_crc[4][256];
    const uint32_t cached = min(a.n_tables, (uint32_t)kSmemTables);
    for (uint32_t i = threadIdx.x; i < cached * 256; i += blockDim.x) s_tab[i] = a.tables[i];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = crc16_byte(0, i);
#pragma unroll
        for (int k = 0; k < 4; k++) { s_crc[k][i] = (uint16_t)c; c = crc16_byte(c, 0); }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t gid = (uint64_t)blockIdx.x * kStageWarps + warp;
    if (gid >= a.n_groups) return;                      // whole warp
    uint32_t lo = 0, hi = a.n_streams;                  // group -> stream
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a.group_prefix[mid] <= gid) lo = mid; else hi = mid;
    }
    const HcaStreamDev& S = a.streams[lo];
    const uint32_t fs = S.frame_size;
    const uint32_t f0 = (uint32_t)(gid - a.group_prefix[lo]) * a.frames_per_group;
    const uint32_t cnt = min(a.frames_per_group, S.frame_count - f0);
    const uint64_t off0 = S.in_off + (uint64_t)f0 * fs;  // same offset in both blobs
    const uint32_t bytes = cnt * fs;
    uint8_t* buf = s_dyn + (size_t)warp * a.group_bytes;
    const uint32_t lead = (uint32_t)(off0 & 15);
    const uint64_t base = off0 - lead;
    const uint32_t nrows = (lead + bytes + 15) >> 4;
    // ---- in: whole 16-byte rows (the blob starts 256-byte aligned and has 64 bytes of slack behind it)
    for (uint32_t row = lane; row < nrows; row += 32) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(buf + row * 16)),
                     "l"(a.in + base + (uint64_t)row * 16));
    }
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // ---- one lane per frame, in place
    if ((uint32_t)lane < cnt) {
        uint8_t* p = buf + lead + (uint32_t)lane * fs;
        const uint8_t* tab = S.cipher < cached ? s_tab + S.cipher * 256 : nullptr;
        const uint8_t* gtab = a.tables + (size_t)S.cipher * 256;
        auto map = [&](uint32_t b) -> uint32_t { return tab ? tab[b] : __ldg(gtab + b); };
        const uint32_t body = fs - 2;                   // bytes covered by the rewritten CRC (hca.cpp:3326)
        uint32_t crc = 0, i = 0;
        const uint32_t head = min(fs, (uint32_t)((4 - ((uint32_t)__cvta_generic_to_shared(p) & 3)) & 3));
        for (; i < head; i++) {
            const uint32_t b = map(p[i]);
            if (i < body) crc = crc16_tab(s_crc[0], crc, b);
            p[i] = (uint8_t)b;
        }
        for (; i + 4 <= body; i += 4) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(p + i);
            const uint32_t o = map(w & 0xFF) | (map((w >> 8) & 0xFF) << 8) | (map((w >> 16) & 0xFF) << 16) | (map(w >> 24) << 24);
            crc = crc16_word(s_crc, crc, o);
            *reinterpret_cast<uint32_t*>(p + i) = o;
        }
        for (; i < body; i++) {
            const uint32_t b = map(p[i]);
            crc = crc16_tab(s_crc[0], crc, b);
            p[i] = (uint8_t)b;
        }
        p[fs - 2] = (uint8_t)(crc >> 8);
        p[fs - 1] = (uint8_t)crc;
    }
    __syncwarp();
    // ---- out: interior rows as 16 bytes, the group's first and last partial rows byte by byte
    const uint32_t end = lead + bytes;
    const uint32_t head_end = lead ? min(16u, end) : 0u;
    const uint32_t r0 = lead ? 1u : 0u;
    const uint32_t r1 = max(end >> 4, r0);
    for (uint32_t row = r0 + lane; row < r1; row += 32)
        *reinterpret_cast<uint4*>(a.out + base + (uint64_t)row * 16) = *reinterpret_cast<const uint4*>(buf + row * 16);
    for (uint32_t k = lead + lane; k < head_end; k += 32) a.out[base + k] = buf[k];
    for (uint32_t k = max(r1 * 16, head_end) + lane; k < end; k += 32) a.out[base + k] = buf[k];
}


// ---- LUT kernel: the streaming path (frames of 128 bytes .. 12 KB, i.e. every real stream)
//
// Persistent CTAs: eight consumer warps and one producer lane around a ring of staging slots in shared memory. One item =
// one group (up to 32 consecutive frames of one stream = one contiguous byte range of the blob). The producer moves
// groups HBM -> shared memory with one bulk copy each (cp.async.bulk, completion on the slot's mbarrier), a consumer
// warp sends its finished slot back with one bulk store, so moving the data costs no issue slots and the next groups
// load while the current ones are worked on. In between, one LANE per frame works in place. Per 32-bit word of the frame:
//   * substitution: four lookups in a copy of the stream's 256-byte table that is replicated once per shared-memory
//     bank ([byte value][lane], 32 KB), so the 32 lanes' data-dependent lookups never conflict;
//   * CRC16 without tables. P = x^16 + x^15 + x^2 + 1 = (x + 1)(x^15 + x + 1). Modulo x + 1 the message reduces to its
//     parity (one XOR per word). Modulo Q = x^15 + x + 1 squaring gives x^(15 * 2^j) = x^(2^j) + 1, in particular
//     x^480 = x^32 + 1: the message is taken in blocks of fifteen 32-bit words and the running value A (15 words in
//     registers) advances by A * x^480 + B = (A << 32) ^ A ^ B, i.e. ONE three-input XOR per word plus two for the
//     word that overflows. A is reduced to 15 bits once per frame (x^30 = x^2 + 1, x^15 = x + 1), the two residues
//     are recombined (R = E ^ (parity(E) ^ parity(M) ? 0x8003 : 0)) and written into the frame's last two bytes.
// tools/crc_model.py restates the arithmetic in Python and checks it against the bitwise CRC.
constexpr int kLutConsumers = 8;                  // warps that work on frames
constexpr int kLutMaxSlots = 9;                    // staging buffers in the ring (as many as shared memory holds, at least 2)
constexpr uint32_t kLutTableBytes = 256 * 32 * 4;
constexpr uint32_t kLutCtrlBytes = 256;            // mbarriers + the item counter
constexpr uint32_t kLutFront = 16;                 // bytes in front of a slot's data (words read before a frame's start)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// running value r (< 2^17) times x^32 plus the next big-endian message word, modulo Q (result < 2^16)
__device__ __forceinline__ uint32_t modq_word(uint32_t r, uint32_t o) {
    uint32_t v = (r << 4) ^ (r << 2) ^ o;                              // x^32 = x^4 + x^2  (x^30 = x^2 + 1)
    uint32_t h = v >> 30; v = (v & 0x3FFFFFFFu) ^ h ^ (h << 2);
    h = v >> 15; v = (v & 0x7FFFu) ^ h ^ (h << 1);
    return v;
}
// the same for one more message byte (r < 2^17; result < 2^15: fully reduced)
__device__ __forceinline__ uint32_t modq_byte(uint32_t r, uint32_t m) {
    const uint32_t v = (r << 8) ^ m, h = v >> 15;
    return (v & 0x7FFFu) ^ h ^ (h << 1);
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

struct CryptGroup {          // one group = up to 32 consecutive frames of one stream = one contiguous byte range
    uint32_t stream, fs, cnt, lead, nrows, bytes;
    uint64_t base;           // 16-byte aligned start in both blobs
};
__device__ __forceinline__ CryptGroup crypt_group(const HcaCryptArgs& a, uint64_t gid) {
    const uint2 e = __ldg(a.group_table + gid);              // host-built: no search (a search is ~13 dependent loads per group)
    const uint32_t lo = e.x;
    const HcaStreamDev& S = a.streams[lo];
    CryptGroup g;
    g.stream = lo;
    g.fs = S.frame_size;
    const uint32_t f0 = e.y;
    g.cnt = min(a.frames_per_group, S.frame_count - f0);
    const uint64_t off0 = S.in_off + (uint64_t)f0 * g.fs;     // same offset in both blobs
    g.bytes = g.cnt * g.fs;
    g.lead = (uint32_t)(off0 & 15);
    g.base = off0 - g.lead;
    g.nrows = (g.lead + g.bytes + 15) >> 4;
    return g;
}

// One lane, one frame, in place in shared memory: substitution + CRC rewrite (see the kernel's header comment).
template <bool kFast>
__device__ __forceinline__ void crypt_frame(uint8_t* dat, uint32_t q0, uint32_t fs, const uint32_t* lut_lane, const uint8_t* gtab) {
    const uint32_t e1 = q0 + fs - 3;                            // last byte covered by the CRC (hca.cpp:3326)
    const uint32_t wE = e1 >> 2, k_tail = (e1 & 3) + 1;         // word of that byte; CRC-covered bytes in it
    const uint32_t N = ((fs - 3) >> 2) + 1;                     // full words in front of it, same for every lane
    uint32_t* wp = reinterpret_cast<uint32_t*>(dat) + (int)(wE - N);   // may start one word in front of the frame
    auto map4 = [&](uint32_t w) -> uint32_t {                   // four substitutions, memory order kept
        uint32_t x0, x1, x2, x3;
        if (kFast) {
            x0 = lut_lane[(w & 0xFF) << 5]; x1 = lut_lane[((w >> 8) & 0xFF) << 5];
            x2 = lut_lane[((w >> 16) & 0xFF) << 5]; x3 = lut_lane[(w >> 24) << 5];
        } else {
            x0 = __ldg(gtab + (w & 0xFF)); x1 = __ldg(gtab + ((w >> 8) & 0xFF));
            x2 = __ldg(gtab + ((w >> 16) & 0xFF)); x3 = __ldg(gtab + (w >> 24));
        }
        return __byte_perm(__byte_perm(x0, x1, 0x3340), __byte_perm(x2, x3, 0x4033), 0x7610);
    };
    uint32_t par = 0, r = 0;
    // words 0 and 1: bytes in front of the frame belong to the neighbour (not stored, zero for the CRC)
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const uint32_t addr = (wE - N + i) * 4;                 // wraps below zero for a word in front of the slot: then drop = 4
        const uint32_t drop = (int32_t)(q0 - addr) <= 0 ? 0u : min(q0 - addr, 4u);
        uint32_t w = 0;                                         // only this frame's bytes are read: the others are the
        if (drop == 0) w = wp[i];                               // neighbour lane's tail, which that lane is rewriting
        else for (uint32_t b = drop; b < 4; b++) w |= (uint32_t)reinterpret_cast<const uint8_t*>(wp + i)[b] << (8 * b);
        uint32_t o = map4(w);
        o = drop >= 4 ? 0u : o & (0xFFFFFFFFu << (8 * drop));
        if (drop == 0) wp[i] = o;
        else for (uint32_t b = drop; b < 4; b++) reinterpret_cast<uint8_t*>(wp + i)[b] = (uint8_t)(o >> (8 * b));
        par ^= o;
        r = modq_word(r, __byte_perm(o, 0, 0x0123));
    }
    const uint32_t blocks = (N - 2) / 15, rest = (N - 2) % 15;
    uint32_t* p = wp + 2;
    if (blocks) {
        uint32_t acc[15];
#pragma unroll
        for (int i = 0; i < 14; i++) acc[i] = 0;
        acc[14] = r;
        for (uint32_t b = 0; b < blocks; b++, p += 15) {
            const uint32_t a0 = acc[0];
#pragma unroll
            for (int i = 0; i < 15; i++) {
                const uint32_t o = map4(p[i]);
                p[i] = o;
                par ^= o;
                const uint32_t be = __byte_perm(o, 0, 0x0123);
                if (i < 13) acc[i] = acc[i] ^ acc[i + 1] ^ be;
                else if (i == 13) acc[13] = acc[13] ^ acc[14] ^ be ^ a0;
                else acc[14] = acc[14] ^ a0 ^ be;
            }
        }
        r = 0;
#pragma unroll
        for (int i = 0; i < 15; i++) r = modq_word(r, acc[i]);
    }
    for (uint32_t i = 0; i < rest; i++) {
        const uint32_t o = map4(p[i]);
        p[i] = o;
        par ^= o;
        r = modq_word(r, __byte_perm(o, 0, 0x0123));
    }
    {   // the word with the last CRC-covered byte: byte stores (what follows is the CRC field and the next frame)
        uint8_t* pb = reinterpret_cast<uint8_t*>(p + rest);
        uint32_t w = 0;
        if (k_tail == 4) w = p[rest];
        else for (uint32_t b = 0; b < k_tail; b++) w |= (uint32_t)pb[b] << (8 * b);
        const uint32_t o = map4(w);
        for (uint32_t b = 0; b < k_tail; b++) {
            const uint32_t m = (o >> (8 * b)) & 0xFF;
            pb[b] = (uint8_t)m;
            par ^= m;
            r = modq_byte(r, m);
        }
    }
    r = modq_byte(modq_byte(r, 0), 0);                          // message * x^16
    const uint32_t t = (__popc(r) ^ __popc(par)) & 1;
    const uint32_t crc = r ^ (t ? 0x8003u : 0u);
    dat[q0 + fs - 2] = (uint8_t)(crc >> 8);
    dat[q0 + fs - 1] = (uint8_t)crc;
}

// Warp kLutConsumers is the producer: one lane keeps every free slot of the ring loading (bulk copies complete on the
// slot's `full` mbarrier). The other warps take items in order from a shared counter, work on the slot in place, send
// it out with a bulk store and hand the slot back through its `empty` mbarrier once the store has read it.
__global__ void __launch_bounds__((kLutConsumers + 1) * 32, 1)
hca_crypt_lut_kernel(HcaCryptArgs a, uint32_t n_slots) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    uint32_t* lut = reinterpret_cast<uint32_t*>(s_dyn);                                   // [256][32]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_dyn + kLutTableBytes);                 // full[kLutMaxSlots], empty[kLutMaxSlots]
    uint32_t* counter = reinterpret_cast<uint32_t*>(s_dyn + kLutTableBytes + 2 * kLutMaxSlots * 8);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t slot_bytes = kLutFront + a.group_bytes;
    uint8_t* slots = s_dyn + kLutTableBytes + kLutCtrlBytes;
    const uint8_t* lut_src = a.tables + (size_t)a.lut_table * 256;
    for (uint32_t i = threadIdx.x; i < 256 * 32; i += blockDim.x) lut[i] = lut_src[i >> 5];
    if (threadIdx.x == 0) {
        for (uint32_t k = 0; k < 2 * kLutMaxSlots; k++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + k)));
        *counter = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint64_t items = a.n_groups > blockIdx.x ? (a.n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // of this CTA
    if (warp == kLutConsumers) {
        if (lane != 0) return;
        for (uint64_t t = 0; t < items; t++) {
            const uint32_t slot = (uint32_t)(t % n_slots), round = (uint32_t)(t / n_slots);
            if (round) mbar_wait(smem_u32(bars + kLutMaxSlots + slot), (round - 1) & 1);   // the slot's previous item is out
            const CryptGroup g = crypt_group(a, blockIdx.x + t * gridDim.x);
            const uint32_t full = smem_u32(bars + slot);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(g.nrows * 16) : "memory");
            // whole 16-byte rows (the blob starts 256-byte aligned and has 64 bytes of slack behind it)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(slots + (size_t)slot * slot_bytes + kLutFront)), "l"(a.in + g.base), "r"(g.nrows * 16), "r"(full) : "memory");
        }
        return;
    }
    const uint32_t* lut_lane = lut + lane;
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(counter, 1u);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (t >= items) break;
        const uint32_t slot = t % n_slots, round = t / n_slots;
        const CryptGroup g = crypt_group(a, blockIdx.x + (uint64_t)t * gridDim.x);
        uint8_t* dat = slots + (size_t)slot * slot_bytes + kLutFront;
        mbar_wait(smem_u32(bars + slot), round & 1);
        // ---- one lane per frame, in place
        if ((uint32_t)lane < g.cnt) {
            const uint32_t cipher = a.streams[g.stream].cipher;
            const uint32_t q0 = g.lead + (uint32_t)lane * g.fs;                        // frame start, bytes from `dat`
            if (cipher == a.lut_table) crypt_frame<true>(dat, q0, g.fs, lut_lane, nullptr);
            else crypt_frame<false>(dat, q0, g.fs, lut_lane, a.tables + (size_t)cipher * 256);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // the lanes' writes, before the bulk store reads them
        __syncwarp();
        // ---- out: interior rows in one bulk store, the group's first and last partial rows byte by byte
        const uint32_t end = g.lead + g.bytes;
        const uint32_t head_end = g.lead ? min(16u, end) : 0u;
        const uint32_t r0 = g.lead ? 1u : 0u;
        const uint32_t r1 = max(end >> 4, r0);
        if (lane == 0 && r1 > r0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(a.out + g.base + (uint64_t)r0 * 16), "r"(smem_u32(dat + r0 * 16)), "r"((r1 - r0) * 16) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        for (uint32_t k = g.lead + lane; k < head_end; k += 32) a.out[g.base + k] = dat[k];
        for (uint32_t k = max(r1 * 16, head_end) + lane; k < end; k += 32) a.out[g.base + k] = dat[k];
        __syncwarp();                                                   // every lane is done with the slot
        if (lane == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");          // ... and so is the store
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bars + kLutMaxSlots + slot)) : "memory");
        }
    }
}

}  // namespace

void launch_hca_crypt(const HcaCryptArgs& a, cudaStream_t s, uint64_t* launches) {
    if (!a.n_frames) return;
    static const bool force_staged = [] { const char* e = getenv("CRI_HCA_CRYPT_STAGED"); return e && *e && *e != '0'; }();
    if (a.n_groups && a.min_frame >= 128 && !force_staged) {
        const size_t fixed = kLutTableBytes + kLutCtrlBytes, slot = kLutFront + a.group_bytes;
        const size_t fit = (227 * 1024 - fixed) / slot;
        const uint32_t n_slots = (uint32_t)std::min<size_t>(fit, kLutMaxSlots);
        if (n_slots >= 2) {
            const size_t smem = fixed + n_slots * slot;
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            cudaFuncSetAttribute(hca_crypt_lut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            const uint64_t ctas = std::min<uint64_t>((a.n_groups + kLutConsumers - 1) / kLutConsumers, (uint64_t)sms);
            hca_crypt_lut_kernel<<<(unsigned)ctas, (kLutConsumers + 1) * 32, smem, s>>>(a, n_slots);
            ++*launches;
            return;
        }
    }
    if (a.n_groups) {
        const size_t smem = (size_t)kStageWarps * a.group_bytes;
        cudaFuncSetAttribute(hca_crypt_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        hca_crypt_staged_kernel<<<(unsigned)((a.n_groups + kStageWarps - 1) / kStageWarps), kStageWarps * 32, smem, s>>>(a);
        ++*launches;
        return;
    }
    hca_crypt_kernel<<<(unsigned)((a.n_frames + kCryptThreads - 1) / kCryptThreads), kCryptThreads, 0, s>>>(a);
    ++*launches;
}

}  // namespace cri
