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
#include <cstdint>

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

}  // namespace

void launch_hca_crypt(const HcaCryptArgs& a, cudaStream_t s, uint64_t* launches) {
    if (!a.n_frames) return;
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
