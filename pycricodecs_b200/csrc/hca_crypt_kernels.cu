// HCA frame en/decryption kernel for sm_100a.
//
// Reference: HcaCrypt (CriCodecs/hca.cpp:3271-3337) maps every byte of every
// frame through a 256-entry substitution table derived from the key
// (cipher_decrypt, :491-497; the inverse table when encrypting, :3315-3320) and
// rewrites the frame's trailing CRC16 over the first frame_size-2 mapped bytes
// (:3326). The header half (signature masks, ciph type, header CRC) is done on
// the host (formats.cpp: crypt_header) and scattered as a patch.
//
// One lane per frame: the CRC is a serial byte recurrence, frames are
// independent, the batch has hundreds of thousands of them. Input and output
// share their byte offsets, so a frame's 4-byte-aligned interior moves as whole
// words (head/tail bytes separately); consecutive loads of a lane fall into the
// same 128-byte line, so HBM traffic is the compulsory 2 x frame_size.
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

}  // namespace

void launch_hca_crypt(const HcaCryptArgs& a, cudaStream_t s, uint64_t* launches) {
    if (!a.n_frames) return;
    hca_crypt_kernel<<<(unsigned)((a.n_frames + kCryptThreads - 1) / kCryptThreads), kCryptThreads, 0, s>>>(a);
    ++*launches;
}

}  // namespace cri
