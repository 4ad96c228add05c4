// HCA decode, second kernel: dequantise -> (HFR, intensity stereo) -> 128-point DCT-IV -> window + overlap-add ->
// PCM16 (clHCA_DecodeBlock_transform, CriCodecs/hca.cpp:1207-1233; imdct_transform :1898-1992;
// clHCA_ReadSamples16 :339-360).
//
// One WARP per unit (a run of consecutive frames of one stream). Coefficient p of a 128-point block lives in lane
// p>>2, register p&3 and never moves: the reference's 7 sum/difference passes pair slots that differ in bit 0..6
// of p, its 7 rotation passes pair bit 6..0 and the window pairs bit 0 (tools/gen_dct.py derives this and emits
// the per-slot rotation factors), so every exchange is a register swap or one __shfl_xor and the overlap state is
// two registers per lane. Input rows (256 B of int16 spectra + 512 B of gains per block) arrive through a
// double-buffered cp.async pipeline, so HBM latency is off the register scoreboard; PCM leaves through a shared
// tile as coalesced 128-byte stores. Every product and sum is rounded separately (__fmul_rn / __fadd_rn; the
// reference build has no FMA; subtractions are folded into the sign of a table constant, x - y == x + (-y)).
//
// Algorithmic bytes per stereo frame: 4096 (int16 spectra) + 1024 (gains) read, 4096 (PCM16) written = 9216.
#include <cstdint>

#include "cri_tables.h"
#include "hca_kernels.h"

namespace cri {
namespace {

__constant__ uint32_t c_intensity[16] = CRI_TBL_INTENSITY_RATIO;
__constant__ uint32_t c_noise_conv[128] = CRI_TBL_SCALE_CONV;

// The v3.0 noise generator is rand()'s LCG, state' = 0x343FD * state + 0x269EC3 (hca.cpp:1616), run once per
// resolution-0 band in stream order. n steps at once: compose the affine map with itself by squaring.
__device__ __forceinline__ uint32_t lcg_jump(uint32_t state, uint32_t n) {
    uint32_t mul = 1, add = 0, m = 0x343FDu, c = 0x269EC3u;
    while (n) {
        if (n & 1) { mul *= m; add = add * m + c; }
        c = c * m + c;
        m *= m;
        n >>= 1;
    }
    return mul * state + add;
}

#include "hca_dct_gen.inc"

constexpr int kWarps = 4;
constexpr unsigned kFull = 0xFFFFFFFFu;

__device__ __forceinline__ int pcm16(float f) {   // hca.cpp:339-360; (int) of an out-of-range float is INT_MIN on x86
    const float v = __fmul_rn(f, 32768.0f);
    int s = __float2int_rz(v);
    if (!(fabsf(v) < 2147483648.0f)) s = INT_MIN;
    return max(-32768, min(32767, s));
}
__device__ __forceinline__ float flip(float v, uint32_t mask) { return __uint_as_float(__float_as_uint(v) ^ mask); }

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_wait_all_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// The 128-point DCT-IV on the warp-distributed block, then window + overlap; `prev` holds the even slots of the
// previous block of this channel (dct[127-j] for the odd slots' dct[j], j >= 64). Returns the 4 wave samples of
// this lane (positions kWinPosA/B) through w[].
struct LaneConsts {
    float wa0, wa1, wb0, wb1;
    uint32_t sgn[5];
    const float4* rot_s;   // shared tables, [pass * 32 + lane]
    const float4* rot_c;
};

__device__ __forceinline__ void dct4_window(float (&x)[4], const LaneConsts& k, float2& prev, float (&w)[4]) {
    // 7 sum/difference passes: slot bit 0, 1 (registers), 2..6 (lanes)
    {
        const float t0 = __fadd_rn(x[0], x[1]), t1 = __fsub_rn(x[0], x[1]), t2 = __fadd_rn(x[2], x[3]), t3 = __fsub_rn(x[2], x[3]);
        x[0] = __fadd_rn(t0, t2); x[2] = __fsub_rn(t0, t2); x[1] = __fadd_rn(t1, t3); x[3] = __fsub_rn(t1, t3);
    }
#pragma unroll
    for (int b = 0; b < 5; b++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const float other = __shfl_xor_sync(kFull, x[r], 1 << b);
            x[r] = __fadd_rn(other, flip(x[r], k.sgn[b]));
        }
    }
    // 7 rotation passes: slot bit 6..2 (lanes), 1, 0 (registers):  v*S + partner*C
#pragma unroll
    for (int st = 0; st < 5; st++) {
        const float4 s4 = k.rot_s[st * 32], c4 = k.rot_c[st * 32];
        const float ss[4] = {s4.x, s4.y, s4.z, s4.w}, cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const float other = __shfl_xor_sync(kFull, x[r], 16 >> st);
            x[r] = __fadd_rn(__fmul_rn(x[r], ss[r]), __fmul_rn(other, cc[r]));
        }
    }
    {
        const float4 s5 = k.rot_s[5 * 32], c5 = k.rot_c[5 * 32], s6 = k.rot_s[6 * 32], c6 = k.rot_c[6 * 32];
        const float y0 = __fadd_rn(__fmul_rn(x[0], s5.x), __fmul_rn(x[2], c5.x));
        const float y2 = __fadd_rn(__fmul_rn(x[2], s5.z), __fmul_rn(x[0], c5.z));
        const float y1 = __fadd_rn(__fmul_rn(x[1], s5.y), __fmul_rn(x[3], c5.y));
        const float y3 = __fadd_rn(__fmul_rn(x[3], s5.w), __fmul_rn(x[1], c5.w));
        x[0] = __fadd_rn(__fmul_rn(y0, s6.x), __fmul_rn(y1, c6.x));
        x[1] = __fadd_rn(__fmul_rn(y1, s6.y), __fmul_rn(y0, c6.y));
        x[2] = __fadd_rn(__fmul_rn(y2, s6.z), __fmul_rn(y3, c6.z));
        x[3] = __fadd_rn(__fmul_rn(y3, s6.w), __fmul_rn(y2, c6.w));
    }
    // window + overlap (hca.cpp:1983-1992): odd slots hold dct[j >= 64], even slots dct[127 - j]
    w[0] = __fadd_rn(__fmul_rn(k.wa0, x[1]), __fmul_rn(k.wb0, prev.x));
    w[1] = __fsub_rn(__fmul_rn(k.wb0, x[1]), __fmul_rn(k.wa0, prev.x));
    w[2] = __fadd_rn(__fmul_rn(k.wa1, x[3]), __fmul_rn(k.wb1, prev.y));
    w[3] = __fsub_rn(__fmul_rn(k.wb1, x[3]), __fmul_rn(k.wa1, prev.y));
    prev = make_float2(x[0], x[2]);
}

// NCH > 0: every stream of the job has NCH channels, all 128 bands coded, no HFR / intensity (quality High and
// Highest): no masks, no joint-stereo code, overlap state in registers. NCH == 0: anything.
template <int NCH>
__global__ void __launch_bounds__(kWarps * 32)
hca_imdct_kernel(HcaDecodeArgs a) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    __shared__ __align__(16) float s_rot_s[7 * 128];
    __shared__ __align__(16) float s_rot_c[7 * 128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t unit = blockIdx.x * kWarps + warp;
    const bool dead = unit >= a.n_units || a.units[unit < a.n_units ? unit : 0].count == 0;
    for (int i = threadIdx.x; i < 7 * 128; i += blockDim.x) {
        s_rot_s[i] = __uint_as_float(kRotS[i]);
        s_rot_c[i] = __uint_as_float(kRotC[i]);
    }
    __syncthreads();
    if (dead) return;
    const HcaUnit u = a.units[unit];
    const HcaStreamDev& S = a.streams[u.stream];
    const int nch = NCH ? NCH : (int)S.channels;
    const int MC = (int)a.max_channels;
    // per-warp shared: 2 input stages (256 B spectra + 512 B gains), PCM tile [MC][128] int16, then (general path)
    // overlap carry [MC][32] float2, HFR / noise scratch [128] float, and for the noise generator the band of every
    // valid rank [128] and the scalefactors [128]
    const size_t per_warp = 2 * 768 + (size_t)MC * 256 + (NCH ? 0 : (size_t)MC * 256 + 512 + 256);
    uint8_t* base = s_dyn + (size_t)warp * per_warp;
    uint8_t* stage = base;
    int16_t* tile = reinterpret_cast<int16_t*>(base + 2 * 768);
    float2* carry = reinterpret_cast<float2*>(base + 2 * 768 + MC * 256);
    float* xs = reinterpret_cast<float*>(base + 2 * 768 + MC * 512);
    uint8_t* vmap = base + 2 * 768 + MC * 512 + 512;
    uint8_t* sfb = vmap + 128;

    LaneConsts k;
    k.wa0 = __uint_as_float(kWinA[2 * lane]); k.wa1 = __uint_as_float(kWinA[2 * lane + 1]);
    k.wb0 = __uint_as_float(kWinB[2 * lane]); k.wb1 = __uint_as_float(kWinB[2 * lane + 1]);
#pragma unroll
    for (int b = 0; b < 5; b++) k.sgn[b] = (lane >> b) & 1 ? 0x80000000u : 0u;
    k.rot_s = reinterpret_cast<const float4*>(s_rot_s) + lane;
    k.rot_c = reinterpret_cast<const float4*>(s_rot_c) + lane;
    const int pa0 = kWinPosA[2 * lane], pa1 = kWinPosA[2 * lane + 1], pb0 = kWinPosB[2 * lane], pb1 = kWinPosB[2 * lane + 1];

    float2 prev_reg[NCH ? NCH : 1];
#pragma unroll
    for (int c = 0; c < (NCH ? NCH : 1); c++) prev_reg[c] = make_float2(0.f, 0.f);
    if (!NCH)
        for (int c = 0; c < nch; c++) carry[c * 32 + lane] = make_float2(0.f, 0.f);
    const int total = S.total_bands, basebands = S.base_bands;
    const int start = S.base_bands + S.stereo_bands;
    // HFR band start + n copies low band start - 1 - min(n, lim): v3.0 stops moving down after half of the groups
    // (hca.cpp:1652-1676); the copying ends at the top band, after the last group, or when the low band would pass 0
    const int lim = (S.v3 ? (int)S.hfr_groups >> 1 : (int)S.hfr_groups) * (int)S.bands_per_hfr;
    int room = min(total - start, (int)S.hfr_groups * (int)S.bands_per_hfr);
    if (lim >= start) room = min(room, start);
    const bool joint = !NCH && S.joint;
    const bool noise_on = !NCH && S.noise && a.sfres != nullptr;
    const uint64_t slot0 = (uint64_t)unit * a.steps;

    // block index -> (step, subframe, channel); a unit has 1 look-back block set (subframe 7 of the frame in front
    // of the run) unless it starts the stream, then count * 8 * nch blocks
    const uint32_t first_step = u.first == 0 ? 1u : 0u;
    auto request = [&](uint32_t step, int sub, int c, int buf) {
        const uint64_t sc = (slot0 + step) * MC + c;
        uint8_t* dst = stage + buf * 768;
        cp_async8(dst + lane * 8, reinterpret_cast<const uint8_t*>(a.quant + (sc * 8 + sub) * 16) + lane * 8);
        cp_async16(dst + 256 + lane * 16, reinterpret_cast<const uint8_t*>(a.gain + sc * 128) + lane * 16);
    };
    int buf = 0;
    request(first_step, first_step == 0 ? 7 : 0, 0, 0);
    cp_commit();

    for (uint32_t step = first_step; step <= u.count; step++) {
        const uint32_t frame = u.first + step - 1;
        const uint64_t slot = slot0 + step;
        for (int sub = (step == 0 ? 7 : 0); sub < 8; sub++) {
            float xl[4] = {0.f, 0.f, 0.f, 0.f};   // primary channel's spectra for intensity stereo
#pragma unroll
            for (int c = 0; c < nch; c++) {
                {   // request the next block's rows, then wait for this block's
                    int nc = c + 1, ns = sub;
                    uint32_t nstep = step;
                    if (nc == nch) { nc = 0; ns++; }
                    if (ns == 8) { ns = 0; nstep++; }
                    if (nstep <= u.count) request(nstep, ns, nc, buf ^ 1);
                    cp_commit();
                    cp_wait_all_but_one();
                }
                const uint2 q2 = *reinterpret_cast<const uint2*>(stage + buf * 768 + lane * 8);
                const float4 g4 = *reinterpret_cast<const float4*>(stage + buf * 768 + 256 + lane * 16);
                buf ^= 1;
                // ---- dequantise: spectra = gain * q  (hca.cpp:1568); bands past the coded count are zero
                float x[4];
                x[0] = __fmul_rn(g4.x, (float)(int)(short)(q2.x & 0xFFFF));
                x[1] = __fmul_rn(g4.y, (float)((int)q2.x >> 16));
                x[2] = __fmul_rn(g4.z, (float)(int)(short)(q2.y & 0xFFFF));
                x[3] = __fmul_rn(g4.w, (float)((int)q2.y >> 16));
                if (!NCH) {
                    const int coded = S.coded[c];
                    const int type = S.type[c];
#pragma unroll
                    for (int r = 0; r < 4; r++)
                        if (4 * lane + r >= coded) x[r] = 0.f;
                    if (noise_on) {
                        // ---- noise fill (hca.cpp:1602-1635): resolution-0 band number i of this channel takes a
                        // random VALID band's value, rescaled by the difference of their scalefactors; draw i of
                        // the channel is draw (subframes before) + (channels before) + i + 1 of the frame
                        const uint32_t* dr = a.draws + slot * MC;
                        const uint32_t mine = dr[c];
                        if (mine) {
                            uint32_t all = 0, before = 0;
                            for (int cc = 0; cc < nch; cc++) {
                                const uint32_t d = dr[cc];
                                all += d;
                                if (cc < c) before += d;
                            }
                            const uint32_t state0 = lcg_jump(a.frame_state[S.frame_base + frame], (uint32_t)sub * all + before);
                            const uint32_t cls4 = reinterpret_cast<const uint32_t*>(a.sfres + (slot * MC + c) * 128)[lane];
                            const uint32_t below = (1u << lane) - 1;
                            int n_rank = 0, v_rank = 0, valid = 0;
#pragma unroll
                            for (int r = 0; r < 4; r++) {
                                const uint32_t nm = __ballot_sync(kFull, (cls4 >> (8 * r)) & 0x40);
                                const uint32_t vm = __ballot_sync(kFull, (cls4 >> (8 * r)) & 0x80);
                                n_rank += __popc(nm & below);
                                v_rank += __popc(vm & below);
                                valid += __popc(vm);
                            }
                            valid = max(valid, 1);           // (a channel that draws has valid bands; keeps bad data in bounds)
                            __syncwarp();
                            reinterpret_cast<float4*>(xs)[lane] = make_float4(x[0], x[1], x[2], x[3]);
#pragma unroll
                            for (int r = 0; r < 4; r++) {
                                const uint32_t b = (cls4 >> (8 * r)) & 0xFF;
                                sfb[4 * lane + r] = (uint8_t)(b & 0x3F);
                                if (b & 0x80) vmap[v_rank++] = (uint8_t)(4 * lane + r);
                            }
                            __syncwarp();
#pragma unroll
                            for (int r = 0; r < 4; r++) {
                                const uint32_t b = (cls4 >> (8 * r)) & 0xFF;
                                if (b & 0x40) {
                                    const uint32_t rnd = lcg_jump(state0, (uint32_t)(++n_rank));
                                    const int pick = valid - 1 - (int)(((rnd & 0x7FFF) * (uint32_t)valid) >> 15);
                                    const int vi = vmap[pick] & 127;
                                    int k = (int)(b & 0x3F) - (int)sfb[vi] + 62;
                                    k &= ~(k >> 31);
                                    x[r] = __fmul_rn(__uint_as_float(c_noise_conv[k]), xs[vi]);
                                }
                            }
                        }
                    }
                    if (joint) {
                        // ---- HFR: mirrored low bands scaled into the high bands (hca.cpp:1638-1683)
                        if (S.bands_per_hfr && type != 2) {
                            __syncwarp();
                            reinterpret_cast<float4*>(xs)[lane] = make_float4(x[0], x[1], x[2], x[3]);
                            __syncwarp();
                            const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                            for (int r = 0; r < 4; r++) {
                                const int p = 4 * lane + r;
                                if (p >= start && p < start + room) x[r] = __fmul_rn(gg[r], xs[start - 1 - min(p - start, lim)]);
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
                }
                float w[4];
                if (NCH) {
                    dct4_window(x, k, prev_reg[NCH ? c : 0], w);
                } else {
                    float2 prev = carry[c * 32 + lane];
                    dct4_window(x, k, prev, w);
                    carry[c * 32 + lane] = prev;
                }
                if (step != 0) {
                    int16_t* t = tile + c * 128;
                    t[pa0] = (int16_t)pcm16(w[0]);
                    t[pb0] = (int16_t)pcm16(w[1]);
                    t[pa1] = (int16_t)pcm16(w[2]);
                    t[pb1] = (int16_t)pcm16(w[3]);
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
                        const uint32_t v = (uint32_t)(uint16_t)tile[i] | ((uint32_t)(uint16_t)tile[128 + i] << 16);
                        *reinterpret_cast<uint32_t*>(dst + n * 4) = v;
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

// Generator state at the start of every frame: one thread walks one stream's per-frame draw counts.
__global__ void hca_noise_scan_kernel(HcaDecodeArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_streams) return;
    const HcaStreamDev& S = a.streams[i];
    if (!S.noise || a.status[i] != 0) return;
    const uint64_t need = ((uint64_t)S.out_samples + S.delay + 1023) / 1024;          // as plan_hca_decode
    const uint32_t frames = (uint32_t)(need < S.frame_count ? need : S.frame_count);
    uint32_t state = 1;                                                                // HCA_DEFAULT_RANDOM, hca.cpp:62
    for (uint32_t f = 0; f < frames; f++) {
        a.frame_state[S.frame_base + f] = state;
        state = lcg_jump(state, 8u * a.frame_draws[S.frame_base + f]);
    }
}

// Intensities a frame did not code keep the previous frame's values (unpack_intensity leaves its channel state alone:
// v <= 2.0 with a first index of 15, hca.cpp:1368-1372; v3.0 with a delta that leaves 0..15, :1410-1412 + :1185). The
// unpack kernels mark such frames (carry = first kept nibble); here one warp looks at one stream and, only if it has a
// marked frame, lane 0 walks its frames in order. General kernels: a frame's slot is unit * steps + step, and the last
// frame of a unit is also the look-back frame (step 0) of the next one.
__global__ void hca_intensity_scan_kernel(HcaDecodeArgs a) {
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= a.n_streams) return;
    const HcaStreamDev& S = a.streams[i];
    const bool fast = a.n_runs != 0;
    const uint32_t frames = fast ? a.dec_prefix[i + 1] - a.dec_prefix[i]
                                 : (uint32_t)min((uint64_t)S.frame_count, ((uint64_t)S.out_samples + S.delay + 1023) / 1024);
    const uint32_t run = a.steps - 1, MC = a.max_channels;
    for (uint32_t c = 1; c < S.channels; c++) {
        if (S.type[c] != 2) continue;
        auto slot_of = [&](uint32_t f) -> uint64_t {
            if (fast) return (uint64_t)a.dec_prefix[i] + f;
            return ((uint64_t)(S.unit_base + f / run) * a.steps + f % run + 1) * MC + c;
        };
        bool any = false;
        for (uint32_t f = lane; f < frames; f += 32) any = any || a.carry[slot_of(f)] != 0;
        if (!__any_sync(0xFFFFFFFFu, any) || lane != 0) continue;
        uint32_t prev = 0;                                                   // channel state starts zeroed (hca.cpp:962)
        for (uint32_t f = 0; f < frames; f++) {
            const uint64_t at = slot_of(f);
            const uint32_t k = a.carry[at];
            uint32_t v = a.inten[at];
            if (k) {
                const uint32_t kept = 0xFFFFFFFFu << (4 * k);
                v = (v & ~kept) | (prev & kept);
                a.inten[at] = v;
            }
            if (!fast && f % run == run - 1 && f + 1 < frames) a.inten[((uint64_t)(S.unit_base + f / run + 1) * a.steps) * MC + c] = v;
            prev = v;
        }
    }
}

template <int NCH>
void launch_one(const HcaDecodeArgs& a, cudaStream_t s) {
    const size_t per_warp = 2 * 768 + (size_t)a.max_channels * 256 + (NCH ? 0 : (size_t)a.max_channels * 256 + 512 + 256);
    const size_t smem = per_warp * kWarps;
    cudaFuncSetAttribute(hca_imdct_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    hca_imdct_kernel<NCH><<<(a.n_units + kWarps - 1) / kWarps, kWarps * 32, smem, s>>>(a);
}

}  // namespace

void launch_hca_noise_scan(const HcaDecodeArgs& a, cudaStream_t s, uint64_t* launches) {
    if (!a.sfres || !a.n_streams) return;
    hca_noise_scan_kernel<<<(a.n_streams + 127) / 128, 128, 0, s>>>(a);
    ++*launches;
}

void launch_hca_intensity_scan(const HcaDecodeArgs& a, cudaStream_t s, uint64_t* launches) {
    if (!a.carry_scan || !a.n_streams) return;
    hca_intensity_scan_kernel<<<(a.n_streams + 3) / 4, 128, 0, s>>>(a);
    ++*launches;
}

void launch_hca_imdct(const HcaDecodeArgs& a, cudaStream_t s, uint64_t* launches) {
    if (!a.n_units) return;
    if (a.uniform == 2) launch_one<2>(a, s);
    else if (a.uniform == 1) launch_one<1>(a, s);
    else launch_one<0>(a, s);
    ++*launches;
}

}  // namespace cri
