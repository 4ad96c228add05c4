// HCA job planning (host) and kernel sequencing. Header parsing is formats.cpp;
// the kernels are hca_dec_kernels.cu / hca_crypt_kernels.cu / hca_enc_kernels.cu.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>

#include "engine.h"
#include "hca_kernels.h"

namespace cri {

#define CU_TRY(ctx, expr)                                                      \
    do {                                                                       \
        cudaError_t e_ = (expr);                                               \
        if (e_ != cudaSuccess) {                                               \
            (ctx)->error = std::string(#expr) + ": " + cudaGetErrorString(e_); \
            return ERR_CUDA;                                                   \
        }                                                                      \
    } while (0)

// one table step of the CRC register with byte b appended (formats.cpp keeps the table)
static uint16_t crc16_append(uint16_t crc, uint8_t b) {
    const unsigned v = ((crc >> 8) ^ b) & 0xFF;
    unsigned t = (v << 1) ^ (v << 2);
    unsigned par = v; par ^= par >> 4; par ^= par >> 2; par ^= par >> 1;
    if (par & 1) t ^= 0x8003;
    return (uint16_t)(((crc << 8) ^ t) & 0xFFFF);
}

// The bit readers keep advancing (and prefetching) when a corrupt frame's codes run past its end: the values they return are
// forced to zero, but the loads go on for at most 16 channels x 1024 codes x 12 bits = 24 KB behind the row. Inside the
// scratch array that lands in other rows; behind the last row it lands in this tail.
constexpr uint64_t kScratchTail = 32 * 1024;

static uint32_t env_u32(const char* name, uint32_t fallback) {
    const char* v = getenv(name);
    if (!v || !*v) return fallback;
    const long x = strtol(v, nullptr, 10);
    return x > 0 ? (uint32_t)x : fallback;
}

// cipher tables are shared between streams that use the same (type, key)
struct CipherPool {
    std::map<std::pair<int, uint64_t>, uint32_t> index;
    std::vector<uint8_t>* bytes;
    explicit CipherPool(std::vector<uint8_t>* b) : bytes(b) {
        bytes->resize(256);
        for (int i = 0; i < 256; i++) (*bytes)[i] = (uint8_t)i;
    }
    uint32_t get(int type, uint64_t key, bool invert = false) {
        if (type == 56 && key == 0) type = 0;
        if (type == 0) return 0;
        const auto k = std::make_pair(invert ? -type : type, key);
        auto it = index.find(k);
        if (it != index.end()) return it->second;
        uint8_t t[256], inv[256];
        cipher_table(type, key, t);
        if (invert) {
            for (int i = 0; i < 256; i++) inv[t[i]] = (uint8_t)i;
            memcpy(t, inv, 256);
        }
        const uint32_t id = (uint32_t)(bytes->size() / 256);
        bytes->insert(bytes->end(), t, t + 256);
        index[k] = id;
        return id;
    }
};

static void fill_stream(HcaStreamDev* s, const HcaInfo& h) {
    memset(s, 0, sizeof *s);
    s->frame_size = h.frame_size;
    s->frame_count = h.frame_count;
    s->delay = h.delay;
    s->channels = (uint8_t)h.channels;
    s->total_bands = (uint8_t)h.total_bands;
    s->base_bands = (uint8_t)h.base_bands;
    s->stereo_bands = (uint8_t)h.stereo_bands;
    s->bands_per_hfr = (uint8_t)h.bands_per_hfr;
    s->hfr_groups = (uint8_t)h.hfr_groups;
    s->min_res = (uint8_t)h.min_res;
    s->max_res = (uint8_t)h.max_res;
    bool joint = h.bands_per_hfr != 0;
    for (unsigned c = 0; c < h.channels; c++) {
        s->type[c] = h.type[c];
        s->coded[c] = (uint8_t)h.coded[c];
        if (h.type[c]) joint = true;
    }
    s->joint = joint ? 1 : 0;
    s->v3 = h.version > 0x0200 ? 1 : 0;
    s->noise = s->v3 && h.min_res == 0 ? 1 : 0;
}

// v3.0 reads hfr_groups extra scalefactors per non-secondary channel and copies them to the top of the table starting
// one entry past the last one read (hca.cpp:1353-1355). That entry is always zero as long as the copies cannot reach
// it, i.e. coded + 2 * hfr_groups < 128; otherwise it carries a value over from the previous frame, which a
// frame-parallel decoder cannot reproduce.
static bool v3_supported(const HcaInfo& h) {
    if (h.version <= 0x0200 || h.hfr_groups == 0) return true;
    for (unsigned c = 0; c < h.channels; c++)
        if (h.type[c] != 2 && h.coded[c] + 2 * h.hfr_groups >= 128) return false;
    return true;
}

int hca_decode_size_one(const uint8_t* d, uint64_t len, HcaInfo* h, uint64_t* size) {
    *size = 0;
    if (parse_hca(d, len, h) != OK) return ERR_HCA_HEADER;
    if ((uint64_t)h->header_size + (uint64_t)h->frame_count * h->frame_size > len) return ERR_HCA_HEADER;   // frames missing
    const uint64_t total = (uint64_t)h->frame_count * 1024;
    if (total < (uint64_t)h->delay + h->padding) return ERR_HCA_HEADER;
    *size = wav_header_size(h->loop_flag) + (total - h->delay - h->padding) * h->channels * 2;
    return OK;
}

int plan_hca_decode(cri_ctx* c, cri_job* j) {
    (void)c;
    HcaJob& J = j->hca;
    std::vector<uint64_t> sizes(j->n, 0);
    std::vector<HcaInfo> infos(j->n);
    std::vector<uint8_t> unsupported(j->n, 0);
    parallel_for(j->n, [&](uint32_t i) {
        const uint8_t* d = j->blob + j->in_off[i];
        const uint64_t len = j->in_off[i + 1] - j->in_off[i];
        HcaInfo& h = infos[i];
        if ((j->status[i] = hca_decode_size_one(d, len, &h, &sizes[i])) != OK) return;
        if (!v3_supported(h)) { j->status[i] = ERR_UNSUPPORTED; unsupported[i] = 1; }              // region stays zero
    });
    for (uint32_t i = 0; i < j->n; i++)
        if (unsupported[i]) j->needs_clear = true;
    finish_layout_public(j, sizes);

    CipherPool pool(&J.cipher_tables);
    J.ath_tables.assign(128, 0);
    const uint32_t run = std::max(1u, env_u32("CRI_HCA_RUN", 16));
    J.max_channels = 1;
    J.streams.assign(j->n, HcaStreamDev{});   // one entry per input stream: the device status array shares the index
    std::vector<uint32_t> needed_frames(j->n, 0);
    bool any_v3 = false, any_noise = false;
    for (uint32_t i = 0; i < j->n; i++) {
        if (j->status[i] != OK) continue;
        const HcaInfo& h = infos[i];
        HcaStreamDev s;
        fill_stream(&s, h);
        const uint32_t samples = h.frame_count * 1024u - h.delay - h.padding;
        const size_t hdr = wav_header_size(h.loop_flag);
        uint8_t hb[0x70];
        const uint32_t ls = h.loop_start_frame * 1024u + h.loop_start_delay - h.delay;
        const uint32_t le = h.loop_end_frame * 1024u + (1024u - h.loop_end_padding) - h.delay;
        write_wav_header(hb, samples, (int)h.channels, (int)h.rate, h.loop_flag, ls, le);
        add_patch_public(j, j->out_off[i], hb, (uint32_t)hdr);
        s.in_off = j->in_off[i] + h.header_size;
        s.out_off = j->out_off[i] + hdr;
        s.out_samples = samples;
        const uint64_t key = mix_subkey(j->keys.empty() ? 0 : j->keys[i], j->subkeys.empty() ? 0 : j->subkeys[i]);
        s.cipher = pool.get((int)h.ciph_type, key);
        if (h.ath_type == 1) {
            s.ath = (uint32_t)(J.ath_tables.size() / 128);
            J.ath_tables.insert(J.ath_tables.end(), h.ath, h.ath + 128);
        }
        J.streams[i] = s;
        J.max_channels = std::max<uint32_t>(J.max_channels, h.channels);
        // the reference stops decoding once it has produced every output sample (hca.cpp:3401)
        needed_frames[i] = (uint32_t)std::min<uint64_t>(h.frame_count, ((uint64_t)samples + h.delay + 1023) / 1024);
        J.streams[i].frame_base = (uint32_t)j->units;
        any_v3 = any_v3 || s.v3;
        for (unsigned ch = 1; ch < h.channels; ch++) J.any_pair = J.any_pair || s.type[ch] == 2;
        any_noise = any_noise || s.noise;
        j->units += needed_frames[i];
    }
    // uniform batch: every decodable stream is mono (1) or stereo (2): discrete channels or one primary/secondary pair
    J.uniform = J.max_channels <= 2 ? J.max_channels : 0;
    bool any_joint = false;
    for (uint32_t i = 0; i < j->n && J.uniform; i++) {
        if (j->status[i] != OK) continue;
        const HcaStreamDev& st = J.streams[i];
        const bool types_ok = st.channels == 1 ? st.type[0] == 0
                                               : (st.type[0] == 0 && st.type[1] == 0) || (st.type[0] == 1 && st.type[1] == 2);
        if (st.channels != J.uniform || !types_ok || st.hfr_groups > 10) J.uniform = 0;   // the fast unpack kernel keeps <= 10 HFR scales (6 bits each) in one word
        any_joint = any_joint || st.joint;
    }
    J.any_joint = any_joint;
    uint32_t max_frame = 8;
    for (const auto& st : J.streams) max_frame = std::max(max_frame, st.frame_size);
    J.scratch_words = ((max_frame + 15) / 16 + 4) * 4;             // whole 16-byte rows + zeroed slack rows for the prefetching reader
    J.n_bytes = 0;
    J.noise_frames = 0;
    if (any_v3) J.uniform = 0;        // v3.0 streams take the general kernels
    if (J.uniform && j->units > 0 && j->units < 0xFFFF0000ull && env_u32("CRI_HCA_GENERAL", 0) == 0) {
        // fast path: flattened frame list cut into runs of run_len frames, one transform lane per (run, channel).
        // run_len is chosen so that the transform kernel's CTAs fill the resident slots of the GPU in whole waves.
        const uint64_t G = j->units;
        const uint64_t cols_per_cta = hca_fast_threads_per_cta() / J.uniform;        // runs per CTA
        const uint64_t slots = (uint64_t)std::max(1, c->sm_count) * hca_fast_ctas_per_sm() * hca_fast_chunks();   // pipelined: one wave per chunk
        uint32_t best = 1;
        const uint32_t forced = env_u32("CRI_HCA_FAST_RUN", 0);
        if (forced) {
            best = forced;
        } else {
            const uint64_t one_wave = (G + cols_per_cta * slots - 1) / (cols_per_cta * slots);   // run_len that fits one wave
            if (one_wave <= 96) {
                best = (uint32_t)std::max<uint64_t>(1, one_wave);
            } else {
                double best_eff = 0.0;
                for (uint32_t R = 12; R <= 32; R++) {
                    const uint64_t runs = (G + R - 1) / R, ctas = (runs + cols_per_cta - 1) / cols_per_cta;
                    const uint64_t waves = (ctas + slots - 1) / slots;
                    const double eff = (double)ctas / (double)(waves * slots) * (1.0 - 1.0 / (8.0 * R));
                    if (eff > best_eff) { best_eff = eff; best = R; }
                }
            }
        }
        J.run_len = best;
        J.n_runs = (uint32_t)((G + best - 1) / best);
        J.total_frames = G;
        J.dec_prefix.assign(j->n + 1, 0);
        for (uint32_t i = 0; i < j->n; i++) J.dec_prefix[i + 1] = J.dec_prefix[i] + needed_frames[i];
        const uint64_t runs_per_warp = 32 / J.uniform;
        const uint64_t warps = (J.n_runs + runs_per_warp - 1) / runs_per_warp;
        J.spec_bytes = warps * J.run_len * 8 * 1024 * sizeof(float4);
        J.s_bytes = (G + 1) * J.scratch_words * sizeof(uint32_t) + kScratchTail;
        J.i_count = G + 1;                                           // intensity nibbles of the frame's secondary channel,
        J.i_bytes = J.i_count * (sizeof(uint32_t) + 1);              // then one "kept from the previous frame" byte each
        J.total_groups = 0;
        J.max_steps = 0;
        return OK;
    }
    for (uint32_t i = 0; i < j->n; i++) {
        if (j->status[i] != OK) continue;
        J.streams[i].unit_base = (uint32_t)J.units.size();
        for (uint32_t f = 0; f < needed_frames[i]; f += run) {
            HcaUnit u{i, f, std::min(run, needed_frames[i] - f)};
            J.units.push_back(u);
        }
    }
    while (J.units.size() % 32) J.units.push_back(HcaUnit{0, 0, 0});
    // general transform kernel's shortcut: no joint tools and all 128 bands coded in every channel
    if (any_joint) J.uniform = 0;
    for (uint32_t i = 0; i < j->n && J.uniform; i++) {
        if (j->status[i] != OK) continue;
        const HcaStreamDev& st = J.streams[i];
        for (unsigned ch = 0; ch < st.channels; ch++)
            if (st.coded[ch] != 128) J.uniform = 0;
    }
    J.max_steps = run + 1;
    J.total_groups = (J.units.size() / 32) * J.max_steps;
    const uint64_t slots = (uint64_t)J.units.size() * J.max_steps;
    J.s_bytes = slots * J.scratch_words * sizeof(uint32_t) + kScratchTail;
    J.q_bytes = slots * J.max_channels * 8 * 16 * sizeof(uint4);
    J.g_bytes = slots * J.max_channels * 128 * sizeof(float);
    J.i_count = slots * J.max_channels;
    J.i_bytes = J.i_count * (sizeof(uint32_t) + 1);
    if (any_noise && j->units < 0xFFFF0000ull) {
        J.noise_frames = j->units;
        J.n_bytes = slots * J.max_channels * (128 + sizeof(uint32_t)) + 2 * J.noise_frames * sizeof(uint32_t);
    }
    return OK;
}

// HcaCrypt (hca.cpp:3271-3337): output has the input's size and layout; frames come from the kernel, the
// rewritten header (and any bytes outside header + frames) from host-built patches.
int plan_hca_crypt(cri_ctx* c, cri_job* j) {
    (void)c;
    HcaJob& J = j->hca;
    std::vector<uint64_t> sizes(j->n, 0);
    CipherPool pool(&J.cipher_tables);
    J.streams.assign(j->n, HcaStreamDev{});
    J.frame_prefix.assign(j->n + 1, 0);
    std::vector<uint8_t> hdr;
    std::vector<HcaInfo> infos(j->n);
    std::vector<uint8_t> parsed(j->n, 0);
    parallel_for(j->n, [&](uint32_t i) {
        const uint64_t len = j->in_off[i + 1] - j->in_off[i];
        HcaInfo& h = infos[i];
        parsed[i] = parse_hca(j->blob + j->in_off[i], len, &h) == OK && (uint64_t)h.header_size + (uint64_t)h.frame_count * h.frame_size <= len;
    });
    for (uint32_t i = 0; i < j->n; i++) {
        const uint64_t len = j->in_off[i + 1] - j->in_off[i];
        sizes[i] = len;
        const HcaInfo& h = infos[i];
        uint64_t frames = 0;
        if (!parsed[i]) {
            j->status[i] = ERR_HCA_HEADER;
        } else {
            const unsigned type = j->encrypt ? j->ciph_type : h.ciph_type;   // hca.cpp:3307
            if (type != 0 && type != 1 && type != 56) {
                j->status[i] = ERR_HCA_HEADER;
            } else {
                const uint64_t key = mix_subkey(j->keys.empty() ? 0 : j->keys[i], j->subkeys.empty() ? 0 : j->subkeys[i]);
                HcaStreamDev s{};
                s.frame_size = h.frame_size;
                s.frame_count = h.frame_count;
                s.in_off = j->in_off[i] + h.header_size;
                s.cipher = pool.get((int)type, key, j->encrypt != 0);
                J.streams[i] = s;
                frames = h.frame_count;
                j->units += frames;
            }
        }
        J.frame_prefix[i + 1] = J.frame_prefix[i] + frames;
    }
    finish_layout_public(j, sizes);   // out_off == in_off
    {   // staged kernel: a warp stages up to 32 consecutive frames (<= 24 KB) of one stream in shared memory
        uint32_t max_fs = 0, min_fs = 0xFFFFFFFFu;
        std::map<uint32_t, uint64_t> table_use;                       // frames per cipher table
        for (const auto& st : J.streams) {
            max_fs = std::max(max_fs, st.frame_size);
            if (st.frame_count) {
                min_fs = std::min(min_fs, st.frame_size);
                table_use[st.cipher] += st.frame_count;
            }
        }
        J.min_frame = min_fs == 0xFFFFFFFFu ? 0 : min_fs;
        J.lut_table = 0;
        uint64_t most = 0;
        for (const auto& kv : table_use)
            if (kv.second > most) { most = kv.second; J.lut_table = kv.first; }
        J.frames_per_group = 0;
        if (max_fs >= 8 && max_fs <= 12288) {
            J.frames_per_group = std::min(32u, std::max(1u, 24576u / max_fs));
            J.group_bytes = ((J.frames_per_group * max_fs + 32 + 15) / 16) * 16;
            J.group_prefix.assign(j->n + 1, 0);
            for (uint32_t i = 0; i < j->n; i++)
                J.group_prefix[i + 1] = J.group_prefix[i] + (J.streams[i].frame_count + J.frames_per_group - 1) / J.frames_per_group;
            J.group_table.reserve(J.group_prefix[j->n]);
            for (uint32_t i = 0; i < j->n; i++)
                for (uint32_t f = 0; f < J.streams[i].frame_count; f += J.frames_per_group) J.group_table.push_back(make_uint2(i, f));
        }
    }
    // the rewritten headers (CryptHeader, hca.cpp:3166-3250: cipher type, CRC) side by side in one buffer, then the patches
    std::vector<uint64_t> hdr_at(j->n + 1, 0);
    for (uint32_t i = 0; i < j->n; i++) hdr_at[i + 1] = hdr_at[i] + (j->status[i] == OK ? (unsigned)be16(j->blob + j->in_off[i] + 6) : 0u);
    hdr.resize(hdr_at[j->n]);
    parallel_for(j->n, [&](uint32_t i) {
        const unsigned hs = (unsigned)(hdr_at[i + 1] - hdr_at[i]);
        if (!hs) return;
        memcpy(hdr.data() + hdr_at[i], j->blob + j->in_off[i], hs);
        crypt_header(hdr.data() + hdr_at[i], hs, j->encrypt ? j->ciph_type : 0);
    });
    for (uint32_t i = 0; i < j->n; i++) {
        if (j->status[i] != OK) continue;
        const uint8_t* d = j->blob + j->in_off[i];
        const uint64_t len = j->in_off[i + 1] - j->in_off[i];
        const unsigned hs = (unsigned)be16(d + 6);
        add_patch_public(j, j->out_off[i], hdr.data() + hdr_at[i], hs);
        const uint64_t body_end = (uint64_t)hs + (uint64_t)J.streams[i].frame_count * J.streams[i].frame_size;
        if (body_end < len) {        // bytes behind the last frame pass through unchanged
            if (j->d_src) j->dev_copies.push_back({j->in_off[i] + body_end, j->out_off[i] + body_end, len - body_end});
            else add_patch_public(j, j->out_off[i] + body_end, d + body_end, (uint32_t)(len - body_end));
        }
    }
    return OK;
}
// HcaEncode (hca.cpp:3459-3489): plan per stream on the host (hca.cpp:2206-2462), frames on the device.
int plan_hca_encode(cri_ctx* c, cri_job* j) {
    (void)c;
    HcaJob& J = j->hca;
    std::vector<uint64_t> sizes(j->n, 0);
    std::vector<WavInfo> wavs(j->n);
    std::vector<HcaEncPlan> plans(j->n);
    parallel_for(j->n, [&](uint32_t i) {
        const uint8_t* d = j->blob + j->in_off[i];
        const uint64_t len = j->in_off[i + 1] - j->in_off[i];
        const int r = parse_wav(d, len, &wavs[i]);
        if (r < 0) { j->status[i] = ERR_WAV_BASE + r; return; }
        const bool looping = wavs[i].looping && !j->adx.force_not_looping;
        int pr;
        if (looping) pr = plan_hca_encode_loop(wavs[i], j->quality, &plans[i]);   // loop chunk + pre / post audio (hca.cpp:2292-2321, 3000-3053)
        else pr = plan_hca_encode((unsigned)wavs[i].channels, (unsigned)wavs[i].rate, wavs[i].total_samples / (unsigned)wavs[i].channels, j->quality, &plans[i]);
        if (pr == ERR_UNSUPPORTED) { j->status[i] = ERR_UNSUPPORTED; return; }
        if (pr < 0 || plans[i].frame_size < 8) { j->status[i] = ERR_HCA_CHANNELS; return; }
        sizes[i] = (uint64_t)plans[i].header_size + (uint64_t)plans[i].frame_count * plans[i].frame_size;
    });
    finish_layout_public(j, sizes);
    J.streams.assign(j->n, HcaStreamDev{});
    J.frame_prefix.assign(j->n + 1, 0);
    J.crc_mul.assign((size_t)j->n * 32, 0);
    J.max_channels = 1;
    uint32_t max_frame = 8;
    std::vector<uint8_t> hdr;
    std::unordered_map<uint32_t, std::vector<uint16_t>> crc_mul_by_size;
    for (uint32_t i = 0; i < j->n; i++) {
        uint64_t frames = 0;
        if (j->status[i] == OK) {
            const HcaEncPlan& p = plans[i];
            HcaStreamDev s{};
            s.in_off = p.loop_flag ? hca_loop_input_offset(j, i, wavs[i], p) : pcm16_offset(j, i, wavs[i]);
            s.out_off = j->out_off[i] + p.header_size;
            s.frame_size = p.frame_size;
            s.frame_count = p.frame_count;
            s.out_samples = p.loop_flag ? p.frame_count * 1024 : p.samples;   // looping: the assembled feeder input, all of it
            s.channels = (uint8_t)p.channels;
            s.total_bands = (uint8_t)p.total_bands;
            s.base_bands = (uint8_t)p.base_bands;
            s.stereo_bands = (uint8_t)p.stereo_bands;
            s.bands_per_hfr = (uint8_t)p.bands_per_hfr;
            s.hfr_groups = (uint8_t)p.hfr_groups;
            s.min_res = 1;
            s.max_res = 15;
            for (unsigned ch = 0; ch < p.channels; ch++) { s.type[ch] = p.type[ch]; s.coded[ch] = (uint8_t)p.coded[ch]; }
            J.streams[i] = s;
            J.max_channels = std::max<uint32_t>(J.max_channels, p.channels);
            max_frame = std::max(max_frame, p.frame_size);
            hdr.assign(p.header_size, 0);
            write_hca_header(hdr.data(), p);
            add_patch_public(j, j->out_off[i], hdr.data(), p.header_size);
            frames = p.frame_count;
            j->units += frames;
            // CRC chunk multipliers: lane l covers bytes [l*chunk, (l+1)*chunk) of the frame body (whole words); appending k bytes
            // multiplies a CRC by x^(8k), i.e. k table steps on the value 1 (the CRC register is x^16-scaled already)
            auto& mul = crc_mul_by_size[p.frame_size];
            if (mul.empty()) {
                mul.resize(32);
                const int body = (int)p.frame_size - 2, chunk = (((body + 3) / 4 + 31) / 32) * 4;
                uint16_t v = 1;
                int done = 0;
                const uint8_t zero = 0;
                for (int l = 31; l >= 0; l--) {      // bytes after lane l's chunk grow as l falls: one running product
                    const int end = std::min(std::min(l * chunk, body) + chunk, body);
                    for (; done < body - end; done++) v = crc16_append(v, zero);
                    mul[l] = v;
                }
            }
            std::copy(mul.begin(), mul.end(), J.crc_mul.begin() + (size_t)i * 32);
        }
        J.frame_prefix[i + 1] = J.frame_prefix[i] + frames;
    }
    J.enc_frame_words = (max_frame + 3) / 4 + 2;
    return OK;
}

template <class T>
static int upload(cri_ctx* c, cudaStream_t s, const std::vector<T>& v, T** d) {
    *d = nullptr;
    if (v.empty()) return OK;
    const int r = pool_alloc(c, (void**)d, v.size() * sizeof(T));
    if (r != OK) return r;
    CU_TRY(c, cudaMemcpyAsync(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return OK;
}

int upload_hca_tables(cri_ctx* c, cri_job* j) {
    HcaJob& J = j->hca;
    cudaStream_t s = j->stream;
    int r = upload(c, s, J.streams, &J.d_streams);
    if (r == OK) r = upload(c, s, J.units, &J.d_units);
    if (r == OK) r = upload(c, s, J.cipher_tables, &J.d_cipher);
    if (r == OK) r = upload(c, s, J.ath_tables, &J.d_ath);
    if (r == OK) r = upload(c, s, J.frame_prefix, &J.d_frame_prefix);
    if (r == OK) r = upload(c, s, J.crc_mul, &J.d_crc_mul);
    if (r == OK) r = upload(c, s, J.dec_prefix, &J.d_dec_prefix);
    if (r == OK) r = upload(c, s, J.group_prefix, &J.d_group_prefix);
    if (r == OK) r = upload(c, s, J.group_table, &J.d_group_table);
    if (r == OK && J.spec_bytes) r = pool_alloc(c, (void**)&J.d_spec, J.spec_bytes);
    if (r == OK && J.q_bytes) r = pool_alloc(c, (void**)&J.d_q, J.q_bytes);
    if (r == OK && J.g_bytes) r = pool_alloc(c, (void**)&J.d_g, J.g_bytes);
    if (r == OK && J.i_bytes) r = pool_alloc(c, (void**)&J.d_i, J.i_bytes);
    if (r == OK && J.s_bytes) r = pool_alloc(c, (void**)&J.d_s, J.s_bytes);
    if (r == OK && J.n_bytes) r = pool_alloc(c, (void**)&J.d_n, J.n_bytes);
    return r;
}

int run_hca(cri_ctx* c, cri_job* j, bool* have_dominant) {
    HcaJob& J = j->hca;
    if (j->kind == CRI_JOB_HCA_DECODE) {
        HcaDecodeArgs a{};
        a.in = j->d_in;
        a.out = j->d_out;
        a.streams = J.d_streams;
        a.units = J.d_units;
        a.cipher = J.d_cipher;
        a.ath = J.d_ath;
        a.quant = reinterpret_cast<uint4*>(J.d_q);
        a.gain = reinterpret_cast<float*>(J.d_g);
        a.scratch = reinterpret_cast<uint32_t*>(J.d_s);
        a.scratch_words = J.scratch_words;
        a.n_units = (uint32_t)J.units.size();
        a.uniform = J.uniform;
        a.inten = reinterpret_cast<uint32_t*>(J.d_i);
        a.carry = J.d_i + J.i_count * sizeof(uint32_t);
        a.carry_scan = J.any_pair ? 1u : 0u;
        a.status = j->d_status;
        a.total_groups = J.total_groups;
        a.steps = J.max_steps;
        a.max_channels = J.max_channels;
        a.dec_prefix = J.d_dec_prefix;
        a.spec = reinterpret_cast<float4*>(J.d_spec);
        a.total_frames = J.total_frames;
        a.n_streams = j->n;
        a.run_len = J.run_len;
        a.n_runs = J.n_runs;
        a.joint = J.any_joint ? 1u : 0u;
        a.force_careful = env_u32("CRI_HCA_CAREFUL", 0);
        a.one2 = 0x3F8000003F800000ull;
        if (J.d_n) {
            const uint64_t slots = (uint64_t)J.units.size() * J.max_steps;
            a.sfres = J.d_n;
            a.draws = reinterpret_cast<uint32_t*>(J.d_n + slots * J.max_channels * 128);
            a.frame_draws = a.draws + slots * J.max_channels;
            a.frame_state = a.frame_draws + J.noise_frames;
        }
        // dominant kernel = the transform (second) kernel: ev[2] sits between the two launches
        if (J.n_runs) {
            HcaFastPipe& P = c->hca_pipe;
            P.chunks = hca_fast_chunks();
            if (P.chunks > 1 && !P.side) {                  // first pipelined job of this context
                int lo = 0, hi = 0;
                cudaDeviceGetStreamPriorityRange(&lo, &hi);
                CU_TRY(c, cudaStreamCreateWithPriority(&P.side, cudaStreamNonBlocking, hi));
                CU_TRY(c, cudaEventCreateWithFlags(&P.start, cudaEventDisableTiming));
                CU_TRY(c, cudaEventCreateWithFlags(&P.done, cudaEventDisableTiming));
                for (auto& e : P.unpacked) CU_TRY(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            }
            launch_hca_decode_fast(a, j->stream, &c->launches, j->ev[2], &P);
        }
        else launch_hca_decode(a, j->stream, &c->launches, j->ev[2]);
        CU_TRY(c, cudaEventRecord(j->ev[3], j->stream));
        *have_dominant = J.total_groups != 0 || J.n_runs != 0;
        return OK;
    }
    if (j->kind == CRI_JOB_HCA_CRYPT) {
        HcaCryptArgs a{};
        a.in = j->d_in;
        a.out = j->d_out;
        a.streams = J.d_streams;
        a.frame_prefix = J.d_frame_prefix;
        a.tables = J.d_cipher;
        a.n_frames = J.frame_prefix.empty() ? 0 : J.frame_prefix.back();
        a.n_streams = j->n;
        a.n_tables = (uint32_t)(J.cipher_tables.size() / 256);
        a.group_prefix = J.d_group_prefix;
        a.group_table = J.d_group_table;
        a.n_groups = J.group_prefix.empty() ? 0 : J.group_prefix.back();
        a.frames_per_group = J.frames_per_group;
        a.group_bytes = J.group_bytes;
        a.lut_table = J.lut_table;
        a.min_frame = J.min_frame;
        CU_TRY(c, cudaEventRecord(j->ev[2], j->stream));
        launch_hca_crypt(a, j->stream, &c->launches);
        CU_TRY(c, cudaEventRecord(j->ev[3], j->stream));
        *have_dominant = a.n_frames != 0;
        return OK;
    }
    if (j->kind == CRI_JOB_HCA_ENCODE) {
        HcaEncodeArgs a{};
        a.in = j->d_in;
        a.out = j->d_out;
        a.streams = J.d_streams;
        a.frame_prefix = J.d_frame_prefix;
        a.crc_mul = J.d_crc_mul;
        a.status = j->d_status;
        a.n_frames = J.frame_prefix.empty() ? 0 : J.frame_prefix.back();
        a.n_streams = j->n;
        a.max_channels = J.max_channels;
        a.frame_words = J.enc_frame_words;
        CU_TRY(c, cudaEventRecord(j->ev[2], j->stream));
        if (launch_hca_encode(a, j->stream, &c->launches) != 0) {
            c->error = "HCA encode: frame size / channel count exceeds the kernel's shared-memory budget";
            return ERR_CUDA;
        }
        CU_TRY(c, cudaEventRecord(j->ev[3], j->stream));
        *have_dominant = a.n_frames != 0;
        return OK;
    }
    return ERR_UNSUPPORTED;
}

void free_hca_tables(cri_ctx* c, cri_job* j) {
    HcaJob& J = j->hca;
    for (void* p : {(void*)J.d_streams, (void*)J.d_units, (void*)J.d_s, (void*)J.d_frame_prefix, (void*)J.d_crc_mul,
                    (void*)J.d_cipher, (void*)J.d_ath, (void*)J.d_q, (void*)J.d_g, (void*)J.d_i, (void*)J.d_dec_prefix, (void*)J.d_spec, (void*)J.d_group_prefix, (void*)J.d_group_table, (void*)J.d_n})
        pool_free(c, p);
}

}  // namespace cri
