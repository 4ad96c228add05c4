// Host-side header / container logic. See formats.h for the reference map.
#include "formats.h"

#include <cmath>
#include <cstring>

#include "cri_tables.h"

namespace cri {

static const uint16_t kCrc[256] = CRI_TBL_CRC16;
static const uint8_t kAthBase[656] = CRI_TBL_ATH_BASE;
static const int16_t kAdxStatic[8] = CRI_TBL_ADX_STATIC_COEF;

uint16_t crc16(const uint8_t* p, size_t n) {  // hca.cpp:205-211
    unsigned s = 0;
    for (size_t i = 0; i < n; i++) s = ((s << 8) ^ kCrc[((s >> 8) ^ p[i]) & 0xFF]) & 0xFFFF;
    return (uint16_t)s;
}

// ------------------------------------------------------------------ WAV
namespace {
constexpr uint32_t fourcc(char a, char b, char c, char d) {
    return (uint32_t)(uint8_t)a | ((uint32_t)(uint8_t)b << 8) | ((uint32_t)(uint8_t)c << 16) | ((uint32_t)(uint8_t)d << 24);
}
}  // namespace

int parse_wav(const uint8_t* d, size_t n, WavInfo* w) {
    *w = WavInfo{};
    if (n < 12 || le32(d) != fourcc('R', 'I', 'F', 'F') || le32(d + 8) != fourcc('W', 'A', 'V', 'E')) return -1;
    const uint32_t riff_size = le32(d + 4);
    size_t at = 12;
    uint32_t walked = 4;  // the reference counts the "WAVE" tag first (pcm.cpp:296)
    bool got_fmt = false, got_data = false;
    uint32_t format = 0, align = 0, bits = 0, data_bytes = 0, valid_bits = 0, sub_format = 0;
    while (walked < riff_size) {
        if (at + 8 > n) return -7;
        const uint32_t tag = le32(d + at);
        const uint32_t body = le32(d + at + 4);
        uint32_t span = body + 8;
        if ((span & 1) && span + walked + 1 <= riff_size) span += 1;  // pad byte only if it still fits (pcm.cpp:300)
        if (tag == fourcc('f', 'm', 't', ' ')) {
            if (body < 16 || at + 24 > n) return -2;
            format = le16(d + at + 8);
            w->channels = (int)le16(d + at + 10);
            w->rate = (int)le32(d + at + 12);
            align = le16(d + at + 20);
            bits = le16(d + at + 22);
            if (format != 1 && format != 3 && format != 0xFFFE) return -3;
            if (format == 0xFFFE) {                       // WAVE_FORMAT_EXTENSIBLE: valid bits + sub-format GUID (pcm.cpp:204-215)
                if (body < 40 || at + 48 > n) return -8;      // no valid-bits field: the reference sees bit depth 0
                valid_bits = le16(d + at + 26);
                sub_format = le32(d + at + 32);
            }
            got_fmt = true;
        } else if (tag == fourcc('s', 'm', 'p', 'l')) {
            if (body < 36 || at + 44 > n) return -4;
            const uint32_t loops = le32(d + at + 36), extra = le32(d + at + 40);
            if ((uint64_t)body < (uint64_t)loops * 24 + extra + 36) return -5;
            w->loop_count = loops;
            if (loops) {
                if (at + 68 > n) return -5;
                w->loop_start = le32(d + at + 52);
                w->loop_end = le32(d + at + 56);
            }
            w->looping = 1;
        } else if (tag == fourcc('d', 'a', 't', 'a')) {
            data_bytes = body;
            w->data_offset = at + 8;
            got_data = true;
        }
        at += span;
        walked += span;
        if (walked > riff_size) return -7;
    }
    if (!got_fmt) return -2;
    if (!got_data) return -6;
    // PCM::load_WAVE (pcm.cpp:291-327): bit depth / container combinations the reference can turn into PCM16; the
    // combinations it mishandles (null source pointer, unconverted buffer) are rejected with its -8
    if (w->channels < 1) return -8;
    const uint32_t depth = format == 0xFFFE ? valid_bits : bits;
    const uint32_t mode = format == 0xFFFE ? (sub_format == 0xFFFE ? 1u : sub_format) : format;
    const uint32_t sample_bytes = align / (uint32_t)w->channels;
    if (mode == 3) {
        if ((depth != 32 && depth != 64) || sample_bytes != depth / 8) return -8;
        w->format = depth == 32 ? WAV_F32 : WAV_F64;
    } else if (mode == 1) {
        if (depth >= 1 && depth <= 8 && sample_bytes == 1) { w->format = WAV_U8; w->shift = (uint8_t)depth; }
        else if (depth >= 9 && depth <= 16 && sample_bytes == 2) w->format = WAV_S16;
        else if (depth >= 17 && depth <= 24 && sample_bytes == 3) { w->format = WAV_S24; w->shift = (uint8_t)(depth - 16); }
        else if (depth == 32 && sample_bytes == 4) w->format = WAV_S32;
        else return -8;
    } else {
        return -3;
    }
    w->sample_bytes = (uint8_t)sample_bytes;
    if (w->data_offset + data_bytes > n) return -7;
    w->total_samples = data_bytes / sample_bytes;
    return 0;
}

static int16_t clamp_like_reference(int v) {               // Clamp<int>(value, 32767), pcm.cpp:155-161
    return (int16_t)(v > 32767 ? 32767 : v < -32768 ? -32768 : v);
}

int16_t wav_sample_s16(const WavInfo& w, const uint8_t* data, size_t index) {
    const uint8_t* p = data + index * w.sample_bytes;
    switch (w.format) {
        case WAV_U8: return (int16_t)(((int)p[0] - (1 << (w.shift - 1))) << 8);
        case WAV_S24: {
            int v = (int)p[0] | ((int)p[1] << 8) | ((int)p[2] << 16);
            if (v & 0x800000) v |= ~0xFFFFFF;
            return (int16_t)((v >> w.shift) & 0xFFFF);
        }
        case WAV_S32: { int32_t v; memcpy(&v, p, 4); return (int16_t)((v >> 16) & 0xFFFF); }
        case WAV_F32: {
            float f; memcpy(&f, p, 4);
            const float s = f * 32767.0f;
            return clamp_like_reference(std::fabs(s) < 2147483648.0f ? (int)s : INT32_MIN);   // cvttss2si: out of range -> INT_MIN
        }
        case WAV_F64: {
            double f; memcpy(&f, p, 8);
            const double s = f * 32767.0;
            return clamp_like_reference(std::fabs(s) < 2147483648.0 ? (int)s : INT32_MIN);
        }
        default: { int16_t v; memcpy(&v, p, 2); return v; }
    }
}

size_t wav_header_size(bool looping) { return looping ? 0x70 : 0x2C; }

void write_wav_header(uint8_t* out, uint32_t samples, int channels, int rate, bool looping, uint32_t loop_start,
                      uint32_t loop_end) {
    const size_t hdr = wav_header_size(looping);
    const uint32_t data = samples * (uint32_t)channels * 2u;
    memset(out, 0, hdr);
    memcpy(out, "RIFF", 4);
    put_le32(out + 4, (uint32_t)(hdr - 8) + data);
    memcpy(out + 8, "WAVEfmt ", 8);
    put_le32(out + 16, 16);
    put_le16(out + 20, 1);
    put_le16(out + 22, (uint32_t)channels);
    put_le32(out + 24, (uint32_t)rate);
    put_le32(out + 28, 2u * (uint32_t)channels * (uint32_t)rate);
    put_le16(out + 32, 2u * (uint32_t)channels);
    put_le16(out + 34, 16);
    size_t at = 36;
    if (looping) {  // one-loop smpl chunk, pcm.cpp:258-265
        memcpy(out + 36, "smpl", 4);
        put_le32(out + 40, 0x3C);
        put_le32(out + 36 + 0x24, 1);
        put_le32(out + 36 + 0x34, loop_start);
        put_le32(out + 36 + 0x38, loop_end);
        at = 104;
    }
    memcpy(out + at, "data", 4);
    put_le32(out + at + 4, data);
}

// ------------------------------------------------------------------ ADX
void adx_coefficients(unsigned highpass, unsigned rate, int coef[2]) {
    // adx.cpp:58-64; PI / SQRT2 are the reference's truncated literals (adx.cpp:6-7).
    const double kPi = 3.141592653589793, kSqrt2 = 1.414213562373095;
    const double a = kSqrt2 - std::cos(2.0 * kPi * (double)(highpass & 0xFFFFu) / (double)rate);
    const double b = kSqrt2 - 1;
    const double c = (a - std::sqrt((a + b) * (a - b))) / b;
    coef[0] = (int)(c * 8192);
    coef[1] = (int)(c * c * -4096);
}

int parse_adx(const uint8_t* d, size_t n, AdxInfo* a) {
    *a = AdxInfo{};
    if (n < 20 || be16(d) != 0x8000) return -1;
    a->data_offset = (int)be16(d + 2);
    a->mode = d[4];
    a->block_size = d[5];
    a->bit_depth = d[6];
    a->channels = d[7];
    a->rate = be32(d + 8);
    a->samples = be32(d + 12);
    a->highpass = be16(d + 16);
    a->version = d[18];
    const int flag = d[19];
    if (a->mode == 0x10 || a->mode == 0x11 || a->version == 6 || !a->block_size || !a->bit_depth) return -2;  // AHX
    if (flag == 8 || flag == 9) return -3;                                                                      // encrypted
    if (a->mode < 2 || a->mode > 4) return -4;
    if (a->version < 3 || a->version > 5) return -5;
    if (((a->block_size - 2) * 8) % a->bit_depth != 0 || a->bit_depth >= 16) return -6;
    if (!a->channels) return -7;

    int cursor = 20;
    bool maybe_loop = false;
    if (a->version == 4) {
        cursor += 4;
        if ((size_t)cursor + 4 * (size_t)a->channels > n) return -1;
        for (int c = 0; c < a->channels; c++) {
            a->history[c][0] = (int16_t)be16(d + cursor + 4 * c);
            a->history[c][1] = (int16_t)be16(d + cursor + 4 * c + 2);
        }
        cursor += a->channels > 1 ? 4 * a->channels : 8;
        maybe_loop = cursor + 24 <= a->data_offset - 2;
    } else if (a->version == 3) {
        maybe_loop = cursor + 24 <= a->data_offset - 2;
    }
    if (maybe_loop) {
        if ((size_t)cursor + 24 > n) return -1;
        const unsigned count = be16(d + cursor + 2);
        if (count) {
            if ((long)cursor + 4 + 20L * count >= (long)a->data_offset - 2) return -8;
            a->loop_start = be32(d + cursor + 8);
            a->loop_end = be32(d + cursor + 16);
            a->looping = 1;
        }
    }
    // "(c)CRI" plus its NUL: 7 bytes ending inside block 0 (adx.cpp:345-348).
    if (a->data_offset < 2 || (size_t)a->data_offset + 5 > n) return -9;
    if (memcmp(d + a->data_offset - 2, "(c)CRI", 7) != 0) return -9;
    a->samples_per_block = (uint32_t)((a->block_size - 2) * 8 / a->bit_depth);
    a->blocks = (uint32_t)ceilf((float)a->samples / (float)a->samples_per_block);
    adx_coefficients(a->highpass, a->rate, a->coef);
    return 0;
}

static int round_up_to(int v, int m) {  // IO.hpp:30-36
    if (m <= 0 || v % m == 0) return v;
    return v + m - v % m;
}

int plan_adx_encode(const WavInfo& w, unsigned bit_depth, unsigned block_size, unsigned mode, unsigned highpass,
                    unsigned filter, unsigned version, AdxEncPlan* p, bool looping) {
    *p = AdxEncPlan{};
    const unsigned ch = (unsigned)w.channels & 0xFFu;  // the reference narrows to unsigned char (adx.cpp:418)
    if (ch < 1) return -10;
    if (bit_depth <= 1 || bit_depth >= 16) return -11;
    if (block_size <= 2 || block_size > 255) return -12;
    if (mode != 2 && mode != 3 && mode != 4) return -13;
    if (filter > 3) return -15;
    if (version != 3 && version != 4 && version != 5) return -16;
    if ((8 * (block_size - 2)) % bit_depth != 0) return -17;
    if (w.total_samples < ch || w.total_samples % ch != 0) return -18;
    p->channels = (int)ch;
    p->mode = (int)mode;
    p->block_size = (int)block_size;
    p->bit_depth = (int)bit_depth;
    p->version = (int)version;
    p->filter = (int)filter;
    p->highpass = highpass & 0xFFFFu;
    p->rate = (uint32_t)w.rate;
    p->samples = w.total_samples / ch;
    p->samples_per_block = (block_size - 2) * 8 / bit_depth;
    // A ragged tail is padded to a multiple of (block_size-2) samples, then cut to whole blocks (adx.cpp:450-452).
    if (p->samples % p->samples_per_block)
        p->frames = (uint32_t)round_up_to((int)p->samples, (int)block_size - 2) / p->samples_per_block;
    else
        p->frames = p->samples / p->samples_per_block;
    int hs = 20 + 6;
    if (version != 3) hs += ch > 1 ? 4 * (int)ch : 8;
    if (looping) {                       // one loop: alignment + count words and a 20-byte ADXLoop (adx.cpp:483-484)
        hs += 4 + 20;
        p->looping = 1;
        p->loop_start = w.loop_start;
        p->loop_end = w.loop_end;
    }
    p->header_size = (hs + 15) / 16 * 16;
    p->out_size = (size_t)p->header_size + (size_t)p->frames * ch * block_size + block_size;
    if (mode == 2) {
        p->coef[0] = kAdxStatic[filter * 2];
        p->coef[1] = kAdxStatic[filter * 2 + 1];
    } else {
        adx_coefficients(p->highpass, p->rate, p->coef);
    }
    return 0;
}

void write_adx_frame(uint8_t* out, const AdxEncPlan& p, const int16_t* first) {
    // adx.cpp:359-379 (header) and :499-502 (EOF block). The caller zero-fills `out`.
    put_be16(out, 0x8000);
    put_be16(out + 2, (uint32_t)p.header_size - 4);
    out[4] = (uint8_t)p.mode;
    out[5] = (uint8_t)p.block_size;
    out[6] = (uint8_t)p.bit_depth;
    out[7] = (uint8_t)p.channels;
    put_be32(out + 8, p.rate);
    put_be32(out + 12, p.samples);
    put_be16(out + 16, p.mode == 2 ? 0 : p.highpass);
    out[18] = (uint8_t)p.version;
    out[19] = 0;
    if (p.version != 3)
        for (int c = 0; c < p.channels; c++) {  // both taps start at the channel's first sample (adx.cpp:471-477)
            put_be16(out + 24 + 4 * c, (uint16_t)first[c]);
            put_be16(out + 26 + 4 * c, (uint16_t)first[c]);
        }
    if (p.looping) {                     // Loop::writeLoops + ADXLoop::writeLoop, adx.cpp:94-107, 133-143
        uint8_t* l = out + 20 + (p.version != 3 ? 4 + (p.channels > 1 ? 4 * p.channels : 8) : 0);
        const uint32_t in_frame = (uint32_t)(p.block_size - 2) * 2;
        const uint32_t align = (uint32_t)round_up_to((int)p.loop_start, (int)(p.channels == 1 ? in_frame * 2 : in_frame)) & 0xFFFFu;
        const uint32_t spf = p.samples_per_block, bs = (uint32_t)p.block_size, ch = (uint32_t)p.channels;
        const uint32_t start = p.loop_start + align, end = p.loop_end + align;
        put_be16(l, align);
        put_be16(l + 2, 1);
        put_be16(l + 4, 0);
        put_be16(l + 6, 1);
        put_be32(l + 8, start);
        put_be32(l + 12, (uint32_t)p.header_size + ((start / spf) * bs) * ch);
        put_be32(l + 16, end);
        put_be32(l + 20, (uint32_t)p.header_size + (uint32_t)round_up_to((int)((end / spf) * bs + (end % spf) / bs), (int)bs) * ch);
    }
    memcpy(out + p.header_size - 6, "(c)CRI", 6);
    uint8_t* eof = out + p.out_size - p.block_size;
    put_be16(eof, 0x8001);
    put_be16(eof + 2, (uint32_t)p.block_size - 4);
}

// ------------------------------------------------------------------ HCA
namespace {
struct ChunkWalker {  // the reference walks chunks in a fixed order with a shrinking byte budget
    const uint8_t* base;
    size_t at;
    size_t left;
    bool is(uint32_t tag, size_t need) const { return left >= need && (be32(base + at) & 0x7F7F7F7Fu) == tag; }
    void take(size_t bytes, bool charge = true) {
        at += bytes;
        if (charge) left -= bytes;
    }
};
constexpr uint32_t tag4(char a, char b, char c, char d) {
    return ((uint32_t)(uint8_t)a << 24) | ((uint32_t)(uint8_t)b << 16) | ((uint32_t)(uint8_t)c << 8) | (uint32_t)(uint8_t)d;
}

void assign_roles(unsigned per_track, unsigned config, uint8_t* t) {  // hca.cpp:887-960 / 2323-2400
    constexpr uint8_t P = 1, S = 2;
    auto pair = [&](int i) { t[i] = P; t[i + 1] = S; };
    switch (per_track) {
        case 2: case 3: pair(0); break;
        case 4: pair(0); if (config == 0) pair(2); break;
        case 5: pair(0); if (config <= 2) pair(3); break;
        case 6: case 7: pair(0); pair(4); break;
        case 8: pair(0); pair(4); pair(6); break;
        default: break;
    }
}
}  // namespace

int parse_hca(const uint8_t* d, size_t n, HcaInfo* h) {
    *h = HcaInfo{};
    if (n < 8) return -1;
    if ((be32(d) & 0x7F7F7F7Fu) != tag4('H', 'C', 'A', 0)) return -2;
    h->version = be16(d + 4);
    h->header_size = be16(d + 6);
    switch (h->version) {
        case 0x0101: case 0x0102: case 0x0103: case 0x0200: case 0x0300: break;
        default: return -2;
    }
    if (n < h->header_size || h->header_size < 8) return -1;
    if (crc16(d, h->header_size)) return -3;
    ChunkWalker c{d, 8, (size_t)h->header_size - 8};

    if (!c.is(tag4('f', 'm', 't', 0), 0x10)) return -2;
    h->channels = d[c.at + 4];
    h->rate = be32(d + c.at + 4) & 0xFFFFFFu;
    h->frame_count = be32(d + c.at + 8);
    h->delay = be16(d + c.at + 12);
    h->padding = be16(d + c.at + 14);
    if (h->channels < 1 || h->channels > 16 || !h->frame_count || h->rate < 1 || h->rate > 0x7FFFFF) return -2;
    c.take(0x10);

    if (c.is(tag4('c', 'o', 'm', 'p'), 0x10)) {
        const uint8_t* q = d + c.at;
        h->frame_size = be16(q + 4);
        h->min_res = q[6]; h->max_res = q[7]; h->tracks = q[8]; h->channel_config = q[9];
        h->total_bands = q[10]; h->base_bands = q[11]; h->stereo_bands = q[12]; h->bands_per_hfr = q[13];
        h->ms_stereo = q[14];
        c.take(0x10);
    } else if (c.is(tag4('d', 'e', 'c', 0), 0x0C)) {
        const uint8_t* q = d + c.at;
        h->frame_size = be16(q + 4);
        h->min_res = q[6]; h->max_res = q[7];
        h->total_bands = q[8] + 1u; h->base_bands = q[9] + 1u;
        h->tracks = q[10] >> 4; h->channel_config = q[10] & 0xF;
        if (q[11] == 0) h->base_bands = h->total_bands;
        h->stereo_bands = h->total_bands - h->base_bands;
        c.take(0x0C);
    } else {
        return -2;
    }
    if (c.is(tag4('v', 'b', 'r', 0), 0x08)) {
        const unsigned vmax = be16(d + c.at + 4);
        if (!(h->frame_size == 0 && vmax > 8 && vmax <= 0x1FF)) return -2;
        c.take(0x08);
    }
    if (c.is(tag4('a', 't', 'h', 0), 0x06)) {
        h->ath_type = be16(d + c.at + 4);
        c.take(6, /*charge=*/false);  // the reference forgets to shrink its budget here (hca.cpp:744-747)
    } else {
        h->ath_type = h->version < 0x0200 ? 1 : 0;
    }
    if (c.is(tag4('l', 'o', 'o', 'p'), 0x10)) {
        const uint8_t* q = d + c.at;
        h->loop_start_frame = be32(q + 4);
        h->loop_end_frame = be32(q + 8);
        h->loop_start_delay = be16(q + 12);
        h->loop_end_padding = be16(q + 14);
        h->loop_flag = 1;
        if (!(h->loop_start_frame <= h->loop_end_frame && h->loop_end_frame < h->frame_count)) return -2;
        c.take(0x10);
    }
    if (c.is(tag4('c', 'i', 'p', 'h'), 0x06)) {
        h->ciph_type = be16(d + c.at + 4);
        if (h->ciph_type != 0 && h->ciph_type != 1 && h->ciph_type != 56) return -2;
        c.take(6);
    }
    if (c.is(tag4('r', 'v', 'a', 0), 0x08)) c.take(8);
    if (c.is(tag4('c', 'o', 'm', 'm'), 0x05)) {
        const unsigned len = d[c.at + 4];
        if (len > c.left) return -2;
        c.take(5 + len);
    }

    if (h->frame_size < 8 || h->frame_size > 0xFFFF) return -2;
    if (h->version <= 0x0200) {
        if (h->min_res != 1 || h->max_res != 15) return -2;
    } else if (h->min_res > h->max_res || h->max_res > 15) {
        return -2;
    }
    if (!h->tracks) h->tracks = 1;
    if (h->tracks > h->channels) return -2;
    if (h->total_bands > 128 || h->base_bands > 128 || h->stereo_bands > 128 || h->base_bands + h->stereo_bands > 128 ||
        h->bands_per_hfr > 128)
        return -2;
    // total < base + stereo makes the reference's group count wrap (hca.cpp:872-874) and its decoder write far outside its
    // scalefactor array: no defined behaviour to reproduce, so such a header is refused
    if (h->total_bands < h->base_bands + h->stereo_bands) return -2;
    if (h->bands_per_hfr) {
        const unsigned rest = h->total_bands - h->base_bands - h->stereo_bands;
        h->hfr_groups = rest / h->bands_per_hfr + (rest % h->bands_per_hfr ? 1 : 0);
    }
    if (h->ath_type == 1) {  // hca.cpp:456-471
        unsigned acc = 0;
        for (int i = 0; i < 128; i++) {
            acc += h->rate;
            const unsigned idx = acc >> 13;
            if (idx >= 654) {
                memset(h->ath + i, 0xFF, (size_t)(128 - i));
                break;
            }
            h->ath[i] = kAthBase[idx];
        }
    } else if (h->ath_type != 0) {
        return -2;
    }
    const unsigned per_track = h->channels / h->tracks;
    if (h->stereo_bands > 0 && per_track > 1)
        for (unsigned t = 0; t < h->tracks; t++) assign_roles(per_track, h->channel_config, h->type + t * per_track);
    for (unsigned i = 0; i < h->channels; i++)
        h->coded[i] = h->type[i] == 2 ? h->base_bands : h->base_bands + h->stereo_bands;
    if (h->ms_stereo) return -2;
    return 0;
}

uint64_t mix_subkey(uint64_t key, unsigned subkey) {  // hca.cpp:3309-3311
    subkey &= 0xFFFFu;
    if (subkey) key *= ((uint64_t)subkey << 16) | (uint16_t)((uint16_t)~subkey + 2u);
    return key;
}

static void nibble_sequence(uint8_t out[16], unsigned seed) {  // hca.cpp:524-534
    const unsigned mul = ((seed & 1) << 3) | 5, add = (seed & 0xE) | 1;
    unsigned v = seed >> 4;
    for (int i = 0; i < 16; i++) {
        v = (v * mul + add) & 0xF;
        out[i] = (uint8_t)v;
    }
}

int cipher_table(int type, uint64_t key, uint8_t t[256]) {
    if (type == 56 && key == 0) type = 0;
    if (type == 0) {
        for (int i = 0; i < 256; i++) t[i] = (uint8_t)i;
        return 0;
    }
    if (type == 1) {  // keyless: one LCG walk that skips the two fixed points
        unsigned v = 0;
        for (int i = 1; i < 255; i++) {
            v = (v * 13 + 11) & 0xFF;
            if (v == 0 || v == 0xFF) v = (v * 13 + 11) & 0xFF;
            t[i] = (uint8_t)v;
        }
        t[0] = 0;
        t[255] = 0xFF;
        return 0;
    }
    if (type != 56) return -2;
    key -= 1;
    uint8_t k[7];
    for (auto& b : k) {
        b = (uint8_t)key;
        key >>= 8;
    }
    const uint8_t seed[16] = {k[1], (uint8_t)(k[1] ^ k[6]), (uint8_t)(k[2] ^ k[3]), k[2],
                              (uint8_t)(k[2] ^ k[1]), (uint8_t)(k[3] ^ k[4]), k[3], (uint8_t)(k[3] ^ k[2]),
                              (uint8_t)(k[4] ^ k[5]), k[4], (uint8_t)(k[4] ^ k[3]), (uint8_t)(k[5] ^ k[6]),
                              k[5], (uint8_t)(k[5] ^ k[4]), (uint8_t)(k[6] ^ k[1]), k[6]};
    uint8_t hi[16], lo[16], grid[256];
    nibble_sequence(hi, k[0]);
    for (int r = 0; r < 16; r++) {
        nibble_sequence(lo, seed[r]);
        for (int c = 0; c < 16; c++) grid[r * 16 + c] = (uint8_t)((hi[r] << 4) | lo[c]);
    }
    unsigned x = 0, fill = 1;
    for (int i = 0; i < 256; i++) {
        x = (x + 17) & 0xFF;
        if (grid[x] != 0 && grid[x] != 0xFF) t[fill++] = grid[x];
    }
    t[0] = 0;
    t[255] = 0xFF;
    return 0;
}

void crypt_header(uint8_t* hd, unsigned header_size, unsigned new_type) {
    auto flip = [](uint8_t* p) {
        for (int i = 0; i < 4; i++)
            if (p[i] & 0x7F) p[i] ^= 0x80;
    };
    ChunkWalker c{hd, 0, header_size};
    if (c.is(tag4('H', 'C', 'A', 0), 0)) { flip(hd + c.at); c.take(8); }
    if (c.is(tag4('f', 'm', 't', 0), 0x10)) { flip(hd + c.at); c.take(0x10); }
    if (c.is(tag4('c', 'o', 'm', 'p'), 0x10)) { flip(hd + c.at); c.take(0x10); }
    else if (c.is(tag4('d', 'e', 'c', 0), 0x0C)) { flip(hd + c.at); c.take(0x0C); }
    if (c.is(tag4('v', 'b', 'r', 0), 0x08)) { flip(hd + c.at); c.take(8); }
    if (c.is(tag4('a', 't', 'h', 0), 0x06)) { flip(hd + c.at); c.take(6, false); }
    if (c.is(tag4('l', 'o', 'o', 'p'), 0x10)) { flip(hd + c.at); c.take(0x10); }
    if (c.is(tag4('c', 'i', 'p', 'h'), 0x06)) { flip(hd + c.at); put_be16(hd + c.at + 4, new_type); c.take(6); }
    if (c.is(tag4('r', 'v', 'a', 0), 0x08)) { flip(hd + c.at); c.take(8); }
    if (c.is(tag4('c', 'o', 'm', 'm'), 0x05)) { flip(hd + c.at); c.take(5 + (size_t)hd[c.at + 4]); }
    if (c.is(tag4('p', 'a', 'd', 0), 0x04)) flip(hd + c.at);
    put_be16(hd + header_size - 2, crc16(hd, header_size - 2));
}

// -------------------------------------------------------- HCA encode plan
static unsigned ceil_div_float(int v, int d) {  // DivideByRoundUp (hca.cpp:182-184) rounds in fp32
    return (unsigned)(int)std::ceil((float)v / (float)d);
}

int plan_hca_encode(unsigned channels, unsigned rate, unsigned samples, unsigned quality, HcaEncPlan* p) {
    *p = HcaEncPlan{};
    if (channels < 1 || channels > 8 || rate < 1) return -3;
    static const uint8_t default_config[9] = {0, 1, 0, 4, 0, 1, 3, 7, 3};
    static const uint8_t config_ok[8] = {0x02, 0x01, 0x16, 0x29, 0x86, 0x08, 0x80, 0x08};  // bit i: config i allowed
    p->channels = channels;
    p->rate = rate;
    p->samples = samples;
    const unsigned pcm_rate = rate * channels * 16;
    unsigned ratio;
    switch (quality) {  // CalculateBitrate; unknown enum values fall back to "High" (chunk.py:73 vs hca.cpp:78)
        case 0: ratio = 4; break;
        case 2: ratio = 8; break;
        case 3: ratio = channels == 1 ? 10 : 12; break;
        case 4: ratio = channels == 1 ? 12 : 16; break;
        default: ratio = 6; break;
    }
    unsigned bitrate = pcm_rate / ratio;
    if (bitrate > pcm_rate / 4) bitrate = pcm_rate / 4;
    p->frame_size = bitrate * 1024 / rate / 8;
    const bool wide = channels <= 1 || pcm_rate / bitrate <= 6;
    const unsigned hfr_ratio = wide ? 6 : 8, cutoff_ratio = wide ? 12 : 16;
    unsigned cutoff = rate / 2;
    if (bitrate < pcm_rate / cutoff_ratio) {
        const unsigned alt = cutoff_ratio * bitrate / (32 * channels);
        if (alt < cutoff) cutoff = alt;
    }
    const unsigned total = (unsigned)std::round((double)cutoff * 256.0 / rate);
    unsigned hfr_start = (unsigned)std::round(((double)hfr_ratio * bitrate * 128.0) / pcm_rate);
    if (hfr_start > total) hfr_start = total;
    const unsigned stereo_start = hfr_ratio == 6 ? hfr_start : (hfr_start + 1) / 2;
    const unsigned hfr_bands = total - hfr_start;
    p->total_bands = total;
    p->base_bands = stereo_start;
    p->stereo_bands = hfr_start - stereo_start;
    p->bands_per_hfr = ceil_div_float((int)hfr_bands, 8);
    if (p->bands_per_hfr) {
        p->hfr_band_count = total - p->base_bands - p->stereo_bands;
        p->hfr_groups = ceil_div_float((int)p->hfr_band_count, (int)p->bands_per_hfr);
    }
    p->channel_config = default_config[channels];
    if (!((config_ok[channels - 1] >> p->channel_config) & 1)) return -3;
    p->frame_count = ceil_div_float((int)(samples + p->delay), 1024);
    p->padding = p->frame_count * 1024 - p->delay - samples;
    if (p->stereo_bands && channels > 1) assign_roles(channels, p->channel_config, p->type);
    for (unsigned c = 0; c < channels; c++)
        p->coded[c] = p->type[c] == 2 ? p->base_bands : p->base_bands + p->stereo_bands;
    return 0;
}

int plan_hca_encode_loop(const WavInfo& w, unsigned quality, HcaEncPlan* p) {
    const unsigned ch = (unsigned)w.channels, column = w.total_samples, n = column / (ch ? ch : 1);
    const int r = plan_hca_encode(ch, (unsigned)w.rate, n, quality, p);
    if (r < 0) return r;
    const unsigned ls = w.loop_start, le = w.loop_end;
    if (w.loop_count != 1 || ls >= le || le > n) return ERR_UNSUPPORTED;
    const unsigned sc = std::min(le, column);                                  // hca.cpp:2443 (compares with ColumnSize)
    unsigned delay = 128 + ((unsigned)round_up_to((int)ls, 1024) - ls);       // :2444
    unsigned lstart = ls + delay, lend = le + delay;                           // CalculateLoopInfo, :2292-2305
    p->loop_start_frame = lstart / 1024;
    p->loop_start_delay = lstart % 1024;
    p->loop_end_frame = lend / 1024;
    p->loop_end_padding = 1024 - lend % 1024;
    if (p->loop_end_padding == 1024) { p->loop_end_frame--; p->loop_end_padding = 0; }
    unsigned input = std::min((unsigned)round_up_to((int)sc, 128), column) + 256;   // :2446-2447
    p->post_samples = input - sc;
    // CalculateHeaderSize, :2307-2321: the loop start frame is moved to a 2048-byte boundary with extra delay frames
    p->header_size = 96;
    {
        const unsigned off = p->header_size + p->frame_size * p->loop_start_frame;
        const unsigned pad_bytes = (unsigned)round_up_to((int)off, 2048) - off;
        const unsigned pad_frames = pad_bytes / p->frame_size;
        delay += pad_frames * 1024;
        p->loop_start_frame += pad_frames;
        p->loop_end_frame += pad_frames;
        p->header_size += pad_bytes % p->frame_size;
    }
    p->delay = delay;
    p->frame_count = ceil_div_float((int)(input + delay), 1024);
    p->padding = p->frame_count * 1024 - delay - input;
    p->loop_flag = 1;
    p->loop_start = ls;
    p->main_samples = sc;
    const unsigned pre = delay - 128;                                          // BufferPreSamples; PreEncode, :3000-3012
    p->pre_zero = pre ? ((pre - 1) / 1024) * 1024 : 0;
    p->pre_first = pre - p->pre_zero;
    // SaveLoopAudio (:3014-3022) copies from 1024-sample input chunks while the main audio is being fed: everything it
    // wants must lie inside the input and inside the chunks visited before the loop end is reached
    const unsigned visited = std::min(n, ((sc + 1023) / 1024) * 1024);
    if (ls + p->post_samples > visited) return ERR_UNSUPPORTED;
    p->samples = sc;
    return 0;
}

void write_hca_header(uint8_t* hd, const HcaEncPlan& p) {  // PackHeader, hca.cpp:3109-3164
    memset(hd, 0, p.header_size);
    put_be32(hd, tag4('H', 'C', 'A', 0));
    put_be16(hd + 4, 0x0200);
    put_be16(hd + 6, p.header_size);
    put_be32(hd + 8, tag4('f', 'm', 't', 0));
    put_be32(hd + 12, p.rate);
    hd[12] = (uint8_t)p.channels;
    put_be32(hd + 16, p.frame_count);
    put_be16(hd + 20, p.delay);
    put_be16(hd + 22, p.padding);
    put_be32(hd + 24, tag4('c', 'o', 'm', 'p'));
    put_be16(hd + 28, p.frame_size);
    hd[30] = 1;
    hd[31] = 15;
    hd[32] = 1;
    hd[33] = (uint8_t)p.channel_config;
    hd[34] = (uint8_t)p.total_bands;
    hd[35] = (uint8_t)p.base_bands;
    hd[36] = (uint8_t)p.stereo_bands;
    hd[37] = (uint8_t)p.bands_per_hfr;
    unsigned at = 40;
    if (p.loop_flag) {
        put_be32(hd + 40, tag4('l', 'o', 'o', 'p'));
        put_be32(hd + 44, p.loop_start_frame);
        put_be32(hd + 48, p.loop_end_frame);
        put_be16(hd + 52, p.loop_start_delay);
        put_be16(hd + 54, p.loop_end_padding);
        at = 56;
    }
    put_be32(hd + at, tag4('c', 'i', 'p', 'h'));
    put_be32(hd + at + 6, tag4('p', 'a', 'd', 0));
    put_be16(hd + p.header_size - 2, crc16(hd, p.header_size - 2));
}

}  // namespace cri
