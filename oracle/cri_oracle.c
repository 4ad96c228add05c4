/* TEST INFRASTRUCTURE ONLY -- "oracle": a plain-C, single-threaded CPU
 * restatement of the reference's ADX / HCA algorithms (Youjose/PyCriCodecs @
 * 0bd67a6, files under CriCodecs/). Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product path (pycricodecs_b200/) never does.
 *
 * PARITY IS PINNED: tests/test_oracle_vs_ref.py compares every entry point
 * byte for byte with the compiled reference (oracle/_ref, built from
 * /root/reference by oracle/Makefile) and tests/golden/ holds digests and
 * vectors generated from that build (tools/make_golden.py).
 *
 * All fp32 arithmetic is written one operation per statement and the file is
 * compiled with -ffp-contract=off and no -march flags, because the reference
 * build (setup.py:13: -std=c++11 -O3) is scalar SSE2 without FMA.
 *
 * Deliberate differences from the reference (its undefined behaviour is not
 * reproduced):
 *   - every buffer access is bounds-checked against the caller's length;
 *   - output memory the reference leaves uninitialised is zero;
 *   - ADX header size uses the real channel count (adx.cpp:482 reads an
 *     uninitialised field); caller's buffers are never modified in place.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "cri_tables.h"

#define API __attribute__((visibility("default")))

static const uint16_t T_CRC[256] = CRI_TBL_CRC16;
static const uint32_t T_DEC_SCALING[64] = CRI_TBL_DEC_SCALING;
static const uint32_t T_DEC_RANGE[16] = CRI_TBL_DEC_RANGE;
static const uint32_t T_SCALE_CONV[128] = CRI_TBL_SCALE_CONV;
static const uint32_t T_INTENSITY[16] = CRI_TBL_INTENSITY_RATIO;
static const uint32_t T_ISIN[7 * 64] = CRI_TBL_IMDCT_SIN;
static const uint32_t T_ICOS[7 * 64] = CRI_TBL_IMDCT_COS;
static const uint32_t T_WINDOW[128] = CRI_TBL_WINDOW;
static const uint8_t T_INVERT[66] = CRI_TBL_INVERT;
static const uint8_t T_MAX_BITS[16] = CRI_TBL_MAX_BITS;
static const uint8_t T_READ_BITS[128] = CRI_TBL_READ_BITS;
static const int8_t T_READ_VALS[128] = CRI_TBL_READ_VALS;
static const uint8_t T_ENC_CURVE[59] = CRI_TBL_ENC_RES_CURVE;
static const uint8_t T_ENC_QBITS[128] = CRI_TBL_ENC_Q_BITS;
static const uint8_t T_ENC_QCODE[128] = CRI_TBL_ENC_Q_CODE;
static const uint32_t T_ENC_INV_STEP[16] = CRI_TBL_ENC_INV_STEP;
static const uint32_t T_ENC_DEAD_ZONE[16] = CRI_TBL_ENC_DEAD_ZONE;
static const uint32_t T_ENC_RATIO_BOUNDS[14] = CRI_TBL_ENC_RATIO_BOUNDS;
static const uint32_t T_ENC_Q_SCALING[64] = CRI_TBL_ENC_Q_SCALING;
static const uint32_t T_MSIN[8 * 128] = CRI_TBL_MDCT_SIN;
static const uint32_t T_MCOS[8 * 128] = CRI_TBL_MDCT_COS;
static const uint8_t T_SHUFFLE[128] = CRI_TBL_ENC_SHUFFLE;
static const int16_t T_ADX_STATIC[8] = CRI_TBL_ADX_STATIC_COEF;
static const uint8_t T_ATH_BASE[656] = CRI_TBL_ATH_BASE;

static inline float bits2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static inline uint32_t be16(const uint8_t* p) { return ((uint32_t)p[0] << 8) | p[1]; }
static inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static inline uint32_t le16(const uint8_t* p) { return p[0] | ((uint32_t)p[1] << 8); }
static inline uint32_t le32(const uint8_t* p) { return p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline void put_be16(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)v; }
static inline void put_be32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
static inline void put_le16(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static inline void put_le32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }

static inline int clamp_sym(int v, int limit) { /* pcm.cpp:155-161: [~limit, limit] */
    if (v > limit) return limit;
    if (v < ~limit) return ~limit;
    return v;
}

static int next_multiple(int value, int multiple) { /* IO.hpp:30-36 */
    if (multiple <= 0) return value;
    if (value % multiple == 0) return value;
    return value + multiple - value % multiple;
}

/* ------------------------------------------------------------------ */
/* MSB-first bit I/O. The reference's BitWriter ORs into the first byte
 * it touches and assigns the rest (IO.cpp:129-181); on zeroed memory that
 * is a plain MSB-first append, which is what this writer does.          */
typedef struct { uint8_t* buf; size_t nbits; size_t pos; } bitw_t;

static void bitw_put(bitw_t* w, uint32_t value, int count) {
    if (count <= 0 || w->pos + (size_t)count > w->nbits) return; /* IO.cpp:131-134: silently dropped */
    for (int i = count - 1; i >= 0; i--) {
        if ((value >> i) & 1u) w->buf[w->pos >> 3] |= (uint8_t)(0x80u >> (w->pos & 7));
        w->pos++;
    }
}

/* vgmstream-style reader used by the HCA decoder (hca.cpp:216-291): a peek
 * past the end returns 0, the cursor still advances. */
typedef struct { const uint8_t* buf; int nbits; int pos; } bitr_t;

static uint32_t bitr_peek(const bitr_t* r, int count) {
    if (count <= 0 || r->pos < 0 || r->pos + count > r->nbits) return 0;
    uint32_t v = 0;
    for (int i = 0; i < count; i++) {
        int p = r->pos + i;
        v = (v << 1) | ((r->buf[p >> 3] >> (7 - (p & 7))) & 1u);
    }
    return v;
}
static uint32_t bitr_read(bitr_t* r, int count) { uint32_t v = bitr_peek(r, count); r->pos += count; return v; }

/* ------------------------------------------------------------------ */
/* WAV (RIFF) ingest: pcm.cpp:291-342, 411-444. 16-bit PCM only here.    */
typedef struct {
    int channels, rate, looping;
    uint32_t loop_start, loop_end;
    const int16_t* pcm;      /* interleaved, little endian host assumed */
    uint32_t total_samples;  /* ColumnSize: samples over all channels    */
} wav_t;

static int wav_parse(const uint8_t* d, size_t n, wav_t* w) {
    memset(w, 0, sizeof *w);
    if (n < 12 || le32(d) != 0x46464952u || le32(d + 8) != 0x45564157u) return -1;
    uint32_t full = le32(d + 4);
    size_t off = 12;
    uint32_t sum = 4;
    int have_fmt = 0, have_data = 0;
    uint32_t comp = 0, block_align = 0, bits = 0, data_size = 0;
    const uint8_t* data_ptr = NULL;
    while (sum < full) {
        if (off + 8 > n) return -7;
        uint32_t sig = le32(d + off);
        uint32_t size = le32(d + off + 4) + 8;
        size += (uint32_t)((size & 1) && (size + sum + (size & 1) <= full));
        if (sig == 0x20746D66u) { /* "fmt " */
            if (le32(d + off + 4) < 16 || off + 24 > n) return -2;
            comp = le16(d + off + 8);
            w->channels = (int)le16(d + off + 10);
            w->rate = (int)le32(d + off + 12);
            block_align = le16(d + off + 20);
            bits = le16(d + off + 22);
            if (comp != 1 && comp != 0xFFFE && comp != 3) return -3;
            have_fmt = 1;
        } else if (sig == 0x6C706D73u) { /* "smpl" */
            uint32_t sz = le32(d + off + 4);
            if (sz < 36 || off + 44 > n) return -4;
            uint32_t nloops = le32(d + off + 36), extra = le32(d + off + 40);
            if ((uint64_t)sz < (uint64_t)nloops * 24 + extra + 36) return -5;
            if (nloops > 0) {
                if (off + 44 + 24 > n) return -5;
                w->loop_start = le32(d + off + 44 + 8);
                w->loop_end = le32(d + off + 44 + 12);
            }
            w->looping = 1;
        } else if (sig == 0x61746164u) { /* "data" */
            data_size = le32(d + off + 4);
            data_ptr = d + off + 8;
            have_data = 1;
        }
        off += size;
        sum += size;
        if (sum > full) return -7;
    }
    if (!have_fmt) return -2;
    if (!have_data) return -6;
    if (comp != 1 || bits != 16 || w->channels < 1 || block_align / (uint32_t)w->channels != 2) return -8;
    if ((size_t)(data_ptr - d) + data_size > n) return -7;
    w->pcm = (const int16_t*)data_ptr;
    w->total_samples = data_size / 2;
    return 0;
}

/* WAV image writer: pcm.cpp:350-375, 547-556 (0x2C header, 0x70 with smpl). */
static size_t wav_image_size(uint32_t samples, int channels, int looping) {
    return (size_t)(looping ? 0x70 : 0x2C) + (size_t)samples * channels * 2;
}
static int16_t* wav_write_header(uint8_t* out, uint32_t samples, int channels, int rate, int looping,
                                 uint32_t loop_start, uint32_t loop_end) {
    size_t hdr = looping ? 0x70 : 0x2C;
    uint32_t data = samples * (uint32_t)channels * 2u;
    memset(out, 0, hdr);
    put_le32(out + 0, 0x46464952u);
    put_le32(out + 4, (uint32_t)(hdr + data - 8));
    put_le32(out + 8, 0x45564157u);
    put_le32(out + 12, 0x20746D66u);
    put_le32(out + 16, 16);
    put_le16(out + 20, 1);
    put_le16(out + 22, (uint32_t)channels);
    put_le32(out + 24, (uint32_t)rate);
    put_le32(out + 28, 2u * (uint32_t)channels * (uint32_t)rate);
    put_le16(out + 32, 2u * (uint32_t)channels);
    put_le16(out + 34, 16);
    size_t pos = 36;
    if (looping) {
        put_le32(out + 36, 0x6C706D73u);
        put_le32(out + 40, 0x3C);
        put_le32(out + 36 + 0x24, 1);
        put_le32(out + 36 + 0x34, loop_start);
        put_le32(out + 36 + 0x38, loop_end);
        pos = 104;
    }
    put_le32(out + pos, 0x61746164u);
    put_le32(out + pos + 4, data);
    return (int16_t*)(out + hdr);
}

/* ================================================================== */
/* ADX                                                                  */
/* ================================================================== */

/* adx.cpp:58-64 with the truncated PI / SQRT2 literals of adx.cpp:6-7. */
API void cri_oracle_adx_coefficients(unsigned highpass, unsigned rate, int* c) {
    const double pi = 3.141592653589793, sqrt2 = 1.414213562373095;
    double a = sqrt2 - cos(2.0 * pi * (double)(highpass & 0xFFFF) / (double)rate);
    double b = sqrt2 - 1;
    double k = (a - sqrt((a + b) * (a - b))) / b;
    c[0] = (int)(k * 8192);
    c[1] = (int)(k * k * -4096);
}

typedef struct {
    int mode, block_size, bit_depth, channels, version, data_offset, looping;
    uint32_t rate, samples, highpass, loop_start, loop_end;
    int16_t hist[256][2];
} adx_hdr_t;

/* adx.cpp:298-358. Returns 0 or the reference's negative code. */
static int adx_parse(const uint8_t* d, size_t n, adx_hdr_t* h) {
    memset(h, 0, sizeof *h);
    if (n < 20) return -1;
    if (be16(d) != 0x8000) return -1;
    h->data_offset = (int)be16(d + 2);
    h->mode = d[4]; h->block_size = d[5]; h->bit_depth = d[6]; h->channels = d[7];
    h->rate = be32(d + 8); h->samples = be32(d + 12); h->highpass = be16(d + 16);
    h->version = d[18];
    int flag = d[19];
    if (h->mode == 0x11 || h->mode == 0x10 || h->version == 6 || h->block_size == 0 || h->bit_depth == 0) return -2;
    if (flag == 8 || flag == 9) return -3;
    if (h->mode != 2 && h->mode != 3 && h->mode != 4) return -4;
    if (h->version != 3 && h->version != 4 && h->version != 5) return -5;
    if (((h->block_size - 2) * 8) % h->bit_depth != 0 || h->bit_depth >= 16) return -6;
    if (h->channels == 0) return -7;
    int base = 20, looping = 0;
    if (h->version == 4) {
        base += 4;
        for (int i = 0; i < h->channels; i++) {
            if ((size_t)base + 4 * i + 4 > n) return -1;
            h->hist[i][0] = (int16_t)be16(d + base + 4 * i);
            h->hist[i][1] = (int16_t)be16(d + base + 4 * i + 2);
        }
        base += h->channels > 1 ? 4 * h->channels : 8;
        if (base + 24 <= h->data_offset - 2) looping = 1;
    } else if (h->version == 3) {
        if (base + 24 <= h->data_offset - 2) looping = 1;
    }
    if (looping) {
        if ((size_t)base + 4 > n) return -1;
        unsigned count = be16(d + base + 2);
        if (!count) looping = 0;
        else if ((long)base + 4 + (long)count * 20 >= (long)h->data_offset - 2) return -8;
        else {
            if ((size_t)base + 24 > n) return -1;
            h->loop_start = be32(d + base + 4 + 4);
            h->loop_end = be32(d + base + 4 + 12);
        }
    }
    h->looping = looping;
    if ((size_t)h->data_offset + 5 > n || h->data_offset < 2) return -9;
    if (memcmp(d + h->data_offset - 2, "(c)CRI", 7) != 0) return -9; /* 7 bytes: includes the NUL, adx.cpp:345-348 */
    return 0;
}

API int cri_oracle_adx_decoded_size(const uint8_t* adx, size_t n, size_t* out_n) {
    adx_hdr_t h;
    int r = adx_parse(adx, n, &h);
    if (r < 0) return r;
    *out_n = wav_image_size(h.samples, h.channels, h.looping);
    return 0;
}

/* adx.cpp:380-415 + 189-214. */
API int cri_oracle_adx_decode(const uint8_t* adx, size_t n, uint8_t* out, size_t out_cap, size_t* out_n) {
    adx_hdr_t h;
    int r = adx_parse(adx, n, &h);
    if (r < 0) return r;
    size_t need = wav_image_size(h.samples, h.channels, h.looping);
    *out_n = need;
    if (out_cap < need) return -100;
    int16_t* pcm = wav_write_header(out, h.samples, h.channels, (int)h.rate, h.looping, h.loop_start, h.loop_end);
    memset(pcm, 0, (size_t)h.samples * h.channels * 2);
    int spb = (h.block_size - 2) * 8 / h.bit_depth;
    unsigned blocks = (unsigned)ceilf((float)h.samples / (float)spb); /* adx.cpp:386 */
    int coef[2];
    cri_oracle_adx_coefficients(h.highpass, h.rate, coef);
    size_t off = (size_t)h.data_offset + 4;
    for (unsigned b = 0; b < blocks; b++) {
        if (off + 2 > n) break;
        if (adx[off] == 0x80 && adx[off + 1] == 0x01) break; /* EOF block, adx.cpp:405-406 */
        for (int c = 0; c < h.channels; c++, off += (size_t)h.block_size) {
            if (off + (size_t)h.block_size > n) return 0; /* truncated stream: rest stays silent */
            int scale = (int)be16(adx + off);
            if (h.mode == 4) scale = (int)(1u << ((12 - scale) & 31));
            else if (h.mode == 2) {
                int pred = (scale >> 13) & 3; /* 3 bits in the stream, 4 pairs in the table (adx.cpp:45) */
                scale = (scale & 0x1FFF) + 1;
                coef[0] = T_ADX_STATIC[pred * 2];
                coef[1] = T_ADX_STATIC[pred * 2 + 1];
            } else scale += 1;
            bitr_t br = { adx + off + 2, (h.block_size - 2) * 8, 0 };
            int h1 = h.hist[c][0], h2 = h.hist[c][1];
            for (int i = 0; i < spb; i++) {
                int q = (int)bitr_read(&br, h.bit_depth);
                int sh = 32 - h.bit_depth;
                q = (int)((uint32_t)q << sh) >> sh; /* sign-extend, IO.cpp:55-64 */
                int s = q * scale + ((coef[0] * h1) >> 12) + ((coef[1] * h2) >> 12);
                s = clamp_sym(s, 0x7FFF);
                size_t idx = (size_t)b * spb + i;
                if (idx < h.samples) pcm[idx * h.channels + c] = (int16_t)s;
                h2 = h1; h1 = s;
            }
            h.hist[c][0] = (int16_t)h1; h.hist[c][1] = (int16_t)h2;
        }
    }
    return 0;
}

static int ilog2_floor(int v) { int r = 0; while (v > 1) { v >>= 1; r++; } return r; } /* adx.cpp:49-56 */

typedef struct { int error; size_t size; int header, frames, spb, channels; uint32_t spc; } adx_plan_t;

/* adx.cpp:416-486: validation + sizes. */
static adx_plan_t adx_plan(const wav_t* w, unsigned bit_depth, unsigned block_size, unsigned mode,
                           unsigned filter, unsigned version) {
    adx_plan_t p; memset(&p, 0, sizeof p);
    unsigned ch = (unsigned)w->channels & 0xFF; /* unsigned char ChannelCount */
    if (ch < 1) { p.error = -10; return p; }
    if (bit_depth <= 1 || bit_depth >= 16) { p.error = -11; return p; }
    if (block_size <= 2 || block_size > 255) { p.error = -12; return p; }
    if (mode != 2 && mode != 3 && mode != 4) { p.error = -13; return p; }
    if (filter > 3) { p.error = -15; return p; }
    if (version != 3 && version != 4 && version != 5) { p.error = -16; return p; }
    if ((8 * (block_size - 2)) % bit_depth != 0) { p.error = -17; return p; }
    if (w->total_samples < ch || w->total_samples % ch != 0) { p.error = -18; return p; }
    p.channels = (int)ch;
    p.spb = (int)((block_size - 2) * 8 / bit_depth);
    p.spc = w->total_samples / ch;
    if (p.spc % (unsigned)p.spb != 0)
        p.frames = next_multiple((int)p.spc, (int)block_size - 2) / p.spb; /* adx.cpp:450-452 */
    else
        p.frames = (int)(p.spc / (unsigned)p.spb);
    int hs = 20 + 6;
    if (version == 4 || version == 5) hs += ch > 1 ? 4 * (int)ch : 8;
    hs = hs % 16 == 0 ? hs : hs + (16 - hs % 16);
    p.header = hs;
    p.size = (size_t)hs + (size_t)p.frames * ch * block_size + block_size;
    return p;
}

/* adx.cpp:416-506 + 215-273. Looping WAVs (smpl chunk) are not restated yet: -99. */
API int cri_oracle_adx_encode(const uint8_t* wav, size_t n, unsigned bit_depth, unsigned block_size,
                              unsigned mode, unsigned highpass, unsigned filter, unsigned version,
                              uint8_t* out, size_t out_cap, size_t* out_n) {
    wav_t w;
    int r = wav_parse(wav, n, &w);
    if (r < 0) return -100 + r;
    if (w.looping && !(version == 5)) return -99;
    adx_plan_t p = adx_plan(&w, bit_depth, block_size, mode, filter, version);
    if (p.error) return p.error;
    *out_n = p.size;
    if (!out) return 0;
    if (out_cap < p.size) return -100;
    highpass &= 0xFFFF;
    memset(out, 0, p.size);
    int ch = p.channels, spb = p.spb;
    int coef[2];
    if (mode == 2) { coef[0] = T_ADX_STATIC[filter * 2]; coef[1] = T_ADX_STATIC[filter * 2 + 1]; }
    else cri_oracle_adx_coefficients(highpass, (unsigned)w.rate, coef);
    unsigned filter_bits = filter << 13;
    size_t padded = (size_t)p.frames * spb; /* samples per channel actually coded */
    int h1[256], h2[256];
    for (int c = 0; c < ch; c++) {
        h1[c] = h2[c] = 0;
        if (version == 4 || version == 5) h1[c] = h2[c] = w.pcm[c];
    }
    /* header, adx.cpp:359-379 */
    put_be16(out, 0x8000); put_be16(out + 2, (uint32_t)p.header - 4);
    out[4] = (uint8_t)mode; out[5] = (uint8_t)block_size; out[6] = (uint8_t)bit_depth; out[7] = (uint8_t)ch;
    put_be32(out + 8, (uint32_t)w.rate); put_be32(out + 12, p.spc);
    put_be16(out + 16, mode == 2 ? 0 : highpass); out[18] = (uint8_t)version; out[19] = 0;
    if (version == 4 || version == 5)
        for (int c = 0; c < ch; c++) { put_be16(out + 24 + 4 * c, (uint16_t)h1[c]); put_be16(out + 26 + 4 * c, (uint16_t)h2[c]); }
    memcpy(out + p.header - 6, "(c)CRI", 6);

    int limit = (1 << (bit_depth - 1)) - 1;
    size_t off = (size_t)p.header;
    for (int f = 0; f < p.frames; f++) {
        for (int c = 0; c < ch; c++, off += block_size) {
            int smp[128 * 8];
            int o1 = h1[c], o2 = h2[c], mn = 0, mx = 0;
            for (int i = 0; i < spb; i++) {
                size_t idx = (size_t)f * spb + i;
                int s = (idx < p.spc && idx < padded) ? w.pcm[idx * ch + c] : 0;
                int res = ((s * 4096) - coef[0] * h1[c] - coef[1] * h2[c]) >> 12;
                if (res < mn) mn = res; else if (res > mx) mx = res;
                smp[i] = s;
                h2[c] = h1[c]; h1[c] = s;
            }
            if (!mn && !mx) continue; /* zero block; history stays RAW (adx.cpp:231-234) */
            int a = mx / limit, b = mn / ~limit;
            unsigned scale = (unsigned)(a > b ? a : b) & 0xFFFF;
            if (scale > 0x1000) scale = 0x1000;
            bitw_t bw = { out + off, (size_t)block_size * 8, 0 };
            if (mode == 4) {
                unsigned power = scale == 0 ? 0 : (unsigned)ilog2_floor((int)scale) + 1;
                scale = (1u << power) & 0xFFFF;
                bitw_put(&bw, (uint32_t)(12 - (int)power) & 0xFFFF, 16);
            } else if (mode == 2) bitw_put(&bw, (filter_bits | (scale & 0x1FFF)) & 0xFFFF, 16);
            else bitw_put(&bw, scale, 16);
            h1[c] = o1; h2[c] = o2;
            for (int i = 0; i < spb; i++) {
                int d = ((smp[i] * 4096) - coef[0] * h1[c] - coef[1] * h2[c]) >> 12;
                if (!scale) scale = 1;
                d = d > 0 ? d + (int)(scale >> 1) : d - (int)(scale >> 1);
                d /= (int)scale;
                d = clamp_sym(d, limit);
                int sim = (d * 4096 * (int)scale + coef[0] * h1[c] + coef[1] * h2[c]) >> 12;
                sim = clamp_sym(sim, 0x7FFF);
                h2[c] = h1[c]; h1[c] = (int16_t)sim;
                bitw_put(&bw, (uint32_t)d & ((1u << bit_depth) - 1), (int)bit_depth);
            }
        }
    }
    put_be16(out + off, 0x8001);            /* EOF block, adx.cpp:499-502 */
    put_be16(out + off + 2, block_size - 4);
    return 0;
}

/* ================================================================== */
/* HCA common                                                           */
/* ================================================================== */
API unsigned cri_oracle_crc16(const uint8_t* p, size_t n) { /* hca.cpp:205-211 */
    unsigned sum = 0;
    for (size_t i = 0; i < n; i++) sum = ((sum << 8) ^ T_CRC[((sum >> 8) ^ p[i]) & 0xFF]) & 0xFFFF;
    return sum;
}

static void nibble_lcg(uint8_t* r, unsigned key) { /* hca.cpp:524-534 */
    unsigned mul = ((key & 1) << 3) | 5, add = (key & 0xE) | 1;
    key >>= 4;
    for (int i = 0; i < 16; i++) { key = (key * mul + add) & 0xF; r[i] = (uint8_t)key; }
}

/* hca.cpp:499-617. Returns 0 or -2 (unknown type). */
API int cri_oracle_cipher_table(int type, uint64_t key, uint8_t* t) {
    if (type == 56 && !key) type = 0;
    if (type == 0) { for (int i = 0; i < 256; i++) t[i] = (uint8_t)i; return 0; }
    if (type == 1) {
        unsigned v = 0;
        for (int i = 1; i < 255; i++) {
            v = (v * 13 + 11) & 0xFF;
            if (v == 0 || v == 0xFF) v = (v * 13 + 11) & 0xFF;
            t[i] = (uint8_t)v;
        }
        t[0] = 0; t[255] = 0xFF;
        return 0;
    }
    if (type != 56) return -2;
    key--;
    uint8_t kc[7];
    for (int i = 0; i < 7; i++) { kc[i] = (uint8_t)key; key >>= 8; }
    const uint8_t seed[16] = {
        kc[1], (uint8_t)(kc[1] ^ kc[6]), (uint8_t)(kc[2] ^ kc[3]), kc[2], (uint8_t)(kc[2] ^ kc[1]), (uint8_t)(kc[3] ^ kc[4]),
        kc[3], (uint8_t)(kc[3] ^ kc[2]), (uint8_t)(kc[4] ^ kc[5]), kc[4], (uint8_t)(kc[4] ^ kc[3]), (uint8_t)(kc[5] ^ kc[6]),
        kc[5], (uint8_t)(kc[5] ^ kc[4]), (uint8_t)(kc[6] ^ kc[1]), kc[6] };
    uint8_t rows[16], cols[16], base[256];
    nibble_lcg(rows, kc[0]);
    for (int r = 0; r < 16; r++) {
        nibble_lcg(cols, seed[r]);
        for (int c = 0; c < 16; c++) base[r * 16 + c] = (uint8_t)((rows[r] << 4) | cols[c]);
    }
    unsigned x = 0, pos = 1;
    for (int i = 0; i < 256; i++) {
        x = (x + 17) & 0xFF;
        if (base[x] != 0 && base[x] != 0xFF) t[pos++] = base[x];
    }
    t[0] = 0; t[255] = 0xFF;
    return 0;
}

static uint64_t mix_subkey(uint64_t key, unsigned subkey) { /* hca.cpp:3309-3311, 3381-3383 */
    subkey &= 0xFFFF;
    if (subkey) key = key * (((uint64_t)subkey << 16) | (uint16_t)((uint16_t)~subkey + 2u));
    return key;
}
API uint64_t cri_oracle_mix_subkey(uint64_t key, unsigned subkey) { return mix_subkey(key, subkey); }

typedef struct {
    unsigned version, header_size, channels, rate, frame_count, delay, padding;
    unsigned frame_size, min_res, max_res, tracks, channel_config, total_bands, base_bands, stereo_bands;
    unsigned bands_per_hfr, ms_stereo, hfr_groups, ath_type, ciph_type;
    unsigned loop_flag, loop_start_frame, loop_end_frame, loop_start_delay, loop_end_padding;
    int ciph_offset;             /* byte offset of the "ciph" chunk, -1 if none */
    uint8_t ath[128];
    uint8_t type[16];            /* 0 discrete, 1 stereo primary, 2 stereo secondary */
    unsigned coded[16];
} hca_hdr_t;

static int sig_is(const uint8_t* p, uint32_t want) { return (be32(p) & 0x7F7F7F7Fu) == want; }

/* hca.cpp:628-984. 0 ok; -1 params, -2 header, -3 checksum (reference codes). */
static int hca_parse(const uint8_t* d, size_t n, hca_hdr_t* h) {
    memset(h, 0, sizeof *h);
    h->ciph_offset = -1;
    if (n < 8) return -1;
    if (!sig_is(d, 0x48434100u)) return -2;
    h->version = be16(d + 4); h->header_size = be16(d + 6);
    if (h->version != 0x0101 && h->version != 0x0102 && h->version != 0x0103 && h->version != 0x0200 && h->version != 0x0300) return -2;
    if (n < h->header_size) return -1;
    if (cri_oracle_crc16(d, h->header_size)) return -3;
    size_t size = h->header_size - 8, p = 8;
    if (size >= 0x10 && sig_is(d + p, 0x666D7400u)) {
        h->channels = d[p + 4]; h->rate = be32(d + p + 4) & 0xFFFFFF; h->frame_count = be32(d + p + 8);
        h->delay = be16(d + p + 12); h->padding = be16(d + p + 14);
        if (h->channels < 1 || h->channels > 16 || h->frame_count == 0 || h->rate < 1 || h->rate > 0x7FFFFF) return -2;
        size -= 0x10; p += 0x10;
    } else return -2;
    if (size >= 0x10 && sig_is(d + p, 0x636F6D70u)) {
        h->frame_size = be16(d + p + 4); h->min_res = d[p + 6]; h->max_res = d[p + 7]; h->tracks = d[p + 8];
        h->channel_config = d[p + 9]; h->total_bands = d[p + 10]; h->base_bands = d[p + 11]; h->stereo_bands = d[p + 12];
        h->bands_per_hfr = d[p + 13]; h->ms_stereo = d[p + 14];
        size -= 0x10; p += 0x10;
    } else if (size >= 0x0C && sig_is(d + p, 0x64656300u)) {
        h->frame_size = be16(d + p + 4); h->min_res = d[p + 6]; h->max_res = d[p + 7];
        h->total_bands = d[p + 8] + 1u; h->base_bands = d[p + 9] + 1u;
        h->tracks = d[p + 10] >> 4; h->channel_config = d[p + 10] & 0xF;
        unsigned stereo_type = d[p + 11];
        if (stereo_type == 0) h->base_bands = h->total_bands;
        h->stereo_bands = h->total_bands - h->base_bands;
        h->bands_per_hfr = 0;
        size -= 0x0C; p += 0x0C;
    } else return -2;
    if (size >= 0x08 && sig_is(d + p, 0x76627200u)) {
        unsigned vmax = be16(d + p + 4);
        if (!(h->frame_size == 0 && vmax > 8 && vmax <= 0x1FF)) return -2;
        size -= 0x08; p += 0x08;
    }
    if (size >= 0x06 && sig_is(d + p, 0x61746800u)) { h->ath_type = be16(d + p + 4); p += 6; /* size not reduced: hca.cpp:744-747 */ }
    else h->ath_type = h->version < 0x0200 ? 1 : 0;
    if (size >= 0x10 && sig_is(d + p, 0x6C6F6F70u)) {
        h->loop_start_frame = be32(d + p + 4); h->loop_end_frame = be32(d + p + 8);
        h->loop_start_delay = be16(d + p + 12); h->loop_end_padding = be16(d + p + 14);
        h->loop_flag = 1;
        if (!(h->loop_start_frame <= h->loop_end_frame && h->loop_end_frame < h->frame_count)) return -2;
        size -= 0x10; p += 0x10;
    }
    if (size >= 0x06 && sig_is(d + p, 0x63697068u)) {
        h->ciph_type = be16(d + p + 4);
        h->ciph_offset = (int)p;
        if (!(h->ciph_type == 0 || h->ciph_type == 1 || h->ciph_type == 56)) return -2;
        size -= 6; p += 6;
    }
    if (size >= 0x08 && sig_is(d + p, 0x72766100u)) { size -= 8; p += 8; }
    if (size >= 0x05 && sig_is(d + p, 0x636F6D6Du)) {
        unsigned len = d[p + 4];
        if (len > size) return -2;
        size -= 5 + len; p += 5 + len;
    }
    if (!(h->frame_size >= 8 && h->frame_size <= 0xFFFF)) return -2;
    if (h->version <= 0x0200) { if (h->min_res != 1 || h->max_res != 15) return -2; }
    else if (h->min_res > h->max_res || h->max_res > 15) return -2;
    if (h->tracks == 0) h->tracks = 1;
    if (h->tracks > h->channels) return -2;
    if (h->total_bands > 128 || h->base_bands > 128 || h->stereo_bands > 128 ||
        h->base_bands + h->stereo_bands > 128 || h->bands_per_hfr > 128) return -2;
    {
        unsigned a = h->total_bands - h->base_bands - h->stereo_bands, b = h->bands_per_hfr;
        h->hfr_groups = b < 1 ? 0 : a / b + (a % b ? 1 : 0);
    }
    if (h->ath_type == 0) memset(h->ath, 0, 128);
    else if (h->ath_type == 1) { /* hca.cpp:456-471 */
        unsigned acc = 0;
        for (int i = 0; i < 128; i++) {
            acc += h->rate;
            unsigned idx = acc >> 13;
            if (idx >= 654) { memset(h->ath + i, 0xFF, (size_t)(128 - i)); break; }
            h->ath[i] = T_ATH_BASE[idx];
        }
    } else return -2;
    {   /* channel roles, hca.cpp:887-970 */
        unsigned cpt = h->channels / h->tracks;
        if (h->stereo_bands > 0 && cpt > 1) {
            for (unsigned t = 0; t < h->tracks; t++) {
                uint8_t* ct = h->type + t * cpt;
                static const uint8_t P = 1, S = 2, D = 0;
                switch (cpt) {
                case 2: ct[0] = P; ct[1] = S; break;
                case 3: ct[0] = P; ct[1] = S; ct[2] = D; break;
                case 4: ct[0] = P; ct[1] = S; if (h->channel_config == 0) { ct[2] = P; ct[3] = S; } else { ct[2] = D; ct[3] = D; } break;
                case 5: ct[0] = P; ct[1] = S; ct[2] = D; if (h->channel_config <= 2) { ct[3] = P; ct[4] = S; } else { ct[3] = D; ct[4] = D; } break;
                case 6: ct[0] = P; ct[1] = S; ct[2] = D; ct[3] = D; ct[4] = P; ct[5] = S; break;
                case 7: ct[0] = P; ct[1] = S; ct[2] = D; ct[3] = D; ct[4] = P; ct[5] = S; ct[6] = D; break;
                case 8: ct[0] = P; ct[1] = S; ct[2] = D; ct[3] = D; ct[4] = P; ct[5] = S; ct[6] = P; ct[7] = S; break;
                default: break;
                }
            }
        }
        for (unsigned c = 0; c < h->channels; c++)
            h->coded[c] = h->type[c] != 2 ? h->base_bands + h->stereo_bands : h->base_bands;
    }
    if (h->ms_stereo) return -2;
    return 0;
}

/* ================================================================== */
/* HCA decode                                                           */
/* ================================================================== */
typedef struct {
    uint8_t intensity[8], sf[128], res[128], noises[128];
    unsigned noise_count, valid_count;
    float gain[128], spectra[8][128], prev[128], wave[8][128];
} hca_ch_t;

typedef struct { hca_hdr_t h; uint8_t cipher[256]; uint32_t random; hca_ch_t ch[16]; } hca_dec_t;

/* hca.cpp:1290-1358 */
static int unpack_scalefactors(const hca_hdr_t* h, hca_ch_t* ch, int c, bitr_t* br) {
    unsigned count = h->coded[c], extra = 0;
    unsigned delta_bits = bitr_read(br, 3);
    if (!(h->type[c] == 2 || h->hfr_groups == 0 || h->version <= 0x0200)) {
        extra = h->hfr_groups;
        count += extra;
        if (count > 128) return -5;
    }
    if (delta_bits >= 6) {
        for (unsigned i = 0; i < count; i++) ch->sf[i] = (uint8_t)bitr_read(br, 6);
    } else if (delta_bits > 0) {
        unsigned escape = (1u << delta_bits) - 1;
        unsigned value = bitr_read(br, 6);
        ch->sf[0] = (uint8_t)value;
        for (unsigned i = 1; i < count; i++) {
            unsigned delta = bitr_read(br, (int)delta_bits);
            if (delta == escape) value = bitr_read(br, 6);
            else {
                int test = (int)value + ((int)delta - (int)(escape >> 1));
                if (test < 0 || test >= 64) return -5;
                value = (value - (escape >> 1) + delta) & 0x3F;
            }
            ch->sf[i] = (uint8_t)value;
        }
    } else memset(ch->sf, 0, 128);
    for (unsigned i = 0; i < extra; i++) ch->sf[127 - i] = ch->sf[count - i]; /* v3 derived HFR scales */
    return 0;
}

/* hca.cpp:1361-1441 */
static int unpack_intensity(const hca_hdr_t* h, hca_ch_t* ch, int c, bitr_t* br) {
    if (h->type[c] == 2) {
        if (h->version <= 0x0200) {
            unsigned v = bitr_peek(br, 4);
            ch->intensity[0] = (uint8_t)v;
            if (v < 15) {
                br->pos += 4;
                for (int i = 1; i < 8; i++) ch->intensity[i] = (uint8_t)bitr_read(br, 4);
            }
        } else {
            unsigned v = bitr_peek(br, 4);
            if (v < 15) {
                br->pos += 4;
                unsigned db = bitr_read(br, 2);
                ch->intensity[0] = (uint8_t)v;
                if (db == 3) {
                    for (int i = 1; i < 8; i++) ch->intensity[i] = (uint8_t)bitr_read(br, 4);
                } else {
                    unsigned bmax = (2u << db) - 1, bits = db + 1;
                    for (int i = 1; i < 8; i++) {
                        unsigned delta = bitr_read(br, (int)bits);
                        if (delta == bmax) v = bitr_read(br, 4);
                        else {
                            v = (uint8_t)(v - (bmax >> 1) + delta);
                            if (v > 15) return -5;
                        }
                        ch->intensity[i] = (uint8_t)v;
                    }
                }
            } else {
                br->pos += 4;
                memset(ch->intensity, 7, 8);
            }
        }
    } else if (h->version <= 0x0200) {
        uint8_t* hfr = ch->sf + 128 - h->hfr_groups;
        for (unsigned i = 0; i < h->hfr_groups; i++) hfr[i] = (uint8_t)bitr_read(br, 6);
    }
    return 0;
}

/* hca.cpp:1444-1507 */
static void resolve_bands(const hca_hdr_t* h, hca_ch_t* ch, int c, unsigned packed_noise) {
    unsigned count = h->coded[c], nn = 0, nv = 0;
    for (unsigned i = 0; i < count; i++) {
        unsigned r = 0, sf = ch->sf[i];
        if (sf > 0) {
            int noise = (int)h->ath[i] + (int)((packed_noise + i) >> 8);
            int pos = noise + 1 - (int)((5 * sf) >> 1);
            if (pos < 0) r = 15;
            else if (pos <= 65) r = T_INVERT[pos];
            else r = 0;
            if (r > h->max_res) r = h->max_res;
            else if (r < h->min_res) r = h->min_res;
            if (r < 1) ch->noises[nn++] = (uint8_t)i;
            else ch->noises[127 - nv++] = (uint8_t)i;
        }
        ch->res[i] = (uint8_t)r;
    }
    ch->noise_count = nn; ch->valid_count = nv;
    memset(ch->res + count, 0, 128 - count);
    for (unsigned i = 0; i < count; i++)
        ch->gain[i] = bits2f(T_DEC_SCALING[ch->sf[i]]) * bits2f(T_DEC_RANGE[ch->res[i]]);
}

/* hca.cpp:1540-1571 */
static void read_spectra(const hca_hdr_t* h, hca_ch_t* ch, int c, bitr_t* br, int sub) {
    unsigned count = h->coded[c];
    for (unsigned i = 0; i < count; i++) {
        unsigned r = ch->res[i];
        int bits = T_MAX_BITS[r];
        unsigned code = bitr_read(br, bits);
        float q;
        if (r > 7) {
            int v = (1 - (int)((code & 1) << 1)) * (int)(code >> 1);
            if (v == 0) br->pos -= 1;
            q = (float)v;
        } else {
            unsigned idx = (r << 4) + code;
            br->pos += (int)T_READ_BITS[idx] - bits;
            q = (float)T_READ_VALS[idx];
        }
        ch->spectra[sub][i] = ch->gain[i] * q;
    }
    memset(&ch->spectra[sub][count], 0, sizeof(float) * (128 - count));
}

/* hca.cpp:1602-1635 (v3.0 only: min_resolution == 0) */
static void fill_noise(hca_dec_t* d, hca_ch_t* ch, int c, int sub) {
    const hca_hdr_t* h = &d->h;
    if (h->min_res > 0) return;
    if (ch->valid_count == 0 || ch->noise_count == 0) return;
    if (!(!h->ms_stereo || h->type[c] == 1)) return;
    uint32_t rnd = d->random;
    for (unsigned i = 0; i < ch->noise_count; i++) {
        rnd = 0x343FDu * rnd + 0x269EC3u;
        unsigned ri = 128 - ch->valid_count + (((rnd & 0x7FFF) * ch->valid_count) >> 15);
        unsigned ni = ch->noises[i], vi = ch->noises[ri];
        int k = (int)ch->sf[ni] - (int)ch->sf[vi] + 62;
        k &= ~(k >> 31);
        ch->spectra[sub][ni] = bits2f(T_SCALE_CONV[k]) * ch->spectra[sub][vi];
    }
    d->random = rnd;
}

/* hca.cpp:1638-1683 */
static void fill_high_bands(const hca_hdr_t* h, hca_ch_t* ch, int c, int sub) {
    if (h->bands_per_hfr == 0 || h->type[c] == 2) return;
    int start = (int)(h->stereo_bands + h->base_bands);
    int high = start, low = start - 1;
    const uint8_t* hfr = ch->sf + 128 - h->hfr_groups;
    int limit = h->version <= 0x0200 ? (int)h->hfr_groups : (int)(h->hfr_groups >> 1);
    for (int g = 0; g < (int)h->hfr_groups; g++) {
        int step = g < limit ? 1 : 0;
        for (unsigned i = 0; i < h->bands_per_hfr; i++) {
            if (high >= (int)h->total_bands || low < 0) break;
            int k = (int)hfr[g] - (int)ch->sf[low] + 63;
            k &= ~(k >> 31);
            ch->spectra[sub][high] = bits2f(T_SCALE_CONV[k]) * ch->spectra[sub][low];
            high += 1; low -= step;
        }
    }
    ch->spectra[sub][high - 1] = 0.0f;
}

/* hca.cpp:1696-1714 */
static void spread_intensity(const hca_hdr_t* h, hca_ch_t* pair, int c, int sub) {
    if (h->type[c] != 1) return;
    float rl = bits2f(T_INTENSITY[pair[1].intensity[sub]]);
    float rr = 2.0f - rl;
    float* l = pair[0].spectra[sub];
    float* r = pair[1].spectra[sub];
    for (unsigned b = h->base_bands; b < h->total_bands; b++) {
        float cl = l[b] * rl;
        float cr = l[b] * rr;
        l[b] = cl; r[b] = cr;
    }
}

/* hca.cpp:1898-2019: 128-point DCT-IV as 7 sum/difference passes followed by
 * 7 rotation passes, then window + overlap with the previous half block. */
static void imdct128(const float* in, float* prev, float* wave, float* dct_out) {
    float a[128], b[128];
    float* src = a; float* dst = b;
    memcpy(a, in, sizeof a);
    for (int half = 64; half >= 1; half >>= 1) {
        int blocks = 64 / half;
        for (int j = 0; j < blocks; j++) {
            const float* s = src + j * 2 * half;
            float* d = dst + j * 2 * half;
            for (int k = 0; k < half; k++) {
                float x = s[2 * k], y = s[2 * k + 1];
                d[k] = x + y;
                d[half + k] = x - y;
            }
        }
        float* t = src; src = dst; dst = t;
    }
    for (int stage = 0; stage < 7; stage++) {
        int half = 1 << stage, blocks = 64 >> stage;
        const uint32_t* ts = T_ISIN + stage * 64;
        const uint32_t* tc = T_ICOS + stage * 64;
        for (int j = 0; j < blocks; j++) {
            const float* s = src + j * 2 * half;
            float* d = dst + j * 2 * half;
            for (int k = 0; k < half; k++) {
                float x = s[k], y = s[half + k];
                float sn = bits2f(ts[j * half + k]), cs = bits2f(tc[j * half + k]);
                float xs = x * sn, yc = y * cs, xc = x * cs, ys = y * sn;
                d[k] = xs - yc;
                d[2 * half - 1 - k] = xc + ys;
            }
        }
        float* t = src; src = dst; dst = t;
    }
    const float* dct = src;
    if (dct_out) memcpy(dct_out, dct, sizeof a);
    float np[128];
    for (int i = 0; i < 64; i++) {
        float w0 = bits2f(T_WINDOW[i]) * dct[i + 64];
        wave[i] = w0 + prev[i];
        float w1 = bits2f(T_WINDOW[i + 64]) * dct[127 - i];
        wave[i + 64] = w1 - prev[i + 64];
        np[i] = bits2f(T_WINDOW[127 - i]) * dct[63 - i];
        np[i + 64] = bits2f(T_WINDOW[63 - i]) * dct[i];
    }
    memcpy(prev, np, sizeof np);
}

API void cri_oracle_imdct(const float* spectra, float* prev, float* wave, float* dct) { imdct128(spectra, prev, wave, dct); }

/* hca.cpp:1149-1233. frame is modified (deciphered) in place. */
static int hca_decode_frame(hca_dec_t* d, uint8_t* frame, int unpack_only, int* bits_read) {
    const hca_hdr_t* h = &d->h;
    bitr_t br = { frame, (int)h->frame_size * 8, 0 };
    if (bitr_read(&br, 16) != 0xFFFF) return -4;
    if (cri_oracle_crc16(frame, h->frame_size)) return -3;
    for (unsigned i = 0; i < h->frame_size; i++) frame[i] = d->cipher[frame[i]];
    unsigned noise_level = bitr_read(&br, 9);
    unsigned boundary = bitr_read(&br, 7);
    unsigned packed = (noise_level << 8) - boundary;
    for (unsigned c = 0; c < h->channels; c++) {
        int r = unpack_scalefactors(h, &d->ch[c], (int)c, &br);
        if (r < 0) return r;
        unpack_intensity(h, &d->ch[c], (int)c, &br);
        resolve_bands(h, &d->ch[c], (int)c, packed);
    }
    for (int sub = 0; sub < 8; sub++)
        for (unsigned c = 0; c < h->channels; c++) read_spectra(h, &d->ch[c], (int)c, &br, sub);
    if (bits_read) *bits_read = br.pos;
    if (unpack_only) return 0;
    for (int sub = 0; sub < 8; sub++) {
        for (unsigned c = 0; c < h->channels; c++) {
            fill_noise(d, &d->ch[c], (int)c, sub);
            fill_high_bands(h, &d->ch[c], (int)c, sub);
        }
        if (h->stereo_bands > 0)
            for (unsigned c = 0; c + 1 < h->channels; c++) spread_intensity(h, &d->ch[c], (int)c, sub);
        for (unsigned c = 0; c < h->channels; c++)
            imdct128(d->ch[c].spectra[sub], d->ch[c].prev, d->ch[c].wave[sub], d->ch[c].spectra[sub]);
    }
    return 0;
}

/* hca.cpp:339-360; (int) of an out-of-range float is INT_MIN on x86-64 (cvttss2si). */
static int16_t pcm16_from_float(float f) {
    float x = f * 32768.0f;
    int s;
    if (!(x > -2147483648.0f && x < 2147483648.0f)) s = INT32_MIN;
    else s = (int)x;
    if (s > 32767) s = 32767; else if (s < -32768) s = -32768;
    return (int16_t)s;
}

static int hca_dec_init(hca_dec_t* d, const uint8_t* file, size_t n, uint64_t key, unsigned subkey) {
    memset(d, 0, sizeof *d);
    int r = hca_parse(file, n, &d->h);
    if (r < 0) return r;
    d->random = 1;
    return cri_oracle_cipher_table((int)d->h.ciph_type, mix_subkey(key, subkey), d->cipher);
}

API int cri_oracle_hca_info(const uint8_t* file, size_t n, unsigned* info /* [16] */) {
    hca_hdr_t h;
    int r = hca_parse(file, n, &h);
    if (r < 0) return -1;
    unsigned v[16] = { h.version, h.header_size, h.channels, h.rate, h.frame_count, h.delay, h.padding, h.frame_size,
                       h.total_bands, h.base_bands, h.stereo_bands, h.bands_per_hfr, h.hfr_groups, h.ciph_type, h.loop_flag, h.min_res };
    memcpy(info, v, sizeof v);
    return 0;
}

API int cri_oracle_hca_decoded_size(const uint8_t* file, size_t n, size_t* out_n) {
    hca_hdr_t h;
    if (hca_parse(file, n, &h) < 0) return -1;
    uint32_t samples = h.frame_count * 1024u - h.delay - h.padding;
    *out_n = wav_image_size(samples, (int)h.channels, (int)h.loop_flag);
    return 0;
}

/* HcaDecode, hca.cpp:3340-3457. Returns 0, -1 (header) or -2 (frame decode). */
API int cri_oracle_hca_decode(const uint8_t* file, size_t n, uint64_t key, unsigned subkey,
                              uint8_t* out, size_t out_cap, size_t* out_n) {
    hca_dec_t* d = (hca_dec_t*)malloc(sizeof *d);
    if (!d) return -100;
    if (hca_dec_init(d, file, n, key, subkey) < 0) { free(d); return -1; }
    const hca_hdr_t* h = &d->h;
    uint32_t samples = h->frame_count * 1024u - h->delay - h->padding;
    size_t need = wav_image_size(samples, (int)h->channels, (int)h->loop_flag);
    *out_n = need;
    if (out_cap < need) { free(d); return -100; }
    if ((size_t)h->header_size + (size_t)h->frame_count * h->frame_size > n) { free(d); return -1; }
    uint32_t ls = h->loop_start_frame * 1024u + h->loop_start_delay - h->delay;
    uint32_t le = h->loop_end_frame * 1024u + (1024u - h->loop_end_padding) - h->delay;
    int16_t* pcm = wav_write_header(out, samples, (int)h->channels, (int)h->rate, (int)h->loop_flag, ls, le);
    memset(pcm, 0, (size_t)samples * h->channels * 2);
    uint8_t* buf = (uint8_t*)malloc(h->frame_size);
    int rc = 0;
    for (unsigned f = 0; f < h->frame_count; f++) {
        int64_t first = (int64_t)f * 1024 - (int64_t)h->delay; /* output index of this frame's sample 0 */
        if (first >= (int64_t)samples) break;                  /* the reference stops once enough samples exist */
        memcpy(buf, file + h->header_size + (size_t)f * h->frame_size, h->frame_size);
        if (hca_decode_frame(d, buf, 0, NULL) < 0) { rc = -2; break; }
        for (int sub = 0; sub < 8; sub++)
            for (int i = 0; i < 128; i++) {
                int64_t o = first + sub * 128 + i;
                if (o < 0 || o >= (int64_t)samples) continue;
                for (unsigned c = 0; c < h->channels; c++)
                    pcm[(size_t)o * h->channels + c] = pcm16_from_float(d->ch[c].wave[sub][i]);
            }
    }
    free(buf); free(d);
    return rc;
}

/* Probe: decode frames [f0,f1) from a reset decoder, raw 1024*channels PCM16 per frame. */
API int cri_oracle_hca_decode_range(const uint8_t* file, size_t n, uint64_t key, unsigned f0, unsigned f1, int16_t* pcm) {
    hca_dec_t* d = (hca_dec_t*)malloc(sizeof *d);
    if (hca_dec_init(d, file, n, key, 0) < 0) { free(d); return -1; }
    const hca_hdr_t* h = &d->h;
    uint8_t* buf = (uint8_t*)malloc(h->frame_size);
    int rc = 0;
    for (unsigned f = f0; f < f1 && f < h->frame_count; f++) {
        memcpy(buf, file + h->header_size + (size_t)f * h->frame_size, h->frame_size);
        if ((rc = hca_decode_frame(d, buf, 0, NULL)) < 0) break;
        int16_t* o = pcm + (size_t)(f - f0) * 1024 * h->channels;
        for (int sub = 0; sub < 8; sub++)
            for (int i = 0; i < 128; i++)
                for (unsigned c = 0; c < h->channels; c++) *o++ = pcm16_from_float(d->ch[c].wave[sub][i]);
    }
    free(buf); free(d);
    return rc;
}

/* Probe: state after unpack of one frame (mirrors ref_hca_unpack_dump). */
API int cri_oracle_hca_unpack(const uint8_t* file, size_t n, uint64_t key, unsigned frame, uint8_t* sf, uint8_t* res,
                              uint8_t* inten, float* gain, float* spectra, int* bits) {
    hca_dec_t* d = (hca_dec_t*)malloc(sizeof *d);
    if (hca_dec_init(d, file, n, key, 0) < 0) { free(d); return -1; }
    const hca_hdr_t* h = &d->h;
    uint8_t* buf = (uint8_t*)malloc(h->frame_size);
    memcpy(buf, file + h->header_size + (size_t)frame * h->frame_size, h->frame_size);
    int rc = hca_decode_frame(d, buf, 1, bits);
    for (unsigned c = 0; c < h->channels; c++) {
        memcpy(sf + 128 * c, d->ch[c].sf, 128); memcpy(res + 128 * c, d->ch[c].res, 128);
        memcpy(inten + 8 * c, d->ch[c].intensity, 8); memcpy(gain + 128 * c, d->ch[c].gain, 512);
        memcpy(spectra + 1024 * c, d->ch[c].spectra, 4096);
    }
    free(buf); free(d);
    return rc;
}

/* ================================================================== */
/* HCA crypt: HcaCrypt hca.cpp:3271-3337 + CryptHeader :3166-3250.      */
/* Works on a copy: `io` holds the file and is transformed in place.    */
/* ================================================================== */
static void toggle_sig(uint8_t* p) { for (int i = 0; i < 4; i++) if (p[i] & 0x7F) p[i] ^= 0x80; }

API int cri_oracle_hca_crypt(uint8_t* io, size_t n, int encrypt, unsigned type, uint64_t key, unsigned subkey) {
    hca_hdr_t h;
    if (hca_parse(io, n, &h) < 0) return -1;
    if ((size_t)h.header_size + (size_t)h.frame_count * h.frame_size > n) return -1;
    unsigned ciph = encrypt ? type : h.ciph_type;
    uint8_t t[256], inv[256];
    if (cri_oracle_cipher_table((int)ciph, mix_subkey(key, subkey), t) < 0) return -1;
    if (encrypt) { for (int i = 0; i < 256; i++) inv[t[i]] = (uint8_t)i; memcpy(t, inv, 256); }
    uint8_t* f = io + h.header_size;
    for (unsigned i = 0; i < h.frame_count; i++, f += h.frame_size) {
        for (unsigned k = 0; k < h.frame_size; k++) f[k] = t[f[k]];
        put_be16(f + h.frame_size - 2, cri_oracle_crc16(f, h.frame_size - 2));
    }
    /* header: flip bit 7 of every non-NUL signature byte of each known chunk, in file order */
    size_t size = h.header_size, p = 0;
    unsigned new_type = encrypt ? type : 0;
    if (sig_is(io + p, 0x48434100u)) { toggle_sig(io + p); p += 8; size -= 8; }
    if (size >= 0x10 && sig_is(io + p, 0x666D7400u)) { toggle_sig(io + p); p += 0x10; size -= 0x10; }
    if (size >= 0x10 && sig_is(io + p, 0x636F6D70u)) { toggle_sig(io + p); p += 0x10; size -= 0x10; }
    else if (size >= 0x0C && sig_is(io + p, 0x64656300u)) { toggle_sig(io + p); p += 0x0C; size -= 0x0C; }
    if (size >= 0x08 && sig_is(io + p, 0x76627200u)) { toggle_sig(io + p); p += 8; size -= 8; }
    if (size >= 0x06 && sig_is(io + p, 0x61746800u)) { toggle_sig(io + p); p += 6; }
    if (size >= 0x10 && sig_is(io + p, 0x6C6F6F70u)) { toggle_sig(io + p); p += 0x10; size -= 0x10; }
    if (size >= 0x06 && sig_is(io + p, 0x63697068u)) { toggle_sig(io + p); put_be16(io + p + 4, new_type); p += 6; size -= 6; }
    if (size >= 0x08 && sig_is(io + p, 0x72766100u)) { toggle_sig(io + p); p += 8; size -= 8; }
    if (size >= 0x05 && sig_is(io + p, 0x636F6D6Du)) { toggle_sig(io + p); unsigned len = io[p + 4]; p += 5 + len; size -= 5 + len; }
    if (size >= 0x04 && sig_is(io + p, 0x70616400u)) { toggle_sig(io + p); }
    put_be16(io + h.header_size - 2, cri_oracle_crc16(io, h.header_size - 2));
    return 0;
}

/* ================================================================== */
/* HCA v2.0 encode: hca.cpp:2206-3164 (VGAudio-derived). Non-looping.   */
/* ================================================================== */
typedef struct {
    unsigned channels, rate, samples, frame_size, frame_count, delay, padding, header_size;
    unsigned total_bands, base_bands, stereo_bands, hfr_groups, bands_per_hfr, hfr_band_count, channel_config;
    uint8_t type[16]; unsigned coded[16];
} hca_plan_t;

static unsigned ceil_div_f(int v, int d) { return (unsigned)(int)ceilf((float)v / (float)d); } /* hca.cpp:182-184 */

static const uint8_t k_default_mapping[9] = { 0, 1, 0, 4, 0, 1, 3, 7, 3 };
static const uint8_t k_valid_mapping[8][8] = {
    {0,1,0,0,0,0,0,0},{1,0,0,0,0,0,0,0},{0,1,1,0,1,0,0,0},{1,0,0,1,0,1,0,0},
    {0,1,1,0,0,0,0,1},{0,0,0,1,0,0,0,0},{0,0,0,0,0,0,0,1},{0,0,0,1,0,0,0,0} };

/* hca.cpp:2206-2290, 2307-2321, 2323-2462 */
static int hca_plan(unsigned channels, unsigned rate, unsigned samples_per_channel, unsigned quality, hca_plan_t* p) {
    memset(p, 0, sizeof *p);
    if (channels < 1 || channels > 8) return -3;
    p->channels = channels; p->rate = rate; p->samples = samples_per_channel; p->delay = 128;
    unsigned pcm_rate = rate * channels * 16, ratio = 6;
    switch (quality) {
    case 0: ratio = 4; break;
    case 1: ratio = 6; break;
    case 2: ratio = 8; break;
    case 3: ratio = channels == 1 ? 10 : 12; break;
    case 4: ratio = channels == 1 ? 12 : 16; break;
    default: ratio = 6; break;
    }
    unsigned bitrate = pcm_rate / ratio, maxrate = pcm_rate / 4;
    if (bitrate > maxrate) bitrate = maxrate;
    unsigned cutoff = rate / 2;
    p->frame_size = bitrate * 1024 / rate / 8;
    unsigned hfr_ratio, cutoff_ratio;
    if (channels <= 1 || pcm_rate / bitrate <= 6) { hfr_ratio = 6; cutoff_ratio = 12; }
    else { hfr_ratio = 8; cutoff_ratio = 16; }
    if (bitrate < pcm_rate / cutoff_ratio) {
        unsigned alt = cutoff_ratio * bitrate / (32 * channels);
        if (alt < cutoff) cutoff = alt;
    }
    unsigned total = (unsigned)round((double)cutoff * 256.0 / rate);
    unsigned hfr_start = (unsigned)round(((double)hfr_ratio * bitrate * 128.0) / pcm_rate);
    if (total < hfr_start) hfr_start = total;
    unsigned stereo_start = hfr_ratio == 6 ? hfr_start : (hfr_start + 1) / 2;
    unsigned hfr_bands = total - hfr_start;
    unsigned per_group = ceil_div_f((int)hfr_bands, 8), groups = 0;
    if (per_group > 0) groups = ceil_div_f((int)hfr_bands, (int)per_group);
    p->total_bands = total; p->base_bands = stereo_start; p->stereo_bands = hfr_start - stereo_start;
    p->hfr_groups = groups; p->bands_per_hfr = per_group;
    if (per_group > 0) {
        p->hfr_band_count = total - p->base_bands - p->stereo_bands;
        p->hfr_groups = ceil_div_f((int)p->hfr_band_count, (int)per_group);
    }
    unsigned cfg = k_default_mapping[channels];
    if (k_valid_mapping[channels - 1][cfg] != 1) return -3;
    p->channel_config = cfg;
    p->header_size = 96;
    unsigned total_samples = samples_per_channel + p->delay;
    p->frame_count = ceil_div_f((int)total_samples, 1024);
    p->padding = p->frame_count * 1024 - p->delay - samples_per_channel;
    /* channel roles, hca.cpp:2323-2412 (track_count is always 1 here) */
    if (!(p->stereo_bands == 0 || channels == 1)) {
        static const uint8_t P = 1, S = 2;
        uint8_t* t = p->type;
        switch (channels) {
        case 2: t[0] = P; t[1] = S; break;
        case 3: t[0] = P; t[1] = S; break;
        case 4: t[0] = P; t[1] = S; if (cfg == 0) { t[2] = P; t[3] = S; } break;
        case 5: t[0] = P; t[1] = S; if (cfg <= 2) { t[3] = P; t[4] = S; } break;
        case 6: t[0] = P; t[1] = S; t[4] = P; t[5] = S; break;
        case 7: t[0] = P; t[1] = S; t[4] = P; t[5] = S; break;
        case 8: t[0] = P; t[1] = S; t[4] = P; t[5] = S; t[6] = P; t[7] = S; break;
        }
    }
    for (unsigned c = 0; c < channels; c++) p->coded[c] = p->type[c] == 2 ? p->base_bands : p->base_bands + p->stereo_bands;
    return 0;
}

typedef struct {
    uint8_t sf[128], res[128], intensity[8];
    float wave[8][128], prev[128], spectra[8][128], scaled[128][8], group_avg[8];
    int quant[8][128], hfr_scale[8], header_bits, delta_bits;
} enc_ch_t;

typedef struct { hca_plan_t p; enc_ch_t ch[8]; int noise_level, boundary; } hca_enc_t;

/* hca.cpp:2481-2553 */
static void mdct128(enc_ch_t* ch, int sub) {
    float in[128], t[128];
    const float* cur = ch->wave[sub];
    for (int i = 0; i < 64; i++) {
        float a = bits2f(T_WINDOW[63 - i]) * -cur[64 + i];
        float b = -bits2f(T_WINDOW[64 + i]) * cur[63 - i];
        float c = bits2f(T_WINDOW[i]) * ch->prev[i];
        float d = -bits2f(T_WINDOW[127 - i]) * ch->prev[127 - i];
        in[i] = a - b;
        in[64 + i] = c - d;
    }
    const uint32_t* ts = T_MSIN + 7 * 128; const uint32_t* tc = T_MCOS + 7 * 128;
    for (int i = 0; i < 64; i++) {
        float a = in[2 * i], b = in[127 - 2 * i];
        float sn = bits2f(ts[i]), cs = bits2f(tc[i]);
        float ac = a * cs, bs = b * sn, as = a * sn, bc = b * cs;
        t[2 * i] = ac + bs;
        t[2 * i + 1] = as - bc;
    }
    for (int stage = 0; stage < 6; stage++) {
        int blocks = 1 << stage, size_bits = 6 - stage, half_bits = size_bits - 1;
        int size = 1 << size_bits, half = 1 << half_bits;
        ts = T_MSIN + half_bits * 128; tc = T_MCOS + half_bits * 128;
        for (int blk = 0; blk < blocks; blk++)
            for (int i = 0; i < half; i++) {
                int fp = (blk * size + i) * 2, bp = fp + size;
                float a = t[fp] - t[bp];
                float b = t[fp + 1] - t[bp + 1];
                float sn = bits2f(ts[i]), cs = bits2f(tc[i]);
                t[fp] = t[fp] + t[bp];
                t[fp + 1] = t[fp + 1] + t[bp + 1];
                float ac = a * cs, bs = b * sn, as = a * sn, bc = b * cs;
                t[bp] = ac + bs;
                t[bp + 1] = as - bc;
            }
    }
    for (int i = 0; i < 128; i++) ch->spectra[sub][i] = t[T_SHUFFLE[i]] * 0.125f;
    memcpy(ch->prev, cur, sizeof ch->prev);
}

API void cri_oracle_mdct(const float* wave, float* prev, float* spectra) {
    static enc_ch_t ch;
    memcpy(ch.wave[0], wave, 512); memcpy(ch.prev, prev, 512);
    mdct128(&ch, 0);
    memcpy(spectra, ch.spectra[0], 512); memcpy(prev, ch.prev, 512);
}

static int find_scalefactor(float v) { /* hca.cpp:2611-2623 */
    unsigned lo = 0, hi = 63;
    while (lo < hi) {
        unsigned mid = (lo + hi) / 2;
        if (bits2f(T_DEC_SCALING[mid]) <= v) lo = mid + 1; else hi = mid;
    }
    return (int)lo;
}

static void header_length(hca_enc_t* e) { /* hca.cpp:2708-2750 */
    const hca_plan_t* p = &e->p;
    for (unsigned c = 0; c < p->channels; c++) {
        enc_ch_t* ch = &e->ch[c];
        int empty = 1;
        for (unsigned i = 0; i < p->coded[c]; i++) if (ch->sf[i]) { empty = 0; break; }
        if (empty) { ch->header_bits = 3; ch->delta_bits = 0; }
        else {
            int best_bits = 6, best_len = 3 + 6 * (int)p->coded[c];
            for (int db = 1; db < 6; db++) {
                int maxd = (1 << (db - 1)) - 1, len = 3 + 6;
                for (unsigned b = 1; b < p->coded[c]; b++) {
                    int delta = (int)ch->sf[b] - (int)ch->sf[b - 1];
                    len += abs(delta) > maxd ? db + 6 : db;
                }
                if (len < best_len) { best_len = len; best_bits = db; }
            }
            ch->header_bits = best_len; ch->delta_bits = best_bits;
        }
        if (p->type[c] == 2) ch->header_bits += 32;
        else if (p->hfr_groups > 0) ch->header_bits += 6 * (int)p->hfr_groups;
    }
}

static int enc_resolution(int sf, int noise) { /* hca.cpp:2752-2761 */
    if (sf == 0) return 0;
    int pos = noise - 5 * sf / 2 + 2;
    if (pos < 0) pos = 0; else if (pos > 58) pos = 58;
    return T_ENC_CURVE[pos];
}

static int used_bits(const hca_enc_t* e, int noise_level, int boundary) { /* hca.cpp:2763-2790 */
    const hca_plan_t* p = &e->p;
    int len = 48;
    for (unsigned c = 0; c < p->channels; c++) {
        const enc_ch_t* ch = &e->ch[c];
        len += ch->header_bits;
        for (int i = 0; i < (int)p->coded[c]; i++) {
            int noise = i < boundary ? noise_level - 1 : noise_level;
            int r = enc_resolution(ch->sf[i], noise);
            if (r >= 8) {
                int bits = T_MAX_BITS[r] - 1;
                float dz = bits2f(T_ENC_DEAD_ZONE[r]);
                for (int j = 0; j < 8; j++) { len += bits; if (fabsf(ch->scaled[i][j]) >= dz) len++; }
            } else {
                float inv = bits2f(T_ENC_INV_STEP[r]);
                float up = inv + 1;
                int down = (int)(inv + 0.5 - 8);
                for (int j = 0; j < 8; j++) {
                    float prod = ch->scaled[i][j] * inv;
                    float sum = prod + up;
                    int q = (int)sum - down;
                    len += T_ENC_QBITS[r * 16 + q];
                }
            }
        }
    }
    return len;
}

static int search_level(const hca_enc_t* e, int avail, int lo, int hi) { /* hca.cpp:2792-2807 */
    int max = hi, mid_value = 0;
    while (lo != hi) {
        int mid = (lo + hi) / 2;
        mid_value = used_bits(e, mid, 0);
        if (mid_value > avail) lo = mid + 1; else hi = mid;
    }
    return lo == max && mid_value > avail ? -1 : lo;
}

static int search_boundary(const hca_enc_t* e, int avail, int level, int lo, int hi) { /* hca.cpp:2834-2850 */
    int max = hi;
    while (abs(hi - lo) > 1) {
        int mid = (lo + hi) / 2;
        int v = used_bits(e, level, mid);
        if (avail < v) hi = mid - 1; else lo = mid;
    }
    if (lo == hi) return lo < max ? lo : -1;
    return used_bits(e, level, hi) > avail ? lo : hi;
}

/* EncodeFrame, hca.cpp:2965-2988. pcm: 1024*channels interleaved. 0 ok, <0 error. */
static int hca_encode_frame(hca_enc_t* e, const int16_t* pcm, uint8_t* out) {
    const hca_plan_t* p = &e->p;
    unsigned nch = p->channels;
    for (unsigned c = 0; c < nch; c++) {                       /* PcmToFloat :2470 + RunMdct :2555 */
        enc_ch_t* ch = &e->ch[c];
        for (int s = 0; s < 1024; s++) ch->wave[s >> 7][s & 127] = (float)pcm[(size_t)s * nch + c] * (1.0f / 32768.0f);
        for (int sub = 0; sub < 8; sub++) mdct128(ch, sub);
    }
    if (p->stereo_bands > 0)                                   /* EncodeIntensityStereo :2561 */
        for (unsigned c = 0; c < nch; c++) {
            if (p->type[c] != 1) continue;
            for (int sub = 0; sub < 8; sub++) {
                float* l = e->ch[c].spectra[sub]; float* r = e->ch[c + 1].spectra[sub];
                float el = 0, er = 0, et = 0;
                for (unsigned b = p->base_bands; b < p->total_bands; b++) {
                    el = el + fabsf(l[b]); er = er + fabsf(r[b]);
                    float s = l[b] + r[b];
                    et = et + fabsf(s);
                }
                et = et * 2;
                float elr = er + el;
                float twice = 2 * el;
                float stored = twice / elr;
                float ratio = elr / et;
                if ((double)ratio < 0.5) ratio = 0.5f;
                else if ((double)ratio > sqrt(2) / 2) ratio = (float)(sqrt(2) / 2);
                int q = 1;
                if (er > 0 || el > 0) { while (q < 13 && bits2f(T_ENC_RATIO_BOUNDS[q]) >= stored) q++; }
                else { q = 0; ratio = 1; }
                e->ch[c + 1].intensity[sub] = (uint8_t)q;
                for (unsigned b = p->base_bands; b < p->total_bands; b++) {
                    float s = l[b] + r[b];
                    l[b] = s * ratio; r[b] = 0;
                }
            }
        }
    for (unsigned c = 0; c < nch; c++) {                       /* CalculateScaleFactors :2625, ScaleSpectra :2639 */
        enc_ch_t* ch = &e->ch[c];
        for (unsigned b = 0; b < p->coded[c]; b++) {
            float mx = 0;
            for (int sub = 0; sub < 8; sub++) { float v = fabsf(ch->spectra[sub][b]); if (mx < v) mx = v; }
            ch->sf[b] = (uint8_t)find_scalefactor(mx);
        }
        memset(ch->sf + p->coded[c], 0, 128 - p->coded[c]);
        for (unsigned b = 0; b < p->coded[c]; b++) {
            int sf = ch->sf[b];
            float k = bits2f(T_ENC_Q_SCALING[sf]);
            for (int sub = 0; sub < 8; sub++) {
                float v = ch->spectra[sub][b] * k;
                if (v > 0.9999999f) v = 0.9999999f; else if (v < -0.9999999f) v = -0.9999999f;
                ch->scaled[b][sub] = sf == 0 ? 0 : v;
            }
        }
    }
    if (p->hfr_groups > 0) {                                   /* CalculateHfrGroupAverages :2656, CalculateHfrScale :2676 */
        int start = (int)(p->stereo_bands + p->base_bands);
        unsigned lim = p->hfr_band_count < p->total_bands - p->hfr_band_count ? p->hfr_band_count : p->total_bands - p->hfr_band_count;
        for (unsigned c = 0; c < nch; c++) {
            enc_ch_t* ch = &e->ch[c];
            if (p->type[c] == 2) continue;
            int band = start;
            for (unsigned g = 0; g < p->hfr_groups; g++) {
                float sum = 0; int count = 0;
                for (unsigned i = 0; i < p->bands_per_hfr && band < 128; band++, i++) {
                    for (int sub = 0; sub < 8; sub++) sum = sum + fabsf(ch->spectra[sub][band]);
                    count += 8;
                }
                ch->group_avg[g] = sum / (float)count;
            }
            unsigned lb = 0;
            for (unsigned g = 0; g < p->hfr_groups; g++) {
                float sum = 0; int count = 0;
                for (unsigned i = 0; i < p->bands_per_hfr && lb < lim; lb++, i++) {
                    for (int sub = 0; sub < 8; sub++) sum = sum + fabsf(ch->scaled[start - (int)lb - 1][sub]);
                    count += 8;
                }
                float avg = sum / (float)count;
                if (avg > 0.0) {
                    double m = 1.0 / (double)avg, s2 = sqrt(2);
                    if (s2 < m) m = s2;
                    ch->group_avg[g] = (float)((double)ch->group_avg[g] * m);
                }
                ch->hfr_scale[g] = find_scalefactor(ch->group_avg[g]);
            }
        }
    }
    header_length(e);
    int avail = (int)p->frame_size * 8;
    {   /* CalculateNoiseLevel :2809-2832 */
        int highest = (int)(p->base_bands + p->stereo_bands) - 1;
        int level = search_level(e, avail, 0, 255);
        while (level < 0) {
            highest -= 2;
            if (highest < 0) return -3;
            for (unsigned c = 0; c < nch; c++) { e->ch[c].sf[highest + 1] = 0; e->ch[c].sf[highest + 2] = 0; }
            header_length(e);
            level = search_level(e, avail, 0, 255);
        }
        e->noise_level = level;
    }
    if (e->noise_level == 0) e->boundary = 0;                  /* CalculateEvaluationBoundary :2852 */
    else {
        int b = search_boundary(e, avail, e->noise_level, 0, 127);
        if (b < 0) return -4;
        e->boundary = b;
    }
    for (unsigned c = 0; c < nch; c++) {                       /* resolutions :2868, quantise :2878 */
        enc_ch_t* ch = &e->ch[c];
        for (int i = 0; i < (int)p->coded[c]; i++)
            ch->res[i] = (uint8_t)enc_resolution(ch->sf[i], i < e->boundary ? e->noise_level - 1 : e->noise_level);
        memset(ch->res + p->coded[c], 0, 128 - p->coded[c]);
        for (unsigned i = 0; i < p->coded[c]; i++) {
            float inv = bits2f(T_ENC_INV_STEP[ch->res[i]]);
            float up = inv + 1;
            int down = (int)(inv + 0.5);
            for (int sub = 0; sub < 8; sub++) {
                float prod = ch->scaled[i][sub] * inv;
                float sum = prod + up;
                ch->quant[sub][i] = (int)sum - down;
            }
        }
    }
    /* PackFrame :2938-2963 */
    memset(out, 0, p->frame_size);
    put_be16(out, 0xFFFF);
    bitw_t bw = { out + 2, (size_t)(p->frame_size - 2) * 8, 0 };
    bitw_put(&bw, (uint32_t)e->noise_level, 9);
    bitw_put(&bw, (uint32_t)e->boundary, 7);
    for (unsigned c = 0; c < nch; c++) {
        enc_ch_t* ch = &e->ch[c];
        int db = ch->delta_bits;
        bitw_put(&bw, (uint32_t)db, 3);
        if (db == 6) { for (unsigned i = 0; i < p->coded[c]; i++) bitw_put(&bw, ch->sf[i], 6); }
        else if (db != 0) {
            bitw_put(&bw, ch->sf[0], 6);
            int maxd = (1 << (db - 1)) - 1, esc = (1 << db) - 1;
            for (unsigned i = 1; i < p->coded[c]; i++) {
                int delta = (int)ch->sf[i] - (int)ch->sf[i - 1];
                if (abs(delta) > maxd) { bitw_put(&bw, (uint32_t)esc, db); bitw_put(&bw, ch->sf[i], 6); }
                else bitw_put(&bw, (uint32_t)(maxd + delta), db);
            }
        }
        if (p->type[c] == 2) { for (int i = 0; i < 8; i++) bitw_put(&bw, ch->intensity[i], 4); }
        else if (p->hfr_groups > 0) { for (unsigned i = 0; i < p->hfr_groups; i++) bitw_put(&bw, (uint32_t)ch->hfr_scale[i] & 0x3F, 6); }
    }
    for (int sub = 0; sub < 8; sub++)
        for (unsigned c = 0; c < nch; c++) {
            enc_ch_t* ch = &e->ch[c];
            for (unsigned i = 0; i < p->coded[c]; i++) {
                int r = ch->res[i], q = ch->quant[sub][i];
                if (r == 0) continue;
                if (r < 8) bitw_put(&bw, T_ENC_QCODE[r * 16 + q + 8], T_ENC_QBITS[r * 16 + q + 8]);
                else {
                    bitw_put(&bw, (uint32_t)abs(q) & ((1u << (T_MAX_BITS[r] - 1)) - 1), T_MAX_BITS[r] - 1);
                    if (q != 0) bitw_put(&bw, q > 0 ? 0 : 1, 1);
                }
            }
        }
    put_be16(out + p->frame_size - 2, cri_oracle_crc16(out, p->frame_size - 2));
    return 0;
}

API int cri_oracle_hca_encoded_size(const uint8_t* wav, size_t n, unsigned quality, size_t* out_n) {
    wav_t w; hca_plan_t p;
    int r = wav_parse(wav, n, &w);
    if (r < 0) return -100 + r;
    if (hca_plan((unsigned)w.channels, (unsigned)w.rate, w.total_samples / (unsigned)w.channels, quality, &p) < 0) return -3;
    *out_n = (size_t)p.header_size + (size_t)p.frame_count * p.frame_size;
    return 0;
}

/* HcaEncode binding hca.cpp:3459-3489 + Encode :3072 + PackHeader :3109. Looping WAVs: -99. */
API int cri_oracle_hca_encode(const uint8_t* wav, size_t n, unsigned quality, uint8_t* out, size_t out_cap, size_t* out_n) {
    wav_t w;
    int r = wav_parse(wav, n, &w);
    if (r < 0) return -100 + r;
    if (w.looping) return -99;
    hca_enc_t* e = (hca_enc_t*)calloc(1, sizeof *e);
    if (!e) return -100;
    if (hca_plan((unsigned)w.channels, (unsigned)w.rate, w.total_samples / (unsigned)w.channels, quality, &e->p) < 0) { free(e); return -3; }
    const hca_plan_t* p = &e->p;
    size_t total = (size_t)p->header_size + (size_t)p->frame_count * p->frame_size;
    *out_n = total;
    if (out_cap < total) { free(e); return -100; }
    memset(out, 0, total);
    int16_t* frame = (int16_t*)malloc(sizeof(int16_t) * 1024 * p->channels);
    int rc = 0;
    for (unsigned f = 0; f < p->frame_count && rc == 0; f++) {
        for (unsigned s = 0; s < 1024; s++) {
            size_t idx = (size_t)f * 1024 + s;
            for (unsigned c = 0; c < p->channels; c++)
                frame[s * p->channels + c] = idx < p->samples ? w.pcm[idx * p->channels + c] : 0;
        }
        if (hca_encode_frame(e, frame, out + p->header_size + (size_t)f * p->frame_size) < 0) rc = -4;
    }
    free(frame);
    if (rc == 0) { /* PackHeader */
        uint8_t* hd = out;
        put_be32(hd, 0x48434100u); put_be16(hd + 4, 0x0200); put_be16(hd + 6, p->header_size);
        put_be32(hd + 8, 0x666D7400u); put_be32(hd + 12, p->rate); hd[12] = (uint8_t)p->channels;
        put_be32(hd + 16, p->frame_count); put_be16(hd + 20, p->delay); put_be16(hd + 22, p->padding);
        put_be32(hd + 24, 0x636F6D70u); put_be16(hd + 28, p->frame_size); hd[30] = 1; hd[31] = 15; hd[32] = 1;
        hd[33] = (uint8_t)p->channel_config; hd[34] = (uint8_t)p->total_bands; hd[35] = (uint8_t)p->base_bands;
        hd[36] = (uint8_t)p->stereo_bands; hd[37] = (uint8_t)p->bands_per_hfr;
        put_be32(hd + 40, 0x63697068u); put_be16(hd + 44, 0);
        put_be32(hd + 46, 0x70616400u);
        put_be16(hd + p->header_size - 2, cri_oracle_crc16(hd, p->header_size - 2));
    }
    free(e);
    return rc;
}
