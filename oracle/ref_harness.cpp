// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Glue that compiles the UNMODIFIED reference sources (found under
// $REF/CriCodecs at build time; nothing is copied into this repo) into one
// shared object, oracle/_ref/CriCodecs.*.so, which is at the same time
//   * the reference's own CPython module `CriCodecs` (PyInit_CriCodecs comes
//     from the included CriCodecs.cpp), and
//   * a ctypes-loadable library exporting the ref_* entry points below, which
//     reach the parts of the reference whose Python bindings are unusable:
//       - AdxEncode's "p" format writes an int into a bool and clobbers
//         blocksize (CriCodecs/adx.cpp:526-527)  -> ref_adx_encode
//       - HcaEncode reads an uninitialised stack clHCA
//         (CriCodecs/hca.cpp:3468, comment_len feeds header_size :2311)
//                                                -> ref_hca_encode
//     plus stage-level probes used to validate the C restatement.
//
// operator new[] is replaced by a zero-filling allocator (bound locally with
// -Wl,-Bsymbolic-functions) because BitWriter::Write ORs into its first byte
// (CriCodecs/IO.cpp:139,143,148) while ADX::Encode only clears the header
// (CriCodecs/adx.cpp:487-488): "bit-exact" is defined on zero-filled memory.
// For the same reason the reference's three malloc(sizeof(clHCA)) calls (hca.cpp:390,
// 3302, 3356) see zero-filled memory here: a v1.x header (`dec` chunk, hca.cpp:710-727)
// never sets ms_stereo, which clHCA_DecodeHeader then tests (hca.cpp:976-977), so on a
// recycled heap block the reference rejects or accepts such a stream at random.
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <cstdlib>
#include <new>

void* operator new[](std::size_t n) { void* p = std::calloc(1, n ? n : 1); if (!p) throw std::bad_alloc(); return p; }
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }

#define malloc(n) calloc(1, (n))
#include "CriCodecs.cpp"
#undef malloc

extern "C" {

// ADX::GetADX (adx.cpp:507-511) on a WAV image. Returns AdxErrorCode (0 ok).
int ref_adx_encode(unsigned char* wav, unsigned bitdepth, unsigned blocksize, unsigned mode,
                   unsigned highpass, unsigned filter, unsigned version, int force_noloop,
                   unsigned char* out, unsigned* out_n) {
    AdxErrorCode = 0;
    ADX adx;
    PCM pcm;
    char r = pcm.LoadDirect(wav);
    if (r < 0) return -100 + r;
    unsigned char* data = adx.GetADX(pcm, bitdepth, blocksize, mode, (unsigned short)highpass, filter, version, force_noloop != 0);
    if (AdxErrorCode) return AdxErrorCode;
    if (out) memcpy(out, data, adx.size);
    *out_n = adx.size;
    delete[] data;
    return 0;
}

// Mirrors the HcaEncode binding (hca.cpp:3467-3486) on a zeroed clHCA.
static clHCA g_enc;
int ref_hca_encode(unsigned char* wav, unsigned quality, unsigned force_noloop,
                   unsigned char* out, unsigned* out_n) {
    memset(&g_enc, 0, sizeof g_enc);
    HcaErrorCode = 0;
    PCM w;
    char r = w.LoadDirect(wav);
    if (r < 0) return -100 + r;
    g_enc.loop_flag = w.wav.chunks.Looping && !force_noloop;
    r = initHCAEncode(w, g_enc, (CriHcaQuality)quality);
    if (r < 0) return -3;
    unsigned total = g_enc.header_size + g_enc.frame_count * g_enc.frame_size;
    *out_n = total;
    if (!out) return 0;
    memset(out, 0, total);
    Encode(g_enc, w, out + g_enc.header_size);
    if (HcaErrorCode < 0) { int e = HcaErrorCode; HcaErrorCode = 0; return e < -4 ? -4 : -4; }
    PackHeader(g_enc, out);
    return 0;
}

// Decode frames [f0, f1) of an HCA image from a freshly reset decoder; writes
// PCM16 interleaved (1024*channels per frame). No delay trimming.
int ref_hca_decode_range(unsigned char* file, unsigned size, unsigned long long key,
                         unsigned f0, unsigned f1, short* pcm) {
    clHCA* h = clHCA_new();
    int hs = clHCA_isOurFile(file, size);
    if (hs < 0) { clHCA_delete(h); return hs; }
    h->keycode = key;
    int r = clHCA_DecodeHeader(h, file, hs);
    if (r < 0) { clHCA_delete(h); return r; }
    clHCA_SetKey(h, key);
    unsigned char* buf = (unsigned char*)malloc(h->frame_size);
    for (unsigned f = f0; f < f1; f++) {
        memcpy(buf, file + hs + (size_t)f * h->frame_size, h->frame_size);
        r = clHCA_DecodeBlock(h, buf, h->frame_size);
        if (r < 0) break;
        clHCA_ReadSamples16(h, pcm + (size_t)(f - f0) * 1024 * h->channels);
    }
    free(buf);
    clHCA_delete(h);
    return r < 0 ? r : 0;
}

// State after clHCA_DecodeBlock_unpack (hca.cpp:1149) of one frame.
int ref_hca_unpack_dump(unsigned char* file, unsigned size, unsigned long long key, unsigned frame,
                        unsigned char* sf, unsigned char* res, unsigned char* inten, float* gain,
                        float* spectra, int* bits) {
    clHCA* h = clHCA_new();
    int hs = clHCA_isOurFile(file, size);
    if (hs < 0) { clHCA_delete(h); return hs; }
    h->keycode = key;
    int r = clHCA_DecodeHeader(h, file, hs);
    if (r < 0) { clHCA_delete(h); return r; }
    clHCA_SetKey(h, key);
    unsigned char* buf = (unsigned char*)malloc(h->frame_size);
    memcpy(buf, file + hs + (size_t)frame * h->frame_size, h->frame_size);
    r = clHCA_DecodeBlock_unpack(h, buf, h->frame_size);
    *bits = r;
    for (unsigned c = 0; c < h->channels; c++) {
        memcpy(sf + 128 * c, h->channel[c].scalefactors, 128);
        memcpy(res + 128 * c, h->channel[c].resolution, 128);
        memcpy(inten + 8 * c, h->channel[c].intensity, 8);
        memcpy(gain + 128 * c, h->channel[c].gain, 128 * sizeof(float));
        memcpy(spectra + 1024 * c, h->channel[c].spectra, 1024 * sizeof(float));
    }
    free(buf);
    clHCA_delete(h);
    return r < 0 ? r : 0;
}

// imdct_transform (hca.cpp:1898) on a lone channel: spectra[128] in,
// prev[128] in/out, wave[128] and dct[128] out.
void ref_imdct(const float* spectra, float* prev, float* wave, float* dct) {
    static stChannel ch;
    memset(&ch, 0, sizeof ch);
    memcpy(ch.spectra[0], spectra, 128 * sizeof(float));
    memcpy(ch.imdct_previous, prev, 128 * sizeof(float));
    imdct_transform(&ch, 0);
    memcpy(wave, ch.wave[0], 128 * sizeof(float));
    memcpy(dct, ch.spectra[0], 128 * sizeof(float));
    memcpy(prev, ch.imdct_previous, 128 * sizeof(float));
}

// mdct_transform (hca.cpp:2529) on a lone channel.
void ref_mdct(const float* wave, float* prev, float* spectra) {
    static stChannel ch;
    memset(&ch, 0, sizeof ch);
    memcpy(ch.wave[0], wave, 128 * sizeof(float));
    memcpy(ch.imdct_previous, prev, 128 * sizeof(float));
    mdct_transform(ch, 0);
    memcpy(spectra, ch.spectra[0], 128 * sizeof(float));
    memcpy(prev, ch.imdct_previous, 128 * sizeof(float));
}

unsigned ref_crc16(const unsigned char* p, unsigned n) { return crc16_checksum(p, n); }

int ref_cipher_table(int type, unsigned long long key, unsigned char* table) {
    return cipher_init(table, type, key);
}

void ref_ath_curve(int type, unsigned sample_rate, unsigned char* curve) { ath_init(curve, type, sample_rate); }

void ref_adx_coefficients(unsigned highpass, unsigned rate, int* c) {
    int* p = c; unsigned short hp = (unsigned short)highpass;
    CalculateCoefficients(p, hp, rate);
}

// Copies a named static table out of the reference (used only to check the
// independently generated tables in tools/gen_tables.py). Returns byte count.
#define TBL(name, sym) if (!strcmp(which, name)) { if (out) memcpy(out, sym, sizeof(sym)); return (int)sizeof(sym); }
int ref_table(const char* which, void* out) {
    TBL("crc", hcacommon_crc_mask_table)
    TBL("ath_base", ath_base_curve)
    TBL("invert", hcadecoder_invert_table)
    TBL("dec_scaling", hcadequantizer_scaling_table_float_hex)
    TBL("dec_range", hcadequantizer_range_table_float_hex)
    TBL("max_bit", hcatbdecoder_max_bit_table)
    TBL("read_bit", hcatbdecoder_read_bit_table)
    TBL("read_val", hcatbdecoder_read_val_table)
    TBL("scale_conv", hcadecoder_scale_conversion_table_hex)
    TBL("intensity_ratio", hcadecoder_intensity_ratio_table_hex)
    TBL("imdct_sin", sin_tables_hex)
    TBL("imdct_cos", cos_tables_hex)
    TBL("window", hcaimdct_window_float_hex)
    TBL("enc_res_curve", ScaleToResolutionCurve)
    TBL("enc_inv_step", QuantizerInverseStepSize)
    TBL("enc_q_bits", QuantizeSpectrumBits)
    TBL("enc_q_value", QuantizeSpectrumValue)
    TBL("enc_ratio_bounds", IntensityRatioBoundsTableHex)
    TBL("enc_dead_zone", QuantizerDeadZoneHex)
    TBL("enc_shuffle", ShuffleTable)
    TBL("enc_q_scaling", QuantizerScalingTableHex)
    TBL("mdct_sin", SinTablesHex)
    TBL("mdct_cos", CosTablesHex)
    TBL("adx_static", StaticCoefficients)
    return -1;
}

}  // extern "C"
