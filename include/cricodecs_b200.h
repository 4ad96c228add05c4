/* cricodecs_b200 -- C-ABI of the B200-native batch engine for the CRI ADX / HCA
 * codec hot paths of Youjose/PyCriCodecs.
 *
 * This is the drop-in boundary: plain pointers and sizes, no CUDA / torch /
 * Python types. Each entry point names the reference interface it replaces
 * (file:line into the reference tree, CriCodecs/ = the C++ extension module,
 * PyCriCodecs/ = its Python front-end). INTEGRATION.md shows the binding a
 * maintainer of the reference would add on top of it.
 *
 * Every function returns 0 on success or a negative status:
 *    -1 ..  -9   ADX decode errors, same numbers as AdxErrorCode (adx.cpp:11-30)
 *   -10 .. -18   ADX encode parameter errors, same numbers (adx.cpp:424-442)
 *  -101 .. -110  WAV ingest errors = -100 + pcm.cpp code (pcm.cpp:22-33)
 *  -201 .. -204  HCA errors = -200 + py_codec_err code (hca.cpp:3252-3268):
 *                header / frame decode (wrong key) / channel config / encode
 *  -300          valid input this build does not handle (WAVs with several
 *                sampler loops; HCA v3.0 layouts whose derived HFR scales
 *                depend on the previous frame)
 *  -301          truncated input or output buffer too small
 *  -400          CUDA failure (no device, launch or copy error); the library
 *                never falls back to a CPU path
 * In batch calls the return value reports call-level failures only; per-stream
 * results are written to `status[i]` and a failed stream leaves its output
 * region zero-filled.
 *
 * Batch layout ("blob + offsets"): stream i occupies bytes
 * [offsets[i], offsets[i+1]) of `blob` (offsets has n+1 entries). Output blobs
 * use the same convention; *_sizes() fills the exact output size of every
 * stream so the caller can allocate once (sizes are computable from headers).
 *
 * Threading: a cri_ctx owns one CUDA stream and its scratch buffers; use one
 * context per host thread. Nothing is global.
 */
#ifndef CRICODECS_B200_H
#define CRICODECS_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define CRI_API __attribute__((visibility("default")))
#else
#define CRI_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cri_ctx cri_ctx;

/* -- context ------------------------------------------------------------ */
CRI_API int cri_ctx_create(int device, cri_ctx** ctx);   /* -400 when no usable CUDA device */
CRI_API void cri_ctx_destroy(cri_ctx* ctx);
CRI_API const char* cri_last_error(const cri_ctx* ctx);  /* text of the last -400 */
CRI_API int cri_version(void);                           /* 0xMMmmpp */
/* Number of engine kernels launched through this context so far (bench's gpu_launches). */
CRI_API uint64_t cri_ctx_launch_count(const cri_ctx* ctx);
/* Device time (ms, CUDA events on the context's stream) of the kernels of the
 * last *_run / *_batch call, and of the dominant kernel alone. */
CRI_API float cri_ctx_last_kernel_ms(const cri_ctx* ctx);
CRI_API float cri_ctx_last_dominant_ms(const cri_ctx* ctx);
/* A context keeps the HBM blocks of finished calls for the next call (sizes
 * repeat from batch to batch); cri_ctx_trim returns the idle ones to the driver. */
CRI_API void cri_ctx_trim(cri_ctx* ctx);
/* Page-locked host buffers. The *_batch calls accept any host pointer; with
 * buffers from cri_host_alloc the copies run at PCIe speed and overlap the
 * kernels (a batch is pipelined in chunks of whole streams). */
CRI_API void* cri_host_alloc(size_t bytes);
CRI_API void cri_host_free(void* p);

/* -- ADX decode: replaces CriCodecs.AdxDecode (adx.cpp:546-558), i.e.
 *    ADX::Decode + ChannelFrame::Decode (adx.cpp:380-415, 189-214) ---------- */
CRI_API int cri_adx_decode_sizes(const uint8_t* blob, const uint64_t* offsets, uint32_t n,
                         uint64_t* out_sizes, int32_t* status);
CRI_API int cri_adx_decode_batch(cri_ctx* ctx, const uint8_t* blob, const uint64_t* offsets, uint32_t n,
                         uint8_t* out_blob, const uint64_t* out_offsets, int32_t* status);

/* -- ADX encode: replaces CriCodecs.AdxEncode (adx.cpp:517-544), i.e.
 *    ADX::Encode + ChannelFrame::Encode (adx.cpp:416-506, 215-273). Argument
 *    order follows the Python wrapper (PyCriCodecs/adx.py:12-14). ------------ */
typedef struct cri_adx_params {
    uint32_t bit_depth;   /* 4  */
    uint32_t block_size;  /* 18 */
    uint32_t encoding;    /* 3 (2 fixed, 3 linear, 4 exponential) */
    uint32_t highpass;    /* 500 */
    uint32_t filter;      /* 0 */
    uint32_t version;     /* 4 */
    uint32_t force_not_looping;
} cri_adx_params;
CRI_API int cri_adx_encode_sizes(const uint8_t* blob, const uint64_t* offsets, uint32_t n, const cri_adx_params* p,
                         uint64_t* out_sizes, int32_t* status);
CRI_API int cri_adx_encode_batch(cri_ctx* ctx, const uint8_t* blob, const uint64_t* offsets, uint32_t n,
                         const cri_adx_params* p, uint8_t* out_blob, const uint64_t* out_offsets, int32_t* status);

/* -- HCA decode: replaces CriCodecs.HcaDecode (hca.cpp:3340-3457): header
 *    parse, key setup, clHCA_DecodeBlock per frame (hca.cpp:1149-1254),
 *    clHCA_ReadSamples16 (hca.cpp:339-360), WAV image. keys / subkeys may be
 *    NULL (all zero) or hold one entry per stream. --------------------------- */
CRI_API int cri_hca_decode_sizes(const uint8_t* blob, const uint64_t* offsets, uint32_t n,
                         uint64_t* out_sizes, int32_t* status);
CRI_API int cri_hca_decode_batch(cri_ctx* ctx, const uint8_t* blob, const uint64_t* offsets, uint32_t n,
                         const uint64_t* keys, const uint16_t* subkeys,
                         uint8_t* out_blob, const uint64_t* out_offsets, int32_t* status);

/* -- HCA crypt: replaces CriCodecs.HcaCrypt (hca.cpp:3271-3337) + CryptHeader
 *    (hca.cpp:3166-3250). encrypt = 0 decrypts with the stream's own ciph type;
 *    encrypt = 1 uses `ciph_type` (56 keyed, 1 keyless). Output size == input
 *    size; the input is never modified (the reference mutates it in place). -- */
CRI_API int cri_hca_crypt_batch(cri_ctx* ctx, const uint8_t* blob, const uint64_t* offsets, uint32_t n,
                        int encrypt, uint32_t ciph_type, const uint64_t* keys, const uint16_t* subkeys,
                        uint8_t* out_blob, int32_t* status);

/* -- HCA encode: replaces CriCodecs.HcaEncode (hca.cpp:3459-3489): planning
 *    (hca.cpp:2206-2462), EncodeFrame (hca.cpp:2965-2988), PackHeader
 *    (hca.cpp:3109-3164). quality: 0 Highest .. 3 Low, 4 Lowest (the Python
 *    enum's Lowest = 5 falls back to High, chunk.py:73). ---------------------- */
CRI_API int cri_hca_encode_sizes(const uint8_t* blob, const uint64_t* offsets, uint32_t n, uint32_t quality,
                         uint64_t* out_sizes, int32_t* status);
/*    A WAV with a sampler loop (smpl chunk) encodes a loop chunk plus pre / post audio frames unless
 *    force_not_looping is set (hca.cpp:2440-2449, 3469): its size depends on the flag. cri_hca_encode_sizes
 *    answers for force_not_looping = 0, the default of HCA.encode (hca.py:255). */
CRI_API int cri_hca_encode_sizes_ex(const uint8_t* blob, const uint64_t* offsets, uint32_t n, uint32_t quality,
                            uint32_t force_not_looping, uint64_t* out_sizes, int32_t* status);
CRI_API int cri_hca_encode_batch(cri_ctx* ctx, const uint8_t* blob, const uint64_t* offsets, uint32_t n,
                         uint32_t quality, uint32_t force_not_looping,
                         uint8_t* out_blob, const uint64_t* out_offsets, int32_t* status);

/* -- Device-resident execution (what bench.py's `value` times): a job keeps
 *    the parsed headers, launch tables and both blobs in HBM; *_run launches
 *    only kernels on the context's stream and returns after they finish. ---- */
typedef struct cri_job cri_job;
enum { CRI_JOB_ADX_DECODE = 1, CRI_JOB_ADX_ENCODE = 2, CRI_JOB_HCA_DECODE = 3, CRI_JOB_HCA_CRYPT = 4, CRI_JOB_HCA_ENCODE = 5 };
typedef struct cri_job_desc {
    int kind;                    /* CRI_JOB_* */
    const uint8_t* blob;         /* host input blob */
    const uint64_t* offsets;     /* n + 1 */
    uint32_t n;
    const uint64_t* keys;        /* HCA decode / crypt, may be NULL */
    const uint16_t* subkeys;     /* may be NULL */
    cri_adx_params adx;          /* ADX encode */
    uint32_t quality;            /* HCA encode */
    int encrypt;                 /* HCA crypt */
    uint32_t ciph_type;          /* HCA crypt */
    /* Device-resident I/O (optional, all NULL for host blobs): `d_blob` is the input blob in the GPU's memory
     * (`blob` is then ignored), `d_out` receives the packed output blob in place of a library-owned buffer, and
     * `stream` (a cudaStream_t) orders every copy and kernel of the job. */
    const uint8_t* d_blob;
    uint8_t* d_out;
    void* stream;
} cri_job_desc;
CRI_API int cri_job_create(cri_ctx* ctx, const cri_job_desc* desc, cri_job** job); /* parses, plans, uploads the input */
CRI_API uint64_t cri_job_out_bytes(const cri_job* job);
CRI_API const uint64_t* cri_job_out_offsets(const cri_job* job);                    /* n + 1 */
CRI_API uint64_t cri_job_units(const cri_job* job);   /* HCA frames (all channels) or ADX blocks (per channel) processed by one run */
CRI_API int cri_job_upload(cri_ctx* ctx, cri_job* job);                             /* host blob -> HBM again (H2D only) */
CRI_API int cri_job_run(cri_ctx* ctx, cri_job* job);                                /* kernels only, inputs resident */
CRI_API int cri_job_download(cri_ctx* ctx, cri_job* job, uint8_t* out_blob, int32_t* status); /* HBM -> host (D2H) + status */
CRI_API void cri_job_destroy(cri_ctx* ctx, cri_job* job);

/* -- Device-pointer batch calls: SURVEY.md section 8(b) item (2), `..._batch(..., cudaStream_t)`. Same meaning
 *    as the host-buffer calls above, but `d_blob` and `d_out` are DEVICE pointers on the context's GPU and every
 *    copy and kernel is ordered on `stream` (a cudaStream_t passed as void*; NULL = the context's own stream), so a
 *    GPU-resident producer / consumer never crosses PCIe with the payload. `offsets`, `out_offsets`, keys and
 *    `status` are HOST arrays. `out_offsets` must be the packed layout that *_sizes() describes (-301 otherwise);
 *    the *_sizes_dev variants read the headers from the device blob. The calls fetch the stream headers (a few
 *    hundred bytes per stream) for planning, enqueue the work and return after `stream` has finished it. ---------- */
CRI_API int cri_sizes_dev(cri_ctx* ctx, int job_kind, const uint8_t* d_blob, const uint64_t* offsets, uint32_t n,
                          const cri_adx_params* adx, uint32_t quality, uint64_t* out_sizes, int32_t* status, void* stream);
CRI_API int cri_adx_decode_batch_dev(cri_ctx* ctx, const uint8_t* d_blob, const uint64_t* offsets, uint32_t n,
                             uint8_t* d_out, const uint64_t* out_offsets, int32_t* status, void* stream);
CRI_API int cri_adx_encode_batch_dev(cri_ctx* ctx, const uint8_t* d_blob, const uint64_t* offsets, uint32_t n,
                             const cri_adx_params* p, uint8_t* d_out, const uint64_t* out_offsets, int32_t* status, void* stream);
CRI_API int cri_hca_decode_batch_dev(cri_ctx* ctx, const uint8_t* d_blob, const uint64_t* offsets, uint32_t n,
                             const uint64_t* keys, const uint16_t* subkeys,
                             uint8_t* d_out, const uint64_t* out_offsets, int32_t* status, void* stream);
CRI_API int cri_hca_crypt_batch_dev(cri_ctx* ctx, const uint8_t* d_blob, const uint64_t* offsets, uint32_t n,
                            int encrypt, uint32_t ciph_type, const uint64_t* keys, const uint16_t* subkeys,
                            uint8_t* d_out, int32_t* status, void* stream);
CRI_API int cri_hca_encode_batch_dev(cri_ctx* ctx, const uint8_t* d_blob, const uint64_t* offsets, uint32_t n,
                             uint32_t quality, uint32_t force_not_looping,
                             uint8_t* d_out, const uint64_t* out_offsets, int32_t* status, void* stream);

/* -- Single-stream conveniences with library-owned output (cri_free). These
 *    are what a 1:1 replacement of the five CriCodecs callables binds. ------- */
CRI_API int cri_adx_decode(cri_ctx* ctx, const uint8_t* in, size_t n, uint8_t** out, size_t* out_n);
CRI_API int cri_adx_encode(cri_ctx* ctx, const uint8_t* in, size_t n, const cri_adx_params* p, uint8_t** out, size_t* out_n);
CRI_API int cri_hca_decode(cri_ctx* ctx, const uint8_t* in, size_t n, uint64_t key, uint16_t subkey, uint8_t** out, size_t* out_n);
CRI_API int cri_hca_crypt(cri_ctx* ctx, const uint8_t* in, size_t n, int encrypt, uint32_t ciph_type, uint64_t key, uint16_t subkey,
                  uint8_t** out, size_t* out_n);
CRI_API int cri_hca_encode(cri_ctx* ctx, const uint8_t* in, size_t n, uint32_t quality, uint32_t force_not_looping,
                   uint8_t** out, size_t* out_n);
CRI_API void cri_free(void* p);

/* -- Host-only helpers (no device needed) ---------------------------------- */
CRI_API uint16_t cri_crc16(const uint8_t* p, size_t n);                     /* hca.cpp:205-211 */
CRI_API int cri_hca_cipher_table(int ciph_type, uint64_t key, uint8_t table[256]); /* hca.cpp:499-617 */
CRI_API uint64_t cri_hca_mix_subkey(uint64_t key, uint16_t subkey);          /* hca.cpp:3309-3311 */
CRI_API void cri_adx_coefficients(uint32_t highpass, uint32_t rate, int32_t coef[2]); /* adx.cpp:58-64 */
CRI_API const char* cri_strerror(int status);                               /* the reference's exception text */

#ifdef __cplusplus
}
#endif
#endif /* CRICODECS_B200_H */
