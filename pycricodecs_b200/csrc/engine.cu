// C-ABI implementation (include/cricodecs_b200.h): contexts, jobs, host<->HBM
// movement. The per-sample work is in *_kernels.cu; the per-stream header work
// is in formats.cpp. There is deliberately no CPU fallback anywhere in this
// file: without a CUDA device every compute entry point returns -400.
#include <cuda_runtime.h>
#include <sys/mman.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/cricodecs_b200.h"
#include "engine.h"

using namespace cri;

// --------------------------------------------------------------- context
#define CU_TRY(ctx, expr)                                                                   \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            (ctx)->error = std::string(#expr) + ": " + cudaGetErrorString(e_);              \
            return ERR_CUDA;                                                                \
        }                                                                                   \
    } while (0)

extern "C" int cri_version(void) { return 0x000100; }

extern "C" int cri_ctx_create(int device, cri_ctx** out) {
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return ERR_CUDA;
    cri_ctx* c = new (std::nothrow) cri_ctx();
    if (!c) return ERR_BUFFER;
    c->device = device;
    bool ok = cudaSetDevice(device) == cudaSuccess && cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (auto& s : c->pipe) ok = ok && cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) {
        delete c;
        return ERR_CUDA;
    }
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    const char* t = getenv("CRI_TRACE");
    c->trace = t && *t && *t != '0';
    const char* pz = getenv("CRI_POISON");
    c->poison = pz && *pz && *pz != '0';
    *out = c;
    return OK;
}

extern "C" void cri_ctx_destroy(cri_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    pool_trim(c);
    for (auto& kv : c->pool.live) cudaFree(kv.first);   // blocks of jobs the caller never destroyed
    for (auto& s : c->pipe) cudaStreamDestroy(s);
    for (auto& p : c->pin_status) cudaFreeHost(p);
    for (auto& e : c->idle_events) cudaEventDestroy(e);
    if (c->pin_stage) cudaFreeHost(c->pin_stage);
    for (auto& sh : c->idle_shadows) munmap(sh.first, sh.second);
    if (c->hca_pipe.side) {
        cudaStreamDestroy(c->hca_pipe.side);
        cudaEventDestroy(c->hca_pipe.start);
        cudaEventDestroy(c->hca_pipe.done);
        for (auto& e : c->hca_pipe.unpacked) cudaEventDestroy(e);
    }
    cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" void cri_ctx_trim(cri_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    pool_trim(c);
}

// Pinned host buffers for callers that want the batch calls to run at PCIe speed
// (pageable memory works too, through the driver's staging copies).
extern "C" void* cri_host_alloc(size_t bytes) {
    void* p = nullptr;
    return cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? p : nullptr;
}
extern "C" void cri_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------ HBM block cache
namespace cri {
static size_t pool_round(size_t bytes) {
    const size_t g = bytes >= (size_t(1) << 20) ? (size_t(2) << 20) : 512;   // 2 MiB granules for large blocks
    return (std::max<size_t>(bytes, 1) + g - 1) / g * g;
}

int pool_alloc(cri_ctx* c, void** p, size_t bytes) {
    *p = nullptr;
    const size_t want = pool_round(bytes);
    DevPool& P = c->pool;
    auto it = P.idle.lower_bound(want);
    if (it != P.idle.end() && it->first <= want + want / 4) {      // close enough in size: reuse
        *p = it->second;
        P.live[*p] = it->first;
        P.idle_bytes -= it->first;
        P.idle.erase(it);
        return OK;
    }
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) {                                        // give the cache back to the driver and retry once
        cudaGetLastError();
        pool_trim(c);
        e = cudaMalloc(p, want);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        c->error = std::string("cudaMalloc: ") + cudaGetErrorString(e);
        *p = nullptr;
        return ERR_CUDA;
    }
    P.live[*p] = want;
    return OK;
}

void pool_free(cri_ctx* c, void* p) {
    if (!p) return;
    if (!c) { cudaFree(p); return; }
    DevPool& P = c->pool;
    auto it = P.live.find(p);
    if (it == P.live.end()) { cudaFree(p); return; }
    P.idle.emplace(it->second, p);
    P.idle_bytes += it->second;
    P.live.erase(it);
}

void pool_trim(cri_ctx* c) {
    for (auto& kv : c->pool.idle) cudaFree(kv.second);
    c->pool.idle.clear();
    c->pool.idle_bytes = 0;
}
}  // namespace cri

extern "C" const char* cri_last_error(const cri_ctx* c) { return c ? c->error.c_str() : "no context"; }
extern "C" uint64_t cri_ctx_launch_count(const cri_ctx* c) { return c ? c->launches : 0; }
extern "C" float cri_ctx_last_kernel_ms(const cri_ctx* c) { return c ? c->last_ms : 0.f; }
extern "C" float cri_ctx_last_dominant_ms(const cri_ctx* c) { return c ? c->last_dominant_ms : 0.f; }
extern "C" void cri_free(void* p) { free(p); }

// ------------------------------------------------------------ host helpers
extern "C" uint16_t cri_crc16(const uint8_t* p, size_t n) { return crc16(p, n); }
extern "C" int cri_hca_cipher_table(int type, uint64_t key, uint8_t t[256]) { return cipher_table(type, key, t) ? ERR_HCA_HEADER : OK; }
extern "C" uint64_t cri_hca_mix_subkey(uint64_t key, uint16_t subkey) { return mix_subkey(key, subkey); }
extern "C" void cri_adx_coefficients(uint32_t highpass, uint32_t rate, int32_t coef[2]) {
    int c[2];
    adx_coefficients(highpass, rate, c);
    coef[0] = c[0];
    coef[1] = c[1];
}

extern "C" const char* cri_strerror(int st) {
    static const char* adx[] = {  // adx.cpp:11-30
        "Invalid ADX file header.", "AHX file provided, unsopported.", "Encrypted ADX detected, unsupported.",
        "Invalid/Unknown encoding mode found.", "Unknown ADX version provided.", "Invalid Bitdepth found on the provided ADX.",
        "ADX does not contain any channels info.", "Invalid ADX header, loop information size is bigger than the header.",
        "Inavlid ADX header, Criware copyright string not found.", "Numbers of Channel cannot exceed 255 or go below 0.",
        "Bitdepth must be between 2 and 15 inclusive.", "Blocksize must be between 3 and 255 inclusive.",
        "EncodingMode must be either 2, 3, or 4.", "HighpassFrequency must be between 0 and 65535 inclusive.",
        "Filter is used with EncodingMode == 2 and must be between 0 and 4 inclusive.", "AdxVersion must be either 3, 4 or 5.",
        "Provided Bitdepth does not fit correctly with the provided BlockSize", "Given WAVE file is not valid for ADX encoding."};
    static const char* wav[] = {  // pcm.cpp:22-33
        "Invalid WAVE file header.", "Invalid WAVE file header. Format info is not present.",
        "Unsupported/Unknown WAVE compression mode.", "Invalid looping sample info data.",
        "Invalid looping sample info data, Number of loops/loop data is larger than the available size.",
        "Data tag is not present.", "Header is not valid.", "PCM Bitdepth does not match compression type.",
        "Filesize exceeds 2GB use python to load in with buffer.", "Filesize is too low to be viable for loading."};
    if (st == 0) return "ok";
    if (st <= -1 && st >= -18) return adx[-st - 1];
    if (st <= -101 && st >= -110) return wav[-st - 101];
    switch (st) {  // hca.cpp:3252-3268
        case ERR_HCA_HEADER: return "Header decoding error, the header is not a valid HCA header.";
        case ERR_HCA_DECODE: return "Decoding error, either an incorrect key or an unknown exception.";
        case ERR_HCA_CHANNELS: return "Error setting up channel configuration.";
        case ERR_HCA_ENCODE: return "Unknown Encoding error.";
        case ERR_UNSUPPORTED: return "Input is valid but not supported by this build of cricodecs_b200.";
        case ERR_BUFFER: return "Truncated input or output buffer too small.";
        case ERR_CUDA: return "CUDA failure (no device, or a launch/copy failed); there is no CPU fallback.";
    }
    return "unknown status";
}

// ---------------------------------------------------------------- planning
// Each planner fills job->status / out_sizes and the kind-specific launch
// tables on the host. Output layout is the packed concatenation of the exact
// per-stream output sizes (failed streams get size 0).

namespace cri {
void finish_layout_public(cri_job* j, const std::vector<uint64_t>& sizes) {
    j->out_off.assign(j->n + 1, j->out_delta);
    for (uint32_t i = 0; i < j->n; i++) j->out_off[i + 1] = j->out_off[i] + sizes[i];
    j->out_off_pub.resize(j->n + 1);
    for (uint32_t i = 0; i <= j->n; i++) j->out_off_pub[i] = j->out_off[i] - j->out_delta;
    j->out_bytes = j->out_off_pub[j->n];
}

void add_patch_public(cri_job* j, uint64_t dst, const uint8_t* bytes, uint32_t n) {
    Patch p{dst, (uint32_t)j->patch_bytes.size(), n};
    j->patch_bytes.insert(j->patch_bytes.end(), bytes, bytes + n);
    j->patches.push_back(p);
}
}  // namespace cri
static void finish_layout(cri_job* j, const std::vector<uint64_t>& sizes) { finish_layout_public(j, sizes); }
static void add_patch(cri_job* j, uint64_t dst, const uint8_t* bytes, uint32_t n) { add_patch_public(j, dst, bytes, n); }

// Fast-path chain lists: a warp's 32 chains belong to 32/nch whole streams of one channel count (1,2,4,..,32); each
// channel-count bucket is padded with idle chains to whole warps. Other channel counts go to the generic list.
struct AdxLists {
    std::vector<AdxChain> bucket[6], generic;
    static int bucket_of(int nch) { for (int b = 0; b < 6; b++) if (nch == (1 << b)) return b; return -1; }
    void add(const std::vector<AdxChain>& stream_chains, bool fast_ok) {
        const int b = fast_ok ? bucket_of((int)stream_chains.size()) : -1;
        auto& dst = b >= 0 ? bucket[b] : generic;
        dst.insert(dst.end(), stream_chains.begin(), stream_chains.end());
    }
    void finish(cri_job* j) {
        j->adx_chains.clear();
        for (int b = 0; b < 6; b++) {
            auto& v = bucket[b];
            while (v.size() % 32) { AdxChain idle{}; idle.channels = (uint8_t)(1 << b); v.push_back(idle); }
            j->adx_chains.insert(j->adx_chains.end(), v.begin(), v.end());
        }
        j->n_fast = (uint32_t)j->adx_chains.size();
        j->n_generic = (uint32_t)generic.size();
        j->adx_chains.insert(j->adx_chains.end(), generic.begin(), generic.end());
    }
};

// Shared by cri_adx_decode_sizes and the planner. A header may promise any sample count; the decoder clips to the blocks
// that are there and zero-fills the rest, but a stream that promises more than twice what its payload can hold (plus a
// block) is refused as truncated rather than answered with gigabytes of silence.
static int adx_decode_size_one(const uint8_t* d, size_t n, AdxInfo* a, uint64_t* size) {
    const int r = parse_adx(d, n, a);
    if (r < 0) return r;
    const uint64_t data = (uint64_t)a->data_offset + 4, frame_bytes = (uint64_t)a->channels * a->block_size;
    const uint64_t avail = n > data && frame_bytes ? (n - data) / frame_bytes : 0;
    if ((uint64_t)a->samples > 2 * (avail + 1) * a->samples_per_block) return ERR_BUFFER;
    *size = wav_header_size(a->looping) + (uint64_t)a->samples * a->channels * 2;
    return OK;
}

static void plan_adx_decode(cri_job* j) {
    std::vector<uint64_t> sizes(j->n, 0);
    std::vector<AdxInfo> infos(j->n);
    parallel_for(j->n, [&](uint32_t i) {
        j->status[i] = adx_decode_size_one(j->blob + j->in_off[i], j->in_off[i + 1] - j->in_off[i], &infos[i], &sizes[i]);
    });
    finish_layout(j, sizes);
    AdxLists lists;
    std::vector<AdxChain> one;
    for (uint32_t i = 0; i < j->n; i++) {
        if (j->status[i] != OK) continue;
        const AdxInfo& a = infos[i];
        one.clear();
        const uint64_t len = j->in_off[i + 1] - j->in_off[i];
        const uint64_t data = (uint64_t)a.data_offset + 4;
        const uint64_t frame_bytes = (uint64_t)a.channels * a.block_size;
        uint64_t avail = len > data ? (len - data) / frame_bytes : 0;  // never read past the caller's buffer
        const uint32_t blocks = (uint32_t)std::min<uint64_t>(a.blocks, avail);
        const size_t hdr = wav_header_size(a.looping);
        uint8_t h[0x70];
        write_wav_header(h, a.samples, a.channels, (int)a.rate, a.looping, a.loop_start, a.loop_end);
        add_patch(j, j->out_off[i], h, (uint32_t)hdr);
        const uint64_t pcm0 = j->out_off[i] + hdr;
        const bool aligned = (pcm0 & 1) == 0;
        const bool is_fast = aligned && a.bit_depth == 4 && a.block_size == 18;
        for (int c = 0; c < a.channels; c++) {
            AdxChain ch{};
            ch.eof_off = j->in_off[i] + data;
            ch.in_off = ch.eof_off + (uint64_t)c * a.block_size;
            ch.out_off = pcm0 + (uint64_t)c * 2;
            ch.blocks = blocks;
            ch.samples = a.samples;
            ch.in_stride = (uint32_t)frame_bytes;
            ch.out_stride = (uint32_t)a.channels;
            ch.coef0 = a.coef[0];
            ch.coef1 = a.coef[1];
            ch.hist1 = a.history[c][0];
            ch.hist2 = a.history[c][1];
            ch.mode = (uint8_t)a.mode;
            ch.bit_depth = (uint8_t)a.bit_depth;
            ch.block_size = (uint8_t)a.block_size;
            ch.channels = (uint8_t)a.channels;
            ch.channel = (uint8_t)c;
            ch.stream = i;
            one.push_back(ch);
        }
        lists.add(one, is_fast);
        j->units += (uint64_t)blocks * a.channels;
        // samples past the last decoded block stay zero: zero-filled by the tail patch below
        if ((uint64_t)blocks * a.samples_per_block < a.samples) j->needs_clear = true;
    }
    lists.finish(j);
}

namespace cri {
uint64_t pcm16_offset(cri_job* j, uint32_t i, const WavInfo& w) {
    const uint64_t data = j->in_off[i] + w.data_offset;
    if (w.format == WAV_S16 && (data & 1) == 0) return data;
    // other encodings, and PCM16 that lands on an odd byte of the blob (an odd-sized stream in front of it): a converted /
    // aligned copy behind the blob, so that a stream's result never depends on its neighbours in the batch
    if (!j->conv_base) j->conv_base = (j->in_bytes + 128 + 255) & ~(uint64_t)255;      // behind the blob and its read slack
    PcmConv c{};
    c.src_off = data;
    c.dst_off = j->conv_base + j->conv_bytes;
    c.count = w.total_samples;
    c.format = w.format;
    c.shift = w.shift;
    j->conv.push_back(c);
    j->conv_bytes += ((uint64_t)w.total_samples * 2 + 255) & ~(uint64_t)255;
    j->conv_max_count = std::max(j->conv_max_count, w.total_samples);
    return c.dst_off;
}
uint64_t hca_loop_input_offset(cri_job* j, uint32_t i, const WavInfo& w, const HcaEncPlan& p) {
    const uint64_t data = j->in_off[i] + w.data_offset;
    if (!j->conv_base) j->conv_base = (j->in_bytes + 128 + 255) & ~(uint64_t)255;
    const uint64_t dst0 = j->conv_base + j->conv_bytes;
    const uint32_t ch = (uint32_t)w.channels;
    uint64_t frames_done = 0;
    auto piece = [&](uint8_t kind, uint64_t first_frame, uint32_t frames) {
        if (!frames) return;
        PcmConv c{};
        c.src_off = data + first_frame * ch * w.sample_bytes;
        c.dst_off = dst0 + frames_done * ch * 2;
        c.count = frames * ch;
        c.format = w.format;
        c.shift = w.shift;
        c.kind = kind;
        c.channels = (uint8_t)ch;
        j->conv.push_back(c);
        j->conv_max_count = std::max(j->conv_max_count, c.count);
        frames_done += frames;
    };
    const uint32_t total = p.frame_count * 1024;
    piece(2, 0, p.pre_zero);
    piece(1, 0, p.pre_first);
    piece(0, 0, p.main_samples);
    piece(0, p.loop_start, p.post_samples);
    if (frames_done < total) piece(2, 0, (uint32_t)(total - frames_done));
    j->conv_bytes += ((uint64_t)total * ch * 2 + 255) & ~(uint64_t)255;
    return dst0;
}
}  // namespace cri

static void plan_adx_encode(cri_job* j) {
    std::vector<uint64_t> sizes(j->n, 0);
    std::vector<WavInfo> wavs(j->n);
    std::vector<AdxEncPlan> plans(j->n);
    const cri_adx_params& q = j->adx;
    parallel_for(j->n, [&](uint32_t i) {
        const uint8_t* d = j->blob + j->in_off[i];
        const size_t len = j->in_off[i + 1] - j->in_off[i];
        int r = parse_wav(d, len, &wavs[i]);
        if (r < 0) { j->status[i] = ERR_WAV_BASE + r; return; }
        // A looping WAV (smpl chunk) encodes a loop table unless version 5 + force flag (adx.cpp:421). One loop is what every
        // tool writes; the reference's multi-loop / zero-loop paths read past its own arrays and stay unsupported.
        const bool looping = wavs[i].looping && !(q.force_not_looping && q.version == 5);
        if (looping && wavs[i].loop_count != 1) { j->status[i] = ERR_UNSUPPORTED; return; }
        r = plan_adx_encode(wavs[i], q.bit_depth, q.block_size, q.encoding, q.highpass, q.filter, q.version, &plans[i], looping);
        if (r < 0) { j->status[i] = r; return; }
        sizes[i] = plans[i].out_size;
    });
    finish_layout(j, sizes);
    AdxLists lists;
    std::vector<AdxChain> one;
    std::vector<uint8_t> tmp;
    for (uint32_t i = 0; i < j->n; i++) {
        if (j->status[i] != OK) continue;
        const AdxEncPlan& p = plans[i];
        one.clear();
        const uint8_t* d = j->blob + j->in_off[i];
        int16_t firsts[256];
        for (int c = 0; c < p.channels; c++) firsts[c] = wav_sample_s16(wavs[i], d + wavs[i].data_offset, (size_t)c);
        // header + EOF block are host-built patches; block payload comes from the kernel
        // (written into a header + one block image: zero-filling a whole output per stream cost 16 ms per 8192 streams)
        AdxEncPlan ends = p;
        ends.out_size = (uint32_t)p.header_size + (uint32_t)p.block_size;
        tmp.assign(ends.out_size, 0);
        write_adx_frame(tmp.data(), ends, firsts);
        add_patch(j, j->out_off[i], tmp.data(), (uint32_t)p.header_size);
        add_patch(j, j->out_off[i] + p.out_size - p.block_size, tmp.data() + p.header_size, (uint32_t)p.block_size);
        const uint64_t pcm0 = pcm16_offset(j, i, wavs[i]);
        const bool is_fast = (pcm0 & 1) == 0 && p.bit_depth == 4 && p.block_size == 18;
        for (int c = 0; c < p.channels; c++) {
            AdxChain ch{};
            ch.in_off = pcm0 + (uint64_t)c * 2;
            ch.out_off = j->out_off[i] + (uint64_t)p.header_size + (uint64_t)c * p.block_size;
            ch.blocks = p.frames;
            ch.samples = p.samples;
            ch.in_stride = (uint32_t)p.channels;
            ch.out_stride = (uint32_t)p.channels * (uint32_t)p.block_size;
            ch.coef0 = p.coef[0];
            ch.coef1 = p.coef[1];
            ch.hist1 = ch.hist2 = p.version == 3 ? (int16_t)0 : firsts[c];
            ch.mode = (uint8_t)p.mode;
            ch.bit_depth = (uint8_t)p.bit_depth;
            ch.block_size = (uint8_t)p.block_size;
            ch.filter = (uint8_t)p.filter;
            ch.channels = (uint8_t)p.channels;
            ch.channel = (uint8_t)c;
            ch.stream = i;
            one.push_back(ch);
        }
        lists.add(one, is_fast);
        j->units += (uint64_t)p.frames * p.channels;
    }
    lists.finish(j);
}

// ------------------------------------------------------------------- jobs
template <class T>
static int upload_vec(cri_ctx* c, cudaStream_t s, const std::vector<T>& v, T** d) {
    *d = nullptr;
    if (v.empty()) return OK;
    const int r = pool_alloc(c, (void**)d, v.size() * sizeof(T));
    if (r != OK) return r;
    CU_TRY(c, cudaMemcpyAsync(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return OK;
}

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// CRI_TRACE=1: one line per device-pointer batch call on stderr with the host-side phases (ms)
static bool trace_on() {
    static const bool on = [] { const char* e = getenv("CRI_TRACE"); return e && *e && *e != '0'; }();
    return on;
}
struct Trace {
    double t0 = 0, last = 0;
    char line[512];
    int at = 0;
    Trace() { if (trace_on()) { t0 = last = now_ms(); line[0] = 0; } }
    void mark(const char* what) {
        if (!trace_on()) return;
        const double t = now_ms();
        at += snprintf(line + at, sizeof(line) - (size_t)at, " %s %.3f", what, t - last);
        if (at > (int)sizeof(line) - 40) at = (int)sizeof(line) - 40;
        last = t;
    }
    void done(const char* call, uint32_t n) { if (trace_on()) fprintf(stderr, "[cri trace] %s n=%u:%s | total %.3f ms\n", call, n, line, now_ms() - t0); }
};
static thread_local Trace* g_trace = nullptr;
static void trace_mark(const char* what) { if (g_trace) g_trace->mark(what); }

// ---------------------------------------------------------- device-resident input
// A device-pointer job plans from headers like every other job; the headers are fetched from the caller's HBM into a
// sparse host shadow of the blob (anonymous mapping: untouched pages cost nothing). Round 0 takes the first kHeadBytes
// of every stream; later rounds take what a stream's own header says is still missing (long HCA / ADX headers, WAV
// chunks in front of or behind the sample data).
namespace {
constexpr uint32_t kHeadBytes = 1024;

struct Valid {            // fetched byte ranges of one stream: a prefix and (WAVs) chunk headers / bodies further in
    uint64_t prefix = 0;
    std::vector<std::pair<uint64_t, uint64_t>> extra;     // [first, last)
    bool has(uint64_t off, uint64_t n) const {
        if (off + n <= prefix) return true;
        for (const auto& e : extra)
            if (off >= e.first && off + n <= e.second) return true;
        return false;
    }
};

constexpr uint32_t fourcc_le(char a, char b, char c, char d) {
    return (uint32_t)(uint8_t)a | ((uint32_t)(uint8_t)b << 8) | ((uint32_t)(uint8_t)c << 16) | ((uint32_t)(uint8_t)d << 24);
}

// What does stream (d, len) still need, given `v`? Appends one [off, off + n) range (stream-relative); nothing = complete.
void missing_ranges(int kind, const uint8_t* d, uint64_t len, const Valid& v, std::vector<std::pair<uint64_t, uint64_t>>* need) {
    auto want = [&](uint64_t from, uint64_t upto) {              // grows the prefix when contiguous with it
        upto = std::min(upto, len);
        if (from <= v.prefix) from = v.prefix;
        if (upto > from) need->emplace_back(from, upto - from);
    };
    if (kind == CRI_JOB_HCA_DECODE || kind == CRI_JOB_HCA_CRYPT) {
        if (v.prefix >= 8) want(0, be16(d + 6));                           // header size field
        return;
    }
    if (kind == CRI_JOB_ADX_DECODE) {
        if (v.prefix >= 4) want(0, (uint64_t)be16(d + 2) + 4 + 4);         // data offset: header, copyright text, first bytes
        return;
    }
    // WAV: the chunk walk of parse_wav (pcm.cpp:291-342): every chunk header, the fmt / smpl bodies, the first sample frame
    if (v.prefix < 12 || le32(d) != fourcc_le('R', 'I', 'F', 'F')) return;
    const uint64_t riff_size = le32(d + 4);
    uint64_t at = 12, walked = 4;
    uint64_t first_frame = 2048;                                           // one sample frame: <= 255 channels x 8 bytes until fmt says
    while (walked < riff_size && at + 8 <= len) {
        if (!v.has(at, 8)) { want(at, at + 512); return; }               // next chunk header (and, likely, its small body)
        const uint32_t tag = le32(d + at), body = le32(d + at + 4);
        uint64_t span = (uint64_t)body + 8;
        if ((span & 1) && span + walked + 1 <= riff_size) span += 1;
        uint64_t n = 8;
        if (tag == fourcc_le('f', 'm', 't', ' ') || tag == fourcc_le('s', 'm', 'p', 'l')) n = std::min<uint64_t>(span, 128);
        else if (tag == fourcc_le('d', 'a', 't', 'a')) n = 8 + first_frame;
        n = std::min(n, len - at);
        if (!v.has(at, n)) { want(at, at + n); return; }
        if (tag == fourcc_le('f', 'm', 't', ' ') && body >= 16) {          // channels x bytes per sample (the planners read the first frame)
            const uint64_t channels = (uint64_t)d[at + 10] | ((uint64_t)d[at + 11] << 8), bits = (uint64_t)d[at + 22] | ((uint64_t)d[at + 23] << 8);
            first_frame = std::min<uint64_t>(std::max<uint64_t>(channels, 1) * std::max<uint64_t>((bits + 7) / 8, 1), 2048);
        }
        at += span;
        walked += span;
    }
}
}  // namespace

static int fetch_segments(cri_ctx* c, cudaStream_t stream, const uint8_t* d_blob, std::vector<Segment>& segs, uint64_t total,
                          uint8_t* shadow, const std::vector<uint64_t>& shadow_off) {
    if (segs.empty()) return OK;
    if (c->pin_stage_cap < total) {
        if (c->pin_stage) cudaFreeHost(c->pin_stage);
        c->pin_stage = nullptr;
        c->pin_stage_cap = 0;
        const size_t cap = std::max<size_t>(total + total / 2, size_t(1) << 20);
        CU_TRY(c, cudaHostAlloc((void**)&c->pin_stage, cap, cudaHostAllocDefault));
        c->pin_stage_cap = cap;
    }
    uint8_t* d_stage = nullptr;
    Segment* d_segs = nullptr;
    int r = pool_alloc(c, (void**)&d_stage, total);
    if (r == OK) r = pool_alloc(c, (void**)&d_segs, segs.size() * sizeof(Segment));
    if (r == OK) r = [&]() -> int {
        CU_TRY(c, cudaMemcpyAsync(d_segs, segs.data(), segs.size() * sizeof(Segment), cudaMemcpyHostToDevice, stream));
        launch_gather_segments(d_stage, d_blob, d_segs, (uint32_t)segs.size(), stream, &c->launches);
        CU_TRY(c, cudaMemcpyAsync(c->pin_stage, d_stage, total, cudaMemcpyDeviceToHost, stream));
        CU_TRY(c, cudaStreamSynchronize(stream));
        return OK;
    }();
    pool_free(c, d_stage);
    pool_free(c, d_segs);
    if (r != OK) return r;
    for (size_t k = 0; k < segs.size(); k++) memcpy(shadow + shadow_off[k], c->pin_stage + segs[k].dst_off, segs[k].bytes);
    return OK;
}

// Builds the host shadow of a device blob: *shadow (munmap it with *bytes) holds every header byte planning reads.
// The shadow is an all-zero anonymous mapping of the blob's size; only fetched ranges are ever written, `dirty` lists them
// and release_shadow zeroes them again, so a context can hand the same mapping (its pages already faulted in) to the
// next job: a fresh mapping costs one page fault per stream.
static void release_shadow(cri_ctx* c, uint8_t* shadow, size_t bytes, const std::vector<std::pair<uint64_t, uint32_t>>& dirty) {
    if (!shadow) return;
    if (!c || c->idle_shadows.size() >= 2) { munmap(shadow, bytes); return; }
    for (const auto& d : dirty) memset(shadow + d.first, 0, d.second);
    c->idle_shadows.emplace_back(shadow, bytes);
}

static int build_shadow(cri_ctx* c, int kind, const uint8_t* d_blob, const uint64_t* off, uint32_t n, cudaStream_t stream,
                        uint8_t** shadow, size_t* bytes, std::vector<std::pair<uint64_t, uint32_t>>* dirty) {
    const uint64_t total_in = n ? off[n] : 0;
    const size_t want = (size_t)std::max<uint64_t>(total_in, 1) + 64;
    *shadow = nullptr;
    for (size_t k = 0; k < c->idle_shadows.size(); k++)
        if (c->idle_shadows[k].second >= want) {
            *shadow = c->idle_shadows[k].first;
            *bytes = c->idle_shadows[k].second;
            c->idle_shadows.erase(c->idle_shadows.begin() + k);
            break;
        }
    if (!*shadow) {
        *bytes = want;
        void* m = mmap(nullptr, *bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (m == MAP_FAILED) return ERR_BUFFER;
        *shadow = (uint8_t*)m;
    }
    std::vector<Valid> valid(n);
    std::vector<Segment> segs;
    std::vector<uint64_t> shadow_off;
    std::vector<std::pair<uint64_t, uint64_t>> need;
    for (int round = 0; round < 12; round++) {
        segs.clear();
        shadow_off.clear();
        uint64_t total = 0;
        std::vector<std::pair<uint32_t, std::pair<uint64_t, uint64_t>>> got;
        for (uint32_t i = 0; i < n; i++) {
            const uint64_t len = off[i + 1] - off[i];
            need.clear();
            if (round == 0) { if (len) need.emplace_back(0, std::min<uint64_t>(len, kHeadBytes)); }
            else missing_ranges(kind, *shadow + off[i], len, valid[i], &need);
            for (auto& rg : need) {
                if (!rg.second) continue;
                Segment sg{off[i] + rg.first, total, (uint32_t)rg.second, 0};
                segs.push_back(sg);
                shadow_off.push_back(off[i] + rg.first);
                total += (rg.second + 15) & ~(uint64_t)15;
                got.push_back({i, rg});
            }
        }
        if (segs.empty()) break;
        const int r = fetch_segments(c, stream, d_blob, segs, total, *shadow, shadow_off);
        for (size_t k = 0; k < segs.size(); k++) dirty->emplace_back(shadow_off[k], segs[k].bytes);
        if (r != OK) return r;
        for (auto& g : got) {
            Valid& v = valid[g.first];
            if (g.second.first == v.prefix) v.prefix += g.second.second;
            else v.extra.emplace_back(g.second.first, g.second.first + g.second.second);
        }
    }
    return OK;
}

static int take_events(cri_ctx* c, cri_job* j) {
    for (auto& e : j->ev) {
        if (!c->idle_events.empty()) { e = c->idle_events.back(); c->idle_events.pop_back(); }
        else CU_TRY(c, cudaEventCreate(&e));
    }
    return OK;
}

// Plan on the host, take HBM from the context's cache, and enqueue every upload on `stream`. Nothing here waits for the GPU
// (a device-pointer job waits for its header fetches).
// `aux` (device-pointer jobs of a pipelined batch call): header fetches, table uploads and the payload copy go there, so
// that they neither wait for nor hold up the kernels of the previous piece on `stream`; `stream` waits for them (`staged`).
static int job_create_on(cri_ctx* c, const cri_job_desc* d, cudaStream_t stream, cri_job** out, const uint64_t* expect_out = nullptr,
                         cudaStream_t aux = nullptr, cudaEvent_t staged = nullptr) {
    *out = nullptr;
    if (!c || !d || (!d->blob && !d->d_blob && d->n) || !d->offsets) return ERR_BUFFER;
    CU_TRY(c, cudaSetDevice(c->device));
    cri_job* j = new (std::nothrow) cri_job();
    if (!j) return ERR_BUFFER;
    j->stream = stream;
    j->kind = d->kind;
    j->n = d->n;
    j->blob = d->blob;
    j->in_off.assign(d->offsets, d->offsets + d->n + 1);
    j->in_bytes = j->in_off[d->n];
    j->status.assign(d->n, OK);
    j->adx = d->adx;
    j->quality = d->quality;
    j->encrypt = d->encrypt;
    j->ciph_type = d->ciph_type;
    if (d->keys) j->keys.assign(d->keys, d->keys + d->n);
    if (d->subkeys) j->subkeys.assign(d->subkeys, d->subkeys + d->n);
    int rc = OK;
    if (d->d_blob) {
        j->d_src = d->d_blob;
        rc = build_shadow(c, d->kind, d->d_blob, d->offsets, d->n, aux ? aux : stream, &j->shadow, &j->shadow_bytes, &j->shadow_dirty);
        j->blob = j->shadow;
        trace_mark("headers");
    }
    if (d->d_out) {      // kernels take a 256-byte aligned blob start: align the pointer down and shift every offset instead
        const uintptr_t addr = reinterpret_cast<uintptr_t>(d->d_out);
        j->out_delta = addr & 255;
        j->d_out = reinterpret_cast<uint8_t*>(addr - j->out_delta);
        j->own_out = false;
    }
    if (rc == OK) switch (d->kind) {
        case CRI_JOB_ADX_DECODE: plan_adx_decode(j); break;
        case CRI_JOB_ADX_ENCODE: plan_adx_encode(j); break;
        case CRI_JOB_HCA_DECODE: rc = plan_hca_decode(c, j); break;
        case CRI_JOB_HCA_CRYPT: rc = plan_hca_crypt(c, j); break;
        case CRI_JOB_HCA_ENCODE: rc = plan_hca_encode(c, j); break;
        default: rc = ERR_UNSUPPORTED;
    }
    trace_mark("plan");
    if (rc == OK && expect_out)      // the caller's buffer is laid out by *_sizes(): it must be the packed layout planned here
        for (uint32_t i = 0; i <= j->n && rc == OK; i++)
            if (expect_out[i] - expect_out[0] != j->out_off_pub[i]) rc = ERR_BUFFER;
    if (rc == OK) rc = [&]() -> int {
        int r = take_events(c, j);
        if (r != OK) return r;
        cudaStream_t run_stream = stream;
        if (aux) { stream = aux; j->stream = aux; }             // uploads below (upload_hca_tables reads j->stream)
        struct Back { cri_job* j; cudaStream_t s; ~Back() { j->stream = s; } } back{j, run_stream};
        // slack: kernels read whole 16-byte rows, up to four rows ahead; then the WAV ingest's conversion region
        const uint64_t in_alloc = j->conv_bytes ? j->conv_base + j->conv_bytes + 128 : std::max<uint64_t>(j->in_bytes, 16) + 128;
        r = pool_alloc(c, (void**)&j->d_in, in_alloc);
        if (r == OK && j->own_out) r = pool_alloc(c, (void**)&j->d_out, std::max<uint64_t>(j->out_bytes, 16) + 16);
        if (r == OK) r = pool_alloc(c, (void**)&j->d_status, sizeof(int32_t) * std::max<uint32_t>(j->n, 1));
        if (r != OK) return r;
        // Every byte of the output blob is written by a kernel or a patch on every run; a job that leaves gaps (streams
        // cut short, unsupported layouts) sets needs_clear and is zero-filled at the start of each run instead.
        if (c->poison && j->out_bytes) CU_TRY(c, cudaMemsetAsync(j->d_out + j->out_delta, 0xA5, j->out_bytes, stream));
        r = upload_vec(c, stream, j->adx_chains, &j->d_adx_chains);
        if (r == OK) r = upload_vec(c, stream, j->conv, &j->d_conv);
        if (r == OK) r = upload_vec(c, stream, j->patches, &j->d_patches);
        if (r == OK) r = upload_vec(c, stream, j->patch_bytes, &j->d_patch_bytes);
        if (r == OK) r = upload_hca_tables(c, j);
        if (r != OK) return r;
        if (j->in_bytes) {
            if (j->d_src) CU_TRY(c, cudaMemcpyAsync(j->d_in, j->d_src, j->in_bytes, cudaMemcpyDeviceToDevice, stream));
            else CU_TRY(c, cudaMemcpyAsync(j->d_in, j->blob, j->in_bytes, cudaMemcpyHostToDevice, stream));
        }
        if (aux && staged) {
            CU_TRY(c, cudaEventRecord(staged, aux));
            CU_TRY(c, cudaStreamWaitEvent(run_stream, staged, 0));
        }
        return OK;
    }();
    trace_mark("alloc+uploads");
    if (rc != OK) {
        cudaStreamSynchronize(stream);
        cri_job_destroy(c, j);
        return rc;
    }
    *out = j;
    return OK;
}

extern "C" int cri_job_create(cri_ctx* c, const cri_job_desc* d, cri_job** out) {
    if (!c) return ERR_CUDA;
    const int rc = job_create_on(c, d, c->stream, out);
    if (rc != OK) return rc;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return OK;
}

extern "C" uint64_t cri_job_out_bytes(const cri_job* j) { return j->out_bytes; }
extern "C" const uint64_t* cri_job_out_offsets(const cri_job* j) { return j->out_off_pub.data(); }
extern "C" uint64_t cri_job_units(const cri_job* j) { return j->units; }

extern "C" int cri_job_upload(cri_ctx* c, cri_job* j) {
    CU_TRY(c, cudaSetDevice(c->device));
    if (j->in_bytes) {
        if (j->d_src) CU_TRY(c, cudaMemcpyAsync(j->d_in, j->d_src, j->in_bytes, cudaMemcpyDeviceToDevice, j->stream));
        else CU_TRY(c, cudaMemcpyAsync(j->d_in, j->blob, j->in_bytes, cudaMemcpyHostToDevice, j->stream));
    }
    CU_TRY(c, cudaStreamSynchronize(j->stream));
    return OK;
}

// Enqueue the kernels of one pass over the job's resident input. No host wait.
static int job_enqueue_run(cri_ctx* c, cri_job* j) {
    cudaStream_t s = j->stream;
    CU_TRY(c, cudaMemsetAsync(j->d_status, 0, sizeof(int32_t) * std::max<uint32_t>(j->n, 1), s));
    if (j->needs_clear && j->out_bytes) CU_TRY(c, cudaMemsetAsync(j->d_out + j->out_delta, 0, j->out_bytes, s));
    CU_TRY(c, cudaEventRecord(j->ev[0], s));
    launch_pcm_convert(j->d_in, j->d_conv, (uint32_t)j->conv.size(), j->conv_max_count, s, &c->launches);
    launch_scatter_patches(j->d_out, j->d_patch_bytes, j->d_patches, (uint32_t)j->patches.size(), s, &c->launches);
    j->have_dominant = false;
    switch (j->kind) {
        case CRI_JOB_ADX_DECODE:
            CU_TRY(c, cudaEventRecord(j->ev[2], s));
            launch_adx_decode(j->d_in, j->d_out, j->d_adx_chains, j->n_fast, j->n_generic, s, &c->launches);
            CU_TRY(c, cudaEventRecord(j->ev[3], s));
            j->have_dominant = true;
            break;
        case CRI_JOB_ADX_ENCODE:
            CU_TRY(c, cudaEventRecord(j->ev[2], s));
            launch_adx_encode(j->d_in, j->d_out, j->d_adx_chains, j->n_fast, j->n_generic, s, &c->launches);
            CU_TRY(c, cudaEventRecord(j->ev[3], s));
            j->have_dominant = true;
            break;
        case CRI_JOB_HCA_DECODE:
        case CRI_JOB_HCA_CRYPT:
        case CRI_JOB_HCA_ENCODE: {
            const int r = run_hca(c, j, &j->have_dominant);
            if (r != OK) return r;
            break;
        }
    }
    for (const auto& dc : j->dev_copies)
        CU_TRY(c, cudaMemcpyAsync(j->d_out + dc.dst_off, j->d_in + dc.src_off, dc.bytes, cudaMemcpyDeviceToDevice, s));
    CU_TRY(c, cudaEventRecord(j->ev[1], s));
    CU_TRY(c, cudaGetLastError());
    return OK;
}

// Wait for the job's stream and fold its device time into the context's counters.
static int job_wait_run(cri_ctx* c, cri_job* j, bool accumulate) {
    CU_TRY(c, cudaStreamSynchronize(j->stream));
    float ms = 0.f, dom = 0.f;
    cudaEventElapsedTime(&ms, j->ev[0], j->ev[1]);
    if (j->have_dominant) cudaEventElapsedTime(&dom, j->ev[2], j->ev[3]);
    c->last_ms = accumulate ? c->last_ms + ms : ms;
    c->last_dominant_ms = accumulate ? c->last_dominant_ms + dom : dom;
    return OK;
}

extern "C" int cri_job_run(cri_ctx* c, cri_job* j) {
    CU_TRY(c, cudaSetDevice(c->device));
    const int r = job_enqueue_run(c, j);
    if (r != OK) return r;
    return job_wait_run(c, j, false);
}

static int job_enqueue_download(cri_ctx* c, cri_job* j, uint8_t* out_blob, int32_t* status, int32_t* landing) {
    if (!landing) {
        j->dev_status.assign(j->n, 0);
        landing = j->dev_status.data();
    }
    j->h_status = landing;
    j->dl_out = out_blob;
    j->dl_status = status;
    if (j->n) CU_TRY(c, cudaMemcpyAsync(landing, j->d_status, sizeof(int32_t) * j->n, cudaMemcpyDeviceToHost, j->stream));
    if (out_blob && j->out_bytes) CU_TRY(c, cudaMemcpyAsync(out_blob, j->d_out + j->out_delta, j->out_bytes, cudaMemcpyDeviceToHost, j->stream));
    return OK;
}

static int job_wait_download(cri_ctx* c, cri_job* j) {
    CU_TRY(c, cudaStreamSynchronize(j->stream));
    for (uint32_t i = 0; i < j->n; i++) {
        const int32_t st = j->status[i] != OK ? j->status[i] : j->h_status[i];
        if (j->dl_status) j->dl_status[i] = st;
        if (st != OK && j->dl_out && j->status[i] == OK)  // a stream that failed on the device leaves silence, not garbage
            memset(j->dl_out + j->out_off_pub[i], 0, j->out_off_pub[i + 1] - j->out_off_pub[i]);
    }
    return OK;
}

extern "C" int cri_job_download(cri_ctx* c, cri_job* j, uint8_t* out_blob, int32_t* status) {
    CU_TRY(c, cudaSetDevice(c->device));
    const int r = job_enqueue_download(c, j, out_blob, status, nullptr);
    if (r != OK) return r;
    return job_wait_download(c, j);
}

extern "C" void cri_job_destroy(cri_ctx* c, cri_job* j) {
    if (!j) return;
    if (c) cudaSetDevice(c->device);
    for (auto& e : j->ev)
        if (e) { if (c) c->idle_events.push_back(e); else cudaEventDestroy(e); }
    if (!j->own_out) j->d_out = nullptr;
    release_shadow(c, j->shadow, j->shadow_bytes, j->shadow_dirty);
    for (void* p : {(void*)j->d_in, (void*)j->d_out, (void*)j->d_status, (void*)j->d_adx_chains, (void*)j->d_conv, (void*)j->d_patches,
                    (void*)j->d_patch_bytes})
        pool_free(c, p);
    free_hca_tables(c, j);
    delete j;
}

// ------------------------------------------------- batch = job in one call
// The batch is cut into chunks of whole streams; each chunk is a job of its own
// on one of kPipeDepth streams, so the upload of chunk k+1, the kernels of
// chunk k and the download of chunk k-1 overlap (two copy engines + SMs), and
// the host plans the next chunk while the GPU works. Chunks are independent:
// no codec on this path carries state from one stream to another.
struct ChunkRun {
    cri_job* job = nullptr;
    uint32_t s0 = 0, s1 = 0;
    bool packed = true;
};

static int chunk_finish(cri_ctx* c, ChunkRun& r, uint8_t* out_blob, const uint64_t* out_offsets, int32_t* status) {
    if (!r.job) return OK;
    cri_job* j = r.job;
    int rc = job_wait_run(c, j, true);
    if (rc == OK) rc = job_wait_download(c, j);
    if (rc == OK && !r.packed)       // caller chose a different layout: place each stream from the staging blob
        for (uint32_t i = 0; i < j->n; i++) {
            const uint64_t sz = j->out_off_pub[i + 1] - j->out_off_pub[i];
            const uint64_t* oo = out_offsets + r.s0;
            if (oo[i + 1] - oo[i] < sz) { if (status) status[r.s0 + i] = ERR_BUFFER; continue; }
            memcpy(out_blob + oo[i], j->staging.data() + j->out_off_pub[i], sz);
        }
    cri_job_destroy(c, j);
    r.job = nullptr;
    return rc;
}

static int run_batch(cri_ctx* c, cri_job_desc d, uint8_t* out_blob, const uint64_t* out_offsets, int32_t* status) {
    if (!c) return ERR_CUDA;
    if (!d.offsets || (!d.blob && d.n)) return ERR_BUFFER;
    const double t_begin = now_ms();
    // chunk count from the larger of the two directions (the copies are what the pipeline hides)
    const uint64_t in_total = d.n ? d.offsets[d.n] - d.offsets[0] : 0;
    const uint64_t out_total = (out_offsets && d.n) ? out_offsets[d.n] - out_offsets[0] : 0;
    // ~16 chunks keeps pipeline fill + drain near 1/8 of the copy time; 8 MiB floor so launches stay amortised
    const uint64_t larger = std::max(in_total, out_total);
    uint64_t target = std::min<uint64_t>(std::max<uint64_t>(larger / 16, uint64_t(8) << 20), uint64_t(256) << 20);
    if (const char* e = getenv("CRI_CHUNK_MB")) target = std::max<uint64_t>(1, strtoull(e, nullptr, 10)) << 20;
    uint32_t n_chunks = (uint32_t)std::min<uint64_t>((larger + target - 1) / target, 64);
    n_chunks = std::max<uint32_t>(1, std::min<uint32_t>(n_chunks, d.n));
    if (!out_offsets) n_chunks = 1;                                  // no caller layout to place later chunks by
    c->last_ms = c->last_dominant_ms = 0.f;

    ChunkRun slots[kPipeDepth];
    std::vector<uint64_t> local;
    int rc = OK;
    double plan_ms = 0;
    for (uint32_t k = 0; k < n_chunks && rc == OK; k++) {
        ChunkRun& r = slots[k % kPipeDepth];
        rc = chunk_finish(c, r, out_blob, out_offsets, status);
        if (rc != OK) break;
        r.s0 = (uint32_t)((uint64_t)d.n * k / n_chunks);
        r.s1 = (uint32_t)((uint64_t)d.n * (k + 1) / n_chunks);
        const uint32_t m = r.s1 - r.s0;
        local.resize(m + 1);
        for (uint32_t i = 0; i <= m; i++) local[i] = d.offsets[r.s0 + i] - d.offsets[r.s0];
        cri_job_desc sub = d;
        sub.blob = d.blob ? d.blob + d.offsets[r.s0] : nullptr;
        sub.offsets = local.data();
        sub.n = m;
        if (d.keys) sub.keys = d.keys + r.s0;
        if (d.subkeys) sub.subkeys = d.subkeys + r.s0;
        const double t0 = now_ms();
        rc = job_create_on(c, &sub, c->pipe[k % kPipeDepth], &r.job);
        plan_ms += now_ms() - t0;
        if (rc != OK) break;
        cri_job* j = r.job;
        rc = job_enqueue_run(c, j);
        if (rc != OK) break;
        r.packed = true;
        if (out_offsets)
            for (uint32_t i = 0; i <= m && r.packed; i++) r.packed = out_offsets[r.s0 + i] - out_offsets[r.s0] == j->out_off_pub[i];
        uint8_t* dst;
        if (r.packed) {
            dst = out_blob ? out_blob + (out_offsets ? out_offsets[r.s0] : 0) : nullptr;
        } else {
            j->staging.resize(j->out_bytes);
            dst = j->staging.data();
        }
        const int slot = k % kPipeDepth;
        if (c->pin_status_cap[slot] < m) {
            cudaFreeHost(c->pin_status[slot]);
            c->pin_status[slot] = nullptr;
            c->pin_status_cap[slot] = 0;
            if (cudaHostAlloc((void**)&c->pin_status[slot], sizeof(int32_t) * m * 2, cudaHostAllocDefault) == cudaSuccess)
                c->pin_status_cap[slot] = (size_t)m * 2;
            else
                cudaGetLastError();      // fall back to the job's pageable landing buffer (correct, just not overlapped)
        }
        rc = job_enqueue_download(c, j, dst, status ? status + r.s0 : nullptr, c->pin_status_cap[slot] >= m ? c->pin_status[slot] : nullptr);
    }
    for (auto& r : slots) {
        if (rc == OK) rc = chunk_finish(c, r, out_blob, out_offsets, status);
        else if (r.job) {
            cudaStreamSynchronize(r.job->stream);
            cri_job_destroy(c, r.job);
            r.job = nullptr;
        }
    }
    if (c->trace)
        fprintf(stderr, "[cri] batch kind=%d n=%u chunks=%u in=%.1f MB out=%.1f MB host-plan=%.2f ms kernels=%.2f ms total=%.2f ms\n",
                d.kind, d.n, n_chunks, in_total / 1e6, out_total / 1e6, plan_ms, c->last_ms, now_ms() - t_begin);
    return rc;
}

static cri_job_desc make_desc(int kind, const uint8_t* blob, const uint64_t* offsets, uint32_t n) {
    cri_job_desc d;
    memset(&d, 0, sizeof d);
    d.kind = kind;
    d.blob = blob;
    d.offsets = offsets;
    d.n = n;
    return d;
}

extern "C" int cri_adx_decode_sizes(const uint8_t* blob, const uint64_t* off, uint32_t n, uint64_t* sizes, int32_t* status) {
    for (uint32_t i = 0; i < n; i++) {
        AdxInfo a;
        sizes[i] = 0;
        const int r = adx_decode_size_one(blob + off[i], off[i + 1] - off[i], &a, &sizes[i]);
        if (status) status[i] = r;
    }
    return OK;
}

extern "C" int cri_adx_decode_batch(cri_ctx* c, const uint8_t* blob, const uint64_t* off, uint32_t n, uint8_t* out,
                                    const uint64_t* out_off, int32_t* status) {
    return run_batch(c, make_desc(CRI_JOB_ADX_DECODE, blob, off, n), out, out_off, status);
}

extern "C" int cri_adx_encode_sizes(const uint8_t* blob, const uint64_t* off, uint32_t n, const cri_adx_params* p,
                                    uint64_t* sizes, int32_t* status) {
    for (uint32_t i = 0; i < n; i++) {
        WavInfo w;
        AdxEncPlan pl;
        sizes[i] = 0;
        int r = parse_wav(blob + off[i], off[i + 1] - off[i], &w);
        if (r < 0) r += ERR_WAV_BASE;
        else {
            const bool looping = w.looping && !(p->force_not_looping && p->version == 5);
            if (looping && w.loop_count != 1) r = ERR_UNSUPPORTED;
            else r = plan_adx_encode(w, p->bit_depth, p->block_size, p->encoding, p->highpass, p->filter, p->version, &pl, looping);
        }
        if (r == OK) sizes[i] = pl.out_size;
        if (status) status[i] = r;
    }
    return OK;
}

extern "C" int cri_adx_encode_batch(cri_ctx* c, const uint8_t* blob, const uint64_t* off, uint32_t n, const cri_adx_params* p,
                                    uint8_t* out, const uint64_t* out_off, int32_t* status) {
    cri_job_desc d = make_desc(CRI_JOB_ADX_ENCODE, blob, off, n);
    d.adx = *p;
    return run_batch(c, d, out, out_off, status);
}

extern "C" int cri_hca_decode_sizes(const uint8_t* blob, const uint64_t* off, uint32_t n, uint64_t* sizes, int32_t* status) {
    parallel_for(n, [&](uint32_t i) {
        HcaInfo h;
        sizes[i] = 0;
        const int r = hca_decode_size_one(blob + off[i], off[i + 1] - off[i], &h, &sizes[i]);
        if (status) status[i] = r;
    });
    return OK;
}

extern "C" int cri_hca_decode_batch(cri_ctx* c, const uint8_t* blob, const uint64_t* off, uint32_t n, const uint64_t* keys,
                                    const uint16_t* subkeys, uint8_t* out, const uint64_t* out_off, int32_t* status) {
    cri_job_desc d = make_desc(CRI_JOB_HCA_DECODE, blob, off, n);
    d.keys = keys;
    d.subkeys = subkeys;
    return run_batch(c, d, out, out_off, status);
}

extern "C" int cri_hca_crypt_batch(cri_ctx* c, const uint8_t* blob, const uint64_t* off, uint32_t n, int encrypt,
                                   uint32_t ciph_type, const uint64_t* keys, const uint16_t* subkeys, uint8_t* out,
                                   int32_t* status) {
    cri_job_desc d = make_desc(CRI_JOB_HCA_CRYPT, blob, off, n);
    d.keys = keys;
    d.subkeys = subkeys;
    d.encrypt = encrypt;
    d.ciph_type = ciph_type;
    return run_batch(c, d, out, off, status);
}

extern "C" int cri_hca_encode_sizes_ex(const uint8_t* blob, const uint64_t* off, uint32_t n, uint32_t quality,
                                       uint32_t force_not_looping, uint64_t* sizes, int32_t* status);
extern "C" int cri_hca_encode_sizes(const uint8_t* blob, const uint64_t* off, uint32_t n, uint32_t quality, uint64_t* sizes,
                                    int32_t* status) {
    return cri_hca_encode_sizes_ex(blob, off, n, quality, 0, sizes, status);
}

extern "C" int cri_hca_encode_sizes_ex(const uint8_t* blob, const uint64_t* off, uint32_t n, uint32_t quality,
                                       uint32_t force_not_looping, uint64_t* sizes, int32_t* status) {
    for (uint32_t i = 0; i < n; i++) {
        WavInfo w;
        HcaEncPlan pl;
        sizes[i] = 0;
        int r = parse_wav(blob + off[i], off[i + 1] - off[i], &w);
        if (r < 0) r += ERR_WAV_BASE;
        else if (w.looping && !force_not_looping) {
            r = plan_hca_encode_loop(w, quality, &pl);
            if (r < 0 && r != ERR_UNSUPPORTED) r = ERR_HCA_CHANNELS;
        } else if (plan_hca_encode((unsigned)w.channels, (unsigned)w.rate, w.total_samples / (unsigned)w.channels, quality, &pl) < 0)
            r = ERR_HCA_CHANNELS;
        if (r == OK) sizes[i] = (uint64_t)pl.header_size + (uint64_t)pl.frame_count * pl.frame_size;
        if (status) status[i] = r;
    }
    return OK;
}

extern "C" int cri_hca_encode_batch(cri_ctx* c, const uint8_t* blob, const uint64_t* off, uint32_t n, uint32_t quality,
                                    uint32_t force_not_looping, uint8_t* out, const uint64_t* out_off, int32_t* status) {
    cri_job_desc d = make_desc(CRI_JOB_HCA_ENCODE, blob, off, n);
    d.quality = quality;
    d.adx.force_not_looping = force_not_looping;
    return run_batch(c, d, out, out_off, status);
}

// ------------------------------------------------- device-pointer batch calls
// One job for the whole batch (there is no PCIe transfer to hide, so no chunking): headers are fetched for planning,
// the payload is copied once inside HBM into the engine's padded input buffer (kernels read whole 16-byte rows past
// the last stream), kernels write straight into the caller's output buffer.
// One call = up to four pieces of whole streams, each a job of its own: while piece k's kernels run on the caller's stream,
// the host fetches and parses piece k + 1's headers and plans it (stream c->pipe[0], which also carries the uploads and the
// payload copy), and piece k - 1 is torn down. Per 8192 HCA streams the host side is ~4 ms (header fetch 1.5, planning 1.5,
// teardown 1.0) against ~6 ms of kernels: unpipelined the GPU idles for 40 % of the call.
struct DevPiece {
    cri_job* job = nullptr;
    cudaEvent_t staged = nullptr, done = nullptr;
    uint32_t s0 = 0;
    int slot = 0;
};

static int dev_piece_finish(cri_ctx* c, DevPiece& p, cudaStream_t st, int32_t* status, bool* zeroed) {
    if (!p.job) return OK;
    cri_job* j = p.job;
    p.job = nullptr;
    int rc = OK;
    if (cudaEventSynchronize(p.done) != cudaSuccess) { c->error = "cudaEventSynchronize"; rc = ERR_CUDA; }
    if (rc == OK) {
        float ms = 0.f, dom = 0.f;
        cudaEventElapsedTime(&ms, j->ev[0], j->ev[1]);
        if (j->have_dominant) cudaEventElapsedTime(&dom, j->ev[2], j->ev[3]);
        c->last_ms += ms;
        c->last_dominant_ms += dom;
        for (uint32_t i = 0; i < j->n; i++) {
            const int32_t v = j->status[i] != OK ? j->status[i] : j->h_status[i];
            if (status) status[p.s0 + i] = v;
            if (j->status[i] == OK && j->h_status[i] != OK && j->out_off[i + 1] > j->out_off[i]) {   // failed on the device: silence, not garbage
                cudaMemsetAsync(j->d_out + j->out_off[i], 0, j->out_off[i + 1] - j->out_off[i], st);
                *zeroed = true;
            }
        }
    } else {
        cudaStreamSynchronize(st);
    }
    cri_job_destroy(c, j);
    return rc;
}

static int run_batch_dev(cri_ctx* c, cri_job_desc d, const uint64_t* out_offsets, int32_t* status) {
    if (!c) return ERR_CUDA;
    if (!d.offsets || (d.n && (!d.d_blob || !d.d_out))) return ERR_BUFFER;
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = d.stream ? (cudaStream_t)d.stream : c->stream;
    cudaStream_t aux = c->pipe[0];
    c->last_ms = c->last_dominant_ms = 0.f;
    Trace tr;
    g_trace = trace_on() ? &tr : nullptr;
    struct Untrace { ~Untrace() { g_trace = nullptr; } } untrace;

    // Only where a piece's kernels shrink with the piece: the HCA frame kernels. An ADX chain is serial over its whole
    // stream, so a quarter of the chains takes as long as all of them (measured: 11.3 -> 16.7 ms per 8192 streams when cut
    // in four), and the crypt kernel is 0.25 ms against ~5 ms of host work that only grows with the number of pieces.
    uint32_t n_pieces = 1;
    if (d.kind == CRI_JOB_HCA_DECODE || d.kind == CRI_JOB_HCA_ENCODE) n_pieces = d.n >= 4096 ? 4 : d.n >= 1024 ? 2 : 1;   // 8192 HCA streams: 1 / 2 / 4 / 8 pieces -> 10.3 / 8.9 / 7.9 / 8.1 ms
    if (const char* e = getenv("CRI_DEV_PIECES")) n_pieces = (uint32_t)std::max(1, atoi(e));
    n_pieces = std::max<uint32_t>(1, std::min<uint32_t>(std::min<uint32_t>(n_pieces, 16), std::max<uint32_t>(d.n, 1)));
    if (!out_offsets) n_pieces = 1;                                  // no caller layout to place later pieces by

    std::vector<cudaEvent_t> events;
    auto event = [&]() -> cudaEvent_t {
        cudaEvent_t e = nullptr;
        if (!c->idle_events.empty()) { e = c->idle_events.back(); c->idle_events.pop_back(); }
        else if (cudaEventCreate(&e) != cudaSuccess) e = nullptr;
        if (e) events.push_back(e);
        return e;
    };
    int rc = OK;
    // the caller's input is ready when the work already on its stream is: the fetch stream starts behind that point
    if (cudaEvent_t in_ready = event()) {
        cudaEventRecord(in_ready, st);
        cudaStreamWaitEvent(aux, in_ready, 0);
    } else rc = ERR_CUDA;

    DevPiece pieces[2];
    std::vector<uint64_t> local, local_out;
    bool zeroed = false;
    for (uint32_t k = 0; k < n_pieces && rc == OK; k++) {
        DevPiece& p = pieces[k & 1];
        const uint32_t s0 = (uint32_t)((uint64_t)d.n * k / n_pieces), s1 = (uint32_t)((uint64_t)d.n * (k + 1) / n_pieces), m = s1 - s0;
        local.resize(m + 1);
        for (uint32_t i = 0; i <= m; i++) local[i] = d.offsets[s0 + i] - d.offsets[s0];
        cri_job_desc sub = d;
        sub.d_blob = d.d_blob + d.offsets[s0];
        sub.offsets = local.data();
        sub.n = m;
        if (d.keys) sub.keys = d.keys + s0;
        if (d.subkeys) sub.subkeys = d.subkeys + s0;
        const uint64_t* expect = nullptr;
        if (out_offsets) {
            sub.d_out = d.d_out + out_offsets[s0];
            expect = out_offsets + s0;
        }
        p.s0 = s0;
        p.slot = (int)(k & 1);
        p.staged = event();
        p.done = event();
        if (!p.staged || !p.done) { rc = ERR_CUDA; break; }
        rc = job_create_on(c, &sub, st, &p.job, expect, aux, p.staged);
        if (rc != OK) { p.job = nullptr; break; }
        cri_job* j = p.job;
        rc = job_enqueue_run(c, j);
        if (rc != OK) break;
        const int slot = p.slot;
        if (c->pin_status_cap[slot] < m) {
            cudaFreeHost(c->pin_status[slot]);
            c->pin_status[slot] = nullptr;
            c->pin_status_cap[slot] = 0;
            if (cudaHostAlloc((void**)&c->pin_status[slot], sizeof(int32_t) * m * 2, cudaHostAllocDefault) == cudaSuccess)
                c->pin_status_cap[slot] = (size_t)m * 2;
            else
                cudaGetLastError();
        }
        rc = job_enqueue_download(c, j, nullptr, nullptr, c->pin_status_cap[slot] >= m ? c->pin_status[slot] : nullptr);
        if (rc == OK && cudaEventRecord(p.done, st) != cudaSuccess) rc = ERR_CUDA;
        trace_mark("enqueue");
        if (rc == OK) rc = dev_piece_finish(c, pieces[(k & 1) ^ 1], st, status, &zeroed);      // the piece in front of this one
        trace_mark("finish-prev");
    }
    for (auto& p : pieces) {
        if (rc == OK) rc = dev_piece_finish(c, p, st, status, &zeroed);
        else if (p.job) {
            cudaStreamSynchronize(st);
            cudaStreamSynchronize(aux);
            cri_job_destroy(c, p.job);
            p.job = nullptr;
        }
    }
    trace_mark("finish-last");
    if (rc == OK && zeroed) CU_TRY(c, cudaStreamSynchronize(st));
    for (cudaEvent_t e : events) c->idle_events.push_back(e);
    tr.done("batch_dev", d.n);
    return rc;
}

static cri_job_desc make_desc_dev(int kind, const uint8_t* d_blob, const uint64_t* offsets, uint32_t n, uint8_t* d_out, void* stream) {
    cri_job_desc d = make_desc(kind, nullptr, offsets, n);
    d.d_blob = d_blob;
    d.d_out = d_out;
    d.stream = stream;
    return d;
}

extern "C" int cri_sizes_dev(cri_ctx* c, int kind, const uint8_t* d_blob, const uint64_t* off, uint32_t n, const cri_adx_params* adx,
                             uint32_t quality, uint64_t* sizes, int32_t* status, void* stream) {
    if (!c) return ERR_CUDA;
    if (!off || (n && !d_blob)) return ERR_BUFFER;
    CU_TRY(c, cudaSetDevice(c->device));
    if (kind == CRI_JOB_HCA_CRYPT) {
        for (uint32_t i = 0; i < n; i++) { sizes[i] = off[i + 1] - off[i]; if (status) status[i] = OK; }
        return OK;
    }
    uint8_t* shadow = nullptr;
    size_t bytes = 0;
    std::vector<std::pair<uint64_t, uint32_t>> dirty;
    int rc = build_shadow(c, kind, d_blob, off, n, stream ? (cudaStream_t)stream : c->stream, &shadow, &bytes, &dirty);
    if (rc == OK) switch (kind) {
        case CRI_JOB_ADX_DECODE: rc = cri_adx_decode_sizes(shadow, off, n, sizes, status); break;
        case CRI_JOB_ADX_ENCODE: rc = adx ? cri_adx_encode_sizes(shadow, off, n, adx, sizes, status) : ERR_BUFFER; break;
        case CRI_JOB_HCA_DECODE: rc = cri_hca_decode_sizes(shadow, off, n, sizes, status); break;
        case CRI_JOB_HCA_ENCODE: rc = cri_hca_encode_sizes_ex(shadow, off, n, quality, adx ? adx->force_not_looping : 0, sizes, status); break;
        default: rc = ERR_UNSUPPORTED;
    }
    release_shadow(c, shadow, bytes, dirty);
    return rc;
}

extern "C" int cri_adx_decode_batch_dev(cri_ctx* c, const uint8_t* d_blob, const uint64_t* off, uint32_t n, uint8_t* d_out,
                                        const uint64_t* out_off, int32_t* status, void* stream) {
    return run_batch_dev(c, make_desc_dev(CRI_JOB_ADX_DECODE, d_blob, off, n, d_out, stream), out_off, status);
}
extern "C" int cri_adx_encode_batch_dev(cri_ctx* c, const uint8_t* d_blob, const uint64_t* off, uint32_t n, const cri_adx_params* p,
                                        uint8_t* d_out, const uint64_t* out_off, int32_t* status, void* stream) {
    cri_job_desc d = make_desc_dev(CRI_JOB_ADX_ENCODE, d_blob, off, n, d_out, stream);
    d.adx = *p;
    return run_batch_dev(c, d, out_off, status);
}
extern "C" int cri_hca_decode_batch_dev(cri_ctx* c, const uint8_t* d_blob, const uint64_t* off, uint32_t n, const uint64_t* keys,
                                        const uint16_t* subkeys, uint8_t* d_out, const uint64_t* out_off, int32_t* status, void* stream) {
    cri_job_desc d = make_desc_dev(CRI_JOB_HCA_DECODE, d_blob, off, n, d_out, stream);
    d.keys = keys;
    d.subkeys = subkeys;
    return run_batch_dev(c, d, out_off, status);
}
extern "C" int cri_hca_crypt_batch_dev(cri_ctx* c, const uint8_t* d_blob, const uint64_t* off, uint32_t n, int encrypt,
                                       uint32_t ciph_type, const uint64_t* keys, const uint16_t* subkeys, uint8_t* d_out,
                                       int32_t* status, void* stream) {
    cri_job_desc d = make_desc_dev(CRI_JOB_HCA_CRYPT, d_blob, off, n, d_out, stream);
    d.keys = keys;
    d.subkeys = subkeys;
    d.encrypt = encrypt;
    d.ciph_type = ciph_type;
    return run_batch_dev(c, d, off, status);
}
extern "C" int cri_hca_encode_batch_dev(cri_ctx* c, const uint8_t* d_blob, const uint64_t* off, uint32_t n, uint32_t quality,
                                        uint32_t force_not_looping, uint8_t* d_out, const uint64_t* out_off, int32_t* status, void* stream) {
    cri_job_desc d = make_desc_dev(CRI_JOB_HCA_ENCODE, d_blob, off, n, d_out, stream);
    d.quality = quality;
    d.adx.force_not_looping = force_not_looping;
    return run_batch_dev(c, d, out_off, status);
}

// ------------------------------------------------------ single-stream API
static int run_single(cri_ctx* c, cri_job_desc d, size_t n, uint8_t** out, size_t* out_n) {
    *out = nullptr;
    *out_n = 0;
    if (!c) return ERR_CUDA;
    const uint64_t off[2] = {0, n};
    d.offsets = off;
    d.n = 1;
    cri_job* j = nullptr;
    int rc = cri_job_create(c, &d, &j);
    if (rc != OK) return rc;
    int32_t st = OK;
    if (j->status[0] != OK) {
        st = j->status[0];
    } else {
        rc = cri_job_run(c, j);
        if (rc == OK) {
            uint8_t* buf = (uint8_t*)malloc(std::max<uint64_t>(j->out_bytes, 1));
            rc = buf ? cri_job_download(c, j, buf, &st) : ERR_BUFFER;
            if (rc == OK && st == OK) {
                *out = buf;
                *out_n = j->out_bytes;
            } else {
                free(buf);
            }
        }
    }
    cri_job_destroy(c, j);
    return rc != OK ? rc : st;
}

extern "C" int cri_adx_decode(cri_ctx* c, const uint8_t* in, size_t n, uint8_t** out, size_t* out_n) {
    return run_single(c, make_desc(CRI_JOB_ADX_DECODE, in, nullptr, 1), n, out, out_n);
}
extern "C" int cri_adx_encode(cri_ctx* c, const uint8_t* in, size_t n, const cri_adx_params* p, uint8_t** out, size_t* out_n) {
    cri_job_desc d = make_desc(CRI_JOB_ADX_ENCODE, in, nullptr, 1);
    d.adx = *p;
    return run_single(c, d, n, out, out_n);
}
extern "C" int cri_hca_decode(cri_ctx* c, const uint8_t* in, size_t n, uint64_t key, uint16_t subkey, uint8_t** out, size_t* out_n) {
    cri_job_desc d = make_desc(CRI_JOB_HCA_DECODE, in, nullptr, 1);
    d.keys = &key;
    d.subkeys = &subkey;
    return run_single(c, d, n, out, out_n);
}
extern "C" int cri_hca_crypt(cri_ctx* c, const uint8_t* in, size_t n, int encrypt, uint32_t ciph_type, uint64_t key,
                             uint16_t subkey, uint8_t** out, size_t* out_n) {
    cri_job_desc d = make_desc(CRI_JOB_HCA_CRYPT, in, nullptr, 1);
    d.keys = &key;
    d.subkeys = &subkey;
    d.encrypt = encrypt;
    d.ciph_type = ciph_type;
    return run_single(c, d, n, out, out_n);
}
extern "C" int cri_hca_encode(cri_ctx* c, const uint8_t* in, size_t n, uint32_t quality, uint32_t force_not_looping,
                              uint8_t** out, size_t* out_n) {
    cri_job_desc d = make_desc(CRI_JOB_HCA_ENCODE, in, nullptr, 1);
    d.quality = quality;
    d.adx.force_not_looping = force_not_looping;
    return run_single(c, d, n, out, out_n);
}
