// Internal definitions behind the opaque handles of include/cricodecs_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/cricodecs_b200.h"
#include "formats.h"
#include "hca_tables_dev.h"
#include "hca_kernels.h"
#include "kernels.h"

// Size-bucketed cache of HBM blocks owned by a context: a batch call needs a
// handful of large buffers whose sizes repeat from call to call, and
// cudaMalloc/cudaFree of gigabyte blocks costs more than the kernels.
struct DevPool {
    std::multimap<size_t, void*> idle;          // size -> block
    std::unordered_map<void*, size_t> live;     // block -> size
    size_t idle_bytes = 0;
};

constexpr int kPipeDepth = 3;                   // chunks of one batch call in flight (copy in / kernels / copy out)

struct cri_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;              // jobs created through cri_job_create
    cudaStream_t pipe[kPipeDepth] = {};         // chunk streams of the one-call batch entry points
    int32_t* pin_status[kPipeDepth] = {};       // page-locked landing buffers of the chunks' status words (a copy to
    size_t pin_status_cap[kPipeDepth] = {};     //  pageable memory would block the host until the chunk's kernels end)
    DevPool pool;
    cri::HcaFastPipe hca_pipe;                  // side stream + events of the pipelined HCA decode (created on first use)
    std::vector<cudaEvent_t> idle_events;       // events of finished jobs, reused by the next ones
    std::vector<std::pair<uint8_t*, size_t>> idle_shadows;   // header shadows of finished device-pointer jobs (all zero again)
    uint8_t* pin_stage = nullptr;               // page-locked landing buffer of the header fetches of device-pointer jobs
    size_t pin_stage_cap = 0;
    bool poison = false;                        // CRI_POISON=1 (tests): output blobs start as 0xA5 so that a byte no kernel writes shows
    bool trace = false;                         // CRI_TRACE=1: phase times of every batch call on stderr
    uint64_t launches = 0;
    float last_ms = 0.f, last_dominant_ms = 0.f;
    std::string error;
};

struct cri_job {
    int kind = 0;
    uint32_t n = 0;
    const uint8_t* blob = nullptr;            // caller's host input (borrowed for the job's lifetime)
    std::vector<uint64_t> in_off, out_off;    // out_off: offsets from d_out (start at out_delta)
    std::vector<uint64_t> out_off_pub;        // the same from the start of the output blob (what callers see)
    uint64_t in_bytes = 0, out_bytes = 0, units = 0;   // out_bytes: size of the output blob (without out_delta)
    uint64_t out_delta = 0;                   // device-pointer jobs: the caller's buffer starts this many bytes behind the
                                              //  256-byte aligned address the kernels take as the blob's start
    std::vector<int32_t> status;              // host-side (header / parameter) status per stream
    std::vector<uint64_t> keys;
    std::vector<uint16_t> subkeys;
    cri_adx_params adx = {};
    uint32_t quality = 1;
    int encrypt = 0;
    uint32_t ciph_type = 0;
    bool needs_clear = false;                 // some output bytes are not produced by kernels / patches

    cudaStream_t stream = nullptr;            // every copy and launch of this job is ordered on it
    cudaEvent_t ev[4] = {};                   // [0],[1] whole run; [2],[3] dominant kernel
    bool have_dominant = false;
    std::vector<int32_t> dev_status;          // landing buffer of d_status (jobs of the create/run/download API)
    int32_t* h_status = nullptr;              // where d_status lands: dev_status.data() or a page-locked chunk buffer
    std::vector<uint8_t> staging;             // download target when the caller's layout is not the packed one
    uint8_t* dl_out = nullptr;                // where the pending download lands
    int32_t* dl_status = nullptr;

    uint8_t* d_in = nullptr;
    uint8_t* d_out = nullptr;
    int32_t* d_status = nullptr;

    // device-pointer jobs (cri_*_batch_dev): the input blob lives in the caller's HBM; `blob` then points at a sparse
    // host shadow that holds only the fetched header bytes, and d_out is the caller's buffer
    const uint8_t* d_src = nullptr;
    uint8_t* shadow = nullptr;
    size_t shadow_bytes = 0;
    std::vector<std::pair<uint64_t, uint32_t>> shadow_dirty;   // fetched ranges: zeroed again when the shadow goes back to the context
    bool own_out = true;
    struct DevCopy { uint64_t src_off, dst_off, bytes; };
    std::vector<DevCopy> dev_copies;          // input bytes that pass through unchanged (in place of host-built patches)

    // WAV ingest: streams whose samples are not PCM16 get a converted copy behind the input blob (region starts at conv_base)
    std::vector<cri::PcmConv> conv;
    cri::PcmConv* d_conv = nullptr;
    uint64_t conv_base = 0, conv_bytes = 0;
    uint32_t conv_max_count = 0;
    std::vector<cri::Patch> patches;          // host-built headers / trailers, scattered on every run
    std::vector<uint8_t> patch_bytes;
    cri::Patch* d_patches = nullptr;
    uint8_t* d_patch_bytes = nullptr;

    // ADX
    std::vector<cri::AdxChain> adx_chains;    // fast-path chains first, then generic
    uint32_t n_fast = 0, n_generic = 0;
    cri::AdxChain* d_adx_chains = nullptr;

    // HCA (hca_engine.cu)
    cri::HcaJob hca;
};

namespace cri {
int pool_alloc(cri_ctx* c, void** p, size_t bytes);
void pool_free(cri_ctx* c, void* p);
void pool_trim(cri_ctx* c);
void finish_layout_public(cri_job* j, const std::vector<uint64_t>& sizes);
void add_patch_public(cri_job* j, uint64_t dst, const uint8_t* bytes, uint32_t n);
// offset (into the device input blob) of stream i's PCM16 samples: the WAV data itself, or a converted copy that the
// conversion kernel fills before the encode kernels run
uint64_t pcm16_offset(cri_job* j, uint32_t i, const cri::WavInfo& w);
// the same for a looping HCA encode: the virtual input of the reference's frame feeder, assembled on the device
uint64_t hca_loop_input_offset(cri_job* j, uint32_t i, const cri::WavInfo& w, const cri::HcaEncPlan& p);
// header checks + exact WAV size of one HCA stream (cri_hca_decode_sizes and the planner agree by construction)
int hca_decode_size_one(const uint8_t* d, uint64_t len, cri::HcaInfo* h, uint64_t* size);
int plan_hca_decode(cri_ctx* c, cri_job* j);
int plan_hca_crypt(cri_ctx* c, cri_job* j);
int plan_hca_encode(cri_ctx* c, cri_job* j);
int upload_hca_tables(cri_ctx* c, cri_job* j);
int run_hca(cri_ctx* c, cri_job* j, bool* have_dominant);
void free_hca_tables(cri_ctx* c, cri_job* j);
}  // namespace cri
