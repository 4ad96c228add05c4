// Internal definitions behind the opaque handles of include/cricodecs_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/cricodecs_b200.h"
#include "formats.h"
#include "hca_tables_dev.h"
#include "kernels.h"

struct cri_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {};       // [0],[1] whole run; [2],[3] dominant kernel
    uint64_t launches = 0;
    float last_ms = 0.f, last_dominant_ms = 0.f;
    std::string error;
};

struct cri_job {
    int kind = 0;
    uint32_t n = 0;
    const uint8_t* blob = nullptr;            // caller's host input (borrowed for the job's lifetime)
    std::vector<uint64_t> in_off, out_off;
    uint64_t in_bytes = 0, out_bytes = 0, units = 0;
    std::vector<int32_t> status;              // host-side (header / parameter) status per stream
    std::vector<uint64_t> keys;
    std::vector<uint16_t> subkeys;
    cri_adx_params adx = {};
    uint32_t quality = 1;
    int encrypt = 0;
    uint32_t ciph_type = 0;
    bool needs_clear = false;                 // some output bytes are not produced by kernels / patches

    uint8_t* d_in = nullptr;
    uint8_t* d_out = nullptr;
    int32_t* d_status = nullptr;

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
void finish_layout_public(cri_job* j, const std::vector<uint64_t>& sizes);
void add_patch_public(cri_job* j, uint64_t dst, const uint8_t* bytes, uint32_t n);
int plan_hca_decode(cri_ctx* c, cri_job* j);
int plan_hca_crypt(cri_ctx* c, cri_job* j);
int plan_hca_encode(cri_ctx* c, cri_job* j);
int upload_hca_tables(cri_ctx* c, cri_job* j);
int run_hca(cri_ctx* c, cri_job* j, bool* have_dominant);
void free_hca_tables(cri_job* j);
}  // namespace cri
