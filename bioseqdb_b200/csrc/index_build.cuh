// index_build.cuh -- interface of the GPU FM-index construction (see index_build.cu).
#pragma once
#include "common.cuh"

struct IndexBuild {
    // in
    const uint8_t* d_pac = nullptr;  // device, l_pac / 4 bytes
    int64_t l_pac = 0;
    bool force_wide = false;         // take the 64-bit-id path regardless of size (tests)
    // out (device allocations owned by the caller after success)
    uint32_t* d_occ = nullptr; uint64_t occ_bytes = 0;
    void* d_sa = nullptr; int sa_bytes = 4;
    uint64_t seq_len = 0, primary = 0, L2[5] = {0, 0, 0, 0, 0};
    double build_ms = 0; uint64_t launches = 0, sort_pass_bytes = 0; int doubling_rounds = 0;
};
int build_index_device(IndexBuild& B, cudaStream_t st);
