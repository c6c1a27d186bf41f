// loader.cuh -- batch text -> NUCLSEQ datum images on the GPU (loader.cu).
#pragma once
#include "common.cuh"

struct LoaderParams {
    const uint8_t* text; const uint64_t* offs; uint64_t n_seqs;   // texts back to back, offs[n_seqs + 1]
    const uint64_t* chunk_off; uint64_t n_chunks;                 // chunk_off[r] = sum of ceil(len / 16) over the sequences before r (n_seqs + 1 entries)
    uint32_t* cnt_amb; uint32_t* cnt_start;                       // n_chunks + 1 each: counts, then exclusive prefix sums
    uint32_t* holes_num;                                          // per sequence
    uint64_t* img_off;                                            // n_seqs + 1: image sizes, then byte offsets
    unsigned long long* first_invalid;                            // smallest text position holding a letter outside "ACGTNWSMKRYBDHV" (~0 = none)
    uint8_t* bytes;                                               // the images (zero-filled before the fill kernels)
};
size_t loader_scan_tmp_elems(uint64_t n);
void launch_loader_scan(const LoaderParams& P, uint32_t* tmp32, uint64_t* tmp64, cudaStream_t st, uint64_t* launches);
void launch_loader_fill(const LoaderParams& P, cudaStream_t st, uint64_t* launches);
