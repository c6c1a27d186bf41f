// debug_kernels.cuh -- launchers of the kernel-level test / microbenchmark entry points.
#pragma once
#include "common.cuh"
void launch_dbg_extend(const DevOpts& o, uint32_t n_jobs, const uint8_t* q, const uint64_t* q_off, const uint8_t* t, const uint64_t* t_off, const int* w,
                       const int* end_bonus, const int* h0, int* out, int* eh, uint32_t max_q, uint32_t* ticket, unsigned long long* cells,
                       int blocks, cudaStream_t st);
void launch_dbg_extend_thread(const DevOpts& o, uint32_t n_jobs, const uint8_t* q, const uint64_t* q_off, const uint8_t* t, const uint64_t* t_off, const int* w,
                              const int* end_bonus, const int* h0, int* out, int reversed, cudaStream_t st);
void launch_dbg_global(const DevOpts& o, uint32_t n_jobs, const uint8_t* q, const uint64_t* q_off, const uint8_t* t, const uint64_t* t_off, const int* w,
                       int* out_score, uint32_t* cigar, uint32_t cig_cap, int* n_cigar, int* eh, uint32_t max_q, uint8_t* z, size_t z_per_warp,
                       uint32_t* ticket, int blocks, cudaStream_t st);
void launch_gather(const uint32_t* occ, uint64_t n_blocks, uint64_t loads_per_group, int blocks, uint32_t* sink, cudaStream_t st);
void launch_dpx(int iters, int blocks, int* sink, cudaStream_t st);
