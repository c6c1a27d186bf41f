// debug_kernels.cu -- kernel-level entry points used by the parity tests and by bench.py's roofline
// microbenchmarks: explicit job lists for the two DP kernels (same device functions as production), the
// random 64-byte gather microbenchmark that gives the seeding roofline denominator, and the DPX issue
// microbenchmark that gives the SW roofline denominator.
#include "pipeline.cuh"
#include "ksw_warp.cuh"
#include "ksw_thread.cuh"
#include "debug_kernels.cuh"

namespace {

__global__ void __launch_bounds__(128) k_dbg_extend(DevOpts o, uint32_t n_jobs, const uint8_t* q, const uint64_t* q_off, const uint8_t* t,
                                                    const uint64_t* t_off, const int* w, const int* end_bonus, const int* h0, int* out,
                                                    int* eh, uint32_t max_q, uint32_t* ticket, unsigned long long* cells_out) {
    __shared__ int smat[25];
    if (threadIdx.x < 25) smat[threadIdx.x] = o.mat[threadIdx.x];
    __syncthreads();
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int* ehh = eh + (size_t)gwarp * 2 * (max_q + 2);
    int* ehe = ehh + (max_q + 2);
    unsigned long long cells = 0, rows = 0;
    for (;;) {
        uint32_t j = next_ticket(ticket);
        if (j >= n_jobs) break;
        ExtOut e = ksw_extend_warp_t<false>(o, (int)(q_off[j + 1] - q_off[j]), q + q_off[j], 1, (int)(t_off[j + 1] - t_off[j]), t + t_off[j], 1,
                                          w[j], end_bonus[j], h0[j], ehh, smat, cells, rows);
        if (lane_id() == 0) {
            int* d = out + (size_t)j * 6;
            d[0] = e.score; d[1] = e.qle; d[2] = e.tle; d[3] = e.gtle; d[4] = e.gscore; d[5] = e.max_off;
        }
    }
    if (cells_out && lane_id() == 0) atomicAdd(cells_out, cells);
}

// the thread-per-extension kernel of the production pre-pass (ksw_thread.cuh) on an explicit job list: one job per thread
struct ByteRows { const uint8_t* t; __device__ __forceinline__ int operator()(int i) const { return (int)t[i]; } };
constexpr int DBG_NCOL = EXT_MEMO_MAXQ + 1;
__global__ void __launch_bounds__(128) k_dbg_extend_thread(DevOpts o, uint32_t n_jobs, const uint8_t* q, const uint64_t* q_off, const uint8_t* t,
                                                           const uint64_t* t_off, const int* w, const int* end_bonus, const int* h0, int* out, int reversed) {
    extern __shared__ uint32_t dbg_smem[];
    uint32_t* eh = dbg_smem + threadIdx.x;
    uint32_t* qn = dbg_smem + DBG_NCOL * 128 + threadIdx.x;
    const uint32_t j = blockIdx.x * 128 + threadIdx.x;
    if (j >= n_jobs) return;
    const int qlen = (int)(q_off[j + 1] - q_off[j]);
    // reversed: the caller stored the query back to front, the loader's left-extension path puts it in order again
    if (reversed) ksw_thread_load_query<128>(qn, q + q_off[j] + qlen, qlen, -1);
    else ksw_thread_load_query<128>(qn, q + q_off[j], qlen, 1);
    ByteRows T; T.t = t + t_off[j];
    uint32_t cells = 0, rows = 0;
    const ExtOut e = ksw_extend_thread<128>(o, eh, qn, qlen, (int)(t_off[j + 1] - t_off[j]), T, w[j], end_bonus[j], h0[j], cells, rows);
    int* d = out + (size_t)j * 6;
    d[0] = e.score; d[1] = e.qle; d[2] = e.tle; d[3] = e.gtle; d[4] = e.gscore; d[5] = e.max_off;
}

__global__ void __launch_bounds__(128) k_dbg_global(DevOpts o, uint32_t n_jobs, const uint8_t* q, const uint64_t* q_off, const uint8_t* t,
                                                    const uint64_t* t_off, const int* w, int* out_score, uint32_t* cigar, uint32_t cig_cap,
                                                    int* n_cigar_out, int* eh, uint32_t max_q, uint8_t* zbuf, size_t z_per_warp, uint32_t* ticket) {
    __shared__ int smat[25];
    if (threadIdx.x < 25) smat[threadIdx.x] = o.mat[threadIdx.x];
    __syncthreads();
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int* ehh = eh + (size_t)gwarp * 2 * (max_q + 2);
    int* ehe = ehh + (max_q + 2);
    uint8_t* z = zbuf + (size_t)gwarp * z_per_warp;
    unsigned long long cells = 0;
    for (;;) {
        uint32_t jb = next_ticket(ticket);
        if (jb >= n_jobs) break;
        const int qlen = (int)(q_off[jb + 1] - q_off[jb]), tlen = (int)(t_off[jb + 1] - t_off[jb]);
        const int ww = w[jb];
        const int n_col = qlen < 2 * ww + 1 ? qlen : 2 * ww + 1;
        int sc = ksw_global_warp(o, qlen, q + q_off[jb], 1, tlen, t + t_off[jb], 1, ww, ehh, ehe, smat, cigar ? z : nullptr, n_col, cells);
        if (lane_id() == 0) {
            out_score[jb] = sc;
            if (cigar) {
                uint32_t* cg = cigar + (size_t)jb * cig_cap;
                int n = 0, which = 0, i = tlen - 1, k = (i + ww + 1 < qlen ? i + ww + 1 : qlen) - 1;
                auto push = [&](uint32_t op, uint32_t len) {
                    if (n == 0 || op != (cg[n - 1] & 0xf)) { if ((uint32_t)n < cig_cap) cg[n] = len << 4 | op; ++n; }
                    else cg[n - 1] += len << 4;
                };
                while (i >= 0 && k >= 0) {
                    which = z[(size_t)i * n_col + (k - (i > ww ? i - ww : 0))] >> (which << 1) & 3;
                    if (which == 0) { push(0, 1); --i; --k; }
                    else if (which == 1) { push(2, 1); --i; }
                    else { push(1, 1); --k; }
                }
                if (i >= 0) push(2, (uint32_t)(i + 1));
                if (k >= 0) push(1, (uint32_t)(k + 1));
                if ((uint32_t)n > cig_cap) n = (int)cig_cap;
                for (int a = 0; a < n >> 1; ++a) { uint32_t tt = cg[a]; cg[a] = cg[n - 1 - a]; cg[n - 1 - a] = tt; }
                n_cigar_out[jb] = n;
            }
        }
        __syncwarp();
    }
}

// random 64-byte block reads: 16 lanes per block, ILP independent loads per lane
template <int ILP>
__global__ void __launch_bounds__(256) k_gather(const uint32_t* __restrict__ occ, uint64_t n_blocks, uint64_t loads_per_group, uint32_t* sink) {
    const uint64_t group = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int idx = threadIdx.x & 15;
    uint64_t s = group * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    uint32_t acc = 0;
    for (uint64_t it = 0; it < loads_per_group; it += ILP) {
        uint32_t v[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
            uint64_t b = __umul64hi(s * 0x2545F4914F6CDD1Dull, n_blocks);   // uniform in [0, n_blocks) without a division
            v[u] = __ldg(occ + (b << 4) + idx);
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) acc ^= v[u];
    }
    if (acc == 0x7fffffffu) *sink = acc;
}

// DPX issue microbenchmark: independent chains of __vimax3_s32 / __viaddmax_s32_relu per thread
__global__ void __launch_bounds__(256) k_dpx(int iters, int seed, int* sink) {
    int a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const int b = seed * 3 + 1, c = seed - 7;
    for (int i = 0; i < iters; ++i) {
        a0 = __vimax3_s32(a0, b, c + i); a1 = __viaddmax_s32_relu(a1, b, c); a2 = __vimax3_s32(a2, c, b - i); a3 = __viaddmax_s32_relu(a3, c, b);
        a4 = __vimax3_s32(a4, b, c - i); a5 = __viaddmax_s32_relu(a5, b, c); a6 = __vimax3_s32(a6, c, b + i); a7 = __viaddmax_s32_relu(a7, c, b);
    }
    int r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (r == 0x12345678) *sink = r;
}

}  // namespace

void launch_dbg_extend(const DevOpts& o, uint32_t n_jobs, const uint8_t* q, const uint64_t* q_off, const uint8_t* t, const uint64_t* t_off, const int* w,
                       const int* end_bonus, const int* h0, int* out, int* eh, uint32_t max_q, uint32_t* ticket, unsigned long long* cells,
                       int blocks, cudaStream_t st) {
    k_dbg_extend<<<blocks, 128, 0, st>>>(o, n_jobs, q, q_off, t, t_off, w, end_bonus, h0, out, eh, max_q, ticket, cells);
}
void launch_dbg_extend_thread(const DevOpts& o, uint32_t n_jobs, const uint8_t* q, const uint64_t* q_off, const uint8_t* t, const uint64_t* t_off, const int* w,
                              const int* end_bonus, const int* h0, int* out, int reversed, cudaStream_t st) {
    const size_t smem = (size_t)(DBG_NCOL + (DBG_NCOL + 6) / 8) * 128 * 4;
    cudaFuncSetAttribute(k_dbg_extend_thread, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_dbg_extend_thread<<<(n_jobs + 127) / 128, 128, smem, st>>>(o, n_jobs, q, q_off, t, t_off, w, end_bonus, h0, out, reversed);
}
void launch_dbg_global(const DevOpts& o, uint32_t n_jobs, const uint8_t* q, const uint64_t* q_off, const uint8_t* t, const uint64_t* t_off, const int* w,
                       int* out_score, uint32_t* cigar, uint32_t cig_cap, int* n_cigar, int* eh, uint32_t max_q, uint8_t* z, size_t z_per_warp,
                       uint32_t* ticket, int blocks, cudaStream_t st) {
    k_dbg_global<<<blocks, 128, 0, st>>>(o, n_jobs, q, q_off, t, t_off, w, out_score, cigar, cig_cap, n_cigar, eh, max_q, z, z_per_warp, ticket);
}
void launch_gather(const uint32_t* occ, uint64_t n_blocks, uint64_t loads_per_group, int blocks, uint32_t* sink, cudaStream_t st) {
    k_gather<8><<<blocks, 256, 0, st>>>(occ, n_blocks, loads_per_group, sink);
}
void launch_dpx(int iters, int blocks, int* sink, cudaStream_t st) { k_dpx<<<blocks, 256, 0, st>>>(iters, 12345, sink); }
