// abi.cu -- the extern "C" boundary (include/bioseqdb_gpu.h) and the host-side orchestration of the
// per-batch kernel pipeline: upload -> seed_smem -> chain_build -> sw_extend -> regs_finalize -> compact
// -> download (+ MAPQ, which needs libm's log).  Pools are bump-allocated on the device
// and grown by re-running the batch when a kernel raises the overflow flag; there is no CPU fallback for
// any stage.
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <chrono>
#include <algorithm>
#include <mutex>
#include "../../include/bioseqdb_gpu.h"
#include "common.cuh"
#include "index_build.cuh"
#include "seed.cuh"
#include "pipeline.cuh"
#include "index_verify.cuh"
#include "tuples.cuh"
#include "loader.cuh"
#include "primitives.cuh"
#include "debug_kernels.cuh"

static thread_local char g_err[1024] = "";
void bsq_set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
// every extern "C" entry point starts with a clean error string: bsq_last_error() after a call describes THAT call
#define BSQ_ENTRY() do { g_err[0] = 0; } while (0)

static_assert(sizeof(bsq_row) == sizeof(RowPub) && sizeof(bsq_row_ext) == sizeof(RowExt), "bsq_row / RowPub mismatch");
static_assert(sizeof(bsq_hole) == 16, "bsq_hole must match bntamb1_t");

namespace {

template <class T> struct DevBuf {
    T* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 16;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    size_t bytes() const { return cap * sizeof(T); }
};

struct Batch {
    uint64_t n = 0; uint32_t max_len = 0; uint64_t total_bases = 0;
    DevBuf<uint8_t> seqs; DevBuf<uint64_t> offs; DevBuf<int64_t> ids;
    DevBuf<uint8_t> datums; DevBuf<uint64_t> datum_off, scan_tmp64;
    DevBuf<uint8_t> ascii;              // the reads as text (what to_text_palloc yields): the row materialisation reads query_subseq from it
    uint64_t res_serial = 0, res_read_base = 0, res_row_base = 0, res_cig_base = 0;
    DevBuf<uint64_t> row_off64;        // the result's view of row_off: 64-bit, counted from the start of the whole result
    uint64_t cig_rebased = 0;          // what has been added to rows_compact[].cigar_off so far (reset when the rows are written again)
    DevBuf<uint64_t> tup_off, tup_tmp; DevBuf<uint32_t> tup_row_read, tup_nholes; DevBuf<int32_t> tup_rm; DevBuf<uint8_t> tup_bytes;   // row materialisation from the resident batch   // the result (and the place in it) this batch's rows went to   // reads handed over as NUCLSEQ datum images (upload_datums)
    DevBuf<Intv> intv; DevBuf<uint32_t> intv_cnt; uint32_t intv_cap = 0;
    DevBuf<Intv> seed_scratch; uint32_t list_cap = 0;
    DevBuf<SeedRec> raw, seeds; DevBuf<ChainTmp> ctmp; DevBuf<uint32_t> ord; DevBuf<ChainRec> chains; DevBuf<uint64_t> srt;
    DevBuf<RegRec> regs; DevBuf<RowDev> rows; DevBuf<RowPub> rows_compact; DevBuf<RowExt> rows_ext; DevBuf<uint32_t> reg_cnt, row_cnt, row_off, scan_tmp;
    DevBuf<ReadBlock> blocks; uint32_t pool_cap = 0;
    DevBuf<uint32_t> cigar; uint32_t cigar_cap = 0;
    DevBuf<double> read_logtab; uint32_t read_logtab_n = 0; uint32_t rseq_cap = 0;
    DevBuf<uint8_t> ext_scratch, fin_scratch, narrow_z, narrow_jobs; DevBuf<uint64_t> wide_jobs;
    DevBuf<ExtMemo> ext_memo; DevBuf<uint8_t> ext_memo_key; DevBuf<uint32_t> ext_memo_perm, ext_memo_hist, ext_todo, chain_todo, fin_todo, seed_todo, seed_pk, seed_u32;   // thread-per-extension pre-pass (extend_plan.cu)
    DevBuf<uint32_t> ctl;  // [0..3] tickets, [4] overflow, [5] pool_top, [6] cigar_top, [8..23] counters (u64 x 8), [24] narrow_cnt, [25] wide_cnt, [26..27] tickets, [56] reads left for sw_extend, [57] reads left for the warp chain kernel, [58] for regs_finalize
    size_t device_bytes() const {
        return seqs.bytes() + offs.bytes() + ids.bytes() + ascii.bytes() + datums.bytes() + datum_off.bytes() + scan_tmp64.bytes() + intv.bytes() + intv_cnt.bytes() + seed_scratch.bytes() + raw.bytes() + seeds.bytes() + ctmp.bytes() +
               ord.bytes() + chains.bytes() + srt.bytes() + regs.bytes() + rows.bytes() + rows_compact.bytes() + rows_ext.bytes() + reg_cnt.bytes() + row_cnt.bytes() + row_off.bytes() +
               scan_tmp.bytes() + blocks.bytes() + cigar.bytes() + read_logtab.bytes() + ext_scratch.bytes() + fin_scratch.bytes() + narrow_z.bytes() +
               narrow_jobs.bytes() + wide_jobs.bytes() + ctl.bytes() + ext_memo.bytes() + ext_memo_key.bytes() + ext_memo_perm.bytes() + ext_memo_hist.bytes() + ext_todo.bytes() + chain_todo.bytes() + fin_todo.bytes() + seed_todo.bytes() + seed_pk.bytes() + seed_u32.bytes();
    }
    bool resident = false, aligned = false;
    // one batch = one lane of the host pipeline: its stream, the host staging that must outlive the async copies,
    // and the state of the attempt in flight (pipeline_enqueue -> pipeline_check)
    cudaStream_t st = nullptr;
    std::vector<bsq_row_ext> ext_tmp;   // extension records fetched only to finish a MAPQ on the host
    uint32_t* ctl_host = nullptr;       // pinned, 16 words: ctl[0..7], total rows (2 words), host-MAPQ flag
    cudaEvent_t ev[5]; bool ev_ok = false;
    cudaEvent_t ev_x[5]; bool evx_ok = false;   // chunked mode: upload begin/end, download begin/end, pipeline queued
    ExtAux ext_aux; bool ext_aux_ok = false;    // side streams of the extension pre-pass
    uint32_t att_rseq_cap = 0; bool idle = true;
    uint64_t out_rows = 0; uint32_t out_cig = 0;
    uint64_t rows_cap = 0;              // capacity of rows_compact (rows of the batch in read order)
    void release() {
        if (ctl_host) { cudaFreeHost(ctl_host); ctl_host = nullptr; }
        if (ev_ok) { for (auto& e : ev) cudaEventDestroy(e); ev_ok = false; }
        if (evx_ok) { for (auto& e : ev_x) cudaEventDestroy(e); evx_ok = false; }
        if (ext_aux_ok) { for (auto& s_ : ext_aux.st) cudaStreamDestroy(s_); for (auto& e : ext_aux.ev) cudaEventDestroy(e); ext_aux_ok = false; }
        seqs.release(); offs.release(); ids.release(); ascii.release(); tup_off.release(); tup_tmp.release(); tup_row_read.release(); tup_nholes.release(); tup_rm.release(); tup_bytes.release(); datums.release(); datum_off.release(); scan_tmp64.release(); intv.release(); intv_cnt.release(); seed_scratch.release(); raw.release(); seeds.release();
        ctmp.release(); ord.release(); chains.release(); srt.release(); regs.release(); rows.release(); rows_compact.release(); rows_ext.release(); reg_cnt.release();
        row_cnt.release(); row_off.release(); scan_tmp.release(); blocks.release(); cigar.release(); ext_scratch.release(); fin_scratch.release();
        ctl.release(); narrow_z.release(); narrow_jobs.release(); wide_jobs.release(); read_logtab.release();
        ext_memo.release(); ext_memo_key.release(); ext_memo_perm.release(); ext_memo_hist.release(); ext_todo.release(); chain_todo.release(); fin_todo.release(); seed_todo.release(); seed_pk.release(); seed_u32.release();
    }
};

}  // namespace

struct bsq_index {
    int device = 0;
    bsq_opts opts; DevOpts dopts;
    float mapQ_coef_len = 50.f; int mapQ_coef_fac = 0;   // bwamem.h: `float mapQ_coef_len; int mapQ_coef_fac;`
    std::vector<uint8_t> pac; std::vector<int64_t> ann_offset, ann_id; std::vector<int32_t> ann_len; std::vector<bsq_hole> holes; std::vector<uint32_t> hole_ann;   // hole_ann: reference row of every hole
    uint8_t* d_pac = nullptr; uint32_t* d_occ = nullptr; void* d_sa = nullptr; int64_t* d_ann_offset = nullptr; int32_t* d_ann_len = nullptr; int64_t* d_ann_id = nullptr;
    bsq_index_meta meta;
    cudaStream_t stream = nullptr;
    Batch batch, batch2;             // batch2 + stream2: second lane of bsq_align_batch's chunk pipeline
    cudaStream_t stream2 = nullptr;
    bsq_timing timing;
    double* d_logtab = nullptr;
    void* d_isa = nullptr;           // inverse SA for the unique-match shortcut of the seeding kernel (built lazily, rows as wide as the SA's)
    uint32_t* d_kmer_z = nullptr;    // sizes-only copy of the prefix table
    void* d_kmer = nullptr; int kmer_k = 0;          // k-mer table of the LAST-like seeding pass (built lazily per device index)
    bool collect_counters = false;
    int kmer_k_cap = 15;                   // lowered when the batch pools did not fit beside the table (kmer_downgrade)
    bool alloc_failed = false;             // the last pipeline_enqueue stopped on cudaErrorMemoryAllocation
    int test_pool_oom = getenv("BSQ_TEST_POOL_OOM") ? atoi(getenv("BSQ_TEST_POOL_OOM")) : 0;   // test hook: that many pool allocations "fail"
    int prio_hi = 0, prio_lo = 0;          // stream priorities of the two lanes of the chunked pipeline
    uint32_t flags = 0;              // BSQ_FLAG_*
    uint64_t lrand_state = 0;        // glibc lrand48 state of the session (one draw per aligned read when the caller passes no ids)
    bool replica_pending = false;    // device arrays allocated by bsq_index_alloc_replica, host mirrors not yet rebuilt (bsq_index_replica_finish)
    uint64_t counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

static void fill_dev_opts(bsq_index* h) {
    const bsq_opts& p = h->opts; DevOpts& o = h->dopts;
    o.a = p.a; o.b = p.b; o.o_del = p.o_del; o.e_del = p.e_del; o.o_ins = p.o_ins; o.e_ins = p.e_ins;
    o.pen_clip5 = p.pen_clip5; o.pen_clip3 = p.pen_clip3; o.w = p.w; o.zdrop = p.zdrop; o.min_seed_len = p.min_seed_len; o.max_occ = p.max_occ;
    o.max_mem_intv = 20; o.split_width = 10;
    o.split_len = (int)(p.min_seed_len * 1.5f + .499);   // (int)(min_seed_len * split_factor + .499), split_factor is a float 1.5
    o.max_chain_gap = 10000; o.max_chain_extend = 1 << 30; o.min_chain_weight = 0;
    o.mask_level = 0.50f; o.drop_ratio = 0.50f; o.mask_level_redun = 0.95f;
    // bwa_fill_scmat(1, 4): filled once by mem_opt_init and never refreshed
    for (int i = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) o.mat[i * 5 + j] = i == j ? 1 : -4; o.mat[i * 5 + 4] = -1; }
    for (int j = 0; j < 5; ++j) o.mat[20 + j] = -1;
    o.mat_max = 1;
    h->mapQ_coef_len = 50.f; h->mapQ_coef_fac = (int)log((double)h->mapQ_coef_len);   // mem_opt_init stores log(50) = 3.912 in an int field: 3
}

static int check_opts(const bsq_opts* o) {
    const int32_t* v = &o->min_seed_len;
    static const char* names[12] = {"min_seed_len", "max_occ", "match_score", "mismatch_penalty", "pen_clip3", "pen_clip5", "zdrop", "bandwidth", "o_del", "e_del", "o_ins", "e_ins"};
    for (int i = 0; i < 12; ++i) if (v[i] < 0) { bsq_set_error("bwa_opt %s must be nonnegative", names[i]); return BSQ_ERR; }   // extension.cpp:205-206
    if (o->e_del == 0 || o->e_ins == 0) { bsq_set_error("bwa_opt e_del / e_ins must be positive (libbwa divides by them)"); return BSQ_ERR; }
    if (o->max_occ == 0) { bsq_set_error("bwa_opt max_occ must be positive (libbwa divides by it)"); return BSQ_ERR; }
    return BSQ_OK;
}

namespace { DevIndex make_dev_index(const bsq_index* h); int ensure_kmer_table(bsq_index* h, const DevIndex& ix); }

extern "C" {

const char* bsq_last_error(void) { return g_err; }

int bsq_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }

void bsq_opts_init(bsq_opts* o) {
    o->min_seed_len = 19; o->max_occ = 500; o->a = 1; o->b = 4; o->pen_clip3 = 5; o->pen_clip5 = 5; o->zdrop = 100; o->w = 100;
    o->o_del = 6; o->e_del = 1; o->o_ins = 6; o->e_ins = 1;
}

bsq_index* bsq_index_new(const bsq_opts* o, int device) {
    BSQ_ENTRY();
    bsq_opts d;
    if (!o) { bsq_opts_init(&d); o = &d; }
    if (check_opts(o) != BSQ_OK) return nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { bsq_set_error("no CUDA device: libbioseqdb_gpu has no CPU fallback"); return nullptr; }
    if (device < 0 || device >= n) { bsq_set_error("device %d out of range (%d devices)", device, n); return nullptr; }
    CUDA_CHECK_NULL(cudaSetDevice(device));
    bsq_index* h = new bsq_index;
    h->device = device; h->opts = *o;
    memset(&h->meta, 0, sizeof(h->meta)); memset(&h->timing, 0, sizeof(h->timing));
    fill_dev_opts(h);
    // The main stream (and lane 0 of the chunked pipeline) outranks lane 1: two chunks share the SMs, the earlier one gets free ones
    // first and is done -- its result on the way to the host -- while the later one still computes.  BSQ_NO_STREAM_PRIO: equal ranks.
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (getenv("BSQ_NO_STREAM_PRIO")) prio_hi = prio_lo;
    h->prio_hi = prio_hi; h->prio_lo = prio_lo;
    if (cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { bsq_set_error("cudaStreamCreate failed"); delete h; return nullptr; }
    h->batch.st = h->stream;
    return h;
}

int bsq_index_set_opts(bsq_index* h, const bsq_opts* o) {
    BSQ_ENTRY();
    if (!h || !o) { bsq_set_error("null argument"); return BSQ_ERR; }
    if (check_opts(o) != BSQ_OK) return BSQ_ERR;
    h->opts = *o; fill_dev_opts(h);
    return BSQ_OK;
}

int bsq_index_set_flags(bsq_index* h, uint32_t flags) {
    BSQ_ENTRY();
    if (!h) { bsq_set_error("null index"); return BSQ_ERR; }
    if (flags & ~(BSQ_FLAG_ROWS_EXT | BSQ_FLAG_TWO_CHUNKS)) { bsq_set_error("unknown flag bits %#x", flags & ~(BSQ_FLAG_ROWS_EXT | BSQ_FLAG_TWO_CHUNKS)); return BSQ_ERR; }
    h->flags = flags;
    return BSQ_OK;
}

int bsq_index_add_ref(bsq_index* h, int64_t id, const uint8_t* pac, uint32_t len, const bsq_hole* holes, uint32_t n_holes) {
    BSQ_ENTRY();
    if (!h) { bsq_set_error("null index"); return BSQ_ERR; }
    if (h->meta.built) { bsq_set_error("index already built"); return BSQ_ERR; }
    // BwaIndex::add_ref_sequence (bwa.cpp:82-105): offset = 4 * bytes so far, byte-rounded append, holes not rebased
    h->ann_offset.push_back((int64_t)h->pac.size() * 4);
    h->ann_len.push_back((int32_t)len);
    h->ann_id.push_back(id);
    size_t nb = (size_t)len / 4 + (len % 4 != 0);
    h->pac.insert(h->pac.end(), pac, pac + nb);
    for (uint32_t i = 0; i < n_holes; ++i) { h->holes.push_back(holes[i]); h->hole_ann.push_back((uint32_t)h->ann_offset.size() - 1); }
    return BSQ_OK;
}

// Batched form of bsq_index_add_ref: n NUCLSEQ datum images exactly as PostgreSQL hands them to iterate_nuclseq_table after
// detoasting (extension.cpp:157-195; layout sequence.h:18-38: varlena length word, holes_num, len, hole records, pac bytes),
// image i at bytes + off[i].  One call instead of one per reference row (500 k rows in BASELINE configs[4]).
int bsq_index_add_ref_datums(bsq_index* h, uint64_t n, const int64_t* ids, const uint8_t* bytes, const uint64_t* off) {
    BSQ_ENTRY();
    if (!h || (n && (!ids || !bytes || !off))) { bsq_set_error("null argument"); return BSQ_ERR; }
    if (h->meta.built) { bsq_set_error("index already built"); return BSQ_ERR; }
    size_t pac_total = 0, holes_total = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t hdr[3]; memcpy(hdr, bytes + off[i], 12);
        const uint64_t need = 12 + (uint64_t)hdr[1] * sizeof(bsq_hole) + ((uint64_t)hdr[2] + 3) / 4;
        if ((hdr[0] >> 2) < need) { bsq_set_error("datum %llu is truncated: varlena size %u, header needs %llu", (unsigned long long)i, hdr[0] >> 2, (unsigned long long)need); return BSQ_ERR; }
        pac_total += ((size_t)hdr[2] + 3) / 4; holes_total += hdr[1];
    }
    h->pac.reserve(h->pac.size() + pac_total); h->holes.reserve(h->holes.size() + holes_total); h->hole_ann.reserve(h->hole_ann.size() + holes_total);
    h->ann_offset.reserve(h->ann_offset.size() + n); h->ann_len.reserve(h->ann_len.size() + n); h->ann_id.reserve(h->ann_id.size() + n);
    for (uint64_t i = 0; i < n; ++i) {
        const uint8_t* d = bytes + off[i];
        uint32_t hdr[3]; memcpy(hdr, d, 12);
        const uint32_t n_holes = hdr[1], len = hdr[2];
        const uint8_t* pac = d + 12 + (size_t)n_holes * sizeof(bsq_hole);
        // BwaIndex::add_ref_sequence (bwa.cpp:82-105), as bsq_index_add_ref
        h->ann_offset.push_back((int64_t)h->pac.size() * 4);
        h->ann_len.push_back((int32_t)len);
        h->ann_id.push_back(ids[i]);
        h->pac.insert(h->pac.end(), pac, pac + ((size_t)len + 3) / 4);
        for (uint32_t k = 0; k < n_holes; ++k) {
            bsq_hole hl; memcpy(&hl, d + 12 + (size_t)k * sizeof(bsq_hole), sizeof(bsq_hole));   // 4-byte aligned only inside a datum (SURVEY B#11)
            h->holes.push_back(hl); h->hole_ann.push_back((uint32_t)h->ann_offset.size() - 1);
        }
    }
    return BSQ_OK;
}

static void free_index_arrays(bsq_index* h) {
    if (h->d_kmer) { cudaFree(h->d_kmer); h->d_kmer = nullptr; }
    if (h->d_kmer_z) { cudaFree(h->d_kmer_z); h->d_kmer_z = nullptr; }
    if (h->d_isa) { cudaFree(h->d_isa); h->d_isa = nullptr; }
    if (h->d_pac) cudaFree(h->d_pac); if (h->d_occ) cudaFree(h->d_occ); if (h->d_sa) cudaFree(h->d_sa);
    if (h->d_ann_offset) cudaFree(h->d_ann_offset); if (h->d_ann_len) cudaFree(h->d_ann_len); if (h->d_ann_id) cudaFree(h->d_ann_id);
    h->d_pac = nullptr; h->d_occ = nullptr; h->d_sa = nullptr; h->d_ann_offset = nullptr; h->d_ann_len = nullptr; h->d_ann_id = nullptr;
}

static int upload_anns(bsq_index* h) {
    size_t na = h->ann_offset.size();
    CUDA_CHECK(cudaMalloc(&h->d_ann_offset, na * 8 + 8)); CUDA_CHECK(cudaMalloc(&h->d_ann_len, na * 4 + 8)); CUDA_CHECK(cudaMalloc(&h->d_ann_id, na * 8 + 8));
    CUDA_CHECK(cudaMemcpyAsync(h->d_ann_offset, h->ann_offset.data(), na * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(h->d_ann_len, h->ann_len.data(), na * 4, cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(h->d_ann_id, h->ann_id.data(), na * 8, cudaMemcpyHostToDevice, h->stream));
    return BSQ_OK;
}

int bsq_index_build(bsq_index* h) {
    BSQ_ENTRY();
    if (!h) { bsq_set_error("null index"); return BSQ_ERR; }
    if (h->pac.empty()) return BSQ_OK;   // bwa.cpp:108-109: empty reference => no index, alignments return nothing
    CUDA_CHECK(cudaSetDevice(h->device));
    free_index_arrays(h);
    CUDA_CHECK(cudaMalloc(&h->d_pac, h->pac.size() + 64));
    CUDA_CHECK(cudaMemcpyAsync(h->d_pac, h->pac.data(), h->pac.size(), cudaMemcpyHostToDevice, h->stream));
    if (upload_anns(h) != BSQ_OK) return BSQ_ERR;
    IndexBuild B;
    B.d_pac = h->d_pac; B.l_pac = (int64_t)h->pac.size() * 4;
    { const char* fw = getenv("BSQ_FORCE_WIDE"); B.force_wide = fw && fw[0] == '1'; }   // test hook: 64-bit index path on small inputs
    if (build_index_device(B, h->stream) != BSQ_OK) return BSQ_ERR;
    h->d_occ = B.d_occ; h->d_sa = B.d_sa;
    bsq_index_meta& m = h->meta;
    m.l_pac = B.l_pac; m.seq_len = B.seq_len; m.primary = B.primary; memcpy(m.L2, B.L2, sizeof(m.L2));
    m.n_anns = h->ann_offset.size(); m.sa_bytes = (uint32_t)B.sa_bytes; m.built = 1;
    m.arr_bytes[BSQ_ARR_PAC] = h->pac.size(); m.arr_bytes[BSQ_ARR_OCC] = B.occ_bytes; m.arr_bytes[BSQ_ARR_SA] = (B.seq_len + 1) * B.sa_bytes;
    m.arr_bytes[BSQ_ARR_ANN_OFFSET] = m.n_anns * 8; m.arr_bytes[BSQ_ARR_ANN_LEN] = m.n_anns * 4; m.arr_bytes[BSQ_ARR_ANN_ID] = m.n_anns * 8;
    m.build_ms = B.build_ms; m.build_launches = B.launches; m.sort_pass_bytes = B.sort_pass_bytes;
    return BSQ_OK;
}

void bsq_index_free(bsq_index* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    h->batch.release(); h->batch2.release();
    free_index_arrays(h);
    if (h->d_logtab) cudaFree(h->d_logtab);
    if (h->d_kmer) cudaFree(h->d_kmer);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    delete h;
}

int bsq_index_get_meta(const bsq_index* h, bsq_index_meta* m) {
    BSQ_ENTRY();
    if (!h || !m) { bsq_set_error("null argument"); return BSQ_ERR; }
    *m = h->meta;
    return BSQ_OK;
}

int bsq_index_device_bytes(const bsq_index* h, uint64_t* bytes) {
    BSQ_ENTRY();
    if (!h || !bytes) { bsq_set_error("null argument"); return BSQ_ERR; }
    uint64_t b = h->batch.device_bytes() + h->batch2.device_bytes();
    if (h->meta.built) {
        const uint64_t n = h->meta.seq_len;
        b += (h->meta.l_pac + 3) / 4 + ((n + 127) / 128 + 1) * 64 + (n + 1) * h->meta.sa_bytes + h->meta.n_anns * 20;
        if (h->d_isa) b += (n + 1) * (uint64_t)h->meta.sa_bytes + 64;
        if (h->d_kmer) b += kmer_table_bytes(h->kmer_k);
        if (h->d_kmer_z) b += kmer_table_bytes(h->kmer_k) / 4;
    }
    *bytes = b;
    return BSQ_OK;
}

static void* index_array(const bsq_index* h, int what) {
    switch (what) {
        case BSQ_ARR_PAC: return h->d_pac; case BSQ_ARR_OCC: return h->d_occ; case BSQ_ARR_SA: return h->d_sa;
        case BSQ_ARR_ANN_OFFSET: return h->d_ann_offset; case BSQ_ARR_ANN_LEN: return h->d_ann_len; case BSQ_ARR_ANN_ID: return h->d_ann_id;
        default: return nullptr;
    }
}

int bsq_index_device_ptr(const bsq_index* h, int what, void** dptr) {
    BSQ_ENTRY();
    if (!h || !dptr || what < 0 || what >= BSQ_ARR_COUNT) { bsq_set_error("bad argument"); return BSQ_ERR; }
    *dptr = index_array(h, what);
    return BSQ_OK;
}

int bsq_index_download(const bsq_index* h, int what, void* dst, uint64_t bytes) {
    BSQ_ENTRY();
    if (!h || !h->meta.built || what < 0 || what >= BSQ_ARR_COUNT) { bsq_set_error("index not built / bad array"); return BSQ_ERR; }
    if (bytes > h->meta.arr_bytes[what]) { bsq_set_error("download of %llu bytes exceeds array size %llu", (unsigned long long)bytes, (unsigned long long)h->meta.arr_bytes[what]); return BSQ_ERR; }
    CUDA_CHECK(cudaSetDevice(h->device));
    CUDA_CHECK(cudaMemcpy(dst, index_array(h, what), bytes, cudaMemcpyDeviceToHost));
    return BSQ_OK;
}

int bsq_index_alloc_replica(bsq_index* h, const bsq_index_meta* m) {
    BSQ_ENTRY();
    if (!h || !m || !m->built) { bsq_set_error("bad argument"); return BSQ_ERR; }
    CUDA_CHECK(cudaSetDevice(h->device));
    free_index_arrays(h);
    h->meta = *m;
    h->replica_pending = true;
    CUDA_CHECK(cudaMalloc(&h->d_pac, m->arr_bytes[BSQ_ARR_PAC] + 64));
    CUDA_CHECK(cudaMalloc(&h->d_occ, m->arr_bytes[BSQ_ARR_OCC] + 64));
    CUDA_CHECK(cudaMalloc(&h->d_sa, m->arr_bytes[BSQ_ARR_SA] + 64));
    CUDA_CHECK(cudaMalloc(&h->d_ann_offset, m->arr_bytes[BSQ_ARR_ANN_OFFSET] + 8));
    CUDA_CHECK(cudaMalloc(&h->d_ann_len, m->arr_bytes[BSQ_ARR_ANN_LEN] + 8));
    CUDA_CHECK(cudaMalloc(&h->d_ann_id, m->arr_bytes[BSQ_ARR_ANN_ID] + 8));
    return BSQ_OK;
}

// Host-side state that has no device copy: the ambiguity holes of the reference rows (copied un-rebased, bwa.cpp:98-104) and the
// row each one came from.  Blob = u64 n_holes | n_holes x bsq_hole | n_holes x u32 row.
int bsq_index_host_state_size(const bsq_index* h, uint64_t* bytes) {
    BSQ_ENTRY();
    if (!h || !bytes) { bsq_set_error("null argument"); return BSQ_ERR; }
    *bytes = 8 + h->holes.size() * (sizeof(bsq_hole) + 4);
    return BSQ_OK;
}

int bsq_index_host_state_get(const bsq_index* h, void* buf, uint64_t bytes) {
    BSQ_ENTRY();
    if (!h || !buf) { bsq_set_error("null argument"); return BSQ_ERR; }
    const uint64_t nh = h->holes.size();
    if (bytes < 8 + nh * (sizeof(bsq_hole) + 4)) { bsq_set_error("host-state buffer too small"); return BSQ_ERR; }
    uint8_t* p = static_cast<uint8_t*>(buf);
    memcpy(p, &nh, 8);
    if (nh) { memcpy(p + 8, h->holes.data(), nh * sizeof(bsq_hole)); memcpy(p + 8 + nh * sizeof(bsq_hole), h->hole_ann.data(), nh * 4); }
    return BSQ_OK;
}

// After the device arrays of a replica have been filled (broadcast / peer copy): rebuild the host mirrors the row
// materialisation and the host adapters read -- pac and the annotation vectors from the device copy, the holes from the
// source's host-state blob -- so that every replica answers exactly like the index it was copied from.
int bsq_index_replica_finish(bsq_index* h, const void* host_state, uint64_t bytes) {
    BSQ_ENTRY();
    if (!h || !h->meta.built) { bsq_set_error("replica not allocated"); return BSQ_ERR; }
    CUDA_CHECK(cudaSetDevice(h->device));
    CUDA_CHECK(cudaDeviceSynchronize());
    const bsq_index_meta& m = h->meta;
    h->pac.resize(m.arr_bytes[BSQ_ARR_PAC]); h->ann_offset.resize(m.n_anns); h->ann_len.resize(m.n_anns); h->ann_id.resize(m.n_anns);
    if (!h->pac.empty()) CUDA_CHECK(cudaMemcpy(h->pac.data(), h->d_pac, h->pac.size(), cudaMemcpyDeviceToHost));
    if (m.n_anns) {
        CUDA_CHECK(cudaMemcpy(h->ann_offset.data(), h->d_ann_offset, m.n_anns * 8, cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(h->ann_len.data(), h->d_ann_len, m.n_anns * 4, cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(h->ann_id.data(), h->d_ann_id, m.n_anns * 8, cudaMemcpyDeviceToHost));
    }
    h->replica_pending = false;
    h->holes.clear(); h->hole_ann.clear();
    if (host_state) {
        uint64_t nh = 0;
        if (bytes < 8) { bsq_set_error("host-state blob truncated"); return BSQ_ERR; }
        const uint8_t* p = static_cast<const uint8_t*>(host_state);
        memcpy(&nh, p, 8);
        if (bytes < 8 + nh * (sizeof(bsq_hole) + 4)) { bsq_set_error("host-state blob truncated"); return BSQ_ERR; }
        h->holes.resize(nh); h->hole_ann.resize(nh);
        if (nh) { memcpy(h->holes.data(), p + 8, nh * sizeof(bsq_hole)); memcpy(h->hole_ann.data(), p + 8 + nh * sizeof(bsq_hole), nh * 4); }
        for (uint32_t a : h->hole_ann) if (a >= m.n_anns) { bsq_set_error("host-state blob names a reference row the index does not have"); return BSQ_ERR; }
    }
    return BSQ_OK;
}

int bsq_index_bwt_plain(const bsq_index* h, uint32_t* out) {
    BSQ_ENTRY();
    if (!h || !h->meta.built) { bsq_set_error("index not built"); return BSQ_ERR; }
    std::vector<uint32_t> occ(h->meta.arr_bytes[BSQ_ARR_OCC] / 4);
    if (bsq_index_download(h, BSQ_ARR_OCC, occ.data(), h->meta.arr_bytes[BSQ_ARR_OCC]) != BSQ_OK) return BSQ_ERR;
    uint64_t nw = (h->meta.seq_len + 15) / 16;
    for (uint64_t i = 0; i < nw; ++i) out[i] = occ[(i >> 3 << 4) + 8 + (i & 7)];
    return BSQ_OK;
}

static __global__ void k_sa_sample(const void* sa, int sa_bytes, uint64_t n_sa, uint64_t* out) {
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_sa; k += (uint64_t)gridDim.x * blockDim.x)
        out[k] = sa_bytes == 4 ? (uint64_t)static_cast<const uint32_t*>(sa)[k * 32] : static_cast<const uint64_t*>(sa)[k * 32];
}

int bsq_index_sa_sampled(const bsq_index* h, uint64_t* out, uint64_t n_sa) {
    BSQ_ENTRY();
    if (!h || !h->meta.built) { bsq_set_error("index not built"); return BSQ_ERR; }
    uint64_t n = h->meta.seq_len;
    if (n_sa != (n + 32) / 32) { bsq_set_error("n_sa must be (seq_len + 32) / 32"); return BSQ_ERR; }
    // rows 0, 32, 64, ... of the full SA, gathered on the device (the full array is 50 GB at 3.1 Gbp)
    CUDA_CHECK(cudaSetDevice(h->device));
    uint64_t* d = nullptr;
    CUDA_CHECK(cudaMalloc(&d, n_sa * 8));
    k_sa_sample<<<148 * 8, 256, 0, h->stream>>>(h->d_sa, (int)h->meta.sa_bytes, n_sa, d);
    cudaError_t e = cudaMemcpyAsync(out, d, n_sa * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    if (e != cudaSuccess) { bsq_set_error("bsq_index_sa_sampled: %s", cudaGetErrorString(e)); return BSQ_ERR; }
    out[0] = (uint64_t)-1;   // bwt_cal_sa: sa[0] = -1
    return BSQ_OK;
}

// Builds the arrays the seeding kernel derives from the index on this device (inverse SA, prefix table) now rather than inside
// the first alignment call; *ms = wall time spent (0 when they were already there).  Every replica calls it after the broadcast.
int bsq_index_prepare(bsq_index* h, float* ms) {
    BSQ_ENTRY();
    if (!h || !h->meta.built) { bsq_set_error("index not built"); return BSQ_ERR; }
    CUDA_CHECK(cudaSetDevice(h->device));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, h->stream);
    const DevIndex ix = make_dev_index(h);
    const int rc = ensure_kmer_table(h, ix);
    cudaEventRecord(e1, h->stream);
    cudaEventSynchronize(e1);
    if (ms) cudaEventElapsedTime(ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return rc;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ batch
namespace {

// ids[i] = the (first + i + 1)-th value glibc lrand48() returns from state x0 (mem_align1 draws one per read):
// x_k = A^k x0 + C (A^k - 1) / (A - 1) mod 2^48 by square-and-multiply on the affine map, id = x_k >> 17
__global__ void k_lrand48_ids(uint64_t x0, uint64_t first, uint64_t n, int64_t* ids) {
    const uint64_t M = (1ull << 48) - 1;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t k = first + i + 1, A = 1, C = 0, ba = 0x5DEECE66DULL, bc = 0xBULL;
        while (k) {
            if (k & 1) { A = (A * ba) & M; C = (C * ba + bc) & M; }
            bc = (bc * ba + bc) & M; ba = (ba * ba) & M;
            k >>= 1;
        }
        ids[i] = (int64_t)(((A * x0 + C) & M) >> 17);
    }
}
uint64_t lrand48_advance(uint64_t x, uint64_t k) {
    const uint64_t M = (1ull << 48) - 1;
    uint64_t A = 1, C = 0, ba = 0x5DEECE66DULL, bc = 0xBULL;
    while (k) {
        if (k & 1) { A = (A * ba) & M; C = (C * ba + bc) & M; }
        bc = (bc * ba + bc) & M; ba = (ba * ba) & M;
        k >>= 1;
    }
    return (A * x + C) & M;
}

// NUCLSEQ datum images -> what mem_align1 sees after to_text_palloc + nst_nt4_table (bwa.cpp:146-149): codes 0..3 from the packed
// bases, 4 wherever a hole (an ambiguous letter) covers the base.  Pass 1: lengths; pass 2 (after the scan): one warp per read.
// stats[0] = longest read, stats[1] = index + 1 of a truncated / oversized image (0 = all fine)
__global__ void k_datum_lens(const uint8_t* bytes, const uint64_t* doff, uint64_t base, uint64_t n, uint64_t* lens, uint32_t* stats) {
    uint32_t mx = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t* d = bytes + (doff[i] - base);
        uint32_t hdr[3]; memcpy(hdr, d, 12);
        const uint64_t need = 12 + (uint64_t)hdr[1] * 16 + ((uint64_t)hdr[2] + 3) / 4;
        if ((hdr[0] >> 2) < need || doff[i + 1] < doff[i] + need || hdr[2] > 0x3fffffffu) { atomicMax(stats + 1, (uint32_t)(i + 1)); hdr[2] = 0; }
        lens[i] = hdr[2];
        mx = mx > hdr[2] ? mx : hdr[2];
    }
    mx = __reduce_max_sync(FULL, mx);
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(stats, mx);
    if (blockIdx.x == 0 && threadIdx.x == 0) lens[n] = 0;
}
__global__ void k_datum_unpack(const uint8_t* bytes, const uint64_t* doff, uint64_t base, uint64_t n, const uint64_t* offs, uint8_t* seqs, uint8_t* ascii) {
    const uint64_t gw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint64_t r = gw; r < n; r += nw) {
        const uint8_t* d = bytes + (doff[r] - base);
        uint32_t n_holes, len; memcpy(&n_holes, d + 4, 4); memcpy(&len, d + 8, 4);
        const uint8_t* pac = d + 12 + (size_t)n_holes * 16;
        uint8_t* q = seqs + offs[r]; uint8_t* a = ascii + offs[r];
        for (uint32_t i = lane; i < len; i += 32) { const uint32_t c = (pac[i >> 2] >> ((~i & 3u) << 1)) & 3u; q[i] = (uint8_t)c; a[i] = (uint8_t)"ACGT"[c]; }
        __syncwarp();
        for (uint32_t k = 0; k < n_holes; ++k) {   // inplace_to_text (sequence.cpp:71-81): holes in order, later ones overwrite
            int64_t ho; int32_t hl; memcpy(&ho, d + 12 + (size_t)k * 16, 8); memcpy(&hl, d + 12 + (size_t)k * 16 + 8, 4);
            const uint8_t amb = d[12 + (size_t)k * 16 + 12];
            for (int64_t i = ho + lane; i < ho + hl && i < (int64_t)len; i += 32) if (i >= 0) { q[i] = 4; a[i] = amb; }
            __syncwarp();
        }
    }
}

__global__ void k_rebase_offs(uint64_t* offs, uint64_t n, uint64_t base) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) offs[i] -= base;
}

__global__ void k_to_nt4(uint8_t* s, uint64_t n) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint8_t c = s[i], v;
        // mem_align1_core: seq[i] < 4 ? seq[i] : nst_nt4_table[seq[i]]
        if (c < 4) v = c;
        else switch (c) {
            case 'A': case 'a': v = 0; break; case 'C': case 'c': v = 1; break; case 'G': case 'g': v = 2; break; case 'T': case 't': v = 3; break;
            case '-': v = 5; break; default: v = 4;
        }
        s[i] = v;
    }
}

// mem_approx_mapq_se on the device.  The only transcendental is log() of small integers
// (alignment length l and sub_n + 1): both come from a table filled by the HOST libm, so every double
// operation here is an IEEE +,-,*,/ evaluated in the reference's order (the library is compiled with
// --fmad=false).  Rows whose l or sub_n fall outside the table get mapq = -1 and are finished on the host.
constexpr int LOGTAB_N = 65536;   // alignments of up to 65535 bases get their MAPQ on the device
struct MapqParams { int a, b, min_seed_len; float coef_len; double coef_fac; const double* logtab; };

__device__ __forceinline__ int approx_mapq_dev(const MapqParams& M, const RowDev& r) {
    int mapq, l, sub = r.sub ? r.sub : M.min_seed_len * M.a;
    sub = r.csub > sub ? r.csub : sub;
    if (sub >= r.score) return 0;
    l = r.qe - r.qb > r.re - r.rb ? r.qe - r.qb : (int)(r.re - r.rb);
    if (l >= LOGTAB_N || r.sub_n + 1 >= LOGTAB_N || l <= 0) return -1;
    const double identity = 1. - (double)(l * M.a - r.score) / (M.a + M.b) / l;
    if (r.score == 0) mapq = 0;
    else {
        double tmp = l < M.coef_len ? 1. : M.coef_fac / M.logtab[l];
        tmp *= identity * identity;
        mapq = (int)(6.02 * (r.score - sub) / M.a * tmp * tmp + .499);
    }
    if (r.sub_n > 0) mapq -= (int)(4.343 * M.logtab[r.sub_n + 1] + .499);
    if (mapq > 60) mapq = 60;
    if (mapq < 0) mapq = 0;
    mapq = (int)(mapq * (1. - r.frac_rep) + .499);
    return mapq;
}

constexpr int MAPQ_HOST = 255;   // mapq value of a row whose MAPQ the host finishes (l or sub_n outside the device log table)
__global__ void k_compact_rows(const ReadBlock* blocks, const RowDev* rows, const uint32_t* row_cnt, const uint32_t* row_off, uint32_t n_reads,
                               RowPub* out, RowExt* ext, uint32_t out_cap, MapqParams M, uint32_t* host_mapq) {
    // thread per row slot of a read: the 120-byte working record becomes the 64-byte public row (+ the 48-byte extension when asked
    // for).  *host_mapq is raised when a row's MAPQ has to be finished on the host.  Rows beyond out_cap are dropped: the host sees
    // the total from the scan, grows the buffer and runs the batch again.
    // thread per read (a read has one row, rarely a handful)
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += gridDim.x * blockDim.x) {
        const uint32_t c = row_cnt[r];
        if (!c || row_off[r] + c > out_cap) continue;
        const RowDev* srow = rows + blocks[r].base;
        for (uint32_t k = 0; k < c; ++k) {
            const RowDev a = srow[k];
            const int mq = a.secondary < 0 ? approx_mapq_dev(M, a) : 0;
            if (mq < 0) atomicExch(host_mapq, 1u);
            RowPub d;
            d.rb = a.rb; d.re = a.re; d.pos = a.pos; d.ref_id = a.ref_id; d.qb = a.qb; d.qe = a.qe; d.rid = a.rid; d.score = a.score; d.NM = a.NM;
            d.cigar_off = a.cigar_off; d.n_cigar = a.n_cigar; d.flag = (uint16_t)a.flag; d.mapq = (uint8_t)(mq < 0 ? MAPQ_HOST : mq); d.is_rev = (uint8_t)a.is_rev;
            out[row_off[r] + k] = d;
            if (ext) {
                RowExt e;
                e.hash = a.hash; e.truesc = a.truesc; e.sub = a.sub; e.csub = a.csub; e.sub_n = a.sub_n; e.w = a.w; e.seedcov = a.seedcov;
                e.secondary = a.secondary; e.seedlen0 = a.seedlen0; e.n_comp = a.n_comp; e.frac_rep = a.frac_rep;
                ext[row_off[r] + k] = e;
            }
        }
    }
}

DevIndex make_dev_index(const bsq_index* h) {
    DevIndex ix;
    ix.pac = h->d_pac; ix.occ = h->d_occ; ix.sa = h->d_sa; ix.ann_offset = h->d_ann_offset; ix.ann_len = h->d_ann_len;
    ix.l_pac = h->meta.l_pac; ix.seq_len = h->meta.seq_len; ix.primary = h->meta.primary;
    for (int i = 0; i < 5; ++i) ix.L2[i] = h->meta.L2[i];
    ix.n_anns = (int)h->meta.n_anns; ix.sa_bytes = (int)h->meta.sa_bytes;
    return ix;
}

uint32_t rseq_cap_for(const bsq_index* h, uint32_t max_len) { return max_len + 4u * (uint32_t)h->opts.w + 16u; }

// reads longer than this would need mem_flt_chained_seeds' local SW (runs when 5.5 ln L <= 0.05 L)
bool needs_seed_sw(uint32_t len) { return len > 0 && 5.5f * log((double)len) <= 0.05f * (double)len; }

int upload_reads(bsq_index* h, Batch& b, const char* seqs, const uint64_t* offs, const int64_t* ids, uint64_t n, uint64_t id_first = 0) {
    b.resident = false; b.aligned = false; b.res_serial = 0;
    if (n >= 0x7fffffffull) { bsq_set_error("batch too large"); return BSQ_ERR; }
    uint32_t max_len = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (offs[i + 1] < offs[i]) { bsq_set_error("read offsets must be non-decreasing"); return BSQ_ERR; }
        uint64_t l = offs[i + 1] - offs[i];
        if (l > 0x3fffffffull) { bsq_set_error("read too long"); return BSQ_ERR; }
        max_len = std::max<uint32_t>(max_len, (uint32_t)l);
    }
    const uint64_t total = n ? offs[n] - offs[0] : 0;
    b.n = n; b.max_len = max_len; b.total_bases = total;
    if (needs_seed_sw(max_len) && b.read_logtab_n < max_len + 1) {
        // log(l_query) for mem_flt_chained_seeds' thresholds: host libm values
        std::vector<double> tab(max_len + 1);
        tab[0] = 0.;
        for (uint32_t i = 1; i <= max_len; ++i) tab[i] = log((double)i);
        CUDA_CHECK(b.read_logtab.ensure(max_len + 1));
        CUDA_CHECK(cudaMemcpy(b.read_logtab.p, tab.data(), (max_len + 1) * sizeof(double), cudaMemcpyHostToDevice));
        b.read_logtab_n = max_len + 1;
    }
    CUDA_CHECK(b.seqs.ensure(total + 64)); CUDA_CHECK(b.offs.ensure(n + 1)); CUDA_CHECK(b.ids.ensure(n + 1));
    // the caller's buffers go to the device as they are (pinned buffers make these copies truly asynchronous; pageable ones are staged
    // by the runtime); the offsets are rebased to the batch's first read on the device
    if (total) CUDA_CHECK(cudaMemcpyAsync(b.seqs.p, seqs + offs[0], total, cudaMemcpyHostToDevice, b.st));
    if (n) {
        CUDA_CHECK(cudaMemcpyAsync(b.offs.p, offs, (n + 1) * 8, cudaMemcpyHostToDevice, b.st));
        if (ids) CUDA_CHECK(cudaMemcpyAsync(b.ids.p, ids, n * 8, cudaMemcpyHostToDevice, b.st));
        else { k_lrand48_ids<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8), 256, 0, b.st>>>(h->lrand_state, id_first, n, b.ids.p); ++h->timing.launches; }
    } else CUDA_CHECK(cudaMemsetAsync(b.offs.p, 0, 8, b.st));
    if (total) {
        CUDA_CHECK(b.ascii.ensure(total + 64));
        CUDA_CHECK(cudaMemcpyAsync(b.ascii.p, b.seqs.p, total, cudaMemcpyDeviceToDevice, b.st));   // the text stays for the row materialisation
        k_to_nt4<<<(unsigned)std::min<uint64_t>((total + 255) / 256, 148 * 16), 256, 0, b.st>>>(b.seqs.p, total); ++h->timing.launches;
    }
    if (n && offs[0]) { k_rebase_offs<<<(unsigned)std::min<uint64_t>((n + 256) / 256, 148 * 8), 256, 0, b.st>>>(b.offs.p, n + 1, offs[0]); ++h->timing.launches; }
    h->timing.h2d_bytes += total + (n + 1) * 8 + (ids ? n * 8 : 0);
    b.resident = true;
    return BSQ_OK;
}

// The same batch handed over as NUCLSEQ datum images (image i at bytes + doff[i], as PostgreSQL stores the query rows): 2 bits per
// base on the wire instead of 8.  The host reads the headers once (longest read, total bases: the pools are sized from them); the
// images are unpacked on the device.
int upload_datums_begin(bsq_index* h, Batch& b, const uint8_t* bytes, const uint64_t* doff, const int64_t* ids, uint64_t n, uint64_t id_first) {
    b.resident = false; b.aligned = false; b.res_serial = 0;
    if (n >= 0x7fffffffull) { bsq_set_error("batch too large"); return BSQ_ERR; }
    if (n && doff[n] < doff[0]) { bsq_set_error("datum offsets must be non-decreasing"); return BSQ_ERR; }
    const uint64_t nbytes = n ? doff[n] - doff[0] : 0;
    b.n = n; b.max_len = 0; b.total_bases = 0;
    CUDA_CHECK(b.offs.ensure(n + 2)); CUDA_CHECK(b.ids.ensure(n + 1)); CUDA_CHECK(b.ctl.ensure(64));
    CUDA_CHECK(b.datums.ensure(nbytes + 64)); CUDA_CHECK(b.datum_off.ensure(n + 1)); CUDA_CHECK(b.scan_tmp64.ensure(prim::scan_tmp_elems(n + 1) + 16));
    if (!b.ctl_host) CUDA_CHECK(cudaHostAlloc(&b.ctl_host, 16 * 4, cudaHostAllocDefault));
    if (n) {
        // images, their offsets and the ids go down as they are; the headers are read ON THE DEVICE (longest read, total bases, validity:
        // four bytes come back) -- walking a million headers in host memory costs more than the whole copy
        CUDA_CHECK(cudaMemcpyAsync(b.datums.p, bytes + doff[0], nbytes, cudaMemcpyHostToDevice, b.st));
        CUDA_CHECK(cudaMemcpyAsync(b.datum_off.p, doff, (n + 1) * 8, cudaMemcpyHostToDevice, b.st));
        if (ids) CUDA_CHECK(cudaMemcpyAsync(b.ids.p, ids, n * 8, cudaMemcpyHostToDevice, b.st));
        else { k_lrand48_ids<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8), 256, 0, b.st>>>(h->lrand_state, id_first, n, b.ids.p); ++h->timing.launches; }
        CUDA_CHECK(cudaMemsetAsync(b.ctl.p + 61, 0, 8, b.st));
        k_datum_lens<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8), 256, 0, b.st>>>(b.datums.p, b.datum_off.p, doff[0], n, b.offs.p, b.ctl.p + 61); ++h->timing.launches;
        prim::device_scan<uint64_t, prim::OpSum, false>(b.offs.p, b.offs.p, n + 1, b.scan_tmp64.p, prim::OpSum(), b.st, &h->timing.launches);
        CUDA_CHECK(cudaMemcpyAsync(b.ctl_host + 12, b.ctl.p + 61, 8, cudaMemcpyDeviceToHost, b.st));
        CUDA_CHECK(cudaMemcpyAsync(b.ctl_host + 14, b.offs.p + n, 8, cudaMemcpyDeviceToHost, b.st));
    } else CUDA_CHECK(cudaMemsetAsync(b.offs.p, 0, 8, b.st));
    h->timing.h2d_bytes += nbytes + (n + 1) * 8 + (ids ? n * 8 : 0);
    return BSQ_OK;
}
// second half: waits for the header pass, sizes the text buffers, unpacks the images
int upload_datums_end(bsq_index* h, Batch& b, uint64_t doff0) {
    const uint64_t n = b.n;
    uint64_t total = 0; uint32_t max_len = 0;
    if (n) {
        CUDA_CHECK(cudaStreamSynchronize(b.st));
        if (b.ctl_host[13]) { bsq_set_error("datum %u is truncated or longer than a read may be", b.ctl_host[13] - 1); return BSQ_ERR; }
        max_len = b.ctl_host[12]; memcpy(&total, b.ctl_host + 14, 8);
    }
    b.max_len = max_len; b.total_bases = total;
    if (needs_seed_sw(max_len) && b.read_logtab_n < max_len + 1) {
        std::vector<double> tab(max_len + 1);
        tab[0] = 0.;
        for (uint32_t i = 1; i <= max_len; ++i) tab[i] = log((double)i);
        CUDA_CHECK(b.read_logtab.ensure(max_len + 1));
        CUDA_CHECK(cudaMemcpy(b.read_logtab.p, tab.data(), (max_len + 1) * sizeof(double), cudaMemcpyHostToDevice));
        b.read_logtab_n = max_len + 1;
    }
    CUDA_CHECK(b.seqs.ensure(total + 64)); CUDA_CHECK(b.ascii.ensure(total + 64));
    if (n) { k_datum_unpack<<<(unsigned)std::min<uint64_t>((n + 7) / 8, 148 * 16), 256, 0, b.st>>>(b.datums.p, b.datum_off.p, doff0, n, b.offs.p, b.seqs.p, b.ascii.p); ++h->timing.launches; }
    b.resident = true;
    return BSQ_OK;
}
int upload_datums(bsq_index* h, Batch& b, const uint8_t* bytes, const uint64_t* doff, const int64_t* ids, uint64_t n, uint64_t id_first = 0) {
    if (upload_datums_begin(h, b, bytes, doff, ids, n, id_first) != BSQ_OK) return BSQ_ERR;
    return upload_datums_end(h, b, n ? doff[0] : 0);
}

// k-mer table of the LAST-like seeding pass: built once per device index (also after a broadcast replica)
int ensure_kmer_table(bsq_index* h, const DevIndex& ix) {
    if (ix.seq_len >= (1ull << 40)) return BSQ_OK;   // the table packs rows in 40 bits
    if (!h->d_isa && !getenv("BSQ_NO_ISA")) {
        CUDA_CHECK(cudaMalloc(&h->d_isa, (ix.seq_len + 1) * (uint64_t)ix.sa_bytes + 64));
        build_isa(ix, h->d_isa, h->stream, &h->timing.launches);
        CUDA_CHECK(cudaStreamSynchronize(h->stream));
    }
    if (h->d_kmer || ix.seq_len < (1u << 16) || getenv("BSQ_NO_KMER")) return BSQ_OK;
    h->kmer_k = kmer_table_depth(ix.seq_len);
    // One level more than ceil(log4 n) where HBM allows it (23 GB at K = 15): a K-mer of the read's own locus then has a second,
    // chance occurrence three times less often, and every such occurrence costs the seeding kernels real bwt_extend steps
    // (measured at 200 M symbols: 13.6 -> 12.3 ms per 1 M reads).  Only for 32-bit indexes; the 64-bit ones keep their HBM for SA + ISA.
    if (h->kmer_k == 14 && ix.seq_len > (1ull << 27)) {
        size_t fr = 0, tot = 0;
        const size_t want = kmer_table_bytes(15) + kmer_table_bytes(15) / 4;     // entries + the size table
        // 32-bit indexes: when the table is a quarter of what is free.  64-bit ones (SA + ISA already hold 16 bytes per symbol): when 40 GB
        // stay free for the batch pools of two lanes (c3: 77 GB free -> 28.6 GB table; measured 28.5 -> 25.9 ms seeding per 1.25 M reads).
        // If the pools do not fit after all, the pipeline drops a level and goes on (kmer_downgrade).
        if (cudaMemGetInfo(&fr, &tot) == cudaSuccess && (ix.sa_bytes == 4 ? fr > 4 * kmer_table_bytes(15) : fr > want + (40ull << 30))) h->kmer_k = 15;
    }
    if (const char* e = getenv("BSQ_KMER_K")) { const int k = atoi(e); if (k >= 8 && k <= 15) h->kmer_k = k; }
    if (h->kmer_k > h->kmer_k_cap) h->kmer_k = h->kmer_k_cap;
    CUDA_CHECK(cudaMalloc(&h->d_kmer, kmer_table_bytes(h->kmer_k)));
    build_kmer_table(ix, h->d_kmer, h->kmer_k, h->stream, &h->timing.launches);
    if (!getenv("BSQ_NO_KMER_Z") && cudaMalloc(&h->d_kmer_z, kmer_table_bytes(h->kmer_k) / 4) == cudaSuccess) build_kmer_sizes(h->d_kmer, h->d_kmer_z, h->kmer_k, h->stream, &h->timing.launches);
    else { h->d_kmer_z = nullptr; cudaGetLastError(); }
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return BSQ_OK;
}

// Device memory ran out while sizing a batch's pools: give back the deepest level of the prefix table (three quarters of it) and let
// ensure_kmer_table rebuild the shallower one.  cudaFree waits for everything already queued, so no kernel loses the table under it.
bool kmer_downgrade(bsq_index* h) {
    if (!h->d_kmer || h->kmer_k <= 8) return false;
    cudaFree(h->d_kmer); h->d_kmer = nullptr;
    if (h->d_kmer_z) { cudaFree(h->d_kmer_z); h->d_kmer_z = nullptr; }
    h->kmer_k_cap = h->kmer_k - 1;
    cudaGetLastError();
    return true;
}

// the seeding kernel's view of a batch: inputs, per-read interval slots, the index's derived arrays, control words
SeedParams seed_params(const bsq_index* h, const Batch& b, const DevIndex& ix, uint32_t n, uint32_t cap, uint32_t* ticket, unsigned long long* n_extend) {
    SeedParams P;
    P.seqs = b.seqs.p; P.offs = b.offs.p; P.n_reads = n; P.out = b.intv.p; P.out_cnt = b.intv_cnt.p; P.cap = cap;
    P.kmer_tab = reinterpret_cast<const uint4*>(h->d_kmer); P.kmer_k = h->kmer_k; P.kmer_ztab = h->d_kmer_z; P.isa = h->d_isa;
    P.scratch = b.seed_scratch.p; P.list_cap = b.list_cap;
    // shared-memory bytes per warp for the read: the bases (padded to 16) and their 2-bit packed copy
    P.read_cap = ((b.max_len + 16) & ~15u) + (((b.max_len >> 4) + 3) << 2) + 16 & ~15u;
    P.lists_in_smem = seed_lists_fit_smem(b.list_cap, P.read_cap, ix.sa_bytes);
    P.ticket = ticket; P.overflow = b.ctl.p + 4; P.n_extend = n_extend;
    P.todo = b.seed_todo.p; P.todo_cnt = b.ctl.p + 59;    // reads the thread-per-read passes leave for seed_smem
    P.pk = b.seed_pk.p; P.rflag = b.seed_u32.p; P.cnt12 = b.seed_u32.p ? b.seed_u32.p + n : nullptr; P.ext12 = b.seed_u32.p ? b.seed_u32.p + 2 * (size_t)n : nullptr;
    return P;
}

// One attempt of the kernel pipeline on the batch's stream: size the pools, launch every stage, read the control
// words back asynchronously.  Nothing here waits for the device.
static int pipeline_enqueue_once(bsq_index* h, Batch& b);
int pipeline_enqueue(bsq_index* h, Batch& b) {
    h->alloc_failed = false;
    int rc = pipeline_enqueue_once(h, b);
    for (int tries = 0; rc != BSQ_OK && h->alloc_failed && tries < 2 && kmer_downgrade(h); ++tries) {
        h->alloc_failed = false;
        h->timing.notes |= BSQ_NOTE_TABLE_DOWNGRADED;
        rc = pipeline_enqueue_once(h, b);
        if (rc == BSQ_OK) bsq_set_error("");
    }
    return rc;
}
static int pipeline_enqueue_once(bsq_index* h, Batch& b) {
    if (!b.resident) { bsq_set_error("no reads uploaded"); return BSQ_ERR; }
    const uint32_t n = (uint32_t)b.n;
    bsq_timing& T = h->timing;
    b.idle = true; b.out_rows = 0; b.out_cig = 0;
    if (n == 0 || !h->meta.built) { b.aligned = true; return BSQ_OK; }
    const DevIndex ix = make_dev_index(h);
    const DevOpts& o = h->dopts;
    cudaStream_t st = b.st;
    if (ensure_kmer_table(h, ix) != BSQ_OK) return BSQ_ERR;
    const uint32_t max_len = std::max<uint32_t>(b.max_len, 1);
    const uint32_t rseq_cap = b.att_rseq_cap = std::max(rseq_cap_for(h, max_len), b.rseq_cap);
    if (b.intv_cap == 0) b.intv_cap = 48 + max_len / 4;
    b.pool_cap = std::max<uint32_t>(b.pool_cap, (uint32_t)std::min<uint64_t>((uint64_t)n * 24 + 4096, 0x7fffffffull));
    b.cigar_cap = std::max<uint32_t>(b.cigar_cap, (uint32_t)std::min<uint64_t>((uint64_t)n * 8 + 4096, 0x7fffffffull));
    if (!b.ev_ok) { for (auto& e : b.ev) cudaEventCreate(&e); b.ev_ok = true; }
    if (!b.ctl_host) CUDA_CHECK(cudaHostAlloc(&b.ctl_host, 16 * 4, cudaHostAllocDefault));
    cudaEvent_t* ev = b.ev;
    // ---- (re)size pools
    const int seed_warps = seed_resident_warps();
    b.list_cap = std::max<uint32_t>(max_len + 1, b.intv_cap);
    const int ext_warps = extend_resident_warps();
    const size_t ext_per_warp = extend_scratch_per_warp(max_len, rseq_cap);
    uint32_t z_cap = 0;
    const size_t fin_per_warp = finalize_scratch_per_warp(max_len, rseq_cap, &z_cap);
    int fin_warps = finalize_resident_warps();
    {   // bound the traceback scratch to ~8 GB
        size_t budget = (size_t)8 << 30;
        int fit = (int)std::max<size_t>(budget / fin_per_warp, 4 * 148);
        fin_warps = std::min(fin_warps, fit) / 4 * 4;
    }
#define ENS(x) do { cudaError_t e_ = (x); if (e_ == cudaSuccess && h->test_pool_oom > 0) { --h->test_pool_oom; e_ = cudaErrorMemoryAllocation; } \
    if (e_ != cudaSuccess) { if (e_ == cudaErrorMemoryAllocation) h->alloc_failed = true; cudaGetLastError(); bsq_set_error("allocating batch pools: %s", cudaGetErrorString(e_)); return BSQ_ERR; } } while (0)
    ENS(b.intv.ensure((size_t)n * b.intv_cap)); ENS(b.intv_cnt.ensure(n));
    ENS(b.seed_scratch.ensure((size_t)seed_warps * 3 * b.list_cap));
    ENS(b.raw.ensure(b.pool_cap)); ENS(b.seeds.ensure(b.pool_cap)); ENS(b.ctmp.ensure(b.pool_cap)); ENS(b.ord.ensure(b.pool_cap));
    ENS(b.chains.ensure(b.pool_cap)); ENS(b.srt.ensure(b.pool_cap)); ENS(b.regs.ensure(b.pool_cap)); ENS(b.rows.ensure(b.pool_cap));
    ENS(b.reg_cnt.ensure(n)); ENS(b.row_cnt.ensure(n)); ENS(b.row_off.ensure(n + 1)); ENS(b.blocks.ensure(n));
    ENS(b.scan_tmp.ensure(prim::scan_tmp_elems(n + 1) + 16));
    ENS(b.cigar.ensure(b.cigar_cap));
    ENS(b.ext_scratch.ensure((size_t)ext_warps * ext_per_warp)); ENS(b.fin_scratch.ensure((size_t)fin_warps * fin_per_warp));
    int narrow_warps = 0;
    const size_t narrow_bytes = narrow_zbuf_bytes(&narrow_warps);
    ENS(b.narrow_z.ensure(narrow_bytes)); ENS(b.narrow_jobs.ensure((size_t)b.pool_cap * 5 * 24)); ENS(b.wide_jobs.ensure(b.pool_cap));
    ENS(b.ctl.ensure(64)); ENS(b.chain_todo.ensure(n)); ENS(b.fin_todo.ensure(n)); ENS(b.seed_todo.ensure(n + 1));
    if (max_len <= 496) { ENS(b.seed_pk.ensure((size_t)n * seed_thread_words(max_len))); ENS(b.seed_u32.ensure(3 * (size_t)n)); }
    // the thread-per-extension pre-pass packs column scores in 16 bits: every value it stores is <= l_query * (a + 1)
    static const bool no_memo = getenv("BSQ_NO_EXT_MEMO") != nullptr;
    // Small batches take the warp-cooperative kernels for the DP stages: the thread-per-extension / thread-per-region kernels are built for
    // throughput (one thread runs a whole DP, ~0.5 ms of latency whatever the batch size), which is what a one-read call of
    // nuclseq_search_bwa would wait for.  BSQ_SMALL_BATCH_READS moves the threshold (0 = thread kernels always; the parity tests use that).
    static const long small_n = getenv("BSQ_SMALL_BATCH_READS") ? atol(getenv("BSQ_SMALL_BATCH_READS")) : 40000;   // measured crossover ~50 k reads (scripts/latency.py)
    const bool small_batch = (long)n < small_n;
    const bool use_memo = !no_memo && !small_batch && max_len <= 512 && (uint64_t)max_len * (uint64_t)(o.a + 1) < 32000 && (uint64_t)n * EXT_MEMO_CHAINS < (1ull << 31);
    if (use_memo) {
        const size_t jobs2 = (size_t)n * EXT_MEMO_CHAINS * 2;
        ENS(b.ext_memo.ensure(jobs2)); ENS(b.ext_memo_key.ensure(jobs2)); ENS(b.ext_memo_perm.ensure(jobs2));
        ENS(b.ext_memo_hist.ensure(6 * EXT_MEMO_BINS)); ENS(b.ext_todo.ensure(n));
    }
    if (!b.ext_aux_ok && !small_batch) {
        for (auto& s_ : b.ext_aux.st) ENS(cudaStreamCreateWithPriority(&s_, cudaStreamNonBlocking, &b == &h->batch ? h->prio_hi : h->prio_lo));
        for (auto& e : b.ext_aux.ev) ENS(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        b.ext_aux_ok = true;
    }
    ENS(cudaMemsetAsync(b.ctl.p, 0, 64 * 4, st));
    unsigned long long* ctr = h->collect_counters ? reinterpret_cast<unsigned long long*>(b.ctl.p + 8) : nullptr;
    cudaEventRecord(ev[0], st);
    {
        SeedParams P = seed_params(h, b, ix, n, b.intv_cap, b.ctl.p + 0, ctr ? ctr + 0 : nullptr);
        // thread per read first (almost every short read); seed_smem, warp per read, takes what it declined
        if (seed_thread_usable(P, o, b.max_len)) T.launches += launch_seed_thread(P, ix, o, b.max_len, b.ctl.p + 60, st);
        else P.todo = nullptr;
        launch_seed(P, ix, o, st, nullptr); ++T.launches;
    }
    cudaEventRecord(ev[1], st);
    {
        ChainParams P;
        P.seqs = b.seqs.p; P.offs = b.offs.p; P.n_reads = n; P.intv = b.intv.p; P.intv_cnt = b.intv_cnt.p; P.intv_cap = b.intv_cap;
        P.raw = b.raw.p; P.ctmp = b.ctmp.p; P.ord = b.ord.p; P.chains = b.chains.p; P.seeds = b.seeds.p; P.pool_cap = b.pool_cap; P.pool_top = b.ctl.p + 5;
        P.blocks = b.blocks.p; P.ticket = b.ctl.p + 1; P.overflow = b.ctl.p + 4; P.counters = ctr ? ctr + 1 : nullptr;
        P.logtab = needs_seed_sw(b.max_len) ? b.read_logtab.p : nullptr; P.sw_cells = nullptr;
        // short reads: a thread-per-read pass takes every read with a handful of seed occurrences, the warp kernel the rest
        static const bool no_thread_chain = getenv("BSQ_NO_CHAIN_THREAD") != nullptr;
        const bool thread_chain = !no_thread_chain && !P.logtab;
        P.todo = thread_chain ? b.chain_todo.p : nullptr; P.todo_cnt = b.ctl.p + 57;
        if (thread_chain) { launch_chain_thread(P, ix, o, st); ++T.launches; }
        launch_chain(P, ix, o, st); ++T.launches;
    }
    cudaEventRecord(ev[2], st);
    {
        ExtendParams P;
        P.seqs = b.seqs.p; P.offs = b.offs.p; P.n_reads = n; P.blocks = b.blocks.p; P.chains = b.chains.p; P.seeds = b.seeds.p; P.srt = b.srt.p;
        P.regs = b.regs.p; P.reg_cnt = b.reg_cnt.p; P.scratch = b.ext_scratch.p; P.scratch_per_warp = ext_per_warp; P.max_len = max_len; P.rseq_cap = rseq_cap;
        P.ticket = b.ctl.p + 2; P.overflow = b.ctl.p + 4; P.need_rseq = b.ctl.p + 7; P.counters = ctr ? ctr + 3 : nullptr;
        P.memo = use_memo ? b.ext_memo.p : nullptr; P.memo_key = b.ext_memo_key.p; P.memo_perm = b.ext_memo_perm.p; P.memo_hist = b.ext_memo_hist.p;
        P.todo = use_memo ? b.ext_todo.p : nullptr; P.todo_cnt = b.ctl.p + 56;
        static const bool one_stream = getenv("BSQ_EXT_ONE_STREAM") != nullptr;
        launch_extend_memo(P, ix, o, st, &T.launches, one_stream ? nullptr : &b.ext_aux);
        if (use_memo) { launch_extend_finish(P, ix, o, st); ++T.launches; }
        launch_extend(P, ix, o, st); ++T.launches;
    }
    cudaEventRecord(ev[3], st);
    {
        FinalizeParams P;
        P.seqs = b.seqs.p; P.offs = b.offs.p; P.ids = b.ids.p; P.n_reads = n; P.blocks = b.blocks.p; P.regs = b.regs.p; P.reg_cnt = b.reg_cnt.p;
        P.rows = b.rows.p; P.row_cnt = b.row_cnt.p; P.cigar_pool = b.cigar.p; P.cigar_cap = b.cigar_cap; P.cigar_top = b.ctl.p + 6;
        P.scratch = b.fin_scratch.p; P.scratch_per_warp = fin_per_warp; P.max_len = max_len; P.z_cap = z_cap; P.ann_id = h->d_ann_id;
        P.narrow_jobs = small_batch ? nullptr : b.narrow_jobs.p; P.narrow_cap = b.pool_cap; P.narrow_cnt = b.ctl.p + 32; P.wide_jobs = b.wide_jobs.p; P.wide_cnt = b.ctl.p + 25;
        P.narrow_z = b.narrow_z.p; P.narrow_warps = narrow_warps;
        { static const bool no_tight = getenv("BSQ_FIN_NO_TIGHT") != nullptr; P.narrow_tight = !no_tight; }
        P.ticket = b.ctl.p + 40; P.overflow = b.ctl.p + 4; P.need_rseq = b.ctl.p + 7; P.counters = ctr ? ctr + 6 : nullptr;
        static const bool no_thread_fin = getenv("BSQ_NO_FIN_THREAD") != nullptr;
        P.todo = (no_thread_fin || small_batch) ? nullptr : b.fin_todo.p; P.todo_cnt = b.ctl.p + 58;
        launch_finalize(P, ix, o, st, rseq_cap, fin_warps, &T.launches, b.ext_aux_ok ? &b.ext_aux : nullptr);
    }
    // compact rows: exclusive scan of row_cnt (n + 1 entries, the last one is a zero pad) -> row_off
    prim::device_scan<uint32_t, prim::OpSum, false>(b.row_cnt.p, b.row_off.p, n, b.scan_tmp.p, prim::OpSum(), st, &T.launches);
    {   // rows in read order + MAPQ (the tail of mem_reg2aln): part of the step, so inside the timed window
        if (!h->d_logtab) {
            std::vector<double> tab(LOGTAB_N);
            tab[0] = 0.;
            for (int i = 1; i < LOGTAB_N; ++i) tab[i] = log((double)i);   // host libm: the values the reference's log() returns
            ENS(cudaMalloc(&h->d_logtab, LOGTAB_N * sizeof(double)));
            ENS(cudaMemcpy(h->d_logtab, tab.data(), LOGTAB_N * sizeof(double), cudaMemcpyHostToDevice));
        }
        b.rows_cap = std::max<uint64_t>(b.rows_cap, (uint64_t)n * 2 + 4096);
        ENS(b.rows_compact.ensure(b.rows_cap));
        // the extension records are always produced on the device: a row the host must finish (MAPQ outside the log table) needs them
        ENS(b.rows_ext.ensure(b.rows_cap));
        MapqParams M; M.a = h->opts.a; M.b = h->opts.b; M.min_seed_len = h->opts.min_seed_len; M.coef_len = h->mapQ_coef_len;
        M.coef_fac = (double)h->mapQ_coef_fac; M.logtab = h->d_logtab;
        b.cig_rebased = 0;
        k_compact_rows<<<(unsigned)std::min<uint32_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(b.blocks.p, b.rows.p, b.row_cnt.p, b.row_off.p, n, b.rows_compact.p, b.rows_ext.p,
                                                (uint32_t)std::min<uint64_t>(b.rows_cap, 0xffffffffull), M, b.ctl.p + 30); ++T.launches;
    }
    cudaEventRecord(ev[4], st);
    ENS(cudaMemcpyAsync(b.ctl_host, b.ctl.p, 8 * 4, cudaMemcpyDeviceToHost, st));
    ENS(cudaMemcpyAsync(b.ctl_host + 8, b.row_off.p + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    ENS(cudaMemcpyAsync(b.ctl_host + 9, b.row_cnt.p + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    ENS(cudaMemcpyAsync(b.ctl_host + 10, b.ctl.p + 30, 4, cudaMemcpyDeviceToHost, st));
    b.idle = false;
    return BSQ_OK;
#undef ENS
}

// Waits for the attempt in flight and looks at its control words: *again = a pool or a scratch area was too small;
// the capacities have been raised and pipeline_enqueue must run once more.
int pipeline_check(bsq_index* h, Batch& b, bool* again) {
    *again = false;
    if (b.idle) return BSQ_OK;
    bsq_timing& T = h->timing;
    CUDA_CHECK(cudaStreamSynchronize(b.st));
    b.idle = true;
    { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { bsq_set_error("kernel failure: %s", cudaGetErrorString(e_)); return BSQ_ERR; } }
    const uint32_t* ctl = b.ctl_host;
    cudaEvent_t* ev = b.ev;
    float ms;
    cudaEventElapsedTime(&ms, ev[0], ev[1]); T.seed += ms;
    cudaEventElapsedTime(&ms, ev[1], ev[2]); T.chain += ms;
    cudaEventElapsedTime(&ms, ev[2], ev[3]); T.extend += ms;
    cudaEventElapsedTime(&ms, ev[3], ev[4]); T.finalize += ms;
    cudaEventElapsedTime(&ms, ev[0], ev[4]); T.total += ms;
    if (ctl[7] > b.att_rseq_cap && ctl[4] != 2) {   // a reference window was larger than planned: grow the per-warp scratch and run again
        b.rseq_cap = ctl[7] + ctl[7] / 8 + 64;
        *again = true;
        return BSQ_OK;
    }
    if (ctl[4] == 0) {
        if (h->collect_counters) {
            uint64_t c8[8];
            CUDA_CHECK(cudaMemcpy(c8, b.ctl.p + 8, sizeof(c8), cudaMemcpyDeviceToHost));
            for (int i = 0; i < 8; ++i) h->counters[i] += c8[i];
        }
        b.out_rows = (uint64_t)ctl[8] + ctl[9]; b.out_cig = ctl[6];
        if (b.out_rows > b.rows_cap) {   // more rows than the compacted buffer holds: grow it and run the batch again
            b.rows_cap = b.out_rows + b.out_rows / 8 + 4096;
            *again = true;
            return BSQ_OK;
        }
        b.aligned = true;
        return BSQ_OK;
    }
    if (ctl[4] == 2) { bsq_set_error("alignment scratch capacity exceeded (reference window / traceback larger than planned)"); return BSQ_ERR; }
    // a pool was too small: grow everything that can overflow and run the batch again
    if (ctl[5] > b.pool_cap) b.pool_cap = (uint32_t)std::min<uint64_t>((uint64_t)ctl[5] + ctl[5] / 8 + 4096, 0x7fffffffull);
    else if (ctl[6] > b.cigar_cap) b.cigar_cap = (uint32_t)std::min<uint64_t>((uint64_t)ctl[6] + ctl[6] / 8 + 4096, 0x7fffffffull);
    else b.intv_cap *= 2;
    *again = true;
    return BSQ_OK;
}

// finishes the batch: checks the attempt in flight (if any) and re-runs it while capacities have to grow
int pipeline_finish(bsq_index* h, Batch& b) {
    for (int attempt = 0; attempt < 12; ++attempt) {
        bool again = false;
        if (pipeline_check(h, b, &again) != BSQ_OK) return BSQ_ERR;
        if (!again) { if (!b.aligned) break; return BSQ_OK; }
        if (pipeline_enqueue(h, b) != BSQ_OK) return BSQ_ERR;
    }
    bsq_set_error("batch pools kept overflowing");
    return BSQ_ERR;
}

int run_pipeline(bsq_index* h, Batch& b) {
    bsq_timing& T = h->timing;
    T.seed = T.chain = T.extend = T.finalize = T.total = 0;
    memset(h->counters, 0, sizeof(h->counters));
    if (pipeline_enqueue(h, b) != BSQ_OK) return BSQ_ERR;
    return pipeline_finish(h, b);
}

// mem_approx_mapq_se on the host: double math with libm log
int approx_mapq(const bsq_index* h, const bsq_row& a, const bsq_row_ext& x) {
    const bsq_opts& o = h->opts;
    int mapq, l, sub = x.sub ? x.sub : o.min_seed_len * o.a;
    double identity;
    sub = x.csub > sub ? x.csub : sub;
    if (sub >= a.score) return 0;
    l = a.qe - a.qb > a.re - a.rb ? a.qe - a.qb : (int)(a.re - a.rb);
    identity = 1. - (double)(l * o.a - a.score) / (o.a + o.b) / l;
    if (a.score == 0) mapq = 0;
    else {   // mapQ_coef_len = 50 > 0
        double tmp;
        tmp = l < h->mapQ_coef_len ? 1. : h->mapQ_coef_fac / log(l);
        tmp *= identity * identity;
        mapq = (int)(6.02 * (a.score - sub) / o.a * tmp * tmp + .499);
    }
    if (x.sub_n > 0) mapq -= (int)(4.343 * log(x.sub_n + 1) + .499);
    if (mapq > 60) mapq = 60;
    if (mapq < 0) mapq = 0;
    mapq = (int)(mapq * (1. - x.frac_rep) + .499);
    return mapq;
}

// A result lives in ONE pinned host block, laid out by capacity: row_off (u64 x (n+1)) | rows (row_cap + 1) |
// cigar (cig_cap + 1) | the device's u32 row offsets (n + 1).  Freed blocks are cached process-wide so that
// steady-state calls do not pay cudaHostAlloc.
struct ResultImpl { bsq_result pub; void* block; size_t bytes; uint64_t row_cap, cig_cap; uint32_t* o32; bsq_row_ext* ext; uint64_t serial; };
// results handed out and not yet freed: bsq_result_tuples recognises its own results (a caller may also pass a bsq_result it assembled)
std::mutex g_live_mu; std::vector<const ResultImpl*> g_live; uint64_t g_result_serial = 0;
struct PinnedCache { void* ptr[4]; size_t bytes[4]; };
PinnedCache g_pinned = {{nullptr, nullptr, nullptr, nullptr}, {0, 0, 0, 0}};
std::mutex g_pinned_mu;

void* pinned_get(size_t need, size_t* got) {
    {
        std::lock_guard<std::mutex> lk(g_pinned_mu);
        for (int i = 0; i < 4; ++i)
            if (g_pinned.ptr[i] && g_pinned.bytes[i] >= need && g_pinned.bytes[i] <= 2 * need + (1 << 20)) {
                void* p = g_pinned.ptr[i]; *got = g_pinned.bytes[i]; g_pinned.ptr[i] = nullptr; g_pinned.bytes[i] = 0; return p;
            }
    }
    void* p = nullptr;
    size_t want = need + need / 8 + 4096;
    if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    *got = want;
    return p;
}
void pinned_put(void* p, size_t bytes) {
    {
        std::lock_guard<std::mutex> lk(g_pinned_mu);
        for (int i = 0; i < 4; ++i) if (!g_pinned.ptr[i]) { g_pinned.ptr[i] = p; g_pinned.bytes[i] = bytes; return; }
    }
    cudaFreeHost(p);
}

// layout by capacity: row_off (u64 x (n+1)) | rows (row_cap + 1) | cigar (cig_cap + 1) | the device's u32 row offsets (n + 1) |
// extension rows (row_cap + 1; only with BSQ_FLAG_ROWS_EXT or when a row's MAPQ was finished on the host)
ResultImpl* result_new(uint64_t n, uint64_t row_cap, uint64_t cig_cap, bool with_ext) {
    const size_t off_rows = ((n + 1) * 8 + 63) & ~(size_t)63;
    const size_t off_cig = (off_rows + (row_cap + 1) * sizeof(bsq_row) + 63) & ~(size_t)63;
    const size_t off_o32 = (off_cig + (cig_cap + 1) * 4 + 63) & ~(size_t)63;
    const size_t off_ext = (off_o32 + (n + 1) * 4 + 63) & ~(size_t)63;
    const size_t need = off_ext + (with_ext ? (row_cap + 1) * sizeof(bsq_row_ext) : 0);
    size_t got = 0;
    void* block = pinned_get(need, &got);
    if (!block) { bsq_set_error("cannot allocate %zu bytes of pinned host memory for the result", need); return nullptr; }
    ResultImpl* R = new ResultImpl;
    R->block = block; R->bytes = got; R->row_cap = row_cap; R->cig_cap = cig_cap;
    R->pub.n_reads = n;
    R->pub.row_off = reinterpret_cast<uint64_t*>(block);
    R->pub.rows = reinterpret_cast<bsq_row*>(static_cast<char*>(block) + off_rows);
    R->pub.cigar = reinterpret_cast<uint32_t*>(static_cast<char*>(block) + off_cig);
    R->pub.n_cigar_words = 0;
    R->o32 = reinterpret_cast<uint32_t*>(static_cast<char*>(block) + off_o32);
    R->ext = with_ext ? reinterpret_cast<bsq_row_ext*>(static_cast<char*>(block) + off_ext) : nullptr;
    R->pub.rows_ext = R->ext;
    { std::lock_guard<std::mutex> lk(g_live_mu); R->serial = ++g_result_serial; g_live.push_back(R); }
    return R;
}

void result_delete(ResultImpl* R) {
    if (!R) return;
    { std::lock_guard<std::mutex> lk(g_live_mu); for (size_t i = 0; i < g_live.size(); ++i) if (g_live[i] == R) { g_live[i] = g_live.back(); g_live.pop_back(); break; } }
    pinned_put(R->block, R->bytes); delete R;
}

// Compacts the aligned batch's rows on the device (+ MAPQ) and starts the copies into the result: the batch's reads are
// reads [read_base, read_base + b.n) of the result, its rows go to row_base, its CIGAR words to cig_base.
// A chunk's place in the whole result, applied on the device ahead of the copy: 64-bit row offsets from the chunk's 32-bit ones, CIGAR
// offsets moved behind the earlier chunks' words (a host loop over a million 64-byte records costs more than the copy itself).
static __global__ void k_result_rebase(const uint32_t* __restrict__ row_off, uint64_t n, uint64_t row_base, uint64_t* __restrict__ row_off64,
                                       RowPub* __restrict__ rows, uint64_t n_rows, uint32_t cig_delta) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, t0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t i = t0; i < n; i += stride) row_off64[i] = row_base + row_off[i];
    if (cig_delta) for (uint64_t i = t0; i < n_rows; i += stride) if (rows[i].n_cigar) rows[i].cigar_off += cig_delta;
}

int download_enqueue(bsq_index* h, Batch& b, ResultImpl* R, uint64_t read_base, uint64_t row_base, uint64_t cig_base) {
    const uint64_t n = b.n;
    if (n == 0 || !h->meta.built || !b.aligned) return BSQ_OK;
    const uint64_t total_rows = b.out_rows; const uint32_t cig_top = b.out_cig;
    if (row_base + total_rows > R->row_cap || cig_base + cig_top > R->cig_cap || cig_base + cig_top > 0xffffffffull) { bsq_set_error("result block too small"); return BSQ_ERR; }
    cudaStream_t st = b.st;
    b.res_serial = R->serial; b.res_read_base = read_base; b.res_row_base = row_base; b.res_cig_base = cig_base;
    CUDA_CHECK(b.row_off64.ensure(n + 1));
    k_result_rebase<<<(unsigned)std::min<uint64_t>((std::max<uint64_t>(n, total_rows) + 255) / 256, 148 * 8), 256, 0, st>>>(
        b.row_off.p, n, row_base, b.row_off64.p, b.rows_compact.p, total_rows, (uint32_t)(cig_base - b.cig_rebased));
    ++h->timing.launches;
    b.cig_rebased = cig_base;
    CUDA_CHECK(cudaMemcpyAsync(R->pub.row_off + read_base, b.row_off64.p, n * 8, cudaMemcpyDeviceToHost, st));
    if (total_rows) CUDA_CHECK(cudaMemcpyAsync(R->pub.rows + row_base, b.rows_compact.p, total_rows * sizeof(bsq_row), cudaMemcpyDeviceToHost, st));
    if (total_rows && R->ext) CUDA_CHECK(cudaMemcpyAsync(R->ext + row_base, b.rows_ext.p, total_rows * sizeof(bsq_row_ext), cudaMemcpyDeviceToHost, st));
    else if (total_rows && b.ctl_host[10]) {   // rare: a MAPQ to finish on the host although the caller did not ask for the extension records
        b.ext_tmp.resize(total_rows);
        CUDA_CHECK(cudaMemcpyAsync(b.ext_tmp.data(), b.rows_ext.p, total_rows * sizeof(bsq_row_ext), cudaMemcpyDeviceToHost, st));
    }
    if (cig_top) CUDA_CHECK(cudaMemcpyAsync(R->pub.cigar + cig_base, b.cigar.p, (size_t)cig_top * 4, cudaMemcpyDeviceToHost, st));
    h->timing.d2h_bytes += n * 8 + 12 + total_rows * (sizeof(bsq_row) + (R->ext ? sizeof(bsq_row_ext) : 0)) + (uint64_t)cig_top * 4;
    return BSQ_OK;
}

// after the copies have landed: 64-bit row offsets, and MAPQ of the rows outside the device log table (very long
// alignments).  n reads starting at read_base, n_rows rows starting at row_base.
void download_finish(bsq_index* h, ResultImpl* R, uint64_t n, uint64_t read_base, uint64_t n_rows, uint64_t row_base, uint64_t cig_base, bool host_mapq,
                     const bsq_row_ext* ext_tmp) {
    if (!h->meta.built) { for (uint64_t i = 0; i < n; ++i) R->pub.row_off[read_base + i] = row_base; return; }
    (void)cig_base;     // row offsets and CIGAR offsets were put in place on the device (k_result_rebase)
    if (n_rows && host_mapq) {
        // alignments longer than the device's log table: mem_approx_mapq_se with libm on the host, from the extension records
        const bsq_row_ext* ext = R->ext ? R->ext + row_base : ext_tmp;
        for (uint64_t i = 0; i < n_rows && ext; ++i)
            if (R->pub.rows[row_base + i].mapq == MAPQ_HOST) R->pub.rows[row_base + i].mapq = (uint8_t)approx_mapq(h, R->pub.rows[row_base + i], ext[i]);
    }
}

int download_result(bsq_index* h, Batch& b, bsq_result** out) {
    if (!b.aligned) { bsq_set_error("no aligned batch to download"); return BSQ_ERR; }
    h->timing.d2h_bytes = 0;
    ResultImpl* R = result_new(b.n, b.out_rows, b.out_cig, (h->flags & BSQ_FLAG_ROWS_EXT) != 0);
    if (!R) return BSQ_ERR;
    if (download_enqueue(h, b, R, 0, 0, 0) != BSQ_OK || cudaStreamSynchronize(b.st) != cudaSuccess) { result_delete(R); if (!*bsq_last_error()) bsq_set_error("download failed"); return BSQ_ERR; }
    download_finish(h, R, b.n, 0, b.out_rows, 0, 0, b.ctl_host && b.ctl_host[10], b.ext_tmp.data());
    R->pub.row_off[b.n] = b.out_rows;
    R->pub.n_cigar_words = b.out_cig;
    *out = &R->pub;
    return BSQ_OK;
}

// bsq_align_batch on a large batch: the reads are cut into chunks that alternate between two lanes (batch + stream
// each), so that a chunk's host->device copy and its result's device->host copy run while the other lane computes.
// Chunk c's rows follow chunk c-1's in the result, so downloads are issued in chunk order; a lane takes chunk c+2 as
// soon as chunk c's download has been queued (stream order protects the device buffers), and the host-side completion
// of a download is deferred until the next chunk's work is in the queue.
// chunk size: half of the batch, within [64 K, 1 M] reads.  Measured on B200, 1 M x 150 bp with the slim wire format (64 MB in, 77 MB
// out): one pass 35.2 ms, 2 chunks 31.8, 4 chunks 35.9 (a chunk pays ~40 kernel tails and runs on half-empty persistent grids, so few
// large chunks win once the copies are small).  BSQ_CHUNK_READS overrides it.
constexpr uint64_t CHUNK_MIN = 1u << 16, CHUNK_MAX = 1u << 20;
uint64_t chunk_reads(uint64_t n, bool two = false) {
    if (two) return std::max((n + 1) / 2, CHUNK_MIN);
    static long long env = -1;
    if (env < 0) { env = 0; if (const char* e = getenv("BSQ_CHUNK_READS")) { const long long x = atoll(e); if (x >= 1024) env = x; } }
    if (env > 0) return (uint64_t)env;
    const uint64_t k = std::max<uint64_t>(2, (n + CHUNK_MAX - 1) / CHUNK_MAX);     // as few chunks as the cap allows, at least two
    return std::max((n + k - 1) / k, CHUNK_MIN);
}

// where a batch's reads come from: ASCII + offsets (bsq_align_batch) or NUCLSEQ datum images (bsq_align_batch_datums)
struct ReadSrc { const char* seqs; const uint64_t* offs; const uint8_t* dbytes; const uint64_t* doff; const int64_t* ids; };
int upload_src(bsq_index* h, Batch& b, const ReadSrc& S, uint64_t s0, uint64_t cnt) {
    const int64_t* ids = S.ids ? S.ids + s0 : nullptr;
    return S.dbytes ? upload_datums(h, b, S.dbytes, S.doff + s0, ids, cnt, s0) : upload_reads(h, b, S.seqs, S.offs + s0, ids, cnt, s0);
}

int align_chunked(bsq_index* h, const ReadSrc& S, uint64_t n, bsq_result** out, bool* fell_back) {
    *fell_back = false;
    if (!h->stream2) {
        CUDA_CHECK(cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, h->prio_lo));
        h->batch2.st = h->stream2;
    }
    static const bool trace = getenv("BSQ_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto mark = [&](const char* what, uint64_t c) {
        if (trace) fprintf(stderr, "[bsq trace] %8.3f ms  %s %llu\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(), what, (unsigned long long)c);
    };
    Batch* lane[2] = {&h->batch, &h->batch2};
    for (Batch* b : lane) if (!b->evx_ok) { for (auto& e : b->ev_x) cudaEventCreate(&e); b->evx_ok = true; }
    // chunk boundaries: uniform chunks, or (BSQ_CHUNK_FRACS="f0,f1,...") fractions of the batch -- the first chunk's upload and the
    // last chunk's download are the only copies nothing hides, while every chunk pays the fixed cost of ~40 kernel tails
    std::vector<uint64_t> starts;
    {
        static const char* fr = getenv("BSQ_CHUNK_FRACS");
        if (fr) {
            double acc = 0; starts.push_back(0);
            for (const char* q = fr; *q;) { char* e; const double f = strtod(q, &e); if (e == q) break; acc += f; starts.push_back(std::min<uint64_t>(n, (uint64_t)(acc * (double)n))); q = *e ? e + 1 : e; }
            if (starts.back() != n) starts.push_back(n);
        } else {
            const uint64_t CHUNK_READS = chunk_reads(n, (h->flags & BSQ_FLAG_TWO_CHUNKS) != 0);
            for (uint64_t s0 = 0; s0 < n; s0 += CHUNK_READS) starts.push_back(s0);
            starts.push_back(n);
        }
    }
    const uint64_t n_chunks = starts.size() - 1;
    auto start_of = [&](uint64_t c) { return starts[std::min<uint64_t>(c, n_chunks)]; };
    auto launch = [&](uint64_t c) -> int {
        Batch& b = *lane[c & 1];
        const uint64_t s0 = start_of(c), cnt = start_of(c + 1) - s0;
        b.intv_cap = std::max(b.intv_cap, lane[(c & 1) ^ 1]->intv_cap);      // capacities learnt by one lane serve the other
        cudaEventRecord(b.ev_x[0], b.st);
        mark("upload begin", c);
        if (upload_src(h, b, S, s0, cnt) != BSQ_OK) return BSQ_ERR;
        mark("upload done (headers back)", c);
        cudaEventRecord(b.ev_x[1], b.st);
        // BSQ_CHUNK_SERIAL: chunk c's kernels start behind chunk c-1's (its copies still overlap them), so that an earlier chunk is
        // finished -- and its download under way -- while the later one computes, instead of both sharing the SMs to the end
        static const bool serial = getenv("BSQ_CHUNK_SERIAL") != nullptr;
        if (serial && c > 0) cudaStreamWaitEvent(b.st, lane[(c & 1) ^ 1]->ev_x[4], 0);
        const int rc2 = pipeline_enqueue(h, b);
        cudaEventRecord(b.ev_x[4], b.st);
        mark("pipeline queued", c);
        return rc2;
    };
    struct Pending { bool on = false; uint64_t c = 0, n = 0, n_rows = 0, row_base = 0, cig_base = 0; bool host_mapq = false; } pend;
    ResultImpl* R = nullptr;
    auto complete = [&](const Pending& q) -> int {       // host side of chunk q.c's download
        Batch& b = *lane[q.c & 1];
        if (cudaEventSynchronize(b.ev_x[3]) != cudaSuccess) { bsq_set_error("result download failed: %s", cudaGetErrorString(cudaGetLastError())); return BSQ_ERR; }
        float ms;
        if (cudaEventElapsedTime(&ms, b.ev_x[2], b.ev_x[3]) == cudaSuccess) h->timing.d2h += ms;
        mark("download arrived", q.c);
        download_finish(h, R, q.n, start_of(q.c), q.n_rows, q.row_base, q.cig_base, q.host_mapq, b.ext_tmp.data());
        mark("download finished on the host", q.c);
        return BSQ_OK;
    };
    uint64_t row_base = 0, cig_base = 0;
    int rc = launch(0);
    if (rc == BSQ_OK && n_chunks > 1) rc = launch(1);
    for (uint64_t c = 0; c < n_chunks && rc == BSQ_OK; ++c) {
        Batch& b = *lane[c & 1];
        if ((rc = pipeline_finish(h, b)) != BSQ_OK) break;
        mark("pipeline finished", c);
        { float ms; if (cudaEventElapsedTime(&ms, b.ev_x[0], b.ev_x[1]) == cudaSuccess) h->timing.h2d += ms; }
        if (!R) {
            // capacity of the result from the first chunk's yield
            // (BSQ_TEST_TIGHT_RESULT makes the estimate too small on purpose so that tests reach the one-plain-pass fallback)
            const double scale = (double)n / (double)b.n * (getenv("BSQ_TEST_TIGHT_RESULT") ? 0.4 : 1.2);
            R = result_new(n, (uint64_t)(b.out_rows * scale) + (getenv("BSQ_TEST_TIGHT_RESULT") ? 16 : 65536), (uint64_t)(b.out_cig * scale) + (getenv("BSQ_TEST_TIGHT_RESULT") ? 16 : 65536), (h->flags & BSQ_FLAG_ROWS_EXT) != 0);
            if (!R) { rc = BSQ_ERR; break; }
        }
        if (row_base + b.out_rows > R->row_cap || cig_base + b.out_cig > R->cig_cap) { *fell_back = true; break; }
        cudaEventRecord(b.ev_x[2], b.st);
        if ((rc = download_enqueue(h, b, R, start_of(c), row_base, cig_base)) != BSQ_OK) break;
        cudaEventRecord(b.ev_x[3], b.st);
        mark("download queued", c);
        Pending mine; mine.on = true; mine.c = c; mine.n = b.n; mine.n_rows = b.out_rows; mine.row_base = row_base; mine.cig_base = cig_base;
        mine.host_mapq = b.ctl_host[10] != 0;   // read now: the lane's control words are reused by the next chunk it takes
        row_base += b.out_rows; cig_base += b.out_cig;
        if (pend.on) { if ((rc = complete(pend)) != BSQ_OK) break; pend.on = false; }
        // chunk c-1's lane is free on the host side now (its download completed): it already runs chunk c+1.
        // This lane takes chunk c+2 behind its download.
        if (c + 2 < n_chunks) { if ((rc = launch(c + 2)) != BSQ_OK) break; }
        pend = mine;
    }
    if (rc == BSQ_OK && !*fell_back && pend.on) rc = complete(pend);
    if (rc != BSQ_OK || *fell_back) {
        cudaStreamSynchronize(h->stream); cudaStreamSynchronize(h->stream2);
        cudaGetLastError();
        h->batch.idle = h->batch2.idle = true; h->batch.aligned = h->batch2.aligned = false;
        result_delete(R);
        return rc;
    }
    R->pub.row_off[n] = row_base;
    R->pub.n_cigar_words = cig_base;
    *out = &R->pub;
    return BSQ_OK;
}

}  // namespace

extern "C" {

int bsq_reads_upload(bsq_index* h, const char* seqs, const uint64_t* offs, const int64_t* ids, uint64_t n) {
    BSQ_ENTRY();
    if (!h || (n && (!seqs || !offs))) { bsq_set_error("null argument"); return BSQ_ERR; }
    CUDA_CHECK(cudaSetDevice(h->device));
    h->timing.launches = 0;
    h->timing.h2d_bytes = 0;
    if (upload_reads(h, h->batch, seqs, offs, ids, n) != BSQ_OK) return BSQ_ERR;
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    return BSQ_OK;
}

int bsq_align_resident(bsq_index* h) {
    BSQ_ENTRY();
    if (!h) { bsq_set_error("null index"); return BSQ_ERR; }
    CUDA_CHECK(cudaSetDevice(h->device));
    h->timing.launches = 0;
    return run_pipeline(h, h->batch);
}

int bsq_result_download(bsq_index* h, bsq_result** out) {
    BSQ_ENTRY();
    if (!h || !out) { bsq_set_error("null argument"); return BSQ_ERR; }
    CUDA_CHECK(cudaSetDevice(h->device));
    return download_result(h, h->batch, out);
}

static int align_batch_src(bsq_index* h, const ReadSrc& S, uint64_t n, bsq_result** out) {
    CUDA_CHECK(cudaSetDevice(h->device));
    cudaEvent_t e0, e1, e2, e3;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
    bsq_timing& T = h->timing;
    T.launches = 0; T.h2d_bytes = T.d2h_bytes = 0; T.h2d = T.d2h = 0; T.notes = 0;
    int rc = BSQ_OK;
    bool chunked = n >= 2 * CHUNK_MIN && n > chunk_reads(n, (h->flags & BSQ_FLAG_TWO_CHUNKS) != 0) && h->meta.built && !getenv("BSQ_NO_CHUNKS");
    cudaEventRecord(e0, h->stream);
    if (chunked) {
        // two lanes, copies overlapped with compute; the stage times are sums over chunks of each lane's stream time
        T.seed = T.chain = T.extend = T.finalize = T.total = 0;
        memset(h->counters, 0, sizeof(h->counters));
        bool fell_back = false;
        rc = align_chunked(h, S, n, out, &fell_back);
        if (fell_back) {   // the result outgrew the estimate: one plain pass; the call succeeds and says so in bsq_timing.notes
            T.notes |= BSQ_NOTE_CHUNK_FALLBACK;
            chunked = false; T.h2d_bytes = T.d2h_bytes = 0; T.h2d = T.d2h = 0;
        }
        else { cudaEventRecord(e3, h->stream); cudaEventSynchronize(e3); cudaEventElapsedTime(&T.total, e0, e3); }
    }
    if (!chunked) {
        rc = upload_src(h, h->batch, S, 0, n);
        cudaEventRecord(e1, h->stream);
        if (rc == BSQ_OK) rc = run_pipeline(h, h->batch);
        cudaEventRecord(e2, h->stream);
        if (rc == BSQ_OK) rc = download_result(h, h->batch, out);
        cudaEventRecord(e3, h->stream);
        cudaEventSynchronize(e3);
        cudaEventElapsedTime(&T.h2d, e0, e1); cudaEventElapsedTime(&T.d2h, e2, e3); cudaEventElapsedTime(&T.total, e0, e3);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaEventDestroy(e3);
    if (rc == BSQ_OK && !S.ids) h->lrand_state = lrand48_advance(h->lrand_state, n);   // the session's stream moved on by one draw per read
    return rc;
}

int bsq_align_batch(bsq_index* h, const char* seqs, const uint64_t* offs, const int64_t* ids, uint64_t n, bsq_result** out) {
    BSQ_ENTRY();
    if (!h || !out || (n && (!seqs || !offs))) { bsq_set_error("null argument"); return BSQ_ERR; }
    const ReadSrc S{seqs, offs, nullptr, nullptr, ids};
    return align_batch_src(h, S, n, out);
}

int bsq_align_batch_datums(bsq_index* h, const uint8_t* bytes, const uint64_t* off, const int64_t* ids, uint64_t n, bsq_result** out) {
    BSQ_ENTRY();
    if (!h || !out || (n && (!bytes || !off))) { bsq_set_error("null argument"); return BSQ_ERR; }
    const ReadSrc S{nullptr, nullptr, bytes, off, ids};
    return align_batch_src(h, S, n, out);
}

int bsq_session_lrand48(bsq_index* h, int set, uint64_t* state) {
    BSQ_ENTRY();
    if (!h || !state) { bsq_set_error("null argument"); return BSQ_ERR; }
    if (set) h->lrand_state = *state & 0xFFFFFFFFFFFFULL; else *state = h->lrand_state;
    return BSQ_OK;
}

// ---- row materialisation

// row -> read index of a batch from its exclusive row offsets (thread per read)
static __global__ void k_row_read(const uint32_t* row_off, const uint32_t* row_cnt, uint32_t n_reads, uint32_t* row_read) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += gridDim.x * blockDim.x)
        for (uint32_t k = 0; k < row_cnt[r]; ++k) row_read[row_off[r] + k] = r;
}
static __global__ void k_add_u64(uint64_t* v, uint64_t n, uint64_t add) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) v[i] += add;
}

static void index_holes_sorted(const bsq_index* h, uint32_t flags, std::vector<TupleHole>& holes, std::vector<int64_t>& maxend) {
    holes.resize(h->holes.size());
    for (size_t i = 0; i < holes.size(); ++i) {
        // reference behaviour: offsets as stored in the row's datum, i.e. relative to its own row (bwa.cpp:100-104); the fix-up rebases them
        const int64_t base = (flags & BSQ_TUPLES_FIX_HOLE_OFFSETS) ? h->ann_offset[h->hole_ann[i]] : 0;
        holes[i].offset = h->holes[i].offset + base; holes[i].end = holes[i].offset + h->holes[i].len; holes[i].idx = (uint32_t)i; holes[i].amb = (uint8_t)h->holes[i].amb;
    }
    std::stable_sort(holes.begin(), holes.end(), [](const TupleHole& a, const TupleHole& b) { return a.offset < b.offset; });
    maxend.resize(holes.size());
    for (size_t i = 0; i < holes.size(); ++i) maxend[i] = i ? std::max(maxend[i - 1], holes[i].end) : holes[i].end;
}

struct TuplesImpl { bsq_tuples pub; void* host = nullptr; size_t host_bytes = 0; bool cached = false; };

// Row materialisation straight from the batch(es) that produced `R` and still sit in HBM: rows in read order (rows_compact), CIGAR
// pool, the reads' text -- nothing is uploaded again.  Returns 1 when the result is not (or no longer) resident: the caller then
// takes the general path that uploads rows, CIGARs and reads.
static int tuples_resident(bsq_index* h, const ResultImpl* R, uint32_t flags, bsq_tuples** out) {
    Batch* lanes[2] = {&h->batch, &h->batch2};
    Batch* piece[2]; int np = 0;
    uint64_t covered = 0;
    for (Batch* b : lanes) if (b->aligned && b->res_serial == R->serial && b->n) piece[np++] = b;
    if (np == 2 && piece[0]->res_read_base > piece[1]->res_read_base) std::swap(piece[0], piece[1]);
    for (int k = 0; k < np; ++k) { if (piece[k]->res_read_base != covered) return 1; covered += piece[k]->n; }
    if (np == 0 || covered != R->pub.n_reads) return 1;
    const uint64_t n_rows = R->pub.row_off[R->pub.n_reads];
    if (n_rows == 0) return 1;
    std::vector<TupleHole> holes; std::vector<int64_t> maxend;
    index_holes_sorted(h, flags, holes, maxend);
    TupleHole* d_holes = nullptr; int64_t* d_maxend = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    TuplesImpl* T = new TuplesImpl();
    T->pub.n_rows = n_rows; T->pub.n_bytes = 0; T->pub.device_ms = 0.f;
    int rc = BSQ_ERR;
    do {
#define TR(x) if ((x) != cudaSuccess) { bsq_set_error("bsq_result_tuples: %s", cudaGetErrorString(cudaGetLastError())); break; }
        TR(cudaEventCreate(&e0)); TR(cudaEventCreate(&e1));
        TR(cudaMalloc(&d_holes, (holes.size() + 1) * sizeof(TupleHole))); TR(cudaMalloc(&d_maxend, (holes.size() + 1) * 8));
        cudaStream_t s0 = piece[0]->st;
        cudaEventRecord(e0, s0);
        if (!holes.empty()) {
            TR(cudaMemcpyAsync(d_holes, holes.data(), holes.size() * sizeof(TupleHole), cudaMemcpyHostToDevice, s0));
            TR(cudaMemcpyAsync(d_maxend, maxend.data(), holes.size() * 8, cudaMemcpyHostToDevice, s0));
        }
        TR(cudaStreamSynchronize(s0));
        TupleParams P[2]; uint64_t nb[2] = {0, 0}; bool bad = false;
        for (int k = 0; k < np && !bad; ++k) {       // sizes of every piece (the pieces run on their own lanes' streams)
            Batch& b = *piece[k]; const uint64_t nr = b.out_rows;
            if (cudaSuccess != b.tup_off.ensure(3 * nr + 2) || cudaSuccess != b.tup_tmp.ensure(tuple_scan_tmp_elems(nr)) || cudaSuccess != b.tup_row_read.ensure(nr + 1) ||
                cudaSuccess != b.tup_nholes.ensure(2 * nr + 2) || cudaSuccess != b.tup_rm.ensure(3 * nr + 3)) { bsq_set_error("bsq_result_tuples: out of device memory"); bad = true; break; }
            TupleParams& q = P[k];
            q.rows = b.rows_compact.p; q.n_rows = nr; q.row_read = b.tup_row_read.p; q.cigar = b.cigar.p - b.cig_rebased; q.seqs = b.ascii.p; q.offs = b.offs.p;
            q.pac = h->d_pac; q.l_pac = h->meta.l_pac; q.ann_offset = h->d_ann_offset;
            q.holes = d_holes; q.hole_maxend = d_maxend; q.n_holes = (uint32_t)holes.size();
            q.nholes = b.tup_nholes.p; q.off = b.tup_off.p; q.ref_match = b.tup_rm.p; q.bytes = nullptr; q.fix_reverse = (flags & BSQ_TUPLES_FIX_REVERSE) != 0;
            if (nr) {
                k_row_read<<<(unsigned)std::min<uint64_t>((b.n + 255) / 256, 148 * 8), 256, 0, b.st>>>(b.row_off.p, b.row_cnt.p, (uint32_t)b.n, b.tup_row_read.p); ++h->timing.launches;
                cudaMemsetAsync(b.tup_off.p, 0, (3 * nr + 1) * 8, b.st);
                launch_tuple_sizes(q, b.tup_tmp.p, b.st, &h->timing.launches);
                cudaMemcpyAsync(b.ctl_host + 14, b.tup_off.p + 3 * nr, 8, cudaMemcpyDeviceToHost, b.st);
            }
        }
        if (bad) break;
        for (int k = 0; k < np; ++k) { TR(cudaStreamSynchronize(piece[k]->st)); if (piece[k]->out_rows) memcpy(&nb[k], piece[k]->ctl_host + 14, 8); }
        const uint64_t n_bytes = nb[0] + nb[1];
        // one pinned block (from the library's cache): off | ref_match | bytes
        const size_t off_b = (3 * n_rows + 1) * 8, rm_b = (n_rows * 12 + 7) & ~(size_t)7;
        size_t got = 0;
        T->host = pinned_get(off_b + rm_b + n_bytes + 64, &got);
        if (!T->host) { bsq_set_error("cannot allocate pinned host memory for the tuples"); break; }
        T->host_bytes = got; T->cached = true;
        T->pub.off = reinterpret_cast<uint64_t*>(T->host);
        T->pub.ref_match = reinterpret_cast<int32_t*>(static_cast<char*>(T->host) + off_b);
        T->pub.bytes = reinterpret_cast<uint8_t*>(static_cast<char*>(T->host) + off_b + rm_b);
        T->pub.n_bytes = n_bytes;
        uint64_t row_base = 0, byte_base = 0;
        for (int k = 0; k < np && !bad; ++k) {
            Batch& b = *piece[k]; const uint64_t nr = b.out_rows;
            if (nr) {
                if (cudaSuccess != b.tup_bytes.ensure(nb[k] + 64)) { bsq_set_error("bsq_result_tuples: out of device memory"); bad = true; break; }
                cudaMemsetAsync(b.tup_bytes.p, 0, nb[k] + 64, b.st);
                P[k].bytes = b.tup_bytes.p;
                launch_tuple_fill(P[k], b.st, &h->timing.launches);
                if (byte_base) { k_add_u64<<<(unsigned)std::min<uint64_t>((3 * nr + 256) / 256, 148 * 8), 256, 0, b.st>>>(b.tup_off.p, 3 * nr + 1, byte_base); ++h->timing.launches; }
                cudaMemcpyAsync(T->pub.off + 3 * row_base, b.tup_off.p, (3 * nr + 1) * 8, cudaMemcpyDeviceToHost, b.st);
                cudaMemcpyAsync(T->pub.ref_match + 3 * row_base, b.tup_rm.p, nr * 12, cudaMemcpyDeviceToHost, b.st);
                cudaMemcpyAsync(T->pub.bytes + byte_base, b.tup_bytes.p, nb[k], cudaMemcpyDeviceToHost, b.st);
            }
            row_base += nr; byte_base += nb[k];
        }
        if (bad) break;
        for (int k = 1; k < np; ++k) TR(cudaStreamSynchronize(piece[k]->st));
        cudaEventRecord(e1, s0);
        TR(cudaStreamSynchronize(s0)); TR(cudaGetLastError());
        T->pub.off[3 * n_rows] = n_bytes;
        cudaEventElapsedTime(&T->pub.device_ms, e0, e1);
        rc = BSQ_OK;
#undef TR
    } while (0);
    cudaFree(d_holes); cudaFree(d_maxend);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (rc != BSQ_OK) { if (T->host) pinned_put(T->host, T->host_bytes); delete T; return BSQ_ERR; }
    *out = &T->pub;
    return BSQ_OK;
}

int bsq_result_tuples(bsq_index* h, const bsq_result* res, const char* seqs, const uint64_t* offs, uint32_t flags, bsq_tuples** out) {
    BSQ_ENTRY();
    if (!h || !res || !out) { bsq_set_error("null argument"); return BSQ_ERR; }
    if (h->replica_pending) { bsq_set_error("replica without host state: call bsq_index_replica_finish after filling its arrays (the hole overlay of ref_subseq lives on the host)"); return BSQ_ERR; }
    {   // a result of this library whose batch is still in HBM needs no upload at all
        const ResultImpl* mine = nullptr;
        { std::lock_guard<std::mutex> lk(g_live_mu); for (const ResultImpl* r : g_live) if (&r->pub == res) { mine = r; break; } }
        static const bool no_res = getenv("BSQ_TUPLES_NO_RESIDENT") != nullptr;
        if (mine && h->meta.built && !no_res) {
            if (cudaSetDevice(h->device) != cudaSuccess) { bsq_set_error("cudaSetDevice failed"); return BSQ_ERR; }
            const int rc = tuples_resident(h, mine, flags, out);
            if (rc != 1) return rc;
        }
    }
    if (!offs) { bsq_set_error("null argument"); return BSQ_ERR; }
    if (!h->meta.built && res->n_reads && res->row_off[res->n_reads]) { bsq_set_error("index not built"); return BSQ_ERR; }
    CUDA_CHECK(cudaSetDevice(h->device));
    const uint64_t n_reads = res->n_reads, n_rows = n_reads ? res->row_off[n_reads] : 0;
    TuplesImpl* T = new TuplesImpl();
    T->pub.n_rows = n_rows; T->pub.n_bytes = 0; T->pub.device_ms = 0.f; T->pub.off = nullptr; T->pub.ref_match = nullptr; T->pub.bytes = nullptr;
    if (n_rows == 0) {
        T->host_bytes = 64;
        if (cudaHostAlloc(&T->host, T->host_bytes, cudaHostAllocDefault) != cudaSuccess) { delete T; bsq_set_error("pinned allocation failed"); return BSQ_ERR; }
        T->pub.off = reinterpret_cast<uint64_t*>(T->host); T->pub.off[0] = 0;
        T->pub.ref_match = reinterpret_cast<int32_t*>(T->pub.off + 1); T->pub.bytes = reinterpret_cast<uint8_t*>(T->pub.off + 2);
        *out = &T->pub;
        return BSQ_OK;
    }
    if (!seqs) { delete T; bsq_set_error("null argument"); return BSQ_ERR; }
    cudaStream_t st = h->stream;
    // host-side preparation: read index of every row; the index's holes sorted by offset with the running maximum of their ends
    std::vector<uint32_t> row_read(n_rows);
    for (uint64_t r = 0; r < n_reads; ++r) for (uint64_t k = res->row_off[r]; k < res->row_off[r + 1]; ++k) row_read[k] = (uint32_t)r;
    std::vector<TupleHole> holes; std::vector<int64_t> maxend;
    index_holes_sorted(h, flags, holes, maxend);
    const uint64_t total = offs[n_reads] - offs[0];
    RowPub* d_rows = nullptr; uint32_t *d_row_read = nullptr, *d_cigar = nullptr, *d_nholes = nullptr; uint8_t *d_seqs = nullptr, *d_bytes = nullptr;
    uint64_t *d_offs = nullptr, *d_off = nullptr, *d_tmp = nullptr; TupleHole* d_holes = nullptr; int64_t* d_maxend = nullptr; int32_t* d_rm = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = BSQ_ERR;
    do {
#define TC(x) if ((x) != cudaSuccess) { bsq_set_error("bsq_result_tuples: %s", cudaGetErrorString(cudaGetLastError())); break; }
        TC(cudaEventCreate(&e0)); TC(cudaEventCreate(&e1));
        TC(cudaMalloc(&d_rows, n_rows * sizeof(RowPub))); TC(cudaMalloc(&d_row_read, n_rows * 4)); TC(cudaMalloc(&d_cigar, (res->n_cigar_words + 1) * 4));
        TC(cudaMalloc(&d_seqs, total + 64)); TC(cudaMalloc(&d_offs, (n_reads + 1) * 8)); TC(cudaMalloc(&d_nholes, n_rows * 8));
        TC(cudaMalloc(&d_off, (3 * n_rows + 1) * 8)); TC(cudaMalloc(&d_tmp, tuple_scan_tmp_elems(n_rows) * 8)); TC(cudaMalloc(&d_rm, n_rows * 12));
        TC(cudaMalloc(&d_holes, (holes.size() + 1) * sizeof(TupleHole))); TC(cudaMalloc(&d_maxend, (holes.size() + 1) * 8));
        std::vector<uint64_t> rel(n_reads + 1);
        for (uint64_t i = 0; i <= n_reads; ++i) rel[i] = offs[i] - offs[0];
        cudaEventRecord(e0, st);
        TC(cudaMemcpyAsync(d_rows, res->rows, n_rows * sizeof(RowPub), cudaMemcpyHostToDevice, st));
        TC(cudaMemcpyAsync(d_row_read, row_read.data(), n_rows * 4, cudaMemcpyHostToDevice, st));
        if (res->n_cigar_words) TC(cudaMemcpyAsync(d_cigar, res->cigar, res->n_cigar_words * 4, cudaMemcpyHostToDevice, st));
        if (total) TC(cudaMemcpyAsync(d_seqs, seqs + offs[0], total, cudaMemcpyHostToDevice, st));
        TC(cudaMemcpyAsync(d_offs, rel.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
        if (!holes.empty()) {
            TC(cudaMemcpyAsync(d_holes, holes.data(), holes.size() * sizeof(TupleHole), cudaMemcpyHostToDevice, st));
            TC(cudaMemcpyAsync(d_maxend, maxend.data(), holes.size() * 8, cudaMemcpyHostToDevice, st));
        }
        TC(cudaMemsetAsync(d_off, 0, (3 * n_rows + 1) * 8, st));
        TupleParams P;
        P.rows = d_rows; P.n_rows = n_rows; P.row_read = d_row_read; P.cigar = d_cigar; P.seqs = d_seqs; P.offs = d_offs;
        P.pac = h->d_pac; P.l_pac = h->meta.l_pac; P.ann_offset = h->d_ann_offset;
        P.holes = d_holes; P.hole_maxend = d_maxend; P.n_holes = (uint32_t)holes.size();
        P.nholes = d_nholes; P.off = d_off; P.ref_match = d_rm; P.bytes = nullptr; P.fix_reverse = (flags & BSQ_TUPLES_FIX_REVERSE) != 0;
        launch_tuple_sizes(P, d_tmp, st, &h->timing.launches);
        uint64_t n_bytes = 0;
        TC(cudaMemcpyAsync(&n_bytes, d_off + 3 * n_rows, 8, cudaMemcpyDeviceToHost, st));
        TC(cudaStreamSynchronize(st));
        TC(cudaMalloc(&d_bytes, n_bytes + 64));
        TC(cudaMemsetAsync(d_bytes, 0, n_bytes + 64, st));
        P.bytes = d_bytes;
        launch_tuple_fill(P, st, &h->timing.launches);
        // one pinned block: off | ref_match | bytes
        const size_t off_b = (3 * n_rows + 1) * 8, rm_b = (n_rows * 12 + 7) & ~(size_t)7;
        T->host_bytes = off_b + rm_b + n_bytes + 64;
        TC(cudaHostAlloc(&T->host, T->host_bytes, cudaHostAllocDefault));
        T->pub.off = reinterpret_cast<uint64_t*>(T->host);
        T->pub.ref_match = reinterpret_cast<int32_t*>(static_cast<char*>(T->host) + off_b);
        T->pub.bytes = reinterpret_cast<uint8_t*>(static_cast<char*>(T->host) + off_b + rm_b);
        T->pub.n_bytes = n_bytes;
        TC(cudaMemcpyAsync(T->pub.off, d_off, off_b, cudaMemcpyDeviceToHost, st));
        TC(cudaMemcpyAsync(T->pub.ref_match, d_rm, n_rows * 12, cudaMemcpyDeviceToHost, st));
        if (n_bytes) TC(cudaMemcpyAsync(T->pub.bytes, d_bytes, n_bytes, cudaMemcpyDeviceToHost, st));
        cudaEventRecord(e1, st);
        TC(cudaStreamSynchronize(st)); TC(cudaGetLastError());
        cudaEventElapsedTime(&T->pub.device_ms, e0, e1);
        rc = BSQ_OK;
#undef TC
    } while (0);
    cudaFree(d_rows); cudaFree(d_row_read); cudaFree(d_cigar); cudaFree(d_seqs); cudaFree(d_offs); cudaFree(d_nholes); cudaFree(d_off); cudaFree(d_tmp);
    cudaFree(d_rm); cudaFree(d_holes); cudaFree(d_maxend); cudaFree(d_bytes);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (rc != BSQ_OK) { if (T->host) cudaFreeHost(T->host); delete T; return BSQ_ERR; }
    *out = &T->pub;
    return BSQ_OK;
}

void bsq_tuples_free(bsq_tuples* t) {
    if (!t) return;
    TuplesImpl* T = reinterpret_cast<TuplesImpl*>(t);
    if (T->host) { if (T->cached) pinned_put(T->host, T->host_bytes); else cudaFreeHost(T->host); }
    delete T;
}

// ---- bulk text -> NUCLSEQ
struct NuclseqsImpl { bsq_nuclseqs pub; void* host = nullptr; };

int bsq_nuclseq_from_text_batch(int device, const char* text, const uint64_t* offs, uint64_t n, bsq_nuclseqs** out) {
    BSQ_ENTRY();
    if (!offs || !out || (n && !text && offs[n] != offs[0])) { bsq_set_error("null argument"); return BSQ_ERR; }
    if (cudaSetDevice(device) != cudaSuccess) { bsq_set_error("no CUDA device %d (there is no CPU fallback)", device); return BSQ_ERR; }
    std::vector<uint64_t> rel(n + 1), chunk_off(n + 1);
    uint64_t chunks = 0;
    for (uint64_t i = 0; i <= n; ++i) {
        rel[i] = offs[i] - offs[0];
        chunk_off[i] = chunks;
        if (i < n) {
            const uint64_t len = offs[i + 1] - offs[i];
            if (len > (uint64_t)(INT32_MAX / 4)) { bsq_set_error("provided sequence is too long"); return BSQ_ERR; }   // extension.cpp:50
            chunks += (len + 15) >> 4;
        }
    }
    const uint64_t total = rel[n];
    if (chunks >= (1ull << 32) - 2 || total >= (1ull << 32)) { bsq_set_error("batch too large for one call (4 G bases at most: the ranks of ambiguous bases are 32-bit)"); return BSQ_ERR; }
    NuclseqsImpl* S = new NuclseqsImpl();
    S->pub.n = n; S->pub.off = nullptr; S->pub.bytes = nullptr; S->pub.n_bytes = 0; S->pub.device_ms = 0.f;
    uint8_t *d_text = nullptr, *d_bytes = nullptr; uint64_t *d_offs = nullptr, *d_chunk_off = nullptr, *d_img = nullptr, *d_tmp64 = nullptr;
    uint32_t *d_amb = nullptr, *d_start = nullptr, *d_holes = nullptr, *d_tmp32 = nullptr; unsigned long long* d_bad = nullptr;
    cudaStream_t st = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = BSQ_ERR;
    do {
#define LC(x) if ((x) != cudaSuccess) { bsq_set_error("bsq_nuclseq_from_text_batch: %s", cudaGetErrorString(cudaGetLastError())); break; }
        LC(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); LC(cudaEventCreate(&e0)); LC(cudaEventCreate(&e1));
        LC(cudaMalloc(&d_text, total + 64)); LC(cudaMalloc(&d_offs, (n + 1) * 8)); LC(cudaMalloc(&d_chunk_off, (n + 1) * 8)); LC(cudaMalloc(&d_img, (n + 2) * 8));
        LC(cudaMalloc(&d_amb, (chunks + 2) * 4)); LC(cudaMalloc(&d_start, (chunks + 2) * 4)); LC(cudaMalloc(&d_holes, (n + 1) * 4));
        LC(cudaMalloc(&d_tmp32, loader_scan_tmp_elems(chunks) * 4)); LC(cudaMalloc(&d_tmp64, loader_scan_tmp_elems(n) * 8)); LC(cudaMalloc(&d_bad, 8));
        cudaEventRecord(e0, st);
        if (total) LC(cudaMemcpyAsync(d_text, text + offs[0], total, cudaMemcpyHostToDevice, st));
        LC(cudaMemcpyAsync(d_offs, rel.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
        LC(cudaMemcpyAsync(d_chunk_off, chunk_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, st));
        LC(cudaMemsetAsync(d_amb, 0, (chunks + 2) * 4, st)); LC(cudaMemsetAsync(d_start, 0, (chunks + 2) * 4, st)); LC(cudaMemsetAsync(d_img, 0, (n + 2) * 8, st));
        LC(cudaMemsetAsync(d_bad, 0xff, 8, st));
        LoaderParams P;
        P.text = d_text; P.offs = d_offs; P.n_seqs = n; P.chunk_off = d_chunk_off; P.n_chunks = chunks; P.cnt_amb = d_amb; P.cnt_start = d_start;
        P.holes_num = d_holes; P.img_off = d_img; P.first_invalid = d_bad; P.bytes = nullptr;
        uint64_t launches = 0;
        if (n) launch_loader_scan(P, d_tmp32, d_tmp64, st, &launches);
        uint64_t n_bytes = 0; unsigned long long bad = ~0ull;
        LC(cudaMemcpyAsync(&n_bytes, d_img + n, 8, cudaMemcpyDeviceToHost, st));
        LC(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, st));
        LC(cudaStreamSynchronize(st)); LC(cudaGetLastError());
        if (bad != ~0ull && (bad >> 63)) { bsq_set_error("sequence at text offset %llu needs a datum of 1 GiB or more (too many ambiguity runs): invalid memory alloc request size", (unsigned long long)(bad & ~(1ull << 63))); break; }
        if (bad != ~0ull) { bsq_set_error("invalid nucleotide in nuclseq_in: '%c'", text[offs[0] + bad]); break; }   // extension.cpp:53-58
        LC(cudaMalloc(&d_bytes, n_bytes + 64));
        LC(cudaMemsetAsync(d_bytes, 0, n_bytes + 64, st));
        P.bytes = d_bytes;
        if (n) launch_loader_fill(P, st, &launches);
        const size_t off_b = (n + 1) * 8;
        LC(cudaHostAlloc(&S->host, off_b + n_bytes + 64, cudaHostAllocDefault));
        S->pub.off = reinterpret_cast<uint64_t*>(S->host);
        S->pub.bytes = reinterpret_cast<uint8_t*>(static_cast<char*>(S->host) + off_b);
        S->pub.n_bytes = n_bytes;
        LC(cudaMemcpyAsync(S->pub.off, d_img, off_b, cudaMemcpyDeviceToHost, st));
        if (n_bytes) LC(cudaMemcpyAsync(S->pub.bytes, d_bytes, n_bytes, cudaMemcpyDeviceToHost, st));
        cudaEventRecord(e1, st);
        LC(cudaStreamSynchronize(st)); LC(cudaGetLastError());
        cudaEventElapsedTime(&S->pub.device_ms, e0, e1);
        rc = BSQ_OK;
#undef LC
    } while (0);
    cudaFree(d_text); cudaFree(d_bytes); cudaFree(d_offs); cudaFree(d_chunk_off); cudaFree(d_img); cudaFree(d_tmp64); cudaFree(d_amb); cudaFree(d_start);
    cudaFree(d_holes); cudaFree(d_tmp32); cudaFree(d_bad);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    if (rc != BSQ_OK) { if (S->host) cudaFreeHost(S->host); delete S; return BSQ_ERR; }
    *out = &S->pub;
    return BSQ_OK;
}

void bsq_nuclseqs_free(bsq_nuclseqs* s) {
    if (!s) return;
    NuclseqsImpl* S = reinterpret_cast<NuclseqsImpl*>(s);
    if (S->host) cudaFreeHost(S->host);
    delete S;
}

void bsq_result_free(bsq_result* r) {
    if (!r) return;
    result_delete(reinterpret_cast<ResultImpl*>(r));   // pub is the first member
}

int bsq_last_timing(const bsq_index* h, bsq_timing* t) {
    BSQ_ENTRY();
    if (!h || !t) { bsq_set_error("null argument"); return BSQ_ERR; }
    *t = h->timing;
    return BSQ_OK;
}

int bsq_set_counters(bsq_index* h, int on) { if (!h) return BSQ_ERR; h->collect_counters = on != 0; return BSQ_OK; }
int bsq_get_counters(const bsq_index* h, uint64_t* out8) { if (!h || !out8) return BSQ_ERR; memcpy(out8, h->counters, sizeof(h->counters)); return BSQ_OK; }

int bsq_index_verify(bsq_index* h, uint64_t n_samples, uint64_t seed, bsq_index_check* out) {
    BSQ_ENTRY();
    if (!h || !out) { bsq_set_error("null argument"); return BSQ_ERR; }
    if (!h->meta.built) { bsq_set_error("index not built"); return BSQ_ERR; }
    CUDA_CHECK(cudaSetDevice(h->device));
    const DevIndex ix = make_dev_index(h);
    unsigned long long* d = nullptr; unsigned long long v[IV_WORDS];
    CUDA_CHECK(cudaMalloc(&d, IV_WORDS * 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, h->stream);
    launch_index_verify(ix, n_samples, seed, 65536, d, h->stream);
    cudaEventRecord(e1, h->stream);
    cudaError_t ce = cudaMemcpyAsync(v, d, sizeof(v), cudaMemcpyDeviceToHost, h->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    if (ce != cudaSuccess) { bsq_set_error("bsq_index_verify: %s", cudaGetErrorString(ce)); return BSQ_ERR; }
    memset(out, 0, sizeof(*out));
    const uint64_t n = ix.seq_len;
    const unsigned __int128 N = n;
    // sum and sum of squares of 0..n, exact in 128 bits (n < 2^41), compared mod 2^64
    const uint64_t want_sum = (uint64_t)(N * (N + 1) / 2), want_sq = (uint64_t)(N * (N + 1) * (2 * N + 1) / 6);
    out->rows = n + 1;
    out->exhaustive = n_samples >= n + 1;
    out->sa_permutation_ok = v[IV_SA_SUM] == want_sum && v[IV_SA_SUMSQ] == want_sq && v[IV_SA_RANGE_BAD] == 0;
    out->sa_out_of_range = v[IV_SA_RANGE_BAD];
    out->order_checked = v[IV_ORDER_CHECKED]; out->order_bad = v[IV_ORDER_BAD]; out->order_undecided = v[IV_ORDER_UNDECIDED];
    out->rows_checked = v[IV_ROWS_CHECKED]; out->bwt_bad = v[IV_BWT_BAD]; out->lf_bad = v[IV_LF_BAD];
    out->occ_blocks_checked = v[IV_OCC_CHECKED]; out->occ_bad = v[IV_OCC_BAD];
    bool l2 = ix.L2[0] == 0;
    for (int c = 0; c < 4; ++c) l2 = l2 && ix.L2[c + 1] - ix.L2[c] == v[IV_TEXT_CNT + c] + v[IV_TEXT_CNT + 3 - c];
    out->l2_ok = l2;
    out->ms = ms;
    return BSQ_OK;
}

// the control words of the last batch's pipeline (queue sizes, tickets; layout at Batch::ctl): read after a call has ended
int bsq_debug_ctl(bsq_index* h, uint32_t* out64) {
    BSQ_ENTRY();
    if (!h || !out64 || !h->batch.ctl.p) { bsq_set_error("no batch"); return BSQ_ERR; }
    CUDA_CHECK(cudaSetDevice(h->device));
    CUDA_CHECK(cudaMemcpy(out64, h->batch.ctl.p, 64 * 4, cudaMemcpyDeviceToHost));
    return BSQ_OK;
}

int bsq_debug_seed(bsq_index* h, const char* seqs, const uint64_t* offs, uint64_t n, uint64_t* out, uint32_t cap, uint32_t* cnt) {
    BSQ_ENTRY();
    if (!h || !h->meta.built) { bsq_set_error("index not built"); return BSQ_ERR; }
    CUDA_CHECK(cudaSetDevice(h->device));
    std::vector<int64_t> ids(n, 0);
    if (upload_reads(h, h->batch, seqs, offs, ids.data(), n) != BSQ_OK) return BSQ_ERR;
    Batch& b = h->batch;
    const DevIndex ix = make_dev_index(h);
    if (ensure_kmer_table(h, ix) != BSQ_OK) return BSQ_ERR;
    b.intv_cap = cap;
    b.list_cap = std::max<uint32_t>(b.max_len + 1, cap);
    const int seed_warps = seed_resident_warps();
    CUDA_CHECK(b.intv.ensure((size_t)n * cap)); CUDA_CHECK(b.intv_cnt.ensure(n)); CUDA_CHECK(b.seed_scratch.ensure((size_t)seed_warps * 3 * b.list_cap));
    CUDA_CHECK(b.ctl.ensure(64));
    CUDA_CHECK(cudaMemsetAsync(b.ctl.p, 0, 64 * 4, h->stream));
    CUDA_CHECK(b.seed_todo.ensure(n + 1));
    if (b.max_len <= 496) { CUDA_CHECK(b.seed_pk.ensure((size_t)n * seed_thread_words(b.max_len))); CUDA_CHECK(b.seed_u32.ensure(3 * (size_t)n)); }
    SeedParams P = seed_params(h, b, ix, (uint32_t)n, cap, b.ctl.p, reinterpret_cast<unsigned long long*>(b.ctl.p + 8));
    if (seed_thread_usable(P, h->dopts, b.max_len)) launch_seed_thread(P, ix, h->dopts, b.max_len, b.ctl.p + 60, h->stream);
    else P.todo = nullptr;
    launch_seed(P, ix, h->dopts, h->stream, nullptr);
    uint32_t ctl[8];
    CUDA_CHECK(cudaMemcpyAsync(ctl, b.ctl.p, sizeof(ctl), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(out, b.intv.p, (size_t)n * cap * sizeof(Intv), cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(cnt, b.intv_cnt.p, n * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(h->counters, b.ctl.p + 8, 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    CUDA_CHECK(cudaGetLastError());
    b.intv_cap = 0; b.resident = false;
    if (ctl[4]) { bsq_set_error("interval capacity %u too small", cap); return BSQ_ERR; }
    return BSQ_OK;
}

static int dbg_common(const bsq_opts* o, int device, bsq_index** tmp) {
    *tmp = bsq_index_new(o, device);
    return *tmp ? BSQ_OK : BSQ_ERR;
}

static int dbg_ksw_extend(const bsq_opts* o, int device, uint64_t n_jobs, const uint8_t* q, const uint64_t* q_off, const uint8_t* t,
                          const uint64_t* t_off, const int32_t* w, const int32_t* end_bonus, const int32_t* h0, int32_t* out, int mode);

int bsq_debug_ksw_extend(const bsq_opts* o, int device, uint64_t n_jobs, const uint8_t* q, const uint64_t* q_off, const uint8_t* t,
                         const uint64_t* t_off, const int32_t* w, const int32_t* end_bonus, const int32_t* h0, int32_t* out) {
    BSQ_ENTRY();
    return dbg_ksw_extend(o, device, n_jobs, q, q_off, t, t_off, w, end_bonus, h0, out, 0);
}

int bsq_debug_ksw_extend_thread(const bsq_opts* o, int device, uint64_t n_jobs, const uint8_t* q, const uint64_t* q_off, const uint8_t* t,
                                const uint64_t* t_off, const int32_t* w, const int32_t* end_bonus, const int32_t* h0, int32_t* out, int reversed) {
    BSQ_ENTRY();
    return dbg_ksw_extend(o, device, n_jobs, q, q_off, t, t_off, w, end_bonus, h0, out, reversed ? 2 : 1);
}

}  // extern "C"

// mode 0: warp-cooperative kernel; 1 / 2: thread-per-extension kernel, query stored forward / back to front
static int dbg_ksw_extend(const bsq_opts* o, int device, uint64_t n_jobs, const uint8_t* q, const uint64_t* q_off, const uint8_t* t,
                          const uint64_t* t_off, const int32_t* w, const int32_t* end_bonus, const int32_t* h0, int32_t* out, int mode) {
    bsq_index* h;
    if (dbg_common(o, device, &h) != BSQ_OK) return BSQ_ERR;
    int rc = BSQ_ERR;
    uint8_t *dq = nullptr, *dt = nullptr; uint64_t *dqo = nullptr, *dto = nullptr; int *dw = nullptr, *deb = nullptr, *dh0 = nullptr, *dout = nullptr, *deh = nullptr;
    uint32_t* dtk = nullptr;
    uint32_t max_q = 0;
    for (uint64_t i = 0; i < n_jobs; ++i) max_q = std::max<uint32_t>(max_q, (uint32_t)(q_off[i + 1] - q_off[i]));
    const int blocks = 148 * 4, warps = blocks * 4;
    do {
#define DC(x) if ((x) != cudaSuccess) { bsq_set_error("debug extend: %s", cudaGetErrorString(cudaGetLastError())); break; }
        DC(cudaMalloc(&dq, q_off[n_jobs] + 16)); DC(cudaMalloc(&dt, t_off[n_jobs] + 16));
        DC(cudaMalloc(&dqo, (n_jobs + 1) * 8)); DC(cudaMalloc(&dto, (n_jobs + 1) * 8));
        DC(cudaMalloc(&dw, n_jobs * 4 + 4)); DC(cudaMalloc(&deb, n_jobs * 4 + 4)); DC(cudaMalloc(&dh0, n_jobs * 4 + 4)); DC(cudaMalloc(&dout, n_jobs * 24 + 4));
        DC(cudaMalloc(&deh, (size_t)warps * 2 * (max_q + 2) * 4)); DC(cudaMalloc(&dtk, 64));
        DC(cudaMemset(dtk, 0, 64));
        DC(cudaMemcpy(dq, q, q_off[n_jobs], cudaMemcpyHostToDevice)); DC(cudaMemcpy(dt, t, t_off[n_jobs], cudaMemcpyHostToDevice));
        DC(cudaMemcpy(dqo, q_off, (n_jobs + 1) * 8, cudaMemcpyHostToDevice)); DC(cudaMemcpy(dto, t_off, (n_jobs + 1) * 8, cudaMemcpyHostToDevice));
        DC(cudaMemcpy(dw, w, n_jobs * 4, cudaMemcpyHostToDevice)); DC(cudaMemcpy(deb, end_bonus, n_jobs * 4, cudaMemcpyHostToDevice));
        DC(cudaMemcpy(dh0, h0, n_jobs * 4, cudaMemcpyHostToDevice));
        if (mode == 0) launch_dbg_extend(h->dopts, (uint32_t)n_jobs, dq, dqo, dt, dto, dw, deb, dh0, dout, deh, max_q, dtk, nullptr, blocks, h->stream);
        else {
            if (max_q > (uint32_t)EXT_MEMO_MAXQ) { bsq_set_error("debug extend (thread kernel): query longer than %d", EXT_MEMO_MAXQ); break; }
            launch_dbg_extend_thread(h->dopts, (uint32_t)n_jobs, dq, dqo, dt, dto, dw, deb, dh0, dout, mode == 2, h->stream);
        }
        DC(cudaStreamSynchronize(h->stream)); DC(cudaGetLastError());
        DC(cudaMemcpy(out, dout, n_jobs * 24, cudaMemcpyDeviceToHost));
        rc = BSQ_OK;
    } while (0);
    cudaFree(dq); cudaFree(dt); cudaFree(dqo); cudaFree(dto); cudaFree(dw); cudaFree(deb); cudaFree(dh0); cudaFree(dout); cudaFree(deh); cudaFree(dtk);
    bsq_index_free(h);
    return rc;
}

extern "C" {

int bsq_debug_ksw_global(const bsq_opts* o, int device, uint64_t n_jobs, const uint8_t* q, const uint64_t* q_off, const uint8_t* t,
                         const uint64_t* t_off, const int32_t* w, int32_t* out_score, uint32_t* cigar, uint32_t cig_cap, int32_t* n_cigar) {
    BSQ_ENTRY();
    bsq_index* h;
    if (dbg_common(o, device, &h) != BSQ_OK) return BSQ_ERR;
    int rc = BSQ_ERR;
    uint8_t *dq = nullptr, *dt = nullptr, *dz = nullptr; uint64_t *dqo = nullptr, *dto = nullptr; int *dw = nullptr, *dsc = nullptr, *dnc = nullptr, *deh = nullptr;
    uint32_t *dtk = nullptr, *dcg = nullptr;
    uint32_t max_q = 0; size_t z_per_warp = 16;
    for (uint64_t i = 0; i < n_jobs; ++i) {
        uint32_t ql = (uint32_t)(q_off[i + 1] - q_off[i]), tl = (uint32_t)(t_off[i + 1] - t_off[i]);
        max_q = std::max<uint32_t>(max_q, ql);
        size_t ncol = std::min<size_t>(ql, 2 * (size_t)w[i] + 1);
        z_per_warp = std::max<size_t>(z_per_warp, ncol * tl + 16);
    }
    const int blocks = 148 * 2, warps = blocks * 4;
    do {
        DC(cudaMalloc(&dq, q_off[n_jobs] + 16)); DC(cudaMalloc(&dt, t_off[n_jobs] + 16));
        DC(cudaMalloc(&dqo, (n_jobs + 1) * 8)); DC(cudaMalloc(&dto, (n_jobs + 1) * 8));
        DC(cudaMalloc(&dw, n_jobs * 4 + 4)); DC(cudaMalloc(&dsc, n_jobs * 4 + 4)); DC(cudaMalloc(&dnc, n_jobs * 4 + 4));
        DC(cudaMalloc(&dcg, (size_t)n_jobs * cig_cap * 4 + 4));
        DC(cudaMalloc(&deh, (size_t)warps * 2 * (max_q + 2) * 4)); DC(cudaMalloc(&dtk, 64)); DC(cudaMalloc(&dz, (size_t)warps * z_per_warp));
        DC(cudaMemset(dtk, 0, 64));
        DC(cudaMemcpy(dq, q, q_off[n_jobs], cudaMemcpyHostToDevice)); DC(cudaMemcpy(dt, t, t_off[n_jobs], cudaMemcpyHostToDevice));
        DC(cudaMemcpy(dqo, q_off, (n_jobs + 1) * 8, cudaMemcpyHostToDevice)); DC(cudaMemcpy(dto, t_off, (n_jobs + 1) * 8, cudaMemcpyHostToDevice));
        DC(cudaMemcpy(dw, w, n_jobs * 4, cudaMemcpyHostToDevice));
        launch_dbg_global(h->dopts, (uint32_t)n_jobs, dq, dqo, dt, dto, dw, dsc, cigar ? dcg : nullptr, cig_cap, dnc, deh, max_q, dz, z_per_warp, dtk, blocks, h->stream);
        DC(cudaStreamSynchronize(h->stream)); DC(cudaGetLastError());
        DC(cudaMemcpy(out_score, dsc, n_jobs * 4, cudaMemcpyDeviceToHost));
        if (cigar) { DC(cudaMemcpy(cigar, dcg, (size_t)n_jobs * cig_cap * 4, cudaMemcpyDeviceToHost)); DC(cudaMemcpy(n_cigar, dnc, n_jobs * 4, cudaMemcpyDeviceToHost)); }
        rc = BSQ_OK;
    } while (0);
#undef DC
    cudaFree(dq); cudaFree(dt); cudaFree(dz); cudaFree(dqo); cudaFree(dto); cudaFree(dw); cudaFree(dsc); cudaFree(dnc); cudaFree(dcg); cudaFree(deh); cudaFree(dtk);
    bsq_index_free(h);
    return rc;
}

int bsq_bench_gather(bsq_index* h, uint64_t n_loads, int reps, double* gbs) {
    BSQ_ENTRY();
    if (!h || !h->meta.built || !gbs) { bsq_set_error("index not built"); return BSQ_ERR; }
    CUDA_CHECK(cudaSetDevice(h->device));
    // the table the random 64-byte reads go to: the largest array the seeding kernels gather from (SURVEY 8d asks for a table at least
    // the size of the index, so that the figure is an HBM number, not an L2 one): prefix table, else the full SA, else Occ
    const uint32_t* table = h->d_occ; uint64_t table_bytes = h->meta.arr_bytes[BSQ_ARR_OCC];
    if (h->meta.arr_bytes[BSQ_ARR_SA] > table_bytes) { table = reinterpret_cast<const uint32_t*>(h->d_sa); table_bytes = h->meta.arr_bytes[BSQ_ARR_SA]; }
    if (h->d_kmer && kmer_table_bytes(h->kmer_k) > table_bytes) { table = reinterpret_cast<const uint32_t*>(h->d_kmer); table_bytes = kmer_table_bytes(h->kmer_k); }
    const uint64_t n_blocks = table_bytes / 64;
    const int blocks = 148 * 8;                       // 8 CTAs of 256 threads per SM
    const uint64_t groups = (uint64_t)blocks * 256 / 16;
    uint64_t per_group = std::max<uint64_t>(8, (n_loads / groups) / 8 * 8);
    uint32_t* sink = nullptr;
    CUDA_CHECK(cudaMalloc(&sink, 64));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch_gather(table, n_blocks, per_group, blocks, sink, h->stream);   // warm-up
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0, h->stream);
        launch_gather(table, n_blocks, per_group, blocks, sink, h->stream);
        cudaEventRecord(e1, h->stream);
        CUDA_CHECK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    *gbs = (double)(per_group * groups) * 64.0 / (best * 1e-3) / 1e9;
    return BSQ_OK;
}

int bsq_bench_dpx(int device, int reps, double* gops) {
    BSQ_ENTRY();
    if (!gops) { bsq_set_error("null argument"); return BSQ_ERR; }
    CUDA_CHECK(cudaSetDevice(device));
    int* sink = nullptr;
    CUDA_CHECK(cudaMalloc(&sink, 64));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, iters = 4096;
    launch_dpx(iters, blocks, sink, 0);
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0, 0);
        launch_dpx(iters, blocks, sink, 0);
        cudaEventRecord(e1, 0);
        CUDA_CHECK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    *gops = (double)blocks * 256 * (double)iters * 8.0 / (best * 1e-3) / 1e9;   // DPX instructions (per thread) per second, in G
    return BSQ_OK;
}


// ---------------------------------------------------------------------------------------- one process, several GPUs
// a PostgreSQL backend is one process with one thread.  bsq_multi is that shape: ONE host thread drives every
// device with streams.  The index built on the first device is copied to the others over NVLink (peer copies of pac / Occ / SA /
// annotations + the host-side state), every device derives its own inverse SA and prefix table, a batch is cut into contiguous
// blocks of reads -- read i keeps lrand48 id i -- and the rows of all blocks land in ONE pinned result in read order.
struct bsq_multi {
    std::vector<bsq_index*> ix;      // ix[0] is the caller's (built) handle, the others are replicas owned here
    bsq_timing timing;
};

bsq_multi* bsq_multi_new(bsq_index* built, const int* devices, int n_devices) {
    BSQ_ENTRY();
    if (!built || !built->meta.built || !devices || n_devices < 1) { bsq_set_error("bsq_multi_new: a built index and a device list are needed"); return nullptr; }
    if (devices[0] != built->device) { bsq_set_error("bsq_multi_new: devices[0] must be the device the index was built on (%d)", built->device); return nullptr; }
    bsq_multi* m = new bsq_multi;
    memset(&m->timing, 0, sizeof(m->timing));
    m->ix.push_back(built);
    uint64_t hs_bytes = 0;
    bsq_index_host_state_size(built, &hs_bytes);
    std::vector<uint8_t> hs(hs_bytes);
    bsq_index_host_state_get(built, hs.data(), hs_bytes);
    bool ok = true;
    for (int k = 1; k < n_devices && ok; ++k) {
        bsq_index* r = bsq_index_new(&built->opts, devices[k]);
        if (!r) { ok = false; break; }
        m->ix.push_back(r);
        r->flags = built->flags;
        if (bsq_index_alloc_replica(r, &built->meta) != BSQ_OK) { ok = false; break; }
        int can = 0;
        cudaDeviceCanAccessPeer(&can, devices[k], built->device);
        if (can) { cudaSetDevice(devices[k]); cudaDeviceEnablePeerAccess(built->device, 0); cudaGetLastError(); }
        for (int what = 0; what < BSQ_ARR_COUNT && ok; ++what) {
            const uint64_t nb = built->meta.arr_bytes[what];
            if (nb && cudaMemcpyPeerAsync(index_array(r, what), devices[k], index_array(built, what), built->device, nb, r->stream) != cudaSuccess) {
                bsq_set_error("bsq_multi_new: peer copy to device %d failed: %s", devices[k], cudaGetErrorString(cudaGetLastError())); ok = false;
            }
        }
    }
    for (size_t k = 1; k < m->ix.size() && ok; ++k) {
        if (bsq_index_replica_finish(m->ix[k], hs.data(), hs_bytes) != BSQ_OK) { ok = false; break; }
        if (bsq_index_prepare(m->ix[k], nullptr) != BSQ_OK) ok = false;
    }
    if (ok && bsq_index_prepare(built, nullptr) != BSQ_OK) ok = false;
    if (!ok) { char keep[1024]; snprintf(keep, sizeof(keep), "%s", g_err); for (size_t k = 1; k < m->ix.size(); ++k) bsq_index_free(m->ix[k]); delete m; bsq_set_error("%s", keep); return nullptr; }
    return m;
}

void bsq_multi_free(bsq_multi* m) {
    if (!m) return;
    for (size_t k = 1; k < m->ix.size(); ++k) bsq_index_free(m->ix[k]);
    delete m;
}

int bsq_multi_devices(const bsq_multi* m) { return m ? (int)m->ix.size() : 0; }

static int multi_align(bsq_multi* m, const ReadSrc& S, uint64_t n, bsq_result** out) {
    const size_t D = m->ix.size();
    std::vector<uint64_t> lo(D + 1);
    for (size_t d = 0; d <= D; ++d) lo[d] = n * d / D;                       // contiguous blocks, as the oracle's thread sharding
    bsq_index* h0 = m->ix[0];
    const uint64_t session = h0->lrand_state;
    std::vector<cudaEvent_t> e0(D), e1(D);
    int rc = BSQ_OK;
    for (size_t d = 0; d < D; ++d) { cudaSetDevice(m->ix[d]->device); cudaEventCreate(&e0[d]); cudaEventCreate(&e1[d]); }
    // phase 1: every device's copies are queued before anything is waited for
    for (size_t d = 0; d < D && rc == BSQ_OK; ++d) {
        bsq_index* h = m->ix[d]; Batch& b = h->batch;
        cudaSetDevice(h->device);
        h->opts = h0->opts; h->dopts = h0->dopts; h->flags = h0->flags; h->lrand_state = session;
        bsq_timing& T = h->timing;
        T.launches = 0; T.h2d_bytes = T.d2h_bytes = 0; T.h2d = T.d2h = 0; T.notes = 0; T.seed = T.chain = T.extend = T.finalize = T.total = 0;
        memset(h->counters, 0, sizeof(h->counters));
        cudaEventRecord(e0[d], b.st);
        const uint64_t cnt = lo[d + 1] - lo[d];
        const int64_t* ids = S.ids ? S.ids + lo[d] : nullptr;
        rc = S.dbytes ? upload_datums_begin(h, b, S.dbytes, S.doff + lo[d], ids, cnt, lo[d]) : upload_reads(h, b, S.seqs, S.offs + lo[d], ids, cnt, lo[d]);
    }
    // phase 2: the kernel pipelines
    for (size_t d = 0; d < D && rc == BSQ_OK; ++d) {
        bsq_index* h = m->ix[d];
        cudaSetDevice(h->device);
        if (S.dbytes) rc = upload_datums_end(h, h->batch, lo[d + 1] > lo[d] ? S.doff[lo[d]] : 0);
        if (rc == BSQ_OK) rc = pipeline_enqueue(h, h->batch);
    }
    for (size_t d = 0; d < D && rc == BSQ_OK; ++d) { cudaSetDevice(m->ix[d]->device); rc = pipeline_finish(m->ix[d], m->ix[d]->batch); }
    // phase 3: one result, every device's rows behind the previous device's
    ResultImpl* R = nullptr;
    if (rc == BSQ_OK) {
        uint64_t rows = 0, cig = 0;
        for (size_t d = 0; d < D; ++d) { rows += m->ix[d]->batch.out_rows; cig += m->ix[d]->batch.out_cig; }
        if (cig > 0xffffffffull) { bsq_set_error("result has more CIGAR words than a 32-bit offset addresses"); rc = BSQ_ERR; }
        else { cudaSetDevice(h0->device); R = result_new(n, rows, cig, (h0->flags & BSQ_FLAG_ROWS_EXT) != 0); if (!R) rc = BSQ_ERR; }
    }
    if (rc == BSQ_OK) {
        uint64_t row_base = 0, cig_base = 0;
        std::vector<uint64_t> rb(D), cb(D);
        for (size_t d = 0; d < D && rc == BSQ_OK; ++d) {
            bsq_index* h = m->ix[d]; Batch& b = h->batch;
            cudaSetDevice(h->device);
            rb[d] = row_base; cb[d] = cig_base;
            rc = download_enqueue(h, b, R, lo[d], row_base, cig_base);
            cudaEventRecord(e1[d], b.st);
            row_base += b.out_rows; cig_base += b.out_cig;
        }
        for (size_t d = 0; d < D && rc == BSQ_OK; ++d) {
            bsq_index* h = m->ix[d]; Batch& b = h->batch;
            cudaSetDevice(h->device);
            if (cudaStreamSynchronize(b.st) != cudaSuccess) { bsq_set_error("device %d: %s", h->device, cudaGetErrorString(cudaGetLastError())); rc = BSQ_ERR; break; }
            download_finish(h, R, b.n, lo[d], b.out_rows, rb[d], cb[d], b.ctl_host && b.ctl_host[10], b.ext_tmp.data());
        }
        if (rc == BSQ_OK) { R->pub.row_off[n] = row_base; R->pub.n_cigar_words = cig_base; }
    }
    // timing: a device's time runs from its first copy to the end of its download; the call's time is the slowest device's
    bsq_timing& MT = m->timing;
    memset(&MT, 0, sizeof(MT));
    for (size_t d = 0; d < D; ++d) {
        bsq_index* h = m->ix[d];
        cudaSetDevice(h->device);
        float ms = 0;
        if (rc == BSQ_OK && cudaEventElapsedTime(&ms, e0[d], e1[d]) == cudaSuccess) MT.total = std::max(MT.total, ms);
        MT.seed = std::max(MT.seed, h->timing.seed); MT.chain = std::max(MT.chain, h->timing.chain); MT.extend = std::max(MT.extend, h->timing.extend);
        MT.finalize = std::max(MT.finalize, h->timing.finalize);
        MT.launches += h->timing.launches; MT.h2d_bytes += h->timing.h2d_bytes; MT.d2h_bytes += h->timing.d2h_bytes;
        cudaEventDestroy(e0[d]); cudaEventDestroy(e1[d]);
    }
    cudaSetDevice(h0->device);
    if (rc != BSQ_OK) { result_delete(R); return BSQ_ERR; }
    if (!S.ids) h0->lrand_state = lrand48_advance(session, n);
    *out = &R->pub;
    return BSQ_OK;
}

int bsq_multi_align_batch(bsq_multi* m, const char* seqs, const uint64_t* offs, const int64_t* ids, uint64_t n, bsq_result** out) {
    BSQ_ENTRY();
    if (!m || !out || (n && (!seqs || !offs))) { bsq_set_error("null argument"); return BSQ_ERR; }
    const ReadSrc S{seqs, offs, nullptr, nullptr, ids};
    return multi_align(m, S, n, out);
}

int bsq_multi_align_batch_datums(bsq_multi* m, const uint8_t* bytes, const uint64_t* off, const int64_t* ids, uint64_t n, bsq_result** out) {
    BSQ_ENTRY();
    if (!m || !out || (n && (!bytes || !off))) { bsq_set_error("null argument"); return BSQ_ERR; }
    const ReadSrc S{nullptr, nullptr, bytes, off, ids};
    return multi_align(m, S, n, out);
}

int bsq_multi_last_timing(const bsq_multi* m, bsq_timing* t) {
    BSQ_ENTRY();
    if (!m || !t) { bsq_set_error("null argument"); return BSQ_ERR; }
    *t = m->timing;
    return BSQ_OK;
}

}  // extern "C"
