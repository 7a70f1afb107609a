// capi.cu — the C ABI of libgenmap_b200.so (include/genmap_b200.h): index handles in HBM and the
// reference-facing gmb_map_frequencies call.  Host logic only; the device work is map_kernel.cu and
// index_build_gpu.cu.  No CPU fallback: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/genmap_b200.h"
#include "gmb_host.h"
#include "index_build_gpu.cuh"
#include "jump_table.cuh"
#include "locate.cuh"
#include "rle.cuh"
#include "map_kernel.cuh"

using namespace gmb;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

int cuda_fail(cudaError_t e, const char* what)
{
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return GMB_ERR_CUDA;
}

#define CU(call)                                              \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)

} // namespace

__global__ void k_decode_bwt(const RankBlock* __restrict__ B, uint64_t n, uint8_t* __restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RankBlock& b = B[i / kBlockBases];
    const uint32_t k = (uint32_t)(i % kBlockBases);
    out[i] = (uint8_t)(1u + (uint32_t)((b.w[k >> 6][0] >> (k & 63)) & 1u) + 2u * (uint32_t)((b.w[k >> 6][1] >> (k & 63)) & 1u));
}

__global__ void k_decode_bwt5(const RankBlock5* __restrict__ B, uint64_t n, uint8_t* __restrict__ out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RankBlock5& b = B[i / kBlockBases5];
    const uint32_t k = (uint32_t)(i % kBlockBases5);
    uint32_t c = 0;
    for (int pl = 0; pl < 3; ++pl) c |= ((b.plane[pl] >> k) & 1u) << pl;
    out[i] = (uint8_t)(1u + c);
}

__global__ void k_mark_sentinels(const uint32_t* __restrict__ S, uint32_t n_seq, uint8_t* __restrict__ out)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_seq) out[S[s]] = 0;
}

struct gmb_index {
    int device = 0;
    int sm_count = 0;
    uint8_t* d_blob = nullptr;
    bool owns_blob = false;
    IndexHeader h{};
    std::vector<uint64_t> limits; // host copy
    // per-handle scratch, grown on demand
    unsigned long long* d_counters = nullptr; // [0] work counter, [1..14] fetch counters (map_kernel.cuh), [15] finished pieces, [16] run count
    uint64_t* d_ranges = nullptr;
    size_t ranges_cap = 0;
    void* d_out = nullptr;
    size_t out_cap = 0;
    uint32_t* d_seq_to_file = nullptr; // --exclude-pseudo: device copy of the caller's mapping
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t s_compute = nullptr, s_copy = nullptr; // host-output pipeline (gmb_map_frequencies_range)
    cudaEvent_t ev_piece[2] = {nullptr, nullptr};
    // jump tables by depth (index 0 unused), built lazily; levels <= kJumpKeep stay cached
    int jump_depth_opt = -1;
    JtEntry* jt_uni[17] = {};
    uint32_t* jt_lof[17] = {};
    JtFull* jt_full[17] = {};    // both intervals per entry (searches that need the interval in SA(T) after the jump)
    bool jt_full_located[17] = {}; // ... with the entries of keys that occur once rewritten as LOCATED entries
    uint64_t jt_epoch = 0;       // bumped whenever a jump table is freed: cached plans holding its address are re-made
    uint64_t plan_n = 0;         // plan the searches as if the text had this many symbols (0 = the index's own n_bwt)
    // search plans by configuration (tables on the device, ready to launch): a map call of a configuration seen
    // before uploads nothing but its work ranges
    std::vector<std::unique_ptr<struct MapPlan>> plans;
    uint64_t plan_clock = 0;
    // Dna5 indices: the N pass of calls whose searches skip the text's N, by (K, E, strands); nfix_off: this index has
    // too many windows with N for it (every call then walks the N children as before)
    std::vector<std::unique_ptr<struct NFix>> nfix;
    bool nfix_off = false;
    // progress of the call in flight (gmb_progress): positions of finished pieces + chunks the kernel has handed out
    // d_counters[15] = positions of the pieces already finished (written in stream order by the host-output pipeline)
    std::atomic<uint64_t> prog_total{0}, prog_chunk{0};
    bool prog_in_pipeline = false;
    cudaStream_t s_progress = nullptr;
};

// Everything a launch needs for one configuration (K, E, block size, table flavour, jump depth, model text size)
struct MapPlan {
    std::string key;
    BlockTables tabs;
    uint32_t plan_depth = 0;
    uint32_t* d_tables = nullptr; // step words | SearchStart entries | substituted-offset sets
    size_t start_off = 0;
    uint64_t jt_epoch = ~0ull;    // ix->jt_epoch the SearchStart entries were written for
    uint64_t last_use = 0;
    const JtFull* e0_table = nullptr; // E = 0, one k-mer per chain, 16-byte entries: table and depth of its only search
    uint32_t e0_depth = 0;
    const uint2* d_keys = nullptr;    // E >= 1, block_kernel.cu: flat key lists by block size (inside d_tables)
    uint32_t key_off[kMaxBlockKmers + 1] = {}, key_n[kMaxBlockKmers + 1] = {};
    ~MapPlan() { if (d_tables) cudaFree(d_tables); }
};

// What a Dna5 call whose searches skip the text's N (MapCtx::skip_n) adds afterwards: the text windows with 1..E N
// (n_win of them, found by one pass over the N mask) were located through the index like csv queries, which gives
//   win_count[i]  the whole count of window win_pos[i] as a query (it overwrites what the search left there), and
//   hits          one entry per (N window, strand, occurrence) whose own window holds no N: by the symmetry of the
//                 Hamming distance (N mismatching everything on either side) these are exactly the alignments of the
//                 N-free queries to text windows with N — the ones the searches skipped; + 1 each.
// Positions in the concatenated text.  Built once per (K, E, strands) and handle, applied after every search kernel
// to the positions of its work ranges.
struct NFix {
    uint32_t K = 0, E = 0;
    bool revcompl = false;
    uint64_t n_win = 0, n_hits = 0;
    uint32_t* d_win_pos = nullptr;
    uint32_t* d_win_count = nullptr;
    uint32_t* d_hits = nullptr;
    double build_ms = 0;
    ~NFix() { if (d_win_pos) cudaFree(d_win_pos); if (d_win_count) cudaFree(d_win_count); if (d_hits) cudaFree(d_hits); }
};

namespace {
constexpr uint32_t kJumpKeep = 12; // all levels up to here together take < 300 MB
constexpr size_t kTableBytes = 256 << 10; // device scratch for the search tables of one call


void fill_ctx(const gmb_index* ix, MapCtx& cx)
{
    const uint8_t* base = ix->d_blob;
    cx.blk[0] = base + ix->h.off_fwd;
    cx.blk[1] = base + ix->h.off_rev;
    cx.sent[0] = reinterpret_cast<const uint32_t*>(base + ix->h.off_sent_fwd);
    cx.sent[1] = reinterpret_cast<const uint32_t*>(base + ix->h.off_sent_rev);
    for (int c = 0; c < 5; ++c) cx.C[c] = (uint32_t)ix->h.C[c];
    cx.n_bwt = (uint32_t)ix->h.n_bwt;
    cx.steps = nullptr; cx.p1_off = nullptr; cx.fl_off = nullptr;
    cx.starts = nullptr;
    cx.K = 0; cx.B = 1; cx.n_search = 0; cx.n_strands = 1; cx.maxv = 65535u;
    cx.sa = ix->h.off_sa ? reinterpret_cast<const uint32_t*>(base + ix->h.off_sa) : nullptr;
    cx.seq_start = reinterpret_cast<const uint32_t*>(base + ix->h.off_seq_start);
    cx.seq_to_file = nullptr;
    cx.n_seq = ix->h.n_seq; cx.own_file = 0; cx.all_files = 0;
    cx.loc_rows = nullptr;
    cx.text = reinterpret_cast<const uint64_t*>(base + ix->h.off_text);
    cx.nmask = ix->h.sigma == 5 ? reinterpret_cast<const uint64_t*>(base + ix->h.off_nmask) : nullptr;
    cx.n_text = ix->h.n_text;
    cx.E = 0;
    cx.skip_n = 0;
}

struct JumpNeeds { bool uni[17] = {}, lof[17] = {}, full[17] = {}; uint32_t top = 0; };

// GMB_LOCATE=0: tables without located entries (every search walks the index; for A/B measurements)
bool locate_enabled()
{
    const char* env = std::getenv("GMB_LOCATE");
    return !(env && env[0] == '0');
}

// use_full: the searches read 16-byte entries holding both intervals (the blocked instantiation, where SA(T) is needed
// after the jump); all_full: every search does (the 16-byte entries of keys that occur once are LOCATED,
// which ends most searches at the table read).  Otherwise 8-byte entries plus a separate array for the interval in SA(T).
JumpNeeds jump_needs(const std::vector<JumpPlan>& plans, bool use_full, bool all_full)
{
    JumpNeeds n;
    for (const JumpPlan& plan : plans)
        for (uint32_t s = 0; s < kMaxSearches; ++s) {
            const uint32_t d = plan.depth[s];
            if (!d) continue;
            if ((plan.need_lof[s] && use_full) || all_full) n.full[d] = true;
            else { n.uni[d] = true; if (plan.need_lof[s]) n.lof[d] = true; }
            n.top = std::max(n.top, d);
        }
    return n;
}

// Drop the big levels (> kJumpKeep) the current call does not use; they are rebuilt in about a second when needed.
// Only done when HBM is short: tables of other configurations stay cached otherwise.  Returns the bytes freed.
size_t evict_stale_jump_tables(gmb_index* ix, const JumpNeeds& n)
{
    size_t freed = 0;
    for (uint32_t d = kJumpKeep + 1; d <= 16; ++d) {
        const size_t e = (size_t)1 << (2 * d);
        if (ix->jt_uni[d] && !n.uni[d]) { cudaFree(ix->jt_uni[d]); ix->jt_uni[d] = nullptr; freed += e * sizeof(JtEntry); }
        if (ix->jt_lof[d] && !n.lof[d]) { cudaFree(ix->jt_lof[d]); ix->jt_lof[d] = nullptr; freed += e * sizeof(uint32_t); }
        if (ix->jt_full[d] && !n.full[d]) { cudaFree(ix->jt_full[d]); ix->jt_full[d] = nullptr; freed += e * sizeof(JtFull); }
    }
    if (freed) ++ix->jt_epoch;
    return freed;
}

// bytes that still have to be allocated for these needs (tables + the transient parent levels of the deepest one)
size_t missing_jump_bytes(const gmb_index* ix, const JumpNeeds& n)
{
    size_t bytes = 0;
    uint32_t deepest_missing = 0;
    for (uint32_t d = 1; d <= 16; ++d) {
        const size_t e = (size_t)1 << (2 * d);
        if (n.uni[d] && !ix->jt_uni[d]) { bytes += e * sizeof(JtEntry); deepest_missing = d; }
        if (n.lof[d] && !ix->jt_lof[d]) { bytes += e * sizeof(uint32_t); deepest_missing = d; }
        if (n.full[d] && !ix->jt_full[d]) { bytes += e * sizeof(JtFull); deepest_missing = d; }
    }
    for (uint32_t d = 1; d < deepest_missing; ++d) // every level below is built as uni + lof on the way
        if (!ix->jt_uni[d] || !ix->jt_lof[d]) bytes += ((size_t)1 << (2 * d)) * (sizeof(JtEntry) + sizeof(uint32_t));
    return bytes;
}

size_t jump_table_bytes(const gmb_index* ix)
{
    size_t bytes = 0;
    for (uint32_t d = 1; d <= 16; ++d) {
        const size_t e = (size_t)1 << (2 * d);
        if (ix->jt_uni[d]) bytes += e * sizeof(JtEntry);
        if (ix->jt_lof[d]) bytes += e * sizeof(uint32_t);
        if (ix->jt_full[d]) bytes += e * sizeof(JtFull);
    }
    return bytes;
}

// cudaMalloc that, when HBM is short, drops the cached tables this call does not use and tries again
template <class T>
cudaError_t jt_alloc(gmb_index* ix, const JumpNeeds& n, T** p, size_t bytes)
{
    cudaError_t err = cudaMalloc(p, bytes);
    if (err == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        if (evict_stale_jump_tables(ix, n)) err = cudaMalloc(p, bytes);
    }
    return err;
}

// make sure the tables these needs name exist on the device (built level by level, cached)
int ensure_jump_tables(gmb_index* ix, const JumpNeeds& n, cudaStream_t stream)
{
    uint32_t top = 0;
    const bool want_located = locate_enabled();
    for (uint32_t d = 1; d <= 16; ++d) {
        if (n.full[d] && ix->jt_full[d] && ix->jt_full_located[d] != want_located) { // built for the other setting of GMB_LOCATE
            cudaFree(ix->jt_full[d]); ix->jt_full[d] = nullptr; ++ix->jt_epoch;
        }
        if ((n.uni[d] && !ix->jt_uni[d]) || (n.lof[d] && !ix->jt_lof[d]) || (n.full[d] && !ix->jt_full[d])) top = d;
    }
    if (top == 0) return GMB_OK; // everything this call needs is cached
    MapCtx cx;
    fill_ctx(ix, cx);
    bool transient[17] = {}; // big parent levels built only to reach a deeper one: dropped again below
    for (uint32_t d = 1; d <= top; ++d) {
        const size_t e = (size_t)1 << (2 * d);
        const bool parent_for_later = d < top; // deeper levels extend this one: needs uni + lof
        const bool want_lof = parent_for_later || n.lof[d];
        if ((parent_for_later || n.uni[d]) && !(ix->jt_uni[d] && (!want_lof || ix->jt_lof[d]))) {
            if (ix->jt_uni[d]) { cudaFree(ix->jt_uni[d]); ix->jt_uni[d] = nullptr; ++ix->jt_epoch; }
            if (ix->jt_lof[d]) { cudaFree(ix->jt_lof[d]); ix->jt_lof[d] = nullptr; ++ix->jt_epoch; }
            transient[d] = d > kJumpKeep && !n.uni[d] && !n.lof[d];
            cudaError_t err = jt_alloc(ix, n, &ix->jt_uni[d], e * sizeof(JtEntry));
            if (err == cudaSuccess && want_lof) err = jt_alloc(ix, n, &ix->jt_lof[d], e * sizeof(uint32_t));
            if (err == cudaSuccess)
                err = build_jump_level(cx, ix->h.sigma, d, ix->jt_uni[d - 1], ix->jt_lof[d - 1], ix->jt_uni[d], ix->jt_lof[d], nullptr, stream);
            if (err != cudaSuccess) return cuda_fail(err, "jump table");
        }
        if (n.full[d] && !ix->jt_full[d]) {
            cudaError_t err = jt_alloc(ix, n, &ix->jt_full[d], e * sizeof(JtFull));
            if (err == cudaSuccess)
                err = build_jump_level(cx, ix->h.sigma, d, ix->jt_uni[d - 1], ix->jt_lof[d - 1], nullptr, nullptr, ix->jt_full[d], stream);
            ix->jt_full_located[d] = want_located;
            if (err == cudaSuccess && want_located) // keys that occur once: position + context instead of intervals
                err = locate_jump_singletons(reinterpret_cast<const uint64_t*>(ix->d_blob + ix->h.off_text),
                                             ix->h.sigma == 5 ? reinterpret_cast<const uint64_t*>(ix->d_blob + ix->h.off_nmask) : nullptr,
                                             ix->h.n_text, cx.seq_start, ix->h.n_seq, d, ix->jt_full[d], stream);
            if (err != cudaSuccess) return cuda_fail(err, "jump table");
        }
    }
    CU(cudaStreamSynchronize(stream));
    for (uint32_t d = kJumpKeep + 1; d <= 16; ++d)
        if (transient[d]) {
            if (ix->jt_uni[d] && !n.uni[d]) { cudaFree(ix->jt_uni[d]); ix->jt_uni[d] = nullptr; ++ix->jt_epoch; }
            if (ix->jt_lof[d] && !n.lof[d]) { cudaFree(ix->jt_lof[d]); ix->jt_lof[d] = nullptr; ++ix->jt_epoch; }
        }
    return GMB_OK;
}
} // namespace

extern "C" {

const char* gmb_last_error(void) { return g_err.c_str(); }
const char* gmb_version(void) { return "genmap-b200 0.1 (index format 2, sm_100a)"; }

int gmb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int gmb_index_build(const uint8_t* codes, const uint64_t* limits, uint32_t n_seq, uint32_t flags, int device,
                    void** blob_out, uint64_t* bytes_out)
{
    if (!codes || !limits || !blob_out || !bytes_out) return fail(GMB_ERR_ARG, "gmb_index_build: NULL argument");
    std::string err;
    const bool with_sa = (flags & GMB_BUILD_WITH_SA) != 0;
    if (flags & GMB_BUILD_ON_GPU) {
        void* p = nullptr;
        uint64_t bytes = 0;
        int rc = build_index_gpu(codes, limits, n_seq, with_sa, device, &p, &bytes, err);
        if (rc != 0) return fail(rc, err);
        *blob_out = p;
        *bytes_out = bytes;
        return GMB_OK;
    }
    Blob b;
    if (!build_index_host(codes, limits, n_seq, with_sa, b, err))
        return fail(err.find("not supported") != std::string::npos || err.find("too") != std::string::npos ? GMB_ERR_UNSUPPORTED : GMB_ERR_ARG, err);
    void* p = std::malloc(b.bytes);
    if (!p) return fail(GMB_ERR_NOMEM, "out of host memory");
    std::memcpy(p, b.data(), b.bytes);
    *blob_out = p;
    *bytes_out = b.bytes;
    return GMB_OK;
}

void gmb_blob_free(void* blob) { std::free(blob); }

int gmb_blob_save(const void* blob, uint64_t bytes, const char* path)
{
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(GMB_ERR_IO, std::string("cannot write ") + path);
    const size_t w = std::fwrite(blob, 1, bytes, f);
    if (std::fclose(f) != 0 || w != bytes) return fail(GMB_ERR_IO, std::string("short write to ") + path);
    return GMB_OK;
}

static int finish_open(gmb_index* ix, gmb_index** out)
{
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, ix->device));
    if (const char* g = std::getenv("GMB_L2_FETCH")) { // tuning knob: L2 fetch granularity hint (32/64/128 bytes)
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)std::atoi(g));
        cudaGetLastError();
    }
    ix->sm_count = prop.multiProcessorCount;
    ix->limits.resize((size_t)ix->h.n_seq + 1);
    CU(cudaMemcpy(ix->limits.data(), ix->d_blob + ix->h.off_limits, ix->limits.size() * 8, cudaMemcpyDeviceToHost));
    CU(cudaMalloc(&ix->d_counters, 24 * sizeof(unsigned long long)));
    CU(cudaMemset(ix->d_counters, 0, 24 * sizeof(unsigned long long)));
    CU(cudaEventCreate(&ix->ev0));
    CU(cudaEventCreate(&ix->ev1));
    CU(cudaStreamCreateWithFlags(&ix->s_progress, cudaStreamNonBlocking));
    *out = ix;
    return GMB_OK;
}

int gmb_index_from_blob(const void* host_blob, uint64_t bytes, int device, gmb_index** out)
{
    if (!host_blob || !out) return fail(GMB_ERR_ARG, "gmb_index_from_blob: NULL argument");
    std::string err;
    if (!validate_blob(static_cast<const uint8_t*>(host_blob), bytes, err)) return fail(GMB_ERR_IO, err);
    if (gmb_device_count() <= device || device < 0) return fail(GMB_ERR_CUDA, "no such CUDA device (the map path has no CPU fallback)");
    CU(cudaSetDevice(device));
    gmb_index* ix = new (std::nothrow) gmb_index;
    if (!ix) return fail(GMB_ERR_NOMEM, "out of host memory");
    ix->device = device;
    std::memcpy(&ix->h, host_blob, sizeof(IndexHeader));
    cudaError_t e = cudaMalloc(&ix->d_blob, ix->h.total_bytes);
    if (e != cudaSuccess) { delete ix; return cuda_fail(e, "cudaMalloc(index blob)"); }
    ix->owns_blob = true;
    e = cudaMemcpy(ix->d_blob, host_blob, ix->h.total_bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { gmb_index_close(ix); return cuda_fail(e, "cudaMemcpy(index blob)"); }
    int rc = finish_open(ix, out);
    if (rc != GMB_OK) gmb_index_close(ix);
    return rc;
}

int gmb_index_build_device(const uint8_t* codes, const uint64_t* limits, uint32_t n_seq, uint32_t flags, int device,
                           gmb_index** out, double* timings_ms)
{
    if (!codes || !limits || !out) return fail(GMB_ERR_ARG, "gmb_index_build_device: NULL argument");
    std::string err;
    uint8_t* d_blob = nullptr;
    IndexHeader h;
    GpuBuildTimings tm;
    int rc = build_index_gpu_device(codes, limits, n_seq, (flags & GMB_BUILD_WITH_SA) != 0, device, &d_blob, &h, &tm, err);
    if (rc != GMB_OK) return fail(rc, err);
    if (timings_ms) { timings_ms[0] = tm.h2d_ms; timings_ms[1] = tm.sort_ms; timings_ms[2] = tm.pack_ms; timings_ms[3] = tm.total_ms; }
    gmb_index* ix = new (std::nothrow) gmb_index;
    if (!ix) { cudaFree(d_blob); return fail(GMB_ERR_NOMEM, "out of host memory"); }
    ix->device = device;
    ix->h = h;
    ix->d_blob = d_blob;
    ix->owns_blob = true;
    rc = finish_open(ix, out);
    if (rc != GMB_OK) gmb_index_close(ix);
    return rc;
}

int gmb_index_adopt_device(void* device_blob, uint64_t bytes, int device, gmb_index** out)
{
    if (!device_blob || !out) return fail(GMB_ERR_ARG, "gmb_index_adopt_device: NULL argument");
    if (gmb_device_count() <= device || device < 0) return fail(GMB_ERR_CUDA, "no such CUDA device");
    CU(cudaSetDevice(device));
    IndexHeader h;
    if (bytes < sizeof(h)) return fail(GMB_ERR_IO, "index blob too small");
    CU(cudaMemcpy(&h, device_blob, sizeof(h), cudaMemcpyDeviceToHost));
    std::string err;
    if (!validate_header(h, bytes, err)) return fail(GMB_ERR_IO, err);
    gmb_index* ix = new (std::nothrow) gmb_index;
    if (!ix) return fail(GMB_ERR_NOMEM, "out of host memory");
    ix->device = device;
    ix->h = h;
    ix->d_blob = static_cast<uint8_t*>(device_blob);
    ix->owns_blob = false;
    int rc = finish_open(ix, out);
    if (rc != GMB_OK) gmb_index_close(ix);
    return rc;
}

int gmb_index_replicate(const gmb_index* src, int device, gmb_index** out)
{
    if (!src || !out) return fail(GMB_ERR_ARG, "gmb_index_replicate: NULL argument");
    if (gmb_device_count() <= device || device < 0) return fail(GMB_ERR_CUDA, "no such CUDA device");
    gmb_index* ix = new (std::nothrow) gmb_index;
    if (!ix) return fail(GMB_ERR_NOMEM, "out of host memory");
    ix->device = device;
    ix->h = src->h;
    ix->jump_depth_opt = src->jump_depth_opt;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&ix->d_blob, src->h.total_bytes);
    if (e != cudaSuccess) { delete ix; return cuda_fail(e, "cudaMalloc(index replica)"); }
    ix->owns_blob = true;
    int can = 0; // direct NVLink / PCIe peer copy when the devices can address each other; staged by the driver otherwise
    if (device != src->device && cudaDeviceCanAccessPeer(&can, device, src->device) == cudaSuccess && can) {
        cudaDeviceEnablePeerAccess(src->device, 0);
        cudaGetLastError(); // already enabled is fine
    }
    e = device == src->device ? cudaMemcpy(ix->d_blob, src->d_blob, src->h.total_bytes, cudaMemcpyDeviceToDevice)
                              : cudaMemcpyPeer(ix->d_blob, device, src->d_blob, src->device, src->h.total_bytes);
    if (e != cudaSuccess) { gmb_index_close(ix); return cuda_fail(e, "cudaMemcpyPeer(index blob)"); }
    int rc = finish_open(ix, out);
    if (rc != GMB_OK) gmb_index_close(ix);
    return rc;
}

int gmb_index_import_reference(const char* dir, void** blob_out, uint64_t* bytes_out)
{
    if (!dir || !blob_out || !bytes_out) return fail(GMB_ERR_ARG, "gmb_index_import_reference: NULL argument");
    Blob b;
    std::string err;
    if (!import_reference_index(dir, b, err)) return fail(GMB_ERR_IO, err);
    void* p = std::malloc(b.bytes);
    if (!p) return fail(GMB_ERR_NOMEM, "out of host memory");
    std::memcpy(p, b.data(), b.bytes);
    *blob_out = p;
    *bytes_out = b.bytes;
    return GMB_OK;
}

int gmb_blob_export_reference(const void* blob, uint64_t bytes, const char* dir, const char* const* ids, uint32_t n_ids,
                              int fasta_directory, uint32_t sampling)
{
    if (!blob || !dir || (!ids && n_ids)) return fail(GMB_ERR_ARG, "gmb_blob_export_reference: NULL argument");
    std::vector<std::string> lines;
    for (uint32_t i = 0; i < n_ids; ++i) lines.emplace_back(ids[i] ? ids[i] : "");
    std::string err;
    if (!export_reference_index(static_cast<const uint8_t*>(blob), bytes, dir, lines, fasta_directory != 0, sampling, err))
        return fail(err.find("write") != std::string::npos ? GMB_ERR_IO : GMB_ERR_UNSUPPORTED, err);
    return GMB_OK;
}

int gmb_index_open(const char* dir, int device, gmb_index** out)
{
    if (!dir || !out) return fail(GMB_ERR_ARG, "gmb_index_open: NULL argument");
    std::string path = std::string(dir);
    if (!path.empty() && path.back() != '/') path += '/';
    const std::string ref_probe = path + "index.lf.drv";
    path += "index.gmb";
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) {
        if (FILE* r = std::fopen(ref_probe.c_str(), "rb")) { // an index written by the reference itself
            std::fclose(r);
            void* blob = nullptr;
            uint64_t bytes = 0;
            int rc = gmb_index_import_reference(dir, &blob, &bytes);
            if (rc != GMB_OK) return rc;
            rc = gmb_index_from_blob(blob, bytes, device, out);
            std::free(blob);
            return rc;
        }
        return fail(GMB_ERR_IO, "cannot open " + path);
    }
    std::fseek(f, 0, SEEK_END);
    const long long sz = std::ftell(f);
    std::fclose(f);
    if (sz < (long long)sizeof(IndexHeader)) return fail(GMB_ERR_IO, "index blob too small: " + path);
    if (gmb_device_count() <= device || device < 0) return fail(GMB_ERR_CUDA, "no such CUDA device (the map path has no CPU fallback)");
    // The blob goes file -> HBM through a small ring of pinned buffers filled by reader threads (pread) while earlier
    // pieces are already on their way to the device: no pinned allocation of the blob's size (15.75 GB with the suffix
    // array at 3 Gbp: the allocation alone took longer than the search), file reads and H2D copies overlap.
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return fail(GMB_ERR_IO, "cannot open " + path);
    IndexHeader h;
    if (::pread(fd, &h, sizeof h, 0) != (ssize_t)sizeof h) { ::close(fd); return fail(GMB_ERR_IO, "short read from " + path); }
    std::string verr;
    if (!validate_header(h, (uint64_t)sz, verr)) { ::close(fd); return fail(GMB_ERR_IO, verr); }
    if (cudaError_t es = cudaSetDevice(device); es != cudaSuccess) { ::close(fd); return cuda_fail(es, "cudaSetDevice"); }
    gmb_index* ix = new (std::nothrow) gmb_index;
    if (!ix) { ::close(fd); return fail(GMB_ERR_NOMEM, "out of host memory"); }
    ix->device = device;
    ix->h = h;
    cudaError_t e = cudaMalloc(&ix->d_blob, h.total_bytes);
    if (e != cudaSuccess) { ::close(fd); delete ix; return cuda_fail(e, "cudaMalloc(index blob)"); }
    ix->owns_blob = true;
    constexpr size_t kPiece = 64u << 20;
    constexpr int kSlots = 4;
    uint8_t* ring = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t done[kSlots] = {};
    e = cudaMallocHost(&ring, kPiece * kSlots);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (int k = 0; k < kSlots && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming);
    const uint64_t total = h.total_bytes, n_pieces = (total + kPiece - 1) / kPiece;
    bool io_error = false;
    for (uint64_t q0 = 0; q0 < n_pieces && e == cudaSuccess && !io_error; q0 += kSlots) {
        const int n = (int)std::min<uint64_t>(kSlots, n_pieces - q0);
        std::vector<std::thread> readers;
        std::vector<char> bad(n, 0);
        for (int k = 0; k < n; ++k) {
            if (q0) { e = cudaEventSynchronize(done[k]); if (e != cudaSuccess) break; } // the slot's previous copy has left it
            readers.emplace_back([&, k] {
                const uint64_t off = (q0 + k) * kPiece, len = std::min<uint64_t>(kPiece, total - off);
                uint64_t got = 0;
                while (got < len) {
                    const ssize_t r = ::pread(fd, ring + k * kPiece + got, len - got, (off_t)(off + got));
                    if (r <= 0) { bad[k] = 1; return; }
                    got += (uint64_t)r;
                }
            });
        }
        for (int k = 0; k < (int)readers.size(); ++k) {
            readers[k].join();
            if (bad[k]) { io_error = true; continue; }
            const uint64_t off = (q0 + k) * kPiece, len = std::min<uint64_t>(kPiece, total - off);
            if (e == cudaSuccess && !io_error) e = cudaMemcpyAsync(ix->d_blob + off, ring + k * kPiece, len, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess && !io_error) e = cudaEventRecord(done[k], st);
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    ::close(fd);
    for (int k = 0; k < kSlots; ++k) if (done[k]) cudaEventDestroy(done[k]);
    if (st) cudaStreamDestroy(st);
    if (ring) cudaFreeHost(ring);
    if (io_error) { gmb_index_close(ix); return fail(GMB_ERR_IO, "short read from " + path); }
    if (e != cudaSuccess) { gmb_index_close(ix); return cuda_fail(e, "loading the index blob"); }
    int rc = finish_open(ix, out);
    if (rc != GMB_OK) gmb_index_close(ix);
    return rc;
}

int gmb_index_close(gmb_index* ix)
{
    if (!ix) return GMB_OK;
    cudaSetDevice(ix->device);
    if (ix->owns_blob && ix->d_blob) cudaFree(ix->d_blob);
    if (ix->d_counters) cudaFree(ix->d_counters);
    if (ix->d_ranges) cudaFree(ix->d_ranges);
    if (ix->d_out) cudaFree(ix->d_out);
    if (ix->d_seq_to_file) cudaFree(ix->d_seq_to_file);
    for (int d = 0; d < 17; ++d) { if (ix->jt_uni[d]) cudaFree(ix->jt_uni[d]); if (ix->jt_lof[d]) cudaFree(ix->jt_lof[d]); if (ix->jt_full[d]) cudaFree(ix->jt_full[d]); }
    ix->plans.clear();
    ix->nfix.clear();
    if (ix->s_progress) cudaStreamDestroy(ix->s_progress);
    if (ix->s_compute) cudaStreamDestroy(ix->s_compute);
    if (ix->s_copy) cudaStreamDestroy(ix->s_copy);
    for (int i = 0; i < 2; ++i) if (ix->ev_piece[i]) cudaEventDestroy(ix->ev_piece[i]);
    if (ix->ev0) cudaEventDestroy(ix->ev0);
    if (ix->ev1) cudaEventDestroy(ix->ev1);
    delete ix;
    return GMB_OK;
}

int gmb_index_get_info(const gmb_index* ix, gmb_index_info* info)
{
    if (!ix || !info) return fail(GMB_ERR_ARG, "gmb_index_get_info: NULL argument");
    std::memset(info, 0, sizeof(*info));
    info->n_text = ix->h.n_text;
    info->n_bwt = ix->h.n_bwt;
    info->n_seq = ix->h.n_seq;
    info->has_sa = ix->h.off_sa != 0;
    info->blob_bytes = ix->h.total_bytes;
    info->rank_block_bytes = block_bytes(ix->h.sigma);
    info->alphabet_size = (int32_t)ix->h.sigma;
    info->device_blob = ix->d_blob;
    info->device = ix->device;
    info->jump_table_bytes = jump_table_bytes(ix);
    return GMB_OK;
}

int gmb_index_set_plan_text_size(gmb_index* ix, uint64_t n_symbols)
{
    if (!ix) return fail(GMB_ERR_ARG, "gmb_index_set_plan_text_size: NULL argument");
    ix->plan_n = n_symbols;
    return GMB_OK;
}

int gmb_progress(gmb_index* ix, uint64_t* done, uint64_t* total)
{
    if (!ix || !done || !total) return fail(GMB_ERR_ARG, "gmb_progress: NULL argument");
    *total = ix->prog_total.load();
    uint64_t d = 0;
    const uint64_t chunk = ix->prog_chunk.load();
    if (chunk && ix->s_progress && ix->d_counters) { // chunks the kernel in flight has handed out + finished pieces
        unsigned long long c[16] = {};
        if (cudaSetDevice(ix->device) == cudaSuccess &&
            cudaMemcpyAsync(c, ix->d_counters, sizeof(c), cudaMemcpyDeviceToHost, ix->s_progress) == cudaSuccess &&
            cudaStreamSynchronize(ix->s_progress) == cudaSuccess)
            d = (uint64_t)c[0] * chunk + (uint64_t)c[15];
        else
            cudaGetLastError();
    }
    *done = d < *total ? d : *total;
    return GMB_OK;
}

int gmb_index_set_jump_depth(gmb_index* ix, int depth)
{
    if (!ix || depth < -1 || depth > 16) return fail(GMB_ERR_ARG, "jump depth must be -1 (auto), 0 (off) or 1..16");
    ix->jump_depth_opt = depth;
    return GMB_OK;
}

} // extern "C"

// stats != nullptr && timed: bracket the kernel with events and wait for it; stats != nullptr && !timed: only
// fill the host-side fields (positions, jump depth) and return without synchronising
namespace {
int ep_many_files(gmb_index* ix, const gmb_params* p, uint64_t text_begin, uint64_t text_len, const uint64_t* chrom_cum,
                  uint32_t n_chrom, const uint64_t (*intervals)[2], uint64_t n_intervals, const uint32_t* seq_to_file,
                  uint32_t n_seq, uint64_t pos_begin, uint64_t pos_end, void* out_device, gmb_map_stats* stats);
bool dna5_nfree(const gmb_index* ix, const gmb_params* p);
int ensure_nfix(gmb_index* ix, const gmb_params* p, const NFix** out);
}

// One pass of the locate path (gmb_map_locations): counting (rows == nullptr: out_device receives two uint32 list
// lengths per position of [pos0, ...)) or filling (rows / off set).
struct LocPass {
    const uint64_t* off;
    uint32_t* rows;
    uint64_t pos0;
    const uint32_t* pos_list = nullptr; // position j stands for text position pos_list[j] (MapLaunch::loc_list)
};

// The search plan of a configuration: step tables, block size, how every search enters through the jump tables —
// built once per handle and configuration, with its tables left on the device.  (Round 1 rebuilt and re-uploaded all
// of it on every call, i.e. once per 32 Mi-position piece of the host-output pipeline.)
// nfree: Dna5 index, the searches skip the text's N (dna5_nfree below)
static int get_plan(gmb_index* ix, const gmb_params* p, bool sync_tables, bool loc, bool nfree, cudaStream_t stream, MapPlan** out)
{
    uint32_t want_b = p->block_kmers;
    if (want_b == 0) { const char* env = std::getenv("GMB_BLOCK_KMERS"); if (env && *env) want_b = (uint32_t)std::atoi(env); }
    const char* env_depth = std::getenv("GMB_JUMP_DEPTH");
    int want = ix->jump_depth_opt;
    if (want < 0 && env_depth && *env_depth) want = std::atoi(env_depth);
    const uint64_t model_n = ix->plan_n ? ix->plan_n : ix->h.n_bwt;
    const char* e1 = std::getenv("GMB_PART_MODEL");
    const char* e2 = std::getenv("GMB_PART_WEIGHTS");
    const char* e3 = std::getenv("GMB_JUMP_VARIANTS");
    char buf[200];
    std::snprintf(buf, sizeof buf, "%u/%u/%u/%d/%d/%d/%d/%llu/", p->K, p->E, want_b, sync_tables ? 1 : 0, loc ? 1 : 0, nfree ? 1 : 0, want,
                  (unsigned long long)model_n);
    const char* e4 = std::getenv("GMB_LOCATE");
    const char* e5 = std::getenv("GMB_BLOCK_KERNEL");
    const char* e6 = std::getenv("GMB_EXACT_KERNEL");
    const std::string key = std::string(buf) + (e1 ? e1 : "") + "/" + (e2 ? e2 : "") + "/" + (e3 ? e3 : "") + "/" + (e4 ? e4 : "") + "/" +
                            (e5 ? e5 : "") + "/" + (e6 ? e6 : "");
    MapPlan* plan = nullptr;
    for (auto& q : ix->plans)
        if (q->key == key) plan = q.get();
    std::string err;
    if (!plan) {
        std::unique_ptr<MapPlan> np(new (std::nothrow) MapPlan);
        if (!np) return fail(GMB_ERR_NOMEM, "out of host memory");
        np->key = key;
        BlockTables& tabs = np->tabs;
        if (!build_block_tables(p->K, p->E, want_b, sync_tables, tabs, err, model_n, block_bases(ix->h.sigma), nfree)) return fail(GMB_ERR_UNSUPPORTED, err);
        // --exclude-pseudo is asked for on indices of several near-identical genomes: every infix hit then stands for about
        // as many real occurrences as there are files, all of them completed window by window, and the planner's iid model
        // (chance hits only) picks blocks that are too large: 10 x 300 Mbp, K = 50, E = 2: 531 M positions/s with its
        // choice, 643 M with 4 k-mers per block, 197 M with 15 (profiles/r02/s19_pangenome_blocks.txt)
        if (p->exclude_pseudo && want_b == 0 && tabs.B > 4 &&
            !build_block_tables(p->K, p->E, 4, sync_tables, tabs, err, model_n, block_bases(ix->h.sigma), nfree)) return fail(GMB_ERR_UNSUPPORTED, err);
        // the tables and the per-chain frame store live in shared memory: shrink the block if they do not fit
        while (tabs.B > 1 && (map_kernel_smem_bytes((uint32_t)tabs.steps.size(), p->E, tabs.B, sync_tables, ix->h.sigma, true) > (200u << 10) ||
                              tabs.steps.size() * 4 + (tabs.B + 1) * kMaxSearches * sizeof(SearchStart) > kTableBytes))
            if (!build_block_tables(p->K, p->E, tabs.B - 1, sync_tables, tabs, err, model_n, block_bases(ix->h.sigma), nfree)) return fail(GMB_ERR_UNSUPPORTED, err);
        if (ix->plans.size() >= 12) { // drop the least recently used plan
            size_t lru = 0;
            for (size_t i = 1; i < ix->plans.size(); ++i) if (ix->plans[i]->last_use < ix->plans[lru]->last_use) lru = i;
            CU(cudaDeviceSynchronize()); // a kernel in flight may still read its tables
            ix->plans.erase(ix->plans.begin() + (long)lru);
        }
        plan = np.get();
        ix->plans.push_back(std::move(np));
    }
    plan->last_use = ++ix->plan_clock;
    *out = plan;
    if (plan->d_tables && plan->jt_epoch == ix->jt_epoch) return GMB_OK;

    // ---- (re)plan the jump-table entries and put the tables on the device -------------------------------------
    const BlockTables& tabs = plan->tabs;
    std::vector<JumpPlan> plans(tabs.B + 1);
    std::vector<SearchStart> starts((size_t)(tabs.B + 1) * kMaxSearches);
    std::vector<uint32_t> variants;
    uint32_t plan_depth = 0;
    uint32_t maxd = want < 0 ? default_jump_depth(model_n) : (uint32_t)want;
    if (maxd > 16) maxd = 16;
    JumpNeeds needs;
    const bool use_full = tabs.B > 1 && !loc; // the blocked instantiation
    const bool all_full = !loc && locate_enabled();
    for (;;) {
        plan_depth = 0;
        for (uint32_t cnt = 0; cnt <= tabs.B; ++cnt) {
            JumpPlan& pl = plans[cnt];
            if (cnt == 0) { pl = JumpPlan(); std::memset(pl.depth, 0, sizeof(pl.depth)); pl.max_depth = 0; continue; }
            plan_jump_tables(tabs.infix[cnt], maxd, pl, p->E, model_n, ix->h.sigma, cnt, tabs.B > 1 && !loc, nfree);
            plan_depth = std::max(plan_depth, pl.max_depth);
        }
        needs = jump_needs(plans, use_full, all_full);
        if (want >= 0 || maxd <= 1) break; // a fixed depth is taken as it is
        // automatic depth: what is missing must fit comfortably in free HBM, if need be after dropping the cached
        // tables this configuration does not use
        size_t free_b = 0, total_b = 0;
        const size_t missing = missing_jump_bytes(ix, needs);
        if (missing == 0) break;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); break; }
        if (missing <= free_b / 10 * 8) break;
        CU(cudaStreamSynchronize(stream));
        if (evict_stale_jump_tables(ix, needs)) {
            if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); break; }
            if (missing_jump_bytes(ix, needs) <= free_b / 10 * 8) break;
        }
        --maxd;
    }
    int rcj = ensure_jump_tables(ix, needs, stream);
    if (rcj != GMB_OK) return rcj;
    for (uint32_t cnt = 1; cnt <= tabs.B; ++cnt)
        for (uint32_t s2 = 0; s2 < kMaxSearches; ++s2) {
            const uint32_t d = plans[cnt].depth[s2];
            SearchStart& S = starts[(size_t)cnt * kMaxSearches + s2];
            std::memset(&S, 0, sizeof(S));
            const bool full = d && ((plans[cnt].need_lof[s2] && use_full) || all_full);
            S.uni = (d && !full) ? ix->jt_uni[d] : nullptr;
            S.lof = (d && !full && plans[cnt].need_lof[s2]) ? ix->jt_lof[d] : nullptr;
            S.full = full ? ix->jt_full[d] : nullptr;
            S.a = plans[cnt].a[s2];
            S.d = d;
            S.n_var = std::max(1u, plans[cnt].n_var[s2]);
        }
    // lay the variant sets of all tables out in one array behind the starts and turn the indices into device pointers
    std::vector<size_t> base_of(tabs.B + 1, 0);
    for (uint32_t cnt = 1; cnt <= tabs.B; ++cnt) {
        base_of[cnt] = variants.size();
        variants.insert(variants.end(), plans[cnt].variants.begin(), plans[cnt].variants.end());
        if (plans[cnt].variants.empty()) variants.push_back(0xffffffffu);
    }
    // block_kernel.cu: every key of every search of one strand as a flat list (the 3^m substitutions of a set of m offsets
    // spelled out as XOR masks on the key window), when every search of every block size enters through 16-byte entries
    KeyLists keylist;
    bool block_ok = p->E >= 1 && !loc && all_full && (ix->h.sigma == 4 || nfree) && p->K + tabs.B - 1 <= 64;
    {
        const char* env = std::getenv("GMB_BLOCK_KERNEL"); // "0": E >= 1 through the general kernel (A/B measurements)
        if (env && env[0] == '0') block_ok = false;
    }
    if (block_ok) block_ok = build_key_lists(tabs, plans, keylist);
    for (uint32_t cnt = 0; cnt <= kMaxBlockKmers; ++cnt) { plan->key_off[cnt] = block_ok ? keylist.off[cnt] : 0; plan->key_n[cnt] = block_ok ? keylist.n[cnt] : 0; }
    if (!block_ok) keylist.xy.clear();
    const size_t step_bytes = tabs.steps.size() * sizeof(uint32_t);
    const size_t start_off = (step_bytes + 15) / 16 * 16;
    const size_t var_off = start_off + starts.size() * sizeof(SearchStart);
    const size_t key_off_b = (var_off + variants.size() * sizeof(uint32_t) + 15) / 16 * 16;
    const size_t total = key_off_b + keylist.xy.size() * sizeof(uint32_t);
    CU(cudaStreamSynchronize(stream)); // nothing in flight reads the old tables while they are replaced
    if (plan->d_tables) { cudaFree(plan->d_tables); plan->d_tables = nullptr; }
    CU(cudaMalloc(&plan->d_tables, total));
    uint8_t* dt = reinterpret_cast<uint8_t*>(plan->d_tables);
    for (uint32_t cnt = 1; cnt <= tabs.B; ++cnt)
        for (uint32_t s2 = 0; s2 < kMaxSearches; ++s2) {
            SearchStart& S = starts[(size_t)cnt * kMaxSearches + s2];
            S.var = reinterpret_cast<const uint32_t*>(dt + var_off) + base_of[cnt] + plans[cnt].var_off[s2];
            S.set0 = plans[cnt].variants.empty() ? 0xffffffffu : variants[base_of[cnt] + plans[cnt].var_off[s2]];
        }
    CU(cudaMemcpyAsync(dt, tabs.steps.data(), step_bytes, cudaMemcpyHostToDevice, stream));
    CU(cudaMemcpyAsync(dt + start_off, starts.data(), starts.size() * sizeof(SearchStart), cudaMemcpyHostToDevice, stream));
    CU(cudaMemcpyAsync(dt + var_off, variants.data(), variants.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    if (!keylist.xy.empty()) CU(cudaMemcpyAsync(dt + key_off_b, keylist.xy.data(), keylist.xy.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    plan->d_keys = keylist.xy.empty() ? nullptr : reinterpret_cast<const uint2*>(dt + key_off_b);
    CU(cudaStreamSynchronize(stream)); // pageable sources: staged before the vectors go out of scope
    plan->start_off = start_off;
    plan->plan_depth = plan_depth;
    plan->e0_table = nullptr;
    plan->e0_depth = 0;
    {
        const char* env = std::getenv("GMB_EXACT_KERNEL"); // "0": E = 0 through the general kernel (A/B measurements)
        const SearchStart& S0 = starts[kMaxSearches];
        if (p->E == 0 && tabs.B == 1 && !loc && S0.full && !(env && env[0] == '0')) { plan->e0_table = S0.full; plan->e0_depth = S0.d; }
    }
    plan->jt_epoch = ix->jt_epoch;
    return GMB_OK;
}

static int map_device_impl(gmb_index* ix, const gmb_params* p_in, uint64_t text_begin, uint64_t text_len,
                           const uint64_t* chrom_cum, uint32_t n_chrom, const uint64_t (*intervals)[2],
                           uint64_t n_intervals, const uint32_t* seq_to_file, uint32_t n_seq,
                           uint64_t pos_begin, uint64_t pos_end, void* out_device, void* cuda_stream,
                           gmb_map_stats* stats, bool timed, const LocPass* loc = nullptr)
{
    if (!ix || !p_in || !chrom_cum || !out_device) return fail(GMB_ERR_ARG, "gmb_map_frequencies: NULL argument");
    gmb_params params = *p_in;
    if (loc) { // one k-mer per chain on tables that keep both intervals in step; no file reduction, no counters
        params.exclude_pseudo = 0; params.count_fetches = 0; params.block_kmers = 1; params.value_bits = 16;
        if (!ix->h.off_sa) return fail(GMB_ERR_UNSUPPORTED, "locations (csv output) need an index built with the full suffix array (GMB_BUILD_WITH_SA)");
    }
    const gmb_params* p = &params;
    const bool sync_tables = p->exclude_pseudo != 0 || loc != nullptr;
    if (p->value_bits != 8 && p->value_bits != 16) return fail(GMB_ERR_ARG, "value_bits must be 8 or 16");
    if (text_begin + text_len > ix->h.n_text) return fail(GMB_ERR_ARG, "text range exceeds the indexed text");
    if (n_chrom == 0 || chrom_cum[0] != 0 || chrom_cum[n_chrom] != text_len) return fail(GMB_ERR_ARG, "chrom_cum_lengths must start at 0 and end at text_len");
    uint32_t n_files = 0, own_file = 0;
    if (p->exclude_pseudo) {
        if (!ix->h.off_sa) return fail(GMB_ERR_UNSUPPORTED, "--exclude-pseudo needs an index built with the full suffix array (`genmap index` without --no-sa / GMB_BUILD_WITH_SA)");
        if (!seq_to_file || n_seq != ix->h.n_seq) return fail(GMB_ERR_ARG, "--exclude-pseudo needs seq_to_file for every indexed sequence");
        for (uint32_t s = 0; s < n_seq; ++s) n_files = std::max(n_files, seq_to_file[s] + 1);
        if (n_files > 64) { // more files than the kernel's 64-bit file mask holds: locate + distinct-file count
            if (cuda_stream) return fail(GMB_ERR_UNSUPPORTED, "--exclude-pseudo with more than 64 FASTA files runs on the default stream only");
            return ep_many_files(ix, p, text_begin, text_len, chrom_cum, n_chrom, intervals, n_intervals, seq_to_file, n_seq,
                                 pos_begin, pos_end, out_device, stats);
        }
        if (p->count_fetches) return fail(GMB_ERR_UNSUPPORTED, "count_fetches is not available with --exclude-pseudo");
        uint32_t s0 = 0; // the sequence containing text_begin tells which file is being mapped
        while (s0 + 1 < n_seq && ix->limits[s0 + 1] <= text_begin) ++s0;
        own_file = seq_to_file[s0];
    }
    if (n_intervals && !intervals) return fail(GMB_ERR_ARG, "intervals is NULL");
    CU(cudaSetDevice(ix->device));
    cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
    // Dna5: with the suffix array in HBM the searches of an E >= 1 call skip the text's N and enter through substituted
    // keys like on a Dna4 index; the alignments to the text windows with N come from the N pass (NFix)
    const NFix* nfix = nullptr;
    if (!loc && dna5_nfree(ix, p)) {
        int rcn = ensure_nfix(ix, p, &nfix);
        if (rcn != GMB_OK) return rcn;
    }
    const bool nfree = nfix != nullptr;
    MapPlan* plan = nullptr;
    {
        int rcp = get_plan(ix, p, sync_tables, loc != nullptr, nfree, stream, &plan);
        if (rcp != GMB_OK) return rcp;
    }
    const BlockTables& tabs = plan->tabs;

    std::vector<WorkRange> ranges;
    build_work_ranges(text_len, p->K, chrom_cum, n_chrom, reinterpret_cast<const uint64_t*>(intervals), n_intervals,
                      pos_begin, pos_end, ranges);
    const uint32_t nr = (uint32_t)ranges.size();
    // whole blocks per work chunk; the two-phase kernel hands a warp one block per lane
    const char* blk_env = std::getenv("GMB_BLOCK_KERNEL"); // "0": never, "2": also for E >= 3 (measurements)
    const bool force_block = blk_env && blk_env[0] == '2';
    const bool block_kernel = plan->d_keys != nullptr && (p->E != 3 || force_block) &&
                              block_kernel_smem_bytes((uint32_t)tabs.steps.size(), p->E, tabs.B, p->exclude_pseudo != 0, ix->h.sigma) <= (200u << 10);
    const uint64_t chunk = block_kernel ? 32ull * tabs.B : std::max<uint64_t>(tabs.B, kChunk / tabs.B * tabs.B);
    std::vector<uint64_t> host_ranges(3 * (size_t)nr + 1); // begin[nr], end[nr], chunk_prefix[nr+1]
    uint64_t total = 0, chunks = 0;
    for (uint32_t i = 0; i < nr; ++i) {
        host_ranges[i] = ranges[i].begin;
        host_ranges[nr + i] = ranges[i].end;
        host_ranges[2 * (size_t)nr + i] = chunks;
        total += ranges[i].end - ranges[i].begin;
        chunks += (ranges[i].end - ranges[i].begin + chunk - 1) / chunk;
    }
    host_ranges[3 * (size_t)nr] = chunks;
    if (stats) { std::memset(stats, 0, sizeof(*stats)); stats->positions = total; }
    if (total == 0) return GMB_OK;

    if (ix->ranges_cap < host_ranges.size()) {
        if (ix->d_ranges) cudaFree(ix->d_ranges);
        ix->d_ranges = nullptr;
        ix->ranges_cap = host_ranges.size() * 2;
        CU(cudaMalloc(&ix->d_ranges, ix->ranges_cap * 8));
    }
    CU(cudaMemcpyAsync(ix->d_ranges, host_ranges.data(), host_ranges.size() * 8, cudaMemcpyHostToDevice, stream));
    CU(cudaMemsetAsync(ix->d_counters, 0, (ix->prog_in_pipeline ? 15 : 16) * sizeof(unsigned long long), stream)); // [15]: see gmb_progress
    ix->prog_chunk.store(chunk);
    if (!ix->prog_in_pipeline) ix->prog_total.store(total);

    MapLaunch L;
    const uint8_t* base = ix->d_blob;
    fill_ctx(ix, L.cx);
    const uint32_t plan_depth = plan->plan_depth;
    L.cx.steps = plan->d_tables;
    L.cx.starts = reinterpret_cast<const SearchStart*>(reinterpret_cast<const uint8_t*>(plan->d_tables) + plan->start_off);
    L.cx.p1_off = nullptr; L.cx.fl_off = nullptr; // the kernel points them at its shared-memory copies
    L.n_step_words = (uint32_t)tabs.steps.size();
    for (uint32_t c2 = 0; c2 <= kMaxBlockKmers; ++c2) { L.p1_off[c2] = tabs.p1_off[c2]; L.fl_off[c2] = tabs.fl_off[c2]; }
    L.chunk = (uint32_t)chunk;
    L.cx.K = p->K;
    L.cx.B = tabs.B;
    L.cx.E = p->E;
    L.cx.n_search = tabs.n_search;
    L.cx.n_strands = p->revcompl ? 2u : 1u;
    L.cx.maxv = p->value_bits == 16 ? 65535u : 255u;
    L.cx.skip_n = nfree ? 1u : 0u;
    L.E = p->E;
    L.sigma = ix->h.sigma;
    L.text = reinterpret_cast<const uint64_t*>(base + ix->h.off_text);
    L.nmask = ix->h.sigma == 5 ? reinterpret_cast<const uint64_t*>(base + ix->h.off_nmask) : nullptr;
    L.text_begin = text_begin;
    L.range_begin = ix->d_ranges;
    L.range_end = ix->d_ranges + nr;
    L.chunk_prefix = ix->d_ranges + 2 * (size_t)nr;
    L.n_ranges = nr;
    L.n_chunks = chunks;
    L.n_work = total;
    L.work_counter = ix->d_counters;
    L.fetch_counter = ix->d_counters + 1;
    L.out = out_device;
    L.value_bits = p->value_bits;
    L.count_fetches = p->count_fetches != 0;
    L.exclude_pseudo = p->exclude_pseudo != 0;
    if (L.exclude_pseudo) {
        if (!ix->d_seq_to_file) CU(cudaMalloc(&ix->d_seq_to_file, (size_t)ix->h.n_seq * 4));
        CU(cudaMemcpyAsync(ix->d_seq_to_file, seq_to_file, (size_t)n_seq * 4, cudaMemcpyHostToDevice, stream));
        L.cx.seq_to_file = ix->d_seq_to_file;
        L.cx.own_file = own_file;
        L.cx.all_files = n_files == 64 ? ~0ull : ((1ull << n_files) - 1ull);
    }

    L.e0_table = plan->e0_table;
    L.e0_depth = plan->e0_depth;
    L.keylist = block_kernel ? plan->d_keys : nullptr;
    L.force_block_kernel = force_block;
    for (uint32_t c2 = 0; c2 <= kMaxBlockKmers; ++c2) { L.key_off[c2] = plan->key_off[c2]; L.key_n[c2] = plan->key_n[c2]; }
    L.cx.loc_rows = loc ? loc->rows : nullptr;
    L.loc_off = loc ? loc->off : nullptr;
    L.loc_pos0 = loc ? loc->pos0 : 0;
    L.loc_list = loc ? loc->pos_list : nullptr;

    if (stats) { stats->jump_depth = plan_depth; stats->kernel_launches = 1; stats->block_kmers = tabs.B; }
    if (stats && timed) CU(cudaEventRecord(ix->ev0, stream));
    CU(loc ? launch_locate_kernel(L, ix->sm_count, stream) : launch_map_kernel(L, ix->sm_count, stream));
    if (nfix) { // the N pass: whole counts of the windows with N, then + 1 per alignment to one of them
        CU(nfix_apply(nfix->d_win_pos, nfix->d_win_count, nfix->n_win, text_begin, L.range_begin, L.range_end, nr, out_device, p->value_bits, stream));
        CU(nfix_apply(nfix->d_hits, nullptr, nfix->n_hits, text_begin, L.range_begin, L.range_end, nr, out_device, p->value_bits, stream));
        if (stats) stats->kernel_launches = 3;
    }
    if (stats && timed) {
        CU(cudaEventRecord(ix->ev1, stream));
        CU(cudaEventSynchronize(ix->ev1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, ix->ev0, ix->ev1));
        stats->kernel_ms = ms;
        stats->kernel_launches = nfix ? 3 : 1;
        stats->jump_depth = plan_depth;
        stats->block_kmers = tabs.B;
        if (L.count_fetches) {
            unsigned long long f[14] = {};
            CU(cudaMemcpy(f, ix->d_counters + 1, sizeof(f), cudaMemcpyDeviceToHost));
            stats->rank_block_fetches = f[0];
            stats->jump_table_reads = f[1];
            for (int k = 0; k < 8; ++k) stats->fetches_by_size[k] = f[2 + k];
            stats->thin_paths = f[10];
            stats->iterations = f[11];
            stats->located_entries = f[12];
            stats->text_reads = f[13];
        }
    }
    return GMB_OK;
}

extern "C" {

int gmb_map_frequencies_device(gmb_index* ix, const gmb_params* p, uint64_t text_begin, uint64_t text_len,
                               const uint64_t* chrom_cum, uint32_t n_chrom, const uint64_t (*intervals)[2],
                               uint64_t n_intervals, const uint32_t* seq_to_file, uint32_t n_seq,
                               uint64_t pos_begin, uint64_t pos_end, void* out_device, void* cuda_stream,
                               gmb_map_stats* stats)
{
    return map_device_impl(ix, p, text_begin, text_len, chrom_cum, n_chrom, intervals, n_intervals, seq_to_file, n_seq,
                           pos_begin, pos_end, out_device, cuda_stream, stats, true);
}

int gmb_map_frequencies_range(gmb_index* ix, const gmb_params* p, uint64_t text_begin, uint64_t text_len,
                              const uint64_t* chrom_cum, uint32_t n_chrom, const uint64_t (*intervals)[2],
                              uint64_t n_intervals, const uint32_t* seq_to_file, uint32_t n_seq,
                              uint64_t pos_begin, uint64_t pos_end, void* out, gmb_map_stats* stats)
{
    if (!ix || !p || !out) return fail(GMB_ERR_ARG, "gmb_map_frequencies: NULL argument");
    if (p->value_bits != 8 && p->value_bits != 16) return fail(GMB_ERR_ARG, "value_bits must be 8 or 16");
    if (pos_end > text_len) pos_end = text_len;
    if (pos_begin > pos_end) return fail(GMB_ERR_ARG, "pos_begin > pos_end");
    CU(cudaSetDevice(ix->device));
    const size_t elem = p->value_bits / 8;
    const size_t bytes = (size_t)(pos_end - pos_begin) * elem;
    if (ix->out_cap < bytes) {
        if (ix->d_out) cudaFree(ix->d_out);
        ix->d_out = nullptr;
        ix->out_cap = 0;
        CU(cudaMalloc(&ix->d_out, bytes ? bytes : 1));
        ix->out_cap = bytes;
    }
    // the kernel indexes its output by file-local position: bias the base so the slice starts at pos_begin
    void* biased = reinterpret_cast<void*>(reinterpret_cast<uintptr_t>(ix->d_out) - (uintptr_t)pos_begin * elem);
    gmb_map_stats local;
    std::memset(&local, 0, sizeof(local));
    // host-output pipeline: pieces that halve towards the end — the copy of the LAST piece is the only one no search
    // overlaps, and every piece costs a launch: half of what is left (at most 64 Mi, at least 2 Mi positions) per piece
    // (ranges of up to 32 Mi positions are not worth a pipeline: one launch, one copy)
    const uint64_t piece = 32ull << 20, kMaxPiece = 64ull << 20, kMinPiece = 2ull << 20;
    bool many_files = false; // --exclude-pseudo beyond 64 files runs unpipelined on the default stream (ep_many_files)
    if (p->exclude_pseudo && seq_to_file)
        for (uint32_t s2 = 0; s2 < n_seq && !many_files; ++s2) many_files = seq_to_file[s2] >= 64;
    if (p->count_fetches || many_files || pos_end - pos_begin <= piece) {
        CU(cudaMemsetAsync(ix->d_out, 0, bytes, nullptr));
        int rc = map_device_impl(ix, p, text_begin, text_len, chrom_cum, n_chrom, intervals, n_intervals, seq_to_file, n_seq,
                                 pos_begin, pos_end, biased, nullptr, &local, true);
        if (rc != GMB_OK) return rc;
        CU(cudaMemcpy(out, ix->d_out, bytes, cudaMemcpyDeviceToHost));
    } else {
        // pipeline: the search of piece i+1 (compute stream) overlaps the device-to-host copy of piece i (copy stream)
        if (!ix->s_compute) {
            CU(cudaStreamCreateWithFlags(&ix->s_compute, cudaStreamNonBlocking));
            CU(cudaStreamCreateWithFlags(&ix->s_copy, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&ix->ev_piece[0], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&ix->ev_piece[1], cudaEventDisableTiming));
        }
        CU(cudaMemsetAsync(ix->d_out, 0, bytes, ix->s_compute));
        CU(cudaMemsetAsync(ix->d_counters + 15, 0, sizeof(unsigned long long), ix->s_compute));
        CU(cudaEventRecord(ix->ev0, ix->s_compute));
        ix->prog_total.store(pos_end - pos_begin);
        struct PipelineFlag { bool& f; ~PipelineFlag() { f = false; } } in_pipeline{ix->prog_in_pipeline};
        ix->prog_in_pipeline = true;
        uint32_t n_piece = 0;
        for (uint64_t b = pos_begin, e = pos_begin; b < pos_end; b = e, ++n_piece) {
            const uint64_t left = pos_end - b;
            e = b + (left <= kMinPiece ? left : std::min(kMaxPiece, std::max(kMinPiece, left / 2)));
            gmb_map_stats st;
            int rc = map_device_impl(ix, p, text_begin, text_len, chrom_cum, n_chrom, intervals, n_intervals, seq_to_file,
                                     n_seq, b, e, biased, ix->s_compute, &st, false);
            if (rc != GMB_OK) { cudaStreamSynchronize(ix->s_compute); cudaStreamSynchronize(ix->s_copy); return rc; }
            local.positions += st.positions;
            local.kernel_launches += st.kernel_launches;
            local.jump_depth = st.jump_depth;
            local.block_kmers = st.block_kmers;
            const unsigned long long done_so_far = local.positions; // staged at once (pageable source)
            CU(cudaMemcpyAsync(ix->d_counters + 15, &done_so_far, sizeof(done_so_far), cudaMemcpyHostToDevice, ix->s_compute));
            CU(cudaEventRecord(ix->ev_piece[n_piece & 1], ix->s_compute));
            CU(cudaStreamWaitEvent(ix->s_copy, ix->ev_piece[n_piece & 1], 0));
            CU(cudaMemcpyAsync(static_cast<uint8_t*>(out) + (b - pos_begin) * elem,
                               static_cast<const uint8_t*>(ix->d_out) + (b - pos_begin) * elem, (e - b) * elem,
                               cudaMemcpyDeviceToHost, ix->s_copy));
        }
        CU(cudaEventRecord(ix->ev1, ix->s_compute));
        CU(cudaStreamSynchronize(ix->s_compute));
        CU(cudaStreamSynchronize(ix->s_copy));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, ix->ev0, ix->ev1));
        local.kernel_ms = ms;
        ix->prog_total.store(local.positions);
    }
    if (stats) *stats = local;
    return GMB_OK;
}

int gmb_map_frequencies(gmb_index* ix, const gmb_params* p, uint64_t text_begin, uint64_t text_len,
                        const uint64_t* chrom_cum, uint32_t n_chrom, const uint64_t (*intervals)[2],
                        uint64_t n_intervals, const uint32_t* seq_to_file, uint32_t n_seq, void* out,
                        gmb_map_stats* stats)
{
    return gmb_map_frequencies_range(ix, p, text_begin, text_len, chrom_cum, n_chrom, intervals, n_intervals,
                                     seq_to_file, n_seq, 0, text_len, out, stats);
}

} // extern "C"

/* ---- locations (csv output) ------------------------------------------------------------------------------- */
namespace {
struct DevBuf { // frees on scope exit
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <class T> T* as() const { return static_cast<T*>(p); }
};
} // namespace

extern "C" {

void gmb_locations_free(gmb_locations* L)
{
    if (!L) return;
    std::free(L->offsets);
    std::free(L->loc);
    L->offsets = nullptr; L->loc = nullptr; L->n_locations = 0;
}

} // extern "C"

namespace {
// The sorted occurrence lists of the positions [pos_begin, pos_begin + m) on the device (the two search passes,
// the scan and the segmented sort of locate_kernel.cu); list offsets also on the host.
struct LocatePiece {
    DevBuf counts, offs, rows, sorted;
    uint64_t m = 0, n_rows = 0;
    uint64_t* h_off = nullptr; // 2 * npos + 1 offsets (only the first 2 * m + 1 belong to the piece); malloc'ed
    double kernel_ms = 0;
    ~LocatePiece() { std::free(h_off); }
};

int locate_piece(gmb_index* ix, const gmb_params* p, uint64_t text_begin, uint64_t text_len, const uint64_t* chrom_cum,
                 uint32_t n_chrom, const uint64_t (*intervals)[2], uint64_t n_intervals, uint64_t pos_begin,
                 uint64_t pos_end, uint64_t max_locations, LocatePiece& P, const uint32_t* pos_list = nullptr)
{
    const uint64_t kMaxPositions = 4ull << 20, kMaxRows = 1ull << 30;
    if (pos_end - pos_begin > kMaxPositions) pos_end = pos_begin + kMaxPositions;
    if (max_locations == 0) max_locations = 64ull << 20;
    if (max_locations > kMaxRows) max_locations = kMaxRows;
    CU(cudaSetDevice(ix->device));
    const uint64_t npos = pos_end - pos_begin, n_lists = 2 * npos;

    // pass 1: list lengths, then their prefix sums
    DevBuf temp;
    CU(P.counts.alloc((n_lists + 1) * 4));
    CU(P.offs.alloc((n_lists + 1) * 8));
    CU(cudaMemsetAsync(P.counts.p, 0, (n_lists + 1) * 4, nullptr));
    gmb_map_stats st1, st2;
    std::memset(&st2, 0, sizeof(st2));
    LocPass pass1{nullptr, nullptr, pos_begin, pos_list};
    int rc = map_device_impl(ix, p, text_begin, text_len, chrom_cum, n_chrom, intervals, n_intervals, nullptr, 0, pos_begin,
                             pos_end, P.counts.p, nullptr, &st1, true, &pass1);
    if (rc != GMB_OK) return rc;
    size_t temp_bytes = 0;
    CU(locate_scan_counts(P.counts.as<uint32_t>(), n_lists, P.offs.as<uint64_t>(), nullptr, temp_bytes, nullptr));
    CU(temp.alloc(temp_bytes));
    CU(locate_scan_counts(P.counts.as<uint32_t>(), n_lists, P.offs.as<uint64_t>(), temp.p, temp_bytes, nullptr));
    P.h_off = static_cast<uint64_t*>(std::malloc((n_lists + 1) * 8));
    if (!P.h_off) return fail(GMB_ERR_NOMEM, "out of host memory");
    CU(cudaMemcpy(P.h_off, P.offs.p, (n_lists + 1) * 8, cudaMemcpyDeviceToHost));

    // keep as many whole positions as fit into max_locations (always at least one)
    uint64_t m = npos;
    if (P.h_off[n_lists] > max_locations) {
        uint64_t lo = 1, hi = npos; // largest m with h_off[2m] <= max_locations, at least 1
        while (lo < hi) {
            const uint64_t mid = (lo + hi + 1) / 2;
            if (P.h_off[2 * mid] <= max_locations) lo = mid; else hi = mid - 1;
        }
        m = lo;
    }
    P.m = m;
    P.n_rows = P.h_off[2 * m];
    if (P.n_rows > (1ull << 31)) return fail(GMB_ERR_UNSUPPORTED, "a single k-mer has more than 2^31 occurrences");
    if (P.n_rows) {
        // pass 2: the same search writes the SA value of every occurrence; then every list is sorted
        DevBuf temp2;
        CU(P.rows.alloc(P.n_rows * 4));
        CU(P.sorted.alloc(P.n_rows * 4));
        LocPass pass2{P.offs.as<uint64_t>(), P.rows.as<uint32_t>(), pos_begin, pos_list};
        rc = map_device_impl(ix, p, text_begin, text_len, chrom_cum, n_chrom, intervals, n_intervals, nullptr, 0, pos_begin,
                             pos_begin + m, P.counts.p, nullptr, &st2, true, &pass2);
        if (rc != GMB_OK) return rc;
        size_t tb = 0;
        CU(locate_sort_lists(P.rows.as<uint32_t>(), P.sorted.as<uint32_t>(), P.n_rows, P.offs.as<uint64_t>(), 2 * m, nullptr, tb, nullptr));
        CU(temp2.alloc(tb));
        CU(locate_sort_lists(P.rows.as<uint32_t>(), P.sorted.as<uint32_t>(), P.n_rows, P.offs.as<uint64_t>(), 2 * m, temp2.p, tb, nullptr));
        CU(cudaStreamSynchronize(nullptr)); // temp2 is released on return
    }
    P.kernel_ms = st1.kernel_ms + st2.kernel_ms;
    return GMB_OK;
}

// --exclude-pseudo with more FASTA files than the 64-bit file mask of the search kernel holds: locate every
// occurrence (locate_piece) and count the distinct files of each k-mer's two sorted lists on the device.
int ep_many_files(gmb_index* ix, const gmb_params* p, uint64_t text_begin, uint64_t text_len, const uint64_t* chrom_cum,
                  uint32_t n_chrom, const uint64_t (*intervals)[2], uint64_t n_intervals, const uint32_t* seq_to_file,
                  uint32_t n_seq, uint64_t pos_begin, uint64_t pos_end, void* out_device, gmb_map_stats* stats)
{
    for (uint32_t s = 1; s < n_seq; ++s)
        if (seq_to_file[s] < seq_to_file[s - 1])
            return fail(GMB_ERR_UNSUPPORTED, "--exclude-pseudo with more than 64 FASTA files needs file ids that do not decrease along the sequences");
    CU(cudaSetDevice(ix->device));
    if (!ix->d_seq_to_file) CU(cudaMalloc(&ix->d_seq_to_file, (size_t)ix->h.n_seq * 4));
    CU(cudaMemcpy(ix->d_seq_to_file, seq_to_file, (size_t)n_seq * 4, cudaMemcpyHostToDevice));
    if (pos_end > text_len) pos_end = text_len;
    if (stats) std::memset(stats, 0, sizeof(*stats));
    gmb_params q = *p;
    q.exclude_pseudo = 0;
    for (uint64_t b = pos_begin; b < pos_end;) {
        LocatePiece P;
        int rc = locate_piece(ix, &q, text_begin, text_len, chrom_cum, n_chrom, intervals, n_intervals, b, pos_end, 0, P);
        if (rc != GMB_OK) return rc;
        CU(locate_distinct_files(P.sorted.as<uint32_t>(), P.offs.as<uint64_t>(), P.m, reinterpret_cast<const uint32_t*>(ix->d_blob + ix->h.off_seq_start),
                                 ix->h.n_seq, ix->d_seq_to_file, out_device, p->value_bits, b, nullptr));
        CU(cudaStreamSynchronize(nullptr));
        if (stats) { stats->kernel_ms += P.kernel_ms; stats->kernel_launches += 3; }
        b += P.m;
    }
    if (stats) stats->positions = pos_end - pos_begin;
    return GMB_OK;
}

// Does this call run with searches that skip the text's N (plus the N pass)?  Dna5 index, E >= 1, plain counts, the
// suffix array in HBM (the N pass locates).  GMB_DNA5_NFREE=0: never (A/B measurements).
bool dna5_nfree(const gmb_index* ix, const gmb_params* p)
{
    const char* env = std::getenv("GMB_DNA5_NFREE");
    return ix->h.sigma == 5 && p->E >= 1 && !p->exclude_pseudo && ix->h.off_sa != 0 && !ix->nfix_off && !(env && env[0] == '0');
}

// The N pass of (K, E, strands), built on first use: *out stays nullptr when the index has too many windows with N
// for it to pay (the call then runs as before, N children walked).
int ensure_nfix(gmb_index* ix, const gmb_params* p, const NFix** out)
{
    *out = nullptr;
    const bool rc = p->revcompl != 0;
    for (auto& q : ix->nfix)
        if (q->K == p->K && q->E == p->E && q->revcompl == rc) { *out = q.get(); return GMB_OK; }
    CU(cudaSetDevice(ix->device));
    CU(cudaDeviceSynchronize()); // the locate passes below use the handle's scratch on the default stream
    const auto t0 = std::chrono::steady_clock::now(); // (every step below ends in a synchronous copy)
    std::unique_ptr<NFix> nf(new (std::nothrow) NFix);
    if (!nf) return fail(GMB_ERR_NOMEM, "out of host memory");
    nf->K = p->K; nf->E = p->E; nf->revcompl = rc;
    const uint64_t n_text = ix->h.n_text;
    // 1. the windows: at most 1/64 of the text (an assembly has a few hundred gaps; every gap edge gives 2 E windows)
    const uint64_t cap = std::max<uint64_t>(1024, std::min<uint64_t>(n_text / 64, 32ull << 20));
    DevBuf win, counter;
    CU(win.alloc(cap * 4));
    CU(counter.alloc(8));
    CU(cudaMemsetAsync(counter.p, 0, 8, nullptr));
    const uint8_t* base = ix->d_blob;
    const uint64_t* nmask = reinterpret_cast<const uint64_t*>(base + ix->h.off_nmask);
    const uint32_t* seq_start = reinterpret_cast<const uint32_t*>(base + ix->h.off_seq_start);
    CU(nfix_collect_windows(nmask, n_text, seq_start, ix->h.n_seq, p->K, p->E, win.as<uint32_t>(), counter.as<unsigned long long>(), cap, nullptr));
    unsigned long long n_win = 0;
    CU(cudaMemcpy(&n_win, counter.p, 8, cudaMemcpyDeviceToHost));
    if (n_win > cap || n_win + p->K - 1 > n_text) { ix->nfix_off = true; return GMB_OK; }
    nf->n_win = n_win;
    std::vector<uint32_t> hits_host;
    if (n_win) {
        CU(cudaMalloc(&nf->d_win_pos, n_win * 4));
        CU(cudaMemcpy(nf->d_win_pos, win.p, n_win * 4, cudaMemcpyDeviceToDevice));
        CU(cudaMalloc(&nf->d_win_count, n_win * 4));
        // 2. locate them like csv queries: "positions" are list indices (LocPass::pos_list), one pseudo chromosome
        gmb_params q = *p;
        q.exclude_pseudo = 0; q.count_fetches = 0;
        const uint64_t list_len = n_win + p->K - 1, cum[2] = {0, list_len};
        const uint64_t kMaxHits = 256ull << 20;
        for (uint64_t b = 0; b < n_win;) {
            LocatePiece P;
            int rcl = locate_piece(ix, &q, 0, list_len, cum, 1, nullptr, 0, b, n_win, 0, P, nf->d_win_pos);
            if (rcl != GMB_OK) return rcl;
            DevBuf hits;
            CU(hits.alloc(P.n_rows * 4));
            CU(cudaMemsetAsync(counter.p, 0, 8, nullptr));
            CU(nfix_collect_hits(P.sorted.as<uint32_t>(), P.offs.as<uint64_t>(), P.m, P.n_rows, seq_start, ix->h.n_seq, nmask, p->K,
                                 nf->d_win_count + b, hits.as<uint32_t>(), counter.as<unsigned long long>(), nullptr));
            unsigned long long n_h = 0;
            CU(cudaMemcpy(&n_h, counter.p, 8, cudaMemcpyDeviceToHost));
            if (hits_host.size() + n_h > kMaxHits) { ix->nfix_off = true; return GMB_OK; }
            const size_t at = hits_host.size();
            hits_host.resize(at + n_h);
            if (n_h) CU(cudaMemcpy(hits_host.data() + at, hits.p, n_h * 4, cudaMemcpyDeviceToHost));
            b += P.m;
        }
    }
    nf->n_hits = hits_host.size();
    if (nf->n_hits) {
        CU(cudaMalloc(&nf->d_hits, nf->n_hits * 4));
        CU(cudaMemcpy(nf->d_hits, hits_host.data(), nf->n_hits * 4, cudaMemcpyHostToDevice));
    }
    CU(cudaDeviceSynchronize());
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    nf->build_ms = ms;
    if (const char* v = std::getenv("GMB_VERBOSE"); v && v[0] == '1')
        std::fprintf(stderr, "[gmb] N pass of (K=%u, E=%u): %llu windows with N, %llu alignments to them, built in %.1f ms\n", p->K, p->E,
                     (unsigned long long)nf->n_win, (unsigned long long)nf->n_hits, ms);
    *out = nf.get();
    ix->nfix.push_back(std::move(nf));
    return GMB_OK;
}
} // namespace

extern "C" {

int gmb_map_locations(gmb_index* ix, const gmb_params* p, uint64_t text_begin, uint64_t text_len,
                      const uint64_t* chrom_cum, uint32_t n_chrom, const uint64_t (*intervals)[2],
                      uint64_t n_intervals, uint64_t pos_begin, uint64_t pos_end, uint64_t max_locations,
                      gmb_locations* out)
{
    if (!ix || !p || !out || !chrom_cum) return fail(GMB_ERR_ARG, "gmb_map_locations: NULL argument");
    std::memset(out, 0, sizeof(*out));
    if (pos_end > text_len) pos_end = text_len;
    if (pos_begin >= pos_end) return fail(GMB_ERR_ARG, "gmb_map_locations: empty position range");
    LocatePiece P;
    int rc = locate_piece(ix, p, text_begin, text_len, chrom_cum, n_chrom, intervals, n_intervals, pos_begin, pos_end, max_locations, P);
    if (rc != GMB_OK) return rc;
    gmb_location* h_loc = static_cast<gmb_location*>(std::malloc(P.n_rows ? P.n_rows * sizeof(gmb_location) : 1));
    if (!h_loc) return fail(GMB_ERR_NOMEM, "out of host memory");
    if (P.n_rows) { // positions inside T -> (sequence, offset)
        DevBuf loc;
        cudaError_t e = loc.alloc(P.n_rows * sizeof(gmb_location));
        if (e == cudaSuccess) e = locate_convert(P.sorted.as<uint32_t>(), P.n_rows, reinterpret_cast<const uint32_t*>(ix->d_blob + ix->h.off_seq_start),
                                                 ix->h.n_seq, loc.p, nullptr);
        if (e == cudaSuccess) e = cudaMemcpy(h_loc, loc.p, P.n_rows * sizeof(gmb_location), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { std::free(h_loc); return cuda_fail(e, "locations"); }
    }
    out->pos_begin = pos_begin;
    out->pos_end = pos_begin + P.m;
    out->n_locations = P.n_rows;
    out->offsets = P.h_off;
    P.h_off = nullptr; // ownership moves to the caller
    out->loc = h_loc;
    out->kernel_ms = P.kernel_ms;
    return GMB_OK;
}

/* ---- runs (track writers) ------------------------------------------------------------------------------------ */
void gmb_runs_free(gmb_runs* R)
{
    if (!R) return;
    std::free(R->start);
    std::free(R->value);
    R->start = nullptr; R->value = nullptr; R->n_runs = 0;
}

int gmb_map_runs(gmb_index* ix, const gmb_params* p, uint64_t text_begin, uint64_t text_len, const uint64_t* chrom_cum,
                 uint32_t n_chrom, const uint64_t (*intervals)[2], uint64_t n_intervals, const uint32_t* seq_to_file,
                 uint32_t n_seq, uint64_t pos_begin, uint64_t pos_end, gmb_runs* out, gmb_map_stats* stats)
{
    if (!ix || !p || !out || !chrom_cum) return fail(GMB_ERR_ARG, "gmb_map_runs: NULL argument");
    std::memset(out, 0, sizeof(*out));
    if (p->value_bits != 8 && p->value_bits != 16) return fail(GMB_ERR_ARG, "value_bits must be 8 or 16");
    if (pos_end > text_len) pos_end = text_len;
    if (pos_begin > pos_end) return fail(GMB_ERR_ARG, "pos_begin > pos_end");
    out->pos_begin = pos_begin; out->pos_end = pos_end;
    if (pos_begin == pos_end) return GMB_OK;
    CU(cudaSetDevice(ix->device));
    const size_t elem = p->value_bits / 8;
    const size_t bytes = (size_t)(pos_end - pos_begin) * elem;
    if (ix->out_cap < bytes) {
        if (ix->d_out) cudaFree(ix->d_out);
        ix->d_out = nullptr;
        ix->out_cap = 0;
        CU(cudaMalloc(&ix->d_out, bytes));
        ix->out_cap = bytes;
    }
    void* biased = reinterpret_cast<void*>(reinterpret_cast<uintptr_t>(ix->d_out) - (uintptr_t)pos_begin * elem);
    CU(cudaMemsetAsync(ix->d_out, 0, bytes, nullptr));
    gmb_map_stats local;
    int rc = map_device_impl(ix, p, text_begin, text_len, chrom_cum, n_chrom, intervals, n_intervals, seq_to_file, n_seq,
                             pos_begin, pos_end, biased, nullptr, &local, true);
    if (rc != GMB_OK) return rc;
    if (stats) *stats = local;
    out->kernel_ms = local.kernel_ms;

    DevBuf cum, temp, starts, values;
    CU(cum.alloc(((size_t)n_chrom + 1) * 8));
    CU(cudaMemcpyAsync(cum.p, chrom_cum, ((size_t)n_chrom + 1) * 8, cudaMemcpyHostToDevice, nullptr));
    unsigned long long* d_count = ix->d_counters + 16;
    CU(cudaEventRecord(ix->ev0, nullptr));
    size_t tb1 = 0, tb2 = 0;
    CU(rle_count(biased, p->value_bits, cum.as<uint64_t>(), n_chrom, pos_begin, pos_end, d_count, nullptr, tb1, nullptr));
    CU(rle_select(biased, p->value_bits, cum.as<uint64_t>(), n_chrom, pos_begin, pos_end, nullptr, d_count, nullptr, tb2, nullptr));
    CU(temp.alloc(std::max(tb1, tb2)));
    CU(rle_count(biased, p->value_bits, cum.as<uint64_t>(), n_chrom, pos_begin, pos_end, d_count, temp.p, tb1, nullptr));
    unsigned long long n_runs = 0;
    CU(cudaMemcpy(&n_runs, d_count, sizeof(n_runs), cudaMemcpyDeviceToHost));
    CU(starts.alloc(n_runs * 8));
    CU(values.alloc(n_runs * 2));
    CU(rle_select(biased, p->value_bits, cum.as<uint64_t>(), n_chrom, pos_begin, pos_end, starts.as<uint64_t>(), d_count, temp.p, tb2, nullptr));
    CU(rle_gather(biased, p->value_bits, starts.as<uint64_t>(), n_runs, values.as<uint16_t>(), nullptr));
    CU(cudaEventRecord(ix->ev1, nullptr));
    uint64_t* h_start = static_cast<uint64_t*>(std::malloc(n_runs * 8 + 8));
    uint16_t* h_value = static_cast<uint16_t*>(std::malloc(n_runs * 2 + 2));
    if (!h_start || !h_value) { std::free(h_start); std::free(h_value); return fail(GMB_ERR_NOMEM, "out of host memory"); }
    cudaError_t e = cudaMemcpy(h_start, starts.p, n_runs * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(h_value, values.p, n_runs * 2, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { std::free(h_start); std::free(h_value); return cuda_fail(e, "cudaMemcpy(runs)"); }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ix->ev0, ix->ev1);
    out->rle_ms = ms;
    out->n_runs = n_runs;
    out->start = h_start;
    out->value = h_value;
    return GMB_OK;
}

int gmb_index_export_sa(gmb_index* ix, uint32_t* out_host)
{
    if (!ix || !out_host) return fail(GMB_ERR_ARG, "gmb_index_export_sa: NULL argument");
    if (!ix->h.off_sa) return fail(GMB_ERR_UNSUPPORTED, "the index holds no suffix array (build it with GMB_BUILD_WITH_SA)");
    CU(cudaSetDevice(ix->device));
    CU(cudaMemcpy(out_host, ix->d_blob + ix->h.off_sa, ix->h.n_bwt * 4, cudaMemcpyDeviceToHost));
    return GMB_OK;
}

int gmb_index_export_bwt(gmb_index* ix, int rev, uint8_t* out_host)
{
    if (!ix || !out_host) return fail(GMB_ERR_ARG, "gmb_index_export_bwt: NULL argument");
    CU(cudaSetDevice(ix->device));
    const uint64_t n = ix->h.n_bwt;
    uint8_t* d = nullptr;
    CU(cudaMalloc(&d, n ? n : 1));
    const uint8_t* Bp = ix->d_blob + (rev ? ix->h.off_rev : ix->h.off_fwd);
    const uint32_t* S = reinterpret_cast<const uint32_t*>(ix->d_blob + (rev ? ix->h.off_sent_rev : ix->h.off_sent_fwd));
    if (ix->h.sigma == 5) k_decode_bwt5<<<(unsigned)((n + 255) / 256), 256>>>(reinterpret_cast<const RankBlock5*>(Bp), n, d);
    else k_decode_bwt<<<(unsigned)((n + 255) / 256), 256>>>(reinterpret_cast<const RankBlock*>(Bp), n, d);
    k_mark_sentinels<<<(ix->h.n_seq + 255) / 256, 256>>>(S, ix->h.n_seq, d);
    cudaError_t e = cudaMemcpy(out_host, d, n, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "export bwt");
    return GMB_OK;
}

} // extern "C"
