// pb_engine.cu — host side of libpetar_b200.so: the C ABI of include/petar_b200.h.
//
// Owns all device state (process-global, lazily created — the reference's functors are
// temporaries re-constructed every tree step, reference src/petar.hpp:894, so nothing can live
// in them), packs FDPS's fp64 AoS inputs into pinned staging in the device formats, plans the
// CTA tasks, and drives H2D -> force kernel -> reduce kernel -> D2H per stream.
//
// Differences to the reference's host code (src/force_gpu_cuda.cu:535-880) by design:
//   * one pinned arena and ONE async H2D per stream and dispatch instead of four blocking copies;
//   * a dispatch is split across several CUDA streams so copy-in, kernels and copy-out of the
//     sub-batches overlap each other and the host's next tree walk;
//   * buffers grow on demand (the reference asserts on fixed 1e6/1e7-entry buffers, :14-16);
//   * every CUDA return code is checked; errors surface through pb_last_error().
#include "petar_b200.h"
#include "pb_device.h"
#include <cub/cub.cuh>
#include <atomic>
#include <emmintrin.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace pb;

namespace {

// ------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------
char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? PB_ERR_NO_DEVICE : PB_ERR_CUDA, \
                        "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------
// state
// ------------------------------------------------------------------------------------------
constexpr int kMaxStreams = 8;

struct Plan {                // layout of one sub-batch inside an arena
    int    n_walk = 0, n_tasks = 0, n_iblocks = 0;
    size_t n_i = 0, n_ide = 0, n_ids = 0, n_part = 0;
    size_t n_lepj = 0, n_lspj = 0;            // direct mode: dispatch-local j store entries
    size_t off_walks = 0, off_tasks = 0, off_iblocks = 0, off_done = 0, off_epi = 0, off_ide = 0, off_ids = 0;
    size_t off_lepj = 0, off_lspj = 0, bytes = 0;
    int    count_only = 0;                    // neighbour search only: EP lists as kind-2 tasks, no SP, eps2 = 0
    int    coords = 0, i_f4 = 2;              // option "coords" the sub-batch was packed for; float4 per packed i-particle
    size_t n_runs = 0, n_idx = 0;             // EP lists as runs: (start, length) pairs in the arena; n_idx entries after expansion on the device
    size_t off_runtab = 0;                    // per walk {first run, number of runs} (int2)
    const int* ext_ide = nullptr;             // index lists living outside the arena (built on the device for the whole step)
    const int* ext_ids = nullptr;
};

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t  ev[4] = {nullptr, nullptr, nullptr, nullptr};   // h2d start, h2d end, kernels end, d2h end
    char*  h_arena = nullptr; char* d_arena = nullptr; size_t cap_arena = 0;
    ForceOut* h_out = nullptr; ForceOut* d_out = nullptr; size_t cap_out = 0;
    double4* d_part4 = nullptr; int* d_partn = nullptr; size_t cap_part = 0;
    Plan plan;
    int  w_begin = 0, w_end = 0;
    bool active = false;
    // neighbour-list emission (count-only dispatches with option "nb_lists")
    unsigned long long* d_pairs = nullptr; unsigned long long* d_pairs_sorted = nullptr; unsigned long long* h_pairs = nullptr;
    size_t cap_pairs = 0;
    unsigned int* d_cursor = nullptr; unsigned int* h_cursor = nullptr;
    void* d_cubtmp = nullptr; size_t cap_cubtmp = 0;
    bool emit = false; int i_base = 0;
    size_t n_pairs_window = 0;         // entries of d_pairs this sub-batch may use (cleared, written, sorted): <= cap_pairs
    long long j_epoch = -1;            // the j publication this stream has already been ordered after
    int* d_ide_x = nullptr; size_t cap_ide_x = 0;   // EP index lists of the sub-batch, expanded on the device from the runs
};

struct Recorded {
    char* d_arena = nullptr;
    int*  d_ide_x = nullptr;     // its own copy of the expanded EP lists (the slot's buffer is reused by later dispatches)
    Plan  plan;
    bool  direct = false;
    int   slot = 0;              // the stream slot the dispatch ran on
};

struct Engine {
    bool inited = false;
    int  rank = 0, device = 0;
    double eps2 = 0.0, rcut2 = 0.0, G = 1.0;
    int opt_coords = 2, opt_streams = 8, opt_jchunk = 0, opt_nr = 0, opt_cull = 1, opt_occ = 2, opt_lead = 3;

    // j store
    float4* d_epj = nullptr; size_t cap_epj = 0; int n_epj = 0;
    float4* d_spj = nullptr; size_t cap_spj = 0; int n_spj = 0;
    float4* h_jstage = nullptr; size_t cap_jstage = 0;     // pinned, in float4 units
    cudaStream_t s_upload = nullptr;
    cudaEvent_t  ev_j_ready = nullptr, ev_send0 = nullptr, ev_send1 = nullptr;
    bool j_published = false, send_timed = false;

    Slot slots[kMaxStreams];
    bool outstanding = false;
    bool count_only = false, out_count_only = false;   // transient: set by pb_dispatch_count_index
    int  out_n_walk = 0, out_n_slots = 0;
    std::vector<int> out_ni;

    bool recording = false;
    std::vector<Recorded> recs;

    // device-side list building (pb_tree_*)
    void* d_cells = nullptr; void* d_groups = nullptr; size_t cap_cells = 0, cap_groups = 0;
    int n_cells = 0, n_groups = 0; double theta = 0.3;
    char* h_tstage = nullptr; size_t cap_tstage = 0;     // pinned staging of the tree
    std::vector<int> grp_n;                       // particles per group
    int2* d_counts = nullptr; size_t cap_counts = 0; std::vector<int2> h_counts;
    int* d_walk_scratch[kMaxStreams] = {nullptr}; int* d_overflow = nullptr;
    int* d_tree_ide = nullptr; int* d_tree_ids = nullptr; size_t cap_tree_ide = 0, cap_tree_ids = 0;
    int2* d_tree_off = nullptr; std::vector<int2> h_tree_off; cudaEvent_t ev_fill = nullptr;
    char* h_corr = nullptr; char* d_corr = nullptr; size_t cap_corr = 0;   // changeover correction: pinned staging + device mirror
    int* d_elem_map = nullptr; int* h_elem_map = nullptr; size_t cap_elem_map = 0; bool has_elem_map = false;   // device walk over a tree with LET elements
    // end of the last kernel of the two most recent dispatches (gap timer, pb_profile.t_gap)
    cudaEvent_t ev_end[2] = {nullptr, nullptr}; int end_cur = 0; bool end_prev_valid = false; int out_first_slot = -1;
    int2* h_counts_p = nullptr; size_t cap_counts_p = 0; int* h_over_p = nullptr;   // pinned landing zone of the count pass
    cudaEvent_t ev_count = nullptr; bool count_pending = false;
    int opt_tree_batch = 256; int tree_last_batches = 0;
    int opt_tree_fill = 0;                                 // pb_tree_force: 0 one step-wide list-fill launch, 1 one per batch on the batch's stream
    long long j_epoch = 0;                                 // bumped whenever ev_j_ready is re-recorded
    long long opt_min_slot_work = 0;                       // > 0: a dispatch is not cut into sub-batches smaller than this many EP-equivalent interactions
    int opt_tree_streams = 4;                              // streams pb_tree_force cycles its batches over
    int opt_tree_spec = 1;                                 // reserve list space from the previous step's lengths and fill in ONE walk pass
    std::vector<int2> prev_counts;                         // list lengths of the previous pb_tree_force
    bool spec_pending = false; int2* d_tree_caps = nullptr; size_t cap_tree_caps = 0;
    int opt_walk_ctas = 148 * 16;                          // CTAs (4 warps each, one warp per i-group at a time) of the tree-walk launches
    int opt_nb_lists = 0;                                  // count-only dispatches also return the neighbour pairs
    int opt_ep_runs = 0;                                   // 1: EP index lists cross PCIe as (start, length) runs and are expanded on the device (measured: does not pay)
    int opt_walk_compact = 1;                              // tree walk classifies on 64-B fp32 records first (exact fp64 re-check when undecided)
    void* d_cellA = nullptr; void* d_cellB = nullptr; size_t cap_cellAB = 0; double coord_max = 0.0;
    int opt_raw_upload = 0;                                // pb_upload_j_range copies the caller's arrays as they are and packs them on the device
    char* d_raw = nullptr; size_t cap_raw = 0;             // device landing zone of the raw arrays
    std::vector<std::pair<const char*, size_t>> registered;   // host ranges page-locked by the library (cudaHostRegister)
    int opt_ws = 1;                                        // persistent launches use the warp-specialised kernel (pb_kernels_ws.cu)
    int opt_raw_result = 0;                                // device-resident step: page-lock the caller's force array once and let the kernel write into it
    int opt_fuse_reduce = 1;                               // device-resident step: the force kernel adds up finished i-blocks and writes the forces to host memory itself
    int opt_sp2i = 1;                                      // SP tasks of groups with >= 2 i-blocks keep two i-particles per lane (sp_pairs_2i)
    int opt_chunk_tile = 1;                                // j chunks are whole 256-entry tiles (0: multiples of 8 entries, the round-1 plan)
    std::vector<unsigned long long> nb_keys;               // (i << 32 | j) of the last retrieved count dispatch, sorted
    long long nb_n_i = 0;

    // device-resident tree step (pb_tree_force_resident): i-particles, plan and forces never visit the host
    int* d_ifirst = nullptr; int* h_ifirst = nullptr; size_t cap_ifirst = 0; long long r_n_i = 0; int r_n_iblk = 0;
    Walk* d_r_walks = nullptr; int3* d_r_goff = nullptr; size_t cap_r_groups = 0;
    float4* d_r_epi = nullptr; ForceOut* d_r_out = nullptr; ForceOut* h_r_out = nullptr; IBlock* d_r_iblocks = nullptr; int* d_r_done = nullptr; size_t cap_r_i = 0;
    Task* d_r_tasks = nullptr; size_t cap_r_tasks = 0;
    double4* d_r_part4 = nullptr; int* d_r_partn = nullptr; size_t cap_r_part = 0;
    int* d_r_meta = nullptr; int* h_r_meta = nullptr;
    long long r_prev_tasks = 0, r_prev_part = 0;
    cudaEvent_t ev_tl[8] = {nullptr}; bool tl_valid = false; float tl_ms[8] = {0};
    int* d_let_idx = nullptr; int* h_let_idx = nullptr; size_t cap_let_idx = 0;      // LET send lists (EP store slots)
    int* d_let_err = nullptr; int* h_let_err = nullptr;                              // an index outside the store, found by the gather kernel
    char* d_raw_let = nullptr; size_t cap_raw_let = 0;                               // raw LET SP rows (option raw_upload)
    float4* h_let_sp = nullptr; size_t cap_let_sp = 0;                               // host-packed LET SP rows (pinned)

    pb_profile prof;
};

Engine E;

// ------------------------------------------------------------------------------------------
// buffers
// ------------------------------------------------------------------------------------------
int ensure_init() {
    if (E.inited) return PB_OK;
    return pb_init(0, -1);
}

int grow_arena(Slot& s, size_t bytes) {
    if (bytes <= s.cap_arena) return PB_OK;
    const size_t cap = align_up(bytes + bytes / 2, 1 << 20);
    CU(cudaStreamSynchronize(s.stream));
    if (s.h_arena) CU(cudaFreeHost(s.h_arena));
    if (s.d_arena) CU(cudaFree(s.d_arena));
    s.h_arena = nullptr; s.d_arena = nullptr; s.cap_arena = 0;
    CU(cudaMallocHost(&s.h_arena, cap));
    CU(cudaMalloc(&s.d_arena, cap));
    s.cap_arena = cap;
    return PB_OK;
}

int grow_out(Slot& s, size_t n) {
    if (n <= s.cap_out) return PB_OK;
    const size_t cap = align_up(n + n / 2, 4096);
    CU(cudaStreamSynchronize(s.stream));
    if (s.h_out) CU(cudaFreeHost(s.h_out));
    if (s.d_out) CU(cudaFree(s.d_out));
    s.h_out = nullptr; s.d_out = nullptr; s.cap_out = 0;
    CU(cudaMallocHost(&s.h_out, cap * sizeof(ForceOut)));
    CU(cudaMalloc(&s.d_out, cap * sizeof(ForceOut)));
    s.cap_out = cap;
    return PB_OK;
}

int grow_part(Slot& s, size_t n) {
    if (n <= s.cap_part) return PB_OK;
    const size_t cap = align_up(n + n / 2, 4096);
    CU(cudaStreamSynchronize(s.stream));
    if (s.d_part4) CU(cudaFree(s.d_part4));
    if (s.d_partn) CU(cudaFree(s.d_partn));
    s.d_part4 = nullptr; s.d_partn = nullptr; s.cap_part = 0;
    CU(cudaMalloc(&s.d_part4, cap * sizeof(double4)));
    CU(cudaMalloc(&s.d_partn, cap * sizeof(int)));
    s.cap_part = cap;
    return PB_OK;
}

int grow_pairs(Slot& s, size_t n) {
    if (!s.d_cursor) {
        CU(cudaMalloc(&s.d_cursor, sizeof(unsigned int)));
        CU(cudaMallocHost(&s.h_cursor, sizeof(unsigned int)));
    }
    if (n <= s.cap_pairs) return PB_OK;
    const size_t cap = align_up(n + n / 2, 4096);
    CU(cudaStreamSynchronize(s.stream));
    if (s.d_pairs) CU(cudaFree(s.d_pairs));
    if (s.d_pairs_sorted) CU(cudaFree(s.d_pairs_sorted));
    if (s.h_pairs) CU(cudaFreeHost(s.h_pairs));
    s.d_pairs = s.d_pairs_sorted = s.h_pairs = nullptr; s.cap_pairs = 0;
    CU(cudaMalloc(&s.d_pairs, cap * sizeof(unsigned long long)));
    CU(cudaMalloc(&s.d_pairs_sorted, cap * sizeof(unsigned long long)));
    CU(cudaMallocHost(&s.h_pairs, cap * sizeof(unsigned long long)));
    s.cap_pairs = cap;
    size_t tmp = 0;
    CU(cub::DeviceRadixSort::SortKeys(nullptr, tmp, s.d_pairs, s.d_pairs_sorted, (int)cap, 0, 64, s.stream));
    if (tmp > s.cap_cubtmp) {
        if (s.d_cubtmp) CU(cudaFree(s.d_cubtmp));
        s.d_cubtmp = nullptr; s.cap_cubtmp = 0;
        CU(cudaMalloc(&s.d_cubtmp, tmp));
        s.cap_cubtmp = tmp;
    }
    return PB_OK;
}

// sort the whole pair buffer of a sub-batch (unused entries are all-ones and stay at the end): queued behind the
// kernel in dispatch, so no host round trip sits between the kernel and the sort
cudaError_t sort_pairs(Slot& s, size_t n_i_dispatch) {
    int bits = 1;
    while (((size_t)1 << bits) <= n_i_dispatch + 1) bits++;
    size_t tmp = s.cap_cubtmp;
    return cub::DeviceRadixSort::SortKeys(s.d_cubtmp, tmp, s.d_pairs, s.d_pairs_sorted, (int)s.n_pairs_window, 0, std::min(64, 32 + bits), s.stream);
}

int grow_jstore(size_t n_epj, size_t n_spj) {
    if (n_epj > E.cap_epj) {
        const size_t cap = align_up(n_epj + n_epj / 4, 4096);
        CU(cudaDeviceSynchronize());
        if (E.d_epj) CU(cudaFree(E.d_epj));
        E.d_epj = nullptr; E.cap_epj = 0;
        CU(cudaMalloc(&E.d_epj, cap * PB_EPJ_DEV_BYTES));
        E.cap_epj = cap;
    }
    if (n_spj > E.cap_spj) {
        const size_t cap = align_up(n_spj + n_spj / 4, 4096);
        CU(cudaDeviceSynchronize());
        if (E.d_spj) CU(cudaFree(E.d_spj));
        E.d_spj = nullptr; E.cap_spj = 0;
        CU(cudaMalloc(&E.d_spj, cap * PB_SPJ_DEV_BYTES));
        E.cap_spj = cap;
    }
    return PB_OK;
}

int grow_jstage(size_t n_float4) {
    if (n_float4 <= E.cap_jstage) return PB_OK;
    const size_t cap = align_up(n_float4 + n_float4 / 4, 1 << 16);
    CU(cudaStreamSynchronize(E.s_upload));
    if (E.h_jstage) CU(cudaFreeHost(E.h_jstage));
    E.h_jstage = nullptr; E.cap_jstage = 0;
    CU(cudaMallocHost(&E.h_jstage, cap * sizeof(float4)));
    E.cap_jstage = cap;
    return PB_OK;
}

// ------------------------------------------------------------------------------------------
// packing fp64 AoS -> device formats
// ------------------------------------------------------------------------------------------
inline double ld(const char* p, size_t off, int k = 0) {
    double v;
    memcpy(&v, p + off + 8 * (size_t)k, 8);
    return v;
}

inline void split(double x, float& hi, float& lo) {
    hi = (float)x;
    lo = (float)(x - (double)hi);
}

void pack_epj(const void* epj, int n, const pb_layout_epj& L, float4* out) {
    const char* base = (const char*)epj;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const char* p = base + (size_t)i * L.stride;
        float4 a, b;
        split(ld(p, L.off_pos, 0), a.x, b.x);
        split(ld(p, L.off_pos, 1), a.y, b.y);
        split(ld(p, L.off_pos, 2), a.z, b.z);
        a.w = (float)ld(p, L.off_mass);
        b.w = (float)ld(p, L.off_rsearch);
        out[2 * (size_t)i]     = a;
        out[2 * (size_t)i + 1] = b;
    }
}

void pack_spj(const void* spj, int n, const pb_layout_spj& L, float4* out) {
    const char* base = (const char*)spj;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const char* p = base + (size_t)i * L.stride;
        float4 a, b, c, d;
        split(ld(p, L.off_pos, 0), a.x, b.x);
        split(ld(p, L.off_pos, 1), a.y, b.y);
        split(ld(p, L.off_pos, 2), a.z, b.z);
        a.w = (float)ld(p, L.off_mass);
        double q[6] = {0, 0, 0, 0, 0, 0};
        if (L.has_quad)
            for (int k = 0; k < 6; k++) q[k] = ld(p, L.off_quad, k);      // xx yy zz xy xz yz
        // traceless form q' = 3 q - tr I, in fp64 (what the kernel's quadrupole formula consumes)
        const double tr = q[0] + q[1] + q[2];
        b.w = (float)(3.0 * q[0] - tr);
        c = make_float4((float)(3.0 * q[1] - tr), (float)(3.0 * q[2] - tr), (float)(3.0 * q[3]), (float)(3.0 * q[4]));
        d = make_float4((float)(3.0 * q[5]), (float)tr, 0.f, 0.f);
        float4* o = out + 4 * (size_t)i;
        o[0] = a; o[1] = b; o[2] = c; o[3] = d;
    }
}

// ------------------------------------------------------------------------------------------
// task planning
// ------------------------------------------------------------------------------------------
struct WalkIn {              // per-walk host inputs (either mode)
    const void* epi; int ni;
    const int* ide; int nej;            // index mode
    const int* ids; int nsj;
    const void* epj; const void* spj;   // direct mode
};
// ide / ids pointing here: the index lists are built on the device straight into the arena
// (pb_tree_force); the host reserves the space and copies nothing
const int g_devlist_marker = 0;

struct Group { int walk, i_first, nib, jsplit; };

// binary decomposition of the walk's i-blocks into groups of 8/4/2/1 warps so that every warp
// of every CTA is busy (a group of nib < 8 blocks splits each j tile across 8/nib warps)
void make_groups(int walk, int ni, std::vector<Group>& out) {
    int nib = (ni + 31) / 32, ib = 0;
    while (nib >= kWarpsPerCta) {
        out.push_back({walk, ib * 32, kWarpsPerCta, 1});
        ib += kWarpsPerCta; nib -= kWarpsPerCta;
    }
    for (int g = kWarpsPerCta / 2; g >= 1; g >>= 1)
        if (nib & g) { out.push_back({walk, ib * 32, g, kWarpsPerCta / g}); ib += g; }
}

// host-side plan of one sub-batch
struct HostPlan {
    Plan p;
    std::vector<Walk>   walks;
    std::vector<Task>   tasks;
    std::vector<IBlock> iblocks;
    std::vector<size_t> lepj_off, lspj_off;   // direct mode: first local j of each walk
    bool use_runs = false;                    // index mode: the EP lists travel run-length coded
    long long run_cursor = 0;                 // runs appended so far to the arena's run section (walks are packed concurrently: atomic adds)
    std::vector<int2> runtab;                 // per walk {first run, number of runs}
    std::vector<std::pair<size_t, int>> zero_blocks;   // {first i, count} of block groups without any chunk (fused reduction: zeroed on the host)
};

void plan_batch(const WalkIn* win, int n_walk, bool direct, int n_streams_active, HostPlan& hp, const int2* ext_off = nullptr) {
    hp.walks.resize(n_walk);
    hp.tasks.clear(); hp.iblocks.clear(); hp.zero_blocks.clear();
    hp.lepj_off.assign(n_walk, 0); hp.lspj_off.assign(n_walk, 0);
    std::vector<Group> groups;
    size_t i_off = 0, ide = 0, ids = 0, lepj = 0, lspj = 0;
    double work = 0.0;                                   // warp-steps: sum nib * (nej + 2 nsj)
    if (hp.use_runs) { hp.runtab.assign(n_walk, make_int2(0, 0)); hp.run_cursor = 0; }
    for (int w = 0; w < n_walk; w++) {
        Walk& W = hp.walks[w];
        W.i_off = (int)i_off; W.ni = win[w].ni;
        // index mode with a null EP list pointer = DENSE walk: its EP list is the whole j store in order
        const bool dense = !direct && win[w].ide == nullptr && win[w].nej > 0;
        W.ej_off = dense ? -1 : (int)ide;  W.nej = win[w].nej;
        W.sj_off = (int)ids;  W.nsj = win[w].nsj;
        if (ext_off) { W.ej_off = ext_off[w].x; W.sj_off = ext_off[w].y; }     // lists live in a step-wide device buffer
        W.ohx = W.ohy = W.ohz = W.olx = W.oly = W.olz = 0.f;
        W.hx = W.hy = W.hz = INFINITY; W.rsi2max = 0.f;
        i_off += (size_t)win[w].ni;
        if (!dense && !ext_off) ide += align_up((size_t)win[w].nej, 4);

        if (!ext_off) ids += align_up((size_t)win[w].nsj, 4);
        if (direct) {
            hp.lepj_off[w] = lepj; hp.lspj_off[w] = lspj;
            lepj += (size_t)win[w].nej; lspj += (size_t)win[w].nsj;
        }
        make_groups(w, win[w].ni, groups);
        work += (double)((win[w].ni + 31) / 32) * ((double)win[w].nej + 2.0 * (double)win[w].nsj);
    }
    // j per warp and task: aim for ~6 waves of 2 CTAs/SM over all concurrently running streams
    int U = E.opt_jchunk;
    if (U <= 0) {
        const double target_tasks = 148.0 * 2.0 * 6.0 / (double)std::max(1, n_streams_active);
        double u = work / (kWarpsPerCta * target_tasks);
        U = (int)std::min(4096.0, std::max(256.0, u));
    }
    U = (int)align_up((size_t)U, kTileJ);
    const int Us = std::max(kTileJ, U / 2);              // SP steps cost ~2x an EP step

    size_t part = 0;
    for (const Group& g : groups) {
        const Walk& W = hp.walks[g.walk];
        const int stride = g.nib * 32;
        const int nce = W.nej > 0 ? (int)((W.nej + (size_t)U * g.jsplit - 1) / ((size_t)U * g.jsplit)) : 0;
        const int ncs = (W.nsj > 0 && !E.count_only) ? (int)((W.nsj + (size_t)Us * g.jsplit - 1) / ((size_t)Us * g.jsplit)) : 0;
        const int part_base = (int)part;
        int chunk = 0;
        for (int kind = 0; kind < 2; kind++) {
            const int nj = kind == 0 ? W.nej : W.nsj;
            const int nc = kind == 0 ? nce : ncs;
            if (nc == 0) continue;
            // equal chunks of whole j tiles: only the last chunk of a list ends in a ragged tile (which the warps that
            // share an i-block then split evenly, see force_kernel)
            const int len = (int)align_up((size_t)(nj + nc - 1) / nc, E.opt_chunk_tile ? kTileJ : 8);
            for (int c = 0; c < nc; c++) {
                const int jb = c * len;
                if (jb >= nj) break;
                Task t;
                t.walk = g.walk; t.i_first = g.i_first; t.nib = g.nib; t.jsplit = g.jsplit;
                t.kind = (kind == 0 && E.count_only) ? 2 : kind; t.j_begin = jb; t.j_count = std::min(len, nj - jb);
                t.part_base = part_base + chunk * stride;
                t.blk0 = (int)hp.iblocks.size(); t.n_chunks = 0; t.pad1 = t.pad2 = 0;
                hp.tasks.push_back(t);
                chunk++;
            }
        }
        for (int c = 0; c < chunk; c++) hp.tasks[hp.tasks.size() - 1 - c].n_chunks = chunk;
        if (chunk == 0) hp.zero_blocks.push_back({(size_t)W.i_off + g.i_first, std::min(g.nib * 32, W.ni - g.i_first)});
        for (int b = 0; b < g.nib; b++) {
            IBlock ib;
            ib.part_base = part_base + b * 32;
            ib.n_chunks = chunk;
            ib.stride = stride;
            ib.out_off = W.i_off + g.i_first + b * 32;
            ib.n_valid = std::min(32, W.ni - (g.i_first + b * 32));
            ib.pad0 = ib.pad1 = ib.pad2 = 0;
            hp.iblocks.push_back(ib);
        }
        part += (size_t)chunk * stride;
    }
    // longest tasks first: CTAs are handed out in blockIdx order
    std::stable_sort(hp.tasks.begin(), hp.tasks.end(), [](const Task& a, const Task& b) {
        const long long wa = (long long)a.nib * a.j_count * (a.kind ? 2 : 1);
        const long long wb = (long long)b.nib * b.j_count * (b.kind ? 2 : 1);
        return wa > wb;
    });

    Plan& p = hp.p;
    p.count_only = E.count_only ? 1 : 0;
    p.coords = E.opt_coords;
    p.i_f4 = (E.opt_coords == 2 && !E.count_only) ? 3 : 2;
    p.n_walk = n_walk; p.n_tasks = (int)hp.tasks.size(); p.n_iblocks = (int)hp.iblocks.size();
    p.n_i = i_off; p.n_ide = ide; p.n_ids = ids; p.n_part = part; p.n_lepj = lepj; p.n_lspj = lspj;
    p.n_runs = 0; p.n_idx = 0;
    if (hp.use_runs) { p.n_idx = ide; p.n_ide = 0; }      // the arena carries the runs (appended while packing, at its end), not the indices
    size_t o = 0;
    p.off_walks = o;   o = align_up(o + sizeof(Walk) * n_walk, 256);
    p.off_tasks = o;   o = align_up(o + sizeof(Task) * p.n_tasks, 256);
    p.off_iblocks = o; o = align_up(o + sizeof(IBlock) * p.n_iblocks, 256);
    p.off_done = o;    o = align_up(o + sizeof(int) * p.n_iblocks, 256);     // chunk counters of the fused reduction: travel as zeros, return to zero
    p.off_runtab = o;  o = align_up(o + (hp.use_runs ? sizeof(int2) * (size_t)n_walk : 0), 256);
    p.off_epi = o;     o = align_up(o + (size_t)p.i_f4 * sizeof(float4) * p.n_i, 256);
    p.off_ide = o;     o = align_up(o + sizeof(int) * p.n_ide, 256);
    p.off_ids = o;     o = align_up(o + sizeof(int) * p.n_ids, 256);
    p.off_lepj = o;    o = align_up(o + (size_t)PB_EPJ_DEV_BYTES * p.n_lepj, 256);
    p.off_lspj = o;    o = align_up(o + (size_t)PB_SPJ_DEV_BYTES * p.n_lspj, 256);
    p.bytes = o;
    if (hp.use_runs) { p.off_ide = o; p.bytes = o + sizeof(int2) * p.n_idx; }   // capacity; the copy ends after the runs actually found
}

// one walk of a sub-batch into its pinned arena: i-particles relative to the walk origin, index lists
void pack_walk(const WalkIn* win, bool direct, const pb_layout_epi& Li, HostPlan& hp, char* arena, int w) {
    const Plan& p = hp.p;
    float4* epi = (float4*)(arena + p.off_epi);
    int* ide = (int*)(arena + p.off_ide);
    int* ids = (int*)(arena + p.off_ids);
    const bool rel = (p.coords != 1);
    const int f4 = p.i_f4;
    Walk& W = hp.walks[w];
    const char* base = (const char*)win[w].epi;
    // Walk origin = mean position of the i-particles (robust against outliers, unlike the box
    // centre), as a hi/lo fp32 pair.  i and j positions are shifted to it by the SAME fp32
    // operation sequence — (x_hi - o_hi) + (x_lo - o_lo) — here for i, in the kernel for j, so
    // that a particle meeting itself (or an exact copy) gives dx == 0 exactly, as in the
    // reference kernel; absolute mode is the same code with a zero origin (then x_rel == x_hi).
    float oh[3] = {0.f, 0.f, 0.f}, ol[3] = {0.f, 0.f, 0.f};
    if (rel && W.ni > 0) {
        double sum[3] = {0.0, 0.0, 0.0};
        for (int i = 0; i < W.ni; i++) {
            const char* q = base + (size_t)i * Li.stride;
            for (int k = 0; k < 3; k++) sum[k] += ld(q, Li.off_pos, k);
        }
        for (int k = 0; k < 3; k++) split(sum[k] / W.ni, oh[k], ol[k]);
    }
    W.ohx = oh[0]; W.ohy = oh[1]; W.ohz = oh[2];
    W.olx = ol[0]; W.oly = ol[1]; W.olz = ol[2];
    float4* e = epi + (size_t)f4 * (size_t)W.i_off;  // per i: {x,y,z,rs}, {xl,yl,zl,0} [, {float(x),float(y),float(z),0}: coords = 2]
    float hmax[3] = {0.f, 0.f, 0.f}, rsmax = 0.f;
    for (int i = 0; i < W.ni; i++) {
        const char* q = base + (size_t)i * Li.stride;
        float r[3], rl[3], xa[3];
        for (int k = 0; k < 3; k++) {
            float xh, xl;
            split(ld(q, Li.off_pos, k), xh, xl);
            rel_hilo(xh, xl, oh[k], ol[k], r[k], rl[k]);
            if (!rel) rl[k] = 0.f;
            xa[k] = xh;
        }
        const float4 v = make_float4(r[0], r[1], r[2], (float)ld(q, Li.off_rsearch));
        e[f4 * i] = v;
        e[f4 * i + 1] = make_float4(rl[0], rl[1], rl[2], 0.f);
        if (f4 == 3) e[f4 * i + 2] = make_float4(xa[0], xa[1], xa[2], 0.f);
        hmax[0] = std::max(hmax[0], std::fabs(v.x)); hmax[1] = std::max(hmax[1], std::fabs(v.y));
        hmax[2] = std::max(hmax[2], std::fabs(v.z)); rsmax = std::max(rsmax, v.w);
    }
    // bounding box of the packed fp32 i-positions about the origin: lets the kernel skip the
    // neighbour test for j segments that are provably out of reach of every i of the walk
    if (rel && E.opt_cull) { W.hx = hmax[0]; W.hy = hmax[1]; W.hz = hmax[2]; }
    else                   { W.hx = W.hy = W.hz = INFINITY; }
    // "near" radius: max r_search of the i-particles, and at least kPrecFactor ulps of the box
    // half-size — inside it a single-float relative coordinate is not accurate enough against
    // the pair separation, so those segments take the exact-dx loop
    const float hbox = std::max(hmax[0], std::max(hmax[1], hmax[2]));
    const float rprec = rel ? 4.0e4f * (std::nextafter(hbox, INFINITY) - hbox) : 0.f;
    const float rnear = std::max(rsmax, rprec);
    W.rsi2max = rnear * rnear;
    if (!direct) {
        if (hp.use_runs) {
            // maximal runs of consecutive indices (FDPS lists are leaf cells in Morton order), found in the one pass the
            // host makes over the list; appended to the arena's run section wherever the cursor stands
            static thread_local std::vector<int2> R;
            R.clear();
            const int* id = win[w].ide;
            const int n = W.nej;
            if (id && n > 0 && W.ej_off >= 0) {
                int start = id[0], len = 1;
                for (int k = 1; k < n; k++) {
                    if (id[k] == start + len) len++;
                    else { R.push_back(make_int2(start, len)); start = id[k]; len = 1; }
                }
                R.push_back(make_int2(start, len));
            }
            const long long at = __atomic_fetch_add(&hp.run_cursor, (long long)R.size(), __ATOMIC_RELAXED);
            hp.runtab[w] = make_int2((int)at, (int)R.size());
            if (!R.empty()) memcpy(reinterpret_cast<int2*>(arena + p.off_ide) + at, R.data(), sizeof(int2) * R.size());
        } else if (W.nej && W.ej_off >= 0 && win[w].ide != &g_devlist_marker) memcpy(ide + W.ej_off, win[w].ide, sizeof(int) * (size_t)W.nej);
        if (W.nsj && win[w].ids != &g_devlist_marker) memcpy(ids + W.sj_off, win[w].ids, sizeof(int) * (size_t)W.nsj);
    } else {
        const size_t e0 = hp.lepj_off[w], s0 = hp.lspj_off[w];
        for (int j = 0; j < W.nej; j++) ide[W.ej_off + j] = (int)(e0 + j);
        for (int j = 0; j < W.nsj; j++) ids[W.sj_off + j] = (int)(s0 + j);
    }
}

// what follows once every walk of the sub-batch is packed: direct-mode j arrays, the tables
void pack_tail(const WalkIn* win, bool direct, const pb_layout_epj* Lj, const pb_layout_spj* Ls, HostPlan& hp, char* arena) {
    const Plan& p = hp.p;
    float4* lepj = (float4*)(arena + p.off_lepj);
    float4* lspj = (float4*)(arena + p.off_lspj);
    if (direct) {
        // per-walk j arrays -> dispatch-local j store (pack_* parallelise internally)
        for (int w = 0; w < p.n_walk; w++) {
            if (win[w].nej) pack_epj(win[w].epj, win[w].nej, *Lj, lepj + 2 * hp.lepj_off[w]);
            if (win[w].nsj) pack_spj(win[w].spj, win[w].nsj, *Ls, lspj + 4 * hp.lspj_off[w]);
        }
    }
    memcpy(arena + p.off_walks, hp.walks.data(), sizeof(Walk) * hp.walks.size());
    memcpy(arena + p.off_tasks, hp.tasks.data(), sizeof(Task) * hp.tasks.size());
    memcpy(arena + p.off_iblocks, hp.iblocks.data(), sizeof(IBlock) * hp.iblocks.size());
    memset(arena + p.off_done, 0, sizeof(int) * hp.iblocks.size());
    if (hp.use_runs) memcpy(arena + p.off_runtab, hp.runtab.data(), sizeof(int2) * hp.runtab.size());
}

// fill the pinned arena of one sub-batch
void pack_batch(const WalkIn* win, bool direct, const pb_layout_epi& Li,
                const pb_layout_epj* Lj, const pb_layout_spj* Ls, HostPlan& hp, char* arena) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int w = 0; w < hp.p.n_walk; w++) pack_walk(win, direct, Li, hp, arena, w);
    pack_tail(win, direct, Lj, Ls, hp, arena);
}

Params make_params(const Plan& p, const Slot* emit) {
    Params prm;
    prm.eps2 = p.count_only ? 0.f : (float)E.eps2;        // SearchNeighborEpEpNoSimd tests r2 without eps
    prm.rcut2 = (float)E.rcut2;
    if (p.coords == 2) {
        // the CPU replay clamps with r_out_32 * r_out_32 in float and takes 1/sqrt in double (src/hard.hpp:1428-1436)
        const float r32 = (float)std::sqrt(E.rcut2);
        prm.rcut2 = r32 * r32;
        prm.rinv_cut = prm.rcut2 > 0.f ? (float)(1.0 / std::sqrt((double)prm.rcut2)) : 0.f;
    } else prm.rinv_cut = 0.f;
    prm.abs_mode = p.count_only ? (p.coords == 1 ? 1 : 0) : p.coords;
    prm.i_f4 = p.i_f4;
    prm.i_base = emit ? emit->i_base : 0;
    prm.pair_cap = emit ? (unsigned int)std::min<size_t>(emit->n_pairs_window, 0xffffffffu) : 0u;
    prm.pairs = emit ? emit->d_pairs : nullptr;
    prm.pair_cursor = emit ? emit->d_cursor : nullptr;
    prm.meta = nullptr;
    prm.iblocks = nullptr; prm.done = nullptr; prm.out = nullptr; prm.G = E.G;
    return prm;
}

// the force kernel reduces finished i-blocks itself and writes the forces to `out_host` (page-locked, device-visible)
bool plan_fused(const Plan& p, const Slot* emit) { return E.opt_fuse_reduce && !p.count_only && !emit && E.opt_occ < 3; }

cudaError_t launch_plan(cudaStream_t st, const Plan& p, char* d_arena, bool direct,
                        double4* part4, int* partn, ForceOut* out, ForceOut* out_host, bool force_only = false, const Slot* emit = nullptr) {
    Params prm = make_params(p, emit);
    const bool fuse = plan_fused(p, emit) && out_host;
    if (fuse) {
        prm.iblocks = (const IBlock*)(d_arena + p.off_iblocks); prm.done = (int*)(d_arena + p.off_done); prm.out = out_host; prm.G = E.G;
    }
    const float4* epj = direct ? (const float4*)(d_arena + p.off_lepj) : E.d_epj;
    const float4* spj = direct ? (const float4*)(d_arena + p.off_lspj) : E.d_spj;
    cudaError_t e = launch_force(st, p.n_tasks, E.opt_nr, E.opt_occ,
                                 (const Walk*)(d_arena + p.off_walks), (const Task*)(d_arena + p.off_tasks),
                                 (const float4*)(d_arena + p.off_epi),
                                 p.ext_ide ? p.ext_ide : (const int*)(d_arena + p.off_ide),
                                 p.ext_ids ? p.ext_ids : (const int*)(d_arena + p.off_ids),
                                 epj, spj, part4, partn, prm, emit != nullptr, E.opt_sp2i != 0);
    if (e != cudaSuccess || force_only || fuse) return e;
    return launch_reduce(st, p.n_iblocks, (const IBlock*)(d_arena + p.off_iblocks), part4, partn, out, E.G);
}

// neighbour pairs of one finished sub-batch (its keys were sorted on the device behind the kernel: ascending i,
// then ascending j): start the copy of the n valid keys; rerun with a larger buffer first if it overflowed
int collect_pairs_begin(Slot& S, size_t n_i_dispatch) {
    unsigned int n = *S.h_cursor;
    while ((size_t)n > S.n_pairs_window) {
        S.n_pairs_window = (size_t)n + n / 4;
        int rc = grow_pairs(S, S.n_pairs_window);
        if (rc != PB_OK) return rc;
        CU(cudaMemsetAsync(S.d_cursor, 0, sizeof(unsigned int), S.stream));
        CU(cudaMemsetAsync(S.d_pairs, 0xff, sizeof(unsigned long long) * S.n_pairs_window, S.stream));
        CU(launch_plan(S.stream, S.plan, S.d_arena, false, S.d_part4, S.d_partn, S.d_out, nullptr, true, &S));
        CU(sort_pairs(S, n_i_dispatch));
        CU(cudaMemcpyAsync(S.h_cursor, S.d_cursor, sizeof(unsigned int), cudaMemcpyDeviceToHost, S.stream));
        CU(cudaStreamSynchronize(S.stream));
        E.prof.n_kernel_launch += 2;
        n = *S.h_cursor;
    }
    if (n) CU(cudaMemcpyAsync(S.h_pairs, S.d_pairs_sorted, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost, S.stream));
    E.prof.d2h_bytes += (long long)(sizeof(unsigned long long) * n);
    return PB_OK;
}
int collect_pairs_end(Slot& S) {
    const unsigned int n = *S.h_cursor;
    if (n == 0) return PB_OK;
    CU(cudaStreamSynchronize(S.stream));
    E.nb_keys.insert(E.nb_keys.end(), S.h_pairs, S.h_pairs + n);
    return PB_OK;
}

int dispatch_common(int n_walk, const WalkIn* win, bool direct, const pb_layout_epi& Li,
                    const pb_layout_epj* Lj, const pb_layout_spj* Ls) {
    if (E.outstanding) return fail(PB_ERR_PROTOCOL, "pb_dispatch_*: previous dispatch not retrieved (tag_max = 1)");
    if (!direct && !E.j_published) return fail(PB_ERR_PROTOCOL, "pb_dispatch_index before pb_upload_j");
    const double t0 = now_s();

    // contiguous sub-batches of ~equal work, one per stream; a dispatch with little work (many ranks sharing
    // the particles) is not cut finer than ~50 us of GPU time per sub-batch — every sub-batch costs ~13 us of
    // enqueueing on the calling thread
    std::vector<double> cum(n_walk + 1, 0.0);
    for (int w = 0; w < n_walk; w++)
        cum[w + 1] = cum[w] + (double)win[w].ni * ((double)win[w].nej + 2.0 * (double)win[w].nsj) + 1.0;
    const double min_work = (double)E.opt_min_slot_work;   // EP-equivalent interactions (option "min_slot_work")
    int n_slots = std::min(E.opt_streams, n_walk);
    if (E.count_only) n_slots = std::min(n_slots, 4);     // neighbour search: ~10x less GPU work per walk group (measured: 4 beats 8)
    if (min_work > 0.0) n_slots = std::min(n_slots, (int)(cum[n_walk] / min_work));
    n_slots = std::max(1, n_slots);
    std::vector<int> cut(n_slots + 1, 0);
    cut[n_slots] = n_walk;
    // the first sub-batch is 1/(1+lead) the size of the others: the GPU is idle until its copy lands
    const double w0 = 1.0 / (1.0 + std::max(0, E.opt_lead));
    const double unit = cum[n_walk] / (n_slots - 1 + w0);
    for (int s = 1; s < n_slots; s++) {
        const double target = unit * (s - 1 + w0);
        cut[s] = (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
        cut[s] = std::max(cut[s], cut[s - 1]);
        cut[s] = std::min(cut[s], n_walk);
    }

    long long n_i = 0, n_ej = 0, n_sj = 0, i_ep = 0, i_sp = 0;
    for (int w = 0; w < n_walk; w++) {
        n_i += win[w].ni; n_ej += win[w].nej; n_sj += win[w].nsj;
        i_ep += (long long)win[w].ni * win[w].nej;
        i_sp += (long long)win[w].ni * win[w].nsj;
    }

    static HostPlan hp[kMaxStreams];
    // plan all sub-batches at once (independent of each other), then pack + enqueue them in turn
    // so that the GPU starts on sub-batch 0 while the host packs sub-batch 1
    const double tp0 = now_s();
    for (int s = 0; s < n_slots; s++) {
        Slot& S = E.slots[s];
        S.w_begin = cut[s]; S.w_end = cut[s + 1];
        S.active = S.w_end > S.w_begin;
    }
    // EP lists as runs of consecutive indices (FDPS lists are leaf cells in Morton order): one pass over the lists finds
    // them — the only time the host reads the EP indices — and only the runs are packed and copied
    const bool use_runs = !direct && E.opt_ep_runs;
    for (int s = 0; s < n_slots; s++) hp[s].use_runs = use_runs;
#pragma omp parallel for schedule(static, 1)      // same team size as the packing loops: no team re-creation
    for (int s = 0; s < n_slots; s++) {
        const Slot& S = E.slots[s];
        if (S.active) plan_batch(win + S.w_begin, S.w_end - S.w_begin, direct, n_slots, hp[s]);
    }
    E.prof.t_plan += now_s() - tp0;
    int first_active = -1, last_active = -1;
    for (int s = 0; s < n_slots; s++)
        if (E.slots[s].active) { if (first_active < 0) first_active = s; last_active = s; }
    E.out_first_slot = first_active;
    int i_base_next = 0;
    if (E.count_only) { E.nb_keys.clear(); E.nb_n_i = 0; }
    for (int s = 0; s < n_slots; s++) {
        Slot& S = E.slots[s];
        if (!S.active) continue;
        int rc;
        if ((rc = grow_arena(S, hp[s].p.bytes)) != PB_OK) return rc;
        if ((rc = grow_out(S, hp[s].p.n_i)) != PB_OK) return rc;
        if ((rc = grow_part(S, hp[s].p.n_part)) != PB_OK) return rc;
        S.emit = E.count_only && E.opt_nb_lists;
        if (S.emit) {                                      // estimate; an overflow re-runs the sub-batch with a larger window
            S.n_pairs_window = 12 * hp[s].p.n_i + 8192;
            if ((rc = grow_pairs(S, S.n_pairs_window)) != PB_OK) return rc;
        }
    }

    // enqueue of one packed sub-batch: its whole input travels in one copy
    double t_enq = 0.0;
    auto enqueue_slot = [&](int s) -> int {
        Slot& S = E.slots[s];
        const double t1 = now_s();
        S.plan = hp[s].p;
        if (hp[s].use_runs) {
            S.plan.n_runs = (size_t)__atomic_load_n(&hp[s].run_cursor, __ATOMIC_ACQUIRE);
            S.plan.bytes = S.plan.off_ide + sizeof(int2) * S.plan.n_runs;
        }
        if (!direct && S.j_epoch != E.j_epoch) {          // once per stream and j publication
            CU(cudaStreamWaitEvent(S.stream, E.ev_j_ready, 0));
            S.j_epoch = E.j_epoch;
        }
        CU(cudaEventRecord(S.ev[0], S.stream));
        CU(cudaMemcpyAsync(S.d_arena, S.h_arena, S.plan.bytes, cudaMemcpyHostToDevice, S.stream));
        CU(cudaEventRecord(S.ev[1], S.stream));
        if (S.emit) {
            S.i_base = i_base_next;
            CU(cudaMemsetAsync(S.d_cursor, 0, sizeof(unsigned int), S.stream));
            CU(cudaMemsetAsync(S.d_pairs, 0xff, sizeof(unsigned long long) * S.n_pairs_window, S.stream));
        }
        i_base_next += (int)S.plan.n_i;
        if (hp[s].use_runs) {
            if (S.plan.n_idx > S.cap_ide_x) {
                CU(cudaStreamSynchronize(S.stream));
                if (S.d_ide_x) CU(cudaFree(S.d_ide_x));
                S.d_ide_x = nullptr; S.cap_ide_x = 0;
                const size_t cap = align_up(S.plan.n_idx + S.plan.n_idx / 2, 4096);
                CU(cudaMalloc(&S.d_ide_x, sizeof(int) * cap));
                S.cap_ide_x = cap;
            }
            CU(launch_expand_runs(S.stream, (const int2*)(S.d_arena + S.plan.off_runtab), (const int2*)(S.d_arena + S.plan.off_ide),
                                  (const Walk*)(S.d_arena + S.plan.off_walks), S.plan.n_walk, S.d_ide_x));
            S.plan.ext_ide = S.d_ide_x;
            E.prof.n_kernel_launch += 1;
        }
        const bool fused = plan_fused(S.plan, S.emit ? &S : nullptr);
        if (fused)                                          // blocks nobody will deliver a chunk for (walks with two empty lists)
            for (const auto& z : hp[s].zero_blocks) memset(S.h_out + z.first, 0, sizeof(ForceOut) * (size_t)z.second);
        CU(launch_plan(S.stream, S.plan, S.d_arena, direct, S.d_part4, S.d_partn, S.d_out, S.h_out, false, S.emit ? &S : nullptr));
        CU(cudaEventRecord(S.ev[2], S.stream));
        if (s == last_active) CU(cudaEventRecord(E.ev_end[E.end_cur], S.stream));
        if (S.emit) {
            CU(sort_pairs(S, (size_t)n_i));
            CU(cudaMemcpyAsync(S.h_cursor, S.d_cursor, sizeof(unsigned int), cudaMemcpyDeviceToHost, S.stream));
            E.prof.n_kernel_launch += 1;
        }
        if (!fused) CU(cudaMemcpyAsync(S.h_out, S.d_out, sizeof(ForceOut) * S.plan.n_i, cudaMemcpyDeviceToHost, S.stream));
        CU(cudaEventRecord(S.ev[3], S.stream));
        E.prof.h2d_bytes += (long long)S.plan.bytes;
        E.prof.d2h_bytes += (long long)(sizeof(ForceOut) * S.plan.n_i);
        E.prof.n_kernel_launch += (S.plan.n_tasks > 0) + (!fused && S.plan.n_iblocks > 0);
        if (E.recording) {
            Recorded r;
            r.plan = S.plan; r.direct = direct; r.slot = s;
            CU(cudaMalloc(&r.d_arena, S.plan.bytes));
            CU(cudaMemcpyAsync(r.d_arena, S.d_arena, S.plan.bytes, cudaMemcpyDeviceToDevice, S.stream));
            if (S.plan.ext_ide && S.plan.n_idx) {              // the expanded EP lists stay with the recording
                CU(cudaMalloc(&r.d_ide_x, sizeof(int) * S.plan.n_idx));
                CU(cudaMemcpyAsync(r.d_ide_x, S.d_ide_x, sizeof(int) * S.plan.n_idx, cudaMemcpyDeviceToDevice, S.stream));
                r.plan.ext_ide = r.d_ide_x;
            }
            E.recs.push_back(r);
        }
        t_enq += now_s() - t1;
        return PB_OK;
    };

    const double tp2 = now_s();
    int rc_enq = PB_OK;
    if (direct || omp_get_max_threads() < 2) {
        // pack and enqueue the sub-batches in turn (the direct mode's j packing parallelises internally)
        for (int s = 0; s < n_slots && rc_enq == PB_OK; s++) {
            Slot& S = E.slots[s];
            if (!S.active) continue;
            pack_batch(win + S.w_begin, direct, Li, Lj, Ls, hp[s], S.h_arena);
            rc_enq = enqueue_slot(s);
        }
    } else {
        // ONE parallel region per dispatch: walks are handed out in order, so the sub-batches complete in order;
        // the calling thread enqueues a sub-batch (copy + kernels) as soon as its last walk is packed and packs
        // walks itself in between, so that enqueueing overlaps the other threads' packing
        std::atomic<int> next_walk{0};
        std::atomic<int> done[kMaxStreams];
        int slot_of_walk_begin[kMaxStreams + 1];
        for (int s = 0; s < n_slots; s++) { done[s].store(0); slot_of_walk_begin[s] = cut[s]; }
        slot_of_walk_begin[n_slots] = n_walk;
        auto pack_one = [&]() -> bool {
            const int w = next_walk.fetch_add(1, std::memory_order_relaxed);
            if (w >= n_walk) return false;
            int s = 0;
            while (w >= slot_of_walk_begin[s + 1]) s++;
            Slot& S = E.slots[s];
            pack_walk(win + S.w_begin, direct, Li, hp[s], S.h_arena, w - S.w_begin);
            done[s].fetch_add(1, std::memory_order_release);
            return true;
        };
#pragma omp parallel
        {
            if (omp_get_thread_num() == 0) {
                int s = 0;
                while (s < n_slots) {
                    Slot& S = E.slots[s];
                    if (!S.active) { s++; continue; }
                    if (done[s].load(std::memory_order_acquire) == S.w_end - S.w_begin) {
                        if (rc_enq == PB_OK) {
                            pack_tail(win + S.w_begin, direct, Lj, Ls, hp[s], S.h_arena);
                            rc_enq = enqueue_slot(s);
                        }
                        s++;
                    } else if (!pack_one()) {
                        _mm_pause();
                    }
                }
            } else {
                while (pack_one()) {}
            }
        }
    }
    if (rc_enq != PB_OK) return rc_enq;
    {
        const double tt = now_s() - tp2;
        E.prof.t_pack += tt - t_enq;
        E.prof.t_enqueue += t_enq;
        E.prof.t_copy -= t_enq;              // enqueue time is not packing time
    }
    for (int s = n_slots; s < kMaxStreams; s++) E.slots[s].active = false;

    E.outstanding = true;
    E.out_count_only = E.count_only;
    E.out_n_walk = n_walk; E.out_n_slots = n_slots;
    E.out_ni.resize(n_walk);
    for (int w = 0; w < n_walk; w++) E.out_ni[w] = win[w].ni;

    E.prof.n_walk += n_walk; E.prof.n_epi += n_i; E.prof.n_epj += n_ej; E.prof.n_spj += n_sj;
    E.prof.n_call += 1; E.prof.n_interaction_ep += i_ep; E.prof.n_interaction_sp += i_sp;
    E.prof.t_copy += now_s() - t0;
    return PB_OK;
}

} // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int pb_abi_version(void) { return PB_ABI_VERSION; }

const char* pb_last_error(void) { return g_err; }

int pb_init(int my_rank, int device) {
    if (E.inited) return PB_OK;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(PB_ERR_NO_DEVICE, "no CUDA device available (%s); libpetar_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    E.rank = my_rank;
    E.device = device >= 0 ? device : my_rank % ndev;        // reference src/force_gpu_cuda.cu:550-553
    if (E.device >= ndev) return fail(PB_ERR_ARG, "device %d out of range (%d devices)", E.device, ndev);
    CU(cudaSetDevice(E.device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, E.device));
    if (prop.major != 10)
        return fail(PB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library contains sm_100a code only",
                    E.device, prop.major, prop.minor);
    CU(cudaStreamCreateWithFlags(&E.s_upload, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&E.ev_j_ready, cudaEventDisableTiming));
    CU(cudaEventCreate(&E.ev_send0));
    CU(cudaEventCreate(&E.ev_send1));
    CU(cudaEventCreate(&E.ev_end[0]));
    CU(cudaEventCreate(&E.ev_end[1]));
    for (int s = 0; s < kMaxStreams; s++) {
        CU(cudaStreamCreateWithFlags(&E.slots[s].stream, cudaStreamNonBlocking));
        for (int k = 0; k < 4; k++) CU(cudaEventCreate(&E.slots[s].ev[k]));
    }
    memset(&E.prof, 0, sizeof(E.prof));
    E.inited = true;
    return PB_OK;
}

void pb_finalize(void) {
    if (!E.inited) return;
    cudaDeviceSynchronize();
    for (auto& r : E.recs) { cudaFree(r.d_arena); cudaFree(r.d_ide_x); }
    E.recs.clear();
    for (int s = 0; s < kMaxStreams; s++) {
        Slot& S = E.slots[s];
        cudaFreeHost(S.h_arena); cudaFree(S.d_arena);
        cudaFreeHost(S.h_out); cudaFree(S.d_out);
        cudaFree(S.d_part4); cudaFree(S.d_partn);
        cudaFree(S.d_pairs); cudaFree(S.d_pairs_sorted); cudaFreeHost(S.h_pairs);
        cudaFree(S.d_cursor); cudaFreeHost(S.h_cursor); cudaFree(S.d_cubtmp); cudaFree(S.d_ide_x);
        for (int k = 0; k < 4; k++) cudaEventDestroy(S.ev[k]);
        cudaStreamDestroy(S.stream);
        S = Slot();
    }
    cudaFree(E.d_epj); cudaFree(E.d_spj); cudaFreeHost(E.h_jstage); cudaFree(E.d_raw); cudaFree(E.d_cellA); cudaFree(E.d_cellB);
    for (auto& r : E.registered) cudaHostUnregister((void*)r.first);
    if (E.d_elem_map) cudaFree(E.d_elem_map);
    if (E.h_elem_map) cudaFreeHost(E.h_elem_map);
    if (E.h_corr) cudaFreeHost(E.h_corr);
    if (E.d_corr) cudaFree(E.d_corr);
    E.h_corr = nullptr; E.d_corr = nullptr; E.cap_corr = 0;
    cudaFree(E.d_cells); cudaFree(E.d_groups); cudaFree(E.d_counts); cudaFree(E.d_overflow); cudaFreeHost(E.h_tstage);
    for (int s = 0; s < kMaxStreams; s++) cudaFree(E.d_walk_scratch[s]);
    cudaFree(E.d_tree_ide); cudaFree(E.d_tree_ids); cudaFree(E.d_tree_off); cudaFree(E.d_tree_caps);
    cudaFreeHost(E.h_counts_p); cudaFreeHost(E.h_over_p);
    cudaFree(E.d_ifirst); cudaFreeHost(E.h_ifirst); cudaFree(E.d_r_walks); cudaFree(E.d_r_goff); cudaFree(E.d_r_epi); cudaFree(E.d_r_out);
    cudaFreeHost(E.h_r_out); cudaFree(E.d_r_iblocks); cudaFree(E.d_r_done); cudaFree(E.d_r_tasks); cudaFree(E.d_r_part4); cudaFree(E.d_r_partn);
    cudaFree(E.d_r_meta); cudaFreeHost(E.h_r_meta); cudaFree(E.d_let_idx); cudaFreeHost(E.h_let_idx); cudaFree(E.d_let_err); cudaFreeHost(E.h_let_err); cudaFree(E.d_raw_let); cudaFreeHost(E.h_let_sp);
    for (int k = 0; k < 8; k++) if (E.ev_tl[k]) cudaEventDestroy(E.ev_tl[k]);
    if (E.ev_count) cudaEventDestroy(E.ev_count);
    if (E.ev_fill) cudaEventDestroy(E.ev_fill);
    cudaEventDestroy(E.ev_j_ready); cudaEventDestroy(E.ev_send0); cudaEventDestroy(E.ev_send1);
    cudaEventDestroy(E.ev_end[0]); cudaEventDestroy(E.ev_end[1]); E.end_prev_valid = false;
    cudaStreamDestroy(E.s_upload);
    const int coords = E.opt_coords, streams = E.opt_streams, jchunk = E.opt_jchunk, nr = E.opt_nr, cull = E.opt_cull, occ = E.opt_occ, lead = E.opt_lead;
    const double eps2 = E.eps2, rcut2 = E.rcut2, G = E.G;
    E = Engine();
    E.opt_coords = coords; E.opt_streams = streams; E.opt_jchunk = jchunk; E.opt_nr = nr; E.opt_cull = cull; E.opt_occ = occ; E.opt_lead = lead;
    E.eps2 = eps2; E.rcut2 = rcut2; E.G = G;
}

int pb_set_params(double eps2, double rcut2, double G) {
    if (!(eps2 >= 0.0) || !(rcut2 >= 0.0)) return fail(PB_ERR_ARG, "pb_set_params: eps2 and rcut2 must be >= 0");
    E.eps2 = eps2; E.rcut2 = rcut2; E.G = G;
    return PB_OK;
}

// Switching raw_upload / raw_result off releases every page-lock taken for them: a caller that wants to free or move a
// registered array switches the option off first (a stale registration would make a later array at the same address look
// page-locked while the driver still maps its old pages).
static void release_registered() {
    if (!E.inited || E.registered.empty()) return;
    cudaDeviceSynchronize();
    for (auto& r : E.registered) cudaHostUnregister((void*)r.first);
    cudaGetLastError();
    E.registered.clear();
}

int pb_set_option(const char* key, long long v) {
    if (!key) return fail(PB_ERR_ARG, "pb_set_option: null key");
    if (!strcmp(key, "coords"))  { if (v < 0 || v > 2) return fail(PB_ERR_ARG, "coords must be 0, 1 or 2"); E.opt_coords = (int)v; return PB_OK; }
    if (!strcmp(key, "streams")) { if (v < 1 || v > kMaxStreams) return fail(PB_ERR_ARG, "streams must be 1..%d", kMaxStreams); E.opt_streams = (int)v; return PB_OK; }
    if (!strcmp(key, "jchunk"))  { if (v < 0 || v > (1 << 20)) return fail(PB_ERR_ARG, "jchunk out of range"); E.opt_jchunk = (int)v; return PB_OK; }
    if (!strcmp(key, "cull"))    { if (v < 0 || v > 1) return fail(PB_ERR_ARG, "cull must be 0 or 1"); E.opt_cull = (int)v; return PB_OK; }
    if (!strcmp(key, "tree_fill")) { if (v < 0 || v > 1) return fail(PB_ERR_ARG, "tree_fill must be 0 or 1"); E.opt_tree_fill = (int)v; return PB_OK; }
    if (!strcmp(key, "min_slot_work")) { if (v < 0) return fail(PB_ERR_ARG, "min_slot_work must be >= 0"); E.opt_min_slot_work = v; return PB_OK; }
    if (!strcmp(key, "tree_streams")) { if (v < 1 || v > kMaxStreams) return fail(PB_ERR_ARG, "tree_streams must be in [1, %d]", kMaxStreams); E.opt_tree_streams = (int)v; return PB_OK; }
    if (!strcmp(key, "tree_spec")) { if (v < 0 || v > 1) return fail(PB_ERR_ARG, "tree_spec must be 0 or 1"); E.opt_tree_spec = (int)v; return PB_OK; }
    if (!strcmp(key, "walk_ctas")) { if (v < 1 || v > 148 * 16) return fail(PB_ERR_ARG, "walk_ctas must be in [1, 2368]"); E.opt_walk_ctas = (int)v; return PB_OK; }
    if (!strcmp(key, "ep_runs")) { if (v < 0 || v > 1) return fail(PB_ERR_ARG, "ep_runs must be 0 or 1"); E.opt_ep_runs = (int)v; return PB_OK; }
    if (!strcmp(key, "walk_compact")) { if (v < 0 || v > 1) return fail(PB_ERR_ARG, "walk_compact must be 0 or 1"); E.opt_walk_compact = (int)v; return PB_OK; }
    if (!strcmp(key, "raw_upload")) { if (v < 0 || v > 1) return fail(PB_ERR_ARG, "raw_upload must be 0 or 1"); E.opt_raw_upload = (int)v; if (!v) release_registered(); return PB_OK; }
    if (!strcmp(key, "ws")) { if (v < 0 || v > 1) return fail(PB_ERR_ARG, "ws must be 0 or 1"); E.opt_ws = (int)v; return PB_OK; }
    if (!strcmp(key, "raw_result")) { if (v < 0 || v > 1) return fail(PB_ERR_ARG, "raw_result must be 0 or 1"); E.opt_raw_result = (int)v; if (!v) release_registered(); return PB_OK; }
    if (!strcmp(key, "fuse_reduce")) { if (v < 0 || v > 1) return fail(PB_ERR_ARG, "fuse_reduce must be 0 or 1"); E.opt_fuse_reduce = (int)v; return PB_OK; }
    if (!strcmp(key, "sp2i")) { if (v < 0 || v > 1) return fail(PB_ERR_ARG, "sp2i must be 0 or 1"); E.opt_sp2i = (int)v; return PB_OK; }
    if (!strcmp(key, "chunk_tile")) { if (v < 0 || v > 1) return fail(PB_ERR_ARG, "chunk_tile must be 0 or 1"); E.opt_chunk_tile = (int)v; return PB_OK; }
    if (!strcmp(key, "nb_lists")) { if (v < 0 || v > 1) return fail(PB_ERR_ARG, "nb_lists must be 0 or 1"); E.opt_nb_lists = (int)v; return PB_OK; }
    if (!strcmp(key, "tree_batch")) { if (v < 1 || v > (1 << 24)) return fail(PB_ERR_ARG, "tree_batch out of range"); E.opt_tree_batch = (int)v; return PB_OK; }
    if (!strcmp(key, "lead"))    { if (v < 0 || v > 15) return fail(PB_ERR_ARG, "lead must be in [0, 15]"); E.opt_lead = (int)v; return PB_OK; }
    if (!strcmp(key, "occupancy")) { if (v < 2 || v > 3) return fail(PB_ERR_ARG, "occupancy must be 2 or 3"); E.opt_occ = (int)v; return PB_OK; }
    if (!strcmp(key, "nr"))      { if (v < 0 || v > 1) return fail(PB_ERR_ARG, "nr must be 0 or 1"); E.opt_nr = (int)v; return PB_OK; }
    return fail(PB_ERR_ARG, "pb_set_option: unknown key '%s'", key);
}

int pb_get_option(const char* key, long long* v) {
    if (!key || !v) return fail(PB_ERR_ARG, "pb_get_option: null argument");
    const struct { const char* k; long long val; } tab[] = {
        {"coords", E.opt_coords}, {"streams", E.opt_streams}, {"jchunk", E.opt_jchunk}, {"cull", E.opt_cull},
        {"tree_fill", E.opt_tree_fill}, {"min_slot_work", E.opt_min_slot_work}, {"tree_streams", E.opt_tree_streams},
        {"tree_spec", E.opt_tree_spec}, {"walk_ctas", E.opt_walk_ctas}, {"nb_lists", E.opt_nb_lists},
        {"tree_batch", E.opt_tree_batch}, {"chunk_tile", E.opt_chunk_tile}, {"ws", E.opt_ws}, {"sp2i", E.opt_sp2i}, {"fuse_reduce", E.opt_fuse_reduce}, {"raw_result", E.opt_raw_result}, {"raw_upload", E.opt_raw_upload}, {"walk_compact", E.opt_walk_compact}, {"ep_runs", E.opt_ep_runs}, {"lead", E.opt_lead}, {"occupancy", E.opt_occ}, {"nr", E.opt_nr}};
    for (const auto& t : tab)
        if (!strcmp(key, t.k)) { *v = t.val; return PB_OK; }
    return fail(PB_ERR_ARG, "pb_get_option: unknown key '%s'", key);
}

int pb_reserve_j(int n_epj, int n_spj, void** d_epj, void** d_spj) {
    int rc = ensure_init();
    if (rc != PB_OK) return rc;
    if (n_epj < 0 || n_spj < 0) return fail(PB_ERR_ARG, "pb_reserve_j: negative size");
    if (E.outstanding) return fail(PB_ERR_PROTOCOL, "pb_reserve_j while a dispatch is outstanding");
    if ((rc = grow_jstore((size_t)n_epj, (size_t)n_spj)) != PB_OK) return rc;
    E.n_epj = n_epj; E.n_spj = n_spj;
    E.j_published = false;
    if (d_epj) *d_epj = E.d_epj;
    if (d_spj) *d_spj = E.d_spj;
    return PB_OK;
}

namespace {
// page-lock [p, p + bytes) once so that it can be the source of an asynchronous DMA; a range registered earlier that
// overlaps a different one (the caller re-allocated its array) is released first.  false: not possible, pack on the host.
bool ensure_registered(const void* ptr, size_t bytes) {
    const char* p = (const char*)ptr;
    if (bytes == 0) return true;
    for (auto& r : E.registered)
        if (p >= r.first && p + bytes <= r.first + r.second) return true;
    for (size_t k = 0; k < E.registered.size();) {
        auto& r = E.registered[k];
        if (p < r.first + r.second && r.first < p + bytes) { cudaHostUnregister((void*)r.first); E.registered.erase(E.registered.begin() + k); }
        else k++;
    }
    if (cudaHostRegister((void*)p, bytes, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return false; }
    E.registered.push_back({p, bytes});
    return true;
}

// the gather kernel of pb_let_gather_epj found an index outside the store (option raw_upload: the list is not read on
// the host): reported by the first call that has synchronised with the device since
int check_let_error() {
    if (E.h_let_err && *E.h_let_err) {
        *E.h_let_err = 0;
        cudaMemset(E.d_let_err, 0, sizeof(int));
        return fail(PB_ERR_ARG, "pb_let_gather_epj: an index was outside the EP store (found on the device)");
    }
    return PB_OK;
}
} // namespace

int pb_upload_j_range(const void* epj, int epj_first, int n_epj, const pb_layout_epj* lepj,
                      const void* spj, int spj_first, int n_spj, const pb_layout_spj* lspj) {
    int rc = ensure_init();
    if (rc != PB_OK) return rc;
    if (n_epj < 0 || n_spj < 0 || epj_first < 0 || spj_first < 0 ||
        (size_t)epj_first + n_epj > E.cap_epj || (size_t)spj_first + n_spj > E.cap_spj)
        return fail(PB_ERR_ARG, "pb_upload_j_range: range outside the reserved store");
    if ((n_epj && (!epj || !lepj)) || (n_spj && (!spj || !lspj))) return fail(PB_ERR_ARG, "pb_upload_j_range: null input");
    const double t0 = now_s();
    const size_t nf4 = 2 * (size_t)n_epj + 4 * (size_t)n_spj;
    if (E.opt_raw_upload) {
        // the arrays travel as they are (page-locked once) and are packed on the device: no host core touches them.
        // They must stay unchanged until the step's first kernels have run (FDPS keeps epj_sorted / spj_sorted for the
        // whole force calculation).
        const size_t be = n_epj ? (size_t)n_epj * lepj->stride : 0, bs = n_spj ? (size_t)n_spj * lspj->stride : 0;
        if (ensure_registered(epj, be) && ensure_registered(spj, bs)) {
            const size_t o_s = align_up(be, 256);
            if (o_s + bs > E.cap_raw) {
                CU(cudaStreamSynchronize(E.s_upload));
                if (E.d_raw) CU(cudaFree(E.d_raw));
                E.d_raw = nullptr; E.cap_raw = 0;
                const size_t cap = align_up(o_s + bs + (o_s + bs) / 4, 1 << 20);
                CU(cudaMalloc(&E.d_raw, cap));
                E.cap_raw = cap;
            }
            E.end_prev_valid = false;
            CU(cudaEventRecord(E.ev_send0, E.s_upload));
            if (n_epj) {
                CU(cudaMemcpyAsync(E.d_raw, epj, be, cudaMemcpyHostToDevice, E.s_upload));
                CU(launch_pack_epj(E.s_upload, E.d_raw, lepj->stride, lepj->off_pos, lepj->off_mass, lepj->off_rsearch, n_epj, E.d_epj + 2 * (size_t)epj_first));
            }
            if (n_spj) {
                CU(cudaMemcpyAsync(E.d_raw + o_s, spj, bs, cudaMemcpyHostToDevice, E.s_upload));
                CU(launch_pack_spj(E.s_upload, E.d_raw + o_s, lspj->stride, lspj->off_pos, lspj->off_mass, lspj->off_quad, lspj->has_quad, n_spj, E.d_spj + 4 * (size_t)spj_first));
            }
            CU(cudaEventRecord(E.ev_send1, E.s_upload));
            E.send_timed = true;
            E.prof.h2d_bytes += (long long)(be + bs);
            E.prof.n_kernel_launch += (n_epj > 0) + (n_spj > 0);
            E.prof.t_copy += now_s() - t0;
            return PB_OK;
        }
    }
    if ((rc = grow_jstage(nf4)) != PB_OK) return rc;
    CU(cudaStreamSynchronize(E.s_upload));               // staging buffer free again
    float4* he = E.h_jstage;
    float4* hs = E.h_jstage + 2 * (size_t)n_epj;
    E.end_prev_valid = false;                             // a new tree step: no gap to the previous dispatch
    CU(cudaEventRecord(E.ev_send0, E.s_upload));
    // packed and copied in pieces: the DMA of one piece runs while the host packs the next
    constexpr int kPiece = 1 << 18;
    const char* be = (const char*)epj;
    for (int o = 0; o < n_epj; o += kPiece) {
        const int m = std::min(kPiece, n_epj - o);
        pack_epj(be + (size_t)o * lepj->stride, m, *lepj, he + 2 * (size_t)o);
        CU(cudaMemcpyAsync(E.d_epj + 2 * ((size_t)epj_first + o), he + 2 * (size_t)o, (size_t)PB_EPJ_DEV_BYTES * m, cudaMemcpyHostToDevice, E.s_upload));
    }
    const char* bs = (const char*)spj;
    for (int o = 0; o < n_spj; o += kPiece) {
        const int m = std::min(kPiece, n_spj - o);
        pack_spj(bs + (size_t)o * lspj->stride, m, *lspj, hs + 4 * (size_t)o);
        CU(cudaMemcpyAsync(E.d_spj + 4 * ((size_t)spj_first + o), hs + 4 * (size_t)o, (size_t)PB_SPJ_DEV_BYTES * m, cudaMemcpyHostToDevice, E.s_upload));
    }
    E.prof.t_copy += now_s() - t0;
    CU(cudaEventRecord(E.ev_send1, E.s_upload));
    E.send_timed = true;
    E.prof.h2d_bytes += (long long)(nf4 * sizeof(float4));
    return PB_OK;
}

int pb_publish_j(void* cuda_stream) {
    int rc = ensure_init();
    if (rc != PB_OK) return rc;
    {
        // order the upload stream after whatever the caller's stream has queued (e.g. a NCCL
        // all-to-all that writes LET entries straight into the store)
        cudaEvent_t ev;
        CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CU(cudaEventRecord(ev, (cudaStream_t)cuda_stream));
        CU(cudaStreamWaitEvent(E.s_upload, ev, 0));
        CU(cudaEventDestroy(ev));
    }
    CU(cudaEventRecord(E.ev_j_ready, E.s_upload)); E.j_epoch++;
    E.j_published = true;
    return PB_OK;
}

int pb_upload_j(const void* epj, int n_epj, const pb_layout_epj* lepj,
                const void* spj, int n_spj, const pb_layout_spj* lspj) {
    int rc = pb_reserve_j(n_epj, n_spj, nullptr, nullptr);
    if (rc != PB_OK) return rc;
    if ((rc = pb_upload_j_range(epj, 0, n_epj, lepj, spj, 0, n_spj, lspj)) != PB_OK) return rc;
    CU(cudaEventRecord(E.ev_j_ready, E.s_upload)); E.j_epoch++;
    E.j_published = true;
    return PB_OK;
}

int pb_let_gather_epj(const int* idx, int n, void* d_out32) {
    int rc = ensure_init();
    if (rc != PB_OK) return rc;
    if (n < 0 || (n && (!idx || !d_out32))) return fail(PB_ERR_ARG, "pb_let_gather_epj: bad argument");
    if (n == 0) return PB_OK;
    if (!E.d_epj) return fail(PB_ERR_PROTOCOL, "pb_let_gather_epj before the j store was reserved");
    if ((size_t)n > E.cap_let_idx) {
        CU(cudaStreamSynchronize(E.s_upload));
        if (E.d_let_idx) CU(cudaFree(E.d_let_idx));
        if (E.h_let_idx) CU(cudaFreeHost(E.h_let_idx));
        E.cap_let_idx = (size_t)n + n / 4 + 1024;
        CU(cudaMalloc(&E.d_let_idx, sizeof(int) * E.cap_let_idx));
        CU(cudaMallocHost(&E.h_let_idx, sizeof(int) * E.cap_let_idx));
    }
    if (!E.d_let_err) {
        CU(cudaMalloc(&E.d_let_err, sizeof(int)));
        CU(cudaMallocHost(&E.h_let_err, sizeof(int)));
        CU(cudaMemset(E.d_let_err, 0, sizeof(int)));
        *E.h_let_err = 0;
    }
    if ((rc = check_let_error()) != PB_OK) return rc;
    // on the upload stream: behind the local particles' copy, ahead of whatever the caller orders after pb_stream_wait_upload
    if (E.opt_raw_upload && ensure_registered(idx, sizeof(int) * (size_t)n)) {
        // the list travels from where it lies (page-locked once); the gather kernel checks the indices
        CU(cudaMemcpyAsync(E.d_let_idx, idx, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, E.s_upload));
    } else {
        for (int k = 0; k < n; k++)
            if (idx[k] < 0 || idx[k] >= E.n_epj) return fail(PB_ERR_ARG, "pb_let_gather_epj: index %d outside the EP store (%d entries)", idx[k], E.n_epj);
        CU(cudaStreamSynchronize(E.s_upload));             // the staging copy of the previous call has been read
        memcpy(E.h_let_idx, idx, sizeof(int) * (size_t)n);
        CU(cudaMemcpyAsync(E.d_let_idx, E.h_let_idx, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, E.s_upload));
    }
    CU(launch_gather_epj(E.s_upload, E.d_epj, E.n_epj, E.d_let_idx, n, (float4*)d_out32, E.d_let_err));
    CU(cudaMemcpyAsync(E.h_let_err, E.d_let_err, sizeof(int), cudaMemcpyDeviceToHost, E.s_upload));
    E.prof.h2d_bytes += (long long)(sizeof(int) * (size_t)n);
    E.prof.n_kernel_launch += 1;
    return PB_OK;
}

int pb_let_pack_spj(const void* spj, int n, const pb_layout_spj* l, void* d_out64) {
    int rc = ensure_init();
    if (rc != PB_OK) return rc;
    if (n < 0 || (n && (!spj || !l || !d_out64))) return fail(PB_ERR_ARG, "pb_let_pack_spj: bad argument");
    if (n == 0) return PB_OK;
    const size_t bs = (size_t)n * l->stride;
    if (E.opt_raw_upload && ensure_registered(spj, bs)) {
        if (bs > E.cap_raw_let) {
            CU(cudaStreamSynchronize(E.s_upload));
            if (E.d_raw_let) CU(cudaFree(E.d_raw_let));
            E.d_raw_let = nullptr; E.cap_raw_let = 0;
            const size_t cap = align_up(bs + bs / 4, 1 << 20);
            CU(cudaMalloc(&E.d_raw_let, cap));
            E.cap_raw_let = cap;
        }
        CU(cudaMemcpyAsync(E.d_raw_let, spj, bs, cudaMemcpyHostToDevice, E.s_upload));
        CU(launch_pack_spj(E.s_upload, E.d_raw_let, l->stride, l->off_pos, l->off_mass, l->off_quad, l->has_quad, n, (float4*)d_out64));
        E.prof.h2d_bytes += (long long)bs;
        E.prof.n_kernel_launch += 1;
        return PB_OK;
    }
    CU(cudaStreamSynchronize(E.s_upload));                 // the staging buffer's previous content has been read
    if ((size_t)n > E.cap_let_sp) {
        if (E.h_let_sp) CU(cudaFreeHost(E.h_let_sp));
        E.cap_let_sp = (size_t)n + n / 4 + 1024;
        CU(cudaMallocHost(&E.h_let_sp, 64 * E.cap_let_sp));
    }
    pack_spj(spj, n, *l, E.h_let_sp);
    CU(cudaMemcpyAsync(d_out64, E.h_let_sp, 64 * (size_t)n, cudaMemcpyHostToDevice, E.s_upload));
    E.prof.h2d_bytes += (long long)(64 * (size_t)n);
    return PB_OK;
}

int pb_stream_wait_upload(void* cuda_stream) {
    int rc = ensure_init();
    if (rc != PB_OK) return rc;
    cudaEvent_t ev;
    CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CU(cudaEventRecord(ev, E.s_upload));
    CU(cudaStreamWaitEvent((cudaStream_t)cuda_stream, ev, 0));
    CU(cudaEventDestroy(ev));
    return PB_OK;
}

int pb_pack_epj_host(const void* epj, int n, const pb_layout_epj* l, void* out32) {
    if (n < 0 || (n && (!epj || !l || !out32))) return fail(PB_ERR_ARG, "pb_pack_epj_host: bad argument");
    if (n) pack_epj(epj, n, *l, (float4*)out32);
    return PB_OK;
}

int pb_pack_epj_host_indexed(const void* epj, const long long* idx, int n, const pb_layout_epj* l, void* out32) {
    if (n < 0 || (n && (!epj || !idx || !l || !out32))) return fail(PB_ERR_ARG, "pb_pack_epj_host_indexed: bad argument");
    const char* base = (const char*)epj;
    float4* out = (float4*)out32;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        const char* p = base + (size_t)idx[i] * l->stride;
        float4 a, b;
        split(ld(p, l->off_pos, 0), a.x, b.x);
        split(ld(p, l->off_pos, 1), a.y, b.y);
        split(ld(p, l->off_pos, 2), a.z, b.z);
        a.w = (float)ld(p, l->off_mass);
        b.w = (float)ld(p, l->off_rsearch);
        out[2 * (size_t)i]     = a;
        out[2 * (size_t)i + 1] = b;
    }
    return PB_OK;
}

int pb_pack_spj_host(const void* spj, int n, const pb_layout_spj* l, void* out64) {
    if (n < 0 || (n && (!spj || !l || !out64))) return fail(PB_ERR_ARG, "pb_pack_spj_host: bad argument");
    if (n) pack_spj(spj, n, *l, (float4*)out64);
    return PB_OK;
}

int pb_dispatch_index(int n_walk,
                      const void* const* epi, const int* n_epi, const pb_layout_epi* lepi,
                      const int* const* id_epj, const int* n_epj,
                      const int* const* id_spj, const int* n_spj) {
    int rc = ensure_init();
    if (rc != PB_OK) return rc;
    if (n_walk < 0) return fail(PB_ERR_ARG, "pb_dispatch_index: n_walk < 0");
    if (n_walk && (!epi || !n_epi || !lepi || !id_epj || !n_epj || !id_spj || !n_spj))
        return fail(PB_ERR_ARG, "pb_dispatch_index: null argument");
    std::vector<WalkIn> win(n_walk);
    for (int w = 0; w < n_walk; w++) {
        if (n_epi[w] < 0 || n_epj[w] < 0 || n_spj[w] < 0) return fail(PB_ERR_ARG, "pb_dispatch_index: negative count in walk %d", w);
        win[w] = {epi[w], n_epi[w], id_epj[w], n_epj[w], id_spj[w], n_spj[w], nullptr, nullptr};
    }
    return dispatch_common(n_walk, win.data(), false, *lepi, nullptr, nullptr);
}

int pb_dispatch_count_index(int n_walk,
                            const void* const* epi, const int* n_epi, const pb_layout_epi* lepi,
                            const int* const* id_epj, const int* n_epj) {
    int rc = ensure_init();
    if (rc != PB_OK) return rc;
    if (n_walk < 0) return fail(PB_ERR_ARG, "pb_dispatch_count_index: n_walk < 0");
    if (n_walk && (!epi || !n_epi || !lepi || !id_epj || !n_epj)) return fail(PB_ERR_ARG, "pb_dispatch_count_index: null argument");
    std::vector<WalkIn> win(n_walk);
    for (int w = 0; w < n_walk; w++) {
        if (n_epi[w] < 0 || n_epj[w] < 0) return fail(PB_ERR_ARG, "pb_dispatch_count_index: negative count in walk %d", w);
        win[w] = {epi[w], n_epi[w], id_epj[w], n_epj[w], nullptr, 0, nullptr, nullptr};
    }
    E.count_only = true;
    rc = dispatch_common(n_walk, win.data(), false, *lepi, nullptr, nullptr);
    E.count_only = false;
    return rc;
}

int pb_dispatch_direct(int n_walk,
                       const void* const* epi, const int* n_epi, const pb_layout_epi* lepi,
                       const void* const* epj, const int* n_epj, const pb_layout_epj* lepj,
                       const void* const* spj, const int* n_spj, const pb_layout_spj* lspj) {
    int rc = ensure_init();
    if (rc != PB_OK) return rc;
    if (n_walk < 0) return fail(PB_ERR_ARG, "pb_dispatch_direct: n_walk < 0");
    if (n_walk && (!epi || !n_epi || !lepi || !epj || !n_epj || !lepj || !spj || !n_spj || !lspj))
        return fail(PB_ERR_ARG, "pb_dispatch_direct: null argument");
    std::vector<WalkIn> win(n_walk);
    for (int w = 0; w < n_walk; w++) {
        if (n_epi[w] < 0 || n_epj[w] < 0 || n_spj[w] < 0) return fail(PB_ERR_ARG, "pb_dispatch_direct: negative count in walk %d", w);
        win[w] = {epi[w], n_epi[w], nullptr, n_epj[w], nullptr, n_spj[w], epj[w], spj[w]};
    }
    return dispatch_common(n_walk, win.data(), true, *lepi, lepj, lspj);
}

int pb_retrieve(int n_walk, const int* ni, void* const* force, const pb_layout_force* L) {
    if (!E.inited || !E.outstanding) return fail(PB_ERR_PROTOCOL, "pb_retrieve without an outstanding dispatch");
    if (n_walk != E.out_n_walk) return fail(PB_ERR_ARG, "pb_retrieve: n_walk %d != dispatched %d", n_walk, E.out_n_walk);
    if (n_walk && (!ni || !force || !L)) return fail(PB_ERR_ARG, "pb_retrieve: null argument");
    for (int w = 0; w < n_walk; w++)
        if (ni[w] != E.out_ni[w]) return fail(PB_ERR_ARG, "pb_retrieve: ni[%d]=%d != dispatched %d", w, ni[w], E.out_ni[w]);
    size_t n_i_dispatch = 0;
    for (int w = 0; w < n_walk; w++) n_i_dispatch += (size_t)ni[w];
    const bool plain = !E.out_count_only && L && L->stride == sizeof(ForceOut) && L->off_acc == 0 && L->off_pot == 24 && L->off_nngb == 32;
    for (int s = 0; s < E.out_n_slots; s++) {
        Slot& S = E.slots[s];
        if (!S.active) continue;
        CU(cudaEventSynchronize(S.ev[3]));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, S.ev[0], S.ev[1])); E.prof.t_send += 1e-3 * ms;
        CU(cudaEventElapsedTime(&ms, S.ev[1], S.ev[2])); E.prof.t_calc += 1e-3 * ms;
        CU(cudaEventElapsedTime(&ms, S.ev[2], S.ev[3])); E.prof.t_recv += 1e-3 * ms;
        if (S.emit) {                                     // the sorted pairs travel while the counts are scattered
            int rc = collect_pairs_begin(S, n_i_dispatch);
            if (rc != PB_OK) return rc;
        }
        const double t0 = now_s();
        const int nw = S.w_end - S.w_begin;
        std::vector<size_t> first(nw + 1, 0);
        for (int k = 0; k < nw; k++) first[k + 1] = first[k] + (size_t)ni[S.w_begin + k];
#pragma omp parallel for schedule(static)
        for (int k = 0; k < nw; k++) {
            const int w = S.w_begin + k;
            const ForceOut* src = S.h_out + first[k];
            char* dst = (char*)force[w];
            if (plain) {
                memcpy(dst, src, sizeof(ForceOut) * (size_t)ni[w]);
            } else {
                for (int i = 0; i < ni[w]; i++) {
                    char* q = dst + (size_t)i * L->stride;
                    if (!E.out_count_only) {              // the neighbour-search functor assigns n_ngb only
                        memcpy(q + L->off_acc, &src[i].ax, 24);
                        memcpy(q + L->off_pot, &src[i].pot, 8);
                    }
                    memcpy(q + L->off_nngb, &src[i].n_ngb, 8);
                }
            }
        }
        E.prof.t_copy += now_s() - t0;
        E.prof.t_unpack += now_s() - t0;
        if (S.emit) {
            int rc = collect_pairs_end(S);
            if (rc != PB_OK) return rc;
            E.nb_n_i += (long long)S.plan.n_i;
        }
        S.active = false;
    }
    if (E.send_timed) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, E.ev_send0, E.ev_send1) == cudaSuccess) E.prof.t_send += 1e-3 * ms;
        E.send_timed = false;
    }
    if (E.out_first_slot >= 0) {
        // GPU idle since the previous walk group's last kernel (same tree step only)
        float ms = 0.f;
        if (E.end_prev_valid && cudaEventElapsedTime(&ms, E.ev_end[E.end_cur ^ 1], E.slots[E.out_first_slot].ev[1]) == cudaSuccess && ms > 0.f)
            E.prof.t_gap += 1e-3 * ms;
        E.end_prev_valid = true;
        E.end_cur ^= 1;
    }
    E.outstanding = false;
    return check_let_error();
}

// ---- direct-sum field query ---------------------------------------------------
int pb_debug_plan(int n_walk, const int* n_epi, const int* n_epj, const int* n_spj, int n_streams_active,
                  int* walks_out, int* tasks_out, int cap_tasks, int* iblocks_out, int cap_iblocks,
                  int* n_iblocks, long long* n_part) {
    if (n_walk < 0 || (n_walk && (!n_epi || !n_epj || !n_spj))) return fail(PB_ERR_ARG, "pb_debug_plan: bad argument");
    static const int dummy = 0;
    std::vector<WalkIn> win(n_walk);
    for (int w = 0; w < n_walk; w++) win[w] = {&dummy, n_epi[w], &dummy, n_epj[w], &dummy, n_spj[w], nullptr, nullptr};
    HostPlan hp;
    plan_batch(win.data(), n_walk, false, std::max(1, n_streams_active), hp);
    for (int w = 0; w < n_walk && walks_out; w++) {
        const Walk& W = hp.walks[w];
        int* o = walks_out + 6 * w;
        o[0] = W.i_off; o[1] = W.ni; o[2] = W.ej_off; o[3] = W.nej; o[4] = W.sj_off; o[5] = W.nsj;
    }
    for (int t = 0; t < (int)hp.tasks.size() && t < cap_tasks && tasks_out; t++) {
        const Task& T = hp.tasks[t];
        int* o = tasks_out + 10 * t;
        o[0] = T.walk; o[1] = T.i_first; o[2] = T.nib; o[3] = T.jsplit; o[4] = T.kind; o[5] = T.j_begin; o[6] = T.j_count; o[7] = T.part_base;
        o[8] = T.blk0; o[9] = T.n_chunks;
    }
    for (int b = 0; b < (int)hp.iblocks.size() && b < cap_iblocks && iblocks_out; b++) {
        const IBlock& B = hp.iblocks[b];
        int* o = iblocks_out + 5 * b;
        o[0] = B.part_base; o[1] = B.n_chunks; o[2] = B.stride; o[3] = B.out_off; o[4] = B.n_valid;
    }
    if (n_iblocks) *n_iblocks = (int)hp.iblocks.size();
    if (n_part) *n_part = (long long)hp.p.n_part;
    return (int)hp.tasks.size();
}

int pb_retrieve_neighbors(long long* n_pairs, int* nb_off, int* nb_idx, long long cap) {
    if (!E.inited) return fail(PB_ERR_PROTOCOL, "pb_retrieve_neighbors before any dispatch");
    if (E.outstanding) return fail(PB_ERR_PROTOCOL, "pb_retrieve_neighbors while a dispatch is outstanding");
    const long long n = (long long)E.nb_keys.size();
    if (n_pairs) *n_pairs = n;
    if (nb_off) {
        // keys are sorted by i: list i ends where the first key of a larger i starts
        const unsigned long long* k0 = E.nb_keys.data();
#pragma omp parallel for schedule(static)
        for (long long i = 0; i <= E.nb_n_i; i++)
            nb_off[i] = (int)(std::lower_bound(k0, k0 + n, (unsigned long long)i << 32) - k0);
    }
    if (nb_idx) {
        if (cap < n) return fail(PB_ERR_ARG, "pb_retrieve_neighbors: %lld pairs do not fit in %lld entries", n, cap);
#pragma omp parallel for schedule(static)
        for (long long k = 0; k < n; k++) nb_idx[k] = (int)(E.nb_keys[k] & 0xffffffffull);
    }
    return PB_OK;
}

int pb_field_at_points(const double* x, const double* y, const double* z, int n_points,
                       const void* ptcl, int n_ptcl, size_t stride, size_t off_pos, size_t off_mass,
                       double G, double* ax, double* ay, double* az, double* pot) {
    int rc = ensure_init();
    if (rc != PB_OK) return rc;
    if (E.outstanding) return fail(PB_ERR_PROTOCOL, "pb_field_at_points while a dispatch is outstanding");
    if (n_points < 0 || n_ptcl < 0 || (n_points && (!x || !y || !z)) || (n_ptcl && !ptcl))
        return fail(PB_ERR_ARG, "pb_field_at_points: bad argument");
    // j = particles with mass > 0 (as CalcForcePPSimd filters, reference src/soft_force.hpp:293-296)
    struct PJ { double mass; double pos[3]; double rs; };
    std::vector<PJ> pj;
    pj.reserve((size_t)n_ptcl);
    for (int i = 0; i < n_ptcl; i++) {
        const char* p = (const char*)ptcl + (size_t)i * stride;
        const double m = ld(p, off_mass);
        if (m > 0.0) pj.push_back({m, {ld(p, off_pos, 0), ld(p, off_pos, 1), ld(p, off_pos, 2)}, 0.0});
    }
    const pb_layout_epj lj = {sizeof(PJ), offsetof(PJ, pos), offsetof(PJ, mass), offsetof(PJ, rs)};
    if ((rc = pb_upload_j(pj.data(), (int)pj.size(), &lj, nullptr, 0, nullptr)) != PB_OK) return rc;

    struct PI { double pos[3]; double rs; };
    std::vector<PI> pi((size_t)n_points);
    for (int i = 0; i < n_points; i++) pi[i] = {{x[i], y[i], z[i]}, 0.0};
    const pb_layout_epi li = {sizeof(PI), offsetof(PI, pos), offsetof(PI, rs)};
    const int chunk = 512;
    const int n_walk = (n_points + chunk - 1) / chunk;
    std::vector<WalkIn> win(n_walk);
    std::vector<int> ni(n_walk);
    std::vector<ForceOut> out((size_t)n_points);
    std::vector<void*> fptr(n_walk);
    for (int w = 0; w < n_walk; w++) {
        ni[w] = std::min(chunk, n_points - w * chunk);
        win[w] = {&pi[(size_t)w * chunk], ni[w], nullptr, (int)pj.size(), nullptr, 0, nullptr, nullptr};   // dense EP list
        fptr[w] = &out[(size_t)w * chunk];
    }
    // eps = 0, no cutoff: CalcForcePP semantics (src/soft_force.hpp:209-211, 311-312)
    const double eps2 = E.eps2, rcut2 = E.rcut2, G0 = E.G;
    E.eps2 = 0.0; E.rcut2 = 0.0; E.G = G;
    const pb_layout_force lf = {sizeof(ForceOut), 0, 24, 32};
    const int batch = 32;                                     // 16k points per dispatch bounds the partial-sum buffer
    for (int w0 = 0; w0 < n_walk && rc == PB_OK; w0 += batch) {
        const int nw = std::min(batch, n_walk - w0);
        rc = dispatch_common(nw, win.data() + w0, false, li, nullptr, nullptr);
        if (rc == PB_OK) rc = pb_retrieve(nw, ni.data() + w0, fptr.data() + w0, &lf);
    }
    E.eps2 = eps2; E.rcut2 = rcut2; E.G = G0;
    // the j store now holds the query's particle set (mass > 0 only, r_search = 0): a dispatch or tree step that follows
    // without a fresh pb_upload_j must fail with PB_ERR_PROTOCOL instead of running against it
    E.j_published = false;
    if (rc != PB_OK) return rc;
    for (int i = 0; i < n_points; i++) {
        if (ax) ax[i] = out[i].ax;
        if (ay) ay[i] = out[i].ay;
        if (az) az[i] = out[i].az;
        if (pot) pot[i] = out[i].pot;
    }
    return PB_OK;
}

// ---- changeover correction ---------------------------------------------------
namespace {
inline long long ldi(const char* p, size_t off) { long long v; memcpy(&v, p + off, sizeof(v)); return v; }
inline CorrJ corr_load(const char* p, const pb_layout_corr& L) {
    CorrJ j;
    j.x = ld(p, L.off_pos, 0); j.y = ld(p, L.off_pos, 1); j.z = ld(p, L.off_pos, 2);
    j.mass = ld(p, L.off_mass); j.r_in = ld(p, L.off_r_in); j.r_out = ld(p, L.off_r_out);
    j.mass_bk = ld(p, L.off_mass_backup); j.status = ld(p, L.off_status); j.id = ldi(p, L.off_id);
    return j;
}
}

int pb_correct_changeover(int n_i, void* ptcl_i, const pb_layout_corr* li,
                          int n_j, const void* ptcl_j, const pb_layout_corr* lj,
                          const int* nb_off, const int* nb_idx, const pb_corr_params* prm) {
    int rc = ensure_init();
    if (rc != PB_OK) return rc;
    if (n_i < 0 || n_j < 0 || !prm || (n_i && (!ptcl_i || !li || !nb_off)) || (n_j && (!ptcl_j || !lj)))
        return fail(PB_ERR_ARG, "pb_correct_changeover: bad argument");
    if (n_i == 0) return PB_OK;
    const long long n_nb = nb_off[n_i];
    if (nb_off[0] != 0 || n_nb < 0 || (n_nb && !nb_idx)) return fail(PB_ERR_ARG, "pb_correct_changeover: bad neighbour offsets");
    int bad = 0;
#pragma omp parallel for reduction(| : bad)
    for (int i = 0; i < n_i; i++) {
        if (nb_off[i + 1] < nb_off[i]) bad |= 1;
        else for (int k = nb_off[i]; k < nb_off[i + 1]; k++) if (nb_idx[k] < 0 || nb_idx[k] >= n_j) bad |= 2;
    }
    if (bad) return fail(PB_ERR_ARG, "pb_correct_changeover: neighbour list %s", (bad & 1) ? "offsets decrease" : "index outside ptcl_j");

    // one staging buffer: [CorrI n_i][CorrJ n_j][nb_off n_i+1][nb_idx n_nb] -> device; [CorrOut n_i] back
    const double t0 = now_s();
    const size_t o_i = 0;
    const size_t o_j = align_up(o_i + sizeof(CorrI) * (size_t)n_i, 256);
    const size_t o_off = align_up(o_j + sizeof(CorrJ) * (size_t)n_j, 256);
    const size_t o_idx = align_up(o_off + sizeof(int) * ((size_t)n_i + 1), 256);
    const size_t o_out = align_up(o_idx + sizeof(int) * (size_t)n_nb, 256);
    const size_t bytes = align_up(o_out + sizeof(CorrOut) * (size_t)n_i, 256);
    if (bytes > E.cap_corr) {
        if (E.h_corr) CU(cudaFreeHost(E.h_corr));
        if (E.d_corr) CU(cudaFree(E.d_corr));
        E.h_corr = nullptr; E.d_corr = nullptr; E.cap_corr = 0;
        const size_t cap = bytes + bytes / 8;
        CU(cudaMallocHost(&E.h_corr, cap));
        CU(cudaMalloc(&E.d_corr, cap));
        E.cap_corr = cap;
    }
    CorrI* hi = (CorrI*)(E.h_corr + o_i);
    CorrJ* hj = (CorrJ*)(E.h_corr + o_j);
    const char* bi = (const char*)ptcl_i;
    const char* bj = (const char*)ptcl_j;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n_i; i++) {
        const char* p = bi + (size_t)i * li->stride;
        CorrI I;
        I.p = corr_load(p, *li);
        I.ax = ld(p, li->off_acc, 0); I.ay = ld(p, li->off_acc, 1); I.az = ld(p, li->off_acc, 2);
        I.pot_tot = ld(p, li->off_pot_tot); I.pot_soft = ld(p, li->off_pot_soft);
        hi[i] = I;
    }
#pragma omp parallel for schedule(static)
    for (int j = 0; j < n_j; j++) hj[j] = corr_load(bj + (size_t)j * lj->stride, *lj);
    memcpy(E.h_corr + o_off, nb_off, sizeof(int) * ((size_t)n_i + 1));
    if (n_nb) memcpy(E.h_corr + o_idx, nb_idx, sizeof(int) * (size_t)n_nb);
    E.prof.t_copy += now_s() - t0;

    cudaStream_t st = E.slots[0].stream;
    CU(cudaMemcpyAsync(E.d_corr, E.h_corr, o_out, cudaMemcpyHostToDevice, st));
    CorrParams P;
    P.eps2 = prm->eps2; P.r_out = prm->r_out; P.G = prm->G; P.status_no_cm = prm->status_no_cm; P.replay_fp32 = prm->replay_fp32 ? 1 : 0;
    CU(launch_corr(st, n_i, (const CorrI*)(E.d_corr + o_i), (const CorrJ*)(E.d_corr + o_j), (const int*)(E.d_corr + o_off),
                   (const int*)(E.d_corr + o_idx), (CorrOut*)(E.d_corr + o_out), P));
    CU(cudaMemcpyAsync(E.h_corr + o_out, E.d_corr + o_out, sizeof(CorrOut) * (size_t)n_i, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    E.prof.n_kernel_launch += 1;
    E.prof.h2d_bytes += (long long)o_out;
    E.prof.d2h_bytes += (long long)(sizeof(CorrOut) * (size_t)n_i);

    const double t1 = now_s();
    const CorrOut* ho = (const CorrOut*)(E.h_corr + o_out);
    char* wi = (char*)ptcl_i;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n_i; i++) {
        char* p = wi + (size_t)i * li->stride;
        memcpy(p + li->off_acc, &ho[i].ax, 24);
        memcpy(p + li->off_pot_tot, &ho[i].pot_tot, 8);
        memcpy(p + li->off_pot_soft, &ho[i].pot_soft, 8);
    }
    E.prof.t_copy += now_s() - t1;
    return PB_OK;
}

// ---- device-side interaction lists --------------------------------------------
namespace {
constexpr int kWalkCap  = 32768;       // frontier capacity per warp (cells of one tree level an i-group touches)
#define kWalkCtas (E.opt_walk_ctas)

int ensure_walk_scratch(int slot) {
    // two frontier buffers of kWalkCap cells per warp of the walk launches actually configured (option "walk_ctas")
    static int scratch_ctas[kMaxStreams] = {0};
    if (!E.d_walk_scratch[slot]) scratch_ctas[slot] = 0;
    if (scratch_ctas[slot] < kWalkCtas) {
        if (E.d_walk_scratch[slot]) { CU(cudaDeviceSynchronize()); CU(cudaFree(E.d_walk_scratch[slot])); E.d_walk_scratch[slot] = nullptr; }
        CU(cudaMalloc(&E.d_walk_scratch[slot], sizeof(int) * (size_t)kWalkCtas * 4 * 2 * kWalkCap));
        scratch_ctas[slot] = kWalkCtas;
    }
    if (!E.d_overflow) { CU(cudaMalloc(&E.d_overflow, 2 * sizeof(int))); CU(cudaMemset(E.d_overflow, 0, 2 * sizeof(int))); }   // [0] frontier, [1] list reservation
    return PB_OK;
}

// wait for a slot's batch and scatter its results (contiguous, group order)
int tree_finish_slot(Slot& S, char* force, const pb_layout_force& L, size_t i_first) {
    CU(cudaEventSynchronize(S.ev[3]));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, S.ev[0], S.ev[1])); E.prof.t_send += 1e-3 * ms;
    CU(cudaEventElapsedTime(&ms, S.ev[1], S.ev[2])); E.prof.t_calc += 1e-3 * ms;
    CU(cudaEventElapsedTime(&ms, S.ev[2], S.ev[3])); E.prof.t_recv += 1e-3 * ms;
    const double t0 = now_s();
    const bool plain = L.stride == sizeof(ForceOut) && L.off_acc == 0 && L.off_pot == 24 && L.off_nngb == 32;
    const size_t n = S.plan.n_i;
    if (plain) {
        memcpy(force + i_first * sizeof(ForceOut), S.h_out, sizeof(ForceOut) * n);
    } else {
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)n; i++) {
            char* q = force + (i_first + (size_t)i) * L.stride;
            memcpy(q + L.off_acc, &S.h_out[i].ax, 24);
            memcpy(q + L.off_pot, &S.h_out[i].pot, 8);
            memcpy(q + L.off_nngb, &S.h_out[i].n_ngb, 8);
        }
    }
    E.prof.t_copy += now_s() - t0;
    E.prof.t_unpack += now_s() - t0;
    S.active = false;
    return PB_OK;
}
} // namespace

namespace {
// the two walk passes, on the compact records (default) or on the fp64 cells alone
cudaError_t walk_count(cudaStream_t st, int g0, int ng, double theta_inv2, int slot) {
    if (E.opt_walk_compact)
        return launch_walk_c(st, false, E.d_cells, E.d_cellA, E.d_cellB, E.d_groups, g0, ng, theta_inv2, E.coord_max, E.d_counts, nullptr, nullptr, nullptr,
                             E.d_walk_scratch[slot], kWalkCap, kWalkCtas, E.d_overflow, E.has_elem_map ? E.d_elem_map : nullptr, E.n_cells, nullptr);
    return launch_walk_count(st, E.d_cells, E.d_groups, g0, ng, theta_inv2, E.d_counts, E.d_walk_scratch[slot], kWalkCap, kWalkCtas, E.d_overflow,
                             E.has_elem_map ? E.d_elem_map : nullptr, E.n_cells);
}
cudaError_t walk_fill(cudaStream_t st, int g0, int ng, double theta_inv2, int slot, int n_ctas, const int2* caps, int2* counts) {
    if (E.opt_walk_compact)
        return launch_walk_c(st, true, E.d_cells, E.d_cellA, E.d_cellB, E.d_groups, g0, ng, theta_inv2, E.coord_max, counts, E.d_tree_off, E.d_tree_ide, E.d_tree_ids,
                             E.d_walk_scratch[slot], kWalkCap, n_ctas, E.d_overflow, E.has_elem_map ? E.d_elem_map : nullptr, E.n_cells, caps);
    return launch_walk_fill(st, E.d_cells, E.d_groups, g0, ng, theta_inv2, E.d_tree_off, E.d_tree_ide, E.d_tree_ids,
                            E.d_walk_scratch[slot], kWalkCap, n_ctas, E.d_overflow, E.has_elem_map ? E.d_elem_map : nullptr, E.n_cells, caps, counts);
}
} // namespace

int pb_tree_stage(int n_cells, int n_groups, pb_tree_cell** cells, pb_tree_group** groups) {
    int rc = ensure_init();
    if (rc != PB_OK) return rc;
    if (n_cells < 0 || n_groups < 0 || !cells || !groups) return fail(PB_ERR_ARG, "pb_tree_stage: bad argument");
    const size_t bc = sizeof(pb_tree_cell) * (size_t)n_cells, bg = sizeof(pb_tree_group) * (size_t)n_groups;
    if (bc + bg > E.cap_tstage) {
        CU(cudaStreamSynchronize(E.s_upload));             // a previous tree may still be travelling from the old buffer
        if (E.h_tstage) CU(cudaFreeHost(E.h_tstage));
        E.h_tstage = nullptr; E.cap_tstage = 0;
        const size_t cap = bc + bg + (bc + bg) / 4 + 4096;
        CU(cudaMallocHost(&E.h_tstage, cap));
        E.cap_tstage = cap;
    }
    *cells = (pb_tree_cell*)E.h_tstage;
    *groups = (pb_tree_group*)(E.h_tstage + bc);
    return PB_OK;
}

int pb_tree_upload(const pb_tree_cell* cells, int n_cells, const pb_tree_group* groups, int n_groups, double theta) {
    return pb_tree_upload_let(cells, n_cells, groups, n_groups, theta, nullptr, 0);
}

int pb_tree_upload_let(const pb_tree_cell* cells, int n_cells, const pb_tree_group* groups, int n_groups, double theta,
                       const int* elem_map, int n_elem) {
    int rc = ensure_init();
    if (rc != PB_OK) return rc;
    if (n_cells < 0 || n_groups < 0 || n_elem < 0 || (n_cells && !cells) || (n_groups && !groups) || !(theta >= 0.0) || (n_elem && !elem_map))
        return fail(PB_ERR_ARG, "pb_tree_upload: bad argument");
    if (elem_map && n_cells && (long long)cells[0].first + cells[0].n > n_elem)
        return fail(PB_ERR_ARG, "pb_tree_upload_let: the root cell spans %lld elements, elem_map has %d", (long long)cells[0].first + cells[0].n, n_elem);
    if (E.outstanding) return fail(PB_ERR_PROTOCOL, "pb_tree_upload while a dispatch is outstanding");
    static const bool trace = getenv("PETAR_B200_TRACE") != nullptr;      // host time of this call's phases, printed once (8th call)
    static int trace_calls = 0;
    double tt[8]; int nt_ = 0;
    tt[nt_++] = now_s();
    CU(cudaDeviceSynchronize());
    tt[nt_++] = now_s();
    if ((size_t)n_cells > E.cap_cells) {
        if (E.d_cells) CU(cudaFree(E.d_cells));
        E.cap_cells = (size_t)n_cells + n_cells / 4 + 1024;
        CU(cudaMalloc(&E.d_cells, E.cap_cells * sizeof(pb_tree_cell)));
    }
    if ((size_t)n_groups > E.cap_groups) {
        if (E.d_groups) CU(cudaFree(E.d_groups));
        E.cap_groups = (size_t)n_groups + n_groups / 4 + 1024;
        CU(cudaMalloc(&E.d_groups, E.cap_groups * sizeof(pb_tree_group)));
    }
    if ((size_t)n_groups > E.cap_counts) {
        if (E.d_counts) CU(cudaFree(E.d_counts));
        E.cap_counts = (size_t)n_groups + n_groups / 4 + 1024;
        CU(cudaMalloc(&E.d_counts, E.cap_counts * sizeof(int2)));
        if (E.d_tree_off) { CU(cudaFree(E.d_tree_off)); E.d_tree_off = nullptr; }
    }
    // stage through pinned memory with all host threads, then one async DMA on the upload stream
    // (a pageable cudaMemcpy of the ~50 MB tree would cost more than the walk itself)
    const size_t bc = sizeof(pb_tree_cell) * (size_t)n_cells, bg = sizeof(pb_tree_group) * (size_t)n_groups;
    if (bc + bg > E.cap_tstage) {
        if (E.h_tstage) CU(cudaFreeHost(E.h_tstage));
        E.cap_tstage = bc + bg + (bc + bg) / 4;
        CU(cudaMallocHost(&E.h_tstage, E.cap_tstage));
    }
    const bool staged = E.h_tstage && (const char*)cells == E.h_tstage && (const char*)groups == E.h_tstage + bc;   // pb_tree_stage buffers
    if (!staged) {
        const size_t chunk = 1 << 20;
        const long long nchunk = (long long)((bc + chunk - 1) / chunk);
#pragma omp parallel for schedule(static)
        for (long long k = 0; k < nchunk; k++)
            memcpy(E.h_tstage + (size_t)k * chunk, (const char*)cells + (size_t)k * chunk, std::min(chunk, bc - (size_t)k * chunk));
        memcpy(E.h_tstage + bc, groups, bg);
    }
    CU(cudaMemcpyAsync(E.d_cells, E.h_tstage, bc, cudaMemcpyHostToDevice, E.s_upload));
    CU(cudaMemcpyAsync(E.d_groups, E.h_tstage + bc, bg, cudaMemcpyHostToDevice, E.s_upload));
    tt[nt_++] = now_s();
    E.has_elem_map = elem_map != nullptr && n_elem > 0;
    if (E.has_elem_map) {
        if ((size_t)n_elem > E.cap_elem_map) {
            if (E.d_elem_map) CU(cudaFree(E.d_elem_map));
            if (E.h_elem_map) CU(cudaFreeHost(E.h_elem_map));
            E.cap_elem_map = (size_t)n_elem + n_elem / 4 + 1024;
            CU(cudaMalloc(&E.d_elem_map, sizeof(int) * E.cap_elem_map));
            CU(cudaMallocHost(&E.h_elem_map, sizeof(int) * E.cap_elem_map));
        }
        if (E.opt_raw_upload && ensure_registered(elem_map, sizeof(int) * (size_t)n_elem))
            CU(cudaMemcpyAsync(E.d_elem_map, elem_map, sizeof(int) * (size_t)n_elem, cudaMemcpyHostToDevice, E.s_upload));   // from where it lies
        else {
            memcpy(E.h_elem_map, elem_map, sizeof(int) * (size_t)n_elem);
            CU(cudaMemcpyAsync(E.d_elem_map, E.h_elem_map, sizeof(int) * (size_t)n_elem, cudaMemcpyHostToDevice, E.s_upload));
        }
        E.prof.h2d_bytes += (long long)(sizeof(int) * (size_t)n_elem);
    }
    tt[nt_++] = now_s();
    E.n_cells = n_cells; E.n_groups = n_groups; E.theta = theta;
    E.grp_n.resize(n_groups);
    for (int g = 0; g < n_groups; g++) E.grp_n[g] = groups[g].n;
    {   // first i-particle of every group in the step's output order (= group order), for the device-resident step
        if ((size_t)n_groups > E.cap_ifirst) {
            if (E.d_ifirst) CU(cudaFree(E.d_ifirst));
            if (E.h_ifirst) CU(cudaFreeHost(E.h_ifirst));
            E.cap_ifirst = (size_t)n_groups + n_groups / 4 + 1024;
            CU(cudaMalloc(&E.d_ifirst, sizeof(int) * E.cap_ifirst));
            CU(cudaMallocHost(&E.h_ifirst, sizeof(int) * E.cap_ifirst));
        }
        long long acc = 0; int nblk = 0;
        for (int g = 0; g < n_groups; g++) { E.h_ifirst[g] = (int)acc; acc += E.grp_n[g]; nblk += (E.grp_n[g] + 31) / 32; }
        if (acc >= (1ll << 31)) return fail(PB_ERR_ARG, "pb_tree_upload: more than 2^31 i-particles");
        E.r_n_i = acc; E.r_n_iblk = nblk;
        if (n_groups) CU(cudaMemcpyAsync(E.d_ifirst, E.h_ifirst, sizeof(int) * (size_t)n_groups, cudaMemcpyHostToDevice, E.s_upload));
    }
    tt[nt_++] = now_s();
    CU(cudaEventRecord(E.ev_j_ready, E.s_upload)); E.j_epoch++;        // dispatch streams wait for j AND tree
    E.prof.h2d_bytes += (long long)(sizeof(pb_tree_cell) * (size_t)n_cells + sizeof(pb_tree_group) * (size_t)n_groups + sizeof(int) * (size_t)n_groups);
    // pass 1 of the device walk starts right away — it needs the tree only, so it runs while the
    // host is still packing this step's j (pb_upload_j): list lengths of every group
    E.count_pending = false;
    if (n_groups > 0) {
        int rc2;
        if ((rc2 = ensure_walk_scratch(0)) != PB_OK) return rc2;
        if (!E.ev_count) CU(cudaEventCreateWithFlags(&E.ev_count, cudaEventDisableTiming));
        if ((size_t)n_groups > E.cap_counts_p) {
            if (E.h_counts_p) CU(cudaFreeHost(E.h_counts_p));
            E.cap_counts_p = E.cap_counts;
            CU(cudaMallocHost(&E.h_counts_p, sizeof(int2) * E.cap_counts_p));
        }
        if (!E.h_over_p) CU(cudaMallocHost(&E.h_over_p, 2 * sizeof(int)));
        const double theta_inv2 = theta > 0.0 ? 1.0 / (theta * theta) : 1e300;
        cudaStream_t s0 = E.slots[0].stream;
        CU(cudaStreamWaitEvent(s0, E.ev_j_ready, 0));
        if (E.opt_walk_compact) {
            if ((size_t)n_cells > E.cap_cellAB) {
                if (E.d_cellA) CU(cudaFree(E.d_cellA));
                if (E.d_cellB) CU(cudaFree(E.d_cellB));
                E.cap_cellAB = E.cap_cells;
                CU(cudaMalloc(&E.d_cellA, 64 * E.cap_cellAB));
                CU(cudaMalloc(&E.d_cellB, 48 * E.cap_cellAB));
            }
            // S = the largest |coordinate| any box of the tree can have: bounds the fp32 rounding of the compact records
            double smax = 0.0;
            if (n_cells > 0)
                for (int k = 0; k < 3; k++)
                    smax = std::max(smax, std::max(std::max(std::fabs(cells[0].out_lo[k]), std::fabs(cells[0].out_hi[k])),
                                                   std::max(std::fabs(cells[0].in_lo[k]), std::fabs(cells[0].in_hi[k]))));
            E.coord_max = 2.0 * smax;
            CU(launch_compact_cells(s0, E.d_cells, n_cells, E.d_cellA, E.d_cellB));
            E.prof.n_kernel_launch += 1;
        }
        for (int k = 0; k < 8; k++) if (!E.ev_tl[k]) CU(cudaEventCreate(&E.ev_tl[k]));
        CU(cudaEventRecord(E.ev_tl[0], s0));                 // tree on the device, walk starts
        E.tl_valid = false;
        E.spec_pending = false;
        CU(cudaMemsetAsync(E.d_overflow, 0, 2 * sizeof(int), s0));         // a frontier overflow of an earlier tree must not stick
        if (E.opt_tree_spec && (int)E.prev_counts.size() == n_groups) {
            // Speculative single pass: the lists of consecutive tree steps have nearly the same lengths, so space is
            // reserved from the previous step's (+12.5 % + 64) and the walk fills the lists right away — while the host
            // is still packing j.  The true lengths come back with it; a list that outgrew its reservation is detected
            // (nothing is written past it) and pb_tree_force then falls back to the exact two-pass fill.
            E.h_tree_off.resize(n_groups);
            std::vector<int2> caps(n_groups);
            size_t tot_e = 0, tot_s = 0;
            for (int g = 0; g < n_groups; g++) {
                const size_t ce = align_up((size_t)E.prev_counts[g].x + E.prev_counts[g].x / 8 + 64, 4);
                const size_t cs = align_up((size_t)E.prev_counts[g].y + E.prev_counts[g].y / 8 + 64, 4);
                E.h_tree_off[g] = make_int2((int)tot_e, (int)tot_s);
                caps[g] = make_int2((int)ce, (int)cs);
                tot_e += ce; tot_s += cs;
            }
            if (tot_e < (1ull << 31) && tot_s < (1ull << 31)) {
                if (tot_e > E.cap_tree_ide) { if (E.d_tree_ide) CU(cudaFree(E.d_tree_ide)); E.cap_tree_ide = tot_e + tot_e / 8 + 4096; CU(cudaMalloc(&E.d_tree_ide, sizeof(int) * E.cap_tree_ide)); }
                if (tot_s > E.cap_tree_ids) { if (E.d_tree_ids) CU(cudaFree(E.d_tree_ids)); E.cap_tree_ids = tot_s + tot_s / 8 + 4096; CU(cudaMalloc(&E.d_tree_ids, sizeof(int) * E.cap_tree_ids)); }
                if (!E.d_tree_off) CU(cudaMalloc(&E.d_tree_off, sizeof(int2) * E.cap_counts));
                if ((size_t)n_groups > E.cap_tree_caps) {
                    if (E.d_tree_caps) CU(cudaFree(E.d_tree_caps));
                    E.cap_tree_caps = E.cap_counts;
                    CU(cudaMalloc(&E.d_tree_caps, sizeof(int2) * E.cap_tree_caps));
                }
                CU(cudaMemcpyAsync(E.d_tree_off, E.h_tree_off.data(), sizeof(int2) * (size_t)n_groups, cudaMemcpyHostToDevice, s0));
                CU(cudaMemcpyAsync(E.d_tree_caps, caps.data(), sizeof(int2) * (size_t)n_groups, cudaMemcpyHostToDevice, s0));
                CU(cudaMemsetAsync(E.d_overflow + 1, 0, sizeof(int), s0));
                CU(walk_fill(s0, 0, n_groups, theta_inv2, 0, kWalkCtas, E.d_tree_caps, E.d_counts));
                E.spec_pending = true;
            }
        }
        if (!E.spec_pending)
            CU(walk_count(s0, 0, n_groups, theta_inv2, 0));
        CU(cudaEventRecord(E.ev_tl[1], s0));                 // walk (count pass, or speculative list fill) done
        CU(cudaMemcpyAsync(E.h_counts_p, E.d_counts, sizeof(int2) * (size_t)n_groups, cudaMemcpyDeviceToHost, s0));
        CU(cudaMemcpyAsync(E.h_over_p, E.d_overflow, 2 * sizeof(int), cudaMemcpyDeviceToHost, s0));
        CU(cudaEventRecord(E.ev_count, s0));
        E.count_pending = true;
        E.prof.n_kernel_launch += 1;
        E.prof.d2h_bytes += (long long)(sizeof(int2) * (size_t)n_groups);
    }
    tt[nt_++] = now_s();
    if (trace && ++trace_calls == 8)
        fprintf(stderr, "petar_b200 trace: pb_tree_upload_let rank %d [ms]: device_sync %.3f, tree copies enqueued %.3f, elem_map %.3f, group sizes %.3f, walk enqueue %.3f (cells %d groups %d elems %d)\n",
                E.rank, 1e3 * (tt[1] - tt[0]), 1e3 * (tt[2] - tt[1]), 1e3 * (tt[3] - tt[2]), 1e3 * (tt[4] - tt[3]), 1e3 * (tt[5] - tt[4]), n_cells, n_groups, n_elem);
    return PB_OK;
}

int pb_tree_force(const void* epi, const pb_layout_epi* lepi, void* force, const pb_layout_force* lforce) {
    int rc = ensure_init();
    if (rc != PB_OK) return rc;
    if (E.outstanding) return fail(PB_ERR_PROTOCOL, "pb_tree_force while a dispatch is outstanding");
    if (!E.j_published) return fail(PB_ERR_PROTOCOL, "pb_tree_force before pb_upload_j");
    if (E.n_groups == 0) return PB_OK;
    if (!E.d_cells) return fail(PB_ERR_PROTOCOL, "pb_tree_force before pb_tree_upload");
    if (!epi || !lepi || !force || !lforce) return fail(PB_ERR_ARG, "pb_tree_force: null argument");
    if (E.n_cells > E.n_spj) return fail(PB_ERR_PROTOCOL, "pb_tree_force: the SP store (%d) must hold one superparticle per cell (%d)", E.n_spj, E.n_cells);
    const double theta_inv2 = E.theta > 0.0 ? 1.0 / (E.theta * E.theta) : 1e300;
    // the batches of a whole step are queued at once here, so four streams already keep the GPU full (measured:
    // eight cost 3 ms per step)
    const int n_slots = std::max(1, std::min(E.opt_streams, E.opt_tree_streams));
    if ((rc = ensure_walk_scratch(0)) != PB_OK) return rc;
    if (!E.ev_fill) CU(cudaEventCreateWithFlags(&E.ev_fill, cudaEventDisableTiming));

    // pass 1 (list lengths) was launched by pb_tree_upload; collect it
    cudaStream_t s0 = E.slots[0].stream;
    if (E.count_pending) {
        CU(cudaEventSynchronize(E.ev_count));
        E.count_pending = false;
        if (*E.h_over_p) return fail(PB_ERR_ARG, "pb_tree_force: tree-walk frontier exceeded %d cells per level", kWalkCap);
        E.h_counts.assign(E.h_counts_p, E.h_counts_p + E.n_groups);
    } else if ((int)E.h_counts.size() != E.n_groups) {
        return fail(PB_ERR_PROTOCOL, "pb_tree_force: call pb_tree_upload for this step first");
    }

    const bool lists_ready = E.spec_pending && E.h_over_p[1] == 0;     // the speculative pass already wrote them
    E.spec_pending = false;
    E.prev_counts = E.h_counts;
    if (lists_ready) {
        CU(cudaStreamWaitEvent(s0, E.ev_j_ready, 0));
        CU(cudaEventRecord(E.ev_fill, s0));
    }
    // pass 2: one launch writes every group's lists into a step-wide device buffer
    const bool split_fill = !lists_ready && E.opt_tree_fill == 1;   // fill each batch's lists on the batch's own stream, just ahead of its force launch
    if (!lists_ready) {
        E.h_tree_off.resize(E.n_groups);
        size_t tot_e = 0, tot_s = 0;
        for (int g = 0; g < E.n_groups; g++) {
            E.h_tree_off[g] = make_int2((int)tot_e, (int)tot_s);
            tot_e += align_up((size_t)E.h_counts[g].x, 4);
            tot_s += align_up((size_t)E.h_counts[g].y, 4);
        }
        if (tot_e >= (1ull << 31) || tot_s >= (1ull << 31)) return fail(PB_ERR_ARG, "pb_tree_force: more than 2^31 list entries in one step");
        if (tot_e > E.cap_tree_ide) { if (E.d_tree_ide) CU(cudaFree(E.d_tree_ide)); E.cap_tree_ide = tot_e + tot_e / 8 + 4096; CU(cudaMalloc(&E.d_tree_ide, sizeof(int) * E.cap_tree_ide)); }
        if (tot_s > E.cap_tree_ids) { if (E.d_tree_ids) CU(cudaFree(E.d_tree_ids)); E.cap_tree_ids = tot_s + tot_s / 8 + 4096; CU(cudaMalloc(&E.d_tree_ids, sizeof(int) * E.cap_tree_ids)); }
        if (!E.d_tree_off) CU(cudaMalloc(&E.d_tree_off, sizeof(int2) * E.cap_counts));   // freed whenever d_counts is re-sized
        CU(cudaMemcpyAsync(E.d_tree_off, E.h_tree_off.data(), sizeof(int2) * (size_t)E.n_groups, cudaMemcpyHostToDevice, s0));
        if (!split_fill) {
            CU(walk_fill(s0, 0, E.n_groups, theta_inv2, 0, kWalkCtas, nullptr, nullptr));
            E.prof.n_kernel_launch += 1;
        } else {
            for (int s = 0; s < n_slots; s++) if ((rc = ensure_walk_scratch(s)) != PB_OK) return rc;
        }
        CU(cudaStreamWaitEvent(s0, E.ev_j_ready, 0));          // j may have been published after the tree (the recommended order)
        CU(cudaEventRecord(E.ev_fill, s0));                    // offsets (and, unless split, the lists) and this step's j are in place
    }
    // pass 3: per batch of groups — plan tasks from the counts, force, reduce
    const char* ebase = (const char*)epi;
    const double tp0 = now_s();
    std::vector<WalkIn> win(E.n_groups);
    std::vector<size_t> grp_i_first(E.n_groups + 1, 0);
    long long n_i = 0, n_ej = 0, n_sj = 0, i_ep = 0, i_sp = 0;
    for (int g = 0; g < E.n_groups; g++) {
        win[g] = {ebase + grp_i_first[g] * lepi->stride, E.grp_n[g], &g_devlist_marker, E.h_counts[g].x,
                  &g_devlist_marker, E.h_counts[g].y, nullptr, nullptr};
        grp_i_first[g + 1] = grp_i_first[g] + (size_t)E.grp_n[g];
        n_i += E.grp_n[g]; n_ej += E.h_counts[g].x; n_sj += E.h_counts[g].y;
        i_ep += (long long)E.grp_n[g] * E.h_counts[g].x; i_sp += (long long)E.grp_n[g] * E.h_counts[g].y;
    }
    const int n_batches = (E.n_groups + E.opt_tree_batch - 1) / E.opt_tree_batch;
    static std::vector<HostPlan> plans;
    if ((int)plans.size() < n_batches) plans.resize(n_batches);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < n_batches; b++) {                      // all batches planned at once, in parallel
        const int g0 = b * E.opt_tree_batch;
        plan_batch(win.data() + g0, std::min(E.opt_tree_batch, E.n_groups - g0), false, n_slots, plans[b], E.h_tree_off.data() + g0);
        plans[b].p.ext_ide = E.d_tree_ide; plans[b].p.ext_ids = E.d_tree_ids;
    }
    E.prof.t_plan += now_s() - tp0;
    E.prof.t_copy += now_s() - tp0;
    size_t slot_i_first[kMaxStreams] = {0};
    for (int b = 0; b < n_batches; b++) {
        const int g0 = b * E.opt_tree_batch;
        const int nb = std::min(E.opt_tree_batch, E.n_groups - g0);
        const int s = b % n_slots;
        Slot& S = E.slots[s];
        if (S.active && (rc = tree_finish_slot(S, (char*)force, *lforce, slot_i_first[s])) != PB_OK) return rc;
        const double t0 = now_s();
        HostPlan& hp = plans[b];
        if ((rc = grow_arena(S, hp.p.bytes)) != PB_OK) return rc;
        if ((rc = grow_out(S, hp.p.n_i)) != PB_OK) return rc;
        if ((rc = grow_part(S, hp.p.n_part)) != PB_OK) return rc;
        pack_batch(win.data() + g0, false, *lepi, nullptr, nullptr, hp, S.h_arena);
        S.plan = hp.p;
        E.prof.t_pack += now_s() - t0;
        E.prof.t_copy += now_s() - t0;
        // only tables + i-particles cross PCIe; the index sections of the arena are filled in place
        const size_t h2d = S.plan.off_ide;
        CU(cudaStreamWaitEvent(S.stream, E.ev_fill, 0));
        if (split_fill) {
            CU(walk_fill(S.stream, g0, nb, theta_inv2, s, std::min(kWalkCtas, (nb + 3) / 4), nullptr, nullptr));
            E.prof.n_kernel_launch += 1;
        }
        CU(cudaEventRecord(S.ev[0], S.stream));
        CU(cudaMemcpyAsync(S.d_arena, S.h_arena, h2d, cudaMemcpyHostToDevice, S.stream));
        CU(cudaEventRecord(S.ev[1], S.stream));
        CU(launch_plan(S.stream, S.plan, S.d_arena, false, S.d_part4, S.d_partn, S.d_out, nullptr));
        CU(cudaEventRecord(S.ev[2], S.stream));
        CU(cudaMemcpyAsync(S.h_out, S.d_out, sizeof(ForceOut) * S.plan.n_i, cudaMemcpyDeviceToHost, S.stream));
        CU(cudaEventRecord(S.ev[3], S.stream));
        E.prof.h2d_bytes += (long long)h2d;
        E.prof.d2h_bytes += (long long)(sizeof(ForceOut) * S.plan.n_i);
        E.prof.n_kernel_launch += (S.plan.n_tasks > 0) + (S.plan.n_iblocks > 0);
        S.active = true;
        slot_i_first[s] = grp_i_first[g0];
        (void)nb;
    }
    for (int s = 0; s < n_slots; s++)
        if (E.slots[s].active && (rc = tree_finish_slot(E.slots[s], (char*)force, *lforce, slot_i_first[s])) != PB_OK) return rc;
    E.tree_last_batches = n_batches;
    E.prof.n_walk += E.n_groups; E.prof.n_epi += n_i; E.prof.n_epj += n_ej; E.prof.n_spj += n_sj;
    E.prof.n_call += n_batches; E.prof.n_interaction_ep += i_ep; E.prof.n_interaction_sp += i_sp;
    return PB_OK;
}

// ---- device-resident tree step ---------------------------------------------------------------------------------
// Everything between "tree and particles are on the device" and "forces are back" runs on the GPU: list building
// (pb_walk.cu), i-particle preparation and task planning (pb_plan.cu), one persistent force launch, the reduction.
// From the second tree step on nothing on the host waits in the middle: list space, task and partial-sum buffers are
// reserved from the previous step's sizes (with a margin), the true sizes come back with the forces, and a step whose
// reservation was too small is detected on the device (nothing runs past a reservation) and repeated exactly.
namespace {
int resident_run(void* force, const pb_layout_force& L, bool exact) {
    cudaStream_t s0 = E.slots[0].stream;
    const int ng = E.n_groups;
    const double theta_inv2 = E.theta > 0.0 ? 1.0 / (E.theta * E.theta) : 1e300;
    int U = E.opt_jchunk > 0 ? E.opt_jchunk : 4096;       // j per warp and task: the persistent launch balances itself, bigger tasks cost less (swept 1024..8192)
    U = (int)align_up((size_t)U, kTileJ);
    const int Us = std::max(kTileJ, U / 2);
    const int coords = E.opt_coords, i_f4 = coords == 2 ? 3 : 2;
    int rc;

    // per-group and per-particle buffers
    if ((size_t)ng > E.cap_r_groups) {
        if (E.d_r_walks) CU(cudaFree(E.d_r_walks));
        if (E.d_r_goff) CU(cudaFree(E.d_r_goff));
        E.cap_r_groups = (size_t)ng + ng / 4 + 1024;
        CU(cudaMalloc(&E.d_r_walks, sizeof(Walk) * E.cap_r_groups));
        CU(cudaMalloc(&E.d_r_goff, sizeof(int3) * E.cap_r_groups));
    }
    if ((size_t)E.r_n_i > E.cap_r_i) {
        if (E.d_r_epi) CU(cudaFree(E.d_r_epi));
        if (E.d_r_out) CU(cudaFree(E.d_r_out));
        if (E.h_r_out) CU(cudaFreeHost(E.h_r_out));
        if (E.d_r_iblocks) CU(cudaFree(E.d_r_iblocks));
        E.cap_r_i = (size_t)E.r_n_i + E.r_n_i / 4 + 4096;
        CU(cudaMalloc(&E.d_r_epi, sizeof(float4) * 3 * E.cap_r_i));
        CU(cudaMalloc(&E.d_r_out, sizeof(ForceOut) * E.cap_r_i));
        CU(cudaMallocHost(&E.h_r_out, sizeof(ForceOut) * E.cap_r_i));
        CU(cudaMalloc(&E.d_r_iblocks, sizeof(IBlock) * (E.cap_r_i / 32 + E.cap_r_groups + 1024)));
        if (E.d_r_done) CU(cudaFree(E.d_r_done));
        CU(cudaMalloc(&E.d_r_done, sizeof(int) * (E.cap_r_i / 32 + E.cap_r_groups + 1024)));
    }
    if ((size_t)E.r_n_iblk > E.cap_r_i / 32 + E.cap_r_groups + 1024) return fail(PB_ERR_ARG, "pb_tree_force_resident: i-block table too small");
    if (!E.d_r_meta) { CU(cudaMalloc(&E.d_r_meta, 8 * sizeof(int))); CU(cudaMallocHost(&E.h_r_meta, 8 * sizeof(int))); }
    if (!E.ev_fill) CU(cudaEventCreateWithFlags(&E.ev_fill, cudaEventDisableTiming));

    long long want_tasks, want_part;
    const int2* d_caps = nullptr;
    if (exact) {
        // list lengths are on the host: exact list offsets, a second walk pass writes the lists, exact buffer sizes
        E.h_tree_off.resize(ng);
        size_t tot_e = 0, tot_s = 0;
        for (int g = 0; g < ng; g++) {
            E.h_tree_off[g] = make_int2((int)tot_e, (int)tot_s);
            tot_e += align_up((size_t)E.h_counts[g].x, 4);
            tot_s += align_up((size_t)E.h_counts[g].y, 4);
        }
        if (tot_e >= (1ull << 31) || tot_s >= (1ull << 31)) return fail(PB_ERR_ARG, "pb_tree_force_resident: more than 2^31 list entries in one step");
        if (tot_e > E.cap_tree_ide) { if (E.d_tree_ide) CU(cudaFree(E.d_tree_ide)); E.cap_tree_ide = tot_e + tot_e / 8 + 4096; CU(cudaMalloc(&E.d_tree_ide, sizeof(int) * E.cap_tree_ide)); }
        if (tot_s > E.cap_tree_ids) { if (E.d_tree_ids) CU(cudaFree(E.d_tree_ids)); E.cap_tree_ids = tot_s + tot_s / 8 + 4096; CU(cudaMalloc(&E.d_tree_ids, sizeof(int) * E.cap_tree_ids)); }
        if (!E.d_tree_off) CU(cudaMalloc(&E.d_tree_off, sizeof(int2) * E.cap_counts));
        CU(cudaMemcpyAsync(E.d_tree_off, E.h_tree_off.data(), sizeof(int2) * (size_t)ng, cudaMemcpyHostToDevice, s0));
        if ((rc = ensure_walk_scratch(0)) != PB_OK) return rc;
        CU(cudaMemsetAsync(E.d_overflow, 0, 2 * sizeof(int), s0));
        CU(walk_fill(s0, 0, ng, theta_inv2, 0, kWalkCtas, nullptr, E.d_counts));
        E.prof.n_kernel_launch += 1;
        long long nt, np, nb;
        plan_sizes_host(E.grp_n.data(), E.h_counts.data(), ng, U, Us, &nt, &np, &nb);
        want_tasks = nt + 64; want_part = np + 4096;
    } else {
        want_tasks = E.r_prev_tasks + E.r_prev_tasks / 4 + 4096;
        want_part = E.r_prev_part + E.r_prev_part / 4 + 65536;
        d_caps = E.d_tree_caps;
    }
    if (want_part >= (1ll << 31)) return fail(PB_ERR_ARG, "pb_tree_force_resident: more than 2^31 partial sums in one step");
    if ((size_t)want_tasks > E.cap_r_tasks) {
        if (E.d_r_tasks) CU(cudaFree(E.d_r_tasks));
        E.cap_r_tasks = (size_t)want_tasks + want_tasks / 8;
        CU(cudaMalloc(&E.d_r_tasks, sizeof(Task) * E.cap_r_tasks));
    }
    if ((size_t)want_part > E.cap_r_part) {
        if (E.d_r_part4) CU(cudaFree(E.d_r_part4));
        if (E.d_r_partn) CU(cudaFree(E.d_r_partn));
        E.cap_r_part = (size_t)want_part + want_part / 8;
        CU(cudaMalloc(&E.d_r_part4, sizeof(double4) * E.cap_r_part));
        CU(cudaMalloc(&E.d_r_partn, sizeof(int) * E.cap_r_part));
    }

    // fused reduction: the force kernel reduces finished i-blocks and writes the forces into page-locked host memory itself
    const bool fuse = E.opt_ws && E.opt_fuse_reduce;
    const bool plain = L.stride == sizeof(ForceOut) && L.off_acc == 0 && L.off_pot == 24 && L.off_nngb == 32;
    bool direct = false;                                   // ... into the caller's own array (option raw_result)
    ForceOut* out_fused = nullptr;
    if (fuse) {
        out_fused = E.h_r_out;
        if (E.opt_raw_result && plain && ensure_registered(force, sizeof(ForceOut) * (size_t)E.r_n_i)) {
            void* dp = nullptr;
            if (cudaHostGetDevicePointer(&dp, force, 0) == cudaSuccess && dp) { out_fused = (ForceOut*)dp; direct = true; }
            else cudaGetLastError();
        }
    }

    CU(cudaStreamWaitEvent(s0, E.ev_j_ready, 0));          // this step's j (local upload + whatever a collective wrote) is in place
    CU(cudaEventRecord(E.ev_tl[2], s0));
    CU(launch_iprep(s0, E.d_groups, ng, E.d_ifirst, E.d_counts, E.d_tree_off, E.d_epj, E.d_r_walks, E.d_r_epi, i_f4, coords, E.opt_cull));
    CU(launch_devplan(s0, E.d_groups, ng, E.d_ifirst, E.d_counts, U, Us, E.d_r_goff, E.d_r_meta,
                   (int)std::min<size_t>(E.cap_r_tasks, 0x7fffffff), (long long)E.cap_r_part, E.d_r_tasks, E.d_r_iblocks, d_caps, out_fused));
    CU(cudaEventRecord(E.ev_tl[3], s0));
    Plan pl; pl.coords = coords; pl.i_f4 = i_f4; pl.count_only = 0;
    Params prm = make_params(pl, nullptr);
    prm.meta = E.d_r_meta;
    if (fuse) {
        CU(cudaMemsetAsync(E.d_r_done, 0, sizeof(int) * (size_t)E.r_n_iblk, s0));
        prm.iblocks = E.d_r_iblocks; prm.done = E.d_r_done; prm.out = out_fused; prm.G = E.G;
    }
    if (E.opt_ws)
        CU(launch_force_ws(s0, 2 * 148, E.opt_nr, E.d_r_walks, E.d_r_tasks, E.d_r_epi, E.d_tree_ide, E.d_tree_ids, E.d_epj, E.d_spj,
                           E.d_r_part4, E.d_r_partn, prm, E.opt_sp2i != 0));
    else
        CU(launch_force_persistent(s0, 2 * 148, E.opt_nr, E.d_r_walks, E.d_r_tasks, E.d_r_epi, E.d_tree_ide, E.d_tree_ids, E.d_epj, E.d_spj,
                                   E.d_r_part4, E.d_r_partn, prm, E.opt_sp2i != 0));
    CU(cudaEventRecord(E.ev_tl[4], s0));
    if (!fuse) CU(launch_reduce(s0, E.r_n_iblk, E.d_r_iblocks, E.d_r_part4, E.d_r_partn, E.d_r_out, E.G, E.d_r_meta));
    CU(cudaEventRecord(E.ev_tl[5], s0));
    if (!fuse) CU(cudaMemcpyAsync(E.h_r_out, E.d_r_out, sizeof(ForceOut) * (size_t)E.r_n_i, cudaMemcpyDeviceToHost, s0));
    CU(cudaMemcpyAsync(E.h_r_meta, E.d_r_meta, 8 * sizeof(int), cudaMemcpyDeviceToHost, s0));
    CU(cudaMemcpyAsync(E.h_counts_p, E.d_counts, sizeof(int2) * (size_t)ng, cudaMemcpyDeviceToHost, s0));
    CU(cudaMemcpyAsync(E.h_over_p, E.d_overflow, 2 * sizeof(int), cudaMemcpyDeviceToHost, s0));
    CU(cudaEventRecord(E.ev_tl[6], s0));
    E.prof.n_kernel_launch += 5;
    E.prof.d2h_bytes += (long long)(sizeof(ForceOut) * (size_t)E.r_n_i + sizeof(int2) * (size_t)ng);
    CU(cudaEventSynchronize(E.ev_tl[6]));

    if (E.h_over_p[0]) return fail(PB_ERR_ARG, "pb_tree_force_resident: tree-walk frontier exceeded %d cells per level", kWalkCap);
    if ((rc = check_let_error()) != PB_OK) return rc;
    const bool retry = (!exact && E.h_over_p[1] != 0) || E.h_r_meta[3] != 0;
    E.h_counts.assign(E.h_counts_p, E.h_counts_p + ng);    // true list lengths of this step, whatever happened
    if (retry) {
        if (exact) return fail(PB_ERR_ARG, "pb_tree_force_resident: task plan does not fit its buffers (flags %d)", E.h_r_meta[3]);
        return 1;                                          // reservation too small: the caller repeats the step exactly
    }
    E.prev_counts = E.h_counts;
    E.r_prev_tasks = E.h_r_meta[5]; E.r_prev_part = E.h_r_meta[1];
    for (int k = 0; k < 6; k++) { float ms = 0.f; if (cudaEventElapsedTime(&ms, E.ev_tl[k], E.ev_tl[k + 1]) == cudaSuccess) E.tl_ms[k] = ms; }
    E.tl_valid = true;
    E.prof.t_calc += 1e-3 * (E.tl_ms[0] + E.tl_ms[2] + E.tl_ms[3] + E.tl_ms[4]);
    E.prof.t_recv += 1e-3 * E.tl_ms[5];

    // forces back to the caller's array (group order)
    const double t0 = now_s();
    const long long n = E.r_n_i;
    char* dst = (char*)force;
    if (direct) {
        // already there
    } else if (plain) {
        const long long chunk = 1 << 15;
#pragma omp parallel for schedule(static)
        for (long long c = 0; c < (n + chunk - 1) / chunk; c++)
            memcpy(dst + (size_t)c * chunk * sizeof(ForceOut), E.h_r_out + c * chunk, sizeof(ForceOut) * (size_t)std::min(chunk, n - c * chunk));
    } else {
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < n; i++) {
            char* q = dst + (size_t)i * L.stride;
            memcpy(q + L.off_acc, &E.h_r_out[i].ax, 24);
            memcpy(q + L.off_pot, &E.h_r_out[i].pot, 8);
            memcpy(q + L.off_nngb, &E.h_r_out[i].n_ngb, 8);
        }
    }
    E.prof.t_copy += now_s() - t0; E.prof.t_unpack += now_s() - t0;
    long long n_ej = 0, n_sj = 0, i_ep = 0, i_sp = 0;
    for (int g = 0; g < ng; g++) {
        n_ej += E.h_counts[g].x; n_sj += E.h_counts[g].y;
        i_ep += (long long)E.grp_n[g] * E.h_counts[g].x; i_sp += (long long)E.grp_n[g] * E.h_counts[g].y;
    }
    E.prof.n_walk += ng; E.prof.n_epi += n; E.prof.n_epj += n_ej; E.prof.n_spj += n_sj;
    E.prof.n_call += 1; E.prof.n_interaction_ep += i_ep; E.prof.n_interaction_sp += i_sp;
    return PB_OK;
}
} // namespace

int pb_tree_force_resident(void* force, const pb_layout_force* lforce) {
    int rc = ensure_init();
    if (rc != PB_OK) return rc;
    if (E.outstanding) return fail(PB_ERR_PROTOCOL, "pb_tree_force_resident while a dispatch is outstanding");
    if (!E.j_published) return fail(PB_ERR_PROTOCOL, "pb_tree_force_resident before the j-particles were published");
    if (E.n_groups == 0) return PB_OK;
    if (!E.d_cells) return fail(PB_ERR_PROTOCOL, "pb_tree_force_resident before pb_tree_upload");
    if (!force || !lforce) return fail(PB_ERR_ARG, "pb_tree_force_resident: null argument");
    if (E.n_cells > E.n_spj) return fail(PB_ERR_PROTOCOL, "pb_tree_force_resident: the SP store (%d) must hold one superparticle per cell (%d)", E.n_spj, E.n_cells);
    if (!E.count_pending) return fail(PB_ERR_PROTOCOL, "pb_tree_force_resident: call pb_tree_upload for this step first");
    E.count_pending = false;
    bool exact = !(E.spec_pending && E.r_prev_tasks > 0);
    E.spec_pending = false;
    if (exact) {
        // first step (or speculation off): the counting pass launched by pb_tree_upload tells the host the list lengths
        CU(cudaEventSynchronize(E.ev_count));
        if (*E.h_over_p) return fail(PB_ERR_ARG, "pb_tree_force_resident: tree-walk frontier exceeded %d cells per level", kWalkCap);
        E.h_counts.assign(E.h_counts_p, E.h_counts_p + E.n_groups);
    }
    rc = resident_run(force, *lforce, exact);
    if (rc == 1) rc = resident_run(force, *lforce, true);   // a reservation was too small: h_counts now holds the true lengths
    return rc;
}

int pb_tree_timeline(float* ms, int n) {
    if (!E.tl_valid) return fail(PB_ERR_PROTOCOL, "pb_tree_timeline: no device-resident step has completed");
    for (int k = 0; k < n && k < 6; k++) ms[k] = E.tl_ms[k];
    return PB_OK;
}

int pb_tree_lists(int* n_ep, int* n_sp, int* id_ep, long long cap_ep, int* id_sp, long long cap_sp) {
    if (!E.inited || E.h_counts.empty()) return fail(PB_ERR_PROTOCOL, "pb_tree_lists before pb_tree_force");
    for (int g = 0; g < E.n_groups; g++) {
        if (n_ep) n_ep[g] = E.h_counts[g].x;
        if (n_sp) n_sp[g] = E.h_counts[g].y;
    }
    if (!id_ep && !id_sp) return PB_OK;
    std::vector<int> he(E.cap_tree_ide ? (size_t)E.h_tree_off.back().x + align_up((size_t)E.h_counts.back().x, 4) : 0);
    std::vector<int> hs(E.cap_tree_ids ? (size_t)E.h_tree_off.back().y + align_up((size_t)E.h_counts.back().y, 4) : 0);
    if (!he.empty()) CU(cudaMemcpy(he.data(), E.d_tree_ide, sizeof(int) * he.size(), cudaMemcpyDeviceToHost));
    if (!hs.empty()) CU(cudaMemcpy(hs.data(), E.d_tree_ids, sizeof(int) * hs.size(), cudaMemcpyDeviceToHost));
    long long ke = 0, ks = 0;
    for (int g = 0; g < E.n_groups; g++) {
        for (int j = 0; j < E.h_counts[g].x; j++, ke++) if (id_ep && ke < cap_ep) id_ep[ke] = he[(size_t)E.h_tree_off[g].x + j];
        for (int j = 0; j < E.h_counts[g].y; j++, ks++) if (id_sp && ks < cap_sp) id_sp[ks] = hs[(size_t)E.h_tree_off[g].y + j];
    }
    return PB_OK;
}

int pb_get_profile(pb_profile* out, int reset) {
    if (out) *out = E.prof;
    if (reset) memset(&E.prof, 0, sizeof(E.prof));
    return PB_OK;
}

// ---- device-resident replay -------------------------------------------------------------------
int pb_record_begin(void) {
    int rc = ensure_init();
    if (rc != PB_OK) return rc;
    CU(cudaDeviceSynchronize());
    for (auto& r : E.recs) { CU(cudaFree(r.d_arena)); if (r.d_ide_x) CU(cudaFree(r.d_ide_x)); }
    E.recs.clear();
    E.recording = true;
    return PB_OK;
}

int pb_record_end(void) {
    E.recording = false;
    return PB_OK;
}

int pb_replay_launches(void) {
    int n = 0;
    for (auto& r : E.recs) n += (r.plan.n_tasks > 0) + (r.plan.n_iblocks > 0);
    return n;
}

int pb_replay(int n_iter, float* ms_total, float* ms_force) {
    if (!E.inited) return fail(PB_ERR_PROTOCOL, "pb_replay before init");
    if (E.outstanding) return fail(PB_ERR_PROTOCOL, "pb_replay while a dispatch is outstanding");
    if (E.recs.empty()) return fail(PB_ERR_PROTOCOL, "pb_replay: nothing recorded");
    if (n_iter < 1) return fail(PB_ERR_ARG, "pb_replay: n_iter < 1");
    // every recorded sub-batch is re-launched on the stream it originally ran on, with that
    // stream's own partial/output buffers, so kernels of different streams overlap as they do
    // in a real dispatch
    CU(cudaDeviceSynchronize());
    bool used[kMaxStreams] = {false};
    for (auto& r : E.recs) {
        int rc;
        if ((rc = grow_part(E.slots[r.slot], r.plan.n_part)) != PB_OK) return rc;
        if ((rc = grow_out(E.slots[r.slot], r.plan.n_i)) != PB_OK) return rc;
        used[r.slot] = true;
    }
    cudaStream_t st0 = E.slots[0].stream;
    cudaEvent_t a = E.slots[0].ev[0], b = E.slots[0].ev[1];
    for (int pass = 0; pass < 2; pass++) {
        const bool force_only = (pass == 1);
        float* dst = force_only ? ms_force : ms_total;
        if (!dst) continue;
        CU(cudaDeviceSynchronize());
        CU(cudaEventRecord(a, st0));
        for (int s = 1; s < kMaxStreams; s++)
            if (used[s]) CU(cudaStreamWaitEvent(E.slots[s].stream, a, 0));
        for (int it = 0; it < n_iter; it++)
            for (auto& r : E.recs) {
                Slot& S = E.slots[r.slot];
                CU(launch_plan(S.stream, r.plan, r.d_arena, r.direct, S.d_part4, S.d_partn, S.d_out, S.h_out, force_only));
            }
        for (int s = 1; s < kMaxStreams; s++)
            if (used[s]) {
                CU(cudaEventRecord(E.slots[s].ev[3], E.slots[s].stream));
                CU(cudaStreamWaitEvent(st0, E.slots[s].ev[3], 0));
            }
        CU(cudaEventRecord(b, st0));
        CU(cudaEventSynchronize(b));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, a, b));
        *dst = ms / n_iter;
        E.prof.n_kernel_launch += (long long)n_iter * (force_only ? (int)E.recs.size() : pb_replay_launches());
    }
    return PB_OK;
}

} // extern "C"
