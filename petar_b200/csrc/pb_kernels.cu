// pb_kernels.cu — hand-written sm_100a kernels of the soft-force hot path.
//
// What they compute is what the reference's two kernels compute
//   force_kernel_ep_ep / dev_gravity_ep_ep   reference src/force_gpu_cuda.cu:222-273 / :50-77
//   force_kernel_ep_sp / dev_gravity_ep_sp   reference src/force_gpu_cuda.cu:465-511 / :276-323
// (and, in fp64, the oracle functors src/soft_force.hpp:38-87, 160-200); how they compute it is
// new:
//
//  * one launch for both interaction kinds; a CTA works on a Task = (i-group of a walk) x
//    (chunk of that walk's EP or SP list); the host sizes tasks so that every CTA runs about the
//    same number of inner-loop steps and there are several waves of CTAs on 148 SMs;
//  * i-particles live in registers (one per lane), relative to the walk's origin;
//  * j tiles (256 entries) are gathered by index with coalesced index reads and 16-byte loads
//    from the L2-resident j store, shifted to the walk origin from their hi/lo fp32 split, and
//    written to shared memory as PAIRS so the inner loop runs on packed fp32x2 instructions
//    (FADD2 / FMUL2 / FFMA2, Blackwell only) — two interactions per issued FP instruction;
//    the gather of tile k+1 and the index read of tile k+2 are in flight during tile k;
//  * reciprocal square roots come from MUFU.RSQ (optionally refined by one Newton step);
//  * forces/potentials are accumulated in fp32 per tile and folded across tiles with a
//    compensated (Kahan-Babuska) sum; per-task results leave the SM as exact hi+lo doubles and a
//    second tiny kernel adds the tasks' partials in fp64 in a fixed order (deterministic, no
//    atomics), applies G and writes ForceSoft-shaped records.
//  * no tensor cores: this is not a contraction.
#include "pb_device.h"
#include "pb_pairs.cuh"

namespace pb {

// ------------------------------------------------------------------------------------------
// Index tiles through TMA: a chunk's index list is contiguous and 16-byte aligned, so each 1 KB
// tile of it is one `cp.async.bulk` (1-D TMA bulk copy, SASS UBLKCP) into a 3-deep shared-memory
// ring, completion signalled on an mbarrier — issued by one thread up to four tiles ahead of the tile being computed, no
// registers held, no per-thread global loads.  (The j records themselves are an indexed gather
// and stay 16-byte loads from L2.)  PB_TMA_IDS=0 builds the plain-load variant.
// ------------------------------------------------------------------------------------------
#ifndef PB_TMA_IDS
#define PB_TMA_IDS 1
#endif

constexpr int kIdRing = 3;      // ring slots
constexpr int kIdDirect = 2;    // tiles 0,1 of a chunk are read directly

struct IdRing {
    alignas(16) int buf[kIdRing][kTileJ];
    alignas(8) unsigned long long bar[kIdRing];
};

struct IdPipe {
    const int* ids;      // nullptr: dense list (element j is store slot dbase + j)
    int j_count, n_tiles, dbase;
    IdRing* ring;

    // ring entry r holds tile r + kIdDirect (the first kIdDirect tiles are plain loads: nothing to wait for
    // in the prologue, and the barrier that publishes the mbarrier init is the one the loop needs anyway)
    __device__ __forceinline__ void issue(int r) const {               // one thread
        const int tile = r + kIdDirect;
        const unsigned bytes = (unsigned)(((min(kTileJ, j_count - tile * kTileJ) + 3) & ~3) * 4);
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&ring->bar[r % kIdRing]);
        const unsigned dst = (unsigned)__cvta_generic_to_shared(&ring->buf[r % kIdRing][0]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst), "l"(ids + (size_t)tile * kTileJ), "r"(bytes), "r"(bar) : "memory");
    }
    // call by all threads; a __syncthreads() must separate it from the first get(tile >= kIdDirect)
    __device__ __forceinline__ void init(const int* ids_, int j_count_, int n_tiles_, int dbase_, IdRing* ring_, int tid) {
        ids = ids_; j_count = j_count_; n_tiles = n_tiles_; dbase = dbase_; ring = ring_;
#if PB_TMA_IDS
        if (tid == 0) {
            for (int b = 0; b < kIdRing; ++b) {
                const unsigned bar = (unsigned)__cvta_generic_to_shared(&ring->bar[b]);
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(1) : "memory");
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            if (ids) for (int r = 0; r < min(kIdRing, n_tiles - kIdDirect); ++r) issue(r);
        }
#endif
    }
    // id of list element tile*kTileJ + tid, or -1 past the end of the chunk
    __device__ __forceinline__ int get(int tile, int tid) const {
        const int j = tile * kTileJ + tid;
        if (tile >= n_tiles) return -1;
        if (!ids) return j < j_count ? dbase + j : -1;
#if PB_TMA_IDS
        if (tile < kIdDirect) return j < j_count ? __ldg(ids + j) : -1;
        const int r = tile - kIdDirect;
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&ring->bar[r % kIdRing]);
        const unsigned parity = (unsigned)((r / kIdRing) & 1);
        asm volatile("{\n"
                     ".reg .pred P1;\n"
                     "LAB_WAIT:\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                     "@P1 bra DONE;\n"
                     "bra LAB_WAIT;\n"
                     "DONE:\n"
                     "}" :: "r"(bar), "r"(parity) : "memory");
        return j < j_count ? ring->buf[r % kIdRing][tid] : -1;
#else
        return j < j_count ? __ldg(ids + j) : -1;
#endif
    }
    // once every thread has read tiles <= k+2 = ring entries <= k (the caller knows: all of them have arrived on the
    // mbarrier of tile k+1, which they do after that read), the slot of entry k is free for entry k + kIdRing
    __device__ __forceinline__ void refill(int k, int tid) const {
#if PB_TMA_IDS
        if (ids && tid == 0 && k + kIdRing + kIdDirect < n_tiles) issue(k + kIdRing);
#endif
    }
};

// ------------------------------------------------------------------------------------------
// the force kernel
// ------------------------------------------------------------------------------------------
// PERSIST: the grid is 2 CTAs per SM and every CTA pulls task numbers from an atomic cursor until the task list is
// used up — the task count lives in device memory (meta[0], cursor meta[4]) because the plan was made on the device
// (pb_plan.cu), so no host round trip sits between the tree walk and the forces.
// TWOI: SP tasks of groups with at least two i-blocks give every warp TWO blocks (one particle of each per lane) and half as
// much of every j tile (sp_pairs_2i in pb_pairs.cuh: +6 % on the SP loop).
// FUSE: the warp that delivers an i-block's last partial sum reduces the block and writes the forces (finish_block, pb_pairs.cuh).
template <int NR, int MINB, bool EMIT = false, bool PERSIST = false, bool TWOI = false, bool FUSE = false>
__global__ void __launch_bounds__(kThreads, MINB)
force_kernel(const Walk* __restrict__ walks, const Task* __restrict__ tasks,
             const float4* __restrict__ epi,
             const int* __restrict__ id_epj, const int* __restrict__ id_spj,
             const float4* __restrict__ epj, const float4* __restrict__ spj,
             double4* __restrict__ part4, int* __restrict__ partn, Params prm)
{
    __shared__ Smem sm;
    __shared__ int near_flag[kTileBufs][kWarpsPerCta];   // per tile buffer, per staging warp (= 16-pair segment)
    __shared__ IdRing ring;                      // index tiles, filled by TMA bulk copies
    __shared__ TileBars bars;                    // tile hand-over (see TileBars)
    __shared__ int jid[EMIT ? kTileBufs : 1][EMIT ? kTileJ : 1];   // store index of every staged j (neighbour-list emission only)
    __shared__ int s_task;

    const int tid   = threadIdx.x;
    const int warp  = tid >> 5, lane = tid & 31;
  for (;;) {
    int task_id = blockIdx.x;
    if (PERSIST) {
        if (tid == 0) s_task = atomicAdd(prm.meta + 4, 1);
        __syncthreads();
        task_id = s_task;
        if (task_id >= prm.meta[0]) break;
    }
    const Task task = tasks[task_id];
    const Walk w    = walks[task.walk];

    // warp role: i-block `ib` of the group, j-split slot `js`
    const bool busy = warp < task.nib * task.jsplit;
    const int  ib   = busy ? warp % task.nib : 0;
    const int  js   = busy ? warp / task.nib : 0;
    const int  ppw  = kTilePairs / task.jsplit;             // pairs of each tile this warp consumes

    // A block with at most 16 (8) real i-particles — the ragged end of a walk — lets 2 (4) lanes share each particle
    // and split every j segment between them instead of idling: lane = (sub, il), il = the particle, sub = the lane's
    // share of the pairs; the shares are added by shuffles after the loops.
    const int  n_blk  = busy ? min(32, w.ni - (task.i_first + ib * 32)) : 32;
    const int  ishift = (n_blk <= 8) ? 2 : (n_blk <= 16) ? 1 : 0;     // isplit = 1 << ishift lanes per particle
    const int  lanes_i = 32 >> ishift;
    const int  il = lane & (lanes_i - 1), sub = lane >> (5 - ishift);

    // i-particle (registers): relative to the walk origin already (host formed x_i - origin in fp64)
    const int  i_loc  = task.i_first + ib * 32 + il;
    const bool ivalid = busy && (i_loc < w.ni);
    float4 pi = make_float4(0.f, 0.f, 0.f, 0.f), pil = make_float4(0.f, 0.f, 0.f, 0.f), pih = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ivalid) {
        const float4* rec = epi + (size_t)prm.i_f4 * (size_t)(w.i_off + i_loc);
        pi  = __ldg(rec);                                        // {x, y, z, r_search}: relative position, hi part
        pil = __ldg(rec + 1);                                    // {xl, yl, zl, -}: lo part
        if (prm.abs_mode == 2) pih = __ldg(rec + 2);             // {float(x), float(y), float(z), -}: absolute, as the CPU replay casts it
    }
    const float rsi2 = ivalid ? pi.w * pi.w : -1.f;

    KSum kx, ky, kz, kp;
    kx.init(); ky.init(); kz.init(); kp.init();
    int cnt = 0;

    const int n_tiles = (task.j_count + kTileJ - 1) / kTileJ;

    // two i-particles per lane (SP tasks only): block pair (b0, b0 + npair), j-split slot js2 of jsp = 2 jsplit
    const bool two_i = TWOI && task.kind == 1 && task.nib >= 2;
    const int  npair = task.nib >> 1, jsp = 2 * task.jsplit;
    const int  b0 = two_i ? warp % npair : 0, js2 = two_i ? warp / npair : 0;
    KSum k2x, k2y, k2z, k2p;
    k2x.init(); k2y.init(); k2z.init(); k2p.init();
    float xi2[2] = {0.f, 0.f}, yi2[2] = {0.f, 0.f}, zi2[2] = {0.f, 0.f};
    if (two_i) {
        const int ia = task.i_first + b0 * 32 + lane, ib2 = ia + npair * 32;
        if (ia < w.ni)  { const float4 q = __ldg(epi + (size_t)prm.i_f4 * (size_t)(w.i_off + ia));  xi2[0] = q.x; yi2[0] = q.y; zi2[0] = q.z; }
        if (ib2 < w.ni) { const float4 q = __ldg(epi + (size_t)prm.i_f4 * (size_t)(w.i_off + ib2)); xi2[1] = q.x; yi2[1] = q.y; zi2[1] = q.z; }
    }

    // Tile pipeline (both kinds): index tiles arrive by TMA up to four tiles ahead, ids are picked up two tiles ahead,
    // the gathered j records one tile ahead; tile k+1 is stored (and its mbarrier arrived on) after tile k was
    // computed, and waited for only at the start of iteration k+1.
    bars.init(tid);
    if (task.kind == 0 || task.kind == 2) {
        const bool count_only = (task.kind == 2);     // neighbour search only (tree_nb): no force math
        const int* ids = (w.ej_off >= 0) ? id_epj + w.ej_off + task.j_begin : nullptr;   // ej_off < 0: dense list
        const int dbase = task.j_begin;
        IdPipe idp;
        idp.init(ids, task.j_count, n_tiles, dbase, &ring, tid);
        int id_cur = idp.get(0, tid);
        EpRegs jr  = ep_load_j(epj, id_cur);
        int id_nxt = idp.get(1, tid);
        __syncthreads();                               // mbarrier inits (tile ring, index ring) visible to every thread
        {
            const bool nr_ = ep_store(sm.ep[0], tid, id_cur, jr, w, prm.abs_mode);
            const unsigned bal = __ballot_sync(0xffffffffu, nr_);
            if (lane == 0) near_flag[0][warp] = (bal != 0u);
            if (EMIT) jid[0][tid] = id_cur;
            bars.arrive(0);
        }
        int cb = 0; unsigned cpar = 0;                 // buffer and mbarrier phase parity of tile k
        for (int k = 0; k < n_tiles; ++k) {
            const bool more = (k + 1 < n_tiles);
            const int nb = (cb + 1 == kTileBufs) ? 0 : cb + 1;
            if (more) jr = ep_load_j(epj, id_nxt);
            const int id_nn = idp.get(k + 2, tid);
            bars.wait(cb, cpar);                       // every thread has stored its j of tile k
            if (k >= 1) idp.refill(k - 1, tid);        // ... and had read the ids of tile k+1 before that
            if (busy) {
                const EpTile& T = sm.ep[cb];
                const int nv  = min(kTileJ, task.j_count - k * kTileJ);
                const int npu = (((nv + 1) >> 1) + kPairUnroll - 1) & ~(kPairUnroll - 1);
                // a ragged (last) tile is shared out evenly, in whole 16-pair segments, between the warps of an i-block
                const int ppk = (nv == kTileJ) ? ppw : max(16, (((npu + task.jsplit - 1) / task.jsplit) + 15) & ~15);
                const int p0  = js * ppk, p1 = min(p0 + ppk, npu);
                float2 ax = bc(0.f), ay = bc(0.f), az = bc(0.f), pt = bc(0.f), cf = bc(0.f);
                // 16-pair segments = the 32 j one staging warp wrote; count only where flagged
                for (int seg0 = p0; seg0 < p1; seg0 += 16) {
                    // this lane's contiguous share of the segment (all of it unless lanes share a particle)
                    const int seg = seg0 + sub * (16 >> ishift);
                    const int e = min(min(seg0 + 16, p1), seg + (16 >> ishift));
                    if (count_only) {
                        if (near_flag[cb][seg0 >> 4])
                            ep_count_pairs<EMIT>(T, seg, e, pi.x, pi.y, pi.z, pil.x, pil.y, pil.z, rsi2, prm.eps2, cf,
                                                 jid[EMIT ? cb : 0], ivalid ? (unsigned int)(prm.i_base + w.i_off + i_loc) : 0xffffffffu, prm);
                    } else if (near_flag[cb][seg0 >> 4]) {
                        if (prm.abs_mode == 2)
                            ep_pairs<NR, 2>(T, seg, e, pi.x, pi.y, pi.z, pil.x, pil.y, pil.z, rsi2, pih.x, pih.y, pih.z, prm.eps2, prm.rcut2, prm.rinv_cut, ax, ay, az, pt, cf);
                        else
                            ep_pairs<NR, 1>(T, seg, e, pi.x, pi.y, pi.z, pil.x, pil.y, pil.z, rsi2, 0.f, 0.f, 0.f, prm.eps2, prm.rcut2, 0.f, ax, ay, az, pt, cf);
                    } else
                        ep_pairs<NR, 0>(T, seg, e, pi.x, pi.y, pi.z, pil.x, pil.y, pil.z, rsi2, 0.f, 0.f, 0.f, prm.eps2, prm.rcut2, 0.f, ax, ay, az, pt, cf);
                }
                kx.add(ax.x + ax.y); ky.add(ay.x + ay.y); kz.add(az.x + az.y); kp.add(pt.x + pt.y);
                cnt += (int)(cf.x + cf.y);
            }
            if (more) {
                const bool nr_ = ep_store(sm.ep[nb], tid, id_nxt, jr, w, prm.abs_mode);
                const unsigned bal = __ballot_sync(0xffffffffu, nr_);
                if (lane == 0) near_flag[nb][warp] = (bal != 0u);
                if (EMIT) jid[nb][tid] = id_nxt;
                bars.arrive(nb);
            }
            id_nxt = id_nn;
            cb = nb; if (nb == 0) cpar ^= 1u;
        }
    } else {
        const int* ids = id_spj + w.sj_off + task.j_begin;
        IdPipe idp;
        idp.init(ids, task.j_count, n_tiles, 0, &ring, tid);
        int id_cur = idp.get(0, tid);
        SpRegs jr  = sp_load_j(spj, id_cur);
        int id_nxt = idp.get(1, tid);
        __syncthreads();                               // mbarrier inits visible
        sp_store(sm.sp[0], tid, id_cur, jr, w, prm.eps2);
        bars.arrive(0);
        int cb = 0; unsigned cpar = 0;
        for (int k = 0; k < n_tiles; ++k) {
            const bool more = (k + 1 < n_tiles);
            const int nb = (cb + 1 == kTileBufs) ? 0 : cb + 1;
            if (more) jr = sp_load_j(spj, id_nxt);
            const int id_nn = idp.get(k + 2, tid);
            bars.wait(cb, cpar);
            if (k >= 1) idp.refill(k - 1, tid);
            if (two_i) {
                const int nv  = min(kTileJ, task.j_count - k * kTileJ);
                const int npu = ((nv + 1) >> 1);
                const int ppk = (nv == kTileJ) ? kTilePairs / jsp : (npu + jsp - 1) / jsp;
                const int p0  = js2 * ppk, p1 = min(p0 + ppk, npu);
                float2 ax[2] = {bc(0.f), bc(0.f)}, ay[2] = {bc(0.f), bc(0.f)}, az[2] = {bc(0.f), bc(0.f)}, pt[2] = {bc(0.f), bc(0.f)};
                sp_pairs_2i<NR>(sm.sp[cb], p0, p1, xi2, yi2, zi2, prm.eps2, ax, ay, az, pt);
                kx.add(ax[0].x + ax[0].y); ky.add(ay[0].x + ay[0].y); kz.add(az[0].x + az[0].y); kp.add(pt[0].x + pt[0].y);
                k2x.add(ax[1].x + ax[1].y); k2y.add(ay[1].x + ay[1].y); k2z.add(az[1].x + az[1].y); k2p.add(pt[1].x + pt[1].y);
            } else if (busy) {
                const int nv  = min(kTileJ, task.j_count - k * kTileJ);
                const int npu = ((nv + 1) >> 1);          // pairs with at least one real j (padding pairs are inert anyway)
                const int ppk = (nv == kTileJ) ? ppw : (npu + task.jsplit - 1) / task.jsplit;   // ragged tile: even shares
                const int p0  = js * ppk, p1 = min(p0 + ppk, npu);
                float2 ax = bc(0.f), ay = bc(0.f), az = bc(0.f), pt = bc(0.f);
                const int plen = (max(p1 - p0, 0) + (1 << ishift) - 1) >> ishift;   // this lane's contiguous share of the warp's pairs
                const int q0 = p0 + sub * plen, q1 = min(p1, q0 + plen);
                sp_pairs<NR>(sm.sp[cb], q0, q1, pi.x, pi.y, pi.z, prm.eps2, ax, ay, az, pt);
                kx.add(ax.x + ax.y); ky.add(ay.x + ay.y); kz.add(az.x + az.y); kp.add(pt.x + pt.y);
            }
            if (more) { sp_store(sm.sp[nb], tid, id_nxt, jr, w, prm.eps2); bars.arrive(nb); }
            id_nxt = id_nn;
            cb = nb; if (nb == 0) cpar ^= 1u;
        }
    }
    __syncthreads();       // every warp is done with the tile buffers (they are reused for the cross-warp combine below)

    if (two_i) {
        // both particles of a lane: combine the jsp warps of a block pair in fixed order js2 = 0,1,..., warp b0 writes
        for (int set = 0; set < 2; ++set) {
            double d0 = set ? k2x.value() : kx.value(), d1 = set ? k2y.value() : ky.value();
            double d2 = set ? k2z.value() : kz.value(), d3 = set ? k2p.value() : kp.value();
            sm.red[warp][0][lane] = d0; sm.red[warp][1][lane] = d1; sm.red[warp][2][lane] = d2; sm.red[warp][3][lane] = d3;
            __syncthreads();
            if (js2 == 0) {
                for (int s = 1; s < jsp; ++s) {
                    const int ww = s * npair + b0;
                    d0 += sm.red[ww][0][lane]; d1 += sm.red[ww][1][lane]; d2 += sm.red[ww][2][lane]; d3 += sm.red[ww][3][lane];
                }
                const int slot = task.part_base + (b0 + set * npair) * 32 + lane;
                part4[slot] = make_double4(d0, d1, d2, d3);
                partn[slot] = 0;
                if (FUSE) finish_block(task.blk0 + b0 + set * npair, task.n_chunks, lane, part4, partn, prm);
            }
            __syncthreads();
        }
    } else {
    // per-warp totals as exact doubles (hi + lo)
    double dax = kx.value(), day = ky.value(), daz = kz.value(), dpt = kp.value();
    for (int o = lanes_i; o < 32; o <<= 1) {              // lanes that shared a particle (isplit > 1): fixed-order butterfly
        dax += __shfl_xor_sync(0xffffffffu, dax, o); day += __shfl_xor_sync(0xffffffffu, day, o);
        daz += __shfl_xor_sync(0xffffffffu, daz, o); dpt += __shfl_xor_sync(0xffffffffu, dpt, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }

    if (task.jsplit > 1) {
        // combine the jsplit warps that share an i-block, in fixed order js = 0,1,...
        if (busy) {
            sm.red[warp][0][lane] = dax; sm.red[warp][1][lane] = day;
            sm.red[warp][2][lane] = daz; sm.red[warp][3][lane] = dpt;
        }
        // neighbour counts ride in a second pass to keep the scratch small
        __syncthreads();
        if (busy && js == 0) {
            for (int s = 1; s < task.jsplit; ++s) {
                const int ww = s * task.nib + ib;
                dax += sm.red[ww][0][lane]; day += sm.red[ww][1][lane];
                daz += sm.red[ww][2][lane]; dpt += sm.red[ww][3][lane];
            }
        }
        __syncthreads();
        int* redn = reinterpret_cast<int*>(&sm.red[0][0][0]);
        if (busy) redn[warp * 32 + lane] = cnt;
        __syncthreads();
        if (busy && js == 0)
            for (int s = 1; s < task.jsplit; ++s) cnt += redn[(s * task.nib + ib) * 32 + lane];
    }

    if (busy && js == 0 && sub == 0) {
        const int slot = task.part_base + ib * 32 + il;
        part4[slot] = make_double4(dax, day, daz, dpt);
        partn[slot] = cnt;
    }
    if (FUSE && busy && js == 0) finish_block(task.blk0 + ib, task.n_chunks, lane, part4, partn, prm);
    }
    if (!PERSIST) break;
    __syncthreads();                                   // everybody is done with this task's shared state
    if (tid == 0) {                                    // the mbarriers are re-initialised by the next task
        for (int b = 0; b < kTileBufs; ++b) {
            const unsigned bar = (unsigned)__cvta_generic_to_shared(&bars.full[b]);
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" :: "r"(bar) : "memory");
        }
        for (int b = 0; b < kIdRing; ++b) {
            const unsigned bar = (unsigned)__cvta_generic_to_shared(&ring.bar[b]);
            asm volatile("mbarrier.inval.shared::cta.b64 [%0];" :: "r"(bar) : "memory");
        }
    }
  }
}

cudaError_t launch_force(cudaStream_t s, int n_tasks, int nr_steps, int min_blocks,
                         const Walk* walks, const Task* tasks,
                         const float4* epi, const int* id_epj, const int* id_spj,
                         const float4* epj, const float4* spj,
                         double4* part4, int* partn, Params p, bool emit_pairs, bool two_i)
{
    if (n_tasks <= 0) return cudaSuccess;
#define PB_LAUNCH(...) force_kernel<__VA_ARGS__><<<n_tasks, kThreads, 0, s>>>(walks, tasks, epi, id_epj, id_spj, epj, spj, part4, partn, p)
    if (emit_pairs)          PB_LAUNCH(0, 2, true);        // neighbour lists: count-only tasks, no rsqrt involved
    else if (min_blocks >= 3) { if (nr_steps >= 1) PB_LAUNCH(1, 3); else PB_LAUNCH(0, 3); }       // occupancy experiment: one i per lane only
    else if (p.out != nullptr) {                           // fused reduction: Params carries the i-block table, the counters and the result array
        if (two_i)          { if (nr_steps >= 1) PB_LAUNCH(1, 2, false, false, true, true); else PB_LAUNCH(0, 2, false, false, true, true); }
        else                { if (nr_steps >= 1) PB_LAUNCH(1, 2, false, false, false, true); else PB_LAUNCH(0, 2, false, false, false, true); }
    }
    else if (two_i)         { if (nr_steps >= 1) PB_LAUNCH(1, 2, false, false, true); else PB_LAUNCH(0, 2, false, false, true); }
    else                    { if (nr_steps >= 1) PB_LAUNCH(1, 2); else PB_LAUNCH(0, 2); }
#undef PB_LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_force_persistent(cudaStream_t s, int n_ctas, int nr_steps,
                                    const Walk* walks, const Task* tasks, const float4* epi, const int* id_epj, const int* id_spj,
                                    const float4* epj, const float4* spj, double4* part4, int* partn, Params p, bool two_i)
{
    if (n_ctas <= 0) return cudaSuccess;
#define PB_LAUNCH(...) force_kernel<__VA_ARGS__><<<n_ctas, kThreads, 0, s>>>(walks, tasks, epi, id_epj, id_spj, epj, spj, part4, partn, p)
    if (two_i) { if (nr_steps >= 1) PB_LAUNCH(1, 2, false, true, true); else PB_LAUNCH(0, 2, false, true, true); }
    else       { if (nr_steps >= 1) PB_LAUNCH(1, 2, false, true); else PB_LAUNCH(0, 2, false, true); }
#undef PB_LAUNCH
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// reduction: one warp per 32-wide i-block adds that block's task partials in fp64, chunk order
// fixed, applies G and writes ForceSoft-shaped records.
//   acc = G * sum ; pot = -G * sum(pot terms)   (signs: see the pair loops above)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
reduce_kernel(int n_iblocks, const IBlock* __restrict__ iblocks,
              const double4* __restrict__ part4, const int* __restrict__ partn,
              ForceOut* __restrict__ out, double G, const int* __restrict__ meta)
{
    const int b    = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= n_iblocks) return;
    if (meta && meta[3] != 0) return;                 // device-made plan did not fit its buffers: nothing was computed
    const IBlock ibk = iblocks[b];
    if (lane >= ibk.n_valid) return;
    double ax = 0.0, ay = 0.0, az = 0.0, pt = 0.0;
    long long n = 0;
#pragma unroll 8
    for (int c = 0; c < ibk.n_chunks; ++c) {          // independent loads: unrolled for memory-level parallelism
        const int slot = ibk.part_base + c * ibk.stride + lane;
        const double4 v = part4[slot];
        ax += v.x; ay += v.y; az += v.z; pt += v.w;
        n += partn[slot];
    }
    ForceOut o;
    o.ax = G * ax; o.ay = G * ay; o.az = G * az; o.pot = -(G * pt); o.n_ngb = n;
    out[ibk.out_off + lane] = o;
}

cudaError_t launch_reduce(cudaStream_t s, int n_iblocks, const IBlock* iblocks,
                          const double4* part4, const int* partn, ForceOut* out, double G, const int* meta)
{
    if (n_iblocks <= 0) return cudaSuccess;
    const int wpb = 8;
    reduce_kernel<<<(n_iblocks + wpb - 1) / wpb, wpb * 32, 0, s>>>(n_iblocks, iblocks, part4, partn, out, G, meta);
    return cudaGetLastError();
}

} // namespace pb
