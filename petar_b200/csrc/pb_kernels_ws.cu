// pb_kernels_ws.cu — the warp-specialised persistent force kernel (sm_100a).
//
// Same arithmetic as pb::force_kernel (pb_pairs.cuh), another division of labour inside the CTA:
//
//   * warps 0..7 are COMPUTE warps.  They never stage j, never meet a block-wide barrier and never fetch a task:
//     they wait for a "tile full" mbarrier, run the pair loops on the tile, arrive on its "tile empty" mbarrier, and
//     at a task's end combine / write their partial sums among themselves (a named barrier for the 256 of them).
//   * warps 8 and 9 are PRODUCERS.  They pull task numbers from the atomic cursor, publish the task and walk records
//     in a two-slot shared-memory ring, and stream the task's j tiles into a four-stage tile ring: coalesced index
//     reads (one tile ahead), 16-byte gathers of the j records from the L2-resident store, origin shift + pair
//     interleave + near test (the staging code of pb_pairs.cuh, four j per lane and tile), then a release-arrive on
//     the stage's mbarrier.  They run up to four tiles (and one task) ahead of the compute warps, so the dependent
//     global loads at a task's start (task -> walk -> index list -> j records) overlap the previous task's arithmetic.
//
// mbarrier phases run on across tasks (both sides count tiles and tasks globally), nothing is re-initialised.
// Launched with 2 CTAs per SM (320 threads, <= 96 registers) by the device-resident tree step (pb_tree_force_resident);
// the functor path's dispatches and the neighbour search keep pb::force_kernel.
//
// Template switches: TWOI — SP tasks of groups with >= 2 i-blocks give every compute warp two blocks (one particle of each
// per lane, sp_pairs_2i) and half as much of every tile; FUSE — the warp that delivers the last partial sum of an i-block
// reduces the block and writes its forces to page-locked host memory (finish_block in pb_pairs.cuh), so the launch is
// followed by neither a reduction kernel nor a D2H copy.
#include "pb_pairs.cuh"

namespace pb {

namespace {

constexpr int kWsStages   = 4;
constexpr int kWsProducers = 2;                              // producer warps per CTA (8 rounds of 32 j per tile are dealt out among them)
constexpr int kWsThreads  = (kWarpsPerCta + kWsProducers) * 32;   // 8 compute warps + the producers
constexpr int kTileBytes  = sizeof(SpTile) > sizeof(EpTile) ? sizeof(SpTile) : sizeof(EpTile);

struct WsCtl {
    alignas(8) unsigned long long full[kWsStages], empty[kWsStages], tfull[2], tempty[2];
    int   near_flag[kWsStages][kWarpsPerCta];
    Task  task[2];
    Walk  walk[2];
    int   next_task[2];
    double red[kWarpsPerCta][4][32];
    int    redn[kWarpsPerCta][32];
    double red2[kWarpsPerCta][4][32];                        // second i-particle of a lane (SP tasks with two i-blocks per warp)
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "WS_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra WS_DONE;\n"
                 "bra WS_WAIT;\n"
                 "WS_DONE:\n"
                 "}" :: "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void compute_bar() {            // the 256 compute threads only
    asm volatile("bar.sync 1, 256;" ::: "memory");
}
__device__ __forceinline__ void producer_bar() {           // the producer warps only
    asm volatile("bar.sync 2, %0;" :: "n"(kWsProducers * 32) : "memory");
}

} // namespace

template <int NR, int TWOI, int FUSE>
__global__ void __maxnreg__(96)
force_kernel_ws(const Walk* __restrict__ walks, const Task* __restrict__ tasks,
                const float4* __restrict__ epi,
                const int* __restrict__ id_epj, const int* __restrict__ id_spj,
                const float4* __restrict__ epj, const float4* __restrict__ spj,
                double4* __restrict__ part4, int* __restrict__ partn, Params prm)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* tiles = smem_raw;                                         // kWsStages tiles
    WsCtl& ctl = *reinterpret_cast<WsCtl*>(smem_raw + (size_t)kWsStages * kTileBytes);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kWsStages; ++s) { mbar_init(&ctl.full[s], kWsProducers * 32); mbar_init(&ctl.empty[s], kThreads); }
        for (int s = 0; s < 2; ++s) { mbar_init(&ctl.tfull[s], 1); mbar_init(&ctl.tempty[s], kThreads); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_tasks = prm.meta[0];

    if (warp >= kWarpsPerCta) {
        // ============================== producer warps ==============================
        // each of the kWsProducers warps stages its share of every tile: rounds q = pw, pw + P, ... of 32 j
        const int pw = warp - kWarpsPerCta;
        unsigned g = 0;                     // tiles produced so far (stage = g % S, phase = (g / S) & 1)
        unsigned tcount = 0;                // tasks published so far
        for (;;) {
            const int ts = tcount & 1;
            if (pw == 0) {
                int t = 0;
                if (lane == 0) t = atomicAdd(prm.meta + 4, 1);
                t = __shfl_sync(0xffffffffu, t, 0);
                if (lane == 0) ctl.next_task[ts] = t;
            }
            producer_bar();                                                    // task number visible to every producer warp
            const int t = ctl.next_task[ts];
            Task task;
            if (t < n_tasks) task = tasks[t]; else { task.kind = -1; task.walk = 0; task.j_count = 0; task.j_begin = 0; task.nib = 1; task.jsplit = 8; task.i_first = 0; task.part_base = 0; task.blk0 = 0; }
            Walk w;
            if (t < n_tasks) w = walks[task.walk];
            if (pw == 0) {
                mbar_wait(&ctl.tempty[ts], ((tcount >> 1) & 1) ^ 1);          // compute warps have read the slot's previous task
                if (lane == 0) {
                    ctl.task[ts] = task;
                    if (t < n_tasks) ctl.walk[ts] = w;
                    __threadfence_block();
                    mbar_arrive(&ctl.tfull[ts]);
                }
            }
            tcount++;
            if (t >= n_tasks) break;

            const int n_tiles = (task.j_count + kTileJ - 1) / kTileJ;
            constexpr int R = 8 / kWsProducers;                                // rounds of 32 j per producer warp and tile
            if (task.kind == 0) {
                const int* ids = (w.ej_off >= 0) ? id_epj + w.ej_off + task.j_begin : nullptr;   // ej_off < 0: dense list
                int id[R], idn[R];
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const int j = (q * kWsProducers + pw) * 32 + lane;
                    idn[q] = (j < task.j_count) ? (ids ? __ldg(ids + j) : task.j_begin + j) : -1;
                }
                for (int k = 0; k < n_tiles; ++k, ++g) {
                    const int s = g % kWsStages;
                    EpRegs r[R];
#pragma unroll
                    for (int q = 0; q < R; ++q) { id[q] = idn[q]; r[q] = ep_load_j(epj, id[q]); }
#pragma unroll
                    for (int q = 0; q < R; ++q) {                               // ids of the next tile travel under this tile's work
                        const int j = (k + 1) * kTileJ + (q * kWsProducers + pw) * 32 + lane;
                        idn[q] = (j < task.j_count) ? (ids ? __ldg(ids + j) : task.j_begin + j) : -1;
                    }
                    mbar_wait(&ctl.empty[s], ((g / kWsStages) & 1) ^ 1);                   // stage free again
                    EpTile& T = *reinterpret_cast<EpTile*>(tiles + (size_t)s * kTileBytes);
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        const int seg = q * kWsProducers + pw;
                        const bool nr_ = ep_store(T, seg * 32 + lane, id[q], r[q], w, prm.abs_mode);
                        const unsigned bal = __ballot_sync(0xffffffffu, nr_);
                        if (lane == 0) ctl.near_flag[s][seg] = (bal != 0u);
                    }
                    mbar_arrive(&ctl.full[s]);
                }
            } else {
                const int* ids = id_spj + w.sj_off + task.j_begin;
                int id[R], idn[R];
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const int j = (q * kWsProducers + pw) * 32 + lane;
                    idn[q] = (j < task.j_count) ? __ldg(ids + j) : -1;
                }
                for (int k = 0; k < n_tiles; ++k, ++g) {
                    const int s = g % kWsStages;
                    SpRegs r[R];
#pragma unroll
                    for (int q = 0; q < R; ++q) { id[q] = idn[q]; r[q] = sp_load_j(spj, id[q]); }
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        const int j = (k + 1) * kTileJ + (q * kWsProducers + pw) * 32 + lane;
                        idn[q] = (j < task.j_count) ? __ldg(ids + j) : -1;
                    }
                    mbar_wait(&ctl.empty[s], ((g / kWsStages) & 1) ^ 1);
                    SpTile& T = *reinterpret_cast<SpTile*>(tiles + (size_t)s * kTileBytes);
#pragma unroll
                    for (int q = 0; q < R; ++q) sp_store(T, (q * kWsProducers + pw) * 32 + lane, id[q], r[q], w, prm.eps2);
                    mbar_arrive(&ctl.full[s]);
                }
            }
        }
        return;
    }

    // ============================== compute warps ==============================
    unsigned g = 0, tcount = 0;
    for (;;) {
        const int ts = tcount & 1;
        mbar_wait(&ctl.tfull[ts], (tcount >> 1) & 1);
        const Task task = ctl.task[ts];
        if (task.kind < 0) break;
        const Walk w = ctl.walk[ts];
        mbar_arrive(&ctl.tempty[ts]);
        tcount++;

        // warp role: i-block `ib` of the group, j-split slot `js` (nib * jsplit == 8)
        const int  ib   = warp % task.nib;
        const int  js   = warp / task.nib;
        const int  ppw  = kTilePairs / task.jsplit;
        // a block with at most 16 (8) real i-particles lets 2 (4) lanes share each particle (see pb::force_kernel)
        const int  n_blk  = min(32, w.ni - (task.i_first + ib * 32));
        const int  ishift = (n_blk <= 8) ? 2 : (n_blk <= 16) ? 1 : 0;
        const int  lanes_i = 32 >> ishift;
        const int  il = lane & (lanes_i - 1), sub = lane >> (5 - ishift);
        const int  i_loc  = task.i_first + ib * 32 + il;
        const bool ivalid = (i_loc < w.ni);
        float4 pi = make_float4(0.f, 0.f, 0.f, 0.f), pil = make_float4(0.f, 0.f, 0.f, 0.f), pih = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ivalid) {
            const float4* rec = epi + (size_t)prm.i_f4 * (size_t)(w.i_off + i_loc);
            pi  = __ldg(rec);
            pil = __ldg(rec + 1);
            if (prm.abs_mode == 2) pih = __ldg(rec + 2);
        }
        const float rsi2 = ivalid ? pi.w * pi.w : -1.f;

        KSum kx, ky, kz, kp;
        kx.init(); ky.init(); kz.init(); kp.init();
        int cnt = 0;
        const int n_tiles = (task.j_count + kTileJ - 1) / kTileJ;

        if (task.kind == 0) {
            for (int k = 0; k < n_tiles; ++k, ++g) {
                const int s = g % kWsStages;
                mbar_wait(&ctl.full[s], (g / kWsStages) & 1);
                const EpTile& T = *reinterpret_cast<const EpTile*>(tiles + (size_t)s * kTileBytes);
                const int nv  = min(kTileJ, task.j_count - k * kTileJ);
                const int npu = (((nv + 1) >> 1) + kPairUnroll - 1) & ~(kPairUnroll - 1);
                const int ppk = (nv == kTileJ) ? ppw : max(16, (((npu + task.jsplit - 1) / task.jsplit) + 15) & ~15);
                const int p0  = js * ppk, p1 = min(p0 + ppk, npu);
                float2 ax = bc(0.f), ay = bc(0.f), az = bc(0.f), pt = bc(0.f), cf = bc(0.f);
                for (int seg0 = p0; seg0 < p1; seg0 += 16) {
                    const int seg = seg0 + sub * (16 >> ishift);
                    const int e = min(min(seg0 + 16, p1), seg + (16 >> ishift));
                    if (ctl.near_flag[s][seg0 >> 4]) {
                        if (prm.abs_mode == 2)
                            ep_pairs<NR, 2>(T, seg, e, pi.x, pi.y, pi.z, pil.x, pil.y, pil.z, rsi2, pih.x, pih.y, pih.z, prm.eps2, prm.rcut2, prm.rinv_cut, ax, ay, az, pt, cf);
                        else
                            ep_pairs<NR, 1>(T, seg, e, pi.x, pi.y, pi.z, pil.x, pil.y, pil.z, rsi2, 0.f, 0.f, 0.f, prm.eps2, prm.rcut2, 0.f, ax, ay, az, pt, cf);
                    } else
                        ep_pairs<NR, 0>(T, seg, e, pi.x, pi.y, pi.z, pil.x, pil.y, pil.z, rsi2, 0.f, 0.f, 0.f, prm.eps2, prm.rcut2, 0.f, ax, ay, az, pt, cf);
                }
                mbar_arrive(&ctl.empty[s]);
                kx.add(ax.x + ax.y); ky.add(ay.x + ay.y); kz.add(az.x + az.y); kp.add(pt.x + pt.y);
                cnt += (int)(cf.x + cf.y);
            }
        } else if (TWOI && task.nib >= 2) {
            // SP task of a group with nib >= 2 i-blocks: every warp takes TWO blocks (b0 and b0 + nib/2, one particle of each
            // per lane) and 1/(2 jsplit) of every j tile, instead of one block and 1/jsplit of the tile
            const int npair = task.nib >> 1, jsp = 2 * task.jsplit;
            const int b0 = warp % npair, js2 = warp / npair;
            const int ia = task.i_first + b0 * 32 + lane, ib2 = ia + npair * 32;
            float xi[2] = {0.f, 0.f}, yi[2] = {0.f, 0.f}, zi[2] = {0.f, 0.f};
            if (ia < w.ni)  { const float4 q = __ldg(epi + (size_t)prm.i_f4 * (size_t)(w.i_off + ia));  xi[0] = q.x; yi[0] = q.y; zi[0] = q.z; }
            if (ib2 < w.ni) { const float4 q = __ldg(epi + (size_t)prm.i_f4 * (size_t)(w.i_off + ib2)); xi[1] = q.x; yi[1] = q.y; zi[1] = q.z; }
            KSum k2x, k2y, k2z, k2p;
            k2x.init(); k2y.init(); k2z.init(); k2p.init();
            for (int k = 0; k < n_tiles; ++k, ++g) {
                const int s = g % kWsStages;
                mbar_wait(&ctl.full[s], (g / kWsStages) & 1);
                const SpTile& T = *reinterpret_cast<const SpTile*>(tiles + (size_t)s * kTileBytes);
                const int nv  = min(kTileJ, task.j_count - k * kTileJ);
                const int npu = ((nv + 1) >> 1);
                const int ppk = (nv == kTileJ) ? kTilePairs / jsp : (npu + jsp - 1) / jsp;
                const int p0  = js2 * ppk, p1 = min(p0 + ppk, npu);
                float2 ax[2] = {bc(0.f), bc(0.f)}, ay[2] = {bc(0.f), bc(0.f)}, az[2] = {bc(0.f), bc(0.f)}, pt[2] = {bc(0.f), bc(0.f)};
                sp_pairs_2i<NR>(T, p0, p1, xi, yi, zi, prm.eps2, ax, ay, az, pt);
                mbar_arrive(&ctl.empty[s]);
                kx.add(ax[0].x + ax[0].y); ky.add(ay[0].x + ay[0].y); kz.add(az[0].x + az[0].y); kp.add(pt[0].x + pt[0].y);
                k2x.add(ax[1].x + ax[1].y); k2y.add(ay[1].x + ay[1].y); k2z.add(az[1].x + az[1].y); k2p.add(pt[1].x + pt[1].y);
            }
            double d0x = kx.value(), d0y = ky.value(), d0z = kz.value(), d0p = kp.value();
            double d1x = k2x.value(), d1y = k2y.value(), d1z = k2z.value(), d1p = k2p.value();
            // combine the warps that share a block pair in fixed order js2 = 0,1,...; warp b0 writes both blocks
            ctl.red[warp][0][lane] = d0x; ctl.red[warp][1][lane] = d0y; ctl.red[warp][2][lane] = d0z; ctl.red[warp][3][lane] = d0p;
            ctl.red2[warp][0][lane] = d1x; ctl.red2[warp][1][lane] = d1y; ctl.red2[warp][2][lane] = d1z; ctl.red2[warp][3][lane] = d1p;
            compute_bar();
            if (js2 == 0) {
                for (int s = 1; s < jsp; ++s) {
                    const int ww = s * npair + b0;
                    d0x += ctl.red[ww][0][lane]; d0y += ctl.red[ww][1][lane]; d0z += ctl.red[ww][2][lane]; d0p += ctl.red[ww][3][lane];
                    d1x += ctl.red2[ww][0][lane]; d1y += ctl.red2[ww][1][lane]; d1z += ctl.red2[ww][2][lane]; d1p += ctl.red2[ww][3][lane];
                }
                const int slot = task.part_base + b0 * 32 + lane;
                part4[slot] = make_double4(d0x, d0y, d0z, d0p);
                partn[slot] = 0;
                part4[slot + npair * 32] = make_double4(d1x, d1y, d1z, d1p);
                partn[slot + npair * 32] = 0;
            }
            compute_bar();                                  // scratch free for the next task
            if (FUSE && js2 == 0) {
                finish_block(task.blk0 + b0, task.n_chunks, lane, part4, partn, prm);
                finish_block(task.blk0 + b0 + npair, task.n_chunks, lane, part4, partn, prm);
            }
            continue;
        } else {
            for (int k = 0; k < n_tiles; ++k, ++g) {
                const int s = g % kWsStages;
                mbar_wait(&ctl.full[s], (g / kWsStages) & 1);
                const SpTile& T = *reinterpret_cast<const SpTile*>(tiles + (size_t)s * kTileBytes);
                const int nv  = min(kTileJ, task.j_count - k * kTileJ);
                const int npu = ((nv + 1) >> 1);
                const int ppk = (nv == kTileJ) ? ppw : (npu + task.jsplit - 1) / task.jsplit;
                const int p0  = js * ppk, p1 = min(p0 + ppk, npu);
                float2 ax = bc(0.f), ay = bc(0.f), az = bc(0.f), pt = bc(0.f);
                const int plen = (max(p1 - p0, 0) + (1 << ishift) - 1) >> ishift;
                const int q0 = p0 + sub * plen, q1 = min(p1, q0 + plen);
                sp_pairs<NR>(T, q0, q1, pi.x, pi.y, pi.z, prm.eps2, ax, ay, az, pt);
                mbar_arrive(&ctl.empty[s]);
                kx.add(ax.x + ax.y); ky.add(ay.x + ay.y); kz.add(az.x + az.y); kp.add(pt.x + pt.y);
            }
        }

        // per-warp totals as exact doubles (hi + lo); lanes that shared a particle: fixed-order butterfly
        double dax = kx.value(), day = ky.value(), daz = kz.value(), dpt = kp.value();
        for (int o = lanes_i; o < 32; o <<= 1) {
            dax += __shfl_xor_sync(0xffffffffu, dax, o); day += __shfl_xor_sync(0xffffffffu, day, o);
            daz += __shfl_xor_sync(0xffffffffu, daz, o); dpt += __shfl_xor_sync(0xffffffffu, dpt, o);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        if (task.jsplit > 1) {
            // combine the jsplit warps that share an i-block, in fixed order js = 0,1,... (compute warps only)
            ctl.red[warp][0][lane] = dax; ctl.red[warp][1][lane] = day;
            ctl.red[warp][2][lane] = daz; ctl.red[warp][3][lane] = dpt;
            ctl.redn[warp][lane] = cnt;
            compute_bar();
            if (js == 0) {
                for (int s = 1; s < task.jsplit; ++s) {
                    const int ww = s * task.nib + ib;
                    dax += ctl.red[ww][0][lane]; day += ctl.red[ww][1][lane];
                    daz += ctl.red[ww][2][lane]; dpt += ctl.red[ww][3][lane];
                    cnt += ctl.redn[ww][lane];
                }
            }
            compute_bar();                                  // scratch free for the next task
        }
        if (js == 0 && sub == 0) {
            const int slot = task.part_base + ib * 32 + il;
            part4[slot] = make_double4(dax, day, daz, dpt);
            partn[slot] = cnt;
        }
        if (FUSE && js == 0) finish_block(task.blk0 + ib, task.n_chunks, lane, part4, partn, prm);
    }
}

size_t ws_smem_bytes() { return (size_t)kWsStages * kTileBytes + sizeof(WsCtl); }

cudaError_t launch_force_ws(cudaStream_t s, int n_ctas, int nr_steps,
                            const Walk* walks, const Task* tasks, const float4* epi, const int* id_epj, const int* id_spj,
                            const float4* epj, const float4* spj, double4* part4, int* partn, Params p, bool two_i)
{
    if (n_ctas <= 0) return cudaSuccess;
    static bool configured = false;
    const size_t smem = ws_smem_bytes();
    if (!configured) {
        cudaError_t e = cudaSuccess;
        const void* fns[] = {(const void*)force_kernel_ws<0, 0, 0>, (const void*)force_kernel_ws<1, 0, 0>, (const void*)force_kernel_ws<0, 1, 0>, (const void*)force_kernel_ws<1, 1, 0>,
                             (const void*)force_kernel_ws<0, 0, 1>, (const void*)force_kernel_ws<1, 0, 1>, (const void*)force_kernel_ws<0, 1, 1>, (const void*)force_kernel_ws<1, 1, 1>};
        for (const void* f : fns)
            if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
#define PB_WS_LAUNCH(NR_, TI_, FU_) force_kernel_ws<NR_, TI_, FU_><<<n_ctas, kWsThreads, smem, s>>>(walks, tasks, epi, id_epj, id_spj, epj, spj, part4, partn, p)
    const int nr = nr_steps >= 1 ? 1 : 0;
    const bool fuse = p.out != nullptr;                   // fused reduction: Params carries the i-block table, counters and the result array
    if (fuse) {
        if (two_i) { if (nr) PB_WS_LAUNCH(1, 1, 1); else PB_WS_LAUNCH(0, 1, 1); }
        else       { if (nr) PB_WS_LAUNCH(1, 0, 1); else PB_WS_LAUNCH(0, 0, 1); }
    } else {
        if (two_i) { if (nr) PB_WS_LAUNCH(1, 1, 0); else PB_WS_LAUNCH(0, 1, 0); }
        else       { if (nr) PB_WS_LAUNCH(1, 0, 0); else PB_WS_LAUNCH(0, 0, 0); }
    }
#undef PB_WS_LAUNCH
    return cudaGetLastError();
}

} // namespace pb
