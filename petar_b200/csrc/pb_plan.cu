// pb_plan.cu — the host-free middle of the device-resident tree step (pb_tree_force_resident).
//
// After the device walk (pb_walk.cu) has produced every i-group's interaction lists, two things the host did in
// round 1 are done here on the GPU, so that a tree step is "upload tree + particles -> wait for forces" with no
// host round trip in between (the round trip cost ~9 ms of a 12.9 ms step at 8 ranks per node):
//
//   iprep_kernel  one warp per i-group: the group's i-particles are its own slice of the EP j store (the local particles
//                 are uploaded in i-group order), so they are taken from there — walk origin = mean position (fp64, fixed
//                 summation order), positions relative to it as two-float pairs by the same rel_hilo() the force kernel
//                 applies to the j side, bounding box / near radius of the walk, the Walk record itself.
//   plan_kernel   one block: per group the binary 8/4/2/1 decomposition of its 32-wide i-blocks and the number of j
//                 chunks (whole 256-entry tiles, about `U` entries per warp) -> exclusive scans -> task / partial-sum /
//                 i-block offsets, totals and capacity checks.
//   emit_kernel   one thread per group: writes the Task and IBlock records at those offsets.
//
// The force kernel then runs PERSISTENT (2 CTAs per SM pulling task numbers from an atomic cursor) because the task
// count lives in device memory.  Plan arithmetic mirrors plan_batch() in pb_engine.cu.
#include "pb_device.h"
#include "petar_b200.h"

namespace pb {

namespace {

__device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

// chunks of one list of nj entries for an i-group whose warps split every tile `jsplit` ways; u = entries per warp-task
__device__ __forceinline__ void chunking(int nj, int u, int jsplit, int& n_chunks, int& len) {
    if (nj <= 0) { n_chunks = 0; len = kTileJ; return; }
    const int nc = ceil_div(nj, u * jsplit);
    len = ceil_div(ceil_div(nj, nc), kTileJ) * kTileJ;
    n_chunks = ceil_div(nj, len);
}

struct GroupPlan { int n_tasks, n_part, n_iblk; };

__device__ __forceinline__ GroupPlan plan_group(int ni, int nej, int nsj, int U, int Us) {
    GroupPlan r = {0, 0, 0};
    int nib = (ni + 31) >> 5;
    r.n_iblk = nib;
    while (nib >= kWarpsPerCta) {
        int ce, cs, l;
        chunking(nej, U, 1, ce, l); chunking(nsj, Us, 1, cs, l);
        r.n_tasks += ce + cs; r.n_part += (ce + cs) * kWarpsPerCta * 32;
        nib -= kWarpsPerCta;
    }
    for (int g = kWarpsPerCta / 2; g >= 1; g >>= 1)
        if (nib & g) {
            int ce, cs, l;
            chunking(nej, U, kWarpsPerCta / g, ce, l); chunking(nsj, Us, kWarpsPerCta / g, cs, l);
            r.n_tasks += ce + cs; r.n_part += (ce + cs) * g * 32;
        }
    return r;
}

} // namespace

// ------------------------------------------------------------------------------------------------------------------
// i-particles and Walk records from the j store
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
iprep_kernel(const pb_tree_group* __restrict__ groups, int n_groups, const int* __restrict__ i_first,
             const int2* __restrict__ counts, const int2* __restrict__ offs, const float4* __restrict__ epj,
             Walk* __restrict__ walks, float4* __restrict__ epi, int i_f4, int coords, int cull)
{
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= n_groups) return;
    const int n = groups[g].n, s0 = groups[g].first, i0 = i_first[g];
    const bool rel = (coords != 1);

    // origin = mean position, accumulated in fp64 lane by lane, then a fixed butterfly: deterministic
    float oh[3] = {0.f, 0.f, 0.f}, ol[3] = {0.f, 0.f, 0.f};
    if (rel && n > 0) {
        double sx = 0.0, sy = 0.0, sz = 0.0;
        for (int k = lane; k < n; k += 32) {
            const float4 a = __ldg(epj + 2 * (size_t)(s0 + k)), b = __ldg(epj + 2 * (size_t)(s0 + k) + 1);
            sx += (double)a.x + (double)b.x; sy += (double)a.y + (double)b.y; sz += (double)a.z + (double)b.z;
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); sz += __shfl_xor_sync(0xffffffffu, sz, o);
        }
        const double m[3] = {sx / n, sy / n, sz / n};
#pragma unroll
        for (int k = 0; k < 3; k++) { oh[k] = (float)m[k]; ol[k] = (float)(m[k] - (double)oh[k]); }
    }
    float hx = 0.f, hy = 0.f, hz = 0.f, rsmax = 0.f;
    for (int k = lane; k < n; k += 32) {
        const float4 a = __ldg(epj + 2 * (size_t)(s0 + k)), b = __ldg(epj + 2 * (size_t)(s0 + k) + 1);
        float x, y, z, xl, yl, zl;
        rel_hilo(a.x, b.x, oh[0], ol[0], x, xl);
        rel_hilo(a.y, b.y, oh[1], ol[1], y, yl);
        rel_hilo(a.z, b.z, oh[2], ol[2], z, zl);
        if (!rel) { xl = 0.f; yl = 0.f; zl = 0.f; }
        float4* rec = epi + (size_t)i_f4 * (size_t)(i0 + k);
        rec[0] = make_float4(x, y, z, b.w);
        rec[1] = make_float4(xl, yl, zl, 0.f);
        if (i_f4 == 3) rec[2] = make_float4(a.x, a.y, a.z, 0.f);
        hx = fmaxf(hx, fabsf(x)); hy = fmaxf(hy, fabsf(y)); hz = fmaxf(hz, fabsf(z)); rsmax = fmaxf(rsmax, b.w);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        hx = fmaxf(hx, __shfl_xor_sync(0xffffffffu, hx, o)); hy = fmaxf(hy, __shfl_xor_sync(0xffffffffu, hy, o));
        hz = fmaxf(hz, __shfl_xor_sync(0xffffffffu, hz, o)); rsmax = fmaxf(rsmax, __shfl_xor_sync(0xffffffffu, rsmax, o));
    }
    if (lane == 0) {
        Walk W;
        const int2 c = counts[g], o = offs[g];
        W.i_off = i0; W.ni = n; W.ej_off = o.x; W.nej = c.x; W.sj_off = o.y; W.nsj = c.y;
        W.ohx = oh[0]; W.ohy = oh[1]; W.ohz = oh[2]; W.olx = ol[0]; W.oly = ol[1]; W.olz = ol[2];
        const float inf = __int_as_float(0x7f800000);
        if (rel && cull) { W.hx = hx; W.hy = hy; W.hz = hz; } else { W.hx = W.hy = W.hz = inf; }
        // "near" radius: max r_search of the i-particles, and at least 4e4 ulps of the box half-size (see pack_walk)
        const float hbox = fmaxf(hx, fmaxf(hy, hz));
        const float rprec = rel ? 4.0e4f * (__int_as_float(__float_as_int(hbox) + 1) - hbox) : 0.f;
        const float rnear = fmaxf(rsmax, rprec);
        W.rsi2max = rnear * rnear;
        walks[g] = W;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// task planning
// ------------------------------------------------------------------------------------------------------------------
// meta[0] tasks to run (0 on any overflow), [1] n_part, [2] n_iblocks, [3] overflow flags (1: tasks, 2: partial sums, 4: a list
// reservation), [4] task cursor (zeroed here), [5] n_tasks
__global__ void __launch_bounds__(1024)
plan_kernel(const pb_tree_group* __restrict__ groups, int n_groups, const int2* __restrict__ counts,
            int U, int Us, int3* __restrict__ goff, int* __restrict__ meta, int cap_tasks, long long cap_part,
            const int2* __restrict__ caps)
{
    __shared__ int s_t[1024], s_p[1024], s_b[1024];
    __shared__ int carry[3];
    __shared__ int list_over;
    const int tid = threadIdx.x;
    if (tid == 0) { carry[0] = carry[1] = carry[2] = 0; list_over = 0; }
    __syncthreads();
    for (int base = 0; base < n_groups; base += 1024) {
        const int g = base + tid;
        GroupPlan gp = {0, 0, 0};
        if (g < n_groups) {
            const int2 c = counts[g];
            gp = plan_group(groups[g].n, c.x, c.y, U, Us);
            if (caps && (c.x > caps[g].x || c.y > caps[g].y)) list_over = 1;   // a list outgrew its reservation: it was not written past it
        }
        s_t[tid] = gp.n_tasks; s_p[tid] = gp.n_part; s_b[tid] = gp.n_iblk;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {                       // Hillis-Steele inclusive scan, three lanes of data
            int a = 0, b = 0, c = 0;
            if (tid >= d) { a = s_t[tid - d]; b = s_p[tid - d]; c = s_b[tid - d]; }
            __syncthreads();
            s_t[tid] += a; s_p[tid] += b; s_b[tid] += c;
            __syncthreads();
        }
        if (g < n_groups) goff[g] = make_int3(carry[0] + s_t[tid] - gp.n_tasks, carry[1] + s_p[tid] - gp.n_part, carry[2] + s_b[tid] - gp.n_iblk);
        __syncthreads();
        if (tid == 1023) { carry[0] += s_t[1023]; carry[1] += s_p[1023]; carry[2] += s_b[1023]; }
        __syncthreads();
    }
    if (tid == 0) {
        int flags = 0;
        if (carry[0] > cap_tasks) flags |= 1;
        if ((long long)carry[1] > cap_part || carry[1] < 0) flags |= 2;
        if (list_over) flags |= 4;
        meta[0] = flags ? 0 : carry[0];      // nothing runs on an overflow: the host re-sizes and repeats the step
        meta[1] = carry[1]; meta[2] = carry[2]; meta[3] = flags; meta[4] = 0; meta[5] = carry[0];
    }
}

__global__ void __launch_bounds__(128)
emit_kernel(const pb_tree_group* __restrict__ groups, int n_groups, const int* __restrict__ i_first,
            const int2* __restrict__ counts, const int3* __restrict__ goff, const int* __restrict__ meta,
            int U, int Us, Task* __restrict__ tasks, IBlock* __restrict__ iblocks, ForceOut* __restrict__ out_fused)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups || meta[3] != 0) return;
    const int ni = groups[g].n, nej = counts[g].x, nsj = counts[g].y, i0 = i_first[g];
    const int3 o = goff[g];
    int t = o.x, part = o.y, ib_out = o.z;
    int nib_left = (ni + 31) >> 5, ib = 0;
    for (int step = 0; nib_left > 0; step++) {
        int nib;
        if (nib_left >= kWarpsPerCta) nib = kWarpsPerCta;
        else { nib = kWarpsPerCta / 2; while (!(nib_left & nib)) nib >>= 1; }
        const int jsplit = kWarpsPerCta / nib, stride = nib * 32, part_base = part;
        int n_chunks_grp;
        { int ce, cs, l; chunking(nej, U, jsplit, ce, l); chunking(nsj, Us, jsplit, cs, l); n_chunks_grp = ce + cs; }
        int chunk = 0;
        for (int kind = 0; kind < 2; kind++) {
            const int nj = kind == 0 ? nej : nsj;
            int nc, len;
            chunking(nj, kind == 0 ? U : Us, jsplit, nc, len);
            for (int c = 0; c < nc; c++) {
                Task T;
                T.walk = g; T.i_first = ib * 32; T.nib = nib; T.jsplit = jsplit; T.kind = kind;
                T.j_begin = c * len; T.j_count = min(len, nj - c * len); T.part_base = part_base + chunk * stride;
                T.blk0 = ib_out; T.n_chunks = n_chunks_grp; T.pad1 = T.pad2 = 0;
                tasks[t++] = T;
                chunk++;
            }
        }
        for (int b = 0; b < nib; b++) {
            IBlock B;
            B.part_base = part_base + b * 32; B.n_chunks = chunk; B.stride = stride;
            B.out_off = i0 + (ib + b) * 32; B.n_valid = min(32, ni - (ib + b) * 32);
            B.pad0 = B.pad1 = B.pad2 = 0;
            iblocks[ib_out++] = B;
            if (chunk == 0 && out_fused) {                 // a group without any list entry: nobody will deliver a last chunk (fused reduction)
                ForceOut z; z.ax = z.ay = z.az = z.pot = 0.0; z.n_ngb = 0;
                for (int l = 0; l < B.n_valid; l++) out_fused[B.out_off + l] = z;
            }
        }
        part += chunk * stride;
        ib += nib; nib_left -= nib;
    }
}

cudaError_t launch_iprep(cudaStream_t s, const void* groups, int n_groups, const int* i_first, const int2* counts, const int2* offs,
                         const float4* epj, Walk* walks, float4* epi, int i_f4, int coords, int cull) {
    if (n_groups <= 0) return cudaSuccess;
    iprep_kernel<<<(n_groups + 3) / 4, 128, 0, s>>>((const pb_tree_group*)groups, n_groups, i_first, counts, offs, epj, walks, epi, i_f4, coords, cull);
    return cudaGetLastError();
}

cudaError_t launch_devplan(cudaStream_t s, const void* groups, int n_groups, const int* i_first, const int2* counts, int U, int Us,
                           int3* goff, int* meta, int cap_tasks, long long cap_part, Task* tasks, IBlock* iblocks, const int2* caps, ForceOut* out_fused) {
    if (n_groups <= 0) return cudaSuccess;
    plan_kernel<<<1, 1024, 0, s>>>((const pb_tree_group*)groups, n_groups, counts, U, Us, goff, meta, cap_tasks, cap_part, caps);
    emit_kernel<<<(n_groups + 127) / 128, 128, 0, s>>>((const pb_tree_group*)groups, n_groups, i_first, counts, goff, meta, U, Us, tasks, iblocks, out_fused);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// LET send rows gathered on the device: out[k] = EP store row idx[k] (32 B rows as two float4)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_epj_kernel(const float4* __restrict__ epj, int n_epj, const int* __restrict__ idx, int n, float4* __restrict__ out, int* __restrict__ err) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n) return;
    const int id = idx[t >> 1];
    if ((unsigned)id >= (unsigned)n_epj) {                // reported by the next call that synchronises with the device
        if (err) atomicExch(err, 1);
        out[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    out[t] = __ldg(epj + 2 * (size_t)id + (t & 1));
}

cudaError_t launch_gather_epj(cudaStream_t s, const float4* epj, int n_epj, const int* idx, int n, float4* out, int* err) {
    if (n <= 0) return cudaSuccess;
    gather_epj_kernel<<<(2 * n + 255) / 256, 256, 0, s>>>(epj, n_epj, idx, n, out, err);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// EP index lists that crossed PCIe as runs (start, length): one warp per walk writes the indices back out, in list
// order, where the force kernel reads them (FDPS's EP lists are leaf cells in Morton order: ~20 consecutive indices
// per run, so 8 B per run instead of 4 B per index)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
expand_runs_kernel(const int2* __restrict__ runtab, const int2* __restrict__ runs, const Walk* __restrict__ walks, int n_walk, int* __restrict__ out) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_walk) return;
    const int2 rt = runtab[w];
    const int ej = walks[w].ej_off;
    if (ej < 0) return;                                   // dense walk: no list
    int* o = out + ej;
    int pos = 0;
    for (int base = 0; base < rt.y; base += 32) {
        int2 r = make_int2(0, 0);
        if (base + lane < rt.y) r = runs[rt.x + base + lane];
        int incl = r.y;                                   // inclusive scan of the run lengths
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
        int* dst = o + pos + incl - r.y;                  // every lane writes out its own run
        for (int q = 0; q < r.y; q++) dst[q] = r.x + q;
        pos += __shfl_sync(0xffffffffu, incl, 31);
    }
}

cudaError_t launch_expand_runs(cudaStream_t s, const int2* runtab, const int2* runs, const Walk* walks, int n_walk, int* out) {
    if (n_walk <= 0) return cudaSuccess;
    expand_runs_kernel<<<(n_walk + 3) / 4, 128, 0, s>>>(runtab, runs, walks, n_walk, out);
    return cudaGetLastError();
}

// host mirror of plan_group, for sizing the task / partial-sum buffers when the list lengths are known on the host
void plan_sizes_host(const int* ni, const int2* counts, int n_groups, int U, int Us, long long* n_tasks, long long* n_part, long long* n_iblk) {
    long long t = 0, p = 0, b = 0;
    auto chunks = [](int nj, int u, int jsplit) {
        if (nj <= 0) return 0;
        const int nc = (nj + u * jsplit - 1) / (u * jsplit);
        const int len = (((nj + nc - 1) / nc) + kTileJ - 1) / kTileJ * kTileJ;
        return (nj + len - 1) / len;
    };
    for (int g = 0; g < n_groups; g++) {
        int nib = (ni[g] + 31) >> 5;
        b += nib;
        while (nib >= kWarpsPerCta) {
            const int c = chunks(counts[g].x, U, 1) + chunks(counts[g].y, Us, 1);
            t += c; p += (long long)c * kWarpsPerCta * 32; nib -= kWarpsPerCta;
        }
        for (int k = kWarpsPerCta / 2; k >= 1; k >>= 1)
            if (nib & k) {
                const int c = chunks(counts[g].x, U, kWarpsPerCta / k) + chunks(counts[g].y, Us, kWarpsPerCta / k);
                t += c; p += (long long)c * k * 32;
            }
    }
    *n_tasks = t; *n_part = p; *n_iblk = b;
}

} // namespace pb
