// ref_cuda_driver.cu — BENCH / TEST INFRASTRUCTURE ONLY: the informational second baseline of
// SURVEY.md §8c ("reference CUDA kernels ... compile -arch=sm_100 to time the *reference design*
// on B200; not a parity oracle").
//
// The reference's device code (structs EpiGPU/EpjGPU/SpjGPU/ForceGPU, dev_gravity_ep_ep/_ep_sp,
// force_kernel_ep_ep, force_kernel_ep_sp — reference src/force_gpu_cuda.cu:12-511) is
// self-contained; its host functors are not (FDPS types).  oracle/Makefile extracts that device
// section from the file where it lies under /root/reference into the git-ignored
// oracle/_ref/ref_cuda_kernels.inc at build time (nothing of it enters this repository), and this
// driver restates the reference's host side around it over the POD mirrors:
//   send phase      reference src/force_gpu_cuda.cu:577-620  (fp64 -> fp32 AoS repack, 2 H2D)
//   dispatch phase  reference :621-699  (ij_disp prefix sums, EPI repack with id_walk, index lists,
//                                        4 H2D, <<<ni_tot_reg/32, 32>>> x 2 kernels)
//   retrieve        reference :831-880  (D2H, widen, assign)
// Compiled natively for sm_100a (the reference's Makefile passes no -arch and would JIT compute_52
// PTX, Makefile.in:177) — this favours the reference.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <chrono>
#include <cuda_runtime.h>

#define USE_QUAD
#define PARTICLE_SIMULATOR_GPU_MULIT_WALK_INDEX
#include "_ref/ref_cuda_kernels.inc"

#include "petar_b200_types.h"

#define RC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { std::fprintf(stderr, "ref_cuda: %s: %s\n", #call, cudaGetErrorString(e_)); return -1; } } while (0)

namespace {
template <class T> struct Buf {
    T* h = nullptr; T* d = nullptr; size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return 0;
        if (h) cudaFreeHost(h);
        if (d) cudaFree(d);
        cap = n + n / 2 + 1024;
        if (cudaMallocHost(&h, cap * sizeof(T)) != cudaSuccess) return -1;
        if (cudaMalloc(&d, cap * sizeof(T)) != cudaSuccess) return -1;
        return 0;
    }
};
Buf<EpiGPU> b_epi; Buf<EpjGPU> b_epj; Buf<SpjGPU> b_spj; Buf<ForceGPU> b_force; Buf<int3> b_disp; Buf<int> b_ide, b_ids;
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}

extern "C" {

// One full tree step the way the reference does it: send all j, then per group of
// `n_walk_limit` walks dispatch + retrieve (blocking copies on the default stream).
// ms_out[0] = kernel time (CUDA events around the two launches, summed over dispatches),
// ms_out[1] = wall-clock of the whole step (packing + copies + kernels).
int refcuda_step(int n_walk, const pb_EPISoft* const* epi, const int* n_epi,
                 const int* const* id_epj, const int* n_epj, const int* const* id_spj, const int* n_spj,
                 const pb_EPJSoft* epj, int n_epj_tot, const pb_SPJQuad* spj, int n_spj_tot,
                 pb_ForceSoft* const* force, double eps2, double rcut2, double G, int n_walk_limit, float* ms_out)
{
    const double t_begin = now_s();
    cudaEvent_t e0, e1;
    RC(cudaEventCreate(&e0)); RC(cudaEventCreate(&e1));
    float ms_k = 0.f;

    // ---- send phase (:577-620) ----
    if (b_epj.ensure(n_epj_tot) || b_spj.ensure(n_spj_tot)) return -1;
#pragma omp parallel for
    for (int i = 0; i < n_epj_tot; i++) {
        b_epj.h[i].pos.x = epj[i].pos.x; b_epj.h[i].pos.y = epj[i].pos.y; b_epj.h[i].pos.z = epj[i].pos.z;
        b_epj.h[i].m = epj[i].mass; b_epj.h[i].r_search = epj[i].r_search;
    }
#pragma omp parallel for
    for (int i = 0; i < n_spj_tot; i++) {
        b_spj.h[i].pos.x = spj[i].pos.x; b_spj.h[i].pos.y = spj[i].pos.y; b_spj.h[i].pos.z = spj[i].pos.z;
        b_spj.h[i].m = spj[i].mass;
        b_spj.h[i].qxx = spj[i].qxx; b_spj.h[i].qyy = spj[i].qyy; b_spj.h[i].qzz = spj[i].qzz;
        b_spj.h[i].qxy = spj[i].qxy; b_spj.h[i].qxz = spj[i].qxz; b_spj.h[i].qyz = spj[i].qyz;
    }
    RC(cudaMemcpy(b_epj.d, b_epj.h, sizeof(EpjGPU) * n_epj_tot, cudaMemcpyHostToDevice));
    RC(cudaMemcpy(b_spj.d, b_spj.h, sizeof(SpjGPU) * n_spj_tot, cudaMemcpyHostToDevice));

    // ---- per walk group: dispatch (:621-699) + retrieve (:831-880) ----
    for (int w0 = 0; w0 < n_walk; w0 += n_walk_limit) {
        const int nw = (n_walk - w0 < n_walk_limit) ? n_walk - w0 : n_walk_limit;
        if (b_disp.ensure(nw + 2)) return -1;
        int3* disp = b_disp.h;
        disp[0] = make_int3(0, 0, 0);
        for (int k = 0; k < nw; k++)
            disp[k + 1] = make_int3(disp[k].x + n_epi[w0 + k], disp[k].y + n_epj[w0 + k], disp[k].z + n_spj[w0 + k]);
        disp[nw + 1] = disp[nw];
        const int ni_tot = disp[nw].x, nej_tot = disp[nw].y, nsj_tot = disp[nw].z;
        int ni_tot_reg = ni_tot;
        if (ni_tot_reg % N_THREAD_GPU) ni_tot_reg = (ni_tot_reg / N_THREAD_GPU + 1) * N_THREAD_GPU;
        // +N_THREAD_GPU: the reference kernel reads id lists up to 31 entries past the end (:100)
        if (b_epi.ensure(ni_tot_reg) || b_force.ensure(ni_tot_reg) || b_ide.ensure(nej_tot + N_THREAD_GPU) || b_ids.ensure(nsj_tot + N_THREAD_GPU)) return -1;
#pragma omp parallel for schedule(dynamic)
        for (int iw = 0; iw < nw; iw++) {
            const int w = w0 + iw;
            for (int i = 0; i < n_epi[w]; i++) {
                EpiGPU& e = b_epi.h[i + disp[iw].x];
                e.pos.x = epi[w][i].pos.x; e.pos.y = epi[w][i].pos.y; e.pos.z = epi[w][i].pos.z;
                e.r_search = epi[w][i].r_search; e.id_walk = iw;
            }
            for (int j = 0; j < n_epj[w]; j++) b_ide.h[j + disp[iw].y] = id_epj[w][j];
            for (int j = 0; j < n_spj[w]; j++) b_ids.h[j + disp[iw].z] = id_spj[w][j];
        }
        for (int i = ni_tot; i < ni_tot_reg; i++) { b_epi.h[i].id_walk = nw; b_epi.h[i].pos = make_float3(0.f, 0.f, 0.f); b_epi.h[i].r_search = 0.f; }
        for (int j = 0; j < N_THREAD_GPU; j++) { b_ide.h[nej_tot + j] = 0; b_ids.h[nsj_tot + j] = 0; }
        RC(cudaMemcpy(b_disp.d, b_disp.h, sizeof(int3) * (nw + 2), cudaMemcpyHostToDevice));
        RC(cudaMemcpy(b_epi.d, b_epi.h, sizeof(EpiGPU) * ni_tot_reg, cudaMemcpyHostToDevice));
        RC(cudaMemcpy(b_ide.d, b_ide.h, sizeof(int) * (nej_tot + N_THREAD_GPU), cudaMemcpyHostToDevice));
        RC(cudaMemcpy(b_ids.d, b_ids.h, sizeof(int) * (nsj_tot + N_THREAD_GPU), cudaMemcpyHostToDevice));
        const int nblocks = ni_tot_reg / N_THREAD_GPU;
        RC(cudaEventRecord(e0));
        force_kernel_ep_ep<<<nblocks, N_THREAD_GPU>>>(b_disp.d, b_epi.d, b_epj.d, b_ide.d, b_force.d, (float)eps2, (float)rcut2, (float)G);
        force_kernel_ep_sp<<<nblocks, N_THREAD_GPU>>>(b_disp.d, b_epi.d, b_spj.d, b_ids.d, b_force.d, (float)eps2, (float)G);
        RC(cudaEventRecord(e1));
        RC(cudaGetLastError());
        RC(cudaMemcpy(b_force.h, b_force.d, sizeof(ForceGPU) * ni_tot, cudaMemcpyDeviceToHost));
        float ms = 0.f;
        RC(cudaEventElapsedTime(&ms, e0, e1));
        ms_k += ms;
        int n_cnt = 0;
        for (int iw = 0; iw < nw; iw++)
            for (int i = 0; i < n_epi[w0 + iw]; i++) {
                pb_ForceSoft& f = force[w0 + iw][i];
                f.acc.x = b_force.h[n_cnt].accp.x; f.acc.y = b_force.h[n_cnt].accp.y; f.acc.z = b_force.h[n_cnt].accp.z;
                f.pot = b_force.h[n_cnt].accp.w; f.n_ngb = b_force.h[n_cnt].nnb;
                n_cnt++;
            }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    ms_out[0] = ms_k;
    ms_out[1] = (float)((now_s() - t_begin) * 1e3);
    return 0;
}

} // extern "C"
