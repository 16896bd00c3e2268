// loopbench.cu — steady-state rate of the force kernel's INNER LOOPS on sm_100a, without staging, barriers or task
// overheads: every warp sweeps a shared-memory j tile (filled once) again and again for its own i-particles.
// Answers (profiles/r2_loopbench.txt): how far the shipped kernel is from its own loops' ceiling, and what the
// alternative loop formulations (i-packed FFMA2, two i-particles per lane) would buy before they are built.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I petar_b200/csrc -I include -o tools/bin/loopbench tools/loopbench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../petar_b200/csrc/pb_kernels.cu"

using namespace pb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// ---- alternative formulations under test ---------------------------------------------------------------------------
// (A) two i-particles per lane, j packed in pairs as shipped: pb::sp_pairs_2i (pb_pairs.cuh; shipped in the warp-specialised kernel)

// (B) i-packed: each lane holds TWO i-particles as one float2, every j enters as a scalar-broadcast (.F32) operand:
//     the j operands are 32-bit register reads instead of 64-bit ones.  j come from the same tile, one at a time.
template <int NR, int UNR>
__device__ __forceinline__ void sp_ipacked(const SpTile& t, int p0, int p1, float2 xi, float2 yi, float2 zi, float eps2,
                                           float2& ax, float2& ay, float2& az, float2& pt) {
    const float2 e2 = bc(eps2);
#pragma unroll UNR
    for (int j = 2 * p0; j < 2 * p1; ++j) {
        const int p = j >> 1, s = j & 1;
        const float* q0 = reinterpret_cast<const float*>(&t.q0[p]);
        const float* q1 = reinterpret_cast<const float*>(&t.q1[p]);
        const float* q2 = reinterpret_cast<const float*>(&t.q2[p]);
        const float* q3 = reinterpret_cast<const float*>(&t.q3[p]);
        const float* q4 = reinterpret_cast<const float*>(&t.q4[p]);
        const float* q5 = reinterpret_cast<const float*>(&t.q5[p]);
        const float xj = q0[s], yj = q0[2 + s], zj = q1[s], mj = q1[2 + s];
        const float qxx = q2[s], qyy = q2[2 + s], qzz = q3[s], qxy = q3[2 + s], qxz = q4[s], qyz = q4[2 + s], mtr = q5[s];
        const float2 dx = __fadd2_rn(xi, bc(-xj)), dy = __fadd2_rn(yi, bc(-yj)), dz = __fadd2_rn(zi, bc(-zj));
        float2 r2 = __ffma2_rn(dx, dx, e2); r2 = __ffma2_rn(dy, dy, r2); r2 = __ffma2_rn(dz, dz, r2);
        const float2 rinv = rsqrt2<NR>(r2);
        float2 qrx = __fmul2_rn(bc(qxx), dx), qry = __fmul2_rn(bc(qxy), dx), qrz = __fmul2_rn(bc(qxz), dx);
        qrx = __ffma2_rn(bc(qxy), dy, qrx); qry = __ffma2_rn(bc(qyy), dy, qry); qrz = __ffma2_rn(bc(qyz), dy, qrz);
        qrx = __ffma2_rn(bc(qxz), dz, qrx); qry = __ffma2_rn(bc(qyz), dz, qry); qrz = __ffma2_rn(bc(qzz), dz, qrz);
        float2 S = __ffma2_rn(qrx, dx, bc(mtr)); S = __ffma2_rn(qry, dy, S); S = __ffma2_rn(qrz, dz, S);
        const float2 rinv2 = __fmul2_rn(rinv, rinv);
        const float2 rinv3 = __fmul2_rn(rinv2, rinv);
        const float2 rinv5 = __fmul2_rn(rinv3, rinv2);
        const float2 mr3   = __fmul2_rn(bc(mj), rinv3);
        const float2 S5    = __fmul2_rn(rinv5, S);
        const float2 S7    = __fmul2_rn(S5, rinv2);
        const float2 A     = __ffma2_rn(bc(2.5f), S7, mr3);
        const float2 nA    = make_float2(-A.x, -A.y);
        ax = __ffma2_rn(nA, dx, ax); ay = __ffma2_rn(nA, dy, ay); az = __ffma2_rn(nA, dz, az);
        ax = __ffma2_rn(rinv5, qrx, ax); ay = __ffma2_rn(rinv5, qry, ay); az = __ffma2_rn(rinv5, qrz, az);
        pt = __ffma2_rn(bc(0.5f), S5, pt);
        pt = __ffma2_rn(bc(mj), rinv, pt);
    }
}

template <int NR, int UNR>
__device__ __forceinline__ void ep_ipacked(const EpTile& t, int p0, int p1, float2 xi, float2 yi, float2 zi, float eps2, float rcut2,
                                           float2& ax, float2& ay, float2& az, float2& pt) {
    const float2 e2 = bc(eps2);
#pragma unroll UNR
    for (int j = 2 * p0; j < 2 * p1; ++j) {
        const int p = j >> 1, s = j & 1;
        const float* a = reinterpret_cast<const float*>(&t.a[p]);
        const float* b = reinterpret_cast<const float*>(&t.b[p]);
        const float xj = a[s], yj = a[2 + s], zj = b[s], mj = b[2 + s];
        const float2 dx = __fadd2_rn(bc(xj), make_float2(-xi.x, -xi.y)), dy = __fadd2_rn(bc(yj), make_float2(-yi.x, -yi.y)), dz = __fadd2_rn(bc(zj), make_float2(-zi.x, -zi.y));
        float2 r2 = __ffma2_rn(dx, dx, e2); r2 = __ffma2_rn(dy, dy, r2); r2 = __ffma2_rn(dz, dz, r2);
        const float2 r2c = make_float2(fmaxf(r2.x, rcut2), fmaxf(r2.y, rcut2));
        const float2 ri = rsqrt2<NR>(r2c);
        const float2 pij = __fmul2_rn(bc(mj), ri);
        const float2 ri2 = __fmul2_rn(ri, ri);
        const float2 mri3 = __fmul2_rn(pij, ri2);
        ax = __ffma2_rn(mri3, dx, ax); ay = __ffma2_rn(mri3, dy, ay); az = __ffma2_rn(mri3, dz, az);
        pt = __fadd2_rn(pt, pij);
    }
}

// (C) the shipped SP loop with the three-register accumulate/contract operations issued as scalar FFMA halves
template <int NR>
__device__ __forceinline__ void sp_pairs_mixed(const SpTile& t, int p0, int p1, float xi, float yi, float zi, float eps2,
                                               float2& ax, float2& ay, float2& az, float2& pt) {
    const float2 vxi = bc(xi), vyi = bc(yi), vzi = bc(zi), e2 = bc(eps2);
#pragma unroll 2
    for (int p = p0; p < p1; ++p) {
        const float4 Q0 = t.q0[p], Q1 = t.q1[p], Q2 = t.q2[p], Q3 = t.q3[p], Q4 = t.q4[p];
        const float2 mtr = t.q5[p];
        const float2 mj = make_float2(Q1.z, Q1.w);
        const float2 qxx = make_float2(Q2.x, Q2.y), qyy = make_float2(Q2.z, Q2.w);
        const float2 qzz = make_float2(Q3.x, Q3.y), qxy = make_float2(Q3.z, Q3.w);
        const float2 qxz = make_float2(Q4.x, Q4.y), qyz = make_float2(Q4.z, Q4.w);
        const float2 dx = __fadd2_rn(vxi, make_float2(-Q0.x, -Q0.y));
        const float2 dy = __fadd2_rn(vyi, make_float2(-Q0.z, -Q0.w));
        const float2 dz = __fadd2_rn(vzi, make_float2(-Q1.x, -Q1.y));
        float2 r2 = __ffma2_rn(dx, dx, e2); r2 = __ffma2_rn(dy, dy, r2); r2 = __ffma2_rn(dz, dz, r2);
        const float2 rinv = rsqrt2<NR>(r2);
        float2 qrx = __fmul2_rn(qxx, dx), qry = __fmul2_rn(qxy, dx), qrz = __fmul2_rn(qxz, dx);
#define SFMA(d, a, b) d.x = fmaf(a.x, b.x, d.x); d.y = fmaf(a.y, b.y, d.y)
        SFMA(qrx, qxy, dy); SFMA(qry, qyy, dy); SFMA(qrz, qyz, dy);
        SFMA(qrx, qxz, dz); SFMA(qry, qyz, dz); SFMA(qrz, qzz, dz);
        float2 S = __ffma2_rn(qrx, dx, mtr); SFMA(S, qry, dy); SFMA(S, qrz, dz);
        const float2 rinv2 = __fmul2_rn(rinv, rinv);
        const float2 rinv3 = __fmul2_rn(rinv2, rinv);
        const float2 rinv5 = __fmul2_rn(rinv3, rinv2);
        const float2 mr3   = __fmul2_rn(mj, rinv3);
        const float2 S5    = __fmul2_rn(rinv5, S);
        const float2 S7    = __fmul2_rn(S5, rinv2);
        const float2 A     = __ffma2_rn(bc(2.5f), S7, mr3);
        const float2 nA    = make_float2(-A.x, -A.y);
        SFMA(ax, nA, dx); SFMA(ay, nA, dy); SFMA(az, nA, dz);
        SFMA(ax, rinv5, qrx); SFMA(ay, rinv5, qry); SFMA(az, rinv5, qrz);
        pt = __ffma2_rn(bc(0.5f), S5, pt);
        SFMA(pt, mj, rinv);
#undef SFMA
    }
}

// ---- the harness kernel ------------------------------------------------------------------------------------------------
// KIND 0: shipped EP far loop, 1: shipped EP near loop, 2: shipped SP loop, 3: SP two i per lane, 4: SP i-packed,
//      5: EP i-packed, 6: SP mixed scalar/packed, 7: SP i-packed unroll 4
template <int KIND, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
loop_kernel(const float* __restrict__ init, int iters, int sync_every, float* __restrict__ out, long long* __restrict__ cycles) {
    __shared__ Smem sm;
    const int tid = threadIdx.x;
    float* raw = reinterpret_cast<float*>(&sm);
    for (int k = tid; k < (int)(sizeof(SpTile) / 4); k += kThreads) raw[k] = init[k];
    __syncthreads();
    const float xi = init[4096 + tid], yi = init[4096 + 256 + tid], zi = init[4096 + 512 + tid];
    const float xi1 = xi + 0.37f, yi1 = yi - 0.21f, zi1 = zi + 0.11f;
    const float eps2 = init[8190], rcut2 = init[8191];
    float2 ax = bc(0.f), ay = bc(0.f), az = bc(0.f), pt = bc(0.f), cf = bc(0.f);
    float2 ax2[2] = {bc(0.f), bc(0.f)}, ay2[2] = {bc(0.f), bc(0.f)}, az2[2] = {bc(0.f), bc(0.f)}, pt2[2] = {bc(0.f), bc(0.f)};
    const float xi2[2] = {xi, xi1}, yi2[2] = {yi, yi1}, zi2[2] = {zi, zi1};
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (KIND == 0) ep_pairs<0, 0>(sm.ep[0], 0, kTilePairs, xi, yi, zi, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, eps2, rcut2, 0.f, ax, ay, az, pt, cf);
        if (KIND == 1) ep_pairs<0, 1>(sm.ep[0], 0, kTilePairs, xi, yi, zi, 1e-9f, 1e-9f, 1e-9f, 1.f, 0.f, 0.f, 0.f, eps2, rcut2, 0.f, ax, ay, az, pt, cf);
        if (KIND == 2) sp_pairs<0>(sm.sp[0], 0, kTilePairs, xi, yi, zi, eps2, ax, ay, az, pt);
        if (KIND == 3) sp_pairs_2i<0>(sm.sp[0], 0, kTilePairs, xi2, yi2, zi2, eps2, ax2, ay2, az2, pt2);
        if (KIND == 4) sp_ipacked<0, 2>(sm.sp[0], 0, kTilePairs, make_float2(xi, xi1), make_float2(yi, yi1), make_float2(zi, zi1), eps2, ax, ay, az, pt);
        if (KIND == 5) ep_ipacked<0, 4>(sm.ep[0], 0, kTilePairs, make_float2(xi, xi1), make_float2(yi, yi1), make_float2(zi, zi1), eps2, rcut2, ax, ay, az, pt);
        if (KIND == 6) sp_pairs_mixed<0>(sm.sp[0], 0, kTilePairs, xi, yi, zi, eps2, ax, ay, az, pt);
        if (KIND == 7) sp_ipacked<0, 4>(sm.sp[0], 0, kTilePairs, make_float2(xi, xi1), make_float2(yi, yi1), make_float2(zi, zi1), eps2, ax, ay, az, pt);
        if (sync_every && (it % sync_every) == sync_every - 1) __syncthreads();
    }
    const long long t1 = clock64();
    float s = ax.x + ax.y + ay.x + ay.y + az.x + az.y + pt.x + pt.y + cf.x + cf.y;
    for (int k = 0; k < 2; k++) s += ax2[k].x + ax2[k].y + ay2[k].x + ay2[k].y + az2[k].x + az2[k].y + pt2[k].x + pt2[k].y;
    out[blockIdx.x * kThreads + tid] = s;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

struct Case { const char* name; int kind; double inter_per_lane_per_tile; double fma_ops_per_tile; };

template <int KIND, int MINB>
void run(const char* name, double inter_per_lane_per_tile, double fma_inst_per_tile, int ctas_per_sm, int sync_every, const float* d_init) {
    const int iters = 400, grid = 148 * ctas_per_sm;
    float* out; long long* cyc;
    CK(cudaMalloc(&out, sizeof(float) * grid * kThreads));
    CK(cudaMalloc(&cyc, sizeof(long long) * grid));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    loop_kernel<KIND, MINB><<<grid, kThreads>>>(d_init, iters, sync_every, out, cyc);
    CK(cudaEventRecord(a));
    loop_kernel<KIND, MINB><<<grid, kThreads>>>(d_init, iters, sync_every, out, cyc);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms = 0.f; CK(cudaEventElapsedTime(&ms, a, b));
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    double cmean = 0; for (auto v : h) cmean += (double)v; cmean /= grid;
    // per SM sub-partition: warps = 8 * ctas_per_sm / 4; each warp issues fma_inst_per_tile FMA-pipe instructions per sweep
    const double warps_per_smsp = 8.0 * ctas_per_sm / 4.0;
    const double clk_per_fma = cmean / (iters * fma_inst_per_tile * warps_per_smsp);
    const double inter = (double)grid * kThreads * iters * inter_per_lane_per_tile;
    printf("%-34s CTAs/SM %d sync %d : %7.3f ms  %8.1f Ginter/s  %.3f clk per FMA-pipe inst per SMSP (pipe floor 2.0 -> %.1f %% of the pipe)\n",
           name, ctas_per_sm, sync_every, ms, inter / (ms * 1e-3) * 1e-9, clk_per_fma, 200.0 / clk_per_fma);
    CK(cudaFree(out)); CK(cudaFree(cyc));
}

int main() {
    std::vector<float> init(8192);
    srand(1);
    for (auto& v : init) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (int k = 4096; k < 4096 + 768; k++) init[k] = 3.f + init[k];          // i-particles away from the j cloud
    init[8190] = 1e-6f; init[8191] = 1e-4f;
    float* d_init; CK(cudaMalloc(&d_init, sizeof(float) * 8192));
    CK(cudaMemcpy(d_init, init.data(), sizeof(float) * 8192, cudaMemcpyHostToDevice));
    for (int sync : {0, 1}) {
        for (int c : {1, 2}) {
            run<0, 2>("EP far (shipped)", 256, 128 * 13, c, sync, d_init);
            run<1, 2>("EP near (shipped)", 256, 128 * 20, c, sync, d_init);
            run<5, 2>("EP i-packed (2 i per lane)", 512, 256 * 13, c, sync, d_init);
            run<2, 2>("SP (shipped)", 256, 128 * 33, c, sync, d_init);
            run<6, 2>("SP mixed scalar/packed", 256, 128 * 33, c, sync, d_init);
            run<3, 2>("SP two i per lane, j pairs", 512, 256 * 33, c, sync, d_init);
            run<4, 2>("SP i-packed (2 i per lane) u2", 512, 256 * 33, c, sync, d_init);
            run<7, 2>("SP i-packed (2 i per lane) u4", 512, 256 * 33, c, sync, d_init);
        }
        printf("\n");
    }
    run<2, 3>("SP (shipped), 3 CTAs/SM (80 regs)", 256, 128 * 33, 3, 0, d_init);
    run<4, 3>("SP i-packed u2, 3 CTAs/SM", 512, 256 * 33, 3, 0, d_init);
    run<0, 3>("EP far (shipped), 3 CTAs/SM", 256, 128 * 13, 3, 0, d_init);
    return 0;
}
