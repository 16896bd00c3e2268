// microbench4.cu — the MEASURED non-tensor FP32 peak of this part (profiles/r2_fp32_peak.json) and what register
// operand forms cost: FFMA2 with 64-bit pair operands vs 32-bit scalar-broadcast (.F32) operands, distinct registers
// per instruction (no operand-reuse-cache hits), scalar FFMA for comparison.  Rates are warp-instructions per clock
// and SM over a whole launch (CUDA events + the SM clock sampled by clock64 on every CTA), 8/16/32 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench4 tools/microbench4.cu && tools/bin/microbench4
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

#define CH 8
#define ITERS 8192

template <int KIND>
__global__ void probe(const float2* __restrict__ in, float* out, long long* cycles) {
    float2 d[CH], x[CH], y[CH];
    float s[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) {
        d[k] = in[threadIdx.x + 32 * k];
        x[k] = in[threadIdx.x + 32 * (k + CH)];
        y[k] = in[threadIdx.x + 32 * (k + 2 * CH)];
        s[k] = in[threadIdx.x + 32 * (k + 3 * CH)].x;
    }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < CH; k++) {
            if (KIND == 0) d[k] = __ffma2_rn(d[k], make_float2(1.0009765625f, 1.0009765625f), make_float2(0.5f, 0.5f)); // FFMA2 reg, imm... (may fold to 1 imm)
            if (KIND == 1) d[k] = __ffma2_rn(x[k], make_float2(2.5f, 2.5f), d[k]);              // FFMA2 64, imm, 64
            if (KIND == 2) d[k] = __ffma2_rn(x[k], y[k], d[k]);                                   // FFMA2 64, 64, 64 (3 distinct)
            if (KIND == 3) d[k] = __ffma2_rn(x[k], make_float2(s[k], s[k]), d[k]);                // FFMA2 64, .F32 (distinct per k), 64
            if (KIND == 4) d[k] = __ffma2_rn(make_float2(s[k], s[k]), make_float2(s[(k + 1) % CH], s[(k + 1) % CH]), d[k]);   // FFMA2 .F32, .F32, 64
            if (KIND == 5) d[k] = __fadd2_rn(d[k], make_float2(s[k], s[k]));                      // FADD2 64, .F32
            if (KIND == 6) d[k] = __fmul2_rn(x[k], make_float2(s[k], s[k]));                      // FMUL2 64, .F32 -> kept alive below
            if (KIND == 7) d[k].x = fmaf(x[k].x, 2.5f, d[k].x);                                   // FFMA reg, imm, reg
            if (KIND == 8) { d[k].x = fmaf(x[k].x, y[k].x, d[k].x); d[k].y = fmaf(x[k].y, y[k].y, d[k].y); }   // 2 FFMA, 3 distinct each
            if (KIND == 9) d[k] = __ffma2_rn(d[k], d[k], x[k]);                                   // FFMA2 a, a, c (2 distinct)
        }
        if (KIND == 6) {
#pragma unroll
            for (int k = 0; k < CH; k++) x[k].x = d[(k + 1) % CH].y;
        }
    }
    const long long t1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < CH; k++) r += d[k].x + d[k].y + x[k].x;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

struct Res { double winst_clk_sm, tflops_event; };

template <int KIND>
Res run(const char* name, int inst_per_step, int flop_per_lane_inst, int warps_per_sm) {
    float* out; long long* cyc; float2* in;
    const int threads = 32 * warps_per_sm, grid = 148;
    cudaMalloc(&out, sizeof(float) * grid * threads);
    cudaMalloc(&cyc, sizeof(long long) * grid);
    cudaMalloc(&in, sizeof(float2) * 32 * (4 * CH + 1) + sizeof(float2) * threads);
    cudaMemset(in, 0, sizeof(float2) * 32 * (4 * CH + 1) + sizeof(float2) * threads);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    probe<KIND><<<grid, threads>>>(in, out, cyc);
    cudaEventRecord(a);
    probe<KIND><<<grid, threads>>>(in, out, cyc);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms = 0.f; cudaEventElapsedTime(&ms, a, b);
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double c = 0; for (auto v : h) c += (double)v; c /= grid;
    const double winst = (double)ITERS * CH * inst_per_step * warps_per_sm;
    Res r;
    r.winst_clk_sm = winst / c;
    r.tflops_event = winst * 32.0 * flop_per_lane_inst * grid / (ms * 1e-3) * 1e-12;
    printf("%-40s warps/SM %2d : %.3f warp-inst/clk/SM  (%.2f clk per inst per SMSP)  %6.2f TFLOP/s by CUDA events (%.0f MHz implied)\n",
           name, warps_per_sm, r.winst_clk_sm, 4.0 / r.winst_clk_sm, r.tflops_event, c / (ms * 1e-3) * 1e-6);
    cudaFree(out); cudaFree(cyc); cudaFree(in);
    return r;
}

int main() {
    double best2 = 0, best1 = 0, best_tf = 0;
    for (int w : {8, 16, 32}) {
        Res r;
        r = run<1>("FFMA2 x_k * imm + d_k", 1, 4, w); if (r.winst_clk_sm > best2) { best2 = r.winst_clk_sm; best_tf = r.tflops_event; }
        run<9>("FFMA2 d_k * d_k + x_k (2 distinct)", 1, 4, w);
        run<2>("FFMA2 x_k * y_k + d_k (3 distinct 64-bit)", 1, 4, w);
        run<3>("FFMA2 x_k * s_k.F32 + d_k (64,32,64)", 1, 4, w);
        run<4>("FFMA2 s_k.F32 * s_k'.F32 + d_k (32,32,64)", 1, 4, w);
        run<5>("FADD2 d_k + s_k.F32", 1, 2, w);
        run<6>("FMUL2 x_k * s_k.F32", 1, 2, w);
        r = run<7>("FFMA x_k * imm + d_k", 1, 2, w); if (r.winst_clk_sm > best1) best1 = r.winst_clk_sm;
        r = run<8>("2 FFMA (halves) x_k * y_k + d_k", 2, 2, w); if (r.winst_clk_sm > best1) best1 = r.winst_clk_sm;
        printf("\n");
    }
    printf("PEAK_JSON {\"ffma2_warp_inst_per_clk_sm\": %.4f, \"ffma_warp_inst_per_clk_sm\": %.4f, \"ffma2_equiv_ffma_per_clk_sm\": %.4f, "
           "\"lanes_per_clk_sm\": %.1f, \"fp32_tflops_measured\": %.2f}\n", best2, best1, 2 * best2, 64 * best2, best_tf);
    return 0;
}
