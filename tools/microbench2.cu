// microbench2.cu — register-operand-read probes for packed FP32 on sm_100a: how the FFMA2 rate depends on
// the number of distinct register operands and on operand-reuse between consecutive instructions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench2 tools/microbench2.cu && ./microbench2
#include <cstdio>
#include <cuda_runtime.h>

#define CH 8
#define ITERS 4096

template <int KIND>
__global__ void probe(const float2* __restrict__ in, float* out, long long* cycles) {
    float2 d[CH], x[CH], y[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) {
        d[k] = in[threadIdx.x + 32 * k];
        x[k] = in[threadIdx.x + 32 * (k + CH)];
        y[k] = in[threadIdx.x + 32 * (k + 2 * CH)];
    }
    const float s0 = in[threadIdx.x + 32 * 3 * CH].x, s1 = in[threadIdx.x + 32 * 3 * CH].y;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < CH; k++) {
            if (KIND == 0) d[k] = __ffma2_rn(x[k], y[k], d[k]);                    // 3 distinct 64-bit operands, no sharing
            if (KIND == 1) d[k] = __ffma2_rn(x[0], y[k], d[k]);                    // slot A shared by all
            if (KIND == 2) d[k] = __ffma2_rn(x[k >> 1], y[k], d[k]);               // slot A shared by consecutive pairs
            if (KIND == 3) d[k] = __ffma2_rn(x[k >> 1], y[k & 1], d[k]);           // slot A shared by pairs, slot B alternates
            if (KIND == 4) d[k] = __ffma2_rn(x[k], make_float2(s0, s0), d[k]);     // scalar-broadcast B
            if (KIND == 5) d[k] = __ffma2_rn(x[0], make_float2(s0, s0), d[k]);     // shared A, scalar-broadcast B
            if (KIND == 6) d[k] = __ffma2_rn(x[k], make_float2(2.5f, 2.5f), d[k]); // immediate B
            if (KIND == 7) d[k] = __fmul2_rn(x[k], y[k]);                          // FMUL2, 2 distinct (dead-code guarded below)
            if (KIND == 8) d[k].x = fmaf(x[k].x, y[k].x, d[k].x);                  // scalar FFMA, 3 distinct
            if (KIND == 9) d[k].x = fmaf(x[0].x, y[k].x, d[k].x);                  // scalar FFMA, shared A
            if (KIND == 10) { d[k].x = fmaf(x[k].x, y[k].x, d[k].x); d[k].y = fmaf(x[k].y, y[k].y, d[k].y); }  // 2 scalar FFMA = 1 FFMA2 of work
            if (KIND == 11) d[k] = __ffma2_rn(d[k], d[k], x[0]);                   // a*a+c
            if (KIND == 12) d[k] = __fadd2_rn(d[k], x[k]);                         // FADD2 2 distinct
            if (KIND == 13) d[k] = __fadd2_rn(d[k], make_float2(s0, s0));          // FADD2 scalar-broadcast
        }
        if (KIND == 7) {
#pragma unroll
            for (int k = 0; k < CH; k++) x[k].x = d[(k + 1) % CH].y;               // keep the FMUL2 alive without adding FMA-pipe work
        }
    }
    const long long t1 = clock64();
    float s = s1;
#pragma unroll
    for (int k = 0; k < CH; k++) s += d[k].x + d[k].y + x[k].x;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int KIND>
void run(const char* name, int inst_per_step, int warps_per_sm) {
    float* out; long long* cyc; long long h = 0; float2* in;
    const int threads = 32 * warps_per_sm;
    cudaMalloc(&out, sizeof(float) * 148 * threads);
    cudaMalloc(&cyc, sizeof(long long));
    cudaMalloc(&in, sizeof(float2) * 32 * (3 * CH + 1) + sizeof(float2) * threads);
    cudaMemset(in, 0, sizeof(float2) * 32 * (3 * CH + 1) + sizeof(float2) * threads);
    probe<KIND><<<148, threads>>>(in, out, cyc);
    probe<KIND><<<148, threads>>>(in, out, cyc);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double winst = (double)ITERS * CH * inst_per_step * warps_per_sm;
    printf("%-44s warps/SM %2d : %.3f warp-inst/clk/SM  (%.2f clk per warp-inst per SMSP)\n", name, warps_per_sm, winst / h, h / (winst / 4));
    cudaFree(out); cudaFree(cyc); cudaFree(in);
}

int main() {
    for (int w : {8, 16}) {
        run<0>("FFMA2 x_k*y_k+d_k (3 distinct)", 1, w);
        run<1>("FFMA2 x*y_k+d_k (A shared by all)", 1, w);
        run<2>("FFMA2 x_{k/2}*y_k+d_k (A shared by pairs)", 1, w);
        run<3>("FFMA2 x_{k/2}*y_{k&1}+d_k", 1, w);
        run<4>("FFMA2 x_k*bcast(s)+d_k", 1, w);
        run<5>("FFMA2 x*bcast(s)+d_k", 1, w);
        run<6>("FFMA2 x_k*imm+d_k", 1, w);
        run<7>("FMUL2 x_k*y_k", 1, w);
        run<8>("FFMA x_k*y_k+d_k", 1, w);
        run<9>("FFMA x*y_k+d_k", 1, w);
        run<10>("2 FFMA (halves) x_k*y_k+d_k", 2, w);
        run<11>("FFMA2 d*d+x", 1, w);
        run<12>("FADD2 d_k+x_k", 1, w);
        run<13>("FADD2 d_k+bcast(s)", 1, w);
        printf("\n");
    }
    return 0;
}
