// microbench3.cu — legacy warp-level MMA (mma.sync m16n8k4 / m16n8k8 tf32) rate on sm_100a, alone and next to FFMA2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench3 tools/microbench3.cu && ./microbench3
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define CH 8

__device__ __forceinline__ void mma_k4(float (&c)[4], unsigned a0, unsigned a1, unsigned b0) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void mma_k8(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int KIND>
__global__ void probe(const float* __restrict__ in, float* out, long long* cycles) {
    float c[CH][4];
    float2 d[CH];
    unsigned a[4], b[2];
    for (int k = 0; k < 4; k++) a[k] = __float_as_uint(in[threadIdx.x + 32 * k]);
    for (int k = 0; k < 2; k++) b[k] = __float_as_uint(in[threadIdx.x + 32 * (4 + k)]);
    for (int k = 0; k < CH; k++) { for (int q = 0; q < 4; q++) c[k][q] = in[threadIdx.x + 32 * (6 + q)]; d[k] = make_float2(c[k][0], c[k][1]); }
    const float2 x = make_float2(in[threadIdx.x + 320], in[threadIdx.x + 352]), y = make_float2(in[threadIdx.x + 384], in[threadIdx.x + 416]);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < CH; k++) {
            if (KIND == 0) mma_k4(c[k], a[0], a[1], b[0]);
            if (KIND == 1) mma_k8(c[k], a[0], a[1], a[2], a[3], b[0], b[1]);
            if (KIND == 2) { mma_k4(c[k], a[0], a[1], b[0]); d[k] = __ffma2_rn(x, y, d[k]); d[k] = __ffma2_rn(d[k], x, y); d[k] = __ffma2_rn(x, d[k], y); d[k] = __ffma2_rn(d[k], d[k], y); }
            if (KIND == 3) { d[k] = __ffma2_rn(x, y, d[k]); d[k] = __ffma2_rn(d[k], x, y); d[k] = __ffma2_rn(x, d[k], y); d[k] = __ffma2_rn(d[k], d[k], y); }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int k = 0; k < CH; k++) s += c[k][0] + c[k][1] + c[k][2] + c[k][3] + d[k].x + d[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int KIND>
void run(const char* name, int inst_per_step, int warps_per_sm) {
    float* out; long long* cyc; long long h = 0; float* in;
    const int threads = 32 * warps_per_sm;
    cudaMalloc(&out, sizeof(float) * 148 * threads);
    cudaMalloc(&cyc, sizeof(long long));
    cudaMalloc(&in, sizeof(float) * (512 + threads));
    cudaMemset(in, 0, sizeof(float) * (512 + threads));
    probe<KIND><<<148, threads>>>(in, out, cyc);
    probe<KIND><<<148, threads>>>(in, out, cyc);
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double winst = (double)ITERS * CH * inst_per_step * warps_per_sm;
    printf("%-40s warps/SM %2d : %.3f warp-inst/clk/SM  (%.2f clk per warp-inst per SMSP), loop %.1f clk per step per warp\n", name, warps_per_sm, winst / h,
           h / (winst / 4), (double)h / (ITERS * CH));
    cudaFree(out); cudaFree(cyc); cudaFree(in);
}

int main() {
    for (int w : {8, 16}) {
        run<0>("mma.sync m16n8k4 tf32", 1, w);
        run<1>("mma.sync m16n8k8 tf32", 1, w);
        run<3>("4 FFMA2", 4, w);
        run<2>("mma m16n8k4 + 4 FFMA2 (5 inst)", 5, w);
        printf("\n");
    }
    return 0;
}
