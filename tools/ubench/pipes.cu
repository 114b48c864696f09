// Micro-benchmark of issue rates on sm_100a: FFMA vs FFMA2 (packed fp32x2), IMAD, IADD3, I2F.S64, MUFU.RCP.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define ILP 8
template <int OP>
__global__ void k(float* out, int n, float a, float b) {
    float x[ILP]; float2 y[ILP]; int z[ILP]; long long w[ILP];
    for (int i = 0; i < ILP; i++) { x[i] = threadIdx.x + i; y[i] = make_float2(x[i], x[i] + 1); z[i] = threadIdx.x * 3 + i; w[i] = (long long)z[i] << 20; }
    const float2 a2 = make_float2(a, b), b2 = make_float2(b, a);
    for (int it = 0; it < n; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) x[i] = fmaf(x[i], a, b);
            if (OP == 1) y[i] = __ffma2_rn(y[i], a2, b2);
            if (OP == 2) z[i] = z[i] * (int)a + (int)b;
            if (OP == 3) z[i] = (z[i] + (int)a) ^ (int)b;
            if (OP == 4) { x[i] = (float)w[i]; w[i] += __float_as_int(x[i]); }
            if (OP == 5) x[i] = __frcp_rn(x[i]) ;
            if (OP == 6) x[i] = __fdividef(1.0f, x[i]);
            if (OP == 7) { x[i] = (float)z[i]; z[i] += __float_as_int(x[i]); }
        }
    }
    float s = 0; for (int i = 0; i < ILP; i++) s += x[i] + y[i].x + y[i].y + z[i] + (float)w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char* name, float* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148 * 4, 512>>>(d, 16, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<OP><<<148 * 4, 512>>>(d, ITERS, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 4 * 512 * (double)ITERS * ILP;   // thread-level ops
    printf("%-10s %.3f ms  %.1f Gop/s thread-level  = %.2f warp-inst/clk/SM @1.965GHz\n", name, ms, ops / ms / 1e6, ops / 32 / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 4 * 512 * 4);
    run<0>("FFMA", d); run<1>("FFMA2", d); run<2>("IMAD", d); run<3>("IADD+LOP", d); run<4>("I2F.S64+", d); run<5>("frcp_rn", d); run<6>("fdividef", d); run<7>("I2F.S32+", d);
    return 0;
}
