// Dependent-issue latency of DADD / predicated DADD / LDS+DADD on this GPU (one warp, clock64 around a long chain).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, const double *in, int n) {
    __shared__ double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = in[i];
    __syncthreads();
    double acc = 0.0;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < n; i++) acc = acc + 1.0000001;
    long long t1 = clock64();
    double acc2 = 0.0;
#pragma unroll 8
    for (int i = 0; i < n; i++) acc2 = acc2 + sm[(i + threadIdx.x * 33) & 4095];
    long long t2 = clock64();
    double acc3 = 0.0;
    int len = n - threadIdx.x * 3;
#pragma unroll 8
    for (int i = 0; i < n; i++) if (i < len) acc3 = acc3 + sm[(i + threadIdx.x * 33) & 4095];
    long long t3 = clock64();
    float f = 0.f;
#pragma unroll 8
    for (int i = 0; i < n; i++) f = f + 1.0000001f;
    long long t4 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
    out[threadIdx.x] = acc + acc2 + acc3 + f;
}
int main() {
    double *out, *in; long long *cyc;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&in, 4096 * 8); cudaMalloc(&cyc, 32);
    cudaMemset(in, 0, 4096 * 8);
    int n = 20000;
    for (int rep = 0; rep < 2; rep++) k<<<1, 32>>>(out, cyc, in, n);
    long long h[4]; cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
    printf("cycles per step: DADD %.2f  LDS+DADD %.2f  predicated LDS+DADD %.2f  FADD %.2f\n", h[0] / (double)n, h[1] / (double)n, h[2] / (double)n, h[3] / (double)n);
    return 0;
}
