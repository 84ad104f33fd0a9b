// dependent-issue latency of DFMA / FADD / IMAD chains and of the K1S-style chain step (2 LDS.64 + DFMA), one warp
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double *out, long long *cyc, double a, double b, int n) {
    double acc = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) acc = __fma_rn(a, b, acc), b += 0.0;
    long long t1 = clock64();
    out[threadIdx.x] = acc; cyc[0] = t1 - t0;
}
__global__ void k_dfma_only(double *out, long long *cyc, double a, double b, int n) {
    double acc = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) acc = __fma_rn(acc, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = acc; cyc[0] = t1 - t0;
}
__global__ void k_fadd(float *out, long long *cyc, float a, int n) {
    float acc = out[threadIdx.x];
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) acc = __fadd_rn(acc, a);
    long long t1 = clock64();
    out[threadIdx.x] = acc; cyc[0] = t1 - t0;
}
template <int U>
__global__ void k_lds_dfma(double *out, long long *cyc, int n) {
    __shared__ double ys[4096 + 64];
    for (int i = threadIdx.x; i < 4096 + 64; i += 32) ys[i] = 1.0 + i * 1e-9;
    __syncwarp();
    double acc = 0.0;
    const double *pa = ys + 32 - threadIdx.x, *pb = ys + 32;
    long long t0 = clock64();
#pragma unroll U
    for (int t = 0; t < n; t++) acc = __fma_rn(pa[t], pb[t], acc);
    long long t1 = clock64();
    out[threadIdx.x] = acc; cyc[0] = t1 - t0;
}
// explicit two-stage pipeline: loads of the next 8 steps in flight while the 8 DFMAs of this block run
__global__ void k_lds_dfma_pipe(double *out, long long *cyc, int n) {
    __shared__ double ys[4096 + 64];
    for (int i = threadIdx.x; i < 4096 + 64; i += 32) ys[i] = 1.0 + i * 1e-9;
    __syncwarp();
    double acc = 0.0;
    const double *pa = ys + 32 - threadIdx.x, *pb = ys + 32;
    long long t0 = clock64();
    double a[8], b[8], c[8], d[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = pa[i]; b[i] = pb[i]; }
    int t = 0;
    for (; t + 16 <= n; t += 16) {
#pragma unroll
        for (int i = 0; i < 8; i++) { c[i] = pa[t + 8 + i]; d[i] = pb[t + 8 + i]; }
#pragma unroll
        for (int i = 0; i < 8; i++) acc = __fma_rn(a[i], b[i], acc);
#pragma unroll
        for (int i = 0; i < 8; i++) { a[i] = pa[t + 16 + i]; b[i] = pb[t + 16 + i]; }
#pragma unroll
        for (int i = 0; i < 8; i++) acc = __fma_rn(c[i], d[i], acc);
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc; cyc[0] = t1 - t0;
}
int main() {
    double *out; long long *cyc; float *fo;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8); cudaMalloc(&fo, 32 * 4);
    cudaMemset(out, 0, 256); cudaMemset(fo, 0, 128);
    long long h;
    const int n = 4096;
    for (int rep = 0; rep < 2; rep++) {
        k_dfma<<<1, 32>>>(out, cyc, 1.0000001, 0.5, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("DFMA chain through addend (+DADD on the side): %.2f cycles/step\n", (double)h / n);
        k_dfma_only<<<1, 32>>>(out, cyc, 1.0000001, 0.5, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("DFMA dependent chain: %.2f cycles/step\n", (double)h / n);
        k_fadd<<<1, 32>>>(fo, cyc, 1.5f, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("FADD dependent chain: %.2f cycles/step\n", (double)h / n);
        k_lds_dfma<4><<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("2 LDS.64 + DFMA, unroll 4: %.2f cycles/step\n", (double)h / n);
        k_lds_dfma<8><<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("2 LDS.64 + DFMA, unroll 8: %.2f cycles/step\n", (double)h / n);
        k_lds_dfma<16><<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("2 LDS.64 + DFMA, unroll 16: %.2f cycles/step\n", (double)h / n);
        k_lds_dfma<32><<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("2 LDS.64 + DFMA, unroll 32: %.2f cycles/step\n", (double)h / n);
        k_lds_dfma_pipe<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        if (rep) printf("2 LDS.64 + DFMA, explicit two-stage pipeline of 8: %.2f cycles/step\n", (double)h / n);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
