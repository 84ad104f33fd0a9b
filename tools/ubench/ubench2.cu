// Loop-shape microbenchmark for the analysis kernel (sm_100a): cycles per sample-step per SMSP of candidate inner loops,
// 16 warps per SM (4 per SMSP, the kernel's occupancy), synthetic samples from registers.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench2 ubench2.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

#define ITER 4096

__device__ __forceinline__ double f2d_int(float f) { // float -> double by integer ops (normal numbers and zero)
    const uint32_t u = __float_as_uint(f);
    const uint32_t a = u & 0x7FFFFFFFu;
    uint32_t hi = (u & 0x80000000u) | ((a >> 3) + 0x38000000u);
    hi = a ? hi : (u & 0x80000000u);
    return __hiloint2double((int)hi, (int)(u << 29));
}

template <int OP>
__global__ void __launch_bounds__(128) k(int32_t *out, int32_t seed, long long *cycles, int iters) {
    int32_t x = seed + threadIdx.x;
    int32_t p0 = 0, p1 = 0, p2 = 0, p3 = 0, mn = 0, mx = 0;
    uint32_t j0 = 0, j1 = 0, j2 = 0, j3 = 0, j4 = 0;
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, fp0 = 0, fp1 = 0, fp2 = 0, fp3 = 0, fmn = 0, fmx = 0;
    double acc[11], ring[12];
    for (int i = 0; i < 11; i++) acc[i] = 0;
    for (int i = 0; i < 12; i++) ring[i] = 0;
    float w = 0.75f;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int s = 0; s < 12; s++) {
            x = x * 1664525 + 1013904223;          // the "sample" (IMAD)
            const int32_t e0 = x >> 16;
            if (OP == 0) { // int pass E: sub chain, min/max, VABSDIFF
                mn = min(mn, e0); mx = max(mx, e0);
                const int32_t e1 = e0 - p0, e2 = e1 - p1, e3 = e2 - p2, e4 = e3 - p3;
                p0 = e0; p1 = e1; p2 = e2; p3 = e3;
                j0 = __sad(e0, 0, j0); j1 = __sad(e1, 0, j1); j2 = __sad(e2, 0, j2); j3 = __sad(e3, 0, j3); j4 = __sad(e4, 0, j4);
            }
            if (OP == 1) { // old float pass E: int sub chain, 5 x (I2F + FADD |.|)
                mn = min(mn, e0); mx = max(mx, e0);
                const int32_t e1 = e0 - p0, e2 = e1 - p1, e3 = e2 - p2, e4 = e3 - p3;
                p0 = e0; p1 = e1; p2 = e2; p3 = e3;
                s0 = __fadd_rn(fabsf((float)e0), s0); s1 = __fadd_rn(fabsf((float)e1), s1); s2 = __fadd_rn(fabsf((float)e2), s2);
                s3 = __fadd_rn(fabsf((float)e3), s3); s4 = __fadd_rn(fabsf((float)e4), s4);
            }
            if (OP == 2) { // float-domain pass E: one I2F, everything else FADD / FMNMX
                const float xf = (float)e0;
                fmn = fminf(fmn, xf); fmx = fmaxf(fmx, xf);
                const float e1 = __fadd_rn(xf, -fp0), e2 = __fadd_rn(e1, -fp1), e3 = __fadd_rn(e2, -fp2), e4 = __fadd_rn(e3, -fp3);
                fp0 = xf; fp1 = e1; fp2 = e2; fp3 = e3;
                s0 = __fadd_rn(fabsf(xf), s0); s1 = __fadd_rn(fabsf(e1), s1); s2 = __fadd_rn(fabsf(e2), s2);
                s3 = __fadd_rn(fabsf(e3), s3); s4 = __fadd_rn(fabsf(e4), s4);
            }
            if (OP == 3 || OP == 4 || OP == 5) { // pass A: y = (f32)x * w -> f64, 11 DFMA with a register ring
                const float yf = __fmul_rn((float)e0, w);
                double y;
                if (OP == 3) y = (double)yf;             // F2F.F64.F32
                else if (OP == 4) y = f2d_int(yf);       // integer conversion
                else y = __hiloint2double(__float_as_int(yf), e0); // no conversion at all (lower bound)
                acc[0] = __fma_rn(y, y, acc[0]);
#pragma unroll
                for (int j = 0; j < 10; j++) acc[j + 1] = __fma_rn(ring[(s - 1 - j + 24) % 12], y, acc[j + 1]);
                ring[s] = y;
            }
        }
    }
    const long long t1 = clock64();
    double r = 0;
    for (int i = 0; i < 11; i++) r += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = mn + mx + j0 + j1 + j2 + j3 + j4 + p0 + p1 + p2 + p3 + (int)(s0 + s1 + s2 + s3 + s4 + fmn + fmx + fp0 + fp1 + fp2 + fp3) + (int)r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name) {
    int32_t *out; long long *cyc;
    const int ctas = 148 * 4;
    cudaMalloc(&out, ctas * 128 * 4); cudaMalloc(&cyc, ctas * 8);
    k<OP><<<ctas, 128>>>(out, 12345, cyc, ITER);
    k<OP><<<ctas, 128>>>(out, 12345, cyc, ITER);
    cudaDeviceSynchronize();
    static long long h[148 * 4]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < ctas; i++) avg += (double)h[i]; avg /= ctas;
    // per SMSP: 4 warps, each ITER * 12 sample-steps
    printf("%-52s %6.2f cycles per sample-step per SMSP (4 warps/SMSP)\n", name, avg / (ITER * 12.0 * 4.0));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("pass E integer (sub chain, min/max, 5 VABSDIFF)");
    run<1>("pass E old (sub chain, min/max, 5 I2F + 5 FADD)");
    run<2>("pass E float domain (I2F, 2 FMNMX, 9 FADD)");
    run<3>("pass A (I2F, FMUL, F2F.F64.F32, 11 DFMA)");
    run<4>("pass A (I2F, FMUL, integer f32->f64, 11 DFMA)");
    run<5>("pass A lower bound (no conversion)");
    return 0;
}
