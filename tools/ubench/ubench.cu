// Instruction-throughput microbenchmark (sm_100a): warp-instructions per cycle per SM for the integer / conversion /
// FP64 operations the analysis kernels are built from.  Every thread runs 8 independent dependency chains of the
// operation, 1024 threads per SM (32 warps) on every SM; the figure is instructions / elapsed SM cycles.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu ; run: ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(1024) k(int32_t *out, int32_t seed, long long *cycles) {
    int32_t a[CHAINS];
    double d[CHAINS];
    float f[CHAINS];
    for (int i = 0; i < CHAINS; i++) { a[i] = seed + threadIdx.x * 7 + i; d[i] = (double)a[i] * 1e-9; f[i] = (float)a[i] * 1e-6f; }
    const int32_t m = seed | 0x0101, q = seed * 3 + 1;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (OP == 0) a[i] = __sad(a[i], q, (unsigned)a[i]);                 // VABSDIFF
            if (OP == 1) a[i] = abs(a[i]) + q;                                   // IABS + IADD (2 instr)
            if (OP == 2) a[i] = __dp2a_lo(a[i], m, a[i]);                        // IDP.2A
            if (OP == 3) a[i] = a[i] * q + a[i];                                 // IMAD
            if (OP == 4) a[i] = (a[i] << 1) ^ (a[i] >> 31);                      // zigzag (3 instr)
            if (OP == 5) f[i] = (float)__float_as_int(f[i]) ;                    // I2F
            if (OP == 6) d[i] = __fma_rn(d[i], 1.0000001, 1e-9);                 // DFMA
            if (OP == 7) a[i] = (a[i] & m) | (a[i] ^ q);                         // LOP3
            if (OP == 8) a[i] = min(a[i], q) + 1;                                // IMNMX + IADD (2 instr)
            if (OP == 9) d[i] = (double)__double2float_rn(d[i]) + 1e-9;          // F2F.F32.F64 + F2F.F64.F32 + DADD (3)
            if (OP == 10) a[i] = __funnelshift_l(a[i], q, a[i] & 31);            // SHF (+LOP)
            if (OP == 11) f[i] = __fadd_rn(f[i], 1.5f);                          // FADD
            if (OP == 12) a[i] = __popc(a[i]) + q;                               // POPC + IADD
            if (OP == 13) a[i] = a[i] + q;                                        // IADD
            if (OP == 14) a[i] = __shfl_xor_sync(0xFFFFFFFFu, a[i], 1);          // SHFL
        }
    }
    const long long t1 = clock64();
    int32_t r = 0;
    for (int i = 0; i < CHAINS; i++) r += a[i] + (int32_t)d[i] + (int32_t)f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int instr_per_op) {
    int32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    k<OP><<<148, 1024>>>(out, 12345, cyc);
    k<OP><<<148, 1024>>>(out, 12345, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += (double)h[i]; avg /= 148;
    const double ops = (double)ITER * CHAINS * 32;  // warp-level ops per SM (32 warps)
    printf("%-34s %7.3f warp-ops/cycle/SM  (%5.3f per SMSP; SASS instr/op %d -> %5.3f instr/cycle/SMSP)\n", name, ops / avg, ops / avg / 4,
           instr_per_op, ops / avg / 4 * instr_per_op);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("VABSDIFF (__sad)", 1);
    run<1>("IABS + IADD", 2);
    run<2>("IDP.2A (dp2a)", 1);
    run<3>("IMAD", 1);
    run<4>("zigzag SHL,SHR,XOR", 3);
    run<5>("I2F", 1);
    run<6>("DFMA", 1);
    run<7>("LOP3", 1);
    run<8>("IMNMX + IADD", 2);
    run<9>("F2F64->32, F2F32->64, DADD", 3);
    run<10>("SHF (+LOP)", 2);
    run<11>("FADD", 1);
    run<12>("POPC + IADD", 2);
    run<13>("IADD", 1);
    run<14>("SHFL", 1);
    return 0;
}
