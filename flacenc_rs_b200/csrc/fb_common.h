// fb_common.h -- shared definitions of the B200 FLAC frame-encode pipeline.
//
// This header is compiled two ways:
//   * by nvcc for sm_100a (the product: libflacenc_b200.so), and
//   * by g++ with FB_EMULATE defined (tests/emu): the same kernel bodies are run on the CPU
//     by executing each barrier-delimited phase for tid = 0..T-1 in turn.  That build exists
//     only so kernel logic can be checked against the oracle on machines without a GPU; it is
//     never linked into the product library and is not a fallback.
//
// Reference citations are relative to /root/reference/.
#pragma once

#include <stdint.h>
#include <string.h>

#include "../../include/flacenc_b200.h"

#if defined(__CUDACC__) && !defined(FB_EMULATE)
#define FB_GPU 1
#define FB_HD __host__ __device__ __forceinline__
#define FB_DEV __device__ __forceinline__
#else
#define FB_GPU 0
#define FB_HD inline
#define FB_DEV inline
#include <math.h>
struct int4 { int32_t x, y, z, w; };
struct int2 { int32_t x, y; };
static inline int32_t min(int32_t a, int32_t b) { return a < b ? a : b; }
static inline int32_t max(int32_t a, int32_t b) { return a > b ? a : b; }
struct float4 { float x, y, z, w; };
#endif

// ---- phase macros: a "phase" is a region between two CTA barriers --------------------------
// Rule for code between phases (executed by every thread on the GPU, once under emulation): it may
// read shared memory for uniform control flow, but a value read there must not be overwritten by
// the next phase unless an FB_SYNC() separates the read from that phase (a fast thread could
// otherwise change the value before a slow thread has read it and the CTA would diverge).
#if FB_GPU
#define FB_PHASE(tid, T) { const int tid = (int)threadIdx.x; (void)tid;
#define FB_PHASE_END } __syncthreads();
#define FB_SYNC() __syncthreads()
#else
#define FB_PHASE(tid, T) for (int tid = 0; tid < (T); ++tid) {
#define FB_PHASE_END }
#define FB_SYNC() ((void)0)
#endif

// ---- warp-scope phase macros ---------------------------------------------------------------------
// FB_WARPS_BEGIN(w, NW) ... FB_WARPS_END : every warp of the CTA runs the enclosed code independently
// (the emulation runs the warps one after the other); inside, FB_WPHASE(lane) ... FB_WPHASE_END is a
// region between two warp barriers.  Values that steer warp-uniform control flow are read from shared
// memory between phases, under the same rule as the CTA-scope macros.
#if FB_GPU
#define FB_WARPS_BEGIN(w, NW) { const int w = (int)(threadIdx.x >> 5); (void)w;
#define FB_WARPS_END } __syncthreads();
#define FB_WPHASE(lane) { const int lane = (int)(threadIdx.x & 31u); (void)lane;
#define FB_WPHASE_END } __syncwarp();
#else
#define FB_WARPS_BEGIN(w, NW) for (int w = 0; w < (NW); ++w) {
#define FB_WARPS_END }
#define FB_WPHASE(lane) for (int lane = 0; lane < 32; ++lane) {
#define FB_WPHASE_END }
#endif

#define FB_RICE_SAT ((1u << 27) - 1u)   // src/rice.rs:51
#define FB_MIN_PRED_BLOCK 64            // src/constant.rs:51 MIN_BLOCK_SIZE_FOR_PREDICTION
#define FB_MAX_ENT_PARTS 64             // src/constant.rs:63

// ---- atomics (plain ops under emulation: phases run one thread at a time) -------------------
#if FB_GPU
FB_DEV void fb_atomic_or(uint32_t *p, uint32_t v) { atomicOr(p, v); }
FB_DEV void fb_atomic_max_u32(uint32_t *p, uint32_t v) { atomicMax(p, v); }
FB_DEV void fb_atomic_min_u32(uint32_t *p, uint32_t v) { atomicMin(p, v); }
FB_DEV void fb_atomic_add_u64(unsigned long long *p, unsigned long long v) { atomicAdd(p, v); }
#else
FB_DEV void fb_atomic_or(uint32_t *p, uint32_t v) { *p |= v; }
FB_DEV void fb_atomic_max_u32(uint32_t *p, uint32_t v) { if (v > *p) *p = v; }
FB_DEV void fb_atomic_min_u32(uint32_t *p, uint32_t v) { if (v < *p) *p = v; }
FB_DEV void fb_atomic_add_u64(unsigned long long *p, unsigned long long v) { *p += v; }
#endif

// ---- EXTENSION beyond the reference (config.ext_lpc_order_search): the quantised coefficient sets of the lower LPC
// orders of a channel variant, written by the analysis kernels and probed by the plan / rice kernels
#define FB_EXT_LPC_MAX 8
struct FbLpcExt {
    int16_t qlp[32];
    int32_t order, shift;   // order after tail-zero truncation (0: no such candidate)
    int32_t precision;      // quantiser precision of this set
};
// the orders tried besides P: P - i * ceil(P / (k + 1)), i = 1..k, while >= 1; returns their number
FB_HD int fb_ext_lpc_orders(int P, int k, int *orders) {
    int n = 0;
    if (k <= 0) return 0;
    const int step = (P + k) / (k + 1);
    for (int i = 1; i <= k && i <= FB_EXT_LPC_MAX; i++) {
        const int o = P - i * step;
        if (o < 1) break;
        orders[n++] = o;
    }
    return n;
}

// ---- job description passed by value to every kernel ----------------------------------------
struct FbJob {
    fb200_config cfg;
    int32_t channels;
    int32_t bps;             // stream bits per sample
    int32_t sample_rate;
    int32_t block_size;
    int32_t nvar;            // channel variants per frame: 4 (L,R,M,S) for stereo, else channels
    int32_t stride;          // planar stride in samples (block_size rounded up to 32)
    int32_t container_bytes; // 2, 3 (packed LE) or 4 (int32)
    int32_t tail_n;          // size of the last frame of the batch (== block_size when full)
    uint32_t n_frames;
    uint32_t first_frame_number;
    uint64_t n_samples;      // per channel
    uint32_t slot_bytes;     // stride of the per-frame output slots (multiple of 16)
    int32_t  pack_in_smem;   // 1: frames are assembled in shared memory, 0: in their global slot
    FbLpcExt *lpc_ext;       // [n_frames * nvar][FB_EXT_LPC_MAX] or nullptr (extension off)
};

// Output of the analysis kernel (K1), one per channel variant.
struct FbAnalysis {
    int32_t  is_constant;
    int32_t  fixed_order;        // ApproxEnt winner, -1 = None
    int32_t  qlp_order;
    int32_t  qlp_shift;
    uint32_t max_abs;            // max |x| of the variant (selects the i32 / i64 residual accumulation, src/lpc.rs:361-374)
    uint32_t pad;
    uint64_t fixed_est[5];
    int16_t  qlp[32];
};

FB_HD int fb_frame_len(const FbJob &J, uint32_t frame) {
    return (frame + 1 == J.n_frames) ? J.tail_n : J.block_size;
}

// bits per sample of a variant (ChannelAssignment::bits_per_sample_offset, src/component/datatype.rs:1145-1171)
FB_HD int fb_variant_bps(const FbJob &J, int v) { return J.bps + ((J.channels == 2 && v == 3) ? 1 : 0); }

// c + a * b with a 64-bit accumulator: one IMAD.WIDE on the GPU (mad.wide.s32)
FB_DEV int64_t fb_mad_wide(int32_t a, int32_t b, int64_t c) {
#if FB_GPU
    long long r;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"((long long)c));
    return (int64_t)r;
#else
    return c + (int64_t)a * (int64_t)b;
#endif
}

// c + lo16(a) * byte0(b) + hi16(a) * byte1(b), all signed: one IDP.2A on the GPU.  With a = a packed 16-bit stereo
// sample pair (left in the low half) and b = FB_PAIR_L / _R / _M / _S it yields L, R, L + R or L - R.
#define FB_PAIR_L 0x0001
#define FB_PAIR_R 0x0100
#define FB_PAIR_M 0x0101
#define FB_PAIR_S 0xFF01
FB_DEV int32_t fb_dp2a_lo(int32_t a, int32_t b, int32_t c) {
#if FB_GPU
    return __dp2a_lo(a, b, c);
#else
    return c + (int32_t)(int16_t)((uint32_t)a & 0xFFFFu) * (int32_t)(int8_t)((uint32_t)b & 0xFFu) +
           (int32_t)(int16_t)((uint32_t)a >> 16) * (int32_t)(int8_t)(((uint32_t)b >> 8) & 0xFFu);
#endif
}
// multiplier bytes and shift of a stereo variant (0 L, 1 R, 2 M, 3 S) formed from a packed pair
FB_HD void fb_pair_mix(int v, int32_t *mb, int32_t *sh) {
    *mb = v == 0 ? FB_PAIR_L : (v == 1 ? FB_PAIR_R : (v == 2 ? FB_PAIR_M : FB_PAIR_S));
    *sh = v == 2 ? 1 : 0;
}

// M = (L + R) >> 1 (arithmetic), S = L - R (src/coding.rs:476-484); unsigned adds so that stale padding
// words (whose results are masked) cannot overflow a signed int
FB_HD int32_t fb_mid(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b) >> 1; }
FB_HD int32_t fb_side(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }

// src/rice.rs:169-171 encode_signbit: (|v| << 1) - (v < 0)
FB_HD uint32_t fb_zigzag(int32_t v) { return ((uint32_t)v << 1) ^ (uint32_t)(v >> 31); }

// src/rice.rs:157-165 finest_partition_order(size, max(64, warmup)); warmup <= 24 so min part is 64
FB_HD int fb_finest_partition_order(int n) {
    uint32_t max_splits = (uint32_t)(n / 64);
    if (max_splits == 0) return 0;
#if defined(__CUDA_ARCH__)
    const int lg = 31 - __clz((int)max_splits);
    const int tz = __ffs(n) - 1; // n >= 64 here
#else
    int lg = 0;
    while ((max_splits >> (lg + 1)) != 0) lg++;
    int tz = 0;
    while (((n >> tz) & 1) == 0 && tz < 31) tz++;
#endif
    int r = lg < tz ? lg : tz;
    return r < 15 ? r : 15;
}

// ---- glibc-compatible log2f ------------------------------------------------------------------
// The reference's estimate_entropy (src/coding.rs:200-227) calls f32::log2, i.e. the platform libm
// (glibc on Linux).  CUDA's log2f rounds differently in the last place, which can flip an order
// decision, so the device evaluates the published table-driven algorithm glibc >= 2.27 uses
// (ARM optimized-routines log2f: 16-entry table, degree-4 polynomial in double).
// tests/test_log2f_compat.py checks this function against the host libm for all 2^31 positive floats.
FB_HD uint32_t fb_f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
FB_HD float fb_u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

#if FB_GPU
__device__ __constant__ double FB_LOG2F_TAB[32] = {
#else
static const double FB_LOG2F_TAB[32] = {
#endif
    0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2, 0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2,
    0x1.49539f0f010bp+0,  -0x1.7418b0a1fb77bp-2, 0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2,
    0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2, 0x1.25e227b0b8eap+0,  -0x1.97c1d1b3b7afp-3,
    0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3, 0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4,
    0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5, 0x1p+0,               0x0p+0,
    0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4,  0x1.ca4b31f026aap-1,  0x1.476a9543891bap-3,
    0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3,  0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2,
    0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2,  0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2};

#if FB_GPU
#define FB_FMA(a, b, c) __fma_rn((a), (b), (c))
#define FB_FMAF(a, b, c) __fmaf_rn((a), (b), (c))
#define FB_DMUL(a, b) __dmul_rn((a), (b))
#define FB_DADD(a, b) __dadd_rn((a), (b))
#define FB_FMUL(a, b) __fmul_rn((a), (b))
#define FB_FADD(a, b) __fadd_rn((a), (b))
#define FB_FDIV(a, b) __fdiv_rn((a), (b))
#define FB_DDIV(a, b) __ddiv_rn((a), (b))
#else
#define FB_FMA(a, b, c) fma((a), (b), (c))
#define FB_FMAF(a, b, c) fmaf((a), (b), (c))
#define FB_DMUL(a, b) ((a) * (b))
#define FB_DADD(a, b) ((a) + (b))
#define FB_FMUL(a, b) ((a) * (b))
#define FB_FADD(a, b) ((a) + (b))
#define FB_FDIV(a, b) ((a) / (b))
#define FB_DDIV(a, b) ((a) / (b))
#endif

FB_DEV float fb_log2f(float x) {
    uint32_t ix = fb_f2u(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0) return fb_u2f(0xff800000u);           // -inf
        if (ix == 0x7f800000u) return x;                        // +inf
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return fb_u2f(0x7fc00000u); // NaN
        ix = fb_f2u(FB_FMUL(x, 8388608.0f));                    // subnormal: scale by 2^23
        ix -= 23u << 23;
    }
    uint32_t tmp = ix - 0x3f330000u;
    int i = (int)((tmp >> 19) & 15u);
    uint32_t top = tmp & 0xff800000u;
    uint32_t iz = ix - top;
    int k = (int32_t)tmp >> 23;
    double invc = FB_LOG2F_TAB[2 * i], logc = FB_LOG2F_TAB[2 * i + 1];
    double z = (double)fb_u2f(iz);
    double r = FB_DADD(FB_DMUL(z, invc), -1.0);
    double y0 = FB_DADD(logc, (double)k);
    double r2 = FB_DMUL(r, r);
    double y = FB_DADD(FB_DMUL(0x1.ecabf496832ep-2, r), -0x1.715479ffae3dep-1);
    y = FB_DADD(FB_DMUL(-0x1.712b6f70a7e4dp-2, r2), y);
    double p = FB_DADD(FB_DMUL(0x1.715475f35c8b8p0, r), y0);
    y = FB_DADD(FB_DMUL(y, r2), p);
    return (float)y;
}

// ---- glibc-compatible powf for positive x ----------------------------------------------------------
// The IRLS refinement of the `experimental` estimator (src/lpc.rs:814-850) weights every sample by
// (max(|err|, 1) / normalizer).max(0.01).powf(-1.2): f32::powf is the platform libm (glibc: the table-driven algorithm of
// ARM's optimized routines -- log2 in double from the 16-entry table above with its own degree-5 polynomial, times y,
// exp2 from a 32-entry table).  Only what that call can meet is implemented: x a positive normal number or +inf, y
// finite and non-zero.  FB_POWF_FMA selects the expression forms glibc's FMA build contracts (the variant it dispatches
// to on every x86-64 CPU with FMA); tests/test_kernel_logic_emu.py sweeps every x >= 0.01 for y = -1.2 against the host.
#if FB_GPU
__device__ __constant__ unsigned long long FB_EXP2F_TAB[32] = {
#else
static const unsigned long long FB_EXP2F_TAB[32] = {
#endif
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, 0x3fef72b83c7d517bull,
    0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, 0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull,
    0x3feedea64c123422ull, 0x3feece086061892dull, 0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull,
    0x3feea47eb03a5585ull, 0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, 0x3feee89f995ad3adull,
    0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, 0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full,
    0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};

#ifndef FB_POWF_FMA
#define FB_POWF_FMA 1
#endif
// a * b + c the way the libm build evaluates it
#if FB_POWF_FMA
#define FB_PW_MAD(a, b, c) FB_FMA((a), (b), (c))
#else
#define FB_PW_MAD(a, b, c) FB_DADD(FB_DMUL((a), (b)), (c))
#endif

FB_DEV float fb_powf_pos(float x, float y) {
    uint32_t ix = fb_f2u(x);
    if (ix == 0x7f800000u) return (fb_f2u(y) & 0x80000000u) ? 0.0f : x; // +inf
    // log2_inline
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const uint32_t top = tmp & 0xff800000u;
    const uint32_t iz = ix - top;
    const int k = (int32_t)top >> 23;
    const double invc = FB_LOG2F_TAB[2 * i], logc = FB_LOG2F_TAB[2 * i + 1];
    const double z = (double)fb_u2f(iz);
    const double r = FB_PW_MAD(z, invc, -1.0);
    const double y0 = FB_DADD(logc, (double)k);
    const double r2 = FB_DMUL(r, r);
    double yy = FB_PW_MAD(0x1.27616c9496e0bp-2, r, -0x1.71969a075c67ap-2);
    const double p = FB_PW_MAD(0x1.ec70a6ca7baddp-2, r, -0x1.7154748bef6c8p-1);
    const double r4 = FB_DMUL(r2, r2);
    double q = FB_PW_MAD(0x1.71547652ab82bp+0, r, y0);
    q = FB_PW_MAD(p, r2, q);
    yy = FB_PW_MAD(yy, r4, q);
    const double ylogx = FB_DMUL((double)y, yy);
    // |y * log2(x)| >= 126: overflow / underflow thresholds of the float result
    uint64_t yb;
    memcpy(&yb, &ylogx, 8);
    if (((yb >> 47) & 0xffffu) >= (0x405f800000000000ull >> 47)) {
        if (ylogx > 0x1.fffffffd1d571p+6) return fb_u2f(0x7f800000u);
        if (ylogx <= -150.0) return 0.0f;
    }
    // exp2_inline
    const double shift = 0x1.8p+47; // 0x1.8p52 / 32
    double kd = FB_DADD(ylogx, shift);
    uint64_t ki;
    memcpy(&ki, &kd, 8);
    kd = FB_DADD(kd, -shift);
    const double rr = FB_DADD(ylogx, -kd);
    uint64_t t = FB_EXP2F_TAB[ki & 31u];
    t += ki << 47; // (52 - 5)
    double sc;
    memcpy(&sc, &t, 8);
    const double zz = FB_PW_MAD(0x1.c6af84b912394p-5, rr, 0x1.ebfce50fac4f3p-3);
    const double rr2 = FB_DMUL(rr, rr);
    double e = FB_PW_MAD(0x1.62e42ff0c52d6p-1, rr, 1.0);
    e = FB_PW_MAD(zz, rr2, e);
    e = FB_DMUL(e, sc);
    return (float)e;
}

// Rust `f32 as usize`: saturating, NaN -> 0 (src/coding.rs:222)
FB_HD uint64_t fb_f32_as_u64(float v) {
    if (!(v == v)) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 18446744073709551616.0f) return 0xFFFFFFFFFFFFFFFFull;
    return (uint64_t)v;
}

// One partition of estimate_entropy (src/coding.rs:216-222); `sum` is the sequential f32 sum of |e|.
FB_DEV uint64_t fb_entropy_partition_bits(float sum, int sample_count) {
    float cnt = (float)sample_count;
    float avg = FB_FDIV(FB_FMUL(sum, 2.0f), FB_FADD(cnt, 0.00001f));
    float geom_p = FB_FDIV(1.0f, FB_FADD(avg, 1.0f));
    float xent = FB_FMAF(avg, -fb_log2f(FB_FADD(1.0f, -geom_p)), -fb_log2f(geom_p));
    return fb_f32_as_u64(FB_FMUL(xent, cnt));
}

// ---- frame header (src/component/bitrepr.rs:359-420, src/component/datatype.rs:1239-1249,
//      1350-1360, 1427-1453) ---------------------------------------------------------------
FB_HD int fb_block_size_tag(int size, int *extra_bits, uint32_t *extra) {
    *extra_bits = 0; *extra = 0;
    if (size == 192) return 1;
    if (size == 576) return 2;
    if (size == 1152) return 3;
    if (size == 2304) return 4;
    if (size == 4608) return 5;
    for (int k = 0; k <= 7; k++) if (size == (256 << k)) return 8 + k;
    if (size <= 256) { *extra_bits = 8; *extra = (uint32_t)(size - 1); return 6; }
    *extra_bits = 16; *extra = (uint32_t)(size - 1); return 7;
}

FB_HD int fb_sample_size_tag(int bits) {
    switch (bits) {
    case 8: return 1; case 12: return 2; case 16: return 4; case 20: return 5; case 24: return 6; case 32: return 7;
    default: return 0;
    }
}

FB_HD int fb_sample_rate_tag(uint32_t f, int *extra_bits, uint32_t *extra) {
    *extra_bits = 0; *extra = 0;
    switch (f) {
    case 88200: return 1; case 176400: return 2; case 192000: return 3; case 8000: return 4;
    case 16000: return 5; case 22050: return 6; case 24000: return 7; case 32000: return 8;
    case 44100: return 9; case 48000: return 10; case 96000: return 11; default: break;
    }
    if (f % 1000 == 0 && f / 1000 <= 255) { *extra_bits = 8; *extra = f / 1000; return 12; }
    if (f % 10 == 0 && f / 10 <= 65535) { *extra_bits = 16; *extra = f / 10; return 14; }
    if (f <= 65535) { *extra_bits = 16; *extra = f; return 13; }
    return 0;
}

// CRC-8/SMBUS: poly 0x07, init 0 (src/component/bitrepr.rs:39,356-357).  One byte step v -> v * x^8 mod P in closed
// form: the low six bits multiply by x^2 + x + 1 without overflowing, bits 6 and 7 contribute 0xC7 and 0x89
// (tests/test_kernel_logic_emu.py checks it against the bit-serial definition).
FB_HD uint32_t fb_crc8_step(uint32_t v) {
    const uint32_t l = v & 63u;
    return ((l ^ (l << 1) ^ (l << 2)) ^ ((v & 64u) ? 0xC7u : 0u) ^ ((v & 128u) ? 0x89u : 0u)) & 0xFFu;
}
FB_HD uint8_t fb_crc8(const uint8_t *d, int len) {
    uint32_t crc = 0;
    for (int i = 0; i < len; i++) crc = fb_crc8_step(crc ^ d[i]);
    return (uint8_t)crc;
}

// Fixed-blocking frame header incl. CRC-8; returns the number of bytes (<= 16).
FB_HD int fb_frame_header(int n, int ch_tag, int bps, int sample_rate, uint32_t frame_number, uint8_t *out,
                           const uint32_t *crc8_tab = nullptr) {
    int bs_bits, sr_bits;
    uint32_t bs_extra, sr_extra;
    int bs_tag = fb_block_size_tag(n, &bs_bits, &bs_extra);
    int sr_tag = fb_sample_rate_tag((uint32_t)sample_rate, &sr_bits, &sr_extra);
    int k = 0;
    out[k++] = 0xFF;
    out[k++] = 0xF8; // fixed blocking (FrameOffset::Frame, src/coding.rs:603-604)
    out[k++] = (uint8_t)((bs_tag << 4) | sr_tag);
    out[k++] = (uint8_t)((ch_tag << 4) | (fb_sample_size_tag(bps) << 1));
    // encode_to_utf8like(frame_number) (src/component/bitrepr.rs:121-159); frame_number < 2^31
    uint32_t v = frame_number;
    int code_bits = 0;
    while (code_bits < 32 && (v >> code_bits) != 0) code_bits++;
    if (code_bits <= 7) {
        out[k++] = (uint8_t)v;
    } else {
        int trailing = (code_bits - 2) / 5;
        const uint8_t heads[7] = {0x80, 0xC0, 0xE0, 0xF0, 0xF8, 0xFC, 0xFE};
        out[k++] = (trailing == 6) ? 0xFE : (uint8_t)(heads[trailing] | (uint8_t)(v >> (6 * trailing)));
        for (int i = trailing - 1; i >= 0; i--) out[k++] = (uint8_t)(0x80 | ((v >> (6 * i)) & 0x3F));
    }
    if (bs_bits == 8) out[k++] = (uint8_t)bs_extra;
    if (bs_bits == 16) { out[k++] = (uint8_t)(bs_extra >> 8); out[k++] = (uint8_t)bs_extra; }
    if (sr_bits == 8) out[k++] = (uint8_t)sr_extra;
    if (sr_bits == 16) { out[k++] = (uint8_t)(sr_extra >> 8); out[k++] = (uint8_t)sr_extra; }
    if (crc8_tab) {
        // table form of the same CRC-8 (256 entries built on the host)
        uint32_t crc = 0;
        for (int i = 0; i < k; i++) crc = crc8_tab[(crc ^ out[i]) & 0xFFu];
        out[k] = (uint8_t)crc;
    } else {
        out[k] = fb_crc8(out, k);
    }
    return k + 1;
}

// Upper bound of a frame in bytes: header (16) + per channel a verbatim subframe at bps+1 + CRC-16.
FB_HD uint32_t fb_max_frame_bytes(int channels, int bps, int block_size) {
    uint64_t bits = 16 * 8 + (uint64_t)channels * (8 + (uint64_t)block_size * (uint64_t)(bps + 1)) + 7 + 16;
    return (uint32_t)(bits >> 3) + 8;
}

// GF(2) helpers for the CRC-16 (poly 0x8005, init 0; src/component/bitrepr.rs:40,270-271)
FB_HD uint32_t fb_crc16_mulmod(uint32_t a, uint32_t b) {
    uint32_t r = 0;
    for (int i = 15; i >= 0; i--) {
        r = (r & 0x8000u) ? (((r << 1) ^ 0x8005u) & 0xFFFFu) : ((r << 1) & 0xFFFFu);
        if ((b >> i) & 1u) r ^= a;
    }
    return r;
}

FB_HD uint32_t fb_crc16_table_entry(uint32_t i) {
    uint32_t c = i << 8;
    for (int b = 0; b < 8; b++) c = (c & 0x8000u) ? (((c << 1) ^ 0x8005u) & 0xFFFFu) : ((c << 1) & 0xFFFFu);
    return c;
}
