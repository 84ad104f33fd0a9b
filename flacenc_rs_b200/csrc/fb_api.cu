// fb_api.cu -- CUDA kernel wrappers and the C ABI of libflacenc_b200.so (include/flacenc_b200.h).
//
// One context = one stream format on one device.  A call encodes a batch of frames in chunks:
//   H2D(pcm) -> K0 ingest -> K1 analyze -> K2 rice -> K3 pack -> K4 scan+gather -> D2H(bytes, sizes)
// on the context's stream.  There is no CPU path: without a usable device every entry point
// returns FB200_ERR_CUDA.  Reference citations are relative to /root/reference/.
#include <cuda_runtime.h>

#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "fb_host.h"
#include "fb_launch.h"
#include "fb_md5.h"

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fb_k0_ingest(FbJob J, const uint8_t *pcm, int32_t *xt, uint32_t *err_flag,
                                                    uint64_t n_items) {
    const uint64_t idx = (uint64_t)blockIdx.x * 256u + threadIdx.x;
    if (idx >= n_items) return;
    uint32_t f;
    int t4;
    fb_k0_item(idx, J.stride / 4, &f, &t4);
    if (f < J.n_frames && t4 < J.stride / 4) fb_k0_quad_any(J, pcm, xt, err_flag, f, t4);
}

__global__ void __launch_bounds__(256) fb_k0_ingest_planar(FbJob J, const int32_t *src, int src_stride, int32_t *xt,
                                                           uint32_t *err_flag) {
    const int t4 = (int)(blockIdx.x * 256u + threadIdx.x);
    if (t4 < J.stride / 4) fb_k0_planar_quad(J, src, src_stride, xt, err_flag, t4);
}

// K0b: plain rows by variant for the frames the generic kernels run on (list == nullptr: every frame, one CTA each)
__global__ void __launch_bounds__(256) fb_k0b_expand(FbJob J, const int32_t *xc, int32_t *xv4, const uint32_t *list,
                                                     const uint32_t *count) {
    const uint32_t total = list ? *count : J.n_frames;
    for (uint32_t i = blockIdx.x; i < total; i += gridDim.x) {
        const uint32_t f = list ? list[i] : i;
        for (int t4 = (int)threadIdx.x; t4 < J.stride / 4; t4 += 256) fb_k0b_expand4(J, xc, xv4, f, t4);
    }
}

// the same from packed 16-bit stereo PCM (pairs mode: there is no planar store), for the frames of the fallback list
__global__ void __launch_bounds__(256) fb_k0b_expand_pairs(FbJob J, const uint8_t *pcm, int32_t *xv4, const uint32_t *list,
                                                           const uint32_t *count) {
    const uint32_t total = *count;
    for (uint32_t i = blockIdx.x; i < total; i += gridDim.x) {
        const uint32_t f = list[i];
        const int n = fb_frame_len(J, f);
        const int32_t *src = reinterpret_cast<const int32_t *>(pcm) + (size_t)f * (size_t)J.block_size;
        for (int t = (int)threadIdx.x; t < J.stride; t += 256) {
            const int32_t w = t < n ? src[t] : 0;
            for (int v = 0; v < 4; v++) {
                int32_t mb, sh;
                fb_pair_mix(v, &mb, &sh);
                xv4[((size_t)f * 4u + (size_t)v) * (size_t)J.stride + (size_t)t] = fb_dp2a_lo(w, mb, 0) >> sh;
            }
        }
    }
}

// K1S: the analysis of K1 for SMALL launches, one CTA of three warps per channel variant.  K1 gives every variant one
// thread that walks the whole frame (the floats must be summed in the reference's sequential order), which is the right
// mapping when there are tens of thousands of variants and a latency floor of two full walks when there are a few
// hundred (BASELINE config 1: 432 variants).  Here the independent chains are spread over lanes and the two passes run
// side by side:
//   * warp 0, autocorrelation (src/lpc.rs:533-548): lane tau owns lag tau -- one sequential f64 FMA chain each, over
//     tiles of y = (f32)x * w that warps 1 and 2 stage as doubles in four shared-memory buffers (named barriers: the
//     chain never waits for memory after the first tile); lane 0 finishes with K1's fb_k1_finish_lpc
//   * warps 1 and 2, estimate_entropy (src/coding.rs:200-227): every second partition each; the lanes take the samples
//     of a partition 32 at a time, the zero-history differences in closed form (wrapping, like the chain of
//     subtractions), |e_k| summed as INTEGERS -- every f32 addition of the reference is exact while the sum stays below
//     2^24, so the order of addition does not matter then; partitions where some order reaches 2^24 (loud 24-bit material)
//     are replayed afterwards one per lane with the sequential f32 additions.  Warp 1 finishes with fb_k1_finish_ent.
// Results are bit-identical to K1's.
#define FB_K1S_TILE 1024
#define FB_K1S_NB 4
#define FB_K1S_THREADS 96
// sequential f32 replay of one estimate partition [t0, t1): the five sums of |e_k|
static __device__ __noinline__ void fb_k1s_replay(const FbJob &J, const int32_t *xt, const uint8_t *pcm, const FbVarRows &rows,
                                                  int32_t mb, int32_t sh, uint32_t f, int t0, int t1, float *s) {
    int32_t h[4];
    for (int i = 0; i < 4; i++) h[i] = t0 - 1 - i >= 0 ? fb_k1c_sample(J, xt, pcm, rows, mb, sh, f, t0 - 1 - i) : 0;
    int32_t pe0 = h[0];
    int32_t pe1 = (int32_t)((uint32_t)h[0] - (uint32_t)h[1]);
    int32_t pe2 = (int32_t)((uint32_t)h[0] - 2u * (uint32_t)h[1] + (uint32_t)h[2]);
    int32_t pe3 = (int32_t)((uint32_t)h[0] - 3u * (uint32_t)h[1] + 3u * (uint32_t)h[2] - (uint32_t)h[3]);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
    for (int t = t0; t < t1; t++) {
        const int32_t e0 = fb_k1c_sample(J, xt, pcm, rows, mb, sh, f, t);
        const int32_t e1 = (int32_t)((uint32_t)e0 - (uint32_t)pe0);
        const int32_t e2 = (int32_t)((uint32_t)e1 - (uint32_t)pe1);
        const int32_t e3 = (int32_t)((uint32_t)e2 - (uint32_t)pe2);
        const int32_t e4 = (int32_t)((uint32_t)e3 - (uint32_t)pe3);
        pe0 = e0; pe1 = e1; pe2 = e2; pe3 = e3;
        s0 = FB_FADD(fabsf((float)e0), s0);
        s1 = FB_FADD(fabsf((float)e1), s1);
        s2 = FB_FADD(fabsf((float)e2), s2);
        s3 = FB_FADD(fabsf((float)e3), s3);
        s4 = FB_FADD(fabsf((float)e4), s4);
    }
    s[0] = s0; s[1] = s1; s[2] = s2; s[3] = s3; s[4] = s4;
}

FB_DEV uint32_t fb_k1s_abs24(uint32_t e) { // min(|e|, 2^24) of a wrapped difference
    const uint32_t a = (int32_t)e < 0 ? 0u - e : e;
    return a < (1u << 24) ? a : (1u << 24);
}

__global__ void __launch_bounds__(FB_K1S_THREADS) fb_k1s_analyze(FbJob J, const int32_t *xt, const uint8_t *pcm,
                                                                 const float *win_full, const float *win_tail, FbAnalysis *ana,
                                                                 fb200_variant_taps *taps_all, uint32_t n_variants) {
    extern __shared__ __align__(16) uint8_t fb_smem_k1s[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t gv = blockIdx.x;
    const uint32_t f = gv / (uint32_t)J.nvar;
    const int v = (int)(gv - f * (uint32_t)J.nvar);
    const FbK1Var V = fb_k1_var(J, f, v, win_full, win_tail);
    const int n = V.n, P = V.P;
    const FbVarRows rows = fb_variant_rows(J, xt, f, v);
    int32_t mb = 0, sh = 0;
    if (pcm) fb_pair_mix(v, &mb, &sh);
    FbAnalysis *out = ana + gv;
    fb200_variant_taps *taps = taps_all ? taps_all + gv : nullptr;
    // FB_K1S_NB tile buffers of y (each with the FB200_MAX_LPC_ORDER samples before the tile in front), filled by warps 1
    // and 2 and walked by warp 0; named barriers 2 + b ("buffer b is full") and 2 + NB + b ("buffer b is free"), 96
    // threads each.  Frames of up to NB tiles (4096 samples) are staged without ever waiting for the chain.
    constexpr int YB = FB200_MAX_LPC_ORDER + FB_K1S_TILE, NB = FB_K1S_NB;
    double *ys = (double *)fb_smem_k1s;                                   // [NB][YB] (+ slack: the chain reads ahead)
    unsigned long long *xb = (unsigned long long *)(ys + NB * YB + 32);   // warp 2 -> warp 1: b0..b4
    int32_t *xm = (int32_t *)(xb + 5);                                    // ... and min, max
    const bool do_a = V.do_lpc && !J.cfg.use_direct_mse;
    const int n_tiles = do_a ? (n + FB_K1S_TILE - 1) / FB_K1S_TILE : 0;

    if (warp == 0) {
        // ---- pass A: lane tau accumulates lag tau
        double acc = 0.0;
        for (int k = 0; k < n_tiles; k++) {
            const int b = k % NB, tile0 = k * FB_K1S_TILE;
            const int tile1 = tile0 + FB_K1S_TILE < n ? tile0 + FB_K1S_TILE : n;
            asm volatile("bar.sync %0, 96;" ::"r"(2 + b) : "memory");
            if ((int)lane <= P) {
                const int lo = P > tile0 ? P : tile0;
                const double *pb = ys + b * YB + FB200_MAX_LPC_ORDER - tile0, *pa = pb - (int)lane;
                // (the loads of 16 steps are in flight ahead of the dependent DFMAs: 9.7 instead of 16.8 cycles per step)
#pragma unroll 16
                for (int t = lo; t < tile1; t++) acc = FB_FMA(pa[t], pb[t], acc);
            }
            if (k + NB < n_tiles) asm volatile("bar.arrive %0, 96;" ::"r"(2 + NB + b) : "memory");
        }
        FbK1Acc<FB200_MAX_LPC_ORDER> A;
#pragma unroll
        for (int i = 0; i <= FB200_MAX_LPC_ORDER; i++) A.acc[i] = __shfl_sync(0xFFFFFFFFu, acc, i);
        if (lane == 0) fb_k1_finish_lpc<FB200_MAX_LPC_ORDER>(J, V, A, out, taps, gv);
        return;
    }

    // ---- warps 1 and 2: stage tile k of y = (f32)x * w as doubles (src/lpc.rs:739-756), eight samples per thread in flight
    auto stage_tile = [&](int k) {
        const int b = k % NB, tile0 = k * FB_K1S_TILE;
        const int tile1 = tile0 + FB_K1S_TILE < n ? tile0 + FB_K1S_TILE : n;
        if (k >= NB) asm volatile("bar.sync %0, 96;" ::"r"(2 + NB + b) : "memory");
        double *yb = ys + b * YB + FB200_MAX_LPC_ORDER - tile0;
        const int lo = tile0 >= FB200_MAX_LPC_ORDER ? tile0 - FB200_MAX_LPC_ORDER : 0;
        for (int t = lo + (int)threadIdx.x - 32; t < tile1; t += 512) {
            int32_t x[8];
            float wv[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int ti = t + 64 * i;
                x[i] = ti < tile1 ? fb_k1c_sample(J, xt, pcm, rows, mb, sh, f, ti) : 0;
                wv[i] = ti < tile1 ? V.win[ti] : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int ti = t + 64 * i;
                if (ti < tile1) yb[ti] = (double)FB_FMUL((float)x[i], wv[i]);
            }
        }
        asm volatile("bar.arrive %0, 96;" ::"r"(2 + b) : "memory");
    };
    for (int k = 0; k < n_tiles && k < NB; k++) stage_tile(k);

    // ---- pass E (warps 1 and 2): min / max over all samples, entropy estimate partition by partition
    const int w = (int)warp - 1;
    int32_t mn = 2147483647, mx = -2147483647 - 1;
    unsigned long long bsum = 0; // lane k < 5: the bits of order k over this warp's partitions
    if (V.do_ent) {
        const int parts = J.cfg.approx_ent_partitions;
        uint32_t flagged = 0; // local partitions (bit i = partition w + 2 i) that need the sequential replay
        for (int p = w, i = 0; p < parts; p += 2, i++) {
            const int t0 = p * V.psize, t1 = t0 + V.psize < n ? t0 + V.psize : n;
            if (t0 >= t1) break; // (empty trailing partitions contribute nothing: NaN -> 0, src/coding.rs:219-222)
            uint32_t j0 = 0, j1 = 0, j2 = 0, j3 = 0, j4 = 0;
            for (int t = t0 + (int)lane; t < t1; t += 32) {
                const uint32_t x0 = (uint32_t)fb_k1c_sample(J, xt, pcm, rows, mb, sh, f, t);
                const uint32_t x1 = t >= 1 ? (uint32_t)fb_k1c_sample(J, xt, pcm, rows, mb, sh, f, t - 1) : 0u;
                const uint32_t x2 = t >= 2 ? (uint32_t)fb_k1c_sample(J, xt, pcm, rows, mb, sh, f, t - 2) : 0u;
                const uint32_t x3 = t >= 3 ? (uint32_t)fb_k1c_sample(J, xt, pcm, rows, mb, sh, f, t - 3) : 0u;
                const uint32_t x4 = t >= 4 ? (uint32_t)fb_k1c_sample(J, xt, pcm, rows, mb, sh, f, t - 4) : 0u;
                mn = (int32_t)x0 < mn ? (int32_t)x0 : mn;
                mx = (int32_t)x0 > mx ? (int32_t)x0 : mx;
                // zero-history k-th differences (src/coding.rs:188-195), wrapping
                const uint32_t d1 = x0 - x1, d1p = x1 - x2, d1pp = x2 - x3, d1ppp = x3 - x4;
                const uint32_t d2 = d1 - d1p, d2p = d1p - d1pp, d2pp = d1pp - d1ppp;
                const uint32_t d3 = d2 - d2p, d3p = d2p - d2pp;
                const uint32_t d4 = d3 - d3p;
                // (sums of clamped terms: at most 2^24 * 2048 rows of a lane -- block sizes are below 2^15 * 32 -- fits)
                j0 += fb_k1s_abs24(x0);
                j1 += fb_k1s_abs24(d1);
                j2 += fb_k1s_abs24(d2);
                j3 += fb_k1s_abs24(d3);
                j4 += fb_k1s_abs24(d4);
                j0 = j0 < (1u << 24) ? j0 : (1u << 24);
                j1 = j1 < (1u << 24) ? j1 : (1u << 24);
                j2 = j2 < (1u << 24) ? j2 : (1u << 24);
                j3 = j3 < (1u << 24) ? j3 : (1u << 24);
                j4 = j4 < (1u << 24) ? j4 : (1u << 24);
            }
            for (int d = 16; d; d >>= 1) { // (32 lanes x 2^24 fits)
                j0 += __shfl_xor_sync(0xFFFFFFFFu, j0, d);
                j1 += __shfl_xor_sync(0xFFFFFFFFu, j1, d);
                j2 += __shfl_xor_sync(0xFFFFFFFFu, j2, d);
                j3 += __shfl_xor_sync(0xFFFFFFFFu, j3, d);
                j4 += __shfl_xor_sync(0xFFFFFFFFu, j4, d);
            }
            if ((j0 | j1 | j2 | j3 | j4) < (1u << 24)) {
                // exact: the f32 sums of the reference are these integers; lane k estimates order k
                const uint32_t js = lane == 0 ? j0 : lane == 1 ? j1 : lane == 2 ? j2 : lane == 3 ? j3 : j4;
                if (lane < 5) bsum += fb_k1_part_bits((float)js, (int)lane, t1, t1 - t0);
            } else {
                flagged |= 1u << i;
            }
        }
        if (flagged) { // (warp-uniform) lane i replays local partition i
            unsigned long long r[5] = {0, 0, 0, 0, 0};
            if ((flagged >> lane) & 1u) {
                const int p = w + 2 * (int)lane;
                const int t0 = p * V.psize, t1 = t0 + V.psize < n ? t0 + V.psize : n;
                float s[5];
                fb_k1s_replay(J, xt, pcm, rows, mb, sh, f, t0, t1, s);
                for (int k = 0; k < 5; k++) r[k] = fb_k1_part_bits(s[k], k, t1, t1 - t0);
            }
            for (int k = 0; k < 5; k++) {
                for (int d = 16; d; d >>= 1) r[k] += __shfl_xor_sync(0xFFFFFFFFu, r[k], d);
                if ((int)lane == k) bsum += r[k];
            }
        }
    } else {
        for (int t = w * 32 + (int)lane; t < n; t += 64) {
            const int32_t e0 = fb_k1c_sample(J, xt, pcm, rows, mb, sh, f, t);
            mn = e0 < mn ? e0 : mn;
            mx = e0 > mx ? e0 : mx;
        }
    }
    for (int d = 16; d; d >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, d));
        mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, d));
    }
    for (int k = NB; k < n_tiles; k++) stage_tile(k); // (frames of more than NB tiles: the chain waits for pass E once)
    if (warp == 2) {
        if (lane < 5) xb[lane] = bsum;
        if (lane == 0) { xm[0] = mn; xm[1] = mx; }
    }
    asm volatile("bar.sync 1, 64;" ::: "memory"); // warps 1 and 2 only
    if (warp == 1) {
        if (lane < 5) bsum += xb[lane];
        FbK1Ent S;
        fb_k1_ent_init(S, n, V.psize, 0);
        S.b0 = __shfl_sync(0xFFFFFFFFu, bsum, 0);
        S.b1 = __shfl_sync(0xFFFFFFFFu, bsum, 1);
        S.b2 = __shfl_sync(0xFFFFFFFFu, bsum, 2);
        S.b3 = __shfl_sync(0xFFFFFFFFu, bsum, 3);
        S.b4 = __shfl_sync(0xFFFFFFFFu, bsum, 4);
        S.xmin = min(mn, xm[0]);
        S.xmax = max(mx, xm[1]);
        if (lane == 0) fb_k1_finish_ent(J, V, S, out, taps);
    }
}

// K1C: direct-MSE LPC estimator, one CTA per channel variant (fb_kernels.cuh)
__global__ void __launch_bounds__(384) fb_k1c_direct_mse(FbJob J, const int32_t *xt, const uint8_t *pcm, const float *win_full,
                                                         const float *win_tail, FbAnalysis *ana, fb200_variant_taps *taps) {
    extern __shared__ __align__(16) uint8_t fb_smem_k1c[];
    fb_k1c_body(J, xt, pcm, win_full, win_tail, ana, taps, blockIdx.x, FB_K1C_TILE, fb_smem_k1c);
}

// K1I: IRLS-MAE refinement of the direct-MSE estimate, one CTA per channel variant (fb_kernels.cuh)
__global__ void __launch_bounds__(384) fb_k1i_irls_mae(FbJob J, const int32_t *xt, const uint8_t *pcm, const float *win_full,
                                                       const float *win_tail, FbAnalysis *ana, fb200_variant_taps *taps) {
    extern __shared__ __align__(16) uint8_t fb_smem_k1i[];
    fb_k1i_body(J, xt, pcm, win_full, win_tail, ana, taps, blockIdx.x, FB_K1I_TILE, fb_smem_k1i);
}

// K1D: direct-MSE estimator for lpc_order 10 and large launches, thread per channel variant (fb_kernels.cuh)
__global__ void __maxnreg__(224) fb_k1d_direct_mse10(FbJob J, const int32_t *xt, const uint8_t *pcm, const float *win_full,
                                                     const float *win_tail, FbAnalysis *ana, fb200_variant_taps *taps,
                                                     uint32_t n_variants) {
    extern __shared__ __align__(16) uint8_t fb_smem_k1d[];
    if (pcm) fb_k1d_warp<true>(J, xt, pcm, win_full, win_tail, ana, taps, n_variants, fb_smem_k1d);
    else fb_k1d_warp<false>(J, xt, pcm, win_full, win_tail, ana, taps, n_variants, fb_smem_k1d);
}

// offsets[i] = *total + exclusive prefix; *total advances by the chunk's bytes (single CTA)
__global__ void __launch_bounds__(FB_K4_THREADS) fb_k4_scan(const uint32_t *frame_bytes, unsigned long long *offsets,
                                                           uint32_t n_frames, unsigned long long *total) {
    __shared__ unsigned long long partials[FB_K4_THREADS];
    const unsigned long long base = *total;
    __syncthreads(); // everybody has read the running total before the last thread replaces it
    fb_k4_scan_body(frame_bytes, offsets, n_frames, partials, base);
    if (threadIdx.x == FB_K4_THREADS - 1) *total = offsets[n_frames];
}

// gather of the frames on the fused path's fallback list only (everything else is stored in place by KP)
__global__ void __launch_bounds__(256) fb_k4_gather_list(const uint8_t *slots, uint32_t slot_bytes, const uint32_t *frame_bytes,
                                                         const unsigned long long *offsets, uint8_t *out,
                                                         unsigned long long out_cap, const uint32_t *list, const uint32_t *count) {
    const uint32_t total = *count;
    for (uint32_t i = blockIdx.x; i < total; i += gridDim.x)
        fb_k4_gather_thread(slots, slot_bytes, frame_bytes, offsets, out, out_cap, list[i], (int)threadIdx.x, 256);
}

__global__ void __launch_bounds__(256) fb_k4_gather(const uint8_t *slots, uint32_t slot_bytes, const uint32_t *frame_bytes,
                                                    const unsigned long long *offsets, uint8_t *out,
                                                    unsigned long long out_cap) {
    fb_k4_gather_thread(slots, slot_bytes, frame_bytes, offsets, out, out_cap, blockIdx.x, (int)threadIdx.x, 256);
}

// parity pins: device builds of fb_log2f / fb_find_shift on caller-provided inputs (fb200_debug_*)
__global__ void __launch_bounds__(256) fb_dbg_log2f(uint32_t first_bits, uint64_t count, uint32_t *out) {
    for (uint64_t i = (uint64_t)blockIdx.x * 256u + threadIdx.x; i < count; i += (uint64_t)gridDim.x * 256u)
        out[i] = fb_f2u(fb_log2f(fb_u2f(first_bits + (uint32_t)i)));
}
__global__ void __launch_bounds__(256) fb_dbg_irls_weight(uint32_t first_bits, uint64_t count, float normalizer, uint32_t *out) {
    for (uint64_t i = (uint64_t)blockIdx.x * 256u + threadIdx.x; i < count; i += (uint64_t)gridDim.x * 256u)
        out[i] = fb_f2u(fb_irls_weight(fb_u2f(first_bits + (uint32_t)i), normalizer));
}
__global__ void __launch_bounds__(256) fb_dbg_find_shift(const double *values, uint64_t count, int precision, int32_t *out) {
    for (uint64_t i = (uint64_t)blockIdx.x * 256u + threadIdx.x; i < count; i += (uint64_t)gridDim.x * 256u)
        out[i] = fb_find_shift(values + i, 1, precision);
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

// Device buffers, events and pinned staging of one chunk in flight.  Set 0 also serves the serial path.
struct ChunkSet {
    DevBuf pcm, xv, xv4, ana, taps, choice, slots, frame_bytes, offsets, out, infos, fb_list, scalars, plan, psubs, poffs, lpc_ext;
    // events: 0 H2D start, 1 H2D end, 2 ingest end, 3 analyze end, 4 rice/fused end, 5 pack/fallback end,
    //         6 gather end (= chunk done), 7 D2H start, 8 D2H end, 9 H2D end on the copy stream (pipelined path),
    //         10 start of the back half (fused kernel onwards) on its stream
    cudaEvent_t ev[11];
    int n_ev = 0;
    uint8_t *pinned = nullptr; // [0, 64): copy of the device scalars; [64, ...): frame sizes of the chunk
    size_t pinned_cap = 0;
    bool d2h_pending = false;  // ev[8] has been recorded for a chunk whose D2H time is not accounted yet
};

#define FB_NSETS_MAX 6

struct fb200_ctx {
    fb200_config cfg;
    int channels, bps, sample_rate, block_size, device;
    cudaStream_t stream = nullptr;   // serial path and compute stream 0
    cudaStream_t s_k1 = nullptr;     // compute stream 1 (pipelined path: chunks alternate)
    cudaStream_t s_in = nullptr, s_out = nullptr; // H2D / D2H streams of the pipelined path
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    ChunkSet sets[FB_NSETS_MAX];
    int nsets = 3; // buffer sets the pipelined host path rotates (FB200_NSETS, 2..6)
    DevBuf win_full, win_tail, ktab;
    int win_tail_n = -1;
    fb200_timing timing;
    std::string last_error;
    uint32_t ktab_chunk = 0; // CRC chunk length the uploaded tables were built for
    bool no_k1d = false;        // FB200_K1D=0: direct MSE always by the CTA-per-variant kernel (tests exercise both)
    int k1_small = -1;          // FB200_K1_SMALL=0/1: never / always analyse with a warp per variant (default: by launch size)
    bool no_pairs = false;      // FB200_KP_PAIRS=0: the pack kernel always stages planes (tests exercise both)
    bool force_generic = false; // FB200_FORCE_GENERIC=1: never use the fused kernel (tests exercise both paths)
    uint32_t serial_mib = 4096;     // FB200_SERIAL_MIB: working-set bound of a chunk of the serial (device-resident) path
    uint64_t pipe_chunk_frames = 0; // FB200_CHUNK_FRAMES: frames per chunk of the pipelined host path (0 = default)
    std::mutex mu;
};

#define FB_CUDA(ctx, call)                                                                           \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            (ctx)->last_error = std::string(#call) + ": " + cudaGetErrorString(e__);                 \
            return FB200_ERR_CUDA;                                                                   \
        }                                                                                            \
    } while (0)

static int fb_reserve(fb200_ctx *ctx, DevBuf &b, size_t bytes) {
    if (bytes <= b.cap) return FB200_OK;
    if (b.p) {
        FB_CUDA(ctx, cudaDeviceSynchronize());
        FB_CUDA(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    FB_CUDA(ctx, cudaMalloc(&b.p, want));
    b.cap = want;
    return FB200_OK;
}

static int fb_reserve_pinned(fb200_ctx *ctx, ChunkSet &S, size_t bytes) {
    if (bytes <= S.pinned_cap) return FB200_OK;
    if (S.pinned) {
        FB_CUDA(ctx, cudaDeviceSynchronize());
        FB_CUDA(ctx, cudaFreeHost(S.pinned));
        S.pinned = nullptr;
        S.pinned_cap = 0;
    }
    FB_CUDA(ctx, cudaHostAlloc((void **)&S.pinned, bytes + 256, cudaHostAllocDefault));
    S.pinned_cap = bytes + 256;
    return FB200_OK;
}

extern "C" {

void fb200_config_default(fb200_config *cfg) { fbh_config_default(cfg); }
int fb200_config_verify(const fb200_config *cfg) { return fbh_config_verify(cfg); }

int fb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char *fb200_strerror(int code) {
    switch (code) {
    case FB200_OK: return "ok";
    case FB200_ERR_CONFIG: return "config/verify error (EncodeError::Config)";
    case FB200_ERR_SOURCE: return "source/argument error (EncodeError::Source)";
    case FB200_ERR_CUDA: return "CUDA error (no CPU fallback)";
    case FB200_ERR_CAPACITY: return "output buffer too small";
    default: return "unknown";
    }
}

static int fb_dbg_run(int device, const void *in, size_t in_bytes, void *out, size_t out_bytes,
                      const std::function<void(const void *, void *)> &launch) {
    if (device < 0 || device >= fb200_device_count()) return FB200_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return FB200_ERR_CUDA;
    void *d_in = nullptr, *d_out = nullptr;
    int rc = FB200_OK;
    if ((in_bytes && cudaMalloc(&d_in, in_bytes) != cudaSuccess) || cudaMalloc(&d_out, out_bytes) != cudaSuccess) rc = FB200_ERR_CUDA;
    if (!rc && in_bytes && cudaMemcpy(d_in, in, in_bytes, cudaMemcpyHostToDevice) != cudaSuccess) rc = FB200_ERR_CUDA;
    if (!rc) {
        launch(d_in, d_out);
        if (cudaGetLastError() != cudaSuccess || cudaMemcpy(out, d_out, out_bytes, cudaMemcpyDeviceToHost) != cudaSuccess)
            rc = FB200_ERR_CUDA;
    }
    cudaFree(d_in);
    cudaFree(d_out);
    return rc;
}

int fb200_debug_log2f(int device, uint32_t first_bits, uint64_t count, uint32_t *out_bits) {
    if (!out_bits || count == 0 || count > (1ull << 28)) return FB200_ERR_SOURCE;
    return fb_dbg_run(device, nullptr, 0, out_bits, (size_t)count * 4u, [&](const void *, void *d_out) {
        fb_dbg_log2f<<<1184, 256>>>(first_bits, count, (uint32_t *)d_out);
    });
}

int fb200_debug_irls_weight(int device, uint32_t first_bits, uint64_t count, float normalizer, uint32_t *out_bits) {
    if (!out_bits || count == 0 || count > (1ull << 28)) return FB200_ERR_SOURCE;
    return fb_dbg_run(device, nullptr, 0, out_bits, (size_t)count * 4u, [&](const void *, void *d_out) {
        fb_dbg_irls_weight<<<1184, 256>>>(first_bits, count, normalizer, (uint32_t *)d_out);
    });
}

int fb200_debug_find_shift(int device, const double *values, uint64_t count, int precision, int32_t *out_shift) {
    if (!values || !out_shift || count == 0 || count > (1ull << 26) || precision < 1 || precision > 15) return FB200_ERR_SOURCE;
    return fb_dbg_run(device, values, (size_t)count * 8u, out_shift, (size_t)count * 4u, [&](const void *d_in, void *d_out) {
        fb_dbg_find_shift<<<296, 256>>>((const double *)d_in, count, precision, (int32_t *)d_out);
    });
}

const char *fb200_version(void) { return "flacenc_b200 0.2.0 (sm_100a)"; }

// ctx == NULL: detail text of the last failed stream-level call (fb200_encode_stream / _streams) on this thread
static thread_local std::string g_stream_error;
const char *fb200_last_error(const fb200_ctx *ctx) { return ctx ? ctx->last_error.c_str() : g_stream_error.c_str(); }

fb200_ctx *fb200_create(const fb200_config *cfg, int channels, int bits_per_sample, int sample_rate, int block_size,
                        int device, int *err) {
    int dummy;
    if (!err) err = &dummy;
    *err = FB200_OK;
    if (!cfg) { *err = FB200_ERR_SOURCE; return nullptr; }
    if ((*err = fbh_config_verify(cfg)) != FB200_OK) return nullptr;
    if ((*err = fbh_format_verify(channels, bits_per_sample, sample_rate, block_size)) != FB200_OK) return nullptr;
    int ndev = fb200_device_count();
    if (device < 0 || device >= ndev) { *err = FB200_ERR_CUDA; return nullptr; }
    fb200_ctx *ctx = new fb200_ctx();
    ctx->cfg = *cfg;
    ctx->channels = channels;
    ctx->bps = bits_per_sample;
    ctx->sample_rate = sample_rate;
    ctx->block_size = block_size;
    ctx->device = device;
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    {
        const char *fg = getenv("FB200_FORCE_GENERIC");
        ctx->force_generic = fg && fg[0] == '1';
        const char *kp = getenv("FB200_KP_PAIRS");
        ctx->no_pairs = kp && kp[0] == '0';
        const char *kd = getenv("FB200_K1D");
        ctx->no_k1d = kd && kd[0] == '0';
        const char *ks = getenv("FB200_K1_SMALL");
        if (ks) ctx->k1_small = ks[0] == '1' ? 1 : 0;
        const char *cf = getenv("FB200_CHUNK_FRAMES");
        if (cf) ctx->pipe_chunk_frames = strtoull(cf, nullptr, 10);
        const char *sm = getenv("FB200_SERIAL_MIB");
        if (sm && atoi(sm) >= 16) ctx->serial_mib = (uint32_t)atoi(sm);
        const char *ns = getenv("FB200_NSETS");
        if (ns) ctx->nsets = std::max(2, std::min(FB_NSETS_MAX, atoi(ns)));
    }
    bool ok = cudaSetDevice(device) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->s_k1, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&ctx->ev_begin) == cudaSuccess && cudaEventCreate(&ctx->ev_end) == cudaSuccess;
    for (int k = 0; ok && k < FB_NSETS_MAX; k++) {
        ChunkSet &S = ctx->sets[k];
        for (int i = 0; ok && i < 11; i++) {
            ok = cudaEventCreate(&S.ev[i]) == cudaSuccess;
            if (ok) S.n_ev = i + 1;
        }
        ok = ok && cudaHostAlloc((void **)&S.pinned, 4096, cudaHostAllocDefault) == cudaSuccess;
        if (ok) S.pinned_cap = 4096;
    }
    if (!ok) {
        *err = FB200_ERR_CUDA;
        fb200_destroy(ctx);
        return nullptr;
    }
    return ctx;
}

void fb200_destroy(fb200_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (ChunkSet &S : ctx->sets) {
        DevBuf *bufs[] = {&S.pcm, &S.xv, &S.xv4, &S.ana, &S.taps, &S.choice, &S.slots, &S.frame_bytes, &S.offsets, &S.out,
                          &S.infos, &S.fb_list, &S.scalars, &S.plan, &S.psubs, &S.poffs, &S.lpc_ext};
        for (DevBuf *b : bufs)
            if (b->p) cudaFree(b->p);
        for (int i = 0; i < S.n_ev; i++) cudaEventDestroy(S.ev[i]);
        if (S.pinned) cudaFreeHost(S.pinned);
    }
    DevBuf *bufs[] = {&ctx->win_full, &ctx->win_tail, &ctx->ktab};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
    if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
    cudaStream_t sts[] = {ctx->stream, ctx->s_k1, ctx->s_in, ctx->s_out};
    for (cudaStream_t st : sts)
        if (st) cudaStreamDestroy(st);
    delete ctx;
}

size_t fb200_max_frame_bytes(const fb200_ctx *ctx) {
    return ctx ? fb_max_frame_bytes(ctx->channels, ctx->bps, ctx->block_size) : 0;
}

int fb200_last_timing(const fb200_ctx *ctx, fb200_timing *t) {
    if (!ctx || !t) return FB200_ERR_SOURCE;
    *t = ctx->timing;
    return FB200_OK;
}

} // extern "C"

// ------------------------------------------------------------------------------------------------
// the pipeline
// ------------------------------------------------------------------------------------------------
namespace {

struct EncodeArgs {
    const void *pcm_host = nullptr;   // host interleaved PCM (or nullptr)
    const void *pcm_dev = nullptr;    // device interleaved PCM (or nullptr)
    const int32_t *planar_host = nullptr;
    int planar_stride = 0;
    int container_bytes = 0;
    uint64_t n_samples = 0;
    uint64_t first_frame = 0;
    uint8_t *out_host = nullptr;
    uint8_t *out_dev = nullptr;
    size_t out_cap = 0;
    uint32_t *frame_sizes = nullptr;  // host
    fb200_frame_info *infos = nullptr; // host
    fb200_variant_taps *taps = nullptr; // host (analyze only)
    size_t taps_cap = 0;
    bool analyze_only = false;
    size_t *n_frames = nullptr;
    size_t *out_len = nullptr;
    size_t *n_variants = nullptr;
};

// everything about a call that does not change from chunk to chunk
struct Plan {
    int cb = 0, nvar = 0, ring = 0, tail_n = 0;
    uint64_t bs = 0, total_frames = 0, stride = 0, slot_bytes = 0;
    uint32_t mb = 0;
    const float *d_win_tail = nullptr;
    FbK2Layout L;
    FbKfLayout KL, KLp, KPL, KPLp; // shared-memory layouts of the plan kernel and of the pack kernel (planes / PCM pairs)
    bool kp_pairs = false;    // 16-bit stereo in a 2-byte container: the pack kernel may stage the PCM itself
    bool pairs = false;       // ... and so may the analysis and plan kernels: no ingest kernel, no planar store
    size_t k2_smem = 0, k3_smem = 0;
    bool fused = false;
};

struct Accum {
    bool fused = false; // ms_k[2..4]: fused = plan kernel | fallback kernels + scan | pack kernel (+ list gather);
                        //             generic = Rice search | assembly | scan + gather
    float ms_h2d = 0, ms_d2h = 0, ms_k[5] = {0, 0, 0, 0, 0};
    uint64_t launches = 0, fused_frames = 0, fallback = 0;
};

int fb_upload_window(fb200_ctx *ctx, DevBuf &buf, int n) {
    std::vector<float> w((size_t)n + 64, 0.f);
    fbh_window_weights(ctx->cfg.window_type, ctx->cfg.tukey_alpha, n, w.data());
    int rc = fb_reserve(ctx, buf, w.size() * sizeof(float));
    if (rc) return rc;
    FB_CUDA(ctx, cudaMemcpyAsync(buf.p, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // `w` dies at scope end
    return FB200_OK;
}

// geometry, window tables, CRC tables, shared-memory opt-in
int fb_make_plan(fb200_ctx *ctx, const EncodeArgs &A, Plan &P) {
    int rc;
    P.cb = A.container_bytes ? A.container_bytes : 4;
    P.bs = (uint64_t)ctx->block_size;
    P.total_frames = (A.n_samples + P.bs - 1) / P.bs;
    P.nvar = ctx->channels == 2 ? 4 : ctx->channels;
    P.stride = (uint64_t)((ctx->block_size + 31) & ~31);
    P.mb = fb_max_frame_bytes(ctx->channels, ctx->bps, ctx->block_size);
    P.slot_bytes = (P.mb + 15u) & ~15u;
    P.ring = fb_k1_ring(ctx->cfg.lpc_order);
    // window tables (src/lpc.rs:217-231: cached per size)
    if (ctx->win_full.p == nullptr && (rc = fb_upload_window(ctx, ctx->win_full, ctx->block_size))) return rc;
    P.tail_n = (A.n_samples % P.bs) ? (int)(A.n_samples % P.bs) : ctx->block_size;
    if (P.tail_n != ctx->block_size && P.tail_n != ctx->win_tail_n) {
        if ((rc = fb_upload_window(ctx, ctx->win_tail, P.tail_n))) return rc;
        ctx->win_tail_n = P.tail_n;
    }
    P.d_win_tail = (P.tail_n == ctx->block_size) ? (const float *)ctx->win_full.p : (const float *)ctx->win_tail.p;

    FbJob J0 = fbh_make_job(ctx->cfg, ctx->channels, ctx->bps, ctx->sample_rate, ctx->block_size, P.cb,
                            std::min<uint64_t>(A.n_samples, P.bs), (uint32_t)A.first_frame);
    const int leaves_max = 1 << fb_finest_partition_order(ctx->block_size);
    const int leaves_tail = 1 << fb_finest_partition_order(P.tail_n);
    P.L = fb_k2_layout(ctx->block_size, std::max(leaves_max, leaves_tail));
    P.k2_smem = P.L.total + 3 * sizeof(FbRiceResult);
    if (P.k2_smem > 227u * 1024u) { // (block sizes 30720 .. 32512 with 256+ finest partitions: without the bank padding)
        P.L = fb_k2_layout(ctx->block_size, std::max(leaves_max, leaves_tail), 31);
        P.k2_smem = P.L.total + 3 * sizeof(FbRiceResult);
    }
    P.k3_smem = fb_k3_smem_bytes(P.mb, ctx->block_size, J0.pack_in_smem);
    if (P.k2_smem > 227u * 1024u || P.k3_smem > 227u * 1024u) {
        ctx->last_error = "internal: shared memory budget exceeded";
        return FB200_ERR_CUDA;
    }
    P.fused = !A.analyze_only && !ctx->force_generic && fbh_fused_ok(J0, P.tail_n, &P.KL);
    P.KPL = fb_kp_layout(ctx->channels, J0.nvar, ctx->bps, ctx->block_size, P.tail_n);
    P.kp_pairs = !ctx->no_pairs && fb_kp_pairs_format(ctx->channels, ctx->bps, P.cb) && A.planar_host == nullptr;
    P.KPLp = fb_kp_layout(ctx->channels, J0.nvar, ctx->bps, ctx->block_size, P.tail_n, P.kp_pairs);
    P.pairs = !ctx->no_pairs && fb_pairs_format(ctx->channels, ctx->bps, P.cb, ctx->block_size) && A.planar_host == nullptr &&
              (P.fused || A.analyze_only);
    P.KLp = fb_kf_layout(ctx->channels, J0.nvar, ctx->bps, ctx->block_size, P.tail_n, false, P.pairs);
    if (P.fused && ctx->ktab_chunk != P.KL.crc_chunk) {
        std::vector<uint32_t> kt(fb_kf_ktab_words(P.KL.crc_chunk, 32u * (uint32_t)J0.nvar));
        fb_kf_build_ktab(P.KL.crc_chunk, 32u * (uint32_t)J0.nvar, kt.data());
        if ((rc = fb_reserve(ctx, ctx->ktab, kt.size() * 4u))) return rc;
        FB_CUDA(ctx, cudaMemcpyAsync(ctx->ktab.p, kt.data(), kt.size() * 4u, cudaMemcpyHostToDevice, ctx->stream));
        FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // `kt` dies at scope end
        ctx->ktab_chunk = P.KL.crc_chunk;
    }
    // Dynamic shared memory opt-in of the kernels of this tap-window size: a per-function, per-device attribute that is
    // shared by every context of the process, so it is raised to the device maximum once and never lowered (a context
    // that set it to its own smaller need would break the launches of another one).
    {
        static std::mutex smem_mu;
        static bool smem_done[64][32];
        std::lock_guard<std::mutex> lock(smem_mu);
        const int di = ctx->device & 63, ri = (P.ring / 4) & 31;
        if (!smem_done[di][ri]) {
            int optin = 0;
            FB_CUDA(ctx, cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
            FB_CUDA(ctx, fb_set_smem(P.ring, FB_KERNEL_KF, optin));
            FB_CUDA(ctx, fb_set_smem(P.ring, FB_KERNEL_K2, optin));
            FB_CUDA(ctx, fb_set_smem(P.ring, FB_KERNEL_K3, optin));
            smem_done[di][ri] = true;
        }
    }
    return FB200_OK;
}

// device buffers of a set for chunks of up to `frames` frames
int fb_reserve_set(fb200_ctx *ctx, const Plan &P, const EncodeArgs &A, ChunkSet &S, uint64_t frames, uint64_t in_bytes,
                   uint64_t sizes_frames, uint64_t out_bytes) {
    int rc;
    if ((rc = fb_reserve(ctx, S.scalars, 64))) return rc;
    if ((rc = fb_reserve(ctx, S.frame_bytes, sizes_frames * 4u))) return rc;
    if ((rc = fb_reserve(ctx, S.offsets, (frames + 1) * 8u))) return rc;
    if ((rc = fb_reserve(ctx, S.xv, (fb_xt_words((int)P.stride, frames * (uint64_t)ctx->channels) + 64) * 4u))) return rc;
    if (!A.analyze_only && (rc = fb_reserve(ctx, S.xv4, (frames * (uint64_t)P.nvar * P.stride + 64) * 4u))) return rc;
    if ((rc = fb_reserve(ctx, S.ana, frames * (uint64_t)P.nvar * sizeof(FbAnalysis)))) return rc;
    if ((ctx->cfg.ext_lpc_order_search > 0 || ctx->cfg.ext_lpc_precision_search > 0) &&
        (rc = fb_reserve(ctx, S.lpc_ext, frames * (uint64_t)P.nvar * FB_EXT_LPC_MAX * sizeof(FbLpcExt))))
        return rc;
    if (in_bytes && (rc = fb_reserve(ctx, S.pcm, in_bytes + 16))) return rc;
    if (A.analyze_only) {
        if ((rc = fb_reserve(ctx, S.taps, frames * (uint64_t)P.nvar * sizeof(fb200_variant_taps)))) return rc;
    } else {
        if ((rc = fb_reserve(ctx, S.choice, frames * (uint64_t)P.nvar * sizeof(fb200_subframe_info)))) return rc;
        if ((rc = fb_reserve(ctx, S.slots, frames * P.slot_bytes))) return rc;
        if ((rc = fb_reserve(ctx, S.fb_list, (frames + 1) * 4u))) return rc;
        if (P.fused) {
            if ((rc = fb_reserve(ctx, S.plan, frames * sizeof(FbKfPlan)))) return rc;
            if ((rc = fb_reserve(ctx, S.psubs, frames * (uint64_t)ctx->channels * sizeof(fb200_subframe_info)))) return rc;
            if ((rc = fb_reserve(ctx, S.poffs, frames * (uint64_t)ctx->channels * (P.KL.U_max + 1) * 4u))) return rc;
        }
        if (A.infos && (rc = fb_reserve(ctx, S.infos, frames * sizeof(fb200_frame_info)))) return rc;
        if (out_bytes && (rc = fb_reserve(ctx, S.out, out_bytes))) return rc;
    }
    return FB200_OK;
}

// Enqueues K0..K4 of one chunk on `st`.  d_pcm: the chunk's interleaved PCM on the device (nullptr: planar frame in
// S.pcm).  Frame sizes go to d_fb[0..nf), bytes to d_out + offsets starting at *d_total (which advances).
// Records S.ev[1..6].
int fb_enqueue_kernels(fb200_ctx *ctx, const Plan &P, const EncodeArgs &A, ChunkSet &S, uint64_t f0, uint64_t ns,
                       const uint8_t *d_pcm, uint32_t *d_fb, uint8_t *d_out, unsigned long long out_cap, cudaStream_t st,
                       Accum &acc, cudaStream_t st_back = nullptr, unsigned long long *d_total_shared = nullptr) {
    // front half (ingest, analysis) on `st`; back half (fused kernel .. gather) on `st_back` when given, so the
    // analysis of the next chunk can overlap the fused kernel of this one
    FbJob J = fbh_make_job(ctx->cfg, ctx->channels, ctx->bps, ctx->sample_rate, ctx->block_size, P.cb, ns,
                           (uint32_t)(A.first_frame + f0));
    if (ctx->cfg.ext_lpc_order_search > 0 || ctx->cfg.ext_lpc_precision_search > 0)
        J.lpc_ext = (FbLpcExt *)S.lpc_ext.p; // (extensions: alternative LPC sets, K1 -> KA / K2)
    const uint32_t nvars = J.n_frames * (uint32_t)J.nvar;
    uint32_t *d_err = (uint32_t *)S.scalars.p;
    unsigned long long *d_total = d_total_shared ? d_total_shared : (unsigned long long *)((uint8_t *)S.scalars.p + 8);
    uint32_t *d_fb_count = (uint32_t *)((uint8_t *)S.scalars.p + 16);
    // pairs mode: 16-bit stereo PCM (16-byte aligned) is read as (left, right) pairs by every kernel: no ingest, no xt
    // (a 16-bit sample in a 2-byte container cannot be out of range, so the range check of the ingest kernel is moot)
    const bool pairs = P.pairs && d_pcm && ((uintptr_t)d_pcm & 15u) == 0;
    const uint8_t *pcm_pairs = pairs ? d_pcm : nullptr;
    FB_CUDA(ctx, cudaEventRecord(S.ev[1], st));
    if (pairs) {
    } else if (A.planar_host) {
        fb_k0_ingest_planar<<<(unsigned)((J.stride / 4 + 255) / 256), 256, 0, st>>>(J, (const int32_t *)S.pcm.p,
                                                                                   A.planar_stride, (int32_t *)S.xv.p, d_err);
    } else {
        // items: ceil(frames / 16) frame groups x ceil(quads / 2) quad pairs x 32 lanes
        const uint64_t n_items = (uint64_t)((J.n_frames + 15u) / 16u) * (uint64_t)((J.stride / 4 + 1) / 2) * 32u;
        fb_k0_ingest<<<(unsigned)((n_items + 255) / 256), 256, 0, st>>>(J, d_pcm, (int32_t *)S.xv.p, d_err, n_items);
    }
    FB_CUDA(ctx, cudaEventRecord(S.ev[2], st));
    // small launches: a CTA per variant (K1S) instead of a thread per variant (K1); same results
    const bool k1_small = ctx->k1_small >= 0 ? ctx->k1_small != 0 : nvars <= 148u * 12u;
    if (k1_small) {
        const uint32_t smem = FB_K1S_NB * (FB200_MAX_LPC_ORDER + FB_K1S_TILE) * 8u + 32u * 8u + 64u;
        fb_k1s_analyze<<<nvars, FB_K1S_THREADS, smem, st>>>(
            J, (const int32_t *)S.xv.p, pcm_pairs, (const float *)ctx->win_full.p, P.d_win_tail, (FbAnalysis *)S.ana.p,
            A.analyze_only ? (fb200_variant_taps *)S.taps.p : nullptr, nvars);
    } else {
        fb_launch_k1(P.ring, J, (const int32_t *)S.xv.p, pcm_pairs, (const float *)ctx->win_full.p, P.d_win_tail,
                     (FbAnalysis *)S.ana.p, A.analyze_only ? (fb200_variant_taps *)S.taps.p : nullptr, nvars, st);
    }
    if (ctx->cfg.use_direct_mse && ctx->cfg.use_lpc) {
        // `experimental` estimator: K1 has skipped its autocorrelation pass; K1C (CTA per variant) or, for the default
        // order and launches large enough to fill the GPU with one thread per variant, K1D fills the LPC half of the records
        fb200_variant_taps *d_taps = A.analyze_only ? (fb200_variant_taps *)S.taps.p : nullptr;
        if (ctx->cfg.mae_optimization_steps > 0) {
            // IRLS-MAE refinement (src/coding.rs:337-345): K1I re-weights and re-solves mae_optimization_steps times
            fb_k1i_irls_mae<<<nvars, fb_k1i_threads(ctx->cfg.lpc_order), fb_k1i_smem_bytes(ctx->cfg.lpc_order, FB_K1I_TILE), st>>>(
                J, (const int32_t *)S.xv.p, pcm_pairs, (const float *)ctx->win_full.p, P.d_win_tail, (FbAnalysis *)S.ana.p, d_taps);
        } else if (ctx->cfg.lpc_order == FB_K1D_P && !k1_small && !ctx->no_k1d) {
            const uint32_t smem = fb_k1_smem_bytes(J.channels, J.nvar, pairs);
            if (smem > 48u * 1024u)
                FB_CUDA(ctx, cudaFuncSetAttribute(fb_k1d_direct_mse10, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            fb_k1d_direct_mse10<<<(fb_k1_slots(J, nvars) + FB_K1_THREADS - 1) / FB_K1_THREADS, FB_K1_THREADS, smem, st>>>(
                J, (const int32_t *)S.xv.p, pcm_pairs, (const float *)ctx->win_full.p, P.d_win_tail, (FbAnalysis *)S.ana.p,
                d_taps, nvars);
        } else {
            fb_k1c_direct_mse<<<nvars, fb_k1c_threads(ctx->cfg.lpc_order), fb_k1c_smem_bytes(ctx->cfg.lpc_order, FB_K1C_TILE), st>>>(
                J, (const int32_t *)S.xv.p, pcm_pairs, (const float *)ctx->win_full.p, P.d_win_tail, (FbAnalysis *)S.ana.p, d_taps);
        }
        acc.launches += 1;
    }
    FB_CUDA(ctx, cudaEventRecord(S.ev[3], st));
    acc.launches += pairs ? 1 : 2;
    if (A.analyze_only) return FB200_OK;
    if (st_back && st_back != st) {
        FB_CUDA(ctx, cudaStreamWaitEvent(st_back, S.ev[3], 0));
        st = st_back;
    }
    FB_CUDA(ctx, cudaEventRecord(S.ev[10], st));
    fb200_frame_info *d_infos = A.infos ? (fb200_frame_info *)S.infos.p : nullptr;
    // plain rows by variant for the generic kernels (K0b)
    const int32_t *xg = (const int32_t *)S.xv4.p;
    if (P.fused) {
        // KA: analysis + plan + frame sizes; frames it cannot reproduce exactly go to the list and are encoded into
        // their slots by the generic kernels.  Then the scan of all frame sizes, KP packs every planned frame
        // straight to its final offset, and the listed frames are gathered from their slots.
        FB_CUDA(ctx, cudaMemsetAsync(d_fb_count, 0, 4, st));
        fb_launch_ka(P.ring, J, (const int32_t *)S.xv.p, pcm_pairs, (const FbAnalysis *)S.ana.p, S.plan.p,
                     (fb200_subframe_info *)S.choice.p, (fb200_subframe_info *)S.psubs.p, (uint32_t *)S.poffs.p, d_fb, d_infos, (uint32_t *)S.fb_list.p,
                     d_fb_count, (const uint32_t *)ctx->ktab.p, pairs ? P.KLp : P.KL, st);
        FB_CUDA(ctx, cudaEventRecord(S.ev[4], st));
        if (pairs)
            fb_k0b_expand_pairs<<<148, 256, 0, st>>>(J, d_pcm, (int32_t *)S.xv4.p, (const uint32_t *)S.fb_list.p, d_fb_count);
        else
            fb_k0b_expand<<<148, 256, 0, st>>>(J, (const int32_t *)S.xv.p, (int32_t *)S.xv4.p, (const uint32_t *)S.fb_list.p,
                                               d_fb_count);
        fb_launch_k2(P.ring, J, xg, (const FbAnalysis *)S.ana.p, (fb200_subframe_info *)S.choice.p,
                     P.L, (const uint32_t *)S.fb_list.p, d_fb_count, 296, P.k2_smem, st);
        fb_launch_k3(P.ring, J, xg, (const fb200_subframe_info *)S.choice.p, (uint8_t *)S.slots.p,
                     d_fb, d_infos, (const uint32_t *)S.fb_list.p, d_fb_count, 148, P.k3_smem, st);
        fb_k4_scan<<<1, FB_K4_THREADS, 0, st>>>(d_fb, (unsigned long long *)S.offsets.p, J.n_frames, d_total);
        FB_CUDA(ctx, cudaEventRecord(S.ev[5], st));
        // (16-byte copies: frame f starts at f * block_size * 4 bytes, so the block size must be a multiple of 4 too)
        const bool kp_pairs = pairs || (P.kp_pairs && d_pcm && ((uintptr_t)d_pcm & 15u) == 0 && (ctx->block_size & 3) == 0);
        fb_launch_kp(P.ring, J, (const int32_t *)S.xv.p, kp_pairs ? d_pcm : nullptr, S.plan.p,
                     (const fb200_subframe_info *)S.psubs.p, (const uint32_t *)S.poffs.p,
                     (const unsigned long long *)S.offsets.p, d_out, out_cap, (const uint32_t *)ctx->ktab.p,
                     kp_pairs ? P.KPLp : P.KPL, st);
        fb_k4_gather_list<<<148, 256, 0, st>>>((const uint8_t *)S.slots.p, J.slot_bytes, d_fb,
                                               (const unsigned long long *)S.offsets.p, d_out, out_cap,
                                               (const uint32_t *)S.fb_list.p, d_fb_count);
        acc.fused_frames += J.n_frames;
        acc.fused = true;
        acc.launches += 7;
    } else {
        fb_k0b_expand<<<J.n_frames, 256, 0, st>>>(J, (const int32_t *)S.xv.p, (int32_t *)S.xv4.p, nullptr, nullptr);
        fb_launch_k2(P.ring, J, xg, (const FbAnalysis *)S.ana.p, (fb200_subframe_info *)S.choice.p,
                     P.L, nullptr, nullptr, nvars, P.k2_smem, st);
        FB_CUDA(ctx, cudaEventRecord(S.ev[4], st));
        fb_launch_k3(P.ring, J, xg, (const fb200_subframe_info *)S.choice.p, (uint8_t *)S.slots.p,
                     d_fb, d_infos, nullptr, nullptr, J.n_frames, P.k3_smem, st);
        FB_CUDA(ctx, cudaEventRecord(S.ev[5], st));
        fb_k4_scan<<<1, FB_K4_THREADS, 0, st>>>(d_fb, (unsigned long long *)S.offsets.p, J.n_frames, d_total);
        fb_k4_gather<<<J.n_frames, 256, 0, st>>>((const uint8_t *)S.slots.p, J.slot_bytes, d_fb,
                                                 (const unsigned long long *)S.offsets.p, d_out, out_cap);
        acc.launches += 5;
    }
    FB_CUDA(ctx, cudaGetLastError());
    // device scalars (error flag, running total, fallback count) for the host
    FB_CUDA(ctx, cudaMemcpyAsync(S.pinned, S.scalars.p, 24, cudaMemcpyDeviceToHost, st));
    FB_CUDA(ctx, cudaEventRecord(S.ev[6], st));
    return FB200_OK;
}

// kernel times of a finished chunk (S.ev[6] has completed)
int fb_harvest_kernel_times(fb200_ctx *ctx, ChunkSet &S, bool analyze_only, Accum &acc) {
    float t;
    const int last = analyze_only ? 2 : 5;
    for (int k = 0; k < last; k++) {
        FB_CUDA(ctx, cudaEventElapsedTime(&t, S.ev[k == 2 ? 10 : 1 + k], S.ev[2 + k]));
        acc.ms_k[k] += t;
    }
    return FB200_OK;
}

void fb_store_timing(fb200_ctx *ctx, const Accum &acc, float total_ms, uint64_t in_bytes, uint64_t out_bytes) {
    fb200_timing &T = ctx->timing;
    T.h2d_ms = acc.ms_h2d;
    T.k_ingest_ms = acc.ms_k[0];
    T.k_analyze_ms = acc.ms_k[1];
    T.k_rice_ms = acc.ms_k[2];
    T.k_pack_ms = acc.fused ? acc.ms_k[4] : acc.ms_k[3];
    T.k_gather_ms = acc.fused ? acc.ms_k[3] : acc.ms_k[4];
    T.kernels_ms = acc.ms_k[0] + acc.ms_k[1] + acc.ms_k[2] + acc.ms_k[3] + acc.ms_k[4];
    T.d2h_ms = acc.ms_d2h;
    T.total_ms = total_ms;
    T.launches = acc.launches;
    T.in_bytes = in_bytes;
    T.out_bytes = out_bytes;
    T.fused_frames = acc.fused_frames - acc.fallback;
    T.fallback_frames = acc.fallback;
}

// ---- serial path: device-resident input/output, single planar frames, analysis taps, small host batches.
// Chunks run back to back on one stream; the output offsets continue across chunks on the device.
int fb_encode_serial(fb200_ctx *ctx, const EncodeArgs &A, const Plan &P) {
    ChunkSet &S = ctx->sets[0];
    cudaStream_t st = ctx->stream;
    int rc;
    // chunking: bound the working set (planar variants + slots) per pass (FB200_SERIAL_MIB, default 4 GiB: an hour of CD
    // stereo is one chunk -- every extra chunk costs each kernel another partly filled last wave)
    const uint64_t per_frame = (uint64_t)P.nvar * P.stride * 4u + P.slot_bytes + 4096u;
    uint64_t chunk_frames = std::max<uint64_t>(64, ((uint64_t)ctx->serial_mib << 20) / per_frame);
    chunk_frames = std::min<uint64_t>(chunk_frames, P.total_frames);
    const uint64_t in_bytes_total = A.n_samples * (uint64_t)ctx->channels * (uint64_t)P.cb;
    uint64_t in_chunk = 0;
    if (A.pcm_host) in_chunk = std::min(chunk_frames * P.bs * (uint64_t)ctx->channels * (uint64_t)P.cb, in_bytes_total);
    else if (A.planar_host) in_chunk = (uint64_t)ctx->channels * (uint64_t)A.planar_stride * 4u;
    if ((rc = fb_reserve_set(ctx, P, A, S, chunk_frames, in_chunk, P.total_frames,
                             (A.out_host && !A.analyze_only) ? std::max<size_t>(A.out_cap, 16) : 0)))
        return rc;
    uint8_t *d_out = A.out_dev ? A.out_dev : (uint8_t *)S.out.p;
    FB_CUDA(ctx, cudaMemsetAsync(S.scalars.p, 0, 64, st));
    Accum acc;
    FB_CUDA(ctx, cudaEventRecord(ctx->ev_begin, st));
    for (uint64_t f0 = 0; f0 < P.total_frames; f0 += chunk_frames) {
        const uint64_t nf = std::min(chunk_frames, P.total_frames - f0);
        const uint64_t s0 = f0 * P.bs;
        const uint64_t ns = std::min(A.n_samples - s0, nf * P.bs);
        FB_CUDA(ctx, cudaEventRecord(S.ev[0], st));
        const uint8_t *d_pcm = nullptr;
        if (A.pcm_host) {
            const uint64_t off = s0 * (uint64_t)ctx->channels * (uint64_t)P.cb;
            const uint64_t len = ns * (uint64_t)ctx->channels * (uint64_t)P.cb;
            FB_CUDA(ctx, cudaMemcpyAsync(S.pcm.p, (const uint8_t *)A.pcm_host + off, len, cudaMemcpyHostToDevice, st));
            d_pcm = (const uint8_t *)S.pcm.p;
        } else if (A.pcm_dev) {
            d_pcm = (const uint8_t *)A.pcm_dev + s0 * (uint64_t)ctx->channels * (uint64_t)P.cb;
        } else {
            FB_CUDA(ctx, cudaMemcpyAsync(S.pcm.p, A.planar_host, in_chunk, cudaMemcpyHostToDevice, st));
        }
        if ((rc = fb_enqueue_kernels(ctx, P, A, S, f0, ns, d_pcm, (uint32_t *)S.frame_bytes.p + f0, d_out,
                                     (unsigned long long)A.out_cap, st, acc)))
            return rc;
        if (A.analyze_only) {
            FB_CUDA(ctx, cudaMemcpyAsync(A.taps + f0 * (uint64_t)P.nvar, S.taps.p,
                                         (size_t)(nf * (uint64_t)P.nvar) * sizeof(fb200_variant_taps),
                                         cudaMemcpyDeviceToHost, st));
            FB_CUDA(ctx, cudaStreamSynchronize(st));
            if ((rc = fb_harvest_kernel_times(ctx, S, true, acc))) return rc;
            continue;
        }
        if (A.infos)
            FB_CUDA(ctx, cudaMemcpyAsync(A.infos + f0, S.infos.p, (size_t)nf * sizeof(fb200_frame_info),
                                         cudaMemcpyDeviceToHost, st));
        // per-chunk timings need the events to have completed; chunks are serial on one stream
        FB_CUDA(ctx, cudaEventSynchronize(S.ev[6]));
        float t;
        FB_CUDA(ctx, cudaEventElapsedTime(&t, S.ev[0], S.ev[1]));
        acc.ms_h2d += t;
        if ((rc = fb_harvest_kernel_times(ctx, S, false, acc))) return rc;
        acc.fallback += *(const uint32_t *)(S.pinned + 16);
    }
    if (A.analyze_only) return FB200_OK;

    FB_CUDA(ctx, cudaStreamSynchronize(st));
    const uint32_t err_flag = *(const uint32_t *)S.pinned;
    const unsigned long long total = *(const unsigned long long *)(S.pinned + 8);
    if (err_flag) {
        ctx->last_error = "input sample out of the range of bits_per_sample";
        return FB200_ERR_CONFIG; // VerifyError (src/source.rs:262-275)
    }
    if (A.out_len) *A.out_len = (size_t)total;
    if (total > A.out_cap) {
        ctx->last_error = "output capacity too small";
        return FB200_ERR_CAPACITY;
    }
    FB_CUDA(ctx, cudaEventRecord(S.ev[7], st));
    if (A.frame_sizes)
        FB_CUDA(ctx, cudaMemcpyAsync(A.frame_sizes, S.frame_bytes.p, P.total_frames * 4u, cudaMemcpyDeviceToHost, st));
    if (A.out_host) FB_CUDA(ctx, cudaMemcpyAsync(A.out_host, d_out, (size_t)total, cudaMemcpyDeviceToHost, st));
    FB_CUDA(ctx, cudaEventRecord(S.ev[8], st));
    FB_CUDA(ctx, cudaEventRecord(ctx->ev_end, st));
    FB_CUDA(ctx, cudaStreamSynchronize(st));
    float t_total = 0;
    FB_CUDA(ctx, cudaEventElapsedTime(&acc.ms_d2h, S.ev[7], S.ev[8]));
    FB_CUDA(ctx, cudaEventElapsedTime(&t_total, ctx->ev_begin, ctx->ev_end));
    fb_store_timing(ctx, acc, t_total, in_bytes_total, total);
    return FB200_OK;
}

// chunk schedule of the pipelined path (see fb_encode_pipelined): the frames before a short final chunk (a quarter of the
// nominal size) are spread evenly over chunks of at most the nominal size
std::vector<std::pair<uint64_t, uint64_t>> fb_chunk_schedule(uint64_t total_frames, uint64_t chunk_frames) {
    std::vector<std::pair<uint64_t, uint64_t>> chunks; // (first frame, frames)
    if (total_frames == 0) return chunks;
    chunk_frames = std::max<uint64_t>(chunk_frames, 1);
    const uint64_t last = std::max<uint64_t>(1, std::min<uint64_t>(chunk_frames / 4, total_frames / 2));
    const uint64_t body = total_frames - last;
    const uint64_t nb = (body + chunk_frames - 1) / chunk_frames;
    uint64_t f = 0;
    for (uint64_t i = 0; i < nb; i++) {
        const uint64_t take = body / nb + (i < body % nb ? 1 : 0);
        chunks.push_back({f, take});
        f += take;
    }
    chunks.push_back({f, last});
    return chunks;
}

// ---- pipelined path: host PCM in, host frame bytes out.  The batch is cut into chunks; chunk c's H2D copy
// (stream s_in), kernels (compute streams, alternating so the latency-bound analysis kernel of one chunk overlaps
// the fused kernel of the previous one) and D2H copy (stream s_out) overlap with those of its neighbours.
// ctx->nsets buffer sets rotate; a set is reused once its D2H copy has been enqueued and is waited for by event.
//
// Several contexts (one per device, same configuration and format) share the work by frame range: chunk c is encoded
// on device c mod N with that device's own streams and buffer sets (what src/par.rs:355-449 does with worker threads).
// Frames are independent, so there is no exchange between the devices; the only ordered state is the byte offset of a
// chunk in the output, which the host learns from the chunk before it (each chunk's bytes are copied straight to
// their final place in the caller's buffer: no per-device staging, no gather pass).
int fb_encode_pipelined(const std::vector<fb200_ctx *> &cs, const EncodeArgs &A, const std::vector<Plan> &Ps, uint64_t chunk_frames) {
    int rc;
    fb200_ctx *ctx = cs[0];              // carries the error text and the timing record of the call
    const Plan &P = Ps[0];
    const uint64_t N = cs.size();
    const uint64_t FB_NSETS = (uint64_t)ctx->nsets;
    // chunk schedule: the H2D stream is the bottleneck, so what matters at the end is how much work is left once the
    // last copy has landed.  One short final chunk (a quarter of the nominal size) keeps that tail small; more than
    // one does not pay, because every chunk costs about a millisecond of kernel latency.  The frames before it are
    // spread evenly over chunks of at most the nominal size.
    const std::vector<std::pair<uint64_t, uint64_t>> chunks = fb_chunk_schedule(P.total_frames, chunk_frames); // (first frame, frames)
    chunk_frames = 0;
    for (auto &c : chunks) chunk_frames = std::max(chunk_frames, c.second);
    const uint64_t nchunks = chunks.size();
    const uint64_t in_bytes_total = A.n_samples * (uint64_t)ctx->channels * (uint64_t)P.cb;
    const uint64_t in_chunk = std::min(chunk_frames * P.bs * (uint64_t)ctx->channels * (uint64_t)P.cb, in_bytes_total);
    const uint64_t n_used = std::min<uint64_t>(N, nchunks); // devices that get at least one chunk
    for (uint64_t d = 0; d < n_used; d++) {
        fb200_ctx *C = cs[d];
        FB_CUDA(ctx, cudaSetDevice(C->device));
        for (uint64_t k = 0; k < FB_NSETS; k++) { // only the sets that rotate
            ChunkSet &S = C->sets[k];
            if ((rc = fb_reserve_set(C, Ps[d], A, S, chunk_frames, in_chunk, chunk_frames, chunk_frames * (uint64_t)P.mb + 16))) {
                ctx->last_error = C->last_error;
                return rc;
            }
            if ((rc = fb_reserve_pinned(C, S, 64 + chunk_frames * 4u))) { ctx->last_error = C->last_error; return rc; }
            S.d2h_pending = false;
        }
    }
    std::vector<Accum> accs(N);
    unsigned long long host_off = 0;
    int result = FB200_OK;
    // FB200_TRACE=1: device timeline per chunk (ms since the start of the call on that chunk's device) on stderr
    const bool trace = getenv("FB200_TRACE") != nullptr;
    std::vector<std::array<float, 6>> tr(trace ? nchunks : 0);
    std::vector<uint64_t> set_chunk(N * FB_NSETS, 0); // chunk whose D2H events a set currently holds
    auto dev_of = [&](uint64_t c) { return c % N; };
    auto set_of = [&](uint64_t c) { return (c / N) % FB_NSETS; };
    auto trace_d2h = [&](fb200_ctx *C, ChunkSet &S, uint64_t c) {
        if (!trace) return;
        cudaEventElapsedTime(&tr[c][4], C->ev_begin, S.ev[7]);
        cudaEventElapsedTime(&tr[c][5], C->ev_begin, S.ev[8]);
    };
    for (uint64_t d = 0; d < n_used; d++) {
        FB_CUDA(ctx, cudaSetDevice(cs[d]->device));
        FB_CUDA(ctx, cudaEventRecord(cs[d]->ev_begin, cs[d]->s_in));
    }

    auto chunk_range = [&](uint64_t c, uint64_t &f0, uint64_t &nf, uint64_t &s0, uint64_t &ns) {
        f0 = chunks[c].first;
        nf = chunks[c].second;
        s0 = f0 * P.bs;
        ns = std::min(A.n_samples - s0, nf * P.bs);
    };
    auto enqueue = [&](uint64_t c) -> int {
        fb200_ctx *C = cs[dev_of(c)];
        ChunkSet &S = C->sets[set_of(c)];
        Accum &acc = accs[dev_of(c)];
        uint64_t f0, nf, s0, ns;
        chunk_range(c, f0, nf, s0, ns);
        FB_CUDA(ctx, cudaSetDevice(C->device));
        if (S.d2h_pending) {
            // the set's previous chunk: its D2H must have drained before the buffers are overwritten
            FB_CUDA(ctx, cudaEventSynchronize(S.ev[8]));
            float t;
            FB_CUDA(ctx, cudaEventElapsedTime(&t, S.ev[7], S.ev[8]));
            acc.ms_d2h += t;
            trace_d2h(C, S, set_chunk[dev_of(c) * FB_NSETS + set_of(c)]);
            S.d2h_pending = false;
        }
        const uint64_t off = s0 * (uint64_t)ctx->channels * (uint64_t)P.cb;
        const uint64_t len = ns * (uint64_t)ctx->channels * (uint64_t)P.cb;
        FB_CUDA(ctx, cudaEventRecord(S.ev[0], C->s_in));
        FB_CUDA(ctx, cudaMemcpyAsync(S.pcm.p, (const uint8_t *)A.pcm_host + off, len, cudaMemcpyHostToDevice, C->s_in));
        cudaStream_t sk = ((c / N) & 1) ? C->s_k1 : C->stream;
        FB_CUDA(ctx, cudaEventRecord(S.ev[9], C->s_in));
        FB_CUDA(ctx, cudaStreamWaitEvent(sk, S.ev[9], 0));
        FB_CUDA(ctx, cudaMemsetAsync(S.scalars.p, 0, 64, sk));
        const int rc2 = fb_enqueue_kernels(C, Ps[dev_of(c)], A, S, f0, ns, (const uint8_t *)S.pcm.p, (uint32_t *)S.frame_bytes.p,
                                           (uint8_t *)S.out.p, (unsigned long long)S.out.cap, sk, acc);
        if (rc2 && C != ctx) ctx->last_error = C->last_error;
        return rc2;
    };
    auto finish = [&](uint64_t c) -> int {
        fb200_ctx *C = cs[dev_of(c)];
        ChunkSet &S = C->sets[set_of(c)];
        Accum &acc = accs[dev_of(c)];
        uint64_t f0, nf, s0, ns;
        chunk_range(c, f0, nf, s0, ns);
        FB_CUDA(ctx, cudaSetDevice(C->device));
        FB_CUDA(ctx, cudaEventSynchronize(S.ev[6]));
        float t;
        FB_CUDA(ctx, cudaEventElapsedTime(&t, S.ev[0], S.ev[9]));
        acc.ms_h2d += t;
        int rc2;
        if ((rc2 = fb_harvest_kernel_times(C, S, false, acc))) { ctx->last_error = C->last_error; return rc2; }
        const uint32_t err_flag = *(const uint32_t *)S.pinned;
        const unsigned long long total = *(const unsigned long long *)(S.pinned + 8);
        acc.fallback += *(const uint32_t *)(S.pinned + 16);
        if (err_flag && result == FB200_OK) {
            ctx->last_error = "input sample out of the range of bits_per_sample";
            result = FB200_ERR_CONFIG; // VerifyError (src/source.rs:262-275)
        }
        if (host_off + total > A.out_cap && result == FB200_OK) {
            ctx->last_error = "output capacity too small";
            result = FB200_ERR_CAPACITY;
        }
        FB_CUDA(ctx, cudaEventRecord(S.ev[7], C->s_out));
        if (result == FB200_OK) {
            if (A.out_host && total)
                FB_CUDA(ctx, cudaMemcpyAsync(A.out_host + host_off, S.out.p, (size_t)total, cudaMemcpyDeviceToHost, C->s_out));
            if (A.frame_sizes) {
                // via pinned staging so the copy stays asynchronous; handed over after the final synchronize
                FB_CUDA(ctx, cudaMemcpyAsync(S.pinned + 64, S.frame_bytes.p, (size_t)nf * 4u, cudaMemcpyDeviceToHost, C->s_out));
            }
            if (A.infos)
                FB_CUDA(ctx, cudaMemcpyAsync(A.infos + f0, S.infos.p, (size_t)nf * sizeof(fb200_frame_info),
                                             cudaMemcpyDeviceToHost, C->s_out));
        }
        FB_CUDA(ctx, cudaEventRecord(S.ev[8], C->s_out));
        S.d2h_pending = true;
        set_chunk[dev_of(c) * FB_NSETS + set_of(c)] = c;
        if (trace) {
            const int idx[4] = {0, 9, 1, 6};
            for (int i = 0; i < 4; i++) cudaEventElapsedTime(&tr[c][i], C->ev_begin, S.ev[idx[i]]);
        }
        host_off += total;
        return FB200_OK;
    };
    // frame sizes staged in a set's pinned area must be copied out before the set is reused
    auto collect_sizes = [&](uint64_t c) -> int {
        if (!A.frame_sizes || result != FB200_OK) return FB200_OK;
        fb200_ctx *C = cs[dev_of(c)];
        ChunkSet &S = C->sets[set_of(c)];
        uint64_t f0, nf, s0, ns;
        chunk_range(c, f0, nf, s0, ns);
        FB_CUDA(ctx, cudaSetDevice(C->device));
        FB_CUDA(ctx, cudaEventSynchronize(S.ev[8]));
        memcpy(A.frame_sizes + f0, S.pinned + 64, (size_t)nf * 4u);
        return FB200_OK;
    };

    const uint64_t in_flight = N * FB_NSETS;       // chunks that own a buffer set at any time
    const uint64_t ahead = N * (FB_NSETS - 1);
    for (uint64_t c = 0; c < nchunks + ahead; c++) {
        if (c >= in_flight && (rc = collect_sizes(c - in_flight))) return rc; // before chunk c reuses that set
        if (c < nchunks && (rc = enqueue(c))) return rc;
        if (c >= ahead && (rc = finish(c - ahead))) return rc;
    }
    for (uint64_t c = nchunks > in_flight ? nchunks - in_flight : 0; c < nchunks; c++)
        if ((rc = collect_sizes(c))) return rc;
    float t_total = 0;
    Accum acc;
    for (uint64_t d = 0; d < n_used; d++) {
        fb200_ctx *C = cs[d];
        FB_CUDA(ctx, cudaSetDevice(C->device));
        FB_CUDA(ctx, cudaEventRecord(C->ev_end, C->s_out));
        FB_CUDA(ctx, cudaStreamSynchronize(C->s_out));
        FB_CUDA(ctx, cudaStreamSynchronize(C->stream));
        FB_CUDA(ctx, cudaStreamSynchronize(C->s_k1));
        for (int k = 0; k < (int)FB_NSETS; k++) {
            ChunkSet &S = C->sets[k];
            if (S.d2h_pending) {
                float t;
                FB_CUDA(ctx, cudaEventElapsedTime(&t, S.ev[7], S.ev[8]));
                accs[d].ms_d2h += t;
                trace_d2h(C, S, set_chunk[d * FB_NSETS + (uint64_t)k]);
                S.d2h_pending = false;
            }
        }
        // the call's device time: the longest span (first copy in .. last copy out) over the devices
        float t_dev = 0;
        FB_CUDA(ctx, cudaEventElapsedTime(&t_dev, C->ev_begin, C->ev_end));
        t_total = std::max(t_total, t_dev);
        const Accum &a = accs[d];
        acc.fused |= a.fused;
        acc.ms_h2d += a.ms_h2d;
        acc.ms_d2h += a.ms_d2h;
        for (int k = 0; k < 5; k++) acc.ms_k[k] += a.ms_k[k];
        acc.launches += a.launches;
        acc.fused_frames += a.fused_frames;
        acc.fallback += a.fallback;
    }
    FB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (A.out_len) *A.out_len = (size_t)host_off;
    if (result != FB200_OK) return result;
    if (trace) {
        for (uint64_t c = 0; c < nchunks; c++)
            fprintf(stderr, "fb200 trace: chunk %llu dev %d frames %llu  h2d %.3f-%.3f  kernels %.3f-%.3f  d2h %.3f-%.3f\n",
                    (unsigned long long)c, cs[dev_of(c)]->device, (unsigned long long)chunks[c].second, tr[c][0], tr[c][1], tr[c][2],
                    tr[c][3], tr[c][4], tr[c][5]);
    }
    fb_store_timing(ctx, acc, t_total, in_bytes_total, host_off);
    return FB200_OK;
}

// checks of an encode call that do not depend on the device; fills the frame / variant counts
int fb_encode_prologue(fb200_ctx *ctx, const EncodeArgs &A, uint64_t *total_frames_out) {
    ctx->last_error.clear();
    const int cb = A.container_bytes;
    if (!A.planar_host && (cb < 1 || cb > 4)) {
        ctx->last_error = "container_bytes must be 1, 2, 3 or 4";
        return FB200_ERR_SOURCE;
    }
    if (cb * 8 < ctx->bps && !A.planar_host) {
        ctx->last_error = "container narrower than bits_per_sample";
        return FB200_ERR_SOURCE;
    }
    const uint64_t bs = (uint64_t)ctx->block_size;
    const uint64_t total_frames = (A.n_samples + bs - 1) / bs;
    if (A.n_frames) *A.n_frames = (size_t)total_frames;
    if (A.out_len) *A.out_len = 0;
    const int nvar = ctx->channels == 2 ? 4 : ctx->channels;
    if (A.n_variants) *A.n_variants = (size_t)(total_frames * (uint64_t)nvar);
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    *total_frames_out = total_frames;
    if (total_frames == 0) return FB200_OK;
    // frame_number < 2^31 (src/coding.rs:587-591)
    if (A.first_frame + total_frames > (1ull << 31)) {
        ctx->last_error = "frame number out of the 31-bit range";
        return FB200_ERR_CONFIG;
    }
    if (A.analyze_only && (uint64_t)A.taps_cap < total_frames * (uint64_t)nvar) return FB200_ERR_CAPACITY;
    return FB200_OK;
}

// frames per chunk of the pipelined host path: sized so that the per-variant analysis kernel still has a few thousand
// threads and the copies of neighbouring chunks overlap the kernels; with several devices every device should get
// about six chunks
uint64_t fb_pipe_chunk_frames(const fb200_ctx *ctx, const Plan &P, uint64_t total_frames, uint64_t n_devices) {
    uint64_t chunk = ctx->pipe_chunk_frames ? ctx->pipe_chunk_frames : 2432;
    if (!ctx->pipe_chunk_frames && n_devices > 1)
        chunk = std::min<uint64_t>(chunk, std::max<uint64_t>(304, total_frames / (n_devices * 6)));
    // a chunk should not be much more than the 40 MB of input that 2432 CD-stereo frames are (64 MiB at most): wide formats (8 channels of
    // 24 bits: 98 KB per frame) get proportionally fewer frames per chunk, so that their copies and kernels overlap too
    const uint64_t in_per_frame = P.bs * (uint64_t)ctx->channels * (uint64_t)P.cb;
    if (!ctx->pipe_chunk_frames) chunk = std::min<uint64_t>(chunk, std::max<uint64_t>(64, (64ull << 20) / std::max<uint64_t>(in_per_frame, 1)));
    const uint64_t per_frame = (uint64_t)P.nvar * P.stride * 4u + 2 * P.slot_bytes + 4096u;
    return std::min<uint64_t>(chunk, std::max<uint64_t>(64, (1024ull << 20) / per_frame));
}

int fb_encode(fb200_ctx *ctx, const EncodeArgs &A) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    FB_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t total_frames = 0;
    int rc = fb_encode_prologue(ctx, A, &total_frames);
    if (rc || total_frames == 0) return rc;
    Plan P;
    rc = fb_make_plan(ctx, A, P);
    if (rc) return rc;
    // host batches of more than one chunk are pipelined
    if (A.pcm_host && !A.analyze_only) {
        const uint64_t chunk = fb_pipe_chunk_frames(ctx, P, total_frames, 1);
        if (total_frames > chunk + chunk / 2) return fb_encode_pipelined({ctx}, A, {P}, chunk);
    }
    return fb_encode_serial(ctx, A, P);
}

// one batch over several contexts (one per device, same configuration and format): frame ranges, chunk by chunk
int fb_encode_multi(const std::vector<fb200_ctx *> &cs, const EncodeArgs &A) {
    fb200_ctx *ctx = cs[0];
    if (cs.size() == 1) return fb_encode(ctx, A);
    std::vector<std::unique_lock<std::mutex>> locks;
    for (fb200_ctx *C : cs) locks.emplace_back(C->mu);
    for (fb200_ctx *C : cs) {
        if (memcmp(&C->cfg, &ctx->cfg, sizeof(fb200_config)) != 0 || C->channels != ctx->channels || C->bps != ctx->bps ||
            C->sample_rate != ctx->sample_rate || C->block_size != ctx->block_size) {
            ctx->last_error = "contexts of one sharded call must share configuration and stream format";
            return FB200_ERR_SOURCE;
        }
    }
    uint64_t total_frames = 0;
    int rc = fb_encode_prologue(ctx, A, &total_frames);
    if (rc || total_frames == 0) return rc;
    std::vector<Plan> Ps(cs.size());
    for (size_t d = 0; d < cs.size(); d++) {
        FB_CUDA(ctx, cudaSetDevice(cs[d]->device));
        if ((rc = fb_make_plan(cs[d], A, Ps[d]))) { ctx->last_error = cs[d]->last_error; return rc; }
    }
    const uint64_t chunk = fb_pipe_chunk_frames(ctx, Ps[0], total_frames, cs.size());
    if (total_frames <= chunk + chunk / 2) { // too small to share
        FB_CUDA(ctx, cudaSetDevice(ctx->device));
        return fb_encode_serial(ctx, A, Ps[0]);
    }
    return fb_encode_pipelined(cs, A, Ps, chunk);
}

} // namespace

extern "C" {

int fb200_encode_interleaved(fb200_ctx *ctx, const void *pcm, int container_bytes, uint64_t n_samples_per_ch,
                             uint64_t first_frame_number, uint8_t *out_bytes, size_t out_cap, uint32_t *frame_sizes,
                             fb200_frame_info *infos, size_t *n_frames, size_t *out_len) {
    if (!ctx || (!pcm && n_samples_per_ch) || (!out_bytes && out_cap)) return FB200_ERR_SOURCE;
    EncodeArgs A;
    A.pcm_host = pcm;
    A.container_bytes = container_bytes;
    A.n_samples = n_samples_per_ch;
    A.first_frame = first_frame_number;
    A.out_host = out_bytes;
    A.out_cap = out_cap;
    A.frame_sizes = frame_sizes;
    A.infos = infos;
    A.n_frames = n_frames;
    A.out_len = out_len;
    return fb_encode(ctx, A);
}

int fb200_encode_device(fb200_ctx *ctx, const void *d_pcm, int container_bytes, uint64_t n_samples_per_ch,
                        uint64_t first_frame_number, uint8_t *d_out, size_t out_cap, uint32_t *frame_sizes,
                        size_t *n_frames, size_t *out_len) {
    if (!ctx || (!d_pcm && n_samples_per_ch) || !d_out) return FB200_ERR_SOURCE;
    EncodeArgs A;
    A.pcm_dev = d_pcm;
    A.container_bytes = container_bytes;
    A.n_samples = n_samples_per_ch;
    A.first_frame = first_frame_number;
    A.out_dev = d_out;
    A.out_cap = out_cap;
    A.frame_sizes = frame_sizes;
    A.n_frames = n_frames;
    A.out_len = out_len;
    return fb_encode(ctx, A);
}

int fb200_encode_planar_frame(fb200_ctx *ctx, const int32_t *planar, int stride, int n, uint32_t frame_number,
                              uint8_t *out, size_t out_cap, size_t *out_len, fb200_frame_info *info) {
    if (!ctx || !planar || !out) return FB200_ERR_SOURCE;
    if (n < 1 || n > ctx->block_size || stride < n) {
        ctx->last_error = "planar frame: need 1 <= n <= block_size and stride >= n";
        return FB200_ERR_SOURCE;
    }
    EncodeArgs A;
    A.planar_host = planar;
    A.planar_stride = stride;
    A.container_bytes = 4;
    A.n_samples = (uint64_t)n;
    A.first_frame = frame_number;
    A.out_host = out;
    A.out_cap = out_cap;
    A.infos = info;
    A.out_len = out_len;
    return fb_encode(ctx, A);
}

int fb200_analyze(fb200_ctx *ctx, const void *pcm, int container_bytes, uint64_t n_samples_per_ch,
                  fb200_variant_taps *taps, size_t taps_cap, size_t *n_variants) {
    if (!ctx || !pcm || !taps) return FB200_ERR_SOURCE;
    EncodeArgs A;
    A.pcm_host = pcm;
    A.container_bytes = container_bytes;
    A.n_samples = n_samples_per_ch;
    A.taps = taps;
    A.taps_cap = taps_cap;
    A.analyze_only = true;
    A.n_variants = n_variants;
    return fb_encode(ctx, A);
}

// Same hot path over several devices: ctxs[0..n_ctx) are contexts of one configuration and stream format on different
// devices; the frames are shared by frame range (chunk c on device c mod n_ctx, what src/par.rs:355-449 does with
// worker threads) and every chunk's bytes land at their final offset in out_bytes.  No collective is involved.
int fb200_encode_interleaved_sharded(fb200_ctx *const *ctxs, int n_ctx, const void *pcm, int container_bytes,
                                     uint64_t n_samples_per_ch, uint64_t first_frame_number, uint8_t *out_bytes,
                                     size_t out_cap, uint32_t *frame_sizes, size_t *n_frames, size_t *out_len) {
    if (!ctxs || n_ctx < 1 || n_ctx > 64 || (!pcm && n_samples_per_ch) || (!out_bytes && out_cap)) return FB200_ERR_SOURCE;
    std::vector<fb200_ctx *> cs;
    for (int i = 0; i < n_ctx; i++) {
        if (!ctxs[i]) return FB200_ERR_SOURCE;
        for (fb200_ctx *c : cs)
            if (c == ctxs[i]) return FB200_ERR_SOURCE; // a context can take part once
        cs.push_back(ctxs[i]);
    }
    EncodeArgs A;
    A.pcm_host = pcm;
    A.container_bytes = container_bytes;
    A.n_samples = n_samples_per_ch;
    A.first_frame = first_frame_number;
    A.out_host = out_bytes;
    A.out_cap = out_cap;
    A.frame_sizes = frame_sizes;
    A.n_frames = n_frames;
    A.out_len = out_len;
    return fb_encode_multi(cs, A);
}

} // extern "C"

namespace {

// ---- contexts kept between stream-level calls (the reference keeps its scratch in thread-local `reusable!` storage,
// src/lib.rs:92-116): device buffers, window and CRC tables are built once per (config, format, device)
struct PoolEntry {
    fb200_config cfg;
    int channels, bps, rate, block, device;
    fb200_ctx *ctx;
};
std::mutex g_pool_mu;
std::vector<PoolEntry> g_pool; // idle contexts
const size_t FB_POOL_MAX = 16;

fb200_ctx *fb_pool_acquire(const fb200_config *cfg, int channels, int bps, int rate, int block, int device, int *err) {
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        for (size_t i = 0; i < g_pool.size(); i++) {
            PoolEntry &e = g_pool[i];
            if (memcmp(&e.cfg, cfg, sizeof(*cfg)) == 0 && e.channels == channels && e.bps == bps && e.rate == rate &&
                e.block == block && e.device == device) {
                fb200_ctx *c = e.ctx;
                g_pool.erase(g_pool.begin() + (long)i);
                *err = FB200_OK;
                return c;
            }
        }
    }
    return fb200_create(cfg, channels, bps, rate, block, device, err);
}

void fb_pool_release(fb200_ctx *ctx) {
    if (!ctx) return;
    fb200_ctx *victim = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        g_pool.push_back({ctx->cfg, ctx->channels, ctx->bps, ctx->sample_rate, ctx->block_size, ctx->device, ctx});
        if (g_pool.size() > FB_POOL_MAX) {
            victim = g_pool.front().ctx;
            g_pool.erase(g_pool.begin());
        }
    }
    if (victim) fb200_destroy(victim);
}

// MD5 over the packed little-endian samples of ceil(bps/8) bytes (src/source.rs:406-429)
void fb_md5_of_pcm(const void *pcm, int container_bytes, int bits_per_sample, uint64_t count, uint8_t md5[16]) {
    FbMd5 h;
    const int bytes_per_sample = (bits_per_sample + 7) / 8;
    if (bytes_per_sample == container_bytes) {
        h.update((const uint8_t *)pcm, (size_t)(count * (uint64_t)container_bytes));
    } else if (container_bytes == 4 && bytes_per_sample == 2) {
        // int32 samples (Fill::fill_interleaved) hashed as 16-bit little-endian values, no packed copy
        h.update_i32_as_le16((const int32_t *)pcm, (size_t)count);
    } else {
        uint8_t tmp[3 * 4096];
        size_t k = 0;
        const uint8_t *p = (const uint8_t *)pcm;
        for (uint64_t i = 0; i < count; i++) {
            for (int b = 0; b < bytes_per_sample; b++) tmp[k++] = p[i * (uint64_t)container_bytes + (uint64_t)b];
            if (k + 4 > sizeof(tmp)) { h.update(tmp, k); k = 0; }
        }
        if (k) h.update(tmp, k);
    }
    h.finish(md5);
}

// "fLaC" + STREAMINFO in front of frames that already sit at out + 42
// (src/component/datatype.rs:514-523, src/coding.rs:676-693, src/component/bitrepr.rs:240-267)
void fb_write_stream_header(uint8_t *p, const uint32_t *sizes, uint64_t n_frames, uint64_t n_samples, int channels,
                            int bits_per_sample, int sample_rate, int block_size, const uint8_t md5[16]) {
    const uint64_t bs = (uint64_t)block_size;
    uint32_t min_block = 0xFFFF, max_block = 0, min_frame = 0xFFFFFFFFu, max_frame = 0;
    for (uint64_t fi = 0; fi < n_frames; fi++) {
        const uint32_t b = (uint32_t)std::min<uint64_t>(bs, n_samples - fi * bs);
        min_block = std::min(min_block, b);
        max_block = std::max(max_block, b);
        min_frame = std::min(min_frame, sizes[fi]);
        max_frame = std::max(max_frame, sizes[fi]);
    }
    if (n_frames > 0) min_block = max_block;
    memcpy(p, "fLaC", 4);
    p[4] = 0x80; p[5] = 0; p[6] = 0; p[7] = 34;
    p[8] = (uint8_t)(min_block >> 8); p[9] = (uint8_t)min_block;
    p[10] = (uint8_t)(max_block >> 8); p[11] = (uint8_t)max_block;
    p[12] = (uint8_t)(min_frame >> 16); p[13] = (uint8_t)(min_frame >> 8); p[14] = (uint8_t)min_frame;
    p[15] = (uint8_t)(max_frame >> 16); p[16] = (uint8_t)(max_frame >> 8); p[17] = (uint8_t)max_frame;
    // 20 bits rate | 3 bits channels-1 | 5 bits bps-1 | 36 bits total samples
    const uint64_t v = ((uint64_t)(uint32_t)sample_rate << 44) | ((uint64_t)(channels - 1) << 41) |
                       ((uint64_t)(bits_per_sample - 1) << 36) | (n_samples & 0xFFFFFFFFFull);
    for (int i = 0; i < 8; i++) p[18 + i] = (uint8_t)(v >> (56 - 8 * i));
    memcpy(p + 26, md5, 16);
}

// the frames of one stream into out + 42 over the given devices (pooled contexts); *frames_len = their bytes
int fb_stream_frames(const fb200_config *cfg, const void *pcm, int container_bytes, uint64_t n_samples, int channels,
                     int bits_per_sample, int sample_rate, int block_size, const int *devices, int nd, uint8_t *out,
                     size_t out_cap, std::vector<uint32_t> &sizes, size_t *frames_len) {
    *frames_len = 0;
    if (n_samples == 0) return FB200_OK;
    std::vector<fb200_ctx *> cs;
    int rc = FB200_OK;
    for (int d = 0; d < nd && rc == FB200_OK; d++) {
        bool dup = false;
        for (int k = 0; k < d; k++) dup |= devices[k] == devices[d];
        if (dup) continue; // a device listed twice works once
        int err = 0;
        fb200_ctx *c = fb_pool_acquire(cfg, channels, bits_per_sample, sample_rate, block_size, devices[d], &err);
        if (!c) rc = err ? err : FB200_ERR_CUDA;
        else cs.push_back(c);
    }
    if (rc == FB200_OK) {
        EncodeArgs A;
        A.pcm_host = pcm;
        A.container_bytes = container_bytes;
        A.n_samples = n_samples;
        A.out_host = out + 42;
        A.out_cap = out_cap - 42;
        A.frame_sizes = sizes.data();
        size_t nf = 0;
        A.n_frames = &nf;
        A.out_len = frames_len;
        rc = fb_encode_multi(cs, A);
        if (rc) g_stream_error = cs[0]->last_error;
    } else {
        g_stream_error = "could not create a context on one of the devices";
    }
    for (fb200_ctx *c : cs) fb_pool_release(c);
    return rc;
}

} // namespace

extern "C" {

// encode_with_fixed_block_size (src/coding.rs:645-695) with par.rs-style sharding: frame ranges over the devices,
// MD5 on its own host thread (src/par.rs:196-277) next to the device work, STREAMINFO finalised on the calling thread.
// The frames are encoded straight into `out` behind the 42 header bytes; contexts are kept between calls.
int fb200_encode_stream(const fb200_config *cfg, const void *pcm, int container_bytes, uint64_t n_samples,
                        int channels, int bits_per_sample, int sample_rate, int block_size, const int *devices,
                        int n_devices, uint8_t *out, size_t out_cap, size_t *out_len) {
    if (!cfg || (!pcm && n_samples) || !out) return FB200_ERR_SOURCE;
    int rc = fbh_config_verify(cfg);
    if (rc) return rc;
    if ((rc = fbh_format_verify(channels, bits_per_sample, sample_rate, block_size))) return rc;
    if (container_bytes < 1 || container_bytes > 4 || container_bytes * 8 < bits_per_sample) return FB200_ERR_SOURCE;
    if (out_cap < 42) return FB200_ERR_CAPACITY;
    int dev0 = 0;
    if (!devices || n_devices <= 0) { devices = &dev0; n_devices = 1; }
    const uint64_t bs = (uint64_t)block_size;
    const uint64_t n_frames = (n_samples + bs - 1) / bs;
    if (n_frames > (1ull << 31)) return FB200_ERR_CONFIG;

    uint8_t md5[16];
    std::thread md5_thread([&] { fb_md5_of_pcm(pcm, container_bytes, bits_per_sample, n_samples * (uint64_t)channels, md5); });
    std::vector<uint32_t> sizes((size_t)std::max<uint64_t>(n_frames, 1));
    size_t frames_len = 0;
    rc = fb_stream_frames(cfg, pcm, container_bytes, n_samples, channels, bits_per_sample, sample_rate, block_size, devices,
                          n_devices, out, out_cap, sizes, &frames_len);
    md5_thread.join();
    if (out_len) *out_len = 42 + frames_len;
    if (rc) return rc;
    fb_write_stream_header(out, sizes.data(), n_frames, n_samples, channels, bits_per_sample, sample_rate, block_size, md5);
    return FB200_OK;
}

// A batch of streams of one format (the 10 h / 8-channel batch of BASELINE config 5 is many files): one MD5 thread per
// stream, all started up front and bounded by the host's cores, while the calling thread feeds the devices stream by
// stream.  rcs[i] (nullable) receives every stream's result; the call returns the first failure.
int fb200_encode_streams(const fb200_config *cfg, int n_streams, const void *const *pcm, const uint64_t *n_samples,
                         int container_bytes, int channels, int bits_per_sample, int sample_rate, int block_size,
                         const int *devices, int n_devices, uint8_t *const *out, const size_t *out_cap, size_t *out_len,
                         int *rcs) {
    if (!cfg || n_streams < 0 || (n_streams && (!pcm || !n_samples || !out || !out_cap))) return FB200_ERR_SOURCE;
    int rc = fbh_config_verify(cfg);
    if (rc) return rc;
    if ((rc = fbh_format_verify(channels, bits_per_sample, sample_rate, block_size))) return rc;
    if (container_bytes < 1 || container_bytes > 4 || container_bytes * 8 < bits_per_sample) return FB200_ERR_SOURCE;
    int dev0 = 0;
    if (!devices || n_devices <= 0) { devices = &dev0; n_devices = 1; }
    std::vector<std::array<uint8_t, 16>> md5((size_t)n_streams);
    // MD5 workers: streams are claimed in order by min(streams, cores) threads
    std::atomic<int> next(0);
    const int nthreads = std::max(1, std::min<int>(n_streams, (int)std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    for (int w = 0; w < nthreads; w++)
        pool.emplace_back([&] {
            for (int i = next.fetch_add(1); i < n_streams; i = next.fetch_add(1))
                if (pcm[i] || n_samples[i] == 0)
                    fb_md5_of_pcm(pcm[i], container_bytes, bits_per_sample, n_samples[i] * (uint64_t)channels, md5[(size_t)i].data());
        });
    int first_rc = FB200_OK;
    std::vector<std::vector<uint32_t>> sizes((size_t)n_streams);
    std::vector<size_t> flen((size_t)n_streams, 0);
    std::vector<int> src((size_t)n_streams, FB200_OK);
    // device feeders: a few host threads per device, each with its own (pooled) context, claim the streams in order --
    // a stream of a batch is encoded on ONE device, the batch is what spreads over the devices.  Several feeders per
    // device let the host-side staging of one stream's pageable copies overlap the kernels of another.
    std::atomic<int> next_enc(0);
    std::mutex err_mu;
    std::string batch_err;
    int per_dev = 4;
    if (const char *e = getenv("FB200_FEEDERS")) per_dev = std::max(1, std::min(16, atoi(e)));
    const int nfeed = std::max(1, std::min(n_streams, per_dev * n_devices));
    std::vector<std::thread> feeders;
    for (int w = 0; w < nfeed; w++)
        feeders.emplace_back([&, w] {
            const int dev = devices[w % n_devices];
            for (int i = next_enc.fetch_add(1); i < n_streams; i = next_enc.fetch_add(1)) {
                int r = FB200_OK;
                const uint64_t nf = (n_samples[i] + (uint64_t)block_size - 1) / (uint64_t)block_size;
                if ((!pcm[i] && n_samples[i]) || !out[i]) r = FB200_ERR_SOURCE;
                else if (out_cap[i] < 42) r = FB200_ERR_CAPACITY;
                else if (nf > (1ull << 31)) r = FB200_ERR_CONFIG;
                else {
                    sizes[(size_t)i].resize((size_t)std::max<uint64_t>(nf, 1));
                    r = fb_stream_frames(cfg, pcm[i], container_bytes, n_samples[i], channels, bits_per_sample, sample_rate,
                                         block_size, &dev, 1, out[i], out_cap[i], sizes[(size_t)i], &flen[(size_t)i]);
                }
                src[(size_t)i] = r;
                if (r) {
                    std::lock_guard<std::mutex> lock(err_mu);
                    if (batch_err.empty()) batch_err = g_stream_error;
                }
            }
        });
    for (auto &t : feeders) t.join();
    if (!batch_err.empty()) g_stream_error = batch_err;
    for (int i = 0; i < n_streams; i++)
        if (src[(size_t)i] && first_rc == FB200_OK) first_rc = src[(size_t)i];
    for (auto &t : pool) t.join();
    for (int i = 0; i < n_streams; i++) {
        if (out_len) out_len[i] = src[(size_t)i] == FB200_OK || src[(size_t)i] == FB200_ERR_CAPACITY ? 42 + flen[(size_t)i] : 0;
        if (rcs) rcs[i] = src[(size_t)i];
        if (src[(size_t)i] == FB200_OK) {
            const uint64_t nf = (n_samples[i] + (uint64_t)block_size - 1) / (uint64_t)block_size;
            fb_write_stream_header(out[i], sizes[(size_t)i].data(), nf, n_samples[i], channels, bits_per_sample, sample_rate,
                                   block_size, md5[(size_t)i].data());
        }
    }
    return first_rc;
}

// The chunk schedule of the (sharded) host path for a batch of total_frames frames and a nominal chunk size: chunk c =
// frames [first[c], first[c] + count[c]) and is encoded on device c mod n_devices.  Host logic only (no device needed);
// returns the number of chunks (fills at most `cap` entries).
size_t fb200_debug_chunk_schedule(uint64_t total_frames, uint64_t chunk_frames, uint64_t *first, uint64_t *count, size_t cap) {
    const auto chunks = fb_chunk_schedule(total_frames, chunk_frames);
    for (size_t i = 0; i < chunks.size() && i < cap; i++) {
        if (first) first[i] = chunks[i].first;
        if (count) count[i] = chunks[i].second;
    }
    return chunks.size();
}

// MD5 of a byte range with the library's own implementation (the floor of every stream-level call: sequential host work)
void fb200_md5(const void *data, size_t len, uint8_t digest[16]) {
    FbMd5 h;
    h.update((const uint8_t *)data, len);
    h.finish(digest);
}

// destroys the contexts kept by the stream-level calls
void fb200_pool_clear(void) {
    std::vector<PoolEntry> victims;
    {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        victims.swap(g_pool);
    }
    for (PoolEntry &e : victims) fb200_destroy(e.ctx);
}

} // extern "C"
