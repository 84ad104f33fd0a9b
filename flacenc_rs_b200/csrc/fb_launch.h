// fb_launch.h -- launchers of the kernels that are instantiated per tap-window size G (fb_inst.cu is
// compiled once per G so the six instantiations build in parallel); fb_api.cu dispatches on
// G = fb_k1_ring(cfg.lpc_order).
#pragma once

#include <cuda_runtime.h>

#include "fb_host.h"

enum { FB_KERNEL_K2 = 2, FB_KERNEL_K3 = 3, FB_KERNEL_KF = 5 };

#define FB_DECLARE_LAUNCHERS(G)                                                                                       \
    void fb_launch_k1_g##G(const FbJob &J, const int32_t *xv, const uint8_t *pcm_pairs, const float *win_full,       \
                           const float *win_tail, FbAnalysis *ana, fb200_variant_taps *taps, uint32_t nvars,         \
                           cudaStream_t st);                                                                         \
    void fb_launch_k2_g##G(const FbJob &J, const int32_t *xv, const FbAnalysis *ana, fb200_subframe_info *choice,     \
                           const FbK2Layout &L, const uint32_t *list, const uint32_t *count, uint32_t grid,          \
                           size_t smem, cudaStream_t st);                                                            \
    void fb_launch_k3_g##G(const FbJob &J, const int32_t *xv, const fb200_subframe_info *choice, uint8_t *slots,      \
                           uint32_t *frame_bytes, fb200_frame_info *infos, const uint32_t *list,                     \
                           const uint32_t *count, uint32_t grid, size_t smem, cudaStream_t st);                      \
    void fb_launch_ka_g##G(const FbJob &J, const int32_t *xt, const uint8_t *pcm_pairs, const FbAnalysis *ana, void *plan, \
                           fb200_subframe_info *vsubs, fb200_subframe_info *psubs, uint32_t *poffs,                 \
                           uint32_t *frame_bytes,                                                                   \
                           fb200_frame_info *infos, uint32_t *fb_list, uint32_t *fb_count, const uint32_t *ktab,    \
                           const FbKfLayout &L, cudaStream_t st);                                                   \
    void fb_launch_kp_g##G(const FbJob &J, const int32_t *xt, const uint8_t *pcm, const void *plan,                  \
                           const fb200_subframe_info *psubs,                                                        \
                           const uint32_t *poffs, const unsigned long long *offsets, uint8_t *out,                  \
                           unsigned long long out_cap, const uint32_t *ktab, const FbKfLayout &L, cudaStream_t st); \
    cudaError_t fb_set_smem_g##G(int kernel, int bytes);

FB_DECLARE_LAUNCHERS(4)
FB_DECLARE_LAUNCHERS(8)
FB_DECLARE_LAUNCHERS(12)
FB_DECLARE_LAUNCHERS(16)
FB_DECLARE_LAUNCHERS(20)
FB_DECLARE_LAUNCHERS(24)

#define FB_FOR_G(ring, CALL)                                                                                          \
    switch (ring) {                                                                                                   \
    case 4: CALL(4); break;                                                                                           \
    case 8: CALL(8); break;                                                                                           \
    case 12: CALL(12); break;                                                                                         \
    case 16: CALL(16); break;                                                                                         \
    case 20: CALL(20); break;                                                                                         \
    default: CALL(24); break;                                                                                         \
    }

#ifndef FB_INST_G // dispatchers, used by fb_api.cu only
static inline void fb_launch_k1(int ring, const FbJob &J, const int32_t *xv, const uint8_t *pcm_pairs, const float *win_full,
                                const float *win_tail, FbAnalysis *ana, fb200_variant_taps *taps, uint32_t nvars,
                                cudaStream_t st) {
#define FB_CALL(G) fb_launch_k1_g##G(J, xv, pcm_pairs, win_full, win_tail, ana, taps, nvars, st)
    FB_FOR_G(ring, FB_CALL)
#undef FB_CALL
}
// list == nullptr: units 0..grid-1 (one CTA each); else the frames of list[0..*count), grid CTAs striding over them
static inline void fb_launch_k2(int ring, const FbJob &J, const int32_t *xv, const FbAnalysis *ana,
                                fb200_subframe_info *choice, const FbK2Layout &L, const uint32_t *list,
                                const uint32_t *count, uint32_t grid, size_t smem, cudaStream_t st) {
#define FB_CALL(G) fb_launch_k2_g##G(J, xv, ana, choice, L, list, count, grid, smem, st)
    FB_FOR_G(ring, FB_CALL)
#undef FB_CALL
}
static inline void fb_launch_k3(int ring, const FbJob &J, const int32_t *xv, const fb200_subframe_info *choice,
                                uint8_t *slots, uint32_t *frame_bytes, fb200_frame_info *infos, const uint32_t *list,
                                const uint32_t *count, uint32_t grid, size_t smem, cudaStream_t st) {
#define FB_CALL(G) fb_launch_k3_g##G(J, xv, choice, slots, frame_bytes, infos, list, count, grid, smem, st)
    FB_FOR_G(ring, FB_CALL)
#undef FB_CALL
}
static inline void fb_launch_ka(int ring, const FbJob &J, const int32_t *xt, const uint8_t *pcm_pairs, const FbAnalysis *ana, void *plan,
                                fb200_subframe_info *vsubs, fb200_subframe_info *psubs, uint32_t *poffs, uint32_t *frame_bytes, fb200_frame_info *infos,
                                uint32_t *fb_list, uint32_t *fb_count, const uint32_t *ktab, const FbKfLayout &L,
                                cudaStream_t st) {
#define FB_CALL(G) fb_launch_ka_g##G(J, xt, pcm_pairs, ana, plan, vsubs, psubs, poffs, frame_bytes, infos, fb_list, fb_count, ktab, L, st)
    FB_FOR_G(ring, FB_CALL)
#undef FB_CALL
}
static inline void fb_launch_kp(int ring, const FbJob &J, const int32_t *xt, const uint8_t *pcm, const void *plan,
                                const fb200_subframe_info *psubs,
                                const uint32_t *poffs, const unsigned long long *offsets, uint8_t *out,
                                unsigned long long out_cap, const uint32_t *ktab, const FbKfLayout &L, cudaStream_t st) {
#define FB_CALL(G) fb_launch_kp_g##G(J, xt, pcm, plan, psubs, poffs, offsets, out, out_cap, ktab, L, st)
    FB_FOR_G(ring, FB_CALL)
#undef FB_CALL
}
static inline cudaError_t fb_set_smem(int ring, int kernel, int bytes) {
    cudaError_t e = cudaSuccess;
#define FB_CALL(G) e = fb_set_smem_g##G(kernel, bytes)
    FB_FOR_G(ring, FB_CALL)
#undef FB_CALL
    return e;
}
#endif
