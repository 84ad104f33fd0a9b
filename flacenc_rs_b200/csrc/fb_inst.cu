// fb_inst.cu -- the kernels that depend on the tap-window size G, compiled once per -DFB_INST_G=<4|8|12|16|20|24>
// (flacenc_rs_b200/build.py runs the six compilations in parallel).
#ifndef FB_INST_G
#error "compile with -DFB_INST_G=<G>"
#endif
#include "fb_launch.h"

#define FB_CAT2(a, b) a##b
#define FB_CAT(a, b) FB_CAT2(a, b)
#define FB_NAME(base) FB_CAT(base, FB_INST_G)

#if FB_INST_G >= 20
#define FB_K1_BOUNDS __maxnreg__(168)
#elif defined(FB_K1_MINB)
#define FB_K1_BOUNDS __launch_bounds__(FB_K1_THREADS, FB_K1_MINB) // (experiment: more resident CTAs, fewer registers)
#else
#define FB_K1_BOUNDS __launch_bounds__(FB_K1_THREADS)
#endif
// analysis: one thread per channel variant, rows staged per warp (fb_kernels.cuh)
__global__ void FB_K1_BOUNDS FB_NAME(fb_k1_analyze_g)(FbJob J, const int32_t *xt, const uint8_t *pcm, const float *win_full,
                                                                          const float *win_tail, FbAnalysis *ana,
                                                                          fb200_variant_taps *taps, uint32_t n_variants) {
    extern __shared__ __align__(16) uint8_t fb_smem[];
    if (pcm) fb_k1_warp<FB_INST_G, true>(J, xt, pcm, win_full, win_tail, ana, taps, n_variants, fb_smem);
    else fb_k1_warp<FB_INST_G, false>(J, xt, pcm, win_full, win_tail, ana, taps, n_variants, fb_smem);
}

// generic Rice search: one CTA per channel variant.  list == nullptr: variant blockIdx.x; else the variants of the
// frames in list[0..*count) (the fused kernel's fallback list, normally empty)
__global__ void __launch_bounds__(FB_K2_THREADS) FB_NAME(fb_k2_rice_g)(FbJob J, const int32_t *xv, const FbAnalysis *ana,
                                                                       fb200_subframe_info *choice, FbK2Layout L,
                                                                       const uint32_t *list, const uint32_t *count) {
    extern __shared__ __align__(16) uint8_t fb_smem[];
    if (!list) {
        fb_k2_body<FB_INST_G>(J, xv, ana, choice, blockIdx.x, fb_smem, L);
        return;
    }
    const uint32_t total = *count * (uint32_t)J.nvar;
    for (uint32_t i = blockIdx.x; i < total; i += gridDim.x) {
        const uint32_t f = list[i / (uint32_t)J.nvar];
        fb_k2_body<FB_INST_G>(J, xv, ana, choice, f * (uint32_t)J.nvar + i % (uint32_t)J.nvar, fb_smem, L);
        __syncthreads();
    }
}

// generic frame assembly: one CTA per frame (same list convention)
__global__ void __launch_bounds__(FB_K3_THREADS) FB_NAME(fb_k3_pack_g)(FbJob J, const int32_t *xv,
                                                                       const fb200_subframe_info *choice, uint8_t *slots,
                                                                       uint32_t *frame_bytes, fb200_frame_info *infos,
                                                                       const uint32_t *list, const uint32_t *count) {
    extern __shared__ __align__(16) uint8_t fb_smem[];
    if (!list) {
        fb_k3_body<FB_INST_G>(J, xv, choice, slots, frame_bytes, infos, blockIdx.x, fb_smem);
        return;
    }
    const uint32_t total = *count;
    for (uint32_t i = blockIdx.x; i < total; i += gridDim.x) {
        fb_k3_body<FB_INST_G>(J, xv, choice, slots, frame_bytes, infos, list[i], fb_smem);
        __syncthreads();
    }
}

// the pack kernel fits 80 registers for tap windows up to 16 (6 CTAs of 128 threads per SM: its shared memory, 36.9 KB
// with the CRC tables taking the place of the planes once the frame is packed, allows as many); the wide windows would spill
#if FB_INST_G <= 16
#ifndef FB_KP_MAXNREG
#define FB_KP_MAXNREG 80
#endif
#define FB_KP_BOUNDS __maxnreg__(FB_KP_MAXNREG)
#else
#define FB_KP_BOUNDS __launch_bounds__(256)
#endif

// the plan kernel of the wide windows: 168 registers = 3 CTAs of 128 threads per SM (what its shared memory allows)
#if FB_INST_G >= 20
#define FB_KA_BOUNDS __maxnreg__(168)
#define FB_KAP_BOUNDS __maxnreg__(168)
#else
#define FB_KA_BOUNDS __launch_bounds__(256)
#ifndef FB_KAP_MAXNREG
#define FB_KAP_MAXNREG 96
#endif
#define FB_KAP_BOUNDS __maxnreg__(FB_KAP_MAXNREG)
#endif

// fused path (fb_fused.cuh), one CTA of 32 * nvar threads per frame: KA = analysis + plan, KP = pack + store
// odd_mode (fb_kf_odd_mode): 0 = every frame has 4-sample aligned units; 1 = all but the last frame (its CTA comes
// first: it is the slow one); 2 = none.  The ODD instance loads its windows sample by sample.
#define FB_KA_KERNEL(NAME, BOUNDS, VMS, BC)                                                                                         \
    __global__ void BOUNDS NAME(FbJob J, const int32_t *xt, const uint8_t *pcm, const FbAnalysis *ana, FbKfPlan *plan,           \
                                fb200_subframe_info *vsubs, fb200_subframe_info *psubs, uint32_t *poffs, uint32_t *frame_bytes,  \
                                fb200_frame_info *infos, uint32_t *fb_list, uint32_t *fb_count, const uint32_t *ktab,            \
                                FbKfLayout L, int odd_mode) {                                                                    \
        extern __shared__ __align__(16) uint8_t fb_smem[];                                                                       \
        const bool odd = odd_mode == 2 || (odd_mode == 1 && blockIdx.x == 0);                                                    \
        const uint32_t f = odd_mode == 1 ? (blockIdx.x == 0 ? J.n_frames - 1u : blockIdx.x - 1u) : blockIdx.x;                   \
        if (odd)                                                                                                                 \
            fb_ka_body<FB_INST_G, true, VMS, BC>(J, xt, pcm, ana, plan, vsubs, psubs, poffs, frame_bytes, infos, fb_list, fb_count,  \
                                             ktab, f, fb_smem, L);                                                               \
        else                                                                                                                     \
            fb_ka_body<FB_INST_G, false, VMS, BC>(J, xt, pcm, ana, plan, vsubs, psubs, poffs, frame_bytes, infos, fb_list, fb_count, \
                                              ktab, f, fb_smem, L);                                                              \
    }
FB_KA_KERNEL(FB_NAME(fb_ka_plan_g), FB_KA_BOUNDS, 0, false)
// 16-bit stereo straight from the packed PCM: one plane of (left, right) pairs
FB_KA_KERNEL(FB_NAME(fb_ka_planp_g), FB_KAP_BOUNDS, FB_VM_PAIRS, false)
// OrderSel::BitCount: the instances that search every fixed order (src/coding.rs:241-262)
FB_KA_KERNEL(FB_NAME(fb_ka_planb_g), FB_KA_BOUNDS, 0, true)
FB_KA_KERNEL(FB_NAME(fb_ka_planpb_g), FB_KAP_BOUNDS, FB_VM_PAIRS, true)

__global__ void FB_KP_BOUNDS FB_NAME(fb_kp_pack_g)(FbJob J, const int32_t *xt, const uint8_t *pcm, const FbKfPlan *plan,
                                                             const fb200_subframe_info *psubs, const uint32_t *poffs,
                                                             const unsigned long long *offsets, uint8_t *out,
                                                             unsigned long long out_cap, const uint32_t *ktab, FbKfLayout L,
                                                             int odd_mode) {
    extern __shared__ __align__(16) uint8_t fb_smem[];
    const bool odd = odd_mode == 2 || (odd_mode == 1 && blockIdx.x == 0);
    const uint32_t f = odd_mode == 1 ? (blockIdx.x == 0 ? J.n_frames - 1u : blockIdx.x - 1u) : blockIdx.x;
    if (odd)
        fb_kp_body<FB_INST_G, true>(J, xt, pcm, plan, psubs, poffs, offsets, out, out_cap, ktab, f, fb_smem, L);
    else
        fb_kp_body<FB_INST_G, false>(J, xt, pcm, plan, psubs, poffs, offsets, out, out_cap, ktab, f, fb_smem, L);
}

// pcm_pairs != nullptr: the rows come straight from the packed 16-bit stereo PCM (fb_pairs_format), xv is not read
void FB_NAME(fb_launch_k1_g)(const FbJob &J, const int32_t *xv, const uint8_t *pcm_pairs, const float *win_full,
                             const float *win_tail, FbAnalysis *ana, fb200_variant_taps *taps, uint32_t nvars, cudaStream_t st) {
    const unsigned grid = 2u * ((fb_k1_slots(J, nvars) + FB_K1_THREADS - 1) / FB_K1_THREADS); // pass A blocks, then pass E blocks
    const uint32_t smem = fb_k1_smem_bytes(J.channels, J.nvar, pcm_pairs != nullptr);
    if (smem > 48u * 1024u)
        cudaFuncSetAttribute(FB_NAME(fb_k1_analyze_g), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    FB_NAME(fb_k1_analyze_g)<<<grid, FB_K1_THREADS, smem, st>>>(J, xv, pcm_pairs, win_full, win_tail, ana, taps, nvars);
}

void FB_NAME(fb_launch_k2_g)(const FbJob &J, const int32_t *xv, const FbAnalysis *ana, fb200_subframe_info *choice,
                             const FbK2Layout &L, const uint32_t *list, const uint32_t *count, uint32_t grid, size_t smem,
                             cudaStream_t st) {
    FB_NAME(fb_k2_rice_g)<<<grid, FB_K2_THREADS, smem, st>>>(J, xv, ana, choice, L, list, count);
}

void FB_NAME(fb_launch_k3_g)(const FbJob &J, const int32_t *xv, const fb200_subframe_info *choice, uint8_t *slots,
                             uint32_t *frame_bytes, fb200_frame_info *infos, const uint32_t *list, const uint32_t *count,
                             uint32_t grid, size_t smem, cudaStream_t st) {
    FB_NAME(fb_k3_pack_g)<<<grid, FB_K3_THREADS, smem, st>>>(J, xv, choice, slots, frame_bytes, infos, list, count);
}

// pcm_pairs != nullptr: L is the pairs layout and the frames are staged from the packed PCM, xt is not read
void FB_NAME(fb_launch_ka_g)(const FbJob &J, const int32_t *xt, const uint8_t *pcm_pairs, const FbAnalysis *ana, void *plan,
                             fb200_subframe_info *vsubs, fb200_subframe_info *psubs,
                             uint32_t *poffs, uint32_t *frame_bytes, fb200_frame_info *infos, uint32_t *fb_list,
                             uint32_t *fb_count, const uint32_t *ktab, const FbKfLayout &L, cudaStream_t st) {
    // the probing instances: OrderSel::BitCount, or the LPC order search extension
    const bool bitcount = (J.cfg.use_fixed && J.cfg.fixed_order_sel == 0) || J.lpc_ext != nullptr;
#define FB_KA_LAUNCH(K, PCM)                                                                                                       \
    FB_NAME(K)<<<J.n_frames, 32 * J.nvar, L.total, st>>>(J, xt, PCM, ana, (FbKfPlan *)plan, vsubs, psubs, poffs, frame_bytes, infos, \
                                                         fb_list, fb_count, ktab, L, fb_kf_odd_mode(J))
    if (pcm_pairs) {
        if (bitcount) FB_KA_LAUNCH(fb_ka_planpb_g, pcm_pairs);
        else FB_KA_LAUNCH(fb_ka_planp_g, pcm_pairs);
    } else {
        if (bitcount) FB_KA_LAUNCH(fb_ka_planb_g, nullptr);
        else FB_KA_LAUNCH(fb_ka_plan_g, nullptr);
    }
#undef FB_KA_LAUNCH
}

void FB_NAME(fb_launch_kp_g)(const FbJob &J, const int32_t *xt, const uint8_t *pcm, const void *plan, const fb200_subframe_info *psubs,
                             const uint32_t *poffs, const unsigned long long *offsets, uint8_t *out,
                             unsigned long long out_cap, const uint32_t *ktab, const FbKfLayout &L, cudaStream_t st) {
    FB_NAME(fb_kp_pack_g)<<<J.n_frames, 32 * J.nvar, L.total, st>>>(J, xt, pcm, (const FbKfPlan *)plan, psubs, poffs, offsets, out,
                                                                     out_cap, ktab, L, fb_kf_odd_mode(J));
}

cudaError_t FB_NAME(fb_set_smem_g)(int kernel, int bytes) {
    switch (kernel) {
    case FB_KERNEL_K2:
        return cudaFuncSetAttribute(FB_NAME(fb_k2_rice_g), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    case FB_KERNEL_K3:
        return cudaFuncSetAttribute(FB_NAME(fb_k3_pack_g), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    case FB_KERNEL_KF: {
        cudaError_t e = cudaFuncSetAttribute(FB_NAME(fb_ka_plan_g), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(FB_NAME(fb_ka_planp_g), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(FB_NAME(fb_ka_planb_g), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(FB_NAME(fb_ka_planpb_g), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(FB_NAME(fb_kp_pack_g), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    }
    default:
        return cudaErrorInvalidValue;
    }
}
