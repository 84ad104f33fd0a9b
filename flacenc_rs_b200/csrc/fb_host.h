// fb_host.h -- host-side logic shared by the CUDA library (fb_api.cu) and the CPU emulation of the
// kernels used by the logic tests: config defaults / verification, job geometry, window table.
// Reference citations are relative to /root/reference/.
#pragma once

#include <math.h>
#include <vector>

#include "fb_fused.cuh"

// config::Encoder::default() (src/config.rs:97-107,143-151,180-191,218-222,257-264,287-297,352-358,411-417)
inline void fbh_config_default(fb200_config *c) {
    memset(c, 0, sizeof(*c));
    c->block_size = 4096;
    c->multithread = 1;
    c->workers = 0;
    c->use_leftside = c->use_rightside = c->use_midside = 1;
    c->use_constant = c->use_fixed = c->use_lpc = 1;
    c->fixed_max_order = 4;
    c->fixed_order_sel = 1;
    c->approx_ent_partitions = 16;
    c->lpc_order = 10;
    c->quant_precision = 15;
    c->use_direct_mse = 0;
    c->mae_optimization_steps = 0;
    c->window_type = 1;
    c->tukey_alpha = 0.4f;
    c->prc_max_parameter = 30;
    c->ext_lpc_order_search = 0;
    c->ext_lpc_precision_search = 0;
}

// Verify for Encoder and children (src/config.rs:109-130,198-204,224-229,299-326,371-387).
// The reference never reaches Fixed::verify / OrderSel::verify (SubFrameCoding::verify skips
// `fixed`), so max_order > 4 is accepted there and clamped by `.take(max_order + 1)`; we accept it
// too.  `partitions` outside 1..=64 would divide by zero / overrun in the reference; it is rejected
// here (documented deviation), as are negative orders and unknown selector/window tags, which the
// Rust type system makes unrepresentable.
inline int fbh_config_verify(const fb200_config *c) {
    if (c->block_size < 32 || c->block_size > 32767) return FB200_ERR_CONFIG;
    if (c->lpc_order < 1 || c->lpc_order > FB200_MAX_LPC_ORDER) return FB200_ERR_CONFIG;
    if (c->quant_precision < 1 || c->quant_precision > 15) return FB200_ERR_CONFIG;
    // use_direct_mse and mae_optimization_steps are accepted like a reference built with the `experimental` feature
    // (src/config.rs:305-321; the covariance-method estimator of src/lpc.rs:852-913 and its IRLS refinement :814-850,
    // which src/coding.rs:337-345 runs only when use_direct_mse is set)
    if (c->use_direct_mse != 0 && c->use_direct_mse != 1) return FB200_ERR_CONFIG;
    if (c->mae_optimization_steps < 0) return FB200_ERR_CONFIG;
    if (c->window_type == 1) {
        if (!(c->tukey_alpha >= 0.0f && c->tukey_alpha <= 1.0f)) return FB200_ERR_CONFIG;
    } else if (c->window_type != 0) {
        return FB200_ERR_CONFIG;
    }
    if (c->prc_max_parameter < 0 || c->prc_max_parameter > 30) return FB200_ERR_CONFIG;
    // extension (not in the reference): lower LPC orders from the same autocorrelation
    if (c->ext_lpc_order_search < 0 || c->ext_lpc_order_search > FB_EXT_LPC_MAX) return FB200_ERR_CONFIG;
    if (c->ext_lpc_precision_search < 0 || c->ext_lpc_precision_search > 4) return FB200_ERR_CONFIG;
    if (c->ext_lpc_order_search + c->ext_lpc_precision_search > FB_EXT_LPC_MAX) return FB200_ERR_CONFIG;
    if ((c->ext_lpc_order_search > 0 || c->ext_lpc_precision_search > 0) && c->use_direct_mse) return FB200_ERR_CONFIG;
    if (c->fixed_max_order < 0) return FB200_ERR_CONFIG;
    if (c->fixed_order_sel != 0 && c->fixed_order_sel != 1) return FB200_ERR_CONFIG;
    if (c->fixed_order_sel == 1 && (c->approx_ent_partitions < 1 || c->approx_ent_partitions > FB_MAX_ENT_PARTS))
        return FB200_ERR_CONFIG;
    return FB200_OK;
}

// Stream format checks of StreamInfo::new / FrameBuf::with_size (src/constant.rs:38-60,
// src/component/verify.rs:51-66,149-151, src/source.rs:167-173): bits per sample 8..=25 and
// a multiple of 4 (or 4n+1), sample rate <= 96000, channels 1..=8, block 32..=32767.
inline int fbh_format_verify(int channels, int bps, int sample_rate, int block_size) {
    if (channels < 1 || channels > FB200_MAX_CHANNELS) return FB200_ERR_CONFIG;
    if (bps < 8 || bps > 25 || !(bps % 4 == 0 || bps % 4 == 1)) return FB200_ERR_CONFIG;
    if (sample_rate < 0 || sample_rate > 96000) return FB200_ERR_CONFIG;
    if (block_size < 32 || block_size > 32767) return FB200_ERR_CONFIG;
    return FB200_OK;
}

// window_weights (src/lpc.rs:96-120): f32 arithmetic, cosf from the host libm.  The table is
// computed once per (length, window) on the host -- like the reference's per-thread WINDOW_CACHE
// (src/lpc.rs:217-231) -- and uploaded; the kernels only multiply by it.
inline void fbh_window_weights(int window_type, float alpha, int len, float *out) {
    if (window_type == 0 || alpha == 0.0f) {
        for (int t = 0; t < len; t++) out[t] = 1.0f;
        return;
    }
    const float pi = 3.14159265358979323846f;
    float max_t = (float)len - 1.0f;
    float alpha_len = alpha * max_t;
    for (int i = 0; i < len; i++) {
        float t = (float)i;
        float w;
        if (t < alpha_len / 2.0f) {
            w = 0.5f * (1.0f - cosf(2.0f * pi * t / alpha_len));
        } else if (t < max_t - alpha_len / 2.0f) {
            w = 1.0f;
        } else {
            w = 0.5f * (1.0f - cosf(2.0f * pi * (max_t - t) / alpha_len));
        }
        out[i] = w;
    }
}

// Geometry of one batch of whole frames (the last may be short).
inline FbJob fbh_make_job(const fb200_config &cfg, int channels, int bps, int sample_rate, int block_size,
                          int container_bytes, uint64_t n_samples, uint32_t first_frame_number) {
    FbJob J;
    memset(&J, 0, sizeof(J));
    J.cfg = cfg;
    J.channels = channels;
    J.bps = bps;
    J.sample_rate = sample_rate;
    J.block_size = block_size;
    J.nvar = channels == 2 ? 4 : channels;
    J.stride = (block_size + 31) & ~31;
    J.container_bytes = container_bytes;
    J.n_samples = n_samples;
    J.n_frames = (uint32_t)((n_samples + (uint64_t)block_size - 1) / (uint64_t)block_size);
    uint64_t rem = n_samples % (uint64_t)block_size;
    J.tail_n = rem ? (int)rem : block_size;
    J.first_frame_number = first_frame_number;
    uint32_t mb = fb_max_frame_bytes(channels, bps, block_size);
    J.slot_bytes = (mb + 15u) & ~15u;
    J.pack_in_smem = (fb_k3_smem_bytes(mb, block_size, 1) <= 200u * 1024u) ? 1 : 0;
    return J;
}

inline int fbh_leaves_max(const FbJob &J) {
    int a = 1 << fb_finest_partition_order(J.block_size);
    int b = 1 << fb_finest_partition_order(J.tail_n);
    return a > b ? a : b;
}

// The fused per-frame kernels (fb_fused.cuh) serve a batch when the frame's working set fits in shared memory (both
// order selectors: ApproxEnt takes the fixed order from K1's entropy estimate, BitCount searches every order).
inline bool fbh_fused_ok(const FbJob &J, int tail_n_call, FbKfLayout *Lout, bool x16 = false) {
    FbKfLayout L = fb_kf_layout(J.channels, J.nvar, J.bps, J.block_size, tail_n_call, x16);
    if (Lout) *Lout = L;
    if (L.total > FB_KF_SMEM_LIMIT) return false;
    // the pack kernel stages channels in groups; it needs room for at least one plane (both planes for stereo)
    const FbKfLayout LP = fb_kp_layout(J.channels, J.nvar, J.bps, J.block_size, tail_n_call);
    return LP.total <= FB_KF_SMEM_LIMIT;
}
