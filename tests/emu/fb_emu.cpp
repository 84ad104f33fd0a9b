// fb_emu.cpp -- CPU emulation of the CUDA kernel bodies, for logic tests on machines without a GPU.
//
// TEST INFRASTRUCTURE.  It compiles flacenc_rs_b200/csrc/fb_kernels.cuh with FB_EMULATE, which turns
// every barrier-delimited phase of a kernel into a loop over the CTA's threads, and drives the same
// K0..K4 sequence as fb_api.cu.  It lets `pytest -m "not gpu"` compare the kernels' logic with the
// oracle byte for byte.  It is not part of libflacenc_b200.so and is not a fallback.
#define FB_EMULATE 1
#include "../../flacenc_rs_b200/csrc/fb_host.h"

#include <stdlib.h>
#include <vector>

#include <thread>
#include <vector>

extern "C" {

float fbemu_log2f(float x) { return fb_log2f(x); }

// fb_log2f (the device code) against the host libm's log2f for the bit patterns [first, first + count), on
// `threads` host threads; returns the number of mismatches (NaN == NaN) and the first mismatching pattern
unsigned long long fbemu_log2f_sweep(uint32_t first, unsigned long long count, int threads, uint32_t *first_bad) {
    if (threads < 1) threads = 1;
    std::vector<unsigned long long> bad((size_t)threads, 0);
    std::vector<uint32_t> fb((size_t)threads, 0xFFFFFFFFu);
    std::vector<std::thread> pool;
    for (int w = 0; w < threads; w++)
        pool.emplace_back([&, w] {
            const unsigned long long a = count * (unsigned long long)w / (unsigned long long)threads;
            const unsigned long long b = count * (unsigned long long)(w + 1) / (unsigned long long)threads;
            for (unsigned long long i = a; i < b; i++) {
                const uint32_t u = first + (uint32_t)i;
                const float x = fb_u2f(u), y = fb_log2f(x), z = log2f(x);
                if (fb_f2u(y) != fb_f2u(z) && !(y != y && z != z)) {
                    if (bad[(size_t)w]++ == 0) fb[(size_t)w] = u;
                }
            }
        });
    for (auto &t : pool) t.join();
    unsigned long long total = 0;
    uint32_t f = 0xFFFFFFFFu;
    for (int w = 0; w < threads; w++) { total += bad[(size_t)w]; if (fb[(size_t)w] < f) f = fb[(size_t)w]; }
    if (first_bad) *first_bad = f;
    return total;
}

// bits of the device-side IRLS weight (fb_irls_weight over fb_powf_pos) for raw errors with bit patterns [first, first + count)
void fbemu_irls_weight_bits(uint32_t first, unsigned long long count, float normalizer, int threads, uint32_t *out) {
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    for (int w = 0; w < threads; w++)
        pool.emplace_back([&, w] {
            const unsigned long long a = count * (unsigned long long)w / (unsigned long long)threads;
            const unsigned long long b = count * (unsigned long long)(w + 1) / (unsigned long long)threads;
            for (unsigned long long i = a; i < b; i++) out[i] = fb_f2u(fb_irls_weight(fb_u2f(first + (uint32_t)i), normalizer));
        });
    for (auto &t : pool) t.join();
}

int fbemu_find_shift(const double *coefs, int n, int precision) { return fb_find_shift(coefs, n, precision); }

// Rice-search statistics: [0] narrow-window runs, [1] full-range reruns, [2] chunk-exact runs, [3] sum of widths
void fbemu_mode_counts(unsigned long long *out4, int reset) {
    for (int i = 0; i < 4; i++) { out4[i] = fb_emu_mode_count[i]; if (reset) fb_emu_mode_count[i] = 0; }
}

// 0 = fused kernel when eligible (like the library), 1 = generic K2/K3 kernels only
static int fb_emu_force_generic = 0;
// 1 = the pack kernel stages 16-bit stereo PCM as (left, right) pairs when the format allows (like the library)
static int fb_emu_kp_pairs = 1;
static unsigned long long fb_emu_fused_frames = 0, fb_emu_fallback_frames = 0;
void fbemu_set_force_generic(int v) { fb_emu_force_generic = v; }
void fbemu_set_kp_pairs(int v) { fb_emu_kp_pairs = v; }
void fbemu_fused_counts(unsigned long long *out2, int reset) {
    out2[0] = fb_emu_fused_frames; out2[1] = fb_emu_fallback_frames;
    if (reset) fb_emu_fused_frames = fb_emu_fallback_frames = 0;
}

void fbemu_config_default(fb200_config *c) { fbh_config_default(c); }
int fbemu_config_verify(const fb200_config *c) { return fbh_config_verify(c); }

// launch geometry helpers shared by the host launchers and the kernels:
// out4 = {lane slots of the analysis launch, variants before an isolated last frame, odd mode of the fused kernels,
//         rows staged per quad by an analysis warp}
int fbemu_launch_geometry(int channels, int bps, int rate, int block_size, uint64_t n_samples, uint32_t *out4) {
    fb200_config cfg;
    fbh_config_default(&cfg);
    const FbJob J = fbh_make_job(cfg, channels, bps, rate, block_size, 4, n_samples, 0);
    const uint32_t nvars = J.n_frames * (uint32_t)J.nvar;
    out4[0] = fb_k1_slots(J, nvars);
    out4[1] = fb_k1_full_variants(J, nvars);
    out4[2] = (uint32_t)fb_kf_odd_mode(J);
    out4[3] = (uint32_t)fb_k1_pitch_rows(J.channels, J.nvar);
    return (int)J.n_frames;
}

int fbemu_crc8(const uint8_t *d, int len) { return fb_crc8(d, len); }

int fbemu_frame_header(int n, int ch_tag, int bps, int rate, uint32_t number, uint8_t *out) {
    return fb_frame_header(n, ch_tag, bps, rate, number, out);
}

struct EmuBuffers {
    std::vector<int32_t> xv, xv4;
    std::vector<float> win_full, win_tail;
    std::vector<FbAnalysis> ana;
    std::vector<FbLpcExt> lpc_ext;
    std::vector<fb200_variant_taps> taps;
    std::vector<fb200_subframe_info> choice;
    std::vector<uint8_t> slots;
    std::vector<uint32_t> frame_bytes;
    std::vector<unsigned long long> offsets;
    std::vector<fb200_frame_info> infos;
    std::vector<uint8_t> stream; // contiguous frame bytes (offsets[n_frames] of them)
};

static int emu_run(const fb200_config *cfg, const void *pcm, const int32_t *planar, int planar_stride,
                   int container_bytes, uint64_t n_samples, int channels, int bps, int rate, int block_size,
                   uint32_t first_frame, bool analyze_only, EmuBuffers &B, FbJob &Jout) {
    int rc = fbh_config_verify(cfg);
    if (rc) return rc;
    rc = fbh_format_verify(channels, bps, rate, block_size);
    if (rc) return rc;
    if (!planar && (container_bytes < 1 || container_bytes > 4 || container_bytes * 8 < bps)) return FB200_ERR_SOURCE;
    FbJob J = fbh_make_job(*cfg, channels, bps, rate, block_size, container_bytes, n_samples, first_frame);
    Jout = J;
    if (J.n_frames == 0) return FB200_OK;
    if ((uint64_t)first_frame + J.n_frames > (1ull << 31)) return FB200_ERR_CONFIG;
    const size_t nvars = (size_t)J.n_frames * J.nvar;
    B.xv.assign(fb_xt_words(J.stride, (uint64_t)J.n_frames * J.channels) + 64, 0x55555555); // poison: padding must never influence results
    B.win_full.assign(J.block_size + 64, 0.f);
    B.win_tail.assign(J.tail_n + 64, 0.f);
    fbh_window_weights(cfg->window_type, cfg->tukey_alpha, J.block_size, B.win_full.data());
    fbh_window_weights(cfg->window_type, cfg->tukey_alpha, J.tail_n, B.win_tail.data());
    B.ana.resize(nvars);
    if (cfg->ext_lpc_order_search > 0 || cfg->ext_lpc_precision_search > 0) {
        FbLpcExt poison;
        memset(&poison, 0x5A, sizeof(poison));
        B.lpc_ext.assign(nvars * FB_EXT_LPC_MAX, poison);
        J.lpc_ext = B.lpc_ext.data();
        Jout = J;
    }
    B.taps.resize(nvars);
    B.choice.resize(nvars);
    uint32_t err_flag = 0;

    // K0
    if (planar) {
        for (int t4 = 0; t4 < J.stride / 4; t4++) fb_k0_planar_quad(J, planar, planar_stride, B.xv.data(), &err_flag, t4);
    } else {
        const uint64_t n_items = (uint64_t)((J.n_frames + 15u) / 16u) * (uint64_t)((J.stride / 4 + 1) / 2) * 32u;
        for (uint64_t idx = 0; idx < n_items; idx++) {
            uint32_t f;
            int t4;
            fb_k0_item(idx, J.stride / 4, &f, &t4);
            if (f < J.n_frames && t4 < J.stride / 4) fb_k0_quad_any(J, (const uint8_t *)pcm, B.xv.data(), &err_flag, f, t4);
        }
    }
    if (err_flag) return FB200_ERR_CONFIG;
    // K1
    for (uint32_t gv = 0; gv < nvars; gv++)
        fb_k1_dispatch(J, B.xv.data(), B.win_full.data(), B.win_tail.data(), B.ana.data(), B.taps.data(), gv);
    // K1C: direct-MSE estimator (tiles of 256 samples here, so that frames span several tiles)
    if (cfg->use_direct_mse && cfg->use_lpc && cfg->mae_optimization_steps > 0) {
        // K1I: IRLS-MAE refinement (same small tiles)
        const int tile = 256;
        std::vector<uint8_t> smem(fb_k1i_smem_bytes(J.cfg.lpc_order, tile) + 64);
        for (uint32_t gv = 0; gv < nvars; gv++) {
            memset(smem.data(), 0xAB, smem.size());
            fb_k1i_body(J, B.xv.data(), nullptr, B.win_full.data(), B.win_tail.data(), B.ana.data(), B.taps.data(), gv, tile,
                        smem.data());
        }
    } else if (cfg->use_direct_mse && cfg->use_lpc) {
        const int tile = 256;
        std::vector<uint8_t> smem(fb_k1c_smem_bytes(J.cfg.lpc_order, tile) + 64);
        for (uint32_t gv = 0; gv < nvars; gv++) {
            memset(smem.data(), 0xAB, smem.size());
            fb_k1c_body(J, B.xv.data(), nullptr, B.win_full.data(), B.win_tail.data(), B.ana.data(), B.taps.data(), gv, tile,
                        smem.data());
        }
    }
    if (analyze_only) return FB200_OK;
    B.slots.assign((size_t)J.n_frames * J.slot_bytes, 0xCD);
    B.frame_bytes.assign(J.n_frames, 0);
    B.infos.resize(J.n_frames);
    // frames for the generic kernels: all of them, or the fused kernel's fallback list
    std::vector<uint32_t> todo;
    FbKfLayout KL;
    const bool fused = !fb_emu_force_generic && fbh_fused_ok(J, J.tail_n, &KL);
    // pairs mode like the library: 16-bit stereo PCM is staged as (left, right) pairs by the plan and pack kernels
    const bool ka_pairs = fused && !planar && fb_emu_kp_pairs && fb_pairs_format(J.channels, J.bps, container_bytes, J.block_size) &&
                          ((uintptr_t)pcm & 15u) == 0;
    if (ka_pairs) KL = fb_kf_layout(J.channels, J.nvar, J.bps, J.block_size, J.tail_n, false, true);
    const uint8_t *pcm8 = (const uint8_t *)pcm;
    std::vector<FbKfPlan> plan;
    std::vector<fb200_subframe_info> psubs;
    std::vector<uint32_t> poffs;
    if (fused) {
        std::vector<uint32_t> list(J.n_frames + 1, 0);
        uint32_t count = 0;
        std::vector<uint8_t> smem(KL.total + 64);
        std::vector<uint32_t> ktab_a(fb_kf_ktab_words(KL.crc_chunk, 32u * (uint32_t)J.nvar));
        fb_kf_build_ktab(KL.crc_chunk, 32u * (uint32_t)J.nvar, ktab_a.data());
        plan.resize(J.n_frames);
        memset(plan.data(), 0xEE, plan.size() * sizeof(FbKfPlan));
        psubs.resize((size_t)J.n_frames * J.channels);
        poffs.assign((size_t)J.n_frames * J.channels * (KL.U_max + 1), 0xDDDDDDDDu);
        for (uint32_t f = 0; f < J.n_frames; f++) {
            memset(smem.data(), 0xAB, smem.size());
#define EMU_KA_ARGS B.ana.data(), plan.data(), B.choice.data(), psubs.data(), poffs.data(), B.frame_bytes.data(), B.infos.data(), list.data(), &count, ktab_a.data(), f, smem.data(), KL
#define EMU_KA(GG) do { const bool odd__ = fb_kf_geom(fb_frame_len(J, f)).leaf_len & 3; \
            if (ka_pairs) { if (odd__) fb_ka_body<GG, true, FB_VM_PAIRS, true>(J, nullptr, pcm8, EMU_KA_ARGS); else fb_ka_body<GG, false, FB_VM_PAIRS, true>(J, nullptr, pcm8, EMU_KA_ARGS); } \
            else { if (odd__) fb_ka_body<GG, true, 0, true>(J, B.xv.data(), nullptr, EMU_KA_ARGS); else fb_ka_body<GG, false, 0, true>(J, B.xv.data(), nullptr, EMU_KA_ARGS); } } while (0)
            switch (fb_k1_ring(J.cfg.lpc_order)) {
            case 4: EMU_KA(4); break;
            case 8: EMU_KA(8); break;
            case 12: EMU_KA(12); break;
            case 16: EMU_KA(16); break;
            case 20: EMU_KA(20); break;
            default: EMU_KA(24); break;
            }
#undef EMU_KA
        }
        todo.assign(list.begin(), list.begin() + count);
        fb_emu_fused_frames += J.n_frames - count;
        fb_emu_fallback_frames += count;
    } else {
        for (uint32_t f = 0; f < J.n_frames; f++) todo.push_back(f);
    }
    // K0b: rows by variant for the generic kernels
    B.xv4.assign(nvars * J.stride + 64, 0x55555555);
    for (uint32_t f : todo)
        for (int t4 = 0; t4 < J.stride / 4; t4++) fb_k0b_expand4(J, B.xv.data(), B.xv4.data(), f, t4);
    const int32_t *xg = B.xv4.data();
    // K2
    {
        int nmax = J.block_size;
        FbK2Layout L = fb_k2_layout(nmax, fbh_leaves_max(J));
        if (L.total + 3 * sizeof(FbRiceResult) > 227u * 1024u) L = fb_k2_layout(nmax, fbh_leaves_max(J), 31); // (like the library)
        std::vector<uint8_t> smem(L.total + 3 * sizeof(FbRiceResult) + 64);
        for (uint32_t f : todo)
          for (uint32_t gv = f * J.nvar; gv < (f + 1) * (uint32_t)J.nvar; gv++) {
            memset(smem.data(), 0xAB, smem.size());
            switch (fb_k1_ring(J.cfg.lpc_order)) {
            case 4: fb_k2_body<4>(J, xg, B.ana.data(), B.choice.data(), gv, smem.data(), L); break;
            case 8: fb_k2_body<8>(J, xg, B.ana.data(), B.choice.data(), gv, smem.data(), L); break;
            case 12: fb_k2_body<12>(J, xg, B.ana.data(), B.choice.data(), gv, smem.data(), L); break;
            case 16: fb_k2_body<16>(J, xg, B.ana.data(), B.choice.data(), gv, smem.data(), L); break;
            case 20: fb_k2_body<20>(J, xg, B.ana.data(), B.choice.data(), gv, smem.data(), L); break;
            default: fb_k2_body<24>(J, xg, B.ana.data(), B.choice.data(), gv, smem.data(), L); break;
            }
          }
    }
    // K3
    {
        uint32_t mb = fb_max_frame_bytes(J.channels, J.bps, J.block_size);
        std::vector<uint8_t> smem(fb_k3_smem_bytes(mb, J.block_size, J.pack_in_smem) + 64);
        for (uint32_t f : todo) {
            memset(smem.data(), 0xEF, smem.size());
#define EMU_K3(GG) fb_k3_body<GG>(J, xg, B.choice.data(), B.slots.data(), B.frame_bytes.data(), B.infos.data(), f, smem.data())
            switch (fb_k1_ring(J.cfg.lpc_order)) {
            case 4: EMU_K3(4); break;
            case 8: EMU_K3(8); break;
            case 12: EMU_K3(12); break;
            case 16: EMU_K3(16); break;
            case 20: EMU_K3(20); break;
            default: EMU_K3(24); break;
            }
#undef EMU_K3
        }
    }
    // K4 scan
    B.offsets.assign(J.n_frames + 1, 0);
    {
        std::vector<unsigned long long> partials(FB_K4_THREADS);
        fb_k4_scan_body(B.frame_bytes.data(), B.offsets.data(), J.n_frames, partials.data());
    }
    // the output stream: KP packs the planned frames in place; the others are gathered from their slots
    const unsigned long long total = B.offsets[J.n_frames];
    B.stream.assign((size_t)total + 16, 0x77);
    if (fused) {
        const bool pairs = !planar && fb_emu_kp_pairs && fb_kp_pairs_format(J.channels, J.bps, container_bytes) &&
                           ((uintptr_t)pcm & 15u) == 0 && (J.block_size & 3) == 0;
        const FbKfLayout KPL = fb_kp_layout(J.channels, J.nvar, J.bps, J.block_size, J.tail_n, pairs);
        std::vector<uint8_t> smem(KPL.total + 64);
        std::vector<uint32_t> ktab(fb_kf_ktab_words(KL.crc_chunk, 32u * (uint32_t)J.nvar));
        fb_kf_build_ktab(KL.crc_chunk, 32u * (uint32_t)J.nvar, ktab.data());
        for (uint32_t f = 0; f < J.n_frames; f++) {
            memset(smem.data(), 0xAB, smem.size());
#define EMU_KP(GG) if (fb_kf_geom(fb_frame_len(J, f)).leaf_len & 3) fb_kp_body<GG, true>(J, B.xv.data(), pairs ? (const uint8_t *)pcm : nullptr, plan.data(), psubs.data(), poffs.data(), B.offsets.data(), B.stream.data(), total, ktab.data(), f, smem.data(), KPL); else fb_kp_body<GG, false>(J, B.xv.data(), pairs ? (const uint8_t *)pcm : nullptr, plan.data(), psubs.data(), poffs.data(), B.offsets.data(), B.stream.data(), total, ktab.data(), f, smem.data(), KPL)
            switch (fb_k1_ring(J.cfg.lpc_order)) {
            case 4: EMU_KP(4); break;
            case 8: EMU_KP(8); break;
            case 12: EMU_KP(12); break;
            case 16: EMU_KP(16); break;
            case 20: EMU_KP(20); break;
            default: EMU_KP(24); break;
            }
#undef EMU_KP
        }
    }
    for (uint32_t f : todo)
        for (int tid = 0; tid < 256; tid++)
            fb_k4_gather_thread(B.slots.data(), J.slot_bytes, B.frame_bytes.data(), B.offsets.data(), B.stream.data(), total, f,
                                tid, 256);
    return FB200_OK;
}

int fbemu_encode_interleaved(const fb200_config *cfg, const void *pcm, int container_bytes, uint64_t n_samples,
                             int channels, int bps, int rate, int block_size, uint32_t first_frame, uint8_t *out,
                             size_t out_cap, uint32_t *frame_sizes, fb200_frame_info *infos, size_t *n_frames,
                             size_t *out_len, int force_global_pack) {
    EmuBuffers B;
    FbJob J;
    (void)force_global_pack;
    int rc = emu_run(cfg, pcm, nullptr, 0, container_bytes, n_samples, channels, bps, rate, block_size, first_frame,
                     false, B, J);
    if (rc) return rc;
    if (n_frames) *n_frames = J.n_frames;
    if (J.n_frames == 0) {
        if (out_len) *out_len = 0;
        return FB200_OK;
    }
    unsigned long long total = B.offsets[J.n_frames];
    if (out_len) *out_len = (size_t)total;
    if (total > out_cap) return FB200_ERR_CAPACITY;
    memcpy(out, B.stream.data(), (size_t)total);
    if (frame_sizes) memcpy(frame_sizes, B.frame_bytes.data(), sizeof(uint32_t) * J.n_frames);
    if (infos) memcpy(infos, B.infos.data(), sizeof(fb200_frame_info) * J.n_frames);
    return FB200_OK;
}

int fbemu_encode_planar_frame(const fb200_config *cfg, const int32_t *planar, int stride, int n, int channels, int bps,
                              int rate, uint32_t frame_number, uint8_t *out, size_t out_cap, size_t *out_len,
                              fb200_frame_info *info) {
    EmuBuffers B;
    FbJob J;
    if (n < 1 || n > 32767) return FB200_ERR_SOURCE;
    int block_size = n < 32 ? 32 : n; // FrameBuf sizes are >= 32; a short frame is a tail of a 32-block
    int rc = emu_run(cfg, nullptr, planar, stride, 4, (uint64_t)n, channels, bps, rate, block_size, frame_number,
                     false, B, J);
    if (rc) return rc;
    unsigned long long total = B.offsets[1];
    if (out_len) *out_len = (size_t)total;
    if (total > out_cap) return FB200_ERR_CAPACITY;
    memcpy(out, B.stream.data(), (size_t)total);
    if (info) *info = B.infos[0];
    return FB200_OK;
}

int fbemu_analyze(const fb200_config *cfg, const void *pcm, int container_bytes, uint64_t n_samples, int channels,
                  int bps, int rate, int block_size, fb200_variant_taps *taps, size_t taps_cap, size_t *n_variants) {
    EmuBuffers B;
    FbJob J;
    int rc = emu_run(cfg, pcm, nullptr, 0, container_bytes, n_samples, channels, bps, rate, block_size, 0, true, B, J);
    if (rc) return rc;
    size_t nv = (size_t)J.n_frames * J.nvar;
    if (n_variants) *n_variants = nv;
    if (nv > taps_cap) return FB200_ERR_CAPACITY;
    if (nv) memcpy(taps, B.taps.data(), nv * sizeof(fb200_variant_taps));
    return FB200_OK;
}

} // extern "C"
