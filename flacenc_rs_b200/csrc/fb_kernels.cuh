// fb_kernels.cuh -- kernel bodies of the B200 FLAC frame-encode pipeline.
//
//   K0 ingest   : interleaved packed PCM -> row-interleaved planar int32 channel rows, range check
//                                                               [thread per (frame, 4 samples)]
//   K1 analyze  : constant detection, fixed-predictor entropy estimate, windowed autocorrelation
//                 (sequential f64 FMA, bit-identical to the scalar reference), Levinson-Durbin,
//                 qlp quantisation                              [thread per channel variant]
//   fused path  : fb_fused.cuh (plan kernel KA, pack kernel KP) -- the default for eligible batches
//   generic path, for the frames KA hands back and for batches that are not eligible:
//   K0b expand  : plain rows per variant (with M and S)
//   K2 rice     : residual of the fixed winner and of the LPC candidate, partitioned Rice search by literal
//                 table replay, exact bit counts, per-variant subframe decision [CTA per channel variant]
//   K3 pack     : stereo decision, frame header + CRC-8, bit packing from a prefix scan of code
//                 lengths, CRC-16, into a per-frame slot          [CTA per frame]
//   K4          : exclusive scan of frame sizes; gather of slots into the contiguous byte stream
//
// Every body is written as barrier-delimited phases (FB_PHASE ... FB_PHASE_END) so the same source
// also runs under the CPU emulation used by the logic tests (see fb_common.h).
// Reference citations are relative to /root/reference/.
#pragma once

#include "fb_common.h"

#if !FB_GPU
static unsigned long long fb_emu_mode_count[4] = {0, 0, 0, 0};
#endif

#define FB_K2_THREADS 128
#define FB_K3_THREADS 256
#define FB_RUN 16 // samples packed sequentially by one thread in K3

// =================================================================================================
// K0: ingest.  Mirrors Fill::fill_le_bytes / fill_interleaved + deinterleave
// (src/source.rs:278-299, src/arrayutils.rs:248-264,345-364), FrameBuf::verify_samples
// (src/source.rs:262-275) and the M/S synthesis of try_stereo_coding (src/coding.rs:476-484).
// Output: the row-interleaved planar store xt (below); M and S are formed on the fly by the consumers.
// =================================================================================================
FB_DEV int32_t fb_load_sample(const uint8_t *pcm, uint64_t idx, int container_bytes) {
    if (container_bytes == 1) {
        return (int32_t)(int8_t)pcm[idx];
    } else if (container_bytes == 2) {
        const uint8_t *p = pcm + idx * 2;
        return (int32_t)(int16_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8));
    } else if (container_bytes == 3) {
        const uint8_t *p = pcm + idx * 3;
        uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
        return (int32_t)(v << 8) >> 8;
    } else {
        const uint8_t *p = pcm + idx * 4;
        return (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
    }
}

// ---- planar sample store "xt" ------------------------------------------------------------------------
// Rows are numbered row = frame * channels + channel and stored interleaved in units of 32 rows at a granularity
// of FB_XT_CH samples:  word offset of sample t of row r =
//     (((r / 32) * (stride / CH) + t / CH) * 32 + r % 32) * CH + t % CH.
// The analysis kernel walks 32 rows in lockstep (lane = variant): its warp-wide 16-byte loads of "the next four
// samples of my row" fall into one 32 * CH * 4-byte block that the following CH / 4 - 1 loads reuse, while a frame's
// own samples still come in runs of CH * 4 bytes per row (adjacent rows = the channels of the frame) for the
// per-frame kernel.
#define FB_XT_CH 4
FB_HD size_t fb_xt_off(int stride, uint32_t row, int t) {
    return ((((size_t)(row >> 5) * (size_t)(stride / FB_XT_CH) + (size_t)(t / FB_XT_CH)) << 5) + (size_t)(row & 31u)) * FB_XT_CH +
           (size_t)(t % FB_XT_CH);
}
// offset of the quad at t (multiple of 4) relative to the row base fb_xt_off(stride, row, 0)
FB_HD size_t fb_xt_quad(int t) { return (size_t)(t / FB_XT_CH) * (32u * FB_XT_CH) + (size_t)(t % FB_XT_CH); }
FB_HD size_t fb_xt_words(int stride, uint64_t rows) { return (size_t)((rows + 31u) >> 5) * 32u * (size_t)stride; }

// K0 work item: four consecutive samples (t = 4*t4 ..) of every channel of frame f.  Samples at or beyond the
// frame's length are stored as zeros.  src/source.rs:262-275 range check -> err_flag.
FB_DEV void fb_k0_quad(const FbJob &J, const uint8_t *pcm, int32_t *xt, uint32_t *err_flag, uint32_t f, int t4) {
    const int n = fb_frame_len(J, f);
    const int t = 4 * t4;
    const int32_t lo = -(int32_t)(1u << (J.bps - 1)), hi = (int32_t)((1u << (J.bps - 1)) - 1u);
    const uint64_t s = (uint64_t)f * (uint64_t)J.block_size + (uint64_t)t; // inter-channel sample index in the batch
    bool bad = false;
    if (J.channels == 2 && J.container_bytes == 2 && t + 4 <= n && (((uintptr_t)pcm + s * 4u) & 15u) == 0) {
        // 16-bit stereo: four (L, R) pairs in one 16-byte load
        const int4 w = *reinterpret_cast<const int4 *>(pcm + s * 4u);
        int4 l, r;
        l.x = (int32_t)(int16_t)((uint32_t)w.x & 0xFFFFu); r.x = w.x >> 16;
        l.y = (int32_t)(int16_t)((uint32_t)w.y & 0xFFFFu); r.y = w.y >> 16;
        l.z = (int32_t)(int16_t)((uint32_t)w.z & 0xFFFFu); r.z = w.z >> 16;
        l.w = (int32_t)(int16_t)((uint32_t)w.w & 0xFFFFu); r.w = w.w >> 16;
        const int32_t mn = min(min(min(l.x, l.y), min(l.z, l.w)), min(min(r.x, r.y), min(r.z, r.w)));
        const int32_t mx = max(max(max(l.x, l.y), max(l.z, l.w)), max(max(r.x, r.y), max(r.z, r.w)));
        bad = (mn < lo) | (mx > hi);
        *reinterpret_cast<int4 *>(xt + fb_xt_off(J.stride, f * 2u, t)) = l;
        *reinterpret_cast<int4 *>(xt + fb_xt_off(J.stride, f * 2u + 1u, t)) = r;
    } else {
        for (int c = 0; c < J.channels; c++) {
            int32_t q[4];
            for (int i = 0; i < 4; i++) {
                int32_t v = 0;
                if (t + i < n) {
                    v = fb_load_sample(pcm, (s + (uint64_t)i) * (uint64_t)J.channels + (uint64_t)c, J.container_bytes);
                    bad |= (v < lo) | (v > hi);
                }
                q[i] = v;
            }
            int4 o;
            o.x = q[0]; o.y = q[1]; o.z = q[2]; o.w = q[3];
            *reinterpret_cast<int4 *>(xt + fb_xt_off(J.stride, f * (uint32_t)J.channels + (uint32_t)c, t)) = o;
        }
    }
    if (bad) fb_atomic_or(err_flag, 1u);
}

// K0 fast path for packed 24-bit samples (container 3): the 12 * CH bytes of a quad are read as 3 * CH aligned words
// and the samples cut out with compile-time shifts.  Requires the quad's first byte 4-byte aligned and a whole quad
// inside the frame (fb_k0_p24_ok); everything else takes fb_k0_quad.
template <int CH>
FB_DEV void fb_k0_quad_p24(const FbJob &J, const uint8_t *pcm, int32_t *xt, uint32_t *err_flag, uint32_t f, int t4) {
    const int t = 4 * t4;
    const int32_t lo = -(int32_t)(1u << (J.bps - 1)), hi = (int32_t)((1u << (J.bps - 1)) - 1u);
    const uint64_t s = (uint64_t)f * (uint64_t)J.block_size + (uint64_t)t;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(pcm + s * (uint64_t)(3 * CH));
    uint32_t w[3 * CH + 1];
#pragma unroll
    for (int i = 0; i < 3 * CH; i++) w[i] = src[i];
    w[3 * CH] = 0;
    bool bad = false;
#pragma unroll
    for (int c = 0; c < CH; c++) {
        int32_t q[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int b = (i * CH + c) * 3; // byte offset inside the quad, compile-time
            const int wi = b >> 2, sh = (b & 3) * 8;
            const uint32_t v = sh == 0 ? w[wi] : ((w[wi] >> sh) | (w[wi + 1] << (32 - sh)));
            q[i] = (int32_t)(v << 8) >> 8; // sign-extend 24 bits (src/source.rs:280-286)
            bad |= (q[i] < lo) | (q[i] > hi);
        }
        int4 o;
        o.x = q[0]; o.y = q[1]; o.z = q[2]; o.w = q[3];
        *reinterpret_cast<int4 *>(xt + fb_xt_off(J.stride, f * (uint32_t)CH + (uint32_t)c, t)) = o;
    }
    if (bad) fb_atomic_or(err_flag, 1u);
}

FB_DEV bool fb_k0_p24_ok(const FbJob &J, const uint8_t *pcm, uint32_t f, int t4) {
    if (J.container_bytes != 3 || 4 * t4 + 4 > fb_frame_len(J, f)) return false;
    // the quad's first byte must be 4-byte aligned: a frame starts at f * block_size * 3 * channels bytes, which is
    // not a multiple of 4 for odd frames of e.g. mono block 1001 (a quad itself is 12 * channels bytes long)
    const uint64_t s = (uint64_t)f * (uint64_t)J.block_size + (uint64_t)(4 * t4);
    return (((uintptr_t)pcm + s * (uint64_t)(3 * J.channels)) & 3u) == 0;
}

FB_DEV void fb_k0_quad_any(const FbJob &J, const uint8_t *pcm, int32_t *xt, uint32_t *err_flag, uint32_t f, int t4) {
    if (fb_k0_p24_ok(J, pcm, f, t4)) {
        switch (J.channels) {
        case 1: fb_k0_quad_p24<1>(J, pcm, xt, err_flag, f, t4); return;
        case 2: fb_k0_quad_p24<2>(J, pcm, xt, err_flag, f, t4); return;
        case 3: fb_k0_quad_p24<3>(J, pcm, xt, err_flag, f, t4); return;
        case 4: fb_k0_quad_p24<4>(J, pcm, xt, err_flag, f, t4); return;
        case 5: fb_k0_quad_p24<5>(J, pcm, xt, err_flag, f, t4); return;
        case 6: fb_k0_quad_p24<6>(J, pcm, xt, err_flag, f, t4); return;
        case 7: fb_k0_quad_p24<7>(J, pcm, xt, err_flag, f, t4); return;
        default: fb_k0_quad_p24<8>(J, pcm, xt, err_flag, f, t4); return;
        }
    }
    fb_k0_quad(J, pcm, xt, err_flag, f, t4);
}

// same for a planar FrameBuf input (fb200_encode_planar_frame): src[ch * src_stride + t], one frame (f = 0)
FB_DEV void fb_k0_planar_quad(const FbJob &J, const int32_t *src, int src_stride, int32_t *xt, uint32_t *err_flag, int t4) {
    const int n = J.tail_n, t = 4 * t4;
    const int32_t lo = -(int32_t)(1u << (J.bps - 1)), hi = (int32_t)((1u << (J.bps - 1)) - 1u);
    bool bad = false;
    for (int c = 0; c < J.channels; c++) {
        int32_t q[4];
        for (int i = 0; i < 4; i++) {
            int32_t v = 0;
            if (t + i < n) {
                v = src[(size_t)c * (size_t)src_stride + (size_t)(t + i)];
                bad |= (v < lo) | (v > hi);
            }
            q[i] = v;
        }
        int4 o;
        o.x = q[0]; o.y = q[1]; o.z = q[2]; o.w = q[3];
        *reinterpret_cast<int4 *>(xt + fb_xt_off(J.stride, (uint32_t)c, t)) = o;
    }
    if (bad) fb_atomic_or(err_flag, 1u);
}

// mapping of a K0 thread to its work item: 16 consecutive frames x 2 quads per warp, so both the PCM reads
// (2 adjacent quads = full 32-byte sectors per frame) and the xt writes (16 frames = adjacent rows) use whole sectors
FB_HD void fb_k0_item(uint64_t idx, int quads_per_frame, uint32_t *f, int *t4) {
    const uint64_t warp = idx >> 5;
    const uint32_t lane = (uint32_t)(idx & 31u);
    const uint64_t pairs = (uint64_t)((quads_per_frame + 1) >> 1);
    const uint64_t fg = warp / pairs, tp = warp - fg * pairs;
    *f = (uint32_t)(fg * 16u + (lane & 15u));
    *t4 = (int)(tp * 2u + (lane >> 4));
}

// The planar samples of a variant: rows of xt.  The stereo variants M and S are not stored; every consumer
// forms sample = (a + m * b) >> sh from the two channel rows with (m, sh) = (0, 0) for a plain channel (pb == pa),
// (1, 1) for M = (L + R) >> 1 and (-1, 0) for S = L - R (src/coding.rs:476-484) -- one code path for all lanes.
struct FbVarRows {
    const int32_t *pa, *pb; // row bases (sample 0); four samples at t (multiple of 4) are at p + fb_xt_quad(t)
    int32_t m, sh;
};

FB_HD FbVarRows fb_variant_rows(const FbJob &J, const int32_t *xt, uint32_t f, int v) {
    FbVarRows r;
    if (J.channels == 2 && v >= 2) {
        r.pa = xt + fb_xt_off(J.stride, f * 2u, 0);
        r.pb = xt + fb_xt_off(J.stride, f * 2u + 1u, 0);
        r.m = v == 2 ? 1 : -1;
        r.sh = v == 2 ? 1 : 0;
    } else {
        r.pa = xt + fb_xt_off(J.stride, f * (uint32_t)J.channels + (uint32_t)v, 0);
        r.pb = r.pa;
        r.m = 0;
        r.sh = 0;
    }
    return r;
}

FB_HD int32_t fb_mix(int32_t a, int32_t b, int32_t m, int32_t sh) {
    return (int32_t)((uint32_t)a + (uint32_t)m * (uint32_t)b) >> sh;
}

// four samples t..t+3 of a variant (t a multiple of 4)
FB_DEV void fb_rows_load4(const FbVarRows &r, int t, int32_t *dst) {
    const size_t o = fb_xt_quad(t);
    const int4 a = *reinterpret_cast<const int4 *>(r.pa + o);
    const int4 b = *reinterpret_cast<const int4 *>(r.pb + o);
    dst[0] = fb_mix(a.x, b.x, r.m, r.sh);
    dst[1] = fb_mix(a.y, b.y, r.m, r.sh);
    dst[2] = fb_mix(a.z, b.z, r.m, r.sh);
    dst[3] = fb_mix(a.w, b.w, r.m, r.sh);
}

// K0b: the generic kernels K2/K3 index plain rows by variant, xv4[(frame * nvar + v) * stride + t] (for stereo
// with the M and S rows, src/coding.rs:476-484).  That copy is produced from xt only for the frames those kernels
// run on.
FB_DEV void fb_k0b_expand4(const FbJob &J, const int32_t *xt, int32_t *xv4, uint32_t f, int t4) {
    for (int v = 0; v < J.nvar; v++) {
        const FbVarRows rows = fb_variant_rows(J, xt, f, v);
        int32_t q[4];
        fb_rows_load4(rows, 4 * t4, q);
        int4 o;
        o.x = q[0]; o.y = q[1]; o.z = q[2]; o.w = q[3];
        *reinterpret_cast<int4 *>(xv4 + ((size_t)f * (size_t)J.nvar + (size_t)v) * (size_t)J.stride + 4 * (size_t)t4) = o;
    }
}

// =================================================================================================
// K1: analyze one channel variant (thread per variant and pass).
//
// Everything whose result depends on floating-point evaluation order is done here in exactly the
// reference's (stable build) order, so the floats are bit-identical to the scalar CPU code:
//   * estimate_entropy's per-partition f32 running sums of |e_k|   (src/coding.rs:200-227,
//     src/arrayutils.rs:496-506), fixed residuals by zero-history differences (src/coding.rs:182-197)
//   * y[t] = (f32)x[t] * w[t] (src/lpc.rs:739-756) and r[tau] += y[t-tau]*y[t] as sequential f64
//     FMAs starting at t = lpc_order for every lag (src/lpc.rs:533-548)
//   * symmetric_levinson_recursion::<f64> (src/lpc.rs:633-705), quantize_parameters (:234-302)
// R = ring size = lpc_order rounded up to a multiple of 4; lags 0..R are accumulated, lags above
// lpc_order are ignored.  Independent variants give the parallelism (32 per warp; the rows a warp walks are
// adjacent 16-byte pieces of the row-interleaved store and are staged together in a shared-memory ring),
// the R+1 independent FMA chains per thread give the ILP.
// =================================================================================================

// src/lpc.rs:633-705
FB_DEV void fb_levinson(const double *coefs, const double *ys, int order, double *dest) {
    for (int i = 0; i < order; i++) dest[i] = 0.0;
    if (order <= 0) return;
    if (coefs[0] == 0.0) return; // digital silence -> all-zero coefficients (:648-658)
    double forward[FB200_MAX_LPC_ORDER + 1], forward_next[FB200_MAX_LPC_ORDER + 1];
    for (int i = 0; i <= FB200_MAX_LPC_ORDER; i++) forward[i] = forward_next[i] = 0.0;
    forward[0] = FB_DDIV(1.0, coefs[0]);
    dest[0] = FB_DDIV(ys[0], coefs[0]);
    for (int n = 1; n < order; n++) {
        double error = 0.0;
        for (int d = 0; d < n; d++) error = FB_FMA(coefs[n - d], forward[d], error);
        double denom = FB_FMA(error, -error, 1.0);
        if (denom == 0.0) continue; // :679-682 (the loading term is never used again)
        double alpha = FB_DDIV(1.0, denom);
        double beta = FB_DMUL(-alpha, error);
        for (int d = 0; d <= n; d++) forward_next[d] = FB_FMA(alpha, forward[d], FB_DMUL(beta, forward[n - d]));
        for (int d = 0; d <= n; d++) forward[d] = forward_next[d];
        double delta = 0.0;
        for (int d = 0; d < n; d++) delta = FB_FMA(coefs[n - d], dest[d], delta);
        double g = FB_DADD(ys[n], -delta);
        for (int d = 0; d <= n; d++) dest[d] = FB_FMA(g, forward[n - d], dest[d]);
    }
}

// src/lpc.rs:234-255 find_shift: shift = clamp((precision - 1) - ceil(log2(max|a|)), 0, 15).
// ceil(log2(x)) is read off the binary representation x = 2^k * (1 + m * 2^-52) and reproduces what the reference's
// libm call returns, including its one inexact case: log2 is a rounded double, so for m != 0 the true value
// k + log2(1 + m * 2^-52) can round back to k itself when |k| >= 4 and m is tiny (log2(16 * (1 + 2^-52)) == 4.0 in
// double arithmetic), and the ceiling is then k, not k + 1.  glibc evaluates k + m * 2^-52 / ln 2 to far more than
// double precision before its final rounding, so "does k + t round to k" decides it (t never comes near a rounding
// boundary: the nearest case is 0.25 % of half an ulp away).  tests/test_kernel_logic_emu.py and the GPU test
// sweep k and m around every power of two against the host libm.
FB_DEV int fb_ceil_log2(double x) { // x >= 0, not NaN
    uint64_t bits;
    memcpy(&bits, &x, 8);
    const int e = (int)((bits >> 52) & 0x7FF);
    const uint64_t mant = bits & 0xFFFFFFFFFFFFFull;
    if (e == 0x7FF) return 32767;      // +inf: the `as i16` cast saturates
    if (e == 0) return -32752;         // zero (log2 = -inf -> i16::MIN + 16) or subnormal (<= -1022): shift saturates at 15
    const int k = e - 1023;
    if (mant == 0) return k;           // an exact power of two
    const double t = FB_DMUL((double)(int64_t)mant, 0x1.71547652b82fep-52); // m * 2^-52 / ln 2
    const double kd = (double)k;
    return FB_DADD(kd, t) == kd ? k : k + 1;
}

FB_DEV int fb_find_shift(const double *coefs, int n, int precision) {
    double max_abs = 0.0;
    bool any = false;
    for (int i = 0; i < n; i++) { // f64::max ignores NaN operands (reduce(T::max))
        double a = coefs[i] < 0 ? -coefs[i] : coefs[i];
        if (a == a && (!any || a > max_abs)) { max_abs = a; any = true; }
    }
    const int abs_log2 = any ? fb_ceil_log2(max_abs) : -32752; // all NaN: Float::max(NaN, lo) = lo
    int shift = (precision - 1) - abs_log2;
    if (shift < 0) shift = 0;
    if (shift > 15) shift = 15;
    return shift;
}

// src/lpc.rs:258-302 quantize_parameter(s); returns the truncated order
FB_DEV int fb_quantize(const double *coefs, int n, int precision, int16_t *q, int *shift_out) {
    for (int i = 0; i < 32; i++) q[i] = 0;
    int shift = fb_find_shift(coefs, n, precision);
    double scale = (double)(1 << shift);
    int lo = -(1 << (precision - 1)), hi = (1 << (precision - 1)) - 1;
    for (int i = 0; i < n; i++) {
        double s = FB_DMUL(coefs[i], scale);
#if FB_GPU
        double r = round(s); // half away from zero, exact
#else
        double r = round(s);
#endif
        if (r < -32768.0) r = -32768.0;
        if (r > 32767.0) r = 32767.0;
        int v = (int)r;
        v = v < lo ? lo : (v > hi ? hi : v);
        q[i] = (int16_t)v;
    }
    int order = FB200_MAX_LPC_ORDER;
    while (order > 0 && q[order - 1] == 0) order--;
    if (order < 1) order = 1;
    *shift_out = shift;
    return order;
}

// K1 makes two sequential passes over a variant, each with straight-line inner bodies (no per-sample branches, so
// the unrolled samples share one register-renamed schedule):
//   pass E  min/max (-> constant, max |x|) and the fixed-predictor entropy estimate, 8 samples per group
//   pass A  windowing and the autocorrelation lags, R samples per group
// On the GPU a warp (32 variants = a contiguous range of xt rows) stages the rows it walks in a shared-memory ring
// with cp.async several groups ahead, so neither pass ever waits for DRAM; the arithmetic below is shared with the
// plain per-thread driver that the CPU emulation uses.

// ---- pass E state: registers of one thread
struct FbK1Ent {
    // Sums of |e_k| of the current estimate partition.  The reference adds them one by one in f32; as long as a sum
    // stays below 2^24 every such addition is exact, so the sum is kept as an INTEGER (i0..i4: one |a - b| + c
    // instruction per sample and order instead of a conversion and an addition) and converted once when the partition
    // ends.  A group of 8 samples is added speculatively; when some lane of the warp would reach 2^24 in some order, the
    // whole warp replays that group -- and the rest of the partition -- with the sequential f32 additions (s0..s4,
    // `fmode`), which are valid in either case.
    uint32_t i0, i1, i2, i3, i4;
    bool fmode;
    float s0, s1, s2, s3, s4;       // running f32 sums of |e_k| of the current estimate partition (float mode)
    int32_t pe0, pe1, pe2, pe3;     // previous e_0..e_3 (zero history)
    int32_t xmin, xmax;
    int pstart, pend, psize, n;
    unsigned long long b0, b1, b2, b3, b4; // estimated bits per order, summed over the closed partitions
};

FB_DEV void fb_k1_ent_init(FbK1Ent &S, int n, int psize, int32_t first) {
    S.s0 = S.s1 = S.s2 = S.s3 = S.s4 = 0.f;
    S.i0 = S.i1 = S.i2 = S.i3 = S.i4 = 0u;
    S.fmode = false;
    S.pe0 = S.pe1 = S.pe2 = S.pe3 = 0;
    S.pstart = 0;
    S.psize = psize;
    S.pend = psize < n ? psize : n;
    S.n = n;
    S.b0 = S.b1 = S.b2 = S.b3 = S.b4 = 0;
    S.xmin = S.xmax = first;
}

// bits of one partition [end - len, end) for order k (src/coding.rs:209-223): the first k samples of the frame are
// not counted, the sum still includes them
FB_DEV unsigned long long fb_k1_part_bits(float sum, int k, int end, int len) {
    if (end < k) return 0;
    const int cnt = (end - k) < len ? (end - k) : len;
    return fb_entropy_partition_bits(sum, cnt);
}

// a partition ends: its five estimates are independent chains (evaluated side by side), then the sums restart
// integer sums -> the f32 sums they stand for (exact: all below 2^24); from here on the partition is in float mode
FB_DEV void fb_k1_ent_to_float(FbK1Ent &S) {
    if (!S.fmode) {
        S.s0 = (float)S.i0; S.s1 = (float)S.i1; S.s2 = (float)S.i2; S.s3 = (float)S.i3; S.s4 = (float)S.i4;
        S.fmode = true;
    }
}

FB_DEV void fb_k1_ent_close(FbK1Ent &S) {
    const int end = S.pend, len = end - S.pstart;
    fb_k1_ent_to_float(S);
    S.i0 = S.i1 = S.i2 = S.i3 = S.i4 = 0u;
    S.fmode = false;
    if (len > 0) {
        S.b0 += fb_k1_part_bits(S.s0, 0, end, len);
        S.b1 += fb_k1_part_bits(S.s1, 1, end, len);
        S.b2 += fb_k1_part_bits(S.s2, 2, end, len);
        S.b3 += fb_k1_part_bits(S.s3, 3, end, len);
        S.b4 += fb_k1_part_bits(S.s4, 4, end, len);
    }
    S.s0 = S.s1 = S.s2 = S.s3 = S.s4 = 0.f;
    S.pstart = end;
    S.pend = (end + S.psize < S.n) ? end + S.psize : S.n;
}

// the same, out of line, for the per-sample checks of the guarded groups (keeps the tail code small)
struct FbK1Bits5 { unsigned long long b[5]; };
#if FB_GPU
static __device__ __noinline__
#else
static inline
#endif
FbK1Bits5 fb_k1_part_bits5(float s0, float s1, float s2, float s3, float s4, int end, int len) {
    FbK1Bits5 r;
    r.b[0] = fb_k1_part_bits(s0, 0, end, len);
    r.b[1] = fb_k1_part_bits(s1, 1, end, len);
    r.b[2] = fb_k1_part_bits(s2, 2, end, len);
    r.b[3] = fb_k1_part_bits(s3, 3, end, len);
    r.b[4] = fb_k1_part_bits(s4, 4, end, len);
    return r;
}
FB_DEV void fb_k1_ent_close_tail(FbK1Ent &S) { // (guarded groups: always in float mode, and they stay in it)
    const int end = S.pend, len = end - S.pstart;
    S.i0 = S.i1 = S.i2 = S.i3 = S.i4 = 0u;
    if (len > 0) {
        const FbK1Bits5 r = fb_k1_part_bits5(S.s0, S.s1, S.s2, S.s3, S.s4, end, len);
        S.b0 += r.b[0]; S.b1 += r.b[1]; S.b2 += r.b[2]; S.b3 += r.b[3]; S.b4 += r.b[4];
    }
    S.s0 = S.s1 = S.s2 = S.s3 = S.s4 = 0.f;
    S.pstart = end;
    S.pend = (end + S.psize < S.n) ? end + S.psize : S.n;
}

// |a| + c on unsigned accumulators (one VABSDIFF-class instruction on the GPU)
FB_DEV uint32_t fb_abs_acc(int32_t a, uint32_t c) {
#if FB_GPU
    return __sad(a, 0, c);
#else
    return c + (a < 0 ? 0u - (uint32_t)a : (uint32_t)a);
#endif
}
// true when `ok` holds in every lane of the warp (the emulation runs one thread at a time: either answer is valid,
// both paths of the caller give the same sums)
FB_DEV bool fb_warp_all(bool ok) {
#if FB_GPU
    return __all_sync(0xFFFFFFFFu, ok);
#else
    return ok;
#endif
}

// 8 samples starting at t0.  GUARDED: samples may lie beyond n and a partition may end at any sample; otherwise all
// 8 are valid and a partition can only end with the group.  zero-history differences, wrapping i32
// (src/coding.rs:188-195); in float mode |e| is taken on the float (the conversion is symmetric; |e| < 2^31 always).
template <bool GUARDED, bool DO_ENT>
FB_DEV void fb_k1_ent_group(FbK1Ent &S, const int32_t *xs, int t0) {
    if (DO_ENT && !GUARDED && !S.fmode) {
        // speculative integer group (see FbK1Ent): inputs are at most 25 bits wide, so |e_4| <= 2^28 and eight of them
        // plus a sum below 2^24 cannot wrap
        uint32_t j0 = S.i0, j1 = S.i1, j2 = S.i2, j3 = S.i3, j4 = S.i4;
        int32_t p0 = S.pe0, p1 = S.pe1, p2 = S.pe2, p3 = S.pe3;
        int32_t mn = S.xmin, mx = S.xmax;
#pragma unroll
        for (int s = 0; s < 8; s++) {
            const int32_t e0 = xs[s];
            mn = e0 < mn ? e0 : mn;
            mx = e0 > mx ? e0 : mx;
            const int32_t e1 = (int32_t)((uint32_t)e0 - (uint32_t)p0);
            const int32_t e2 = (int32_t)((uint32_t)e1 - (uint32_t)p1);
            const int32_t e3 = (int32_t)((uint32_t)e2 - (uint32_t)p2);
            const int32_t e4 = (int32_t)((uint32_t)e3 - (uint32_t)p3);
            p0 = e0; p1 = e1; p2 = e2; p3 = e3;
            j0 = fb_abs_acc(e0, j0);
            j1 = fb_abs_acc(e1, j1);
            j2 = fb_abs_acc(e2, j2);
            j3 = fb_abs_acc(e3, j3);
            j4 = fb_abs_acc(e4, j4);
        }
        if (fb_warp_all((j0 | j1 | j2 | j3 | j4) < (1u << 24))) {
            S.i0 = j0; S.i1 = j1; S.i2 = j2; S.i3 = j3; S.i4 = j4;
            S.pe0 = p0; S.pe1 = p1; S.pe2 = p2; S.pe3 = p3;
            S.xmin = mn; S.xmax = mx;
            if (t0 + 8 == S.pend) fb_k1_ent_close(S);
            return;
        }
        fb_k1_ent_to_float(S); // the sums before this group, then the group again in f32
    }
    if (DO_ENT && GUARDED) fb_k1_ent_to_float(S);
#pragma unroll
    for (int s = 0; s < 8; s++) {
        const int t = t0 + s;
        if (GUARDED && t >= S.n) break;
        const int32_t xt = xs[s];
        S.xmin = xt < S.xmin ? xt : S.xmin;
        S.xmax = xt > S.xmax ? xt : S.xmax;
        if (DO_ENT) {
            const int32_t e0 = xt;
            const int32_t e1 = (int32_t)((uint32_t)e0 - (uint32_t)S.pe0);
            const int32_t e2 = (int32_t)((uint32_t)e1 - (uint32_t)S.pe1);
            const int32_t e3 = (int32_t)((uint32_t)e2 - (uint32_t)S.pe2);
            const int32_t e4 = (int32_t)((uint32_t)e3 - (uint32_t)S.pe3);
            S.pe0 = e0; S.pe1 = e1; S.pe2 = e2; S.pe3 = e3;
            S.s0 = FB_FADD(fabsf((float)e0), S.s0);
            S.s1 = FB_FADD(fabsf((float)e1), S.s1);
            S.s2 = FB_FADD(fabsf((float)e2), S.s2);
            S.s3 = FB_FADD(fabsf((float)e3), S.s3);
            S.s4 = FB_FADD(fabsf((float)e4), S.s4);
            if (GUARDED && t + 1 == S.pend) fb_k1_ent_close_tail(S);
        }
    }
    if (DO_ENT && !GUARDED && t0 + 8 == S.pend) fb_k1_ent_close(S);
}

// ---- pass A state
template <int R>
struct FbK1Acc {
    double acc[R + 1];              // autocorrelation lags 0..R
    double ring[R];                 // y[t-1] .. y[t-R]
};

// R samples starting at t0 (a multiple of R).  y[t] = (f32)x[t] * w[t] (src/lpc.rs:739-756); r[lag] += y[t-lag] * y[t]
// as sequential f64 FMAs starting at t = lpc_order for every lag (src/lpc.rs:533-548).
// SKIP = R - lpc_order (0..3): the lags above lpc_order are not accumulated at all.
template <int R, bool GUARDED, int SKIP>
FB_DEV void fb_k1_acc_group(FbK1Acc<R> &S, const int32_t *xs, const float *ws, int t0, int n, int P) {
#pragma unroll
    for (int s = 0; s < R; s++) {
        const int t = t0 + s;
        if (GUARDED && t >= n) break;
        const double y = (double)FB_FMUL((float)xs[s], ws[s]);
        if (!GUARDED || t >= P) {
            S.acc[0] = FB_FMA(y, y, S.acc[0]);
#pragma unroll
            for (int j = 0; j < R - SKIP; j++) {
                // logical y[t-1-j] lives in ring[(s-1-j) mod R]; static after unrolling
                S.acc[j + 1] = FB_FMA(S.ring[(s - 1 - j + 2 * R) % R], y, S.acc[j + 1]);
            }
        }
        S.ring[s] = y; // overwrites y[t-R]
    }
}

// window weights of the R samples at t0; GUARDED: nothing is read at or beyond n (the tables end shortly after n)
template <int R, bool GUARDED>
FB_DEV void fb_k1_win_load(const float *win, int t0, int n, float *ws) {
#pragma unroll
    for (int i = 0; i < R; i += 4) {
        float4 v;
        v.x = v.y = v.z = v.w = 0.f;
        if (!GUARDED || t0 + i < n) v = *reinterpret_cast<const float4 *>(win + t0 + i);
        ws[i] = v.x; ws[i + 1] = v.y; ws[i + 2] = v.z; ws[i + 3] = v.w;
    }
}

// R must equal fb_k1_ring(J.cfg.lpc_order); one kernel instantiation per R keeps the register
// allocation of the common small orders independent of the order-24 case.
FB_HD int fb_k1_ring(int lpc_order) { return (lpc_order + 3) & ~3; }

// ---- what a variant's analysis needs besides the samples
struct FbK1Var {
    int n, bps_v, P, psize;
    bool do_ent, do_lpc;
    const float *win;
};

FB_DEV FbK1Var fb_k1_var(const FbJob &J, uint32_t f, int v, const float *win_full, const float *win_tail) {
    FbK1Var V;
    V.n = fb_frame_len(J, f);
    V.bps_v = fb_variant_bps(J, v);
    V.P = J.cfg.lpc_order;
    const bool too_short = V.n < FB_MIN_PRED_BLOCK;
    V.do_ent = !too_short && J.cfg.use_fixed && J.cfg.fixed_order_sel == 1;
    V.do_lpc = !too_short && J.cfg.use_lpc;
    const int parts = J.cfg.approx_ent_partitions;
    V.psize = (V.n + parts - 1) / parts;
    V.win = (V.n == J.block_size) ? win_full : win_tail;
    return V;
}

// results of pass E: constant flag, max |x|, ApproxEnt order (src/coding.rs:264-285, :396-401)
FB_DEV void fb_k1_finish_ent(const FbJob &J, const FbK1Var &V, const FbK1Ent &S, FbAnalysis *out, fb200_variant_taps *taps) {
    const bool allsame = S.xmin == S.xmax; // src/arrayutils.rs:382-389
    out->is_constant = allsame ? 1 : 0;
    {
        const uint32_t a = S.xmin < 0 ? (uint32_t)(-(int64_t)S.xmin) : (uint32_t)S.xmin;
        const uint32_t b = S.xmax < 0 ? (uint32_t)(-(int64_t)S.xmax) : (uint32_t)S.xmax;
        out->max_abs = a > b ? a : b;
        out->pad = 0;
    }
    out->fixed_order = -1;
    for (int k = 0; k < 5; k++) out->fixed_est[k] = 0;
    if (taps) { // (the autocorrelation / LPC fields belong to fb_k1_finish_lpc, which may run in another thread)
        taps->is_constant = allsame ? 1 : 0;
        taps->fixed_order = -1;
        for (int k = 0; k < 5; k++) taps->fixed_est_bits[k] = 0;
    }
    if (V.do_ent) {
        // estimate_entropy per order + bits_per_sample * order; first minimum wins; accepted only
        // if below the verbatim size
        const int n_orders = (J.cfg.fixed_max_order < 4 ? J.cfg.fixed_max_order : 4) + 1;
        const uint64_t verbatim_bits = 8 + (uint64_t)V.n * (uint64_t)V.bps_v;
        const unsigned long long est[5] = {S.b0, S.b1, S.b2, S.b3, S.b4};
        int best = -1;
        uint64_t best_bits = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            if (k < n_orders) {
                const uint64_t bits = est[k] + (uint64_t)V.bps_v * (uint64_t)k;
                out->fixed_est[k] = bits;
                if (taps) taps->fixed_est_bits[k] = bits;
                if (best < 0 || bits < best_bits) { best = k; best_bits = bits; }
            }
        }
        if (best >= 0 && best_bits < verbatim_bits) out->fixed_order = best;
        if (taps) taps->fixed_order = out->fixed_order;
    }
}

// results of pass A: Levinson, quantiser
template <int R>
FB_DEV void fb_k1_finish_lpc(const FbJob &J, const FbK1Var &V, const FbK1Acc<R> &A, FbAnalysis *out, fb200_variant_taps *taps,
                              uint32_t gv) {
    out->qlp_order = 0;
    out->qlp_shift = 0;
    if (taps) {
        for (int i = 0; i <= FB200_MAX_LPC_ORDER; i++) taps->autocorr[i] = 0.0;
        for (int i = 0; i < FB200_MAX_LPC_ORDER; i++) taps->lpc[i] = 0.0;
        for (int i = 0; i < 32; i++) taps->qlp[i] = 0;
        taps->qlp_order = 0;
        taps->qlp_shift = 0;
    }
    if (!V.do_lpc) {
        for (int i = 0; i < 32; i++) out->qlp[i] = 0;
        return;
    }
    double corr[FB200_MAX_LPC_ORDER + 1];
#pragma unroll
    for (int i = 0; i <= R; i++)
        if (i <= FB200_MAX_LPC_ORDER) corr[i] = A.acc[i];
    double lpc[FB200_MAX_LPC_ORDER];
    fb_levinson(corr, corr + 1, V.P, lpc);
    int16_t q[32];
    int shift;
    int order = fb_quantize(lpc, V.P, J.cfg.quant_precision, q, &shift);
    out->qlp_order = order;
    out->qlp_shift = shift;
    for (int i = 0; i < 32; i++) out->qlp[i] = i < order ? q[i] : (int16_t)0;
    if (taps) {
        for (int i = 0; i <= V.P; i++) taps->autocorr[i] = corr[i];
        for (int i = 0; i < V.P; i++) taps->lpc[i] = lpc[i];
        for (int i = 0; i < 32; i++) taps->qlp[i] = out->qlp[i];
        taps->qlp_order = order;
        taps->qlp_shift = shift;
    }
    if (J.lpc_ext) {
        // EXTENSIONS: the Levinson solutions of the lower orders on the same autocorrelation, quantised alike, then the
        // order-P solution quantised with fewer bits; unused entries get order 0
        int orders[FB_EXT_LPC_MAX];
        const int no = fb_ext_lpc_orders(V.P, J.cfg.ext_lpc_order_search, orders);
        FbLpcExt *ext = J.lpc_ext + (size_t)gv * FB_EXT_LPC_MAX;
        double lpc_p[FB200_MAX_LPC_ORDER];
        for (int i = 0; i < FB200_MAX_LPC_ORDER; i++) lpc_p[i] = lpc[i];
        int np = 0; // precision candidates written so far
        for (int k = 0; k < FB_EXT_LPC_MAX; k++) {
            int ok = 0, sk = 0, pk = J.cfg.quant_precision;
            if (k < no) {
                fb_levinson(corr, corr + 1, orders[k], lpc);
                ok = fb_quantize(lpc, orders[k], pk, q, &sk);
            } else if (np < J.cfg.ext_lpc_precision_search && J.cfg.quant_precision - (np + 1) >= 1) {
                np++;
                pk = J.cfg.quant_precision - np;
                ok = fb_quantize(lpc_p, V.P, pk, q, &sk);
            }
            ext[k].order = ok;
            ext[k].shift = sk;
            ext[k].precision = pk;
            for (int i = 0; i < 32; i++) ext[k].qlp[i] = i < ok ? q[i] : (int16_t)0;
        }
    }
}

template <int R, int SKIP>
FB_DEV void fb_k1_pass_a(FbK1Acc<R> &A, const FbVarRows &rows, const FbK1Var &V) {
    int32_t xs[R];
    float ws[R];
    const int n = V.n;
    for (int t0 = 0; t0 < n; t0 += R) {
        const bool fast = t0 > 0 && t0 + R <= n;
#pragma unroll
        for (int i = 0; i < R; i += 4) fb_rows_load4(rows, t0 + i, xs + i);
        if (fast) {
            fb_k1_win_load<R, false>(V.win, t0, n, ws);
            fb_k1_acc_group<R, false, SKIP>(A, xs, ws, t0, n, V.P);
        } else {
            fb_k1_win_load<R, true>(V.win, t0, n, ws);
            fb_k1_acc_group<R, true, SKIP>(A, xs, ws, t0, n, V.P);
        }
    }
}

// plain driver: one thread walks its variant straight from xt (CPU emulation; the GPU kernel is fb_k1_warp below)
template <int R>
FB_DEV void fb_k1_thread(const FbJob &J, const int32_t *xt, const float *win_full, const float *win_tail,
                         FbAnalysis *ana, fb200_variant_taps *taps_all, uint32_t gv) {
    const uint32_t f = gv / (uint32_t)J.nvar;
    const int v = (int)(gv - f * (uint32_t)J.nvar);
    const FbK1Var V = fb_k1_var(J, f, v, win_full, win_tail);
    const FbVarRows rows = fb_variant_rows(J, xt, f, v);
    FbAnalysis *out = ana + gv;
    fb200_variant_taps *taps = taps_all ? taps_all + gv : nullptr;
    const int n = V.n;

    FbK1Ent S;
    {
        int32_t q[4];
        fb_rows_load4(rows, 0, q);
        fb_k1_ent_init(S, n, V.psize, q[0]);
    }
    {
        int32_t xs[8];
        const bool whole = (V.psize & 7) == 0;
        for (int t0 = 0; t0 < n; t0 += 8) {
            fb_rows_load4(rows, t0, xs);
            fb_rows_load4(rows, t0 + 4, xs + 4);
            const bool fast = t0 + 8 <= n && (whole || !V.do_ent);
            if (V.do_ent) {
                if (fast) fb_k1_ent_group<false, true>(S, xs, t0);
                else fb_k1_ent_group<true, true>(S, xs, t0);
            } else {
                if (fast) fb_k1_ent_group<false, false>(S, xs, t0);
                else fb_k1_ent_group<true, false>(S, xs, t0);
            }
        }
    }
    fb_k1_finish_ent(J, V, S, out, taps);

    FbK1Acc<R> A;
#pragma unroll
    for (int i = 0; i <= R; i++) A.acc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < R; i++) A.ring[i] = 0.0;
    if (V.do_lpc && !J.cfg.use_direct_mse) { // (direct MSE: K1C estimates the LPC)
        switch (R - V.P) {
        case 0: fb_k1_pass_a<R, 0>(A, rows, V); break;
        case 1: fb_k1_pass_a<R, 1>(A, rows, V); break;
        case 2: fb_k1_pass_a<R, 2>(A, rows, V); break;
        default: fb_k1_pass_a<R, 3>(A, rows, V); break;
        }
    }
    fb_k1_finish_lpc<R>(J, V, A, out, taps, gv);
}

// rows of xt that the 32 variants of one warp can span, rounded to the staging pitch (16, 32 or 48 rows)
FB_HD int fb_k1_pitch_rows(int channels, int nvar) {
    const int frames = (32 % nvar == 0) ? 32 / nvar : (32 + nvar - 1) / nvar + 1;
    const int rows = frames * channels;
    return rows <= 16 ? 16 : (rows <= 32 ? 32 : 48);
}
#define FB_K1_THREADS 128
#ifndef FB_K1_NQ
#define FB_K1_NQ 32 // quads of every staged row in a warp's ring (128 samples)
#endif
#ifndef FB_K1_GCA
#define FB_K1_GCA 2 // pass A, ring lengths 12 and 16: groups per staged chunk
#endif
#ifndef FB_K1_GCE
#define FB_K1_GCE 4 // pass E: groups of 8 samples per staged chunk
#endif
// lane slots of a launch: the variants in order; a shorter last frame is moved to the next warp boundary
FB_HD uint32_t fb_k1_full_variants(const FbJob &J, uint32_t n_variants) {
    return (J.n_frames > 1 && J.tail_n != J.block_size) ? n_variants - (uint32_t)J.nvar : n_variants;
}
FB_HD uint32_t fb_k1_slots(const FbJob &J, uint32_t n_variants) {
    const uint32_t n_full = fb_k1_full_variants(J, n_variants);
    return n_full < n_variants ? ((n_full + 31u) & ~31u) + (n_variants - n_full) : n_variants;
}
FB_HD uint32_t fb_k1_smem_bytes(int channels, int nvar, bool pairs = false) {
    const uint32_t rows = pairs ? 8u : (uint32_t)fb_k1_pitch_rows(channels, nvar); // pairs: a warp's 32 variants are 8 frames
    return (uint32_t)(FB_K1_THREADS / 32) * FB_K1_NQ * rows * 16u;
}
// 16-bit stereo in a 2-byte container with a block size that is a multiple of 4 and 16-byte aligned PCM: the analysis,
// plan and pack kernels read the packed PCM itself (as (left, right) pairs) and the planar store is never built
FB_HD bool fb_pairs_format(int channels, int bps, int container_bytes, int block_size) {
    return channels == 2 && bps == 16 && container_bytes == 2 && (block_size & 3) == 0;
}

#if FB_GPU
// ---- GPU driver: a warp stages the xt rows of its 32 variants in a shared-memory ring (cp.async, NG - 1 groups
// ahead) and every lane then walks its variant out of the ring.  A "quad" is four samples of every staged row: in xt
// the quads of up to 32 consecutive rows are contiguous, and they keep that order in the ring, so the copy is
// 16 bytes per lane and the reads of the lanes (16 bytes each from row a and row b) are conflict free.
struct FbK1Stage {
    uint32_t ring;          // shared-space address of the warp's ring
    uint32_t pitch;         // bytes per staged quad
    // copy role: this lane copies quads qsub, qsub + qstep, ... of every group, 16 bytes of one row (two rows when
    // more than 32 rows are staged); lanes without a row have cp_quads = 0
    const int32_t *g0;      // global address of (row, quad qsub)
    uint32_t d0;            // byte offset of (row, quad qsub) inside a group slot
    uint32_t g1_off;        // second row: word offset from g0 (0: none), staged 512 bytes further
    int qsub, qstep;
    int cp_quads;           // quads stored per row (0: nothing to copy)
    // read role
    uint32_t ra, rb;        // byte offsets of this lane's a / b row inside a staged quad
    int32_t m, sh;          // sample = (a + m * b) >> sh
    // pairs mode (16-bit stereo straight from the packed PCM, no planar store): a staged "row" is a frame, a quad holds
    // four (left, right) pairs, the lane copies quads qsub, qsub + 4, ... of frame lane & 7 and forms its variant's
    // samples as dp2a(pair, mb) >> sh.  The last quad of a frame whose length is not a multiple of 4 is copied with
    // `last_bytes` source bytes and zero fill (nothing is read beyond the PCM buffer).
    bool pairs;
    int32_t mb;
    uint32_t last_bytes;
};

template <int N>
FB_DEV void fb_k1_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

FB_DEV void fb_k1_cp16(uint32_t dst, const int32_t *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// copies the QC quads starting at quad q0 to the chunk slot at shared address `slot_addr`; always commits one
// cp.async group.  The common case (16-row pitch, whole chunk inside the rows) is a straight run of copies at constant
// offsets.
template <int QC>
FB_DEV void fb_k1_stage_issue(const FbK1Stage &T, uint32_t slot_addr, int q0) {
    const int32_t *src = T.g0 + (size_t)q0 * (32u * FB_XT_CH);
    const uint32_t dst = slot_addr + T.d0;
    if (q0 + QC <= T.cp_quads) { // warp-uniform unless some lanes have no row at all
        if (T.qstep == 2) {
#pragma unroll
            for (int k = 0; k < QC / 2; k++) fb_k1_cp16(dst + (uint32_t)k * 512u, src + (size_t)k * (2u * 32u * FB_XT_CH));
            if ((QC & 1) && T.qsub == 0) fb_k1_cp16(dst + (uint32_t)(QC / 2) * 512u, src + (size_t)(QC / 2) * (2u * 32u * FB_XT_CH));
        } else {
#pragma unroll
            for (int k = 0; k < QC; k++) {
                fb_k1_cp16(dst + (uint32_t)k * T.pitch, src + (size_t)k * (32u * FB_XT_CH));
                if (T.g1_off) fb_k1_cp16(dst + (uint32_t)k * T.pitch + 512u, src + (size_t)k * (32u * FB_XT_CH) + T.g1_off);
            }
        }
    } else {
#pragma unroll 1
        for (int i = T.qsub; i < QC; i += T.qstep) {
            if (q0 + i < T.cp_quads) {
                const int k = i - T.qsub;
                fb_k1_cp16(dst + (uint32_t)k * T.pitch, src + (size_t)k * (32u * FB_XT_CH));
                if (T.g1_off) fb_k1_cp16(dst + (uint32_t)k * T.pitch + 512u, src + (size_t)k * (32u * FB_XT_CH) + T.g1_off);
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// pairs mode: the lane copies quads qsub, qsub + 4, ... of the chunk (four lanes cover 64 contiguous bytes of a frame)
template <int QC>
FB_DEV void fb_k1_stage_issue_pairs(const FbK1Stage &T, uint32_t slot_addr, int q0) {
#pragma unroll
    for (int k = 0; 4 * k < QC; k++) {
        const int qi = T.qsub + 4 * k, q = q0 + qi;
        if (qi < QC && q < T.cp_quads) {
            const uint32_t bytes = (q + 1 == T.cp_quads) ? T.last_bytes : 16u;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(slot_addr + T.d0 + (uint32_t)(4 * k) * T.pitch),
                         "l"(T.g0 + (size_t)(q0 + 4 * k) * 4u), "r"(bytes) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int QC, bool PAIRS>
FB_DEV void fb_k1_stage_issue_any(const FbK1Stage &T, uint32_t slot_addr, int q0) {
    if (PAIRS) fb_k1_stage_issue_pairs<QC>(T, slot_addr, q0);
    else fb_k1_stage_issue<QC>(T, slot_addr, q0);
}

// the lane's 4 * QG samples of the group slot at shared address `slot_addr`
template <int QG, bool PAIRS>
FB_DEV void fb_k1_stage_read(const FbK1Stage &T, uint32_t slot_addr, int32_t *xs) {
    if (PAIRS) {
#pragma unroll
        for (int i = 0; i < QG; i++) {
            int4 a;
            asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w)
                         : "r"(slot_addr + (uint32_t)i * T.pitch + T.ra));
            xs[4 * i + 0] = fb_dp2a_lo(a.x, T.mb, 0) >> T.sh;
            xs[4 * i + 1] = fb_dp2a_lo(a.y, T.mb, 0) >> T.sh;
            xs[4 * i + 2] = fb_dp2a_lo(a.z, T.mb, 0) >> T.sh;
            xs[4 * i + 3] = fb_dp2a_lo(a.w, T.mb, 0) >> T.sh;
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < QG; i++) {
        const uint32_t base = slot_addr + (uint32_t)i * T.pitch;
        int4 a, b;
        asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(base + T.ra));
        asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(base + T.rb));
        xs[4 * i + 0] = fb_mix(a.x, b.x, T.m, T.sh);
        xs[4 * i + 1] = fb_mix(a.y, b.y, T.m, T.sh);
        xs[4 * i + 2] = fb_mix(a.z, b.z, T.m, T.sh);
        xs[4 * i + 3] = fb_mix(a.w, b.w, T.m, T.sh);
    }
}

// One pass of the warp over its rows in groups of 4 * QG samples, staged in chunks of GC groups.  body(g, xs) gets
// group g's samples of this lane.  NC chunk slots: NC - 1 copies in flight while one chunk is read.
template <int QG, int GC, bool PAIRS, typename Body>
FB_DEV void fb_k1_stream(const FbK1Stage &T, int groups, Body body) {
    constexpr int QC = QG * GC;
    constexpr int NC = FB_K1_NQ / QC;
    constexpr int D = NC - 1;
    static_assert(NC >= 2, "ring too small");
    const uint32_t group_bytes = (uint32_t)QG * T.pitch, slot_bytes = (uint32_t)GC * group_bytes;
    const uint32_t ring_end = T.ring + (uint32_t)NC * slot_bytes;
#ifndef FB_X_NOSTAGE
#pragma unroll 1
    for (int c = 0; c < D; c++) fb_k1_stage_issue_any<QC, PAIRS>(T, T.ring + (uint32_t)c * slot_bytes, c * QC);
#else
    for (uint32_t o = (threadIdx.x & 31u) * 4u; o < (uint32_t)NC * slot_bytes; o += 128u)
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(T.ring + o), "r"(0) : "memory");
    __syncwarp();
#endif
    uint32_t slot_r = T.ring, slot_w = T.ring + (uint32_t)D * slot_bytes;
    int q_w = D * QC;
#pragma unroll 1
    for (int g0 = 0; g0 < groups; g0 += GC) {
#ifndef FB_X_NOSTAGE
        fb_k1_cp_wait<D - 1>(); // this chunk has landed (this lane's copies) ...
        __syncwarp();           // ... and everybody's; all lanes are also done reading the previous chunk
        fb_k1_stage_issue_any<QC, PAIRS>(T, slot_w, q_w);
#endif
        q_w += QC;
        slot_w += slot_bytes;
        slot_w = slot_w == ring_end ? T.ring : slot_w;
#pragma unroll 1
        for (int gi = 0; gi < GC; gi++) {
            if (g0 + gi >= groups) break;
            int32_t xs[4 * QG];
            fb_k1_stage_read<QG, PAIRS>(T, slot_r + (uint32_t)gi * group_bytes, xs);
            body(g0 + gi, xs);
        }
        slot_r += slot_bytes;
        slot_r = slot_r == ring_end ? T.ring : slot_r;
    }
    fb_k1_cp_wait<0>();
    __syncwarp(); // the ring may be refilled by the next pass
}

template <int R, int SKIP, bool PAIRS>
FB_DEV void fb_k1_warp_pass_a(const FbK1Stage &T, FbK1Acc<R> &A, const FbK1Var &V, int n_w, bool uniform) {
    fb_k1_stream<R / 4, (R <= 8 ? 4 : (R <= 16 ? FB_K1_GCA : 1)), PAIRS>(T, (n_w + R - 1) / R, [&](int g, const int32_t *xs) {
        const int t0 = g * R;
        float ws[R];
        if (uniform && g > 0 && t0 + R <= n_w) {
#ifdef FB_X_NOWIN
#pragma unroll
            for (int i = 0; i < R; i++) ws[i] = 1.0f;
#else
            fb_k1_win_load<R, false>(V.win, t0, V.n, ws);
#endif
            fb_k1_acc_group<R, false, SKIP>(A, xs, ws, t0, V.n, V.P);
        } else {
            fb_k1_win_load<R, true>(V.win, t0, V.n, ws);
            fb_k1_acc_group<R, true, SKIP>(A, xs, ws, t0, V.n, V.P);
        }
    });
}

// what a lane of the analysis kernels knows about its variant and its share of the warp's staging
struct FbK1Lane {
    bool valid;      // the lane has a variant of its own (surplus lanes shadow the last one: they still copy)
    uint32_t gve;    // index of the variant
    int n_w;         // frame length of the warp's first frame
    bool uniform;    // all lanes of the warp walk frames of that length
    FbK1Var V;
    FbK1Stage T;
};

// Variants to lanes for block `blk` of 128 lane slots, and the staging roles.  false: the whole warp has nothing to do.
// PAIRS: pairs mode (fb_pairs_format): the rows come from pcm, xt is not read
template <bool PAIRS>
FB_DEV bool fb_k1_lane_setup(const FbJob &J, const int32_t *xt, const uint8_t *pcm, const float *win_full, const float *win_tail,
                             uint32_t n_variants, uint32_t blk, uint8_t *smem, FbK1Lane &W) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t nvar = (uint32_t)J.nvar, ch = (uint32_t)J.channels;
    // Variants to lanes: in order, 32 per warp -- except that the variants of a shorter last frame start a warp of
    // their own (fb_k1_slots), so that every warp walks frames of one length and can use the unguarded groups.
    const uint32_t slot = blk * (uint32_t)FB_K1_THREADS + threadIdx.x, slot0 = slot - lane;
    const uint32_t n_full = fb_k1_full_variants(J, n_variants); // variants before the isolated last frame
    const uint32_t tail_slot = (n_full + 31u) & ~31u;
    const bool tail_warp = n_full < n_variants && slot0 >= tail_slot;
    const uint32_t gv0 = tail_warp ? n_full + (slot0 - tail_slot) : slot0;
    const uint32_t lim = tail_warp ? n_variants : n_full;        // end of the variants this warp may take
    if (gv0 >= lim) return false; // the whole warp has nothing to do
    const uint32_t gv = gv0 + lane;
    W.valid = gv < lim;
    const bool valid = W.valid;
    const uint32_t gve = valid ? gv : lim - 1u; // surplus lanes shadow the last variant (they still copy)
    const uint32_t f = gve / nvar;
    const int v = (int)(gve - f * nvar);
    const uint32_t gvl = gv0 + 31u < lim ? gv0 + 31u : lim - 1u;
    const uint32_t f_lo = gv0 / nvar, f_hi = gvl / nvar;
    const uint32_t rlo = f_lo * ch;
    const uint32_t NR = (f_hi - f_lo + 1u) * ch;
    W.n_w = fb_frame_len(J, f_lo);                       // frames of a batch only get shorter at its end
    W.uniform = fb_frame_len(J, f_hi) == W.n_w;          // all lanes walk frames of one length
    W.V = fb_k1_var(J, f, v, win_full, win_tail);
    W.gve = gve;
    FbK1Stage &T = W.T;
    T.pairs = PAIRS;
    T.mb = 0;
    T.last_bytes = 16u;
    if (PAIRS) {
        // rows = the warp's (up to 8) frames; lane & 7 = frame to copy, lane >> 3 = first quad of every four
        T.pitch = 8u * 16u;
        T.ring = (uint32_t)__cvta_generic_to_shared(smem) + warp * (uint32_t)FB_K1_NQ * T.pitch;
        fb_pair_mix(v, &T.mb, &T.sh);
        T.m = 0;
        T.ra = T.rb = (f - f_lo) * 16u;
        const uint32_t fi = lane & 7u;
        T.qsub = (int)(lane >> 3);
        T.qstep = 4;
        const uint32_t fc = f_lo + fi <= f_hi ? f_lo + fi : f_lo;
        const int n_c = fb_frame_len(J, fc);
        T.g0 = reinterpret_cast<const int32_t *>(pcm) + (size_t)fc * (size_t)J.block_size + (size_t)T.qsub * 4u;
        T.d0 = fi * 16u + (uint32_t)T.qsub * T.pitch;
        T.cp_quads = f_lo + fi <= f_hi ? (n_c + 3) >> 2 : 0;
        T.last_bytes = (n_c & 3) ? (uint32_t)(n_c & 3) * 4u : 16u;
        T.g1_off = 0;
    } else {
        const uint32_t prow = (uint32_t)fb_k1_pitch_rows(J.channels, J.nvar);
        T.pitch = prow * 16u;
        T.ring = (uint32_t)__cvta_generic_to_shared(smem) + warp * (uint32_t)FB_K1_NQ * T.pitch;
        uint32_t row_a, row_b;
        if (ch == 2u && v >= 2) {
            row_a = f * 2u; row_b = row_a + 1u;
            T.m = v == 2 ? 1 : -1;
            T.sh = v == 2 ? 1 : 0;
        } else {
            row_a = row_b = f * ch + (uint32_t)v;
            T.m = 0; T.sh = 0;
        }
        T.ra = (row_a - rlo) * 16u;
        T.rb = (row_b - rlo) * 16u;
        uint32_t ri = lane;
        T.qsub = 0;
        T.qstep = 1;
        if (prow == 16u) { ri = lane & 15u; T.qsub = (int)(lane >> 4); T.qstep = 2; }
        const uint32_t rc = ri < NR ? ri : 0u; // lanes beyond the rows copy nothing
        T.g0 = xt + fb_xt_off(J.stride, rlo + rc, 0) + (size_t)T.qsub * (32u * FB_XT_CH);
        T.d0 = ri * 16u + (uint32_t)T.qsub * T.pitch;
        T.cp_quads = ri < NR ? J.stride / FB_XT_CH : 0;
        // more than 32 rows: lane ri also copies row ri + 32, which lies one 32-row unit further in xt; the lanes
        // whose second row does not exist get offset 0 and skip it
        T.g1_off = 0;
        if (prow > 32u) {
            const uint32_t r2 = ri + 32u < NR ? ri + 32u : rc;
            T.g1_off = (uint32_t)(fb_xt_off(J.stride, rlo + r2, 0) - fb_xt_off(J.stride, rlo + rc, 0));
        }
    }

    return true;
}

template <int R, bool PAIRS>
FB_DEV void fb_k1_warp(const FbJob &J, const int32_t *xt, const uint8_t *pcm, const float *win_full, const float *win_tail,
                       FbAnalysis *ana, fb200_variant_taps *taps_all, uint32_t n_variants, uint8_t *smem) {
    // The grid holds every block of 128 variants twice: the first half runs pass A (the longer one, so it is scheduled
    // first), the second half pass E.  Half-length CTAs pack the SMs' slots better at the end of a launch, and small
    // batches (latency-bound: one thread walks a whole frame) finish in roughly half the time.
    const uint32_t nblk = gridDim.x >> 1;
    const bool role_a = blockIdx.x < nblk;
    const uint32_t blk = role_a ? blockIdx.x : blockIdx.x - nblk;
    FbK1Lane W;
    if (!fb_k1_lane_setup<PAIRS>(J, xt, pcm, win_full, win_tail, n_variants, blk, smem, W)) return;
    const FbK1Stage &T = W.T;
    const FbK1Var &V = W.V;
    const bool valid = W.valid, uniform = W.uniform;
    const int n_w = W.n_w;
    FbAnalysis *out = ana + W.gve;
    fb200_variant_taps *taps = taps_all ? taps_all + W.gve : nullptr;

#ifdef FB_X_ONLY_A
    if (!role_a) return;
#endif
#ifdef FB_X_ONLY_E
    if (role_a) return;
#endif
    // ---- pass E
    if (!role_a) {
        FbK1Ent S;
        fb_k1_ent_init(S, V.n, V.psize, 0);
        S.xmin = 2147483647;
        S.xmax = -2147483647 - 1;
        const bool ent_w = J.cfg.use_fixed && J.cfg.fixed_order_sel == 1 && n_w >= FB_MIN_PRED_BLOCK;
        const bool whole = uniform && (!ent_w || (V.psize & 7) == 0);
        if (ent_w)
            fb_k1_stream<2, FB_K1_GCE, PAIRS>(T, (n_w + 7) / 8, [&](int g, const int32_t *xs) {
                if (whole && g * 8 + 8 <= n_w) fb_k1_ent_group<false, true>(S, xs, g * 8);
                else fb_k1_ent_group<true, true>(S, xs, g * 8);
            });
        else
            fb_k1_stream<2, FB_K1_GCE, PAIRS>(T, (n_w + 7) / 8, [&](int g, const int32_t *xs) {
                if (whole && g * 8 + 8 <= n_w) fb_k1_ent_group<false, false>(S, xs, g * 8);
                else fb_k1_ent_group<true, false>(S, xs, g * 8);
            });
        if (valid) fb_k1_finish_ent(J, V, S, out, taps);
        return;
    }

    // ---- pass A
    FbK1Acc<R> A;
#pragma unroll
    for (int i = 0; i <= R; i++) A.acc[i] = 0.0;
#pragma unroll
    for (int i = 0; i < R; i++) A.ring[i] = 0.0;
    if (J.cfg.use_lpc && !J.cfg.use_direct_mse && n_w >= FB_MIN_PRED_BLOCK) { // (direct MSE: K1C / K1D estimate the LPC)
        switch (R - V.P) {
        case 0: fb_k1_warp_pass_a<R, 0, PAIRS>(T, A, V, n_w, uniform); break;
        case 1: fb_k1_warp_pass_a<R, 1, PAIRS>(T, A, V, n_w, uniform); break;
        case 2: fb_k1_warp_pass_a<R, 2, PAIRS>(T, A, V, n_w, uniform); break;
        default: fb_k1_warp_pass_a<R, 3, PAIRS>(T, A, V, n_w, uniform); break;
        }
    }
    if (valid) fb_k1_finish_lpc<R>(J, V, A, out, taps, W.gve);
}
#endif

#if !FB_GPU
inline void fb_k1_dispatch(const FbJob &J, const int32_t *xt, const float *win_full, const float *win_tail,
                           FbAnalysis *ana, fb200_variant_taps *taps, uint32_t gv) {
    switch (fb_k1_ring(J.cfg.lpc_order)) {
    case 4: fb_k1_thread<4>(J, xt, win_full, win_tail, ana, taps, gv); break;
    case 8: fb_k1_thread<8>(J, xt, win_full, win_tail, ana, taps, gv); break;
    case 12: fb_k1_thread<12>(J, xt, win_full, win_tail, ana, taps, gv); break;
    case 16: fb_k1_thread<16>(J, xt, win_full, win_tail, ana, taps, gv); break;
    case 20: fb_k1_thread<20>(J, xt, win_full, win_tail, ana, taps, gv); break;
    default: fb_k1_thread<24>(J, xt, win_full, win_tail, ana, taps, gv); break;
    }
}
#endif

// =================================================================================================
// K1C: the `experimental` direct-MSE (covariance method) LPC estimator of one channel variant, CTA per variant
// (src/lpc.rs:852-913 weighted_lpc_with_direct_mse with NoWeight; config key qlpc.use_direct_mse).
//   y[t] = (f32)x[t] * w[t]                                              (src/lpc.rs:739-756)
//   r[tau]  += y[t - tau] * y[t],        t = P .. n-1,  tau = 0..P        (src/lpc.rs:533-548, order P + 1)
//   C[i][j] += y[t - i] * y[t - j],      t = P-1 .. n-2, i <= j < P       (src/lpc.rs:573-600 on y[0..n-1))
// every sum a sequential f64 FMA chain in t, one chain per thread (P (P + 1) / 2 + P + 1 threads); then one thread
// solves C a = r[1..P] by Cholesky (nalgebra's operation order, restated in oracle/flacenc_oracle.c fo_solve_sym;
// parity with the crate is unpinned there), regularising the diagonal by 1, 1, 2, 4 ... while C is not positive definite,
// and quantises the coefficients (src/lpc.rs:273-302).  K1's pass A is skipped for such configurations; this kernel
// runs after K1 and fills the LPC half of FbAnalysis.  The frame is walked in tiles of FB_K1C_TILE samples staged as
// doubles in shared memory (history of P samples carried from tile to tile), so any block size fits.
// =================================================================================================
#define FB_K1C_TILE 2048
FB_HD int fb_k1c_chains(int P) { return P * (P + 1) / 2 + P + 1; }
FB_HD int fb_k1c_threads(int P) { return (fb_k1c_chains(P) + 31) & ~31; }
// shared memory: y tile (history + tile) | per-thread accumulators (the emulation keeps them here) | C | r | coefficients
FB_HD uint32_t fb_k1c_smem_bytes(int P, int tile) {
    return (uint32_t)((FB200_MAX_LPC_ORDER + tile) * 8 + fb_k1c_threads(P) * 8 + (P * P + 2 * (P + 1)) * 8 + 64);
}

// sample t of variant v of frame f: from the planar store, or (pcm != nullptr) from packed 16-bit stereo pairs
FB_DEV int32_t fb_k1c_sample(const FbJob &J, const int32_t *xt, const uint8_t *pcm, const FbVarRows &rows, int32_t mb, int32_t sh,
                             uint32_t f, int t) {
    if (pcm) {
        const int32_t w = reinterpret_cast<const int32_t *>(pcm)[(size_t)f * (size_t)J.block_size + (size_t)t];
        return fb_dp2a_lo(w, mb, 0) >> sh;
    }
    const size_t o = fb_xt_quad(t & ~3) + (size_t)(t & 3);
    return fb_mix(rows.pa[o], rows.pb[o], rows.m, rows.sh);
}

// src/lpc.rs:76-87 solve_sym_mut = nalgebra Cholesky::new + solve_mut, in place on the n x n matrix L (row major,
// lower triangle) and the right-hand side v; false when the matrix is not positive definite
FB_DEV bool fb_solve_sym(double *L, int n, double *v) {
    for (int j = 0; j < n; j++) {
        for (int k = 0; k < j; k++) {
            const double factor = -L[j * n + k];
            for (int r = j; r < n; r++) L[r * n + j] = FB_DADD(FB_DMUL(factor, L[r * n + k]), L[r * n + j]);
        }
        const double diag = L[j * n + j];
        if (diag == 0.0 || !(diag >= 0.0)) return false;
#if FB_GPU
        const double denom = __dsqrt_rn(diag);
#else
        const double denom = sqrt(diag);
#endif
        L[j * n + j] = denom;
        for (int r = j + 1; r < n; r++) L[r * n + j] = FB_DDIV(L[r * n + j], denom);
    }
    for (int i = 0; i < n; i++) {
        const double diag = L[i * n + i];
        if (diag == 0.0) return false;
        const double coeff = FB_DDIV(v[i], diag);
        v[i] = coeff;
        for (int r = i + 1; r < n; r++) v[r] = FB_DADD(FB_DMUL(-coeff, L[r * n + i]), v[r]);
    }
    for (int i = n - 1; i >= 0; i--) {
        // dot(L[i+1.., i], v[i+1..]) with nalgebra's eight interleaved accumulators
        const int m = n - 1 - i;
        double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, res = 0.0;
        int q = 0;
        for (; m - q >= 8; q += 8)
            for (int k = 0; k < 8; k++) acc[k] = FB_DADD(acc[k], FB_DMUL(L[(i + 1 + q + k) * n + i], v[i + 1 + q + k]));
        res = FB_DADD(res, FB_DADD(acc[0], acc[4]));
        res = FB_DADD(res, FB_DADD(acc[1], acc[5]));
        res = FB_DADD(res, FB_DADD(acc[2], acc[6]));
        res = FB_DADD(res, FB_DADD(acc[3], acc[7]));
        for (; q < m; q++) res = FB_DADD(res, FB_DMUL(L[(i + 1 + q) * n + i], v[i + 1 + q]));
        const double diag = L[i * n + i];
        if (diag == 0.0) return false;
        v[i] = FB_DDIV(FB_DADD(v[i], -res), diag);
    }
    return true;
}

// src/lpc.rs:886-896: solve Cm xy = rv[1..P], regularising the diagonal of Cm while the factorisation fails
FB_DEV void fb_k1c_solve(double *Cm, const double *rv, int P, double *xy) {
    double L[FB200_MAX_LPC_ORDER * FB200_MAX_LPC_ORDER];
    double regularizer = 0.0;
    for (;;) {
        for (int i = 0; i < P * P; i++) L[i] = Cm[i];
        for (int i = 0; i < P; i++) xy[i] = rv[1 + i];
        if (fb_solve_sym(L, P, xy)) break;
        const double old = regularizer;
        const double twice = FB_DADD(regularizer, regularizer);
        regularizer = twice > 1.0 ? twice : 1.0;
        for (int i = 0; i < P; i++) Cm[i * P + i] = FB_DADD(Cm[i * P + i], FB_DADD(regularizer, -old));
    }
}

FB_DEV void fb_k1c_body(const FbJob &J, const int32_t *xt, const uint8_t *pcm, const float *win_full, const float *win_tail,
                        FbAnalysis *ana, fb200_variant_taps *taps_all, uint32_t gv, int tile, uint8_t *smem) {
    const int P = J.cfg.lpc_order;
    const int T = fb_k1c_threads(P);
    const uint32_t f = gv / (uint32_t)J.nvar;
    const int v = (int)(gv - f * (uint32_t)J.nvar);
    const FbK1Var V = fb_k1_var(J, f, v, win_full, win_tail);
    const int n = V.n;
    FbAnalysis *out = ana + gv;
    fb200_variant_taps *taps = taps_all ? taps_all + gv : nullptr;
    if (!V.do_lpc) return; // K1 has stored the "no LPC" record already
    const FbVarRows rows = fb_variant_rows(J, xt, f, v);
    int32_t mb = 0, sh = 0;
    if (pcm) fb_pair_mix(v, &mb, &sh);
    double *ys = (double *)smem;                                   // ys[FB200_MAX_LPC_ORDER + (t - tile0)] = y[t]
    double *accs = ys + FB200_MAX_LPC_ORDER + tile;                // [T]
    double *Cm = accs + T;                                         // [P * P]
    double *rv = Cm + P * P;                                       // [P + 1]
    double *xy = rv + P + 1;                                       // [P + 1]
    const int ncov = P * (P + 1) / 2;
#if FB_GPU
    double acc = 0.0;
#define FB_K1C_ACC acc
#else
#define FB_K1C_ACC accs[tid]
    FB_PHASE(tid, T)
        accs[tid] = 0.0;
    FB_PHASE_END
#endif
    for (int tile0 = 0; tile0 < n; tile0 += tile) {
        const int tile1 = tile0 + tile < n ? tile0 + tile : n;
        FB_PHASE(tid, T)
            // history: the last P samples of the previous tile (zeros before the frame: never used, the sums start at P-1)
            if (tid < FB200_MAX_LPC_ORDER) ys[tid] = tile0 == 0 ? 0.0 : ys[tile + tid];
        FB_PHASE_END
        FB_PHASE(tid, T)
            for (int t = tile0 + tid; t < tile1; t += T) {
                const int32_t x = fb_k1c_sample(J, xt, pcm, rows, mb, sh, f, t);
                ys[FB200_MAX_LPC_ORDER + (t - tile0)] = (double)FB_FMUL((float)x, V.win[t]);
            }
        FB_PHASE_END
        FB_PHASE(tid, T)
            int a = 0, b = 0, t_lo, t_hi; // the chain adds y[t - a] * y[t - b] for t in [t_lo, t_hi)
            bool active = true;
            if (tid < ncov) {
                // pair (i, j), i <= j, in row-major order of the upper triangle
                int i = 0, rem = tid;
                while (rem >= P - i) { rem -= P - i; i++; }
                a = i; b = i + rem;
                t_lo = P - 1; t_hi = n - 1;
            } else if (tid < ncov + P + 1) {
                a = tid - ncov; b = 0;
                t_lo = P; t_hi = n;
            } else {
                active = false; t_lo = t_hi = 0;
            }
            if (active) {
                const int lo = t_lo > tile0 ? t_lo : tile0, hi = t_hi < tile1 ? t_hi : tile1;
                const double *pa = ys + FB200_MAX_LPC_ORDER - tile0 - a, *pb = ys + FB200_MAX_LPC_ORDER - tile0 - b;
                double s = FB_K1C_ACC;
#pragma unroll 8
                for (int t = lo; t < hi; t++) s = FB_FMA(pa[t], pb[t], s);
                FB_K1C_ACC = s;
            }
        FB_PHASE_END
    }
    FB_PHASE(tid, T)
        if (tid < ncov) {
            int i = 0, rem = tid;
            while (rem >= P - i) { rem -= P - i; i++; }
            const int j = i + rem;
            Cm[i * P + j] = FB_K1C_ACC;
            Cm[j * P + i] = FB_K1C_ACC;
        } else if (tid < ncov + P + 1) {
            rv[tid - ncov] = FB_K1C_ACC;
        }
    FB_PHASE_END
#undef FB_K1C_ACC
    FB_PHASE(tid, T)
        if (tid == 0) {
            if (taps) {
                for (int i = 0; i <= FB200_MAX_LPC_ORDER; i++) taps->autocorr[i] = i <= P ? rv[i] : 0.0;
            }
            double lpc[FB200_MAX_LPC_ORDER];
            fb_k1c_solve(Cm, rv, P, xy);
            for (int i = 0; i < FB200_MAX_LPC_ORDER; i++) lpc[i] = i < P ? xy[i] : 0.0;
            int16_t q[32];
            int shift;
            const int order = fb_quantize(lpc, P, J.cfg.quant_precision, q, &shift);
            out->qlp_order = order;
            out->qlp_shift = shift;
            for (int i = 0; i < 32; i++) out->qlp[i] = i < order ? q[i] : (int16_t)0;
            if (taps) {
                for (int i = 0; i < FB200_MAX_LPC_ORDER; i++) taps->lpc[i] = lpc[i];
                for (int i = 0; i < 32; i++) taps->qlp[i] = out->qlp[i];
                taps->qlp_order = order;
                taps->qlp_shift = shift;
            }
        }
    FB_PHASE_END
}

// =================================================================================================
// K1I: the `experimental` IRLS-MAE refinement of the direct-MSE estimate (src/lpc.rs:814-850 lpc_with_irls_mae; config
// key qlpc.mae_optimization_steps > 0 with use_direct_mse), CTA per channel variant.  steps + 1 weighted solutions:
//   solution k:  r[tau]  += y[u - tau] * f32(w_k[u] * y[u]),          u = P .. n-1              (src/lpc.rs:533-548, VecWeight)
//                C[i][j] += y[u - 1 - i] * f32(w_k[u] * y[u - 1 - j]), u = P .. n-1, i <= j < P   (:573-600, ShiftedWeight<1>)
//                (u = t + 1 of the reference's loop over the signal without its last sample), chains as in K1C
//   raw error:   e_k[t] = fma chain in f32 over the taps of -(f32)x[t] + sum_j f32(a_k[j]) * (f32)x[t - 1 - j]   (:606-618)
//   score:       S_k = sequential f32 sum of |e_k[t]|; the smallest wins, the earliest on ties                   (:839-843)
//   weights:     w_0 = 1;  w_{k+1}[t] = powf(max(max(|e_k[t]|, 1) / max|x|, 0.01), -1.2) for t >= P               (:827-828, :845-847)
// Nothing per sample is stored between solutions: pass k over the frame (k = 0 .. steps + 1) recomputes e_{k-1} from the
// f32 coefficients kept in shared memory, which gives both the weights of solution k and, summed by one otherwise idle
// thread while the chain threads work, the score of solution k - 1.  Pass steps + 1 only scores.  powf is glibc's
// (fb_powf_pos, exhaustively equal for this exponent); weights of 1 make pass 0 the plain direct-MSE estimate.
// =================================================================================================
#define FB_K1I_TILE 1024
FB_HD int fb_k1i_threads(int P) { return (fb_k1c_chains(P) + 1 + 31) & ~31; }
FB_HD uint32_t fb_k1i_smem_bytes(int P, int tile) {
    const int T = fb_k1i_threads(P);
    return (uint32_t)((FB200_MAX_LPC_ORDER + tile) * 16 + tile * 8 + T * 8 + (P * P + 2 * (P + 1)) * 8 +
                      FB200_MAX_LPC_ORDER * 20 + T * 4 + 64);
}

FB_DEV float fb_irls_weight(float err, float normalizer) {
    float a = fabsf(err);
    a = a > 1.0f ? a : 1.0f; // f32::max: a NaN error counts as 1
    float r = FB_FDIV(a, normalizer);
    r = r > 0.01f ? r : 0.01f;
    return fb_powf_pos(r, -1.2f);
}

FB_DEV void fb_k1i_body(const FbJob &J, const int32_t *xt, const uint8_t *pcm, const float *win_full, const float *win_tail,
                        FbAnalysis *ana, fb200_variant_taps *taps_all, uint32_t gv, int tile, uint8_t *smem) {
    constexpr int H = FB200_MAX_LPC_ORDER;
    const int P = J.cfg.lpc_order;
    const int steps = J.cfg.mae_optimization_steps;
    const int T = fb_k1i_threads(P);
    const uint32_t f = gv / (uint32_t)J.nvar;
    const int v = (int)(gv - f * (uint32_t)J.nvar);
    const FbK1Var V = fb_k1_var(J, f, v, win_full, win_tail);
    const int n = V.n;
    FbAnalysis *out = ana + gv;
    fb200_variant_taps *taps = taps_all ? taps_all + gv : nullptr;
    if (!V.do_lpc) return; // K1 has stored the "no LPC" record already
    const FbVarRows rows = fb_variant_rows(J, xt, f, v);
    int32_t mb = 0, sh = 0;
    if (pcm) fb_pair_mix(v, &mb, &sh);
    double *ys = (double *)smem;                    // ys[H + (t - tile0)] = (f64)y[t]
    double *accs = ys + H + tile;                   // [T]
    double *Cm = accs + T;                          // [P * P]
    double *rv = Cm + P * P;                        // [P + 1]
    double *xy = rv + P + 1;                        // [P + 1]
    double *lpc_cur = xy + P + 1;                   // [H] solution k
    double *lpc_best = lpc_cur + H;                 // [H]
    float *yf = (float *)(lpc_best + H);            // [H + tile] y[t]
    float *xf = yf + H + tile;                      // [H + tile] (f32)x[t]
    float *wt = xf + H + tile;                      // [tile] w_k[t]
    float *ae = wt + tile;                          // [tile] |e_{k-1}[t]|
    float *cf = ae + tile;                          // [H] f32(a_{k-1}[j])
    int32_t *red = (int32_t *)(cf + H);             // [T] peak reduction
    float *sc = (float *)(red + T);                 // [0] running score, [1] best score, [2] normalizer, [3] have_best
    const int ncov = P * (P + 1) / 2, nch = fb_k1c_chains(P);

    // normalizer = (f32) max |x[t]| (src/lpc.rs:827)
    FB_PHASE(tid, T)
        int32_t m = 0;
        for (int t = tid; t < n; t += T) {
            const int32_t x = fb_k1c_sample(J, xt, pcm, rows, mb, sh, f, t);
            const int32_t a = x < 0 ? -x : x;
            m = a > m ? a : m;
        }
        red[tid] = m;
    FB_PHASE_END
    FB_PHASE(tid, T)
        if (tid == 0) {
            int32_t m = 0;
            for (int i = 0; i < T; i++) m = red[i] > m ? red[i] : m;
            sc[0] = 0.0f;
            sc[1] = 3.402823466e+38f; // f32::MAX
            sc[2] = (float)m;
            sc[3] = 0.0f;
            for (int i = 0; i < H; i++) { lpc_cur[i] = 0.0; lpc_best[i] = 0.0; cf[i] = 0.0f; }
        }
    FB_PHASE_END

    for (int k = 0; k <= steps + 1; k++) {
#if FB_GPU
        double acc = 0.0;
#define FB_K1I_ACC acc
#else
#define FB_K1I_ACC accs[tid]
        FB_PHASE(tid, T)
            accs[tid] = 0.0;
        FB_PHASE_END
#endif
        for (int tile0 = 0; tile0 < n; tile0 += tile) {
            const int tile1 = tile0 + tile < n ? tile0 + tile : n;
            FB_PHASE(tid, T)
                // history: the last H samples of the previous tile (zeros before the frame are never used)
                if (tid < H) {
                    ys[tid] = tile0 == 0 ? 0.0 : ys[tile + tid];
                    yf[tid] = tile0 == 0 ? 0.0f : yf[tile + tid];
                    xf[tid] = tile0 == 0 ? 0.0f : xf[tile + tid];
                }
            FB_PHASE_END
            FB_PHASE(tid, T)
                for (int t = tile0 + tid; t < tile1; t += T) {
                    const float x = (float)fb_k1c_sample(J, xt, pcm, rows, mb, sh, f, t);
                    const float y = FB_FMUL(x, V.win[t]);
                    xf[H + (t - tile0)] = x;
                    yf[H + (t - tile0)] = y;
                    ys[H + (t - tile0)] = (double)y;
                }
            FB_PHASE_END
            if (k > 0) {
                FB_PHASE(tid, T)
                    const float normalizer = sc[2];
                    for (int t = tile0 + tid; t < tile1; t += T) {
                        float a = 0.0f, w = 1.0f;
                        if (t >= P) {
                            const float *xp = xf + H + (t - tile0);
                            float e = xp[0] == 0.0f ? 0.0f : -xp[0]; // (f32)(-x[t])
                            for (int j = 0; j < P; j++) e = FB_FMAF(cf[j], xp[-1 - j], e);
                            a = fabsf(e);
                            w = fb_irls_weight(e, normalizer);
                        }
                        ae[t - tile0] = a;
                        wt[t - tile0] = w;
                    }
                FB_PHASE_END
            }
            FB_PHASE(tid, T)
                if (tid < nch && k <= steps) {
                    int a, b; // the chain adds y[u - a] * f32(w[u] * y[u - b]) for u in [P, n)
                    if (tid < ncov) {
                        // pair (i, j), i <= j, in row-major order of the upper triangle
                        int i = 0, rem = tid;
                        while (rem >= P - i) { rem -= P - i; i++; }
                        a = i + 1; b = i + rem + 1;
                    } else {
                        a = tid - ncov; b = 0;
                    }
                    const int lo = P > tile0 ? P : tile0, hi = tile1;
                    const double *pa = ys + H - tile0 - a;
                    double s = FB_K1I_ACC;
                    if (k == 0) { // weights of 1: the products are exact
                        const double *pb = ys + H - tile0 - b;
#pragma unroll 8
                        for (int u = lo; u < hi; u++) s = FB_FMA(pa[u], pb[u], s);
                    } else {
                        const float *pb = yf + H - tile0 - b, *pw = wt - tile0;
#pragma unroll 8
                        for (int u = lo; u < hi; u++) s = FB_FMA(pa[u], (double)FB_FMUL(pw[u], pb[u]), s);
                    }
                    FB_K1I_ACC = s;
                } else if (tid == nch && k > 0) {
                    float s = sc[0];
                    for (int t = tile0; t < tile1; t++) s = FB_FADD(s, ae[t - tile0]);
                    sc[0] = s;
                }
            FB_PHASE_END
        }
        FB_PHASE(tid, T)
            if (k <= steps) {
                if (tid < ncov) {
                    int i = 0, rem = tid;
                    while (rem >= P - i) { rem -= P - i; i++; }
                    const int j = i + rem;
                    Cm[i * P + j] = FB_K1I_ACC;
                    Cm[j * P + i] = FB_K1I_ACC;
                } else if (tid < nch) {
                    rv[tid - ncov] = FB_K1I_ACC;
                }
            }
        FB_PHASE_END
#undef FB_K1I_ACC
        FB_PHASE(tid, T)
            if (tid == 0) {
                if (k > 0 && sc[0] < sc[1]) { // src/lpc.rs:840-843: solution k - 1 is the best so far
                    sc[1] = sc[0];
                    sc[3] = 1.0f;
                    for (int i = 0; i < H; i++) lpc_best[i] = lpc_cur[i];
                }
                sc[0] = 0.0f;
                if (k <= steps) {
                    if (k == 0 && taps) {
                        for (int i = 0; i <= FB200_MAX_LPC_ORDER; i++) taps->autocorr[i] = i <= P ? rv[i] : 0.0;
                    }
                    fb_k1c_solve(Cm, rv, P, xy);
                    for (int i = 0; i < H; i++) {
                        lpc_cur[i] = i < P ? xy[i] : 0.0;
                        cf[i] = (float)lpc_cur[i];
                    }
                }
            }
        FB_PHASE_END
    }
    FB_PHASE(tid, T)
        if (tid == 0) {
            // (the reference unwraps None when no score was below f32::MAX; the last solution stands in here)
            const double *lpc = sc[3] != 0.0f ? lpc_best : lpc_cur;
            double lq[FB200_MAX_LPC_ORDER];
            for (int i = 0; i < FB200_MAX_LPC_ORDER; i++) lq[i] = lpc[i];
            int16_t q[32];
            int shift;
            const int order = fb_quantize(lq, P, J.cfg.quant_precision, q, &shift);
            out->qlp_order = order;
            out->qlp_shift = shift;
            for (int i = 0; i < 32; i++) out->qlp[i] = i < order ? q[i] : (int16_t)0;
            if (taps) {
                for (int i = 0; i < FB200_MAX_LPC_ORDER; i++) taps->lpc[i] = lq[i];
                for (int i = 0; i < 32; i++) taps->qlp[i] = out->qlp[i];
                taps->qlp_order = order;
                taps->qlp_shift = shift;
            }
        }
    FB_PHASE_END
}

#if FB_GPU
// K1D: the same estimator for the default lpc_order (10) and LARGE launches, thread per channel variant like K1: the 55
// covariance chains and the 11 autocorrelation chains of a variant live in the registers of one thread next to a ring of
// its last 11 windowed samples, so a sample costs one shared-memory read and 66 DFMAs instead of 66 x (2 reads + 1 DFMA)
// spread over a CTA (K1C).  Chains are summed in the same order: results are bit-identical to K1C.
#define FB_K1D_P 10
template <bool PAIRS>
FB_DEV void fb_k1d_warp(const FbJob &J, const int32_t *xt, const uint8_t *pcm, const float *win_full, const float *win_tail,
                        FbAnalysis *ana, fb200_variant_taps *taps_all, uint32_t n_variants, uint8_t *smem) {
    constexpr int P = FB_K1D_P;
    FbK1Lane W;
    if (!fb_k1_lane_setup<PAIRS>(J, xt, pcm, win_full, win_tail, n_variants, blockIdx.x, smem, W)) return;
    const FbK1Var &V = W.V;
    FbAnalysis *out = ana + W.gve;
    fb200_variant_taps *taps = taps_all ? taps_all + W.gve : nullptr;
    double C[P * (P + 1) / 2], Rr[P + 1], ring[P + 1]; // ring[k] = y[t - k]
#pragma unroll
    for (int i = 0; i < P * (P + 1) / 2; i++) C[i] = 0.0;
#pragma unroll
    for (int i = 0; i <= P; i++) { Rr[i] = 0.0; ring[i] = 0.0; }
    if (J.cfg.use_lpc && W.n_w >= FB_MIN_PRED_BLOCK) {
        const int n = V.n; // (every lane of the warp: fb_k1_slots isolates a shorter last frame in a warp of its own)
        const float *win = V.win;
        fb_k1_stream<2, 4, PAIRS>(W.T, (W.n_w + 7) / 8, [&](int g, const int32_t *xs) {
            const int t0 = g * 8;
            float ws[8];
            fb_k1_win_load<8, false>(win, t0, 0, ws); // (the table is zero-padded beyond n)
            const bool mid = t0 >= P && t0 + 8 <= n - 1; // P <= t <= n - 2 for all eight: every chain takes every sample
#pragma unroll
            for (int s = 0; s < 8; s++) {
                const int t = t0 + s;
                const double y = (double)FB_FMUL((float)xs[s], ws[s]);
#pragma unroll
                for (int k = P; k >= 1; k--) ring[k] = ring[k - 1];
                ring[0] = y;
                if (mid || (t >= P && t <= n - 1)) {
#pragma unroll
                    for (int tau = 0; tau <= P; tau++) Rr[tau] = FB_FMA(ring[tau], ring[0], Rr[tau]);
                }
                if (mid || (t >= P - 1 && t <= n - 2)) {
                    int c = 0;
#pragma unroll
                    for (int i = 0; i < P; i++)
#pragma unroll
                        for (int j = i; j < P; j++, c++) C[c] = FB_FMA(ring[i], ring[j], C[c]);
                }
            }
        });
    }
    if (!W.valid || !V.do_lpc) return; // (K1 has stored the "no LPC" record already)
    // solve C a = r[1..P] (src/lpc.rs:886-896) and quantise, like K1C's last phase
    double Cm[P * P], L[P * P], xy[P], lpc[FB200_MAX_LPC_ORDER];
    {
        int c = 0;
#pragma unroll
        for (int i = 0; i < P; i++)
#pragma unroll
            for (int j = i; j < P; j++, c++) { Cm[i * P + j] = C[c]; Cm[j * P + i] = C[c]; }
    }
    double regularizer = 0.0;
    for (;;) {
        for (int i = 0; i < P * P; i++) L[i] = Cm[i];
        for (int i = 0; i < P; i++) xy[i] = Rr[1 + i];
        if (fb_solve_sym(L, P, xy)) break;
        const double old = regularizer;
        const double twice = FB_DADD(regularizer, regularizer);
        regularizer = twice > 1.0 ? twice : 1.0;
        for (int i = 0; i < P; i++) Cm[i * P + i] = FB_DADD(Cm[i * P + i], FB_DADD(regularizer, -old));
    }
    for (int i = 0; i < FB200_MAX_LPC_ORDER; i++) lpc[i] = i < P ? xy[i] : 0.0;
    int16_t q[32];
    int shift;
    const int order = fb_quantize(lpc, P, J.cfg.quant_precision, q, &shift);
    out->qlp_order = order;
    out->qlp_shift = shift;
    for (int i = 0; i < 32; i++) out->qlp[i] = i < order ? q[i] : (int16_t)0;
    if (taps) {
        for (int i = 0; i <= FB200_MAX_LPC_ORDER; i++) taps->autocorr[i] = i <= P ? Rr[i] : 0.0;
        for (int i = 0; i < FB200_MAX_LPC_ORDER; i++) taps->lpc[i] = lpc[i];
        for (int i = 0; i < 32; i++) taps->qlp[i] = out->qlp[i];
        taps->qlp_order = order;
        taps->qlp_shift = shift;
    }
}
#endif

// =================================================================================================
// K2: per-variant residual coding search and subframe decision (CTA per variant).
// =================================================================================================

// residual of sample t for a candidate: kind 0 = fixed order `order`, kind 1 = LPC.
// fixed (src/coding.rs:182-197): zero-history k-th difference == sum_j (-1)^j C(k,j) x[t-j] for t >= k.
// lpc (src/lpc.rs:306-390): x[t] - ((sum_j q[j] x[t-1-j]) >> shift); the i32 and i64 paths of the
// reference agree modulo 2^32, so the sum is always formed in 64 bits and truncated.
FB_DEV int32_t fb_residual_at(const int32_t *x, int t, int kind, int order, const int16_t *q, int shift) {
    if (t < order) return 0;
    if (kind == 0) {
        uint32_t a = (uint32_t)x[t];
        switch (order) {
        case 0: return (int32_t)a;
        case 1: return (int32_t)(a - (uint32_t)x[t - 1]);
        case 2: return (int32_t)(a - 2u * (uint32_t)x[t - 1] + (uint32_t)x[t - 2]);
        case 3: return (int32_t)(a - 3u * (uint32_t)x[t - 1] + 3u * (uint32_t)x[t - 2] - (uint32_t)x[t - 3]);
        default:
            return (int32_t)(a - 4u * (uint32_t)x[t - 1] + 6u * (uint32_t)x[t - 2] - 4u * (uint32_t)x[t - 3] +
                             (uint32_t)x[t - 4]);
        }
    }
    int64_t acc = 0;
    for (int j = 0; j < order; j++) acc += (int64_t)q[j] * (int64_t)x[t - 1 - j];
    return (int32_t)(uint32_t)((uint64_t)(int64_t)x[t] - (uint64_t)(acc >> shift));
}

// ---- run-based residuals: one thread produces the zigzag residuals of FB_RUN consecutive samples
// from a register window of the signal, so every tap is one multiply-add with static operands
// (no per-tap loads, no loop overhead).  G = number of taps evaluated (order rounded up to 4, 8, 12,
// 16 or 24); coefficients beyond `order` are zero.  Results equal fb_residual_at() sample for sample.

// u[i] for t = t0 + i, i < FB_RUN; samples at t >= n give 0.  x must be readable up to index
// ((n + 3) & ~3) - 1 (the planar stride is a multiple of 32).
template <int G>
FB_DEV void fb_run_residual_lpc(const int32_t *x, int n, int t0, const int16_t *q, int order, int shift,
                                uint32_t *u) {
    int32_t win[G + FB_RUN]; // win[i] = x[t0 - G + i]
    if (t0 >= G) {
#pragma unroll
        for (int i = 0; i < G + FB_RUN; i += 4) {
            const int4 v = *reinterpret_cast<const int4 *>(x + t0 - G + i); // t0, G multiples of 4: aligned
            win[i] = v.x; win[i + 1] = v.y; win[i + 2] = v.z; win[i + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < G + FB_RUN; i++) {
            const int idx = t0 - G + i;
            win[i] = (idx >= 0) ? x[idx] : 0;
        }
    }
    int32_t qq[G];
#pragma unroll
    for (int j = 0; j < G; j++) qq[j] = (j < order) ? (int32_t)q[j] : 0;
#pragma unroll
    for (int i = 0; i < FB_RUN; i++) {
        const int t = t0 + i;
        int64_t acc = 0;
#pragma unroll
        for (int j = 0; j < G; j++) acc = fb_mad_wide(qq[j], win[G + i - 1 - j], acc);
        const int32_t e = (int32_t)(uint32_t)((uint64_t)(int64_t)win[G + i] - (uint64_t)(acc >> shift));
        u[i] = (t >= order && t < n) ? fb_zigzag(e) : 0u;
    }
}

FB_DEV void fb_run_residual_fixed(const int32_t *x, int n, int t0, int order, uint32_t *u) {
    uint32_t win[4 + FB_RUN]; // win[i] = x[t0 - 4 + i]
    if (t0 >= 4) {
#pragma unroll
        for (int i = 0; i < 4 + FB_RUN; i += 4) {
            const int4 v = *reinterpret_cast<const int4 *>(x + t0 - 4 + i);
            win[i] = (uint32_t)v.x; win[i + 1] = (uint32_t)v.y; win[i + 2] = (uint32_t)v.z; win[i + 3] = (uint32_t)v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4 + FB_RUN; i++) {
            const int idx = t0 - 4 + i;
            win[i] = (idx >= 0) ? (uint32_t)x[idx] : 0u;
        }
    }
#pragma unroll
    for (int i = 0; i < FB_RUN; i++) {
        const int t = t0 + i;
        const uint32_t a = win[4 + i], b = win[3 + i], c = win[2 + i], d = win[1 + i], e4 = win[i];
        uint32_t e;
        switch (order) {
        case 0: e = a; break;
        case 1: e = a - b; break;
        case 2: e = a - 2u * b + c; break;
        case 3: e = a - 3u * b + 3u * c - d; break;
        default: e = a - 4u * b + 6u * c - 4u * d + e4; break;
        }
        u[i] = (t >= order && t < n) ? fb_zigzag((int32_t)e) : 0u;
    }
}

// G >= order is fixed per kernel instantiation (fb_k1_ring(cfg.lpc_order)), so the register allocation
// of the default order-10 build does not pay for the order-24 window.
template <int G>
FB_DEV void fb_run_residual(const int32_t *x, int n, int t0, int kind, int order, const int16_t *q, int shift,
                            uint32_t *u) {
    if (kind == 0) fb_run_residual_fixed(x, n, t0, order, u);
    else fb_run_residual_lpc<G>(x, n, t0, q, order, shift, u);
}

// index of u[t] in shared memory: one pad word per 32 keeps both the per-sample (stride 1) and the
// per-run (stride 16) access patterns free of bank conflicts
FB_HD int fb_uidx(int t) { return t + (t >> 5); }

// shared-memory layout of K2 (offsets in bytes), identical on host and device
struct FbK2Layout {
    uint32_t off_u, off_tbl_a, off_tbl_b, off_part, off_lvl_params, off_lvl_bits, off_unit_sum, off_misc, total;
    int32_t u_shift;
};

// index of residual t in K2's buffer: one pad word per 32 (u_shift 5) keeps strided readers off one bank; the few large
// block sizes with 256 or more finest partitions (30720 .. 32512) only fit without the padding (u_shift 31)
FB_HD int fb_uidx2(int t, int u_shift) { return t + (t >> u_shift); }
FB_HD FbK2Layout fb_k2_layout(int n_max, int leaves_max, int u_shift = 5) {
    FbK2Layout L;
    uint32_t o = 0;
    L.u_shift = u_shift;
    L.off_u = o;          o += (uint32_t)((n_max + (n_max >> u_shift) + 4 + 3) & ~3) * 4u; // padded, see fb_uidx2
    L.off_tbl_a = o;      o += (uint32_t)leaves_max * 32u * 4u;
    L.off_tbl_b = o;      o += (uint32_t)leaves_max * 32u * 4u;
    uint32_t units = (uint32_t)(leaves_max > FB_K2_THREADS ? leaves_max : FB_K2_THREADS);
    L.off_part = o;       o += units * 32u * 4u;
    L.off_unit_sum = o;   o += units * 8u;
    L.off_lvl_bits = o;   o += 16u * 8u;
    L.off_lvl_params = o; o += 2u * (uint32_t)leaves_max;
    o = (o + 15u) & ~15u;
    L.off_misc = o;       o += 256u;
    L.total = o;
    return L;
}

// result of one residual search, kept in shared memory between candidates
struct FbRiceResult {
    int32_t  part_order;
    int32_t  rice2;
    uint64_t res_bits; // Residual::count_bits()
    uint64_t code_bits;
    uint8_t  params[FB200_MAX_RICE_PARTS];
};

struct FbK2Misc {
    uint32_t maxu;
    uint32_t pmin, pmax;
    uint32_t fail;
    int32_t  win_a, win_b;
    int32_t  mode;
    int32_t  best_level;
    unsigned long long sum_q;
    uint32_t any_gt14;
    uint32_t pad;
};

// Partitioned Rice parameter search for one candidate (src/rice.rs:246-298) + Residual::count_bits
// (src/component/bitrepr.rs:532-544).
//
// The reference evaluates, for every finest partition, the cost of all 31 parameters
// (bits[p] = sum(u >> p) + len*(p+1) + 4, saturated to 2^27-1) and merges tables bottom-up.
// bits[p] is convex in p, so the minimiser of every node of the partition tree lies between the
// smallest and the largest leaf minimiser.  The kernel therefore evaluates only a window [a, b] of
// parameters around per-leaf estimates, then *verifies* on every leaf that the window brackets the
// minimum (bits[a] > bits[a+1] or a == 0; bits[b-1] <= bits[b] or b == max_p; nothing saturated).
// If the check fails the search is repeated over the full range, which reproduces the reference's
// tables entry for entry.  When a residual is >= 2^27 the reference's 16-sample chunked saturating
// u32 accumulation (src/rice.rs:75-98) is replayed literally (mode 2).  Either way the chosen
// partition order, parameters and bit count are identical to the reference's.
template <int G>
FB_DEV void fb_k2_rice_search(const FbJob &J, const int32_t *x, int n, int kind, int order, const int16_t *q,
                              int shift, uint8_t *smem, const FbK2Layout &L, FbRiceResult *res) {
    const int T = FB_K2_THREADS;
    uint32_t *u = (uint32_t *)(smem + L.off_u);
    uint32_t *tbl_a = (uint32_t *)(smem + L.off_tbl_a);
    uint32_t *tbl_b = (uint32_t *)(smem + L.off_tbl_b);
    uint32_t *part = (uint32_t *)(smem + L.off_part);
    unsigned long long *unit_sum = (unsigned long long *)(smem + L.off_unit_sum);
    unsigned long long *lvl_bits = (unsigned long long *)(smem + L.off_lvl_bits);
    uint8_t *lvl_params = smem + L.off_lvl_params;
    FbK2Misc *M = (FbK2Misc *)(smem + L.off_misc);

    const int warm = order;
    const int max_p = J.cfg.prc_max_parameter;
    const int o0 = fb_finest_partition_order(n);
    const int leaves = 1 << o0;
    const int leaf_len = n >> o0;
    int nsub = 1;
    while (leaves * nsub * 2 <= T && (leaf_len / (nsub * 2)) >= 16) nsub *= 2;
    const int units = leaves * nsub;

    // ---- phase 1: residual -> zigzag u[] (warm-up slots 0), block maximum
    FB_PHASE(tid, T)
        if (tid == 0) { M->maxu = 0; M->pmin = 31; M->pmax = 0; M->fail = 0; M->sum_q = 0; M->any_gt14 = 0; }
    FB_PHASE_END
    FB_PHASE(tid, T)
        uint32_t mx = 0;
        const int nruns = (n + FB_RUN - 1) / FB_RUN;
        for (int run = tid; run < nruns; run += T) {
            uint32_t uu[FB_RUN];
            fb_run_residual<G>(x, n, run * FB_RUN, kind, order, q, shift, uu);
#pragma unroll
            for (int i = 0; i < FB_RUN; i++) {
                const int t = run * FB_RUN + i;
                if (t < n) u[fb_uidx2(t, L.u_shift)] = uu[i];
                mx = uu[i] > mx ? uu[i] : mx;
            }
        }
        if (mx) fb_atomic_max_u32(&M->maxu, mx);
    FB_PHASE_END

    // ---- phase 2: per-unit sums of u (unit = contiguous piece of a leaf).  Integer sums do not
    // depend on the order, so every lane starts at a different offset of its piece (no bank conflicts).
    FB_PHASE(tid, T)
        for (int unit = tid; unit < units; unit += T) {
            int leaf = unit / nsub, sub = unit - leaf * nsub;
            int lstart = leaf * leaf_len;
            int a0 = lstart + (int)(((long long)leaf_len * sub) / nsub);
            int a1 = lstart + (int)(((long long)leaf_len * (sub + 1)) / nsub);
            if (a0 < warm) a0 = warm;
            const int m = a1 - a0;
            int t = a0 + (m > 0 ? (tid % m) : 0);
            unsigned long long sacc = 0;
            for (int i = 0; i < m; i++) {
                sacc += u[fb_uidx2(t, L.u_shift)];
                t = (t + 1 < a1) ? t + 1 : a0;
            }
            unit_sum[unit] = sacc;
        }
    FB_PHASE_END

    // ---- phase 3: per-leaf parameter estimate floor(log2(mean)) -> window bounds
    FB_PHASE(tid, T)
        for (int leaf = tid; leaf < leaves; leaf += T) {
            unsigned long long sacc = 0;
            for (int s = 0; s < nsub; s++) sacc += unit_sum[leaf * nsub + s];
            int cnt = leaf_len - (leaf == 0 ? warm : 0);
            unsigned long long mean = cnt > 0 ? sacc / (unsigned long long)cnt : 0;
            uint32_t pe = 0;
            while (pe < 31 && (mean >> (pe + 1)) != 0) pe++;
            fb_atomic_min_u32(&M->pmin, pe);
            fb_atomic_max_u32(&M->pmax, pe);
        }
    FB_PHASE_END
    FB_PHASE(tid, T)
        if (tid == 0) {
            int a = (int)M->pmin - 2, b = (int)M->pmax + 1;
            if (b > max_p) b = max_p;
            if (a > b - 1) a = b - 1;
            if (a < 0) a = 0;
            int mode = 0;
            if (M->maxu >= (1u << 27)) { mode = 2; a = 0; b = max_p; }
            M->win_a = a; M->win_b = b; M->mode = mode;
        }
    FB_PHASE_END

    for (int attempt = 0; attempt < 2; attempt++) {
        const int a = M->win_a, b = M->win_b, mode = M->mode;
        const int W = b - a + 1;
#if !FB_GPU
        fb_emu_mode_count[mode]++; // emulation-only statistics: which search mode ran
        fb_emu_mode_count[3] += (unsigned long long)W;
#endif

        if (mode != 2) {
            // ---- phase 4: per-unit partial sums S(p) = sum(u >> p), p in [a, b]
            FB_PHASE(tid, T)
                for (int unit = tid; unit < units; unit += T) {
                    int leaf = unit / nsub, sub = unit - leaf * nsub;
                    int lstart = leaf * leaf_len;
                    int a0 = lstart + (int)(((long long)leaf_len * sub) / nsub);
                    int a1 = lstart + (int)(((long long)leaf_len * (sub + 1)) / nsub);
                    if (a0 < warm) a0 = warm;
                    const int m = a1 - a0;
                    const int tstart = a0 + (m > 0 ? (tid % m) : 0);
                    for (int pc = 0; pc < W; pc += 4) {
                        unsigned long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
                        const int p0 = a + pc;
                        int t = tstart;
                        for (int i = 0; i < m; i++) {
                            uint32_t v = u[fb_uidx2(t, L.u_shift)] >> p0;
                            c0 += v; c1 += v >> 1; c2 += v >> 2; c3 += v >> 3;
                            t = (t + 1 < a1) ? t + 1 : a0;
                        }
                        unsigned long long c[4] = {c0, c1, c2, c3};
                        for (int k = 0; k < 4 && pc + k < W; k++)
                            part[unit * 32 + pc + k] = c[k] > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)c[k];
                    }
                }
            FB_PHASE_END
            // ---- phase 5: leaf tables  min(S + 4 + cnt*(p+1), 2^27-1)
            FB_PHASE(tid, T)
                for (int e = tid; e < leaves * W; e += T) {
                    int leaf = e / W, j = e - leaf * W;
                    unsigned long long sacc = 0;
                    for (int s = 0; s < nsub; s++) sacc += part[(leaf * nsub + s) * 32 + j];
                    int cnt = leaf_len - (leaf == 0 ? warm : 0);
                    sacc += 4ull + (unsigned long long)cnt * (unsigned long long)(a + j + 1);
                    tbl_a[leaf * 32 + j] = sacc > FB_RICE_SAT ? FB_RICE_SAT : (uint32_t)sacc;
                }
            FB_PHASE_END
        } else {
            // ---- mode 2: literal replay of PrcBitTable::from_errors (src/rice.rs:65-105):
            // u32 wrapping adds, clamp after every 16 samples, offset added at the end
            FB_PHASE(tid, T)
                for (int e = tid; e < leaves * W; e += T) {
                    int leaf = e / W, j = e - leaf * W;
                    int p = a + j;
                    int start = leaf * leaf_len;
                    int end = start + leaf_len;
                    if (start < warm) start = warm;
                    uint32_t accv = 0;
                    for (int c = start; c < end; c += 16) {
                        int ce = c + 16 < end ? c + 16 : end;
                        for (int t = c; t < ce; t++) accv += u[fb_uidx2(t, L.u_shift)] >> p;
                        accv = accv < FB_RICE_SAT ? accv : FB_RICE_SAT;
                    }
                    accv += 4u + (uint32_t)(end - start) * (uint32_t)(p + 1);
                    tbl_a[leaf * 32 + j] = accv < FB_RICE_SAT ? accv : FB_RICE_SAT;
                }
            FB_PHASE_END
        }

        // ---- phase 6: bracket verification on the leaves (narrow window only)
        if (mode == 0) {
            FB_PHASE(tid, T)
                for (int leaf = tid; leaf < leaves; leaf += T) {
                    const uint32_t *row = tbl_a + leaf * 32;
                    bool ok = true;
                    for (int j = 0; j < W; j++) ok = ok && (row[j] < FB_RICE_SAT);
                    if (a > 0) ok = ok && (W >= 2) && (row[0] > row[1]);
                    if (b < max_p) ok = ok && (W >= 2) && (row[W - 2] <= row[W - 1]);
                    if (!ok) M->fail = 1; // benign race: every writer stores 1
                }
            FB_PHASE_END
        }

        // ---- phase 7: tree search, finest level first (src/rice.rs:276-291)
        if (!(mode == 0 && M->fail)) {
            uint32_t *cur = tbl_a, *nxt = tbl_b;
            int pbase = 0;
            for (int lvl = o0; lvl >= 0; lvl--) {
                const int nodes = 1 << lvl;
                FB_PHASE(tid, T)
                    if (tid == 0) lvl_bits[lvl] = 0;
                FB_PHASE_END
                FB_PHASE(tid, T)
                    unsigned long long local = 0;
                    for (int node = tid; node < nodes; node += T) {
                        const uint32_t *row = cur + node * 32;
                        // minimizer: smallest (bits << 5 | p) (src/rice.rs:117-141)
                        uint32_t best = row[0], bp = (uint32_t)a;
                        for (int j = 1; j < W; j++)
                            if (row[j] < best) { best = row[j]; bp = (uint32_t)(a + j); }
                        lvl_params[pbase + node] = (uint8_t)bp;
                        local += best;
                        if (mode == 0 && best >= FB_RICE_SAT) M->fail = 1;
                    }
                    if (local) fb_atomic_add_u64(&lvl_bits[lvl], local);
                FB_PHASE_END
                if (lvl > 0) {
                    // merge pairs: min(a + b - 4, 2^27-1) with u32 wrapping (src/rice.rs:144-152)
                    FB_PHASE(tid, T)
                        for (int e = tid; e < (nodes / 2) * W; e += T) {
                            int node = e / W, j = e - node * W;
                            uint32_t v = cur[(2 * node) * 32 + j] + cur[(2 * node + 1) * 32 + j] - 4u;
                            if (mode == 0 && v >= FB_RICE_SAT) M->fail = 1;
                            nxt[node * 32 + j] = v < FB_RICE_SAT ? v : FB_RICE_SAT;
                        }
                    FB_PHASE_END
                    uint32_t *tmp = cur; cur = nxt; nxt = tmp;
                }
                pbase += nodes;
            }
        }

        // every thread must have read the flag before thread 0 clears it below
        const bool failed = (mode == 0) && (M->fail != 0);
        FB_SYNC();
        if (failed) {
            // widen to the full parameter range and repeat (exact by construction)
            FB_PHASE(tid, T)
                if (tid == 0) { M->win_a = 0; M->win_b = max_p; M->mode = 1; M->fail = 0; }
            FB_PHASE_END
            continue;
        }
        break;
    }

    // ---- phase 8: pick the partition order: strictly smaller total wins while going coarser
    FB_PHASE(tid, T)
        if (tid == 0) {
            unsigned long long min_bits = lvl_bits[o0];
            int best = o0;
            for (int lvl = o0 - 1; lvl >= 0; lvl--)
                if (lvl_bits[lvl] < min_bits) { min_bits = lvl_bits[lvl]; best = lvl; }
            M->best_level = best;
            res->part_order = best;
            res->code_bits = min_bits;
        }
    FB_PHASE_END
    {
        const int best = M->best_level;
        int pbase = 0;
        for (int lvl = o0; lvl > best; lvl--) pbase += 1 << lvl;
        const int nparts = 1 << best;
        const int plen = n >> best;
        FB_PHASE(tid, T)
            for (int j = tid; j < nparts; j += T) {
                uint8_t p = lvl_params[pbase + j];
                res->params[j] = p;
                if (p > 14) M->any_gt14 = 1;
            }
            // exact quotient sum for Residual::count_bits (src/component/datatype.rs:2328-2335)
            unsigned long long local = 0;
            const float inv_plen = 1.0f / (float)plen;
            for (int t = tid; t < n; t += T) {
                // partition index t / plen without an integer division (t < 2^15: one correction step each way)
                int pj = (int)((float)t * inv_plen);
                if ((pj + 1) * plen <= t) pj++;
                if (pj * plen > t) pj--;
                if (t >= warm) local += u[fb_uidx2(t, L.u_shift)] >> lvl_params[pbase + pj];
            }
            if (local) fb_atomic_add_u64(&M->sum_q, local);
        FB_PHASE_END
        FB_PHASE(tid, T)
            if (tid == 0) {
                unsigned long long sum_p = 0;
                for (int j = 0; j < nparts; j++) sum_p += res->params[j];
                int rice2 = M->any_gt14 ? 1 : 0;
                res->rice2 = rice2;
                // src/component/bitrepr.rs:532-544
                res->res_bits = 2 + 4 + (unsigned long long)nparts * (rice2 ? 5 : 4) + M->sum_q +
                                (unsigned long long)(n - warm) + sum_p * (unsigned long long)plen -
                                (unsigned long long)warm * res->params[0];
            }
        FB_PHASE_END
    }
}

// K2 body: one CTA per channel variant.  Implements fixed_lpc / estimated_qlpc / encode_subframe
// (src/coding.rs:298-418) on top of K1's analysis.
template <int G>
FB_DEV void fb_k2_body(const FbJob &J, const int32_t *xv, const FbAnalysis *ana, fb200_subframe_info *choice,
                       uint32_t gv, uint8_t *smem, const FbK2Layout &L) {
    const int T = FB_K2_THREADS;
    const uint32_t f = gv / (uint32_t)J.nvar;
    const int v = (int)(gv - f * (uint32_t)J.nvar);
    const int n = fb_frame_len(J, f);
    const int bps_v = fb_variant_bps(J, v);
    const int32_t *x = xv + (size_t)gv * (size_t)J.stride;
    const FbAnalysis &A = ana[gv];
    fb200_subframe_info *out = choice + gv;
    FbRiceResult *res_fixed = (FbRiceResult *)(smem + L.total);
    FbRiceResult *res_lpc = res_fixed + 1;
    FbRiceResult *res_tmp = res_fixed + 2;
    const unsigned long long verbatim_bits = 8ull + (unsigned long long)n * (unsigned long long)bps_v;
    const bool too_short = n < FB_MIN_PRED_BLOCK;

    // common fields
    FB_PHASE(tid, T)
        if (tid == 0) {
            out->type = FB200_SF_VERBATIM;
            out->order = 0;
            out->bits_per_sample = bps_v;
            out->precision = 0;
            out->shift = 0;
            out->partition_order = 0;
            out->rice2 = 0;
            out->reserved = n;
            out->bits = verbatim_bits;
        }
        for (int i = tid; i < 32; i += T) out->qlp[i] = 0;
    FB_PHASE_END

    if (J.cfg.use_constant && A.is_constant) {
        FB_PHASE(tid, T)
            if (tid == 0) { out->type = FB200_SF_CONSTANT; out->bits = 8ull + (unsigned long long)bps_v; }
        FB_PHASE_END
        return;
    }
    if (too_short) return;

    // ---- fixed candidate (src/coding.rs:298-331)
    int kf = -1;
    unsigned long long fixed_bits = 0;
    if (J.cfg.use_fixed) {
        if (J.cfg.fixed_order_sel == 1) {
            kf = A.fixed_order;
            if (kf >= 0) fb_k2_rice_search<G>(J, x, n, 0, kf, nullptr, 0, smem, L, res_fixed);
        } else {
            // OrderSel::BitCount: exact Rice search per order, key = bps*order + code_bits,
            // first minimum wins (src/coding.rs:241-262)
            const int n_orders = (J.cfg.fixed_max_order < 4 ? J.cfg.fixed_max_order : 4) + 1;
            unsigned long long best_key = 0;
            for (int k = 0; k < n_orders; k++) {
                fb_k2_rice_search<G>(J, x, n, 0, k, nullptr, 0, smem, L, res_tmp);
                unsigned long long key = (unsigned long long)bps_v * (unsigned long long)k + res_tmp->code_bits;
                bool better = (kf < 0) || key < best_key;
                if (better) {
                    kf = k;
                    best_key = key;
                    FB_PHASE(tid, T)
                        for (int i = tid; i < (int)sizeof(FbRiceResult); i += T)
                            ((uint8_t *)res_fixed)[i] = ((const uint8_t *)res_tmp)[i];
                    FB_PHASE_END
                }
            }
            if (!(best_key < verbatim_bits)) kf = -1;
        }
        if (kf >= 0) fixed_bits = 8ull + (unsigned long long)bps_v * (unsigned long long)kf + res_fixed->res_bits;
    }
    const unsigned long long baseline_bits =
        kf >= 0 ? (fixed_bits < verbatim_bits ? fixed_bits : verbatim_bits) : verbatim_bits;

    // ---- LPC candidate (src/coding.rs:360-381)
    bool lpc_ok = false;
    unsigned long long lpc_bits = 0;
    const int16_t *lq = A.qlp; // the LPC candidate's coefficients, order after truncation, shift
    int lo = A.qlp_order, ls = A.qlp_shift, lp = J.cfg.quant_precision;
    if (J.cfg.use_lpc) {
        fb_k2_rice_search<G>(J, x, n, 1, A.qlp_order, A.qlp, A.qlp_shift, smem, L, res_lpc);
        lpc_bits = 8ull + (unsigned long long)bps_v * (unsigned long long)A.qlp_order + 4ull + 5ull +
                   (unsigned long long)J.cfg.quant_precision * (unsigned long long)A.qlp_order + res_lpc->res_bits;
        if (J.lpc_ext) {
            // EXTENSIONS (config.ext_lpc_*_search): the alternative sets of K1; fewest bits win, the earlier set on ties
            const FbLpcExt *ext = J.lpc_ext + (size_t)gv * FB_EXT_LPC_MAX;
            for (int k = 0; k < FB_EXT_LPC_MAX && ext[k].order > 0; k++) {
                fb_k2_rice_search<G>(J, x, n, 1, ext[k].order, ext[k].qlp, ext[k].shift, smem, L, res_tmp);
                const unsigned long long bits = 8ull + (unsigned long long)bps_v * (unsigned long long)ext[k].order + 4ull + 5ull +
                                                (unsigned long long)ext[k].precision * (unsigned long long)ext[k].order +
                                                res_tmp->res_bits;
                if (bits < lpc_bits) {
                    lpc_bits = bits;
                    lq = ext[k].qlp; lo = ext[k].order; ls = ext[k].shift; lp = ext[k].precision;
                    FB_PHASE(tid, T)
                        for (int i = tid; i < (int)sizeof(FbRiceResult); i += T)
                            ((uint8_t *)res_lpc)[i] = ((const uint8_t *)res_tmp)[i];
                    FB_PHASE_END
                }
            }
        }
        lpc_ok = lpc_bits < baseline_bits;
    }

    // ---- decision (src/coding.rs:403-416)
    int pick = -1; // 0 fixed, 1 lpc
    if (lpc_ok) pick = 1;
    else if (kf >= 0) pick = 0;
    if (pick == 1 && !(lpc_bits < verbatim_bits)) pick = -1;
    if (pick == 0 && !(fixed_bits < verbatim_bits)) pick = -1;
    if (pick < 0) return; // verbatim (already written)

    const FbRiceResult *R = pick == 1 ? res_lpc : res_fixed;
    FB_PHASE(tid, T)
        if (tid == 0) {
            out->type = pick == 1 ? FB200_SF_LPC : FB200_SF_FIXED;
            out->order = pick == 1 ? lo : kf;
            out->precision = pick == 1 ? lp : 0;
            out->shift = pick == 1 ? ls : 0;
            out->partition_order = R->part_order;
            out->rice2 = R->rice2;
            out->bits = pick == 1 ? lpc_bits : fixed_bits;
        }
        if (pick == 1)
            for (int i = tid; i < 32; i += T) out->qlp[i] = lq[i];
        for (int i = tid; i < (1 << R->part_order); i += T) out->rice_params[i] = R->params[i];
    FB_PHASE_END
}

// =================================================================================================
// K3: frame assembly (CTA per frame).
// =================================================================================================

// MSB-first bit writer of a thread-private run of the frame (src/bitsink.rs semantics).  Words are
// big-endian 32-bit groups of the byte stream; the buffer is zero-initialised, so runs of zero bits
// are skipped.  Only the first and last word of a run can be shared with a neighbour (atomic OR);
// interior words are owned by the run (plain store).
struct FbBitRun {
    uint32_t *words;
    uint32_t w_first, w_last, cur_w, acc, pos;
};

FB_DEV void fb_run_init(FbBitRun &r, uint32_t *words, uint32_t pos_start, uint32_t pos_end) {
    r.words = words;
    r.w_first = pos_start >> 5;
    r.w_last = pos_end > pos_start ? ((pos_end - 1) >> 5) : r.w_first;
    r.cur_w = r.w_first;
    r.acc = 0;
    r.pos = pos_start;
}

FB_DEV void fb_run_flush(FbBitRun &r) {
    if (r.acc) {
        if (r.cur_w == r.w_first || r.cur_w == r.w_last) fb_atomic_or(&r.words[r.cur_w], r.acc);
        else r.words[r.cur_w] = r.acc;
        r.acc = 0;
    }
}

FB_DEV void fb_run_skip(FbBitRun &r, uint32_t q) {
    r.pos += q;
    uint32_t nw = r.pos >> 5;
    if (nw != r.cur_w) { fb_run_flush(r); r.cur_w = nw; }
}

// append the k (1..32) low bits of v (v < 2^k)
FB_DEV void fb_run_put(FbBitRun &r, uint32_t v, uint32_t k) {
    uint32_t space = 32u - (r.pos & 31u);
    if (k <= space) {
        r.acc |= (k == 32u) ? v : (v << (space - k));
        r.pos += k;
        if (k == space) { fb_run_flush(r); r.cur_w++; }
    } else {
        uint32_t rem = k - space; // 1..31
        r.acc |= v >> rem;
        fb_run_flush(r);
        r.cur_w++;
        r.acc = v << (32u - rem);
        r.pos += k;
    }
}

struct FbK3Sub {
    int32_t  variant;       // index into the frame's variants
    int32_t  type, order, bps, precision, shift, part_order, rice2;
    uint32_t start_bit;     // first bit of the subframe
    uint32_t res_bit;       // first bit of the residual section (method field)
    uint32_t code_bit;      // first bit after the 6-bit residual header
};

struct FbK3Shared {
    FbK3Sub sub[FB200_MAX_CHANNELS];
    uint8_t header[16];
    int32_t header_len;
    int32_t ch_tag;
    uint32_t data_bytes;    // bytes covered by the CRC-16
    uint32_t crc_xpow[9];
    uint32_t crc_tab[256];
    uint32_t crc_part[FB_K3_THREADS];
    uint32_t scan_part[FB_K3_THREADS];
};

FB_HD uint32_t fb_k3_smem_bytes(uint32_t frame_bytes_max, int block_size, int pack_in_smem) {
    uint32_t runs = (uint32_t)((block_size + FB_RUN - 1) / FB_RUN);
    uint32_t o = (uint32_t)((sizeof(FbK3Shared) + 15) & ~(size_t)15);
    o += ((runs + 1 + 3) & ~3u) * 4u;                                   // run offsets
    o += (uint32_t)((block_size + (block_size >> 5) + 4 + 3) & ~3) * 4u; // zigzag residuals (fb_uidx layout)
    if (pack_in_smem) o += ((frame_bytes_max + 3u) & ~3u) + 16u;        // frame words
    return o;
}

template <int G>
FB_DEV void fb_k3_body(const FbJob &J, const int32_t *xv, const fb200_subframe_info *choice, uint8_t *slots,
                       uint32_t *frame_bytes, fb200_frame_info *infos, uint32_t f, uint8_t *smem) {
    const int T = FB_K3_THREADS;
    FbK3Shared *S = (FbK3Shared *)smem;
    uint32_t off = (uint32_t)((sizeof(FbK3Shared) + 15) & ~(size_t)15);
    const int n = fb_frame_len(J, f);
    const uint32_t runs_cap = (uint32_t)((J.block_size + FB_RUN - 1) / FB_RUN);
    uint32_t *run_off = (uint32_t *)(smem + off);
    off += ((runs_cap + 1 + 3) & ~3u) * 4u;
    uint32_t *ubuf = (uint32_t *)(smem + off);
    off += (uint32_t)((J.block_size + (J.block_size >> 5) + 4 + 3) & ~3) * 4u;
    uint8_t *slot = slots + (size_t)f * (size_t)J.slot_bytes;
    uint32_t *words = J.pack_in_smem ? (uint32_t *)(smem + off) : (uint32_t *)slot;
    const uint32_t max_bytes = fb_max_frame_bytes(J.channels, J.bps, J.block_size);
    const uint32_t max_words = (max_bytes + 3u) / 4u;
    const fb200_subframe_info *var = choice + (size_t)f * (size_t)J.nvar;
    const int nruns = (n + FB_RUN - 1) / FB_RUN;

    // ---- phase 0: zero the word buffer, CRC table; thread 0: stereo decision, header, offsets
    FB_PHASE(tid, T)
        for (uint32_t w = (uint32_t)tid; w < max_words; w += T) words[w] = 0;
        for (int i = tid; i < 256; i += T) S->crc_tab[i] = fb_crc16_table_entry((uint32_t)i);
        if (tid == 0) {
            int ch_tag = J.channels - 1;
            int sel[FB200_MAX_CHANNELS];
            for (int c = 0; c < J.channels; c++) sel[c] = c;
            if (J.channels == 2) {
                // try_stereo_coding (src/coding.rs:469-527): strict <, order I, L/S, R/S, M/S
                unsigned long long bl = var[0].bits, br = var[1].bits, bm = var[2].bits, bs = var[3].bits;
                unsigned long long min_bits = bl + br;
                if (J.cfg.use_leftside && bl + bs < min_bits) { min_bits = bl + bs; ch_tag = 8; }
                if (J.cfg.use_rightside && br + bs < min_bits) { min_bits = br + bs; ch_tag = 9; }
                if (J.cfg.use_midside && bm + bs < min_bits) { min_bits = bm + bs; ch_tag = 10; }
                // select_channels (src/component/datatype.rs:1171-1184)
                if (ch_tag == 8) { sel[0] = 0; sel[1] = 3; }
                else if (ch_tag == 9) { sel[0] = 3; sel[1] = 1; }
                else if (ch_tag == 10) { sel[0] = 2; sel[1] = 3; }
            }
            S->ch_tag = ch_tag;
            S->header_len = fb_frame_header(n, ch_tag, J.bps, J.sample_rate, J.first_frame_number + f, S->header);
            uint32_t bit = (uint32_t)S->header_len * 8u;
            for (int c = 0; c < J.channels; c++) {
                const fb200_subframe_info &V = var[sel[c]];
                FbK3Sub &D = S->sub[c];
                D.variant = sel[c];
                D.type = V.type; D.order = V.order; D.bps = V.bits_per_sample;
                D.precision = V.precision; D.shift = V.shift; D.part_order = V.partition_order; D.rice2 = V.rice2;
                D.start_bit = bit;
                uint32_t hb = 8;
                if (V.type == FB200_SF_FIXED) hb += (uint32_t)(V.order * V.bits_per_sample);
                if (V.type == FB200_SF_LPC)
                    hb += (uint32_t)(V.order * V.bits_per_sample) + 4u + 5u + (uint32_t)(V.precision * V.order);
                D.res_bit = bit + hb;
                D.code_bit = D.res_bit + 6u;
                bit += (uint32_t)V.bits;
            }
            S->data_bytes = (bit + 7u) >> 3;
        }
    FB_PHASE_END

    // ---- phase 1: header bytes and the fixed-position leading fields of every subframe
    FB_PHASE(tid, T)
        if (tid == 0) {
            FbBitRun r;
            fb_run_init(r, words, 0, 1); // every word through the atomic path
            r.w_last = 0xFFFFFFFFu;
            for (int i = 0; i < S->header_len; i++) {
                r.w_first = r.cur_w; // force atomic OR on each flush
                fb_run_put(r, S->header[i], 8);
            }
            r.w_first = r.cur_w;
            fb_run_flush(r);
        }
        if (tid >= 1 && tid <= J.channels) {
            const FbK3Sub &D = S->sub[tid - 1];
            const int32_t *x = xv + ((size_t)f * (size_t)J.nvar + (size_t)D.variant) * (size_t)J.stride;
            const fb200_subframe_info &V = var[D.variant];
            FbBitRun r;
            fb_run_init(r, words, D.start_bit, D.start_bit + 1);
            r.w_last = 0xFFFFFFFFu;
            const uint32_t mask = D.bps >= 32 ? 0xFFFFFFFFu : ((1u << D.bps) - 1u);
#define FB_PUT_ATOMIC(val, nb) do { r.w_first = r.cur_w; fb_run_put(r, (uint32_t)(val), (uint32_t)(nb)); } while (0)
            if (D.type == FB200_SF_CONSTANT) {
                FB_PUT_ATOMIC(0x00, 8);
                FB_PUT_ATOMIC((uint32_t)x[0] & mask, D.bps);
            } else if (D.type == FB200_SF_VERBATIM) {
                FB_PUT_ATOMIC(0x02, 8);
            } else if (D.type == FB200_SF_FIXED) {
                FB_PUT_ATOMIC(0x10 | (D.order << 1), 8);
                for (int t = 0; t < D.order; t++) FB_PUT_ATOMIC((uint32_t)x[t] & mask, D.bps);
                FB_PUT_ATOMIC(((D.rice2 ? 1 : 0) << 4) | D.part_order, 6);
            } else {
                FB_PUT_ATOMIC(0x40 | ((D.order - 1) << 1), 8);
                for (int t = 0; t < D.order; t++) FB_PUT_ATOMIC((uint32_t)x[t] & mask, D.bps);
                FB_PUT_ATOMIC(D.precision - 1, 4);
                FB_PUT_ATOMIC((uint32_t)D.shift & 31u, 5);
                const uint32_t pmask = (1u << D.precision) - 1u;
                for (int j = 0; j < D.order; j++) FB_PUT_ATOMIC((uint32_t)(int32_t)V.qlp[j] & pmask, D.precision);
                FB_PUT_ATOMIC(((D.rice2 ? 1 : 0) << 4) | D.part_order, 6);
            }
#undef FB_PUT_ATOMIC
            r.w_first = r.cur_w;
            fb_run_flush(r);
        }
    FB_PHASE_END

    // ---- per subframe: samples
    for (int c = 0; c < J.channels; c++) {
        const FbK3Sub D = S->sub[c];
        const int32_t *x = xv + ((size_t)f * (size_t)J.nvar + (size_t)D.variant) * (size_t)J.stride;
        const fb200_subframe_info &V = var[D.variant];
        if (D.type == FB200_SF_CONSTANT) continue;
        if (D.type == FB200_SF_VERBATIM) {
            // Verbatim::write (src/component/bitrepr.rs:463-470): bps bits per sample, fixed positions
            FB_PHASE(tid, T)
                const uint32_t mask = D.bps >= 32 ? 0xFFFFFFFFu : ((1u << D.bps) - 1u);
                for (int run = tid; run < nruns; run += T) {
                    int t0 = run * FB_RUN, t1 = t0 + FB_RUN < n ? t0 + FB_RUN : n;
                    uint32_t p0 = D.start_bit + 8u + (uint32_t)t0 * (uint32_t)D.bps;
                    uint32_t p1 = D.start_bit + 8u + (uint32_t)t1 * (uint32_t)D.bps;
                    FbBitRun r;
                    fb_run_init(r, words, p0, p1);
                    for (int t = t0; t < t1; t++) fb_run_put(r, (uint32_t)x[t] & mask, (uint32_t)D.bps);
                    fb_run_flush(r);
                }
            FB_PHASE_END
            continue;
        }
        // Residual::write (src/component/bitrepr.rs:550-597): per partition a 4/5-bit parameter,
        // then per sample q zeros, a one, and the p low bits.
        const int kind = D.type == FB200_SF_LPC ? 1 : 0;
        const int warm = D.order;
        const int plen = n >> D.part_order;
        const uint32_t pbits = D.rice2 ? 5u : 4u;
        // pass A: bits of each run (codes + the parameter fields that start inside it)
        FB_PHASE(tid, T)
            for (int run = tid; run < nruns; run += T) {
                int t0 = run * FB_RUN, t1 = t0 + FB_RUN < n ? t0 + FB_RUN : n;
                uint32_t bits = 0;
                int pj = t0 / plen;
                int pstart = pj * plen;
                if (pstart < warm) pstart = warm; // partition 0 starts after the warm-up
                uint32_t rp = V.rice_params[pj];
                int pnext = (pj + 1) * plen;
                uint32_t uu16[FB_RUN];
                fb_run_residual<G>(x, n, t0, kind, D.order, V.qlp, D.shift, uu16);
#pragma unroll
                for (int i = 0; i < FB_RUN; i++) {
                    const int t = t0 + i;
                    if (t < t1) {
                        ubuf[fb_uidx(t)] = uu16[i];
                        if (t == pnext) { pj++; pstart = pnext; pnext += plen; rp = V.rice_params[pj]; }
                        if (t >= warm) {
                            if (t == pstart) bits += pbits;
                            bits += (uu16[i] >> rp) + 1u + rp;
                        }
                    }
                }
                run_off[run] = bits;
            }
        FB_PHASE_END
        // exclusive scan of run_off[0..nruns) (three steps: slice sums, scan of T partials, fix-up)
        {
            const int per = (nruns + T - 1) / T;
            FB_PHASE(tid, T)
                uint32_t s = 0;
                for (int i = tid * per; i < (tid + 1) * per && i < nruns; i++) s += run_off[i];
                S->scan_part[tid] = s;
            FB_PHASE_END
            FB_PHASE(tid, T)
                if (tid == 0) {
                    uint32_t s = 0;
                    for (int i = 0; i < T; i++) { uint32_t v2 = S->scan_part[i]; S->scan_part[i] = s; s += v2; }
                }
            FB_PHASE_END
            FB_PHASE(tid, T)
                uint32_t s = S->scan_part[tid];
                for (int i = tid * per; i < (tid + 1) * per && i < nruns; i++) { uint32_t v2 = run_off[i]; run_off[i] = s; s += v2; }
                if (tid == T - 1) run_off[nruns] = s;
            FB_PHASE_END
        }
        // pass B: pack
        FB_PHASE(tid, T)
            for (int run = tid; run < nruns; run += T) {
                int t0 = run * FB_RUN, t1 = t0 + FB_RUN < n ? t0 + FB_RUN : n;
                uint32_t p0 = D.code_bit + run_off[run];
                uint32_t p1 = D.code_bit + run_off[run + 1];
                if (p1 == p0) continue;
                FbBitRun r;
                fb_run_init(r, words, p0, p1);
                int pj = t0 / plen;
                int pstart = pj * plen;
                if (pstart < warm) pstart = warm;
                uint32_t rp = V.rice_params[pj];
                int pnext = (pj + 1) * plen;
                for (int t = t0; t < t1; t++) {
                    if (t == pnext) { pj++; pstart = pnext; pnext += plen; rp = V.rice_params[pj]; }
                    if (t < warm) continue;
                    if (t == pstart) fb_run_put(r, rp, pbits);
                    const uint32_t uu = ubuf[fb_uidx(t)];
                    fb_run_skip(r, uu >> rp);
                    fb_run_put(r, (uu & ((1u << rp) - 1u)) | (1u << rp), rp + 1u);
                }
                fb_run_flush(r);
            }
        FB_PHASE_END
    }

    // ---- CRC-16 over data_bytes (Frame::write, src/component/bitrepr.rs:289-320).
    // Chunks of L bytes are aligned to the END of the data (leading zero bytes do not change a
    // CRC with init 0); partial CRCs are combined pairwise with x^(8*L*2^k) mod P.
    const uint32_t B = S->data_bytes;
    const uint32_t Lc = (B + (uint32_t)T - 1u) / (uint32_t)T;
    FB_PHASE(tid, T)
        if (tid == 0) {
            // x^(8*Lc) mod P by square-and-multiply, then successive squares
            uint32_t result = 1, base = 2; // polynomial "x"
            uint32_t e = 8u * Lc;
            while (e) { if (e & 1u) result = fb_crc16_mulmod(result, base); base = fb_crc16_mulmod(base, base); e >>= 1; }
            S->crc_xpow[0] = result;
            for (int k = 1; k < 9; k++) S->crc_xpow[k] = fb_crc16_mulmod(S->crc_xpow[k - 1], S->crc_xpow[k - 1]);
        }
        // chunk tid covers bytes [B - (T - tid)*Lc, B - (T - tid - 1)*Lc) intersected with [0, B)
        long long lo = (long long)B - (long long)(T - tid) * (long long)Lc;
        long long hi = lo + (long long)Lc;
        if (lo < 0) lo = 0;
        uint32_t crc = 0;
        for (long long i = lo; i < hi; i++) {
            uint32_t byte = (words[i >> 2] >> (24u - 8u * (uint32_t)(i & 3))) & 0xFFu;
            crc = ((crc << 8) & 0xFFFFu) ^ S->crc_tab[((crc >> 8) ^ byte) & 0xFFu];
        }
        S->crc_part[tid] = crc;
    FB_PHASE_END
    for (int k = 0; (1 << k) < T; k++) {
        FB_PHASE(tid, T)
            const int span = 1 << (k + 1);
            if ((tid % span) == 0) {
                uint32_t left = S->crc_part[tid], right = S->crc_part[tid + (1 << k)];
                S->crc_part[tid] = fb_crc16_mulmod(left, S->crc_xpow[k]) ^ right;
            }
        FB_PHASE_END
    }
    FB_PHASE(tid, T)
        if (tid == 0) {
            uint32_t crc = S->crc_part[0];
            // the two CRC bytes follow the (byte-aligned) data, big-endian
            for (int i = 0; i < 2; i++) {
                uint32_t pos = B + (uint32_t)i;
                uint32_t byte = (crc >> (8 * (1 - i))) & 0xFFu;
                fb_atomic_or(&words[pos >> 2], byte << (24u - 8u * (pos & 3u)));
            }
            frame_bytes[f] = B + 2u;
            if (infos) {
                fb200_frame_info &I = infos[f];
                I.channel_assignment = S->ch_tag;
                I.block_size = n;
                I.frame_number = J.first_frame_number + f;
                I.frame_bytes = B + 2u;
            }
        }
        if (infos) {
            // decision records of the emitted subframes
            fb200_frame_info &I = infos[f];
            for (int c = 0; c < J.channels; c++) {
                const uint8_t *src = (const uint8_t *)&var[S->sub[c].variant];
                uint8_t *dst = (uint8_t *)&I.sub[c];
                for (int i = tid; i < (int)sizeof(fb200_subframe_info); i += T) dst[i] = src[i];
            }
        }
    FB_PHASE_END
    // ---- store: big-endian words -> bytes of the slot
    FB_PHASE(tid, T)
        const uint32_t total = B + 2u;
        const uint32_t nwords = (total + 3u) / 4u;
        uint32_t *dstw = (uint32_t *)slot;
        for (uint32_t w = (uint32_t)tid; w < nwords; w += T) {
            uint32_t v = words[w];
            dstw[w] = ((v & 0xFFu) << 24) | ((v & 0xFF00u) << 8) | ((v >> 8) & 0xFF00u) | (v >> 24);
        }
    FB_PHASE_END
}

// =================================================================================================
// K4: exclusive scan of frame sizes (single CTA) and gather into a contiguous stream.
// =================================================================================================
#define FB_K4_THREADS 1024

FB_DEV void fb_k4_scan_body(const uint32_t *frame_bytes, unsigned long long *offsets, uint32_t n_frames,
                            unsigned long long *partials /* smem, T entries */, unsigned long long base = 0) {
    const int T = FB_K4_THREADS;
    const uint32_t per = (n_frames + (uint32_t)T - 1u) / (uint32_t)T;
    FB_PHASE(tid, T)
        unsigned long long s = 0;
        for (uint32_t i = (uint32_t)tid * per; i < ((uint32_t)tid + 1u) * per && i < n_frames; i++) s += frame_bytes[i];
        partials[tid] = s;
    FB_PHASE_END
#if FB_GPU
    {
        // exclusive scan of the T partials: shuffles inside each warp, then over the 32 warp totals
        const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
        __shared__ unsigned long long warp_tot[FB_K4_THREADS / 32];
        const unsigned long long v = partials[tid];
        unsigned long long inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const unsigned long long w = warp_tot[lane];
            unsigned long long winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, winc, d);
                if (lane >= d) winc += o;
            }
            warp_tot[lane] = winc - w; // exclusive
        }
        __syncthreads();
        partials[tid] = warp_tot[warp] + inc - v;
        __syncthreads();
    }
#else
    FB_PHASE(tid, T)
        if (tid == 0) {
            unsigned long long s = 0;
            for (int i = 0; i < T; i++) { unsigned long long v = partials[i]; partials[i] = s; s += v; }
        }
    FB_PHASE_END
#endif
    FB_PHASE(tid, T)
        unsigned long long s = base + partials[tid];
        for (uint32_t i = (uint32_t)tid * per; i < ((uint32_t)tid + 1u) * per && i < n_frames; i++) {
            offsets[i] = s;
            s += frame_bytes[i];
        }
        if (tid == T - 1) offsets[n_frames] = s;
    FB_PHASE_END
}

// copies frame f from its (16-byte aligned) slot to out + offsets[f]: bytes up to the first 4-byte aligned
// destination address, then whole destination words assembled from two source words, then the tail bytes
FB_DEV void fb_k4_gather_thread(const uint8_t *slots, uint32_t slot_bytes, const uint32_t *frame_bytes,
                                const unsigned long long *offsets, uint8_t *out, unsigned long long out_cap,
                                uint32_t f, int tid, int T) {
    const uint8_t *src = slots + (size_t)f * (size_t)slot_bytes;
    const unsigned long long o = offsets[f];
    const uint32_t len = frame_bytes[f];
    if (o + len > out_cap) return; // capacity error is reported by the host from offsets[n_frames]
    uint8_t *dst = out + o;
    uint32_t head = (uint32_t)((4u - (uint32_t)((uintptr_t)dst & 3u)) & 3u);
    if (head > len) head = len;
    const uint32_t nw = (len - head) >> 2;
    if ((uint32_t)tid < head) dst[tid] = src[tid];
    const uint32_t *srcw = (const uint32_t *)src;
    uint32_t *dstw = (uint32_t *)(dst + head);
    const uint32_t sh = (head & 3u) * 8u;
    for (uint32_t k = (uint32_t)tid; k < nw; k += (uint32_t)T) {
        const uint32_t wi = (head >> 2) + k;
        uint32_t v = srcw[wi];
        if (sh) v = (v >> sh) | (srcw[wi + 1] << (32u - sh)); // wi + 1 still holds frame bytes when sh != 0
        dstw[k] = v;
    }
    const uint32_t done = head + 4u * nw;
    if ((uint32_t)tid < len - done) dst[done + (uint32_t)tid] = src[done + (uint32_t)tid];
}
