// fb_fused.cuh -- the fused path: per-frame plan kernel KA (Rice search of every channel variant, decisions, frame
// plan) and per-frame pack kernel KP (bit packing, CRC-16, store at the final stream offset).
//
// KA: one CTA per frame, one warp per channel variant (L, R, M, S for stereo).  The frame's independent
// channels are staged once in shared memory (M and S are formed on the fly: src/coding.rs:476-484);
// everything else runs out of shared memory and registers:
//
//   per variant (one warp), for the fixed winner and for the LPC candidate
//     pass 1  residual -> zigzag in register-window runs of 16 samples (src/lpc.rs:306-390,
//             src/coding.rs:182-197, src/rice.rs:169-171).  The samples of a "unit" (<= 112 samples, a
//             leaf of the finest Rice partitioning or a piece of one) are never stored: they are added
//             into a bit-sliced (carry-save) counter of 7 words, w_j holding bit j of the per-bit-plane
//             population counts.  Because the counter is exact per bit plane,
//                 sum_t (u_t >> p)  ==  sum_j (w_j >> p) << j          for every p,
//             so any entry of the reference's PrcBitTable (src/rice.rs:65-105) costs 7 shifts per unit
//             instead of one shift per sample.
//     pass 2  per finest partition: the exact minimiser, found by walking the convex cost function (done
//             in registers right after the unit's last run when a unit is a whole partition)
//     pass 3  bottom-up partition tree (src/rice.rs:246-298) over the parameter window
//             [min leaf minimiser, max leaf minimiser] -- the minimiser of every merged node lies in it
//             (sum of convex functions) -- in passes of FB_KF_COLS parameters
//     pass 4  partition order, parameters, exact Residual::count_bits (src/component/bitrepr.rs:532-544)
//             and the bit length of every unit (again from the counters)
//   subframe decision (src/coding.rs:384-418); then per frame: stereo decision (src/coding.rs:454-527),
//   header + CRC-8, bit offsets of all units by a warp scan; the plan and the frame size go to global memory.
// KP (after the scan of the frame sizes): stages the channels again (16-bit stereo: the packed PCM itself, as
//   (left, right) pairs), every thread packs one unit (residuals recomputed, every code OR-ed on its own into the
//   zeroed word buffer), CRC-16, store at out + offsets[f].
//
// Frames whose units do not start on multiples of 4 samples (finest partitions of odd tail frames) run through a
// second template instance of both bodies (ODD: sample-by-sample window loads) inside the same launches.
// Whatever KA cannot reproduce exactly -- a residual >= 2^26 (the reference's 16-sample chunked saturating
// accumulation matters, src/rice.rs:75-98) or a saturated table minimum -- is not guessed: the frame is appended to
// a fallback list and redone by the generic K2/K3 kernels (fb_kernels.cuh), which replay the reference literally.
// Results are byte-identical either way.
#pragma once

#include "fb_kernels.cuh"

#ifndef FB_KF_RUN
#define FB_KF_RUN 8        // samples per residual run (register window of G + FB_KF_RUN samples)
#endif
#define FB_KF_COLS 4       // Rice parameters evaluated per tree pass
#define FB_KF_ROW 5        // row stride of the tree tables (one pad word: conflict-free column reads)
#define FB_KF_NWORDS 7     // bit-sliced counter words per unit (counts <= 127)
#define FB_KF_UNIT_MAX 112 // samples per unit (14 runs of 8; the counters hold counts <= 127)
#define FB_KF_SMEM_LIMIT (225u * 1024u) // dynamic shared memory one CTA may ask for (227 KiB on sm_100a, minus slack)

// 16-byte global -> shared copy that does not pass through registers (LDGSTS), so a thread keeps all the copies of
// its staging loop in flight at once
#if FB_GPU
FB_DEV void fb_copy16_async(int32_t *dst_smem, const int32_t *src_global) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src_global) : "memory");
}
FB_DEV void fb_copy8_async(void *dst_smem, const void *src_global) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src_global) : "memory");
}
FB_DEV void fb_copy4_async(void *dst_smem, const void *src_global) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src_global) : "memory");
}
FB_DEV void fb_copy_async_wait() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
#else
FB_DEV void fb_copy16_async(int32_t *dst, const int32_t *src) { memcpy(dst, src, 16); }
FB_DEV void fb_copy8_async(void *dst, const void *src) { memcpy(dst, src, 8); }
FB_DEV void fb_copy4_async(void *dst, const void *src) { memcpy(dst, src, 4); }
FB_DEV void fb_copy_async_wait() {}
#endif

#if FB_GPU
FB_DEV void fb_atomic_or_u32(uint32_t *p, uint32_t v) { atomicOr(p, v); }
FB_DEV int fb_clz32(uint32_t v) { return __clz((int)v); }
FB_DEV int fb_clz64(unsigned long long v) { return __clzll((long long)v); }
#else
FB_DEV void fb_atomic_or_u32(uint32_t *p, uint32_t v) { *p |= v; }
FB_DEV int fb_clz32(uint32_t v) { return v ? __builtin_clz(v) : 32; }
FB_DEV int fb_clz64(unsigned long long v) { return v ? __builtin_clzll(v) : 64; }
#endif

// ---- geometry of one frame length -----------------------------------------------------------------
struct FbKfGeom {
    int n, o0, leaves, leaf_len;
    int m, lgm;   // units per leaf (power of two)
    int U, lgU;   // units per variant = leaves * m, a power of two >= 32
    int unit_len; // samples per unit (multiple of 4, <= FB_KF_UNIT_MAX); the last unit of a leaf may be shorter
};

FB_HD FbKfGeom fb_kf_geom(int n) {
    FbKfGeom g;
    g.n = n;
    g.o0 = fb_finest_partition_order(n);
    g.leaves = 1 << g.o0;
    g.leaf_len = n >> g.o0;
    int m = 1, lgm = 0;
    while (g.leaves * m < 32) { m <<= 1; lgm++; }
    while (((((g.leaf_len + m - 1) >> lgm) + 3) & ~3) > FB_KF_UNIT_MAX) { m <<= 1; lgm++; } // (m = 1 << lgm)
    g.m = m;
    g.lgm = lgm;
    g.unit_len = (((g.leaf_len + m - 1) >> lgm) + 3) & ~3;
    g.U = g.leaves * m;
    g.lgU = g.o0 + lgm;
    return g;
}

// which frames of a job need the ODD kernel instances (units that do not start on multiples of 4 samples):
// 0 = none, 1 = only the last (shorter) frame, 2 = all
FB_HD int fb_kf_odd_mode(const FbJob &J) {
    if (fb_kf_geom(J.block_size).leaf_len & 3) return 2;
    if (J.n_frames > 0 && J.tail_n != J.block_size && (fb_kf_geom(J.tail_n).leaf_len & 3)) return J.n_frames > 1 ? 1 : 2;
    return 0;
}

// sample range [t0, t1) of unit u
FB_HD void fb_kf_unit_range(const FbKfGeom &g, int u, int *t0, int *t1) {
    const int leaf = u >> g.lgm, k = u & (g.m - 1);
    const int ls = leaf * g.leaf_len, le = ls + g.leaf_len;
    int a = ls + k * g.unit_len, b = a + g.unit_len;
    if (a > le) a = le;
    if (b > le) b = le;
    *t0 = a;
    *t1 = b;
}

// index of sample t in a staged plane: 4 pad words per 64 samples keep 16-byte loads of lanes that are
// 64 samples apart on different banks
FB_HD int fb_xidx(int t) { return t + ((t >> 6) << 2); }
// A kernel may stage streams of at most 16 bits per sample as int16 (half the shared memory, more CTAs per SM);
// fb_xidx is then a halfword index (8 pad bytes per 64 samples keep the 8-byte loads conflict-free).  The flag
// travels in bit 2 of the variant mode `vm`.
#define FB_VM_X16 4
#define FB_VM_PAIRS 8 // one plane of 16-bit stereo pairs (left in the low half); bits 0-1: 0 left, 1 right, 2 mid, 3 side

// ---- CRC-16 tables (poly 0x8005, init 0, MSB first; src/component/bitrepr.rs:40,270-271) ------------
// Built on the host once per context and read by the kernel:
//   [0, 1024)            slicing-by-4 tables: T[k][b] = CRC of byte b followed by k zero bytes
//   [1280, 1536)         CRC-8 (poly 0x07, init 0) byte table for the frame header
//   [1536, 1536 + Lm+4)  xb[i] = x^(8 * i) mod P, i <= Lm   (Lm = largest CRC chunk in bytes, see fb_kf_crc_chunk)
//   [XP, XP + Lm/4 * T)  xp[m - 1][j] = x^(8 * 4m * j) mod P, j < T: the shifts for chunks of 4m bytes.  A frame of B
//                        bytes is cut into chunks of Lc = 4 * ceil(B / 4T) bytes, so that (nearly) all T threads of
//                        the CTA get one, however short the frame is
#define FB_KTAB_C8 1280
#define FB_KTAB_XB 1536
FB_HD uint32_t fb_kf_crc_chunk(int channels, int bps, int block_size, int threads) {
    const uint32_t mb = fb_max_frame_bytes(channels, bps, block_size);
    return ((mb + (uint32_t)threads - 1u) / (uint32_t)threads + 3u) & ~3u;
}
FB_HD uint32_t fb_kf_ktab_xp(uint32_t Lm) { return FB_KTAB_XB + Lm + 4u; }
FB_HD uint32_t fb_kf_ktab_words(uint32_t Lm, uint32_t T) { return fb_kf_ktab_xp(Lm) + (Lm / 4u) * T; }
inline void fb_kf_build_ktab(uint32_t Lc, uint32_t T, uint32_t *t) {
    for (uint32_t b = 0; b < 256; b++) {
        uint32_t c = fb_crc16_table_entry(b);
        t[b] = c;
        for (int k = 1; k < 4; k++) {
            c = ((c << 8) & 0xFFFFu) ^ fb_crc16_table_entry((c >> 8) & 0xFFu);
            t[k * 256 + b] = c;
        }
    }
    for (uint32_t b = 0; b < 256; b++) {
        const uint8_t one = (uint8_t)b;
        t[FB_KTAB_C8 + b] = fb_crc8(&one, 1);
    }
    // x^(8*i): start from 1 and multiply by x^8 (= 0x100 reduced: as a 16-bit polynomial x^8 is 0x0100)
    uint32_t v = 1;
    for (uint32_t i = 0; i <= Lc + 3; i++) { t[FB_KTAB_XB + i] = v; v = fb_crc16_mulmod(v, 0x0100u); }
    for (uint32_t m = 1; m <= Lc / 4u; m++) {
        const uint32_t step = t[FB_KTAB_XB + 4u * m];
        uint32_t *row = t + fb_kf_ktab_xp(Lc) + (m - 1u) * T;
        v = 1;
        for (uint32_t j = 0; j < T; j++) { row[j] = v; v = fb_crc16_mulmod(v, step); }
    }
}

// ---- shared-memory layout (bytes), identical on host and device ---------------------------------
struct FbKfLayout {
    uint32_t x_stride;   // 32-bit words per staged plane
    uint32_t x16;        // planes hold int16 samples
    uint32_t off_x;      // channels planes
    uint32_t off_keep;   // per warp: unit_bits[2][U+1], results of both candidates
    uint32_t keep_bytes;
    uint32_t k_unit_bits, k_res, res_stride;
    uint32_t off_choice; // pack kernel: channels x fb200_subframe_info
    uint32_t off_frame;  // FbKfFrame
    uint32_t off_ana;    // plan kernel: the frame's FbAnalysis records (K1's results)
    uint32_t off_scratch; // per warp scratch, aliased by the frame words during packing
    uint32_t scratch_bytes;
    uint32_t s_words, s_tbl_a, s_tbl_b, s_best_val, s_best_p, s_lvl_bits, s_misc;
    uint32_t words_bytes; // bytes of the frame word buffer
    uint32_t U_max, leaves_max;
    uint32_t crc_chunk;   // Lc
    uint32_t off_crc_tab; // pack kernel: the four CRC-16 slicing tables (4 KiB)
    uint32_t group_ch;    // pack kernel: channels staged at a time (all of them unless the frame is too large)
    uint32_t total;
};

struct FbKfRes {
    int32_t  part_order;
    int32_t  rice2;
    unsigned long long res_bits; // Residual::count_bits()
    uint8_t  params[FB200_MAX_RICE_PARTS];
};

struct FbKfMisc {
    uint32_t ormask, pmin, pmax, fail;
    int32_t  best_level, any_gt14;
};

struct FbKfSub {
    int32_t  variant, cand;  // cand: 0 fixed, 1 lpc (index of the unit_bits / result set)
    int32_t  type, order, bps, precision, shift, part_order, rice2;
    uint32_t start_bit, res_bit, code_bit;
};

struct FbKfVHead { // what the frame-level decisions need of a variant's record (the full record is in global memory)
    unsigned long long bits;
    int16_t type, order, bps, precision, shift, part_order, rice2, pad;
};

struct FbKfFrame {
    FbKfSub sub[FB200_MAX_CHANNELS];
    uint8_t header[16];
    int32_t header_len, ch_tag;
    uint32_t data_bytes;
    uint32_t frame_fail;
    int32_t  cand[FB200_MAX_CHANNELS]; // per variant: result set of the chosen coding (0 fixed, 1 lpc)
    uint32_t crc_acc, crc_last;
    FbKfVHead vh[FB200_MAX_CHANNELS]; // plan kernel only
};

FB_HD uint32_t fb_align16(uint32_t v) { return (v + 15u) & ~15u; }

FB_HD FbKfLayout fb_kf_layout(int channels, int nvar, int bps, int block_size, int tail_n, bool x16 = false, bool pairs = false) {
    FbKfLayout L;
    const FbKfGeom ga = fb_kf_geom(block_size), gb = fb_kf_geom(tail_n);
    const uint32_t U = (uint32_t)(ga.U > gb.U ? ga.U : gb.U);
    const uint32_t leaves = (uint32_t)(ga.leaves > gb.leaves ? ga.leaves : gb.leaves);
    L.U_max = U;
    L.leaves_max = leaves;
    L.crc_chunk = fb_kf_crc_chunk(channels, bps, block_size, 32 * nvar);
    L.off_crc_tab = 0;
    L.group_ch = (uint32_t)channels;
    L.x16 = (x16 && bps <= 16) ? 1u : 0u;
    L.x_stride = (uint32_t)((fb_xidx(block_size + 32) + 8 + 3) & ~3);
    if (L.x16) L.x_stride = ((L.x_stride + 1u) / 2u + 3u) & ~3u;
    if (pairs) L.x16 = 2; // one plane of (left, right) pairs: the frame's packed PCM itself
    uint32_t o = 0;
    L.off_x = o;        o += fb_align16((pairs ? 1u : (uint32_t)channels) * L.x_stride * 4u);
    // kept per warp
    uint32_t k = 0;
    L.k_unit_bits = k;  k += fb_align16(2u * (U + 1u) * 4u);
    L.res_stride = fb_align16((uint32_t)offsetof(FbKfRes, params) + leaves); // params[] holds at most `leaves` entries
    L.k_res = k;        k += 2u * L.res_stride;
    L.keep_bytes = k;
    L.off_keep = o;     o += (uint32_t)nvar * k;
    L.off_choice = 0;   // (the plan kernel keeps the variants' records in global memory)
    L.off_frame = o;    o += fb_align16((uint32_t)sizeof(FbKfFrame));
    L.off_ana = o;      o += fb_align16((uint32_t)nvar * (uint32_t)sizeof(FbAnalysis));
    // scratch per warp
    uint32_t s = 0;
    L.s_words = s;      s += fb_align16(FB_KF_NWORDS * U * 4u);
    uint32_t ta = leaves * FB_KF_ROW * 4u;
    {
        uint32_t levels = 1; // also the level-sum exchange (levels x 32 lanes) and the offset scan's 32 words
        while ((1u << (levels - 1u)) < leaves) levels++;
        if (ta < levels * 32u * 4u) ta = levels * 32u * 4u;
    }
    L.s_tbl_a = s;      s += fb_align16(ta);
    L.s_tbl_b = s;      // (no second table: the tree levels are merged in place)
    L.s_best_val = s;   s += fb_align16(2u * leaves * 4u);
    L.s_best_p = s;     s += fb_align16(2u * leaves);
    L.s_lvl_bits = s;   s += 16u * 8u;
    L.s_misc = s;       s += fb_align16((uint32_t)sizeof(FbKfMisc));
    L.scratch_bytes = s;
    L.words_bytes = fb_align16(((fb_max_frame_bytes(channels, bps, block_size) + 3u) & ~3u) + 16u);
    L.off_scratch = o;  o += (uint32_t)nvar * s; // (the pack kernel's layout puts the frame words here instead)
    L.total = o;
    return L;
}

// Shared memory of the pack kernel KP: planes (int16 when the stream has at most 16 bits per sample), the frame's
// plan, the chosen subframe records, the unit offsets and the frame's word buffer.  Returned in the same struct.
// pairs: 16-bit stereo whose packed PCM is at hand (2-byte container, 16-byte aligned): the frame's PCM itself is
// staged as one plane of (left, right) pairs (x16 = 2) -- plain 16-byte asynchronous copies, no conversion.
FB_HD bool fb_kp_pairs_format(int channels, int bps, int container_bytes) {
    return channels == 2 && bps == 16 && container_bytes == 2;
}
FB_HD FbKfLayout fb_kp_layout(int channels, int nvar, int bps, int block_size, int tail_n, bool pairs = false) {
    FbKfLayout L = fb_kf_layout(channels, nvar, bps, block_size, tail_n, !pairs);
    if (pairs) L.x16 = 2;
    // everything but the planes, then as many planes as fit next to it (channels other than a stereo pair are
    // independent, so they can be staged and packed group by group)
    const uint32_t fixed = fb_align16((uint32_t)channels * (L.U_max + 1u) * 4u) +
                           fb_align16((uint32_t)channels * (uint32_t)sizeof(fb200_subframe_info)) +
                           fb_align16((uint32_t)sizeof(FbKfFrame)) + L.words_bytes;
    const uint32_t plane = L.x_stride * 4u;
    uint32_t gc = (uint32_t)channels;
    while (gc > 1u && fixed + fb_align16(gc * plane) > FB_KF_SMEM_LIMIT) gc--;
    if (channels == 2) gc = 2; // M and S need both planes
    L.group_ch = gc;
    uint32_t o = 0;
    // the planes; the four CRC-16 slicing tables (4 KiB) take their place once the frame is packed
    const uint32_t planes = fb_align16((pairs ? 1u : gc) * L.x_stride * 4u);
    L.off_x = o;        L.off_crc_tab = o;  o += planes > 4096u ? planes : 4096u;
    L.off_keep = o;     o += fb_align16((uint32_t)channels * (L.U_max + 1u) * 4u);   // unit offsets
    L.off_choice = o;   o += fb_align16((uint32_t)channels * (uint32_t)sizeof(fb200_subframe_info));
    L.off_frame = o;    o += fb_align16((uint32_t)sizeof(FbKfFrame));
    L.off_scratch = o;  o += L.words_bytes;                                            // frame words
    L.total = o;
    return L;
}

// ---- bit-sliced counters ----------------------------------------------------------------------------
// full adder on 32 independent bit planes: (h, l) = a + b + c
#define FB_CSA(h, l, a, b, c) do { const uint32_t a__ = (a), b__ = (b), c__ = (c); const uint32_t x__ = a__ ^ b__; \
                                   (h) = (a__ & b__) | (x__ & c__); (l) = x__ ^ c__; } while (0)

// adds FB_KF_RUN (8 or 16) values into the counter words cw[0..6] (weights 1, 2, 4, 8, 16, 32, 64);
// counts stay <= 127 by construction (units of at most 112 samples)
FB_DEV void fb_kf_csa_run(uint32_t *cw, const uint32_t *d) {
    uint32_t twoA, twoB, fourA, fourB, eightA;
    uint32_t ones = cw[0], twos = cw[1], fours = cw[2];
    FB_CSA(twoA, ones, ones, d[0], d[1]);
    FB_CSA(twoB, ones, ones, d[2], d[3]);
    FB_CSA(fourA, twos, twos, twoA, twoB);
    FB_CSA(twoA, ones, ones, d[4], d[5]);
    FB_CSA(twoB, ones, ones, d[6], d[7]);
    FB_CSA(fourB, twos, twos, twoA, twoB);
    FB_CSA(eightA, fours, fours, fourA, fourB);
#if FB_KF_RUN == 16
    uint32_t eightB, sixteen, eights = cw[3];
    FB_CSA(twoA, ones, ones, d[8], d[9]);
    FB_CSA(twoB, ones, ones, d[10], d[11]);
    FB_CSA(fourA, twos, twos, twoA, twoB);
    FB_CSA(twoA, ones, ones, d[12], d[13]);
    FB_CSA(twoB, ones, ones, d[14], d[15]);
    FB_CSA(fourB, twos, twos, twoA, twoB);
    FB_CSA(eightB, fours, fours, fourA, fourB);
    FB_CSA(sixteen, eights, eights, eightA, eightB);
    cw[3] = eights;
    uint32_t c = cw[4] & sixteen; cw[4] ^= sixteen;
#else
    // ripple the carry of weight 8 into the 8/16/32/64 words (half adders)
    uint32_t c0 = cw[3] & eightA; cw[3] ^= eightA;
    uint32_t c = cw[4] & c0; cw[4] ^= c0;
#endif
    cw[0] = ones; cw[1] = twos; cw[2] = fours;
    uint32_t c2 = cw[5] & c;      cw[5] ^= c;
    cw[6] ^= c2;
}

// sum_t (u_t >> p) of one unit from its counter words: in registers (cw[j]) or in shared memory
// ([word][unit], stride U).  The 32-bit forms are exact when every u < 2^24 (112 * 2^24 < 2^31).
FB_DEV unsigned long long fb_kf_evalr64(const uint32_t *cw, int p) {
    unsigned long long s = 0;
#pragma unroll
    for (int j = 0; j < FB_KF_NWORDS; j++) s += (unsigned long long)(cw[j] >> p) << j;
    return s;
}
FB_DEV uint32_t fb_kf_evalr32(const uint32_t *cw, int p) {
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < FB_KF_NWORDS; j++) s += (cw[j] >> p) << j;
    return s;
}
#if FB_GPU
static __device__ __noinline__
#else
inline
#endif
unsigned long long fb_kf_eval64(const uint32_t *words, int U, int unit, int p) {
    unsigned long long s = 0;
#pragma unroll
    for (int j = 0; j < FB_KF_NWORDS; j++) s += (unsigned long long)(words[j * U + unit] >> p) << j;
    return s;
}
FB_DEV uint32_t fb_kf_eval32(const uint32_t *words, int U, int unit, int p) {
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < FB_KF_NWORDS; j++) s += (words[j * U + unit] >> p) << j;
    return s;
}
FB_DEV unsigned long long fb_kf_eval(const uint32_t *words, int U, int unit, int p, bool small) {
    return small ? (unsigned long long)fb_kf_eval32(words, U, unit, p) : fb_kf_eval64(words, U, unit, p);
}

// Smallest minimiser of the convex cost f over [0, max_p], starting the walk at p (any start gives the same
// result).  Going up requires a strict decrease, going down accepts ties, so ties resolve to the smallest p
// like PrcBitTable::minimizer (src/rice.rs:117-141).
template <class F>
FB_DEV int fb_kf_walk(F f, int p, int max_p, unsigned long long *fmin) {
    // one evaluation site: first the start, then upwards, then (if the first step up failed) downwards
    unsigned long long fc = 0;
    int q = p, dir = 0; // dir 0: evaluating the start, +1: trying p + 1, -1: trying p - 1
    bool moved = false;
    for (;;) {
        const unsigned long long fq = f(q);
        if (dir == 0) {
            fc = fq;
            dir = 1;
        } else if (dir > 0) {
            if (fq < fc) { fc = fq; p = q; moved = true; }
            else if (moved) break;
            else dir = -1;
        } else {
            if (fq <= fc) { fc = fq; p = q; }
            else break;
        }
        if (dir > 0 && p >= max_p) {
            if (moved) break;
            dir = -1;
        }
        if (dir < 0 && p <= 0) break;
        q = p + dir;
    }
    *fmin = fc;
    return p;
}

// starting point of the walk: ~ log2(mean)
FB_DEV int fb_kf_pstart(unsigned long long s0, int cnt, int max_p) {
    int p = 0;
    if (cnt > 0 && s0 > (unsigned long long)cnt) p = (63 - fb_clz64(s0)) - (31 - fb_clz32((uint32_t)cnt));
    if (p < 0) p = 0;
    return p > max_p ? max_p : p;
}

// ---- residual runs out of the staged planes ------------------------------------------------------
// vm: 0 = plane xa as is, 2 = mid, 3 = side.  A run's window is win[i] = x[t0 - G + i], i < G + FB_KF_RUN:
// the G history samples are loaded once per unit and then slide in registers, the 16 new samples come
// from 16-byte shared-memory loads (scalar loads when the unit is not 4-aligned).

// Variant mode `vm`: bits 0-1 = 0 plain channel (xb == xa), 2 mid, 3 side; bit 2 = int16 planes.  Every sample is
// formed as (a + m * b) >> sh with (m, sh) = (0, 0), (1, 1), (-1, 0): one branch-free path for all variants
// (src/coding.rs:476-484 for M and S).
FB_DEV void fb_vm_mix(int vm, int32_t *m, int32_t *sh) {
    const int k = vm & 3;
    *m = k == 2 ? 1 : (k == 3 ? -1 : 0);
    *sh = k == 2 ? 1 : 0;
}

// The mixing constants of a variant mode, computed once per variant (not per load): planes: sample = (a + m * b) >> sh;
// pairs: sample = dp2a(pair, multiplier bytes) >> sh (fb_pair_mix)
struct FbMix { int32_t a, sh; };
template <int VMS>
FB_DEV FbMix fb_mix_of(int vm) {
    FbMix mx;
    if ((VMS & FB_VM_PAIRS) && (VMS == FB_VM_PAIRS || (vm & FB_VM_PAIRS))) fb_pair_mix(vm & 3, &mx.a, &mx.sh);
    else fb_vm_mix(vm, &mx.a, &mx.sh);
    return mx;
}

// four samples at plane offset o = fb_xidx(t) (t a multiple of 4, inside the plane incl. its slack)

// VMS: the plane formats a caller can meet (the plan kernel only ever sees 32-bit planes: VMS = 0 drops the rest)
#define FB_VMS_ALL (FB_VM_X16 | FB_VM_PAIRS)
template <int VMS = FB_VMS_ALL>
FB_DEV void fb_kf_load4_at(const int32_t *xa, const int32_t *xb, int vm, const FbMix &mx, int o, int32_t *dst) {
    const int32_t m = mx.a, sh = mx.sh;
    if ((VMS & FB_VM_PAIRS) && (VMS == FB_VM_PAIRS || (vm & FB_VM_PAIRS))) {
        const int32_t mb = mx.a;
        const int4 w = *reinterpret_cast<const int4 *>(xa + o);
        dst[0] = fb_dp2a_lo(w.x, mb, 0) >> sh;
        dst[1] = fb_dp2a_lo(w.y, mb, 0) >> sh;
        dst[2] = fb_dp2a_lo(w.z, mb, 0) >> sh;
        dst[3] = fb_dp2a_lo(w.w, mb, 0) >> sh;
    } else if ((VMS & FB_VM_X16) && (vm & FB_VM_X16)) {
        const int2 wa = *reinterpret_cast<const int2 *>(reinterpret_cast<const int16_t *>(xa) + o);
        const int2 wb = *reinterpret_cast<const int2 *>(reinterpret_cast<const int16_t *>(xb) + o);
        dst[0] = fb_mix((int32_t)(int16_t)(uint32_t)wa.x, (int32_t)(int16_t)(uint32_t)wb.x, m, sh);
        dst[1] = fb_mix(wa.x >> 16, wb.x >> 16, m, sh);
        dst[2] = fb_mix((int32_t)(int16_t)(uint32_t)wa.y, (int32_t)(int16_t)(uint32_t)wb.y, m, sh);
        dst[3] = fb_mix(wa.y >> 16, wb.y >> 16, m, sh);
    } else {
        const int4 wa = *reinterpret_cast<const int4 *>(xa + o);
        const int4 wb = *reinterpret_cast<const int4 *>(xb + o);
        dst[0] = fb_mix(wa.x, wb.x, m, sh);
        dst[1] = fb_mix(wa.y, wb.y, m, sh);
        dst[2] = fb_mix(wa.z, wb.z, m, sh);
        dst[3] = fb_mix(wa.w, wb.w, m, sh);
    }
}

template <int VMS = FB_VMS_ALL>
FB_DEV void fb_kf_load4(const int32_t *xa, const int32_t *xb, int vm, const FbMix &mx, int t, int32_t *dst) {
    fb_kf_load4_at<VMS>(xa, xb, vm, mx, fb_xidx(t), dst);
}

template <int VMS = FB_VMS_ALL>
FB_DEV int32_t fb_kf_load1(const int32_t *xa, const int32_t *xb, int vm, const FbMix &mx, int t) {
    const int o = fb_xidx(t);
    const int32_t m = mx.a, sh = mx.sh;
    if ((VMS & FB_VM_PAIRS) && (VMS == FB_VM_PAIRS || (vm & FB_VM_PAIRS))) {
        const int32_t mb = mx.a;
        return fb_dp2a_lo(xa[o], mb, 0) >> sh;
    }
    if ((VMS & FB_VM_X16) && (vm & FB_VM_X16))
        return fb_mix(reinterpret_cast<const int16_t *>(xa)[o], reinterpret_cast<const int16_t *>(xb)[o], m, sh);
    return fb_mix(xa[o], xb[o], m, sh);
}

// win[idx] = v for a run-time idx: a chain of predicated moves, so the window stays in registers
template <int G>
FB_DEV void fb_kf_win_set(int32_t *win, int idx, int32_t v) {
#pragma unroll
    for (int j = 0; j < G + FB_KF_RUN; j++) win[j] = (j == idx) ? v : win[j];
}

// ODD instantiations serve frames whose units do not start on multiples of 4 samples (finest partitions of odd
// tail frames): the same arithmetic with sample-by-sample window loads.  They are separate template instances, so
// the hot loops of the regular frames carry none of that code.
// win[0..G) = x[ta - G .. ta), zeros before the start of the frame (ta a multiple of 4 unless ODD)
template <int G, int VMS = FB_VMS_ALL, bool ODD = false>
FB_DEV void fb_kf_history(const int32_t *xa, const int32_t *xb, int vm, const FbMix &mx, int ta, int32_t *win) {
    if (ODD) {
#pragma unroll 1
        for (int i = 0; i < G; i++) fb_kf_win_set<G>(win, i, ta - G + i >= 0 ? fb_kf_load1<VMS>(xa, xb, vm, mx, ta - G + i) : 0);
        return;
    }
#pragma unroll
    for (int i = 0; i < G; i += 4) {
        const int t = ta - G + i;
        if (t >= 0) fb_kf_load4<VMS>(xa, xb, vm, mx, t, win + i);
        else { win[i] = 0; win[i + 1] = 0; win[i + 2] = 0; win[i + 3] = 0; }
    }
}

// win[G..G+RUN) = x[t0 .. t0+RUN) (t0 a multiple of 4 unless ODD); samples at t >= n are don't-cares (masked)
template <int G, int VMS = FB_VMS_ALL, bool ODD = false>
FB_DEV void fb_kf_fetch_run(const int32_t *xa, const int32_t *xb, int vm, const FbMix &mx, int t0, int32_t *win) {
    if (ODD) {
#pragma unroll 1
        for (int i = 0; i < FB_KF_RUN; i++) fb_kf_win_set<G>(win, G + i, fb_kf_load1<VMS>(xa, xb, vm, mx, t0 + i));
        return;
    }
    if ((t0 & 15) == 0) {
        // a run that starts on a multiple of 16 lies inside one 64-sample block of the padded plane: one index
        const int o = fb_xidx(t0);
#pragma unroll
        for (int i = 0; i < FB_KF_RUN; i += 4) fb_kf_load4_at<VMS>(xa, xb, vm, mx, o + i, win + G + i);
    } else {
#pragma unroll
        for (int i = 0; i < FB_KF_RUN; i += 4) fb_kf_load4<VMS>(xa, xb, vm, mx, t0 + i, win + G + i);
    }
}

template <int G>
FB_DEV void fb_kf_slide(int32_t *win) {
#pragma unroll
    for (int i = 0; i < G; i++) win[i] = win[i + FB_KF_RUN];
}

// description of one residual candidate of a variant
struct FbKfCand {
    int kind, order, shift; // kind 0: fixed predictor of `order`, 1: LPC
    bool narrow;            // LPC: the reference's i32 accumulation is safe (src/lpc.rs:361-374), else 64-bit
    int32_t fc[4];          // fixed: e[t] = x[t] + fc0 x[t-1] + fc1 x[t-2] + fc2 x[t-3] + fc3 x[t-4] (wrapping i32)
    const int16_t *q;
};

FB_DEV void fb_kf_fixed_coefs(int order, int32_t *fc) {
    // zero-history k-th differences (src/coding.rs:182-197) == binomial predictors
    // (-1)^j C(k, j):  k=1: -1 | k=2: -2 1 | k=3: -3 3 -1 | k=4: -4 6 -4 1
    const int k = order < 0 ? 0 : (order > 4 ? 4 : order);
    fc[0] = -k;
    fc[1] = k * (k - 1) / 2;
    fc[2] = -(k * (k - 1) * (k - 2)) / 6;
    fc[3] = k == 4 ? 1 : 0;
}

// zigzag residuals of the run t0..t0+15 from its window; samples outside [lo, hi) give 0.
// ROT: the zigzag value rotated right by one bit instead, v ^ ((v >> 31) & 0x7FFFFFFF) = sign in bit 31 above |v| or
// |v| - 1 -- two instructions instead of three.  The plan kernel only counts bit planes, so it rotates the counter words
// of a unit back (bit plane b of the rotated values is plane b + 1 of the zigzag values) instead of every sample.
template <bool ROT>
FB_DEV uint32_t fb_kf_zz(int32_t v) {
    return ROT ? ((uint32_t)v ^ ((uint32_t)(v >> 31) & 0x7FFFFFFFu)) : fb_zigzag(v);
}
template <int G, bool ROT = false>
FB_DEV void fb_kf_run_u(const int32_t *win, int t0, int lo, int hi, const FbKfCand &cd, const int32_t *qq, uint32_t *u) {
    // bit i of vmask: sample t0 + i lies in [lo, hi)
    const int a = lo - t0, b = hi - t0;
    const uint32_t full = (1u << FB_KF_RUN) - 1u;
    const uint32_t m_hi = b >= FB_KF_RUN ? full : (b <= 0 ? 0u : ((1u << b) - 1u));
    const uint32_t m_lo = a <= 0 ? 0u : (a >= FB_KF_RUN ? full : ((1u << a) - 1u));
    const uint32_t vmask = m_hi & ~m_lo;
    if (cd.kind == 0) {
#pragma unroll
        for (int i = 0; i < FB_KF_RUN; i++) {
            const uint32_t e = (uint32_t)win[G + i] + (uint32_t)cd.fc[0] * (uint32_t)win[G + i - 1] +
                               (uint32_t)cd.fc[1] * (uint32_t)win[G + i - 2] + (uint32_t)cd.fc[2] * (uint32_t)win[G + i - 3] +
                               (uint32_t)cd.fc[3] * (uint32_t)win[G + i - 4];
            u[i] = fb_kf_zz<ROT>((int32_t)e);
        }
    } else if (cd.narrow && G >= 8 && cd.order <= G - 2) {
        // the common orders (e.g. 10 with G = 12) leave the last two taps zero: skip them
#pragma unroll
        for (int i = 0; i < FB_KF_RUN; i++) {
            uint32_t acc = 0;
#pragma unroll
            for (int j = 0; j < G - 2; j++) acc += (uint32_t)qq[j] * (uint32_t)win[G + i - 1 - j];
            u[i] = fb_kf_zz<ROT>((int32_t)((uint32_t)win[G + i] - (uint32_t)((int32_t)acc >> cd.shift)));
        }
    } else if (cd.narrow) {
#pragma unroll
        for (int i = 0; i < FB_KF_RUN; i++) {
            uint32_t acc = 0;
#pragma unroll
            for (int j = 0; j < G; j++) acc += (uint32_t)qq[j] * (uint32_t)win[G + i - 1 - j];
            u[i] = fb_kf_zz<ROT>((int32_t)((uint32_t)win[G + i] - (uint32_t)((int32_t)acc >> cd.shift)));
        }
    } else {
#pragma unroll
        for (int i = 0; i < FB_KF_RUN; i++) {
            int64_t acc = 0;
#pragma unroll
            for (int j = 0; j < G; j++) acc = fb_mad_wide(qq[j], win[G + i - 1 - j], acc);
            u[i] = fb_kf_zz<ROT>((int32_t)(uint32_t)((uint64_t)(int64_t)win[G + i] - (uint64_t)(acc >> cd.shift)));
        }
    }
    if (vmask != full) {
#pragma unroll
        for (int i = 0; i < FB_KF_RUN; i++) u[i] = ((vmask >> i) & 1u) ? u[i] : 0u;
    }
}

// =====================================================================================================
// Rice search of one candidate by one warp.  xa/xb/vm select the variant's samples.  Writes res and
// unit_bits[0..U) (bits of every unit without the parameter fields).
// Sets M->fail when the frame must be redone by the literal path.
// =====================================================================================================
template <int G, bool ODD, int VMS>
FB_DEV void fb_kf_search(const FbJob &J, const FbKfGeom &g, const int32_t *xa, const int32_t *xb, int vm,
                         const FbKfCand &cd, uint8_t *scratch, const FbKfLayout &L, uint32_t *unit_bits, FbKfRes *res) {
    uint32_t *words = (uint32_t *)(scratch + L.s_words);
    uint32_t *tbl_a = (uint32_t *)(scratch + L.s_tbl_a);
    uint32_t *best_val = (uint32_t *)(scratch + L.s_best_val);
    uint8_t *best_p = scratch + L.s_best_p;
    unsigned long long *lvl_bits = (unsigned long long *)(scratch + L.s_lvl_bits);
    FbKfMisc *M = (FbKfMisc *)(scratch + L.s_misc);
    const int n = g.n, U = g.U, warm = cd.order, max_p = J.cfg.prc_max_parameter;
    const int slots = U >> 5;

    FB_WPHASE(lane)
        if (lane == 0) { M->ormask = 0; M->pmin = 31; M->pmax = 0; M->any_gt14 = 0; }
    FB_WPHASE_END

    // ---- pass 1: residuals -> bit-sliced counters per unit (+ pass 2 in registers when unit == leaf)
    FB_WPHASE(lane)
        const FbMix mx = fb_mix_of<VMS>(vm);
        int32_t qq[G];
#pragma unroll
        for (int j = 0; j < G; j++) qq[j] = (cd.kind == 1 && j < cd.order) ? (int32_t)cd.q[j] : 0;
        uint32_t orm = 0;
        for (int s = 0; s < slots; s++) {
            const int unit = s * 32 + lane;
            int ta, tb;
            fb_kf_unit_range(g, unit, &ta, &tb);
            uint32_t cw[FB_KF_NWORDS];
#pragma unroll
            for (int j = 0; j < FB_KF_NWORDS; j++) cw[j] = 0;
            const int lo = ta > warm ? ta : warm;
            if (tb > ta) {
                int32_t win[G + FB_KF_RUN];
                fb_kf_history<G, VMS, ODD>(xa, xb, vm, mx, ta, win);
                for (int t0 = ta; t0 < tb; t0 += FB_KF_RUN) {
                    uint32_t uu[FB_KF_RUN];
                    fb_kf_fetch_run<G, VMS, ODD>(xa, xb, vm, mx, t0, win);
                    fb_kf_run_u<G, true>(win, t0, lo, tb, cd, qq, uu);
                    fb_kf_csa_run(cw, uu);
                    fb_kf_slide<G>(win);
                }
#pragma unroll
                for (int j = 0; j < FB_KF_NWORDS; j++) cw[j] = (cw[j] << 1) | (cw[j] >> 31); // back to zigzag bit planes
            }
            uint32_t orm_u = 0;
#pragma unroll
            for (int j = 0; j < FB_KF_NWORDS; j++) { words[j * U + unit] = cw[j]; orm_u |= cw[j]; }
            orm |= orm_u;
            if (g.m == 1) {
                // the unit is a finest partition: its exact minimiser, from the registers
                const int cnt = g.leaf_len - (unit == 0 ? warm : 0);
                unsigned long long fmin;
                int p;
                if (orm_u < (1u << 24)) {
                    p = fb_kf_pstart(fb_kf_evalr32(cw, 0), cnt, max_p);
                    p = fb_kf_walk([&](int pp) { return (unsigned long long)(fb_kf_evalr32(cw, pp) + (uint32_t)cnt * (uint32_t)(pp + 1)); },
                                   p, max_p, &fmin);
                } else {
                    p = fb_kf_pstart(fb_kf_evalr64(cw, 0), cnt, max_p);
                    p = fb_kf_walk([&](int pp) { return fb_kf_evalr64(cw, pp) + (unsigned long long)cnt * (unsigned long long)(pp + 1); },
                                   p, max_p, &fmin);
                }
                if (fmin + 4ull >= (unsigned long long)FB_RICE_SAT) M->fail = 1; // benign race: every writer stores 1
                fb_atomic_min_u32(&M->pmin, (uint32_t)p);
                fb_atomic_max_u32(&M->pmax, (uint32_t)p);
            }
        }
        if (orm) fb_atomic_or_u32(&M->ormask, orm);
    FB_WPHASE_END

    // a residual >= 2^27: the reference's chunked saturating sums are order dependent -> literal path
    if (M->ormask >= (1u << 27)) {
        FB_WPHASE(lane)
            if (lane == 0) M->fail = 1;
        FB_WPHASE_END
        return;
    }
    const bool small = M->ormask < (1u << 24);

    // ---- pass 2 (units smaller than a finest partition): minimiser per partition from shared memory
    if (g.m > 1) {
        FB_WPHASE(lane)
            for (int leaf = lane; leaf < g.leaves; leaf += 32) {
                const int cnt = g.leaf_len - (leaf == 0 ? warm : 0);
                auto cost = [&](int pp) {
                    unsigned long long f = (unsigned long long)cnt * (unsigned long long)(pp + 1);
                    for (int k = 0; k < g.m; k++) f += fb_kf_eval(words, U, leaf * g.m + k, pp, small);
                    return f;
                };
                unsigned long long fmin;
                int p = fb_kf_pstart(cost(0) - (unsigned long long)cnt, cnt, max_p);
                p = fb_kf_walk(cost, p, max_p, &fmin);
                if (fmin + 4ull >= (unsigned long long)FB_RICE_SAT) M->fail = 1;
                fb_atomic_min_u32(&M->pmin, (uint32_t)p);
                fb_atomic_max_u32(&M->pmax, (uint32_t)p);
            }
        FB_WPHASE_END
    }
    if (M->fail) return;

    // ---- pass 3: partition tree over [pmin, pmax], FB_KF_COLS parameters at a time.  One lane per node:
    // it forms the node's row (leaf: from the counters; inner node: min(a + b - 4, 2^27-1), src/rice.rs:144-152)
    // and keeps the node's minimum; windows are visited in ascending p, so ties keep the smallest p.
    const int pa0 = (int)M->pmin, pb = (int)M->pmax;
    for (int pa = pa0; pa <= pb; pa += FB_KF_COLS) {
        const int Wc = (pb - pa + 1) < FB_KF_COLS ? (pb - pa + 1) : FB_KF_COLS;
        const bool first = pa == pa0;
        FB_WPHASE(lane)
            for (int leaf = lane; leaf < g.leaves; leaf += 32) {
                const int cnt = g.leaf_len - (leaf == 0 ? warm : 0);
                uint32_t bv = 0xFFFFFFFFu;
                int bp = pa;
                for (int j = 0; j < Wc; j++) {
                    const int p = pa + j;
                    unsigned long long f = 4ull + (unsigned long long)cnt * (unsigned long long)(p + 1);
                    for (int k = 0; k < g.m; k++) f += fb_kf_eval(words, U, leaf * g.m + k, p, small);
                    const uint32_t v = f > FB_RICE_SAT ? FB_RICE_SAT : (uint32_t)f;
                    tbl_a[leaf * FB_KF_ROW + j] = v;
                    if (v < bv) { bv = v; bp = p; }
                }
                const int idx = g.leaves - 1 + leaf;
                if (first || bv < best_val[idx]) { best_val[idx] = bv; best_p[idx] = (uint8_t)bp; }
            }
        FB_WPHASE_END
        // the levels are merged IN PLACE in one table: node i of a level reads rows 2i and 2i + 1 and writes row i.  Nodes
        // are taken 32 at a time in ascending order, all reads of a batch before its writes: a batch starting at node b
        // reads rows >= 2b, and everything written so far lies below row b + 32 <= 2b for b >= 32 (for b = 0 the batch's
        // own reads are done before its writes).
        for (int lvl = g.o0 - 1; lvl >= 0; lvl--) {
            const int nodes = 1 << lvl;
            for (int node0 = 0; node0 < nodes; node0 += 32) {
                FB_WPHASE(lane)
                    const int node = node0 + lane;
                    uint32_t v[FB_KF_COLS];
                    uint32_t bv = 0xFFFFFFFFu;
                    int bp = pa;
                    if (node < nodes) {
                        const uint32_t *ra = tbl_a + (2 * node) * FB_KF_ROW, *rb = ra + FB_KF_ROW;
                        for (int j = 0; j < Wc; j++) {
                            uint32_t x = ra[j] + rb[j] - 4u;
                            x = x < FB_RICE_SAT ? x : FB_RICE_SAT;
                            v[j] = x;
                            if (x < bv) { bv = x; bp = pa + j; }
                        }
                    }
#if FB_GPU
                    __syncwarp(); // (the emulation runs the lanes in ascending order, which is safe as it is: lane k writes
                                  //  row k after reading rows 2k and 2k + 1, and no lane below k reads row k later)
#endif
                    if (node < nodes) {
                        uint32_t *ro = tbl_a + node * FB_KF_ROW;
                        for (int j = 0; j < Wc; j++) ro[j] = v[j];
                        const int idx = nodes - 1 + node;
                        if (first || bv < best_val[idx]) { best_val[idx] = bv; best_p[idx] = (uint8_t)bp; }
                    }
                FB_WPHASE_END
            }
        }
    }

    // ---- pass 4a: totals per level.  GPU: two warp-wide integer reductions per level (low / high 16 bits of the lane
    // partials, each < 2^30); the emulation exchanges the partials through shared memory instead.
#if FB_GPU
    for (int lvl = 0; lvl <= g.o0; lvl++) {
        const int nodes = 1 << lvl;
        const int lane = (int)(threadIdx.x & 31u);
        uint32_t part = 0; // <= 8 nodes per lane, each < 2^27
        for (int node = lane; node < nodes; node += 32) {
            const uint32_t v = best_val[nodes - 1 + node];
            if (v >= FB_RICE_SAT) M->fail = 1;
            part += v;
        }
        const uint32_t lo = __reduce_add_sync(0xFFFFFFFFu, part & 0xFFFFu);
        const uint32_t hi = __reduce_add_sync(0xFFFFFFFFu, part >> 16);
        if (lane == 0) lvl_bits[lvl] = ((unsigned long long)hi << 16) + (unsigned long long)lo;
    }
    __syncwarp();
#else
    uint32_t *lx = tbl_a; // (o0 + 1) x 32 exchange words
    FB_WPHASE(lane)
        for (int lvl = 0; lvl <= g.o0; lvl++) {
            const int nodes = 1 << lvl;
            uint32_t part = 0; // <= 8 nodes per lane, each < 2^27
            for (int node = lane; node < nodes; node += 32) {
                const uint32_t v = best_val[nodes - 1 + node];
                if (v >= FB_RICE_SAT) M->fail = 1;
                part += v;
            }
            lx[lvl * 32 + lane] = part;
        }
    FB_WPHASE_END
    FB_WPHASE(lane)
        if (lane <= g.o0) {
            unsigned long long s = 0;
            for (int i = 0; i < 32; i++) s += lx[lane * 32 + ((i + lane) & 31)];
            lvl_bits[lane] = s;
        }
    FB_WPHASE_END
#endif
    if (M->fail) return;
    // ---- pass 4b: partition order: strictly smaller total wins while going coarser (src/rice.rs:276-291)
    FB_WPHASE(lane)
        if (lane == 0) {
            unsigned long long min_bits = lvl_bits[g.o0];
            int best = g.o0;
            for (int lvl = g.o0 - 1; lvl >= 0; lvl--)
                if (lvl_bits[lvl] < min_bits) { min_bits = lvl_bits[lvl]; best = lvl; }
            M->best_level = best;
            res->part_order = best;
        }
    FB_WPHASE_END
    // ---- pass 4c: parameters, bits of every unit (for the packer)
    {
        const int best = M->best_level;
        const int nparts = 1 << best;
        const int ush = g.lgU - best; // units per partition = 1 << ush
        FB_WPHASE(lane)
            for (int j = lane; j < nparts; j += 32) {
                const uint8_t p = best_p[nparts - 1 + j];
                res->params[j] = p;
                if (p > 14) M->any_gt14 = 1;
            }
            for (int unit = lane; unit < U; unit += 32) {
                const int p = best_p[nparts - 1 + (unit >> ush)];
                int ta, tb;
                fb_kf_unit_range(g, unit, &ta, &tb);
                if (ta < warm) ta = warm;
                const int cnt = tb > ta ? tb - ta : 0;
                unit_bits[unit] = (uint32_t)(fb_kf_eval(words, U, unit, p, small) + (unsigned long long)cnt * (unsigned long long)(p + 1));
            }
        FB_WPHASE_END
        FB_WPHASE(lane)
            if (lane == 0) {
                const int rice2 = M->any_gt14 ? 1 : 0;
                res->rice2 = rice2;
                // src/component/bitrepr.rs:532-544: 2 + 4 + parts*(4|5) + sum q + (n - w) + sum p*len - w*p0.
                // Every node minimum is S + cnt*(p+1) + 4 (unsaturated, checked above), so the code bits of the
                // chosen partitioning are the level total minus the 4-bit offsets.
                res->res_bits = 6ull + (unsigned long long)nparts * (rice2 ? 5ull : 4ull) + lvl_bits[best] -
                                4ull * (unsigned long long)nparts;
            }
        FB_WPHASE_END
    }
}

// =====================================================================================================
// One variant (one warp): fixed_lpc / estimated_qlpc / encode_subframe (src/coding.rs:298-418) on top
// of K1's analysis.  Writes the decision record `out` (shared memory) and S->cand[v].
// =====================================================================================================
template <int G, bool ODD, int VMS, bool BC>
FB_DEV void fb_kf_variant(const FbJob &J, const FbKfGeom &g, const int32_t *xs, const FbAnalysis &A, int v,
                          uint8_t *smem, const FbKfLayout &L, fb200_subframe_info *out, const FbLpcExt *ext = nullptr) {
    FbKfFrame *S = (FbKfFrame *)(smem + L.off_frame);
    uint8_t *scratch = smem + L.off_scratch + (uint32_t)v * L.scratch_bytes;
    uint8_t *keep = smem + L.off_keep + (uint32_t)v * L.keep_bytes;
    uint32_t *unit_bits = (uint32_t *)(keep + L.k_unit_bits);
    FbKfRes *res[2] = {(FbKfRes *)(keep + L.k_res), (FbKfRes *)(keep + L.k_res + L.res_stride)};
    FbKfMisc *M = (FbKfMisc *)(scratch + L.s_misc);
    const int n = g.n;
    const int bps_v = fb_variant_bps(J, v);
    const unsigned long long verbatim_bits = 8ull + (unsigned long long)n * (unsigned long long)bps_v;
    // sample planes of this variant: 32-bit planes, or (VMS = FB_VM_PAIRS) the one plane of 16-bit stereo pairs
    int vm = 0;
    const int32_t *xa = xs + (size_t)v * L.x_stride, *xb = xa;
    if (VMS == FB_VM_PAIRS) { vm = FB_VM_PAIRS | v; xa = xb = xs; }
    else if (J.channels == 2 && v >= 2) { vm |= v; xa = xs; xb = xs + L.x_stride; }

    FB_WPHASE(lane)
        if (lane == 0) {
            M->fail = 0;
            out->type = FB200_SF_VERBATIM;
            out->order = 0;
            out->bits_per_sample = bps_v;
            out->precision = 0;
            out->shift = 0;
            out->partition_order = 0;
            out->rice2 = 0;
            out->reserved = n;
            out->bits = verbatim_bits;
            FbKfVHead &H = S->vh[v];
            H.bits = verbatim_bits; H.type = FB200_SF_VERBATIM; H.order = 0; H.bps = (int16_t)bps_v; H.precision = 0; H.shift = 0;
            H.part_order = 0; H.rice2 = 0; H.pad = 0;
            S->cand[v] = 0;
        }
        out->qlp[lane] = 0;
    FB_WPHASE_END

    if (J.cfg.use_constant && A.is_constant) {
        FB_WPHASE(lane)
            if (lane == 0) {
                out->type = FB200_SF_CONSTANT; out->bits = 8ull + (unsigned long long)bps_v;
                S->vh[v].type = FB200_SF_CONSTANT; S->vh[v].bits = 8ull + (unsigned long long)bps_v;
            }
        FB_WPHASE_END
        return;
    }
    if (n < FB_MIN_PRED_BLOCK) return;

    // candidates: 0 = fixed (src/coding.rs:298-331), 1 = LPC (src/coding.rs:360-381).  The fixed order is the winner of
    // K1's entropy estimate (OrderSel::ApproxEnt) or, for OrderSel::BitCount (src/coding.rs:241-262), the first order
    // minimising bps * order + code_bits of its exact Rice search: the orders are searched one after the other into
    // result set 0 and the winner is searched again below unless it was the last one.  BC is a template parameter: the
    // plan kernels of the default selector do not contain this loop at all (they are sensitive to their code size).
    int kf = J.cfg.use_fixed ? A.fixed_order : -1;
    int set0_order = -1;
    if (BC && J.cfg.use_fixed && J.cfg.fixed_order_sel == 0) {
        const int n_probe = (J.cfg.fixed_max_order < 4 ? J.cfg.fixed_max_order : 4) + 1;
        unsigned long long best_key = 0;
        kf = -1;
#if FB_GPU
#pragma unroll 1
#endif
        for (int k = 0; k < n_probe; k++) {
            FbKfCand cd;
            cd.kind = 0; cd.q = A.qlp; cd.order = k; cd.shift = 0; cd.narrow = true;
            fb_kf_fixed_coefs(k, cd.fc);
            fb_kf_search<G, ODD, VMS>(J, g, xa, xb, vm, cd, scratch, L, unit_bits, res[0]);
            if (M->fail) return;
            set0_order = k;
            // code_bits of the search = the level total = res_bits without the 6 header bits and the RICE2 extra bits
            const unsigned long long code_bits =
                res[0]->res_bits - 6ull - (res[0]->rice2 ? (unsigned long long)(1 << res[0]->part_order) : 0ull);
            const unsigned long long key = (unsigned long long)bps_v * (unsigned long long)k + code_bits;
            if (kf < 0 || key < best_key) { kf = k; best_key = key; }
        }
        if (!(best_key < verbatim_bits)) kf = -1;
    }
    unsigned long long cbits[2] = {0, 0};
#if FB_GPU
#pragma unroll 1
#endif
    for (int c = 0; c < 2; c++) {
        if (c == 0 ? (kf < 0) : !J.cfg.use_lpc) continue;
        if (BC && c == 0 && set0_order == kf) { // (the last order probed won, its results are in place)
            cbits[0] = 8ull + (unsigned long long)bps_v * (unsigned long long)kf + res[0]->res_bits;
            continue;
        }
        FbKfCand cd;
        cd.kind = c;
        cd.q = A.qlp;
        if (c == 0) {
            cd.order = kf; cd.shift = 0; cd.narrow = true;
            fb_kf_fixed_coefs(kf, cd.fc);
        } else {
            cd.order = A.qlp_order; cd.shift = A.qlp_shift;
            cd.fc[0] = cd.fc[1] = cd.fc[2] = cd.fc[3] = 0;
            // src/lpc.rs:361-374: i32 accumulation when max|x| * sum|q| < 2^31 - 1
            unsigned long long sumabs = 0;
            for (int j = 0; j < A.qlp_order; j++) sumabs += (unsigned long long)(A.qlp[j] < 0 ? -A.qlp[j] : A.qlp[j]);
            cd.narrow = (unsigned long long)A.max_abs * sumabs < 0x7FFFFFFFull;
        }
        fb_kf_search<G, ODD, VMS>(J, g, xa, xb, vm, cd, scratch, L, unit_bits + (size_t)c * (L.U_max + 1), res[c]);
        if (M->fail) return;
        cbits[c] = c == 0 ? 8ull + (unsigned long long)bps_v * (unsigned long long)kf + res[0]->res_bits
                          : 8ull + (unsigned long long)bps_v * (unsigned long long)A.qlp_order + 4ull + 5ull +
                                (unsigned long long)J.cfg.quant_precision * (unsigned long long)A.qlp_order + res[1]->res_bits;
    }
    int win_prec = J.cfg.quant_precision; // (BC: precision of the winning LPC set)
    if (BC && ext && J.cfg.use_lpc) {
        // EXTENSION (config.ext_lpc_order_search): the lower-order coefficient sets K1 left in `ext` are searched into
        // result set 1 one after the other; the fewest subframe bits win (the higher order on ties), the winner is
        // searched again unless it was the last one, and its coefficients replace the staged analysis record.
        int best = -1, last = -1; // -1: the primary set
        unsigned long long best_bits = cbits[1];
#if FB_GPU
#pragma unroll 1
#endif
        for (int k = 0; k <= FB_EXT_LPC_MAX; k++) {
            int pick_set;
            if (k < FB_EXT_LPC_MAX && ext[k].order > 0) pick_set = k;       // probe set k
            else if (best != last) pick_set = best;                           // the winner again, then done
            else break;
            const int16_t *qs = pick_set < 0 ? A.qlp : ext[pick_set].qlp;
            const int o = pick_set < 0 ? A.qlp_order : ext[pick_set].order;
            const int sh = pick_set < 0 ? A.qlp_shift : ext[pick_set].shift;
            const int pr = pick_set < 0 ? J.cfg.quant_precision : ext[pick_set].precision;
            FbKfCand cd;
            cd.kind = 1; cd.q = qs; cd.order = o; cd.shift = sh;
            cd.fc[0] = cd.fc[1] = cd.fc[2] = cd.fc[3] = 0;
            unsigned long long sumabs = 0;
            for (int j = 0; j < o; j++) sumabs += (unsigned long long)(qs[j] < 0 ? -qs[j] : qs[j]);
            cd.narrow = (unsigned long long)A.max_abs * sumabs < 0x7FFFFFFFull;
            fb_kf_search<G, ODD, VMS>(J, g, xa, xb, vm, cd, scratch, L, unit_bits + (size_t)(L.U_max + 1), res[1]);
            if (M->fail) return;
            const unsigned long long bits = 8ull + (unsigned long long)bps_v * (unsigned long long)o + 4ull + 5ull +
                                            (unsigned long long)pr * (unsigned long long)o + res[1]->res_bits;
            const bool probing = k < FB_EXT_LPC_MAX && ext[k].order > 0;
            last = pick_set;
            if (!probing) break;
            if (bits < best_bits) { best = pick_set; best_bits = bits; }
        }
        if (best >= 0) {
            FbAnalysis &Aw = const_cast<FbAnalysis &>(A); // (the staged copy in shared memory; this warp owns it)
            FB_WPHASE(lane)
                Aw.qlp[lane] = ext[best].qlp[lane];
                if (lane == 0) { Aw.qlp_order = ext[best].order; Aw.qlp_shift = ext[best].shift; }
            FB_WPHASE_END
            cbits[1] = best_bits;
            win_prec = ext[best].precision;
        }
    }
    const unsigned long long fixed_bits = cbits[0], lpc_bits = cbits[1];
    const unsigned long long baseline_bits =
        kf >= 0 ? (fixed_bits < verbatim_bits ? fixed_bits : verbatim_bits) : verbatim_bits;
    const bool lpc_ok = J.cfg.use_lpc && lpc_bits < baseline_bits;

    // decision (src/coding.rs:403-416)
    int pick = -1;
    if (lpc_ok) pick = 1;
    else if (kf >= 0) pick = 0;
    if (pick == 1 && !(lpc_bits < verbatim_bits)) pick = -1;
    if (pick == 0 && !(fixed_bits < verbatim_bits)) pick = -1;
    if (pick < 0) return;

    const FbKfRes *R = res[pick];
    FB_WPHASE(lane)
        if (lane == 0) {
            out->type = pick == 1 ? FB200_SF_LPC : FB200_SF_FIXED;
            out->order = pick == 1 ? A.qlp_order : kf;
            out->precision = pick == 1 ? J.cfg.quant_precision : 0;
            out->shift = pick == 1 ? A.qlp_shift : 0;
            out->partition_order = R->part_order;
            out->rice2 = R->rice2;
            out->bits = pick == 1 ? lpc_bits : fixed_bits;
            FbKfVHead &H = S->vh[v];
            H.type = pick == 1 ? FB200_SF_LPC : FB200_SF_FIXED; H.order = (int16_t)(pick == 1 ? A.qlp_order : kf);
            H.precision = (int16_t)(pick == 1 ? J.cfg.quant_precision : 0);
            H.shift = (int16_t)(pick == 1 ? A.qlp_shift : 0); H.part_order = (int16_t)R->part_order; H.rice2 = (int16_t)R->rice2;
            H.bits = pick == 1 ? lpc_bits : fixed_bits;
            S->cand[v] = pick;
        }
        if (pick == 1) out->qlp[lane] = A.qlp[lane];
        for (int i = lane; i < (1 << R->part_order); i += 32) out->rice_params[i] = R->params[i];
    FB_WPHASE_END
    if (BC && pick == 1 && win_prec != J.cfg.quant_precision) { // (precision search extension: the winner's precision)
        FB_WPHASE(lane)
            if (lane == 0) { out->precision = win_prec; S->vh[v].precision = (int16_t)win_prec; }
        FB_WPHASE_END
    }
}

// ORs the k (1..32) low bits of v (v < 2^k) into the MSB-first bit stream at bit position pos.  The buffer is
// zero-initialised and every bit is written once, so concurrent writers only ever share words, never bits.
FB_DEV void fb_or_bits(uint32_t *words, uint32_t pos, uint32_t v, uint32_t k) {
    const uint32_t o = pos & 31u;
    const unsigned long long w = (unsigned long long)v << (64u - o - k);
    fb_atomic_or(&words[pos >> 5], (uint32_t)(w >> 32));
    const uint32_t lo = (uint32_t)w;
    if (lo) fb_atomic_or(&words[(pos >> 5) + 1u], lo);
}

// ---- MSB-first bit writer with a 64-bit accumulator (src/bitsink.rs semantics) ----------------------
// A writer owns the bit range [pos_start, pos_end) of the zero-initialised word buffer; only its first and
// last word can be shared with a neighbour (atomic OR), interior words are stored plainly.
struct FbBitW {
    uint32_t *words;
    uint32_t w_first, w_last, cur_w, fill; // fill < 32: valid bits at the top of acc
    unsigned long long acc;
};

FB_DEV void fb_bw_init(FbBitW &r, uint32_t *words, uint32_t pos_start, uint32_t pos_end) {
    r.words = words;
    r.w_first = pos_start >> 5;
    r.w_last = pos_end > pos_start ? ((pos_end - 1) >> 5) : r.w_first;
    r.cur_w = r.w_first;
    r.fill = pos_start & 31u;
    r.acc = 0;
}

FB_DEV void fb_bw_store(FbBitW &r, uint32_t w) {
    if (w) {
        if (r.cur_w == r.w_first || r.cur_w == r.w_last) fb_atomic_or(&r.words[r.cur_w], w);
        else r.words[r.cur_w] = w;
    }
}

// append the k (1..32) low bits of v (v < 2^k)
FB_DEV void fb_bw_put(FbBitW &r, uint32_t v, uint32_t k) {
    r.acc |= (unsigned long long)v << (64u - r.fill - k);
    r.fill += k;
    if (r.fill >= 32u) {
        fb_bw_store(r, (uint32_t)(r.acc >> 32));
        r.acc <<= 32;
        r.fill -= 32u;
        r.cur_w++;
    }
}

// append q zero bits
FB_DEV void fb_bw_skip(FbBitW &r, uint32_t q) {
    const uint32_t total = r.fill + q;
    if (total >= 32u) {
        fb_bw_store(r, (uint32_t)(r.acc >> 32));
        r.acc = 0; // fill < 32, so everything pending was in the high word
        r.cur_w += total >> 5;
    }
    r.fill = total & 31u;
}

FB_DEV void fb_bw_finish(FbBitW &r) {
    if (r.fill) fb_bw_store(r, (uint32_t)(r.acc >> 32));
    r.acc = 0;
    r.fill = 0;
}

// =====================================================================================================
// The fused path is two kernels, each one CTA (32 * nvar threads) per frame, so that all warps resident on an SM
// run the same (smaller) code -- a single kernel with every phase inlined thrashed the instruction caches:
//   KA  stage the channels, analyse every variant (one warp each), subframe + stereo decisions, frame header,
//       bit offsets of all units; writes the frame's PLAN (header, subframe records, unit offsets) and its size
//   KP  (after the scan of the frame sizes) stage the channels again, pack the frame from the plan into shared
//       memory, CRC-16, and store it at its final offset of the output stream (no per-frame slot, no gather)
// Frames KA cannot reproduce exactly are appended to the fallback list (plan state 1) and skipped by KP.
// =====================================================================================================
struct FbKfPlan {               // global, one per frame; the leading part mirrors FbKfFrame
    FbKfSub sub[FB200_MAX_CHANNELS];
    uint8_t header[16];
    int32_t header_len, ch_tag;
    uint32_t data_bytes;
    uint32_t state;             // 0: planned by KA, 1: left to the generic kernels
};

// stages channels [c0, c1) of frame f into planes 0 .. c1-c0-1
template <int G>
FB_DEV void fb_kf_stage(const FbJob &J, const int32_t *xv, uint32_t f, int n, int32_t *xs, const FbKfLayout &L, int tid, int T,
                        int c0, int c1) {
    // quad i of every channel together: the rows of one frame are adjacent in xt (one 32-byte sector for stereo)
    const int n4 = (n + 3) >> 2;
    if (L.x16) {
#pragma unroll 8
        for (int i = tid; i < n4; i += T) {
            for (int c = c0; c < c1; c++) {
                const int32_t *src = xv + fb_xt_off(J.stride, f * (uint32_t)J.channels + (uint32_t)c, 0);
                const int4 v = *reinterpret_cast<const int4 *>(src + fb_xt_quad(4 * i));
                int2 w;
                w.x = (int32_t)(((uint32_t)v.x & 0xFFFFu) | ((uint32_t)v.y << 16));
                w.y = (int32_t)(((uint32_t)v.z & 0xFFFFu) | ((uint32_t)v.w << 16));
                *reinterpret_cast<int2 *>(reinterpret_cast<int16_t *>(xs + (size_t)(c - c0) * L.x_stride) + fb_xidx(4 * i)) = w;
            }
        }
    } else {
        for (int i = tid; i < n4; i += T) {
            for (int c = c0; c < c1; c++) {
                const int32_t *src = xv + fb_xt_off(J.stride, f * (uint32_t)J.channels + (uint32_t)c, 0);
                fb_copy16_async(xs + (size_t)(c - c0) * L.x_stride + fb_xidx(4 * i), src + fb_xt_quad(4 * i));
            }
        }
    }
}

// stages frame f's packed 16-bit stereo PCM as one plane of (left, right) pairs
FB_DEV void fb_kp_stage_pairs(const FbJob &J, const uint8_t *pcm, uint32_t f, int n, int32_t *xs, int tid, int T) {
    const int32_t *src = reinterpret_cast<const int32_t *>(pcm + (size_t)f * (size_t)J.block_size * 4u);
    const int nq = n >> 2;
    for (int i = tid; i < nq; i += T) fb_copy16_async(xs + fb_xidx(4 * i), src + 4 * i);
    if (tid < 4 && (n & 3)) { // partial last quad: the samples that exist, zeros behind them (like xt)
        const int t = 4 * nq + tid;
        if (t < n) fb_copy4_async(xs + fb_xidx(t), src + t);
        else xs[fb_xidx(t)] = 0;
    }
}

FB_DEV void fb_kf_to_fallback(uint32_t *fb_list, uint32_t *fb_count, FbKfPlan *plan, uint32_t f) {
#if FB_GPU
    const uint32_t slot_i = atomicAdd(fb_count, 1u);
#else
    const uint32_t slot_i = (*fb_count)++;
#endif
    fb_list[slot_i] = f;
    plan[f].state = 1;
}

// ---- KA: analysis and plan.  psubs: [frame][channels] chosen subframe records; poffs: [frame][channels][U_max+1]
// VMS = 0: 32-bit planes staged from the planar store xv; VMS = FB_VM_PAIRS: the frame's packed 16-bit stereo PCM
// (pcm) staged as one plane of pairs, xv is not read
template <int G, bool ODD = false, int VMS = 0, bool BC = false>
FB_DEV void fb_ka_body(const FbJob &J, const int32_t *xv, const uint8_t *pcm, const FbAnalysis *ana, FbKfPlan *plan, fb200_subframe_info *vsubs,
                       fb200_subframe_info *psubs, uint32_t *poffs, uint32_t *frame_bytes, fb200_frame_info *infos, uint32_t *fb_list,
                       uint32_t *fb_count, const uint32_t *ktab, uint32_t f, uint8_t *smem, const FbKfLayout &L) {
    const int NW = J.nvar;
    const int T = 32 * NW;
    const int n = fb_frame_len(J, f);
    const FbKfGeom g = fb_kf_geom(n);
    int32_t *xs = (int32_t *)(smem + L.off_x);
    // decision records of the frame's variants: global memory (written and re-read by this CTA only, so they stay in
    // L2; shared memory is what limits the CTAs per SM)
    fb200_subframe_info *choice = vsubs + (size_t)f * (size_t)J.nvar;
    FbKfFrame *S = (FbKfFrame *)(smem + L.off_frame);

    // units must start on multiples of 4 samples (16-byte window loads): frames whose finest partitions are not a
    // multiple of 4 long (odd tail frames, odd block sizes) are left to the generic kernels
    if (!ODD && (g.leaf_len & 3) != 0) {
        FB_PHASE(tid, T)
            if (tid == 0) fb_kf_to_fallback(fb_list, fb_count, plan, f);
        FB_PHASE_END
        return;
    }

    // ---- stage the independent channels from the row-interleaved store xt (16-byte asynchronous copies)
    FB_PHASE(tid, T)
        if (VMS == FB_VM_PAIRS) fb_kp_stage_pairs(J, pcm, f, n, xs, tid, T);
        else fb_kf_stage<G>(J, xv, f, n, xs, L, tid, T, 0, J.channels);
        {
            // K1's results for the frame's variants: read often and early, so they come along into shared memory
            static_assert(sizeof(FbAnalysis) % 8 == 0, "async copy granularity");
            const uint64_t *src = (const uint64_t *)(ana + (size_t)f * (size_t)J.nvar);
            uint64_t *dst = (uint64_t *)(smem + L.off_ana);
            for (int i = tid; i < J.nvar * (int)(sizeof(FbAnalysis) / 8); i += T) fb_copy8_async(dst + i, src + i);
        }
        if (tid == 0) S->frame_fail = 0;
        fb_copy_async_wait();
    FB_PHASE_END

    // ---- analysis: one warp per variant
    FB_WARPS_BEGIN(w, NW)
        fb_kf_variant<G, ODD, VMS, BC>(J, g, xs, ((const FbAnalysis *)(smem + L.off_ana))[w], w, smem, L, &choice[w],
                                       (BC && J.lpc_ext) ? J.lpc_ext + ((size_t)f * (size_t)J.nvar + (size_t)w) * FB_EXT_LPC_MAX : nullptr);
        const FbKfMisc *M = (const FbKfMisc *)(smem + L.off_scratch + (uint32_t)w * L.scratch_bytes + L.s_misc);
        FB_WPHASE(lane)
            if (lane == 0 && M->fail) S->frame_fail = 1; // benign race between warps
            if (lane == 0) choice[w].reserved = (int32_t)((const FbAnalysis *)(smem + L.off_ana))[w].max_abs;
        FB_WPHASE_END
    FB_WARPS_END

    if (S->frame_fail) {
        // not reproducible here: hand the frame to the literal kernels
        FB_PHASE(tid, T)
            if (tid == 0) fb_kf_to_fallback(fb_list, fb_count, plan, f);
        FB_PHASE_END
        return;
    }

    // ---- stereo decision, header, subframe offsets (thread 0)
    FB_PHASE(tid, T)
        if (tid == 0) {
            int ch_tag = J.channels - 1;
            int sel[FB200_MAX_CHANNELS];
            for (int c = 0; c < J.channels; c++) sel[c] = c;
            if (J.channels == 2) {
                // try_stereo_coding (src/coding.rs:469-527): strict <, order I, L/S, R/S, M/S
                const unsigned long long bl = S->vh[0].bits, br = S->vh[1].bits, bm = S->vh[2].bits, bs = S->vh[3].bits;
                unsigned long long min_bits = bl + br;
                if (J.cfg.use_leftside && bl + bs < min_bits) { min_bits = bl + bs; ch_tag = 8; }
                if (J.cfg.use_rightside && br + bs < min_bits) { min_bits = br + bs; ch_tag = 9; }
                if (J.cfg.use_midside && bm + bs < min_bits) { min_bits = bm + bs; ch_tag = 10; }
                if (ch_tag == 8) { sel[0] = 0; sel[1] = 3; }
                else if (ch_tag == 9) { sel[0] = 3; sel[1] = 1; }
                else if (ch_tag == 10) { sel[0] = 2; sel[1] = 3; }
            }
            S->ch_tag = ch_tag;
            S->header_len = fb_frame_header(n, ch_tag, J.bps, J.sample_rate, J.first_frame_number + f, S->header);
            uint32_t bit = (uint32_t)S->header_len * 8u;
            for (int c = 0; c < J.channels; c++) {
                const FbKfVHead &V = S->vh[sel[c]];
                FbKfSub &D = S->sub[c];
                D.variant = sel[c];
                D.cand = S->cand[sel[c]];
                D.type = V.type; D.order = V.order; D.bps = V.bps;
                D.precision = V.precision; D.shift = V.shift; D.part_order = V.part_order; D.rice2 = V.rice2;
                D.start_bit = bit;
                uint32_t hb = 8;
                if (V.type == FB200_SF_FIXED) hb += (uint32_t)(V.order * V.bps);
                if (V.type == FB200_SF_LPC)
                    hb += (uint32_t)(V.order * V.bps) + 4u + 5u + (uint32_t)(V.precision * V.order);
                D.res_bit = bit + hb;
                D.code_bit = D.res_bit + 6u;
                bit += (uint32_t)V.bits;
            }
            S->data_bytes = (bit + 7u) >> 3;
        }
    FB_PHASE_END

    // ---- the plan, in one warp-parallel block: warp c scans the unit bits of subframe c (exclusive scan incl. the
    // parameter fields) straight into the global unit offsets and copies the subframe's record; all warps share the
    // copies of the frame plan and of the optional info record
    FB_WARPS_BEGIN(w, NW)
        if (w < J.channels) {
            const FbKfSub &D = S->sub[w];
            if (D.type == FB200_SF_FIXED || D.type == FB200_SF_LPC) {
                uint8_t *keep = smem + L.off_keep + (uint32_t)D.variant * L.keep_bytes;
                const uint32_t *ub = (const uint32_t *)(keep + L.k_unit_bits) + (size_t)D.cand * (L.U_max + 1);
                uint32_t *xch = (uint32_t *)(smem + L.off_scratch + (uint32_t)w * L.scratch_bytes + L.s_tbl_a); // free again
                uint32_t *po = poffs + ((size_t)f * (size_t)J.channels + (size_t)w) * (L.U_max + 1);
                const int per = g.U >> 5;
                const int ush = g.lgU - D.part_order;
                const uint32_t pbits = D.rice2 ? 5u : 4u;
                FB_WPHASE(lane)
                    uint32_t s = 0;
                    for (int i = 0; i < per; i++) {
                        const int unit = lane * per + i;
                        s += ub[unit] + (((unit & ((1 << ush) - 1)) == 0) ? pbits : 0u);
                    }
                    xch[lane] = s;
                FB_WPHASE_END
                FB_WPHASE(lane)
                    uint32_t s = 0;
                    for (int i = 0; i < lane; i++) s += xch[i];
                    for (int i = 0; i < per; i++) {
                        const int unit = lane * per + i;
                        po[unit] = s;
                        s += ub[unit] + (((unit & ((1 << ush) - 1)) == 0) ? pbits : 0u);
                    }
                    if (lane == 31) po[g.U] = s;
                FB_WPHASE_END
            }
            FB_WPHASE(lane)
                const uint32_t *src = (const uint32_t *)&choice[D.variant];
                uint32_t *dst = (uint32_t *)&psubs[(size_t)f * (size_t)J.channels + (size_t)w];
                for (int i = lane; i < (int)(sizeof(fb200_subframe_info) / 4); i += 32) dst[i] = src[i];
            FB_WPHASE_END
        }
        FB_WPHASE(lane)
            const int tid = w * 32 + lane;
            if (tid == 0) frame_bytes[f] = S->data_bytes + 2u;
            {
                // FbKfPlan mirrors the head of FbKfFrame up to and including frame_fail (= state 0 here)
                const uint32_t *src = (const uint32_t *)S;
                uint32_t *dst = (uint32_t *)&plan[f];
                for (int i = tid; i < (int)(sizeof(FbKfPlan) / 4); i += T) dst[i] = src[i];
            }
            if (infos) {
                fb200_frame_info &I = infos[f];
                if (tid == 0) {
                    I.channel_assignment = S->ch_tag;
                    I.block_size = n;
                    I.frame_number = J.first_frame_number + f;
                    I.frame_bytes = S->data_bytes + 2u;
                }
                for (int c = 0; c < J.channels; c++) {
                    const uint32_t *src = (const uint32_t *)&choice[S->sub[c].variant];
                    uint32_t *dst = (uint32_t *)&I.sub[c];
                    // word 7 is `reserved`: the block size, like the generic kernels (internally it carries max |x|)
                    for (int i = tid; i < (int)(sizeof(fb200_subframe_info) / 4); i += T) dst[i] = i == 7 ? (uint32_t)n : src[i];
                }
            }
        FB_WPHASE_END
    FB_WARPS_END
}

// planes and variant mode of a subframe's variant in the pack kernel
FB_DEV int fb_kp_planes(const FbJob &J, const FbKfLayout &L, const int32_t *xs, int variant, int c0, const int32_t **xa,
                        const int32_t **xb) {
    if (L.x16 == 2) { *xa = *xb = xs; return FB_VM_PAIRS | variant; }
    int vm = L.x16 ? FB_VM_X16 : 0;
    *xa = *xb = xs + (size_t)(variant - c0) * L.x_stride;
    if (J.channels == 2 && variant >= 2) { vm |= variant; *xa = xs; *xb = xs + L.x_stride; }
    return vm;
}

// ---- KP: pack a planned frame and store it at out + offsets[f].  pcm: the batch's packed PCM (only read when the
// layout says pairs)
template <int G, bool ODD = false>
FB_DEV void fb_kp_body(const FbJob &J, const int32_t *xv, const uint8_t *pcm, const FbKfPlan *plan,
                       const fb200_subframe_info *psubs, const uint32_t *poffs, const unsigned long long *offsets, uint8_t *out, unsigned long long out_cap,
                       const uint32_t *ktab, uint32_t f, uint8_t *smem, const FbKfLayout &L) {
    const int NW = J.nvar;
    const int T = 32 * NW;
    if (plan[f].state != 0) return; // the generic kernels own this frame
    const int n = fb_frame_len(J, f);
    const FbKfGeom g = fb_kf_geom(n);
    int32_t *xs = (int32_t *)(smem + L.off_x);
    FbKfFrame *S = (FbKfFrame *)(smem + L.off_frame);
    fb200_subframe_info *psub = (fb200_subframe_info *)(smem + L.off_choice);   // [channels]
    uint32_t *poff = (uint32_t *)(smem + L.off_keep);                           // [channels][U_max + 1]
    uint32_t *words = (uint32_t *)(smem + L.off_scratch);
    uint32_t *crc_tab = (uint32_t *)(smem + L.off_crc_tab);
    const uint32_t max_words = (fb_max_frame_bytes(J.channels, J.bps, J.block_size) + 3u) / 4u + 2u;

    // ---- the plan, the CRC tables; clear the word buffer; stage the first group of channels
    const int GC = (int)L.group_ch;
    static_assert(sizeof(FbKfPlan) % 16 == 0 && sizeof(fb200_subframe_info) % 8 == 0, "async copy granularity");
    FB_PHASE(tid, T)
        // the small records fly (cp.async) while the planes are staged through registers
        {
            const int32_t *src = (const int32_t *)&plan[f];
            int32_t *dst = (int32_t *)S;
            for (int i = tid; i < (int)(sizeof(FbKfPlan) / 16); i += T) fb_copy16_async(dst + 4 * i, src + 4 * i);
        }
        {
            const uint64_t *src = (const uint64_t *)&psubs[(size_t)f * (size_t)J.channels];
            uint64_t *dst = (uint64_t *)psub;
            for (int i = tid; i < J.channels * (int)(sizeof(fb200_subframe_info) / 8); i += T) fb_copy8_async(dst + i, src + i);
        }
        {
            const uint32_t *src = poffs + (size_t)f * (size_t)J.channels * (L.U_max + 1);
            for (int i = tid; i < J.channels * (int)(L.U_max + 1); i += T) fb_copy4_async(poff + i, src + i);
        }
        if (L.x16 == 2) fb_kp_stage_pairs(J, pcm, f, n, xs, tid, T);
        else fb_kf_stage<G>(J, xv, f, n, xs, L, tid, T, 0, GC < J.channels ? GC : J.channels);
        for (uint32_t w = (uint32_t)tid; w < max_words; w += (uint32_t)T) words[w] = 0;
        if (tid == 0) { S->crc_acc = 0; S->crc_last = 0; } // (beyond the part of S that the plan copy fills)
        fb_copy_async_wait();
    FB_PHASE_END

    // ---- frame header, subframe heads, and the samples of every unit -- per group of staged channels
    for (int c0 = 0; c0 < J.channels; c0 += GC) {
    const int c1 = c0 + GC < J.channels ? c0 + GC : J.channels;
    if (c0 > 0) {
        FB_PHASE(tid, T)
            fb_kf_stage<G>(J, xv, f, n, xs, L, tid, T, c0, c1);
            fb_copy_async_wait();
        FB_PHASE_END
    }
    FB_PHASE(tid, T)
        // frame header bytes: one lane each, in the last warp (the first warps write the subframe heads)
        if (c0 == 0 && (tid >> 5) == NW - 1 && (tid & 31) < S->header_len)
            fb_or_bits(words, 8u * (uint32_t)(tid & 31), S->header[tid & 31], 8u);
#define FB_KF_X(t) ((uint32_t)fb_kf_load1(xa, xb, vm, mx, (t)))
        // subframe heads (src/component/bitrepr.rs:443-528): every field has a known bit position, so the lanes of
        // warp w write the fields of subframe c0 + w side by side -- type byte, warm-up samples, precision / shift,
        // coefficients, residual method and partition order
        if ((tid >> 5) < c1 - c0) {
            const int sc = c0 + (tid >> 5), lane = tid & 31;
            const FbKfSub &D = S->sub[sc];
            const fb200_subframe_info &V = psub[sc];
            const int32_t *xa, *xb;
            const int vm = fb_kp_planes(J, L, xs, D.variant, c0, &xa, &xb);
            const FbMix mx = fb_mix_of<FB_VMS_ALL>(vm);
            const uint32_t mask = D.bps >= 32 ? 0xFFFFFFFFu : ((1u << D.bps) - 1u);
            const uint32_t base = D.start_bit;
            const bool lpc = D.type == FB200_SF_LPC, fixed = D.type == FB200_SF_FIXED;
            const int warm = (lpc || fixed) ? D.order : (D.type == FB200_SF_CONSTANT ? 1 : 0);
            const int ncoef = lpc ? D.order : 0;
            const int nitems = 1 + warm + (lpc ? 1 : 0) + ncoef + ((lpc || fixed) ? 1 : 0);
            for (int item = lane; item < nitems; item += 32) {
                if (item == 0) {
                    const uint32_t tb = D.type == FB200_SF_CONSTANT ? 0x00u
                                       : D.type == FB200_SF_VERBATIM ? 0x02u
                                       : fixed ? (0x10u | ((uint32_t)D.order << 1)) : (0x40u | ((uint32_t)(D.order - 1) << 1));
                    fb_or_bits(words, base, tb, 8u);
                } else if (item <= warm) {
                    const int t = item - 1;
                    fb_or_bits(words, base + 8u + (uint32_t)t * (uint32_t)D.bps, FB_KF_X(t) & mask, (uint32_t)D.bps);
                } else {
                    const uint32_t after_warm = base + 8u + (uint32_t)warm * (uint32_t)D.bps;
                    int k = item - 1 - warm;
                    if (lpc && k == 0) {
                        fb_or_bits(words, after_warm, ((uint32_t)(D.precision - 1) << 5) | ((uint32_t)D.shift & 31u), 9u);
                    } else {
                        if (lpc) k--;
                        if (k < ncoef) {
                            const uint32_t pmask = (1u << D.precision) - 1u;
                            fb_or_bits(words, after_warm + 9u + (uint32_t)k * (uint32_t)D.precision,
                                       (uint32_t)(int32_t)V.qlp[k] & pmask, (uint32_t)D.precision);
                        } else {
                            fb_or_bits(words, D.res_bit, ((D.rice2 ? 1u : 0u) << 4) | (uint32_t)D.part_order, 6u);
                        }
                    }
                }
            }
        }
        for (int item = tid; item < (c1 - c0) * g.U; item += T) {
            const int c = c0 + (item >> g.lgU), unit = item & (g.U - 1);
            const FbKfSub &D = S->sub[c];
            if (D.type == FB200_SF_CONSTANT) continue;
            int ta, tb;
            fb_kf_unit_range(g, unit, &ta, &tb);
            const int32_t *xa, *xb;
            const int vm = fb_kp_planes(J, L, xs, D.variant, c0, &xa, &xb);
            const FbMix mx = fb_mix_of<FB_VMS_ALL>(vm);
            if (D.type == FB200_SF_VERBATIM) {
                // Verbatim::write (src/component/bitrepr.rs:463-470): bps bits per sample at fixed positions
                if (tb <= ta) continue;
                const uint32_t mask = D.bps >= 32 ? 0xFFFFFFFFu : ((1u << D.bps) - 1u);
                FbBitW r;
                fb_bw_init(r, words, D.start_bit + 8u + (uint32_t)ta * (uint32_t)D.bps,
                           D.start_bit + 8u + (uint32_t)tb * (uint32_t)D.bps);
                for (int t = ta; t < tb; t++) fb_bw_put(r, FB_KF_X(t) & mask, (uint32_t)D.bps);
                fb_bw_finish(r);
                continue;
            }
            // Residual::write (src/component/bitrepr.rs:550-597)
            const fb200_subframe_info &V = psub[c];
            const uint32_t *ub = poff + (size_t)c * (L.U_max + 1);
            const uint32_t p0 = D.code_bit + ub[unit], p1 = D.code_bit + ub[unit + 1];
            if (p1 == p0) continue;
            const int ush = g.lgU - D.part_order;
            const uint32_t rp = V.rice_params[unit >> ush];
            // Every code is OR-ed into the zero-initialised word buffer on its own: the q leading zeros need no
            // write, the (p + 1)-bit tail "1 rrr" lands at bit position pos + q, and pos advances by q + p + 1.
            // Only the running position is sequential; the writes of a run are independent of one another.
            uint32_t pos = p0;
            if ((unit & ((1 << ush) - 1)) == 0) {
                const uint32_t pbits = D.rice2 ? 5u : 4u;
                fb_or_bits(words, pos, rp, pbits);
                pos += pbits;
            }
            const int warm = D.order;
            const int lo = ta > warm ? ta : warm;
            FbKfCand cd;
            cd.kind = D.type == FB200_SF_LPC ? 1 : 0;
            cd.order = D.order; cd.shift = D.shift; cd.q = V.qlp;
            fb_kf_fixed_coefs(cd.kind == 0 ? D.order : 0, cd.fc);
            int32_t qq[G];
            unsigned long long sumabs = 0;
#pragma unroll
            for (int j = 0; j < G; j++) {
                qq[j] = (cd.kind == 1 && j < D.order) ? (int32_t)V.qlp[j] : 0;
                sumabs += (unsigned long long)(qq[j] < 0 ? -qq[j] : qq[j]);
            }
            cd.narrow = cd.kind == 0 ||
                        (unsigned long long)(uint32_t)V.reserved * sumabs < 0x7FFFFFFFull; // reserved: max |x| of the variant
            const uint32_t rmask = (1u << rp) - 1u, rone = 1u << rp;
            if (tb > lo) {
                int32_t win[G + FB_KF_RUN];
                fb_kf_history<G, FB_VMS_ALL, ODD>(xa, xb, vm, mx, ta, win);
                for (int t0 = ta; t0 < tb; t0 += FB_KF_RUN) {
                    uint32_t uu[FB_KF_RUN];
                    fb_kf_fetch_run<G, FB_VMS_ALL, ODD>(xa, xb, vm, mx, t0, win);
                    fb_kf_run_u<G>(win, t0, lo, tb, cd, qq, uu);
#pragma unroll
                    for (int i = 0; i < FB_KF_RUN; i++) {
                        const int t = t0 + i;
                        if (t >= lo && t < tb) {
                            const uint32_t q = uu[i] >> rp;
                            fb_or_bits(words, pos + q, (uu[i] & rmask) | rone, rp + 1u);
                            pos += q + rp + 1u;
                        }
                    }
                    fb_kf_slide<G>(win);
                }
            }
        }
#undef FB_KF_X
    FB_PHASE_END
    } // channel groups

    // ---- the CRC tables replace the planes (every residual is packed: the phase above has ended)
    FB_PHASE(tid, T)
        for (int i = tid; i < 256; i += T) fb_copy16_async((int32_t *)crc_tab + 4 * i, (const int32_t *)ktab + 4 * i);
        fb_copy_async_wait();
    FB_PHASE_END

    // ---- CRC-16 over data_bytes (Frame::write, src/component/bitrepr.rs:289-320).  Chunks of Lc bytes from
    // the start of the frame, one per thread, four bytes per step (slicing tables); the chunk CRCs are shifted
    // to the end of the data with x^(8*Lc*k) (table xp) and one final x^(8*len(last chunk)) (table xb):
    //   crc = (xor_{i<K-1} crc_i * xp[K-2-i]) * xb[len_last]  xor  crc_{K-1}
    const uint32_t B = S->data_bytes;
    uint32_t Lc = 4u * ((B + 4u * (uint32_t)T - 1u) / (4u * (uint32_t)T)); // <= L.crc_chunk: B <= fb_max_frame_bytes
    Lc = Lc ? Lc : 4u;
    const uint32_t K = (B + Lc - 1u) / Lc; // >= 1 chunks, K <= T by the choice of Lc
    FB_PHASE(tid, T)
        if ((uint32_t)tid < K) {
            const uint32_t b0 = (uint32_t)tid * Lc;
            const uint32_t b1 = b0 + Lc < B ? b0 + Lc : B;
            uint32_t crc = 0;
            uint32_t i = b0;
            for (; i + 4u <= b1; i += 4u) { // b0 is a multiple of 4: whole big-endian words
                const uint32_t w = words[i >> 2];
                crc = crc_tab[768 + (((crc >> 8) ^ (w >> 24)) & 0xFFu)] ^ crc_tab[512 + ((crc ^ (w >> 16)) & 0xFFu)] ^
                      crc_tab[256 + ((w >> 8) & 0xFFu)] ^ crc_tab[w & 0xFFu];
            }
            for (; i < b1; i++) {
                const uint32_t byte = (words[i >> 2] >> (24u - 8u * (i & 3u))) & 0xFFu;
                crc = ((crc << 8) & 0xFFFFu) ^ crc_tab[((crc >> 8) ^ byte) & 0xFFu];
            }
            if ((uint32_t)tid + 1u == K) {
                S->crc_last = crc;
            } else {
                const uint32_t sh = fb_crc16_mulmod(crc, ktab[fb_kf_ktab_xp(L.crc_chunk) + (Lc / 4u - 1u) * (uint32_t)T + (K - 2u - (uint32_t)tid)]);
#if FB_GPU
                if (sh) atomicXor(&S->crc_acc, sh);
#else
                S->crc_acc ^= sh;
#endif
            }
        }
    FB_PHASE_END
    FB_PHASE(tid, T)
        if (tid == 0) {
            const uint32_t len_last = B - (K - 1u) * Lc;
            const uint32_t crc = fb_crc16_mulmod(S->crc_acc, ktab[FB_KTAB_XB + len_last]) ^ S->crc_last;
            // the two CRC bytes follow the (byte-aligned) data, big-endian
            for (int i = 0; i < 2; i++) {
                const uint32_t pos = B + (uint32_t)i;
                const uint32_t byte = (crc >> (8 * (1 - i))) & 0xFFu;
                fb_atomic_or(&words[pos >> 2], byte << (24u - 8u * (pos & 3u)));
            }
        }
    FB_PHASE_END
    // ---- store: the frame's bytes (big-endian words in shared memory) at out + offsets[f].  Head bytes up to the first
    // 4-byte aligned destination address, whole words assembled from two shared-memory words, tail bytes.
    FB_PHASE(tid, T)
        const uint32_t len = B + 2u;
        const unsigned long long o = offsets[f];
        if (o + len <= out_cap) { // a capacity error is reported by the host from the total
            uint8_t *dst = out + o;
            uint32_t head = (uint32_t)((4u - (uint32_t)((uintptr_t)dst & 3u)) & 3u);
            if (head > len) head = len;
            const uint32_t nw = (len - head) >> 2;
            if ((uint32_t)tid < head) dst[tid] = (uint8_t)(words[0] >> (24u - 8u * (uint32_t)tid));
            uint32_t *dstw = (uint32_t *)(dst + head);
            const uint32_t sh = head * 8u; // head <= 3: stream word k starts at byte head + 4k
            for (uint32_t k = (uint32_t)tid; k < nw; k += (uint32_t)T) {
                uint32_t v = words[k];
                if (sh) v = (v << sh) | (words[k + 1] >> (32u - sh));
                dstw[k] = ((v & 0xFFu) << 24) | ((v & 0xFF00u) << 8) | ((v >> 8) & 0xFF00u) | (v >> 24);
            }
            const uint32_t done = head + 4u * nw;
            if ((uint32_t)tid < len - done) {
                const uint32_t p = done + (uint32_t)tid;
                dst[p] = (uint8_t)(words[p >> 2] >> (24u - 8u * (p & 3u)));
            }
        }
    FB_PHASE_END
}
