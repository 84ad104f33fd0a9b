// fb_fused.cuh -- KF: the fused per-frame kernel (Rice search of every channel variant + frame assembly).
//
// One CTA per frame, one warp per channel variant (L, R, M, S for stereo).  The frame's independent
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
//     pass 2  per finest partition: the exact minimiser, found by walking the convex cost function
//     pass 3  bottom-up partition tree (src/rice.rs:246-298) over the parameter window
//             [min leaf minimiser, max leaf minimiser] -- the minimiser of every merged node lies in it
//             (sum of convex functions) -- in passes of 8 parameters
//     pass 4  partition order, parameters, exact Residual::count_bits (src/component/bitrepr.rs:532-544)
//             and the bit length of every unit (again from the counters)
//   subframe decision (src/coding.rs:384-418); then per frame: stereo decision (src/coding.rs:454-527),
//   header + CRC-8, bit offsets of all units by a scan, packing (one thread per unit; residuals are
//   recomputed from the staged samples), CRC-16, store.
//
// Whatever this kernel cannot reproduce exactly -- a residual >= 2^27 (the reference's 16-sample chunked
// saturating accumulation matters, src/rice.rs:75-98) or a saturated table minimum -- is not guessed: the
// frame is appended to a fallback list and redone by the generic K2/K3 kernels (fb_kernels.cuh), which
// replay the reference literally.  Results are byte-identical either way.
#pragma once

#include "fb_kernels.cuh"

#define FB_KF_COLS 8       // Rice parameters evaluated per tree pass
#define FB_KF_ROW 9        // row stride of the tree tables (one pad word: conflict-free column reads)
#define FB_KF_NWORDS 7     // bit-sliced counter words per unit (counts <= 127)
#define FB_KF_UNIT_MAX 112 // samples per unit (7 runs of 16)

// ---- warp-scope phase macros (see fb_common.h for the CTA-scope ones) --------------------------
// FB_WARPS_BEGIN(w, NW) ... FB_WARPS_END : every warp of the CTA runs the enclosed code independently
// (the emulation runs the warps one after the other); inside, FB_WPHASE(lane) ... FB_WPHASE_END is a
// region between two warp barriers.  Values that steer warp-uniform control flow are read from shared
// memory between phases, under the same rule as the CTA-scope macros.
#if FB_GPU
#define FB_WARPS_BEGIN(w, NW) { const int w = (int)(threadIdx.x >> 5); (void)w;
#define FB_WARPS_END } __syncthreads();
#define FB_WSYNC() __syncwarp()
#else
#define FB_WARPS_BEGIN(w, NW) for (int w = 0; w < (NW); ++w) {
#define FB_WARPS_END }
#define FB_WSYNC() ((void)0)
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
    int m;        // units per leaf (power of two)
    int U;        // units per variant = leaves * m, a power of two >= 32
    int lgU;
    int unit_len; // samples per unit (multiple of 4, <= FB_KF_UNIT_MAX); the last unit of a leaf may be shorter
};

FB_HD FbKfGeom fb_kf_geom(int n) {
    FbKfGeom g;
    g.n = n;
    g.o0 = fb_finest_partition_order(n);
    g.leaves = 1 << g.o0;
    g.leaf_len = n >> g.o0;
    int m = 1;
    while (g.leaves * m < 32) m <<= 1;
    while (((((g.leaf_len + m - 1) / m) + 3) & ~3) > FB_KF_UNIT_MAX) m <<= 1;
    g.m = m;
    g.unit_len = (((g.leaf_len + m - 1) / m) + 3) & ~3;
    g.U = g.leaves * m;
    g.lgU = 0;
    while ((1 << g.lgU) < g.U) g.lgU++;
    return g;
}

// sample range [t0, t1) of unit u
FB_HD void fb_kf_unit_range(const FbKfGeom &g, int u, int *t0, int *t1) {
    const int leaf = u / g.m, k = u - leaf * g.m;
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

// ---- shared-memory layout (bytes), identical on host and device ---------------------------------
struct FbKfLayout {
    uint32_t x_stride;   // words per staged plane
    uint32_t off_x;      // channels planes
    uint32_t off_keep;   // per warp: unit_bits[2][U+1], xch[64], results
    uint32_t keep_bytes;
    uint32_t k_unit_bits, k_xch, k_res;
    uint32_t off_choice; // nvar x fb200_subframe_info
    uint32_t off_frame;  // FbKfFrame
    uint32_t off_scratch; // per warp scratch, aliased by the frame words during packing
    uint32_t scratch_bytes;
    uint32_t s_words, s_tbl_a, s_tbl_b, s_best_val, s_best_p, s_lvl_bits, s_misc;
    uint32_t words_bytes; // bytes of the frame word buffer
    uint32_t U_max, leaves_max;
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
    unsigned long long sum_bits;
};

struct FbKfSub {
    int32_t  variant, cand;  // cand: 0 fixed, 1 lpc (index of the unit_bits / result set)
    int32_t  type, order, bps, precision, shift, part_order, rice2;
    uint32_t start_bit, res_bit, code_bit;
};

struct FbKfFrame {
    FbKfSub sub[FB200_MAX_CHANNELS];
    uint8_t header[16];
    int32_t header_len, ch_tag;
    uint32_t data_bytes;
    uint32_t frame_fail;
    int32_t  cand[FB200_MAX_CHANNELS]; // per variant: result set of the chosen coding (0 fixed, 1 lpc)
    uint32_t crc_xpow[9];
    uint32_t crc_tab[256];
    uint32_t crc_part[256];
};

FB_HD uint32_t fb_align16(uint32_t v) { return (v + 15u) & ~15u; }

FB_HD FbKfLayout fb_kf_layout(int channels, int nvar, int bps, int block_size, int tail_n) {
    FbKfLayout L;
    const FbKfGeom ga = fb_kf_geom(block_size), gb = fb_kf_geom(tail_n);
    const uint32_t U = (uint32_t)(ga.U > gb.U ? ga.U : gb.U);
    const uint32_t leaves = (uint32_t)(ga.leaves > gb.leaves ? ga.leaves : gb.leaves);
    L.U_max = U;
    L.leaves_max = leaves;
    L.x_stride = (uint32_t)((fb_xidx(block_size + 32) + 8 + 3) & ~3);
    uint32_t o = 0;
    L.off_x = o;        o += fb_align16((uint32_t)channels * L.x_stride * 4u);
    // kept per warp
    uint32_t k = 0;
    L.k_unit_bits = k;  k += fb_align16(2u * (U + 1u) * 4u);
    L.k_xch = k;        k += 64u * 8u;
    L.k_res = k;        k += fb_align16(2u * (uint32_t)sizeof(FbKfRes));
    L.keep_bytes = k;
    L.off_keep = o;     o += (uint32_t)nvar * k;
    L.off_choice = o;   o += fb_align16((uint32_t)nvar * (uint32_t)sizeof(fb200_subframe_info));
    L.off_frame = o;    o += fb_align16((uint32_t)sizeof(FbKfFrame));
    // scratch per warp
    uint32_t s = 0;
    L.s_words = s;      s += fb_align16(FB_KF_NWORDS * U * 4u);
    L.s_tbl_a = s;      s += fb_align16(leaves * FB_KF_ROW * 4u + 64u * 4u); // also level-sum exchange (16 x 32)
    if (leaves * FB_KF_ROW < 16u * 32u) s = L.s_tbl_a + fb_align16(16u * 32u * 4u);
    L.s_tbl_b = s;      s += fb_align16((leaves / 2u + 1u) * FB_KF_ROW * 4u);
    L.s_best_val = s;   s += fb_align16(2u * leaves * 4u);
    L.s_best_p = s;     s += fb_align16(2u * leaves);
    L.s_lvl_bits = s;   s += 16u * 8u;
    L.s_misc = s;       s += fb_align16((uint32_t)sizeof(FbKfMisc));
    L.scratch_bytes = s;
    L.words_bytes = fb_align16(((fb_max_frame_bytes(channels, bps, block_size) + 3u) & ~3u) + 16u);
    const uint32_t scratch_total = (uint32_t)nvar * s;
    L.off_scratch = o;  o += scratch_total > L.words_bytes ? scratch_total : L.words_bytes;
    L.total = o;
    return L;
}

// ---- bit-sliced counters ----------------------------------------------------------------------------
// full adder on 32 independent bit planes: (h, l) = a + b + c
#define FB_CSA(h, l, a, b, c) do { const uint32_t a__ = (a), b__ = (b), c__ = (c); const uint32_t x__ = a__ ^ b__; \
                                   (h) = (a__ & b__) | (x__ & c__); (l) = x__ ^ c__; } while (0)

// adds 16 values into the counter words cw[0..6] (weights 1, 2, 4, 8, 16, 32, 64)
FB_DEV void fb_kf_csa16(uint32_t *cw, const uint32_t *d) {
    uint32_t twoA, twoB, fourA, fourB, eightA, eightB, sixteen;
    uint32_t ones = cw[0], twos = cw[1], fours = cw[2], eights = cw[3];
    FB_CSA(twoA, ones, ones, d[0], d[1]);
    FB_CSA(twoB, ones, ones, d[2], d[3]);
    FB_CSA(fourA, twos, twos, twoA, twoB);
    FB_CSA(twoA, ones, ones, d[4], d[5]);
    FB_CSA(twoB, ones, ones, d[6], d[7]);
    FB_CSA(fourB, twos, twos, twoA, twoB);
    FB_CSA(eightA, fours, fours, fourA, fourB);
    FB_CSA(twoA, ones, ones, d[8], d[9]);
    FB_CSA(twoB, ones, ones, d[10], d[11]);
    FB_CSA(fourA, twos, twos, twoA, twoB);
    FB_CSA(twoA, ones, ones, d[12], d[13]);
    FB_CSA(twoB, ones, ones, d[14], d[15]);
    FB_CSA(fourB, twos, twos, twoA, twoB);
    FB_CSA(eightB, fours, fours, fourA, fourB);
    FB_CSA(sixteen, eights, eights, eightA, eightB);
    cw[0] = ones; cw[1] = twos; cw[2] = fours; cw[3] = eights;
    // ripple the carry into the 16/32/64 words (half adders); counts stay <= 127 by construction
    uint32_t c = cw[4] & sixteen; cw[4] ^= sixteen;
    uint32_t c2 = cw[5] & c;      cw[5] ^= c;
    cw[6] ^= c2;
}

// sum_t (u_t >> p) of one unit from its counter words (stored [word][unit], stride U)
FB_DEV unsigned long long fb_kf_eval(const uint32_t *words, int U, int unit, int p) {
    unsigned long long s = 0;
#pragma unroll
    for (int j = 0; j < FB_KF_NWORDS; j++) s += (unsigned long long)(words[j * U + unit] >> p) << j;
    return s;
}

// M = (L + R) >> 1 (arithmetic), S = L - R (src/coding.rs:476-484); unsigned adds so that stale padding
// words (whose results are masked) cannot overflow a signed int
FB_HD int32_t fb_mid(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b) >> 1; }
FB_HD int32_t fb_side(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }

// ---- residual runs out of the staged planes ------------------------------------------------------
// vm: 0 = plane xa as is, 2 = mid (xa + xb) >> 1, 3 = side xa - xb   (src/coding.rs:476-484)
// win[i] = x[t0 - G + i], i < G + FB_RUN; samples outside [0, n) read as 0 (their results are masked anyway)
template <int G>
FB_DEV void fb_kf_window(const int32_t *xa, const int32_t *xb, int vm, int t0, int n, int32_t *win) {
    if (((t0 & 3) == 0) && t0 >= G) {
#pragma unroll
        for (int i = 0; i < G + FB_RUN; i += 4) {
            const int o = fb_xidx(t0 - G + i);
            const int4 a = *reinterpret_cast<const int4 *>(xa + o);
            if (vm < 2) {
                win[i] = a.x; win[i + 1] = a.y; win[i + 2] = a.z; win[i + 3] = a.w;
            } else {
                const int4 b = *reinterpret_cast<const int4 *>(xb + o);
                if (vm == 2) {
                    win[i] = fb_mid(a.x, b.x); win[i + 1] = fb_mid(a.y, b.y);
                    win[i + 2] = fb_mid(a.z, b.z); win[i + 3] = fb_mid(a.w, b.w);
                } else {
                    win[i] = fb_side(a.x, b.x); win[i + 1] = fb_side(a.y, b.y);
                    win[i + 2] = fb_side(a.z, b.z); win[i + 3] = fb_side(a.w, b.w);
                }
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < G + FB_RUN; i++) {
            const int t = t0 - G + i;
            int32_t v = 0;
            if (t >= 0 && t < n) {
                const int o = fb_xidx(t);
                v = xa[o];
                if (vm == 2) v = fb_mid(v, xb[o]);
                else if (vm == 3) v = fb_side(v, xb[o]);
            }
            win[i] = v;
        }
    }
}

// zigzag residuals of the run t0..t0+15 from its window; samples outside [lo, hi) give 0.
// kind 0: fixed predictor of `order` (zero-history differences, wrapping i32);
// kind 1: LPC, narrow = the reference's i32 accumulation is safe (src/lpc.rs:361-374), else 64-bit.
template <int G>
FB_DEV void fb_kf_run_u(const int32_t *win, int t0, int lo, int hi, int kind, int order, const int32_t *qq, int shift,
                        bool narrow, uint32_t *u) {
    if (kind == 0) {
#pragma unroll
        for (int i = 0; i < FB_RUN; i++) {
            const uint32_t a = (uint32_t)win[G + i], b = (uint32_t)win[G + i - 1], c = (uint32_t)win[G + i - 2],
                           d = (uint32_t)win[G + i - 3], e4 = (uint32_t)win[G + i - 4];
            uint32_t e;
            switch (order) {
            case 0: e = a; break;
            case 1: e = a - b; break;
            case 2: e = a - 2u * b + c; break;
            case 3: e = a - 3u * b + 3u * c - d; break;
            default: e = a - 4u * b + 6u * c - 4u * d + e4; break;
            }
            const int t = t0 + i;
            u[i] = (t >= lo && t < hi) ? fb_zigzag((int32_t)e) : 0u;
        }
    } else if (narrow) {
#pragma unroll
        for (int i = 0; i < FB_RUN; i++) {
            uint32_t acc = 0;
#pragma unroll
            for (int j = 0; j < G; j++) acc += (uint32_t)qq[j] * (uint32_t)win[G + i - 1 - j];
            const int32_t e = (int32_t)((uint32_t)win[G + i] - (uint32_t)((int32_t)acc >> shift));
            const int t = t0 + i;
            u[i] = (t >= lo && t < hi) ? fb_zigzag(e) : 0u;
        }
    } else {
#pragma unroll
        for (int i = 0; i < FB_RUN; i++) {
            int64_t acc = 0;
#pragma unroll
            for (int j = 0; j < G; j++) acc = fb_mad_wide(qq[j], win[G + i - 1 - j], acc);
            const int32_t e = (int32_t)(uint32_t)((uint64_t)(int64_t)win[G + i] - (uint64_t)(acc >> shift));
            const int t = t0 + i;
            u[i] = (t >= lo && t < hi) ? fb_zigzag(e) : 0u;
        }
    }
}

// description of one residual candidate of a variant
struct FbKfCand {
    int kind, order, shift;
    bool narrow;
    const int16_t *q;
};

// =====================================================================================================
// Rice search of one candidate by one warp.  xa/xb/vm select the variant's samples.  Writes res and
// unit_bits[0..U] (bits of every unit without the parameter fields; [U] unused here).
// Sets M->fail when the frame must be redone by the literal path.
// =====================================================================================================
template <int G>
FB_DEV void fb_kf_search(const FbJob &J, const FbKfGeom &g, const int32_t *xa, const int32_t *xb, int vm,
                         const FbKfCand &cd, uint8_t *scratch, const FbKfLayout &L, uint32_t *unit_bits,
                         uint32_t *xch, FbKfRes *res) {
    uint32_t *words = (uint32_t *)(scratch + L.s_words);
    uint32_t *tbl_a = (uint32_t *)(scratch + L.s_tbl_a);
    uint32_t *tbl_b = (uint32_t *)(scratch + L.s_tbl_b);
    uint32_t *best_val = (uint32_t *)(scratch + L.s_best_val);
    uint8_t *best_p = scratch + L.s_best_p;
    unsigned long long *lvl_bits = (unsigned long long *)(scratch + L.s_lvl_bits);
    FbKfMisc *M = (FbKfMisc *)(scratch + L.s_misc);
    const int n = g.n, U = g.U, warm = cd.order, max_p = J.cfg.prc_max_parameter;
    const int slots = U >> 5;

    FB_WPHASE(lane)
        if (lane == 0) { M->ormask = 0; M->pmin = 31; M->pmax = 0; M->any_gt14 = 0; M->sum_bits = 0; }
    FB_WPHASE_END

    // ---- pass 1: residuals -> bit-sliced counters per unit
    FB_WPHASE(lane)
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
            for (int t0 = ta; t0 < tb; t0 += FB_RUN) {
                int32_t win[G + FB_RUN];
                uint32_t uu[FB_RUN];
                fb_kf_window<G>(xa, xb, vm, t0, n, win);
                fb_kf_run_u<G>(win, t0, lo, tb, cd.kind, cd.order, qq, cd.shift, cd.narrow, uu);
                fb_kf_csa16(cw, uu);
            }
#pragma unroll
            for (int j = 0; j < FB_KF_NWORDS; j++) { words[j * U + unit] = cw[j]; orm |= cw[j]; }
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

    // ---- pass 2: exact minimiser of every finest partition (walk on the convex cost)
    FB_WPHASE(lane)
        for (int leaf = lane; leaf < g.leaves; leaf += 32) {
            const int cnt = g.leaf_len - (leaf == 0 ? warm : 0);
            unsigned long long s0 = 0;
            for (int k = 0; k < g.m; k++) s0 += fb_kf_eval(words, U, leaf * g.m + k, 0);
            // starting point ~ log2(mean); any start gives the same result
            int p = 0;
            if (s0 > (unsigned long long)cnt && cnt > 0) p = (63 - fb_clz64(s0)) - (31 - fb_clz32((uint32_t)cnt));
            if (p < 0) p = 0;
            if (p > max_p) p = max_p;
            unsigned long long fc = 0;
            for (int k = 0; k < g.m; k++) fc += fb_kf_eval(words, U, leaf * g.m + k, p);
            fc += (unsigned long long)cnt * (unsigned long long)(p + 1);
            bool moved = false;
            while (p < max_p) {
                unsigned long long fu = 0;
                for (int k = 0; k < g.m; k++) fu += fb_kf_eval(words, U, leaf * g.m + k, p + 1);
                fu += (unsigned long long)cnt * (unsigned long long)(p + 2);
                if (fu < fc) { fc = fu; p++; moved = true; } else break;
            }
            if (!moved) {
                while (p > 0) {
                    unsigned long long fd = 0;
                    for (int k = 0; k < g.m; k++) fd += fb_kf_eval(words, U, leaf * g.m + k, p - 1);
                    fd += (unsigned long long)cnt * (unsigned long long)p;
                    if (fd <= fc) { fc = fd; p--; } else break;
                }
            }
            if (fc + 4ull >= (unsigned long long)FB_RICE_SAT) M->fail = 1; // benign race: every writer stores 1
            fb_atomic_min_u32(&M->pmin, (uint32_t)p);
            fb_atomic_max_u32(&M->pmax, (uint32_t)p);
        }
    FB_WPHASE_END
    if (M->fail) return;

    // ---- pass 3: partition tree over [pmin, pmax], FB_KF_COLS parameters at a time
    const int pa0 = (int)M->pmin, pb = (int)M->pmax;
    for (int pa = pa0; pa <= pb; pa += FB_KF_COLS) {
        const int Wc = (pb - pa + 1) < FB_KF_COLS ? (pb - pa + 1) : FB_KF_COLS;
        const bool first = pa == pa0;
        // leaf tables: min(S + cnt*(p+1) + 4, 2^27-1)   (src/rice.rs:65-105)
        FB_WPHASE(lane)
            for (int e = lane; e < g.leaves * FB_KF_COLS; e += 32) {
                const int leaf = e >> 3, j = e & 7;
                if (j < Wc) {
                    const int p = pa + j;
                    const int cnt = g.leaf_len - (leaf == 0 ? warm : 0);
                    unsigned long long f = 4ull + (unsigned long long)cnt * (unsigned long long)(p + 1);
                    for (int k = 0; k < g.m; k++) f += fb_kf_eval(words, U, leaf * g.m + k, p);
                    tbl_a[leaf * FB_KF_ROW + j] = f > FB_RICE_SAT ? FB_RICE_SAT : (uint32_t)f;
                }
            }
        FB_WPHASE_END
        uint32_t *cur = tbl_a, *nxt = tbl_b;
        for (int lvl = g.o0; lvl >= 0; lvl--) {
            const int nodes = 1 << lvl;
            FB_WPHASE(lane)
                // minimiser: smallest (bits, p)  (src/rice.rs:117-141); windows are visited in ascending p
                for (int node = lane; node < nodes; node += 32) {
                    const uint32_t *row = cur + node * FB_KF_ROW;
                    uint32_t bv = row[0];
                    int bp = pa;
                    for (int j = 1; j < Wc; j++)
                        if (row[j] < bv) { bv = row[j]; bp = pa + j; }
                    const int idx = nodes - 1 + node;
                    if (first || bv < best_val[idx]) { best_val[idx] = bv; best_p[idx] = (uint8_t)bp; }
                }
                // merge pairs: min(a + b - 4, 2^27-1)  (src/rice.rs:144-152)
                if (lvl > 0) {
                    for (int e = lane; e < (nodes >> 1) * FB_KF_COLS; e += 32) {
                        const int node = e >> 3, j = e & 7;
                        if (j < Wc) {
                            const uint32_t v = cur[(2 * node) * FB_KF_ROW + j] + cur[(2 * node + 1) * FB_KF_ROW + j] - 4u;
                            nxt[node * FB_KF_ROW + j] = v < FB_RICE_SAT ? v : FB_RICE_SAT;
                        }
                    }
                }
            FB_WPHASE_END
            uint32_t *tmp = cur; cur = nxt; nxt = tmp;
        }
    }

    // ---- pass 4a: totals per level (lane partials -> exchange -> one lane per level)
    uint32_t *lx = tbl_a; // 16 x 32 exchange words
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
    // ---- pass 4c: parameters, bits of every unit, Residual::count_bits
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
            unsigned long long local = 0;
            for (int unit = lane; unit < U; unit += 32) {
                const int p = best_p[nparts - 1 + (unit >> ush)];
                int ta, tb;
                fb_kf_unit_range(g, unit, &ta, &tb);
                if (ta < warm) ta = warm;
                const int cnt = tb > ta ? tb - ta : 0;
                const unsigned long long b = fb_kf_eval(words, U, unit, p) + (unsigned long long)cnt * (unsigned long long)(p + 1);
                unit_bits[unit] = (uint32_t)b;
                local += b;
            }
            xch[2 * lane] = (uint32_t)local;
            xch[2 * lane + 1] = (uint32_t)(local >> 32);
        FB_WPHASE_END
        FB_WPHASE(lane)
            if (lane == 0) {
                unsigned long long total = 0;
                for (int i = 0; i < 32; i++) total += (unsigned long long)xch[2 * i] | ((unsigned long long)xch[2 * i + 1] << 32);
                const int rice2 = M->any_gt14 ? 1 : 0;
                res->rice2 = rice2;
                // src/component/bitrepr.rs:532-544: 2 + 4 + parts*(4|5) + sum q + (n - w) + sum p*len - w*p0
                res->res_bits = 6ull + (unsigned long long)nparts * (rice2 ? 5ull : 4ull) + total;
            }
        FB_WPHASE_END
    }
}

// =====================================================================================================
// One variant (one warp): fixed_lpc / estimated_qlpc / encode_subframe (src/coding.rs:298-418) on top
// of K1's analysis.  Writes the decision record `out` (shared memory) and *cand_out (which unit_bits set).
// =====================================================================================================
template <int G>
FB_DEV void fb_kf_variant(const FbJob &J, const FbKfGeom &g, const int32_t *xs, const FbAnalysis &A, int v,
                          uint8_t *smem, const FbKfLayout &L, fb200_subframe_info *out) {
    FbKfFrame *S = (FbKfFrame *)(smem + L.off_frame);
    uint8_t *scratch = smem + L.off_scratch + (uint32_t)v * L.scratch_bytes;
    uint8_t *keep = smem + L.off_keep + (uint32_t)v * L.keep_bytes;
    uint32_t *unit_bits = (uint32_t *)(keep + L.k_unit_bits);
    uint32_t *xch = (uint32_t *)(keep + L.k_xch);
    FbKfRes *res = (FbKfRes *)(keep + L.k_res);
    FbKfMisc *M = (FbKfMisc *)(scratch + L.s_misc);
    const int n = g.n;
    const int bps_v = fb_variant_bps(J, v);
    const unsigned long long verbatim_bits = 8ull + (unsigned long long)n * (unsigned long long)bps_v;
    // sample planes of this variant
    int vm = 0;
    const int32_t *xa = xs + (size_t)v * L.x_stride, *xb = xa;
    if (J.channels == 2 && v >= 2) { vm = v; xa = xs; xb = xs + L.x_stride; }

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
            S->cand[v] = 0;
        }
        out->qlp[lane] = 0;
    FB_WPHASE_END

    if (J.cfg.use_constant && A.is_constant) {
        FB_WPHASE(lane)
            if (lane == 0) { out->type = FB200_SF_CONSTANT; out->bits = 8ull + (unsigned long long)bps_v; }
        FB_WPHASE_END
        return;
    }
    if (n < FB_MIN_PRED_BLOCK) return;

    // fixed candidate (ApproxEnt winner from K1; src/coding.rs:298-331)
    const int kf = J.cfg.use_fixed ? A.fixed_order : -1;
    unsigned long long fixed_bits = 0;
    if (kf >= 0) {
        FbKfCand cd;
        cd.kind = 0; cd.order = kf; cd.shift = 0; cd.narrow = true; cd.q = nullptr;
        fb_kf_search<G>(J, g, xa, xb, vm, cd, scratch, L, unit_bits, xch, &res[0]);
        if (M->fail) return;
        fixed_bits = 8ull + (unsigned long long)bps_v * (unsigned long long)kf + res[0].res_bits;
    }
    const unsigned long long baseline_bits =
        kf >= 0 ? (fixed_bits < verbatim_bits ? fixed_bits : verbatim_bits) : verbatim_bits;

    // LPC candidate (src/coding.rs:360-381)
    bool lpc_ok = false;
    unsigned long long lpc_bits = 0;
    if (J.cfg.use_lpc) {
        FbKfCand cd;
        cd.kind = 1; cd.order = A.qlp_order; cd.shift = A.qlp_shift; cd.q = A.qlp;
        // src/lpc.rs:361-374: i32 accumulation when max|x| * sum|q| < 2^31 - 1
        unsigned long long sumabs = 0;
        for (int j = 0; j < A.qlp_order; j++) sumabs += (unsigned long long)(A.qlp[j] < 0 ? -A.qlp[j] : A.qlp[j]);
        cd.narrow = (unsigned long long)A.max_abs * sumabs < 0x7FFFFFFFull;
        fb_kf_search<G>(J, g, xa, xb, vm, cd, scratch, L, unit_bits + (L.U_max + 1), xch, &res[1]);
        if (M->fail) return;
        lpc_bits = 8ull + (unsigned long long)bps_v * (unsigned long long)A.qlp_order + 4ull + 5ull +
                   (unsigned long long)J.cfg.quant_precision * (unsigned long long)A.qlp_order + res[1].res_bits;
        lpc_ok = lpc_bits < baseline_bits;
    }

    // decision (src/coding.rs:403-416)
    int pick = -1;
    if (lpc_ok) pick = 1;
    else if (kf >= 0) pick = 0;
    if (pick == 1 && !(lpc_bits < verbatim_bits)) pick = -1;
    if (pick == 0 && !(fixed_bits < verbatim_bits)) pick = -1;
    if (pick < 0) return;

    const FbKfRes *R = &res[pick];
    FB_WPHASE(lane)
        if (lane == 0) {
            out->type = pick == 1 ? FB200_SF_LPC : FB200_SF_FIXED;
            out->order = pick == 1 ? A.qlp_order : kf;
            out->precision = pick == 1 ? J.cfg.quant_precision : 0;
            out->shift = pick == 1 ? A.qlp_shift : 0;
            out->partition_order = R->part_order;
            out->rice2 = R->rice2;
            out->bits = pick == 1 ? lpc_bits : fixed_bits;
            S->cand[v] = pick;
        }
        if (pick == 1) out->qlp[lane] = A.qlp[lane];
        for (int i = lane; i < (1 << R->part_order); i += 32) out->rice_params[i] = R->params[i];
    FB_WPHASE_END
}

// =====================================================================================================
// KF body: one CTA (32 * nvar threads) per frame.
// =====================================================================================================
template <int G>
FB_DEV void fb_kf_body(const FbJob &J, const int32_t *xv, const FbAnalysis *ana, uint8_t *slots, uint32_t *frame_bytes,
                       fb200_frame_info *infos, uint32_t *fb_list, uint32_t *fb_count, uint32_t f, uint8_t *smem,
                       const FbKfLayout &L) {
    const int NW = J.nvar;
    const int T = 32 * NW;
    const int n = fb_frame_len(J, f);
    const FbKfGeom g = fb_kf_geom(n);
    int32_t *xs = (int32_t *)(smem + L.off_x);
    fb200_subframe_info *choice = (fb200_subframe_info *)(smem + L.off_choice);
    FbKfFrame *S = (FbKfFrame *)(smem + L.off_frame);
    uint32_t *words = (uint32_t *)(smem + L.off_scratch);
    uint8_t *slot = slots + (size_t)f * (size_t)J.slot_bytes;

    // ---- stage the independent channels (coalesced 16-byte loads; rows are padded to a multiple of 32)
    FB_PHASE(tid, T)
        const int n4 = (n + 3) >> 2;
        for (int c = 0; c < J.channels; c++) {
            const int32_t *src = xv + ((size_t)f * (size_t)J.nvar + (size_t)c) * (size_t)J.stride;
            int32_t *dst = xs + (size_t)c * L.x_stride;
            for (int i = tid; i < n4; i += T) {
                const int4 v = *reinterpret_cast<const int4 *>(src + 4 * i);
                *reinterpret_cast<int4 *>(dst + fb_xidx(4 * i)) = v;
            }
        }
        if (tid == 0) S->frame_fail = 0;
    FB_PHASE_END

    // ---- analysis: one warp per variant
    FB_WARPS_BEGIN(w, NW)
        fb_kf_variant<G>(J, g, xs, ana[(size_t)f * (size_t)J.nvar + (size_t)w], w, smem, L, &choice[w]);
        const FbKfMisc *M = (const FbKfMisc *)(smem + L.off_scratch + (uint32_t)w * L.scratch_bytes + L.s_misc);
        FB_WPHASE(lane)
            if (lane == 0 && M->fail) S->frame_fail = 1; // benign race between warps
        FB_WPHASE_END
    FB_WARPS_END

    if (S->frame_fail) {
        // not reproducible here: hand the frame to the literal kernels
        FB_PHASE(tid, T)
            if (tid == 0) {
#if FB_GPU
                const uint32_t slot_i = atomicAdd(fb_count, 1u);
#else
                const uint32_t slot_i = (*fb_count)++;
#endif
                fb_list[slot_i] = f;
            }
        FB_PHASE_END
        return;
    }

    // ---- stereo decision, header, subframe offsets (thread 0); CRC table; clear the word buffer
    const uint32_t max_words = (fb_max_frame_bytes(J.channels, J.bps, J.block_size) + 3u) / 4u + 2u;
    FB_PHASE(tid, T)
        for (int i = tid; i < 256; i += T) S->crc_tab[i] = fb_crc16_table_entry((uint32_t)i);
        if (tid == 0) {
            int ch_tag = J.channels - 1;
            int sel[FB200_MAX_CHANNELS];
            for (int c = 0; c < J.channels; c++) sel[c] = c;
            if (J.channels == 2) {
                // try_stereo_coding (src/coding.rs:469-527): strict <, order I, L/S, R/S, M/S
                const unsigned long long bl = choice[0].bits, br = choice[1].bits, bm = choice[2].bits, bs = choice[3].bits;
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
                const fb200_subframe_info &V = choice[sel[c]];
                FbKfSub &D = S->sub[c];
                D.variant = sel[c];
                D.cand = S->cand[sel[c]];
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
    // the scratch of the analysis is dead from here on; the frame words alias it
    FB_PHASE(tid, T)
        for (uint32_t w = (uint32_t)tid; w < max_words; w += (uint32_t)T) words[w] = 0;
    FB_PHASE_END

    // ---- bit offsets of the units of every coded subframe: exclusive scan incl. the parameter fields
    // (warp c scans subframe c; channels <= nvar)
    FB_WARPS_BEGIN(w, NW)
        if (w < J.channels && (S->sub[w].type == FB200_SF_FIXED || S->sub[w].type == FB200_SF_LPC)) {
            const FbKfSub &D = S->sub[w];
            uint8_t *keep = smem + L.off_keep + (uint32_t)D.variant * L.keep_bytes;
            uint32_t *ub = (uint32_t *)(keep + L.k_unit_bits) + (size_t)D.cand * (L.U_max + 1);
            uint32_t *xch = (uint32_t *)(smem + L.off_keep + (uint32_t)w * L.keep_bytes + L.k_xch);
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
                    const uint32_t b = ub[unit] + (((unit & ((1 << ush) - 1)) == 0) ? pbits : 0u);
                    ub[unit] = s;
                    s += b;
                }
                if (lane == 31) ub[g.U] = s;
            FB_WPHASE_END
        }
    FB_WARPS_END

    // ---- frame header, subframe heads, and the samples of every unit
    FB_PHASE(tid, T)
        if (tid == 0) {
            FbBitRun r;
            fb_run_init(r, words, 0, 1);
            r.w_last = 0xFFFFFFFFu;
            for (int i = 0; i < S->header_len; i++) {
                r.w_first = r.cur_w; // every flush through the atomic path
                fb_run_put(r, S->header[i], 8);
            }
            r.w_first = r.cur_w;
            fb_run_flush(r);
        }
        if (tid < J.channels) {
            const FbKfSub &D = S->sub[tid];
            const fb200_subframe_info &V = choice[D.variant];
            int vm = 0;
            const int32_t *xa = xs + (size_t)D.variant * L.x_stride, *xb = xa;
            if (J.channels == 2 && D.variant >= 2) { vm = D.variant; xa = xs; xb = xs + L.x_stride; }
            FbBitRun r;
            fb_run_init(r, words, D.start_bit, D.start_bit + 1);
            r.w_last = 0xFFFFFFFFu;
            const uint32_t mask = D.bps >= 32 ? 0xFFFFFFFFu : ((1u << D.bps) - 1u);
#define FB_PUT_ATOMIC(val, nb) do { r.w_first = r.cur_w; fb_run_put(r, (uint32_t)(val), (uint32_t)(nb)); } while (0)
#define FB_KF_X(t) ((uint32_t)(vm == 0 ? xa[fb_xidx(t)] : (vm == 2 ? fb_mid(xa[fb_xidx(t)], xb[fb_xidx(t)]) : fb_side(xa[fb_xidx(t)], xb[fb_xidx(t)]))))
            if (D.type == FB200_SF_CONSTANT) {
                FB_PUT_ATOMIC(0x00, 8);
                FB_PUT_ATOMIC(FB_KF_X(0) & mask, D.bps);
            } else if (D.type == FB200_SF_VERBATIM) {
                FB_PUT_ATOMIC(0x02, 8);
            } else if (D.type == FB200_SF_FIXED) {
                FB_PUT_ATOMIC(0x10 | (D.order << 1), 8);
                for (int t = 0; t < D.order; t++) FB_PUT_ATOMIC(FB_KF_X(t) & mask, D.bps);
                FB_PUT_ATOMIC(((D.rice2 ? 1 : 0) << 4) | D.part_order, 6);
            } else {
                FB_PUT_ATOMIC(0x40 | ((D.order - 1) << 1), 8);
                for (int t = 0; t < D.order; t++) FB_PUT_ATOMIC(FB_KF_X(t) & mask, D.bps);
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
        for (int item = tid; item < J.channels * g.U; item += T) {
            const int c = item >> g.lgU, unit = item & (g.U - 1);
            const FbKfSub &D = S->sub[c];
            if (D.type == FB200_SF_CONSTANT) continue;
            int ta, tb;
            fb_kf_unit_range(g, unit, &ta, &tb);
            int vm = 0;
            const int32_t *xa = xs + (size_t)D.variant * L.x_stride, *xb = xa;
            if (J.channels == 2 && D.variant >= 2) { vm = D.variant; xa = xs; xb = xs + L.x_stride; }
            if (D.type == FB200_SF_VERBATIM) {
                // Verbatim::write (src/component/bitrepr.rs:463-470): bps bits per sample at fixed positions
                if (tb <= ta) continue;
                const uint32_t mask = D.bps >= 32 ? 0xFFFFFFFFu : ((1u << D.bps) - 1u);
                FbBitRun r;
                fb_run_init(r, words, D.start_bit + 8u + (uint32_t)ta * (uint32_t)D.bps,
                            D.start_bit + 8u + (uint32_t)tb * (uint32_t)D.bps);
                for (int t = ta; t < tb; t++) fb_run_put(r, FB_KF_X(t) & mask, (uint32_t)D.bps);
                fb_run_flush(r);
                continue;
            }
            // Residual::write (src/component/bitrepr.rs:550-597)
            const fb200_subframe_info &V = choice[D.variant];
            const uint32_t *ub = (const uint32_t *)(smem + L.off_keep + (uint32_t)D.variant * L.keep_bytes + L.k_unit_bits) +
                                 (size_t)D.cand * (L.U_max + 1);
            const uint32_t p0 = D.code_bit + ub[unit], p1 = D.code_bit + ub[unit + 1];
            if (p1 == p0) continue;
            const int ush = g.lgU - D.part_order;
            const uint32_t rp = V.rice_params[unit >> ush];
            FbBitRun r;
            fb_run_init(r, words, p0, p1);
            if ((unit & ((1 << ush) - 1)) == 0) fb_run_put(r, rp, D.rice2 ? 5u : 4u);
            const int warm = D.order;
            const int lo = ta > warm ? ta : warm;
            const int kind = D.type == FB200_SF_LPC ? 1 : 0;
            int32_t qq[G];
            unsigned long long sumabs = 0;
#pragma unroll
            for (int j = 0; j < G; j++) {
                qq[j] = (kind == 1 && j < D.order) ? (int32_t)V.qlp[j] : 0;
                sumabs += (unsigned long long)(qq[j] < 0 ? -qq[j] : qq[j]);
            }
            const bool narrow = kind == 0 ||
                                (unsigned long long)ana[(size_t)f * (size_t)J.nvar + (size_t)D.variant].max_abs * sumabs < 0x7FFFFFFFull;
            const uint32_t rmask = (1u << rp) - 1u, rone = 1u << rp;
            for (int t0 = ta; t0 < tb; t0 += FB_RUN) {
                int32_t win[G + FB_RUN];
                uint32_t uu[FB_RUN];
                fb_kf_window<G>(xa, xb, vm, t0, n, win);
                fb_kf_run_u<G>(win, t0, lo, tb, kind, D.order, qq, D.shift, narrow, uu);
#pragma unroll
                for (int i = 0; i < FB_RUN; i++) {
                    const int t = t0 + i;
                    if (t >= lo && t < tb) {
                        fb_run_skip(r, uu[i] >> rp);
                        fb_run_put(r, (uu[i] & rmask) | rone, rp + 1u);
                    }
                }
            }
            fb_run_flush(r);
        }
#undef FB_KF_X
    FB_PHASE_END

    // ---- CRC-16 over data_bytes (Frame::write, src/component/bitrepr.rs:289-320): chunks aligned to the
    // END of the data (leading zero bytes do not change a CRC with init 0), combined pairwise with
    // x^(8*Lc*2^k) mod P.  TC = largest power of two <= T threads take part.
    int TC = 32;
    while (TC * 2 <= T) TC *= 2;
    const uint32_t B = S->data_bytes;
    const uint32_t Lc = (B + (uint32_t)TC - 1u) / (uint32_t)TC;
    FB_PHASE(tid, T)
        if (tid == 0) {
            uint32_t result = 1, base = 2;
            uint32_t e = 8u * Lc;
            while (e) { if (e & 1u) result = fb_crc16_mulmod(result, base); base = fb_crc16_mulmod(base, base); e >>= 1; }
            S->crc_xpow[0] = result;
            for (int k = 1; k < 9; k++) S->crc_xpow[k] = fb_crc16_mulmod(S->crc_xpow[k - 1], S->crc_xpow[k - 1]);
        }
        if (tid < TC) {
            long long lo = (long long)B - (long long)(TC - tid) * (long long)Lc;
            const long long hi = lo + (long long)Lc;
            if (lo < 0) lo = 0;
            uint32_t crc = 0;
            for (long long i = lo; i < hi; i++) {
                const uint32_t byte = (words[i >> 2] >> (24u - 8u * (uint32_t)(i & 3))) & 0xFFu;
                crc = ((crc << 8) & 0xFFFFu) ^ S->crc_tab[((crc >> 8) ^ byte) & 0xFFu];
            }
            S->crc_part[tid] = crc;
        }
    FB_PHASE_END
    for (int k = 0; (1 << k) < TC; k++) {
        FB_PHASE(tid, T)
            const int span = 1 << (k + 1);
            if (tid < TC && (tid % span) == 0) {
                const uint32_t left = S->crc_part[tid], right = S->crc_part[tid + (1 << k)];
                S->crc_part[tid] = fb_crc16_mulmod(left, S->crc_xpow[k]) ^ right;
            }
        FB_PHASE_END
    }
    FB_PHASE(tid, T)
        if (tid == 0) {
            const uint32_t crc = S->crc_part[0];
            for (int i = 0; i < 2; i++) {
                const uint32_t pos = B + (uint32_t)i;
                const uint32_t byte = (crc >> (8 * (1 - i))) & 0xFFu;
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
            fb200_frame_info &I = infos[f];
            for (int c = 0; c < J.channels; c++) {
                const uint8_t *src = (const uint8_t *)&choice[S->sub[c].variant];
                uint8_t *dst = (uint8_t *)&I.sub[c];
                for (int i = tid; i < (int)sizeof(fb200_subframe_info); i += T) dst[i] = src[i];
            }
        }
    FB_PHASE_END
    // ---- store: big-endian words -> bytes of the slot
    FB_PHASE(tid, T)
        const uint32_t nwords = (B + 2u + 3u) / 4u;
        uint32_t *dstw = (uint32_t *)slot;
        for (uint32_t w = (uint32_t)tid; w < nwords; w += (uint32_t)T) {
            const uint32_t v = words[w];
            dstw[w] = ((v & 0xFFu) << 24) | ((v & 0xFF00u) << 8) | ((v >> 8) & 0xFF00u) | (v >> 24);
        }
    FB_PHASE_END
}
