/*
 * flacenc_oracle.h -- CPU restatement of flacenc-rs's per-frame hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (flacenc_rs_b200/csrc, libflacenc_b200.so) never links, includes or calls anything here.
 *
 * Flavour: the reference's *stable* build (scalar `fakesimd`, sequential evaluation order;
 * /root/reference/src/fakesimd.rs, src/arrayutils.rs:427-439).
 *
 * Parity pinning: every known-answer test the reference holds for this path is replayed in
 * tests/test_oracle_kat.py (see the table in DESIGN.md).  The Rust crate itself cannot be
 * compiled in this image (no cargo/rustc), and the reference has no golden .flac streams, so
 * whole-frame bytes are pinned by (i) those KATs, (ii) an independently written decoder
 * (fo_decode_stream) that must reproduce the PCM and both CRCs, (iii) the reference's MD5 KATs.
 * libm last-ULP behaviour (cosf/log2f/log2) is the host glibc's: "parity unpinned" for that
 * detail only.
 *
 * All paths cited are relative to /root/reference/.
 */
#ifndef FLACENC_ORACLE_H
#define FLACENC_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FO_MAX_LPC_ORDER 24        /* src/constant.rs:118 */
#define FO_MAX_FIXED_ORDER 4       /* src/constant.rs:94  */
#define FO_MAX_RICE_PARAMETER 30   /* src/constant.rs:140 */
#define FO_MAX_RICE_PARTS 256      /* n <= 32767, min partition 64 -> order <= 8 */
#define FO_MAX_CHANNELS 8
#define FO_MIN_BLOCK_SIZE 32
#define FO_MAX_BLOCK_SIZE 32767
#define FO_RICE_SAT ((1u << 27) - 1u) /* src/rice.rs:51 */

/* POD mirror of config::Encoder (src/config.rs:85-432).  Same field order as
 * fb200_config in include/flacenc_b200.h so tests can share one ctypes.Structure. */
typedef struct fo_config {
    int32_t block_size;            /* Encoder.block_size, default 4096 */
    int32_t multithread;           /* Encoder.multithread, default 1 (feature "par") */
    int32_t workers;               /* Encoder.workers, 0 = None */
    int32_t use_leftside;          /* StereoCoding */
    int32_t use_rightside;
    int32_t use_midside;
    int32_t use_constant;          /* SubFrameCoding */
    int32_t use_fixed;
    int32_t use_lpc;
    int32_t fixed_max_order;       /* Fixed.max_order, default 4 */
    int32_t fixed_order_sel;       /* 0 = BitCount, 1 = ApproxEnt */
    int32_t approx_ent_partitions; /* OrderSel::ApproxEnt.partitions, default 16 */
    int32_t lpc_order;             /* Qlpc.lpc_order, default 10 */
    int32_t quant_precision;       /* Qlpc.quant_precision, default 15 */
    int32_t use_direct_mse;        /* `experimental` feature: covariance-method LPC (src/lpc.rs:852-903) */
    int32_t mae_optimization_steps;/* `experimental` feature: IRLS refinement steps of the direct-MSE estimate (src/lpc.rs:814-850) */
    int32_t window_type;           /* 0 = Rectangle, 1 = Tukey */
    float   tukey_alpha;           /* default 0.4 */
    int32_t prc_max_parameter;     /* Prc.max_parameter, default 30 */
    int32_t ext_lpc_order_search;  /* EXTENSION beyond the reference (0 = off): lower LPC orders tried, see fo_ext_lpc_orders */
    int32_t ext_lpc_precision_search; /* EXTENSION (0 = off): lower quantiser precisions tried for the order-P coefficients */
} fo_config;

enum { FO_SF_CONSTANT = 0, FO_SF_VERBATIM = 1, FO_SF_FIXED = 2, FO_SF_LPC = 3 };
enum { FO_CH_INDEPENDENT = 0, FO_CH_LEFT_SIDE = 8, FO_CH_RIGHT_SIDE = 9, FO_CH_MID_SIDE = 10 };

/* Decision record for one subframe (what component::SubFrame carries). */
typedef struct fo_subframe {
    int32_t  type;                 /* FO_SF_* */
    int32_t  order;                /* fixed order / truncated qlpc order */
    int32_t  bps;                  /* bits per sample of this subframe (side: +1) */
    int32_t  n;                    /* block size */
    int32_t  precision;            /* lpc only */
    int32_t  shift;                /* lpc only */
    int32_t  part_order;           /* residual partition order */
    int32_t  rice2;                /* 1 when any parameter > 14 */
    int16_t  qlp[32];              /* lpc only, first `order` valid */
    uint8_t  rice_params[FO_MAX_RICE_PARTS];
    uint64_t sum_quotients;
    uint64_t code_bits;            /* PrcParameter.code_bits (the search's estimate) */
    uint64_t bits;                 /* BitRepr::count_bits() of the subframe */
    int32_t *residual;             /* n entries, warm-up slots zero; owned; NULL for const/verbatim */
    const int32_t *samples;        /* borrowed pointer to the coded signal */
} fo_subframe;

typedef struct fo_frame {
    int32_t channels;
    int32_t n;                     /* block size of this frame */
    int32_t bps;                   /* stream bits per sample */
    int32_t sample_rate;
    int32_t ch_assignment;         /* 0 independent, 8/9/10 = LS/RS/MS (header tag) */
    uint32_t frame_number;
    fo_subframe sub[FO_MAX_CHANNELS];
    int32_t *ms_buf;               /* owned: M then S, stride n (stereo only) */
} fo_frame;

/* ---- config (src/config.rs) ---- */
void fo_config_default(fo_config *cfg);
int  fo_config_verify(const fo_config *cfg);  /* 0 ok, 1 = VerifyError */

/* ---- lpc.rs ---- */
void   fo_window_weights(int window_type, float alpha, int len, float *out);
void   fo_fill_windowed_signal(const int32_t *signal, const float *window, int n, float *out);
void   fo_auto_correlation_f64(int order, const float *signal, int n, double *dest);
void   fo_auto_correlation_f32(int order, const float *signal, int n, float *dest);
void   fo_levinson_f64(const double *coefs, const double *ys, int order, double *dest);
void   fo_levinson_f32(const float *coefs, const float *ys, int order, float *dest);
int    fo_find_shift(const double *coefs, int n, int precision);
/* host libm tables for the parity pins of the device helpers: out[i] = bits of log2f(float with bits first + i)
 * (the f32::log2 of estimate_entropy), on `threads` host threads; shifts[i] = fo_find_shift of {values[i]} */
void   fo_log2f_bits(uint32_t first, uint64_t count, int threads, uint32_t *out);
void   fo_find_shift_each(const double *values, uint64_t count, int precision, int32_t *shifts);
/* returns the truncated order; q_out has FO_MAX_LPC_ORDER entries */
int    fo_quantize_parameters(const double *coefs, int n, int precision, int16_t *q_out, int *shift_out);
void   fo_compute_error(const int16_t *q, int order, int shift, const int32_t *signal, int n, int32_t *errors);
void   fo_lpc_from_autocorr(const int32_t *signal, int n, int window_type, float alpha, int lpc_order,
                            double *coefs_out, double *corr_out /* nullable, lpc_order+1 */);

/* `experimental` feature (src/lpc.rs:573-600, 76-87, 852-913; IRLS refinement :606-618, :814-850) */
void   fo_lagged_outer_prod_sum(int order, const float *signal, int len, double *dest);
int    fo_solve_sym(const double *mat, int n, double *v);
void   fo_lpc_with_direct_mse(const int32_t *signal, int n, int window_type, float alpha, int lpc_order,
                              double *coefs_out, double *corr_out /* nullable */, double *covar_out /* nullable */);
void   fo_compute_raw_errors(const int32_t *signal, int n, const double *coefs, int lpc_order, float *errors);
int    fo_ext_lpc_orders(int lpc_order, int k, int *orders /* up to k + 1 */); /* EXTENSION, see fo_config */
float  fo_irls_weight(float err, float normalizer);
void   fo_irls_weight_bits(uint32_t first_bits, uint64_t count, float normalizer, int threads, uint32_t *out_bits);
void   fo_lpc_with_irls_mae(const int32_t *signal, int n, int window_type, float alpha, int lpc_order, int steps,
                            double *coefs_out, float *sums_out /* nullable, steps + 1 */);

/* ---- rice.rs ---- */
uint32_t fo_encode_signbit(int32_t v);
int32_t  fo_decode_signbit(uint32_t v);
int      fo_finest_partition_order(int size, int min_part_size);
void     fo_bit_table_from_errors(const uint32_t *errors, int n, uint32_t offset, uint32_t *table32);
void     fo_bit_table_minimizer(const uint32_t *table32, int max_p, int *p_out, uint32_t *bits_out);
void     fo_bit_table_merge(const uint32_t *a, const uint32_t *b, uint32_t offset, uint32_t *out);
/* returns partition order; ps has FO_MAX_RICE_PARTS entries */
int      fo_find_partitioned_rice_parameter(const int32_t *signal, int n, int warmup, int max_p,
                                            uint8_t *ps, uint64_t *code_bits);

/* ---- coding.rs ---- */
void     fo_fixed_lpc_errors(const int32_t *signal, int n, int32_t *errors5 /* 5*n */);
uint64_t fo_estimate_entropy(const int32_t *errors, int n, int warmup, int partitions);
/* select_order_and_encode_residual: returns chosen order or -1 (None). order_sel: 0 BitCount, 1 ApproxEnt */
int      fo_select_order(int order_sel, int partitions, int max_p, const int32_t *errors, int n_orders, int n,
                         int bps, uint64_t baseline_bits, uint64_t *bits_out);
void     fo_encode_subframe(const fo_config *cfg, const int32_t *samples, int n, int bps, fo_subframe *out);
/* encode_fixed_size_frame on a planar FrameBuf (samples[ch*stride + t], t < n). 0 ok, 1 verify error */
int      fo_encode_frame(const fo_config *cfg, const int32_t *planar, int channels, int stride, int n,
                         int bps, int sample_rate, uint32_t frame_number, fo_frame *out);
void     fo_frame_free(fo_frame *f);

/* ---- component/bitrepr.rs ---- */
uint8_t  fo_crc8(const uint8_t *data, size_t len);
uint16_t fo_crc16(const uint8_t *data, size_t len);
int      fo_encode_utf8like(uint64_t val, uint8_t *out7);            /* bytes written, -1 on error */
int      fo_frame_header_bytes(int n, int ch_tag, int bps, int sample_rate, int variable, uint64_t number,
                               uint8_t *out16);                        /* incl. CRC-8 */
uint64_t fo_frame_count_bits(const fo_frame *f);
int64_t  fo_subframe_write(const fo_subframe *sf, uint8_t *out, size_t cap); /* test hook: bits written */
/* writes the frame (header, subframes, padding, CRC-16); returns bytes or -1 if cap too small */
int64_t  fo_frame_write(const fo_frame *f, uint8_t *out, size_t cap);

/* ---- whole path: batch of frames from interleaved PCM (i32 samples) ---- */
/* Mirrors the loop of encode_with_fixed_block_size (src/coding.rs:662-674) without MD5/STREAMINFO.
 * nthreads > 1 mirrors par.rs (frames are independent; results ordered by frame number).
 * frame_sizes has ceil(n_samples/block_size) entries. Returns total bytes, -1 = verify error, -2 = cap. */
int64_t  fo_encode_frames(const fo_config *cfg, const int32_t *interleaved, uint64_t n_samples_per_ch,
                          int channels, int bps, int sample_rate, int block_size, uint32_t first_frame_number,
                          int nthreads, uint8_t *out, size_t cap, uint32_t *frame_sizes);
/* Whole stream ("fLaC" + STREAMINFO with MD5 + frames), src/coding.rs:645-695. */
int64_t  fo_encode_stream(const fo_config *cfg, const int32_t *interleaved, uint64_t n_samples_per_ch,
                          int channels, int bps, int sample_rate, int block_size, int nthreads,
                          uint8_t *out, size_t cap);

/* ---- source.rs ---- */
void     fo_md5(const uint8_t *data, size_t len, uint8_t digest[16]);
void     fo_md5_of_samples(const int32_t *interleaved, size_t count, int bytes_per_sample, uint8_t digest[16]);

/* ---- independent decoder (tier-1 check); written from the FLAC format, not from parser.rs ---- */
typedef struct fo_stream_info {
    uint32_t min_block, max_block, min_frame, max_frame, sample_rate, channels, bps;
    uint64_t total_samples;
    uint8_t  md5[16];
    uint64_t n_frames;
} fo_stream_info;
/* Decodes a whole stream; out_interleaved may be NULL to just count. Returns samples per channel,
 * or <0: -1 bad magic/metadata, -2 sync/header, -3 crc8, -4 crc16, -5 subframe, -6 capacity. */
int64_t  fo_decode_stream(const uint8_t *data, size_t len, fo_stream_info *info,
                          int32_t *out_interleaved, uint64_t cap_samples_per_ch);
/* Decodes consecutive frames without a stream header (needs bps/channels for tag-0 cases). */
int64_t  fo_decode_frames(const uint8_t *data, size_t len, int channels, int bps,
                          int32_t *out_interleaved, uint64_t cap_samples_per_ch, uint64_t *n_frames);

#ifdef __cplusplus
}
#endif
#endif
