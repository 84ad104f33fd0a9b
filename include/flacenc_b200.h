/*
 * flacenc_b200.h -- C ABI of the B200-native backend for flacenc-rs's per-frame encode path.
 *
 * The reference (yotarok/flacenc-rs, a pure Rust crate) has no FFI/plugin interface; the seam this
 * library sits behind is its public Rust API (paths relative to /root/reference/):
 *
 *   encode_with_fixed_block_size(config, src, block_size) -> Stream     src/coding.rs:645-695
 *   encode_fixed_size_frame(config, framebuf, frame_number, stream_info) src/coding.rs:581-606
 *   par::encode_with_fixed_block_size (frame-parallel worker pool)       src/par.rs:355-449
 *
 * A Rust shim (`extern "C"` block + build.rs running nvcc) binds exactly the
 * entry points below.  Plain pointers and sizes only; the caller owns every buffer.  There is no
 * CPU fallback: every encode entry point runs hand-written sm_100a kernels and fails with
 * FB200_ERR_CUDA when no device is usable.
 *
 * Return codes mirror the reference's error enums (src/error.rs:458-463):
 *   FB200_OK, FB200_ERR_CONFIG  (EncodeError::Config / VerifyError, incl. out-of-range samples and
 *   frame_number >= 2^31, src/coding.rs:587-593), FB200_ERR_SOURCE (SourceError / bad argument),
 *   FB200_ERR_CUDA, FB200_ERR_CAPACITY (output buffer too small).
 */
#ifndef FLACENC_B200_H
#define FLACENC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB200_OK 0
#define FB200_ERR_CONFIG 1
#define FB200_ERR_SOURCE 2
#define FB200_ERR_CUDA 3
#define FB200_ERR_CAPACITY 4

#define FB200_MAX_CHANNELS 8       /* src/constant.rs:60 */
#define FB200_MAX_LPC_ORDER 24     /* src/constant.rs:118 */
#define FB200_MAX_RICE_PARTS 256   /* block <= 32767, partitions >= 64 samples -> order <= 8 */

/* POD mirror of config::Encoder and its children (src/config.rs:85-432). */
typedef struct fb200_config {
    int32_t block_size;             /* Encoder.block_size (default 4096); informative, see fb200_create */
    int32_t multithread;            /* Encoder.multithread: accepted, no effect (frames are always parallel) */
    int32_t workers;                /* Encoder.workers: accepted, no effect */
    int32_t use_leftside;           /* StereoCoding (src/config.rs:137-151) */
    int32_t use_rightside;
    int32_t use_midside;
    int32_t use_constant;           /* SubFrameCoding (src/config.rs:167-191) */
    int32_t use_fixed;
    int32_t use_lpc;
    int32_t fixed_max_order;        /* Fixed.max_order (default 4; values > 4 act as 4, src/coding.rs:309-313) */
    int32_t fixed_order_sel;        /* OrderSel: 0 = BitCount, 1 = ApproxEnt (default) */
    int32_t approx_ent_partitions;  /* OrderSel::ApproxEnt.partitions (default 16, 1..=64) */
    int32_t lpc_order;              /* Qlpc.lpc_order (default 10, 1..=24) */
    int32_t quant_precision;        /* Qlpc.quant_precision (default 15, 1..=15) */
    int32_t use_direct_mse;         /* Qlpc.use_direct_mse (`experimental` feature, src/config.rs:276-285): covariance-method
                                       LPC (src/lpc.rs:852-913) instead of autocorrelation + Levinson; 0 (default) or 1 */
    int32_t mae_optimization_steps; /* `experimental` IRLS-MAE refinement steps (src/lpc.rs:814-850); used when use_direct_mse */
    int32_t window_type;            /* Window: 0 = Rectangle, 1 = Tukey (default) */
    float   tukey_alpha;            /* Window::Tukey.alpha (default 0.4, 0..=1) */
    int32_t prc_max_parameter;      /* Prc.max_parameter (default 30, 0..=30) */
    /* ---- extensions beyond the reference (opt-in: any non-zero value gives frames the reference would not produce; they
     *      decode to the same samples and are never larger) ---- */
    int32_t ext_lpc_order_search;   /* 0 (default, off) .. 8: besides lpc_order P also try the k lower orders P - i * ceil(P / (k + 1)),
                                       i = 1..k (those >= 1), each the Levinson solution of that order on the same autocorrelation;
                                       the LPC candidate with the fewest subframe bits wins, the higher order on ties.
                                       Autocorrelation estimator only (rejected with use_direct_mse). */
    int32_t ext_lpc_precision_search; /* 0 (default, off) .. 4: also quantise the order-P coefficients with quant_precision - 1 ..
                                       quant_precision - k bits (those >= 1) and take the cheapest; candidates are ranked after
                                       the lower orders.  ext_lpc_order_search + ext_lpc_precision_search <= 8. */
} fb200_config;

/* Subframe types (component::SubFrame, src/component/datatype.rs:1782-1795). */
enum { FB200_SF_CONSTANT = 0, FB200_SF_VERBATIM = 1, FB200_SF_FIXED = 2, FB200_SF_LPC = 3 };

/* Everything needed to rebuild component::SubFrame on the Rust side through the public verified
 * constructors (Constant::new / Verbatim::new / FixedLpc::new / Lpc::new / Residual::new,
 * src/component/datatype.rs:1849,1927,1995,2088,2292) without re-running the analysis. */
typedef struct fb200_subframe_info {
    int32_t  type;            /* FB200_SF_* */
    int32_t  order;           /* fixed order, or qlpc order after tail-zero truncation */
    int32_t  bits_per_sample; /* of this subframe (side channel: stream bps + 1) */
    int32_t  precision;       /* LPC only */
    int32_t  shift;           /* LPC only */
    int32_t  partition_order; /* FIXED/LPC: Residual.partition_order */
    int32_t  rice2;           /* 1 when any Rice parameter > 14 (5-bit parameters) */
    int32_t  reserved;
    int16_t  qlp[32];         /* LPC only: quantized coefficients, first `order` valid */
    uint8_t  rice_params[FB200_MAX_RICE_PARTS];
    uint64_t bits;            /* BitRepr::count_bits() of the subframe */
} fb200_subframe_info;

typedef struct fb200_frame_info {
    int32_t  channel_assignment; /* header tag: channels-1 (independent), 8 = L/S, 9 = R/S, 10 = M/S */
    int32_t  block_size;         /* samples per channel in this frame */
    uint32_t frame_number;
    uint32_t frame_bytes;
    fb200_subframe_info sub[FB200_MAX_CHANNELS];
} fb200_frame_info;

/* Debug/parity taps of the float tier (fb200_analyze): one entry per channel variant.
 * For stereo streams the variants of a frame are L, R, M, S (in that order); otherwise one per channel. */
typedef struct fb200_variant_taps {
    double   autocorr[FB200_MAX_LPC_ORDER + 1]; /* src/lpc.rs:533-548, lags 0..lpc_order */
    double   lpc[FB200_MAX_LPC_ORDER];          /* src/lpc.rs:633-705 */
    int16_t  qlp[32];                           /* src/lpc.rs:273-302 */
    int32_t  qlp_order;
    int32_t  qlp_shift;
    int32_t  is_constant;                       /* src/arrayutils.rs:382-389 */
    int32_t  fixed_order;                       /* winner of select_order (-1 = None), src/coding.rs:230-288 */
    uint64_t fixed_est_bits[5];                 /* per-order selector cost incl. bps*order */
} fb200_variant_taps;

typedef struct fb200_ctx fb200_ctx;

/* Timing of the last encode call on a context, in milliseconds (CUDA events on the context's stream). */
typedef struct fb200_timing {
    float h2d_ms, kernels_ms, d2h_ms, total_ms;
    /* fused path: k_rice_ms = analysis/plan kernel, k_pack_ms = pack kernel (+ gather of listed frames),
     * k_gather_ms = generic kernels over the fallback list + scan of the frame sizes;
     * generic path: k_rice_ms = Rice search, k_pack_ms = frame assembly, k_gather_ms = scan + gather */
    float k_ingest_ms, k_analyze_ms, k_rice_ms, k_pack_ms, k_gather_ms;
    uint64_t launches;            /* kernels launched by the call */
    uint64_t in_bytes, out_bytes; /* PCM bytes consumed / frame bytes produced */
    uint64_t fused_frames;        /* frames encoded by the fused per-frame kernel */
    uint64_t fallback_frames;     /* frames that kernel handed to the generic (literal replay) kernels */
} fb200_timing;

/* ---- configuration (src/config.rs) ---- */
void fb200_config_default(fb200_config *cfg);          /* config::Encoder::default() */
int  fb200_config_verify(const fb200_config *cfg);     /* Verify::verify -> FB200_OK / FB200_ERR_CONFIG */

/* ---- lifecycle ---- */
int  fb200_device_count(void);                         /* usable CUDA devices (0 if none) */
/* Replaces `Verified<config::Encoder>` + StreamInfo::new + FrameBuf::with_size for one stream format
 * on one device (src/coding.rs:656-660).  Verifies the config and the format (channels 1..=8,
 * bits_per_sample 8..=24, block_size 32..=32767, src/constant.rs:38-60).  *err receives FB200_*. */
fb200_ctx *fb200_create(const fb200_config *cfg, int channels, int bits_per_sample, int sample_rate,
                        int block_size, int device, int *err);
void fb200_destroy(fb200_ctx *ctx);
/* Upper bound of one encoded frame in bytes (verbatim subframes), for sizing out buffers. */
size_t fb200_max_frame_bytes(const fb200_ctx *ctx);

/* ---- the hot path ---- */
/* Replaces the frame loop of encode_with_fixed_block_size / par.rs (src/coding.rs:662-674,
 * src/par.rs:377-403): encodes ceil(n_samples_per_ch / block_size) frames from interleaved PCM in
 * HOST memory (what Fill::fill_le_bytes / fill_interleaved receive, src/source.rs:278-299):
 * container_bytes = 2 or 3 -> packed little-endian samples, 4 -> int32 samples.
 * Frame i gets number first_frame_number + i; the last frame may be short.
 * out_bytes receives the concatenated frames, frame_sizes[i] each frame's length.
 * infos (nullable) receives one decision record per frame. */
int fb200_encode_interleaved(fb200_ctx *ctx, const void *pcm, int container_bytes,
                             uint64_t n_samples_per_ch, uint64_t first_frame_number,
                             uint8_t *out_bytes, size_t out_cap, uint32_t *frame_sizes,
                             fb200_frame_info *infos, size_t *n_frames, size_t *out_len);

/* Same work with input and output resident in device memory (no host copies): d_pcm and d_out are
 * device pointers on the context's device; frame_sizes/out_len are host pointers filled after a
 * stream synchronize.  Used to separate kernel throughput from PCIe in measurements. */
int fb200_encode_device(fb200_ctx *ctx, const void *d_pcm, int container_bytes,
                        uint64_t n_samples_per_ch, uint64_t first_frame_number,
                        uint8_t *d_out, size_t out_cap, uint32_t *frame_sizes,
                        size_t *n_frames, size_t *out_len);

/* Replaces encode_fixed_size_frame (src/coding.rs:581-606) on a FrameBuf layout:
 * planar[ch * stride + t], t < n (src/source.rs:115-127,251-253). */
int fb200_encode_planar_frame(fb200_ctx *ctx, const int32_t *planar, int stride, int n,
                              uint32_t frame_number, uint8_t *out, size_t out_cap, size_t *out_len,
                              fb200_frame_info *info);

/* Parity taps: runs ingest + the analysis kernel only and returns one record per channel variant
 * (frames * variants_per_frame entries; variants_per_frame = 4 for stereo, else channels). */
int fb200_analyze(fb200_ctx *ctx, const void *pcm, int container_bytes, uint64_t n_samples_per_ch,
                  fb200_variant_taps *taps, size_t taps_cap, size_t *n_variants);

/* The same over several devices of one box: ctxs[0..n_ctx) are contexts of ONE configuration and stream format on
 * different devices.  Frames are independent, so they are shared by frame range -- chunk c of the batch goes to
 * device c mod n_ctx, each device with its own copy / compute streams, what src/par.rs:355-449 does with worker
 * threads -- and every chunk's bytes are copied straight to their final offset in out_bytes.  No collective.
 * fb200_last_timing(ctxs[0]) reports the call: total_ms = the longest per-device span, kernel times summed. */
int fb200_encode_interleaved_sharded(fb200_ctx *const *ctxs, int n_ctx, const void *pcm, int container_bytes,
                                     uint64_t n_samples_per_ch, uint64_t first_frame_number, uint8_t *out_bytes,
                                     size_t out_cap, uint32_t *frame_sizes, size_t *n_frames, size_t *out_len);

/* ---- stream level: host assembly around the hot path ---- */
/* Replaces encode_with_fixed_block_size end to end (src/coding.rs:645-695): "fLaC" + STREAMINFO
 * (min/max block and frame size, total samples, MD5 of the packed LE samples, src/source.rs:406-429)
 * + frames.  Frames are sharded by contiguous frame range over `n_devices` devices (NULL/0 = device 0),
 * one host thread and one context per device (what par.rs does with worker threads); MD5 runs on its
 * own host thread like par.rs's ParContext (src/par.rs:196-277).  No collective is involved. */
int fb200_encode_stream(const fb200_config *cfg, const void *pcm, int container_bytes,
                        uint64_t n_samples_per_ch, int channels, int bits_per_sample, int sample_rate,
                        int block_size, const int *devices, int n_devices,
                        uint8_t *out, size_t out_cap, size_t *out_len);

/* A batch of streams of one format (src/par.rs:196-277 runs one MD5 thread per stream; a batch of files gives as many
 * independent MD5 chains as there are files): the MD5 of every stream runs on its own host thread, bounded by the
 * host's cores, while the calling thread feeds the devices stream by stream.  pcm[i] / n_samples[i] / out[i] /
 * out_cap[i] describe stream i; out_len[i] and rcs[i] (nullable) receive its length and result.  Returns the first
 * failure (FB200_OK if none). */
int fb200_encode_streams(const fb200_config *cfg, int n_streams, const void *const *pcm, const uint64_t *n_samples,
                         int container_bytes, int channels, int bits_per_sample, int sample_rate, int block_size,
                         const int *devices, int n_devices, uint8_t *const *out, const size_t *out_cap, size_t *out_len,
                         int *rcs);
/* MD5 (RFC 1321) of a byte range with the library's implementation: the sequential floor of the stream-level calls. */
void fb200_md5(const void *data, size_t len, uint8_t digest[16]);
/* The stream-level calls keep their device contexts between calls (like the reference's thread-local scratch,
 * src/lib.rs:92-116); this destroys them. */
void fb200_pool_clear(void);

/* ---- parity pins (diagnostics) ----
 * Device builds of the two scalar float helpers whose last bit decides encoder choices, evaluated on
 * caller-provided inputs so tests can compare them with the host libm the reference uses:
 *   log2f of estimate_entropy (src/coding.rs:200-227): out_bits[i] = bits of log2f(float with bits first_bits + i)
 *   find_shift (src/lpc.rs:234-255) of the one-coefficient set {values[i]}: out_shift[i]
 *   the IRLS weight (src/lpc.rs:828, powf(-1.2) inside) of the raw error with bits first_bits + i: out_bits[i] */
int fb200_debug_log2f(int device, uint32_t first_bits, uint64_t count, uint32_t *out_bits);
int fb200_debug_irls_weight(int device, uint32_t first_bits, uint64_t count, float normalizer, uint32_t *out_bits);
/* The chunk schedule of the (sharded) host path: chunk c = frames [first[c], first[c] + count[c]) of a batch of
 * total_frames frames cut with a nominal chunk size; chunk c is encoded on device c mod n.  Host logic only (no device
 * needed).  Returns the number of chunks; fills at most cap entries. */
size_t fb200_debug_chunk_schedule(uint64_t total_frames, uint64_t chunk_frames, uint64_t *first, uint64_t *count, size_t cap);
int fb200_debug_find_shift(int device, const double *values, uint64_t count, int precision, int32_t *out_shift);

/* ---- diagnostics ---- */
int  fb200_last_timing(const fb200_ctx *ctx, fb200_timing *t);
const char *fb200_strerror(int code);
const char *fb200_last_error(const fb200_ctx *ctx); /* detail text of the last failure on ctx; ctx == NULL: of the last
                                                       failed stream-level call on the calling thread */
const char *fb200_version(void);

#ifdef __cplusplus
}
#endif
#endif
