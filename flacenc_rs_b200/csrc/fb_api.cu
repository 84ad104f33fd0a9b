// fb_api.cu -- CUDA kernel wrappers and the C ABI of libflacenc_b200.so (include/flacenc_b200.h).
//
// One context = one stream format on one device.  A call encodes a batch of frames in chunks:
//   H2D(pcm) -> K0 ingest -> K1 analyze -> K2 rice -> K3 pack -> K4 scan+gather -> D2H(bytes, sizes)
// on the context's stream.  There is no CPU path: without a usable device every entry point
// returns FB200_ERR_CUDA.  Reference citations are relative to /root/reference/.
#include <cuda_runtime.h>

#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "fb_host.h"
#include "fb_launch.h"
#include "fb_md5.h"

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fb_k0_ingest(FbJob J, const uint8_t *pcm, int32_t *xv, uint32_t *err_flag) {
    uint64_t s = (uint64_t)blockIdx.x * 256u + threadIdx.x;
    if (s < J.n_samples) fb_k0_sample(J, pcm, xv, err_flag, s);
}

__global__ void __launch_bounds__(256) fb_k0_ingest_planar(FbJob J, const int32_t *src, int src_stride, int32_t *xv,
                                                           uint32_t *err_flag) {
    int t = (int)(blockIdx.x * 256u + threadIdx.x);
    if (t < J.tail_n) fb_k0_planar_sample(J, src, src_stride, xv, err_flag, t);
}

// offsets[i] = *total + exclusive prefix; *total advances by the chunk's bytes (single CTA)
__global__ void __launch_bounds__(FB_K4_THREADS) fb_k4_scan(const uint32_t *frame_bytes, unsigned long long *offsets,
                                                           uint32_t n_frames, unsigned long long *total) {
    __shared__ unsigned long long partials[FB_K4_THREADS];
    fb_k4_scan_body(frame_bytes, offsets, n_frames, partials);
    const unsigned long long base = *total;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i <= n_frames; i += FB_K4_THREADS) offsets[i] += base;
    __syncthreads();
    if (threadIdx.x == 0) *total = offsets[n_frames];
}

__global__ void __launch_bounds__(256) fb_k4_gather(const uint8_t *slots, uint32_t slot_bytes, const uint32_t *frame_bytes,
                                                    const unsigned long long *offsets, uint8_t *out,
                                                    unsigned long long out_cap) {
    fb_k4_gather_thread(slots, slot_bytes, frame_bytes, offsets, out, out_cap, blockIdx.x, (int)threadIdx.x, 256);
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct fb200_ctx {
    fb200_config cfg;
    int channels, bps, sample_rate, block_size, device;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[16];
    int n_ev = 0;
    DevBuf pcm, xv, win_full, win_tail, ana, taps, choice, slots, frame_bytes, offsets, out, infos, scalars, fb_list, ktab;
    int win_tail_n = -1;
    void *pinned = nullptr; // small pinned staging: err flag, total bytes
    size_t pinned_cap = 0;
    fb200_timing timing;
    std::string last_error;
    int k2_smem_set = 0, k3_smem_set = 0, kf_smem_set = 0;
    uint32_t ktab_chunk = 0; // CRC chunk length the uploaded tables were built for
    bool force_generic = false; // FB200_FORCE_GENERIC=1: never use the fused kernel (tests exercise both paths)
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
        FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        FB_CUDA(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    FB_CUDA(ctx, cudaMalloc(&b.p, want));
    b.cap = want;
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

const char *fb200_version(void) { return "flacenc_b200 0.1.0 (sm_100a)"; }

const char *fb200_last_error(const fb200_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }

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
    }
    bool ok = cudaSetDevice(device) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < 16; i++) {
        ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
        if (ok) ctx->n_ev = i + 1;
    }
    ctx->pinned_cap = 4096;
    ok = ok && cudaHostAlloc(&ctx->pinned, ctx->pinned_cap, cudaHostAllocDefault) == cudaSuccess;
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
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->pcm, &ctx->xv, &ctx->win_full, &ctx->win_tail, &ctx->ana, &ctx->taps, &ctx->choice,
                      &ctx->slots, &ctx->frame_bytes, &ctx->offsets, &ctx->out, &ctx->infos, &ctx->scalars, &ctx->fb_list, &ctx->ktab};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    for (int i = 0; i < ctx->n_ev; i++) cudaEventDestroy(ctx->ev[i]);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
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

int fb_upload_window(fb200_ctx *ctx, DevBuf &buf, int n) {
    std::vector<float> w((size_t)n + 64, 0.f);
    fbh_window_weights(ctx->cfg.window_type, ctx->cfg.tukey_alpha, n, w.data());
    int rc = fb_reserve(ctx, buf, w.size() * sizeof(float));
    if (rc) return rc;
    FB_CUDA(ctx, cudaMemcpyAsync(buf.p, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // `w` dies at scope end
    return FB200_OK;
}

int fb_encode(fb200_ctx *ctx, const EncodeArgs &A) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->last_error.clear();
    FB_CUDA(ctx, cudaSetDevice(ctx->device));
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
    if (total_frames == 0) return FB200_OK;
    // frame_number < 2^31 (src/coding.rs:587-591)
    if (A.first_frame + total_frames > (1ull << 31)) {
        ctx->last_error = "frame number out of the 31-bit range";
        return FB200_ERR_CONFIG;
    }
    if (A.analyze_only && (uint64_t)A.taps_cap < total_frames * (uint64_t)nvar) return FB200_ERR_CAPACITY;

    // chunking: bound the working set (planar variants + slots) to ~2 GiB per pass
    const uint64_t stride = (uint64_t)((ctx->block_size + 31) & ~31);
    const uint32_t mb = fb_max_frame_bytes(ctx->channels, ctx->bps, ctx->block_size);
    const uint64_t per_frame = (uint64_t)nvar * stride * 4u + ((mb + 15u) & ~15u) + 4096u;
    uint64_t chunk_frames = std::max<uint64_t>(64, (2048ull << 20) / per_frame);
    chunk_frames = std::min<uint64_t>(chunk_frames, total_frames);

    // device scalars: [0] err flag (u32), [8] running total bytes (u64), [16] fallback-list length (u32)
    int rc;
    if ((rc = fb_reserve(ctx, ctx->scalars, 64))) return rc;
    FB_CUDA(ctx, cudaMemsetAsync(ctx->scalars.p, 0, 64, ctx->stream));
    uint32_t *d_err = (uint32_t *)ctx->scalars.p;
    unsigned long long *d_total = (unsigned long long *)((uint8_t *)ctx->scalars.p + 8);
    uint32_t *d_fb_count = (uint32_t *)((uint8_t *)ctx->scalars.p + 16);

    if ((rc = fb_reserve(ctx, ctx->frame_bytes, total_frames * 4u))) return rc;
    if ((rc = fb_reserve(ctx, ctx->offsets, (chunk_frames + 1) * 8u))) return rc;
    if ((rc = fb_reserve(ctx, ctx->xv, (chunk_frames * (uint64_t)nvar * stride + 64) * 4u))) return rc;
    if ((rc = fb_reserve(ctx, ctx->ana, chunk_frames * (uint64_t)nvar * sizeof(FbAnalysis)))) return rc;
    if (A.analyze_only) {
        if ((rc = fb_reserve(ctx, ctx->taps, chunk_frames * (uint64_t)nvar * sizeof(fb200_variant_taps)))) return rc;
    } else {
        if ((rc = fb_reserve(ctx, ctx->choice, chunk_frames * (uint64_t)nvar * sizeof(fb200_subframe_info)))) return rc;
        if ((rc = fb_reserve(ctx, ctx->slots, chunk_frames * (uint64_t)((mb + 15u) & ~15u)))) return rc;
        if ((rc = fb_reserve(ctx, ctx->fb_list, (chunk_frames + 1) * 4u))) return rc;
        if (A.infos && (rc = fb_reserve(ctx, ctx->infos, chunk_frames * sizeof(fb200_frame_info)))) return rc;
        if (A.out_host && (rc = fb_reserve(ctx, ctx->out, A.out_cap ? A.out_cap : 16))) return rc;
    }
    uint8_t *d_out = A.out_dev ? A.out_dev : (uint8_t *)ctx->out.p;

    // window tables (src/lpc.rs:217-231: cached per size)
    if (ctx->win_full.p == nullptr && (rc = fb_upload_window(ctx, ctx->win_full, ctx->block_size))) return rc;
    const int tail_n = (A.n_samples % bs) ? (int)(A.n_samples % bs) : ctx->block_size;
    if (tail_n != ctx->block_size && tail_n != ctx->win_tail_n) {
        if ((rc = fb_upload_window(ctx, ctx->win_tail, tail_n))) return rc;
        ctx->win_tail_n = tail_n;
    }
    const float *d_win_tail = (tail_n == ctx->block_size) ? (const float *)ctx->win_full.p : (const float *)ctx->win_tail.p;

    // host input staging area on the device
    const uint64_t in_bytes_total = A.n_samples * (uint64_t)ctx->channels * (uint64_t)cb;
    if (A.pcm_host) {
        uint64_t chunk_in = chunk_frames * bs * (uint64_t)ctx->channels * (uint64_t)cb;
        if ((rc = fb_reserve(ctx, ctx->pcm, std::min(chunk_in, in_bytes_total) + 16))) return rc;
    } else if (A.planar_host) {
        if ((rc = fb_reserve(ctx, ctx->pcm, (uint64_t)ctx->channels * (uint64_t)A.planar_stride * 4u + 16))) return rc;
    }

    // opt-in to large dynamic shared memory once
    FbJob J0 = fbh_make_job(ctx->cfg, ctx->channels, ctx->bps, ctx->sample_rate, ctx->block_size, cb ? cb : 4,
                            std::min<uint64_t>(A.n_samples, chunk_frames * bs), (uint32_t)A.first_frame);
    const int leaves_max = 1 << fb_finest_partition_order(ctx->block_size); // tail leaves <= block leaves? not always:
    int leaves_tail = 1 << fb_finest_partition_order(tail_n);
    const FbK2Layout L = fb_k2_layout(ctx->block_size, std::max(leaves_max, leaves_tail));
    const size_t k2_smem = L.total + 3 * sizeof(FbRiceResult);
    const size_t k3_smem = fb_k3_smem_bytes(mb, ctx->block_size, J0.pack_in_smem);
    if (k2_smem > 227u * 1024u || k3_smem > 227u * 1024u) {
        ctx->last_error = "internal: shared memory budget exceeded";
        return FB200_ERR_CUDA;
    }
    const int ring = fb_k1_ring(ctx->cfg.lpc_order);
    FbKfLayout KL;
    const bool fused = !A.analyze_only && !ctx->force_generic && fbh_fused_ok(J0, tail_n, &KL);
    if (fused && ctx->ktab_chunk != KL.crc_chunk) {
        std::vector<uint32_t> kt(fb_kf_ktab_words(KL.crc_chunk));
        fb_kf_build_ktab(KL.crc_chunk, kt.data());
        if ((rc = fb_reserve(ctx, ctx->ktab, kt.size() * 4u))) return rc;
        FB_CUDA(ctx, cudaMemcpyAsync(ctx->ktab.p, kt.data(), kt.size() * 4u, cudaMemcpyHostToDevice, ctx->stream));
        FB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // `kt` dies at scope end
        ctx->ktab_chunk = KL.crc_chunk;
    }
    if (fused && (int)KL.total > ctx->kf_smem_set) {
        FB_CUDA(ctx, fb_set_smem(ring, FB_KERNEL_KF, (int)KL.total));
        ctx->kf_smem_set = (int)KL.total;
    }
    if ((int)k2_smem > ctx->k2_smem_set) {
        FB_CUDA(ctx, fb_set_smem(ring, FB_KERNEL_K2, (int)k2_smem));
        ctx->k2_smem_set = (int)k2_smem;
    }
    if ((int)k3_smem > ctx->k3_smem_set) {
        FB_CUDA(ctx, fb_set_smem(ring, FB_KERNEL_K3, (int)k3_smem));
        ctx->k3_smem_set = (int)k3_smem;
    }

    cudaStream_t st = ctx->stream;
    float ms_h2d = 0, ms_k[5] = {0, 0, 0, 0, 0};
    uint64_t launches = 0, n_chunks = 0, fused_frames = 0;
    uint32_t *h_fb_counts = (uint32_t *)((uint8_t *)ctx->pinned + 64); // per-chunk fallback counts
    const uint64_t max_counts = (ctx->pinned_cap - 64) / 4;
    FB_CUDA(ctx, cudaEventRecord(ctx->ev[0], st)); // start of the call

    for (uint64_t f0 = 0; f0 < total_frames; f0 += chunk_frames) {
        const uint64_t nf = std::min(chunk_frames, total_frames - f0);
        const uint64_t s0 = f0 * bs;
        const uint64_t ns = std::min(A.n_samples - s0, nf * bs);
        FbJob J = fbh_make_job(ctx->cfg, ctx->channels, ctx->bps, ctx->sample_rate, ctx->block_size, cb ? cb : 4, ns,
                               (uint32_t)(A.first_frame + f0));
        const uint32_t nvars = J.n_frames * (uint32_t)J.nvar;
        uint32_t *d_fb = (uint32_t *)ctx->frame_bytes.p + f0;

        FB_CUDA(ctx, cudaEventRecord(ctx->ev[1], st));
        const uint8_t *d_pcm = nullptr;
        if (A.pcm_host) {
            const uint64_t off = s0 * (uint64_t)ctx->channels * (uint64_t)cb;
            const uint64_t len = ns * (uint64_t)ctx->channels * (uint64_t)cb;
            FB_CUDA(ctx, cudaMemcpyAsync(ctx->pcm.p, (const uint8_t *)A.pcm_host + off, len, cudaMemcpyHostToDevice, st));
            d_pcm = (const uint8_t *)ctx->pcm.p;
        } else if (A.pcm_dev) {
            d_pcm = (const uint8_t *)A.pcm_dev + s0 * (uint64_t)ctx->channels * (uint64_t)cb;
        } else {
            FB_CUDA(ctx, cudaMemcpyAsync(ctx->pcm.p, A.planar_host,
                                         (uint64_t)ctx->channels * (uint64_t)A.planar_stride * 4u,
                                         cudaMemcpyHostToDevice, st));
        }
        FB_CUDA(ctx, cudaEventRecord(ctx->ev[2], st));
        // K0
        if (A.planar_host) {
            fb_k0_ingest_planar<<<(unsigned)((J.tail_n + 255) / 256), 256, 0, st>>>(
                J, (const int32_t *)ctx->pcm.p, A.planar_stride, (int32_t *)ctx->xv.p, d_err);
        } else {
            fb_k0_ingest<<<(unsigned)((ns + 255) / 256), 256, 0, st>>>(J, d_pcm, (int32_t *)ctx->xv.p, d_err);
        }
        FB_CUDA(ctx, cudaEventRecord(ctx->ev[3], st));
        // K1
        fb_launch_k1(ring, J, (const int32_t *)ctx->xv.p, (const float *)ctx->win_full.p, d_win_tail,
                     (FbAnalysis *)ctx->ana.p, A.analyze_only ? (fb200_variant_taps *)ctx->taps.p : nullptr, nvars, st);
        FB_CUDA(ctx, cudaEventRecord(ctx->ev[4], st));
        launches += 2;
        if (A.analyze_only) {
            FB_CUDA(ctx, cudaMemcpyAsync(A.taps + f0 * (uint64_t)nvar, ctx->taps.p,
                                         (size_t)nvars * sizeof(fb200_variant_taps), cudaMemcpyDeviceToHost, st));
            FB_CUDA(ctx, cudaStreamSynchronize(st));
            continue;
        }
        fb200_frame_info *d_infos = A.infos ? (fb200_frame_info *)ctx->infos.p : nullptr;
        if (fused) {
            // KF: Rice search + frame assembly per frame; frames it cannot reproduce exactly go to the list
            FB_CUDA(ctx, cudaMemsetAsync(d_fb_count, 0, 4, st));
            fb_launch_kf(ring, J, (const int32_t *)ctx->xv.p, (const FbAnalysis *)ctx->ana.p, (uint8_t *)ctx->slots.p, d_fb,
                         d_infos, (uint32_t *)ctx->fb_list.p, d_fb_count, (const uint32_t *)ctx->ktab.p, KL, st);
            FB_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
            fb_launch_k2(ring, J, (const int32_t *)ctx->xv.p, (const FbAnalysis *)ctx->ana.p,
                         (fb200_subframe_info *)ctx->choice.p, L, (const uint32_t *)ctx->fb_list.p, d_fb_count, 296,
                         k2_smem, st);
            fb_launch_k3(ring, J, (const int32_t *)ctx->xv.p, (const fb200_subframe_info *)ctx->choice.p,
                         (uint8_t *)ctx->slots.p, d_fb, d_infos, (const uint32_t *)ctx->fb_list.p, d_fb_count, 148,
                         k3_smem, st);
            FB_CUDA(ctx, cudaEventRecord(ctx->ev[6], st));
            if (n_chunks < max_counts)
                FB_CUDA(ctx, cudaMemcpyAsync(&h_fb_counts[n_chunks], d_fb_count, 4, cudaMemcpyDeviceToHost, st));
            n_chunks++;
            fused_frames += J.n_frames;
            launches += 3;
        } else {
            fb_launch_k2(ring, J, (const int32_t *)ctx->xv.p, (const FbAnalysis *)ctx->ana.p,
                         (fb200_subframe_info *)ctx->choice.p, L, nullptr, nullptr, nvars, k2_smem, st);
            FB_CUDA(ctx, cudaEventRecord(ctx->ev[5], st));
            fb_launch_k3(ring, J, (const int32_t *)ctx->xv.p, (const fb200_subframe_info *)ctx->choice.p,
                         (uint8_t *)ctx->slots.p, d_fb, d_infos, nullptr, nullptr, J.n_frames, k3_smem, st);
            FB_CUDA(ctx, cudaEventRecord(ctx->ev[6], st));
        }
        // K4
        fb_k4_scan<<<1, FB_K4_THREADS, 0, st>>>(d_fb, (unsigned long long *)ctx->offsets.p, J.n_frames, d_total);
        fb_k4_gather<<<J.n_frames, 256, 0, st>>>((const uint8_t *)ctx->slots.p, J.slot_bytes, d_fb,
                                                 (const unsigned long long *)ctx->offsets.p, d_out,
                                                 (unsigned long long)A.out_cap);
        FB_CUDA(ctx, cudaEventRecord(ctx->ev[7], st));
        launches += fused ? 2 : 4;
        FB_CUDA(ctx, cudaGetLastError());
        if (A.infos) {
            FB_CUDA(ctx, cudaMemcpyAsync(A.infos + f0, ctx->infos.p, (size_t)J.n_frames * sizeof(fb200_frame_info),
                                         cudaMemcpyDeviceToHost, st));
        }
        // per-chunk timings need the events to have completed; chunks are serial on one stream
        FB_CUDA(ctx, cudaEventSynchronize(ctx->ev[7]));
        float t;
        FB_CUDA(ctx, cudaEventElapsedTime(&t, ctx->ev[1], ctx->ev[2])); ms_h2d += t;
        for (int k = 0; k < 5; k++) {
            FB_CUDA(ctx, cudaEventElapsedTime(&t, ctx->ev[2 + k], ctx->ev[3 + k]));
            ms_k[k] += t;
        }
    }
    if (A.analyze_only) return FB200_OK;

    // results: error flag + total, sizes, bytes
    FB_CUDA(ctx, cudaEventRecord(ctx->ev[8], st));
    FB_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, ctx->scalars.p, 16, cudaMemcpyDeviceToHost, st));
    FB_CUDA(ctx, cudaStreamSynchronize(st));
    const uint32_t err_flag = *(const uint32_t *)ctx->pinned;
    const unsigned long long total = *(const unsigned long long *)((const uint8_t *)ctx->pinned + 8);
    if (err_flag) {
        ctx->last_error = "input sample out of the range of bits_per_sample";
        return FB200_ERR_CONFIG; // VerifyError (src/source.rs:262-275)
    }
    if (A.out_len) *A.out_len = (size_t)total;
    if (total > A.out_cap) {
        ctx->last_error = "output capacity too small";
        return FB200_ERR_CAPACITY;
    }
    if (A.frame_sizes)
        FB_CUDA(ctx, cudaMemcpyAsync(A.frame_sizes, ctx->frame_bytes.p, total_frames * 4u, cudaMemcpyDeviceToHost, st));
    if (A.out_host) FB_CUDA(ctx, cudaMemcpyAsync(A.out_host, d_out, (size_t)total, cudaMemcpyDeviceToHost, st));
    FB_CUDA(ctx, cudaEventRecord(ctx->ev[9], st));
    FB_CUDA(ctx, cudaStreamSynchronize(st));

    float t_d2h = 0, t_total = 0;
    FB_CUDA(ctx, cudaEventElapsedTime(&t_d2h, ctx->ev[8], ctx->ev[9]));
    FB_CUDA(ctx, cudaEventElapsedTime(&t_total, ctx->ev[0], ctx->ev[9]));
    ctx->timing.h2d_ms = ms_h2d;
    ctx->timing.k_ingest_ms = ms_k[0];
    ctx->timing.k_analyze_ms = ms_k[1];
    ctx->timing.k_rice_ms = ms_k[2];
    ctx->timing.k_pack_ms = ms_k[3];
    ctx->timing.k_gather_ms = ms_k[4];
    ctx->timing.kernels_ms = ms_k[0] + ms_k[1] + ms_k[2] + ms_k[3] + ms_k[4];
    ctx->timing.d2h_ms = t_d2h;
    ctx->timing.total_ms = t_total;
    ctx->timing.launches = launches;
    uint64_t fallback = 0;
    for (uint64_t i = 0; i < n_chunks && i < max_counts; i++) fallback += h_fb_counts[i];
    ctx->timing.fused_frames = fused_frames - fallback;
    ctx->timing.fallback_frames = fallback;
    ctx->timing.in_bytes = in_bytes_total;
    ctx->timing.out_bytes = total;
    return FB200_OK;
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

// encode_with_fixed_block_size (src/coding.rs:645-695) with par.rs-style sharding: contiguous frame
// ranges over the devices (one host thread + context per device), MD5 on its own host thread
// (src/par.rs:196-277), STREAMINFO finalised on the calling thread.
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
    const int nd = (int)std::min<uint64_t>((uint64_t)n_devices, std::max<uint64_t>(n_frames, 1));

    // MD5 over the packed little-endian samples of ceil(bps/8) bytes (src/source.rs:406-429)
    uint8_t md5[16];
    std::thread md5_thread([&] {
        FbMd5 h;
        const int bytes_per_sample = (bits_per_sample + 7) / 8;
        const uint64_t count = n_samples * (uint64_t)channels;
        if (bytes_per_sample == container_bytes) {
            h.update((const uint8_t *)pcm, (size_t)(count * (uint64_t)container_bytes));
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
    });

    struct Shard {
        uint64_t f0 = 0, nf = 0;
        std::vector<uint8_t> bytes;
        std::vector<uint32_t> sizes;
        size_t len = 0;
        int rc = FB200_OK;
    };
    std::vector<Shard> shards((size_t)nd);
    std::vector<std::thread> workers;
    const size_t max_frame_bytes = fb_max_frame_bytes(channels, bits_per_sample, block_size);
    for (int d = 0; d < nd; d++) {
        Shard &S = shards[(size_t)d];
        S.f0 = (n_frames * (uint64_t)d + (uint64_t)nd - 1) / (uint64_t)nd; // ceil(F*g/G), SURVEY.md 8(e)
        uint64_t f1 = (n_frames * (uint64_t)(d + 1) + (uint64_t)nd - 1) / (uint64_t)nd;
        S.nf = f1 - S.f0;
        if (S.nf == 0) continue;
        workers.emplace_back([&, d] {
            Shard &S2 = shards[(size_t)d];
            int err = 0;
            fb200_ctx *ctx = fb200_create(cfg, channels, bits_per_sample, sample_rate, block_size, devices[d], &err);
            if (!ctx) { S2.rc = err; return; }
            const uint64_t s0 = S2.f0 * bs;
            const uint64_t ns = std::min(n_samples - s0, S2.nf * bs);
            S2.bytes.resize((size_t)(S2.nf * max_frame_bytes));
            S2.sizes.resize((size_t)S2.nf);
            size_t nf_out = 0;
            S2.rc = fb200_encode_interleaved(
                ctx, (const uint8_t *)pcm + s0 * (uint64_t)channels * (uint64_t)container_bytes, container_bytes, ns,
                S2.f0, S2.bytes.data(), S2.bytes.size(), S2.sizes.data(), nullptr, &nf_out, &S2.len);
            fb200_destroy(ctx);
        });
    }
    for (auto &w : workers) w.join();
    md5_thread.join();
    for (auto &S : shards)
        if (S.rc) return S.rc;

    // STREAMINFO (src/component/datatype.rs:514-523, src/coding.rs:676-693, src/component/bitrepr.rs:240-267)
    uint32_t min_block = 0xFFFF, max_block = 0, min_frame = 0xFFFFFFFFu, max_frame = 0;
    size_t total = 42;
    for (auto &S : shards) {
        for (uint64_t i = 0; i < S.nf; i++) {
            uint64_t fi = S.f0 + i;
            uint32_t b = (uint32_t)std::min<uint64_t>(bs, n_samples - fi * bs);
            min_block = std::min(min_block, b);
            max_block = std::max(max_block, b);
            min_frame = std::min(min_frame, S.sizes[(size_t)i]);
            max_frame = std::max(max_frame, S.sizes[(size_t)i]);
        }
        total += S.len;
    }
    if (n_frames > 0) min_block = max_block;
    if (out_len) *out_len = total;
    if (total > out_cap) return FB200_ERR_CAPACITY;
    uint8_t *p = out;
    memcpy(p, "fLaC", 4);
    p[4] = 0x80; p[5] = 0; p[6] = 0; p[7] = 34;
    p[8] = (uint8_t)(min_block >> 8); p[9] = (uint8_t)min_block;
    p[10] = (uint8_t)(max_block >> 8); p[11] = (uint8_t)max_block;
    p[12] = (uint8_t)(min_frame >> 16); p[13] = (uint8_t)(min_frame >> 8); p[14] = (uint8_t)min_frame;
    p[15] = (uint8_t)(max_frame >> 16); p[16] = (uint8_t)(max_frame >> 8); p[17] = (uint8_t)max_frame;
    // 20 bits rate | 3 bits channels-1 | 5 bits bps-1 | 36 bits total samples
    uint64_t v = ((uint64_t)(uint32_t)sample_rate << 44) | ((uint64_t)(channels - 1) << 41) |
                 ((uint64_t)(bits_per_sample - 1) << 36) | (n_samples & 0xFFFFFFFFFull);
    for (int i = 0; i < 8; i++) p[18 + i] = (uint8_t)(v >> (56 - 8 * i));
    memcpy(p + 26, md5, 16);
    size_t o = 42;
    for (auto &S : shards) {
        if (S.len) memcpy(out + o, S.bytes.data(), S.len);
        o += S.len;
    }
    return FB200_OK;
}

} // extern "C"
