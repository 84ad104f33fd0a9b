/*
 * flacenc_decoder.c -- independent FLAC decoder + MD5 used as the tier-1 ("decodes bit-exactly")
 * check.  TEST INFRASTRUCTURE ONLY (see flacenc_oracle.h).
 *
 * Written from the FLAC format definition (frame header, subframe types, partitioned Rice
 * residual, inter-channel decorrelation, CRC-8/CRC-16), deliberately NOT from the reference's
 * src/component/parser.rs / decode.rs, so that an encoder-side misreading of the format cannot
 * cancel out.  It plays the role claxon / `flac -d` play in the reference's own integration tests
 * (src/test_helper.rs:131-185, pytools/reporter.py:123-148).  It supports more than the encoder
 * emits (wasted bits, escaped partitions, variable blocking).
 */
#include "flacenc_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- MD5 (RFC 1321) ---------- */

typedef struct {
    uint32_t a, b, c, d;
    uint64_t len;
    uint8_t buf[64];
    size_t fill;
} fo_md5_ctx;

static const uint32_t MD5_K[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
    0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
    0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
    0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
    0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
    0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
    0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
    0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
static const uint8_t MD5_S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22,
                                  5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                                  4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                  6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};

static void md5_block(fo_md5_ctx *c, const uint8_t *p) {
    uint32_t m[16];
    for (int i = 0; i < 16; i++)
        m[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) |
               ((uint32_t)p[4 * i + 3] << 24);
    uint32_t a = c->a, b = c->b, cc = c->c, d = c->d;
    for (int i = 0; i < 64; i++) {
        uint32_t f;
        int g;
        if (i < 16) {
            f = (b & cc) | (~b & d);
            g = i;
        } else if (i < 32) {
            f = (d & b) | (~d & cc);
            g = (5 * i + 1) & 15;
        } else if (i < 48) {
            f = b ^ cc ^ d;
            g = (3 * i + 5) & 15;
        } else {
            f = cc ^ (b | ~d);
            g = (7 * i) & 15;
        }
        uint32_t tmp = d;
        d = cc;
        cc = b;
        uint32_t x = a + f + MD5_K[i] + m[g];
        b = b + ((x << MD5_S[i]) | (x >> (32 - MD5_S[i])));
        a = tmp;
    }
    c->a += a;
    c->b += b;
    c->c += cc;
    c->d += d;
}

static void md5_init(fo_md5_ctx *c) {
    c->a = 0x67452301;
    c->b = 0xefcdab89;
    c->c = 0x98badcfe;
    c->d = 0x10325476;
    c->len = 0;
    c->fill = 0;
}

static void md5_update(fo_md5_ctx *c, const uint8_t *data, size_t len) {
    c->len += len;
    if (c->fill) {
        size_t take = 64 - c->fill;
        if (take > len) take = len;
        memcpy(c->buf + c->fill, data, take);
        c->fill += take;
        data += take;
        len -= take;
        if (c->fill == 64) {
            md5_block(c, c->buf);
            c->fill = 0;
        }
    }
    while (len >= 64) {
        md5_block(c, data);
        data += 64;
        len -= 64;
    }
    if (len) {
        memcpy(c->buf, data, len);
        c->fill = len;
    }
}

static void md5_final(fo_md5_ctx *c, uint8_t digest[16]) {
    uint64_t bits = c->len * 8;
    uint8_t pad[72];
    size_t padlen = (c->fill < 56) ? (56 - c->fill) : (120 - c->fill);
    memset(pad, 0, sizeof(pad));
    pad[0] = 0x80;
    md5_update(c, pad, padlen);
    uint8_t lenb[8];
    for (int i = 0; i < 8; i++) lenb[i] = (uint8_t)(bits >> (8 * i));
    md5_update(c, lenb, 8);
    uint32_t v[4] = {c->a, c->b, c->c, c->d};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) digest[4 * i + j] = (uint8_t)(v[i] >> (8 * j));
}

void fo_md5(const uint8_t *data, size_t len, uint8_t digest[16]) {
    fo_md5_ctx c;
    md5_init(&c);
    md5_update(&c, data, len);
    md5_final(&c, digest);
}

/* src/source.rs:406-418  Context::fill_interleaved: little-endian, ceil(bps/8) bytes per sample */
void fo_md5_of_samples(const int32_t *interleaved, size_t count, int bytes_per_sample, uint8_t digest[16]) {
    fo_md5_ctx c;
    md5_init(&c);
    uint8_t tmp[4096];
    size_t k = 0;
    for (size_t i = 0; i < count; i++) {
        uint32_t v = (uint32_t)interleaved[i];
        for (int b = 0; b < bytes_per_sample; b++) tmp[k++] = (uint8_t)(v >> (8 * b));
        if (k + 4 > sizeof(tmp)) {
            md5_update(&c, tmp, k);
            k = 0;
        }
    }
    if (k) md5_update(&c, tmp, k);
    md5_final(&c, digest);
}

/* ---------------------------------------------------------------- bit reader -------------- */

typedef struct {
    const uint8_t *p;
    size_t len;
    uint64_t pos; /* in bits */
    int err;
} br_t;

static uint32_t br_bit(br_t *b) {
    uint64_t byte = b->pos >> 3;
    if (byte >= b->len) {
        b->err = 1;
        return 0;
    }
    uint32_t v = (b->p[byte] >> (7 - (b->pos & 7))) & 1u;
    b->pos++;
    return v;
}

static uint64_t br_bits(br_t *b, int n) {
    uint64_t v = 0;
    for (int i = 0; i < n; i++) v = (v << 1) | br_bit(b);
    return v;
}

static int64_t br_sbits(br_t *b, int n) {
    uint64_t v = br_bits(b, n);
    if (n < 64 && (v >> (n - 1)) & 1) v |= ~0ull << n;
    return (int64_t)v;
}

static uint64_t br_unary(br_t *b) { /* number of 0 bits before the next 1 */
    uint64_t q = 0;
    while (!b->err && br_bit(b) == 0) q++;
    return q;
}

/* ---------------------------------------------------------------- frame decoding ---------- */

static int decode_residual(br_t *b, int n, int pred_order, int64_t *res) {
    int method = (int)br_bits(b, 2);
    if (method > 1) return -5;
    int pbits = method ? 5 : 4;
    int escape = method ? 31 : 15;
    int porder = (int)br_bits(b, 4);
    int nparts = 1 << porder;
    if ((n % nparts) != 0 && porder > 0) return -5;
    int plen = n >> porder;
    int t = pred_order;
    for (int p = 0; p < nparts; p++) {
        int param = (int)br_bits(b, pbits);
        int count = plen - (p == 0 ? pred_order : 0);
        if (count < 0) return -5;
        if (param == escape) {
            int raw = (int)br_bits(b, 5);
            for (int i = 0; i < count; i++) res[t++] = raw ? br_sbits(b, raw) : 0;
        } else {
            for (int i = 0; i < count; i++) {
                uint64_t q = br_unary(b);
                uint64_t r = param ? br_bits(b, param) : 0;
                uint64_t u = (q << param) | r;
                res[t++] = (u & 1) ? -(int64_t)((u >> 1) + 1) : (int64_t)(u >> 1);
                if (b->err) return -5;
            }
        }
    }
    return b->err ? -5 : 0;
}

static int decode_subframe(br_t *b, int n, int bps, int64_t *out, int64_t *res) {
    if (br_bit(b)) return -5; /* padding bit */
    int type = (int)br_bits(b, 6);
    int wasted = 0;
    if (br_bit(b)) wasted = 1 + (int)br_unary(b);
    bps -= wasted;
    if (bps <= 0) return -5;
    if (type == 0) {
        int64_t v = br_sbits(b, bps);
        for (int t = 0; t < n; t++) out[t] = v;
    } else if (type == 1) {
        for (int t = 0; t < n; t++) out[t] = br_sbits(b, bps);
    } else if (type >= 8 && type <= 12) {
        int order = type - 8;
        if (order > n) return -5;
        for (int t = 0; t < order; t++) out[t] = br_sbits(b, bps);
        int rc = decode_residual(b, n, order, res);
        if (rc) return rc;
        for (int t = order; t < n; t++) {
            int64_t pred = 0;
            switch (order) {
            case 1: pred = out[t - 1]; break;
            case 2: pred = 2 * out[t - 1] - out[t - 2]; break;
            case 3: pred = 3 * out[t - 1] - 3 * out[t - 2] + out[t - 3]; break;
            case 4: pred = 4 * out[t - 1] - 6 * out[t - 2] + 4 * out[t - 3] - out[t - 4]; break;
            default: break;
            }
            /* FLAC samples are <= 32 bits; wrap like a 32-bit decoder so wrapped encoder residuals invert */
            out[t] = (int64_t)(int32_t)(uint32_t)(uint64_t)(pred + res[t]);
        }
    } else if (type >= 32) {
        int order = type - 31;
        if (order > n) return -5;
        for (int t = 0; t < order; t++) out[t] = br_sbits(b, bps);
        int prec = (int)br_bits(b, 4) + 1;
        if (prec == 16) return -5;
        int shift = (int)br_sbits(b, 5);
        if (shift < 0) return -5;
        int64_t coef[32];
        for (int j = 0; j < order; j++) coef[j] = br_sbits(b, prec);
        int rc = decode_residual(b, n, order, res);
        if (rc) return rc;
        for (int t = order; t < n; t++) {
            int64_t acc = 0;
            for (int j = 0; j < order; j++) acc += coef[j] * out[t - 1 - j];
            out[t] = (int64_t)(int32_t)(uint32_t)(uint64_t)((acc >> shift) + res[t]);
        }
    } else {
        return -5;
    }
    if (wasted)
        for (int t = 0; t < n; t++) out[t] = (int64_t)((uint64_t)out[t] << wasted);
    return b->err ? -5 : 0;
}

static int64_t read_utf8like(br_t *b) {
    uint32_t first = (uint32_t)br_bits(b, 8);
    int extra;
    uint64_t v;
    if (first < 0x80) return first;
    if ((first & 0xE0) == 0xC0) { extra = 1; v = first & 0x1F; }
    else if ((first & 0xF0) == 0xE0) { extra = 2; v = first & 0x0F; }
    else if ((first & 0xF8) == 0xF0) { extra = 3; v = first & 0x07; }
    else if ((first & 0xFC) == 0xF8) { extra = 4; v = first & 0x03; }
    else if ((first & 0xFE) == 0xFC) { extra = 5; v = first & 0x01; }
    else if (first == 0xFE) { extra = 6; v = 0; }
    else return -1;
    for (int i = 0; i < extra; i++) {
        uint32_t c = (uint32_t)br_bits(b, 8);
        if ((c & 0xC0) != 0x80) return -1;
        v = (v << 6) | (c & 0x3F);
    }
    return (int64_t)v;
}

/* decodes one frame starting at byte offset *off; appends n samples/ch to out (if not NULL) */
static int decode_frame(const uint8_t *data, size_t len, size_t *off, int stream_channels, int stream_bps,
                        int32_t *out, uint64_t out_pos, uint64_t cap, int *n_out, uint64_t *number_out,
                        int *variable_out) {
    br_t b = {data + *off, len - *off, 0, 0};
    uint32_t sync = (uint32_t)br_bits(&b, 14);
    if (sync != 0x3FFE) return -2;
    if (br_bit(&b)) return -2;
    int variable = (int)br_bit(&b);
    int bs_code = (int)br_bits(&b, 4);
    int sr_code = (int)br_bits(&b, 4);
    int ch_code = (int)br_bits(&b, 4);
    int ss_code = (int)br_bits(&b, 3);
    if (br_bit(&b)) return -2;
    int64_t number = read_utf8like(&b);
    if (number < 0) return -2;
    int n;
    if (bs_code == 0) return -2;
    else if (bs_code == 1) n = 192;
    else if (bs_code <= 5) n = 576 << (bs_code - 2);
    else if (bs_code == 6) n = (int)br_bits(&b, 8) + 1;
    else if (bs_code == 7) n = (int)br_bits(&b, 16) + 1;
    else n = 256 << (bs_code - 8);
    if (sr_code == 12) (void)br_bits(&b, 8);
    else if (sr_code == 13 || sr_code == 14) (void)br_bits(&b, 16);
    else if (sr_code == 15) return -2;
    size_t hdr_bytes = (size_t)(b.pos >> 3);
    uint32_t crc8 = (uint32_t)br_bits(&b, 8);
    if (b.err) return -2;
    if (fo_crc8(b.p, hdr_bytes) != crc8) return -3;
    int bps;
    switch (ss_code) {
    case 0: bps = stream_bps; break;
    case 1: bps = 8; break;
    case 2: bps = 12; break;
    case 4: bps = 16; break;
    case 5: bps = 20; break;
    case 6: bps = 24; break;
    case 7: bps = 32; break;
    default: return -2;
    }
    int channels = ch_code < 8 ? ch_code + 1 : 2;
    if (ch_code > 10) return -2;
    if (stream_channels && channels != stream_channels) return -2;
    int64_t *chan = (int64_t *)malloc(sizeof(int64_t) * (size_t)n * (size_t)(channels + 1));
    int64_t *res = chan + (size_t)n * (size_t)channels;
    int rc = 0;
    for (int ch = 0; ch < channels && !rc; ch++) {
        int sbps = bps;
        if ((ch_code == 8 && ch == 1) || (ch_code == 9 && ch == 0) || (ch_code == 10 && ch == 1)) sbps++;
        rc = decode_subframe(&b, n, sbps, chan + (size_t)ch * n, res);
    }
    if (!rc) {
        b.pos = (b.pos + 7) & ~7ull;
        size_t body_bytes = (size_t)(b.pos >> 3);
        uint32_t crc16 = (uint32_t)br_bits(&b, 16);
        if (b.err) rc = -2;
        else if (fo_crc16(b.p, body_bytes) != crc16) rc = -4;
    }
    if (!rc) {
        int64_t *c0 = chan, *c1 = chan + n;
        if (ch_code == 8) {
            for (int t = 0; t < n; t++) c1[t] = c0[t] - c1[t];
        } else if (ch_code == 9) {
            for (int t = 0; t < n; t++) c0[t] = c0[t] + c1[t];
        } else if (ch_code == 10) {
            for (int t = 0; t < n; t++) {
                int64_t mid = c0[t], side = c1[t];
                mid = (int64_t)((uint64_t)mid << 1) | (side & 1);
                c0[t] = (mid + side) >> 1;
                c1[t] = (mid - side) >> 1;
            }
        }
        if (out) {
            if (out_pos + (uint64_t)n > cap) rc = -6;
            else
                for (int t = 0; t < n; t++)
                    for (int ch = 0; ch < channels; ch++)
                        out[(out_pos + (uint64_t)t) * (uint64_t)channels + (uint64_t)ch] = (int32_t)chan[(size_t)ch * n + t];
        }
    }
    free(chan);
    if (rc) return rc;
    *off += (size_t)(b.pos >> 3);
    *n_out = n;
    if (number_out) *number_out = (uint64_t)number;
    if (variable_out) *variable_out = variable;
    return 0;
}

int64_t fo_decode_frames(const uint8_t *data, size_t len, int channels, int bps, int32_t *out,
                         uint64_t cap, uint64_t *n_frames) {
    size_t off = 0;
    uint64_t pos = 0, frames = 0;
    while (off < len) {
        int n = 0;
        int rc = decode_frame(data, len, &off, channels, bps, out, pos, cap, &n, NULL, NULL);
        if (rc) return rc;
        pos += (uint64_t)n;
        frames++;
    }
    if (n_frames) *n_frames = frames;
    return (int64_t)pos;
}

int64_t fo_decode_stream(const uint8_t *data, size_t len, fo_stream_info *info, int32_t *out, uint64_t cap) {
    if (len < 42 || memcmp(data, "fLaC", 4) != 0) return -1;
    size_t off = 4;
    int last = 0, have_info = 0;
    fo_stream_info si;
    memset(&si, 0, sizeof(si));
    while (!last) {
        if (off + 4 > len) return -1;
        last = data[off] >> 7;
        int type = data[off] & 0x7F;
        size_t blen = ((size_t)data[off + 1] << 16) | ((size_t)data[off + 2] << 8) | data[off + 3];
        off += 4;
        if (off + blen > len) return -1;
        if (type == 0) {
            if (blen != 34) return -1;
            br_t b = {data + off, blen, 0, 0};
            si.min_block = (uint32_t)br_bits(&b, 16);
            si.max_block = (uint32_t)br_bits(&b, 16);
            si.min_frame = (uint32_t)br_bits(&b, 24);
            si.max_frame = (uint32_t)br_bits(&b, 24);
            si.sample_rate = (uint32_t)br_bits(&b, 20);
            si.channels = (uint32_t)br_bits(&b, 3) + 1;
            si.bps = (uint32_t)br_bits(&b, 5) + 1;
            si.total_samples = br_bits(&b, 36);
            memcpy(si.md5, data + off + 18, 16);
            have_info = 1;
        }
        off += blen;
    }
    if (!have_info) return -1;
    uint64_t frames = 0;
    int64_t n = fo_decode_frames(data + off, len - off, (int)si.channels, (int)si.bps, out, cap, &frames);
    si.n_frames = frames;
    if (info) *info = si;
    return n;
}
