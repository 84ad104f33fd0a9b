// fb_md5.h -- MD5 (RFC 1321) for the STREAMINFO signature.  Host-side stream assembly only
// (the reference hashes the packed little-endian samples on a dedicated thread, src/par.rs:196-277,
// src/source.rs:406-429); MD5 is sequential per stream, so it stays on a host core.
#pragma once

#include <stddef.h>
#include <stdint.h>
#include <string.h>

class FbMd5 {
  public:
    FbMd5() : a_(0x67452301u), b_(0xefcdab89u), c_(0x98badcfeu), d_(0x10325476u), len_(0), fill_(0) {}

    void update(const uint8_t *data, size_t len) {
        len_ += len;
        if (fill_) {
            size_t take = 64 - fill_;
            if (take > len) take = len;
            memcpy(buf_ + fill_, data, take);
            fill_ += take;
            data += take;
            len -= take;
            if (fill_ == 64) { block(buf_); fill_ = 0; }
        }
        for (; len >= 64; data += 64, len -= 64) block(data);
        if (len) { memcpy(buf_, data, len); fill_ = len; }
    }

    // int32 samples hashed as their packed 16-bit little-endian form (what Context::fill_interleaved feeds the hash for
    // 16-bit streams, src/source.rs:406-418) without a packed copy: the 16 message words of a block are assembled from 32
    // samples as they are needed
    void update_i32_as_le16(const int32_t *x, size_t count) {
        while (count && fill_) { // finish a partial block byte by byte
            const uint16_t v = (uint16_t)*x++;
            const uint8_t b[2] = {(uint8_t)v, (uint8_t)(v >> 8)};
            update(b, 2);
            count--;
        }
        for (; count >= 32; x += 32, count -= 32) {
            uint32_t m[16];
            for (int k = 0; k < 16; k++) m[k] = ((uint32_t)x[2 * k] & 0xFFFFu) | ((uint32_t)x[2 * k + 1] << 16);
            len_ += 64;
            block((const uint8_t *)m);
        }
        for (; count; count--) {
            const uint16_t v = (uint16_t)*x++;
            const uint8_t b[2] = {(uint8_t)v, (uint8_t)(v >> 8)};
            update(b, 2);
        }
    }

    void finish(uint8_t digest[16]) {
        const uint64_t bits = len_ * 8;
        uint8_t pad[72] = {0x80};
        size_t padlen = (fill_ < 56) ? (56 - fill_) : (120 - fill_);
        update(pad, padlen);
        uint8_t lenb[8];
        for (int i = 0; i < 8; i++) lenb[i] = (uint8_t)(bits >> (8 * i));
        update(lenb, 8);
        const uint32_t v[4] = {a_, b_, c_, d_};
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) digest[4 * i + j] = (uint8_t)(v[i] >> (8 * j));
    }

  private:
    static uint32_t rol(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }

    void block(const uint8_t *p) {
        uint32_t m[16];
        memcpy(m, p, 64); // little-endian host (x86-64)
        uint32_t a = a_, b = b_, c = c_, d = d_;
#define FB_MD5_F(x, y, z) ((z) ^ ((x) & ((y) ^ (z))))
#define FB_MD5_G(x, y, z) ((y) ^ ((z) & ((x) ^ (y))))
#define FB_MD5_H(x, y, z) ((x) ^ (y) ^ (z))
#define FB_MD5_I(x, y, z) ((y) ^ ((x) | ~(z)))
#define FB_MD5_STEP(f, w, x, y, z, k, t, s) w = x + rol(w + f(x, y, z) + m[k] + t, s)
        FB_MD5_STEP(FB_MD5_F, a, b, c, d, 0, 0xd76aa478u, 7);   FB_MD5_STEP(FB_MD5_F, d, a, b, c, 1, 0xe8c7b756u, 12);
        FB_MD5_STEP(FB_MD5_F, c, d, a, b, 2, 0x242070dbu, 17);  FB_MD5_STEP(FB_MD5_F, b, c, d, a, 3, 0xc1bdceeeu, 22);
        FB_MD5_STEP(FB_MD5_F, a, b, c, d, 4, 0xf57c0fafu, 7);   FB_MD5_STEP(FB_MD5_F, d, a, b, c, 5, 0x4787c62au, 12);
        FB_MD5_STEP(FB_MD5_F, c, d, a, b, 6, 0xa8304613u, 17);  FB_MD5_STEP(FB_MD5_F, b, c, d, a, 7, 0xfd469501u, 22);
        FB_MD5_STEP(FB_MD5_F, a, b, c, d, 8, 0x698098d8u, 7);   FB_MD5_STEP(FB_MD5_F, d, a, b, c, 9, 0x8b44f7afu, 12);
        FB_MD5_STEP(FB_MD5_F, c, d, a, b, 10, 0xffff5bb1u, 17); FB_MD5_STEP(FB_MD5_F, b, c, d, a, 11, 0x895cd7beu, 22);
        FB_MD5_STEP(FB_MD5_F, a, b, c, d, 12, 0x6b901122u, 7);  FB_MD5_STEP(FB_MD5_F, d, a, b, c, 13, 0xfd987193u, 12);
        FB_MD5_STEP(FB_MD5_F, c, d, a, b, 14, 0xa679438eu, 17); FB_MD5_STEP(FB_MD5_F, b, c, d, a, 15, 0x49b40821u, 22);
        FB_MD5_STEP(FB_MD5_G, a, b, c, d, 1, 0xf61e2562u, 5);   FB_MD5_STEP(FB_MD5_G, d, a, b, c, 6, 0xc040b340u, 9);
        FB_MD5_STEP(FB_MD5_G, c, d, a, b, 11, 0x265e5a51u, 14); FB_MD5_STEP(FB_MD5_G, b, c, d, a, 0, 0xe9b6c7aau, 20);
        FB_MD5_STEP(FB_MD5_G, a, b, c, d, 5, 0xd62f105du, 5);   FB_MD5_STEP(FB_MD5_G, d, a, b, c, 10, 0x02441453u, 9);
        FB_MD5_STEP(FB_MD5_G, c, d, a, b, 15, 0xd8a1e681u, 14); FB_MD5_STEP(FB_MD5_G, b, c, d, a, 4, 0xe7d3fbc8u, 20);
        FB_MD5_STEP(FB_MD5_G, a, b, c, d, 9, 0x21e1cde6u, 5);   FB_MD5_STEP(FB_MD5_G, d, a, b, c, 14, 0xc33707d6u, 9);
        FB_MD5_STEP(FB_MD5_G, c, d, a, b, 3, 0xf4d50d87u, 14);  FB_MD5_STEP(FB_MD5_G, b, c, d, a, 8, 0x455a14edu, 20);
        FB_MD5_STEP(FB_MD5_G, a, b, c, d, 13, 0xa9e3e905u, 5);  FB_MD5_STEP(FB_MD5_G, d, a, b, c, 2, 0xfcefa3f8u, 9);
        FB_MD5_STEP(FB_MD5_G, c, d, a, b, 7, 0x676f02d9u, 14);  FB_MD5_STEP(FB_MD5_G, b, c, d, a, 12, 0x8d2a4c8au, 20);
        FB_MD5_STEP(FB_MD5_H, a, b, c, d, 5, 0xfffa3942u, 4);   FB_MD5_STEP(FB_MD5_H, d, a, b, c, 8, 0x8771f681u, 11);
        FB_MD5_STEP(FB_MD5_H, c, d, a, b, 11, 0x6d9d6122u, 16); FB_MD5_STEP(FB_MD5_H, b, c, d, a, 14, 0xfde5380cu, 23);
        FB_MD5_STEP(FB_MD5_H, a, b, c, d, 1, 0xa4beea44u, 4);   FB_MD5_STEP(FB_MD5_H, d, a, b, c, 4, 0x4bdecfa9u, 11);
        FB_MD5_STEP(FB_MD5_H, c, d, a, b, 7, 0xf6bb4b60u, 16);  FB_MD5_STEP(FB_MD5_H, b, c, d, a, 10, 0xbebfbc70u, 23);
        FB_MD5_STEP(FB_MD5_H, a, b, c, d, 13, 0x289b7ec6u, 4);  FB_MD5_STEP(FB_MD5_H, d, a, b, c, 0, 0xeaa127fau, 11);
        FB_MD5_STEP(FB_MD5_H, c, d, a, b, 3, 0xd4ef3085u, 16);  FB_MD5_STEP(FB_MD5_H, b, c, d, a, 6, 0x04881d05u, 23);
        FB_MD5_STEP(FB_MD5_H, a, b, c, d, 9, 0xd9d4d039u, 4);   FB_MD5_STEP(FB_MD5_H, d, a, b, c, 12, 0xe6db99e5u, 11);
        FB_MD5_STEP(FB_MD5_H, c, d, a, b, 15, 0x1fa27cf8u, 16); FB_MD5_STEP(FB_MD5_H, b, c, d, a, 2, 0xc4ac5665u, 23);
        FB_MD5_STEP(FB_MD5_I, a, b, c, d, 0, 0xf4292244u, 6);   FB_MD5_STEP(FB_MD5_I, d, a, b, c, 7, 0x432aff97u, 10);
        FB_MD5_STEP(FB_MD5_I, c, d, a, b, 14, 0xab9423a7u, 15); FB_MD5_STEP(FB_MD5_I, b, c, d, a, 5, 0xfc93a039u, 21);
        FB_MD5_STEP(FB_MD5_I, a, b, c, d, 12, 0x655b59c3u, 6);  FB_MD5_STEP(FB_MD5_I, d, a, b, c, 3, 0x8f0ccc92u, 10);
        FB_MD5_STEP(FB_MD5_I, c, d, a, b, 10, 0xffeff47du, 15); FB_MD5_STEP(FB_MD5_I, b, c, d, a, 1, 0x85845dd1u, 21);
        FB_MD5_STEP(FB_MD5_I, a, b, c, d, 8, 0x6fa87e4fu, 6);   FB_MD5_STEP(FB_MD5_I, d, a, b, c, 15, 0xfe2ce6e0u, 10);
        FB_MD5_STEP(FB_MD5_I, c, d, a, b, 6, 0xa3014314u, 15);  FB_MD5_STEP(FB_MD5_I, b, c, d, a, 13, 0x4e0811a1u, 21);
        FB_MD5_STEP(FB_MD5_I, a, b, c, d, 4, 0xf7537e82u, 6);   FB_MD5_STEP(FB_MD5_I, d, a, b, c, 11, 0xbd3af235u, 10);
        FB_MD5_STEP(FB_MD5_I, c, d, a, b, 2, 0x2ad7d2bbu, 15);  FB_MD5_STEP(FB_MD5_I, b, c, d, a, 9, 0xeb86d391u, 21);
#undef FB_MD5_STEP
#undef FB_MD5_F
#undef FB_MD5_G
#undef FB_MD5_H
#undef FB_MD5_I
        a_ += a; b_ += b; c_ += c; d_ += d;
    }

    uint32_t a_, b_, c_, d_;
    uint64_t len_;
    uint8_t buf_[64];
    size_t fill_;
};
