/*
 * bits.h -- host-side bit I/O for the small serial parts of a packet: headers, stability map,
 * motion side-info (bs.c:21-267).  MSB-first; the writer ORs into zeroed memory like the reference.
 * Coefficient data never goes through here -- that is hzcc_enc.cu / hzcc_dec.cu on the GPU.
 */
#pragma once
#include <stdint.h>
#include <string.h>

namespace dsv {

struct BitWriter {
    uint8_t *buf;
    uint64_t pos; /* bits */
    explicit BitWriter(uint8_t *b = nullptr) : buf(b), pos(0) {}
    void align() { pos = (pos + 7) & ~(uint64_t) 7; }
    unsigned byte_pos() const { return (unsigned) (pos >> 3); }
    void put_bit(int b)
    {
        if (b) {
            buf[pos >> 3] |= (uint8_t) (0x80u >> (pos & 7));
        }
        pos++;
    }
    void put_bits(unsigned n, uint32_t v)
    {
        while (n--) {
            put_bit((v >> n) & 1);
        }
    }
    /* unsigned interleaved exp-Golomb (bs.c:128-145) */
    void put_ueg(uint32_t v)
    {
        uint32_t x = v + 1;
        int n = 31;
        while (!(x >> n)) {
            n--;
        }
        for (int i = n - 1; i >= 0; i--) {
            pos++; /* the '0' continuation flag */
            put_bit((x >> i) & 1);
        }
        put_bit(1);
    }
    /* signed (bs.c:159-175) */
    void put_seg(int v)
    {
        uint32_t m = (uint32_t) (v < 0 ? -v : v);
        put_ueg(m);
        if (m) {
            put_bit(v < 0);
        }
    }
    void concat(const uint8_t *data, unsigned len)
    {
        memcpy(buf + (pos >> 3), data, len);
        pos += (uint64_t) len * 8;
    }
};

/* zero-bit run-length coder (bs.c:221-267): UEG(number of zeros before each one), final run flushed */
struct RleWriter {
    BitWriter bw;
    unsigned nz;
    explicit RleWriter(uint8_t *b) : bw(b), nz(0) {}
    void put(int bit)
    {
        if (bit) {
            bw.put_ueg(nz);
            nz = 0;
        } else {
            nz++;
        }
    }
    unsigned finish()
    {
        bw.put_ueg(nz);
        nz = 0;
        bw.align();
        return bw.byte_pos();
    }
};

struct BitReader {
    const uint8_t *buf;
    uint64_t pos;
    uint64_t limit; /* bits available; reads past it return 0 (the reference would read out of bounds) */
    BitReader(const uint8_t *b = nullptr, uint64_t nbytes = 0) : buf(b), pos(0), limit(nbytes * 8) {}
    void align() { pos = (pos + 7) & ~(uint64_t) 7; }
    unsigned byte_pos() const { return (unsigned) (pos >> 3); }
    void skip_bytes(unsigned n) { pos += (uint64_t) n * 8; }
    bool exhausted() const { return pos >= limit; }
    unsigned get_bit()
    {
        unsigned r = 0;
        if (pos < limit) {
            r = (buf[pos >> 3] >> (7 - (pos & 7))) & 1;
        }
        pos++;
        return r;
    }
    uint32_t get_bits(unsigned n)
    {
        uint32_t v = 0;
        while (n--) {
            v = (v << 1) | get_bit();
        }
        return v;
    }
    uint32_t get_ueg() /* bs.c:147-157; terminates on truncated input */
    {
        uint32_t v = 1;
        while (!get_bit()) {
            if (exhausted()) {
                break;
            }
            v = (v << 1) | get_bit();
        }
        return v - 1;
    }
    int get_seg() /* bs.c:177-188 */
    {
        int v = (int) get_ueg();
        return (v && get_bit()) ? -v : v;
    }
};

struct RleReader {
    BitReader br;
    int nz;
    RleReader(const uint8_t *b, uint64_t nbytes) : br(b, nbytes), nz(0) {}
    int get() /* bs.c:257-267 */
    {
        if (nz == 0) {
            nz = (int) br.get_ueg();
            return nz == 0;
        }
        nz--;
        return nz == 0;
    }
};

} // namespace dsv
