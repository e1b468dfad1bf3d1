// Raw DEFLATE (RFC 1951) decoder for BGZF blocks: host ingest of fragments.tsv.gz (replaces the
// `gzip -c -d` pipe of lib/python/tools/io.py:197-217 together with the block-parallel driver in api.cu).
//
// A BGZF block inflates to at most 64 KiB, straight into its place in the caller's buffer, so the "window" is the
// output itself.  Decoding is table driven: a 10-bit first-level table for literal/length codes and an 8-bit one for
// distance codes, second-level tables for the longer codes, a 64-bit bit buffer refilled with one unaligned load, and
// word-wise match copies while there is slack before the end of the block's output.  Anything unexpected (malformed
// header, a code that is not in the tables, output that does not end exactly at the block's ISIZE) returns false and
// the caller hands the block to zlib, which then either decodes it or produces the error.
#include <stdint.h>
#include <string.h>

#include <mutex>

namespace cratac {

namespace {

constexpr int LL_BITS = 10, D_BITS = 8, CL_BITS = 7;
constexpr int LL_TABLE = (1 << LL_BITS) + 9216, D_TABLE = (1 << D_BITS) + 3840;   // worst-case second-level space

enum : uint32_t { K_LITERAL = 0, K_SYMBOL = 1, K_END = 2, K_SUBTABLE = 3, K_INVALID = 4 };

// entry: bits 0-7 code bits to consume, 8-10 kind, 11-15 extra bits (or sub-table bits), 16-31 base (or sub-table offset)
inline uint32_t make_entry(uint32_t len, uint32_t kind, uint32_t extra, uint32_t base) { return len | (kind << 8) | (extra << 11) | (base << 16); }
inline uint32_t e_len(uint32_t e) { return e & 0xFF; }
inline uint32_t e_kind(uint32_t e) { return (e >> 8) & 7; }
inline uint32_t e_extra(uint32_t e) { return (e >> 11) & 31; }
inline uint32_t e_base(uint32_t e) { return e >> 16; }

const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

inline uint32_t reverse_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; i++) { r = (r << 1) | (code & 1); code >>= 1; }
    return r;
}

enum TableKind { T_LITLEN, T_DIST, T_CODELEN };

inline uint32_t symbol_entry(TableKind kind, int sym, uint32_t len) {
    if (kind == T_CODELEN) return make_entry(len, K_SYMBOL, 0, (uint32_t)sym);
    if (kind == T_LITLEN) {
        if (sym < 256) return make_entry(len, K_LITERAL, 0, (uint32_t)sym);
        if (sym == 256) return make_entry(len, K_END, 0, 0);
        if (sym > 285) return make_entry(len, K_INVALID, 0, 0);
        return make_entry(len, K_SYMBOL, LEN_EXTRA[sym - 257], LEN_BASE[sym - 257]);
    }
    if (sym > 29) return make_entry(len, K_INVALID, 0, 0);
    return make_entry(len, K_SYMBOL, DIST_EXTRA[sym], DIST_BASE[sym]);
}

// Canonical Huffman code lengths -> decode table.  false: over-subscribed, or the tables would not fit.
bool build_table(TableKind kind, const uint8_t* lens, int n_sym, int primary_bits, uint32_t* table, int capacity) {
    int count[16] = {0};
    for (int s = 0; s < n_sym; s++) count[lens[s]]++;
    const int primary = 1 << primary_bits;
    for (int i = 0; i < primary; i++) table[i] = make_entry(1, K_INVALID, 0, 0);
    if (count[0] == n_sym) return true;                       // no code at all: every lookup is invalid
    uint32_t next_code[16];
    {
        long left = 1;                                        // Kraft: unused code space after each length
        for (int len = 1; len <= 15; len++) {
            left = (left << 1) - count[len];
            if (left < 0) return false;                       // over-subscribed
        }
        uint32_t code = 0;
        int prev = 0;                                         // bl_count[0] = 0 (RFC 1951, 3.2.2)
        for (int len = 1; len <= 15; len++) { code = (code + (uint32_t)prev) << 1; next_code[len] = code; prev = count[len]; }
    }
    // first pass over the long codes: how many bits each first-level prefix needs below it
    uint8_t sub_bits[1 << 11];                 // first-level size of the largest table
    memset(sub_bits, 0, (size_t)primary);
    uint32_t codes[320];
    for (int s = 0; s < n_sym; s++) {
        const int len = lens[s];
        if (!len) continue;
        const uint32_t r = reverse_bits(next_code[len]++, len);
        codes[s] = r;
        if (len > primary_bits) {
            const uint32_t prefix = r & (uint32_t)(primary - 1);
            if (len - primary_bits > sub_bits[prefix]) sub_bits[prefix] = (uint8_t)(len - primary_bits);
        }
    }
    int used = primary;
    for (int p = 0; p < primary; p++) {
        if (!sub_bits[p]) continue;
        const int size = 1 << sub_bits[p];
        if (used + size > capacity) return false;
        table[p] = make_entry((uint32_t)primary_bits, K_SUBTABLE, sub_bits[p], (uint32_t)used);
        for (int i = 0; i < size; i++) table[used + i] = make_entry(1, K_INVALID, 0, 0);
        used += size;
    }
    for (int s = 0; s < n_sym; s++) {
        const int len = lens[s];
        if (!len) continue;
        const uint32_t r = codes[s];
        if (len <= primary_bits) {
            const uint32_t e = symbol_entry(kind, s, (uint32_t)len);
            for (uint32_t i = r; i < (uint32_t)primary; i += 1u << len) table[i] = e;
        } else {
            const uint32_t head = table[r & (uint32_t)(primary - 1)];
            const uint32_t bits = e_extra(head), base = e_base(head);
            const uint32_t e = symbol_entry(kind, s, (uint32_t)(len - primary_bits));
            for (uint32_t i = r >> primary_bits; i < (1u << bits); i += 1u << (len - primary_bits)) table[base + i] = e;
        }
    }
    return true;
}

struct FixedTables {
    uint32_t ll[LL_TABLE], d[D_TABLE];
    bool ok;
};
FixedTables* fixed_tables() {
    static FixedTables t;
    static std::once_flag once;
    std::call_once(once, [] {
        uint8_t lens[288];
        for (int i = 0; i < 144; i++) lens[i] = 8;
        for (int i = 144; i < 256; i++) lens[i] = 9;
        for (int i = 256; i < 280; i++) lens[i] = 7;
        for (int i = 280; i < 288; i++) lens[i] = 8;
        uint8_t dl[32];
        for (int i = 0; i < 32; i++) dl[i] = 5;
        t.ok = build_table(T_LITLEN, lens, 288, LL_BITS, t.ll, LL_TABLE) && build_table(T_DIST, dl, 32, D_BITS, t.d, D_TABLE);
    });
    return &t;
}

struct BitReader {
    const uint8_t *in, *end;
    uint64_t buf = 0;
    int cnt = 0;                 // valid bits in buf
    bool overrun = false;        // bits were asked for beyond the input
    inline void refill() {
        if (end - in >= 8) {
            uint64_t w;
            memcpy(&w, in, 8);
            buf |= w << cnt;
            in += (63 - cnt) >> 3;
            cnt |= 56;
        } else {
            while (cnt <= 56 && in < end) { buf |= (uint64_t)*in++ << cnt; cnt += 8; }
        }
    }
    inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
    inline void drop(int n) {
        if (n > cnt) { overrun = true; n = cnt; }
        buf >>= n;
        cnt -= n;
    }
    inline uint32_t take(int n) { uint32_t v = peek(n); drop(n); return v; }
};

inline uint32_t lookup(const uint32_t* table, int primary_bits, BitReader& br) {
    uint32_t e = table[br.peek(primary_bits)];
    if (e_kind(e) == K_SUBTABLE) {
        br.drop(primary_bits);
        e = table[e_base(e) + br.peek((int)e_extra(e))];
    }
    br.drop((int)e_len(e));
    return e;
}

}  // namespace

// `in_len` bytes of raw DEFLATE -> exactly `out_len` bytes.  scratch: LL_TABLE + D_TABLE uint32 (per thread).
bool inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len, uint32_t* scratch) {
    uint32_t* dyn_ll = scratch;
    uint32_t* dyn_d = scratch + LL_TABLE;
    BitReader br;
    br.in = in;
    br.end = in + in_len;
    uint8_t* const out0 = out;
    uint8_t* const out_end = out + out_len;
    for (;;) {
        br.refill();
        const uint32_t last = br.take(1), type = br.take(2);
        if (type == 0) {
            br.drop(br.cnt & 7);                               // to the byte boundary
            br.refill();
            if (br.cnt < 32) return false;
            const uint32_t len = br.take(16), nlen = br.take(16);
            if ((len ^ nlen) != 0xFFFFu) return false;
            // bytes still in the bit buffer belong to the stored data
            size_t n = len;
            while (n && br.cnt >= 8) {
                if (out >= out_end) return false;
                *out++ = (uint8_t)br.take(8);
                n--;
            }
            if (n) br.buf = 0;                                 // byte aligned and drained: drop the look-ahead bits as well
            if ((size_t)(br.end - br.in) < n || (size_t)(out_end - out) < n) return false;
            memcpy(out, br.in, n);
            br.in += n;
            out += n;
        } else if (type == 1 || type == 2) {
            const uint32_t *ll, *dt;
            if (type == 1) {
                FixedTables* f = fixed_tables();
                if (!f->ok) return false;
                ll = f->ll;
                dt = f->d;
            } else {
                br.refill();
                const uint32_t hlit = br.take(5) + 257, hdist = br.take(5) + 1, hclen = br.take(4) + 4;
                if (hlit > 286 || hdist > 30) return false;
                uint8_t cl[19];
                memset(cl, 0, sizeof(cl));
                for (uint32_t i = 0; i < hclen; i++) {
                    if (br.cnt < 3) br.refill();
                    cl[CL_ORDER[i]] = (uint8_t)br.take(3);
                }
                uint32_t clt[1 << CL_BITS];
                if (!build_table(T_CODELEN, cl, 19, CL_BITS, clt, 1 << CL_BITS)) return false;
                uint8_t lens[288 + 32];
                uint32_t i = 0;
                while (i < hlit + hdist) {
                    br.refill();
                    const uint32_t e = clt[br.peek(CL_BITS)];
                    if (e_kind(e) != K_SYMBOL) return false;
                    br.drop((int)e_len(e));
                    const uint32_t sym = e_base(e);
                    if (sym < 16) { lens[i++] = (uint8_t)sym; continue; }
                    uint32_t rep, val = 0;
                    if (sym == 16) {
                        if (i == 0) return false;
                        val = lens[i - 1];
                        rep = 3 + br.take(2);
                    } else if (sym == 17) rep = 3 + br.take(3);
                    else rep = 11 + br.take(7);
                    if (i + rep > hlit + hdist) return false;
                    while (rep--) lens[i++] = (uint8_t)val;
                }
                if (br.overrun || lens[256] == 0) return false;
                if (!build_table(T_LITLEN, lens, (int)hlit, LL_BITS, dyn_ll, LL_TABLE)) return false;
                if (!build_table(T_DIST, lens + hlit, (int)hdist, D_BITS, dyn_d, D_TABLE)) return false;
                ll = dyn_ll;
                dt = dyn_d;
            }
            // Fast loop: at least 8 input bytes ahead (every refill is one unaligned load of real data, so no bit is
            // ever missing) and room for the longest match plus the slack of the word copies: no per-symbol checks.
            bool end_of_block = false;
            constexpr uint32_t LL_MASK = (1u << LL_BITS) - 1, D_MASK = (1u << D_BITS) - 1;
            while (br.end - br.in >= 8 && out_end - out >= 258 + 16) {
                {
                    uint64_t w;
                    memcpy(&w, br.in, 8);
                    br.buf |= w << br.cnt;
                    br.in += (63 - br.cnt) >> 3;
                    br.cnt |= 56;
                }
                uint32_t e = ll[br.buf & LL_MASK];
                if (e_kind(e) == K_SUBTABLE) {
                    br.buf >>= LL_BITS; br.cnt -= LL_BITS;
                    e = ll[e_base(e) + (uint32_t)(br.buf & ((1u << e_extra(e)) - 1))];
                }
                br.buf >>= e_len(e); br.cnt -= (int)e_len(e);
                const uint32_t kind = e_kind(e);
                if (kind == K_LITERAL) {
                    *out++ = (uint8_t)e_base(e);
                    e = ll[br.buf & LL_MASK];
                    if (e_kind(e) != K_LITERAL) continue;
                    br.buf >>= e_len(e); br.cnt -= (int)e_len(e);
                    *out++ = (uint8_t)e_base(e);
                    e = ll[br.buf & LL_MASK];
                    if (e_kind(e) != K_LITERAL) continue;
                    br.buf >>= e_len(e); br.cnt -= (int)e_len(e);
                    *out++ = (uint8_t)e_base(e);
                    continue;
                }
                if (kind != K_SYMBOL) {
                    if (kind == K_END) { end_of_block = true; break; }
                    return false;
                }
                const uint32_t length = e_base(e) + (uint32_t)(br.buf & ((1u << e_extra(e)) - 1));
                br.buf >>= e_extra(e); br.cnt -= (int)e_extra(e);
                uint32_t de = dt[br.buf & D_MASK];
                if (e_kind(de) == K_SUBTABLE) {
                    br.buf >>= D_BITS; br.cnt -= D_BITS;
                    de = dt[e_base(de) + (uint32_t)(br.buf & ((1u << e_extra(de)) - 1))];
                }
                br.buf >>= e_len(de); br.cnt -= (int)e_len(de);
                if (e_kind(de) != K_SYMBOL) return false;
                const uint32_t dist = e_base(de) + (uint32_t)(br.buf & ((1u << e_extra(de)) - 1));
                br.buf >>= e_extra(de); br.cnt -= (int)e_extra(de);
                if (dist > (size_t)(out - out0)) return false;
                const uint8_t* src = out - dist;
                uint8_t* dst = out;
                out += length;
                if (dist >= 8) {
                    do { uint64_t w; memcpy(&w, src, 8); memcpy(dst, &w, 8); src += 8; dst += 8; } while (dst < out);
                } else if (dist == 1) {
                    memset(dst, *src, length);
                } else {
                    do { *dst++ = *src++; } while (dst < out);
                }
            }
            if (!end_of_block)
            for (;;) {
                br.refill();                                   // >= 56 bits unless the input is about to end
                uint32_t e = lookup(ll, LL_BITS, br);
                uint32_t kind = e_kind(e);
                if (kind == K_LITERAL) {
                    if (out >= out_end) return false;
                    *out++ = (uint8_t)e_base(e);
                    // up to two more literals from the bits already in the buffer (<= 3 x 15 bits of the 56)
                    e = ll[br.peek(LL_BITS)];
                    if (e_kind(e) != K_LITERAL || out >= out_end) continue;
                    br.drop((int)e_len(e));
                    *out++ = (uint8_t)e_base(e);
                    e = ll[br.peek(LL_BITS)];
                    if (e_kind(e) != K_LITERAL || out >= out_end) continue;
                    br.drop((int)e_len(e));
                    *out++ = (uint8_t)e_base(e);
                    continue;
                }
                if (kind == K_END) break;
                if (kind != K_SYMBOL) return false;
                const uint32_t length = e_base(e) + br.take((int)e_extra(e));        // 15 + 5 bits at most so far
                if (br.cnt < 28) br.refill();
                const uint32_t de = lookup(dt, D_BITS, br);
                if (e_kind(de) != K_SYMBOL) return false;
                const uint32_t dist = e_base(de) + br.take((int)e_extra(de));
                if (dist > (size_t)(out - out0) || length > (size_t)(out_end - out)) return false;
                const uint8_t* src = out - dist;
                if (dist >= 8 && (size_t)(out_end - out) >= length + 8) {
                    // word copies may run up to 7 bytes past the match, still inside this block's output
                    uint8_t* dst = out;
                    const uint8_t* stop = out + length;
                    do { uint64_t w; memcpy(&w, src, 8); memcpy(dst, &w, 8); src += 8; dst += 8; } while (dst < stop);
                } else {
                    for (uint32_t k = 0; k < length; k++) out[k] = src[k];
                }
                out += length;
            }
            if (br.overrun) return false;
        } else {
            return false;
        }
        if (br.overrun) return false;
        if (last) break;
    }
    return out == out_end;
}

size_t inflate_scratch_words() { return (size_t)LL_TABLE + D_TABLE; }

}  // namespace cratac
