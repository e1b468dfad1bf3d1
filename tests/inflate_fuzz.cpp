// Sanitizer fuzz of csrc/inflate.cpp (built and run by tests/test_library.py): blocks of a BGZF file are decoded from
// exact-size heap copies after random damage, truncation, or with a wrong expected size; AddressSanitizer and UBSan
// see any read or write outside the two buffers.  usage: inflate_fuzz file.gz iterations
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <vector>
#include <random>
namespace cratac { bool inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len, uint32_t* scratch); size_t inflate_scratch_words(); }
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> buf(n); if (fread(buf.data(), 1, n, f) != (size_t)n) return 1; fclose(f);
    struct B { size_t off, cs; uint32_t is; }; std::vector<B> bl; size_t p = 0;
    while (p + 18 <= (size_t)n) { unsigned xlen = buf[p+10] | (buf[p+11] << 8); unsigned bs = (buf[p+16] | (buf[p+17] << 8)) + 1; uint32_t is; memcpy(&is, &buf[p+bs-4], 4); bl.push_back({p+12+xlen, bs-12-xlen-8, is}); p += bs; }
    std::mt19937_64 rng(7); std::vector<uint32_t> scr(cratac::inflate_scratch_words());
    long ok = 0, bad = 0; int iters = atoi(argv[2]);
    for (int it = 0; it < iters; it++) {
        const B& b = bl[rng() % bl.size()];
        // exact-size heap copies so that the sanitizer sees any over-read / over-write
        size_t cs = b.cs; int mode = rng() % 4;
        if (mode == 3 && cs > 4) cs = rng() % cs;                      // truncated input
        uint8_t* in = (uint8_t*)malloc(cs ? cs : 1); memcpy(in, &buf[b.off], cs);
        if (mode == 1 || mode == 2) for (int k = 0; k < 1 + (int)(rng() % 3); k++) if (cs) in[rng() % cs] ^= (uint8_t)(1 + rng() % 255);
        size_t os = b.is; if (mode == 2) os = rng() % (b.is + 300);    // wrong expected size
        uint8_t* out = (uint8_t*)malloc(os ? os : 1);
        if (cratac::inflate_raw(in, cs, out, os, scr.data())) ok++; else bad++;
        free(in); free(out);
    }
    printf("ok=%ld declined=%ld\n", ok, bad);
}
