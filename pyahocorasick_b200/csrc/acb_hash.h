/*
 * acb_hash.h -- the gram hash shared by the host (filter construction, acb_host.cpp)
 * and the device (probing, acb_device.cu).  Both sides MUST compute identical values.
 *
 * A gram is g <= 16 consecutive bytes.  It is read as NW = ceil(g/4) little-endian
 * 32-bit windows w_0..w_{NW-1}; bytes past g inside the last window are cancelled by
 * giving that window a multiplier whose low 8*(4*NW-g) bits are zero (multiplication
 * mod 2^32 then ignores the high bytes of the window), so the probe loop needs no
 * masking instruction.
 *
 *      h  = sum_k  w_k * mul[k]      (mod 2^32)
 *      bit index = h >> (32 - log2_bits)
 *
 * The shared-memory bitmap (hash 1) and the anchor table (hash 2, the "tag") use different odd
 * multiplier sets so that their false positives are independent.
 */
#ifndef ACB_HASH_H_INCLUDED
#define ACB_HASH_H_INCLUDED

#include <stdint.h>

#define ACB_MAX_GRAM 16
#define ACB_MAX_WINDOWS 4

#if defined(__CUDACC__)
#define ACB_HD __host__ __device__ __forceinline__
#else
#define ACB_HD static inline
#endif

/* odd 32-bit constants (golden-ratio / murmur / xxhash finalizer primes) */
#define ACB_S1_M0 0x9E3779B1u
#define ACB_S1_M1 0x85EBCA77u
#define ACB_S1_M2 0xC2B2AE3Du
#define ACB_S1_M3 0x27D4EB2Fu
#define ACB_S2_M0 0x165667B1u
#define ACB_S2_M1 0xD3A2646Du
#define ACB_S2_M2 0xFD7046C5u
#define ACB_S2_M3 0xB55A4F09u
#define ACB_TAGMAP_MIX 0x9E3779B1u   /* tag bitmap (global memory): bit index = (tag * ACB_TAGMAP_MIX) >> (32 - log2_bits3) */

/* fill mul[0..3] for a gram of g bytes; stage = 1 or 2 */
ACB_HD void acb_hash_multipliers(int g, int stage, uint32_t mul[ACB_MAX_WINDOWS]) {
    const uint32_t base1[4] = {ACB_S1_M0, ACB_S1_M1, ACB_S1_M2, ACB_S1_M3};
    const uint32_t base2[4] = {ACB_S2_M0, ACB_S2_M1, ACB_S2_M2, ACB_S2_M3};
    int nw = (g + 3) / 4;
    for (int k = 0; k < ACB_MAX_WINDOWS; k++) {
        uint32_t m = (stage == 1) ? base1[k] : base2[k];
        if (k >= nw) m = 0;
        else if (k == nw - 1) {
            int unused = 4 * nw - g;          /* high bytes of the last window to ignore */
            m = (unused >= 4) ? 0u : (m << (8 * unused));
        }
        mul[k] = m;
    }
}

/* hash of the gram whose windows are already loaded */
ACB_HD uint32_t acb_hash_windows(const uint32_t w[ACB_MAX_WINDOWS], const uint32_t mul[ACB_MAX_WINDOWS], int nw) {
    uint32_t h = 0;
    for (int k = 0; k < nw; k++) h += w[k] * mul[k];
    return h;
}

/* hash of g bytes at p (bytes beyond the gram are never read) -- host + slow device path */
ACB_HD uint32_t acb_hash_bytes(const uint8_t *p, int g, const uint32_t mul[ACB_MAX_WINDOWS]) {
    uint32_t h = 0;
    int nw = (g + 3) / 4;
    for (int k = 0; k < nw; k++) {
        uint32_t w = 0;
        for (int b = 0; b < 4; b++) {
            int i = 4 * k + b;
            if (i < g) w |= (uint32_t)p[i] << (8 * b);
        }
        h += w * mul[k];
    }
    return h;
}

/* The same sum with 64-bit products (its low half IS acb_hash_bytes): when the gram fills its windows exactly
 * (g % 4 == 0, no byte is cancelled through a shifted multiplier) the high half is a second, well mixed hash that
 * costs the probe loop nothing -- mad.wide.u32 instead of mad.lo.u32 -- and supplies the first Bloom bit. */
ACB_HD uint64_t acb_hash_bytes_wide(const uint8_t *p, int g, const uint32_t mul[ACB_MAX_WINDOWS]) {
    uint64_t h = 0;
    int nw = (g + 3) / 4;
    for (int k = 0; k < nw; k++) {
        uint32_t w = 0;
        for (int b = 0; b < 4; b++) {
            int i = 4 * k + b;
            if (i < g) w |= (uint32_t)p[i] << (8 * b);
        }
        h += (uint64_t)w * mul[k];
    }
    return h;
}

/* does the stage-1 filter of a gram of g bytes take its first Bloom bit from the high half? */
ACB_HD int acb_hash_is_wide(int g) { return (g % 4) == 0; }

/* the two bit positions (0..31) of a gram inside its stage-1 word; hw = acb_hash_bytes_wide() */
ACB_HD uint32_t acb_stage1_bit_a(uint64_t hw, int g, int log2_bits) {
    return acb_hash_is_wide(g) ? ((uint32_t)(hw >> 32) & 31u) : (((uint32_t)hw >> (32 - log2_bits)) & 31u);
}
ACB_HD uint32_t acb_stage1_bit_b(uint64_t hw) { return (uint32_t)hw & 31u; }

/* ---- PAIR placement (gram 4, stride 1, 1-byte letters) -----------------------------------------------------
 * Positions x (even) and x+1 share the three bytes text[x+1 .. x+3].  ONE 64-bit product of the window at x+1,
 *      pr = (uint64) window(x+1) * (ACB_PAIR_M << 8),     hc = low half (the shifted multiplier drops the window's
 *      fourth byte: hc depends on the three common bytes only),     hb = high half (all four bytes),
 * selects
 *   level 1:  ONE WORD of 2^(n-5) words, index hc >> (37-n), one shared-memory load per TWO positions, in which
 *             role 0 (the gram G starts at the even position x; its bytes 1..3 are the common ones) owns bit
 *             31 - (G & 31) -- the low five bits of the gram's first byte, the one byte hc does not see -- and
 *             role 1 (G starts at x+1, G IS the window; bytes 0..2 common) owns bit 31 - (hb & 31).  The kernel
 *             shifts the word LEFT by the raw byte / by hb (wrap shifts use the low five bits), so the tested bit
 *             lands in bit 31 and is added into a per-POSITION pass mask through the carry.  Every key gram is
 *             entered under both roles (a key may start at an even or an odd position).
 *   level 2:  a second bitmap of 2^k bits right behind level 1 (shared memory as well), a blocked Bloom filter with
 *             two bits per gram keyed by the anchor tag (hash 2 of the gram, role-independent): word
 *             tag >> (37-k), bits (tag >> (32-k)) & 31 and (tag >> (27-k)) & 31.  Only positions that pass level 1
 *             (about 2 % on random text against 10 k keys) compute their tag at all. */
#define ACB_PAIR_M  0x9E3779B1u
ACB_HD uint32_t acb_pair_mul(void) { return ACB_PAIR_M << 8; }
/* level-1 placement of gram G (little-endian word) in one role: *word (index from the start of the bitmap) |= *bit */
ACB_HD void acb_pair_place(uint32_t G, int role, int log2_bits, uint32_t *word, uint32_t *bit) {
    const uint32_t common = role ? G : (G >> 8);
    const uint64_t pr = (uint64_t)common * (uint64_t)(ACB_PAIR_M << 8);
    const uint32_t hc = (uint32_t)pr;
    const uint32_t amount = role ? (uint32_t)(((uint64_t)G * (uint64_t)(ACB_PAIR_M << 8)) >> 32) : G;
    *word = hc >> (37 - log2_bits);
    *bit = 1u << (31u - (amount & 31u));
}
/* level-2 placement of the anchor tag: *word counts from the start of level 2 */
ACB_HD void acb_pair_place2(uint32_t tag, int log2_bits2, uint32_t *word, uint32_t *bits) {
    *word = tag >> (37 - log2_bits2);
    *bits = (1u << ((tag >> (32 - log2_bits2)) & 31u)) | (1u << ((tag >> (27 - log2_bits2)) & 31u));
}

#endif
