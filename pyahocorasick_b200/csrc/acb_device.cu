/*
 * acb_device.cu -- sm_100a scan kernels and the device half of the C ABI (include/acb200.h).
 *
 * ACB_ALGO_FILTER (the fast path) is two launches per <= 2 GiB segment of the batch:
 *
 *  acb_filter_kernel<NW,STRIDE,WIDE>   persistent, one CTA per SM; streams the haystack bytes once.
 *      Start-anchored search.  Every probe position is tested against the stage-1 gram filter in shared memory
 *      (a blocked Bloom filter, two bits per gram in one word).  The rare survivors are hashed a second time
 *      and tested against stage 2 (shared memory) and, for large key sets, stage 3 (global memory), then queued
 *      per warp.  Between work units a warp resolves its queue through the anchor table in global memory (one
 *      32-byte slot): a UNIQUE anchor carries the only key that can match there, which is compared with the
 *      text directly; a MULTI anchor (keys sharing that prefix) walks the trie through the column-major goto
 *      table.  No failure links are followed: an occurrence is found exactly once, from its first byte, so the
 *      result set equals what the reference produces by walking fail chains at every position
 *      (src/AutomatonSearchIter.c:157-197, src/Automaton.c:693-714).
 *  acb_verify_kernel                   resolves the candidates a warp had to spill to the global list because
 *      its queue was full (normally none), re-arms the counters, flags an overflowed list.
 *
 *  acb_dfa_kernel                      (ACB_ALGO_DFA)
 *      The textbook automaton: goto, else fail until root (src/trie.c:177-194), outputs from CSR lists.  One
 *      lane per 64-byte span with a max_key-1 byte warm-up.  Slower (every byte is a dependent L2 lookup) but
 *      insensitive to key-set shape; also used to cross-check the filter kernel on the GPU.
 *
 *  acb_long_kernel                     (ACB_ALGO_LONG)
 *      iter_long: the reference's longest-match walk (src/AutomatonSearchIterLong.c:89-153) replayed letter by
 *      letter on the flattened tables, one lane per haystack.
 *
 * Match records are compacted per warp in shared memory and appended to the global buffer with one atomicAdd
 * per warp flush; acb_sort_matches_device puts them into the reference's order with one radix sort.
 */
#include "acb_internal.h"
#include "acb_hash.h"

#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <new>
#include <vector>

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            acb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,      \
                          __LINE__, cudaGetErrorString(_e));                                 \
            return ACB_ECUDA;                                                                \
        }                                                                                    \
    } while (0)

namespace {

constexpr int kThreads      = 1024;               /* filter kernel: one CTA per SM                 */
constexpr int kWarps        = kThreads / 32;
#ifndef ACB_BLOCK_BYTES
#define ACB_BLOCK_BYTES 4096
#endif
constexpr int kBlockBytes   = ACB_BLOCK_BYTES;    /* work unit per warp grab (filter kernel)       */
constexpr int kQueueCap     = 128;                /* stage-1 survivors queued per warp (smem)      */
constexpr int kStageCap     = 32;                 /* match records staged per warp (smem)          */
constexpr uint32_t kFull    = 0xffffffffu;
constexpr int32_t  kTermBit = 0x40000000;         /* goto entry flag: child ends a key             */
constexpr int32_t  kIdMask  = 0x3fffffff;
constexpr long long kSegBytes = 1LL << 31;        /* candidates are uint32 offsets into a segment  */

constexpr int kVerThreads   = 256;
constexpr int kVerWarps     = kVerThreads / 32;

constexpr int kDfaSpan      = 64;                 /* bytes per lane in the DFA kernel              */
constexpr int kDfaThreads   = 256;

std::atomic<long long> g_launches{0};
thread_local float g_last_ms = 0.f;
std::atomic<int> g_timing{0};

struct ScanParams {
    const uint8_t *hay;
    long long total;
    const long long *offsets;      /* nullptr => fixed stride */
    long long n_hay;
    long long stride_bytes;
    const uint8_t *cls;
    const int32_t *gto;            /* flagged goto (kTermBit) */
    const int32_t *fail;
    const int32_t *letter_fail;
    const int32_t *key_of;
    const int32_t *out_ptr;
    const int32_t *out_idx;
    const int32_t *key_len;
    int32_t S;
    int32_t L;
    int32_t gram;
    int32_t max_key_bytes;
    const uint32_t *bm1;
    const uint32_t *bm2;
    const uint32_t *bm3;           /* stage 3, global memory; log3 == 0: unused */
    const uint4 *anchors;          /* 2 x uint4 per slot */
    int32_t log1, log2, log3, logA;
    uint32_t mul1[ACB_MAX_WINDOWS];
    uint32_t mul2[ACB_MAX_WINDOWS];
    acb_match *out;
    long long cap;
    unsigned long long *count;
    /* filter -> verify hand-over */
    long long seg_begin, seg_end;  /* byte range of this launch pair */
    long long n_blocks;            /* work units in the segment */
    uint2 *cand;                   /* candidates: {probe position relative to seg_begin, hash2(gram)|1} */
    unsigned long long cand_cap;
    unsigned long long *cand_count;
    unsigned int *work_ctr;        /* [0] next work unit, [1] filter CTAs done, [2] verify CTAs done */
    int stride_shift;              /* log2(stride_bytes) when it is a power of two, else -1 */
    int letter_shift;              /* log2(L) */
    unsigned long long *timeline;  /* debug (ACB_TIMELINE=1): 6 globaltimer stamps per warp, else nullptr */
    int inline_resolve;            /* 1: warps resolve their own candidates between work units; 0: all go to the list */
    int filter_flags;              /* ACB_FILTER_* of the table */
};

/* ---------------------------------------------------------------- helpers */

/* aligned 32-bit word at byte offset a (a % 4 == 0), zero filled past the end of the buffer */
__device__ __forceinline__ uint32_t load_word(const uint8_t *hay, long long a, long long total) {
    if (a + 4 <= total) return __ldg(reinterpret_cast<const uint32_t *>(hay + a));
    uint32_t v = 0;
    for (int b = 0; b < 4; b++) if (a + b < total) v |= (uint32_t)hay[a + b] << (8 * b);
    return v;
}

__device__ __forceinline__ void find_haystack(const ScanParams &p, long long q, long long &h, long long &hs, long long &he) {
    if (p.offsets == nullptr) {
        if (p.stride_shift >= 0) h = q >> p.stride_shift;
        else if (p.total <= 0xffffffffLL) h = (long long)((uint32_t)q / (uint32_t)p.stride_bytes);
        else h = q / p.stride_bytes;
        hs = h * p.stride_bytes;
        he = hs + p.stride_bytes;
    } else {
        long long lo = 0, hi = p.n_hay;       /* largest h with offsets[h] <= q */
        while (hi - lo > 1) {
            long long mid = (lo + hi) >> 1;
            if (__ldg(p.offsets + mid) <= q) lo = mid; else hi = mid;
        }
        h = lo;
        hs = __ldg(p.offsets + lo);
        he = __ldg(p.offsets + lo + 1);
    }
}

struct WarpStage {            /* per-warp match staging in shared memory */
    acb_match *buf;
    int *cnt;
};

__device__ __forceinline__ void emit(const ScanParams &p, const WarpStage &ws, int32_t h, int32_t e, int32_t k) {
    acb_match m;
    m.hay_id = h;
    m.end_index = e;
    m.key_id = k;
    int slot = atomicAdd(ws.cnt, 1);
    if (slot < kStageCap) {
        ws.buf[slot] = m;
    } else {                                  /* staging full: straight to global */
        unsigned long long g = atomicAdd(p.count, 1ULL);
        if (g < (unsigned long long)p.cap) p.out[g] = m;
    }
}

/* all 32 lanes must call; flushes the staged records with one global atomic */
__device__ __forceinline__ void flush_stage(const ScanParams &p, const WarpStage &ws, int lane) {
    __syncwarp();
    int n = *ws.cnt;
    if (n > kStageCap) n = kStageCap;
    if (n > 0) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(p.count, (unsigned long long)n);
        base = __shfl_sync(kFull, base, 0);
        if (lane < n && base + lane < (unsigned long long)p.cap) p.out[base + lane] = ws.buf[lane];
    }
    __syncwarp();
    if (lane == 0) *ws.cnt = 0;
    __syncwarp();
}

/* stage 3 (MULTI anchors only): walk the trie from the root at `start` */
__device__ __forceinline__ void walk_from(const ScanParams &p, const WarpStage &ws, long long start,
                                          long long h, long long hs, long long he) {
    const int L = p.L;
    int32_t st = 0;
    for (long long i = start; i < he; ++i) {
        int c = __ldg(p.cls + p.hay[i]);
        int32_t nx = __ldg(p.gto + (long long)c * p.S + st);
        if (nx < 0) break;
        st = nx & kIdMask;
        if (nx & kTermBit) {
            int32_t k = __ldg(p.key_of + st);
            emit(p, ws, (int32_t)h, (int32_t)((i - hs + 1) / L - 1), k);
        }
    }
}

/* six aligned words covering the 20 text bytes from x on (zero fill past the end of the buffer) */
__device__ __forceinline__ void load_text(const ScanParams &p, long long x, uint32_t (&w)[6]) {
    const long long x0 = x & ~3LL;
    if (x0 + 24 <= p.total) {
        const uint32_t *a = reinterpret_cast<const uint32_t *>(p.hay + x0);
#pragma unroll
        for (int i = 0; i < 6; i++) w[i] = __ldg(a + i);
    } else {
#pragma unroll
        for (int i = 0; i < 6; i++) w[i] = load_word(p.hay, x0 + 4 * i, p.total);
    }
}

/* do the n (<= 20) text bytes at x (words w = load_text(x)) equal the packed bytes kw? */
__device__ __forceinline__ bool text_equals(const uint32_t (&w)[6], long long x, int n, const uint32_t kw[5]) {
    const int sh = (int)(x & 3) * 8;
    uint32_t diff = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
        uint32_t t = __funnelshift_r(w[i], w[i + 1], sh);
        int nb = n - 4 * i;                                   /* bytes of this word that count */
        uint32_t mask = nb >= 4 ? 0xffffffffu : (nb <= 0 ? 0u : ((1u << (8 * nb)) - 1u));
        diff |= (t ^ kw[i]) & mask;
    }
    return diff == 0;
}

/* ------------------------------------------------------- the filter kernel */
/* Stage 1.  Persistent CTAs; every warp grabs 4 KiB work units.  Per iteration a lane owns 32
 * consecutive bytes (two 16-byte loads, prefetched one iteration ahead), hashes the gram at
 * each probe position and tests it against the bitmap in shared memory.  Survivors are queued
 * per warp in shared memory together with hash2 of their gram (re-read through L1) and written,
 * 32 at a time, to the global candidate list.  Work units that lie completely inside the buffer
 * run a variant without any bounds check (GUARD = false). */

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define ACB_STAMP(i) do { if (p.timeline && lane == 0) p.timeline[((size_t)blockIdx.x * kWarps + warp) * 6 + (i)] = gtime(); } while (0)

__device__ __forceinline__ unsigned long long mad_wide(uint32_t a, uint32_t b, unsigned long long c) {
    unsigned long long d;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}

__device__ __forceinline__ uint32_t lds_word(uint32_t saddr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

__device__ __forceinline__ void resolve(const ScanParams &p, const WarpStage &ws, uint2 c) {
    const long long q = p.seg_begin + (long long)c.x;
    const uint32_t tag = c.y;
    const uint32_t amask = (1u << p.logA) - 1u;
    uint32_t slot = tag >> (32 - p.logA);
    long long h = -1, hs = 0, he = 0;
    /* the text at the probe position is fetched together with the anchor slot, not after it: for stride-1
       filters (j == 0) that is the text every entry compares with, so the two L2 round trips overlap */
    uint32_t tq[6];
    load_text(p, q, tq);
    for (;;) {
        const uint4 e0 = __ldg(p.anchors + 2 * (size_t)slot);             /* both halves of the 32-byte slot at once */
        const uint4 e1 = __ldg(p.anchors + 2 * (size_t)slot + 1);
        if (e0.x == 0u) break;                                 /* empty slot ends the probe sequence */
        if (e0.x == tag) {
            const uint32_t kw[5] = {e0.w, e1.x, e1.y, e1.z, e1.w};
            const int j = (int)(e0.z & 0xffu), len = (int)((e0.z >> 8) & 0xffu);
            const int32_t kid = (int32_t)e0.y;
            if (h < 0) find_haystack(p, q, h, hs, he);
            const long long start = q - j;
            if (start >= hs) {
                if (kid >= 0) {                                /* UNIQUE: the only key that can match at start */
                    bool eq = false;
                    if (start + len <= he) {
                        if (j == 0) eq = text_equals(tq, q, len, kw);
                        else { uint32_t ts[6]; load_text(p, start, ts); eq = text_equals(ts, start, len, kw); }
                    }
                    if (eq) emit(p, ws, (int32_t)h, (int32_t)(((start + len - hs) >> p.letter_shift) - 1), kid);
                } else if (q + len <= he && text_equals(tq, q, len, kw)) {   /* MULTI: exact gram, then the trie */
                    walk_from(p, ws, start, h, hs, he);
                }
            }
            if (e0.z & 0x10000u) break;                        /* no further entry carries this tag */
        }
        slot = (slot + 1) & amask;
    }
}

/* per-warp state of the in-kernel stage 2/3 (all in shared memory except the ring positions) */
struct WarpResolve {
    uint2 *queue;      /* candidates {pos, tag} that passed both bitmaps, ring of kQueueCap */
    WarpStage ws;
};

/* Resolve the warp's queued candidates (they passed both bitmaps) through the anchor table, 32 at a
 * time.  Called between work units, never from the probe loop, so that none of the probe loop's
 * registers are live across it.  While this warp waits on L2 the other warps keep filtering. */
__device__ __noinline__ void drain_queue(const ScanParams &p, const WarpResolve &wr, int &qhead, int qtail,
                                         bool final_flush, int lane) {
    while (qtail - qhead >= 32 || (final_flush && qtail > qhead)) {
        __syncwarp();
        const int n = (qtail - qhead >= 32) ? 32 : (qtail - qhead);
        if (lane < n) resolve(p, wr.ws, wr.queue[(qhead + lane) & (kQueueCap - 1)]);
        qhead += n;
        flush_stage(p, wr.ws, lane);
    }
}

struct FilterCtx {
    const uint8_t *seg;        /* hay + seg_begin */
    uint32_t seg_len;          /* bytes of this segment: positions >= seg_len belong to the next launch */
    long long total_rel;       /* total - seg_begin: bytes that exist from seg onwards */
    uint32_t sbm;              /* shared-memory address of the bitmap */
    uint32_t mul_word;         /* umulhi(h, mul_word) = stage-1 word index (7/8 of the words) */
    const uint32_t *bm3;       /* stage-3 bitmap in global memory (large key sets), log3 == 0: unused */
    int log3;
    uint32_t sbm2;             /* shared-memory address of the stage-2 bitmap (last 1/8) */
    int sh_word2, sh_bit2;     /* tag >> sh_word2 = stage-2 word index, tag >> sh_bit2 = its bit index */
    uint32_t four;             /* == 4, opaque to the compiler so the address is one IMAD (FMA pipe) */
    uint32_t two;              /* == 2, same trick for the hit accumulator */
    int sh_bit;                /* h >> sh_bit: bit index (low 5 bits, wrap shift) */
    uint32_t lt_mask;
    int lane;
    uint2 *queue;
    uint2 *cand;
    unsigned long long cand_cap;
    unsigned long long *cand_count;
};

/* move queued candidates to the global list, 32 at a time (all of them when `all` is set) */
__device__ __forceinline__ void spill_queue(const FilterCtx &c, int &qhead, int qtail, bool all) {
    while (qtail - qhead >= 32 || (all && qtail > qhead)) {
        __syncwarp();
        const int n = (qtail - qhead >= 32) ? 32 : (qtail - qhead);
        unsigned long long g = 0;
        if (c.lane == 0) g = atomicAdd(c.cand_count, (unsigned long long)n);
        g = __shfl_sync(kFull, g, 0);
        if (c.lane < n && g + c.lane < c.cand_cap) c.cand[g + c.lane] = c.queue[(qhead + c.lane) & (kQueueCap - 1)];
        qhead += n;
        __syncwarp();
    }
}

template <bool GUARD>
__device__ __forceinline__ uint4 ld_chunk(const FilterCtx &c, uint32_t rel) {
    if (!GUARD || (long long)rel + 16 <= c.total_rel) return __ldg(reinterpret_cast<const uint4 *>(c.seg + rel));
    uint32_t w[4] = {0, 0, 0, 0};
    for (int i = 0; i < 16; i++) if ((long long)rel + i < c.total_rel) w[i >> 2] |= (uint32_t)c.seg[rel + i] << (8 * (i & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

template <bool GUARD>
__device__ __forceinline__ uint32_t ld_word(const FilterCtx &c, uint32_t rel) {       /* rel % 4 == 0 */
    if (!GUARD || (long long)rel + 4 <= c.total_rel) return __ldg(reinterpret_cast<const uint32_t *>(c.seg + rel));
    uint32_t v = 0;
    for (int b = 0; b < 4; b++) if ((long long)rel + b < c.total_rel) v |= (uint32_t)c.seg[rel + b] << (8 * b);
    return v;
}

constexpr int kFChunk = 32;                       /* bytes per lane per iteration   */
constexpr int kFWarpBytes = 32 * kFChunk;         /* 1 KiB per warp iteration       */
constexpr int kFIters = kBlockBytes / kFWarpBytes;

/* W[off + g] for g in 0..7 without dynamic register indexing: a 3-level select tree */
template <int N>
__device__ __forceinline__ uint32_t sel8(const uint32_t (&W)[N], int off, int g) {
    const uint32_t a0 = (g & 1) ? W[off + 1] : W[off + 0], a1 = (g & 1) ? W[off + 3] : W[off + 2];
    const uint32_t a2 = (g & 1) ? W[off + 5] : W[off + 4], a3 = (g & 1) ? W[off + 7] : W[off + 6];
    const uint32_t b0 = (g & 2) ? a1 : a0, b1 = (g & 2) ? a3 : a2;
    return (g & 4) ? b1 : b0;
}

template <int NW, int STRIDE, bool WIDE, bool GUARD>
__device__ __forceinline__ void filter_unit(const FilterCtx &c, const uint32_t (&mul)[NW], const uint32_t (&mul2)[NW],
                                            uint32_t rel0, int &qhead, int &qtail) {
    constexpr int kProbes = kFChunk / STRIDE;
    const int lane = c.lane;
    uint4 cur0 = ld_chunk<GUARD>(c, rel0 + lane * kFChunk);
    uint4 cur1 = ld_chunk<GUARD>(c, rel0 + lane * kFChunk + 16);
#pragma unroll 1
    for (int it = 0; it < kFIters; ++it) {
        const uint32_t pos0 = rel0 + it * kFWarpBytes + lane * kFChunk;
        if (GUARD && rel0 + it * kFWarpBytes >= c.seg_len) break;                  /* warp-uniform */
        /* prefetch the next iteration; past the unit only lane 0's first chunk is needed (look-ahead) */
        uint4 nxt0 = make_uint4(0, 0, 0, 0), nxt1 = make_uint4(0, 0, 0, 0);
        if (it + 1 < kFIters) {
            nxt0 = ld_chunk<GUARD>(c, pos0 + kFWarpBytes);
            nxt1 = ld_chunk<GUARD>(c, pos0 + kFWarpBytes + 16);
        } else if (lane == 0) {
            nxt0 = ld_chunk<GUARD>(c, pos0 + kFWarpBytes);
        }
        uint32_t W[8 + NW];
        W[0] = cur0.x; W[1] = cur0.y; W[2] = cur0.z; W[3] = cur0.w;
        W[4] = cur1.x; W[5] = cur1.y; W[6] = cur1.z; W[7] = cur1.w;
        {   /* look-ahead words: the next lane's first chunk; lane 31 takes lane 0's next-iteration chunk */
            const uint32_t cw[4] = {cur0.x, cur0.y, cur0.z, cur0.w};
            const uint32_t nw[4] = {nxt0.x, nxt0.y, nxt0.z, nxt0.w};
#pragma unroll
            for (int k = 0; k < NW; k++) {
                uint32_t a = __shfl_down_sync(kFull, cw[k], 1);
                uint32_t b = __shfl_sync(kFull, nw[k], 0);
                W[8 + k] = (lane == 31) ? b : a;
            }
        }
        /* One bitmap probe per position.  The hit bit is shifted into `acc` with a multiply-add (FMA pipe, which
           has room; c.two == 2 is opaque to the compiler), so probe i ends up in bit kProbes-1-i.  Blocked Bloom,
           k = 2: both bits of the gram must be set in its word (wrap shifts use the low 5 bits of the hash). */
        uint32_t acc = 0;
#pragma unroll
        for (int t = 0; t < kFChunk; t += STRIDE) {
            uint32_t h = 0, ha;
            if (WIDE) {                          /* 64-bit products: low half = hash1, high half -> first bit */
                unsigned long long hw = 0;
#pragma unroll
                for (int k = 0; k < NW; k++) {
                    const int wi = (t >> 2) + k;
                    const uint32_t w = ((t & 3) == 0) ? W[wi] : __funnelshift_r(W[wi], W[wi + 1], (t & 3) * 8);
                    hw = mad_wide(w, mul[k], hw);
                }
                h = (uint32_t)hw;
                ha = (uint32_t)(hw >> 32);
            } else {
#pragma unroll
                for (int k = 0; k < NW; k++) {
                    const int wi = (t >> 2) + k;
                    const uint32_t w = ((t & 3) == 0) ? W[wi] : __funnelshift_r(W[wi], W[wi + 1], (t & 3) * 8);
                    h += w * mul[k];
                }
                ha = h >> c.sh_bit;
            }
            const uint32_t word = lds_word(__umulhi(h, c.mul_word) * c.four + c.sbm);
            const uint32_t both = __funnelshift_r(word, 0u, ha) & __funnelshift_r(word, 0u, h) & 1u;
            acc = acc * c.two + both;
        }
        uint32_t hits = __brev(acc) >> (32 - kProbes);                             /* bit i = probe i */
        if (GUARD && pos0 + kFChunk > c.seg_len) {                                 /* probes that start past the segment */
            const int valid = (pos0 >= c.seg_len) ? 0 : ((int)(c.seg_len - pos0) + STRIDE - 1) / STRIDE;
            hits = (valid <= 0) ? 0u : ((valid >= 32) ? hits : (hits & ((1u << valid) - 1u)));
        }
        /* queue the survivors (ballot-ranked append into the warp's ring) with hash2 of their gram */
        unsigned any = __ballot_sync(kFull, hits != 0);
        while (any) {
            bool keep = false;
            uint2 cand = make_uint2(0, 0);
            if (hits) {
                const int t = (__ffs(hits) - 1) * STRIDE;                          /* byte offset inside the lane's 32 */
                const uint32_t x = pos0 + t;
                hits &= hits - 1;
                /* hash2 of the gram, from the words still in registers (8-way select on t / 4) */
                const int g = t >> 2, sh = (t & 3) * 8;
                uint32_t tag = 0, w0 = sel8(W, 0, g);
#pragma unroll
                for (int k = 0; k < NW; k++) {
                    const uint32_t w1 = sel8(W, k + 1, g);
                    tag += __funnelshift_r(w0, w1, sh) * mul2[k];
                    w0 = w1;
                }
                tag |= 1u;
                /* stage 2: second bitmap (hash2), also in shared memory; only its survivors are queued */
                const uint32_t word2 = lds_word((tag >> c.sh_word2) * 4u + c.sbm2);
                keep = (__funnelshift_r(word2, 0u, tag >> c.sh_bit2) & 1u) != 0;
                if (keep && c.log3) {                                              /* stage 3 (large key sets only): bitmap in L2 */
                    const uint32_t i3 = (tag * ACB_S3_MIX) >> (32 - c.log3);
                    keep = ((__ldg(c.bm3 + (i3 >> 5)) >> (i3 & 31)) & 1u) != 0;
                }
                cand = make_uint2(x, tag);
            }
            const unsigned km = __ballot_sync(kFull, keep);
            if (keep) c.queue[(qtail + __popc(km & c.lt_mask)) & (kQueueCap - 1)] = cand;
            qtail += __popc(km);
            if (qtail - qhead > kQueueCap - 32) {                                  /* ring nearly full: spill 32 to the global list */
                __syncwarp();
                unsigned long long g = 0;
                if (lane == 0) g = atomicAdd(c.cand_count, 32ULL);
                g = __shfl_sync(kFull, g, 0);
                if (g + lane < c.cand_cap) c.cand[g + lane] = c.queue[(qhead + lane) & (kQueueCap - 1)];
                qhead += 32;
                __syncwarp();
            }
            any = __ballot_sync(kFull, hits != 0);
        }
        cur0 = nxt0;
        cur1 = nxt1;
    }
}

template <int NW, int STRIDE, bool WIDE>
__global__ void __launch_bounds__(kThreads, 1) acb_filter_kernel(const __grid_constant__ ScanParams p) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int nwords = 1 << (p.log1 - 5);
    uint32_t *s_bm = smem;
    uint2 *s_queue = reinterpret_cast<uint2 *>(s_bm + nwords);           /* kWarps * kQueueCap */
    acb_match *s_stage = reinterpret_cast<acb_match *>(s_queue + kWarps * kQueueCap);   /* kWarps * kStageCap */
    int *s_cnt = reinterpret_cast<int *>(s_stage + kWarps * kStageCap);  /* kWarps */

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    ACB_STAMP(0);
    {   /* both bitmaps -> shared memory (stage 1 in the first 7/8 of the words, stage 2 in the last 1/8), with
           cp.async so that all of a thread's 16-byte pieces are in flight at once instead of one L2 round trip each */
        const int n1 = 7 * (nwords / 8);
        const uint4 *src1 = reinterpret_cast<const uint4 *>(p.bm1), *src2 = reinterpret_cast<const uint4 *>(p.bm2);
        const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(s_bm);
        for (int i = tid; i < n1 / 4; i += kThreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sdst + 16u * i), "l"(src1 + i));
        for (int i = tid; i < (nwords - n1) / 4; i += kThreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sdst + 16u * (n1 / 4 + i)), "l"(src2 + i));
        asm volatile("cp.async.commit_group;");
        if (tid < kWarps) s_cnt[tid] = 0;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    ACB_STAMP(1);
    WarpResolve wr;
    wr.queue = s_queue + warp * kQueueCap;
    wr.ws.buf = s_stage + warp * kStageCap;
    wr.ws.cnt = s_cnt + warp;

    FilterCtx c;
    c.seg = p.hay + p.seg_begin;
    c.seg_len = (uint32_t)(p.seg_end - p.seg_begin);                     /* <= 2^31 */
    c.total_rel = p.total - p.seg_begin;
    c.sbm = (uint32_t)__cvta_generic_to_shared(s_bm);
    c.mul_word = 7u << (p.log1 - 8);
    c.bm3 = p.bm3;
    c.log3 = p.log3;
    c.sbm2 = c.sbm + 4u * (7u << (p.log1 - 8));
    c.sh_word2 = 40 - p.log1;
    c.sh_bit2 = 35 - p.log1;
    c.four = 4u + (uint32_t)(p.log1 >> 8);                               /* always 4 */
    c.two = 2u + (uint32_t)(p.log1 >> 8);                                /* always 2 */
    c.sh_bit = 32 - p.log1;
    c.lt_mask = (1u << lane) - 1u;
    c.lane = lane;
    c.queue = s_queue + warp * kQueueCap;
    c.cand = p.cand;
    c.cand_cap = p.cand_cap;
    c.cand_count = p.cand_count;
    uint32_t mul[NW], mul2[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) { mul[k] = p.mul1[k]; mul2[k] = p.mul2[k]; }
    int qhead = 0, qtail = 0;                    /* ring positions; the queue persists across work units */
    /* units below this index need no bounds checks: unit, look-ahead and gram re-reads stay inside both
       the segment and the buffer */
    long long interior_end = c.total_rel - 64 < (long long)c.seg_len ? c.total_rel - 64 : (long long)c.seg_len;
    const long long n_interior = interior_end < kBlockBytes ? 0 : interior_end / kBlockBytes;

    bool first_unit = true;
    for (;;) {
        unsigned int blk = 0;
        if (lane == 0) blk = atomicAdd(p.work_ctr, 1u);
        blk = __shfl_sync(kFull, blk, 0);
        if ((long long)blk >= p.n_blocks) break;
        if (first_unit) { ACB_STAMP(2); first_unit = false; }
        const uint32_t rel0 = blk * (uint32_t)kBlockBytes;
        if ((long long)blk < n_interior) filter_unit<NW, STRIDE, WIDE, false>(c, mul, mul2, rel0, qhead, qtail);
        else filter_unit<NW, STRIDE, WIDE, true>(c, mul, mul2, rel0, qhead, qtail);
        if (qtail - qhead >= 32) {
            if (p.inline_resolve) drain_queue(p, wr, qhead, qtail, false, lane);
            else spill_queue(c, qhead, qtail, false);
        }
    }
    ACB_STAMP(3);
    if (p.inline_resolve) drain_queue(p, wr, qhead, qtail, true, lane);  /* leftovers */
    else spill_queue(c, qhead, qtail, true);
    ACB_STAMP(4);
    /* the last CTA to leave re-arms the work counter, so a launch needs no memset before it */
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        unsigned int done = atomicAdd(p.work_ctr + 1, 1u);
        if (done == gridDim.x - 1) {
            p.work_ctr[0] = 0u;
            p.work_ctr[1] = 0u;
            __threadfence();
        }
    }
    ACB_STAMP(5);
}

/* ------------------------------------------------------- the verify kernel */
/* Overflow path of stage 3: resolves the candidates that warps had to spill to the global list
 * (normally none).  Each candidate is resolved through the anchor table (open addressing, 32-byte slots):
 * UNIQUE anchor -> the single key that can match is compared with the text;
 * MULTI anchor  -> exact gram compare, then a trie walk from the root. */

__global__ void __launch_bounds__(kVerThreads) acb_verify_kernel(const __grid_constant__ ScanParams p) {
    __shared__ acb_match s_stage[kVerWarps * kStageCap];
    __shared__ int s_cnt[kVerWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < kVerWarps) s_cnt[tid] = 0;
    __syncthreads();
    WarpStage ws;
    ws.buf = s_stage + warp * kStageCap;
    ws.cnt = s_cnt + warp;
    const unsigned long long found = *p.cand_count;
    const unsigned long long n = found < p.cand_cap ? found : p.cand_cap;
    const unsigned long long gwarp = (unsigned long long)blockIdx.x * kVerWarps + warp;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * kVerWarps;
    for (unsigned long long w0 = gwarp * 32; w0 < n; w0 += nwarps * 32) {
        if (w0 + lane < n) resolve(p, ws, p.cand[w0 + lane]);
        flush_stage(p, ws, lane);
    }
    /* last CTA out: re-arm the candidate counter; flag an overflowed candidate list in *count */
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        unsigned int done = atomicAdd(p.work_ctr + 2, 1u);
        if (done == gridDim.x - 1) {
            if (found > p.cand_cap) *p.count = ~0ULL;       /* results incomplete: caller must retry */
            *p.cand_count = 0ULL;
            p.work_ctr[2] = 0u;
            __threadfence();
        }
    }
}

/* ---------------------------------------------------------- the DFA kernel */

__global__ void __launch_bounds__(kDfaThreads) acb_dfa_kernel(const ScanParams p) {
    const long long span = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long a = span * kDfaSpan;
    if (a >= p.total) return;
    const long long b = (a + kDfaSpan < p.total) ? a + kDfaSpan : p.total;
    long long h, hs, he;
    find_haystack(p, a, h, hs, he);
    /* with variable offsets, position a may sit in a run of empty haystacks: find_haystack
       returns the last h with offsets[h] <= a, which is the non-empty one containing a */
    long long i = a - p.max_key_bytes;               /* warm-up start, letter aligned */
    if (i < hs) i = hs;
    int32_t st = 0;
    const int L = p.L;
    for (; i < b; ++i) {
        while (i >= he) {                             /* crossed into the next haystack(s) */
            h += 1; hs = he;
            he = (p.offsets == nullptr) ? hs + p.stride_bytes : __ldg(p.offsets + h + 1);
            st = 0;
        }
        const long long col = (long long)__ldg(p.cls + p.hay[i]) * p.S;
        int32_t nx;
        while ((nx = __ldg(p.gto + col + st)) < 0 && st != 0) st = __ldg(p.fail + st);   /* src/trie.c:182-190 */
        st = (nx < 0) ? 0 : (nx & kIdMask);
        if (i >= a && st != 0 && ((i + 1 - hs) % L) == 0) {
            const int32_t o0 = __ldg(p.out_ptr + st), o1 = __ldg(p.out_ptr + st + 1);
            for (int32_t o = o0; o < o1; ++o) {
                const int32_t k = __ldg(p.out_idx + o);
                const long long kb = (long long)__ldg(p.key_len + k) * L;
                if (i + 1 - kb < hs) continue;        /* cannot happen (state resets at hs); defensive */
                unsigned long long g = atomicAdd(p.count, 1ULL);
                if (g < (unsigned long long)p.cap) {
                    acb_match m;
                    m.hay_id = (int32_t)h;
                    m.end_index = (int32_t)((i - hs + 1) / L - 1);
                    m.key_id = k;
                    p.out[g] = m;
                }
            }
        }
    }
}

/* ------------------------------------------------------- the iter_long kernel */
/* ACB_ALGO_LONG: the reference's longest-match iterator (src/AutomatonSearchIterLong.c:89-153) is a
 * sequential state machine per haystack (after every reported match it restarts from the root at the
 * match's last letter), so one lane replays it per haystack: trie edges only (`goto`, letter by letter),
 * letter-level fail links, and the reference's early return when a non-terminal state's fail state ends
 * a key (:122-126).  Records of one haystack come out in increasing end_index. */
__device__ __forceinline__ int32_t letter_step(const ScanParams &p, int32_t st, const uint8_t *letter) {
    for (int b = 0; b < p.L; b++) {
        const int32_t nx = __ldg(p.gto + (long long)__ldg(p.cls + letter[b]) * p.S + st);
        if (nx < 0) return -1;
        st = nx & kIdMask;
    }
    return st;
}

__global__ void __launch_bounds__(kDfaThreads) acb_long_kernel(const __grid_constant__ ScanParams p) {
    const long long h = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= p.n_hay) return;
    const long long hs = p.offsets ? __ldg(p.offsets + h) : h * p.stride_bytes;
    const long long he = p.offsets ? __ldg(p.offsets + h + 1) : hs + p.stride_bytes;
    const long long n = (he - hs) / p.L;                       /* letters */
    const uint8_t *text = p.hay + hs;
    int32_t state = 0, last_node = -1;
    long long index = -1, last_index = -1;
    for (;;) {
        if (last_node >= 0) {                                   /* return_output */
            unsigned long long g = atomicAdd(p.count, 1ULL);
            if (g < (unsigned long long)p.cap) {
                acb_match m;
                m.hay_id = (int32_t)h;
                m.end_index = (int32_t)last_index;
                m.key_id = __ldg(p.key_of + last_node);
                p.out[g] = m;
            }
            state = 0;                                          /* start over: no overlapped results */
            index = last_index;
            last_node = -1;
            last_index = -1;
        }
        index += 1;
        bool emit = false;
        while (index < n) {
            const int32_t nx = letter_step(p, state, text + index * p.L);
            if (nx >= 0) {
                if (__ldg(p.key_of + nx) >= 0) {
                    last_node = nx;
                    last_index = index;
                } else {
                    const int32_t fl = __ldg(p.letter_fail + nx);
                    if (fl > 0 && __ldg(p.key_of + fl) >= 0) { last_node = fl; last_index = index; emit = true; break; }
                }
                state = nx;
                index += 1;
            } else {
                if (last_node >= 0) { emit = true; break; }
                for (;;) {
                    state = __ldg(p.letter_fail + state);
                    if (state < 0) { state = 0; index += 1; break; }
                    if (letter_step(p, state, text + index * p.L) >= 0) break;
                }
            }
        }
        if (!emit && last_node < 0) break;                      /* StopIteration */
    }
}

} // namespace

/* ------------------------------------------------------------- the table */

struct acb_table {
    int device = 0;
    int sm_count = 0;
    int32_t S = 0, K = 0, L = 1, n_keys = 0, gram = 1, stride = 1, log1 = 13, log2 = 15, log3 = 0, logA = 10, filter_flags = 0;
    int32_t min_key_bytes = 0, max_key_bytes = 0;
    uint32_t mul1[ACB_MAX_WINDOWS], mul2[ACB_MAX_WINDOWS];
    uint8_t *d_cls = nullptr;
    int32_t *d_lfail = nullptr;
    int32_t *d_goto = nullptr, *d_fail = nullptr, *d_keyof = nullptr, *d_outptr = nullptr, *d_outidx = nullptr, *d_keylen = nullptr;
    uint32_t *d_bm1 = nullptr, *d_bm2 = nullptr, *d_bm3 = nullptr, *d_anchors = nullptr;
    unsigned int *d_work = nullptr;
    uint2 *d_cand = nullptr;                 /* candidate list (filter -> verify) */
    unsigned long long cand_cap = 0;
    unsigned long long *d_cand_count = nullptr;
    bool cand_worst_case = false;
    long long dev_bytes = 0;
    std::vector<int32_t> key_len;            /* host copy, for sorting records */
    /* workspace of acb_scan_host */
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint8_t *w_hay = nullptr; size_t w_hay_cap = 0;
    long long *w_off = nullptr; size_t w_off_cap = 0;
    acb_match *w_out = nullptr; size_t w_out_cap = 0;
    unsigned long long *w_count = nullptr;
    unsigned long long *h_count = nullptr;   /* pinned */
    acb_match *h_out = nullptr; size_t h_out_cap = 0;   /* pinned staging for the records */
    unsigned long long h_out_n = 0;                      /* records of the last scan held in h_out */
    void *d_sort = nullptr; size_t sort_cap = 0;         /* radix-sort scratch */
};

extern "C" int acb_device_count(int32_t *n) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { acb_set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e)); if (n) *n = 0; return ACB_ECUDA; }
    if (n) *n = c;
    return ACB_OK;
}

template <typename T>
static int upload(T **dst, const T *src, size_t n, long long &acc) {
    size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    bytes = (bytes + 15) & ~(size_t)15;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(dst), bytes));
    CUDA_TRY(cudaMemset(*dst, 0, bytes));
    if (n) CUDA_TRY(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    acc += (long long)bytes;
    return ACB_OK;
}

extern "C" void acb_table_free(acb_table *tb) {
    if (!tb) return;
    cudaSetDevice(tb->device);
    cudaFree(tb->d_lfail); cudaFree(tb->d_cls); cudaFree(tb->d_goto); cudaFree(tb->d_fail); cudaFree(tb->d_keyof);
    cudaFree(tb->d_outptr); cudaFree(tb->d_outidx); cudaFree(tb->d_keylen); cudaFree(tb->d_bm1); cudaFree(tb->d_bm2); cudaFree(tb->d_bm3); cudaFree(tb->d_anchors);
    cudaFree(tb->d_sort); cudaFree(tb->d_work); cudaFree(tb->d_cand); cudaFree(tb->d_cand_count); cudaFree(tb->w_hay); cudaFree(tb->w_off); cudaFree(tb->w_out); cudaFree(tb->w_count);
    if (tb->h_count) cudaFreeHost(tb->h_count);
    if (tb->h_out) cudaFreeHost(tb->h_out);
    if (tb->ev0) cudaEventDestroy(tb->ev0);
    if (tb->ev1) cudaEventDestroy(tb->ev1);
    if (tb->stream) cudaStreamDestroy(tb->stream);
    delete tb;
}

extern "C" int acb_table_upload(const acb_trie *t, int device, acb_table **out) {
    if (!t || !out) { acb_set_error("bad argument"); return ACB_EINVAL; }
    *out = nullptr;
    acb_flat_view f;
    int rc = acb_trie_flat_view(t, &f);
    if (rc != ACB_OK) return rc;
    if (f.n_states > kIdMask) { acb_set_error("too many states for the device table (%d)", f.n_states); return ACB_ERANGE; }
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { acb_set_error("no such CUDA device %d (have %d)", device, ndev); return ACB_ECUDA; }
    CUDA_TRY(cudaSetDevice(device));
    acb_table *tb = new (std::nothrow) acb_table();
    if (!tb) { acb_set_error("out of memory"); return ACB_ENOMEM; }
    tb->device = device;
    cudaDeviceProp prop;
    rc = ACB_OK;
    do {
        if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { acb_set_error("cudaGetDeviceProperties failed"); rc = ACB_ECUDA; break; }
        tb->sm_count = prop.multiProcessorCount;
        tb->S = f.n_states; tb->K = f.n_classes; tb->L = f.letter_bytes; tb->n_keys = f.n_keys;
        tb->gram = f.gram_bytes; tb->stride = f.stride; tb->log1 = f.log2_bits1; tb->log2 = f.log2_bits2; tb->log3 = f.log2_bits3; tb->logA = f.log2_anchor_slots; tb->filter_flags = f.filter_flags;
        tb->min_key_bytes = f.min_key_bytes; tb->max_key_bytes = f.max_key_bytes;
        acb_hash_multipliers(tb->gram, 1, tb->mul1);
        acb_hash_multipliers(tb->gram, 2, tb->mul2);
        /* goto entries get a flag bit when the child ends a key, saving a key_of lookup per step */
        std::vector<int32_t> flagged;
        try {
            tb->key_len.assign(f.key_len, f.key_len + f.n_keys);
            flagged.resize((size_t)f.n_classes * f.n_states);
        } catch (const std::exception &) {                   /* nothing may cross the C ABI */
            acb_set_error("out of host memory while staging the tables");
            rc = ACB_ENOMEM;
            break;
        }
        for (size_t i = 0; i < flagged.size(); i++) {
            int32_t v = f.goto_cm[i];
            flagged[i] = (v >= 0 && f.key_of[v] >= 0) ? (v | kTermBit) : v;
        }
        if ((rc = upload(&tb->d_cls, f.byte_class, 256, tb->dev_bytes))) break;
        if ((rc = upload(&tb->d_goto, flagged.data(), flagged.size(), tb->dev_bytes))) break;
        if ((rc = upload(&tb->d_fail, f.fail, (size_t)f.n_states, tb->dev_bytes))) break;
        if ((rc = upload(&tb->d_lfail, f.letter_fail, (size_t)f.n_states, tb->dev_bytes))) break;
        if ((rc = upload(&tb->d_keyof, f.key_of, (size_t)f.n_states, tb->dev_bytes))) break;
        if ((rc = upload(&tb->d_outptr, f.out_ptr, (size_t)f.n_states + 1, tb->dev_bytes))) break;
        if ((rc = upload(&tb->d_outidx, f.out_idx, (size_t)f.out_ptr[f.n_states], tb->dev_bytes))) break;
        if ((rc = upload(&tb->d_keylen, f.key_len, (size_t)f.n_keys, tb->dev_bytes))) break;
        if ((rc = upload(&tb->d_bm1, f.bitmap1, (size_t)7 << (f.log2_bits1 - 8), tb->dev_bytes))) break;
        if ((rc = upload(&tb->d_bm2, f.bitmap2, (size_t)1 << (f.log2_bits1 - 8), tb->dev_bytes))) break;
        if ((rc = upload(&tb->d_bm3, f.bitmap3, f.log2_bits3 ? ((size_t)1 << (f.log2_bits3 - 5)) : 1, tb->dev_bytes))) break;
        if ((rc = upload(&tb->d_anchors, f.anchors, (size_t)8 << f.log2_anchor_slots, tb->dev_bytes))) break;
        unsigned int zero[4] = {0, 0, 0, 0};   /* work counters, re-armed by the kernels themselves */
        if ((rc = upload(&tb->d_work, zero, 4, tb->dev_bytes))) break;
        unsigned long long zero64 = 0;
        if ((rc = upload(&tb->d_cand_count, &zero64, 1, tb->dev_bytes))) break;
    } while (0);
    if (rc != ACB_OK) { acb_table_free(tb); return rc; }
    *out = tb;
    return ACB_OK;
}

extern "C" int64_t acb_table_device_bytes(const acb_table *tb) { return tb ? tb->dev_bytes : 0; }
extern "C" int acb_table_reserve_candidates(acb_table *tb, int worst_case) {
    if (!tb) return ACB_EINVAL;
    tb->cand_worst_case = worst_case != 0;
    return ACB_OK;
}
extern "C" int64_t acb_launch_count(void) { return g_launches.load(); }
extern "C" int acb_set_kernel_timing(int enabled) { g_timing.store(enabled ? 1 : 0); return ACB_OK; }
extern "C" float acb_last_kernel_ms(void) { return g_last_ms; }

/* ------------------------------------------------------------- launching */

static size_t filter_smem_bytes(int log1) {
    return ((size_t)1 << (log1 - 3)) + (size_t)kWarps * kQueueCap * sizeof(uint2) +
           (size_t)kWarps * kStageCap * sizeof(acb_match) + (size_t)kWarps * sizeof(int);
}

template <int NW, int STRIDE, bool WIDE>
static int launch_filter_w(const ScanParams &p, int grid, cudaStream_t s) {
    auto kern = acb_filter_kernel<NW, STRIDE, WIDE>;
    const size_t smem = filter_smem_bytes(p.log1);
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreads, smem, s>>>(p);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    return ACB_OK;
}

/* WIDE follows from the gram length (acb_hash_is_wide) */
template <int NW, int STRIDE>
static int launch_filter_t(const ScanParams &p, int grid, cudaStream_t s) {
    if (p.filter_flags & ACB_FILTER_WIDE) {
        if (p.gram != 4 * NW) { acb_set_error("WIDE filter with gram %d", p.gram); return ACB_EINVAL; }
        return launch_filter_w<NW, STRIDE, true>(p, grid, s);
    }
    return launch_filter_w<NW, STRIDE, false>(p, grid, s);
}

template <int NW>
static int launch_filter_s(const ScanParams &p, int stride, int grid, cudaStream_t s) {
    switch (stride) {
        case 1:  return launch_filter_t<NW, 1>(p, grid, s);
        case 2:  return launch_filter_t<NW, 2>(p, grid, s);
        case 4:  return launch_filter_t<NW, 4>(p, grid, s);
        case 8:  return launch_filter_t<NW, 8>(p, grid, s);
        case 16: return launch_filter_t<NW, 16>(p, grid, s);
    }
    acb_set_error("unsupported filter stride %d", stride);
    return ACB_EINVAL;
}

static int launch_filter(const ScanParams &p, int stride, int grid, cudaStream_t s) {
    switch ((p.gram + 3) / 4) {
        case 1: return launch_filter_s<1>(p, stride, grid, s);
        case 2: return launch_filter_s<2>(p, stride, grid, s);
        case 3: return launch_filter_s<3>(p, stride, grid, s);
        case 4: return launch_filter_s<4>(p, stride, grid, s);
    }
    acb_set_error("unsupported gram length %d", p.gram);
    return ACB_EINVAL;
}

/* candidate list capacity for a segment of `seg` bytes: 1/8 of the probe positions by default,
 * every probe position once a scan has reported an overflow (acb_scan_host retries that way) */
static int ensure_candidates(acb_table *tb, long long seg, bool worst_case) {
    unsigned long long probes = (unsigned long long)(seg / tb->stride + 1);
    unsigned long long want = worst_case ? probes : std::max<unsigned long long>(1ULL << 20, probes / 8);
    if (tb->d_cand && tb->cand_cap >= want) return ACB_OK;
    if (tb->d_cand) { cudaFree(tb->d_cand); tb->d_cand = nullptr; tb->cand_cap = 0; }
    CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&tb->d_cand), (size_t)want * sizeof(uint2)));
    tb->cand_cap = want;
    return ACB_OK;
}

extern "C" int acb_scan_device(acb_table *tb, const uint8_t *d_hay, int64_t total_bytes,
                               const int64_t *d_offsets, int64_t n_hay, int64_t stride_bytes,
                               acb_match *d_out, int64_t cap, int64_t *d_count, void *stream, int algo) {
    if (!tb || !d_count || total_bytes < 0 || n_hay < 0 || cap < 0 || (cap > 0 && !d_out)) { acb_set_error("bad argument"); return ACB_EINVAL; }
    if (n_hay > 0x7fffffffLL) { acb_set_error("more than 2^31-1 haystacks in one batch"); return ACB_ERANGE; }
    if (!d_offsets) {
        if (stride_bytes <= 0 || stride_bytes % tb->L || stride_bytes * n_hay != total_bytes) {
            acb_set_error("fixed-stride batch needs stride_bytes > 0, a multiple of letter_bytes, and n_hay*stride == total_bytes");
            return ACB_EINVAL;
        }
        if (stride_bytes / tb->L > 0x7fffffffLL) { acb_set_error("haystack longer than 2^31-1 letters"); return ACB_ERANGE; }
    }
    if (total_bytes == 0 || n_hay == 0) return ACB_OK;
    if (reinterpret_cast<uintptr_t>(d_hay) & 15) { acb_set_error("d_hay must be 16-byte aligned"); return ACB_EINVAL; }
    CUDA_TRY(cudaSetDevice(tb->device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);

    ScanParams p;
    memset(&p, 0, sizeof(p));
    p.hay = d_hay; p.total = total_bytes; p.offsets = reinterpret_cast<const long long *>(d_offsets);
    p.n_hay = n_hay; p.stride_bytes = stride_bytes;
    p.cls = tb->d_cls; p.gto = tb->d_goto; p.fail = tb->d_fail; p.letter_fail = tb->d_lfail; p.key_of = tb->d_keyof;
    p.out_ptr = tb->d_outptr; p.out_idx = tb->d_outidx; p.key_len = tb->d_keylen;
    p.S = tb->S; p.L = tb->L; p.gram = tb->gram; p.filter_flags = tb->filter_flags; p.max_key_bytes = tb->max_key_bytes;
    p.bm1 = tb->d_bm1; p.bm2 = tb->d_bm2; p.bm3 = tb->d_bm3; p.anchors = reinterpret_cast<const uint4 *>(tb->d_anchors);
    p.log1 = tb->log1; p.log2 = tb->log2; p.log3 = tb->log3; p.logA = tb->logA;
    memcpy(p.mul1, tb->mul1, sizeof(p.mul1));
    memcpy(p.mul2, tb->mul2, sizeof(p.mul2));
    p.out = d_out; p.cap = cap; p.count = reinterpret_cast<unsigned long long *>(d_count);
    p.work_ctr = tb->d_work;
    p.stride_shift = -1;
    if (!d_offsets) for (int b = 0; b < 62; b++) if ((1LL << b) == stride_bytes) p.stride_shift = b;
    p.letter_shift = tb->L == 4 ? 2 : (tb->L == 2 ? 1 : 0);
    {
        static int inl = -1;
        if (inl < 0) { const char *e = getenv("ACB_INLINE_RESOLVE"); inl = e ? atoi(e) : 1; }
        p.inline_resolve = inl;
    }

    if (algo == ACB_ALGO_AUTO) algo = ACB_ALGO_FILTER;
    if (tb->n_keys == 0) return ACB_OK;                     /* empty key set: nothing can match */
    const bool timing = g_timing.load() != 0;
    if (timing) {
        if (!tb->ev0) { CUDA_TRY(cudaEventCreate(&tb->ev0)); CUDA_TRY(cudaEventCreate(&tb->ev1)); }
        CUDA_TRY(cudaEventRecord(tb->ev0, s));
    }
    if (algo == ACB_ALGO_FILTER) {
        int rc = ensure_candidates(tb, std::min<long long>(total_bytes, kSegBytes), tb->cand_worst_case);
        if (rc != ACB_OK) return rc;
        p.cand = tb->d_cand; p.cand_cap = tb->cand_cap; p.cand_count = tb->d_cand_count;
        for (long long seg = 0; seg < total_bytes; seg += kSegBytes) {
            p.seg_begin = seg;
            p.seg_end = std::min<long long>(seg + kSegBytes, total_bytes);
            p.n_blocks = (p.seg_end - p.seg_begin + kBlockBytes - 1) / kBlockBytes;
            int grid = (int)std::min<long long>(tb->sm_count, p.n_blocks);
            unsigned long long *d_tl = nullptr;
            const size_t n_tl = (size_t)grid * kWarps * 6;
            if (getenv("ACB_TIMELINE")) {                       /* diagnostic: per-warp phase timestamps */
                cudaMalloc(reinterpret_cast<void **>(&d_tl), n_tl * 8);
                cudaMemset(d_tl, 0, n_tl * 8);
                p.timeline = d_tl;
            }
            rc = launch_filter(p, tb->stride, grid, s);
            if (rc != ACB_OK) return rc;
            if (d_tl) {
                std::vector<unsigned long long> tl(n_tl);
                cudaStreamSynchronize(s);
                cudaMemcpy(tl.data(), d_tl, n_tl * 8, cudaMemcpyDeviceToHost);
                cudaFree(d_tl);
                p.timeline = nullptr;
                unsigned long long t0 = ~0ULL;
                for (size_t w = 0; w < n_tl / 6; w++) if (tl[w * 6]) t0 = std::min(t0, tl[w * 6]);
                const char *nm[6] = {"start", "bitmap ready", "first unit claimed", "stream done", "drained", "exit"};
                for (int k = 0; k < 6; k++) {
                    std::vector<double> v;
                    for (size_t w = 0; w < n_tl / 6; w++) if (tl[w * 6 + k]) v.push_back((double)(tl[w * 6 + k] - t0) / 1000.0);
                    if (v.empty()) continue;
                    std::sort(v.begin(), v.end());
                    fprintf(stderr, "[timeline] %-20s n=%6zu  min %8.2f  p50 %8.2f  p90 %8.2f  max %8.2f us\n", nm[k], v.size(),
                            v.front(), v[v.size() / 2], v[v.size() * 9 / 10], v.back());
                }
            }
            if (getenv("ACB_DEBUG")) {                          /* diagnostic: size of the spilled candidate list */
                unsigned long long cc = 0, mc = 0;
                cudaStreamSynchronize(s);
                cudaMemcpy(&cc, tb->d_cand_count, sizeof(cc), cudaMemcpyDeviceToHost);
                cudaMemcpy(&mc, p.count, sizeof(mc), cudaMemcpyDeviceToHost);
                fprintf(stderr, "[acb_scan_device] segment %lld..%lld: %llu candidates spilled (cap %llu), %llu matches so far, gram %d stride %d log3 %d\n",
                        p.seg_begin, p.seg_end, cc, (unsigned long long)p.cand_cap, mc, p.gram, tb->stride, p.log3);
            }
            acb_verify_kernel<<<p.inline_resolve ? tb->sm_count : tb->sm_count * 6, kVerThreads, 0, s>>>(p);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { acb_set_error("verify kernel launch failed: %s", cudaGetErrorString(e)); return ACB_ECUDA; }
            g_launches.fetch_add(1);
        }
    } else if (algo == ACB_ALGO_DFA) {
        long long spans = (total_bytes + kDfaSpan - 1) / kDfaSpan;
        long long grid = (spans + kDfaThreads - 1) / kDfaThreads;
        if (grid > 0x7fffffffLL) { acb_set_error("batch too large for one launch"); return ACB_ERANGE; }
        acb_dfa_kernel<<<(unsigned)grid, kDfaThreads, 0, s>>>(p);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { acb_set_error("DFA kernel launch failed: %s", cudaGetErrorString(e)); return ACB_ECUDA; }
        g_launches.fetch_add(1);
    } else if (algo == ACB_ALGO_LONG) {
        long long grid = (n_hay + kDfaThreads - 1) / kDfaThreads;
        acb_long_kernel<<<(unsigned)grid, kDfaThreads, 0, s>>>(p);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { acb_set_error("iter_long kernel launch failed: %s", cudaGetErrorString(e)); return ACB_ECUDA; }
        g_launches.fetch_add(1);
    } else {
        acb_set_error("unknown algo %d", algo);
        return ACB_EINVAL;
    }
    if (timing) {
        CUDA_TRY(cudaEventRecord(tb->ev1, s));
        CUDA_TRY(cudaEventSynchronize(tb->ev1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, tb->ev0, tb->ev1));
        g_last_ms = ms;
    }
    return ACB_OK;
}

extern "C" int acb_copy_records(acb_table *tb, acb_match *out, int64_t n) {
    if (!tb || n < 0 || (n && !out) || (unsigned long long)n > tb->h_out_n) { acb_set_error("bad argument"); return ACB_EINVAL; }
    if (n) memcpy(out, tb->h_out, (size_t)n * sizeof(acb_match));
    return ACB_OK;
}

/* Zero-copy hand-over of the records.  The pinned staging buffer of the last acb_scan_host can be taken by the
 * caller (no memcpy, no page faults of a fresh destination); it comes back through acb_release_records into a
 * small process-wide pool from which the next scan that needs a staging buffer is served.  A buffer that is
 * never released is simply not reused. */
namespace {
struct PinnedBuf { acb_match *p; size_t cap; };
std::mutex g_pool_mu;
std::vector<PinnedBuf> g_pool;
constexpr size_t kPoolMax = 4;

acb_match *pool_take(size_t need, size_t *cap) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    size_t best = g_pool.size();
    for (size_t i = 0; i < g_pool.size(); i++)
        if (g_pool[i].cap >= need && (best == g_pool.size() || g_pool[i].cap < g_pool[best].cap)) best = i;
    if (best == g_pool.size()) return nullptr;
    acb_match *p = g_pool[best].p;
    *cap = g_pool[best].cap;
    g_pool.erase(g_pool.begin() + (long)best);
    return p;
}
} // namespace

extern "C" int acb_take_records(acb_table *tb, acb_match **ptr, int64_t *n, int64_t *cap) {
    if (!tb || !ptr || !n || !cap) { acb_set_error("bad argument"); return ACB_EINVAL; }
    *ptr = nullptr; *n = 0; *cap = 0;
    if (tb->h_out_n == 0 || !tb->h_out) return ACB_OK;       /* nothing to hand over */
    *ptr = tb->h_out;
    *n = (int64_t)tb->h_out_n;
    *cap = (int64_t)tb->h_out_cap;
    tb->h_out = nullptr;                                     /* the next scan gets a buffer from the pool or a new one */
    tb->h_out_cap = 0;
    tb->h_out_n = 0;
    return ACB_OK;
}

extern "C" void acb_release_records(acb_match *ptr, int64_t cap) {
    if (!ptr || cap <= 0) return;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (g_pool.size() < kPoolMax) { g_pool.push_back({ptr, (size_t)cap}); return; }
    }
    cudaFreeHost(ptr);
}

/* ------------------------------------------------------------ record sort */
/* Reference order: haystack, then end_index ascending, then longest key first.  One
 * 64-bit radix key per record: hay_id | end_index | (max_len - len), packed into the fewest bits. */
namespace {
__global__ void acb_sortkey_kernel(const acb_match *rec, long long n, const int32_t *key_len, int be, int bl,
                                   int max_len, unsigned long long *keys) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const acb_match m = rec[i];
    const unsigned long long inv = (unsigned long long)(max_len - __ldg(key_len + m.key_id));
    keys[i] = ((unsigned long long)(uint32_t)m.hay_id << (be + bl)) | ((unsigned long long)(uint32_t)m.end_index << bl) | inv;
}
int bits_for(unsigned long long v) { int b = 1; while (b < 64 && (v >> b)) b++; return b; }
} // namespace

extern "C" int acb_sort_matches_device(acb_table *tb, acb_match *d_records, int64_t n, int64_t n_hay,
                                       int64_t max_hay_letters, void *stream) {
    if (!tb || n < 0 || (n && !d_records)) { acb_set_error("bad argument"); return ACB_EINVAL; }
    if (n <= 1) return ACB_OK;
    CUDA_TRY(cudaSetDevice(tb->device));
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int max_len = tb->max_key_bytes / tb->L;
    const int bh = bits_for((unsigned long long)std::max<int64_t>(n_hay - 1, 1));
    const int be = bits_for((unsigned long long)std::max<int64_t>(max_hay_letters, 1));
    const int bl = bits_for((unsigned long long)max_len);
    if (bh + be + bl > 64) { acb_set_error("sort key does not fit 64 bits (%d+%d+%d)", bh, be, bl); return ACB_ERANGE; }
    size_t temp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const acb_match *)nullptr, (acb_match *)nullptr, (int)n, 0, bh + be + bl, s);
    const size_t need = (size_t)n * (2 * sizeof(unsigned long long) + sizeof(acb_match)) + temp + 1024;
    if (tb->sort_cap < need) {
        if (tb->d_sort) { cudaFree(tb->d_sort); tb->d_sort = nullptr; tb->sort_cap = 0; }
        CUDA_TRY(cudaMalloc(&tb->d_sort, need + need / 4));
        tb->sort_cap = need + need / 4;
    }
    if (n > 0x7fffffffLL) { acb_set_error("too many records to sort on the device"); return ACB_ERANGE; }
    char *base = reinterpret_cast<char *>(tb->d_sort);
    unsigned long long *k0 = reinterpret_cast<unsigned long long *>(base);
    unsigned long long *k1 = k0 + n;
    acb_match *r1 = reinterpret_cast<acb_match *>(k1 + n);
    void *tmp = reinterpret_cast<void *>((reinterpret_cast<uintptr_t>(r1 + n) + 255) & ~(uintptr_t)255);
    acb_sortkey_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_records, n, tb->d_keylen, be, bl, max_len, k0);
    CUDA_TRY(cudaGetLastError());
    g_launches.fetch_add(1);
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp, temp, k0, k1, d_records, r1, (int)n, 0, bh + be + bl, s));
    CUDA_TRY(cudaMemcpyAsync(d_records, r1, (size_t)n * sizeof(acb_match), cudaMemcpyDeviceToDevice, s));
    return ACB_OK;
}

/* ------------------------------------------------------- host-buffer scan */

template <typename T>
static int ensure(T **buf, size_t *cap, size_t need) {
    if (*cap >= need && *buf) return ACB_OK;
    if (*buf) { cudaFree(*buf); *buf = nullptr; *cap = 0; }
    size_t n = std::max<size_t>(need + need / 4, 1024);
    CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(buf), n * sizeof(T)));
    *cap = n;
    return ACB_OK;
}

extern "C" int acb_scan_host(acb_table *tb, const uint8_t *hay, int64_t total_bytes,
                             const int64_t *offsets, int64_t n_hay, int64_t stride_bytes,
                             acb_match *out, int64_t cap, int64_t *n_found, int algo, int sort) {
    if (!tb || !n_found || total_bytes < 0 || n_hay < 0 || cap < 0) { acb_set_error("bad argument"); return ACB_EINVAL; }
    *n_found = 0;
    tb->h_out_n = 0;
    if (total_bytes == 0 || n_hay == 0) return ACB_OK;
    CUDA_TRY(cudaSetDevice(tb->device));
    if (!tb->stream) CUDA_TRY(cudaStreamCreateWithFlags(&tb->stream, cudaStreamNonBlocking));
    if (!tb->w_count) CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&tb->w_count), sizeof(unsigned long long)));
    if (!tb->h_count) CUDA_TRY(cudaMallocHost(reinterpret_cast<void **>(&tb->h_count), sizeof(unsigned long long)));
    int rc;
    if ((rc = ensure(&tb->w_hay, &tb->w_hay_cap, (size_t)total_bytes + 64))) return rc;
    if (offsets && (rc = ensure(&tb->w_off, &tb->w_off_cap, (size_t)n_hay + 1))) return rc;
    if ((rc = ensure(&tb->w_out, &tb->w_out_cap, (size_t)std::max<int64_t>(cap, 1)))) return rc;
    cudaStream_t s = tb->stream;
    static const bool trace = getenv("ACB_TRACE") != nullptr;          /* phase timing to stderr (adds syncs) */
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = trace ? now() : 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
    CUDA_TRY(cudaMemcpyAsync(tb->w_hay, hay, (size_t)total_bytes, cudaMemcpyHostToDevice, s));
    if (offsets) CUDA_TRY(cudaMemcpyAsync(tb->w_off, offsets, (size_t)(n_hay + 1) * sizeof(long long), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemsetAsync(tb->w_count, 0, sizeof(unsigned long long), s));
    if (trace) { cudaStreamSynchronize(s); t1 = now(); }
    rc = acb_scan_device(tb, tb->w_hay, total_bytes, offsets ? reinterpret_cast<const int64_t *>(tb->w_off) : nullptr,
                         n_hay, stride_bytes, tb->w_out, cap, reinterpret_cast<int64_t *>(tb->w_count), s, algo);
    if (rc != ACB_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(tb->h_count, tb->w_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (trace) t2 = now();
    unsigned long long n = *tb->h_count;
    if (n == ~0ULL) {                                           /* candidate list overflowed: redo with room for every probe */
        if (tb->cand_worst_case) { acb_set_error("candidate list overflow even at worst-case capacity"); return ACB_ECUDA; }
        tb->cand_worst_case = true;
        return acb_scan_host(tb, hay, total_bytes, offsets, n_hay, stride_bytes, out, cap, n_found, algo, sort);
    }
    *n_found = (int64_t)n;
    if (n > (unsigned long long)cap) {
        acb_set_error("match buffer too small: %llu matches, capacity %lld", n, (long long)cap);
        return ACB_EOVERFLOW;
    }
    if (n) {
        if (tb->h_out_cap < n) {                               /* pinned staging: D2H at full PCIe rate */
            if (tb->h_out) { acb_release_records(tb->h_out, (int64_t)tb->h_out_cap); tb->h_out = nullptr; tb->h_out_cap = 0; }
            size_t got = 0;
            if (acb_match *p = pool_take((size_t)n, &got)) {  /* a buffer some caller has given back */
                tb->h_out = p;
                tb->h_out_cap = got;
            } else {
                size_t want = (size_t)n + (size_t)n / 4 + 1024;
                CUDA_TRY(cudaMallocHost(reinterpret_cast<void **>(&tb->h_out), want * sizeof(acb_match)));
                tb->h_out_cap = want;
            }
        }
        bool host_sort = sort != 0;
        if (sort) {                                            /* radix sort on the device when the key fits 64 bits */
            const int64_t max_letters = (offsets ? total_bytes : stride_bytes) / tb->L;
            if (acb_sort_matches_device(tb, tb->w_out, (int64_t)n, n_hay, max_letters, s) == ACB_OK) host_sort = false;
        }
        if (trace) { cudaStreamSynchronize(s); t3 = now(); }
        CUDA_TRY(cudaMemcpyAsync(tb->h_out, tb->w_out, (size_t)n * sizeof(acb_match), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        if (host_sort) {
            const int32_t *kl = tb->key_len.data();
            std::sort(tb->h_out, tb->h_out + n, [kl](const acb_match &a, const acb_match &b) {
                if (a.hay_id != b.hay_id) return a.hay_id < b.hay_id;
                if (a.end_index != b.end_index) return a.end_index < b.end_index;
                return kl[a.key_id] > kl[b.key_id];            /* longest first: fail-chain order */
            });
        }
        if (out) memcpy(out, tb->h_out, (size_t)n * sizeof(acb_match));   /* out == NULL: fetch with acb_copy_records */
    }
    tb->h_out_n = n;
    if (trace) {
        t4 = now();
        fprintf(stderr, "[acb_scan_host] %lld B: h2d %.3f ms, scan %.3f ms, sort %.3f ms, d2h+copy %.3f ms (%llu records)\n",
                (long long)total_bytes, t1 - t0, t2 - t1, t3 ? t3 - t2 : 0.0, t3 ? t4 - t3 : t4 - t2, n);
    }
    return ACB_OK;
}
