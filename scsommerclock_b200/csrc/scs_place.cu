// scs_place.cu -- mutation placement (reference: run_PT_test.py:386-448,
// map_mutations_gt + _normalize_log_probs :368-383) as a two-pass tcgen05 GEMM.
//
// For SNV i and candidate row b of the clade matrix S (2n-1 tree nodes + one
// empty "FP" row) the reference computes
//     probs[i,b] = sum_c  S[b,c]*log(1-FN)*x + (1-S[b,c])*log(FP)*x
//                       + (1-S[b,c])*log(1-FP)*(1-x) + S[b,c]*log(FN)*(1-x)
// over the observed cells c (x in {0,1}, NaN cells dropped by nansum), then
// M[i,:] = softmax_b(probs[i,:]) and Y[b] = sum_i M[i,b].  Everything that does
// not depend on b cancels in the softmax, leaving
//     logit[i,b] = a1 * C1[i,b] + a0 * C0[i,b],
//     C1 = X1 . S^T,  C0 = X0 . S^T   (integer counts),
//     a1 = log((1-FN)/FP),  a0 = log(FN/(1-FP)),
// where X1/X0 are the 0/1 indicator planes "cell observed mutated / observed
// wild type".  C1 and C0 are exact integer GEMMs: unsigned 8-bit operands with
// s32 accumulators in TMEM (tcgen05.mma kind::i8), so the only floating-point
// rounding in the whole placement is in the epilogue, which works on *exact
// integer differences* against a per-row reference column.
//
// Pass 1 emits, per (row, 64-column half tile), the reference counts of the best
// column and sum_b exp2(logit - logit_ref).  A small merge kernel turns those into
// one (reference, log2-sum) record per SNV.  Pass 2 recomputes the tiles and
// accumulates softmax column sums in registers; the SNV x branch matrix never
// exists in HBM.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>

#include "scs_internal.h"
#include "scs_ptx.cuh"

namespace scs {

// ------------------------------------------------------------------ tiling
constexpr int BLOCK_M = 128;  // SNV rows per tile (= TMEM lanes)
constexpr int BLOCK_N = 128;  // clade rows (GEMM columns) per tile
constexpr int BLOCK_K = 128;  // cells per pipeline stage (bytes, 8-bit operands)
constexpr int UMMA_K = 32;    // cells per tcgen05.mma (8-bit kinds)
constexpr int TILE_BYTES = BLOCK_M * BLOCK_K;  // one 128 x 128 operand tile
constexpr int NUM_EPI_WARPS = 8;  // 4 TMEM lane quarters x 2 column halves
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BARRIER_BYTES = 256;
constexpr int SMEM_RED_BYTES = 4 * BLOCK_N * 4;

// Two ways to get the two count matrices out of the tensor cores:
//  MODE_PLANES  two unsigned-8-bit 0/1 planes (mutated / wild type), two kind::i8 MMAs
//               per K step, s32 accumulators: exact by construction, any cell count.
//  MODE_RADIX   one FP8 (e5m2) plane holding 1 for "mutated" and 4096 for "wild type",
//               one kind::f8f6f4 MMA per K step, f32 accumulator = C1 + 4096 * C0.  Both
//               counts stay below 4096 per accumulator because K is cut into chunks of
//               31 * 128 = 3968 cells, each with its own TMEM columns, so every partial sum
//               is an integer < 2^24 and exact in fp32.  Half the tensor work and half the
//               genotype bytes; up to 4 chunks (15,872 cells).
enum { MODE_PLANES = 0, MODE_RADIX = 1 };
constexpr int RADIX_KB_PER_CHUNK = 31;
constexpr int RADIX_MAX_CHUNKS = 4;
constexpr uint8_t E5M2_ONE = 0x3C;    // 1.0
constexpr uint8_t E5M2_4096 = 0x6C;   // 2^12

template <int MODE>
struct ModeCfg {
    static constexpr int A_TILES = MODE == MODE_PLANES ? 2 : 1;
    static constexpr int NUM_STAGES = MODE == MODE_PLANES ? 4 : 6;
    static constexpr int STAGE_BYTES = (A_TILES + 1) * TILE_BYTES;
    static constexpr int SMEM_TOTAL = NUM_STAGES * STAGE_BYTES + SMEM_BARRIER_BYTES + SMEM_RED_BYTES + 1024;
    // pass 1 with the count cache: 8 x 4 KB staging buffers (one per epilogue warp) in place of `red`
    static constexpr int SMEM_CACHE = NUM_STAGES * STAGE_BYTES + SMEM_BARRIER_BYTES + NUM_EPI_WARPS * 4096 + 1024;
};
static_assert(ModeCfg<0>::SMEM_CACHE <= 232448 && ModeCfg<1>::SMEM_CACHE <= 232448, "count-cache staging does not fit 227 KB");

// Instruction descriptors (bit layout of the tcgen05 instruction descriptor): D format at
// [4,6) (1 = f32, 2 = s32), A format [7,10), B format [10,13), both K-major (bits 15, 16 = 0),
// N >> 3 at [17,23), M >> 4 at [24,29).
constexpr uint32_t IDESC_I8 = (2u << 4) | (uint32_t(BLOCK_N >> 3) << 17) | (uint32_t(BLOCK_M >> 4) << 24);            // u8 x u8 -> s32
constexpr uint32_t IDESC_E5M2 = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(BLOCK_N >> 3) << 17) | (uint32_t(BLOCK_M >> 4) << 24);   // e5m2 x e5m2 -> f32

enum { PASS_DUMP = 0, PASS_STATS = 1, PASS_COLSUM = 2 };

// One task = one column tile against a run of row blocks of the same group (dataset).
struct TaskEntry {
    int j;          // global column tile
    int r0, r1;     // global row blocks [r0, r1)
    int jl;         // column tile index inside its group (indexes the pass-1 partials)
    int run_local;  // run index inside its group (indexes the pass-2 partials)
    int nvalid;     // valid clade rows in this tile (1..128)
    int group;
    int pad;
};
// Per-group (per-dataset) error-model constants, log2 units (see fill_constants).
struct GroupConst {
    float a1h, a1l, a0h, a0l;  // split so that a?h * (count difference) is exact
    float a1f, a0f;            // plain fp32 copies, used only to rank columns
    float pad0, pad1;
};

struct PlaceParams {
    int M, N;            // single-group problems: valid SNV rows / valid clade rows (2-CTA kernel)
    int M_pad, N_pad;    // padded totals over all groups
    int num_k_blocks;
    int num_row_blocks, num_col_tiles;
    int run_len, num_runs, sb_width, num_tasks;
    int l2_hints;              // 1: clade tiles evict-last, genotype tiles evict-first
    int nchunk;                // radix mode: K chunks (accumulators) per tile
    const TaskEntry* tasks;    // [num_tasks], the 1-CTA kernel's schedule
    const GroupConst* gconst;  // [groups]
    uint2* stats;              // pass 1 out: [num_col_tiles*2][M_pad] {ref counts, sum}
    const uint2* rowref;       // pass 2 in : [M_pad] {ref counts, log2 sum}
    float* colpart;            // pass 2 out: [num_runs][N_pad]
    int* dump;                 // debug     : [M_pad][N_pad][2] raw counts
    uint32_t* cache;           // pass 1 out (optional): packed counts C1 | C0 << 16, 64 KB per (row block, column tile) (see cache_store_and_stats)
};

__device__ __forceinline__ float count_to_float(uint32_t c) {
    // exact for 0 <= c < 2^23: plant the integer in the mantissa of 2^23
    return __uint_as_float(c | 0x4B000000u) - 8388608.0f;
}

// 32 consecutive columns of this thread's row: number of clade cells observed mutated (f1)
// and observed wild type (f0, times F0_SCALE<MODE>), as exact integer-valued floats.
// `taddr` = TMEM base + lane quarter + column offset inside a slot; `use0` = the tile's first
// slot-use number (slot of accumulator c is (use0 + c) & 3, at column 128 * slot).
template <int MODE>
constexpr float F0_SCALE = MODE == MODE_RADIX ? 4096.0f : 1.0f;   // radix mode leaves C0 multiplied by its radix

template <int MODE>
__device__ __forceinline__ void load_counts(uint32_t taddr, uint32_t use0, int nchunk, float (&f1)[32], float (&f0)[32]) {
    uint32_t raw[32];
    if (MODE == MODE_PLANES) {
        ptx::tmem_ld_x32(taddr + (use0 & 3) * BLOCK_N, raw);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) f1[i] = count_to_float(raw[i]);
        ptx::tmem_ld_x32(taddr + ((use0 + 1) & 3) * BLOCK_N, raw);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) f0[i] = count_to_float(raw[i]);
    } else {
        // first chunk initialises, further chunks (cells beyond 3968) accumulate.
        // v = C1 + 4096 * C0 < 2^24 is split on the FMA pipe (FRND shares the quarter-rate pipe with
        // ex2 and was half of its load): adding 1.5 * 2^35 with round-toward-zero drops the bits
        // below 4096 (ulp there is 2^12), so t - magic = 4096 * C0 and v - that = C1, all exact.
        ptx::tmem_ld_x32(taddr + (use0 & 3) * BLOCK_N, raw);
        ptx::tmem_ld_wait();
        const float2 magic = make_float2(51539607552.0f, 51539607552.0f), nmagic = make_float2(-51539607552.0f, -51539607552.0f);
        const float2 neg1 = make_float2(-1.0f, -1.0f);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            const float2 v = make_float2(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1]));
            const float2 hi = __fadd2_rn(__fadd2_rz(v, magic), nmagic);     // 4096 * C0
            const float2 lo = __ffma2_rn(neg1, hi, v);                      // C1
            f0[i] = hi.x;
            f0[i + 1] = hi.y;
            f1[i] = lo.x;
            f1[i + 1] = lo.y;
        }
#pragma unroll 1
        for (int c = 1; c < nchunk; ++c) {
            ptx::tmem_ld_x32(taddr + ((use0 + c) & 3) * BLOCK_N, raw);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const float2 v = make_float2(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1]));
                const float2 hi = __fadd2_rn(__fadd2_rz(v, magic), nmagic);
                const float2 lo = __ffma2_rn(neg1, hi, v);
                f0[i] += hi.x;
                f0[i + 1] += hi.y;
                f1[i] += lo.x;
                f1[i + 1] += lo.y;
            }
        }
    }
}

// 2^x for x <= ~20 or very negative (flushes to 0): one MUFU, no range fix-up.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Pass 1, 32 columns of one row: running reference column (refv = its rank value, b1/b0 = its counts)
// and sum of 2^(logit - logit_ref) relative to it.  The best column of the group replaces the
// reference when it beats it by more than 2^16 (then the running sum is rescaled by the exactly
// formed difference).  f0 carries the wild-type counts times a power of two that a0h/a0l/a0f undo.
__device__ __forceinline__ void stats_cols32(const GroupConst& gc, float a0h, float a0l, float a0f, const float (&f1)[32],
                                             const float (&f0)[32], int nv, float& refv, float& b1, float& b0, float& sum) {
    // nv >= 32 except in the last column tile
    float cm = -INFINITY, c1 = 0.f, c0 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float v = fmaf(gc.a1f, f1[i], a0f * f0[i]);
        const bool take = (i < nv) && (v > cm);
        cm = take ? v : cm;
        c1 = take ? f1[i] : c1;
        c0 = take ? f0[i] : c0;
    }
    if (cm > refv + 16.0f) {
        if (refv != -INFINITY) {
            const float e1 = b1 - c1, e0 = b0 - c0;
            sum *= ex2_approx(fmaf(gc.a1h, e1, a0h * e0) + fmaf(gc.a1l, e1, a0l * e0));
        }
        b1 = c1;
        b0 = c0;
        refv = cm;
    }
    if (nv >= 32) {
        const float2 nb1 = make_float2(-b1, -b1), nb0 = make_float2(-b0, -b0);
        const float2 a1h2 = make_float2(gc.a1h, gc.a1h), a0h2 = make_float2(a0h, a0h);
        const float2 a1l2 = make_float2(gc.a1l, gc.a1l), a0l2 = make_float2(a0l, a0l);
        float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            const float2 e1 = __fadd2_rn(make_float2(f1[i], f1[i + 1]), nb1);
            const float2 e0 = __fadd2_rn(make_float2(f0[i], f0[i + 1]), nb0);
            const float2 sdiff = __ffma2_rn(a1h2, e1, __fmul2_rn(a0h2, e0));   // exact operands, one rounding
            const float2 lo = __ffma2_rn(a1l2, e1, __fmul2_rn(a0l2, e0));
            const float2 x = __fadd2_rn(sdiff, lo);
            sum2 = __fadd2_rn(sum2, make_float2(ex2_approx(x.x), ex2_approx(x.y)));
        }
        sum += sum2.x + sum2.y;
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float e1 = f1[i] - b1;
            const float e0 = f0[i] - b0;
            const float sdiff = fmaf(gc.a1h, e1, a0h * e0);
            const float lo = fmaf(gc.a1l, e1, a0l * e0);
            sum += (i < nv) ? ex2_approx(sdiff + lo) : 0.f;
        }
    }
}

// ---- epilogue building blocks shared by the 1-CTA and the 2-CTA kernels -------------
// One tile (this thread's row, its 64-column half) from TMEM.
template <int PASS, int MODE>
__device__ __forceinline__ void epilogue_tile(const PlaceParams& p, const GroupConst& gc, uint32_t taddr, uint32_t use0, int row,
                                              int jl, int hh, int col0, int n_valid, float (&colacc)[64], uint2 rr) {
    float f1[32], f0[32];
    // wild-type counts arrive scaled by a power of two: fold it into their constants (exact)
    constexpr float F0S = F0_SCALE<MODE>, F0I = 1.0f / F0_SCALE<MODE>;
    const float a0h = gc.a0h * F0I, a0l = gc.a0l * F0I, a0f = gc.a0f * F0I;
    if (PASS == PASS_DUMP) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            load_counts<MODE>(taddr + ch * 32, use0, p.nchunk, f1, f0);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const size_t o = (size_t(row) * p.N_pad + col0 + ch * 32 + i) * 2;
                p.dump[o] = int(f1[i]);
                p.dump[o + 1] = int(f0[i] * F0I);
            }
        }
    } else if (PASS == PASS_STATS) {
        // One sweep over the 64 columns, 32 at a time, with a running reference column
        // (b1, b0): the best column of a group replaces it when it beats it by more than
        // 2^16 (then the running sum is rescaled by the exactly formed difference).
        float refv = -INFINITY, b1 = 0.f, b0 = 0.f, sum = 0.f;
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
            if (ch * 32 >= n_valid) break;
            load_counts<MODE>(taddr + ch * 32, use0, p.nchunk, f1, f0);
            stats_cols32(gc, a0h, a0l, a0f, f1, f0, n_valid - ch * 32, refv, b1, b0, sum);
        }
        const uint32_t pk = uint32_t(b1) | (uint32_t(b0 * F0I) << 16);
        p.stats[size_t(jl * 2 + hh) * p.M_pad + row] = make_uint2(pk, __float_as_uint(sum));
    } else {
        // rr = rowref[row], loaded by the caller one tile ahead
        const float r1f = float(rr.x & 0xFFFFu);
        const float r0f = float(rr.x >> 16) * F0S;
        const float L = __uint_as_float(rr.y);     // +inf for padded / dropped rows: 2^(-inf) = 0
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            load_counts<MODE>(taddr + ch * 32, use0, p.nchunk, f1, f0);
            const float2 nr1 = make_float2(-r1f, -r1f), nr0 = make_float2(-r0f, -r0f), nL = make_float2(-L, -L);
            const float2 a1h2 = make_float2(gc.a1h, gc.a1h), a0h2 = make_float2(a0h, a0h);
            const float2 a1l2 = make_float2(gc.a1l, gc.a1l), a0l2 = make_float2(a0l, a0l);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const float2 e1 = __fadd2_rn(make_float2(f1[i], f1[i + 1]), nr1);
                const float2 e0 = __fadd2_rn(make_float2(f0[i], f0[i + 1]), nr0);
                const float2 sdiff = __ffma2_rn(a1h2, e1, __fmul2_rn(a0h2, e0));
                const float2 tl = __ffma2_rn(a0l2, e0, nL);
                const float2 lo = __ffma2_rn(a1l2, e1, tl);
                const float2 x = __fadd2_rn(sdiff, lo);
                const float2 acc = __fadd2_rn(make_float2(colacc[ch * 32 + i], colacc[ch * 32 + i + 1]),
                                              make_float2(ex2_approx(x.x), ex2_approx(x.y)));
                colacc[ch * 32 + i] = acc.x;
                colacc[ch * 32 + i + 1] = acc.y;
            }
        }
    }
}

// ---- pass 1 with the count cache (scs_place_kernel<PASS_STATS, MODE, 2>) -----------------------
// Same sweep as epilogue_tile<PASS_STATS>; in addition every 32-column group of counts is packed as
// C1 | C0 << 16 (two mantissa plants and one byte permute per entry) and written to the cache through
// shared memory by the bulk-copy engine: direct 16-byte stores from a warp whose lanes own different
// rows touch 32 lines per instruction and half-fill every sector (measured: +17 % cycles on pass 1,
// against +9 % this way).  ncu shows the MMA warp never waiting for a TMEM slot or for operands in
// either variant: what the extra work costs is tensor-pipe duty under the power cap, so the epilogue
// adds as few instructions as it can.
// Cache layout: [row block][column tile][lane quarter q][column half hh][ch][32 rows][32 words], i.e.
// one contiguous 4 KB sub-block per (epilogue warp, 32-column group); inside a row the eight 16-byte
// chunks are XOR-swizzled with (row & 7) so that the shared-memory staging stores are conflict free.
constexpr int CACHE_SUB_WORDS = 32 * 32;

template <int MODE>
__device__ __forceinline__ void epilogue_tile_cache(const PlaceParams& p, const GroupConst& gc, uint32_t taddr, uint32_t use0, int row,
                                                    int jl, int j, int q, int hh, int lane, int n_valid, uint8_t* stage) {
    constexpr float F0I = 1.0f / F0_SCALE<MODE>;
    const float a0h = gc.a0h * F0I, a0l = gc.a0l * F0I, a0f = gc.a0f * F0I;
    uint32_t* tile = p.cache + (size_t(row >> 7) * (p.N_pad >> 7) + j) * (BLOCK_M * BLOCK_N);
    const uint32_t st_row = ptx::smem_u32(stage) + lane * 128;
    const float2 two23 = make_float2(8388608.0f, 8388608.0f), inv = make_float2(F0I, F0I);
    float refv = -INFINITY, b1 = 0.f, b0 = 0.f, sum = 0.f;
    float f1[32], f0[32];
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
        // (column groups beyond the last clade row hold zero counts -- the operand is zero padded -- and are
        // cached like the rest; the sweep over the cache drops their sums)
        load_counts<MODE>(taddr + ch * 32, use0, p.nchunk, f1, f0);
        if (lane == 0) ptx::bulk_wait_read0();     // the copy engine is done reading this warp's staging buffer
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; k += 2) {
                const float2 m1 = __fadd2_rn(make_float2(f1[4 * c + k], f1[4 * c + k + 1]), two23);          // C1 + 2^23
                const float2 m0 = __ffma2_rn(make_float2(f0[4 * c + k], f0[4 * c + k + 1]), inv, two23);     // C0 + 2^23
                w[k] = __byte_perm(__float_as_uint(m1.x), __float_as_uint(m0.x), 0x5410);
                w[k + 1] = __byte_perm(__float_as_uint(m1.y), __float_as_uint(m0.y), 0x5410);
            }
            ptx::st_shared_v4(st_row + ((c ^ (lane & 7)) << 4), w[0], w[1], w[2], w[3]);
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            // streaming: the counts are read back once, after the whole pass -- keep the operand slabs in L2
            ptx::bulk_store_hint(tile + ((q * 2 + hh) * 2 + ch) * CACHE_SUB_WORDS, stage, CACHE_SUB_WORDS * 4, ptx::L2_EVICT_FIRST);
            ptx::bulk_commit();
        }
        if (ch * 32 < n_valid) stats_cols32(gc, a0h, a0l, a0f, f1, f0, n_valid - ch * 32, refv, b1, b0, sum);
    }
    p.stats[size_t(jl * 2 + hh) * p.M_pad + row] = make_uint2(uint32_t(b1) | (uint32_t(b0 * F0I) << 16), __float_as_uint(sum));
}

// ---- radix mode, one K chunk (<= 3968 cells): the same epilogue, software pipelined ------
// With short K the tensor work is negligible and the epilogue warps (two per scheduler) are
// latency bound, so every TMEM load is issued before the arithmetic on the previous one.
constexpr int STATS_W = 32;   // columns per TMEM load in the pipelined pass-1 epilogue
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[16]) { ptx::tmem_ld_x16(taddr, r); }
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[32]) { ptx::tmem_ld_x32(taddr, r); }
__device__ __forceinline__ void tmem_wait_cols(uint32_t (&r)[16]) { ptx::tmem_ld_wait16(r); }
__device__ __forceinline__ void tmem_wait_cols(uint32_t (&r)[32]) { ptx::tmem_ld_wait32(r); }

template <int W>
__device__ __forceinline__ void decode_cols(const uint32_t (&raw)[W], float (&f1)[W], float (&f0)[W]) {
    const float2 magic = make_float2(51539607552.0f, 51539607552.0f), nmagic = make_float2(-51539607552.0f, -51539607552.0f);
    const float2 neg1 = make_float2(-1.0f, -1.0f);
#pragma unroll
    for (int i = 0; i < W; i += 2) {
        const float2 v = make_float2(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1]));
        const float2 hi = __fadd2_rn(__fadd2_rz(v, magic), nmagic);     // 4096 * C0 (see load_counts)
        const float2 lo = __ffma2_rn(neg1, hi, v);                      // C1
        f0[i] = hi.x;
        f0[i + 1] = hi.y;
        f1[i] = lo.x;
        f1[i + 1] = lo.y;
    }
}

struct StatsState {
    float refv, b1, b0, sum;   // running reference column (rank value, counts) and sum relative to it
};

// pass 1, W (16 or 32) columns: same running-reference rule as epilogue_tile<PASS_STATS>
template <int W>
__device__ __forceinline__ void stats_cols(const GroupConst& gc, float a0h, float a0l, float a0f, const uint32_t (&raw)[W], int nv,
                                           StatsState& st) {
    float f1[W], f0[W];
    decode_cols<W>(raw, f1, f0);
    // Best column by rank value a1*C1 + a0*C0 (plain fp32: it only ranks).  The column index rides
    // in the low mantissa bits of the key, so one max tree yields both the value and the position,
    // and the winner's packed counts are picked from `raw` by a select tree -- a third fewer
    // instructions than carrying (value, C1, C0) through W compare/selects.
    constexpr uint32_t IDX_MASK = W - 1;
    float key[W];
    {
        const float2 a1f2 = make_float2(gc.a1f, gc.a1f), a0f2 = make_float2(a0f, a0f);
#pragma unroll
        for (int i = 0; i < W; i += 2) {
            const float2 v = __ffma2_rn(a1f2, make_float2(f1[i], f1[i + 1]), __fmul2_rn(a0f2, make_float2(f0[i], f0[i + 1])));
            key[i] = __uint_as_float((__float_as_uint(v.x) & ~IDX_MASK) | uint32_t(i));
            key[i + 1] = __uint_as_float((__float_as_uint(v.y) & ~IDX_MASK) | uint32_t(i + 1));
        }
        if (nv < W) {
#pragma unroll
            for (int i = 0; i < W; ++i)
                if (i >= nv) key[i] = -3.0e38f;
        }
    }
#pragma unroll
    for (int n = W / 2; n >= 1; n >>= 1) {
#pragma unroll
        for (int i = 0; i < n; ++i) key[i] = fmaxf(key[2 * i], key[2 * i + 1]);
    }
    const float cm = key[0];
    if (cm > st.refv + 16.0f) {
        const uint32_t idx = __float_as_uint(cm) & IDX_MASK;
        uint32_t sel[W];
#pragma unroll
        for (int i = 0; i < W; ++i) sel[i] = raw[i];
#pragma unroll
        for (int n = W / 2, bit = 1; n >= 1; n >>= 1, bit <<= 1) {
            const bool pb = (idx & uint32_t(bit)) != 0;
#pragma unroll
            for (int i = 0; i < n; ++i) sel[i] = pb ? sel[2 * i + 1] : sel[2 * i];
        }
        const float vsel = __uint_as_float(sel[0]);
        const float c0 = __fadd_rn(__fadd_rz(vsel, 51539607552.0f), -51539607552.0f);   // 4096 * C0
        const float c1 = vsel - c0;
        if (st.refv != -INFINITY) {
            const float e1 = st.b1 - c1, e0 = st.b0 - c0;
            st.sum *= ex2_approx(fmaf(gc.a1h, e1, a0h * e0) + fmaf(gc.a1l, e1, a0l * e0));
        }
        st.b1 = c1;
        st.b0 = c0;
        st.refv = cm;
    }
    if (nv >= W) {
        const float2 nb1 = make_float2(-st.b1, -st.b1), nb0 = make_float2(-st.b0, -st.b0);
        const float2 a1h2 = make_float2(gc.a1h, gc.a1h), a0h2 = make_float2(a0h, a0h);
        const float2 a1l2 = make_float2(gc.a1l, gc.a1l), a0l2 = make_float2(a0l, a0l);
        float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < W; i += 2) {
            const float2 e1 = __fadd2_rn(make_float2(f1[i], f1[i + 1]), nb1);
            const float2 e0 = __fadd2_rn(make_float2(f0[i], f0[i + 1]), nb0);
            const float2 sdiff = __ffma2_rn(a1h2, e1, __fmul2_rn(a0h2, e0));   // exact operands, one rounding
            const float2 lo = __ffma2_rn(a1l2, e1, __fmul2_rn(a0l2, e0));
            const float2 x = __fadd2_rn(sdiff, lo);
            sum2 = __fadd2_rn(sum2, make_float2(ex2_approx(x.x), ex2_approx(x.y)));
        }
        st.sum += sum2.x + sum2.y;
    } else {
#pragma unroll
        for (int i = 0; i < W; ++i) {
            const float e1 = f1[i] - st.b1;
            const float e0 = f0[i] - st.b0;
            const float sdiff = fmaf(gc.a1h, e1, a0h * e0);
            const float lo = fmaf(gc.a1l, e1, a0l * e0);
            st.sum += (i < nv) ? ex2_approx(sdiff + lo) : 0.f;
        }
    }
}

// pass 2, 16 columns starting at accumulator OFF
template <int OFF>
__device__ __forceinline__ void colsum16(const GroupConst& gc, float a0h, float a0l, const uint32_t (&raw)[16], float r1f, float r0f,
                                         float L, float (&colacc)[64]) {
    float f1[16], f0[16];
    decode_cols<16>(raw, f1, f0);
    const float2 nr1 = make_float2(-r1f, -r1f), nr0 = make_float2(-r0f, -r0f), nL = make_float2(-L, -L);
    const float2 a1h2 = make_float2(gc.a1h, gc.a1h), a0h2 = make_float2(a0h, a0h);
    const float2 a1l2 = make_float2(gc.a1l, gc.a1l), a0l2 = make_float2(a0l, a0l);
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
        const float2 e1 = __fadd2_rn(make_float2(f1[i], f1[i + 1]), nr1);
        const float2 e0 = __fadd2_rn(make_float2(f0[i], f0[i + 1]), nr0);
        const float2 sdiff = __ffma2_rn(a1h2, e1, __fmul2_rn(a0h2, e0));
        const float2 tl = __ffma2_rn(a0l2, e0, nL);
        const float2 lo = __ffma2_rn(a1l2, e1, tl);
        const float2 x = __fadd2_rn(sdiff, lo);
        const float2 acc = __fadd2_rn(make_float2(colacc[OFF + i], colacc[OFF + i + 1]), make_float2(ex2_approx(x.x), ex2_approx(x.y)));
        colacc[OFF + i] = acc.x;
        colacc[OFF + i + 1] = acc.y;
    }
}

// `taddr` = TMEM address of this thread's row, this warp's 64-column half, in the tile's slot
template <int PASS>
__device__ __forceinline__ void epilogue_tile_r1(const PlaceParams& p, const GroupConst& gc, uint32_t taddr, int row, int jl, int hh,
                                                 int n_valid, float (&colacc)[64], uint2 rr) {
    constexpr float F0S = 4096.0f, F0I = 1.0f / 4096.0f;
    const float a0h = gc.a0h * F0I, a0l = gc.a0l * F0I, a0f = gc.a0f * F0I;
    if (n_valid <= 0) {                    // (uniform over the warp) nothing to do in this half
        if (PASS == PASS_STATS) p.stats[size_t(jl * 2 + hh) * p.M_pad + row] = make_uint2(0u, 0u);
        return;
    }
    if (PASS == PASS_STATS) {
        constexpr int W = STATS_W;
        uint32_t wa[W], wb[W];
        StatsState st = {-INFINITY, 0.f, 0.f, 0.f};
        tmem_ld_cols(taddr, wa);
#pragma unroll
        for (int c = 0; c < 64 / W; c += 2) {
            tmem_wait_cols(wa);
            if (n_valid > (c + 1) * W) tmem_ld_cols(taddr + (c + 1) * W, wb);
            stats_cols<W>(gc, a0h, a0l, a0f, wa, n_valid - c * W, st);
            if (n_valid > (c + 1) * W) {
                tmem_wait_cols(wb);
                if ((c + 2) * W < 64 && n_valid > (c + 2) * W) tmem_ld_cols(taddr + (c + 2) * W, wa);
                stats_cols<W>(gc, a0h, a0l, a0f, wb, n_valid - (c + 1) * W, st);
            }
            if (n_valid <= (c + 2) * W) break;
        }
        const uint32_t pk = uint32_t(st.b1) | (uint32_t(st.b0 * F0I) << 16);
        p.stats[size_t(jl * 2 + hh) * p.M_pad + row] = make_uint2(pk, __float_as_uint(st.sum));
    } else {
        const float r1f = float(rr.x & 0xFFFFu);
        const float r0f = float(rr.x >> 16) * F0S;
        const float L = __uint_as_float(rr.y);     // +inf for padded / dropped rows: 2^(-inf) = 0
        uint32_t ra[16], rb[16];
        ptx::tmem_ld_x16(taddr, ra);
        ptx::tmem_ld_wait16(ra);
        if (n_valid > 16) ptx::tmem_ld_x16(taddr + 16, rb);
        colsum16<0>(gc, a0h, a0l, ra, r1f, r0f, L, colacc);
        if (n_valid > 16) {
            ptx::tmem_ld_wait16(rb);
            if (n_valid > 32) ptx::tmem_ld_x16(taddr + 32, ra);
            colsum16<16>(gc, a0h, a0l, rb, r1f, r0f, L, colacc);
        }
        if (n_valid > 32) {
            ptx::tmem_ld_wait16(ra);
            if (n_valid > 48) ptx::tmem_ld_x16(taddr + 48, rb);
            colsum16<32>(gc, a0h, a0l, ra, r1f, r0f, L, colacc);
        }
        if (n_valid > 48) {
            ptx::tmem_ld_wait16(rb);
            colsum16<48>(gc, a0h, a0l, rb, r1f, r0f, L, colacc);
        }
    }
}

// End of a run (pass 2): rows -> one number per column: lanes, then the four lane quarters.
__device__ __forceinline__ void epilogue_flush_columns(const PlaceParams& p, float* red, const float (&colacc)[64], int q, int hh,
                                                       int lane, int run, int j) {
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        float v = colacc[i];
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        if (lane == (i & 31)) red[q * BLOCK_N + hh * 64 + i] = v;
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(NUM_EPI_WARPS * 32) : "memory");
    const int et = threadIdx.x - 64;
    if (et < BLOCK_N) {
        const float v = ((red[et] + red[BLOCK_N + et]) + red[2 * BLOCK_N + et]) + red[3 * BLOCK_N + et];
        p.colpart[size_t(run) * p.N_pad + j * BLOCK_N + et] = v;
    }
    asm volatile("bar.sync 1, %0;\n" ::"n"(NUM_EPI_WARPS * 32) : "memory");
}

// 10 warps are allocated as 12 (warp slots come in fours), which caps the kernel at
// 65536 / 384 = 168 registers per thread; the 64-accumulator pass-2 epilogue spills a few.
template <int PASS, int MODE, int R1 = 0>   // R1 = 1: radix mode with one K chunk -> pipelined epilogue; R1 = 2: pass 1 that also fills the count cache
__global__ void __launch_bounds__(NUM_THREADS, 1)
scs_place_kernel(const __grid_constant__ CUtensorMap tm_x1, const __grid_constant__ CUtensorMap tm_x0,
                 const __grid_constant__ CUtensorMap tm_s, const PlaceParams p) {
    using Cfg = ModeCfg<MODE>;
    constexpr int NUM_STAGES = Cfg::NUM_STAGES;
    constexpr int STAGE_BYTES = Cfg::STAGE_BYTES;
    constexpr int A_TILES = Cfg::A_TILES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NUM_STAGES * STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + NUM_STAGES;
    // TMEM is a rotating pool of four 128-column accumulator slots.  A tile takes SPT
    // consecutive slots (2 planes, or one per K chunk); the MMA warp may start a slot as
    // soon as the epilogue of the tile that used it last has drained it, so the epilogue
    // overlaps the next tile's MMAs whenever SPT < 4.
    uint64_t* tfull_bar = bars + 2 * NUM_STAGES;        // [4] tile n -> n & 3
    uint64_t* sempty_bar = bars + 2 * NUM_STAGES + 4;   // [4] one per slot
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * NUM_STAGES + 8);
    float* red = reinterpret_cast<float*>(smem + NUM_STAGES * STAGE_BYTES + SMEM_BARRIER_BYTES);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp_idx == 0 && ptx::elect_one()) {
        ptx::prefetch_tmap(&tm_x1);
        if (A_TILES == 2) ptx::prefetch_tmap(&tm_x0);
        ptx::prefetch_tmap(&tm_s);
        for (int s = 0; s < NUM_STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 4; ++a) {
            ptx::mbar_init(&tfull_bar[a], 1);
            ptx::mbar_init(&sempty_bar[a], NUM_EPI_WARPS);
        }
        ptx::fence_mbar_init();
    }
    if (warp_idx == 1) {
        ptx::tmem_alloc(tmem_ptr_smem, TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp_idx == 0) {
        // ======================= TMA producer (one lane) =======================
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const uint64_t pol_x = p.l2_hints ? ptx::L2_EVICT_FIRST : ptx::L2_EVICT_NORMAL;
            const uint64_t pol_s = p.l2_hints ? ptx::L2_EVICT_LAST : ptx::L2_EVICT_NORMAL;
            for (int t = blockIdx.x; t < p.num_tasks; t += gridDim.x) {
                const int j = p.tasks[t].j, r0 = p.tasks[t].r0, r1 = p.tasks[t].r1;
                for (int r = r0; r < r1; ++r) {
                    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* st = smem + stage * STAGE_BYTES;
                        ptx::mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                        // the clade slab is re-read by this CTA for every row block of the run and shared
                        // by the whole super-block: keep it; genotype rows stream through once per super-block
                        ptx::tma_load_2d_hint(&tm_x1, &full_bar[stage], st, kb * BLOCK_K, r * BLOCK_M, pol_x);
                        if (A_TILES == 2)
                            ptx::tma_load_2d_hint(&tm_x0, &full_bar[stage], st + TILE_BYTES, kb * BLOCK_K, r * BLOCK_M, pol_x);
                        ptx::tma_load_2d_hint(&tm_s, &full_bar[stage], st + A_TILES * TILE_BYTES, kb * BLOCK_K, j * BLOCK_N, pol_s);
                        if (++stage == NUM_STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ======================= MMA issuer (one lane) =========================
        if (ptx::elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const int spt = MODE == MODE_PLANES ? 2 : p.nchunk;   // slots per tile
            uint32_t tile_n = 0;                                   // tiles issued by this CTA
            for (int t = blockIdx.x; t < p.num_tasks; t += gridDim.x) {
                const int r0 = p.tasks[t].r0, r1 = p.tasks[t].r1;
                for (int r = r0; r < r1; ++r, ++tile_n) {
                    const uint32_t use0 = tile_n * spt;            // running slot-use counter
                    if (MODE == MODE_PLANES) {
                        // both planes accumulate over the whole K loop: claim both slots up front
                        for (int c = 0; c < 2; ++c) ptx::mbar_wait(&sempty_bar[(use0 + c) & 3], (((use0 + c) >> 2) & 1) ^ 1);
                        ptx::tc_fence_after();
                    }
                    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                        int chunk = 0, kin = kb;
                        if (MODE == MODE_RADIX) {
                            chunk = kb / RADIX_KB_PER_CHUNK;
                            kin = kb - chunk * RADIX_KB_PER_CHUNK;
                            if (kin == 0) {                        // first K block of a chunk: claim its slot
                                ptx::mbar_wait(&sempty_bar[(use0 + chunk) & 3], (((use0 + chunk) >> 2) & 1) ^ 1);
                                ptx::tc_fence_after();
                            }
                        }
                        ptx::mbar_wait(&full_bar[stage], phase);
                        ptx::tc_fence_after();
                        const uint32_t st = ptx::smem_u32(smem + stage * STAGE_BYTES);
                        const uint64_t dx1 = ptx::make_kmajor_sw128_desc(st);
                        const uint64_t ds = ptx::make_kmajor_sw128_desc(st + A_TILES * TILE_BYTES);
                        if (MODE == MODE_PLANES) {
                            const uint64_t dx0 = ptx::make_kmajor_sw128_desc(st + TILE_BYTES);
                            const uint32_t d1 = tmem_base + ((use0 & 3) * BLOCK_N);
                            const uint32_t d0 = tmem_base + (((use0 + 1) & 3) * BLOCK_N);
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                                // advance 32 bytes along K inside the 128-byte swizzle row: +2 in 16-byte units
                                const uint64_t koff = uint64_t(k * (UMMA_K >> 4));
                                const uint32_t accum = (kb | k) != 0;
                                ptx::mma_i8_ss(d1, dx1 + koff, ds + koff, IDESC_I8, accum);
                                ptx::mma_i8_ss(d0, dx0 + koff, ds + koff, IDESC_I8, accum);
                            }
                        } else {
                            const uint32_t d = tmem_base + (((use0 + chunk) & 3) * BLOCK_N);
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                                const uint64_t koff = uint64_t(k * (UMMA_K >> 4));
                                ptx::mma_f8_ss(d, dx1 + koff, ds + koff, IDESC_E5M2, (kin | k) != 0);
                            }
                        }
                        ptx::mma_commit(&empty_bar[stage]);
                        if (++stage == NUM_STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    ptx::mma_commit(&tfull_bar[tile_n & 3]);
                }
            }
        }
    } else {
        // ======================= epilogue warps ================================
        const int ew = warp_idx - 2;
        const int q = warp_idx & 3;   // TMEM lane quarter this warp may touch
        const int hh = ew >> 2;       // column half of the tile
        const uint32_t lane_base = uint32_t(q * 32) << 16;
        const int spt = MODE == MODE_PLANES ? 2 : p.nchunk;
        uint32_t tile_n = 0;
        // task entry and its group's constants are fetched one task ahead (two dependent loads)
        TaskEntry te_next = {};
        GroupConst gc_next = {};
        if (R1 == 1 && int(blockIdx.x) < p.num_tasks) {      // (short tasks only: long-K kernels are short of registers)
            te_next = p.tasks[blockIdx.x];
            gc_next = p.gconst[te_next.group];
        }
        for (int t = blockIdx.x; t < p.num_tasks; t += gridDim.x) {
            const TaskEntry te = R1 == 1 ? te_next : p.tasks[t];
            const GroupConst gc = R1 == 1 ? gc_next : p.gconst[te.group];
            if (R1 == 1 && t + int(gridDim.x) < p.num_tasks) {
                te_next = p.tasks[t + gridDim.x];
                gc_next = p.gconst[te_next.group];
            }
            const int j = te.j, r0 = te.r0, r1 = te.r1;
            const int col0 = j * BLOCK_N + hh * 64;
            const int n_valid = max(0, min(64, te.nvalid - hh * 64));
            float colacc[64];
            if (PASS == PASS_COLSUM) {
#pragma unroll
                for (int i = 0; i < 64; ++i) colacc[i] = 0.f;
            }
            // pass 2 needs rowref[row]: fetched one tile ahead so that the DRAM latency hides behind a tile
            uint2 rr = make_uint2(0u, 0u), rr_next = make_uint2(0u, 0u);
            if (PASS == PASS_COLSUM) rr = p.rowref[r0 * BLOCK_M + q * 32 + lane];
            for (int r = r0; r < r1; ++r, ++tile_n) {
                const int row = r * BLOCK_M + q * 32 + lane;
                if (PASS == PASS_COLSUM && r + 1 < r1) rr_next = p.rowref[row + BLOCK_M];
                ptx::mbar_wait(&tfull_bar[tile_n & 3], (tile_n >> 2) & 1);
                ptx::tc_fence_after();
                const uint32_t use0 = tile_n * spt;
                const uint32_t taddr = tmem_base + lane_base + hh * 64;
                if (R1 == 2)
                    epilogue_tile_cache<MODE>(p, gc, taddr, use0, row, te.jl, j, q, hh, lane, n_valid, reinterpret_cast<uint8_t*>(red) + ew * (CACHE_SUB_WORDS * 4));
                else if (R1 == 1)
                    epilogue_tile_r1<PASS>(p, gc, taddr + (use0 & 3) * BLOCK_N, row, te.jl, hh, n_valid, colacc, rr);
                else
                    epilogue_tile<PASS, MODE>(p, gc, taddr, use0, row, te.jl, hh, col0, n_valid, colacc, rr);
                rr = rr_next;
                // this tile's slots are drained: hand them back to the MMA warp
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0)
                    for (int c = 0; c < spt; ++c) ptx::mbar_arrive(&sempty_bar[(use0 + c) & 3]);
            }
            if (PASS == PASS_COLSUM) epilogue_flush_columns(p, red, colacc, q, hh, lane, te.run_local, j);
        }
        if (R1 == 2 && lane == 0) ptx::bulk_wait0();   // the staging buffers must outlive the copies
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp_idx == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------- operand expansion
// Row geometry of a (possibly multi-group) plan: every group's SNVs are padded to whole
// 256-row units in the operand planes; group_of_unit[u] says whose rows unit u holds.
struct RowMap {
    const int* group_of_unit;       // [M_pad / 256]
    const long long* group_row0;    // [G] first padded (operand) row of the group
    const long long* group_src0;    // [G] first row of the group in the caller's packed matrix
    const int* group_nsnv;          // [G]
};

// 2-bit genotype codes (0 wild type, 1 het, 2 hom, 3 missing; 4 cells per byte,
// cell c in bits [2*(c%4), +2)) -> 0/1 byte plane(s), K-major, zero padded.  One thread
// turns 16 cells (one 32-bit word of codes) into one 16-byte store per plane.  Padding rows
// of the planes are never written (they were zeroed when the plan was created).
__device__ __forceinline__ void expand_16_cells(const uint8_t* __restrict__ gt, int64_t gt_pitch, int n_cells, const RowMap& rm,
                                                uint8_t* __restrict__ x1, uint8_t* __restrict__ x0, int K_pad, int radix, int64_t row,
                                                int kq) {
    const int g = rm.group_of_unit[row >> 8];
    const int64_t local = row - rm.group_row0[g];
    if (local >= rm.group_nsnv[g]) return;
    const int64_t src = rm.group_src0[g] + local;
    uint32_t w = 0;
    if (kq * 4 < gt_pitch && kq * 16 < n_cells) w = *reinterpret_cast<const uint32_t*>(gt + src * gt_pitch + kq * 4);
    const int left = n_cells - kq * 16;                 // cells of this group that exist (may be <= 0 or >= 16)
    if (left < 16) w &= left <= 0 ? 0u : ((1u << (2 * left)) - 1u);      // trailing codes read as 0 ...
    uint32_t o1[4], o0[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        // spread the four 2-bit codes of byte q into the low bits of four bytes
        uint32_t x = (w >> (8 * q)) & 0xFFu;
        x = (x | (x << 12)) & 0x000F000Fu;
        x = (x | (x << 6)) & 0x03030303u;
        const uint32_t lo = x & 0x01010101u, hi = (x >> 1) & 0x01010101u;
        uint32_t mut = lo ^ hi;                          // code 1 or 2
        uint32_t wt = (lo | hi) ^ 0x01010101u;           // code 0
        if (left < 16) {                                 // ... so mask them out of the wild-type plane
            const int rem = left - 4 * q;
            const uint32_t valid = rem >= 4 ? 0x01010101u : (rem <= 0 ? 0u : (0x01010101u >> (8 * (4 - rem))));
            wt &= valid;
        }
        if (radix) {
            o1[q] = mut * uint32_t(E5M2_ONE) + wt * uint32_t(E5M2_4096);   // byte fields: 1.0 / 4096.0 / 0, no carries
            o0[q] = 0;
        } else {
            o1[q] = mut;
            o0[q] = wt;
        }
    }
    *reinterpret_cast<uint4*>(x1 + row * K_pad + kq * 16) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
    if (!radix) *reinterpret_cast<uint4*>(x0 + row * K_pad + kq * 16) = make_uint4(o0[0], o0[1], o0[2], o0[3]);
}

// Wide matrices: blockIdx.y walks the rows, threads cover the 16-cell groups of a row.
__global__ void scs_expand_genotypes_kernel(const uint8_t* __restrict__ gt, int64_t gt_pitch, int64_t M_pad, int n_cells, RowMap rm,
                                            uint8_t* __restrict__ x1, uint8_t* __restrict__ x0, int K_pad, int radix) {
    const int kq = blockIdx.x * blockDim.x + threadIdx.x;
    if (kq * 16 >= K_pad) return;
    for (int64_t row = blockIdx.y; row < M_pad; row += gridDim.y) expand_16_cells(gt, gt_pitch, n_cells, rm, x1, x0, K_pad, radix, row, kq);
}

// Narrow matrices (replicate batches, ~100 cells): flattened (row, group) index so that every thread works.
__global__ void scs_expand_genotypes_flat_kernel(const uint8_t* __restrict__ gt, int64_t gt_pitch, int64_t M_pad, int n_cells, RowMap rm,
                                                 uint8_t* __restrict__ x1, uint8_t* __restrict__ x0, int K_pad, int radix) {
    const uint32_t kq_per_row = uint32_t(K_pad / 16);
    const uint32_t total = uint32_t(M_pad) * kq_per_row;            // the host guarantees this fits 32 bits
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const uint32_t row = idx / kq_per_row;
        expand_16_cells(gt, gt_pitch, n_cells, rm, x1, x0, K_pad, radix, int64_t(row), int(idx - row * kq_per_row));
    }
}

// clade bit masks [G * n_rows][words] (bit c%32 of word c/32 set when cell c is below the
// node) -> 0/1 byte matrix, K-major; group g's rows start at operand row g * N_pad1.
__global__ void scs_expand_clades_kernel(const uint32_t* __restrict__ bits, int64_t pitch_words, int64_t total_rows, int n_rows,
                                         int N_pad1, int n_cells, uint8_t* __restrict__ s, int K_pad, uint32_t one) {
    const int kq = blockIdx.x * blockDim.x + threadIdx.x;
    if (kq * 4 >= K_pad) return;
    for (int64_t src = blockIdx.y; src < total_rows; src += gridDim.y) {
        const int64_t g = src / n_rows;
        const int64_t row = g * N_pad1 + (src - g * n_rows);
        uint32_t o = 0;
        if (kq * 4 < n_cells) {
            const uint32_t w = bits[src * pitch_words + (kq >> 3)] >> ((kq & 7) * 4);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool in = kq * 4 + i < n_cells;
                o |= ((in && ((w >> i) & 1u)) ? one : 0u) << (8 * i);
            }
        }
        *reinterpret_cast<uint32_t*>(s + row * K_pad + kq * 4) = o;
    }
}

// ------------------------------------------------------------ small kernels
// Merge the per-(row, half tile) partials of pass 1 into one record per SNV:
// the reference counts of the globally best column and log2 sum_b 2^(logit-ref).
__global__ void scs_merge_stats_kernel(const uint2* __restrict__ stats, int num_parts, int64_t M_pad, RowMap rm,
                                       const uint8_t* __restrict__ row_keep, const double2* __restrict__ gconst_d,
                                       uint2* __restrict__ rowref, float4* __restrict__ rowref_f) {
    const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (row >= M_pad) return;
    const int g = rm.group_of_unit[row >> 8];
    const int64_t local = row - rm.group_row0[g];
    if (local >= rm.group_nsnv[g] || (row_keep != nullptr && row_keep[rm.group_src0[g] + local] == 0)) {
        rowref[row] = make_uint2(0u, __float_as_uint(INFINITY));  // 2^(x - inf) = 0: row contributes nothing
        if (rowref_f) rowref_f[row] = make_float4(-8388608.0f, -8388608.0f, -INFINITY, 0.f);
        return;
    }
    const double a1 = gconst_d[g].x, a0 = gconst_d[g].y;
    double best = -1e300;
    uint32_t best_pk = 0;
    for (int s = 0; s < num_parts; ++s) {
        const uint2 v = stats[size_t(s) * M_pad + row];
        if (__uint_as_float(v.y) > 0.f) {
            const double l = a1 * double(v.x & 0xFFFFu) + a0 * double(v.x >> 16);
            if (l > best) {
                best = l;
                best_pk = v.x;
            }
        }
    }
    double total = 0.0;
    for (int s = 0; s < num_parts; ++s) {
        const uint2 v = stats[size_t(s) * M_pad + row];
        const float sm = __uint_as_float(v.y);
        if (sm > 0.f) {
            const double l = a1 * double(v.x & 0xFFFFu) + a0 * double(v.x >> 16);
            total += double(sm) * exp2(l - best);
        }
    }
    rowref[row] = make_uint2(best_pk, __float_as_uint(float(log2(total))));
    // the same record for the cached pass 2: reference counts offset by 2^23, negated (see scs_colsum_cached_kernel)
    if (rowref_f) rowref_f[row] = make_float4(-(8388608.0f + float(best_pk & 0xFFFFu)), -(8388608.0f + float(best_pk >> 16)), -float(log2(total)), 0.f);
}

// Pass 2 from the count cache (single-group plans with long K): the tensor work of a second GEMM
// pass costs 2*K flops per (SNV, clade row) -- 140 ms at 10k cells -- while reading the 4-byte packed
// counts back costs 4 bytes, so beyond ~2,000 cells the HBM round trip of the *counts* (not of the
// floating-point SNV x branch matrix: the softmax terms are still formed on the fly, with the same
// arithmetic on exact integer differences as the GEMM epilogue) is several times cheaper.
// Cache layout (written by scs_place_kernel<PASS_STATS, MODE, 2>): 64 KB per tile as 16 contiguous 4 KB
// sub-blocks [lane quarter q][32-column group cs][32 rows][8 x 16 bytes], the 16-byte chunks of a row
// XOR-swizzled with (row & 7).  Block = one column tile over one run of row blocks; warp w takes the
// 32-column group cs = w and walks the four sub-blocks of every tile.  One load instruction covers
// four whole rows (512 contiguous bytes): lane = (row offset lane >> 3, logical chunk lane & 7), so a
// lane keeps the same four clade columns throughout; four loads in flight per thread (64 bytes: ~64 KB
// per SM at full occupancy).  The arithmetic is the packed fp32x2 form of the GEMM epilogue's;
// rowref_f holds the NEGATED record {-(2^23 + R1), -(2^23 + R0), -log2 sum}.  Partial column sums land
// in colpart[run][col] like the GEMM pass 2.
__global__ void __launch_bounds__(BLOCK_N)
scs_colsum_cached_kernel(const uint32_t* __restrict__ cache, const float4* __restrict__ rowref_f, const GroupConst* __restrict__ gconst,
                         int num_row_blocks, int64_t N_pad, int n_cols, int run_len, float* __restrict__ colpart) {
    constexpr int UNROLL = 4;
    const int j = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int rsub = lane >> 3, chunk = lane & 7;
    const GroupConst gc = gconst[0];
    const int rb0 = blockIdx.y * run_len, rb1 = min(num_row_blocks, rb0 + run_len);
    const int64_t ct = N_pad / BLOCK_N;
    const float2 a1h2 = make_float2(gc.a1h, gc.a1h), a0h2 = make_float2(gc.a0h, gc.a0h);
    const float2 a1l2 = make_float2(gc.a1l, gc.a1l), a0l2 = make_float2(gc.a0l, gc.a0l);
    // row k * 4 + rsub of a sub-block: (row & 7) = (k & 1) * 4 + rsub; offsets in uint4 units
    const int off_even = rsub * 8 + (chunk ^ rsub), off_odd = rsub * 8 + (chunk ^ (4 + rsub));
    float2 acc[UNROLL][2];      // one accumulator pair per load slot: short fp32 chains
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc[u][0] = acc[u][1] = make_float2(0.f, 0.f);
    for (int rb = rb0; rb < rb1; ++rb) {
        const uint4* tile = reinterpret_cast<const uint4*>(cache + (size_t(rb) * ct + j) * (BLOCK_M * BLOCK_N));
#pragma unroll 1
        for (int qh = 0; qh < 8; ++qh) {          // sub-block q = qh >> 1, its rows 16 * (qh & 1) .. + 15
            const int q = qh >> 1, k0 = (qh & 1) * 4;
            const uint4* src = tile + (q * 4 + w) * (CACHE_SUB_WORDS / 4) + k0 * 32;
            const float4* ref = rowref_f + size_t(rb) * BLOCK_M + q * 32 + k0 * 4 + rsub;
            uint4 v[UNROLL];
            float4 rr[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                v[u] = __ldcs(src + u * 32 + ((u & 1) ? off_odd : off_even));     // read once
                rr[u] = __ldg(ref + 4 * u);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const float2 n1 = make_float2(rr[u].x, rr[u].x), n0 = make_float2(rr[u].y, rr[u].y), nL = make_float2(rr[u].z, rr[u].z);
                const uint32_t qv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    // count + 2^23 as a float straight from the bits; the reference already carries the 2^23
                    const float2 e1 = __fadd2_rn(make_float2(__uint_as_float(__byte_perm(qv[2 * h], 0x4B000000u, 0x7610)),
                                                             __uint_as_float(__byte_perm(qv[2 * h + 1], 0x4B000000u, 0x7610))), n1);
                    const float2 e0 = __fadd2_rn(make_float2(__uint_as_float(__byte_perm(qv[2 * h], 0x4B000000u, 0x7632)),
                                                             __uint_as_float(__byte_perm(qv[2 * h + 1], 0x4B000000u, 0x7632))), n0);
                    const float2 sdiff = __ffma2_rn(a1h2, e1, __fmul2_rn(a0h2, e0));   // exact operands, one rounding
                    const float2 lo = __ffma2_rn(a1l2, e1, __ffma2_rn(a0l2, e0, nL));
                    const float2 x = __fadd2_rn(sdiff, lo);
                    acc[u][h] = __fadd2_rn(acc[u][h], make_float2(ex2_approx(x.x), ex2_approx(x.y)));
                }
            }
        }
    }
    float out[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float2 t = __fadd2_rn(__fadd2_rn(acc[0][h], acc[1][h]), __fadd2_rn(acc[2][h], acc[3][h]));
        out[2 * h] = t.x;
        out[2 * h + 1] = t.y;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {     // the four row offsets of a load sit in lanes 8 apart
        out[i] += __shfl_xor_sync(0xffffffffu, out[i], 8);
        out[i] += __shfl_xor_sync(0xffffffffu, out[i], 16);
    }
    if (rsub == 0) {
        const int col = j * BLOCK_N + w * 32 + chunk * 4;
        // clade rows beyond the last valid one (zero rows of the padded operand) carry no weight
#pragma unroll
        for (int i = 0; i < 4; ++i) out[i] = col + i < n_cols ? out[i] : 0.f;
        *reinterpret_cast<float4*>(colpart + size_t(blockIdx.y) * N_pad + col) = make_float4(out[0], out[1], out[2], out[3]);
    }
}

// Y[g][b] = sum over the group's runs of the per-run column partials, fixed order, fp64.
__global__ void scs_reduce_colpart_kernel(const float* __restrict__ colpart, const int* __restrict__ group_runs, int64_t total_rows,
                                          int n_rows, int N_pad1, int64_t N_pad, double* __restrict__ y) {
    const int64_t o = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;   // index into the compact [G][n_rows] output
    if (o >= total_rows) return;
    const int64_t g = o / n_rows;
    const int64_t col = g * N_pad1 + (o - g * n_rows);
    const int runs = group_runs[g];
    double acc = 0.0;
    for (int r = 0; r < runs; ++r) acc += double(colpart[size_t(r) * N_pad + col]);
    y[o] = acc;
}

}  // namespace scs
#include "scs_place_interval.cuh"
#include "scs_place_iv4.cuh"
#include "scs_place_iv5.cuh"
#include "scs_place_percell.cuh"
namespace scs {

// ---------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    return fn;
}

static int make_u8_tmap(CUtensorMap* tm, void* base, uint64_t rows, uint64_t k_pad, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return set_error(SCS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {k_pad, rows};
    cuuint64_t strides[1] = {k_pad};
    cuuint32_t box[2] = {BLOCK_K, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(SCS_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", int(r));
    return SCS_OK;
}

template <int PASS>
static void launch_place(int mode, int grid, cudaStream_t st, const CUtensorMap& a1, const CUtensorMap& a0, const CUtensorMap& b,
                         const PlaceParams& p) {
    if (PASS == PASS_STATS && p.cache != nullptr) {
        // pass 1 that also fills the count cache: same pipeline, 32 KB of staging buffers in place of `red`
        constexpr int R = PASS == PASS_STATS ? 2 : 0;
        if (mode == MODE_RADIX)
            scs_place_kernel<PASS, MODE_RADIX, R><<<grid, NUM_THREADS, ModeCfg<MODE_RADIX>::SMEM_CACHE, st>>>(a1, a0, b, p);
        else
            scs_place_kernel<PASS, MODE_PLANES, R><<<grid, NUM_THREADS, ModeCfg<MODE_PLANES>::SMEM_CACHE, st>>>(a1, a0, b, p);
    } else if (mode == MODE_RADIX && p.nchunk == 1 && PASS != PASS_DUMP)
        scs_place_kernel<PASS, MODE_RADIX, (PASS != PASS_DUMP)><<<grid, NUM_THREADS, ModeCfg<MODE_RADIX>::SMEM_TOTAL, st>>>(a1, a0, b, p);
    else if (mode == MODE_RADIX)
        scs_place_kernel<PASS, MODE_RADIX><<<grid, NUM_THREADS, ModeCfg<MODE_RADIX>::SMEM_TOTAL, st>>>(a1, a0, b, p);
    else
        scs_place_kernel<PASS, MODE_PLANES><<<grid, NUM_THREADS, ModeCfg<MODE_PLANES>::SMEM_TOTAL, st>>>(a1, a0, b, p);
}

template <int PASS>
static void set_place_smem() {
    cudaFuncSetAttribute(scs_place_kernel<PASS, MODE_PLANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, ModeCfg<MODE_PLANES>::SMEM_TOTAL);
    cudaFuncSetAttribute(scs_place_kernel<PASS, MODE_RADIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, ModeCfg<MODE_RADIX>::SMEM_TOTAL);
    if (PASS == PASS_STATS) {
        cudaFuncSetAttribute(scs_place_kernel<PASS, MODE_RADIX, PASS == PASS_STATS ? 2 : 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             ModeCfg<MODE_RADIX>::SMEM_CACHE);
        cudaFuncSetAttribute(scs_place_kernel<PASS, MODE_PLANES, PASS == PASS_STATS ? 2 : 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             ModeCfg<MODE_PLANES>::SMEM_CACHE);
    }
    if (PASS != PASS_DUMP)
        cudaFuncSetAttribute(scs_place_kernel<PASS, MODE_RADIX, (PASS != PASS_DUMP)>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             ModeCfg<MODE_RADIX>::SMEM_TOTAL);
}

struct Placement {
    scs_ctx* ctx = nullptr;
    int n_groups = 1;
    int64_t n_snv = 0;            // total over groups
    int n_cells = 0, n_rows = 0;  // per group
    int64_t M_pad = 0, N_pad = 0;
    int K_pad = 0, N_pad1 = 0;
    std::vector<int64_t> g_nsnv, g_row0, g_src0;
    std::vector<int> g_runs;
    // device tables
    int* d_group_of_unit = nullptr;
    long long *d_group_row0 = nullptr, *d_group_src0 = nullptr;
    int *d_group_nsnv = nullptr, *d_group_runs = nullptr;
    TaskEntry* d_tasks = nullptr;
    GroupConst* d_gconst = nullptr;
    double2* d_gconst_d = nullptr;
    RowMap rowmap;
    int num_parts = 0;
    uint8_t *x1 = nullptr, *x0 = nullptr, *s = nullptr;
    uint8_t* row_keep = nullptr;  // optional, device
    bool have_row_keep = false;
    uint2* stats = nullptr;
    uint2* rowref = nullptr;
    float4* rowref_f = nullptr;   // cached pass 2 only
    uint32_t* cache = nullptr;    // [M_pad][N_pad] packed counts, allocated on first use
    int pass2_mode = 0;           // 0 undecided, 1 recompute (second GEMM pass), 2 count cache
    float* colpart = nullptr;
    double* y = nullptr;
    int* dump = nullptr;
    CUtensorMap tm_x1, tm_x0, tm_s;
    PlaceParams prm;
    int grid = 0;
    uint8_t* stage_gt = nullptr;   // device staging for host-side genotype uploads
    size_t stage_gt_bytes = 0;
    uint32_t* stage_mask = nullptr;
    size_t stage_mask_bytes = 0;
    long long launches = 0;
    int mode = MODE_PLANES;
    int colpart_rows = 0;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // brackets of the four kernels of run()
    cudaEvent_t ev_prep[2] = {nullptr, nullptr};                         // brackets of the last scs_placement_prepare's kernels
    bool prep_timed = false;
    bool timed = false;
    // ---- interval path (scs_place_interval.cuh): chosen by set_clades when the masks are a laminar family
    int algo = 0;                  // 0 undecided (no clades yet), 1 GEMM, 2 interval
    bool operands_valid = false;   // the GEMM operand plane(s) hold the current genotypes
    bool gemm_ready = false;       // ensure_gemm_buffers has run (operand planes, clade operand, pass-1 partials, tensor maps)
    std::vector<uint32_t> h_bits;  // interval plans small enough for scs_placement_debug_counts: the clade masks, for the dump's GEMM
    int64_t h_bits_pitch = 0;
    uint8_t* gt_own = nullptr;     // plan-owned copy of the packed genotypes (set_genotypes)
    size_t gt_own_bytes = 0;
    const uint8_t* gt_src = nullptr;   // what the interval kernel reads: gt_own or the caller's bound buffer
    int64_t gt_pitch = 0;
    uint16_t* d_perm = nullptr;
    uint32_t* d_lohi = nullptr;
    int* d_slot_row = nullptr;     // clade row of every accumulator slot (-1: none)
    uint4* d_off4 = nullptr;       // variant 3: byte offsets into the prefix counts, per pair of rows
    uint32_t* d_leafmask = nullptr;   // variant 3: positions that have a leaf row, per word of a permuted genotype row
    uint32_t* gt_perm = nullptr;   // variant 3: the genotype rows in depth-first cell order [n_snv][ceil(n_cells/16)]
    size_t gt_perm_bytes = 0;
    bool gt_perm_valid = false;
    int iv_leaf_off = 0;
    int iv_slots = 0, iv_variant = 0, iv_nodes = 0, iv_leaf_pos = 0, iv_stride = 0, iv_entries = 0;
    float* ipart = nullptr;
    size_t ipart_bytes = 0;
    int iv_threads = 0, iv_grid = 0, iv_rows_per_run = 0, iv_runs = 0;
    bool iv_dcache = false, iv_perm_smem = false;
    size_t iv_smem = 0;
    int iv4_np = 0;                // variant 4 (scs_place_iv4.cuh): pairs of non-leaf rows per thread
    float4* iv4_leafw = nullptr;   // variant 4, leaf split: per-SNV leaf weights
    size_t iv4_leafw_bytes = 0;
    // ---- batches of small trees (scs_place_iv5.cuh), algo 3: chosen by scs_placement_set_trees
    int iv5_L = 0, iv5_stride = 0, iv5_ntasks = 0, iv5_grid = 0;
    uint16_t* iv5_perm = nullptr;
    uint32_t* iv5_lohi = nullptr;
    int *iv5_row_slot = nullptr, *iv5_meta = nullptr, *iv5_task0 = nullptr;
    Iv5Task* iv5_tasks = nullptr;
    float* iv5_part = nullptr;
    // ---- scs_placement_prepare: row filter + missing counts (+ cell-order pass) with the counters kept on the device
    std::vector<uint16_t> h_perm;  // host copy of the cell order (positions -> cells)
    uint8_t* d_prep = nullptr;     // mode 1: [kept u64 | pad][active mask, rw words (16-byte padded)][missing per POSITION, 16 rw u32]
    size_t d_prep_bytes = 0;       // mode 2: the scratch layout of gt_reduce_launch (scs_decode.cu)
    int prep_mode = 0;             // 0: nothing prepared yet
};

// Radix packing is the default wherever it applies (<= 15,872 cells); its integer exactness is
// asserted bit-for-bit by tests/test_gpu_parity.py::test_counts_bit_exact_* in both modes.
static int g_default_mode = MODE_RADIX;

static int round_up(int64_t v, int m) { return int((v + m - 1) / m * m); }

static float round_to_bits(double v, int bits) {
    if (v == 0.0) return 0.f;
    int e;
    const double m = frexp(v, &e);  // v = m * 2^e, 0.5 <= |m| < 1
    const double sc = ldexp(1.0, bits);
    return float(ldexp(nearbyint(m * sc) / sc, e));
}

}  // namespace scs

using namespace scs;

extern "C" {

int scs_placement_create(scs_ctx* ctx, int64_t n_snv, int32_t n_cells, int32_t n_rows, scs_placement** out) {
    return scs_placement_create_batch(ctx, 1, &n_snv, n_cells, n_rows, out);
}

int scs_placement_create_batch(scs_ctx* ctx, int32_t n_groups, const int64_t* n_snv_g, int32_t n_cells, int32_t n_rows,
                               scs_placement** out) {
    if (!ctx || !out || !n_snv_g) return set_error(SCS_ERR_ARG, "null argument");
    if (n_groups <= 0 || n_cells <= 0 || n_rows <= 0) return set_error(SCS_ERR_ARG, "empty problem (groups=%d n_cells=%d n_rows=%d)", n_groups, n_cells, n_rows);
    if (n_cells > 65535) return set_error(SCS_ERR_ARG, "n_cells %d exceeds the 16-bit count packing (65535)", n_cells);
    int64_t total = 0, m_pad = 0;
    for (int g = 0; g < n_groups; ++g) {
        if (n_snv_g[g] <= 0) return set_error(SCS_ERR_ARG, "group %d has no SNVs", g);
        total += n_snv_g[g];
        m_pad += (n_snv_g[g] + 2 * BLOCK_M - 1) / (2 * BLOCK_M) * (2 * BLOCK_M);
    }
    if (m_pad > (int64_t(1) << 30)) return set_error(SCS_ERR_ARG, "too many SNV rows for one placement plan; shard by SNV rows / replicates");
    SCS_CUDA(cudaSetDevice(ctx->device));
    Placement* pl = new Placement();
    pl->ctx = ctx;
    pl->n_groups = n_groups;
    pl->n_snv = total;
    pl->n_cells = n_cells;
    pl->n_rows = n_rows;
    pl->M_pad = m_pad;                               // every group padded to row-block pairs (2-CTA kernel)
    pl->N_pad1 = round_up(n_rows, BLOCK_N);
    pl->N_pad = int64_t(pl->N_pad1) * n_groups;
    pl->K_pad = round_up(n_cells, BLOCK_K);
    if (pl->N_pad > (int64_t(1) << 30)) {
        delete pl;
        return set_error(SCS_ERR_ARG, "too many clade rows for one placement plan");
    }
    PlaceParams& p = pl->prm;
    memset(&p, 0, sizeof(p));
    p.M = int(n_snv_g[0]);
    p.N = n_rows;
    p.M_pad = int(pl->M_pad);
    p.N_pad = int(pl->N_pad);
    p.num_k_blocks = pl->K_pad / BLOCK_K;
    // Operand mode: radix packing halves the tensor work but needs K <= 4 chunks of 3968 cells.
    {
        const int nchunk = (p.num_k_blocks + RADIX_KB_PER_CHUNK - 1) / RADIX_KB_PER_CHUNK;
        int mode = nchunk <= RADIX_MAX_CHUNKS ? g_default_mode : MODE_PLANES;
        if (const char* e = getenv("SCS_PLACE_MODE")) {
            if (!strcmp(e, "radix") && nchunk <= RADIX_MAX_CHUNKS) mode = MODE_RADIX;
            if (!strcmp(e, "planes")) mode = MODE_PLANES;
        }
        pl->mode = mode;
        p.nchunk = mode == MODE_RADIX ? nchunk : 1;
    }
    p.num_row_blocks = int(pl->M_pad / BLOCK_M);
    p.num_col_tiles = int(pl->N_pad / BLOCK_N);
    const int ct1 = pl->N_pad1 / BLOCK_N;           // column tiles per group
    pl->num_parts = 2 * ct1;
    // Super-block: the w column tiles whose clade slab the concurrently running CTAs share.
    // With sm_count CTAs in flight the live L2 footprint is  w * slab + (sm_count / w) * 2 planes
    // * rows, smallest near w ~ sqrt(sm_count).  B200's L2 is two ~63 MB halves that replicate
    // lines read from both dies, so this matters: measured at 10k cells x 1M SNVs
    // (profiles/r01_l2_sweep.txt) w = 32 (41 MB slab) runs at a 48% L2 hit rate and 6.6 TB/s of
    // DRAM reads, w = 12..16 at ~225 ms per pass instead of ~330 ms.
    const size_t slab = size_t(BLOCK_N) * pl->K_pad;
    int sbw = std::max(1, int(std::lround(std::sqrt(double(ctx->sm_count)))));
    if (const char* e = getenv("SCS_SLAB_MB")) sbw = int(std::max<size_t>(1, (size_t(std::max(1, atoi(e))) << 20) / slab));
    p.sb_width = std::min(sbw, ct1);
    p.l2_hints = 1;
    if (const char* e = getenv("SCS_L2_HINTS")) p.l2_hints = atoi(e) != 0;
    // Runs: long enough to amortise the end-of-run column reduction, short enough
    // that every SM gets several tasks.
    int run_len = 32;
    if (const char* e = getenv("SCS_RUN_LEN")) run_len = std::max(1, atoi(e));
    while (run_len > 1 && (int64_t(p.num_row_blocks + run_len - 1) / run_len) * ct1 < 4 * ctx->sm_count) run_len >>= 1;
    p.run_len = run_len;

    // ---- the task table: group -> super-block -> run -> column tile
    std::vector<TaskEntry> tasks;
    std::vector<int> unit_group(size_t(pl->M_pad / (2 * BLOCK_M)));
    pl->g_nsnv.assign(n_snv_g, n_snv_g + n_groups);
    pl->g_row0.resize(n_groups);
    pl->g_src0.resize(n_groups);
    pl->g_runs.resize(n_groups);
    int64_t row0 = 0, src0 = 0;
    int max_runs = 0;
    for (int g = 0; g < n_groups; ++g) {
        const int64_t rows_pad = (n_snv_g[g] + 2 * BLOCK_M - 1) / (2 * BLOCK_M) * (2 * BLOCK_M);
        pl->g_row0[g] = row0;
        pl->g_src0[g] = src0;
        for (int64_t u = row0 / (2 * BLOCK_M); u < (row0 + rows_pad) / (2 * BLOCK_M); ++u) unit_group[size_t(u)] = g;
        const int rb0 = int(row0 / BLOCK_M), rbn = int(rows_pad / BLOCK_M);
        const int runs = (rbn + run_len - 1) / run_len;
        pl->g_runs[g] = runs;
        max_runs = std::max(max_runs, runs);
        for (int sb0 = 0; sb0 < ct1; sb0 += p.sb_width) {
            const int w = std::min(p.sb_width, ct1 - sb0);
            for (int run = 0; run < runs; ++run)
                for (int jj = 0; jj < w; ++jj) {
                    TaskEntry te;
                    te.jl = sb0 + jj;
                    te.j = g * ct1 + te.jl;
                    te.r0 = rb0 + run * run_len;
                    te.r1 = rb0 + std::min(rbn, (run + 1) * run_len);
                    te.run_local = run;
                    te.nvalid = std::max(0, std::min(BLOCK_N, n_rows - te.jl * BLOCK_N));
                    te.group = g;
                    te.pad = 0;
                    tasks.push_back(te);
                }
        }
        row0 += rows_pad;
        src0 += n_snv_g[g];
    }
    p.num_runs = max_runs;
    p.num_tasks = int(tasks.size());
    pl->grid = std::min(ctx->sm_count, p.num_tasks);
    pl->colpart_rows = max_runs;

    int rc = SCS_OK;
    do {
        std::vector<long long> row0v(pl->g_row0.begin(), pl->g_row0.end()), src0v(pl->g_src0.begin(), pl->g_src0.end());
        std::vector<int> nsnvv(n_groups);
        for (int g = 0; g < n_groups; ++g) nsnvv[g] = int(n_snv_g[g]);
        // (the GEMM's own buffers -- operand planes, clade operand, pass-1 partials: 15 GB at 10k cells x 1M SNVs -- are
        // allocated by ensure_gemm_buffers when the GEMM path is first taken; interval plans never hold them)
        if (cudaMalloc(&pl->rowref, size_t(pl->M_pad) * sizeof(uint2)) != cudaSuccess ||
            cudaMalloc(&pl->colpart, size_t(pl->colpart_rows) * pl->N_pad * sizeof(float)) != cudaSuccess ||
            cudaMalloc(&pl->y, size_t(n_groups) * n_rows * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&pl->row_keep, size_t(total)) != cudaSuccess ||
            cudaMalloc(&pl->d_group_of_unit, unit_group.size() * sizeof(int)) != cudaSuccess ||
            cudaMalloc(&pl->d_group_row0, size_t(n_groups) * sizeof(long long)) != cudaSuccess ||
            cudaMalloc(&pl->d_group_src0, size_t(n_groups) * sizeof(long long)) != cudaSuccess ||
            cudaMalloc(&pl->d_group_nsnv, size_t(n_groups) * sizeof(int)) != cudaSuccess ||
            cudaMalloc(&pl->d_group_runs, size_t(n_groups) * sizeof(int)) != cudaSuccess ||
            cudaMalloc(&pl->d_tasks, tasks.size() * sizeof(TaskEntry)) != cudaSuccess ||
            cudaMalloc(&pl->d_gconst, size_t(n_groups) * sizeof(GroupConst)) != cudaSuccess ||
            cudaMalloc(&pl->d_gconst_d, size_t(n_groups) * sizeof(double2)) != cudaSuccess) {
            rc = set_error(SCS_ERR_OOM, "device allocation failed for placement plan (%s)", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        cudaMemcpy(pl->d_group_of_unit, unit_group.data(), unit_group.size() * sizeof(int), cudaMemcpyHostToDevice);
        cudaMemcpy(pl->d_group_row0, row0v.data(), size_t(n_groups) * sizeof(long long), cudaMemcpyHostToDevice);
        cudaMemcpy(pl->d_group_src0, src0v.data(), size_t(n_groups) * sizeof(long long), cudaMemcpyHostToDevice);
        cudaMemcpy(pl->d_group_nsnv, nsnvv.data(), size_t(n_groups) * sizeof(int), cudaMemcpyHostToDevice);
        cudaMemcpy(pl->d_group_runs, pl->g_runs.data(), size_t(n_groups) * sizeof(int), cudaMemcpyHostToDevice);
        cudaMemcpy(pl->d_tasks, tasks.data(), tasks.size() * sizeof(TaskEntry), cudaMemcpyHostToDevice);
        pl->rowmap.group_of_unit = pl->d_group_of_unit;
        pl->rowmap.group_row0 = pl->d_group_row0;
        pl->rowmap.group_src0 = pl->d_group_src0;
        pl->rowmap.group_nsnv = pl->d_group_nsnv;
        p.rowref = pl->rowref;
        p.colpart = pl->colpart;
        p.tasks = pl->d_tasks;
        p.gconst = pl->d_gconst;
        set_place_smem<PASS_DUMP>();
        set_place_smem<PASS_STATS>();
        set_place_smem<PASS_COLSUM>();
        for (int i = 0; i < 5; ++i) cudaEventCreate(&pl->ev[i]);
        for (int i = 0; i < 2; ++i) cudaEventCreate(&pl->ev_prep[i]);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            rc = set_error(SCS_ERR_CUDA, "placement plan setup failed: %s", cudaGetErrorString(e));
            break;
        }
    } while (0);
    if (rc != SCS_OK) {
        scs_placement_destroy(reinterpret_cast<scs_placement*>(pl));
        return rc;
    }
    *out = reinterpret_cast<scs_placement*>(pl);
    return SCS_OK;
}

int scs_placement_destroy(scs_placement* h) {
    Placement* pl = reinterpret_cast<Placement*>(h);
    if (!pl) return SCS_OK;
    cudaFree(pl->x1);
    cudaFree(pl->x0);
    cudaFree(pl->s);
    cudaFree(pl->stats);
    cudaFree(pl->rowref);
    cudaFree(pl->rowref_f);
    cudaFree(pl->cache);
    cudaFree(pl->gt_own);
    cudaFree(pl->d_perm);
    cudaFree(pl->d_lohi);
    cudaFree(pl->d_slot_row);
    cudaFree(pl->d_off4);
    cudaFree(pl->d_leafmask);
    cudaFree(pl->gt_perm);
    cudaFree(pl->ipart);
    cudaFree(pl->iv4_leafw);
    cudaFree(pl->iv5_perm);
    cudaFree(pl->iv5_lohi);
    cudaFree(pl->iv5_row_slot);
    cudaFree(pl->iv5_meta);
    cudaFree(pl->iv5_task0);
    cudaFree(pl->iv5_tasks);
    cudaFree(pl->iv5_part);
    cudaFree(pl->d_prep);
    cudaFree(pl->colpart);
    cudaFree(pl->y);
    cudaFree(pl->row_keep);
    cudaFree(pl->dump);
    cudaFree(pl->stage_gt);
    cudaFree(pl->stage_mask);
    cudaFree(pl->d_group_of_unit);
    cudaFree(pl->d_group_row0);
    cudaFree(pl->d_group_src0);
    cudaFree(pl->d_group_nsnv);
    cudaFree(pl->d_group_runs);
    cudaFree(pl->d_tasks);
    cudaFree(pl->d_gconst);
    cudaFree(pl->d_gconst_d);
    for (int i = 0; i < 5; ++i)
        if (pl->ev[i]) cudaEventDestroy(pl->ev[i]);
    for (int i = 0; i < 2; ++i)
        if (pl->ev_prep[i]) cudaEventDestroy(pl->ev_prep[i]);
    delete pl;
    return SCS_OK;
}

// The GEMM path's buffers and tensor maps, on first use: operand plane(s) X, clade operand S (both zero padded) and the
// pass-1 partials.  A plan that stays on an interval path never calls this.
static int ensure_gemm_buffers(Placement* pl, cudaStream_t st) {
    if (pl->gemm_ready) return SCS_OK;
    const size_t plane = size_t(pl->M_pad) * pl->K_pad;
    const size_t sbytes = size_t(pl->N_pad) * pl->K_pad;
    if (cudaMalloc(&pl->x1, plane) != cudaSuccess || (pl->mode == MODE_PLANES && cudaMalloc(&pl->x0, plane) != cudaSuccess) ||
        cudaMalloc(&pl->s, sbytes) != cudaSuccess ||
        cudaMalloc(&pl->stats, size_t(pl->num_parts) * pl->M_pad * sizeof(uint2)) != cudaSuccess) {
        const char* why = cudaGetErrorString(cudaGetLastError());
        cudaFree(pl->x1);
        cudaFree(pl->x0);
        cudaFree(pl->s);
        cudaFree(pl->stats);
        pl->x1 = pl->x0 = pl->s = nullptr;
        pl->stats = nullptr;
        return set_error(SCS_ERR_OOM, "device allocation failed for the GEMM operands of a placement plan (%s)", why);
    }
    SCS_CUDA(cudaMemsetAsync(pl->x1, 0, plane, st));
    if (pl->x0) SCS_CUDA(cudaMemsetAsync(pl->x0, 0, plane, st));
    SCS_CUDA(cudaMemsetAsync(pl->s, 0, sbytes, st));
    pl->prm.stats = pl->stats;
    int rc;
    if ((rc = make_u8_tmap(&pl->tm_x1, pl->x1, pl->M_pad, pl->K_pad, BLOCK_M)) != SCS_OK) return rc;
    if (pl->x0) {
        if ((rc = make_u8_tmap(&pl->tm_x0, pl->x0, pl->M_pad, pl->K_pad, BLOCK_M)) != SCS_OK) return rc;
    } else {
        pl->tm_x0 = pl->tm_x1;
    }
    if ((rc = make_u8_tmap(&pl->tm_s, pl->s, pl->N_pad, pl->K_pad, BLOCK_N)) != SCS_OK) return rc;
    pl->gemm_ready = true;
    return SCS_OK;
}

// GEMM operand plane(s) from the packed matrix
static int expand_operands(Placement* pl, const uint8_t* d_gt, int64_t pitch_bytes, cudaStream_t st) {
    {
        const int erc = ensure_gemm_buffers(pl, st);
        if (erc != SCS_OK) return erc;
    }
    const int64_t work = pl->M_pad * (pl->K_pad / 16);
    if (pl->K_pad / 16 < 64 && work < (int64_t(1) << 31)) {
        const unsigned grid = unsigned(std::min<int64_t>((work + 255) / 256, int64_t(pl->ctx->sm_count) * 64));
        scs_expand_genotypes_flat_kernel<<<grid, 256, 0, st>>>(d_gt, pitch_bytes, pl->M_pad, pl->n_cells, pl->rowmap, pl->x1, pl->x0,
                                                               pl->K_pad, pl->mode == MODE_RADIX);
    } else {
        dim3 block(128), grid((pl->K_pad / 16 + 127) / 128, unsigned(std::min<int64_t>(pl->M_pad, 32768)));
        scs_expand_genotypes_kernel<<<grid, block, 0, st>>>(d_gt, pitch_bytes, pl->M_pad, pl->n_cells, pl->rowmap, pl->x1, pl->x0, pl->K_pad,
                                                             pl->mode == MODE_RADIX);
    }
    pl->launches++;
    pl->operands_valid = true;
    SCS_CUDA(cudaGetLastError());
    return SCS_OK;
}

static int set_genotypes_impl(Placement* pl, const uint8_t* d_gt, int64_t pitch_bytes, const uint8_t* d_row_keep, cudaStream_t st, bool bind) {
    if (!pl || !d_gt) return set_error(SCS_ERR_ARG, "null argument");
    if (pitch_bytes < (pl->n_cells + 3) / 4) return set_error(SCS_ERR_ARG, "genotype row pitch %lld is smaller than ceil(n_cells/4)", (long long)pitch_bytes);
    if ((pitch_bytes & 3) || (reinterpret_cast<uintptr_t>(d_gt) & 3)) return set_error(SCS_ERR_ARG, "packed genotype rows must be 4-byte aligned (pitch %lld)", (long long)pitch_bytes);
    pl->operands_valid = false;
    pl->gt_perm_valid = false;
    pl->gt_src = nullptr;
    if (pl->algo == 3) {
        // batches of small trees: the kernel reads the caller's packed rows as they are (they must stay alive and
        // unchanged until the run has finished, like scs_placement_bind_genotypes)
        pl->gt_src = d_gt;
        pl->gt_pitch = pitch_bytes;
        pl->gt_perm_valid = true;
        if (d_row_keep) {
            SCS_CUDA(cudaMemcpyAsync(pl->row_keep, d_row_keep, size_t(pl->n_snv), cudaMemcpyDeviceToDevice, st));
            pl->have_row_keep = true;
        } else {
            pl->have_row_keep = false;
        }
        return SCS_OK;
    }
    // GEMM path: the operand plane(s).  Clades not seen yet: a single dataset keeps the packed matrix (below) and builds
    // the planes at the first run if the masks turn out to need the GEMM; a batch has no such copy and builds them now.
    bool need_operands = pl->algo == 1 || (pl->algo == 0 && pl->n_groups != 1);
    // interval path (or clades not seen yet): the packed matrix itself -- bound (the caller keeps the buffer
    // alive and unchanged until the run has finished) or copied into the plan
    if (pl->algo != 1 && pl->n_groups == 1) {
        if (bind && (pl->algo == 2 || d_gt == pl->stage_gt)) {      // (the staging buffer of set_genotypes_host is the plan's own)
            pl->gt_src = d_gt;
        } else {
            const size_t bytes = size_t(pl->n_snv) * size_t(pitch_bytes);
            if (bytes > pl->gt_own_bytes) {
                cudaFree(pl->gt_own);
                pl->gt_own = nullptr;
                pl->gt_own_bytes = 0;
                if (cudaMalloc(&pl->gt_own, bytes) != cudaSuccess) {
                    cudaGetLastError();
                    if (pl->algo == 2) return set_error(SCS_ERR_OOM, "device allocation of %zu bytes for the packed genotypes failed", bytes);
                } else {
                    pl->gt_own_bytes = bytes;
                }
            }
            if (pl->gt_own) {
                SCS_CUDA(cudaMemcpyAsync(pl->gt_own, d_gt, bytes, cudaMemcpyDeviceToDevice, st));
                pl->gt_src = pl->gt_own;
            }
        }
        pl->gt_pitch = pitch_bytes;
        if (!pl->gt_src) need_operands = true;      // (no room for a copy of the packed matrix)
    }
    if (need_operands) {
        int rc = expand_operands(pl, d_gt, pitch_bytes, st);
        if (rc != SCS_OK) return rc;
    }
    if (d_row_keep) {
        SCS_CUDA(cudaMemcpyAsync(pl->row_keep, d_row_keep, size_t(pl->n_snv), cudaMemcpyDeviceToDevice, st));
        pl->have_row_keep = true;
    } else {
        pl->have_row_keep = false;
    }
    SCS_CUDA(cudaGetLastError());
    return SCS_OK;
}

// Host-only: the cell order and the intervals the interval path would use for these masks.
int scs_clade_intervals(const uint32_t* bits, int64_t pitch_words, int32_t n_rows, int32_t n_cells, uint16_t* perm, uint32_t* lohi,
                        int32_t* laminar) {
    if (!bits || !laminar || n_rows <= 0 || n_cells <= 0 || n_cells > 65535 || pitch_words < (n_cells + 31) / 32)
        return set_error(SCS_ERR_ARG, "bad argument");
    std::vector<uint16_t> pv;
    std::vector<uint32_t> lv;
    *laminar = clade_intervals(bits, pitch_words, n_rows, n_cells, pv, lv) ? 1 : 0;
    if (*laminar) {
        if (perm) memcpy(perm, pv.data(), pv.size() * sizeof(uint16_t));
        if (lohi) memcpy(lohi, lv.data(), lv.size() * sizeof(uint32_t));
    }
    return SCS_OK;
}

int scs_placement_set_genotypes(scs_placement* h, const uint8_t* d_gt, int64_t pitch_bytes, const uint8_t* d_row_keep, void* stream) {
    return set_genotypes_impl(reinterpret_cast<Placement*>(h), d_gt, pitch_bytes, d_row_keep, reinterpret_cast<cudaStream_t>(stream), false);
}

int scs_placement_bind_genotypes(scs_placement* h, const uint8_t* d_gt, int64_t pitch_bytes, const uint8_t* d_row_keep, void* stream) {
    return set_genotypes_impl(reinterpret_cast<Placement*>(h), d_gt, pitch_bytes, d_row_keep, reinterpret_cast<cudaStream_t>(stream), true);
}

// Decide the algorithm for these clade masks (host copy `bits`): interval path when they are a laminar
// family, the plan is a single dataset of at least 100 cells and the working set fits shared memory;
// SCS_PLACE_ALGO = gemm | interval | auto overrides ("interval": whenever the masks are laminar, at any size).
constexpr int IV_AUTO_MIN_CELLS = 100;
static int choose_algo(Placement* pl, const uint32_t* bits, int64_t pitch_words, cudaStream_t st) {
    pl->algo = 1;
    pl->gt_perm_valid = false;      // (the cell order may change with the masks)
    const char* e = getenv("SCS_PLACE_ALGO");
    if (e && !strcmp(e, "gemm")) return SCS_OK;
    if (pl->n_groups != 1) return SCS_OK;
    // Where each algorithm wins was measured, not assumed (tools/gpu_r2_s18.sh: bench.py --algo gemm | interval, one PT test
    // over 100k SNVs, ms per step GEMM / interval): 120 cells 0.47 / 0.46, 250 0.52 / 0.49, 400 0.51 / 0.45, 500 0.57 / 0.48,
    // 1,000 0.86 / 0.62, 2,000 1.86 / 0.96, 3,900 4.46 / 1.71, 10,000 173 / 26.  (Round 1 had the GEMM ahead below 3,968
    // cells: the per-row barriers of interval variant 3, then the pre-pass and the flush of thousands of tiny runs, were
    // what lost there -- variant 4, the sub-block pre-pass and whole-wave runs removed them.)  "auto" takes the interval
    // path for every tree-shaped matrix of at least 100 cells; below that the step is all fixed cost either way and the GEMM
    // spares the host-side interval analysis.  SCS_PLACE_ALGO=interval takes it at any size.
    const bool forced = e && (!strcmp(e, "interval") || !strcmp(e, "interval1") || !strcmp(e, "interval3"));
    if (!forced && pl->n_cells < IV_AUTO_MIN_CELLS) return SCS_OK;
    const size_t SMEM_MAX = 227 * 1024;
    // cheapest test first: does even the smallest working set (prefix counts + accumulators) fit shared memory?
    if (interval_smem_bytes(pl->n_cells, pl->n_rows, false, false) > SMEM_MAX) return SCS_OK;
    std::vector<uint16_t> perm;
    std::vector<uint32_t> lohi;
    if (!clade_intervals(bits, pitch_words, pl->n_rows, pl->n_cells, perm, lohi)) return SCS_OK;
    const int n_cells = pl->n_cells, n_rows = pl->n_rows;
    int threads = n_rows >= 4096 ? 1024 : (n_rows >= 1024 ? 512 : 256);
    if (const char* t = getenv("SCS_IV_THREADS")) threads = std::max(32, std::min(1024, atoi(t) / 32 * 32));

    // ---- variant 3 (leaves out of the sweeps, count cache in shared memory) when it fits
    std::vector<int> leaf_row_of_pos(size_t(n_cells), -1), general;
    for (int r = 0; r < n_rows; ++r) {
        const int lo = int(lohi[size_t(r)] & 0xFFFFu), hi = int(lohi[size_t(r)] >> 16);
        if (hi - lo == 1 && leaf_row_of_pos[size_t(lo)] < 0) leaf_row_of_pos[size_t(lo)] = r;
        else general.push_back(r);
    }
    // the sweeps walk the rows sorted by (lo, hi): neighbouring lanes then look up neighbouring prefix counts
    // instead of 32 random shared-memory banks
    auto by_interval = [&](int a, int b) {
        const uint32_t la = lohi[size_t(a)] & 0xFFFFu, lb = lohi[size_t(b)] & 0xFFFFu;
        return la != lb ? la < lb : (lohi[size_t(a)] >> 16) < (lohi[size_t(b)] >> 16);
    };
    std::stable_sort(general.begin(), general.end(), by_interval);
    const int n_gen = int(general.size()), n_slots3 = (n_gen + 1) & ~1;
    const int rw = (n_cells + 15) / 16;                     // 32-bit words of a permuted row: one thread each
    // one thread per word at least; when the words fill more than half of the default CTA, exactly one thread per
    // word: every thread then carries the same per-position work and nobody idles at the barriers around it
    // (measured at 10k cells: 640 threads 11.12 ms, 768 11.20 ms, 1024 11.66 ms per 250k SNVs)
    const int rw32 = (rw + 31) / 32 * 32;
    int threads3 = std::max(threads, rw32);
    if (!getenv("SCS_IV_THREADS") && 2 * rw32 >= threads) threads3 = rw32;
    const bool v3 = threads3 <= 1024 && interval3_smem_bytes(n_cells, n_slots3) <= SMEM_MAX && !(e && !strcmp(e, "interval1"));
    std::vector<int> slot_row;
    size_t smem = 0;
    int per_sm = 1;
    if (v3) {
        threads = threads3;
        // variant 4 (scs_place_iv4.cuh: the same tables, every per-row quantity in registers) when the shape allows:
        // one thread per word of a permuted row, at most 8 pairs of non-leaf rows per thread, at most 640 threads
        pl->iv4_np = 0;
        int t4 = 0;
        if (!(e && !strcmp(e, "interval3")) && n_slots3 >= 2) {
            const int n_pairs = n_slots3 / 2;
            t4 = std::max(rw32, (((n_pairs + IV4_MAX_NP - 1) / IV4_MAX_NP) + 31) / 32 * 32);
            if (const char* tt = getenv("SCS_IV_THREADS")) t4 = std::max(t4, std::min(1024, atoi(tt) / 32 * 32));
            const int np = (n_pairs + t4 - 1) / t4;
            if (t4 <= IV4_MAX_THREADS && np >= 1 && np <= IV4_MAX_NP && iv4_smem_bytes(rw, t4) <= SMEM_MAX) {
                const int occ = iv4_occupancy(np, t4, iv4_smem_bytes(rw, t4));
                if (occ > 0) {
                    pl->iv4_np = np;
                    per_sm = occ;
                }
            }
        }
        const bool v4 = pl->iv4_np > 0;
        // byte offset of prefix count P[k] in shared memory: variant 3 keeps P[0] at word 3; variant 4 keeps a zero word
        // at byte 0 and swizzles the 16-byte chunks of P[1..] (iv4_p_offset) so that its stores are bank-conflict free
        auto poff = [&](uint32_t k) -> uint32_t { return v4 ? iv4_p_offset(k) : k * 4u; };
        std::vector<uint4> off4(size_t(n_slots3 / 2), make_uint4(poff(0), poff(0), poff(0), poff(0)));
        for (int sidx = 0; sidx < n_gen; ++sidx) {
            const uint32_t lh = lohi[size_t(general[size_t(sidx)])];
            uint4& o = off4[size_t(sidx >> 1)];
            if (sidx & 1) {
                o.z = poff(lh & 0xFFFFu);
                o.w = poff(lh >> 16);
            } else {
                o.x = poff(lh & 0xFFFFu);
                o.y = poff(lh >> 16);
            }
        }
        int n_leaf_pos = 0;
        std::vector<uint32_t> leafmask(size_t(rw), 0u);
        for (int k = 0; k < n_cells; ++k)
            if (leaf_row_of_pos[size_t(k)] >= 0) {
                leafmask[size_t(k >> 4)] |= 1u << (2 * (k & 15));
                ++n_leaf_pos;
            }
        const int leaf_off = (n_slots3 + 3) & ~3;           // the leaf section of a partial row starts 16-byte aligned
        slot_row.assign(size_t(leaf_off + 16 * rw), -1);
        for (int sidx = 0; sidx < n_gen; ++sidx) slot_row[size_t(sidx)] = general[size_t(sidx)];
        for (int k = 0; k < n_cells; ++k) slot_row[size_t(leaf_off + k)] = leaf_row_of_pos[size_t(k)];
        cudaFree(pl->d_off4);
        cudaFree(pl->d_leafmask);
        pl->d_off4 = nullptr;
        pl->d_leafmask = nullptr;
        if (cudaMalloc(&pl->d_off4, std::max<size_t>(1, off4.size()) * sizeof(uint4)) != cudaSuccess ||
            cudaMalloc(&pl->d_leafmask, size_t(rw) * sizeof(uint32_t)) != cudaSuccess)
            return set_error(SCS_ERR_OOM, "device allocation failed");
        SCS_CUDA(cudaMemcpyAsync(pl->d_off4, off4.data(), off4.size() * sizeof(uint4), cudaMemcpyHostToDevice, st));
        SCS_CUDA(cudaMemcpyAsync(pl->d_leafmask, leafmask.data(), leafmask.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        SCS_CUDA(cudaStreamSynchronize(st));
        pl->iv_variant = 3;
        pl->iv_nodes = n_gen;
        pl->iv_slots = n_slots3;
        pl->iv_leaf_pos = n_leaf_pos;
        pl->iv_stride = leaf_off + 16 * rw;
        pl->iv_leaf_off = leaf_off;
        if (v4) {
            threads = t4;
            smem = iv4_smem_bytes(rw, t4);
            pl->iv_variant = 4;
        }
        if (pl->iv4_np == 0) {
            smem = interval3_smem_bytes(n_cells, n_slots3);
            if (threads <= 640) {
                cudaFuncSetAttribute(scs_place_interval3_kernel<640, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scs_place_interval3_kernel<640, 2>, threads, smem);
            } else {
                cudaFuncSetAttribute(scs_place_interval3_kernel<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scs_place_interval3_kernel<1024, 1>, threads, smem);
            }
        }
    } else {
        // ---- variant 1: every row through the sweeps, counts recomputed in the second sweep
        std::vector<int> all(size_t(n_rows), 0);
        std::iota(all.begin(), all.end(), 0);
        std::stable_sort(all.begin(), all.end(), by_interval);
        const int n_slots1 = (n_rows + 1) & ~1;
        std::vector<uint32_t> lohi_sorted(size_t(n_slots1), 0u);
        for (int sidx = 0; sidx < n_rows; ++sidx) lohi_sorted[size_t(sidx)] = lohi[size_t(all[size_t(sidx)])];
        slot_row.assign(size_t(n_slots1), -1);
        for (int sidx = 0; sidx < n_rows; ++sidx) slot_row[size_t(sidx)] = all[size_t(sidx)];
        cudaFree(pl->d_lohi);
        pl->d_lohi = nullptr;
        if (cudaMalloc(&pl->d_lohi, size_t(n_slots1) * sizeof(uint32_t)) != cudaSuccess) return set_error(SCS_ERR_OOM, "device allocation failed");
        SCS_CUDA(cudaMemcpyAsync(pl->d_lohi, lohi_sorted.data(), lohi_sorted.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        SCS_CUDA(cudaStreamSynchronize(st));
        pl->iv_variant = 1;
        pl->iv_nodes = n_rows;
        pl->iv_slots = n_slots1;
        pl->iv_leaf_pos = 0;
        pl->iv_stride = n_slots1;
        pl->iv_perm_smem = interval_smem_bytes(n_cells, n_rows, false, true) <= SMEM_MAX;
        smem = interval_smem_bytes(n_cells, n_rows, false, pl->iv_perm_smem);
        cudaFuncSetAttribute(scs_place_interval_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scs_place_interval_kernel<false>, threads, smem);
    }
    cudaFree(pl->d_perm);
    cudaFree(pl->d_slot_row);
    pl->d_perm = nullptr;
    pl->d_slot_row = nullptr;
    if (cudaMalloc(&pl->d_perm, size_t(n_cells) * sizeof(uint16_t)) != cudaSuccess ||
        cudaMalloc(&pl->d_slot_row, slot_row.size() * sizeof(int)) != cudaSuccess)
        return set_error(SCS_ERR_OOM, "device allocation failed");
    pl->h_perm = perm;
    pl->prep_mode = 0;
    SCS_CUDA(cudaMemcpyAsync(pl->d_perm, perm.data(), perm.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, st));
    SCS_CUDA(cudaMemcpyAsync(pl->d_slot_row, slot_row.data(), slot_row.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    SCS_CUDA(cudaStreamSynchronize(st));          // the host vectors are stack-local
    pl->iv_entries = int(slot_row.size());
    pl->iv_smem = smem;
    pl->iv_threads = threads;
    per_sm = std::max(1, per_sm);
    const int64_t slots = int64_t(pl->ctx->sm_count) * per_sm;
    // runs: short fp32 accumulation chains (<= 256 rows), a whole number of waves of the resident CTAs where the matrix
    // allows it (a run flushes every accumulator: thousands of 17-row runs -- the round-1 rule "four runs per CTA" at 1,000
    // cells x 100k SNVs -- cost more in partial sums, 70 MB, than the rows themselves)
    const int64_t waves = std::max<int64_t>(1, (pl->n_snv + 256 * slots - 1) / (256 * slots));
    int rpr = int(std::min<int64_t>(256, std::max<int64_t>(16, (pl->n_snv + slots * waves - 1) / (slots * waves))));
    if (const char* t = getenv("SCS_IV_ROWS_PER_RUN")) rpr = std::max(1, atoi(t));
    pl->iv_rows_per_run = rpr;
    pl->iv_runs = int((pl->n_snv + rpr - 1) / rpr);
    pl->iv_grid = int(std::min<int64_t>(pl->iv_runs, slots));
    const size_t pbytes = size_t(pl->iv_runs) * size_t(pl->iv_stride) * sizeof(float);
    if (pbytes > pl->ipart_bytes) {
        cudaFree(pl->ipart);
        pl->ipart = nullptr;
        pl->ipart_bytes = 0;
        if (cudaMalloc(&pl->ipart, pbytes) != cudaSuccess) return set_error(SCS_ERR_OOM, "device allocation of %zu bytes failed", pbytes);
        pl->ipart_bytes = pbytes;
    }
    SCS_CUDA(cudaGetLastError());
    pl->algo = 2;
    return SCS_OK;
}

int scs_placement_set_clades(scs_placement* h, const uint32_t* d_bits, int64_t pitch_words, void* stream) {
    Placement* pl = reinterpret_cast<Placement*>(h);
    if (!pl || !d_bits) return set_error(SCS_ERR_ARG, "null argument");
    if (pitch_words < (pl->n_cells + 31) / 32) return set_error(SCS_ERR_ARG, "clade mask pitch %lld is smaller than ceil(n_cells/32)", (long long)pitch_words);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t total_rows = int64_t(pl->n_groups) * pl->n_rows;
    if (pl->n_groups == 1) {
        // which algorithm these masks allow is decided on a host copy (tree-shaped masks -> interval path)
        std::vector<uint32_t> hb(size_t(pl->n_rows) * size_t(pitch_words));
        SCS_CUDA(cudaMemcpyAsync(hb.data(), d_bits, hb.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SCS_CUDA(cudaStreamSynchronize(st));
        const int prev = pl->algo;
        int rc = choose_algo(pl, hb.data(), pitch_words, st);
        if (rc != SCS_OK) return rc;
        if (prev == 2 && pl->algo == 1) pl->operands_valid = false;   // genotypes were set for the interval path only
        pl->h_bits.clear();
        if (pl->algo == 2 && size_t(pl->M_pad) * pl->N_pad * 2 <= (size_t(1) << 28)) {
            pl->h_bits.swap(hb);
            pl->h_bits_pitch = pitch_words;
        }
    } else {
        pl->algo = 1;
    }
    if (pl->algo == 1) {
        // the GEMM's clade operand (interval plans never read it)
        const int erc = ensure_gemm_buffers(pl, st);
        if (erc != SCS_OK) return erc;
        dim3 block(128), grid((pl->K_pad / 4 + 127) / 128, unsigned(std::min<int64_t>(total_rows, 32768)));
        scs_expand_clades_kernel<<<grid, block, 0, st>>>(d_bits, pitch_words, total_rows, pl->n_rows, pl->N_pad1, pl->n_cells, pl->s, pl->K_pad,
                                                          pl->mode == MODE_RADIX ? uint32_t(E5M2_ONE) : 1u);
        pl->launches++;
        SCS_CUDA(cudaGetLastError());
    }
    return SCS_OK;
}

// Clade rows given as TREES (scs_tree_parse_batch's outputs on the device): every group's rows are the nodes of a binary
// tree, so the batch can take the interval path for small trees (scs_place_iv5.cuh) -- no operand planes, no GEMM.  Falls
// back to scs_placement_set_clades (the masks alone, tcgen05 GEMM) for more than 512 cells or SCS_PLACE_ALGO=gemm.
int scs_placement_set_trees(scs_placement* h, const uint32_t* d_bits, int64_t pitch_words, const int32_t* d_children, const int32_t* d_info,
                            void* stream) {
    Placement* pl = reinterpret_cast<Placement*>(h);
    if (!pl || !d_bits || !d_children || !d_info) return set_error(SCS_ERR_ARG, "null argument");
    if (pitch_words < (pl->n_cells + 31) / 32) return set_error(SCS_ERR_ARG, "clade mask pitch %lld is smaller than ceil(n_cells/32)", (long long)pitch_words);
    const char* e = getenv("SCS_PLACE_ALGO");
    if (pl->n_cells > IV5_MAX_CELLS || pl->n_rows > 4 * IV5_MAX_CELLS || (e && !strcmp(e, "gemm")))
        return scs_placement_set_clades(h, d_bits, pitch_words, stream);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int G = pl->n_groups;
    if (pl->iv5_L == 0) {
        const int L = pl->n_cells <= 128 ? 8 : (pl->n_cells <= 256 ? 16 : 32);
        const int stride = IV5_MAX_NS * L + 16 * L + 1;
        int rows_per_task = 2048;
        if (const char* ev = getenv("SCS_IV5_ROWS")) rows_per_task = std::max(IV5_THREADS / L, atoi(ev));
        std::vector<Iv5Task> tasks;
        std::vector<int> task0(size_t(G) + 1, 0);
        for (int g = 0; g < G; ++g) {
            task0[size_t(g)] = int(tasks.size());
            for (int64_t r = 0; r < pl->g_nsnv[size_t(g)]; r += rows_per_task)
                tasks.push_back(Iv5Task{g, int(std::min<int64_t>(rows_per_task, pl->g_nsnv[size_t(g)] - r)), pl->g_src0[size_t(g)] + r});
        }
        task0[size_t(G)] = int(tasks.size());
        if (cudaMalloc(&pl->iv5_perm, size_t(G) * 16 * L * sizeof(uint16_t)) != cudaSuccess ||
            cudaMalloc(&pl->iv5_lohi, size_t(G) * IV5_MAX_NS * L * sizeof(uint32_t)) != cudaSuccess ||
            cudaMalloc(&pl->iv5_row_slot, size_t(G) * pl->n_rows * sizeof(int)) != cudaSuccess ||
            cudaMalloc(&pl->iv5_meta, size_t(G) * 4 * sizeof(int)) != cudaSuccess ||
            cudaMalloc(&pl->iv5_task0, task0.size() * sizeof(int)) != cudaSuccess ||
            cudaMalloc(&pl->iv5_tasks, tasks.size() * sizeof(Iv5Task)) != cudaSuccess ||
            cudaMalloc(&pl->iv5_part, tasks.size() * size_t(stride) * sizeof(float)) != cudaSuccess)
            return set_error(SCS_ERR_OOM, "device allocation failed for the batch interval path (%s)", cudaGetErrorString(cudaGetLastError()));
        SCS_CUDA(cudaMemcpy(pl->iv5_task0, task0.data(), task0.size() * sizeof(int), cudaMemcpyHostToDevice));
        SCS_CUDA(cudaMemcpy(pl->iv5_tasks, tasks.data(), tasks.size() * sizeof(Iv5Task), cudaMemcpyHostToDevice));
        pl->iv5_L = L;
        pl->iv5_stride = stride;
        pl->iv5_ntasks = int(tasks.size());
        pl->iv5_grid = int(std::min<int64_t>(int64_t(tasks.size()), int64_t(pl->ctx->sm_count) * IV5_MIN_CTAS));
    }
    const size_t smem = (size_t(7) * pl->n_rows + 2) * sizeof(int);
    if (smem > 48 * 1024) cudaFuncSetAttribute(scs_iv5_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    scs_iv5_tables_kernel<<<G, 64, smem, st>>>(d_bits, pitch_words, d_children, d_info, pl->n_rows, pl->n_cells, pl->iv5_L, pl->iv5_perm,
                                                pl->iv5_lohi, pl->iv5_row_slot, pl->iv5_meta);
    SCS_CUDA(cudaGetLastError());
    pl->launches++;
    pl->algo = 3;
    pl->operands_valid = false;
    pl->gt_perm_valid = false;
    return SCS_OK;
}

int scs_placement_set_genotypes_host(scs_placement* h, const uint8_t* gt, int64_t pitch_bytes, const uint8_t* row_keep, void* stream) {
    Placement* pl = reinterpret_cast<Placement*>(h);
    if (!pl || !gt) return set_error(SCS_ERR_ARG, "null argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t bytes = size_t(pl->n_snv) * pitch_bytes + (row_keep ? size_t(pl->n_snv) : 0);
    if (bytes > pl->stage_gt_bytes) {
        cudaFree(pl->stage_gt);
        pl->stage_gt = nullptr;
        pl->stage_gt_bytes = 0;
        if (cudaMalloc(&pl->stage_gt, bytes) != cudaSuccess) return set_error(SCS_ERR_OOM, "device staging allocation of %zu bytes failed", bytes);
        pl->stage_gt_bytes = bytes;
    }
    SCS_CUDA(cudaMemcpyAsync(pl->stage_gt, gt, size_t(pl->n_snv) * pitch_bytes, cudaMemcpyHostToDevice, st));
    uint8_t* d_keep = nullptr;
    if (row_keep) {
        d_keep = pl->stage_gt + size_t(pl->n_snv) * pitch_bytes;
        SCS_CUDA(cudaMemcpyAsync(d_keep, row_keep, size_t(pl->n_snv), cudaMemcpyHostToDevice, st));
    }
    return set_genotypes_impl(pl, pl->stage_gt, pitch_bytes, d_keep, st, true);
}

int scs_placement_set_clades_host(scs_placement* h, const uint32_t* bits, int64_t pitch_words, void* stream) {
    Placement* pl = reinterpret_cast<Placement*>(h);
    if (!pl || !bits) return set_error(SCS_ERR_ARG, "null argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t bytes = size_t(pl->n_groups) * pl->n_rows * pitch_words * 4;
    if (bytes > pl->stage_mask_bytes) {
        cudaFree(pl->stage_mask);
        pl->stage_mask = nullptr;
        pl->stage_mask_bytes = 0;
        if (cudaMalloc(&pl->stage_mask, bytes) != cudaSuccess) return set_error(SCS_ERR_OOM, "device staging allocation of %zu bytes failed", bytes);
        pl->stage_mask_bytes = bytes;
    }
    SCS_CUDA(cudaMemcpyAsync(pl->stage_mask, bits, bytes, cudaMemcpyHostToDevice, st));
    return scs_placement_set_clades(h, pl->stage_mask, pitch_words, stream);
}

static void make_constants(int n_cells, double FP, double FN, GroupConst* gc, double2* gd) {
    const double LOG2E = 1.4426950408889634074;
    const double a1 = (log1p(-FN) - log(FP)) * LOG2E;   // log2((1-FN)/FP)
    const double a0 = (log(FN) - log1p(-FP)) * LOG2E;   // log2(FN/(1-FP))
    int cbits = 1;
    while ((1 << cbits) <= n_cells) ++cbits;             // counts and their differences fit cbits (+sign)
    const int hbits = std::max(4, 24 - cbits);
    gc->a1h = round_to_bits(a1, hbits);
    gc->a0h = round_to_bits(a0, hbits);
    gc->a1l = float(a1 - double(gc->a1h));
    gc->a0l = float(a0 - double(gc->a0h));
    gc->a1f = float(a1);
    gc->a0f = float(a0);
    gc->pad0 = gc->pad1 = 0.f;
    gd->x = a1;
    gd->y = a0;
}

// How the softmax column sums (pass 2) are formed.  "recompute": a second GEMM pass, nothing but
// 16 bytes per SNV kept between the passes.  "cache": pass 1 also writes the packed integer counts
// (4 bytes per SNV x clade row) and pass 2 is an HBM-bound sweep over them -- cheaper than the tensor
// work once K exceeds ~2,000 cells (measured at 10k cells: ~15 ms instead of 140 ms), at the price of
// M_pad * N_pad * 4 bytes of HBM (80 GB for 1M SNVs x 20k clade rows; callers with less memory to
// spare walk the SNV rows in chunks, engine.StreamingPlacement).  SCS_PLACE_PASS2 = recompute | cache
// overrides; the default takes the cache when the plan is a single dataset with K >= 2560 cells and
// the buffer fits the free device memory with 4 GB to spare, else recomputes.
static void choose_pass2_mode(Placement* pl) {
    pl->pass2_mode = 1;
    const char* e = getenv("SCS_PLACE_PASS2");
    if (e && !strcmp(e, "recompute")) return;
    const bool forced = e && !strcmp(e, "cache");
    if (pl->n_groups != 1) return;
    if (!forced && pl->K_pad < 2560) return;
    const size_t bytes = size_t(pl->M_pad) * size_t(pl->N_pad) * sizeof(uint32_t);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || free_b < bytes + (size_t(4) << 30)) return;
    if (cudaMalloc(&pl->cache, bytes) != cudaSuccess || cudaMalloc(&pl->rowref_f, size_t(pl->M_pad) * sizeof(float4)) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(pl->cache);
        cudaFree(pl->rowref_f);
        pl->cache = nullptr;
        pl->rowref_f = nullptr;
        return;
    }
    pl->prm.cache = pl->cache;
    pl->pass2_mode = 2;
}

// get_gt_tree's filters (run_PT_test.py:350-358) for a plan whose clades are set, asynchronous on `stream`, counters on
// the device.  Interval path (variants 3 / 4): ONE kernel reads the caller's matrix once and produces the row flags, the
// missing counts and the rows in depth-first cell order (scs_permute_reduce_kernel).  GEMM path / variant 1: the two
// reduction kernels of scs_gt_reduce, then the operand planes.
int scs_placement_prepare(scs_placement* h, const uint8_t* d_gt, int64_t pitch_bytes, const uint8_t* col_active, int32_t filter_rows,
                          int32_t accumulate, void* stream) {
    Placement* pl = reinterpret_cast<Placement*>(h);
    if (!pl || !d_gt || !col_active) return set_error(SCS_ERR_ARG, "null argument");
    if (pl->algo == 0) return set_error(SCS_ERR_ARG, "scs_placement_prepare before scs_placement_set_clades");
    if (pitch_bytes < (pl->n_cells + 3) / 4) return set_error(SCS_ERR_ARG, "genotype row pitch %lld is smaller than ceil(n_cells/4)", (long long)pitch_bytes);
    if ((pitch_bytes & 3) || (reinterpret_cast<uintptr_t>(d_gt) & 3)) return set_error(SCS_ERR_ARG, "packed genotype rows must be 4-byte aligned (pitch %lld)", (long long)pitch_bytes);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int n_cells = pl->n_cells, rw = (n_cells + 15) / 16;
    const bool fused = pl->algo == 2 && pl->iv_variant >= 3;
    const int mode = fused ? 1 : 2;
    if (accumulate && pl->prep_mode != mode) return set_error(SCS_ERR_ARG, "scs_placement_prepare: nothing to accumulate onto");
    const size_t off_mask = 16, off_miss = off_mask + size_t(rw + 3) / 4 * 16;
    const size_t need = fused ? off_miss + size_t(16) * rw * sizeof(uint32_t) : gt_reduce_scratch_bytes(n_cells);
    if (need > pl->d_prep_bytes) {
        if (accumulate) return set_error(SCS_ERR_ARG, "scs_placement_prepare: counter buffer changed size while accumulating");
        cudaFree(pl->d_prep);
        pl->d_prep = nullptr;
        pl->d_prep_bytes = 0;
        if (cudaMalloc(&pl->d_prep, need) != cudaSuccess) return set_error(SCS_ERR_OOM, "device allocation of %zu bytes failed", need);
        pl->d_prep_bytes = need;
    }
    pl->prep_mode = mode;
    cudaEventRecord(pl->ev_prep[0], st);
    pl->prep_timed = true;
    if (!fused) {
        int rc = gt_reduce_launch(pl->ctx, d_gt, pitch_bytes, pl->n_snv, n_cells, col_active, filter_rows, pl->row_keep, pl->d_prep, accumulate != 0, st);
        if (rc != SCS_OK) return rc;
        pl->launches += rw <= 32 ? 1 : 2;
        rc = set_genotypes_impl(pl, d_gt, pitch_bytes, nullptr, st, true);
        pl->have_row_keep = true;
        cudaEventRecord(pl->ev_prep[1], st);
        return rc;
    }
    if (!accumulate) {
        std::vector<uint32_t> amask(size_t(rw), 0u);      // active cells (source column order): low bit of every 2-bit code
        for (int c = 0; c < n_cells; ++c)
            if (col_active[c]) amask[size_t(c >> 4)] |= 1u << (2 * (c & 15));
        SCS_CUDA(cudaMemsetAsync(pl->d_prep, 0, need, st));
        {
            const int urc = upload_small(pl->ctx, pl->d_prep + off_mask, amask.data(), size_t(rw) * 4, st);      // (a kernel over pinned staging, not the copy engine, which may be busy with the next chunk of the matrix)
            if (urc != SCS_OK) return urc;
            pl->launches++;
        }
    }
    const size_t bytes = size_t(pl->n_snv) * rw * sizeof(uint32_t);
    if (bytes > pl->gt_perm_bytes) {
        cudaFree(pl->gt_perm);
        pl->gt_perm = nullptr;
        pl->gt_perm_bytes = 0;
        if (cudaMalloc(&pl->gt_perm, bytes) != cudaSuccess) return set_error(SCS_ERR_OOM, "device allocation of %zu bytes failed", bytes);
        pl->gt_perm_bytes = bytes;
    }
    const int pt = (rw + 31) / 32 * 32;
    const char* pe = getenv("SCS_PREP_ROWS");          // 1: the one-row-per-step kernel (kept for comparison); default: four rows per step
    if (pe && atoi(pe) == 1) {
        const unsigned pgrid = unsigned(std::min<int64_t>(pl->n_snv, int64_t(pl->ctx->sm_count) * std::max(1, 2048 / pt)));
        scs_permute_reduce_kernel<<<pgrid, pt, 2 * rw * sizeof(uint32_t), st>>>(
            d_gt, pitch_bytes, pl->n_snv, n_cells, rw, pl->d_perm, reinterpret_cast<const uint32_t*>(pl->d_prep + off_mask), filter_rows,
            pl->gt_perm, pl->row_keep, reinterpret_cast<unsigned long long*>(pl->d_prep), reinterpret_cast<unsigned int*>(pl->d_prep + off_miss));
    } else {
        const int64_t groups = (pl->n_snv + 3) / 4;
        // narrow rows: several sub-blocks (one per group of four rows) share a 256-thread block, each seeing >= 32 groups
        const int nsub = pt >= 256 ? 1 : std::min(PERM4_MAX_SUB, 256 / pt);
        int64_t sub_blocks = std::min<int64_t>(groups, int64_t(pl->ctx->sm_count) * std::max(1, 2048 / pt));
        if (nsub > 1) sub_blocks = std::min<int64_t>(sub_blocks, std::max<int64_t>(1, (groups + 31) / 32));
        const unsigned pgrid = unsigned((sub_blocks + nsub - 1) / nsub);
        scs_permute_reduce4_kernel<<<pgrid, pt * nsub, size_t(nsub) * 2 * (4 * rw + 4) * sizeof(uint32_t), st>>>(
            d_gt, pitch_bytes, pl->n_snv, n_cells, rw, pl->d_perm, reinterpret_cast<const uint32_t*>(pl->d_prep + off_mask), filter_rows,
            pl->gt_perm, pl->row_keep, reinterpret_cast<unsigned long long*>(pl->d_prep), reinterpret_cast<unsigned int*>(pl->d_prep + off_miss), pt);
    }
    SCS_CUDA(cudaGetLastError());
    cudaEventRecord(pl->ev_prep[1], st);
    pl->launches++;
    pl->operands_valid = false;
    pl->gt_src = d_gt;             // bound: a later switch to general clade masks builds the GEMM operands from it
    pl->gt_pitch = pitch_bytes;
    pl->gt_perm_valid = true;
    pl->have_row_keep = true;
    return SCS_OK;
}

// Waits for the prepare call(s) on `stream` and returns the kept-SNV count and, per cell (column), the fraction of kept
// SNVs where the cell is missing.
int scs_placement_prepare_result(scs_placement* h, int64_t* n_kept, double* missing_frac, void* stream) {
    Placement* pl = reinterpret_cast<Placement*>(h);
    if (!pl || !n_kept || !missing_frac) return set_error(SCS_ERR_ARG, "null argument");
    if (pl->prep_mode == 0) return set_error(SCS_ERR_ARG, "scs_placement_prepare_result before scs_placement_prepare");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int n_cells = pl->n_cells, rw = (n_cells + 15) / 16;
    if (pl->prep_mode == 2) return gt_reduce_fetch(pl->d_prep, n_cells, n_kept, missing_frac, st);
    const size_t off_miss = 16 + size_t(rw + 3) / 4 * 16;
    std::vector<unsigned int> miss(size_t(16) * rw);
    unsigned long long kept = 0;
    SCS_CUDA(cudaMemcpyAsync(&kept, pl->d_prep, sizeof(kept), cudaMemcpyDeviceToHost, st));
    SCS_CUDA(cudaMemcpyAsync(miss.data(), pl->d_prep + off_miss, miss.size() * 4, cudaMemcpyDeviceToHost, st));
    SCS_CUDA(cudaStreamSynchronize(st));
    *n_kept = int64_t(kept);
    for (int k = 0; k < n_cells; ++k) missing_frac[pl->h_perm[size_t(k)]] = kept > 0 ? double(miss[size_t(k)]) / double(kept) : 0.0;
    return SCS_OK;
}

// FP / FN: one value per group (per dataset).
int scs_placement_run_batch(scs_placement* h, const double* FP, const double* FN, double* d_soft_weights, void* stream) {
    Placement* pl = reinterpret_cast<Placement*>(h);
    if (!pl || !FP || !FN) return set_error(SCS_ERR_ARG, "null argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    std::vector<GroupConst> gc(pl->n_groups);
    std::vector<double2> gd(pl->n_groups);
    for (int g = 0; g < pl->n_groups; ++g) {
        if (!(FP[g] > 0.0 && FP[g] < 1.0 && FN[g] > 0.0 && FN[g] < 1.0))
            return set_error(SCS_ERR_ARG, "error rates must lie in (0,1): group %d FP=%g FN=%g", g, FP[g], FN[g]);
        make_constants(pl->n_cells, FP[g], FN[g], &gc[g], &gd[g]);
    }
    {
        // (a kernel reading page-locked staging, not the copy engine: see scs::upload_small)
        int rc = upload_small(pl->ctx, pl->d_gconst, gc.data(), gc.size() * sizeof(GroupConst), st);
        if (rc == SCS_OK) rc = upload_small(pl->ctx, pl->d_gconst_d, gd.data(), gd.size() * sizeof(double2), st);
        if (rc != SCS_OK) return rc;
        pl->launches += 2;          // (the two upload kernels)
    }
    if (pl->algo == 0) return set_error(SCS_ERR_ARG, "scs_placement_run before scs_placement_set_clades");
    if (pl->algo == 3) {
        if (!pl->gt_perm_valid || !pl->gt_src) return set_error(SCS_ERR_ARG, "scs_placement_run before scs_placement_set_genotypes");
        for (int g = 0; g < pl->n_groups; ++g)      // (the kernel parks its unused row slots at counts (0, 65535): needs log(FN / (1 - FP)) < 0)
            if (!(FN[g] + FP[g] < 1.0))
                return set_error(SCS_ERR_ARG, "group %d: FN + FP >= 1 (FP=%g FN=%g) is not supported on the small-tree path; SCS_PLACE_ALGO=gemm", g, FP[g], FN[g]);
        Iv5Params q;
        q.gt = pl->gt_src;
        q.pitch = pl->gt_pitch;
        q.n_cols = pl->n_cells;
        q.perm = pl->iv5_perm;
        q.row_keep = pl->have_row_keep ? pl->row_keep : nullptr;
        q.tasks = pl->iv5_tasks;
        q.n_tasks = pl->iv5_ntasks;
        q.n_rows = pl->n_rows;
        q.nl_lohi = pl->iv5_lohi;
        q.meta = pl->iv5_meta;
        q.gconst = pl->d_gconst;
        q.part = pl->iv5_part;
        q.stride = pl->iv5_stride;
        double* y = d_soft_weights ? d_soft_weights : pl->y;
        const size_t smem = iv5_smem_bytes(pl->iv5_L, pl->iv5_stride);
        cudaEventRecord(pl->ev[0], st);
        cudaEventRecord(pl->ev[1], st);
        cudaEventRecord(pl->ev[2], st);
        iv5_launch(pl->iv5_L, (std::max(pl->n_cells - 2, 1) + pl->iv5_L - 1) / pl->iv5_L, q, pl->iv5_grid, smem, st);
        cudaEventRecord(pl->ev[3], st);          // ms[2] = the placement kernel
        scs_iv5_reduce_kernel<<<pl->n_groups, 128, 0, st>>>(pl->iv5_part, pl->iv5_stride, pl->iv5_task0, pl->iv5_row_slot, pl->iv5_meta, pl->n_rows, y);
        cudaEventRecord(pl->ev[4], st);
        pl->timed = true;
        pl->launches += 2;
        SCS_CUDA(cudaGetLastError());
        return SCS_OK;
    }
    if (pl->algo == 2) {
        if (!pl->gt_src) return set_error(SCS_ERR_ARG, "scs_placement_run before scs_placement_set_genotypes");
        IntervalParams ip;
        ip.gt = pl->gt_src;
        ip.pitch = pl->gt_pitch;
        ip.n_snv = pl->n_snv;
        ip.n_cells = pl->n_cells;
        ip.n_nodes = pl->iv_nodes;
        ip.n_slots = pl->iv_slots;
        ip.row_words = (pl->n_cells + 15) / 16;
        ip.perm = pl->d_perm;
        ip.lohi = pl->d_lohi;
        ip.row_keep = pl->have_row_keep ? pl->row_keep : nullptr;
        ip.gconst = pl->d_gconst;
        ip.part = pl->ipart;
        ip.rows_per_run = pl->iv_rows_per_run;
        ip.num_runs = pl->iv_runs;
        ip.perm_smem = pl->iv_perm_smem ? 1 : 0;
        double* y = d_soft_weights ? d_soft_weights : pl->y;
        cudaEventRecord(pl->ev[0], st);
        if (pl->iv_variant >= 3) {
            const int rw = ip.row_words;
            if (!pl->gt_perm_valid) {
                // the rows in depth-first cell order (one streaming pass; the walk below then works on whole words)
                const size_t bytes = size_t(pl->n_snv) * rw * sizeof(uint32_t);
                if (bytes > pl->gt_perm_bytes) {
                    cudaFree(pl->gt_perm);
                    pl->gt_perm = nullptr;
                    pl->gt_perm_bytes = 0;
                    if (cudaMalloc(&pl->gt_perm, bytes) != cudaSuccess) return set_error(SCS_ERR_OOM, "device allocation of %zu bytes failed", bytes);
                    pl->gt_perm_bytes = bytes;
                }
                const int pt = (rw + 31) / 32 * 32;
                const unsigned pgrid = unsigned(std::min<int64_t>(pl->n_snv, int64_t(pl->ctx->sm_count) * std::max(1, 2048 / pt)));
                scs_permute_cells_kernel<<<pgrid, pt, 2 * rw * sizeof(uint32_t), st>>>(pl->gt_src, pl->gt_pitch, pl->n_snv, pl->n_cells, rw, pl->d_perm,
                                                                                   pl->gt_perm);
                pl->gt_perm_valid = true;
                pl->launches++;
            }
            cudaEventRecord(pl->ev[1], st);      // ms[0] = the permutation pass (0 when the permuted rows were still valid)
            cudaEventRecord(pl->ev[2], st);
            Interval3Params q;
            ip.gt = reinterpret_cast<const uint8_t*>(pl->gt_perm);
            ip.pitch = int64_t(rw) * 4;
            q.b = ip;
            q.leafmask = pl->d_leafmask;
            q.off4 = pl->d_off4;
            q.n_leaf_pos = pl->iv_leaf_pos;
            q.part_stride = pl->iv_stride;
            q.leaf_off = pl->iv_leaf_off;
            if (pl->iv_variant == 4) {
                Iv4Params q4;
                q4.gt = pl->gt_perm;
                q4.n_snv = pl->n_snv;
                q4.rw = rw;
                q4.n_pairs = pl->iv_slots / 2;
                q4.pad_pair = (pl->iv_nodes & 1) ? pl->iv_slots / 2 - 1 : -1;
                q4.off4 = pl->d_off4;
                q4.leafmask = pl->d_leafmask;
                q4.n_leaf_pos = pl->iv_leaf_pos;
                q4.row_keep = ip.row_keep;
                q4.gconst = pl->d_gconst;
                q4.part = pl->ipart;
                q4.part_stride = pl->iv_stride;
                q4.leaf_off = pl->iv_leaf_off;
                q4.rows_per_run = pl->iv_rows_per_run;
                q4.num_runs = pl->iv_runs;
                q4.leafw = nullptr;
                // leaf split: the leaf accumulators leave the placement kernel (where they live in shared memory: 8 LDS / STS
                // of 16 bytes per thread and row on top of the selects and adds) for scs_iv4_leaf_kernel (registers); measured
                // at 10,000 cells x 1M SNVs: 20.9 -> 20.3 ms for the pair of kernels.  Default from 2,048 cells on (below that
                // the second launch costs what the split saves); SCS_IV4_LEAF_SPLIT=0 / 1 overrides.
                bool split = pl->n_cells >= 2048;
                if (const char* ev = getenv("SCS_IV4_LEAF_SPLIT")) split = atoi(ev) != 0;
                {
                    if (split) {
                        const size_t lbytes = size_t(pl->n_snv) * sizeof(float4);
                        if (lbytes > pl->iv4_leafw_bytes) {
                            cudaFree(pl->iv4_leafw);
                            pl->iv4_leafw = nullptr;
                            pl->iv4_leafw_bytes = 0;
                            if (cudaMalloc(&pl->iv4_leafw, lbytes) != cudaSuccess) return set_error(SCS_ERR_OOM, "device allocation of %zu bytes failed", lbytes);
                            pl->iv4_leafw_bytes = lbytes;
                        }
                        SCS_CUDA(cudaMemsetAsync(pl->iv4_leafw, 0, lbytes, st));      // (rows that are filtered out keep zero weights)
                        q4.leafw = pl->iv4_leafw;
                        pl->launches++;
                    }
                }
                iv4_launch(pl->iv4_np, q4, pl->iv_grid, pl->iv_threads, pl->iv_smem, st);
            } else if (pl->iv_threads <= 640) scs_place_interval3_kernel<640, 2><<<pl->iv_grid, pl->iv_threads, pl->iv_smem, st>>>(q);
            else scs_place_interval3_kernel<1024, 1><<<pl->iv_grid, pl->iv_threads, pl->iv_smem, st>>>(q);
        } else {
            cudaEventRecord(pl->ev[1], st);
            cudaEventRecord(pl->ev[2], st);
            scs_place_interval_kernel<false><<<pl->iv_grid, pl->iv_threads, pl->iv_smem, st>>>(ip);
        }
        cudaEventRecord(pl->ev[3], st);          // ms[2] = the placement kernel
        scs_reduce_interval_kernel<<<unsigned((pl->iv_entries + 31) / 32), 32 * IV_REDUCE_SLICES, 0, st>>>(pl->ipart, pl->iv_runs, pl->iv_entries,
                                                                                                           pl->iv_stride, pl->d_slot_row, y);
        cudaEventRecord(pl->ev[4], st);
        pl->timed = true;
        pl->launches += 2;
        SCS_CUDA(cudaGetLastError());
        return SCS_OK;
    }
    if (!pl->operands_valid) {
        // the masks changed from tree-shaped to general after the genotypes were set: build the operands now
        if (!pl->gt_src) return set_error(SCS_ERR_ARG, "scs_placement_run before scs_placement_set_genotypes");
        int rc = expand_operands(pl, pl->gt_src, pl->gt_pitch, st);
        if (rc != SCS_OK) return rc;
    }
    if (pl->pass2_mode == 0) choose_pass2_mode(pl);
    const PlaceParams& p = pl->prm;
    const bool cached = pl->pass2_mode == 2;
    cudaEventRecord(pl->ev[0], st);
    launch_place<PASS_STATS>(pl->mode, pl->grid, st, pl->tm_x1, pl->tm_x0, pl->tm_s, p);
    cudaEventRecord(pl->ev[1], st);
    scs_merge_stats_kernel<<<unsigned((pl->M_pad + 255) / 256), 256, 0, st>>>(pl->stats, pl->num_parts, pl->M_pad, pl->rowmap,
                                                                              pl->have_row_keep ? pl->row_keep : nullptr, pl->d_gconst_d,
                                                                              pl->rowref, cached ? pl->rowref_f : nullptr);
    cudaEventRecord(pl->ev[2], st);
    if (cached) {
        dim3 grid(unsigned((pl->n_rows + BLOCK_N - 1) / BLOCK_N), unsigned(p.num_runs));
        scs_colsum_cached_kernel<<<grid, BLOCK_N, 0, st>>>(pl->cache, pl->rowref_f, pl->d_gconst, p.num_row_blocks, pl->N_pad, pl->n_rows,
                                                           p.run_len, pl->colpart);
    } else {
        launch_place<PASS_COLSUM>(pl->mode, pl->grid, st, pl->tm_x1, pl->tm_x0, pl->tm_s, p);
    }
    cudaEventRecord(pl->ev[3], st);
    double* y = d_soft_weights ? d_soft_weights : pl->y;
    const int64_t total_rows = int64_t(pl->n_groups) * pl->n_rows;
    scs_reduce_colpart_kernel<<<unsigned((total_rows + 255) / 256), 256, 0, st>>>(pl->colpart, pl->d_group_runs, total_rows, pl->n_rows,
                                                                                  pl->N_pad1, pl->N_pad, y);
    cudaEventRecord(pl->ev[4], st);
    pl->timed = true;
    pl->launches += 4;
    SCS_CUDA(cudaGetLastError());
    return SCS_OK;
}

int scs_placement_run(scs_placement* h, double FP, double FN, double* d_soft_weights, void* stream) {
    Placement* pl = reinterpret_cast<Placement*>(h);
    if (!pl) return set_error(SCS_ERR_ARG, "null plan");
    std::vector<double> fp(pl->n_groups, FP), fn(pl->n_groups, FN);
    return scs_placement_run_batch(h, fp.data(), fn.data(), d_soft_weights, stream);
}

int scs_placement_run_host(scs_placement* h, double FP, double FN, double* soft_weights, void* stream) {
    Placement* pl = reinterpret_cast<Placement*>(h);
    if (!pl || !soft_weights) return set_error(SCS_ERR_ARG, "null argument");
    int rc = scs_placement_run(h, FP, FN, nullptr, stream);
    if (rc != SCS_OK) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    SCS_CUDA(cudaMemcpyAsync(soft_weights, pl->y, size_t(pl->n_groups) * pl->n_rows * sizeof(double), cudaMemcpyDeviceToHost, st));
    SCS_CUDA(cudaStreamSynchronize(st));
    return SCS_OK;
}

// Debug / parity aid: raw integer counts C1, C0 for every (SNV, clade row).
// Only meant for small problems (the buffer is M_pad x N_pad x 2 ints).
int scs_placement_debug_counts(scs_placement* h, int32_t* counts_host, void* stream) {
    Placement* pl = reinterpret_cast<Placement*>(h);
    if (!pl || !counts_host) return set_error(SCS_ERR_ARG, "null argument");
    if (pl->n_groups != 1) return set_error(SCS_ERR_ARG, "the debug count dump is for single-group plans");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    std::vector<GroupConst> gc(1);
    std::vector<double2> gd(1);
    make_constants(pl->n_cells, 0.01, 0.1, &gc[0], &gd[0]);     // the dump pass reads no constants; keep the table defined
    SCS_CUDA(cudaMemcpyAsync(pl->d_gconst, gc.data(), sizeof(GroupConst), cudaMemcpyHostToDevice, st));
    const size_t n = size_t(pl->M_pad) * pl->N_pad * 2;
    if (n > (size_t(1) << 28)) return set_error(SCS_ERR_ARG, "debug count dump is limited to small problems");
    if (!pl->dump && cudaMalloc(&pl->dump, n * sizeof(int)) != cudaSuccess) return set_error(SCS_ERR_OOM, "debug dump allocation failed");
    if (!pl->operands_valid) {       // the plan runs the interval path: the GEMM operands were never built
        if (!pl->gt_src) return set_error(SCS_ERR_ARG, "set the genotypes before asking for the count dump");
        int rc = expand_operands(pl, pl->gt_src, pl->gt_pitch, st);
        if (rc != SCS_OK) return rc;
    }
    if (pl->algo == 2) {             // ... and neither was the clade operand
        if (pl->h_bits.empty()) return set_error(SCS_ERR_ARG, "no clade masks kept for the count dump");
        uint32_t* d_bits = nullptr;
        const size_t bbytes = pl->h_bits.size() * sizeof(uint32_t);
        if (cudaMalloc(&d_bits, bbytes) != cudaSuccess) return set_error(SCS_ERR_OOM, "debug dump allocation failed");
        cudaMemcpyAsync(d_bits, pl->h_bits.data(), bbytes, cudaMemcpyHostToDevice, st);
        dim3 block(128), grid((pl->K_pad / 4 + 127) / 128, unsigned(std::min<int64_t>(pl->n_rows, 32768)));
        scs_expand_clades_kernel<<<grid, block, 0, st>>>(d_bits, pl->h_bits_pitch, pl->n_rows, pl->n_rows, pl->N_pad1, pl->n_cells, pl->s, pl->K_pad,
                                                          pl->mode == MODE_RADIX ? uint32_t(E5M2_ONE) : 1u);
        pl->launches++;
        cudaStreamSynchronize(st);
        cudaFree(d_bits);
        SCS_CUDA(cudaGetLastError());
    }
    PlaceParams p = pl->prm;
    p.dump = pl->dump;
    launch_place<PASS_DUMP>(pl->mode, pl->grid, st, pl->tm_x1, pl->tm_x0, pl->tm_s, p);
    pl->launches++;
    SCS_CUDA(cudaGetLastError());
    std::vector<int> tmp(n);
    SCS_CUDA(cudaMemcpyAsync(tmp.data(), pl->dump, n * sizeof(int), cudaMemcpyDeviceToHost, st));
    SCS_CUDA(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < pl->n_snv; ++i)
        memcpy(counts_host + size_t(i) * pl->n_rows * 2, tmp.data() + size_t(i) * pl->N_pad * 2, size_t(pl->n_rows) * 2 * sizeof(int));
    return SCS_OK;
}

// Analysis aid: the pass-1 partials ([parts][M_pad] {ref counts, sum}) and the merged per-SNV
// records ([M_pad] {ref counts, log2 sum}) of the last run, copied to the host.
int scs_placement_debug_stats(scs_placement* h, void* stats_host, void* rowref_host) {
    Placement* pl = reinterpret_cast<Placement*>(h);
    if (!pl) return set_error(SCS_ERR_ARG, "null argument");
    SCS_CUDA(cudaDeviceSynchronize());
    if (stats_host && !pl->stats) return set_error(SCS_ERR_ARG, "no pass-1 partials: the plan has not run the GEMM path");
    if (stats_host) SCS_CUDA(cudaMemcpy(stats_host, pl->stats, size_t(pl->num_parts) * pl->M_pad * sizeof(uint2), cudaMemcpyDeviceToHost));
    if (rowref_host) SCS_CUDA(cudaMemcpy(rowref_host, pl->rowref, size_t(pl->M_pad) * sizeof(uint2), cudaMemcpyDeviceToHost));
    return SCS_OK;
}

int scs_placement_info(scs_placement* h, int64_t* info, int32_t n) {
    Placement* pl = reinterpret_cast<Placement*>(h);
    if (!pl || !info) return set_error(SCS_ERR_ARG, "null argument");
    const int64_t v[21] = {pl->M_pad, pl->N_pad, pl->K_pad, pl->prm.num_tasks, pl->prm.run_len, pl->prm.sb_width, pl->grid, pl->launches, pl->prm.num_runs,
                           (pl->mode == MODE_RADIX ? ModeCfg<MODE_RADIX>::SMEM_TOTAL : ModeCfg<MODE_PLANES>::SMEM_TOTAL),
                           pl->mode, pl->prm.nchunk, 0 /* (was: 2-CTA kernel, removed) */, pl->n_groups, pl->pass2_mode,
                           pl->algo, pl->algo == 3 ? pl->iv5_grid : pl->iv_grid, pl->algo == 3 ? IV5_THREADS : pl->iv_threads,
                           pl->algo == 3 ? int64_t(iv5_smem_bytes(pl->iv5_L, pl->iv5_stride)) : int64_t(pl->iv_smem), pl->algo == 3 ? pl->iv5_ntasks : pl->iv_runs,
                           pl->algo == 3 ? 5 : pl->iv_variant};
    for (int i = 0; i < n && i < 21; ++i) info[i] = v[i];
    return SCS_OK;
}

int scs_placement_last_times(scs_placement* h, float* ms) {
    Placement* pl = reinterpret_cast<Placement*>(h);
    if (!pl || !ms) return set_error(SCS_ERR_ARG, "null argument");
    if (!pl->timed) return set_error(SCS_ERR_ARG, "scs_placement_run has not been called yet");
    SCS_CUDA(cudaEventSynchronize(pl->ev[4]));
    for (int i = 0; i < 4; ++i) SCS_CUDA(cudaEventElapsedTime(&ms[i], pl->ev[i], pl->ev[i + 1]));
    return SCS_OK;
}

int scs_placement_prepare_time(scs_placement* h, float* ms) {
    Placement* pl = reinterpret_cast<Placement*>(h);
    if (!pl || !ms) return set_error(SCS_ERR_ARG, "null argument");
    if (!pl->prep_timed) return set_error(SCS_ERR_ARG, "scs_placement_prepare has not been called yet");
    SCS_CUDA(cudaEventSynchronize(pl->ev_prep[1]));
    SCS_CUDA(cudaEventElapsedTime(ms, pl->ev_prep[0], pl->ev_prep[1]));
    return SCS_OK;
}

const void* scs_placement_device_result(scs_placement* h) {
    Placement* pl = reinterpret_cast<Placement*>(h);
    return pl ? pl->y : nullptr;
}

// map_mutations_gt with one false-negative rate per cell (the pd.Series branch, run_PT_test.py:413-418): scs_place_percell.cuh.
int scs_place_percell(scs_ctx* ctx, const uint8_t* d_gt, int64_t pitch_bytes, int64_t n_snv, int32_t n_cells, const uint8_t* d_row_keep,
                      const uint32_t* clade_bits, int64_t pitch_words, int32_t n_rows, double FP, const double* FN_cells,
                      double* d_soft_weights, float* kernel_ms, void* stream) {
    if (!ctx || !clade_bits || !FN_cells || !d_soft_weights || (n_snv > 0 && !d_gt)) return set_error(SCS_ERR_ARG, "null argument");
    if (n_snv < 0 || n_cells <= 0 || n_cells > 65535 || n_rows <= 0) return set_error(SCS_ERR_ARG, "bad problem size");
    if (pitch_words < (n_cells + 31) / 32) return set_error(SCS_ERR_ARG, "clade mask pitch %lld is smaller than ceil(n_cells/32)", (long long)pitch_words);
    if (pitch_bytes < (n_cells + 3) / 4) return set_error(SCS_ERR_ARG, "genotype row pitch %lld is smaller than ceil(n_cells/4)", (long long)pitch_bytes);
    if ((pitch_bytes & 3) || (reinterpret_cast<uintptr_t>(d_gt) & 3)) return set_error(SCS_ERR_ARG, "packed genotype rows must be 4-byte aligned (pitch %lld)", (long long)pitch_bytes);
    if (!(FP > 0.0 && FP < 1.0)) return set_error(SCS_ERR_ARG, "FP must lie in (0, 1)");
    SCS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    std::vector<uint16_t> perm;
    std::vector<uint32_t> lohi;
    if (!clade_intervals(clade_bits, pitch_words, n_rows, n_cells, perm, lohi))
        return set_error(SCS_ERR_ARG, "per-cell rates need tree-shaped clade masks (the rows are not a laminar family)");
    // cells that belong to no clade (the outgroup column) carry no weight; every other cell needs a rate in (0, 1)
    const int words = (n_cells + 31) / 32;
    std::vector<uint32_t> covered(size_t(words), 0u);
    for (int r = 0; r < n_rows; ++r)
        for (int w = 0; w < words; ++w) covered[size_t(w)] |= clade_bits[size_t(r) * pitch_words + w];
    const double l2e = 1.4426950408889634074;
    std::vector<double2> wv(static_cast<size_t>(n_cells));
    for (int k = 0; k < n_cells; ++k) {
        const int c = perm[size_t(k)];
        if (!((covered[size_t(c >> 5)] >> (c & 31)) & 1u)) {
            wv[size_t(k)] = make_double2(0.0, 0.0);
            continue;
        }
        const double fn = FN_cells[c];
        if (!(fn > 0.0 && fn < 1.0)) return set_error(SCS_ERR_ARG, "FN of cell column %d must lie in (0, 1)", c);
        wv[size_t(k)] = make_double2((std::log(1.0 - fn) - std::log(FP)) * l2e, (std::log(fn) - std::log(1.0 - FP)) * l2e);
    }
    // clade rows sorted by interval (neighbouring lanes then read neighbouring prefix sums); slot -> caller's row
    std::vector<int> slot_row(static_cast<size_t>(n_rows));
    std::iota(slot_row.begin(), slot_row.end(), 0);
    std::stable_sort(slot_row.begin(), slot_row.end(), [&](int a, int b) {
        const uint32_t la = lohi[size_t(a)], lb = lohi[size_t(b)];
        return (la & 0xffffu) != (lb & 0xffffu) ? (la & 0xffffu) < (lb & 0xffffu) : (la >> 16) < (lb >> 16);
    });
    std::vector<uint32_t> lohi_sorted(static_cast<size_t>(n_rows));
    for (int b = 0; b < n_rows; ++b) lohi_sorted[size_t(b)] = lohi[size_t(slot_row[size_t(b)])];
    int max_smem = 0;
    SCS_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    // which kernel: trees of 1,024 .. 12,287 cells keep every thread's clade rows in registers (scs_place_percell2_kernel<R>,
    // 512 threads, R = rows per thread); smaller and bigger trees the general kernel.  SCS_PERCELL_KERNEL = 1 forces the latter.
    int R2 = 0;
    {
        const char* e = getenv("SCS_PERCELL_KERNEL");
        const int need = (n_rows + PC2_THREADS - 1) / PC2_THREADS;
        if (!(e && atoi(e) == 1) && n_rows >= 2048 && need <= PC2_MAX_R) {
            R2 = (need + 7) / 8 * 8;
            if (percell2_smem_bytes(n_cells, R2) > size_t(max_smem)) R2 = 0;
        }
    }
    // general kernel: a thread per ~8 clade rows / cells, 64..1024 threads; accumulators and cell order in shared memory when they fit
    int T = std::max((n_rows + 7) / 8, (n_cells + 7) / 8);
    T = std::min(1024, std::max(64, (T + 31) / 32 * 32));
    bool acc_smem = true, perm_smem = true;
    size_t smem = 0;
    int per_sm = 0;
    void (*kern2)(PercellParams) = nullptr;
    if (R2 > 0) {
        T = PC2_THREADS;
        smem = percell2_smem_bytes(n_cells, R2);
        switch (R2) {
            case 8: kern2 = scs_place_percell2_kernel<8>; break;
            case 16: kern2 = scs_place_percell2_kernel<16>; break;
            case 24: kern2 = scs_place_percell2_kernel<24>; break;
            case 32: kern2 = scs_place_percell2_kernel<32>; break;
            case 40: kern2 = scs_place_percell2_kernel<40>; break;
            default: kern2 = scs_place_percell2_kernel<48>; break;
        }
        SCS_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        SCS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern2, T, smem));
    } else {
        if (percell_smem_bytes(n_cells, n_rows, acc_smem, perm_smem) > size_t(max_smem)) perm_smem = false;
        if (percell_smem_bytes(n_cells, n_rows, acc_smem, perm_smem) > size_t(max_smem)) acc_smem = false;
        smem = percell_smem_bytes(n_cells, n_rows, acc_smem, perm_smem);
        if (smem > size_t(max_smem)) return set_error(SCS_ERR_ARG, "per-cell placement: %d cells do not fit shared memory", n_cells);
        SCS_CUDA(cudaFuncSetAttribute(scs_place_percell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        SCS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scs_place_percell_kernel, T, smem));
    }
    per_sm = std::max(1, per_sm);
    const int grid = int(std::max<int64_t>(1, std::min<int64_t>(n_snv, int64_t(ctx->sm_count) * per_sm)));
    // workspace: tables + per-CTA partial sums, one allocation
    const size_t off_lohi = (size_t(n_cells) * sizeof(uint16_t) + 15) & ~size_t(15);
    const size_t off_w = off_lohi + ((size_t(n_rows) * sizeof(uint32_t) + 15) & ~size_t(15));
    const size_t off_slot = off_w + size_t(n_cells) * sizeof(double2);
    const size_t off_part = off_slot + ((size_t(n_rows) * sizeof(int) + 15) & ~size_t(15));
    const size_t total = off_part + size_t(grid) * n_rows * sizeof(double);
    uint8_t* ws = nullptr;
    if (cudaMalloc(&ws, total) != cudaSuccess) {
        cudaGetLastError();
        return set_error(SCS_ERR_OOM, "device allocation of %zu bytes failed", total);
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int rc = SCS_OK;
    do {
        if (cudaMemcpyAsync(ws, perm.data(), size_t(n_cells) * sizeof(uint16_t), cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(ws + off_lohi, lohi_sorted.data(), size_t(n_rows) * sizeof(uint32_t), cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(ws + off_slot, slot_row.data(), size_t(n_rows) * sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(ws + off_w, wv.data(), size_t(n_cells) * sizeof(double2), cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemsetAsync(ws + off_part, 0, size_t(grid) * n_rows * sizeof(double), st) != cudaSuccess) {
            rc = set_error(SCS_ERR_CUDA, "per-cell placement set-up failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        PercellParams q;
        q.gt = d_gt;
        q.pitch = pitch_bytes;
        q.n_snv = n_snv;
        q.n_cells = n_cells;
        q.n_rows = n_rows;
        q.row_words = (n_cells + 15) / 16;
        q.perm = reinterpret_cast<const uint16_t*>(ws);
        q.lohi = reinterpret_cast<const uint32_t*>(ws + off_lohi);
        q.w = reinterpret_cast<const double2*>(ws + off_w);
        q.row_keep = d_row_keep;
        q.part = reinterpret_cast<double*>(ws + off_part);
        q.acc_smem = acc_smem ? 1 : 0;
        q.perm_smem = perm_smem ? 1 : 0;
        if (kernel_ms) {
            cudaEventCreate(&ev0);
            cudaEventCreate(&ev1);
            cudaEventRecord(ev0, st);
        }
        if (n_snv > 0) {
            if (kern2) kern2<<<grid, T, smem, st>>>(q);
            else scs_place_percell_kernel<<<grid, T, smem, st>>>(q);
        }
        scs_percell_reduce_kernel<<<(n_rows + 255) / 256, 256, 0, st>>>(q.part, grid, n_rows, reinterpret_cast<const int*>(ws + off_slot), d_soft_weights);
        if (kernel_ms) cudaEventRecord(ev1, st);
        const cudaError_t e = cudaStreamSynchronize(st);           // the workspace is freed below
        if (e != cudaSuccess) {
            rc = set_error(SCS_ERR_CUDA, "per-cell placement failed: %s", cudaGetErrorString(e));
            break;
        }
        if (kernel_ms) cudaEventElapsedTime(kernel_ms, ev0, ev1);
    } while (0);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    cudaFree(ws);
    return rc;
}

}  // extern "C"
