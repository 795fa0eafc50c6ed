// scs_place_iv4.cuh -- the interval placement with everything per-clade-row in REGISTERS.
//
// Same algorithm and the same arithmetic as scs_place_interval3_kernel (scs_place_interval.cuh: rows in
// depth-first cell order, word-wise prefix counts, counts of a clade = one difference of prefix counts,
// leaves outside the sweeps, softmax terms from exact integer differences), restructured around what ncu
// showed limits that kernel (profiles/r01_ncu_interval_final_250k_summary.txt: 21.1k warp instructions per
// SNV, issue slots 45 % busy, short-scoreboard stalls on shared memory, four barriers per row):
//
//   * a thread OWNS its clade rows for the whole kernel: NP pairs of non-leaf rows per thread (pair
//     j = t + i T, the same assignment the sweeps of variant 3 used).  Their prefix-count offsets, their
//     counts / softmax terms of the current SNV and their accumulators live in registers -- no per-row
//     reload of the offset table from L1/L2, no count cache and no accumulator traffic in shared memory;
//     the sweeps are straight-line code over NP unrolled pairs with all their shared-memory loads
//     (2 per row: P[hi], P[lo]) issued back to back;
//   * the row loop is software pipelined over three rows so that ONE pair of barriers serves them all:
//         phase 1:  E(r-1) accumulate   | C(r) counts + best column      | A(r+1) word popcounts, warp scan
//         ---- barrier
//         phase 2:  D(r) terms + row sum                                  | B(r+1) block scan, prefix counts
//         ---- barrier
//     (every hand-over through shared memory -- scan totals, P, best column, row sum -- is written in one
//     phase and read in the next, each buffer has a single writer phase, so no double buffering);
//   * shared memory shrinks to the prefix counts, the leaf accumulators and 160 words of scratch
//     (82 KB at 10,000 cells instead of 160 KB).
//
// Shapes: one CTA of T threads per run of SNV rows, T = words of a permuted row rounded up to a warp
// (<= 640: ~100 registers per thread), NP = ceil(pairs / T) <= 8.  Anything else (more than 10,240 cells,
// more non-leaf rows than 16 T) keeps variant 3.  Small trees simply get small CTAs and several of
// them per SM (1,000 cells: 64 threads), which is what makes this path the faster one from 100 cells on (scs_place.cu, choose_algo).
#pragma once

#include "scs_place_interval.cuh"

namespace scs {

struct Iv4Params {
    const uint32_t* gt;        // rows in depth-first cell order [n_snv][rw], positions beyond n_cells = 3 (missing)
    long long n_snv;
    int rw;                    // 32-bit words of a row: ceil(n_cells / 16), <= blockDim.x
    int n_pairs;               // pairs of non-leaf rows (sorted by (lo, hi)); ceil(n_pairs / blockDim.x) == NP
    int pad_pair;              // the pair whose second slot is padding (odd number of rows), or -1
    const uint4* off4;         // [n_pairs] byte offsets into P: (lo0, hi0, lo1, hi1) * 4
    const uint32_t* leafmask;  // [rw] bit 2e set when position 16 w + e has a leaf row
    int n_leaf_pos;
    const uint8_t* row_keep;   // optional [n_snv]
    const GroupConst* gconst;
    float* part;               // [num_runs][part_stride]: non-leaf accumulators, then (at leaf_off) 16 rw leaf accumulators
    int part_stride, leaf_off;
    int rows_per_run, num_runs;
    float4* leafw;             // SPLIT kernels: [n_snv] (w_wt, w_mut, w_miss, 0) of every kept row for scs_iv4_leaf_kernel (zero elsewhere)
};

constexpr int IV4_SCRATCH_WORDS = 5 * 32;   // (at the start of shared memory) [0,32) scan totals, [32,64) leaf totals, [64,96) best rank value, [96,128) its counts, [128,160) row sums
constexpr int IV4_MAX_THREADS = 640;
constexpr int IV4_MAX_NP = 8;

// Shared-memory byte offset of prefix count P[k].  Word 0 holds P[0] = 0; P[1..] follows from word 4 in 16-byte chunks of
// four counts, chunk c at physical chunk c ^ ((c >> 3) & 3): a thread writes the four chunks 4 t .. 4 t + 3 of its word with
// 16-byte stores, and without the swizzle lanes t, t + 2, t + 4, t + 6 of a quarter warp would hit the same banks (64-byte
// lane stride, 4-way conflict: ncu counted 3.5k shared-memory wavefronts per SNV against 1.8k ideal).  Readers never pay for
// it: the offsets of a clade row are computed once on the host.
__host__ __device__ inline uint32_t iv4_p_offset(uint32_t k) {
    if (k == 0u) return 0u;
    const uint32_t m = k - 1u, c = m >> 2;
    return 4u * (4u + 4u * (c ^ ((c >> 3) & 3u)) + (m & 3u));
}

__device__ __forceinline__ float iv4_warp_max(float v) {
    float m;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
    return m;
}

static inline size_t iv4_smem_bytes(int rw, int threads) {
    return 4 * (size_t(iv3_p_words(rw)) + 16 * size_t(threads) + IV4_SCRATCH_WORDS);
}

template <int NP, bool SPLIT>
__global__ void __launch_bounds__(IV4_MAX_THREADS, 1) scs_place_iv4_kernel(const Iv4Params q) {
    extern __shared__ uint32_t iv_smem[];
    // thread index and CTA size are read ONCE (volatile: ptxas otherwise re-reads the special registers -- S2R, ~20 cycles --
    // and re-derives lane / warp / shared-memory addresses all over the row loop instead of keeping them in registers)
    int t, T;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
    asm volatile("mov.u32 %0, %%ntid.x;" : "=r"(T));
    const int lane = t & 31, warp = t >> 5, nwarp = T >> 5;
    const int rw = q.rw;
    // shared memory: [scratch, 160 words][P[0] = 0 and padding, 4 words][P[1..], 16 rw words, swizzled][leaf accumulators, 16 T words]
    // -- everything the row loop addresses without a thread-dependent index sits at a compile-time offset
    uint32_t* const scr = iv_smem;
    const uint32_t* const P = iv_smem + IV4_SCRATCH_WORDS;      // prefix counts at their swizzled byte offsets (iv4_p_offset); word 0 = P[0] = 0
    const bool own = t < rw;                        // this thread owns word t = positions 16 t .. 16 t + 15
    float4* const accL = reinterpret_cast<float4*>(iv_smem + IV4_SCRATCH_WORDS + iv3_p_words(rw)) + t;     // [4][T]: positions 16 t + 4 i .. + 3 at [i][t]
    const GroupConst gc = q.gconst[0];
    const uint32_t lmask = own ? q.leafmask[t] : 0u;
    const uint32_t* gw = q.gt;
    if (t == 0) iv_smem[IV4_SCRATCH_WORDS] = 0u;    // P[0]: never written again (visible after the first barrier)
    // this thread's four 16-byte chunks of P: chunk 4 t + i sits at physical chunk 4 t + (i ^ ((t >> 1) & 3))
    uint4* const pdst = reinterpret_cast<uint4*>(iv_smem + IV4_SCRATCH_WORDS + 4) + 4 * t;
    const int psw = (t >> 1) & 3;

    // this thread's clade rows: NP pairs, offsets packed lo | hi << 16 (bytes into P: < 64 KB)
    uint32_t offx[NP], offy[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int j = t + i * T;
        const uint4 o = j < q.n_pairs ? q.off4[j] : make_uint4(0u, 0u, 0u, 0u);
        offx[i] = o.x | (o.y << 16);
        offy[i] = o.z | (o.w << 16);
    }
    // only the last pair of a thread can be missing or half padding (NP = ceil(n_pairs / T))
    const int j_last = t + (NP - 1) * T;
    const bool valid_x = j_last < q.n_pairs, valid_y = valid_x && j_last != q.pad_pair;

    const float2 a1f2 = make_float2(gc.a1f, gc.a1f), a0f2 = make_float2(gc.a0f, gc.a0f);
    const float2 a1h2 = make_float2(gc.a1h, gc.a1h), a0h2 = make_float2(gc.a0h, gc.a0h);
    const float2 a1l2 = make_float2(gc.a1l, gc.a1l), a0l2 = make_float2(gc.a0l, gc.a0l);
    const float2 n23 = make_float2(-8388608.0f, -8388608.0f);
    const float n_leaf = float(q.n_leaf_pos);

    for (int run = blockIdx.x; run < q.num_runs; run += gridDim.x) {
        float2 acc[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) acc[i] = make_float2(0.f, 0.f);
        if (!SPLIT && own) {
#pragma unroll
            for (int i = 0; i < 4; ++i) accL[i * T] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const long long r0 = (long long)run * q.rows_per_run;
        const long long r1 = r0 + q.rows_per_run < q.n_snv ? r0 + q.rows_per_run : q.n_snv;

        // ---- pipeline state: row "cur" is in stages C/D, "nxt" in A/B, "prev" in E
        uint32_t w_cur = own ? gw[size_t(r0) * rw + t] : 0xFFFFFFFFu;
        uint32_t keep_cur = q.row_keep != nullptr ? q.row_keep[r0] : 1u;
        uint32_t w_nxt = 0xFFFFFFFFu, keep_nxt = 0u;
        if (r0 + 1 < r1) {
            if (own) w_nxt = gw[size_t(r0 + 1) * rw + t];
            keep_nxt = q.row_keep != nullptr ? q.row_keep[r0 + 1] : 1u;
        }
        uint32_t w_prev = 0u, keep_prev = 0u;
        uint32_t ltot_cur = 0u, ltot_nxt = 0u;      // leaf totals (mutated | wild type << 16) of the rows in C/D and A/B
        uint32_t sc_incl = 0u, sc_tot = 0u;         // A -> B: this word's inclusive warp scan and its own popcounts
        uint32_t dx[NP], dy[NP];                    // C -> D: packed counts; D -> E: softmax terms (float bits)
        float tl_wt = 0.f, tl_mut = 0.f, tl_miss = 0.f, leaf_total = 0.f;      // D -> E: the three leaf terms and their share of the row sum

        // stage A: popcounts of this word, inclusive scan over the warp, leaf totals; warp totals -> scratch
        auto stage_a = [&](uint32_t w) {
            const uint32_t mm = (w ^ (w >> 1)) & 0x55555555u, wm = ~(w | (w >> 1)) & 0x55555555u;
            const uint32_t tot = uint32_t(__popc(mm)) | (uint32_t(__popc(wm)) << 16);
            uint32_t leaf_sum = uint32_t(__popc(mm & lmask)) | (uint32_t(__popc(wm & lmask)) << 16);
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            leaf_sum = __reduce_add_sync(0xffffffffu, leaf_sum);      // (both halves stay below 2^16: no carry between them)
            if (lane == 31) {
                scr[warp] = incl;
                scr[32 + warp] = leaf_sum;
            }
            sc_incl = incl;
            sc_tot = tot;
        };
        // stage B: scan of the warp totals, this word's sixteen prefix counts -> P; returns the row's leaf totals
        auto stage_b = [&](uint32_t w) -> uint32_t {
            const uint32_t mm = (w ^ (w >> 1)) & 0x55555555u, wm = ~(w | (w >> 1)) & 0x55555555u;
            // prefix before this warp and the row's leaf totals: one integer warp reduction each over the warp totals
            const uint32_t wbase = __reduce_add_sync(0xffffffffu, lane < warp ? scr[lane] : 0u);
            const uint32_t ltot = __reduce_add_sync(0xffffffffu, lane < nwarp ? scr[32 + lane] : 0u);
            if (own) {
                const uint32_t base = wbase + sc_incl - sc_tot;                      // prefix before this word
                // the sixteen inclusive prefix counts of the word, four positions per 32-bit operation: plane bits of positions
                // 4 q + r gathered into byte q (A_r), byte-wise running sums over r, and one multiply for the carry-in of the
                // groups q' < q; PRMT then zips a mutated-count byte and a wild-type-count byte into the packed 16 + 16 format
                // (selector nibbles 8 | q replicate the -- always clear -- sign bit of a count byte: zero bytes)
                uint32_t pm[4], pw[4];
                {
                    const uint32_t a0 = mm & 0x01010101u, a1 = (mm >> 2) & 0x01010101u, a2 = (mm >> 4) & 0x01010101u, a3 = (mm >> 6) & 0x01010101u;
                    const uint32_t s1 = a0 + a1, s2 = s1 + a2, s3 = s2 + a3, ex = s3 * 0x01010100u;
                    pm[0] = ex + a0; pm[1] = ex + s1; pm[2] = ex + s2; pm[3] = ex + s3;
                }
                {
                    const uint32_t a0 = wm & 0x01010101u, a1 = (wm >> 2) & 0x01010101u, a2 = (wm >> 4) & 0x01010101u, a3 = (wm >> 6) & 0x01010101u;
                    const uint32_t s1 = a0 + a1, s2 = s1 + a2, s3 = s2 + a3, ex = s3 * 0x01010100u;
                    pw[0] = ex + a0; pw[1] = ex + s1; pw[2] = ex + s2; pw[3] = ex + s3;
                }
                uint32_t pv[16];
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        uint32_t z;
                        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(z) : "r"(pm[rr]), "r"(pw[rr]), "r"(uint32_t(qq | ((8 | qq) << 4) | ((4 + qq) << 8) | ((8 | qq) << 12))));
                        pv[4 * qq + rr] = base + z;
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) pdst[i ^ psw] = make_uint4(pv[4 * i], pv[4 * i + 1], pv[4 * i + 2], pv[4 * i + 3]);
            }
            return ltot;
        };
        // stage E: normalised terms into the accumulators
        auto stage_e = [&](uint32_t w, long long row) {
            float total = lane < nwarp ? __uint_as_float(scr[128 + lane]) : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
            const float inv = 1.0f / (total + leaf_total);        // >= 1: the best column's own term
            const float2 inv2 = make_float2(inv, inv);
#pragma unroll
            for (int i = 0; i < NP; ++i) acc[i] = __ffma2_rn(make_float2(__uint_as_float(dx[i]), __uint_as_float(dy[i])), inv2, acc[i]);
            if (SPLIT) {
                // leaves are accumulated by scs_iv4_leaf_kernel from the three weights of the row
                if (t == 0) q.leafw[row] = make_float4(tl_wt * inv, tl_mut * inv, tl_miss * inv, 0.f);
            } else if (own) {
                // leaves: the term by the position's code (positions without a leaf row collect numbers nobody reads)
                const uint32_t mm = (w ^ (w >> 1)) & 0x55555555u, wm = ~(w | (w >> 1)) & 0x55555555u;
                const float w_wt = tl_wt * inv, w_mut = tl_mut * inv, w_miss = tl_miss * inv;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 a = accL[i * T];
                    float sel[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int e = 4 * i + k;
                        sel[k] = ((mm >> (2 * e)) & 1u) ? w_mut : (((wm >> (2 * e)) & 1u) ? w_wt : w_miss);
                    }
                    const float2 lo2 = __fadd2_rn(make_float2(a.x, a.y), make_float2(sel[0], sel[1]));
                    const float2 hi2 = __fadd2_rn(make_float2(a.z, a.w), make_float2(sel[2], sel[3]));
                    accL[i * T] = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
                }
            }
        };

        if (keep_cur) stage_a(w_cur);
        __syncthreads();
        if (keep_cur) ltot_cur = stage_b(w_cur);
        __syncthreads();

        for (long long r = r0; r < r1; ++r) {
            // ================= phase 1: E(r-1) | C(r) | A(r+1)
            if (keep_prev) stage_e(w_prev, r - 1);
            if (keep_cur) {
                // counts of this thread's rows: both counts of a row in one subtraction (no borrow: P is monotone in each half)
                float bv = -INFINITY;
                uint32_t bd = 0u;
#pragma unroll
                for (int i = 0; i < NP; ++i) {
                    dx[i] = iv_lds_off(P, offx[i] >> 16) - iv_lds_off(P, offx[i] & 0xFFFFu);
                    dy[i] = iv_lds_off(P, offy[i] >> 16) - iv_lds_off(P, offy[i] & 0xFFFFu);
                }
                // best of this thread's rows: two interleaved chains (even / odd pairs) instead of one chain of 2 NP dependent
                // compare/selects, merged at the end; on ties the earlier row wins (a later one must be strictly greater),
                // which is what one sequential scan would give
                float bv1 = -INFINITY;
                uint32_t bd1 = 0u;
#pragma unroll
                for (int i = 0; i < NP; ++i) {
                    const float2 c1 = __fadd2_rn(iv_lo16(dx[i], dy[i]), n23), c0 = __fadd2_rn(iv_hi16(dx[i], dy[i]), n23);
                    float2 v = __ffma2_rn(a1f2, c1, __fmul2_rn(a0f2, c0));
                    if (i == NP - 1) {
                        if (!valid_x) v.x = -INFINITY;
                        if (!valid_y) v.y = -INFINITY;
                    }
                    const bool y_wins = v.y > v.x;
                    const float pvv = y_wins ? v.y : v.x;
                    const uint32_t pdd = y_wins ? dy[i] : dx[i];
                    if (i & 1) {
                        if (pvv > bv1) {
                            bv1 = pvv;
                            bd1 = pdd;
                        }
                    } else {
                        if (pvv > bv) {
                            bv = pvv;
                            bd = pdd;
                        }
                    }
                }
                // merge: chain 0 holds pairs 0, 2, .. and chain 1 pairs 1, 3, ..; the earlier ROW wins a tie.  Rows are not
                // compared by index here: equal rank values with different counts are told apart below by the counts, and equal
                // counts are the same number either way
                if (bv1 > bv || (bv1 == bv && bd1 > bd)) {
                    bv = bv1;
                    bd = bd1;
                }
                // across threads: the largest rank value, ties to the larger packed counts -- a total order, so the result does
                // not depend on the reduction shape (two warp-wide reductions: redux.sync.max.f32, then max over the tied counts)
                {
                    const float m = iv4_warp_max(bv);
                    bd = __reduce_max_sync(0xffffffffu, bv == m ? bd : 0u);
                    bv = m;
                }
                if (lane == 0) {
                    scr[64 + warp] = __float_as_uint(bv);
                    scr[96 + warp] = bd;
                }
            }
            const bool has_next = r + 1 < r1;
            uint32_t w_nn = 0xFFFFFFFFu, keep_nn = 0u;
            if (r + 2 < r1) {                       // two rows ahead: consumed by stage A of the next iteration
                if (own) w_nn = gw[size_t(r + 2) * rw + t];
                keep_nn = q.row_keep != nullptr ? q.row_keep[r + 2] : 1u;
            }
            if (has_next && keep_nxt) stage_a(w_nxt);
            __syncthreads();

            // ================= phase 2: D(r) | B(r+1)
            if (keep_cur) {
                float bv = lane < nwarp ? __uint_as_float(scr[64 + lane]) : -INFINITY;
                uint32_t bd = lane < nwarp ? scr[96 + lane] : 0u;
                {
                    const float m = iv4_warp_max(bv);
                    bd = __reduce_max_sync(0xffffffffu, bv == m ? bd : 0u);
                    bv = m;
                }
                // the three leaf columns compete too: mutated (C1, C0) = (1, 0), wild type (0, 1), missing (0, 0)
                const float n_mut = float(ltot_cur & 0xFFFFu), n_wt = float(ltot_cur >> 16), n_miss = n_leaf - n_mut - n_wt;
                if (n_mut > 0.f && (gc.a1f > bv || (gc.a1f == bv && 1u > bd))) {
                    bv = gc.a1f;
                    bd = 1u;
                }
                if (n_wt > 0.f && (gc.a0f > bv || (gc.a0f == bv && 0x10000u > bd))) {
                    bv = gc.a0f;
                    bd = 0x10000u;
                }
                if (n_miss > 0.f && 0.f > bv) {
                    bv = 0.f;
                    bd = 0u;
                }
                // reference counts: e = count - reference is an exact integer difference
                const float rf1 = float(bd & 0xFFFFu), rf0 = float(bd >> 16);
                const float2 nb1_2 = make_float2(-(8388608.0f + rf1), -(8388608.0f + rf1)), nb0_2 = make_float2(-(8388608.0f + rf0), -(8388608.0f + rf0));
                {
                    const float e1m = 1.f - rf1, e1z = -rf1, e0w = 1.f - rf0, e0z = -rf0;
                    tl_wt = ex2_approx(fmaf(gc.a1h, e1z, gc.a0h * e0w) + fmaf(gc.a1l, e1z, gc.a0l * e0w));
                    tl_mut = ex2_approx(fmaf(gc.a1h, e1m, gc.a0h * e0z) + fmaf(gc.a1l, e1m, gc.a0l * e0z));
                    tl_miss = ex2_approx(fmaf(gc.a1h, e1z, gc.a0h * e0z) + fmaf(gc.a1l, e1z, gc.a0l * e0z));
                }
                leaf_total = fmaf(n_mut, tl_mut, fmaf(n_wt, tl_wt, n_miss * tl_miss));
                // terms relative to the best column (in place of the counts), row sum
                float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < NP; ++i) {
                    const float2 e1 = __fadd2_rn(iv_lo16(dx[i], dy[i]), nb1_2), e0 = __fadd2_rn(iv_hi16(dx[i], dy[i]), nb0_2);
                    const float2 sdiff = __ffma2_rn(a1h2, e1, __fmul2_rn(a0h2, e0));      // exact operands, one rounding
                    const float2 lo = __ffma2_rn(a1l2, e1, __fmul2_rn(a0l2, e0));
                    const float2 x = __fadd2_rn(sdiff, lo);
                    float2 tt = make_float2(ex2_approx(x.x), ex2_approx(x.y));
                    if (i == NP - 1) {
                        if (!valid_x) tt.x = 0.f;
                        if (!valid_y) tt.y = 0.f;
                    }
                    s2 = __fadd2_rn(s2, tt);
                    dx[i] = __float_as_uint(tt.x);
                    dy[i] = __float_as_uint(tt.y);
                }
                float sum = s2.x + s2.y;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                if (lane == 0) scr[128 + warp] = __float_as_uint(sum);
            }
            if (has_next && keep_nxt) ltot_nxt = stage_b(w_nxt);
            __syncthreads();

            keep_prev = keep_cur;
            w_prev = w_cur;
            keep_cur = has_next ? keep_nxt : 0u;
            w_cur = w_nxt;
            ltot_cur = ltot_nxt;
            keep_nxt = keep_nn;
            w_nxt = w_nn;
        }
        if (keep_prev) stage_e(w_prev, r1 - 1);

        float* out = q.part + size_t(run) * q.part_stride;
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            const int j = t + i * T;
            if (j < q.n_pairs) reinterpret_cast<float2*>(out)[j] = acc[i];
        }
        if (!SPLIT && own) {
#pragma unroll
            for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(out + q.leaf_off)[4 * t + i] = accL[i * T];
        }
    }
}

// The leaf accumulators of the SPLIT variant: thread = one word (16 positions) of the rows of one run, accumulators in
// registers, the row's three weights read once per row (the same address for the whole block).  Same order of additions per
// leaf as the fused kernel (rows in order).
__global__ void __launch_bounds__(128) scs_iv4_leaf_kernel(const Iv4Params q) {
    const int wd = blockIdx.x * blockDim.x + threadIdx.x;
    const int run = blockIdx.y;
    if (wd >= q.rw) return;
    const long long r0 = (long long)run * q.rows_per_run;
    const long long r1 = r0 + q.rows_per_run < q.n_snv ? r0 + q.rows_per_run : q.n_snv;
    float acc[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = 0.f;
    const uint32_t* gp = q.gt + size_t(r0) * q.rw + wd;
    uint32_t w = r0 < r1 ? *gp : 0u;
    for (long long r = r0; r < r1; ++r) {
        const float4 lw = __ldg(q.leafw + r);
        const uint32_t wc = w;
        gp += q.rw;
        if (r + 1 < r1) w = *gp;
        const uint32_t mm = (wc ^ (wc >> 1)) & 0x55555555u, wm = ~(wc | (wc >> 1)) & 0x55555555u;
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] += ((mm >> (2 * e)) & 1u) ? lw.y : (((wm >> (2 * e)) & 1u) ? lw.x : lw.z);
    }
    float* out = q.part + size_t(run) * q.part_stride + q.leaf_off + 16 * wd;
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(out)[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
}

template <int NP>
static void iv4_launch_np(const Iv4Params& q, int grid, int threads, size_t smem, cudaStream_t st) {
    if (q.leafw) {
        cudaFuncSetAttribute(scs_place_iv4_kernel<NP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        scs_place_iv4_kernel<NP, true><<<grid, threads, smem, st>>>(q);
        scs_iv4_leaf_kernel<<<dim3(unsigned((q.rw + 127) / 128), unsigned(q.num_runs)), 128, 0, st>>>(q);
    } else {
        scs_place_iv4_kernel<NP, false><<<grid, threads, smem, st>>>(q);
    }
}

// Resident CTAs per SM for this shape (0 when the shape is not supported); also raises the kernel's dynamic shared-memory limit.
template <int NP>
static int iv4_occupancy_np(int threads, size_t smem) {
    int per_sm = 0;
    if (cudaFuncSetAttribute(scs_place_iv4_kernel<NP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scs_place_iv4_kernel<NP, false>, threads, smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return per_sm;
}

static int iv4_occupancy(int np, int threads, size_t smem) {
    switch (np) {
        case 1: return iv4_occupancy_np<1>(threads, smem);
        case 2: return iv4_occupancy_np<2>(threads, smem);
        case 3: return iv4_occupancy_np<3>(threads, smem);
        case 4: return iv4_occupancy_np<4>(threads, smem);
        case 5: return iv4_occupancy_np<5>(threads, smem);
        case 6: return iv4_occupancy_np<6>(threads, smem);
        case 7: return iv4_occupancy_np<7>(threads, smem);
        case 8: return iv4_occupancy_np<8>(threads, smem);
    }
    return 0;
}

static void iv4_launch(int np, const Iv4Params& q, int grid, int threads, size_t smem, cudaStream_t st) {
    switch (np) {
        case 1: iv4_launch_np<1>(q, grid, threads, smem, st); break;
        case 2: iv4_launch_np<2>(q, grid, threads, smem, st); break;
        case 3: iv4_launch_np<3>(q, grid, threads, smem, st); break;
        case 4: iv4_launch_np<4>(q, grid, threads, smem, st); break;
        case 5: iv4_launch_np<5>(q, grid, threads, smem, st); break;
        case 6: iv4_launch_np<6>(q, grid, threads, smem, st); break;
        case 7: iv4_launch_np<7>(q, grid, threads, smem, st); break;
        case 8: iv4_launch_np<8>(q, grid, threads, smem, st); break;
    }
}

}  // namespace scs
