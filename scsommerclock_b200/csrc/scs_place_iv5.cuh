// scs_place_iv5.cuh -- the interval placement for BATCHES OF SMALL TREES (replicate sweeps: thousands of (matrix, tree)
// pairs of <= 512 cells), a few lanes per SNV instead of a CTA per SNV run.
//
// Same algorithm and arithmetic as scs_place_iv4.cuh (run_PT_test.py:386-448 evaluated through per-SNV prefix counts in
// a depth-first cell order: both counts of a clade are one difference; leaves outside the generic rows; softmax terms
// from exact integer differences against the best column).  What differs is the shape: at 100 cells a row is 7 words and
// a tree has ~100 internal nodes, so
//   * L = 8, 16 or 32 lanes (16 L >= cells) own one SNV; a warp walks 32 / L SNVs of the SAME group in lockstep;
//   * lane l holds word l of the row in depth-first cell order = the codes of leaves 16 l .. 16 l + 15, and owns the
//     accumulators of exactly those leaves (registers) plus up to 16 internal-node rows (interval bounds + accumulators
//     in registers, row j = i L + l);
//   * the prefix counts of a row go through a double-buffered table in shared memory (132 words per SNV at L = 8), one
//     warp barrier per row, no block barrier anywhere in the row loop;
//   * the zero rows of S (rows beyond the tree's nodes, the reference's final "FP" row) all have the counts (0, 0) of a
//     missing leaf: they ride on the missing-leaf term with a multiplicity instead of taking row slots.
// Per group the depth-first order and the intervals come from the tree's child table (scs_iv5_tables_kernel: one thread
// walks a <= 1,022-node tree).  The rows are NOT rewritten in that order by a pre-pass: the L lanes of an SNV load the
// source row's words and each lane gathers its sixteen 2-bit fields with one shuffle + one rotate each (a separate
// permutation pass cost 1.4 ms per 12.5M rows -- as much as a third of this kernel -- plus 0.8 GB of traffic).  Per-task
// partial sums are reduced in fixed order in float64 (scs_iv5_reduce_kernel): deterministic.
#pragma once

#include "scs_place_interval.cuh"

namespace scs {

constexpr int IV5_MAX_NS = 16;          // internal-node rows per lane
constexpr int IV5_THREADS = 128;
constexpr int IV5_MIN_CTAS = 4;          // resident CTAs per SM the register budget is capped for (128 registers per thread)
constexpr int IV5_MAX_CELLS = 512;

struct Iv5Task {
    int group;
    int n_rows;             // SNV rows of this task
    long long row0;         // first row (index into the permuted matrix and row_keep)
};

struct Iv5Params {
    const uint8_t* gt;          // the packed genotype rows as the caller holds them (source column order), `pitch` bytes per row
    long long pitch;
    int n_cols;
    const uint16_t* perm;       // [G][16 L]: source column at depth-first position p (0xFFFF: none)
    const uint8_t* row_keep;    // optional [total rows]
    const Iv5Task* tasks;
    int n_tasks;
    int n_rows;                 // clade rows per group
    const uint32_t* nl_lohi;    // [G][IV5_MAX_NS * L]: lo | hi << 16 of internal-node row j (positions)
    const int* meta;            // [G][4]: leaves, internal rows, zero rows, status
    const GroupConst* gconst;   // [G]
    float* part;                // [n_tasks][stride]: internal rows (NS L), leaves (16 L), the zero-row accumulator
    int stride;
};

// -------------------------------------------------------------------------------------------------------------------
// Tables of one group from its child table: depth-first leaf order, interval of every node, slot of every clade row.
//   perm[g][16 L]       source column at depth-first position p, 0xFFFF beyond the leaves
//   nl_lohi[g][NS L]    interval of internal-node row j
//   row_slot[g][n_rows] where a clade row's accumulator sits in a partial-sum row: j | NS L + p (leaf) | NS L + 16 L (zero row)
//   meta[g][4]          leaves, internal rows, zero rows, status (1: the child table is not a forest of binary trees over single-cell leaves)
__global__ void scs_iv5_tables_kernel(const uint32_t* __restrict__ bits, long long pitch_words, const int* __restrict__ children,
                                      const int* __restrict__ info, int n_rows, int n_cols, int L, uint16_t* __restrict__ perm,
                                      uint32_t* __restrict__ nl_lohi, int* __restrict__ row_slot, int* __restrict__ meta) {
    extern __shared__ int iv5_tab_smem[];
    const int g = blockIdx.x, t = threadIdx.x;
    int* lo = iv5_tab_smem;
    int* hi = lo + n_rows;
    int* stack = hi + n_rows;
    int* is_child = stack + n_rows + 2;
    int* ch = is_child + n_rows;               // [2 n_rows] the child table
    int* leafcol = ch + 2 * n_rows;            // [n_rows] column of a leaf row (-1: its mask is not a single cell)
    __shared__ int s_status, s_leaves, s_internal;
    const int nodes = min(max(info[g * 4 + 0], 0), n_rows);
    const int NSL = IV5_MAX_NS * L, NPOS = 16 * L;
    // everything thread 0's walk touches is staged in shared memory first (the walk is one chain of dependent loads)
    for (int i = t; i < 2 * n_rows; i += blockDim.x) ch[i] = children[size_t(g) * n_rows * 2 + i];
    for (int r = t; r < n_rows; r += blockDim.x) {
        is_child[r] = 0;
        lo[r] = hi[r] = 0;
        int col = -1, cnt = 0;
        if (r < nodes)
            for (int wd = 0; wd < (n_cols + 31) / 32; ++wd) {
                const uint32_t b = bits[(size_t(g) * n_rows + r) * pitch_words + wd];
                if (b) {
                    cnt += __popc(b);
                    col = wd * 32 + __ffs(b) - 1;
                }
            }
        leafcol[r] = (cnt == 1 && col < n_cols) ? col : -1;
    }
    for (int p = t; p < NPOS; p += blockDim.x) perm[size_t(g) * NPOS + p] = 0xFFFFu;
    for (int j = t; j < NSL; j += blockDim.x) nl_lohi[size_t(g) * NSL + j] = 0u;
    __syncthreads();
    if (t == 0) {
        int status = 0;
        for (int r = 0; r < nodes; ++r)
            for (int c = 0; c < 2; ++c) {
                const int k = ch[2 * r + c];
                if (k >= nodes || k == r) status = 1;
                else if (k >= 0) {
                    if (is_child[k]) status = 1;      // two parents
                    is_child[k] = 1;
                }
            }
        int next = 0, steps = 0;
        for (int r0 = 0; r0 < nodes && status == 0; ++r0) {
            if (is_child[r0]) continue;
            int sp = 0;
            stack[sp++] = r0 << 1;
            while (sp > 0 && status == 0) {
                if (++steps > 4 * n_rows + 8) {
                    status = 1;         // a cycle
                    break;
                }
                const int top = stack[sp - 1], row = top >> 1;
                const int c0 = ch[2 * row], c1 = ch[2 * row + 1];
                if ((top & 1) == 0) {
                    lo[row] = next;
                    if (c0 < 0 && c1 < 0) {
                        // a leaf: its column is the single bit of its mask
                        const int col = leafcol[row];
                        if (col < 0 || next >= NPOS) {
                            status = 1;
                        } else {
                            perm[size_t(g) * NPOS + next] = uint16_t(col);
                            ++next;
                        }
                        hi[row] = next;
                        --sp;
                    } else if (c0 < 0 || c1 < 0) {
                        status = 1;     // unary node
                    } else {
                        stack[sp - 1] = (row << 1) | 1;
                        stack[sp++] = c1 << 1;
                        stack[sp++] = c0 << 1;
                    }
                } else {
                    hi[row] = next;
                    --sp;
                }
            }
        }
        int n_int = 0;
        for (int r = 0; r < n_rows && status == 0; ++r) {
            int slot;
            if (r >= nodes) {
                slot = NSL + NPOS;
            } else if (ch[2 * r] < 0) {
                slot = NSL + lo[r];
            } else {
                if (n_int >= NSL) {
                    status = 1;
                    break;
                }
                nl_lohi[size_t(g) * NSL + n_int] = uint32_t(lo[r]) | (uint32_t(hi[r]) << 16);
                slot = n_int++;
            }
            row_slot[size_t(g) * n_rows + r] = slot;
        }
        s_status = status;
        s_leaves = next;
        s_internal = n_int;
    }
    __syncthreads();
    if (t == 0) {
        meta[g * 4 + 0] = s_status ? 0 : s_leaves;
        meta[g * 4 + 1] = s_status ? 0 : s_internal;
        meta[g * 4 + 2] = s_status ? 0 : n_rows - nodes;
        meta[g * 4 + 3] = s_status;
    }
}

// -------------------------------------------------------------------------------------------------------------------
// L lanes per SNV, NS (even) internal-node rows per lane, processed in pairs with packed fp32x2 arithmetic.
template <int L, int NS>
__global__ void __launch_bounds__(IV5_THREADS, IV5_MIN_CTAS) scs_place_iv5_kernel(const Iv5Params q) {
    constexpr int NSUB = IV5_THREADS / L;           // SNVs in flight per CTA
    constexpr int PW = 16 * L + 4;                  // words of one prefix-count table: P[0] at word 3, P[k] at 3 + k
    constexpr int NSL = IV5_MAX_NS * L, NPOS = 16 * L, NPAIR = NS / 2;
    extern __shared__ __align__(16) uint32_t iv5_smem[];
    // [2][NSUB][PW] prefix tables | [16][L] gather table (source lane), [16][L] (rotation) | after the row loop: [NSUB][stride] accumulators
    const int t = threadIdx.x, sl = t % L, sub = t / L;
    uint32_t* const ptab = iv5_smem + sub * PW;
    uint32_t* const g_lane = iv5_smem + 2 * NSUB * PW;
    uint32_t* const g_rot = g_lane + 16 * L;
    float* const stage = reinterpret_cast<float*>(g_rot + 16 * L);
    const int src_words = (q.n_cols + 15) / 16;
    for (int task = blockIdx.x; task < q.n_tasks; task += gridDim.x) {
        const Iv5Task tk = q.tasks[task];
        const int g = tk.group;
        const GroupConst gc = q.gconst[g];
        const int n_leaves = q.meta[g * 4 + 0], n_int = q.meta[g * 4 + 1];
        const float n_zero = float(q.meta[g * 4 + 2]);
        // gather table of this group: position p = 16 l + e takes the code of source column c = perm[p], i.e. bits
        // 2 (c % 16) .. of word c / 16: lane c / 16, rotated left by 2 e - 2 (c % 16) so that the field lands at bits 2 e ..
        uint32_t padmask = 0u;                       // positions without a cell read "missing" (3)
        for (int k = t; k < 16 * L; k += IV5_THREADS) {
            const int e = k / L, l = k % L;
            const uint32_t c = q.perm[size_t(g) * NPOS + 16 * l + e];
            g_lane[k] = c == 0xFFFFu ? uint32_t(l) : (c >> 4);
            g_rot[k] = c == 0xFFFFu ? 0u : uint32_t((2 * e - 2 * int(c & 15u)) & 31);
        }
#pragma unroll
        for (int e = 0; e < 16; ++e)
            if (q.perm[size_t(g) * NPOS + 16 * sl + e] == 0xFFFFu) padmask |= 3u << (2 * e);
        // this lane's internal-node rows: byte offsets of P[lo], P[hi]; row j = i L + sl
        // A slot without a row (j >= n_int) points at two constants in front of the table, words 0 and 1 = 0 and 0xFFFF0000:
        // its "counts" are (0, 65535), its rank value ~ -3e5 and its term exactly 0 -- no predicates in the row loop.
        uint32_t off[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const int j = i * L + sl;
            const uint32_t lh = q.nl_lohi[size_t(g) * NSL + j];
            off[i] = j < n_int ? ((4u * (3u + (lh & 0xFFFFu))) | ((4u * (3u + (lh >> 16))) << 16)) : ((4u * 0u) | ((4u * 1u) << 16));
        }
        // leaf positions of this lane's word that exist
        const int npos = min(max(n_leaves - 16 * sl, 0), 16);
        const uint32_t lmask = npos >= 16 ? 0x55555555u : (0x55555555u & ((1u << (2 * npos)) - 1u));
        float2 acc[NPAIR];
        float accL[16], acc_zero = 0.f;
#pragma unroll
        for (int i = 0; i < NPAIR; ++i) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int e = 0; e < 16; ++e) accL[e] = 0.f;
        if (sl == 0) {
            ptab[3] = 0u;                            // P[0] of both buffers
            ptab[NSUB * PW + 3] = 0u;
            ptab[0] = 0u;                            // the constants of the slots without a row
            ptab[1] = 0xFFFF0000u;
            ptab[NSUB * PW + 0] = 0u;
            ptab[NSUB * PW + 1] = 0xFFFF0000u;
        }
        __syncthreads();                             // gather table
        const float2 a1f2 = make_float2(gc.a1f, gc.a1f), a0f2 = make_float2(gc.a0f, gc.a0f);
        const float2 a1h2 = make_float2(gc.a1h, gc.a1h), a0h2 = make_float2(gc.a0h, gc.a0h);
        const float2 a1l2 = make_float2(gc.a1l, gc.a1l), a0l2 = make_float2(gc.a0l, gc.a0l);
        const float2 n23 = make_float2(-8388608.0f, -8388608.0f);
        const float n_leaf_f = float(n_leaves);
        int buf = 0;
        const int n = tk.n_rows;
        const uint8_t* const src = q.gt + tk.row0 * q.pitch;
        const bool has_word = sl < src_words;
        uint32_t s_next = (sub < n && has_word) ? reinterpret_cast<const uint32_t*>(src + size_t(sub) * q.pitch)[sl] : 0xFFFFFFFFu;
        uint32_t k_next = sub < n ? (q.row_keep ? q.row_keep[tk.row0 + sub] : 1u) : 0u;
        // every sub-warp makes the same number of trips (rows beyond the task: keep = 0), so the warp stays converged
        const int trips = (n + NSUB - 1) / NSUB;
        for (int it = 0; it < trips; ++it) {
            const uint32_t ws = s_next, keep = k_next;
            {
                const int rn = (it + 1) * NSUB + sub;
                s_next = (rn < n && has_word) ? reinterpret_cast<const uint32_t*>(src + size_t(rn) * q.pitch)[sl] : 0xFFFFFFFFu;
                k_next = rn < n ? (q.row_keep ? q.row_keep[tk.row0 + rn] : 1u) : 0u;
            }
            // this lane's word in depth-first cell order: sixteen 2-bit fields gathered from the sub-warp's source words
            uint32_t w = padmask;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const uint32_t x = __shfl_sync(0xffffffffu, ws, int(g_lane[e * L + sl]), L);
                w |= __funnelshift_l(x, x, g_rot[e * L + sl]) & (3u << (2 * e));
            }
            const uint32_t m2 = (w ^ (w >> 1)) & 0x55555555u, z2 = ~(w | (w >> 1)) & 0x55555555u;
            const uint32_t mm = m2 & lmask, wm = z2 & lmask;
            const uint32_t tot = uint32_t(__popc(mm)) | (uint32_t(__popc(wm)) << 16);
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < L; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o, L);
                if (sl >= o) incl += v;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, L - 1, L);
            const uint32_t base = incl - tot;
            // the sixteen inclusive prefix counts of this word (SWAR, as in scs_place_iv4.cuh) -> P[16 sl + 1 ..]
            uint32_t* const P = ptab + buf * (NSUB * PW);
            {
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
                uint4* dst = reinterpret_cast<uint4*>(P + 4 + 16 * sl);
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    uint32_t pv[4];
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        uint32_t z;
                        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(z) : "r"(pm[rr]), "r"(pw[rr]), "r"(uint32_t(qq | ((8 | qq) << 4) | ((4 + qq) << 8) | ((8 | qq) << 12))));
                        pv[rr] = base + z;
                    }
                    dst[qq] = make_uint4(pv[0], pv[1], pv[2], pv[3]);
                }
            }
            __syncwarp();
            // counts of this lane's internal-node rows; the largest rank value, then the largest packed counts among the
            // rows that reach it (a total order: the result does not depend on lane assignment)
            uint32_t d[NS];
            float2 rank[NPAIR];
            float bv = -INFINITY;
#pragma unroll
            for (int i = 0; i < NPAIR; ++i) {
                d[2 * i] = iv_lds_off(P, off[2 * i] >> 16) - iv_lds_off(P, off[2 * i] & 0xFFFFu);
                d[2 * i + 1] = iv_lds_off(P, off[2 * i + 1] >> 16) - iv_lds_off(P, off[2 * i + 1] & 0xFFFFu);
                const float2 c1 = __fadd2_rn(iv_lo16(d[2 * i], d[2 * i + 1]), n23), c0 = __fadd2_rn(iv_hi16(d[2 * i], d[2 * i + 1]), n23);
                const float2 v = __ffma2_rn(a1f2, c1, __fmul2_rn(a0f2, c0));
                rank[i] = v;
                bv = fmaxf(bv, fmaxf(v.x, v.y));
            }
#pragma unroll
            for (int o = L / 2; o > 0; o >>= 1) bv = fmaxf(bv, __shfl_xor_sync(0xffffffffu, bv, o, L));
            uint32_t bd = 0u;
#pragma unroll
            for (int i = 0; i < NPAIR; ++i) {
                if (rank[i].x == bv) bd = max(bd, d[2 * i]);
                if (rank[i].y == bv) bd = max(bd, d[2 * i + 1]);
            }
#pragma unroll
            for (int o = L / 2; o > 0; o >>= 1) bd = max(bd, __shfl_xor_sync(0xffffffffu, bd, o, L));
            // the leaf columns and the zero rows compete too: mutated (1, 0), wild type (0, 1), missing / zero row (0, 0)
            const float n_mut = float(total & 0xFFFFu), n_wt = float(total >> 16), n_miss = n_leaf_f - n_mut - n_wt + n_zero;
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
            const float rf1 = float(bd & 0xFFFFu), rf0 = float(bd >> 16);
            const float2 nb1_2 = make_float2(-(8388608.0f + rf1), -(8388608.0f + rf1)), nb0_2 = make_float2(-(8388608.0f + rf0), -(8388608.0f + rf0));
            float tl_wt, tl_mut, tl_miss;
            {
                const float e1m = 1.f - rf1, e1z = -rf1, e0w = 1.f - rf0, e0z = -rf0;
                tl_wt = ex2_approx(fmaf(gc.a1h, e1z, gc.a0h * e0w) + fmaf(gc.a1l, e1z, gc.a0l * e0w));
                tl_mut = ex2_approx(fmaf(gc.a1h, e1m, gc.a0h * e0z) + fmaf(gc.a1l, e1m, gc.a0l * e0z));
                tl_miss = ex2_approx(fmaf(gc.a1h, e1z, gc.a0h * e0z) + fmaf(gc.a1l, e1z, gc.a0l * e0z));
            }
            float2 s2 = make_float2(0.f, 0.f);
            float2 tt[NPAIR];
#pragma unroll
            for (int i = 0; i < NPAIR; ++i) {
                const float2 e1 = __fadd2_rn(iv_lo16(d[2 * i], d[2 * i + 1]), nb1_2), e0 = __fadd2_rn(iv_hi16(d[2 * i], d[2 * i + 1]), nb0_2);      // exact integer differences
                const float2 sdiff = __ffma2_rn(a1h2, e1, __fmul2_rn(a0h2, e0));
                const float2 lo = __ffma2_rn(a1l2, e1, __fmul2_rn(a0l2, e0));
                const float2 x = __fadd2_rn(sdiff, lo);
                const float2 tv = make_float2(ex2_approx(x.x), ex2_approx(x.y));      // (a slot without a row: 2^-3e5 = 0)
                tt[i] = tv;
                s2 = __fadd2_rn(s2, tv);
            }
            float sum = s2.x + s2.y;
#pragma unroll
            for (int o = L / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o, L);
            sum += fmaf(n_mut, tl_mut, fmaf(n_wt, tl_wt, n_miss * tl_miss));
            const float inv = keep ? 1.0f / sum : 0.f;           // filtered rows (and the padding trips) add nothing
            const float2 inv2 = make_float2(inv, inv);
#pragma unroll
            for (int i = 0; i < NPAIR; ++i) acc[i] = __ffma2_rn(tt[i], inv2, acc[i]);
            const float w_wt = tl_wt * inv, w_mut = tl_mut * inv, w_miss = tl_miss * inv;
            acc_zero += w_miss;
            // leaves: the term by the position's code (positions beyond the leaves collect numbers nobody reads)
#pragma unroll
            for (int e = 0; e < 16; ++e) accL[e] += ((m2 >> (2 * e)) & 1u) ? w_mut : (((z2 >> (2 * e)) & 1u) ? w_wt : w_miss);
            buf ^= 1;
        }
        // ---- the CTA's partial sums: sub-warps in fixed order
        float* mine = stage + size_t(sub) * q.stride;
#pragma unroll
        for (int i = 0; i < NPAIR; ++i) {
            mine[(2 * i) * L + sl] = acc[i].x;
            mine[(2 * i + 1) * L + sl] = acc[i].y;
        }
        for (int i = NS; i < IV5_MAX_NS; ++i) mine[i * L + sl] = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) mine[NSL + 16 * sl + e] = accL[e];
        if (sl == 0) mine[NSL + NPOS] = acc_zero;
        __syncthreads();
        float* out = q.part + size_t(task) * q.stride;
        for (int k = t; k < q.stride; k += IV5_THREADS) {
            float s = 0.f;
#pragma unroll
            for (int u = 0; u < NSUB; ++u) s += stage[size_t(u) * q.stride + k];
            out[k] = s;
        }
        __syncthreads();
    }
}

template <int L>
static void iv5_launch_l(int ns, const Iv5Params& q, int grid, size_t smem, cudaStream_t st) {
    if (ns <= 4) scs_place_iv5_kernel<L, 4><<<grid, IV5_THREADS, smem, st>>>(q);
    else if (ns <= 8) scs_place_iv5_kernel<L, 8><<<grid, IV5_THREADS, smem, st>>>(q);
    else if (ns <= 12) scs_place_iv5_kernel<L, 12><<<grid, IV5_THREADS, smem, st>>>(q);
    else if (ns <= 14) scs_place_iv5_kernel<L, 14><<<grid, IV5_THREADS, smem, st>>>(q);
    else scs_place_iv5_kernel<L, 16><<<grid, IV5_THREADS, smem, st>>>(q);
}

// ns: internal-node rows per lane any group of the plan can need (ceil((cells - 2) / L))
static void iv5_launch(int L, int ns, const Iv5Params& q, int grid, size_t smem, cudaStream_t st) {
    if (L == 8) iv5_launch_l<8>(ns, q, grid, smem, st);
    else if (L == 16) iv5_launch_l<16>(ns, q, grid, smem, st);
    else iv5_launch_l<32>(ns, q, grid, smem, st);
}

static inline size_t iv5_smem_bytes(int L, int stride) {
    const int nsub = IV5_THREADS / L, pw = 16 * L + 4;
    return 4 * (size_t(2) * nsub * pw + size_t(32) * L + size_t(nsub) * stride);
}

// Y[g][row] = sum over the group's tasks (fixed order, float64) of the partial sum at the row's slot; NaN for a group whose
// tables could not be built.
__global__ void scs_iv5_reduce_kernel(const float* __restrict__ part, int stride, const int* __restrict__ task0, const int* __restrict__ row_slot,
                                      const int* __restrict__ meta, int n_rows, double* __restrict__ y) {
    const int g = blockIdx.x;
    const int t0 = task0[g], t1 = task0[g + 1];
    const bool bad = meta[g * 4 + 3] != 0;
    for (int r = threadIdx.x; r < n_rows; r += blockDim.x) {
        const int slot = row_slot[size_t(g) * n_rows + r];
        double s = 0.0;
        for (int k = t0; k < t1; ++k) s += double(part[size_t(k) * stride + slot]);
        y[size_t(g) * n_rows + r] = bad ? NAN : s;
    }
}

}  // namespace scs
