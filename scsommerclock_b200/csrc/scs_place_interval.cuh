// scs_place_interval.cuh -- placement without the GEMM, for clade matrices that come from a tree.
//
// The clades of a tree are a laminar family (any two are nested or disjoint), so there is an order
// of the cells -- a depth-first order of the leaves -- in which every clade is ONE interval
// [lo, hi).  With the per-SNV prefix counts P[k] = (mutated, wild type) among the first k cells in
// that order, both counts of a clade are one difference,
//     (C1, C0)[i, b] = P_i[hi_b] - P_i[lo_b],
// i.e. O(cells + clade rows) work per SNV instead of the O(cells x clade rows) contraction
// (run_PT_test.py:421-430 evaluates the dense product; the numbers are the same integers).  At
// 10k cells that is ~10,000x fewer operations than the GEMM.  Nothing but the packed 2-bit
// genotypes (2.5 GB at 10k x 1M, read once) and the per-run partial sums touches HBM; what bounds
// the kernels is instruction issue and shared-memory latency, not the tensor pipe or HBM.
//
// Two kernels walk SNV rows, one CTA per run of rows, accumulators in shared memory (each owned by
// one thread: no atomics), flushed per run to part[run][slot]; a fixed-order fp64 reduction over
// the runs gives the soft weights (short fp32 chains, deterministic):
//   scs_place_interval3_kernel  the main one (trees up to ~14,000 cells): rows rewritten in
//       depth-first cell order first (scs_permute_cells_kernel), word-wise prefix counts, the
//       leaves handled outside the sweeps, per-row count cache in shared memory -- see below;
//   scs_place_interval_kernel   every clade row through two sweeps with a running reference column
//       (same rule and arithmetic as stats_cols32 in scs_place.cu), cells gathered through the
//       cell order, counts recomputed in the second sweep: for trees whose count cache does not
//       fit shared memory.
//
// Whether a clade matrix is laminar is decided from the bit masks themselves (clade_intervals), so
// the C ABI is unchanged; which plans take this path is decided in choose_algo (scs_place.cu).
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

namespace scs {

struct IntervalParams {
    const uint8_t* gt;         // packed 2-bit genotypes [n_snv][pitch]
    long long pitch;
    long long n_snv;
    int n_cells, n_nodes;
    int row_words;             // 32-bit words of a row that hold cells: ceil(n_cells / 16)
    const uint16_t* perm;      // [n_cells] column of the cell at depth-first position k
    const uint32_t* lohi;      // [n_slots] lo | hi << 16 (positions, hi exclusive; lo = hi for an empty row), the clade rows
                               // sorted by (lo, hi) so that neighbouring lanes look up neighbouring prefix counts
    int n_slots;               // n_nodes rounded up to even (slot n_nodes, if any, is padding and carries no weight)
    const uint8_t* row_keep;   // optional [n_snv]
    const GroupConst* gconst;
    float* part;               // [num_runs][n_slots], in sorted clade-row order
    int rows_per_run, num_runs;
    int perm_smem;             // 1: the cell order is staged in shared memory (it fits), 0: read from global / L2
};

constexpr int IV_SCRATCH_WORDS = 5 * 32;

__host__ __device__ inline int iv_p_words(int n_cells) { return (n_cells + 1 + 3) & ~3; }

static inline size_t interval_smem_bytes(int n_cells, int n_nodes, bool dcache, bool perm_smem) {
    const size_t n_slots = size_t(n_nodes + 1) & ~size_t(1);
    return 4 * (size_t(iv_p_words(n_cells)) + n_slots * (dcache ? 2 : 1) + 2 * size_t((n_cells + 15) / 16) + IV_SCRATCH_WORDS +
                (perm_smem ? size_t((n_cells + 1) / 2) : 0));
}

__device__ __forceinline__ void iv_issue_row(const IntervalParams& p, long long r, uint32_t* dst, int t, int T) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.gt + size_t(r) * p.pitch);
    for (int w = t; w < p.row_words; w += T) {
        const uint32_t d = ptx::smem_u32(dst + w);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(src + w) : "memory");
    }
}

// lexicographic "greater" on (rank value, mutated count, wild-type count): a total order, so the
// block-wide reference column does not depend on the reduction shape
__device__ __forceinline__ bool iv_better(float v, float c1, float c0, float bv, float b1, float b0) {
    return v > bv || (v == bv && (c1 > b1 || (c1 == b1 && c0 > b0)));
}

template <bool DCACHE>
__global__ void __launch_bounds__(1024, 1) scs_place_interval_kernel(const IntervalParams p) {
    extern __shared__ uint32_t iv_smem[];
    const int T = blockDim.x, t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarp = T >> 5;
    uint32_t* P = iv_smem;
    float* acc = reinterpret_cast<float*>(P + iv_p_words(p.n_cells));
    uint32_t* dc = reinterpret_cast<uint32_t*>(acc + p.n_slots);
    uint32_t* rowbuf = dc + (DCACHE ? p.n_slots : 0);
    uint32_t* scr = rowbuf + 2 * p.row_words;      // [0,32) scan totals, [32,64) rank values, [64,96) / [96,128) reference counts, [128,160) sums
    const GroupConst gc = p.gconst[0];
    const int E = ((p.n_cells + T - 1) / T) | 1;   // depth-first positions per thread (contiguous; odd: the P[k + 1] stores of a warp spread over all banks)
    const int k0 = min(t * E, p.n_cells), k1 = min(k0 + E, p.n_cells);
    // Nothing row-independent is fetched from global memory inside the row loop: with one CTA per SM in
    // lockstep between barriers, every exposed L2 round trip (~0.5 us) would be a tenth of a row's time.
    const uint16_t* perm = p.perm;
    if (p.perm_smem) {
        uint16_t* ps = reinterpret_cast<uint16_t*>(scr + IV_SCRATCH_WORDS);
        for (int k = t; k < p.n_cells; k += T) ps[k] = p.perm[k];
        perm = ps;                                  // (visible after the first barrier of the row loop)
    }
    // this thread's first two clade rows stay in registers; the rest are fetched two iterations ahead
    const uint32_t lh_first0 = t < p.n_nodes ? p.lohi[t] : 0u, lh_first1 = t + T < p.n_nodes ? p.lohi[t + T] : 0u;

    for (int run = blockIdx.x; run < p.num_runs; run += gridDim.x) {
        for (int b = t; b < p.n_nodes; b += T) acc[b] = 0.f;
        const long long r0 = (long long)run * p.rows_per_run;
        const long long r1 = r0 + p.rows_per_run < p.n_snv ? r0 + p.rows_per_run : p.n_snv;
        int buf = 0;
        iv_issue_row(p, r0, rowbuf, t, T);
        uint32_t keep_next = p.row_keep != nullptr ? p.row_keep[r0] : 1u;
        for (long long r = r0; r < r1; ++r, buf ^= 1) {
            asm volatile("cp.async.wait_all;\n" ::: "memory");
            __syncthreads();      // row r has landed; everybody is done with the previous row's P and scratch
            const uint32_t keep = keep_next;
            if (r + 1 < r1) {
                iv_issue_row(p, r + 1, rowbuf + (buf ^ 1) * p.row_words, t, T);
                if (p.row_keep != nullptr) keep_next = p.row_keep[r + 1];    // consumed one row later
            }
            if (keep == 0u) continue;      // (uniform over the block)

            // ---- 1. prefix counts in depth-first order
            const uint8_t* rb8 = reinterpret_cast<const uint8_t*>(rowbuf + buf * p.row_words);
            uint32_t run_sum = 0;
            for (int k = k0; k < k1; ++k) {
                const uint32_t col = perm[k];
                const uint32_t code = (uint32_t(rb8[col >> 2]) >> ((col & 3) * 2)) & 3u;
                run_sum += code == 3u ? 0u : (code == 0u ? 0x10000u : 1u);
                P[k + 1] = run_sum;
            }
            uint32_t incl = run_sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) scr[warp] = incl;
            __syncthreads();
            const uint32_t wt = lane < nwarp ? scr[lane] : 0u;
            uint32_t winc = wt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += v;
            }
            const uint32_t base = __shfl_sync(0xffffffffu, winc - wt, warp) + incl - run_sum;
            for (int k = k0; k < k1; ++k) P[k + 1] += base;
            if (t == 0) P[0] = 0u;
            __syncthreads();

            // ---- 2. one sweep over the clade rows: running reference column and sum relative to it
            float refv = -INFINITY, r1f = 0.f, r0f = 0.f, sum = 0.f;
            uint32_t lh = lh_first0, lh_n1 = lh_first1;
            for (int b = t; b < p.n_nodes; b += T) {
                const uint32_t lh_n2 = b + 2 * T < p.n_nodes ? p.lohi[b + 2 * T] : 0u;
                const uint32_t d = P[lh >> 16] - P[lh & 0xFFFFu];      // both counts at once (no borrow: P is monotone in each half)
                lh = lh_n1;
                lh_n1 = lh_n2;
                if (DCACHE) dc[b] = d;
                const float c1 = __uint_as_float(__byte_perm(d, 0x4B000000u, 0x7610)) - 8388608.0f;
                const float c0 = __uint_as_float(__byte_perm(d, 0x4B000000u, 0x7632)) - 8388608.0f;
                const float v = fmaf(gc.a1f, c1, gc.a0f * c0);
                if (v > refv + 16.0f) {
                    if (refv != -INFINITY) {
                        const float e1 = r1f - c1, e0 = r0f - c0;
                        sum *= ex2_approx(fmaf(gc.a1h, e1, gc.a0h * e0) + fmaf(gc.a1l, e1, gc.a0l * e0));
                    }
                    r1f = c1;
                    r0f = c0;
                    refv = v;
                }
                const float e1 = c1 - r1f, e0 = c0 - r0f;
                sum += ex2_approx(fmaf(gc.a1h, e1, gc.a0h * e0) + fmaf(gc.a1l, e1, gc.a0l * e0));
            }
            // block-wide reference column
            float bv = refv, b1 = r1f, b0 = r0f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o), o1 = __shfl_xor_sync(0xffffffffu, b1, o),
                            o0 = __shfl_xor_sync(0xffffffffu, b0, o);
                if (iv_better(ov, o1, o0, bv, b1, b0)) {
                    bv = ov;
                    b1 = o1;
                    b0 = o0;
                }
            }
            if (lane == 0) {
                scr[32 + warp] = __float_as_uint(bv);
                scr[64 + warp] = __float_as_uint(b1);
                scr[96 + warp] = __float_as_uint(b0);
            }
            __syncthreads();
            bv = lane < nwarp ? __uint_as_float(scr[32 + lane]) : -INFINITY;
            b1 = lane < nwarp ? __uint_as_float(scr[64 + lane]) : 0.f;
            b0 = lane < nwarp ? __uint_as_float(scr[96 + lane]) : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o), o1 = __shfl_xor_sync(0xffffffffu, b1, o),
                            o0 = __shfl_xor_sync(0xffffffffu, b0, o);
                if (iv_better(ov, o1, o0, bv, b1, b0)) {
                    bv = ov;
                    b1 = o1;
                    b0 = o0;
                }
            }
            // every thread's sum relative to the block's reference (exact differences), then the block sum
            if (refv != -INFINITY) {
                const float e1 = r1f - b1, e0 = r0f - b0;
                sum *= ex2_approx(fmaf(gc.a1h, e1, gc.a0h * e0) + fmaf(gc.a1l, e1, gc.a0l * e0));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0) scr[128 + warp] = __float_as_uint(sum);
            __syncthreads();
            float total = lane < nwarp ? __uint_as_float(scr[128 + lane]) : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
            const float L = log2f(total);        // total >= 1: the reference column itself contributes 2^0

            // ---- 3. second sweep: softmax terms into the per-clade-row accumulators
            for (int b = t; b < p.n_nodes; b += T) {
                uint32_t d;
                if (DCACHE) {
                    d = dc[b];
                } else {
                    const uint32_t lh = p.lohi[b];
                    d = P[lh >> 16] - P[lh & 0xFFFFu];
                }
                const float e1 = (__uint_as_float(__byte_perm(d, 0x4B000000u, 0x7610)) - 8388608.0f) - b1;
                const float e0 = (__uint_as_float(__byte_perm(d, 0x4B000000u, 0x7632)) - 8388608.0f) - b0;
                const float x = fmaf(gc.a1h, e1, gc.a0h * e0) + fmaf(gc.a1l, e1, fmaf(gc.a0l, e0, -L));
                acc[b] += ex2_approx(x);
            }
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        float* out = p.part + size_t(run) * p.n_slots;
        for (int b = t; b < p.n_nodes; b += T) out[b] = acc[b];
    }
}

// ---- the main variant (trees up to ~14,000 cells): the genotype rows are first rewritten in depth-first
// cell order (scs_permute_cells_kernel, one streaming pass), so that a thread owns one 32-bit word = 16
// consecutive positions and all per-position work is bit arithmetic on two masks (mutated / wild type):
//   * prefix counts: block scan over the words' popcounts, then 16 unrolled adds and four 16-byte stores
//     per thread (~4 instructions per position, against ~30 for gathering through the cell order);
//   * single-cell rows (the leaves: half of a tree's clade rows) stay out of the sweeps.  A leaf's counts
//     are its cell's genotype, so a row has only THREE distinct leaf terms (mutated / wild type /
//     missing): they are evaluated once per SNV, their share of the row sum is n_mut t_mut + n_wt t_wt +
//     n_miss t_miss from the leaf totals (popcounts under the leaf mask), and each leaf's accumulator
//     gets its term by two selects on the mask bits (~6 instructions per (SNV, leaf));
//   * the other rows (internal nodes, empty rows, duplicates), sorted by (lo, hi), go through three
//     sweeps over PAIRS of rows in packed fp32x2 arithmetic:
//       sweep 1  counts d = P[hi] - P[lo] -> cache; best column of the row by rank value (no exponentials)
//       sweep 2  t = 2^(logit - logit_best) from exact integer differences -> cache (in place); row sum
//       sweep 3  accumulators += t / sum
// off4[pair] = byte offsets into P (lo0, hi0, lo1, hi1) * 4; leafmask[w] has bit 2e set when position
// 16w + e has a leaf row.
struct Interval3Params {
    IntervalParams b;          // gt = the permuted matrix [n_snv][row_words] (pitch = 4 row_words); n_nodes / n_slots describe
                               // the NON-leaf rows; part rows hold [n_slots non-leaf accumulators, padded to leaf_off][16 row_words leaf accumulators]
    const uint4* off4;         // [n_slots / 2]
    const uint32_t* leafmask;  // [row_words]
    int n_leaf_pos;            // positions that have a leaf row
    int part_stride;           // leaf_off + 16 row_words
    int leaf_off;              // n_slots rounded up to a multiple of 4 (16-byte aligned leaf section)
};

__device__ __forceinline__ float2 iv_lo16(uint32_t a, uint32_t b) {      // (2^23 + low half of a, 2^23 + low half of b)
    return make_float2(__uint_as_float(__byte_perm(a, 0x4B000000u, 0x7610)), __uint_as_float(__byte_perm(b, 0x4B000000u, 0x7610)));
}
__device__ __forceinline__ float2 iv_hi16(uint32_t a, uint32_t b) {
    return make_float2(__uint_as_float(__byte_perm(a, 0x4B000000u, 0x7632)), __uint_as_float(__byte_perm(b, 0x4B000000u, 0x7632)));
}
__device__ __forceinline__ uint32_t iv_lds_off(const uint32_t* base, uint32_t byte_off) {
    return *reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(base) + byte_off);
}

constexpr int IV3_SCRATCH_WORDS = 6 * 32;

// P[0] sits at word 3 so that P[16 w + 1], the first of a thread's sixteen prefix counts, is 16-byte aligned
__host__ __device__ inline int iv3_p_words(int row_words) { return 4 + 16 * row_words; }

static inline size_t interval3_smem_bytes(int n_cells, int n_slots) {
    const size_t rw = size_t((n_cells + 15) / 16);
    // P | non-leaf accumulators | count / term cache | leaf accumulators | scratch
    return 4 * (size_t(iv3_p_words(int(rw))) + 2 * size_t(n_slots) + 16 * rw + IV3_SCRATCH_WORDS);
}

// Rows rewritten in depth-first cell order: out[r][k] = code of cell perm[k] (positions beyond n_cells: 3 = missing).
// One CTA per row at a time, thread = one output word; its sixteen source columns stay in registers.
__global__ void scs_permute_cells_kernel(const uint8_t* __restrict__ gt, long long pitch, long long n_snv, int n_cells, int row_words,
                                         const uint16_t* __restrict__ perm, uint32_t* __restrict__ out) {
    extern __shared__ uint32_t pm_smem[];          // two row buffers
    const int t = threadIdx.x;
    uint32_t cols[8];
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
        const int k = 16 * t + e;
        const uint32_t c0 = (t < row_words && k < n_cells) ? perm[k] : 0xFFFFu, c1 = (t < row_words && k + 1 < n_cells) ? perm[k + 1] : 0xFFFFu;
        cols[e >> 1] = c0 | (c1 << 16);
    }
    int buf = 0;
    // the next row's word is fetched while this row is gathered (nothing between a global load and its use but work)
    uint32_t w_next = (t < row_words && blockIdx.x < n_snv) ? reinterpret_cast<const uint32_t*>(gt + size_t(blockIdx.x) * pitch)[t] : 0u;
    for (long long r = blockIdx.x; r < n_snv; r += gridDim.x, buf ^= 1) {
        uint32_t* rb = pm_smem + buf * row_words;
        if (t < row_words) {
            rb[t] = w_next;
            if (r + gridDim.x < n_snv) w_next = reinterpret_cast<const uint32_t*>(gt + size_t(r + gridDim.x) * pitch)[t];
        }
        __syncthreads();           // (two buffers: the previous row's readers are at most one barrier behind)
        if (t < row_words) {
            const uint8_t* rb8 = reinterpret_cast<const uint8_t*>(rb);
            uint32_t w = 0u;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const uint32_t col = (cols[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu;
                const uint32_t code = col == 0xFFFFu ? 3u : ((uint32_t(rb8[col >> 2]) >> ((col & 3) * 2)) & 3u);
                w |= code << (2 * e);
            }
            out[size_t(r) * row_words + t] = w;
        }
    }
}

// The pre-pass of a plan on the interval path, fused: ONE read of the caller's packed matrix gives (1) the rows in
// depth-first cell order (as scs_permute_cells_kernel), (2) the row filter of get_gt_tree (run_PT_test.py:352: an SNV
// stays when an active -- non-outgroup -- cell carries a call) and (3) the per-cell missing counts over the kept rows
// (:358), which used to be three passes over 2.5 GB.  The row's barrier doubles as the block-wide OR of "has a call";
// missing counts are sixteen byte-wide SWAR counters per thread over its OUTPUT positions (flushed every 255 kept rows),
// added to missing_pos[position] once per CTA; the host maps positions back to cells.
__global__ void scs_permute_reduce_kernel(const uint8_t* __restrict__ gt, long long pitch, long long n_snv, int n_cells, int row_words,
                                          const uint16_t* __restrict__ perm, const uint32_t* __restrict__ active_mask, int filter_rows,
                                          uint32_t* __restrict__ out, uint8_t* __restrict__ row_keep,
                                          unsigned long long* __restrict__ n_kept, unsigned int* __restrict__ missing_pos) {
    extern __shared__ uint32_t pm_smem[];          // two row buffers
    const int t = threadIdx.x;
    const bool own = t < row_words;
    uint32_t cols[8];
    uint32_t valid = 0u;                           // bit 2e: output position 16 t + e is a cell
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
        const int k = 16 * t + e;
        const uint32_t c0 = (own && k < n_cells) ? perm[k] : 0xFFFFu, c1 = (own && k + 1 < n_cells) ? perm[k + 1] : 0xFFFFu;
        cols[e >> 1] = c0 | (c1 << 16);
        if (own && k < n_cells) valid |= 1u << (2 * e);
        if (own && k + 1 < n_cells) valid |= 1u << (2 * e + 2);
    }
    const uint32_t am = own ? active_mask[t] : 0u;
    unsigned int total[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) total[k] = 0u;
    uint32_t a0 = 0u, a1 = 0u, a2 = 0u, a3 = 0u;
    int pending = 0;
    unsigned int kept_cnt = 0u;
    auto flush = [&]() {
#pragma unroll
        for (int b = 0; b < 4; ++b) {              // byte b of accumulator q holds position 4 b + q
            total[4 * b + 0] += (a0 >> (8 * b)) & 0xFFu;
            total[4 * b + 1] += (a1 >> (8 * b)) & 0xFFu;
            total[4 * b + 2] += (a2 >> (8 * b)) & 0xFFu;
            total[4 * b + 3] += (a3 >> (8 * b)) & 0xFFu;
        }
        a0 = a1 = a2 = a3 = 0u;
        pending = 0;
    };
    int buf = 0;
    uint32_t w_next = (own && blockIdx.x < n_snv) ? reinterpret_cast<const uint32_t*>(gt + size_t(blockIdx.x) * pitch)[t] : 0u;
    for (long long r = blockIdx.x; r < n_snv; r += gridDim.x, buf ^= 1) {
        uint32_t* rb = pm_smem + buf * row_words;
        const uint32_t ws = w_next;
        if (own) {
            rb[t] = ws;
            if (r + gridDim.x < n_snv) w_next = reinterpret_cast<const uint32_t*>(gt + size_t(r + gridDim.x) * pitch)[t];
        }
        // a cell holds 1 or 2 exactly when its two bits differ (the mask keeps the low bit of the active cells only)
        const int has_call = (((ws ^ (ws >> 1)) & am) != 0u) ? 1 : 0;
        const int keep = __syncthreads_or(has_call) || !filter_rows;      // also the barrier between filling and gathering the row
        if (own) {
            const uint8_t* rb8 = reinterpret_cast<const uint8_t*>(rb);
            uint32_t w = 0u;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const uint32_t col = (cols[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu;
                const uint32_t code = col == 0xFFFFu ? 3u : ((uint32_t(rb8[col >> 2]) >> ((col & 3) * 2)) & 3u);
                w |= code << (2 * e);
            }
            out[size_t(r) * row_words + t] = w;
            if (keep) {
                const uint32_t m = w & (w >> 1) & valid;          // bit 2e set <=> position e missing
                a0 += m & 0x01010101u;
                a1 += (m >> 2) & 0x01010101u;
                a2 += (m >> 4) & 0x01010101u;
                a3 += (m >> 6) & 0x01010101u;
            }
        }
        if (keep) {                                               // (uniform over the block)
            if (++pending == 255) flush();
            ++kept_cnt;
        }
        if (t == 0) row_keep[r] = uint8_t(keep);
    }
    flush();
    if (own) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (total[k]) atomicAdd(&missing_pos[16 * t + k], total[k]);
    }
    if (t == 0 && kept_cnt) atomicAdd(n_kept, (unsigned long long)kept_cnt);
}

// The same pre-pass, FOUR rows per step.  ncu on the one-row kernel (profiles/r02_ncu_iv4_first_summary.txt): ALU pipe 72 %
// busy, ~6 instructions and one 3.4-way bank-conflicted byte load per (row, cell).  The cell order is the same for every row,
// so the rows of a group are staged byte-interleaved -- shared-memory word b holds source byte b of rows 4g .. 4g+3 (a 4 x 4
// byte transpose in registers, one 16-byte store per thread) -- and ONE 32-bit load then serves a cell of all four rows.
// Round 2: the fields are collected byte-sliced (load + rotate + LOP3 per cell quadruple, then a 4 x 4 byte transpose of four
// accumulators) instead of four funnel shifts + three shifts per cell: ncu had counted 9 instructions per (row, cell), ALU
// pipe 61 % busy, for a kernel that should stream.
// Row flags of the four rows are OR-ed over the block through one shared word per group (three rotate, cleared two groups
// ahead) in front of the group's single barrier.
// Narrow rows (round 2): a block of `nsub` sub-blocks of `pt` threads each (pt = row words rounded up to a warp), every
// sub-block walking its own groups of four rows -- at 1,000 cells a 64-thread block per group left the SMs mostly empty and
// spent its time in per-block set-up and in 1,008 global atomics per block (0.39 ms for 25 MB, 130 GB/s); now 256 threads
// share the barrier, the missing counts of the sub-blocks are combined in shared memory first, and the grid is sized so
// that a sub-block sees at least 32 groups.
constexpr int PERM4_THREADS_MAX = 1024;
constexpr int PERM4_MAX_SUB = 8;
__global__ void scs_permute_reduce4_kernel(const uint8_t* __restrict__ gt, long long pitch, long long n_snv, int n_cells, int row_words,
                                           const uint16_t* __restrict__ perm, const uint32_t* __restrict__ active_mask, int filter_rows,
                                           uint32_t* __restrict__ out, uint8_t* __restrict__ row_keep,
                                           unsigned long long* __restrict__ n_kept, unsigned int* __restrict__ missing_pos, int pt) {
    extern __shared__ uint32_t pm_smem[];          // per sub-block: two buffers of 4 row_words + 4 words (the last four: all ones = "missing" for padding)
    __shared__ uint32_t s_flags[3 * PERM4_MAX_SUB];
    const int nsub = blockDim.x / pt;
    const int t = threadIdx.x % pt, sub = threadIdx.x / pt, lane = threadIdx.x & 31;
    const bool own = t < row_words;
    const int buf_words = 4 * row_words + 4;
    uint32_t* const my_smem = pm_smem + sub * 2 * buf_words;
    uint32_t* const my_flags = s_flags + 3 * sub;
    // per output position 16 t + e: byte address of the staged word that holds its source byte of the four rows, and the
    // rotation that moves its 2-bit field from bit 2 (c % 4) of every byte lane to bit 2 (e % 4)
    uint32_t addr[16], rot[16];
    uint32_t valid = 0u;                           // bit 2e: output position 16 t + e is a cell
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const int k = 16 * t + e;
        if (own && k < n_cells) {
            const uint32_t c = perm[k];
            addr[e] = 4u * (c >> 2);
            rot[e] = uint32_t((2 * int(c & 3u) - 2 * (e & 3)) & 31);
            valid |= 1u << (2 * e);
        } else {
            addr[e] = 4u * uint32_t(4 * row_words);      // the all-ones word: code 3 whatever the rotation
            rot[e] = 0u;
        }
    }
    const uint32_t am = own ? active_mask[t] : 0u;
    unsigned int total[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) total[k] = 0u;
    uint32_t a0 = 0u, a1 = 0u, a2 = 0u, a3 = 0u;
    int pending = 0;
    unsigned int kept_cnt = 0u;
    auto flush = [&]() {
#pragma unroll
        for (int b = 0; b < 4; ++b) {              // byte b of accumulator q holds position 4 b + q
            total[4 * b + 0] += (a0 >> (8 * b)) & 0xFFu;
            total[4 * b + 1] += (a1 >> (8 * b)) & 0xFFu;
            total[4 * b + 2] += (a2 >> (8 * b)) & 0xFFu;
            total[4 * b + 3] += (a3 >> (8 * b)) & 0xFFu;
        }
        a0 = a1 = a2 = a3 = 0u;
        pending = 0;
    };
    if (t < 3) my_flags[t] = 0u;
    if (t < 8) my_smem[(t >> 2) * buf_words + 4 * row_words + (t & 3)] = 0xFFFFFFFFu;
    __syncthreads();
    const long long n_groups = (n_snv + 3) >> 2;
    const long long gstride = (long long)gridDim.x * nsub;
    // the rows of the next TWO groups are in flight while one is worked on: with one 640-thread block per SM (80 registers)
    // a single group ahead left ~10 KB per SM in flight -- 1.4 TB/s by Little's law, which is what the kernel ran at.
    // Addresses are walked, not recomputed (one 64-bit add per group instead of a 64-bit multiply per row), and rows beyond
    // the matrix are handled by a row count per group instead of a 64-bit comparison per row.
    const long long pitch_w = pitch >> 2;                      // (the ABI asks for 4-byte aligned rows)
    uint32_t wn[4], wn2[4];
    auto rows_of = [&](long long g) -> int {                   // valid rows of group g: 4, or fewer at the end, or none
        const long long left = n_snv - 4 * g;
        return left >= 4 ? 4 : (left > 0 ? int(left) : 0);
    };
    auto load_group = [&](const uint32_t* p, int nv, uint32_t (&w)[4]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = (own && i < nv) ? p[size_t(i) * pitch_w] : 0xFFFFFFFFu;      // (rows beyond the matrix read as missing)
    };
    const long long gfirst = (long long)blockIdx.x * nsub + sub;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(gt) + gfirst * 4 * pitch_w + t;
    const long long src_step = gstride * 4 * pitch_w;
    uint32_t* dst = out + gfirst * 4 * row_words + t;
    const long long dst_step = gstride * 4 * row_words;
    load_group(src, rows_of(gfirst), wn);
    load_group(src + src_step, rows_of(gfirst + gstride), wn2);
    src += 2 * src_step;                                       // -> the group two trips ahead
    int buf = 0, fb = 0;
    // every sub-block makes the trips of sub-block 0 (the barrier is block wide); a trip beyond the matrix has no rows
    for (long long g0 = (long long)blockIdx.x * nsub; g0 < n_groups; g0 += gstride, buf ^= 1, fb = fb == 2 ? 0 : fb + 1) {
        const long long g = g0 + sub;
        const int nv = rows_of(g);
        uint32_t* rb = my_smem + buf * buf_words;
        const uint32_t w0 = wn[0], w1 = wn[1], w2 = wn[2], w3 = wn[3];
#pragma unroll
        for (int i = 0; i < 4; ++i) wn[i] = wn2[i];
        load_group(src, rows_of(g + 2 * gstride), wn2);
        src += src_step;
        uint32_t flags = 0u;
        if (own) {
            // 4 x 4 byte transpose: word k of the result holds byte k of the four rows
            const uint32_t l01 = __byte_perm(w0, w1, 0x5140), h01 = __byte_perm(w0, w1, 0x7362);
            const uint32_t l23 = __byte_perm(w2, w3, 0x5140), h23 = __byte_perm(w2, w3, 0x7362);
            reinterpret_cast<uint4*>(rb)[t] = make_uint4(__byte_perm(l01, l23, 0x5410), __byte_perm(l01, l23, 0x7632),
                                                         __byte_perm(h01, h23, 0x5410), __byte_perm(h01, h23, 0x7632));
            // a cell holds 1 or 2 exactly when its two bits differ (the mask keeps the low bit of the active cells only;
            // a row beyond the matrix is all ones: no such cell)
            flags = ((((w0 ^ (w0 >> 1)) & am) != 0u) ? 1u : 0u) | ((((w1 ^ (w1 >> 1)) & am) != 0u) ? 2u : 0u) |
                    ((((w2 ^ (w2 >> 1)) & am) != 0u) ? 4u : 0u) | ((((w3 ^ (w3 >> 1)) & am) != 0u) ? 8u : 0u);
        }
        flags = __reduce_or_sync(0xffffffffu, flags);
        if (lane == 0 && flags) atomicOr(&my_flags[fb], flags);
        __syncthreads();           // the group is staged, its flags are complete
        const uint32_t keep4 = (filter_rows ? my_flags[fb] : 0xFu) & ((1u << nv) - 1u);      // kept AND inside the matrix
        if (t == 0) my_flags[fb == 0 ? 2 : fb - 1] = 0u;      // the word of group g + 2 (last read one group ago, next written after the next barrier)
        if (own) {
            // byte-sliced gather: one load serves a cell of all four rows (byte lane i = row i); a rotation puts the field
            // at bit 2 (e % 4) of every lane (source and target offsets both lie inside the lane, what wraps around is masked
            // off), so accumulator q collects positions 4 q .. 4 q + 3 of the four rows: load + rotate + one LOP3 per cell
            // quadruple.  A 4 x 4 byte transpose then turns the four accumulators into the four output words.
            uint32_t acc4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const uint32_t y = iv_lds_off(rb, addr[e]);
                acc4[e >> 2] |= __funnelshift_r(y, y, rot[e]) & (0x03030303u << (2 * (e & 3)));
            }
            const uint32_t l01 = __byte_perm(acc4[0], acc4[1], 0x5140), h01 = __byte_perm(acc4[0], acc4[1], 0x7362);
            const uint32_t l23 = __byte_perm(acc4[2], acc4[3], 0x5140), h23 = __byte_perm(acc4[2], acc4[3], 0x7362);
            const uint32_t ov[4] = {__byte_perm(l01, l23, 0x5410), __byte_perm(l01, l23, 0x7632), __byte_perm(h01, h23, 0x5410),
                                    __byte_perm(h01, h23, 0x7632)};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i < nv) dst[size_t(i) * row_words] = ov[i];
            // missing cells of the kept rows (bit 2e of m_i: position e of row i is missing).  The four rows are added
            // pairwise while they are still 2-bit fields (0 .. 2 fits), then spread into the byte counters: half the
            // instructions of four rows done one by one
            const uint32_t m0 = (keep4 & 1u) ? (ov[0] & (ov[0] >> 1) & valid) : 0u, m1 = (keep4 & 2u) ? (ov[1] & (ov[1] >> 1) & valid) : 0u;
            const uint32_t m2 = (keep4 & 4u) ? (ov[2] & (ov[2] >> 1) & valid) : 0u, m3 = (keep4 & 8u) ? (ov[3] & (ov[3] >> 1) & valid) : 0u;
            const uint32_t s01 = m0 + m1, s23 = m2 + m3;
            a0 += (s01 & 0x03030303u) + (s23 & 0x03030303u);
            a1 += ((s01 >> 2) & 0x03030303u) + ((s23 >> 2) & 0x03030303u);
            a2 += ((s01 >> 4) & 0x03030303u) + ((s23 >> 4) & 0x03030303u);
            a3 += ((s01 >> 6) & 0x03030303u) + ((s23 >> 6) & 0x03030303u);
        }
        dst += dst_step;
        {
            const int nk = __popc(keep4);                         // (uniform over the sub-block)
            pending += nk;
            kept_cnt += nk;
        }
        if (pending > 250) flush();
        if (t < nv) row_keep[4 * g + t] = uint8_t((keep4 >> t) & 1u);
    }
    flush();
    if (nsub == 1) {
        if (own) {
#pragma unroll
            for (int k = 0; k < 16; ++k)
                if (total[k]) atomicAdd(&missing_pos[16 * t + k], total[k]);
        }
        if (t == 0 && kept_cnt) atomicAdd(n_kept, (unsigned long long)kept_cnt);
        return;
    }
    // several sub-blocks: their counts meet in shared memory (the staging buffers are free now), one global atomic per position
    __syncthreads();
    unsigned int* tot_s = pm_smem;                 // [16 row_words] + 1 (kept rows); nsub >= 2 sub-blocks own >= 16 row_words + 16 words
    for (int k = threadIdx.x; k < 16 * row_words + 1; k += blockDim.x) tot_s[k] = 0u;
    __syncthreads();
    if (own) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (total[k]) atomicAdd(&tot_s[16 * t + k], total[k]);
    }
    if (t == 0 && kept_cnt) atomicAdd(&tot_s[16 * row_words], kept_cnt);
    __syncthreads();
    for (int k = threadIdx.x; k < 16 * row_words; k += blockDim.x)
        if (tot_s[k]) atomicAdd(&missing_pos[k], tot_s[k]);
    if (threadIdx.x == 0 && tot_s[16 * row_words]) atomicAdd(n_kept, (unsigned long long)tot_s[16 * row_words]);
}

// MAXT: largest CTA this instantiation is launched with (register budget); UNR: unroll factor of the three sweeps.
// At 10k cells the CTA has 640 threads (one per word): 96 registers per thread are available, and two pairs of rows in
// flight per thread hide more of the shared-memory / ex2 latency that bounds the sweeps.
template <int MAXT, int UNR>
__global__ void __launch_bounds__(MAXT, 1) scs_place_interval3_kernel(const Interval3Params q) {
    const IntervalParams& p = q.b;
    extern __shared__ uint32_t iv_smem[];
    const int T = blockDim.x, t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarp = T >> 5;
    const int rw = p.row_words;                    // <= T (host)
    uint32_t* P = iv_smem + 3;
    float2* acc2 = reinterpret_cast<float2*>(iv_smem + iv3_p_words(rw));           // 16-byte aligned
    uint2* dc2 = reinterpret_cast<uint2*>(reinterpret_cast<float*>(acc2) + p.n_slots);
    float4* accL4 = reinterpret_cast<float4*>(reinterpret_cast<uint32_t*>(dc2) + p.n_slots);   // n_slots even: 8-byte aligned ...
    uint32_t* scr = reinterpret_cast<uint32_t*>(accL4) + 16 * rw;      // [0,32) scan totals, [32,64) rank values, [64,96) their counts, [128,160) sums, [160,192) leaf totals
    const GroupConst gc = p.gconst[0];
    const bool own = t < rw;                       // this thread owns word t = positions 16 t .. 16 t + 15
    const int n_pairs = p.n_slots >> 1;
    const int j_pad = (p.n_nodes & 1) ? n_pairs - 1 : -1;         // the pair whose second slot is padding
    const uint32_t lmask = own ? q.leafmask[t] : 0u;
    const uint32_t* gw = reinterpret_cast<const uint32_t*>(p.gt);
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    const uint4 of_first0 = t < n_pairs ? q.off4[t] : zero4, of_first1 = t + T < n_pairs ? q.off4[t + T] : zero4;
    const float2 a1f2 = make_float2(gc.a1f, gc.a1f), a0f2 = make_float2(gc.a0f, gc.a0f);
    const float2 a1h2 = make_float2(gc.a1h, gc.a1h), a0h2 = make_float2(gc.a0h, gc.a0h);
    const float2 a1l2 = make_float2(gc.a1l, gc.a1l), a0l2 = make_float2(gc.a0l, gc.a0l);
    const float2 n23 = make_float2(-8388608.0f, -8388608.0f);

    for (int run = blockIdx.x; run < p.num_runs; run += gridDim.x) {
        for (int j = t; j < n_pairs; j += T) acc2[j] = make_float2(0.f, 0.f);
        if (own) {
#pragma unroll
            for (int i = 0; i < 4; ++i) accL4[4 * t + i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const long long r0 = (long long)run * p.rows_per_run;
        const long long r1 = r0 + p.rows_per_run < p.n_snv ? r0 + p.rows_per_run : p.n_snv;
        // the row's word and its keep flag are fetched one row ahead, straight into registers
        uint32_t w_next = own ? gw[size_t(r0) * rw + t] : 0xFFFFFFFFu;
        uint32_t keep_next = p.row_keep != nullptr ? p.row_keep[r0] : 1u;
        for (long long r = r0; r < r1; ++r) {
            const uint32_t w = w_next, keep = keep_next;
            if (r + 1 < r1) {
                if (own) w_next = gw[size_t(r + 1) * rw + t];
                if (p.row_keep != nullptr) keep_next = p.row_keep[r + 1];
            }
            if (keep == 0u) continue;              // (uniform over the block)

            // ---- prefix counts: mutated = exactly one bit of the code set (1 het, 2 hom), wild type = none, 3 = missing
            const uint32_t mm = (w ^ (w >> 1)) & 0x55555555u, wm = ~(w | (w >> 1)) & 0x55555555u;
            const uint32_t tot = uint32_t(__popc(mm)) | (uint32_t(__popc(wm)) << 16);
            uint32_t leaf_sum = uint32_t(__popc(mm & lmask)) | (uint32_t(__popc(wm & lmask)) << 16);
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) leaf_sum += __shfl_xor_sync(0xffffffffu, leaf_sum, o);
            // (no barrier needed here: whoever gets this far has passed the previous row's last barrier, which
            // every thread reaches only after its last read of P, the cache and the scratch words written below)
            if (lane == 31) {
                scr[warp] = incl;
                scr[160 + warp] = leaf_sum;
            }
            __syncthreads();
            const uint32_t wt = lane < nwarp ? scr[lane] : 0u;
            uint32_t winc = wt, ltot = lane < nwarp ? scr[160 + lane] : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += v;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ltot += __shfl_xor_sync(0xffffffffu, ltot, o);
            const uint32_t wbase = __shfl_sync(0xffffffffu, winc - wt, warp);                   // prefix before this warp
            if (own) {
                uint32_t run_sum = wbase + incl - tot;                                          // prefix before this word
                uint32_t pv[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    run_sum += ((mm >> (2 * e)) & 1u) + (((wm >> (2 * e)) & 1u) << 16);
                    pv[e] = run_sum;
                }
                uint4* dst = reinterpret_cast<uint4*>(P + 16 * t + 1);
#pragma unroll
                for (int i = 0; i < 4; ++i) dst[i] = make_uint4(pv[4 * i], pv[4 * i + 1], pv[4 * i + 2], pv[4 * i + 3]);
            }
            if (t == 0) P[0] = 0u;
            __syncthreads();
            const float n_mut = float(ltot & 0xFFFFu), n_wt = float(ltot >> 16), n_miss = float(q.n_leaf_pos) - n_mut - n_wt;

            // ---- sweep 1: counts into the cache, best column by rank value
            float bv = -INFINITY;
            uint32_t bd = 0u;
            {
                uint4 of = of_first0, of_n1 = of_first1;
#pragma unroll(UNR)
                for (int j = t; j < n_pairs; j += T) {
                    const uint4 of_n2 = j + 2 * T < n_pairs ? q.off4[j + 2 * T] : zero4;
                    const uint32_t d0 = iv_lds_off(P, of.y) - iv_lds_off(P, of.x);      // both counts at once (no borrow: P is monotone in each half)
                    const uint32_t d1 = iv_lds_off(P, of.w) - iv_lds_off(P, of.z);
                    of = of_n1;
                    of_n1 = of_n2;
                    dc2[j] = make_uint2(d0, d1);
                    const float2 c1 = __fadd2_rn(iv_lo16(d0, d1), n23), c0 = __fadd2_rn(iv_hi16(d0, d1), n23);
                    float2 v = __ffma2_rn(a1f2, c1, __fmul2_rn(a0f2, c0));
                    if (j == j_pad) v.y = -INFINITY;
                    if (v.x > bv) {          // (first maximum in this thread's fixed order)
                        bv = v.x;
                        bd = d0;
                    }
                    if (v.y > bv) {
                        bv = v.y;
                        bd = d1;
                    }
                }
            }
            // across threads: ties go to the larger packed counts -- a total order, so the result is the same
            // for every reduction shape
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const uint32_t od = __shfl_xor_sync(0xffffffffu, bd, o);
                if (ov > bv || (ov == bv && od > bd)) {
                    bv = ov;
                    bd = od;
                }
            }
            if (lane == 0) {
                scr[32 + warp] = __float_as_uint(bv);
                scr[64 + warp] = bd;
            }
            __syncthreads();
            bv = lane < nwarp ? __uint_as_float(scr[32 + lane]) : -INFINITY;
            bd = lane < nwarp ? scr[64 + lane] : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const uint32_t od = __shfl_xor_sync(0xffffffffu, bd, o);
                if (ov > bv || (ov == bv && od > bd)) {
                    bv = ov;
                    bd = od;
                }
            }
            // the three leaf columns compete too: mutated (C1, C0) = (1, 0), wild type (0, 1), missing (0, 0)
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
            // the three leaf terms and their share of the row sum
            float t_leaf[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float e1 = (c == 1 ? 1.f : 0.f) - rf1, e0 = (c == 0 ? 1.f : 0.f) - rf0;
                t_leaf[c] = ex2_approx(fmaf(gc.a1h, e1, gc.a0h * e0) + fmaf(gc.a1l, e1, gc.a0l * e0));
            }
            const float leaf_total = fmaf(n_mut, t_leaf[1], fmaf(n_wt, t_leaf[0], n_miss * t_leaf[2]));

            // ---- sweep 2: terms relative to the best column, in place; row sum
            float2 s2 = make_float2(0.f, 0.f);
#pragma unroll(UNR)
            for (int j = t; j < n_pairs; j += T) {
                const uint2 d = dc2[j];
                const float2 e1 = __fadd2_rn(iv_lo16(d.x, d.y), nb1_2), e0 = __fadd2_rn(iv_hi16(d.x, d.y), nb0_2);
                const float2 sdiff = __ffma2_rn(a1h2, e1, __fmul2_rn(a0h2, e0));      // exact operands, one rounding
                const float2 lo = __ffma2_rn(a1l2, e1, __fmul2_rn(a0l2, e0));
                const float2 x = __fadd2_rn(sdiff, lo);
                float2 tt = make_float2(ex2_approx(x.x), ex2_approx(x.y));
                if (j == j_pad) tt.y = 0.f;
                s2 = __fadd2_rn(s2, tt);
                dc2[j] = make_uint2(__float_as_uint(tt.x), __float_as_uint(tt.y));
            }
            float sum = s2.x + s2.y;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0) scr[128 + warp] = __float_as_uint(sum);
            __syncthreads();
            float total = lane < nwarp ? __uint_as_float(scr[128 + lane]) : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
            const float inv = 1.0f / (total + leaf_total);        // >= 1: the best column's own term
            const float2 inv2 = make_float2(inv, inv);

            // ---- sweep 3: normalised terms into the accumulators (each owned by one thread)
#pragma unroll(UNR)
            for (int j = t; j < n_pairs; j += T) {
                const uint2 tb = dc2[j];
                acc2[j] = __ffma2_rn(make_float2(__uint_as_float(tb.x), __uint_as_float(tb.y)), inv2, acc2[j]);
            }
            if (own) {
                // leaves: the term by the position's code (positions without a leaf row collect numbers nobody reads)
                const float w_wt = t_leaf[0] * inv, w_mut = t_leaf[1] * inv, w_miss = t_leaf[2] * inv;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 a = accL4[4 * t + i];
                    float* av = reinterpret_cast<float*>(&a);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int e = 4 * i + k;
                        av[k] += ((mm >> (2 * e)) & 1u) ? w_mut : (((wm >> (2 * e)) & 1u) ? w_wt : w_miss);
                    }
                    accL4[4 * t + i] = a;
                }
            }
        }
        float* out = p.part + size_t(run) * q.part_stride;
        for (int j = t; j < n_pairs; j += T) reinterpret_cast<float2*>(out)[j] = acc2[j];
        if (own) {
#pragma unroll
            for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(out + q.leaf_off)[4 * t + i] = accL4[4 * t + i];
        }
    }
}

// y[row of slot s] = sum over the runs of part[run][s], fixed order, fp64 (slots without a row -- padding,
// positions without a leaf row -- are skipped)
constexpr int IV_REDUCE_SLICES = 8;
__global__ void scs_reduce_interval_kernel(const float* __restrict__ part, int num_runs, int n_entries, int stride,
                                           const int* __restrict__ row_of_slot, double* __restrict__ y) {
    // block = 32 entries x 8 slices of the runs; a slice sums runs s, s + 8, ... (4 independent chains), the slices are
    // combined in fixed order: deterministic, and thousands of runs are no longer one thread's sequential walk
    __shared__ double sm[IV_REDUCE_SLICES][32];
    const int e = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int b = blockIdx.x * 32 + e;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (b < n_entries) {
        int r = sl;
        for (; r + 3 * IV_REDUCE_SLICES < num_runs; r += 4 * IV_REDUCE_SLICES) {
            a0 += double(part[size_t(r) * stride + b]);
            a1 += double(part[size_t(r + IV_REDUCE_SLICES) * stride + b]);
            a2 += double(part[size_t(r + 2 * IV_REDUCE_SLICES) * stride + b]);
            a3 += double(part[size_t(r + 3 * IV_REDUCE_SLICES) * stride + b]);
        }
        for (; r < num_runs; r += IV_REDUCE_SLICES) a0 += double(part[size_t(r) * stride + b]);
    }
    sm[sl][e] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (sl == 0 && b < n_entries) {
        const int row = row_of_slot[b];
        if (row >= 0) {
            double s = sm[0][e];
#pragma unroll
            for (int k = 1; k < IV_REDUCE_SLICES; ++k) s += sm[k][e];
            y[row] = s;
        }
    }
}

// ---- host: is the clade matrix a laminar family, and in which cell order is every row an interval?
// Rows sorted by decreasing size; a cell's signature is its membership bit vector over those rows;
// cells sorted by signature (descending).  For a laminar family every row is then contiguous: two
// members of a row agree on all larger rows (each of which contains the row or misses it entirely),
// so anything sorted between them shares that prefix and, having a 0 where they have a 1, would
// sort after both.  The check afterwards (hi - lo == size) makes the decision exact either way.
static bool clade_intervals(const uint32_t* bits, int64_t pitch_words, int n_rows, int n_cells, std::vector<uint16_t>& perm,
                            std::vector<uint32_t>& lohi) {
    const int words = (n_cells + 31) / 32;
    std::vector<int> cnt(size_t(n_rows), 0);
    for (int r = 0; r < n_rows; ++r) {
        const uint32_t* row = bits + size_t(r) * pitch_words;
        int c = 0;
        for (int w = 0; w < words; ++w) {
            uint32_t v = row[w];
            if (w == words - 1 && (n_cells & 31)) v &= (1u << (n_cells & 31)) - 1u;
            c += __builtin_popcount(v);
        }
        cnt[size_t(r)] = c;
    }
    std::vector<int> order;
    for (int r = 0; r < n_rows; ++r)
        if (cnt[size_t(r)] > 0) order.push_back(r);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cnt[size_t(a)] > cnt[size_t(b)]; });
    const size_t sw = (order.size() + 63) / 64;
    std::vector<uint64_t> sig(size_t(n_cells) * std::max<size_t>(sw, 1), 0);
    for (size_t ri = 0; ri < order.size(); ++ri) {
        const uint32_t* row = bits + size_t(order[ri]) * pitch_words;
        const uint64_t bit = uint64_t(1) << (63 - (ri & 63));
        for (int w = 0; w < words; ++w) {
            uint32_t v = row[w];
            if (w == words - 1 && (n_cells & 31)) v &= (1u << (n_cells & 31)) - 1u;
            while (v) {
                const int c = w * 32 + __builtin_ctz(v);
                v &= v - 1;
                sig[size_t(c) * sw + (ri >> 6)] |= bit;
            }
        }
    }
    std::vector<int> cells(size_t(n_cells), 0);
    std::iota(cells.begin(), cells.end(), 0);
    if (sw > 0)
        std::stable_sort(cells.begin(), cells.end(), [&](int a, int b) {
            const uint64_t *sa = &sig[size_t(a) * sw], *sb = &sig[size_t(b) * sw];
            for (size_t w = 0; w < sw; ++w)
                if (sa[w] != sb[w]) return sa[w] > sb[w];
            return false;
        });
    std::vector<int> pos(size_t(n_cells), 0);
    perm.resize(size_t(n_cells));
    for (int k = 0; k < n_cells; ++k) {
        perm[size_t(k)] = uint16_t(cells[size_t(k)]);
        pos[size_t(cells[size_t(k)])] = k;
    }
    lohi.assign(size_t(n_rows), 0u);
    for (int r = 0; r < n_rows; ++r) {
        if (cnt[size_t(r)] == 0) continue;
        const uint32_t* row = bits + size_t(r) * pitch_words;
        int lo = n_cells, hi = 0;
        for (int w = 0; w < words; ++w) {
            uint32_t v = row[w];
            if (w == words - 1 && (n_cells & 31)) v &= (1u << (n_cells & 31)) - 1u;
            while (v) {
                const int c = w * 32 + __builtin_ctz(v);
                v &= v - 1;
                lo = std::min(lo, pos[size_t(c)]);
                hi = std::max(hi, pos[size_t(c)] + 1);
            }
        }
        if (hi - lo != cnt[size_t(r)]) return false;      // not an interval: the family is not laminar
        lohi[size_t(r)] = uint32_t(lo) | (uint32_t(hi) << 16);
    }
    return true;
}

}  // namespace scs
