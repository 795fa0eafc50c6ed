// scs_place_interval.cuh -- placement without the GEMM, for clade matrices that come from a tree.
//
// The clades of a tree are a laminar family (any two are nested or disjoint), so there is an order
// of the cells -- a depth-first order of the leaves -- in which every clade is ONE interval
// [lo, hi).  With the per-SNV prefix counts P[k] = (mutated, wild type) among the first k cells in
// that order, both counts of a clade are one difference,
//     (C1, C0)[i, b] = P_i[hi_b] - P_i[lo_b],
// i.e. O(cells + clade rows) work per SNV instead of the O(cells x clade rows) contraction
// (run_PT_test.py:421-430 evaluates the dense product; the numbers are the same integers).  At
// 10k cells that is ~10,000x fewer operations than the GEMM, and the kernel is bound by instruction
// issue (about 40 instructions per (SNV, clade row): two shared-memory lookups, the exact-difference
// exponent of the GEMM epilogue, ex2, accumulate), not by the tensor pipe or HBM: the packed 2-bit
// genotypes (2.5 GB at 10k x 1M) are read exactly once, nothing else touches HBM.
//
// One CTA walks SNV rows; per row it
//   1. gathers the row's 2-bit codes in depth-first order and scans them into P (packed u32:
//      mutated | wild type << 16, so one subtraction yields both counts),
//   2. sweeps the clade rows once with a running reference column (same rule and the same
//      arithmetic on exact integer differences as stats_cols32 in scs_place.cu), merges the
//      per-thread (reference, sum) pairs over the block -> reference column and log2-sum of the row,
//   3. sweeps the clade rows again and adds 2^(logit - logit_ref - log2sum) to per-clade-row
//      accumulators in shared memory (each owned by one thread: no atomics).
// Accumulators are flushed per run of rows to part[run][row]; a fixed-order fp64 reduction over the
// runs gives the soft weights (short fp32 chains, deterministic).
//
// Whether a clade matrix is laminar is decided from the bit masks themselves (clade_intervals), so
// the C ABI is unchanged: any caller that passes tree-shaped masks gets this path, anything else
// goes through the GEMM.
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
    const uint32_t* lohi;      // [n_nodes] lo | hi << 16 (positions, hi exclusive; lo = hi for an empty row)
    const uint8_t* row_keep;   // optional [n_snv]
    const GroupConst* gconst;
    float* part;               // [num_runs][n_nodes]
    int rows_per_run, num_runs;
    int perm_smem;             // 1: the cell order is staged in shared memory (it fits), 0: read from global / L2
};

constexpr int IV_SCRATCH_WORDS = 5 * 32;

__host__ __device__ inline int iv_p_words(int n_cells) { return (n_cells + 1 + 3) & ~3; }

static inline size_t interval_smem_bytes(int n_cells, int n_nodes, bool dcache, bool perm_smem) {
    return 4 * (size_t(iv_p_words(n_cells)) + size_t(n_nodes) * (dcache ? 2 : 1) + 2 * size_t((n_cells + 15) / 16) + IV_SCRATCH_WORDS +
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
    uint32_t* dc = reinterpret_cast<uint32_t*>(acc + p.n_nodes);
    uint32_t* rowbuf = dc + (DCACHE ? p.n_nodes : 0);
    uint32_t* scr = rowbuf + 2 * p.row_words;      // [0,32) scan totals, [32,64) rank values, [64,96) / [96,128) reference counts, [128,160) sums
    const GroupConst gc = p.gconst[0];
    const int E = (p.n_cells + T - 1) / T;         // depth-first positions per thread (contiguous)
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
        float* out = p.part + size_t(run) * p.n_nodes;
        for (int b = t; b < p.n_nodes; b += T) out[b] = acc[b];
    }
}

// y[b] = sum over the runs of part[run][b], fixed order, fp64
__global__ void scs_reduce_interval_kernel(const float* __restrict__ part, int num_runs, int n_nodes, double* __restrict__ y) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_nodes) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int r = 0;
    for (; r + 3 < num_runs; r += 4) {
        a0 += double(part[size_t(r) * n_nodes + b]);
        a1 += double(part[size_t(r + 1) * n_nodes + b]);
        a2 += double(part[size_t(r + 2) * n_nodes + b]);
        a3 += double(part[size_t(r + 3) * n_nodes + b]);
    }
    for (; r < num_runs; ++r) a0 += double(part[size_t(r) * n_nodes + b]);
    y[b] = (a0 + a1) + (a2 + a3);
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
