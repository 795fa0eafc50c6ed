// scs_place_percell.cuh -- map_mutations_gt with ONE FALSE-NEGATIVE RATE PER CELL: the pd.Series branch of
// run_PT_test.py:413-418 (TP_m = S * log(1 - FN_i), FN_m = S * log(FN_i), FP scalar).
//
// With per-cell rates the logit of (SNV i, clade row b) is no longer a function of two integer counts:
//
//     logit[i, b] - const(i) = sum over cells c of clade b:  x1[i, c] * log((1 - FN_c) / FP)  +  x0[i, c] * log(FN_c / (1 - FP))
//
// (x1 / x0: cell observed mutated / wild type; every term that does not depend on b cancels in the normalisation
// :368-383).  For tree-shaped clade masks -- every call the reference makes -- the sum over a clade is still a DIFFERENCE
// OF PREFIX SUMS in the depth-first cell order of clade_intervals (scs_place_interval.cuh), now of per-cell weights instead
// of 0/1 indicators, so the work per SNV stays O(cells + clade rows):
//
//   per SNV row, one CTA:  row bytes -> shared memory (coalesced);  every thread sums the weights of its contiguous block of
//   depth-first positions, a block-wide scan turns them into prefix sums P[0..n] (float64 in shared memory: the sums reach
//   ~10^4 log2 units at 10,000 cells and the logits are differences of them);  sweep 1 over the clade rows keeps a running
//   (max, sum of 2^(d - max)) per thread, combined over the block;  sweep 2 adds 2^(d - max) / sum to the row's accumulator.
//   Accumulators: fp32 in shared memory, flushed to a per-CTA float64 vector every 256 SNVs (short fp32 chains, fixed order);
//   a last kernel adds the per-CTA vectors in CTA order.  The softmax terms are exp2f of float64 differences.
//
// This is the general-weights sibling of the count kernels, built for completeness of map_mutations_gt's interface (nothing on
// the reference's command-line path passes per-cell rates); it shares their row order, reduction discipline and C ABI style.
#pragma once

#include <cstdint>

namespace scs {

struct PercellParams {
    const uint8_t* gt;         // packed 2-bit genotypes [n_snv][pitch]
    long long pitch;
    long long n_snv;
    int n_cells, n_rows;
    int row_words;             // ceil(n_cells / 16)
    const uint16_t* perm;      // [n_cells] column of the cell at depth-first position k
    const uint32_t* lohi;      // [n_rows] lo | hi << 16 (hi exclusive; lo = hi: empty row, logit 0)
    const double2* w;          // [n_cells] by depth-first position: {log2((1-FN_c)/FP), log2(FN_c/(1-FP))}; {0,0} for cells in no clade
    const uint8_t* row_keep;   // optional [n_snv]
    double* part;              // [gridDim.x][n_rows], zeroed by the caller
    int acc_smem;              // 1: fp32 accumulators in shared memory, 0: straight into `part` (trees too big for that)
    int perm_smem;             // 1: the cell order staged in shared memory
};

constexpr int PC_FLUSH_ROWS = 256;

static inline size_t percell_smem_bytes(int n_cells, int n_rows, bool acc_smem, bool perm_smem) {
    size_t b = size_t(n_cells + 2) / 2 * 2 * sizeof(double);      // P
    b += 2 * 32 * sizeof(double);                                   // scan totals / (max, sum) pairs
    b += size_t((n_cells + 15) / 16) * sizeof(uint32_t);            // the row
    if (acc_smem) b += size_t(n_rows) * sizeof(float);
    if (perm_smem) b += size_t(n_cells + 1) / 2 * 2 * sizeof(uint16_t);
    return b + 16;
}

__device__ __forceinline__ double pc_shfl_up(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double pc_shfl_xor(double v, int d) { return __shfl_xor_sync(0xffffffffu, v, d); }

// (m, s) <- softmax state of the union: s counts 2^(d - m)
__device__ __forceinline__ void pc_combine(double& m, double& s, double m2, double s2) {
    if (m2 > m) {
        s = s * double(exp2f(float(m - m2))) + s2;
        m = m2;
    } else {
        s += s2 * double(exp2f(float(m2 - m)));
    }
}

__global__ void __launch_bounds__(1024, 1) scs_place_percell_kernel(const PercellParams p) {
    extern __shared__ double pc_smem[];
    const int T = blockDim.x, t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarp = T >> 5;
    double* P = pc_smem;                                           // [n_cells + 1]
    double* scr = P + size_t(p.n_cells + 2) / 2 * 2;               // [64]
    uint32_t* rowbuf = reinterpret_cast<uint32_t*>(scr + 64);      // [row_words]
    float* acc = reinterpret_cast<float*>(rowbuf + p.row_words);   // [n_rows] when acc_smem
    uint16_t* perm_s = reinterpret_cast<uint16_t*>(acc + (p.acc_smem ? p.n_rows : 0));
    const uint16_t* perm = p.perm;
    if (p.perm_smem) {
        for (int k = t; k < p.n_cells; k += T) perm_s[k] = p.perm[k];
        perm = perm_s;
    }
    if (p.acc_smem)
        for (int b = t; b < p.n_rows; b += T) acc[b] = 0.f;
    double* part = p.part + size_t(blockIdx.x) * p.n_rows;
    const int E = (p.n_cells + T - 1) / T;                         // depth-first positions per thread (contiguous)
    const int k0 = min(t * E, p.n_cells), k1 = min(k0 + E, p.n_cells);
    if (t == 0) P[0] = 0.0;
    __syncthreads();
    int since_flush = 0;
    for (long long r = blockIdx.x; r < p.n_snv; r += gridDim.x) {
        if (p.row_keep && !p.row_keep[r]) continue;                // (block-uniform)
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.gt + size_t(r) * p.pitch);
        for (int w = t; w < p.row_words; w += T) rowbuf[w] = src[w];
        __syncthreads();
        // ---- prefix sums of the per-cell weights in depth-first order
        double run = 0.0;
        for (int k = k0; k < k1; ++k) {
            const int c = perm[k];
            const uint32_t code = (rowbuf[c >> 4] >> (2 * (c & 15))) & 3u;
            const double2 wk = p.w[k];
            run += code == 0u ? wk.y : (code == 3u ? 0.0 : wk.x);
            P[k + 1] = run;
        }
        double incl = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double v = pc_shfl_up(incl, d);
            if (lane >= d) incl += v;
        }
        if (lane == 31) scr[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            double v = lane < nwarp ? scr[lane] : 0.0, x = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double y = pc_shfl_up(x, d);
                if (lane >= d) x += y;
            }
            scr[32 + lane] = x - v;                                // exclusive offset of every warp
        }
        __syncthreads();
        const double off = scr[32 + warp] + (incl - run);
        for (int k = k0; k < k1; ++k) P[k + 1] += off;
        __syncthreads();
        // ---- sweep 1: max and sum of 2^(d - max) over the clade rows
        double m = 0.0, s = 0.0;                                   // (the all-zero 'FP' row guarantees a logit of 0)
        bool any = false;
        for (int b = t; b < p.n_rows; b += T) {
            const uint32_t lh = p.lohi[b];
            const double d = P[lh >> 16] - P[lh & 0xffffu];
            if (!any) {
                m = d;
                s = 1.0;
                any = true;
            } else if (d > m) {
                s = s * double(exp2f(float(m - d))) + 1.0;
                m = d;
            } else {
                s += double(exp2f(float(d - m)));
            }
        }
        if (!any) {
            m = -1.0e300;
            s = 0.0;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const double m2 = pc_shfl_xor(m, d), s2 = pc_shfl_xor(s, d);
            pc_combine(m, s, m2, s2);
        }
        if (lane == 0) {
            scr[warp] = m;
            scr[32 + warp] = s;
        }
        __syncthreads();
        m = scr[0];
        s = scr[32];
        for (int wv = 1; wv < nwarp; ++wv) pc_combine(m, s, scr[wv], scr[32 + wv]);      // same order in every thread
        const float inv = float(1.0 / s);
        // ---- sweep 2: the normalised terms onto the accumulators (thread t owns rows t, t + T, ...)
        if (p.acc_smem) {
            for (int b = t; b < p.n_rows; b += T) {
                const uint32_t lh = p.lohi[b];
                const double d = P[lh >> 16] - P[lh & 0xffffu];
                acc[b] += exp2f(float(d - m)) * inv;
            }
            if (++since_flush == PC_FLUSH_ROWS) {
                for (int b = t; b < p.n_rows; b += T) {
                    part[b] += double(acc[b]);
                    acc[b] = 0.f;
                }
                since_flush = 0;
            }
        } else {
            for (int b = t; b < p.n_rows; b += T) {
                const uint32_t lh = p.lohi[b];
                const double d = P[lh >> 16] - P[lh & 0xffffu];
                part[b] += double(exp2f(float(d - m)) * inv);
            }
        }
        __syncthreads();                                           // P, rowbuf and scr are rewritten by the next row
    }
    if (p.acc_smem && since_flush > 0)
        for (int b = t; b < p.n_rows; b += T) part[b] += double(acc[b]);
}

// y[b] = sum over the CTAs' vectors, in CTA order
__global__ void scs_percell_reduce_kernel(const double* part, int n_parts, int n_rows, double* y) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_rows) return;
    double s = 0.0;
    for (int g = 0; g < n_parts; ++g) s += part[size_t(g) * n_rows + b];
    y[b] = s;
}

}  // namespace scs
