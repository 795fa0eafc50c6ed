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
//   ~10^4 log2 units at 10,000 cells and the logits are differences of them);  three sweeps over the clade rows (sorted by
//   interval start, four rows in flight per thread): the largest logit, the normaliser sum of 2^(d - max), and 2^(d - max) / sum
//   onto the row's accumulator -- the logit differences are re-formed from the float64 prefix sums in every sweep (two 64-bit
//   shared-memory loads and an add: cheaper than keeping 20 values per thread alive at 1024 threads).
//   Accumulators: fp32 in shared memory, flushed to a per-CTA float64 vector every 256 SNVs (short fp32 chains, fixed order);
//   a last kernel adds the per-CTA vectors in CTA order.  The softmax terms are ex2.approx of float64 differences.
//   Trees of 1,024 .. 12,287 cells take scs_place_percell2_kernel<R> below: the same arithmetic with every thread's clade
//   rows in registers (accumulators and terms) and the interval table in shared memory.
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
    const uint32_t* lohi;      // [n_rows] lo | hi << 16 (hi exclusive; lo = hi: empty row, logit 0), the clade rows SORTED by (lo, hi):
                               // slot -> caller's row in the reduction kernel's slot_row
    const double2* w;          // [n_cells] by depth-first position: {log2((1-FN_c)/FP), log2(FN_c/(1-FP))}; {0,0} for cells in no clade
    const uint8_t* row_keep;   // optional [n_snv]
    double* part;              // [gridDim.x][n_rows] by slot, zeroed by the caller
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

// One SNV row: its bytes -> shared memory, the weights of this thread's contiguous block of depth-first positions summed
// up, a block-wide scan -> P[1..n_cells] (P[0] = 0 was written once).  Ends with a barrier: P is complete for every thread.
__device__ __forceinline__ void pc_prefix_sums(const PercellParams& p, long long r, double* P, double* scr, uint32_t* rowbuf,
                                               const uint16_t* perm, int t, int T, int lane, int warp, int nwarp, int k0, int k1) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.gt + size_t(r) * p.pitch);
    for (int w = t; w < p.row_words; w += T) rowbuf[w] = src[w];
    __syncthreads();
    double run = 0.0;
    const double* wd = reinterpret_cast<const double*>(p.w);
    // (the weight loads come from L2 -- 80 KB per SNV, more than the L1 left beside the shared memory -- and are the longest
    // latency of the phase: eight positions' loads in flight per round trip)
#pragma unroll 8
    for (int k = k0; k < k1; ++k) {
        const int c = perm[k];
        const uint32_t code = (rowbuf[c >> 4] >> (2 * (c & 15))) & 3u;
        const double wk = __ldg(wd + 2 * k + (code == 0u ? 1 : 0));     // {observed mutated, observed wild type}
        run += code == 3u ? 0.0 : wk;
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
}

// block-wide maximum / sum of one double per thread (scr: 32 doubles; one barrier; the same bits in every thread)
__device__ __forceinline__ double pc_block_max(double m, double* scr, int lane, int warp, int nwarp) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const double v = pc_shfl_xor(m, o);
        m = v > m ? v : m;
    }
    if (lane == 0) scr[warp] = m;
    __syncthreads();
    m = lane < nwarp ? scr[lane] : -1.0e300;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const double v = pc_shfl_xor(m, o);
        m = v > m ? v : m;
    }
    return m;
}
__device__ __forceinline__ double pc_block_sum(double s, double* scr, int lane, int warp, int nwarp) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += pc_shfl_xor(s, o);
    if (lane == 0) scr[warp] = s;
    __syncthreads();
    s = lane < nwarp ? scr[lane] : 0.0;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += pc_shfl_xor(s, o);            // (a butterfly of commutative adds: the same bits in every lane and warp)
    return s;
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
    const int E = ((p.n_cells + T - 1) / T) | 1;                   // depth-first positions per thread (contiguous; odd: the 64-bit P[k + 1] accesses of a warp then spread over all banks)
    const int k0 = min(t * E, p.n_cells), k1 = min(k0 + E, p.n_cells);
    if (t == 0) P[0] = 0.0;
    __syncthreads();
    int since_flush = 0;
    for (long long r = blockIdx.x; r < p.n_snv; r += gridDim.x) {
        if (p.row_keep && !p.row_keep[r]) continue;                // (block-uniform)
        pc_prefix_sums(p, r, P, scr, rowbuf, perm, t, T, lane, warp, nwarp, k0, k1);
        // ---- sweep 1: the largest logit (four rows in flight per thread; the rows are sorted by interval start, so the lanes of
        // a warp read neighbouring prefix sums)
        double m = -1.0e300;
        for (int b0 = t; b0 < p.n_rows; b0 += 4 * T) {
            uint32_t lh[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) lh[u] = b0 + u * T < p.n_rows ? p.lohi[b0 + u * T] : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double d = P[lh[u] >> 16] - P[lh[u] & 0xffffu];
                if (b0 + u * T < p.n_rows && d > m) m = d;         // (no NaNs here: a plain compare, not the library maximum's NaN rules)
            }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const double v = pc_shfl_xor(m, o);
            m = v > m ? v : m;
        }
        if (lane == 0) scr[warp] = m;
        __syncthreads();
        m = lane < nwarp ? scr[lane] : -1.0e300;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const double v = pc_shfl_xor(m, o);
            m = v > m ? v : m;
        }
        // ---- sweep 2: the normaliser, sum of 2^(d - max)
        float sf = 0.f;
        for (int b0 = t; b0 < p.n_rows; b0 += 4 * T) {
            uint32_t lh[4];
            float e[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) lh[u] = b0 + u * T < p.n_rows ? p.lohi[b0 + u * T] : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) e[u] = ex2_approx(float(P[lh[u] >> 16] - P[lh[u] & 0xffffu] - m));
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (b0 + u * T < p.n_rows) sf += e[u];
        }
        double s = double(sf);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) s += pc_shfl_xor(s, o);
        if (lane == 0) scr[32 + warp] = s;
        __syncthreads();
        s = lane < nwarp ? scr[32 + lane] : 0.0;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) s += pc_shfl_xor(s, o);            // (a butterfly of commutative adds: the same bits in every lane and warp)
        const float inv = float(1.0 / s);
        // ---- sweep 3: the normalised terms onto the accumulators (thread t owns slots t, t + T, ...)
        if (p.acc_smem) {
            for (int b0 = t; b0 < p.n_rows; b0 += 4 * T) {
                uint32_t lh[4];
                float e[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) lh[u] = b0 + u * T < p.n_rows ? p.lohi[b0 + u * T] : 0u;
#pragma unroll
                for (int u = 0; u < 4; ++u) e[u] = ex2_approx(float(P[lh[u] >> 16] - P[lh[u] & 0xffffu] - m)) * inv;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (b0 + u * T < p.n_rows) acc[b0 + u * T] += e[u];
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
                part[b] += double(ex2_approx(float(d - m)) * inv);
            }
        }
        __syncthreads();                                           // P, rowbuf and scr are rewritten by the next row
    }
    if (p.acc_smem && since_flush > 0)
        for (int b = t; b < p.n_rows; b += T) part[b] += double(acc[b]);
}

// ---- trees of 1,024 .. 12,287 cells: 512 threads, every thread's clade rows (slots t, t + 512, ...: R of them) live in
// REGISTERS -- the accumulators across SNVs and, within an SNV, the terms 2^(d - max) between the sweep that sums them and the
// update that needs their sum -- and the interval table sits in shared memory where the accumulators were.  Two sweeps over
// the prefix sums instead of three, no table reads from L2 inside the SNV loop, one FFMA per (SNV, row) for the update.
constexpr int PC2_THREADS = 512;
constexpr int PC2_MAX_R = 48;

static inline size_t percell2_smem_bytes(int n_cells, int R) {
    return size_t(n_cells + 2) / 2 * 2 * sizeof(double) + 64 * sizeof(double) + size_t((n_cells + 15) / 16) * sizeof(uint32_t) +
           size_t(R) * PC2_THREADS * sizeof(uint32_t) + size_t(n_cells + 1) / 2 * 2 * sizeof(uint16_t) + 16;
}

template <int R>
__global__ void __launch_bounds__(PC2_THREADS, 1) scs_place_percell2_kernel(const PercellParams p) {
    extern __shared__ double pc_smem[];
    constexpr int T = PC2_THREADS, nwarp = T / 32;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    double* P = pc_smem;                                           // [n_cells + 1]
    double* scr = P + size_t(p.n_cells + 2) / 2 * 2;               // [64]
    uint32_t* rowbuf = reinterpret_cast<uint32_t*>(scr + 64);      // [row_words]
    uint32_t* lohi_s = rowbuf + p.row_words;                       // [R * T], zero beyond n_rows
    uint16_t* perm_s = reinterpret_cast<uint16_t*>(lohi_s + R * T);
    for (int k = t; k < p.n_cells; k += T) perm_s[k] = p.perm[k];
    for (int b = t; b < R * T; b += T) lohi_s[b] = b < p.n_rows ? p.lohi[b] : 0u;
    const int nvalid = t < p.n_rows ? (p.n_rows - t + T - 1) / T : 0;      // this thread's slots: t + j T, j < nvalid
    float acc[R];
#pragma unroll
    for (int j = 0; j < R; ++j) acc[j] = 0.f;
    double* part = p.part + size_t(blockIdx.x) * p.n_rows;
    const int E = ((p.n_cells + T - 1) / T) | 1;                   // (odd: see the general kernel)
    const int k0 = min(t * E, p.n_cells), k1 = min(k0 + E, p.n_cells);
    if (t == 0) P[0] = 0.0;
    __syncthreads();
    int since_flush = 0;
    for (long long r = blockIdx.x; r < p.n_snv; r += gridDim.x) {
        if (p.row_keep && !p.row_keep[r]) continue;                // (block-uniform)
        pc_prefix_sums(p, r, P, scr, rowbuf, perm_s, t, T, lane, warp, nwarp, k0, k1);
        // ---- sweep 1: the largest logit
        double m = -1.0e300;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const uint32_t lh = lohi_s[t + j * T];
            const double d = P[lh >> 16] - P[lh & 0xffffu];
            if (j < nvalid && d > m) m = d;
        }
        m = pc_block_max(m, scr, lane, warp, nwarp);
        // ---- sweep 2: the terms 2^(d - max), kept, and their sum
        float e[R];
        float sf = 0.f;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const uint32_t lh = lohi_s[t + j * T];
            const float x = ex2_approx(float(P[lh >> 16] - P[lh & 0xffffu] - m));
            e[j] = j < nvalid ? x : 0.f;
            sf += e[j];
        }
        const double s = pc_block_sum(double(sf), scr + 32, lane, warp, nwarp);
        const float inv = float(1.0 / s);
#pragma unroll
        for (int j = 0; j < R; ++j) acc[j] = fmaf(e[j], inv, acc[j]);
        if (++since_flush == PC_FLUSH_ROWS) {
#pragma unroll
            for (int j = 0; j < R; ++j) {
                if (j < nvalid) part[t + j * T] += double(acc[j]);
                acc[j] = 0.f;
            }
            since_flush = 0;
        }
        __syncthreads();                                           // P, rowbuf and scr are rewritten by the next row
    }
    if (since_flush > 0) {
#pragma unroll
        for (int j = 0; j < R; ++j)
            if (j < nvalid) part[t + j * T] += double(acc[j]);
    }
}

// y[row of slot b] = sum over the CTAs' vectors, in CTA order
__global__ void scs_percell_reduce_kernel(const double* part, int n_parts, int n_rows, const int* slot_row, double* y) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_rows) return;
    double s = 0.0;
    for (int g = 0; g < n_parts; ++g) s += part[size_t(g) * n_rows + b];
    y[slot_row[b]] = s;
}

}  // namespace scs
