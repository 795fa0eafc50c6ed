// scs_decode.cu -- VCF genotype decoding into the packed SNV x cell observation
// matrix (reference: run_PT_test.py:27-207, get_mut_df with filter=True) and the
// row/column reductions get_gt_tree applies to it (run_PT_test.py:350-358).
//
// The host hands over the decompressed VCF text and the sample columns to keep
// (the #CHROM line, renaming and the -excl/-incl regexes are header work and stay
// on the host).  One CTA decodes one record line:
//   * thread 0 walks the nine fixed columns (CHROM POS ID REF ALT QUAL FILTER INFO
//     FORMAT) and derives the GT / TG sub-field indices from FORMAT (:87-93);
//   * all threads count tabs in their slice of the line, a block scan turns that
//     into sample ordinals, and whoever owns the tab in front of a sample parses
//     its GT (and TG) sub-field (:89-119);
//   * the per-line decisions of :147-181 (skip lines without any call, ISA
//     violations, majority alt, FILTER) are taken from shared-memory tallies;
//   * the kept samples are packed four to a byte.
// Kept lines are then compacted in file order.
#include <cstdio>
#include <vector>

#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/reduce.h>
#include <thrust/scan.h>

#include "scs_internal.h"

namespace scs {

constexpr int DEC_THREADS = 128;

// per-line status
enum : uint8_t { LINE_HEADER = 0, LINE_NOALT = 1, LINE_FILTERED = 2, LINE_KEPT = 3, LINE_ERROR = 4 };
// parse error detail (first one wins)
enum : int { PERR_NONE = 0, PERR_COLUMNS = 1, PERR_GT = 2, PERR_SUBFIELD = 3, PERR_CHROM = 4, PERR_REFLAST = 5, PERR_NOGT = 6 };

struct DecodeParams {
    const char* text;
    int64_t nbytes;
    const int64_t* line_start;  // [n_lines + 1]
    int64_t n_lines;
    int n_samples;              // sample columns in the file
    const int* incl_idx;        // [n_incl] file column of every kept sample
    int n_incl;
    int pitch;                  // bytes per packed row
    uint8_t* tmp;               // [n_lines][n_samples] kind | alt << 2
    uint8_t* packed_all;        // [n_lines][pitch]
    uint8_t* status;            // [n_lines]
    int64_t* pos_all;           // [n_lines][2]
    int* err;                   // {code, line}
};

__device__ __forceinline__ bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

__global__ void scs_mark_newlines_kernel(const char* __restrict__ text, int64_t n, unsigned long long* __restrict__ counts) {
    // counts[b] = number of '\n' in block b's slice (exclusive-scanned afterwards)
    const int64_t per = 4096;
    const int64_t s = int64_t(blockIdx.x) * per;
    int c = 0;
    for (int64_t i = s + threadIdx.x; i < min(n, s + per); i += blockDim.x) c += text[i] == '\n';
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    __shared__ int sh[32];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < int(blockDim.x >> 5); ++i) t += sh[i];
        counts[blockIdx.x] = t;
    }
}

__global__ void scs_fill_line_starts_kernel(const char* __restrict__ text, int64_t n, const unsigned long long* __restrict__ offs,
                                            int64_t* __restrict__ line_start) {
    // one thread walks each 4 KiB slice: line k+1 starts after the k-th newline
    const int64_t b = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t per = 4096;
    const int64_t s = b * per;
    if (s >= n) return;
    unsigned long long k = offs[b];
    for (int64_t i = s; i < min(n, s + per); ++i)
        if (text[i] == '\n') line_start[++k] = i + 1;
}

__global__ void __launch_bounds__(DEC_THREADS) scs_decode_lines_kernel(const DecodeParams p) {
    __shared__ int s_tab[9];
    __shared__ int s_cnt[DEC_THREADS];
    __shared__ int s_hist[10];
    __shared__ int s_sum_het, s_sum_hom, s_any_true, s_err;
    __shared__ int s_gt_idx, s_tg_idx, s_filter_ok, s_ab_alt, s_mode;
    const int tid = threadIdx.x;
    for (int64_t line = blockIdx.x; line < p.n_lines; line += gridDim.x) {
        int64_t s = p.line_start[line], e = p.line_start[line + 1];
        if (e > s && p.text[e - 1] == '\n') --e;
        while (e > s && is_space(p.text[e - 1])) --e;      // line.strip() (:82)
        while (s < e && is_space(p.text[s])) ++s;
        __syncthreads();
        if (tid < 10) s_hist[tid] = 0;
        if (tid == 0) {
            s_sum_het = s_sum_hom = s_any_true = 0;
            s_err = PERR_NONE;
            s_gt_idx = s_tg_idx = -1;
            s_filter_ok = 0;
        }
        if (e == s || p.text[s] == '#') {                   // :47, :78
            if (tid == 0) p.status[line] = LINE_HEADER;
            continue;
        }
        const char* t = p.text + s;
        const int len = int(e - s);
        // ---- fixed columns (thread 0)
        if (tid == 0) {
            int nt = 0;
            for (int i = 0; i < len && nt < 9; ++i)
                if (t[i] == '\t') s_tab[nt++] = i;
            if (nt < 9) {
                s_err = PERR_COLUMNS;
            } else {
                // CHROM: int(chrom.replace('chr', ''))  (:172)
                long long chrom = 0, pos = 0;
                bool ok = true, any = false;
                for (int i = 0; i < s_tab[0];) {
                    if (i + 2 < s_tab[0] && t[i] == 'c' && t[i + 1] == 'h' && t[i + 2] == 'r') {
                        i += 3;
                        continue;
                    }
                    if (t[i] >= '0' && t[i] <= '9') {
                        chrom = chrom * 10 + (t[i] - '0');
                        any = true;
                    } else {
                        ok = false;
                    }
                    ++i;
                }
                bool anyp = false;
                for (int i = s_tab[0] + 1; i < s_tab[1]; ++i) {
                    if (t[i] >= '0' && t[i] <= '9') {
                        pos = pos * 10 + (t[i] - '0');
                        anyp = true;
                    } else {
                        ok = false;
                    }
                }
                if (!ok || !any || !anyp) s_err = PERR_CHROM;
                p.pos_all[line * 2] = chrom;
                p.pos_all[line * 2 + 1] = pos;
                // FILTER (:175): contains "PASS", or equals "s50" / "singleton"
                const int f0 = s_tab[5] + 1, f1 = s_tab[6];
                const int fl = f1 - f0;
                bool pass = false;
                for (int i = f0; i + 4 <= f1; ++i)
                    if (t[i] == 'P' && t[i + 1] == 'A' && t[i + 2] == 'S' && t[i + 3] == 'S') pass = true;
                if (fl == 3 && t[f0] == 's' && t[f0 + 1] == '5' && t[f0 + 2] == '0') pass = true;
                if (fl == 9) {
                    const char* sg = "singleton";
                    bool eq = true;
                    for (int i = 0; i < 9; ++i) eq &= t[f0 + i] == sg[i];
                    pass |= eq;
                }
                s_filter_ok = pass;
                // FORMAT (:87): index of the GT and TG tokens
                int idx = 0, tok = s_tab[7] + 1;
                for (int i = s_tab[7] + 1; i <= s_tab[8]; ++i) {
                    if (i == s_tab[8] || t[i] == ':') {
                        if (i - tok == 2 && t[tok] == 'G' && t[tok + 1] == 'T' && s_gt_idx < 0) s_gt_idx = idx;
                        if (i - tok == 2 && t[tok] == 'T' && t[tok + 1] == 'G' && s_tg_idx < 0) s_tg_idx = idx;
                        ++idx;
                        tok = i + 1;
                    }
                }
                if (s_gt_idx < 0) s_err = PERR_NOGT;
            }
        }
        __syncthreads();
        if (s_err != PERR_NONE) {
            if (tid == 0) {
                p.status[line] = LINE_ERROR;
                if (atomicCAS(&p.err[0], 0, s_err) == 0) p.err[1] = int(line);
            }
            continue;
        }
        // ---- tab ordinals: slice the line, count, scan
        const int body0 = s_tab[8];                          // the tab in front of sample 0
        const int blen = len - body0;
        const int per = (blen + DEC_THREADS - 1) / DEC_THREADS;
        const int c0 = body0 + tid * per, c1 = min(len, c0 + per);
        int ntab = 0;
        for (int i = c0; i < c1; ++i) ntab += t[i] == '\t';
        s_cnt[tid] = ntab;
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int i = 0; i < DEC_THREADS; ++i) {
                const int v = s_cnt[i];
                s_cnt[i] = run;
                run += v;
            }
            if (run != p.n_samples) s_err = PERR_COLUMNS;
        }
        __syncthreads();
        if (s_err != PERR_NONE) {
            if (tid == 0) {
                p.status[line] = LINE_ERROR;
                if (atomicCAS(&p.err[0], 0, s_err) == 0) p.err[1] = int(line);
            }
            continue;
        }
        // ---- per-sample GT / TG (:89-119)
        const int gt_idx = s_gt_idx, tg_idx = s_tg_idx;
        int smp = s_cnt[tid];
        int l_het = 0, l_hom = 0, l_true = 0;
        for (int i = c0; i < c1; ++i) {
            if (t[i] != '\t') continue;
            // record of sample `smp` starts at i+1
            int q = i + 1, sub = 0, g0 = -1, g1 = -1, tg0 = -1, tg1 = -1;
            int fs = q;
            for (;; ++q) {
                const bool end = q >= len || t[q] == '\t';
                if (end || t[q] == ':') {
                    if (sub == gt_idx) { g0 = fs; g1 = q; }
                    if (sub == tg_idx) { tg0 = fs; tg1 = q; }
                    ++sub;
                    fs = q + 1;
                    if (end) break;
                }
            }
            uint8_t rec = 0;
            if (g0 < 0 || (tg_idx >= 0 && tg0 < 0)) {
                atomicCAS(&s_err, PERR_NONE, PERR_SUBFIELD);
            } else {
                const int gl = g1 - g0;
                auto gc = [&](int k) { char c = t[g0 + k]; return c == '/' ? '|' : c; };
                bool has0 = false;
                for (int k = 0; k < gl; ++k) has0 |= t[g0 + k] == '0';
                const bool miss = gl == 3 && gc(0) == '.' && gc(1) == '|' && gc(2) == '.';
                const bool wt = gl == 3 && gc(0) == '0' && gc(1) == '|' && gc(2) == '0';
                if (miss) {
                    rec = 3;
                } else if (wt) {
                    rec = 0;
                } else {
                    const char last = gl > 0 ? t[g0 + gl - 1] : 'x';
                    if (last < '0' || last > '9') {
                        atomicCAS(&s_err, PERR_NONE, PERR_GT);    // int(gt[-1]) raises in the reference
                    } else {
                        const int alt = last - '0';
                        rec = uint8_t((has0 ? 1 : 2) | (alt << 2));
                        atomicAdd(&s_hist[alt], 1);
                        if (has0) l_het += alt; else l_hom += alt;
                    }
                }
                if (tg_idx >= 0) {                                 // :100-104
                    const int tl = tg1 - tg0;
                    int zeros = 0;
                    for (int k = 0; k < tl; ++k) zeros += t[tg0 + k] == '0';
                    if (zeros == 1) {
                        l_true = 1;
                    } else if (tl == 3 && t[tg0 + 1] == '|' && (t[tg0] == '1' || t[tg0] == '2') && (t[tg0 + 2] == '1' || t[tg0 + 2] == '2')) {
                        l_true = 1;
                    }
                }
            }
            p.tmp[line * p.n_samples + smp] = rec;
            ++smp;
        }
        if (l_het) atomicAdd(&s_sum_het, l_het);
        if (l_hom) atomicAdd(&s_sum_hom, l_hom);
        if (l_true) atomicOr(&s_any_true, 1);
        __syncthreads();
        // ---- line-level decisions (:147-181)
        if (tid == 0 && s_err == PERR_NONE) {
            if (!(s_sum_het > 0 || s_sum_hom > 0 || s_any_true)) {
                s_mode = -1;                                      // no call at all: skipped (:147)
            } else {
                int n_alts = 0, first = -1, best = -1, bestc = 0;
                bool all_one = true;
                for (int a = 0; a < 10; ++a)
                    if (s_hist[a] > 0) {
                        ++n_alts;
                        if (first < 0) first = a;
                        if (s_hist[a] != 1) all_one = false;
                        if (s_hist[a] > bestc) { bestc = s_hist[a]; best = a; }   // np.argmax: first maximum
                    }
                if (n_alts == 0) {
                    s_mode = 0;                                   // nothing called: zeros / NaN (:155)
                } else if (n_alts == 2 && all_one) {
                    s_mode = 1;                                   // ISA violation (:159)
                } else {
                    s_mode = 2;                                   // majority alt (:163)
                    s_ab_alt = best;
                    if (best == 0) s_err = PERR_REFLAST;
                }
            }
        }
        __syncthreads();
        if (s_err != PERR_NONE) {
            if (tid == 0) {
                p.status[line] = LINE_ERROR;
                if (atomicCAS(&p.err[0], 0, s_err) == 0) p.err[1] = int(line);
            }
            continue;
        }
        const int mode = s_mode;
        if (mode < 0) {
            if (tid == 0) p.status[line] = LINE_NOALT;
            continue;
        }
        const bool kept = s_filter_ok && mode != 1;
        if (tid == 0) p.status[line] = kept ? LINE_KEPT : LINE_FILTERED;
        if (!kept) continue;
        // ---- pack the kept samples, four per byte
        const int ab = s_ab_alt;
        for (int b = tid; b < p.pitch; b += DEC_THREADS) {
            uint32_t byte = 0;
            for (int k = 0; k < 4; ++k) {
                const int oc = b * 4 + k;
                if (oc >= p.n_incl) break;
                const uint8_t rec = p.tmp[line * p.n_samples + p.incl_idx[oc]];
                const int kind = rec & 3, alt = rec >> 2;
                uint32_t code;
                if (kind == 3) code = 3;
                else if (mode == 2 && kind != 0 && alt == ab) code = uint32_t(kind);   // het -> 1, hom -> 2
                else code = 0;
                byte |= code << (2 * k);
            }
            p.packed_all[line * p.pitch + b] = uint8_t(byte);
        }
    }
}

__global__ void scs_status_flags_kernel(const uint8_t* __restrict__ status, int64_t n, long long* __restrict__ flag) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = status[i] == LINE_KEPT;
}

__global__ void scs_compact_rows_kernel(const uint8_t* __restrict__ status, const long long* __restrict__ dest, int64_t n_lines,
                                        const uint8_t* __restrict__ packed_all, const int64_t* __restrict__ pos_all, int pitch,
                                        uint8_t* __restrict__ packed, int64_t* __restrict__ pos) {
    for (int64_t line = blockIdx.x; line < n_lines; line += gridDim.x) {
        if (status[line] != LINE_KEPT) continue;
        const long long d = dest[line];
        for (int b = threadIdx.x; b < pitch; b += blockDim.x) packed[d * pitch + b] = packed_all[line * pitch + b];
        if (threadIdx.x == 0) {
            pos[d * 2] = pos_all[line * 2];
            pos[d * 2 + 1] = pos_all[line * 2 + 1];
        }
    }
}

__global__ void scs_count_status_kernel(const uint8_t* __restrict__ status, int64_t n, unsigned long long* __restrict__ counts) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&counts[status[i]], 1ull);
}

// get_gt_tree's reductions (run_PT_test.py:350-358) on the packed matrix:
// row_keep[i] = any active cell of SNV i carries code 1 or 2 (sum(axis=1) > 0 with
// the outgroup column dropped); missing[c] = number of kept SNVs where cell c is NaN.
__global__ void scs_row_keep_kernel(const uint8_t* __restrict__ gt, int64_t pitch, int64_t n_rows, int n_words, int lanes_per_row,
                                    const uint32_t* __restrict__ active_mask, int filter_rows, uint8_t* __restrict__ row_keep,
                                    unsigned long long* __restrict__ n_kept) {
    // `lanes_per_row` (a power of two <= 32) lanes share a row, 16 cells (one 32-bit word of
    // codes) per lane and step: a full warp for wide matrices, several rows per warp for narrow
    // ones.  A cell holds 1 or 2 exactly when its two code bits differ; active_mask has bit 2k of
    // word j set when cell 16j+k takes part in the filter.
    const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int shift = 31 - __clz(lanes_per_row);
    const int64_t row = t >> shift;
    const int sub = int(t) & (lanes_per_row - 1);
    uint32_t any = 0;
    if (row < n_rows) {   // (no early return: every thread reaches the block-wide count below)
        const uint32_t* r = reinterpret_cast<const uint32_t*>(gt + row * pitch);
        for (int j = sub; j < n_words; j += lanes_per_row) {
            const uint32_t v = r[j];
            any |= ((v ^ (v >> 1)) & 0x55555555u) & active_mask[j];
        }
    }
    for (int o = lanes_per_row >> 1; o > 0; o >>= 1) any |= __shfl_xor_sync(0xffffffffu, any, o);
    const bool kept = row < n_rows && sub == 0 && (!filter_rows || any != 0);
    if (row < n_rows && sub == 0) row_keep[row] = uint8_t(kept);
    const int votes = __syncthreads_count(kept);                                            // one atomic per block
    if (threadIdx.x == 0 && votes) atomicAdd(n_kept, (unsigned long long)votes);             // integer: order-independent
}

constexpr int MISS_REPL = 8;   // copies of the per-cell counters (spreads same-address atomics), summed on the host

constexpr int MISS_SLABS = 8;   // row slabs per block of the wide missing-count kernel

__global__ void __launch_bounds__(64 * MISS_SLABS)
scs_missing_count_kernel(const uint8_t* __restrict__ gt, int64_t pitch, int64_t n_rows, int n_cols, int n_words,
                         const uint8_t* __restrict__ row_keep, unsigned int* __restrict__ missing) {
    // Wide matrices: a thread owns one 32-bit word column (16 cells) over a slab of rows; loads are
    // coalesced across the warp and eight rows are in flight per thread (a dropped row's word is
    // loaded and masked rather than skipped, which keeps the loads independent).  "Missing"
    // (code 3) is both bits set.  The 16 counters are four words of byte-wide SWAR counters,
    // flushed to 32-bit registers before they can wrap.  Block = 64 word columns x MISS_SLABS row
    // slabs, combined in shared memory: one atomic per cell and block (atomic throughput, not HBM,
    // bounded the first version), on one of MISS_REPL counter copies.
    constexpr int UNROLL = 8;
    __shared__ unsigned int part[MISS_SLABS - 1][64 * 16];
    const int j = blockIdx.x * 64 + threadIdx.x;
    const int64_t n_slabs = int64_t(gridDim.y) * MISS_SLABS;
    const int64_t slab = (n_rows + n_slabs - 1) / n_slabs;
    const int64_t r0 = (int64_t(blockIdx.y) * MISS_SLABS + threadIdx.y) * slab, r1 = min(n_rows, r0 + slab);
    unsigned int total[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) total[k] = 0;
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int pending = 0;
    auto flush = [&]() {
#pragma unroll
        for (int b = 0; b < 4; ++b) {      // byte b of accumulator q holds cell 4*b + q
            total[4 * b + 0] += (a0 >> (8 * b)) & 0xFFu;
            total[4 * b + 1] += (a1 >> (8 * b)) & 0xFFu;
            total[4 * b + 2] += (a2 >> (8 * b)) & 0xFFu;
            total[4 * b + 3] += (a3 >> (8 * b)) & 0xFFu;
        }
        a0 = a1 = a2 = a3 = 0;
        pending = 0;
    };
    auto add = [&](uint32_t v, bool keep) {
        const uint32_t m = keep ? (v & (v >> 1)) : 0u;    // bit 2k set <=> cell k missing
        a0 += m & 0x01010101u;
        a1 += (m >> 2) & 0x01010101u;
        a2 += (m >> 4) & 0x01010101u;
        a3 += (m >> 6) & 0x01010101u;
    };
    if (j < n_words) {
        int64_t r = r0;
        for (; r + UNROLL <= r1; r += UNROLL) {      // full batches: plain loads, all issued before the first use
            uint32_t v[UNROLL];
            uint8_t keep[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                v[u] = __ldg(reinterpret_cast<const uint32_t*>(gt + (r + u) * pitch) + j);
                keep[u] = __ldg(row_keep + r + u);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) add(v[u], keep[u] != 0);
            pending += UNROLL;
            if (pending > 255 - UNROLL) flush();
        }
        for (; r < r1; ++r) {                         // (< UNROLL rows: cannot overflow the byte counters)
            add(__ldg(reinterpret_cast<const uint32_t*>(gt + r * pitch) + j), row_keep[r] != 0);
        }
        flush();
    }
    if (threadIdx.y > 0) {
#pragma unroll
        for (int k = 0; k < 16; ++k) part[threadIdx.y - 1][k * 64 + threadIdx.x] = total[k];
    }
    __syncthreads();
    if (threadIdx.y == 0 && j < n_words) {
        unsigned int* dst = missing + size_t(blockIdx.y % MISS_REPL) * n_cols;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            unsigned int t = total[k];
#pragma unroll
            for (int y = 0; y < MISS_SLABS - 1; ++y) t += part[y][k * 64 + threadIdx.x];
            if (j * 16 + k < n_cols && t) atomicAdd(&dst[j * 16 + k], t);
        }
    }
}

// Narrow matrices (<= 512 cells, replicate batches): row filter AND missing counts in one pass over the
// matrix.  2^lane_shift lanes share a row, one 32-bit word (16 cells) each; the lanes of a row OR their
// "mutated" bits with shuffles, lane 0 of the row writes row_keep, and every lane adds the missing
// (code 3) bits of a kept row to sixteen byte-wide SWAR counters (flushed before they can wrap).  Eight
// independent rows per lane are in flight per iteration; loop bounds are warp-uniform so that every
// lane reaches the shuffles.
__global__ void __launch_bounds__(256, 4)
scs_gt_reduce_narrow_kernel(const uint8_t* __restrict__ gt, int64_t pitch, int64_t n_rows, int n_words, int lane_shift, int n_cols,
                            const uint32_t* __restrict__ active_mask, int filter_rows, uint8_t* __restrict__ row_keep,
                            unsigned long long* __restrict__ n_kept, unsigned int* __restrict__ missing,
                            const long long* __restrict__ group_row0 = nullptr) {
    constexpr int UNROLL = 8;
    if (group_row0 != nullptr) {
        // a batch of datasets (replicates), one per blockIdx.y: rows [group_row0[g], group_row0[g + 1]) with their own counters
        const long long g0 = group_row0[blockIdx.y];
        n_rows = group_row0[blockIdx.y + 1] - g0;
        gt += g0 * pitch;
        row_keep += g0;
        n_kept += blockIdx.y;
        missing += size_t(blockIdx.y) * n_cols;
    }
    __shared__ unsigned int s_miss[512];
    __shared__ unsigned int s_kept;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_miss[i] = 0;
    if (threadIdx.x == 0) s_kept = 0;
    __syncthreads();
    const int lanes = 1 << lane_shift;
    const int sub = threadIdx.x & (lanes - 1);
    const int64_t gtid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t nslots = (int64_t(gridDim.x) * blockDim.x) >> lane_shift;
    const int64_t wslot = (gtid & ~int64_t(31)) >> lane_shift;       // first row slot of this warp
    const int64_t my = (gtid >> lane_shift) - wslot;                 // this lane's slot inside the warp
    const bool has = sub < n_words;
    const uint32_t am = has ? active_mask[sub] : 0u;
    unsigned int total[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) total[k] = 0;
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int pending = 0;
    unsigned int kept_cnt = 0;
    auto flush = [&]() {
#pragma unroll
        for (int b = 0; b < 4; ++b) {      // byte b of accumulator q holds cell 4*b + q
            total[4 * b + 0] += (a0 >> (8 * b)) & 0xFFu;
            total[4 * b + 1] += (a1 >> (8 * b)) & 0xFFu;
            total[4 * b + 2] += (a2 >> (8 * b)) & 0xFFu;
            total[4 * b + 3] += (a3 >> (8 * b)) & 0xFFu;
        }
        a0 = a1 = a2 = a3 = 0;
        pending = 0;
    };
    const int subc = min(sub, n_words - 1);                            // lanes beyond the row re-read its last word, masked below
    const int rows_per_warp = 32 >> lane_shift;
    for (int64_t base = wslot; base < n_rows; base += nslots * UNROLL) {
        uint32_t v[UNROLL];
        if (base + rows_per_warp - 1 + (UNROLL - 1) * nslots < n_rows) {   // (warp-uniform) whole batch inside: plain loads
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                v[u] = __ldg(reinterpret_cast<const uint32_t*>(gt + (base + my + u * nslots) * pitch) + subc);
        } else {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int64_t r = base + my + u * nslots;
                v[u] = r < n_rows ? __ldg(reinterpret_cast<const uint32_t*>(gt + r * pitch) + subc) : 0u;
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t r = base + my + u * nslots;
            const uint32_t w = has ? v[u] : 0u;
            uint32_t any = ((w ^ (w >> 1)) & 0x55555555u) & am;      // a cell holds 1 or 2 exactly when its two bits differ
            for (int o = lanes >> 1; o > 0; o >>= 1) any |= __shfl_xor_sync(0xffffffffu, any, o);
            const bool kept = r < n_rows && (!filter_rows || any != 0);
            if (sub == 0 && r < n_rows) {
                row_keep[r] = uint8_t(kept);
                kept_cnt += kept;
            }
            const uint32_t m = kept ? (w & (w >> 1)) : 0u;           // bit 2k set <=> cell k missing
            a0 += m & 0x01010101u;
            a1 += (m >> 2) & 0x01010101u;
            a2 += (m >> 4) & 0x01010101u;
            a3 += (m >> 6) & 0x01010101u;
        }
        pending += UNROLL;
        if (pending > 255 - UNROLL) flush();
    }
    flush();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        unsigned int c = total[k];
        for (int o = 16; o >= lanes; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if ((threadIdx.x & 31) < lanes && has && c) atomicAdd(&s_miss[sub * 16 + k], c);
    }
    for (int o = 16; o > 0; o >>= 1) kept_cnt += __shfl_xor_sync(0xffffffffu, kept_cnt, o);
    if ((threadIdx.x & 31) == 0 && kept_cnt) atomicAdd(&s_kept, kept_cnt);
    __syncthreads();
    for (int i = threadIdx.x; i < n_words * 16; i += blockDim.x)
        if (i < n_cols && s_miss[i]) atomicAdd(&missing[i], s_miss[i]);
    if (threadIdx.x == 0 && s_kept) atomicAdd(n_kept, (unsigned long long)s_kept);     // integers: order-independent
}

// ---- row filter + per-cell missing counts, split into an asynchronous launch and a fetch so that the placement
// plan (scs_placement_prepare, scs_place.cu) can run them on its own counters without a host round trip per call.
// scratch layout: [kept counter (8 B, padded to 16)] [active mask words] [missing counters x MISS_REPL]
size_t gt_reduce_scratch_bytes(int n_cells) {
    const int n_words = (n_cells + 15) / 16;
    return 16 + size_t(n_words + 3) / 4 * 16 + size_t(n_cells) * 4 * MISS_REPL;
}

int gt_reduce_launch(scs_ctx* ctx, const uint8_t* d_gt, int64_t pitch_bytes, int64_t n_snv, int n_cells, const uint8_t* col_active,
                     int filter_rows, uint8_t* d_row_keep, uint8_t* sc, bool accumulate, cudaStream_t st) {
    const int n_words = (n_cells + 15) / 16;
    const size_t off_mask = 16, off_miss = off_mask + size_t(n_words + 3) / 4 * 16;
    const size_t bytes = gt_reduce_scratch_bytes(n_cells);
    unsigned long long* d_kept = reinterpret_cast<unsigned long long*>(sc);
    uint32_t* d_act = reinterpret_cast<uint32_t*>(sc + off_mask);
    unsigned int* d_miss = reinterpret_cast<unsigned int*>(sc + off_miss);
    if (!accumulate) {
        // active cells as a mask over the low bit of every 2-bit code
        std::vector<uint32_t> amask(n_words, 0u);
        for (int c = 0; c < n_cells; ++c)
            if (col_active[c]) amask[c >> 4] |= 1u << (2 * (c & 15));
        SCS_CUDA(cudaMemsetAsync(sc, 0, bytes, st));
        {
            const int urc = upload_small(ctx, d_act, amask.data(), size_t(n_words) * 4, st);      // (not the copy engine: scs_internal.h)
            if (urc != SCS_OK) return urc;
        }
    }
    if (n_snv > 0) {
        if (n_words <= 32) {
            int shift = 0;
            while ((1 << shift) < n_words) ++shift;
            const int64_t want = (n_snv * (int64_t(1) << shift) + 255) / 256;
            const unsigned grid = unsigned(std::min<int64_t>(std::max<int64_t>(want, 1), int64_t(ctx->sm_count) * 4));
            scs_gt_reduce_narrow_kernel<<<grid, 256, 0, st>>>(d_gt, pitch_bytes, n_snv, n_words, shift, n_cells, d_act, filter_rows, d_row_keep, d_kept, d_miss);
        } else {
            scs_row_keep_kernel<<<unsigned((n_snv * 32 + 255) / 256), 256, 0, st>>>(d_gt, pitch_bytes, n_snv, n_words, 32, d_act, filter_rows, d_row_keep, d_kept);
            // 32..256 rows per thread: at least ~2 blocks per SM on small inputs, few atomics on big ones
            const unsigned gx = unsigned((n_words + 63) / 64);
            const int64_t per_thread = std::max<int64_t>(32, std::min<int64_t>(256, n_snv / (int64_t(ctx->sm_count) * 2 * MISS_SLABS)));
            const int64_t gy = std::max<int64_t>(1, (n_snv + per_thread * MISS_SLABS - 1) / (per_thread * MISS_SLABS));
            dim3 grid(gx, unsigned(std::min<int64_t>(gy, 65535))), block(64, MISS_SLABS);
            scs_missing_count_kernel<<<grid, block, 0, st>>>(d_gt, pitch_bytes, n_snv, n_cells, n_words, d_row_keep, d_miss);
        }
    }
    SCS_CUDA(cudaGetLastError());
    return SCS_OK;
}

int gt_reduce_fetch(const uint8_t* sc, int n_cells, int64_t* n_kept, double* missing_frac, cudaStream_t st) {
    const int n_words = (n_cells + 15) / 16;
    const size_t off_miss = 16 + size_t(n_words + 3) / 4 * 16;
    std::vector<unsigned int> miss(size_t(n_cells) * MISS_REPL);
    unsigned long long kept = 0;
    SCS_CUDA(cudaMemcpyAsync(&kept, sc, sizeof(kept), cudaMemcpyDeviceToHost, st));
    SCS_CUDA(cudaMemcpyAsync(miss.data(), sc + off_miss, miss.size() * 4, cudaMemcpyDeviceToHost, st));
    SCS_CUDA(cudaStreamSynchronize(st));
    *n_kept = int64_t(kept);
    for (int c = 0; c < n_cells; ++c) {
        unsigned long long m = 0;
        for (int rep = 0; rep < MISS_REPL; ++rep) m += miss[size_t(rep) * n_cells + c];
        missing_frac[c] = kept > 0 ? double(m) / double(kept) : 0.0;
    }
    return SCS_OK;
}

// Row filter + missing counts for a batch of datasets stacked in one matrix (replicate sweeps): per-group counters, all
// outputs stay on the device.  d_group_row0: [n_groups + 1] row offsets on the device.  Narrow matrices only (<= 512 cells).
int gt_reduce_groups_launch(scs_ctx* ctx, const uint8_t* d_gt, int64_t pitch_bytes, int n_groups, const long long* d_group_row0,
                            int64_t max_rows, int n_cells, const uint32_t* d_active_mask, int filter_rows, uint8_t* d_row_keep,
                            unsigned long long* d_n_kept, unsigned int* d_missing, cudaStream_t st) {
    const int n_words = (n_cells + 15) / 16;
    if (n_words > 32) return set_error(SCS_ERR_ARG, "grouped row filter: at most 512 cells per dataset (got %d)", n_cells);
    int shift = 0;
    while ((1 << shift) < n_words) ++shift;
    SCS_CUDA(cudaMemsetAsync(d_n_kept, 0, size_t(n_groups) * sizeof(unsigned long long), st));
    SCS_CUDA(cudaMemsetAsync(d_missing, 0, size_t(n_groups) * n_cells * sizeof(unsigned int), st));
    const int64_t want = (max_rows * (int64_t(1) << shift) + 255) / 256;
    const unsigned gx = unsigned(std::max<int64_t>(1, std::min<int64_t>(want, std::max<int64_t>(1, int64_t(ctx->sm_count) * 8 / std::max(1, n_groups)))));
    dim3 grid(gx, unsigned(n_groups));
    scs_gt_reduce_narrow_kernel<<<grid, 256, 0, st>>>(d_gt, pitch_bytes, 0, n_words, shift, n_cells, d_active_mask, filter_rows, d_row_keep, d_n_kept,
                                                      d_missing, d_group_row0);
    SCS_CUDA(cudaGetLastError());
    return SCS_OK;
}

struct Vcf {
    scs_ctx* ctx = nullptr;
    int64_t n_lines = 0, n_kept = 0, n_noalt = 0, n_filtered = 0;
    int n_incl = 0, pitch = 0;
    uint8_t* packed = nullptr;   // [n_kept][pitch]
    int64_t* pos = nullptr;      // [n_kept][2]
};

}  // namespace scs

using namespace scs;

extern "C" {

int scs_vcf_decode(scs_ctx* ctx, const char* text, int64_t nbytes, int32_t n_samples, const int32_t* incl_idx, int32_t n_incl,
                   scs_vcf** out) {
    if (!ctx || !text || !incl_idx || !out) return set_error(SCS_ERR_ARG, "null argument");
    if (nbytes <= 0 || n_samples <= 0 || n_incl <= 0) return set_error(SCS_ERR_ARG, "empty VCF input (bytes=%lld samples=%d kept=%d)", (long long)nbytes, n_samples, n_incl);
    for (int i = 0; i < n_incl; ++i)
        if (incl_idx[i] < 0 || incl_idx[i] >= n_samples) return set_error(SCS_ERR_ARG, "kept sample column %d out of range", incl_idx[i]);
    SCS_CUDA(cudaSetDevice(ctx->device));
    Vcf* v = new Vcf();
    v->ctx = ctx;
    v->n_incl = n_incl;
    v->pitch = ((n_incl + 3) / 4 + 15) / 16 * 16;
    char* d_text = nullptr;
    unsigned long long* d_cnt = nullptr;
    int64_t* d_ls = nullptr;
    int* d_incl = nullptr;
    uint8_t *d_tmp = nullptr, *d_packed_all = nullptr, *d_status = nullptr;
    int64_t* d_pos_all = nullptr;
    int* d_err = nullptr;
    long long* d_flag = nullptr;
    unsigned long long* d_hist = nullptr;
    int rc = SCS_OK;
    const int64_t nblk = (nbytes + 4095) / 4096;
    do {
#define DEC_TRY(call)                                                                                      \
    {                                                                                                      \
        cudaError_t _e = (call);                                                                           \
        if (_e != cudaSuccess) {                                                                           \
            rc = set_error(_e == cudaErrorMemoryAllocation ? SCS_ERR_OOM : SCS_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(_e)); \
            break;                                                                                         \
        }                                                                                                  \
    }
        DEC_TRY(cudaMalloc(&d_text, size_t(nbytes) + 1));
        DEC_TRY(cudaMemcpy(d_text, text, size_t(nbytes), cudaMemcpyHostToDevice));
        DEC_TRY(cudaMalloc(&d_cnt, size_t(nblk + 1) * sizeof(unsigned long long)));
        DEC_TRY(cudaMemset(d_cnt, 0, size_t(nblk + 1) * sizeof(unsigned long long)));
        scs_mark_newlines_kernel<<<unsigned(nblk), 128>>>(d_text, nbytes, d_cnt);
        thrust::exclusive_scan(thrust::device, thrust::device_ptr<unsigned long long>(d_cnt), thrust::device_ptr<unsigned long long>(d_cnt) + nblk + 1,
                               thrust::device_ptr<unsigned long long>(d_cnt));
        unsigned long long n_nl = 0;
        DEC_TRY(cudaMemcpy(&n_nl, d_cnt + nblk, sizeof(n_nl), cudaMemcpyDeviceToHost));
        const bool ends_nl = text[nbytes - 1] == '\n';
        const int64_t n_lines = int64_t(n_nl) + (ends_nl ? 0 : 1);
        v->n_lines = n_lines;
        DEC_TRY(cudaMalloc(&d_ls, size_t(n_nl + 2) * sizeof(int64_t)));
        {
            const int64_t zero = 0;
            DEC_TRY(cudaMemcpy(d_ls, &zero, sizeof(zero), cudaMemcpyHostToDevice));
            scs_fill_line_starts_kernel<<<unsigned((nblk + 127) / 128), 128>>>(d_text, nbytes, d_cnt, d_ls);
            if (!ends_nl) DEC_TRY(cudaMemcpy(d_ls + n_nl + 1, &nbytes, sizeof(int64_t), cudaMemcpyHostToDevice));
        }
        DEC_TRY(cudaMalloc(&d_incl, size_t(n_incl) * sizeof(int)));
        DEC_TRY(cudaMemcpy(d_incl, incl_idx, size_t(n_incl) * sizeof(int), cudaMemcpyHostToDevice));
        DEC_TRY(cudaMalloc(&d_tmp, size_t(n_lines) * n_samples));
        DEC_TRY(cudaMalloc(&d_packed_all, size_t(n_lines) * v->pitch));
        DEC_TRY(cudaMalloc(&d_status, size_t(n_lines)));
        DEC_TRY(cudaMalloc(&d_pos_all, size_t(n_lines) * 2 * sizeof(int64_t)));
        DEC_TRY(cudaMalloc(&d_err, 2 * sizeof(int)));
        DEC_TRY(cudaMemset(d_err, 0, 2 * sizeof(int)));
        DEC_TRY(cudaMemset(d_pos_all, 0, size_t(n_lines) * 2 * sizeof(int64_t)));
        DecodeParams p;
        p.text = d_text;
        p.nbytes = nbytes;
        p.line_start = d_ls;
        p.n_lines = n_lines;
        p.n_samples = n_samples;
        p.incl_idx = d_incl;
        p.n_incl = n_incl;
        p.pitch = v->pitch;
        p.tmp = d_tmp;
        p.packed_all = d_packed_all;
        p.status = d_status;
        p.pos_all = d_pos_all;
        p.err = d_err;
        const unsigned grid = unsigned(std::min<int64_t>(n_lines, int64_t(ctx->sm_count) * 16));
        scs_decode_lines_kernel<<<grid, DEC_THREADS>>>(p);
        DEC_TRY(cudaGetLastError());
        int err[2];
        DEC_TRY(cudaMemcpy(err, d_err, sizeof(err), cudaMemcpyDeviceToHost));
        if (err[0] != PERR_NONE) {
            static const char* msg[] = {"", "wrong number of tab-separated columns", "GT does not end in an allele digit",
                                        "sample record has fewer sub-fields than FORMAT", "CHROM/POS is not an integer (after dropping 'chr')",
                                        "the majority 'alt' allele index is 0 (GT written reference-last, e.g. 1|0): the reference's result is undefined here",
                                        "FORMAT has no GT field"};
            rc = set_error(SCS_ERR_PARSE, "VCF line %d: %s", err[1] + 1, msg[err[0]]);
            break;
        }
        DEC_TRY(cudaMalloc(&d_flag, size_t(n_lines + 1) * sizeof(long long)));
        DEC_TRY(cudaMemset(d_flag, 0, size_t(n_lines + 1) * sizeof(long long)));
        scs_status_flags_kernel<<<unsigned((n_lines + 255) / 256), 256>>>(d_status, n_lines, d_flag);
        thrust::exclusive_scan(thrust::device, thrust::device_ptr<long long>(d_flag), thrust::device_ptr<long long>(d_flag) + n_lines + 1,
                               thrust::device_ptr<long long>(d_flag));
        long long n_kept = 0;
        DEC_TRY(cudaMemcpy(&n_kept, d_flag + n_lines, sizeof(n_kept), cudaMemcpyDeviceToHost));
        v->n_kept = n_kept;
        DEC_TRY(cudaMalloc(&d_hist, 8 * sizeof(unsigned long long)));
        DEC_TRY(cudaMemset(d_hist, 0, 8 * sizeof(unsigned long long)));
        scs_count_status_kernel<<<unsigned((n_lines + 255) / 256), 256>>>(d_status, n_lines, d_hist);
        unsigned long long hist[8];
        DEC_TRY(cudaMemcpy(hist, d_hist, sizeof(hist), cudaMemcpyDeviceToHost));
        v->n_noalt = int64_t(hist[LINE_NOALT]);
        v->n_filtered = int64_t(hist[LINE_FILTERED]);
        DEC_TRY(cudaMalloc(&v->packed, std::max<size_t>(16, size_t(n_kept) * v->pitch)));
        DEC_TRY(cudaMalloc(&v->pos, std::max<size_t>(16, size_t(n_kept) * 2 * sizeof(int64_t))));
        if (n_kept > 0) {
            scs_compact_rows_kernel<<<grid, 128>>>(d_status, d_flag, n_lines, d_packed_all, d_pos_all, v->pitch, v->packed, v->pos);
            DEC_TRY(cudaGetLastError());
        }
        DEC_TRY(cudaDeviceSynchronize());
#undef DEC_TRY
    } while (0);
    cudaFree(d_text);
    cudaFree(d_cnt);
    cudaFree(d_ls);
    cudaFree(d_incl);
    cudaFree(d_tmp);
    cudaFree(d_packed_all);
    cudaFree(d_status);
    cudaFree(d_pos_all);
    cudaFree(d_err);
    cudaFree(d_flag);
    cudaFree(d_hist);
    if (rc != SCS_OK) {
        scs_vcf_destroy(reinterpret_cast<scs_vcf*>(v));
        return rc;
    }
    *out = reinterpret_cast<scs_vcf*>(v);
    return SCS_OK;
}

int scs_vcf_destroy(scs_vcf* h) {
    Vcf* v = reinterpret_cast<Vcf*>(h);
    if (!v) return SCS_OK;
    cudaFree(v->packed);
    cudaFree(v->pos);
    delete v;
    return SCS_OK;
}

int scs_vcf_info(scs_vcf* h, int64_t* info, int32_t n) {
    Vcf* v = reinterpret_cast<Vcf*>(h);
    if (!v || !info) return set_error(SCS_ERR_ARG, "null argument");
    const int64_t a[6] = {v->n_kept, v->n_incl, v->pitch, v->n_lines, v->n_noalt, v->n_filtered};
    for (int i = 0; i < n && i < 6; ++i) info[i] = a[i];
    return SCS_OK;
}

const uint8_t* scs_vcf_device_genotypes(scs_vcf* h) {
    Vcf* v = reinterpret_cast<Vcf*>(h);
    return v ? v->packed : nullptr;
}

int scs_vcf_copy(scs_vcf* h, uint8_t* packed, int64_t* pos) {
    Vcf* v = reinterpret_cast<Vcf*>(h);
    if (!v) return set_error(SCS_ERR_ARG, "null argument");
    if (packed && v->n_kept) SCS_CUDA(cudaMemcpy(packed, v->packed, size_t(v->n_kept) * v->pitch, cudaMemcpyDeviceToHost));
    if (pos && v->n_kept) SCS_CUDA(cudaMemcpy(pos, v->pos, size_t(v->n_kept) * 2 * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return SCS_OK;
}

int scs_gt_reduce(scs_ctx* ctx, const uint8_t* d_gt, int64_t pitch_bytes, int64_t n_snv, int32_t n_cells, const uint8_t* col_active,
                  int32_t filter_rows, uint8_t* d_row_keep, int64_t* n_kept, double* missing_frac, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (!ctx || !d_gt || !col_active || !d_row_keep || !n_kept || !missing_frac) return set_error(SCS_ERR_ARG, "null argument");
    if ((pitch_bytes & 3) || pitch_bytes * 4 < n_cells) return set_error(SCS_ERR_ARG, "packed genotype pitch %lld must be a multiple of 4 and cover %d cells", (long long)pitch_bytes, n_cells);
    SCS_CUDA(cudaSetDevice(ctx->device));
    uint8_t* sc = static_cast<uint8_t*>(ctx_scratch(ctx, gt_reduce_scratch_bytes(n_cells)));
    if (!sc) return set_error(SCS_ERR_OOM, "device scratch allocation failed");
    int rc = gt_reduce_launch(ctx, d_gt, pitch_bytes, n_snv, n_cells, col_active, filter_rows, d_row_keep, sc, false, st);
    if (rc != SCS_OK) return rc;
    return gt_reduce_fetch(sc, n_cells, n_kept, missing_frac, st);
}

}  // extern "C"
