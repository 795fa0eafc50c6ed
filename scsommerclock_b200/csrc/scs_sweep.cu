// scs_sweep.cu -- the per-replicate model set-up of a simulation sweep, batched on the device.
//
// run_poisson_tree_test (run_PT_test.py:703-728) does, per (VCF, tree) pair and after the placement:
//     add_br_weights   (:451-515)  branch weights from the per-cell missing rates and the clade of every node,
//     get_model_data   (:518-635)  the branch order of the LRT model (post-order over the internal nodes; of two internal
//                                  children the one with fewer soft mutations first, a leaf before an internal sibling),
// then get_LRT_poisson for every w_max.  For a sweep of thousands of small trees these are thousands of tiny host jobs
// (and host round trips: the branch order needs the soft weights).  Here they are two kernels over ALL replicates that
// write straight into the device buffers of an LRT plan (scs_lrt_plan_device_inputs), fed by the per-group counters of
// scs_gt_reduce_groups and the soft weights of the batched placement -- nothing comes back to the host between the upload
// of the packed matrices and the download of the test results.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "scs_internal.h"

namespace scs {

constexpr double SW_LMIN = 1e-6;          // LAMBDA_MIN, run_PT_test.py:21
constexpr int SW_THREADS = 128;

// One CTA per replicate.  Dynamic shared memory: diff[n_cells] doubles, p[n_rows] doubles, red[SW_THREADS] doubles.
// missing rate of a cell (:357-358): clip(missing count / kept SNVs, LAMBDA_MIN, 1 - FN - 0.2); only the cells of the tree
// (the clade of its root row) enter the sums (:475-478 index by the tree's leaf names).
__global__ void __launch_bounds__(SW_THREADS)
scs_sweep_weights_kernel(const uint32_t* __restrict__ bits, long long pitch_words, int n_rows, int n_cells, const int* __restrict__ info,
                         const unsigned long long* __restrict__ n_kept, const unsigned int* __restrict__ missing,
                         const double* __restrict__ FP, const double* __restrict__ FN, const double* __restrict__ w_max, int n_w,
                         double* __restrict__ wn /*[groups][n_w][n_rows]*/, int* __restrict__ status) {
    extern __shared__ double sw_dyn[];
    double* diff = sw_dyn;
    double* pado = diff + n_cells;
    double* red = pado + n_rows;
    __shared__ int s_root;
    const int g = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarp = SW_THREADS >> 5;
    const int n_nodes = info[4 * g], n_leaves = info[4 * g + 1];
    const uint32_t* gb = bits + size_t(g) * n_rows * pitch_words;
    const int words = (n_cells + 31) / 32;
    if (t == 0) s_root = n_rows;
    __syncthreads();
    // root = first node whose clade holds every cell of the tree (:497)
    for (int i = t; i < n_nodes; i += SW_THREADS) {
        int c = 0;
        for (int w = 0; w < words; ++w) c += __popc(gb[size_t(i) * pitch_words + w]);
        if (c == n_leaves) atomicMin(&s_root, i);
    }
    __syncthreads();
    const int root = s_root;
    if (root >= n_nodes) {
        if (t == 0) status[g] = 1;        // no node contains every leaf: not a rooted tree
        return;
    }
    const double fp = FP[g], fn = FN[g];
    const double kept = double(n_kept[g]);
    const double cap = 1.0 - fn - 0.2;
    double part = 0.0;
    for (int c = t; c < n_cells; c += SW_THREADS) {
        double d = 0.0;
        if ((gb[size_t(root) * pitch_words + (c >> 5)] >> (c & 31)) & 1u) {
            double ms = kept > 0.0 ? double(missing[size_t(g) * n_cells + c]) / kept : 0.0;
            ms = fmin(fmax(ms, SW_LMIN), cap);                               // np.clip(.., LAMBDA_MIN, 1 - FN - 0.2)
            const double l_tn = log((1.0 - ms) * (1.0 - fp) + ms);           // :475
            const double l_do = log((1.0 - ms) * fn + ms);                   // :477
            d = l_do - l_tn;
            part += l_tn;
        }
        diff[c] = d;
    }
    red[t] = part;
    __syncthreads();
    double sum_tn = 0.0;
    for (int i = 0; i < SW_THREADS; ++i) sum_tn += red[i];                   // fixed order: the same bits on every run
    __syncthreads();
    // p_ADO of every node (:483): one warp per node, lanes over the words of its mask
    for (int i = warp; i < n_nodes; i += nwarp) {
        double acc = 0.0;
        for (int w = 0; w < words; ++w) {
            const uint32_t b = gb[size_t(i) * pitch_words + w];
            const int c = w * 32 + lane;
            if (((b >> lane) & 1u) && c < n_cells) acc += diff[c];
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            const double xlog = sum_tn + acc;
            pado[i] = xlog < -708.3964185322641 ? 1e-6 : fmax(exp(xlog), 1e-6);   // the reference catches the underflow (:485)
        }
    }
    __syncthreads();
    // weights (:492) and their normalisation to the number of branches without the root (:498-502)
    for (int wi = 0; wi < n_w; ++wi) {
        const double wmax = w_max[wi];
        double s = 0.0;
        for (int i = t; i < n_nodes; i += SW_THREADS)
            if (i != root) s += fmin(1.0 / ((1.0 - pado[i]) * pado[i]), wmax);
        red[t] = s;
        __syncthreads();
        double sum = 0.0;
        for (int i = 0; i < SW_THREADS; ++i) sum += red[i];
        __syncthreads();
        double* dst = wn + (size_t(g) * n_w + wi) * n_rows;
        for (int i = t; i < n_rows; i += SW_THREADS) {
            double v = 0.0;
            if (i < n_nodes) v = i == root ? -1.0 : double(n_nodes - 1) * fmin(1.0 / ((1.0 - pado[i]) * pado[i]), wmax) / sum;
            dst[i] = v;
        }
    }
    if (t == 0) status[g] = 0;
}

// get_model_data's branch order (:569-611) for every replicate, and the LRT plan's inputs.  One CTA per replicate: thread 0
// walks the tree (post-order over the internal nodes below the root holder), the block then writes, for every w_max, the
// problem's child_int, gather index (branch -> row of the soft-weight vector) and branch weights.
// children: [groups][n_rows][2] rows of the two children (-1: leaf), row 0.. in tree.iter_descendants() order.
__global__ void __launch_bounds__(64)
scs_sweep_branch_order_kernel(const int* __restrict__ children, int n_rows, const int* __restrict__ info, const double* __restrict__ soft,
                              const double* __restrict__ wn, int n_w, int n_br, int* __restrict__ d_child, int* __restrict__ d_gather,
                              double* __restrict__ d_w, int* __restrict__ scratch /*[groups][3 n_rows]*/, int* __restrict__ status) {
    const int g = blockIdx.x, t = threadIdx.x;
    const int n_nodes = info[4 * g], n_leaves = info[4 * g + 1];
    // the child table, the soft weights and the walk's work arrays live in shared memory: thread 0's walk is a chain of
    // dependent loads, ~100 us per launch from global memory (L2 latency per step) whatever the number of groups
    extern __shared__ double bo_dyn[];
    double* y = bo_dyn;                                        // [n_rows]
    int* ch = reinterpret_cast<int*>(y + n_rows);              // [2 n_rows]
    int* stack = ch + 2 * n_rows;                              // [n_rows] node stack, [n_rows] post-order index of internal nodes,
    int* int_index = stack + n_rows;                           // [n_rows] branch order (node rows)
    int* order = int_index + n_rows;
    for (int i = t; i < n_rows; i += blockDim.x) y[i] = soft[size_t(g) * n_rows + i];
    for (int i = t; i < 2 * n_rows; i += blockDim.x) ch[i] = children[size_t(g) * n_rows * 2 + i];
    (void)scratch;
    __syncthreads();
    __shared__ int s_ok;
    if (t == 0) {
        int ok = (2 * (n_leaves - 1) == n_br && n_nodes == 2 * n_leaves - 1) ? 1 : 0;
        if (ok) {
            // iterative post-order from row 0 (the single child of the root holder: every other node hangs below it)
            int sp = 0, n_int = 0, k = 0;
            stack[sp++] = 0;                           // bit 30 marks "children done"
            while (sp > 0) {
                const int top = stack[--sp];
                const int v = top & 0x3FFFFFFF;
                const int a0 = ch[2 * v], b0 = ch[2 * v + 1];
                if (a0 < 0) continue;                  // a leaf
                if (!(top & 0x40000000)) {
                    stack[sp++] = v | 0x40000000;
                    stack[sp++] = b0;                  // children[0] is walked first: push it last
                    stack[sp++] = a0;
                    continue;
                }
                int a = a0, b = b0;
                const bool ta = ch[2 * a] < 0, tb = ch[2 * b] < 0;
                if (ta == tb) {
                    if (!ta && y[b] < y[a]) {          // two internal children: fewer mutations first (stable)
                        a = b0;
                        b = a0;
                    }
                } else if (tb) {                       // one leaf: the leaf first
                    a = b0;
                    b = a0;
                }
                if (k + 2 > n_br) {
                    ok = 0;
                    break;
                }
                int_index[v] = n_int++;
                order[k++] = a;
                order[k++] = b;
            }
            if (k != n_br) ok = 0;
        }
        s_ok = ok;
        if (!ok) status[g] = 2;                        // the tree does not have the planned number of leaves
    }
    __syncthreads();
    if (!s_ok) return;
    for (int idx = t; idx < n_w * n_br; idx += blockDim.x) {
        const int wi = idx / n_br, k = idx - wi * n_br;
        const int row = order[k];
        const size_t dst = (size_t(g) * n_w + wi) * n_br + k;
        d_child[dst] = ch[2 * row] < 0 ? -1 : int_index[row];
        d_gather[dst] = g * n_rows + row;
        d_w[dst] = wn[(size_t(g) * n_w + wi) * n_rows + row];
    }
}

// float.__repr__: the shortest digits that round-trip, fixed notation for 1e-4 <= |v| < 1e16, else d.ddde-XX
static std::string py_repr(double v) {
    if (std::isnan(v)) return "nan";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    if (v == 0.0) return std::signbit(v) ? "-0.0" : "0.0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);      // shortest round-trip digits
    std::string sci(buf, r.ptr);
    const size_t e = sci.find('e');
    std::string mant = sci.substr(0, e);
    const int exp10 = atoi(sci.c_str() + e + 1);
    std::string sign;
    if (mant[0] == '-') {
        sign = "-";
        mant.erase(0, 1);
    }
    std::string digits;
    for (char c : mant)
        if (c != '.') digits.push_back(c);
    if (exp10 < -4 || exp10 >= 16) {
        std::string out = sign + digits.substr(0, 1);
        if (digits.size() > 1) out += "." + digits.substr(1);
        char eb[16];
        snprintf(eb, sizeof(eb), "e%c%02d", exp10 < 0 ? '-' : '+', exp10 < 0 ? -exp10 : exp10);
        return out + eb;
    }
    std::string out = sign;
    if (exp10 < 0) {
        out += "0.";
        out.append(size_t(-exp10 - 1), '0');
        out += digits;
    } else if (size_t(exp10) + 1 >= digits.size()) {
        out += digits;
        out.append(size_t(exp10) + 1 - digits.size(), '0');
        out += ".0";
    } else {
        out += digits.substr(0, size_t(exp10) + 1) + "." + digits.substr(size_t(exp10) + 1);
    }
    return out;
}

}  // namespace scs

using namespace scs;

extern "C" {

int scs_gt_reduce_groups(scs_ctx* ctx, const uint8_t* d_gt, int64_t pitch_bytes, int32_t n_groups, const int64_t* d_group_row0,
                         int64_t max_rows, int32_t n_cells, const uint32_t* d_active_mask, int32_t filter_rows, uint8_t* d_row_keep,
                         uint64_t* d_n_kept, uint32_t* d_missing, void* stream) {
    if (!ctx || !d_gt || !d_group_row0 || !d_active_mask || !d_row_keep || !d_n_kept || !d_missing) return set_error(SCS_ERR_ARG, "null argument");
    if (n_groups <= 0 || n_cells <= 0 || (pitch_bytes & 3) || pitch_bytes * 4 < n_cells) return set_error(SCS_ERR_ARG, "bad shape");
    SCS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    static_assert(sizeof(long long) == sizeof(int64_t) && sizeof(unsigned long long) == sizeof(uint64_t), "64-bit integer types");
    return gt_reduce_groups_launch(ctx, d_gt, pitch_bytes, n_groups, reinterpret_cast<const long long*>(d_group_row0), max_rows, n_cells,
                                   d_active_mask, filter_rows, d_row_keep, reinterpret_cast<unsigned long long*>(d_n_kept), d_missing, st);
}

int64_t scs_sweep_model_workspace(int32_t n_groups, int32_t n_rows, int32_t n_w) {
    // wn [G][n_w][n_rows] + constants [2 G + n_w] doubles, then [G][3 n_rows] ints (padded to a multiple of 8 bytes)
    const int64_t dbl = int64_t(n_groups) * n_w * n_rows + 2 * int64_t(n_groups) + n_w;
    const int64_t ints = int64_t(n_groups) * 3 * n_rows;
    return dbl * 8 + (ints * 4 + 7) / 8 * 8;
}

int scs_sweep_model(scs_ctx* ctx, int32_t n_groups, int32_t n_rows, int32_t n_cells, int64_t pitch_words, const uint32_t* d_bits,
                    const int32_t* d_children, const int32_t* d_info, const uint64_t* d_n_kept, const uint32_t* d_missing,
                    const double* FP, const double* FN, const double* w_max, int32_t n_w, const double* d_soft, int32_t n_br,
                    scs_lrt_plan* plan, int32_t* d_status, void* d_workspace, void* stream) {
    if (!ctx || !d_bits || !d_children || !d_info || !d_n_kept || !d_missing || !FP || !FN || !w_max || !d_soft || !plan || !d_status || !d_workspace)
        return set_error(SCS_ERR_ARG, "null argument");
    if (n_groups <= 0 || n_rows <= 0 || n_cells <= 0 || n_w <= 0 || n_br < 2) return set_error(SCS_ERR_ARG, "bad shape");
    if (reinterpret_cast<uintptr_t>(d_workspace) & 7) return set_error(SCS_ERR_ARG, "the workspace must be 8-byte aligned");
    SCS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    double* d_wn = static_cast<double*>(d_workspace);
    double* d_consts = d_wn + size_t(n_groups) * n_w * n_rows;
    int* d_scratch = reinterpret_cast<int*>(d_consts + 2 * size_t(n_groups) + n_w);
    std::vector<double> consts(size_t(2) * n_groups + n_w);
    for (int g = 0; g < n_groups; ++g) {
        consts[size_t(g)] = FP[g];
        consts[size_t(n_groups) + g] = FN[g];
    }
    for (int w = 0; w < n_w; ++w) consts[size_t(2) * n_groups + w] = w_max[w];
    {
        const int urc = upload_small(ctx, d_consts, consts.data(), consts.size() * sizeof(double), st);      // (not the copy engine: scs_internal.h)
        if (urc != SCS_OK) return urc;
    }
    int32_t *d_child = nullptr, *d_gather = nullptr;
    double* d_w = nullptr;
    int rc = scs_lrt_plan_device_inputs(plan, int64_t(n_groups) * n_rows, &d_child, &d_gather, &d_w);
    if (rc != SCS_OK) return rc;
    const size_t smem = (size_t(n_cells) + n_rows + SW_THREADS) * sizeof(double);
    if (smem > 48 * 1024) cudaFuncSetAttribute(scs_sweep_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    scs_sweep_weights_kernel<<<n_groups, SW_THREADS, smem, st>>>(d_bits, pitch_words, n_rows, n_cells, d_info,
                                                                  reinterpret_cast<const unsigned long long*>(d_n_kept), d_missing, d_consts,
                                                                  d_consts + n_groups, d_consts + 2 * size_t(n_groups), n_w, d_wn, d_status);
    const size_t bo_smem = size_t(n_rows) * (sizeof(double) + 5 * sizeof(int));
    if (bo_smem > 48 * 1024) cudaFuncSetAttribute(scs_sweep_branch_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bo_smem));
    scs_sweep_branch_order_kernel<<<n_groups, 64, bo_smem, st>>>(d_children, n_rows, d_info, d_soft, d_wn, n_w, n_br, d_child, d_gather, d_w, d_scratch,
                                                                 d_status);
    SCS_CUDA(cudaGetLastError());
    return SCS_OK;
}

// The model line of run_poisson_tree_test's TSV (:722-724) for every replicate: "{muts}\t{FN:.4f}\t{FP:.4f}" followed, per
// w_max, by "\t{LR:0>5.3f}\t{dof}\t{p_val}\tH{hyp}".  out: [groups][n_w][8] as the LRT kernel writes it (a non-zero status
// prints the reference's failure tuple: nan).  Rows are written back to back into buf; row_off[g] .. row_off[g + 1] delimits
// row g.  Returns SCS_ERR_ARG when buf is too small (row_off[n_groups] then holds the size needed).
static void format_rows_block(int g0, int g1, int n_w, const double* out, const int64_t* n_kept, const double* FN, const double* FP, std::string& s,
                              std::vector<int64_t>& len) {
    s.reserve(size_t(g1 - g0) * (32 + size_t(n_w) * 48));
    char tmp[96];
    for (int g = g0; g < g1; ++g) {
        const size_t before = s.size();
        snprintf(tmp, sizeof(tmp), "%lld\t%.4f\t%.4f", (long long)n_kept[g], FN[g], FP[g]);
        s += tmp;
        for (int w = 0; w < n_w; ++w) {
            const double* o = out + (size_t(g) * n_w + w) * 8;
            const bool failed = o[7] != 0.0 || !std::isfinite(o[0]);
            if (failed) {
                s += "\t00nan\tnan\tnan\tH0";          // f"{nan:0>5.3f}" is '00nan', int(nan < 0.05) is 0
                continue;
            }
            snprintf(tmp, sizeof(tmp), "%.3f", o[0]);
            std::string lr(tmp);
            if (lr.size() < 5) lr.insert(0, 5 - lr.size(), '0');
            s += '\t';
            s += lr;
            snprintf(tmp, sizeof(tmp), "\t%lld\t", (long long)llround(o[1]));
            s += tmp;
            s += py_repr(o[3]);
            s += (o[3] < 0.05) ? "\tH1" : "\tH0";
        }
        len[size_t(g)] = int64_t(s.size() - before);
    }
}

int scs_format_pt_rows(int32_t n_groups, int32_t n_w, const double* out, const int64_t* n_kept, const double* FN, const double* FP, char* buf,
                       int64_t buf_len, int64_t* row_off) {
    if (!out || !n_kept || !FN || !FP || !buf || !row_off) return set_error(SCS_ERR_ARG, "null argument");
    if (n_groups < 0 || n_w < 0) return set_error(SCS_ERR_ARG, "bad shape");
    // blocks of rows on host threads (number formatting is ~0.2 us per field: 2 ms per 300 rows of ten w_max on one thread)
    int n_thr = 1;
    if (n_groups >= 64) {
        const unsigned hw = std::thread::hardware_concurrency();
        n_thr = int(std::max(1u, std::min(std::min(hw, 8u), unsigned(n_groups / 32))));
    }
    std::vector<std::string> parts{size_t(n_thr)};
    std::vector<int64_t> len(size_t(n_groups), 0);
    auto block = [&](int t) {
        const int g0 = int(int64_t(n_groups) * t / n_thr), g1 = int(int64_t(n_groups) * (t + 1) / n_thr);
        format_rows_block(g0, g1, n_w, out, n_kept, FN, FP, parts[size_t(t)], len);
    };
    if (n_thr == 1) {
        block(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 1; t < n_thr; ++t) th.emplace_back(block, t);
        block(0);
        for (auto& x : th) x.join();
    }
    int64_t total = 0;
    for (int g = 0; g < n_groups; ++g) {
        row_off[g] = total;
        total += len[size_t(g)];
    }
    row_off[n_groups] = total;
    if (total > buf_len) return set_error(SCS_ERR_ARG, "row buffer too small: %lld bytes needed", (long long)total);
    int64_t at = 0;
    for (const std::string& part : parts) {
        memcpy(buf + at, part.data(), part.size());
        at += int64_t(part.size());
    }
    return SCS_OK;
}

}  // extern "C"
