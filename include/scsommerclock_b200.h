/*
 * scsommerclock_b200.h -- C ABI of the B200-native Poisson Tree (PT) test hot path.
 *
 * The reference (cbg-ethz/scSomMerClock, run_PT_test.py) is pure Python; the
 * functions below are what a maintainer would bind (ctypes)
 * in place of the NumPy/SciPy bodies of
 *
 *   get_mut_df          run_PT_test.py:27-207    -> scs_vcf_decode*
 *   map_mutations_gt    run_PT_test.py:386-448   -> scs_placement_* / scs_map_mutations / scs_place_percell
 *   _normalize_log_probs run_PT_test.py:368-383  -> fused in the placement epilogue
 *   add_br_weights      run_PT_test.py:451-515   -> scs_branch_weights
 *   get_LRT_poisson     run_PT_test.py:638-700   -> scs_lrt_poisson_batch
 *
 * Plain pointers and sizes only.  Functions whose buffer arguments are named
 * d_* take DEVICE pointers (memory of the current CUDA primary context, e.g. a
 * torch tensor's data_ptr()); everything else is HOST memory.  `stream` is a
 * cudaStream_t passed as void* (NULL = the legacy default stream).
 *
 * Every function returns SCS_OK (0) or a negative error code; the message for
 * the calling thread's last error is scs_last_error().  There is no CPU
 * fallback: without a usable sm_100 device scs_ctx_create fails.
 */
#ifndef SCSOMMERCLOCK_B200_H
#define SCSOMMERCLOCK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCS_OK 0
#define SCS_ERR_ARG (-1)     /* invalid argument                               */
#define SCS_ERR_CUDA (-2)    /* CUDA runtime / driver failure                  */
#define SCS_ERR_OOM (-3)     /* device or host allocation failed               */
#define SCS_ERR_DEVICE (-4)  /* no sm_100 device: the hot path cannot run      */
#define SCS_ERR_PARSE (-5)   /* malformed VCF input                            */
#define SCS_ERR_NUMERIC (-6) /* solver failed to converge                      */

/* 2-bit genotype codes of the packed SNV x cell observation matrix; they are the
 * values of the reference's `muts` DataFrame (run_PT_test.py:156-167): 0 wild
 * type, 1 heterozygous, 2 homozygous, NaN missing.  Four cells per byte, cell c
 * in bits [2*(c%4), 2*(c%4)+2) of byte c/4 of its row. */
#define SCS_GT_WT 0
#define SCS_GT_HET 1
#define SCS_GT_HOM 2
#define SCS_GT_MISSING 3

typedef struct scs_ctx scs_ctx;
typedef struct scs_placement scs_placement;
typedef struct scs_vcf scs_vcf;
typedef struct scs_lrt_plan scs_lrt_plan;

/* ------------------------------------------------------------------ context */
const char* scs_version(void);
const char* scs_last_error(void);
int scs_ctx_create(int32_t device, scs_ctx** out);
int scs_ctx_destroy(scs_ctx* ctx);
/* info[0..5] = sm_count, cc_major, cc_minor, total_mem_MiB, l2_bytes, device */
int scs_ctx_info(scs_ctx* ctx, int64_t* info, int32_t n);
/* Host -> device upload of a SMALL page-locked buffer (cudaHostAlloc / cudaHostRegister / torch pin_memory) by a kernel that
 * reads it over PCIe, not by the copy engine: the device's one H2D queue serves all streams, so a small upload issued while
 * another stream's 100 MB matrix copy is in flight would wait for it.  The buffer must stay unchanged until the kernel has
 * run.  Asynchronous on `stream`; falls back to cudaMemcpyAsync when the memory is not page-locked. */
int scs_upload_async(void* d_dst, const void* h_pinned, int64_t bytes, void* stream);

/* ------------------------------------------------------------------ VCF decode
 * get_mut_df (run_PT_test.py:27-207, filter=True).  `text` is the decompressed
 * VCF (header lines included; they are skipped like the reference does),
 * n_samples the number of sample columns of the #CHROM line and incl_idx the
 * file columns that survive -excl/-incl (run_PT_test.py:57-76, :187-189), in
 * output order.  The decoded matrix stays on the device. */
int scs_vcf_decode(scs_ctx* ctx, const char* text, int64_t nbytes, int32_t n_samples, const int32_t* incl_idx,
                   int32_t n_incl, scs_vcf** out);
int scs_vcf_destroy(scs_vcf* vcf);
/* info[0..5] = kept SNVs, kept samples, row pitch (bytes), lines seen, lines
 * skipped for lack of any call (:147), lines dropped by FILTER / ISA (:175-181) */
int scs_vcf_info(scs_vcf* vcf, int64_t* info, int32_t n);
const uint8_t* scs_vcf_device_genotypes(scs_vcf* vcf);
/* Host copies: packed [kept][pitch] and pos [kept][2] = (chrom, pos); either may be NULL. */
int scs_vcf_copy(scs_vcf* vcf, uint8_t* packed, int64_t* pos);
/* get_gt_tree's filters (run_PT_test.py:350-358) on a device-resident packed
 * matrix: d_row_keep[i] = 1 when SNV i has a call in an active cell (all ones when
 * filter_rows == 0: the reference only filters when the outgroup is a column);
 * missing_frac[c] = fraction of kept SNVs where cell c is missing.  Everything (scratch reset, kernels, the
 * read-back) is ordered on `stream`, which is synchronised before the call returns. */
int scs_gt_reduce(scs_ctx* ctx, const uint8_t* d_gt, int64_t pitch_bytes, int64_t n_snv, int32_t n_cells,
                  const uint8_t* col_active, int32_t filter_rows, uint8_t* d_row_keep, int64_t* n_kept,
                  double* missing_frac, void* stream);

/* ---------------------------------------------------------------- placement
 * map_mutations_gt (run_PT_test.py:386-448): for every SNV, the normalised
 * likelihood of sitting on each row of the clade matrix S (the 2n-1 tree nodes in
 * ete3 level order plus a final all-zero "FP" row), summed over SNVs:
 *     soft_weights[b] = sum_i M[i,b],   M[i,:] = softmax_b(probs[i,:]).
 * A plan owns the device-resident operands and scratch for one (n_snv, n_cells,
 * n_rows) shape and can be re-run with new genotypes / clades / error rates. */
int scs_placement_create(scs_ctx* ctx, int64_t n_snv, int32_t n_cells, int32_t n_rows, scs_placement** out);
/* A batch of independent datasets ("groups": simulation replicates, each with its own
 * SNVs and its own tree over the same number of cells) in one plan and one launch per
 * pass.  Genotypes and row_keep are the groups' rows concatenated in group order, clade
 * masks are [n_groups * n_rows][pitch_words], results are [n_groups][n_rows]. */
int scs_placement_create_batch(scs_ctx* ctx, int32_t n_groups, const int64_t* n_snv, int32_t n_cells, int32_t n_rows,
                               scs_placement** out);
int scs_placement_destroy(scs_placement* plan);
/* d_gt: packed 2-bit genotypes [n_snv][pitch_bytes]; d_row_keep: optional
 * [n_snv] bytes, 0 drops the SNV (the reference's `muts_red.sum(axis=1) > 0`
 * filter, run_PT_test.py:352). */
int scs_placement_set_genotypes(scs_placement* plan, const uint8_t* d_gt, int64_t pitch_bytes,
                                const uint8_t* d_row_keep, void* stream);
int scs_placement_set_genotypes_host(scs_placement* plan, const uint8_t* gt, int64_t pitch_bytes,
                                     const uint8_t* row_keep, void* stream);
/* Same as scs_placement_set_genotypes without the plan taking a copy where it does not need one:
 * on the interval path (see scs_placement_set_clades) the kernel reads the caller's packed
 * matrix directly, so `d_gt` must stay alive and unchanged until the run that uses it has
 * finished.  Call it after the clades are set; on the GEMM path it builds the operand planes
 * like scs_placement_set_genotypes. */
int scs_placement_bind_genotypes(scs_placement* plan, const uint8_t* d_gt, int64_t pitch_bytes,
                                 const uint8_t* d_row_keep, void* stream);
/* get_gt_tree's filters (run_PT_test.py:350-358) and the plan's own genotype set-up in one call, for a plan whose
 * clades are set: d_row_keep / n_kept / missing_frac as scs_gt_reduce computes them (col_active: host, [n_cells]
 * bytes; filter_rows as there), kept in the plan, followed by what scs_placement_bind_genotypes would do.  On the
 * interval path all of it is ONE pass over d_gt (row flags, per-cell missing counts and the rows rewritten in the
 * plan's depth-first cell order come out of the same read); on the GEMM path it is the reduction kernels plus the
 * operand planes.  Asynchronous on `stream`, no host synchronisation: the counters stay on the device until
 * scs_placement_prepare_result.  accumulate != 0 adds to the counters of the previous prepare calls instead of
 * starting from zero (a matrix walked in SNV-row chunks through one chunk-sized plan).  d_gt must stay alive and
 * unchanged until the run that uses it has finished.  Plans holding several groups get the totals over all groups. */
int scs_placement_prepare(scs_placement* plan, const uint8_t* d_gt, int64_t pitch_bytes, const uint8_t* col_active,
                          int32_t filter_rows, int32_t accumulate, void* stream);
/* Waits for `stream`, then: *n_kept = SNVs kept so far, missing_frac[c] = fraction of them where cell c is missing. */
int scs_placement_prepare_result(scs_placement* plan, int64_t* n_kept, double* missing_frac, void* stream);
/* d_bits: clade membership bit mask [n_rows][pitch_words] of uint32, bit c%32 of
 * word c/32 set when cell c descends from the node (S, run_PT_test.py:389-397).
 * The masks also decide HOW the counts S x genotypes are formed: the clades of a tree are a
 * laminar family, so in a depth-first order of the cells every row is one interval and both
 * counts of a (SNV, clade) pair are a difference of per-SNV prefix counts -- the "interval
 * path" (csrc/scs_place_interval.cuh), used for single-dataset plans of at least 100 cells
 * whose masks pass that test (measured faster than the GEMM from there on).  Any other 0/1 matrix,
 * and every replicate batch given as masks only, goes through the tcgen05 GEMM (batches given as trees:
 * scs_placement_set_trees).  SCS_PLACE_ALGO=gemm forces the GEMM,
 * SCS_PLACE_ALGO=interval takes the interval path for tree-shaped masks at any size. */
int scs_placement_set_clades(scs_placement* plan, const uint32_t* d_bits, int64_t pitch_words, void* stream);
/* The clade rows of every group given as TREES: d_bits as for scs_placement_set_clades plus the child table and node
 * counts scs_tree_parse_batch produces (device copies: children[g][n_rows][2], info[g][4]).  Plans of up to 512 cells then
 * take the interval path for batches of small trees (a few lanes per SNV, no GEMM); anything else is handed to
 * scs_placement_set_clades.  A group whose child table is not a binary tree over single-cell leaves gets NaN weights. */
int scs_placement_set_trees(scs_placement* plan, const uint32_t* d_bits, int64_t pitch_words, const int32_t* d_children,
                            const int32_t* d_info, void* stream);
int scs_placement_set_clades_host(scs_placement* plan, const uint32_t* bits, int64_t pitch_words, void* stream);
/* Host-only helper (no device work): *laminar = 1 when the masks are a laminar family, and then
 * perm[k] (optional, [n_cells]) = cell at depth-first position k and lohi[r] (optional, [n_rows]) =
 * lo | hi << 16, the interval of positions row r covers (lo = hi = 0 for an empty row). */
int scs_clade_intervals(const uint32_t* bits, int64_t pitch_words, int32_t n_rows, int32_t n_cells, uint16_t* perm,
                        uint32_t* lohi, int32_t* laminar);
/* Asynchronous on `stream`.  d_soft_weights: [n_rows] doubles on the device, or
 * NULL to leave the result in the plan (scs_placement_device_result). */
int scs_placement_run(scs_placement* plan, double FP, double FN, double* d_soft_weights, void* stream);
/* One (FP, FN) pair per group. */
int scs_placement_run_batch(scs_placement* plan, const double* FP, const double* FN, double* d_soft_weights, void* stream);
/* Same, then copies [n_rows] doubles to host memory and synchronises. */
int scs_placement_run_host(scs_placement* plan, double FP, double FN, double* soft_weights, void* stream);
const void* scs_placement_device_result(scs_placement* plan);
/* Device durations (CUDA events on the launching stream) of the four kernels of the
 * last scs_placement_run: ms[0] GEMM pass 1 (row statistics), ms[1] merge, ms[2]
 * pass 2 (column sums: second GEMM pass or count-cache sweep), ms[3] reduction; on the
 * interval path ms[0] is the pass that rewrites the genotype rows in depth-first cell order
 * (zero when they were still valid), ms[1] is zero, ms[2] the placement kernel.  Waits for
 * that run to finish. */
int scs_placement_last_times(scs_placement* plan, float* ms);
/* Device ms of the last scs_placement_prepare's kernels (waits for them). */
int scs_placement_prepare_time(scs_placement* plan, float* ms);
/* info[0..9] = M_pad, N_pad, K_pad, tasks, run_len, superblock_width, grid,
 * kernel launches so far, runs, dynamic smem bytes, operand mode (0 = two u8
 * planes / kind::i8, 1 = one radix-packed e5m2 plane / kind::f8f6f4), K chunks,
 * 1 when the 2-CTA (cta_group::2) kernel is used, groups, pass-2 mode (0 not run yet,
 * 1 second GEMM pass, 2 sweep over the packed integer counts that pass 1 cached in HBM --
 * chosen per plan, see SCS_PLACE_PASS2 in DESIGN.md), algorithm (0 no clades yet, 1 GEMM,
 * 2 interval path), and for the interval path its grid, threads per CTA, dynamic smem bytes,
 * number of row runs and kernel variant (3: leaves outside the sweeps + count cache in shared
 * memory, 1: every row through the sweeps -- trees too large for the cache) */
int scs_placement_info(scs_placement* plan, int64_t* info, int32_t n);
/* Parity aid for small problems: counts[i][b][0/1] = number of cells of clade b
 * observed mutated / observed wild type for SNV i (exact integers). */
int scs_placement_debug_counts(scs_placement* plan, int32_t* counts, void* stream);
/* Analysis aid: pass-1 partials ([2 * column tiles][M_pad] of {u32 ref counts, f32 sum}) and merged
 * per-SNV records ([M_pad] of {u32 ref counts, f32 log2 sum}) of the last run; either may be NULL. */
int scs_placement_debug_stats(scs_placement* plan, void* stats, void* rowref);

/* One-shot host call mirroring map_mutations_gt's signature: host buffers in,
 * soft_weights[n_rows] out (H2D copy, both GEMM passes, D2H copy). */
int scs_map_mutations(scs_ctx* ctx, const uint8_t* gt, int64_t pitch_bytes, const uint8_t* row_keep,
                      int64_t n_snv, int32_t n_cells, const uint32_t* clade_bits, int64_t pitch_words,
                      int32_t n_rows, double FP, double FN, double* soft_weights);

/* map_mutations_gt with ONE FALSE-NEGATIVE RATE PER CELL -- the `isinstance(FN, pd.Series)` branch of
 * run_PT_test.py:413-418 (TP_m = S * log(1 - FN_i), FN_m = S * log(FN_i); FP stays a scalar).  d_gt / d_row_keep: the
 * packed matrix and (optionally, may be NULL) the row flags of scs_gt_reduce on the device; clade_bits: HOST masks
 * [n_rows][pitch_words] over the matrix columns (the rows of S, zero rows and the final 'FP' row included); FN_cells:
 * HOST, [n_cells], indexed by matrix column -- columns that belong to no clade (the outgroup) are ignored, every other
 * rate must lie in (0, 1).  d_soft_weights[n_rows] (device, float64) receives M.sum(axis=0).  The clade masks must be
 * tree-shaped (a laminar family: every call the reference makes), else SCS_ERR_ARG: the logits are differences of
 * per-SNV prefix sums of per-cell weights in a depth-first cell order (csrc/scs_place_percell.cuh).  kernel_ms (optional,
 * may be NULL): device time of the placement + reduction kernels.  Ordered on `stream`, which is synchronised before
 * the call returns. */
int scs_place_percell(scs_ctx* ctx, const uint8_t* d_gt, int64_t pitch_bytes, int64_t n_snv, int32_t n_cells,
                      const uint8_t* d_row_keep, const uint32_t* clade_bits, int64_t pitch_words, int32_t n_rows,
                      double FP, const double* FN_cells, double* d_soft_weights, float* kernel_ms, void* stream);

/* -------------------------------------------------------------- branch weights
 * add_br_weights (run_PT_test.py:451-515): clade_bits holds the tree's nodes
 * only ([n_nodes][pitch_words], ete3 level order); missing_rate[c] is the clipped
 * per-cell missing rate MS_i (run_PT_test.py:358), negative for cells that are not
 * leaves of the tree.  For every w_max: weights[w][node] (inverse variance,
 * clipped; -1 for the root) and weights_norm[w][node] (scaled to sum to the
 * number of branches; -1 for the root).  root_row receives the root's row. */
int scs_branch_weights(scs_ctx* ctx, const uint32_t* clade_bits, int64_t pitch_words, int32_t n_nodes,
                       int32_t n_cells, const double* missing_rate, double FP, double FN, const double* w_max,
                       int32_t n_w, double* weights, double* weights_norm, int32_t* root_row);

/* ------------------------------------------------------------------------ LRT
 * get_LRT_poisson (run_PT_test.py:638-700), batched: problem p owns branches
 * [br_off[p], br_off[p+1]) of Y / w / child_int.  Branch k hangs below internal
 * node k/2 (get_model_data's order, run_PT_test.py:569-622); child_int[k] is the
 * internal node below it, or -1 for a leaf.  out[p][8] = {-2logLR, dof, on_bound,
 * p-value, ll_H0, ll_H1, Newton iterations, status}; x_out (optional, may be NULL) receives the fitted
 * branch lengths.  status: 0 converged; 1 iteration limit; 2 degenerate input (the reference's scale_fac is 0 or
 * there are no counts); 3 numerical breakdown of the Newton system; 4 ll_H1 < 0 -- the reference's scale_fac is
 * negative there and its minimiser runs on a sign-flipped objective, so no meaningful value exists to reproduce.
 * Every non-zero status comes with NaN in the first five fields (the reference's failure tuple, :678).
 * A problem whose solve ends with status 1 or 3 is solved a second time by path following over eight barrier parameters
 * (the reference retries a failed trust-constr solve with SLSQP, :671-678); only what fails twice keeps its status. */
int scs_lrt_poisson_batch(scs_ctx* ctx, int32_t n_prob, const int64_t* br_off, const double* Y, const double* w,
                          const int32_t* child_int, double* out, double* x_out);
/* The same batch as a reusable, device-resident plan (topology and workspace are
 * uploaded once).  With a gather index the counts are taken from a per-node vector
 * (e.g. the placement's soft weights, still on the device): Y[k] = d_Y[index[k]]. */
int scs_lrt_plan_create(scs_ctx* ctx, int32_t n_prob, const int64_t* br_off, const int32_t* child_int,
                        scs_lrt_plan** out);
int scs_lrt_plan_destroy(scs_lrt_plan* plan);
int scs_lrt_plan_set_weights(scs_lrt_plan* plan, const double* w);
int scs_lrt_plan_set_gather(scs_lrt_plan* plan, const int32_t* index, int64_t src_len);
/* Asynchronous on `stream`; d_out ([n_prob][8]) and d_x may be NULL (results stay in the plan). */
int scs_lrt_plan_run(scs_lrt_plan* plan, const double* d_Y, double* d_out, double* d_x, void* stream);
int scs_lrt_plan_run_host(scs_lrt_plan* plan, const double* Y, double* out, double* x_out, void* stream);
const void* scs_lrt_plan_device_result(scs_lrt_plan* plan);

/* Placeholder-topology plan of n_prob problems with n_br branches each, and the device buffers a producer kernel
 * (scs_sweep_model) fills in place of scs_lrt_plan_set_weights / set_gather / the topology upload. */
int scs_lrt_plan_create_uniform(scs_ctx* ctx, int32_t n_prob, int32_t n_br, scs_lrt_plan** out);
int scs_lrt_plan_device_inputs(scs_lrt_plan* plan, int64_t src_len, int32_t** d_child, int32_t** d_gather, double** d_w);

/* ------------------------------------------------------------ replicate sweeps
 * The reference runs run_poisson_tree_test (run_PT_test.py:703-728) once per (VCF, tree) pair; a simulation sweep is
 * thousands of pairs over the same samples.  These entry points do the per-pair work of read_tree, add_br_weights and
 * get_model_data for ALL pairs of a batch at once, so that a sweep costs a handful of calls instead of a Python loop.
 *
 * scs_tree_parse_batch (host only, `threads` host threads, 0 = as many as there are cores up to 16): read_tree
 * (:210-222, :281-285, :296-313) for CellPhy / plain Newick texts -- clean-up, cell -> tumcell / outgcell -> healthycell
 * renames, `[n]` tags dropped when strip_tags != 0 (CellPhy files), Newick parsing, re-rooting on `outgroup`, its
 * removal, removal of leaves that are not samples -- and the rows of S (:389-397).  Texts and sample names are passed
 * as one character buffer plus offsets ([n + 1] each).  Per tree t: bits[t][n_rows][pitch_words] = membership masks over
 * the sample columns, one row per node of tree.iter_descendants() (level order), zero rows after them;
 * children[t][n_rows][2] = rows of a node's two children (-1 -1: leaf or unused row); info[t][4] = {nodes, leaves, 0, 0}.
 * SCS_ERR_PARSE names the first tree that is malformed, lacks the outgroup leaf or is not binary. */
int scs_tree_parse_batch(int32_t n_trees, const char* text, const int64_t* text_off, int32_t strip_tags, int32_t n_samples,
                         const char* sample_names, const int64_t* name_off, const char* outgroup, int32_t n_rows,
                         int64_t pitch_words, uint32_t* bits, int32_t* children, int32_t* info, int32_t threads);
/* scs_gt_reduce for a batch of datasets stacked in one device matrix (<= 512 cells): group g owns rows
 * [d_group_row0[g], d_group_row0[g + 1]) (device array); d_n_kept[g] and d_missing[g][n_cells] (counts, not fractions)
 * stay on the device.  d_active_mask: device, ceil(n_cells / 16) words, bit 2 (c % 16) of word c / 16 set for every
 * active (non-outgroup) cell c.  max_rows: the largest group (sizes the grid).  Asynchronous on `stream`. */
int scs_gt_reduce_groups(scs_ctx* ctx, const uint8_t* d_gt, int64_t pitch_bytes, int32_t n_groups, const int64_t* d_group_row0,
                         int64_t max_rows, int32_t n_cells, const uint32_t* d_active_mask, int32_t filter_rows, uint8_t* d_row_keep,
                         uint64_t* d_n_kept, uint32_t* d_missing, void* stream);
/* add_br_weights (:451-515) and get_model_data's branch order (:569-611) for every group, on the device, written into
 * the inputs of `plan` (created with scs_lrt_plan_create_uniform(n_groups * n_w, n_br)): problem g * n_w + w is group g
 * at w_max[w].  d_bits / d_children / d_info: device copies of scs_tree_parse_batch's outputs; d_n_kept / d_missing: from
 * scs_gt_reduce_groups; d_soft: [n_groups][n_rows] soft weights of the batched placement; FP, FN: host, one per group;
 * d_status[g] != 0 flags a group whose tree has no root row or not n_br / 2 + 1 leaves.  d_workspace: caller-owned
 * device scratch of scs_sweep_model_workspace(n_groups, n_rows, n_w) bytes, 8-byte aligned (one per concurrent
 * call).  Asynchronous on `stream`. */
int64_t scs_sweep_model_workspace(int32_t n_groups, int32_t n_rows, int32_t n_w);
int scs_sweep_model(scs_ctx* ctx, int32_t n_groups, int32_t n_rows, int32_t n_cells, int64_t pitch_words, const uint32_t* d_bits,
                    const int32_t* d_children, const int32_t* d_info, const uint64_t* d_n_kept, const uint32_t* d_missing,
                    const double* FP, const double* FN, const double* w_max, int32_t n_w, const double* d_soft, int32_t n_br,
                    scs_lrt_plan* plan, int32_t* d_status, void* d_workspace, void* stream);
/* The model line of the reference's TSV (:722-724) for every group, host only: "{muts}\t{FN:.4f}\t{FP:.4f}" and, per
 * w_max, "\t{LR:0>5.3f}\t{dof}\t{p_val}\tH{hyp}" with Python's number formatting.  out: [n_groups][n_w][8] as
 * scs_lrt_plan_run writes it.  Rows are written back to back into buf; row g is buf[row_off[g] .. row_off[g + 1]). */
int scs_format_pt_rows(int32_t n_groups, int32_t n_w, const double* out, const int64_t* n_kept, const double* FN, const double* FP,
                       char* buf, int64_t buf_len, int64_t* row_off);

#ifdef __cplusplus
}
#endif
#endif /* SCSOMMERCLOCK_B200_H */
