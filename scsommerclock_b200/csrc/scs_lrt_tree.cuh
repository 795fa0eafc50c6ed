// scs_lrt_tree.cuh -- the clock-constrained Poisson MLE + LRT (run_PT_test.py:638-700) for big trees:
// one thread-block CLUSTER per (dataset, w_max) problem.  Included by scs_lrt.cu (same arithmetic as scs_lrt_kernel:
// Newton on the internal-node heights of the barrier problem, tree-structured elimination for the Newton system).
//
// Why a second kernel: with one CTA per problem and the working set in global memory, a 10,000-leaf tree spent ~190 us
// per Newton iteration, almost all of it in the 2 x 33 level-synchronous steps of the elimination (every level = one
// block barrier + three dependent L2 round trips), on 10 of 148 SMs.  Here
//   * the topology is prepared on the host when the plan is made: internal nodes renumbered in LEVEL ORDER (level = 1 +
//     max level of the internal children), so a level is a contiguous index range and children have smaller indices;
//   * everything above a cut level ("top region": all levels >= qc, at most ~7,000 nodes = every level but the lowest of a
//     10,000-leaf coalescent tree) is eliminated out of the LEADER CTA's shared memory -- (e, 1/pivot, rhs) + 16-bit
//     relative child/parent links, 32 bytes per node; the rows are produced by all CTAs into global memory (L2) and the
//     leader pulls its region in with batched loads (writing them straight into the leader's shared memory through
//     DSMEM was measured slower: ~20 B/clk of remote stores into one SM).  Levels with <= 32 nodes (the long thin upper
//     part of a coalescent tree) are walked by ONE warp with warp barriers only;
//   * the per-branch / per-node passes (barrier gradient and curvature, step bound, line search) are split over the
//     CTAs of the cluster (8 for <= 18 problems); cluster-wide sums go through per-CTA slots in global memory and one
//     hardware cluster barrier each, in a fixed order (deterministic);
//   * level-0 nodes (no internal child) are eliminated where they are produced; the levels
//     between 0 and the cut (only for trees beyond ~10,800 leaves) are swept cluster-wide in global memory.
// The iterate is double-buffered (h_in / h_out) and the step is applied lazily (h = h_in + a * x when read), which
// saves one cluster barrier per iteration.
#pragma once
#include <cooperative_groups.h>

namespace scs {

namespace cg = cooperative_groups;

constexpr int LRT_TREE_CMAX = 8;          // CTAs per cluster
constexpr int LRT_TREE_LEVSM = 1024;      // level offsets kept in shared memory up to this many levels (+1)
constexpr int LRT_TREE_SLOT_DOUBLES = 2 * LRT_TREE_CMAX * 2;   // [parity][rank][2] per problem

struct LrtTreeProblem {
    int br_off;      // offset into the per-branch arrays
    int n_br;
    int node_off;    // offset into the per-node arrays
    int nlev;        // number of levels
    int lev_off;     // offset of this problem's nlev + 1 level offsets
    int qc;          // first level of the top region (>= 1)
    int qw;          // first level from which every level has <= 32 nodes (>= qc)
    int s_cut;       // first node of the top region = levoff[qc]
};

struct LrtTreeArgs {
    const LrtTreeProblem* probs;
    const double* Ysrc;      // counts, gathered through `gather`
    const int* gather;       // [branch] index into Ysrc (level-order branch numbering)
    const double* w;         // [branch] weights (level-order numbering)
    const int* child;        // [branch] internal node below the branch, or -1
    const int* up;           // [node] branch above the node, or -1 (root)
    const int* orig;         // [branch] the caller's branch index
    const int* levoff;
    double *Y, *wq, *l, *dl;                 // per branch
    double *cD, *cR;                         // per branch, zero where the child is a leaf or in the top region (never written there)
    double *hA, *hB, *x, *G, *e, *D, *r;     // per node
    double* slots;
    double* out;
    double* x_out;
    double mu;
    int max_iter;
    int start_mode;          // 0: max-based feasible start (scs_lrt_kernel's), 1: mean-based
    long long* prof;         // tools only (SCS_LRT_PROF): per-phase clock cycles of problem 0's leader, thread 0; nullptr otherwise
};

template <int T>
struct LrtTreeCtx {
    cg::cluster_group cl;
    int C, rank, par;
    double* slots;
    double* sh;
};

// CTA-wide then cluster-wide: a summed, b summed (B_MIN = false) or minimised.  Contains one cluster barrier when C > 1
// (and block barriers always), so it also orders everything written before it against everything read after it.
template <int T, bool B_MIN>
__device__ __forceinline__ void tree_reduce2(LrtTreeCtx<T>& cx, double& a, double& b) {
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        const double t = __shfl_xor_sync(0xffffffffu, b, o);
        b = B_MIN ? fmin(b, t) : b + t;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) {
        cx.sh[2 * warp] = a;
        cx.sh[2 * warp + 1] = b;
    }
    __syncthreads();
    a = cx.sh[0];
    b = cx.sh[1];
#pragma unroll
    for (int i = 1; i < T / 32; ++i) {
        a += cx.sh[2 * i];
        b = B_MIN ? fmin(b, cx.sh[2 * i + 1]) : b + cx.sh[2 * i + 1];
    }
    if (cx.C == 1) return;
    double* s = cx.slots + size_t(cx.par) * LRT_TREE_CMAX * 2;
    if (threadIdx.x == 0) {
        s[2 * cx.rank] = a;
        s[2 * cx.rank + 1] = b;
    }
    cx.cl.sync();
    a = __ldcg(s);
    b = __ldcg(s + 1);
    for (int i = 1; i < cx.C; ++i) {
        a += __ldcg(s + 2 * i);
        const double t = __ldcg(s + 2 * i + 1);
        b = B_MIN ? fmin(b, t) : b + t;
    }
    cx.par ^= 1;
}

template <int T>
__device__ __forceinline__ void tree_sync(LrtTreeCtx<T>& cx) {
    if (cx.C == 1) __syncthreads();
    else cx.cl.sync();
}

// barrier gradient and curvature of one branch of length lk (scs_lrt_kernel's expressions)
__device__ __forceinline__ void tree_branch(double lk, double wy, double wq, double U, double mu, double& g, double& d) {
    const double il = tree_rcp(lk), ia = tree_rcp(lk - LRT_LMIN), ib = tree_rcp(U - lk);
    g = wq - wy * il + mu * (ib - ia);
    d = wy * il * il + mu * (ia * ia + ib * ib);
}

__device__ __forceinline__ double tree_back(double r, double e, double D, double xp) { return (r + e * xp) * D; }

// phase timer for tools/lrt_timing.py: adds the cycles since the last mark to slot `i`
#define LRT_MARK(i)                                  \
    if (prof) {                                      \
        const long long _t = clock64();              \
        prof[i] += _t - t_mark;                      \
        t_mark = _t;                                 \
    }

template <int T>
__global__ void __launch_bounds__(T, 1) scs_lrt_tree_kernel(LrtTreeArgs A) {
    __shared__ double sh[2 * (T / 32)];
    extern __shared__ double lrt_tree_dyn[];
    LrtTreeCtx<T> cx{cg::this_cluster(), 1, 0, 0, nullptr, sh};
    cx.C = int(cx.cl.num_blocks());
    cx.rank = int(cx.cl.block_rank());
    const int C = cx.C, rank = cx.rank;
    const int p = blockIdx.x / C;
    cx.slots = A.slots + size_t(p) * LRT_TREE_SLOT_DOUBLES;
    const LrtTreeProblem pb = A.probs[p];
    const int nb = pb.n_br, m1 = nb >> 1, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s_cut = pb.s_cut, ntop = m1 - s_cut, nlev = pb.nlev, qc = pb.qc, qw = pb.qw;
    const int* __restrict__ child = A.child + pb.br_off;
    const int* __restrict__ up = A.up + pb.node_off;
    const int* __restrict__ gath = A.gather + pb.br_off;
    const double* __restrict__ wraw = A.w + pb.br_off;
    double *Y = A.Y + pb.br_off, *wq = A.wq + pb.br_off, *l = A.l + pb.br_off, *dl = A.dl + pb.br_off;
    double *hA = A.hA + pb.node_off, *hB = A.hB + pb.node_off, *x = A.x + pb.node_off, *G = A.G + pb.node_off;
    double *e = A.e + pb.node_off, *D = A.D + pb.node_off, *r = A.r + pb.node_off;
    double *cD = A.cD + pb.br_off, *cR = A.cR + pb.br_off;      // Schur contributions of nodes below the cut, per up branch
    constexpr int NB = 3;                                        // nodes per thread and batch in the per-node passes
    // shared memory of the top region (used in the leader CTA; the others hold the same layout unused)
    double* e_t = lrt_tree_dyn;
    double* D_t = e_t + ntop;
    double* r_t = D_t + ntop;
    uint2* topo_t = reinterpret_cast<uint2*>(r_t + ntop);          // {child0 | child1 << 16, parent | flags << 16}: top-relative, 0xFFFF = none
    int* lev_s = reinterpret_cast<int*>(topo_t + ntop);
    const bool lev_in_smem = nlev + 1 <= LRT_TREE_LEVSM;
    const int* levoff = A.levoff + pb.lev_off;
    if (lev_in_smem) {
        for (int q = tid; q <= nlev; q += T) lev_s[q] = levoff[q];
        levoff = lev_s;
    }
    if (rank == 0) {
        for (int ip = tid; ip < ntop; ip += T) {
            const int i = s_cut + ip;
            const int c0 = child[2 * i], c1 = child[2 * i + 1], u = up[i];
            const uint32_t a0 = c0 >= s_cut ? uint32_t(c0 - s_cut) : 0xFFFFu, a1 = c1 >= s_cut ? uint32_t(c1 - s_cut) : 0xFFFFu;
            const uint32_t fl = ((c0 >= 0 && c0 < s_cut) ? 1u : 0u) | ((c1 >= 0 && c1 < s_cut) ? 2u : 0u);      // children below the cut
            topo_t[ip] = make_uint2(a0 | (a1 << 16), (u >= 0 ? uint32_t((u >> 1) - s_cut) : 0xFFFFu) | (fl << 16));
        }
    }
    __syncthreads();
    const int n0 = levoff[1];                                      // level-0 nodes: [0, n0)
    const int s0 = int((long long)m1 * rank / C), s1 = int((long long)m1 * (rank + 1) / C);     // this CTA's nodes / branches 2 s0 .. 2 s1
    double* res = A.out + size_t(p) * 8;

    // ---- sums, scale factor (run_PT_test.py:639-643)
    double U = 0.0, ll_H1 = 0.0, wsum = 0.0, zero = 0.0;
    for (int k = 2 * s0 + tid; k < 2 * s1; k += T) {
        const double y = A.Ysrc[gath[k]], wk = wraw[k];
        Y[k] = y;
        U += y;
        ll_H1 += (y * log(y > 0.0 ? y : 1.0) - y) * wk;
        wsum += wk;
    }
    tree_reduce2<T, false>(cx, U, ll_H1);
    tree_reduce2<T, false>(cx, wsum, zero);
    double sf = 0.0;
    {
        const double steps[6] = {10000.0, 1000.0, 100.0, 10.0, 1.0, 0.1};
        for (int i = 0; i < 6; ++i) {
            sf = floor(ll_H1 / steps[i]) * steps[i];
            if (sf != 0.0) break;
        }
    }
    {
        const bool degenerate = !(sf != 0.0) || !(U > 2.0 * nb * LRT_LMIN) || !isfinite(ll_H1);
        if (degenerate || sf < 0.0) {             // every CTA of the cluster takes this branch together; statuses as in scs_lrt_kernel
            if (tid == 0 && rank == 0) {
                for (int i = 0; i < 7; ++i) res[i] = NAN;
                if (!degenerate) res[5] = ll_H1;
                res[7] = degenerate ? 2.0 : 4.0;
            }
            return;
        }
    }

    // level q's share of this CTA (bottom levels are swept cluster-wide)
    auto level_range = [&](int q, int& a0, int& a1) {
        const int b0 = levoff[q], b1 = levoff[q + 1];
        a0 = b0 + int((long long)(b1 - b0) * rank / C);
        a1 = b0 + int((long long)(b1 - b0) * (rank + 1) / C);
    };

    // ---- strictly feasible start: heights from the counts themselves, level by level
    {
        const double dmin = 1e-3;
        // start_mode 0: h = max over children of (child height + max(Y, dmin)) -- every branch at least as long as its count;
        // start_mode 1: the MEAN of the two instead of the max (closer to the clock optimum, whose heights are weighted
        // means of the path sums, not their maxima), kept feasible by the floor max(child heights) + dmin
        const bool mean_start = A.start_mode == 1;
        auto start_h = [&](double a0, double a1, double floor_h) { return mean_start ? fmax(0.5 * (a0 + a1), floor_h) : fmax(a0, a1); };
        for (int s = s0 + tid; s < min(s1, n0); s += T) hA[s] = start_h(fmax(Y[2 * s], dmin), fmax(Y[2 * s + 1], dmin), dmin);
        tree_sync(cx);
        for (int q = 1; q < qc; ++q) {
            int a0, a1;
            level_range(q, a0, a1);
            for (int i = a0 + tid; i < a1; i += T) {
                const int j0 = child[2 * i], j1 = child[2 * i + 1];
                const double hc0 = j0 >= 0 ? hA[j0] : 0.0, hc1 = j1 >= 0 ? hA[j1] : 0.0;
                hA[i] = start_h(hc0 + fmax(Y[2 * i], dmin), hc1 + fmax(Y[2 * i + 1], dmin), fmax(hc0, hc1) + dmin);
            }
            tree_sync(cx);
        }
        if (rank == 0) {
            // e_t / D_t: child height (if the child is below the cut; leaves 0) + max(Y, dmin) of the two child branches;
            // r_t: the floor from the children below the cut, replaced by the node's height when its level is reached
            for (int ip = tid; ip < ntop; ip += T) {
                const int i = s_cut + ip;
                const int c0 = child[2 * i], c1 = child[2 * i + 1];
                const double hc0 = (c0 >= 0 && c0 < s_cut) ? hA[c0] : 0.0, hc1 = (c1 >= 0 && c1 < s_cut) ? hA[c1] : 0.0;
                e_t[ip] = fmax(Y[2 * i], dmin) + hc0;
                D_t[ip] = fmax(Y[2 * i + 1], dmin) + hc1;
                r_t[ip] = fmax(hc0, hc1) + dmin;
            }
            __syncthreads();
            auto node_h = [&](int ip) {
                const uint32_t cc = topo_t[ip].x;
                const uint32_t c0 = cc & 0xFFFFu, c1 = cc >> 16;
                const double hc0 = c0 != 0xFFFFu ? r_t[c0] : 0.0, hc1 = c1 != 0xFFFFu ? r_t[c1] : 0.0;
                r_t[ip] = start_h(e_t[ip] + hc0, D_t[ip] + hc1, fmax(r_t[ip], fmax(hc0, hc1) + dmin));
            };
            for (int q = qc; q < qw; ++q) {
                for (int ip = levoff[q] - s_cut + tid; ip < levoff[q + 1] - s_cut; ip += T) node_h(ip);
                __syncthreads();
            }
            if (warp == 0)
                for (int q = qw; q < nlev; ++q) {
                    const int ip = levoff[q] - s_cut + lane;
                    if (ip < levoff[q + 1] - s_cut) node_h(ip);
                    __syncwarp();
                }
            __syncthreads();
            for (int ip = tid; ip < ntop; ip += T) hA[s_cut + ip] = r_t[ip];
        }
        tree_sync(cx);
        double lmaxp = 0.0, z = 0.0;
        for (int k = 2 * s0 + tid; k < 2 * s1; k += T) {
            const int j = child[k];
            lmaxp = fmax(lmaxp, hA[k >> 1] - (j >= 0 ? hA[j] : 0.0));
        }
        lmaxp = -lmaxp;
        tree_reduce2<T, true>(cx, z, lmaxp);
        const double lmax = -lmaxp;
        const double sc = lmax >= 0.9 * U ? 0.9 * U / lmax : 1.0;
        if (sc != 1.0)
            for (int s = s0 + tid; s < s1; s += T) hA[s] = hA[s] * sc;
        // from here on the objective only needs w/scale_fac and w*Y/scale_fac
        const double isf = 1.0 / sf;
        for (int k = 2 * s0 + tid; k < 2 * s1; k += T) {
            const double wk = wraw[k] * isf;
            wq[k] = wk;
            Y[k] = wk * Y[k];
        }
        tree_sync(cx);
    }

    const double mu = A.mu;
    double *hin = hA, *hout = hB;
    bool use_x = false;          // the iterate is hin + a_prev * x (the step of the previous iteration, applied when read)
    double a_prev = 0.0;
    auto hcur = [&](int v) { return use_x ? fma(a_prev, x[v], hin[v]) : hin[v]; };
    int it = 0, status = 1;
    double f0 = 0.0;
    long long* prof = (A.prof && p == 0 && rank == 0 && tid == 0) ? A.prof : nullptr;
    long long t_mark = prof ? clock64() : 0;
    for (; it < A.max_iter; ++it) {
        // ---- per node: lengths of its two child branches and of its up branch, barrier gradient / curvature, the node's
        // row of the Newton system.  Nodes below the cut publish their Schur contribution to the parent's row in the
        // per-branch slots cD / cR (level 0 right here, the other levels when their own row is final).  NB nodes per
        // thread at a time, phase by phase: index loads, then the height loads that depend on them, then arithmetic --
        // three nodes in sequence would be three times two dependent L2 round trips.
        double f0p = 0.0;
        for (int sb = s0 + tid; sb < s1; sb += NB * T) {
            int u[NB];
            int2 ch[NB];
            double hs[NB], hp[NB], h0[NB], h1[NB], yu[NB], wu[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int s = sb + b * T;
                const bool v = s < s1;
                u[b] = v ? up[s] : -1;
                ch[b] = v ? *reinterpret_cast<const int2*>(child + 2 * s) : make_int2(-1, -1);
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int s = sb + b * T;
                const bool v = s < s1;
                hs[b] = v ? hcur(s) : 1.0;
                hp[b] = u[b] >= 0 ? hcur(u[b] >> 1) : 0.0;
                h0[b] = ch[b].x >= 0 ? hcur(ch[b].x) : 0.0;
                h1[b] = ch[b].y >= 0 ? hcur(ch[b].y) : 0.0;
                yu[b] = u[b] >= 0 ? Y[u[b]] : 0.0;
                wu[b] = u[b] >= 0 ? wq[u[b]] : 0.0;
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int s = sb + b * T;
                if (s >= s1) continue;
                hout[s] = hs[b];
                const double2 wy2 = *reinterpret_cast<const double2*>(Y + 2 * s), wk2 = *reinterpret_cast<const double2*>(wq + 2 * s);
                const double lk0 = hs[b] - h0[b], lk1 = hs[b] - h1[b];
                double g0, d0, g1, d1;
                tree_branch(lk0, wy2.x, wk2.x, U, mu, g0, d0);
                tree_branch(lk1, wy2.y, wk2.y, U, mu, g1, d1);
                *reinterpret_cast<double2*>(l + 2 * s) = make_double2(lk0, lk1);
                double gs = g0 + g1, ds = d0 + d1;
                if (it == 0) {
                    f0p += wk2.x * lk0 - wy2.x * log(lk0) - mu * log((lk0 - LRT_LMIN) * (U - lk0));
                    f0p += wk2.y * lk1 - wy2.y * log(lk1) - mu * log((lk1 - LRT_LMIN) * (U - lk1));
                }
                double eu = 0.0;
                if (u[b] >= 0) {
                    double g, d;
                    tree_branch(hp[b] - hs[b], yu[b], wu[b], U, mu, g, d);
                    gs -= g;
                    ds += d;
                    eu = d;
                }
                G[s] = gs;
                e[s] = eu;
                r[s] = -gs;
                if (s < n0) {
                    const double Di = tree_rcp(ds);
                    D[s] = Di;
                    if (u[b] >= 0) {
                        const double t = eu * Di;
                        cD[u[b]] = eu * t;
                        cR[u[b]] = t * -gs;
                    }
                } else {
                    D[s] = ds;
                }
            }
        }
        LRT_MARK(0)
        if (it == 0) {
            double z = 0.0;
            tree_reduce2<T, false>(cx, f0p, z);
            f0 = f0p;
        } else {
            tree_sync(cx);
        }
        LRT_MARK(1)
        // ---- leaf-to-root elimination: levels below the cut cluster-wide in global memory ...
        for (int q = 1; q < qc; ++q) {
            int a0, a1;
            level_range(q, a0, a1);
            for (int i = a0 + tid; i < a1; i += T) {
                const double2 c2 = *reinterpret_cast<const double2*>(cD + 2 * i), r2 = *reinterpret_cast<const double2*>(cR + 2 * i);
                const double Di = tree_rcp((D[i] - c2.x) - c2.y), rv = (r[i] + r2.x) + r2.y;
                D[i] = Di;
                r[i] = rv;
                const int ui = up[i];
                if (ui >= 0) {
                    const double ei = e[i], t = ei * Di;
                    cD[ui] = ei * t;
                    cR[ui] = t * rv;
                }
            }
            tree_sync(cx);
        }
        LRT_MARK(2)
        // ... the top region in the leader's shared memory, both directions
        if (rank == 0) {
            // pull the region's rows in from global memory, FB nodes per thread at a time (all loads of a batch in flight
            // together), and fold in the contributions of children below the cut, which are final
            constexpr int FB = 4;
            for (int ib = tid; ib < ntop; ib += FB * T) {
                double ev[FB], Dv[FB], rv[FB], a0[FB], a1[FB], b0[FB], b1[FB];
#pragma unroll
                for (int b = 0; b < FB; ++b) {
                    const int ip = ib + b * T;
                    const bool v = ip < ntop;
                    const int i = s_cut + (v ? ip : 0);
                    const uint32_t fl = v ? topo_t[ip].y >> 16 : 0u;
                    ev[b] = v ? e[i] : 0.0;
                    Dv[b] = v ? D[i] : 0.0;
                    rv[b] = v ? r[i] : 0.0;
                    a0[b] = (fl & 1u) ? cD[2 * i] : 0.0;
                    b0[b] = (fl & 1u) ? cR[2 * i] : 0.0;
                    a1[b] = (fl & 2u) ? cD[2 * i + 1] : 0.0;
                    b1[b] = (fl & 2u) ? cR[2 * i + 1] : 0.0;
                }
#pragma unroll
                for (int b = 0; b < FB; ++b) {
                    const int ip = ib + b * T;
                    if (ip < ntop) {
                        e_t[ip] = ev[b];
                        D_t[ip] = (Dv[b] - a0[b]) - a1[b];
                        r_t[ip] = (rv[b] + b0[b]) + b1[b];
                    }
                }
            }
            __syncthreads();
            LRT_MARK(3)
            auto node_up = [&](int ip) {
                const uint32_t cc = topo_t[ip].x;
                const uint32_t c0 = cc & 0xFFFFu, c1 = cc >> 16;
                double Dv = D_t[ip], rv = r_t[ip];
                if (c0 != 0xFFFFu) {
                    const double ej = e_t[c0], t = ej * D_t[c0];
                    Dv -= ej * t;
                    rv += t * r_t[c0];
                }
                if (c1 != 0xFFFFu) {
                    const double ej = e_t[c1], t = ej * D_t[c1];
                    Dv -= ej * t;
                    rv += t * r_t[c1];
                }
                D_t[ip] = tree_rcp(Dv);
                r_t[ip] = rv;
            };
            auto node_down = [&](int ip) {
                const uint32_t pr = topo_t[ip].y & 0xFFFFu;
                r_t[ip] = tree_back(r_t[ip], e_t[ip], D_t[ip], pr != 0xFFFFu ? r_t[pr] : 0.0);
            };
            for (int q = qc; q < qw; ++q) {
                for (int ip = levoff[q] - s_cut + tid; ip < levoff[q + 1] - s_cut; ip += T) node_up(ip);
                __syncthreads();
            }
            LRT_MARK(4)
            if (warp == 0) {
                for (int q = qw; q < nlev; ++q) {
                    const int ip = levoff[q] - s_cut + lane;
                    if (ip < levoff[q + 1] - s_cut) node_up(ip);
                    __syncwarp();
                }
                for (int q = nlev - 1; q >= qw; --q) {
                    const int ip = levoff[q] - s_cut + lane;
                    if (ip < levoff[q + 1] - s_cut) node_down(ip);
                    __syncwarp();
                }
            }
            __syncthreads();
            LRT_MARK(5)
            for (int q = qw - 1; q >= qc; --q) {
                for (int ip = levoff[q] - s_cut + tid; ip < levoff[q + 1] - s_cut; ip += T) node_down(ip);
                __syncthreads();
            }
            LRT_MARK(6)
            for (int ip = tid; ip < ntop; ip += T) x[s_cut + ip] = r_t[ip];
        }
        tree_sync(cx);
        LRT_MARK(7)
        // ... and back down through the levels below the cut (level 0 is done where its x is needed)
        for (int q = qc - 1; q >= 1; --q) {
            int a0, a1;
            level_range(q, a0, a1);
            for (int i = a0 + tid; i < a1; i += T) {
                const int ui = up[i];
                x[i] = tree_back(r[i], e[i], D[i], ui >= 0 ? x[ui >> 1] : 0.0);
            }
            tree_sync(cx);
        }
        LRT_MARK(8)
        // ---- Newton decrement and the largest step that stays strictly inside the bounds (same batching as above)
        double dec = 0.0, amax = 1.0;
        for (int sb = s0 + tid; sb < s1; sb += NB * T) {
            int u[NB];
            int2 ch[NB];
            double xs[NB], x0[NB], x1[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int s = sb + b * T;
                const bool v = s < s1;
                u[b] = (v && s < n0) ? up[s] : -1;
                ch[b] = v ? *reinterpret_cast<const int2*>(child + 2 * s) : make_int2(-1, -1);
            }
            double rj0[NB], ej0[NB], Dj0[NB], rj1[NB], ej1[NB], Dj1[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int s = sb + b * T;
                const bool v = s < s1;
                // level-0 node: x from its own row and the parent's x; above: already in x
                xs[b] = !v ? 0.0 : (s < n0 ? (u[b] >= 0 ? x[u[b] >> 1] : 0.0) : x[s]);
                const int j0 = ch[b].x, j1 = ch[b].y;
                x0[b] = j0 >= n0 ? x[j0] : 0.0;
                x1[b] = j1 >= n0 ? x[j1] : 0.0;
                const bool f0 = j0 >= 0 && j0 < n0, f1 = j1 >= 0 && j1 < n0;
                rj0[b] = f0 ? r[j0] : 0.0;
                ej0[b] = f0 ? e[j0] : 0.0;
                Dj0[b] = f0 ? D[j0] : 0.0;
                rj1[b] = f1 ? r[j1] : 0.0;
                ej1[b] = f1 ? e[j1] : 0.0;
                Dj1[b] = f1 ? D[j1] : 0.0;
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const int s = sb + b * T;
                if (s >= s1) continue;
                double xv = xs[b];
                if (s < n0) {
                    xv = tree_back(r[s], e[s], D[s], xv);
                    x[s] = xv;
                }
                dec -= G[s] * xv;
                const int j0 = ch[b].x, j1 = ch[b].y;
                const double xc0 = (j0 >= 0 && j0 < n0) ? tree_back(rj0[b], ej0[b], Dj0[b], xv) : x0[b];      // leaf: 0
                const double xc1 = (j1 >= 0 && j1 < n0) ? tree_back(rj1[b], ej1[b], Dj1[b], xv) : x1[b];
                const double dk0 = xv - xc0, dk1 = xv - xc1;
                *reinterpret_cast<double2*>(dl + 2 * s) = make_double2(dk0, dk1);
                const double2 l2 = *reinterpret_cast<const double2*>(l + 2 * s);
                if (dk0 != 0.0) amax = fmin(amax, 0.99 * (dk0 < 0.0 ? l2.x - LRT_LMIN : U - l2.x) / fabs(dk0));
                if (dk1 != 0.0) amax = fmin(amax, 0.99 * (dk1 < 0.0 ? l2.y - LRT_LMIN : U - l2.y) / fabs(dk1));
            }
        }
        LRT_MARK(9)
        tree_reduce2<T, true>(cx, dec, amax);
        LRT_MARK(10)
        if (!(dec == dec) || dec < -1e-9 * fmax(1.0, fabs(f0))) {
            status = 3;
            a_prev = 0.0;
            use_x = false;
            double* t = hin; hin = hout; hout = t;
            break;
        }
        if (dec <= 1e-12 * fmax(1.0, fabs(f0))) {       // converged: take the step and stop (see scs_lrt_kernel)
            a_prev = amax;
            use_x = true;
            double* t = hin; hin = hout; hout = t;
            status = 0;
            ++it;
            break;
        }
        // backtracking (Armijo) on the barrier objective
        double a = amax, f1 = f0;
        bool ok = false;
        for (int bt = 0; bt < 40; ++bt) {
            double fp = 0.0, z = 0.0;
            for (int k = 2 * s0 + tid; k < 2 * s1; k += T) {
                const double lk = l[k] + a * dl[k];
                fp += wq[k] * lk - Y[k] * log(lk) - mu * log((lk - LRT_LMIN) * (U - lk));
            }
            tree_reduce2<T, false>(cx, fp, z);
            f1 = fp;
            if (prof) prof[14] += 1;
            if (f1 <= f0 - 1e-4 * a * dec) {
                ok = true;
                break;
            }
            a *= 0.5;
        }
        {
            double* t = hin; hin = hout; hout = t;
        }
        if (!ok) {          // no step decreases the objective any more: at the numerical optimum
            a_prev = 0.0;
            use_x = false;
            status = 0;
            ++it;
            break;
        }
        a_prev = a;
        use_x = true;
        f0 = f1;
        LRT_MARK(11)
        if (prof) prof[15] += 1;
    }

    // ---- statistic and p-value (run_PT_test.py:683-695), from the unscaled counts and weights
    double ll_H0 = 0.0, sOB = 0.0;
    for (int k = 2 * s0 + tid; k < 2 * s1; k += T) {
        const int j = child[k];
        const double lk = hcur(k >> 1) - (j >= 0 ? hcur(j) : 0.0);
        const double y = A.Ysrc[gath[k]];
        ll_H0 += (y * log(lk) - lk) * wraw[k];
        sOB += (lk <= 1e-3) ? 1.0 : 0.0;
        if (A.x_out) A.x_out[pb.br_off + A.orig[pb.br_off + k]] = lk;
    }
    tree_reduce2<T, false>(cx, ll_H0, sOB);
    const int on_bound = int(sOB + 0.5);
    const double LR = -2.0 * (ll_H0 - ll_H1);
    const double dof = rint(wsum - double(m1));
    double num = 0.0, den = 0.0;
    const double lg_n = lgamma(double(on_bound) + 1.0);
    for (int k = rank * T + tid; k <= on_bound; k += C * T) {
        const double df = dof - double(k);
        if (df > 0.0) {
            double pv = gamma_q(0.5 * df, 0.5 * LR);
            if (on_bound > 0) pv = fmin(fmax(pv, 1e-100), 1.0);
            const double lw = lg_n - lgamma(double(k) + 1.0) - lgamma(double(on_bound - k) + 1.0) - double(on_bound) * 0.6931471805599453;
            const double wk = exp(lw);
            num += wk * pv;
            den += wk;
        }
    }
    tree_reduce2<T, false>(cx, num, den);
    if (tid == 0 && rank == 0) {
        res[0] = LR;
        res[1] = dof;
        res[2] = double(on_bound);
        res[3] = den > 0.0 ? num / den : NAN;
        res[4] = ll_H0;
        res[5] = ll_H1;
        res[6] = double(it);
        res[7] = double(status);
    }
}

}  // namespace scs
