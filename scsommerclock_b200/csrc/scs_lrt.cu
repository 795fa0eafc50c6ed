// scs_lrt.cu -- clock-constrained Poisson branch-length MLE + likelihood-ratio
// test (reference: run_PT_test.py:638-700, get_LRT_poisson) and branch weights
// (reference: run_PT_test.py:451-502, add_br_weights), batched on the GPU.
//
// The reference minimises  -sum_b w_b (Y_b log l_b - l_b) / scale_fac  over branch
// lengths l subject to  constr @ l = 0  (every root-to-tip path has the same
// length, built in get_model_data :518-635) and  LAMBDA_MIN <= l_b <= sum(Y)  with
// SciPy's trust-constr, an interior-point method that stops on the central path
// at barrier parameter ~0.1 * 0.2^10 (it terminates when the Lagrangian gradient
// drops below gtol = 1e-8).  The kernel solves that same barrier problem exactly:
// the equality constraints are eliminated by parametrising the ultrametric tree
// by its internal-node heights h (l_b = h_parent - h_child), which leaves a
// smooth strictly convex problem in m-1 unknowns whose Hessian has the sparsity
// of the tree, so each Newton system is solved in O(m) by leaf-to-root
// elimination.  One CTA per (dataset, w_max) problem.
//
// Branch k hangs below internal node k/2 (get_model_data adds the two children of
// each internal node, in post-order, as consecutive branches) and child_int[k] is
// the internal node below branch k, or -1 for a leaf.
#include <cmath>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "scs_internal.h"

namespace scs {

constexpr int LRT_MAX_THREADS = 512;   // CTA size: 128 for small trees, up to 512 for 10k-cell trees
#define LRT_THREADS (int(blockDim.x))
constexpr double LRT_LMIN = 1e-6;           // LAMBDA_MIN, run_PT_test.py:21
constexpr double LRT_MU = 0.1 * 1.024e-7;   // 0.1 * 0.2^10, see above
constexpr int LRT_MAX_ITER = 200;

struct LrtProblem {
    int br_off;      // offset into per-branch arrays
    int n_br;        // 2 * (leaves - 1)
    int node_off;    // offset into per-internal-node workspace
};

struct LrtWork {
    // per branch
    double *l, *g, *d, *dl, *wq;
    // per internal node
    double *h, *Hd, *e, *D, *r, *x, *G;
    int *up, *order, *level, *levoff;
};

// Thread group of one problem: a whole CTA (WARP = false; big trees) or a single warp
// (WARP = true; small trees, where CTA barriers were a quarter of the stall samples).
template <bool WARP>
__device__ __forceinline__ void grp_sync() {
    if (WARP) __syncwarp();
    else __syncthreads();
}

template <bool WARP>
__device__ __forceinline__ int grp_any(int pred) {
    if (WARP) return __any_sync(0xffffffffu, pred);
    return __syncthreads_or(pred);
}

// two reductions in one pass: a is summed, b is summed (B_MIN = false) or minimised (B_MIN = true)
template <bool WARP, bool B_MIN>
__device__ __forceinline__ void grp_reduce2(double& a, double& b, double* sh) {
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        const double t = __shfl_xor_sync(0xffffffffu, b, o);
        b = B_MIN ? fmin(b, t) : b + t;
    }
    if (WARP) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) {
        sh[2 * warp] = a;
        sh[2 * warp + 1] = b;
    }
    __syncthreads();
    a = sh[0];
    b = sh[1];
    for (int i = 1; i < (LRT_THREADS >> 5); ++i) {
        a += sh[2 * i];
        b = B_MIN ? fmin(b, sh[2 * i + 1]) : b + sh[2 * i + 1];
    }
}

template <bool WARP>
__device__ __forceinline__ double grp_sum(double v, double* sh) {
    double z = 0.0;
    grp_reduce2<WARP, false>(v, z, sh);
    return v;
}

// 1/d for the moderate, strictly positive magnitudes of this solver (lengths, slacks, pivots: no denormals, infinities or
// zeros): the hardware seed (~20 bits) and two Newton steps, ~1 ulp -- a third of the instructions of the IEEE division,
// which is what the single SM that eliminates the top region is bound by.
__device__ __forceinline__ double tree_rcp(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    return fma(y, e, y);
}

// Regularised upper incomplete gamma Q(a, x) = chi2.sf(2x, 2a): power series of
// P for x < a+1, Lentz continued fraction of Q otherwise.
__device__ double gamma_q(double a, double x) {
    if (!(a > 0.0)) return NAN;
    if (x <= 0.0) return 1.0;
    const double lg = lgamma(a);
    if (x < a + 1.0) {
        double ap = a, del = 1.0 / a, sum = del;
        for (int n = 0; n < 10000; ++n) {
            ap += 1.0;
            del *= x / ap;
            sum += del;
            if (fabs(del) < fabs(sum) * 1e-17) break;
        }
        return 1.0 - sum * exp(-x + a * log(x) - lg);
    }
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, dd = 1.0 / b, hh = dd;
    for (int i = 1; i < 10000; ++i) {
        const double an = -double(i) * (double(i) - a);
        b += 2.0;
        dd = an * dd + b;
        if (fabs(dd) < tiny) dd = tiny;
        c = b + an / c;
        if (fabs(c) < tiny) c = tiny;
        dd = 1.0 / dd;
        const double del = dd * c;
        hh *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return exp(-x + a * log(x) - lg) * hh;
}

// out[p*8 ..] = {LR, dof, on_bound, p_value, ll_H0, ll_H1, newton iterations, status}
template <bool WARP>
__global__ void __launch_bounds__(WARP ? 32 : LRT_MAX_THREADS)
scs_lrt_kernel(const LrtProblem* __restrict__ probs, const double* __restrict__ Ysrc, const int* __restrict__ gather,
               double* __restrict__ Yall, const double* __restrict__ wall, const int* __restrict__ child_all, LrtWork ws,
               double* __restrict__ out, double* __restrict__ x_out, int smem_mode, double mu_in, int n_stages, int max_iter) {
    __shared__ double sh[WARP ? 2 : 2 * (LRT_MAX_THREADS / 32)];
    __shared__ int s_nlev;
    const LrtProblem pb = probs[blockIdx.x];
    const int nb = pb.n_br, m1 = nb / 2, tid = threadIdx.x;
    extern __shared__ double lrt_dyn[];
    // Y: counts, later w*Y/scale_fac; wq: w/scale_fac; D holds 1/pivot
    double *Y, *wq, *l, *g, *d, *dl, *h, *Hd, *e, *D, *r, *x, *G;
    const double* wraw;
    const int* child;
    int *up, *order, *level, *levoff;
    if (smem_mode) {
        // small trees: the whole working set (a few KB) lives in shared memory.  Arrays whose
        // lifetimes do not overlap share storage (more resident problems per SM): g is dead once
        // the node gradient is formed and dl is written after the solve; d is dead once Hd/e are
        // formed and the pivots D and right-hand sides r are written after that; level is only
        // needed to build the level order, x only during the Newton solve.
        double* q = lrt_dyn;
        Y = q; q += nb;
        wq = q; q += nb;
        l = q; q += nb;
        g = dl = q; q += nb;
        const int mn = m1 + 2;
        d = D = q; q += mn;      // d[0..nb) spans D and r (2 * mn >= nb)
        r = q; q += mn;
        h = q; q += mn;
        Hd = q; q += mn;
        e = q; q += mn;
        x = q; q += mn;
        G = q; q += mn;
        int* iq = reinterpret_cast<int*>(q);
        int* cq = iq; iq += nb;
        up = iq; iq += mn;
        order = iq; iq += mn;
        levoff = iq;
        level = reinterpret_cast<int*>(x);
        for (int k = tid; k < nb; k += LRT_THREADS) {
            wq[k] = wall[pb.br_off + k];
            cq[k] = child_all[pb.br_off + k];
        }
        wraw = wq;
        child = cq;
    } else {
        Y = Yall + pb.br_off;
        wraw = wall + pb.br_off;
        wq = ws.wq + pb.br_off;
        child = child_all + pb.br_off;
        l = ws.l + pb.br_off; g = ws.g + pb.br_off; d = ws.d + pb.br_off; dl = ws.dl + pb.br_off;
        h = ws.h + pb.node_off; Hd = ws.Hd + pb.node_off; e = ws.e + pb.node_off; D = ws.D + pb.node_off;
        r = ws.r + pb.node_off; x = ws.x + pb.node_off; G = ws.G + pb.node_off;
        up = ws.up + pb.node_off; order = ws.order + pb.node_off; level = ws.level + pb.node_off; levoff = ws.levoff + pb.node_off;
    }
    // counts in branch order: straight copy, or a gather from per-node soft weights
    for (int k = tid; k < nb; k += LRT_THREADS) Y[k] = gather ? Ysrc[gather[pb.br_off + k]] : Ysrc[pb.br_off + k];
    grp_sync<WARP>();
    double* res = out + size_t(blockIdx.x) * 8;

    // ---- sums, scale factor (run_PT_test.py:639-643)
    double U = 0.0, ll_H1 = 0.0, wsum = 0.0, zero = 0.0;
    for (int k = tid; k < nb; k += LRT_THREADS) {
        const double y = Y[k];
        U += y;
        ll_H1 += (y * log(y > 0.0 ? y : 1.0) - y) * wraw[k];
        wsum += wraw[k];
    }
    grp_reduce2<WARP, false>(U, ll_H1, sh);
    grp_reduce2<WARP, false>(wsum, zero, sh);
    double sf = 0.0;
    {
        const double steps[6] = {10000.0, 1000.0, 100.0, 10.0, 1.0, 0.1};
        for (int i = 0; i < 6; ++i) {
            sf = floor(ll_H1 / steps[i]) * steps[i];
            if (sf != 0.0) break;
        }
    }
    if (!(sf != 0.0) || !(U > 2.0 * nb * LRT_LMIN) || !isfinite(ll_H1)) {
        if (tid == 0) {
            for (int i = 0; i < 7; ++i) res[i] = NAN;
            res[7] = 2.0;   // degenerate input: the reference divides by scale_fac == 0
        }
        return;
    }
    if (sf < 0.0) {
        // ll_H1 < 0 (very sparse counts, most Y < e): the reference's scale_fac = (ll_H1 // 10000) * 10000 = -10000 flips
        // the sign of its objective, so its "minimize" call MAXIMISES the negative log-likelihood and returns whatever
        // vertex of the feasible polytope trust-constr stops at (LR in the hundreds or thousands for a 30-leaf tree,
        // tests/golden/lrt_sweep.npz holds examples).  That number is an accident of the optimiser, not a statistic: it
        // is reported as a failure with its own status instead of being imitated or silently "converged".
        if (tid == 0) {
            for (int i = 0; i < 7; ++i) res[i] = NAN;
            res[5] = ll_H1;
            res[7] = 4.0;
        }
        return;
    }

    // ---- topology: up-branch of every internal node, levels, level order.  All of it in parallel:
    // with 10,000-leaf trees the obvious thread-0 loops were 5 of the kernel's 7 ms.
    for (int i = tid; i < m1; i += LRT_THREADS) {
        up[i] = -1;
        level[i] = 0;
    }
    grp_sync<WARP>();
    for (int k = tid; k < nb; k += LRT_THREADS)
        if (child[k] >= 0) up[child[k]] = k;
    // level = 1 + max(level of internal children): relax until nothing changes (<= depth + 1 sweeps;
    // children have smaller indices, and a sweep may already see this sweep's updates -- harmless)
    for (int sweep = 0; sweep <= m1; ++sweep) {
        int ch = 0;
        for (int i = tid; i < m1; i += LRT_THREADS) {
            int lv = 0;
            const int c0 = child[2 * i], c1 = child[2 * i + 1];
            if (c0 >= 0) lv = max(lv, level[c0] + 1);
            if (c1 >= 0) lv = max(lv, level[c1] + 1);
            if (lv != level[i]) {
                level[i] = lv;
                ch = 1;
            }
        }
        if (!grp_any<WARP>(ch)) break;   // also the barrier between sweeps
    }
    // counting sort by level: levoff[q] = first slot of level q in `order` (order inside a level is
    // irrelevant: the nodes of a level are independent)
    for (int q = tid; q <= m1 + 1; q += LRT_THREADS) levoff[q] = 0;
    grp_sync<WARP>();
    for (int i = tid; i < m1; i += LRT_THREADS) atomicAdd(&levoff[level[i] + 1], 1);
    grp_sync<WARP>();
    int nlev = 0;
    if (tid == 0) {
        for (int q = 0; q < m1; ++q) {
            if (levoff[q + 1] == 0) break;
            levoff[q + 1] += levoff[q];
            nlev = q + 1;
        }
        if (!WARP) s_nlev = nlev;
    }
    grp_sync<WARP>();
    nlev = WARP ? __shfl_sync(0xffffffffu, nlev, 0) : s_nlev;
    // scatter with per-level cursors: reuse G (doubles, unused yet) as ints
    int* cursor = reinterpret_cast<int*>(G);
    for (int q = tid; q < nlev; q += LRT_THREADS) cursor[q] = levoff[q];
    grp_sync<WARP>();
    for (int i = tid; i < m1; i += LRT_THREADS) order[atomicAdd(&cursor[level[i]], 1)] = i;
    grp_sync<WARP>();

    // ---- strictly feasible start: heights from the counts themselves, level by level
    {
        const double dmin = 1e-3;
        for (int q = 0; q < nlev; ++q) {
            for (int s = levoff[q] + tid; s < levoff[q + 1]; s += LRT_THREADS) {
                const int i = order[s];
                double hv = 0.0;
                for (int c = 0; c < 2; ++c) {
                    const int k = 2 * i + c;
                    const double hc = child[k] >= 0 ? h[child[k]] : 0.0;
                    hv = fmax(hv, hc + fmax(Y[k], dmin));
                }
                h[i] = hv;
            }
            grp_sync<WARP>();
        }
        double lmaxp = 0.0, z = 0.0;
        for (int k = tid; k < nb; k += LRT_THREADS) {
            const double hc = child[k] >= 0 ? h[child[k]] : 0.0;
            lmaxp = fmax(lmaxp, h[k >> 1] - hc);
        }
        lmaxp = -lmaxp;
        grp_reduce2<WARP, true>(z, lmaxp, sh);
        const double lmax = -lmaxp;
        const double sc = lmax >= 0.9 * U ? 0.9 * U / lmax : 1.0;
        if (sc != 1.0)
            for (int i = tid; i < m1; i += LRT_THREADS) h[i] *= sc;
        // from here on the objective only needs w/scale_fac and w*Y/scale_fac
        const double isf = 1.0 / sf;
        for (int k = tid; k < nb; k += LRT_THREADS) {
            const double wk = wraw[k] * isf;
            wq[k] = wk;
            Y[k] = wk * Y[k];
        }
        grp_sync<WARP>();
    }

    // n_stages == 1: the barrier problem at mu_in straight from the feasible start (fewest Newton steps).
    // n_stages > 1 is the SECOND ATTEMPT after a failed solve (the reference retries with SLSQP, run_PT_test.py:671-678):
    // path following -- the same Newton iteration at mu_in * 5^(n_stages - 1), ..., mu_in * 5, mu_in, each stage started
    // from the previous stage's optimum -- slower, but it walks to the same (unique) optimum along the central path.
    int it = 0, status = 1, it_total = 0;
    double f0 = 0.0;      // barrier objective at the current iterate (carried over from the line search)
    for (int stage = 0; stage < n_stages; ++stage) {
    double mu = mu_in;
    for (int q = stage; q < n_stages - 1; ++q) mu *= 5.0;
    status = 1;
    for (it = 0; it < max_iter; ++it) {
        // branch lengths, barrier gradient and curvature
        double f0p = 0.0;
        for (int k = tid; k < nb; k += LRT_THREADS) {
            const double hc = child[k] >= 0 ? h[child[k]] : 0.0;
            const double lk = h[k >> 1] - hc;
            const double a = lk - LRT_LMIN, b = U - lk;
            const double il = tree_rcp(lk), ia = tree_rcp(a), ib = tree_rcp(b);
            const double wy = Y[k];
            l[k] = lk;
            g[k] = wq[k] - wy * il + mu * (ib - ia);
            d[k] = wy * il * il + mu * (ia * ia + ib * ib);
            if (it == 0) f0p += wq[k] * lk - wy * log(lk) - mu * log(a * b);
        }
        if (it == 0) f0 = grp_sum<WARP>(f0p, sh);
        else grp_sync<WARP>();
        for (int i = tid; i < m1; i += LRT_THREADS) {
            const int u = up[i];
            G[i] = g[2 * i] + g[2 * i + 1] - (u >= 0 ? g[u] : 0.0);
            Hd[i] = d[2 * i] + d[2 * i + 1] + (u >= 0 ? d[u] : 0.0);
            e[i] = u >= 0 ? d[u] : 0.0;
        }
        grp_sync<WARP>();
        // leaf-to-root elimination, level by level (D holds the reciprocal pivots)
        for (int q = 0; q < nlev; ++q) {
            const int a0 = levoff[q], a1 = levoff[q + 1];
            for (int s = a0 + tid; s < a1; s += LRT_THREADS) {
                const int i = order[s];
                double Dv = Hd[i], rv = -G[i];
                for (int c = 0; c < 2; ++c) {
                    const int j = child[2 * i + c];
                    if (j >= 0) {
                        const double t = e[j] * D[j];
                        Dv -= e[j] * t;
                        rv += t * r[j];
                    }
                }
                D[i] = tree_rcp(Dv);
                r[i] = rv;
            }
            grp_sync<WARP>();
        }
        // root-to-leaf back substitution
        for (int q = nlev - 1; q >= 0; --q) {
            const int a0 = levoff[q], a1 = levoff[q + 1];
            for (int s = a0 + tid; s < a1; s += LRT_THREADS) {
                const int i = order[s];
                const int u = up[i];
                x[i] = (u >= 0 ? r[i] + e[i] * x[u >> 1] : r[i]) * D[i];
            }
            grp_sync<WARP>();
        }
        // Newton decrement and the largest step that stays strictly inside the bounds
        double dec = 0.0, amax = 1.0;
        for (int i = tid; i < m1; i += LRT_THREADS) dec -= G[i] * x[i];
        for (int k = tid; k < nb; k += LRT_THREADS) {
            const double xc = child[k] >= 0 ? x[child[k]] : 0.0;
            const double dk = x[k >> 1] - xc;
            dl[k] = dk;
            if (dk != 0.0) {
                const double room = dk < 0.0 ? l[k] - LRT_LMIN : U - l[k];
                amax = fmin(amax, 0.99 * room / fabs(dk));
            }
        }
        grp_reduce2<WARP, true>(dec, amax, sh);
        if (!(dec == dec) || dec < -1e-9 * fmax(1.0, fabs(f0))) {
            status = 3;   // NaN in the Newton system, or an ascent direction (the barrier objective is convex: cannot happen in exact arithmetic)
            break;
        }
        // Newton converges quadratically: once the decrement (~ twice the distance to the optimum
        // of the scaled objective) is below 1e-12 the step about to be taken leaves ~1e-24, far
        // under the rounding level of the objective, so take it and stop -- no line search, whose
        // Armijo test could only compare rounding noise there.
        if (dec <= 1e-12 * fmax(1.0, fabs(f0))) {
            for (int i = tid; i < m1; i += LRT_THREADS) h[i] += amax * x[i];
            grp_sync<WARP>();
            status = 0;
            ++it;
            break;
        }
        // backtracking (Armijo) on the barrier objective
        double a = amax, f1 = f0;
        bool ok = false;
        for (int bt = 0; bt < 40; ++bt) {
            double fp = 0.0;
            for (int k = tid; k < nb; k += LRT_THREADS) {
                const double lk = l[k] + a * dl[k];
                fp += wq[k] * lk - Y[k] * log(lk) - mu * log((lk - LRT_LMIN) * (U - lk));
            }
            f1 = grp_sum<WARP>(fp, sh);
            if (f1 <= f0 - 1e-4 * a * dec) {
                ok = true;
                break;
            }
            a *= 0.5;
        }
        if (ok)
            for (int i = tid; i < m1; i += LRT_THREADS) h[i] += a * x[i];
        grp_sync<WARP>();
        if (!ok) {          // no step decreases the objective any more: at the numerical optimum
            status = 0;
            ++it;
            break;
        }
        f0 = f1;
    }
    it_total += it;
    if (status != 0) break;
    }
    it = it_total;

    // ---- statistic and p-value (run_PT_test.py:683-695), from the unscaled counts and weights
    double ll_H0 = 0.0, sOB = 0.0;
    for (int k = tid; k < nb; k += LRT_THREADS) {
        const double hc = child[k] >= 0 ? h[child[k]] : 0.0;
        const double lk = h[k >> 1] - hc;
        const double y = gather ? Ysrc[gather[pb.br_off + k]] : Ysrc[pb.br_off + k];
        ll_H0 += (y * log(lk) - lk) * wall[pb.br_off + k];
        sOB += (lk <= 1e-3) ? 1.0 : 0.0;
        if (x_out) x_out[pb.br_off + k] = lk;
    }
    grp_reduce2<WARP, false>(ll_H0, sOB, sh);
    const int on_bound = int(sOB + 0.5);
    const double LR = -2.0 * (ll_H0 - ll_H1);
    const double dof = rint(wsum - double(m1));
    // mixture over the number of parameters on the bound, binomial weights, in log space
    double num = 0.0, den = 0.0;
    const double lg_n = lgamma(double(on_bound) + 1.0);
    for (int k = tid; k <= on_bound; k += LRT_THREADS) {
        const double df = dof - double(k);
        if (df > 0.0) {
            double pv = gamma_q(0.5 * df, 0.5 * LR);
            if (on_bound > 0) pv = fmin(fmax(pv, 1e-100), 1.0);   // the reference clips only inside the mixture (:690 vs :695)
            // weights relative to the central binomial coefficient keep everything finite
            const double lw = lg_n - lgamma(double(k) + 1.0) - lgamma(double(on_bound - k) + 1.0) - double(on_bound) * 0.6931471805599453;
            const double wk = exp(lw);
            num += wk * pv;
            den += wk;
        }
    }
    grp_reduce2<WARP, false>(num, den, sh);
    if (tid == 0) {
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

#include "scs_lrt_tree.cuh"

namespace scs {

// ------------------------------------------------------------ branch weights
// add_br_weights (run_PT_test.py:451-502): for every node, the probability that
// all of its cells dropped out / went missing while the others read wild type:
//   p_ADO = max(exp(sum_{c in clade} l_DO[c] + sum_{c not in clade} l_TN[c]), 1e-6)
// One warp per node row of the clade bit mask.
__global__ void scs_branch_padO_kernel(const uint32_t* __restrict__ bits, int64_t pitch_words, int n_nodes, int n_cells,
                                       const double* __restrict__ l_do_minus_tn, double sum_tn, double* __restrict__ p_ado,
                                       int* __restrict__ clade_size) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_nodes) return;
    const int words = (n_cells + 31) / 32;
    double acc = 0.0;
    int cnt = 0;
    for (int wd = 0; wd < words; ++wd) {
        const uint32_t b = bits[size_t(row) * pitch_words + wd];
        const int c = wd * 32 + lane;
        if (((b >> lane) & 1u) && c < n_cells) {
            acc += l_do_minus_tn[c];
            cnt++;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) {
        const double xlog = sum_tn + acc;
        // the reference catches the FloatingPointError of an underflowing exp (:485)
        const double pv = xlog < -708.3964185322641 ? 1e-6 : fmax(exp(xlog), 1e-6);
        p_ado[row] = pv;
        clade_size[row] = cnt;
    }
}

}  // namespace scs

using namespace scs;

extern "C" {

struct LrtPlanImpl {
    scs_ctx* ctx = nullptr;
    int n_prob = 0;
    size_t nbr = 0, nnode = 0;
    double* d_buf = nullptr;
    int* i_buf = nullptr;
    LrtProblem* d_probs = nullptr;
    double *dYsrc = nullptr, *dY = nullptr, *dw = nullptr, *d_out = nullptr, *d_x = nullptr;
    int *d_child = nullptr, *d_gather = nullptr;
    bool have_gather = false, have_w = false;
    size_t src_len = 0;
    LrtWork ws;
    long long launches = 0;
    int threads = 128;
    int smem_bytes = 0;    // > 0: per-CTA working set in shared memory (small trees)
    bool warp_mode = false;   // one warp per problem (trees up to 256 leaves)
    int n_stages = 1;          // > 1: path following over that many barrier parameters (the second attempt)
    int max_iter = LRT_MAX_ITER;
    // cluster-per-problem kernel (scs_lrt_tree.cuh): topology renumbered in level order on the host
    bool tree_mode = false;
    int tree_cluster = 1, tree_threads = 512, tree_smem = 0;
    std::vector<int> h_orig;          // [branch, level-order numbering] -> the caller's branch index (global, with br_off)
    LrtTreeProblem* t_probs = nullptr;
    int *t_child = nullptr, *t_up = nullptr, *t_orig = nullptr, *t_levoff = nullptr, *t_gather = nullptr;
    double *t_w = nullptr, *t_hB = nullptr, *t_slots = nullptr, *t_cD = nullptr, *t_cR = nullptr;
    int* t_ibuf = nullptr;
    long long* t_prof = nullptr;
    double* t_dbuf = nullptr;
};

// Level-order renumbering of every problem of the batch for scs_lrt_tree_kernel.  top_cap: nodes the leader CTA's shared
// memory holds (32 bytes each).
static int lrt_tree_prepare(LrtPlanImpl* pl, int n_prob, const int64_t* br_off, const int32_t* child_int, const std::vector<LrtProblem>& probs) {
    int top_cap = (227 * 1024 - 1024 - LRT_TREE_LEVSM * 4) / 32;
    if (const char* ev = getenv("SCS_LRT_TOP_CAP")) top_cap = std::max(0, std::min(top_cap, atoi(ev)));     // test knob: forces levels below the cut
    const size_t nbr = pl->nbr, nnode = pl->nnode;
    std::vector<LrtTreeProblem> tp((size_t(n_prob)));
    std::vector<int> t_child(nbr), t_orig(nbr), t_up(nnode, -1), t_levoff;
    int max_top = 0, max_lev_smem = 0;
    for (int p = 0; p < n_prob; ++p) {
        const int nb = probs[p].n_br, m1 = nb / 2;
        const int32_t* ch = child_int + br_off[p];
        std::vector<int> level(size_t(m1), 0), pos(size_t(m1), 0);
        int nlev = 0;
        for (int i = 0; i < m1; ++i) {          // children have smaller indices (validated by the caller)
            int lv = 0;
            for (int c = 0; c < 2; ++c)
                if (ch[2 * i + c] >= 0) lv = std::max(lv, level[size_t(ch[2 * i + c])] + 1);
            level[size_t(i)] = lv;
            nlev = std::max(nlev, lv + 1);
        }
        std::vector<int> off(size_t(nlev) + 1, 0);
        for (int i = 0; i < m1; ++i) off[size_t(level[size_t(i)]) + 1]++;
        for (int q = 0; q < nlev; ++q) off[size_t(q) + 1] += off[size_t(q)];
        std::vector<int> cur(off.begin(), off.end() - 1);
        for (int i = 0; i < m1; ++i) pos[size_t(i)] = cur[size_t(level[size_t(i)])]++;
        int* tc = t_child.data() + probs[p].br_off;
        int* to = t_orig.data() + probs[p].br_off;
        int* tu = t_up.data() + probs[p].node_off;
        for (int i = 0; i < m1; ++i)
            for (int c = 0; c < 2; ++c) {
                const int k = 2 * pos[size_t(i)] + c, j = ch[2 * i + c];
                tc[k] = j >= 0 ? pos[size_t(j)] : -1;
                to[k] = 2 * i + c;
                if (j >= 0) tu[pos[size_t(j)]] = k;
            }
        int qc = 1;
        while (qc < nlev && m1 - off[size_t(qc)] > top_cap) ++qc;
        int qw = nlev;
        while (qw > qc && off[size_t(qw)] - off[size_t(qw) - 1] <= 32) --qw;
        LrtTreeProblem& t = tp[size_t(p)];
        t.br_off = probs[p].br_off;
        t.n_br = nb;
        t.node_off = probs[p].node_off;
        t.nlev = nlev;
        t.lev_off = int(t_levoff.size());
        t.qc = qc;
        t.qw = qw;
        t.s_cut = off[size_t(std::min(qc, nlev))];
        t_levoff.insert(t_levoff.end(), off.begin(), off.end());
        max_top = std::max(max_top, m1 - t.s_cut);
        if (nlev + 1 <= LRT_TREE_LEVSM) max_lev_smem = std::max(max_lev_smem, nlev + 1);
    }
    pl->h_orig.resize(nbr);
    for (int p = 0; p < n_prob; ++p)
        for (int k = 0; k < probs[p].n_br; ++k) pl->h_orig[size_t(probs[p].br_off) + k] = probs[p].br_off + t_orig[size_t(probs[p].br_off) + k];
    pl->tree_smem = max_top * 32 + max_lev_smem * 4;
    const int cl_max = std::max(1, pl->ctx->sm_count / std::max(1, n_prob));
    pl->tree_cluster = cl_max >= 8 ? 8 : (cl_max >= 4 ? 4 : (cl_max >= 2 ? 2 : 1));
    if (const char* ev = getenv("SCS_LRT_CLUSTER")) {
        const int c = atoi(ev);
        if (c == 1 || c == 2 || c == 4 || c == 8) pl->tree_cluster = c;
    }
    const int64_t per_cta = (pl->nbr / size_t(n_prob)) / size_t(pl->tree_cluster);
    pl->tree_threads = per_cta >= 1024 ? 512 : 256;
    const size_t ints = 3 * nbr + nnode + t_levoff.size();
    const size_t dbl = 3 * nbr + nnode + size_t(n_prob) * LRT_TREE_SLOT_DOUBLES;
    if (cudaMalloc(&pl->t_ibuf, ints * sizeof(int)) != cudaSuccess || cudaMalloc(&pl->t_dbuf, dbl * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&pl->t_probs, size_t(n_prob) * sizeof(LrtTreeProblem)) != cudaSuccess)
        return set_error(SCS_ERR_OOM, "device allocation failed for the LRT batch");
    int* iq = pl->t_ibuf;
    pl->t_child = iq; iq += nbr;
    pl->t_orig = iq; iq += nbr;
    pl->t_gather = iq; iq += nbr;
    pl->t_up = iq; iq += nnode;
    pl->t_levoff = iq;
    double* dq = pl->t_dbuf;
    pl->t_cD = dq; dq += nbr;      // (first: 16-byte aligned pairs)
    pl->t_cR = dq; dq += nbr;
    pl->t_w = dq; dq += nbr;
    pl->t_hB = dq; dq += nnode;
    pl->t_slots = dq;
    cudaError_t e = cudaMemcpy(pl->t_child, t_child.data(), nbr * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(pl->t_orig, t_orig.data(), nbr * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(pl->t_gather, pl->h_orig.data(), nbr * sizeof(int), cudaMemcpyHostToDevice);     // no gather: the caller's order
    if (e == cudaSuccess) e = cudaMemcpy(pl->t_up, t_up.data(), nnode * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(pl->t_levoff, t_levoff.data(), t_levoff.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(pl->t_probs, tp.data(), size_t(n_prob) * sizeof(LrtTreeProblem), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(pl->t_slots, 0, size_t(n_prob) * LRT_TREE_SLOT_DOUBLES * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(pl->t_cD, 0, 2 * nbr * sizeof(double));
    if (e != cudaSuccess) return set_error(SCS_ERR_CUDA, "LRT plan upload failed: %s", cudaGetErrorString(e));
    return SCS_OK;
}

static int lrt_plan_create_impl(scs_ctx* ctx, int32_t n_prob, const int64_t* br_off, const int32_t* child_int, scs_lrt_plan** out, bool allow_tree) {
    if (!ctx || !br_off || !child_int || !out) return set_error(SCS_ERR_ARG, "null argument");
    if (n_prob <= 0) return set_error(SCS_ERR_ARG, "empty LRT batch");
    SCS_CUDA(cudaSetDevice(ctx->device));
    const int64_t total_br = br_off[n_prob];
    if (total_br > (int64_t(1) << 30)) return set_error(SCS_ERR_ARG, "batch too large: split it");
    std::vector<LrtProblem> probs(n_prob);
    int64_t node_off = 0, max_nb = 0;
    for (int p = 0; p < n_prob; ++p) {
        const int64_t nb = br_off[p + 1] - br_off[p];
        if (nb < 2 || (nb & 1)) return set_error(SCS_ERR_ARG, "problem %d has %lld branches; a binary tree has an even number >= 2", p, (long long)nb);
        for (int64_t k = 0; k < nb; ++k) {
            const int c = child_int[br_off[p] + k];
            if (c >= int(k / 2) || c < -1) return set_error(SCS_ERR_ARG, "problem %d branch %lld: child node %d is not below its parent %lld in post-order", p, (long long)k, c, (long long)(k / 2));
        }
        max_nb = std::max<int64_t>(max_nb, nb);
        probs[p].br_off = int(br_off[p]);
        probs[p].n_br = int(nb);
        probs[p].node_off = int(node_off);
        node_off += nb / 2 + 2;   // levoff needs (levels + 1) <= nodes + 1 entries
    }
    LrtPlanImpl* pl = new LrtPlanImpl();
    pl->ctx = ctx;
    pl->n_prob = n_prob;
    pl->nbr = size_t(total_br);
    pl->nnode = size_t(node_off);
    pl->threads = max_nb >= 8192 ? 512 : (max_nb >= 2048 ? 256 : (max_nb > 256 ? 128 : 64));
    // small trees: one warp per problem -- no CTA barriers in the level sweeps, shuffle-only reductions
    pl->warp_mode = max_nb <= 512;
    if (const char* ev = getenv("SCS_LRT_WARP")) pl->warp_mode = atoi(ev) != 0 && max_nb <= 1024;
    if (pl->warp_mode) pl->threads = 32;
    // the per-problem working set lives in shared memory while it fits one SM (~3,000 branches);
    // beyond 1024 branches only a handful of CTAs are resident anyway
    if (max_nb <= 3000) {
        const size_t mn = size_t(max_nb / 2 + 2);
        pl->smem_bytes = int((4 * size_t(max_nb) + 7 * mn) * sizeof(double) + (size_t(max_nb) + 3 * mn) * sizeof(int));
    }
    const size_t nbr = pl->nbr, nnode = pl->nnode;
    const size_t dbl = 9 * nbr /*Ysrc Y w l g d dl x wq*/ + 7 * nnode + size_t(n_prob) * 8;
    const size_t ints = 2 * nbr /*child gather*/ + 4 * nnode;
    if (cudaMalloc(&pl->d_buf, dbl * sizeof(double)) != cudaSuccess || cudaMalloc(&pl->i_buf, ints * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&pl->d_probs, size_t(n_prob) * sizeof(LrtProblem)) != cudaSuccess) {
        scs_lrt_plan_destroy(reinterpret_cast<scs_lrt_plan*>(pl));
        return set_error(SCS_ERR_OOM, "device allocation failed for the LRT batch");
    }
    double* q = pl->d_buf;
    pl->dYsrc = q; q += nbr;
    pl->dY = q; q += nbr;
    pl->dw = q; q += nbr;
    pl->ws.l = q; q += nbr;
    pl->ws.g = q; q += nbr;
    pl->ws.d = q; q += nbr;
    pl->ws.dl = q; q += nbr;
    pl->d_x = q; q += nbr;
    pl->ws.wq = q; q += nbr;
    pl->ws.h = q; q += nnode;
    pl->ws.Hd = q; q += nnode;
    pl->ws.e = q; q += nnode;
    pl->ws.D = q; q += nnode;
    pl->ws.r = q; q += nnode;
    pl->ws.x = q; q += nnode;
    pl->ws.G = q; q += nnode;
    pl->d_out = q;
    int* iq = pl->i_buf;
    pl->d_child = iq; iq += nbr;
    pl->d_gather = iq; iq += nbr;
    pl->ws.up = iq; iq += nnode;
    pl->ws.order = iq; iq += nnode;
    pl->ws.level = iq; iq += nnode;
    pl->ws.levoff = iq;
    cudaError_t e = cudaMemcpy(pl->d_child, child_int, nbr * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(pl->d_probs, probs.data(), size_t(n_prob) * sizeof(LrtProblem), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        scs_lrt_plan_destroy(reinterpret_cast<scs_lrt_plan*>(pl));
        return set_error(SCS_ERR_CUDA, "LRT plan upload failed: %s", cudaGetErrorString(e));
    }
    // trees too big for one warp: a thread-block cluster per problem on a level-ordered topology (scs_lrt_tree.cuh);
    // SCS_LRT_KERNEL = tree | cta | warp forces a kernel (tests play them against each other)
    pl->tree_mode = !pl->warp_mode;
    if (const char* ev = getenv("SCS_LRT_MAX_ITER")) pl->max_iter = std::max(1, atoi(ev));      // test knob: makes the first attempt fail
    if (const char* ev = getenv("SCS_LRT_KERNEL")) {
        const std::string k(ev);
        if (k == "tree") pl->tree_mode = true, pl->warp_mode = false;
        else if (k == "cta") pl->tree_mode = false, pl->warp_mode = false;
        else if (k == "warp" && max_nb <= 1024) pl->tree_mode = false, pl->warp_mode = true;
        if (!pl->warp_mode && pl->threads == 32) pl->threads = max_nb >= 8192 ? 512 : (max_nb >= 2048 ? 256 : (max_nb > 256 ? 128 : 64));
        if (pl->warp_mode) pl->threads = 32;
    }
    if (!allow_tree && pl->tree_mode) {
        pl->tree_mode = false;
        if (pl->threads == 32 && !pl->warp_mode) pl->threads = 128;
    }
    if (pl->tree_mode) {
        const int rc = lrt_tree_prepare(pl, n_prob, br_off, child_int, probs);
        if (rc != SCS_OK) {
            scs_lrt_plan_destroy(reinterpret_cast<scs_lrt_plan*>(pl));
            return rc;
        }
    }
    *out = reinterpret_cast<scs_lrt_plan*>(pl);
    return SCS_OK;
}

int scs_lrt_plan_create(scs_ctx* ctx, int32_t n_prob, const int64_t* br_off, const int32_t* child_int, scs_lrt_plan** out) {
    return lrt_plan_create_impl(ctx, n_prob, br_off, child_int, out, true);
}

// A plan of n_prob problems with n_br branches each whose topologies, gather indices and weights are produced ON THE DEVICE
// (scs_sweep_model, scs_sweep.cu): created around a placeholder topology (a caterpillar tree) that is overwritten before the
// first run.  The kernel's shape decisions (threads per problem, shared-memory working set) depend on n_br only.
int scs_lrt_plan_create_uniform(scs_ctx* ctx, int32_t n_prob, int32_t n_br, scs_lrt_plan** out) {
    if (!ctx || !out) return set_error(SCS_ERR_ARG, "null argument");
    if (n_prob <= 0 || n_br < 2 || (n_br & 1)) return set_error(SCS_ERR_ARG, "bad LRT batch shape (problems=%d branches=%d)", n_prob, n_br);
    std::vector<int64_t> off(size_t(n_prob) + 1);
    for (int p = 0; p <= n_prob; ++p) off[size_t(p)] = int64_t(p) * n_br;
    std::vector<int32_t> child(size_t(n_prob) * n_br, -1);
    for (int p = 0; p < n_prob; ++p)
        for (int i = 1; i < n_br / 2; ++i) child[size_t(p) * n_br + 2 * i + 1] = i - 1;
    return scs_lrt_plan_create(ctx, n_prob, off.data(), child.data(), out);
}

// Device buffers a producer kernel fills instead of scs_lrt_plan_set_weights / set_gather / the topology upload:
// child_int [n_br total], gather index [n_br total] into a source vector of src_len doubles, weights [n_br total].
int scs_lrt_plan_device_inputs(scs_lrt_plan* h, int64_t src_len, int32_t** d_child, int32_t** d_gather, double** d_w) {
    LrtPlanImpl* pl = reinterpret_cast<LrtPlanImpl*>(h);
    if (!pl || !d_child || !d_gather || !d_w) return set_error(SCS_ERR_ARG, "null argument");
    *d_child = pl->d_child;
    *d_gather = pl->d_gather;
    *d_w = pl->dw;
    pl->have_gather = true;
    pl->have_w = true;
    pl->src_len = size_t(src_len);
    if (pl->tree_mode) {            // the level order was made for the placeholder topology: back to one CTA per problem
        pl->tree_mode = false;
        if (pl->threads == 32 && !pl->warp_mode) pl->threads = 128;
    }
    return SCS_OK;
}

int scs_lrt_plan_destroy(scs_lrt_plan* h) {
    LrtPlanImpl* pl = reinterpret_cast<LrtPlanImpl*>(h);
    if (!pl) return SCS_OK;
    cudaFree(pl->d_buf);
    cudaFree(pl->i_buf);
    cudaFree(pl->d_probs);
    cudaFree(pl->t_ibuf);
    cudaFree(pl->t_dbuf);
    cudaFree(pl->t_probs);
    cudaFree(pl->t_prof);
    delete pl;
    return SCS_OK;
}

int scs_lrt_plan_set_weights(scs_lrt_plan* h, const double* w) {
    LrtPlanImpl* pl = reinterpret_cast<LrtPlanImpl*>(h);
    if (!pl || !w) return set_error(SCS_ERR_ARG, "null argument");
    SCS_CUDA(cudaMemcpy(pl->dw, w, pl->nbr * sizeof(double), cudaMemcpyHostToDevice));
    if (pl->tree_mode) {
        std::vector<double> wp(pl->nbr);
        for (size_t k = 0; k < pl->nbr; ++k) wp[k] = w[size_t(pl->h_orig[k])];
        SCS_CUDA(cudaMemcpy(pl->t_w, wp.data(), pl->nbr * sizeof(double), cudaMemcpyHostToDevice));
    }
    pl->have_w = true;
    return SCS_OK;
}

int scs_lrt_plan_set_gather(scs_lrt_plan* h, const int32_t* index, int64_t src_len) {
    LrtPlanImpl* pl = reinterpret_cast<LrtPlanImpl*>(h);
    if (!pl) return set_error(SCS_ERR_ARG, "null argument");
    if (!index) {
        pl->have_gather = false;
        if (pl->tree_mode) SCS_CUDA(cudaMemcpy(pl->t_gather, pl->h_orig.data(), pl->nbr * sizeof(int), cudaMemcpyHostToDevice));
        return SCS_OK;
    }
    for (size_t k = 0; k < pl->nbr; ++k)
        if (index[k] < 0 || index[k] >= src_len) return set_error(SCS_ERR_ARG, "gather index %d out of range [0,%lld)", index[k], (long long)src_len);
    SCS_CUDA(cudaMemcpy(pl->d_gather, index, pl->nbr * sizeof(int), cudaMemcpyHostToDevice));
    if (pl->tree_mode) {
        std::vector<int> gp(pl->nbr);
        for (size_t k = 0; k < pl->nbr; ++k) gp[k] = index[size_t(pl->h_orig[k])];
        SCS_CUDA(cudaMemcpy(pl->t_gather, gp.data(), pl->nbr * sizeof(int), cudaMemcpyHostToDevice));
    }
    pl->have_gather = true;
    pl->src_len = size_t(src_len);
    return SCS_OK;
}

int scs_lrt_plan_run(scs_lrt_plan* h, const double* d_Y, double* d_out, double* d_x, void* stream) {
    LrtPlanImpl* pl = reinterpret_cast<LrtPlanImpl*>(h);
    if (!pl || !d_Y) return set_error(SCS_ERR_ARG, "null argument");
    if (!pl->have_w) return set_error(SCS_ERR_ARG, "scs_lrt_plan_set_weights has not been called");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    double mu = LRT_MU;
    if (const char* e = getenv("SCS_LRT_MU")) mu = atof(e) > 0.0 ? atof(e) : LRT_MU;      // analysis knob (tools/lrt_sweep_probe.py)
    if (pl->tree_mode) {
        LrtTreeArgs a;
        a.probs = pl->t_probs;
        a.Ysrc = d_Y;
        a.gather = pl->t_gather;
        a.w = pl->t_w;
        a.child = pl->t_child;
        a.up = pl->t_up;
        a.orig = pl->t_orig;
        a.levoff = pl->t_levoff;
        a.Y = pl->dY; a.wq = pl->ws.wq; a.l = pl->ws.l; a.dl = pl->ws.dl;
        a.hA = pl->ws.h; a.hB = pl->t_hB; a.x = pl->ws.x; a.G = pl->ws.G; a.e = pl->ws.e; a.D = pl->ws.D; a.r = pl->ws.r;
        a.slots = pl->t_slots;
        a.cD = pl->t_cD; a.cR = pl->t_cR;
        a.out = d_out ? d_out : pl->d_out;
        a.x_out = d_x ? d_x : pl->d_x;
        a.mu = mu;
        a.max_iter = pl->max_iter;
        a.prof = nullptr;
        a.start_mode = 0;
        if (const char* ev = getenv("SCS_LRT_START")) a.start_mode = atoi(ev) == 1 ? 1 : 0;      // experiment knob (tools/lrt_timing.py)
        if (getenv("SCS_LRT_PROF")) {          // tools/lrt_timing.py: cycles per phase, printed after the launch (synchronises)
            if (!pl->t_prof && cudaMalloc(&pl->t_prof, 16 * sizeof(long long)) != cudaSuccess) pl->t_prof = nullptr;
            if (pl->t_prof) cudaMemsetAsync(pl->t_prof, 0, 16 * sizeof(long long), st);
            a.prof = pl->t_prof;
        }
        auto launch_tree = [&](auto kern) -> cudaError_t {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pl->tree_smem);
            if (e != cudaSuccess) return e;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(unsigned(pl->n_prob * pl->tree_cluster));
            cfg.blockDim = dim3(unsigned(pl->tree_threads));
            cfg.dynamicSmemBytes = size_t(pl->tree_smem);
            cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = unsigned(pl->tree_cluster);
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            return cudaLaunchKernelEx(&cfg, kern, a);
        };
        SCS_CUDA(pl->tree_threads == 512 ? launch_tree(scs_lrt_tree_kernel<512>) : launch_tree(scs_lrt_tree_kernel<256>));
        if (a.prof) {
            long long h[16];
            if (cudaMemcpyAsync(h, a.prof, sizeof(h), cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess) {
                static const char* nm[12] = {"rows", "sync", "bottom-up", "fold", "top-up-cta", "thin", "top-down-cta", "x-out+sync", "bottom-down", "step", "reduce", "linesearch"};
                fprintf(stderr, "lrt phases (cycles, problem 0, %lld iterations, %lld line-search tries):", h[15] + 1, h[14]);
                for (int i = 0; i < 12; ++i) fprintf(stderr, " %s %lld", nm[i], h[i]);
                fprintf(stderr, "\n");
            }
        }
        pl->launches++;
        return SCS_OK;
    }
    auto launch = [&](auto kern) {
        if (pl->smem_bytes > 0) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pl->smem_bytes);
            cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        }
        kern<<<pl->n_prob, pl->threads, pl->smem_bytes, st>>>(pl->d_probs, d_Y, pl->have_gather ? pl->d_gather : nullptr, pl->dY, pl->dw,
                                                              pl->d_child, pl->ws, d_out ? d_out : pl->d_out, d_x ? d_x : pl->d_x,
                                                              pl->smem_bytes > 0, mu, pl->n_stages, pl->max_iter);
    };
    if (pl->warp_mode) launch(scs_lrt_kernel<true>);
    else launch(scs_lrt_kernel<false>);
    pl->launches++;
    SCS_CUDA(cudaGetLastError());
    return SCS_OK;
}

int scs_lrt_plan_run_host(scs_lrt_plan* h, const double* Y, double* out, double* x_out, void* stream) {
    LrtPlanImpl* pl = reinterpret_cast<LrtPlanImpl*>(h);
    if (!pl || !Y || !out) return set_error(SCS_ERR_ARG, "null argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t n = pl->have_gather ? pl->src_len : pl->nbr;
    if (n > pl->nbr) return set_error(SCS_ERR_ARG, "gather source longer than the batch");
    SCS_CUDA(cudaMemcpyAsync(pl->dYsrc, Y, n * sizeof(double), cudaMemcpyHostToDevice, st));
    int rc = scs_lrt_plan_run(h, pl->dYsrc, nullptr, nullptr, stream);
    if (rc != SCS_OK) return rc;
    SCS_CUDA(cudaMemcpyAsync(out, pl->d_out, size_t(pl->n_prob) * 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (x_out) SCS_CUDA(cudaMemcpyAsync(x_out, pl->d_x, pl->nbr * sizeof(double), cudaMemcpyDeviceToHost, st));
    SCS_CUDA(cudaStreamSynchronize(st));
    return SCS_OK;
}

const void* scs_lrt_plan_device_result(scs_lrt_plan* h) {
    LrtPlanImpl* pl = reinterpret_cast<LrtPlanImpl*>(h);
    return pl ? pl->d_out : nullptr;
}

int scs_lrt_poisson_batch(scs_ctx* ctx, int32_t n_prob, const int64_t* br_off, const double* Y, const double* w,
                          const int32_t* child_int, double* out, double* x_out) {
    if (!ctx || !br_off || !Y || !w || !child_int || !out) return set_error(SCS_ERR_ARG, "null argument");
    if (n_prob <= 0) return SCS_OK;
    scs_lrt_plan* pl = nullptr;
    int rc = scs_lrt_plan_create(ctx, n_prob, br_off, child_int, &pl);
    if (rc != SCS_OK) return rc;
    rc = scs_lrt_plan_set_weights(pl, w);
    if (rc == SCS_OK) rc = scs_lrt_plan_run_host(pl, Y, out, x_out, nullptr);
    scs_lrt_plan_destroy(pl);
    if (rc != SCS_OK) return rc;
    // ---- second attempt for the problems whose solve failed (status 1: iteration limit, 3: numerical breakdown), as the
    // reference retries a failed trust-constr solve with SLSQP (run_PT_test.py:671-678): path following over eight barrier
    // parameters down to the same final one (scs_lrt_kernel, n_stages = 8) -- the same unique optimum, reached along the
    // central path.  What fails twice keeps its failure tuple (:676-678).  Degenerate inputs (status 2, 4) are not retried.
    std::vector<int> again;
    for (int p = 0; p < n_prob; ++p)
        if (out[size_t(p) * 8 + 7] == 1.0 || out[size_t(p) * 8 + 7] == 3.0) again.push_back(p);
    if (again.empty() || getenv("SCS_LRT_NO_RETRY")) return SCS_OK;
    std::vector<int64_t> off2(again.size() + 1, 0);
    for (size_t k = 0; k < again.size(); ++k) off2[k + 1] = off2[k] + (br_off[again[k] + 1] - br_off[again[k]]);
    std::vector<double> Y2(size_t(off2.back())), w2(size_t(off2.back())), out2(again.size() * 8), x2(size_t(off2.back()));
    std::vector<int32_t> ch2(size_t(off2.back()));
    for (size_t k = 0; k < again.size(); ++k) {
        const int64_t a = br_off[again[k]], n = br_off[again[k] + 1] - a;
        memcpy(Y2.data() + off2[k], Y + a, size_t(n) * sizeof(double));
        memcpy(w2.data() + off2[k], w + a, size_t(n) * sizeof(double));
        memcpy(ch2.data() + off2[k], child_int + a, size_t(n) * sizeof(int32_t));
    }
    scs_lrt_plan* pl2 = nullptr;
    rc = lrt_plan_create_impl(ctx, int(again.size()), off2.data(), ch2.data(), &pl2, false);
    if (rc != SCS_OK) return rc;
    {
        LrtPlanImpl* q = reinterpret_cast<LrtPlanImpl*>(pl2);
        q->n_stages = 8;
        q->max_iter = LRT_MAX_ITER;
    }
    rc = scs_lrt_plan_set_weights(pl2, w2.data());
    if (rc == SCS_OK) rc = scs_lrt_plan_run_host(pl2, Y2.data(), out2.data(), x2.data(), nullptr);
    scs_lrt_plan_destroy(pl2);
    if (rc != SCS_OK) return rc;
    for (size_t k = 0; k < again.size(); ++k) {
        if (out2[k * 8 + 7] != 0.0) continue;               // failed twice: the failure tuple stays
        const double first_its = out[size_t(again[k]) * 8 + 6];
        memcpy(out + size_t(again[k]) * 8, out2.data() + k * 8, 8 * sizeof(double));
        out[size_t(again[k]) * 8 + 6] += std::isfinite(first_its) ? first_its : 0.0;      // Newton iterations of both attempts
        if (x_out) memcpy(x_out + br_off[again[k]], x2.data() + off2[k], size_t(off2[k + 1] - off2[k]) * sizeof(double));
    }
    return SCS_OK;
}

int scs_branch_weights(scs_ctx* ctx, const uint32_t* clade_bits, int64_t pitch_words, int32_t n_nodes, int32_t n_cells,
                       const double* missing_rate, double FP, double FN, const double* w_max, int32_t n_w,
                       double* weights, double* weights_norm, int32_t* root_row) {
    if (!ctx || !clade_bits || !missing_rate || !w_max || !weights || !weights_norm) return set_error(SCS_ERR_ARG, "null argument");
    if (n_nodes < 2 || n_cells < 1 || n_w < 1) return set_error(SCS_ERR_ARG, "empty problem");
    SCS_CUDA(cudaSetDevice(ctx->device));
    std::vector<double> diff(n_cells);
    double sum_tn = 0.0;
    for (int c = 0; c < n_cells; ++c) {
        const double ms = missing_rate[c];
        if (ms < 0.0) {   // cell not part of the tree
            diff[c] = 0.0;
            continue;
        }
        const double l_tn = log((1.0 - ms) * (1.0 - FP) + ms);   // :475
        const double l_do = log((1.0 - ms) * FN + ms);           // :477
        diff[c] = l_do - l_tn;
        sum_tn += l_tn;
    }
    uint32_t* d_bits = nullptr;
    double *d_diff = nullptr, *d_p = nullptr;
    int* d_cnt = nullptr;
    const size_t mask_bytes = size_t(n_nodes) * pitch_words * 4;
    int rc = SCS_OK;
    std::vector<double> p(n_nodes);
    std::vector<int> cnt(n_nodes);
    do {
        if (cudaMalloc(&d_bits, mask_bytes) != cudaSuccess || cudaMalloc(&d_diff, n_cells * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&d_p, n_nodes * sizeof(double)) != cudaSuccess || cudaMalloc(&d_cnt, n_nodes * sizeof(int)) != cudaSuccess) {
            rc = set_error(SCS_ERR_OOM, "device allocation failed for branch weights");
            break;
        }
        cudaError_t e = cudaMemcpy(d_bits, clade_bits, mask_bytes, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_diff, diff.data(), n_cells * sizeof(double), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
            const int threads = 128, blocks = (n_nodes * 32 + threads - 1) / threads;
            scs_branch_padO_kernel<<<blocks, threads>>>(d_bits, pitch_words, n_nodes, n_cells, d_diff, sum_tn, d_p, d_cnt);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpy(p.data(), d_p, n_nodes * sizeof(double), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(cnt.data(), d_cnt, n_nodes * sizeof(int), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = set_error(SCS_ERR_CUDA, "branch weight kernel failed: %s", cudaGetErrorString(e));
    } while (0);
    cudaFree(d_bits);
    cudaFree(d_diff);
    cudaFree(d_p);
    cudaFree(d_cnt);
    if (rc != SCS_OK) return rc;
    // root = first node that holds every cell of the tree (:497); it is dropped from the normalisation
    int n_tree_cells = 0;
    for (int c = 0; c < n_cells; ++c) n_tree_cells += missing_rate[c] >= 0.0;
    int root = -1;
    for (int i = 0; i < n_nodes; ++i)
        if (cnt[i] == n_tree_cells) {
            root = i;
            break;
        }
    if (root < 0) return set_error(SCS_ERR_ARG, "no node of the clade mask contains every cell: not a rooted tree");
    if (root_row) *root_row = root;
    for (int wi = 0; wi < n_w; ++wi) {
        double sum = 0.0;
        for (int i = 0; i < n_nodes; ++i) {
            const double pa = p[i];
            double wv = 1.0 / ((1.0 - pa) * pa);                  // :492 inverse variance
            if (wv > w_max[wi]) wv = w_max[wi];
            weights[size_t(wi) * n_nodes + i] = wv;
            if (i != root) sum += wv;
        }
        for (int i = 0; i < n_nodes; ++i)
            weights_norm[size_t(wi) * n_nodes + i] = i == root ? -1.0 : double(n_nodes - 1) * weights[size_t(wi) * n_nodes + i] / sum;   // :502
        weights[size_t(wi) * n_nodes + root] = -1.0;                                                                                   // :513
    }
    return SCS_OK;
}

}  // extern "C"
