// pnp_solver.cuh -- per-cell linear algebra of the implicit integrator (K2+K3 building blocks).
//
// One warp owns one cell.  The Newton matrix  A = Mass/gamma - dF/dy  is block
// tridiagonal (+ one wall block); it is FACTORED by a block-Thomas sweep with
// partial pivoting inside the NB x NB blocks and the factors are kept in global
// memory (L2) so that the modified-Newton iterations of the following steps only
// run the cheap SOLVE sweeps (VODE/CVODE re-use policy):
//
//   factor_sweep   lane j owns one column of the augmented block row
//                  [A_D' | I | u_g]  (2*NB+1 <= 32 columns, NB registers per lane);
//                  Gauss-Jordan with row pivoting turns it into [I | inv_i | W_i[:,g]];
//                  inv_i = A_D'^{-1}, W_i = inv_i*A_U (A_U = diag + one column (g)),
//                  both go to global memory; W columns are handed to the A_D lanes of
//                  the next node by warp shuffles for the Schur update A_D' = A_D - A_L*W.
//   residual_pass  lane per node: rhs = F(y) - Mass*(y+psi)/gamma  (K1 arithmetic on the
//                  shared-memory state)
//   forward_solve  lane r owns row r:  z_i = inv_i*(rhs_i - A_L z_{i-1})
//   backward_sweep lane r owns row r:  d_i = z_i - W_i d_{i+1};  y += d; weighted max norms
//
// No tensor cores: the blocks are 9..13 wide and the chain over nodes is sequential.
#pragma once
#include <cuda_pipeline.h>
#include "pnp_device.cuh"

namespace catint {

template <int NB, bool ST>
struct WarpState {
    double* y;      // current Newton iterate              [n*NB]   (shared or global)
    double* psi;    // rl1*zn1 - zn0 (mass rows)           [n*NB]
    double* zb;     // rhs -> z -> delta                   [n*NB]
    double* zn;     // Nordsieck array                     [LMAX][N] (global)
    double* ewt;    // error weights                       [N]
    double* fac;    // per node record [ inv_i: NB x NBP | per row l, a (A_L), ud, ua (A_U): NB x 4 ]  (global)
    double* W1;     // inv_1*A_U1' (dense because of the wall block)   [NB][NBP]
    double* V0;     // inv_0*A_E                           [NB][NBP]
    double* ring;   // shared: RING node records staged by cp.async ahead of the solve sweeps
    double* scratch;            // shared, per warp
    const CellSpecies* sp;      // shared, per warp
    const DevTables* tb;        // shared, per block
    CellScalars cs;
    int N;          // n*NB
    int lane;
    int vlane, vstride;   // element-wise passes: first index and stride (a warp pair splits them)
    int bar_id;           // named barrier of the warp pair (PAIR mode)
};

// barrier of the two warps that share a cell (64 threads, hardware barrier bar_id)
__device__ __forceinline__ void pair_barrier(int bar_id) {
    asm volatile("bar.sync %0, 64;" :: "r"(bar_id) : "memory");
}

template <int NB, bool ST>
__host__ __device__ constexpr int scratch_doubles() { return 2 * (NB + 2) + 4 * NB + 2; }
// sub-arrays of the per-cell workspace start on 32-byte boundaries
__host__ __device__ constexpr size_t align4(size_t doubles) { return (doubles + 3) & ~size_t(3); }
// padded row length of the stored blocks (even -> 16-byte aligned rows, double2 loads)
template <int NB, bool ST>
__host__ __device__ constexpr int padded() { return NB + (NB & 1); }
// doubles per node of the stored factors: inverse block (padded rows) + 4 coefficients per row
template <int NB, bool ST>
__host__ __device__ constexpr int fac_rec() { return NB * padded<NB, ST>() + NB * 4; }
constexpr int RING = 8;   // node records in flight in the solve sweeps (power of two)

// ---------------------------------------------------------------------------
// Gauss-Jordan elimination with threshold partial pivoting; lane j owns column j
// (columns >= NB are right-hand sides).  The pivot column travels through shared
// memory (one writer, broadcast reads).  A row swap is a warp-uniform branch, so
// it only costs when it happens.  Returns false on a zero/non-finite pivot.
template <int NB, bool ST>
__device__ __forceinline__ bool gauss_jordan(double (&A)[NB], int lane, double* pivbuf) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        double* buf = pivbuf + (k & 1) * (NB + 2);
        if (lane == k) {
            // magnitude keys: high word of |a| with the row index in the low 4 bits
            int best = ((__double2hiint(A[k]) & 0x7fffffff) & ~0xf) | k;
            const int diag = best;
#pragma unroll
            for (int r = k + 1; r < NB; ++r) {
                const int key = ((__double2hiint(A[r]) & 0x7fffffff) & ~0xf) | r;
                best = max(best, key);
            }
            // keep the diagonal unless another entry is more than 8x larger (3 exponent steps)
            int p = best & 0xf;
            if (diag + (3 << 20) >= best) p = k;
#pragma unroll
            for (int r = 0; r < NB; ++r) buf[r] = A[r];
            buf[NB] = (double)p;
        }
        __syncwarp();
        const int p = (int)buf[NB];
        double col[NB];
#pragma unroll
        for (int r = 0; r < NB; ++r) col[r] = buf[r];
        if (p != k) {           // warp-uniform
            double ck = col[k], ak = A[k];
#pragma unroll
            for (int r = k + 1; r < NB; ++r) {
                if (r == p) {
                    double t = col[r]; col[r] = ck; ck = t;
                    t = A[r]; A[r] = ak; ak = t;
                }
            }
            col[k] = ck; A[k] = ak;
        }
        const double ck = col[k];
        const double inv = __drcp_rn(ck);
        ok = ok && (ck != 0.0) && (fabs(inv) < 1e300);
        const double ak = A[k] * inv;
        A[k] = ak;
#pragma unroll
        for (int r = 0; r < NB; ++r)
            if (r != k) A[r] = fma(-col[r], ak, A[r]);
    }
    return ok;
}

// ---------------------------------------------------------------------------
// residual of row r at node i:  transport rows dc/dt, algebraic rows -constraint
// (the g-row of interior nodes is returned UNSCALED)
template <int NB, bool ST>
__device__ __forceinline__ double row_residual(const WarpState<NB, ST>& ws, const double* y, int i, int r) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    const int n = ws.cs.n;
    const DevTables& tb = *ws.tb;
    const double* y0 = y + (size_t)i * NB;
    constexpr int P = S + 1;      // phi unknown / row (Stern mode only)
    if (i == 0) {
        const WallCoef w = wall_coef(ws.cs);
        const double* y1 = y + NB;
        const double* y2 = y + 2 * NB;
        if (r < S)
            return (ws.sp->D[r] * ((y2[r] - y0[r]) * w.w0 + ws.sp->bq[r] * y1[r] * y1[S]) + ws.sp->J[r]) * w.ih0;
        if (r == S) return tb.use_migration ? -(y0[S] - y1[S] - (y1[S] - y2[S]) * w.ext) : -y0[S];
        // Robin (Stern layer) wall condition: eps*g_0 = -C_S*((phiM-phiPZC) - phi_0)
        return -((ws.cs.eps / ws.cs.cstern) * y0[S] + ws.cs.phi_wall - y0[P]);
    }
    if (i == n - 1) {
        if (r < S) return ws.sp->cb[r] - y0[r];
        if (!ST) return (tb.use_migration ? ws.cs.g_bulk : 0.0) - y0[S];
        if (r == S) return -y0[P];                                           // phi(L) = 0
        const NodeCoef k = interior_coef(ws.cs, i - 1);
        return -(y0[P] - (y0 - NB)[P] - y0[S] * k.hi);                        // phi recursion, h_{n-2}
    }
    const NodeCoef k = interior_coef(ws.cs, i);
    const double* ym = y0 - NB;
    const double* yp = y0 + NB;
    if (r < S) {
        double R = 0.0;
        for (int rr = 0; rr < tb.R; ++rr) {
            const double nu = tb.nu[rr][r];
            if (nu != 0.0) R += nu * net_rate(tb, rr, y0);
        }
        const double cm = ym[r], c0 = y0[r], cp = yp[r];
        return ws.sp->D[r] * (k.am * cm - (k.am + k.ap) * c0 + k.ap * cp
                              + ws.sp->bq[r] * k.ac * (cp * yp[S] - cm * ym[S])) + R;
    }
    if (r == S) {
        if (!tb.use_migration) return -y0[S];
        double rho = 0.0;
        for (int s = 0; s < S; ++s) rho += ws.sp->qe[s] * y0[s];
        return -(y0[S] - yp[S] - rho * k.hi);
    }
    return -(y0[P] - ym[P] - y0[S] * k.him);                                  // phi_i = phi_{i-1} + g_i*h_{i-1}
}

// Jacobian coefficients of row r (= lane) at node i, published through shared memory:
//   interior: sl,sa = dF_r/dc_{r,i-1}, dF_r/dg_{i-1};  sud,sua = dF_r/dc_{r,i+1}, dF_r/dg_{i+1}
//   wall:     sa = dF_r/dy_{r,0} (diagonal), sud,sua w.r.t. node 1, sl = dF_r/dy_{r,2} (extra block)
template <int NB, bool ST>
__device__ __forceinline__ void node_coeffs(const WarpState<NB, ST>& ws, const double* y, int i,
                                            double* sl, double* sa, double* sud, double* sua) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    const int r = ws.lane;
    const int n = ws.cs.n;
    const bool mig = ws.tb->use_migration;
    if (r >= NB) return;
    double l = 0.0, a = 0.0, ud = 0.0, ua = 0.0;
    const double* y0 = y + (size_t)i * NB;
    if (i == 0) {
        const WallCoef w = wall_coef(ws.cs);
        const double* y1 = y + NB;
        if (r < S) {
            const double Dr = ws.sp->D[r], bq = ws.sp->bq[r];
            ud = Dr * bq * y1[S] * w.ih0;
            ua = Dr * bq * y1[r] * w.ih0;
            l = Dr * w.w0 * w.ih0;
            a = -l;
        } else if (r == S) {
            if (mig) { ud = 1.0 + w.ext; l = -w.ext; }
            a = -1.0;
        } else {
            a = 1.0;                       // dF_phi/dphi_0 (the (phi,g) entry is added in the assembly)
        }
    } else if (i < n - 1) {
        const NodeCoef k = interior_coef(ws.cs, i);
        if (r < S) {
            const double Dr = ws.sp->D[r], bq = ws.sp->bq[r];
            const double* ym = y0 - NB;
            const double* yp = y0 + NB;
            l = Dr * (k.am - bq * k.ac * ym[S]);
            a = -Dr * bq * k.ac * ym[r];
            ud = Dr * (k.ap + bq * k.ac * yp[S]);
            ua = Dr * bq * k.ac * yp[r];
        } else if (r == S) {
            if (mig) ud = 1.0;
        } else {
            l = 1.0;                       // dF_phi/dphi_{i-1}
        }
    } else if (ST && r == S + 1) {
        l = 1.0;                           // bulk node, phi recursion row: dF/dphi_{n-2}
    }
    sl[r] = l; sa[r] = a; sud[r] = ud; sua[r] = ua;
}

// column j (< NB) of A_D = Mass*inv_gamma - dF_i/dy_i at an interior node, g-row scaled by sg
template <int NB, bool ST>
__device__ __forceinline__ void interior_diag_column(const WarpState<NB, ST>& ws, const double* yi, int j,
                                                     const NodeCoef& k, double inv_gamma, double sg,
                                                     double (&A)[NB]) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    const DevTables& tb = *ws.tb;
#pragma unroll
    for (int r = 0; r < NB; ++r) A[r] = 0.0;
    if (j < S) {
        for (int t = tb.tbeg[j]; t < tb.tbeg[j + 1]; ++t) {
            double v = tb.tcoef[t];
            if (tb.ti1[t] >= 0) v *= yi[tb.ti1[t]];
            if (tb.ti2[t] >= 0) v *= yi[tb.ti2[t]];
            if (tb.ti3[t] >= 0) v *= yi[tb.ti3[t]];
            const double* nur = tb.nu[tb.tr[t]];
#pragma unroll
            for (int r = 0; r < S; ++r) A[r] = fma(-nur[r], v, A[r]);
        }
        const double dd = inv_gamma + ws.sp->D[j] * (k.am + k.ap);
#pragma unroll
        for (int r = 0; r < S; ++r)
            if (r == j) A[r] += dd;
        if (tb.use_migration) A[S] = -tb.z[j];          // -(q_j*h/eps)*sg
    } else if (j == S) {
        A[S] = tb.use_migration ? sg : 1.0;
        if (ST) A[NB - 1] = -k.him;                // -dF_phi/dg_i
    } else {
        A[NB - 1] = 1.0;                           // -dF_phi/dphi_i
    }
}

// ---------------------------------------------------------------------------
// Factorisation sweep.  Lanes: D = 0..NB-1 (columns of A_D'), I = NB..2NB-1 (identity ->
// inverse), G = 2NB (g-column of A_U -> W[:,g]).  Stores inv_i (all nodes), the sparse
// coefficients of A_L/A_U (la), V_0 and the dense W_1.
// One-warp mode: all nodes top-down (mid < 0).  Twisted mode (warp pair): this is the TOP half, nodes
// 0..mid; at the coupling node `mid` the Schur complement of the bottom half arrives through the
// shared block xch = W^b_{mid+1} (written by factor_bottom, separated by a block barrier).
template <int NB, bool ST>
__device__ bool factor_sweep(WarpState<NB, ST>& ws, double inv_gamma, int mid = -1, const double* xch = nullptr) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    constexpr int NBP = padded<NB, ST>();
    const int lane = ws.lane;
    const int n = ws.cs.n;
    const bool mig = ws.tb->use_migration;
    double* pivbuf = ws.scratch;
    double* sl = ws.scratch + 2 * (NB + 2);
    double* sa = sl + NB; double* sud = sa + NB; double* sua = sud + NB;
    const bool isD = lane < NB, isI = lane >= NB && lane < 2 * NB, isG = lane == 2 * NB;
    const int j = isD ? lane : lane - NB;
    // source lane of the W column that D-lane j needs for the next Schur update
    const int wsrc = isD ? (j < S ? lane + NB : 2 * NB) : lane;
    double A[NB], Wp[NB];
#pragma unroll
    for (int r = 0; r < NB; ++r) Wp[r] = 0.0;
    bool ok = true;
    const double* y = ws.y;

    // Stern mode: the bulk node couples phi_{n-1} to phi_{n-2} and is eliminated like any other node
    const int n_fac = (mid >= 0) ? mid + 1 : (ST ? n : n - 1);
    for (int i = 0; i < n_fac; ++i) {
        __syncwarp();
        node_coeffs<NB, ST>(ws, y, i, sl, sa, sud, sua);
        __syncwarp();
        const double* yi = y + (size_t)i * NB;
        double* rec = ws.fac + (size_t)i * fac_rec<NB, ST>();
        double* invcol = rec + j;                                   // column j of inv_i (I lanes)
        const bool bulk = (i == n - 1);
        const NodeCoef k = bulk ? interior_coef(ws.cs, i - 1)
                                : ((i > 0) ? interior_coef(ws.cs, i) : NodeCoef{0, 0, 0, 1, 1});
        const double sg = (i > 0 && !bulk && mig) ? grow_scale(ws.cs, k.hi) : 1.0;
        if (lane < NB) {
            double4 v;
            v.x = sl[lane]; v.y = sa[lane];
            v.z = (lane == S) ? sud[lane] * sg : sud[lane];
            v.w = sua[lane];
            reinterpret_cast<double4*>(rec + NB * NBP)[lane] = v;
        }
#pragma unroll
        for (int r = 0; r < NB; ++r) A[r] = 0.0;

        if (i == 0) {
            // A_D0 is diag(mass*inv_gamma - sa); A_U0 = diag(-sud) + g-column(-sua); A_E = diag(-sl)
            if (isD || isI) {
#pragma unroll
                for (int r = 0; r < NB; ++r)
                    if (r == j) A[r] = isD ? ((j < S ? inv_gamma : 0.0) - sa[r]) : 1.0;
                if (ST && isD && j == S) A[NB - 1] = ws.cs.eps / ws.cs.cstern;      // -dF_phi/dg_0 (Robin row)
            } else if (isG) {
#pragma unroll
                for (int r = 0; r < NB; ++r) A[r] = r < S ? -sua[r] : -sud[r];
            }
            ok = gauss_jordan<NB, ST>(A, lane, pivbuf) && ok;
            double Wc[NB];
#pragma unroll
            for (int r = 0; r < NB; ++r) Wc[r] = A[r];
            if (isI) {
                const double ud = -sud[j], ae = -sl[j];
                double* v0col = ws.V0 + j;
#pragma unroll
                for (int r = 0; r < NB; ++r) {
                    invcol[r * NBP] = A[r];
                    v0col[r * NBP] = A[r] * ae;
                    if (j < S) Wc[r] = A[r] * ud;
                }
            }
#pragma unroll
            for (int r = 0; r < NB; ++r) Wp[r] = __shfl_sync(FULL, Wc[r], wsrc);
            // the I lanes keep V_0 columns for the modified A_U of node 1
            if (isI) {
                const double ae = -sl[j];
#pragma unroll
                for (int r = 0; r < NB; ++r) Wp[r] = A[r] * ae;
            }
            continue;
        }

        // ---- interior node: column j of [A_D' | I | u_g] ----
        double Dsave[NB];
        const bool couple = (i == mid);
        if (couple) pair_barrier(ws.bar_id);   // the bottom half has published W^b_{mid+1}
        if (isD) {
            if (!bulk) {
                interior_diag_column<NB, ST>(ws, yi, j, k, inv_gamma, sg, A);
            } else {
                // Stern bulk node: c rows identity; row S: phi_{n-1} = 0; row P: phi recursion with h_{n-2}
#pragma unroll
                for (int r = 0; r < S; ++r)
                    if (r == j) A[r] = 1.0;
                if (j == S) A[NB - 1] = -k.hi;
                if (j == NB - 1) { A[S] = 1.0; A[NB - 1] = 1.0; }
            }
            const double wg = Wp[S];
#pragma unroll
            for (int r = 0; r < S; ++r) A[r] += sl[r] * Wp[r] + sa[r] * wg;     // A_D - A_L*W_{i-1}
            if (ST) A[NB - 1] += sl[NB - 1] * Wp[NB - 1];                       // phi row: A_L = -1 on the diagonal
            if (couple) {
                // - A_U * W^b_{mid+1}:  A_U = -(diag ud + ua e_g^T), g row: -ud*sg
                const double* xc = xch + j;
                const double xg = xc[S * NBP];
#pragma unroll
                for (int r = 0; r < S; ++r) A[r] += sud[r] * xc[r * NBP] + sua[r] * xg;
                A[S] += sud[S] * sg * xg;
            }
            if (i == 1) {
#pragma unroll
                for (int r = 0; r < NB; ++r) Dsave[r] = A[r];
            }
        } else if (isI) {
#pragma unroll
            for (int r = 0; r < NB; ++r)
                if (r == j) A[r] = 1.0;
        } else if (isG) {
#pragma unroll
            for (int r = 0; r < NB; ++r) A[r] = r < S ? -sua[r] : -sud[r] * sg;
        }
        ok = gauss_jordan<NB, ST>(A, lane, pivbuf) && ok;
        double Wc[NB];
#pragma unroll
        for (int r = 0; r < NB; ++r) Wc[r] = A[r];
        if (isI) {
            const double ud = -sud[j];
#pragma unroll
            for (int r = 0; r < NB; ++r) {
                invcol[r * NBP] = A[r];
                if (j < S) Wc[r] = A[r] * ud;
            }
        }
        if (i == 1) {
            // A_U of node 1 is modified by the wall block: A_U1' = A_U1 - A_L1*V_0 (dense in
            // general), so W_1 = inv_1*A_U1' comes from a second elimination [A_D1' | A_U1']
            __syncwarp();
#pragma unroll
            for (int r = 0; r < NB; ++r) A[r] = 0.0;
            if (isD) {
#pragma unroll
                for (int r = 0; r < NB; ++r) A[r] = Dsave[r];
            } else if (isI) {
                const double vg = Wp[S];
#pragma unroll
                for (int r = 0; r < NB; ++r) {
                    double v = 0.0;
                    if (r == j) v = (r < S) ? -sud[r] : -sud[r] * sg;
                    if (j == S && r < S) v = -sua[r];
                    if (r < S) v += sl[r] * Wp[r] + sa[r] * vg;
                    if (ST && r == NB - 1) v += sl[r] * Wp[r];
                    A[r] = v;
                }
            }
            ok = gauss_jordan<NB, ST>(A, lane, pivbuf) && ok;
#pragma unroll
            for (int r = 0; r < NB; ++r) Wc[r] = A[r];
            if (isI) {
                double* w1col = ws.W1 + j;
#pragma unroll
                for (int r = 0; r < NB; ++r) w1col[r * NBP] = A[r];
            }
            const int src1 = isD ? lane + NB : lane;
#pragma unroll
            for (int r = 0; r < NB; ++r) Wp[r] = __shfl_sync(FULL, Wc[r], src1);
            continue;
        }
#pragma unroll
        for (int r = 0; r < NB; ++r) Wp[r] = __shfl_sync(FULL, Wc[r], wsrc);
    }
    // bulk node (default Poisson BCs): identity rows, no coupling
    if (!ST && mid < 0) {
        const int i = n - 1;
        double* rec = ws.fac + (size_t)i * fac_rec<NB, ST>();
        for (int e = lane; e < NB * NBP; e += 32) rec[e] = (e / NBP == e % NBP) ? 1.0 : 0.0;
        if (lane < NB) reinterpret_cast<double4*>(rec + NB * NBP)[lane] = make_double4(0, 0, 0, 0);
    }
    __syncwarp();
    return __all_sync(FULL, ok);
}

// ---------------------------------------------------------------------------
// BOTTOM half of the twisted factorisation (second warp of a pair): nodes n-1 down to mid+1 are
// eliminated upwards,  A_D'_i = A_D,i - A_U,i * W^b_{i+1},  W^b_i = inv_i * A_L,i  with
// A_L = -(diag l + a e_g^T).  Same lane roles and node records as factor_sweep.  Publishes
// W^b_{mid+1} in xch for the coupling node and joins the block barrier.
template <int NB, bool ST>
__device__ bool factor_bottom(WarpState<NB, ST>& ws, double inv_gamma, int mid, double* xch) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    constexpr int NBP = padded<NB, ST>();
    const int lane = ws.lane;
    const int n = ws.cs.n;
    const bool mig = ws.tb->use_migration;
    double* pivbuf = ws.scratch;
    double* sl = ws.scratch + 2 * (NB + 2);
    double* sa = sl + NB; double* sud = sa + NB; double* sua = sud + NB;
    const bool isD = lane < NB, isI = lane >= NB && lane < 2 * NB, isG = lane == 2 * NB;
    const int j = isD ? lane : lane - NB;
    const int wsrc = isD ? (j != S ? lane + NB : 2 * NB) : lane;
    double A[NB], Wp[NB];
#pragma unroll
    for (int r = 0; r < NB; ++r) Wp[r] = 0.0;
    bool ok = true;
    const double* y = ws.y;
    for (int i = n - 1; i > mid; --i) {
        __syncwarp();
        node_coeffs<NB, ST>(ws, y, i, sl, sa, sud, sua);
        __syncwarp();
        const double* yi = y + (size_t)i * NB;
        double* rec = ws.fac + (size_t)i * fac_rec<NB, ST>();
        double* invcol = rec + j;
        const bool bulk = (i == n - 1);
        const NodeCoef k = bulk ? interior_coef(ws.cs, i - 1) : interior_coef(ws.cs, i);
        const double sg = (!bulk && mig) ? grow_scale(ws.cs, k.hi) : 1.0;
        if (lane < NB) {
            double4 v;
            v.x = sl[lane]; v.y = sa[lane];
            v.z = (lane == S) ? sud[lane] * sg : sud[lane];
            v.w = sua[lane];
            reinterpret_cast<double4*>(rec + NB * NBP)[lane] = v;
        }
#pragma unroll
        for (int r = 0; r < NB; ++r) A[r] = 0.0;
        if (isD) {
            if (!bulk) {
                interior_diag_column<NB, ST>(ws, yi, j, k, inv_gamma, sg, A);
                // - A_U * W^b_{i+1}
                const double wg = Wp[S];
#pragma unroll
                for (int r = 0; r < S; ++r) A[r] += sud[r] * Wp[r] + sua[r] * wg;
                A[S] += sud[S] * sg * wg;
            } else {
#pragma unroll
                for (int r = 0; r < S; ++r)
                    if (r == j) A[r] = 1.0;
                if (!ST) {
                    if (j == S) A[S] = 1.0;
                } else {
                    if (j == S) A[NB - 1] = -k.hi;
                    if (j == NB - 1) { A[S] = 1.0; A[NB - 1] = 1.0; }
                }
            }
        } else if (isI) {
#pragma unroll
            for (int r = 0; r < NB; ++r)
                if (r == j) A[r] = 1.0;
        } else if (isG) {
            // g-column of A_L: -a_r on the transport rows
#pragma unroll
            for (int r = 0; r < NB; ++r) A[r] = r < S ? -sa[r] : 0.0;
        }
        ok = gauss_jordan<NB, ST>(A, lane, pivbuf) && ok;
        double Wc[NB];
#pragma unroll
        for (int r = 0; r < NB; ++r) Wc[r] = A[r];
        if (isI) {
            const double lj = -sl[j];
#pragma unroll
            for (int r = 0; r < NB; ++r) {
                invcol[r * NBP] = A[r];
                if (j != S) Wc[r] = A[r] * lj;
            }
        }
#pragma unroll
        for (int r = 0; r < NB; ++r) Wp[r] = __shfl_sync(FULL, Wc[r], wsrc);
    }
    if (isD) {
#pragma unroll
        for (int r = 0; r < NB; ++r) xch[r * NBP + j] = Wp[r];
    }
    pair_barrier(ws.bar_id);
    return __all_sync(FULL, ok);
}

// ---------------------------------------------------------------------------
// rhs = F(y) - Mass*(y+psi)*inv_gamma for every unknown; one lane per node.
// The g-row of interior nodes carries the same scale as in factor_sweep.
template <int NB, bool ST>
__device__ void residual_pass(WarpState<NB, ST>& ws, double inv_gamma) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    const int n = ws.cs.n;
    const DevTables& tb = *ws.tb;
    const double* y = ws.y;
    for (int i = ws.vlane; i < n; i += ws.vstride) {
        const double* y0 = y + (size_t)i * NB;
        double* out = ws.zb + (size_t)i * NB;
        if (i == 0 || i == n - 1) {
#pragma unroll
            for (int r = 0; r < NB; ++r) {
                double v = row_residual<NB, ST>(ws, y, i, r);
                out[r] = v;
            }
            continue;
        }
        const NodeCoef k = interior_coef(ws.cs, i);
        const double* ym = y0 - NB;
        const double* yp = y0 + NB;
        const double gm = ym[S], gp = yp[S];
        double F[S];
        double rho = 0.0;
#pragma unroll
        for (int r = 0; r < S; ++r) {
            const double cm = ym[r], c0 = y0[r], cp = yp[r];
            F[r] = ws.sp->D[r] * (k.am * cm - (k.am + k.ap) * c0 + k.ap * cp
                                  + ws.sp->bq[r] * k.ac * (cp * gp - cm * gm));
            rho = fma(ws.sp->qe[r], c0, rho);
        }
        for (int rr = 0; rr < tb.R; ++rr) {
            // educt/product indices of one reaction as two packed words (one LDS each, uniform)
            const unsigned ew = *reinterpret_cast<const unsigned*>(tb.ed[rr]);
            const unsigned pw = *reinterpret_cast<const unsigned*>(tb.pr[rr]);
            const int ne = tb.ned[rr], np = tb.npr[rr];
            double f = tb.kf[rr], b = tb.kr[rr];
#pragma unroll
            for (int e = 0; e < MAXRT; ++e) {
                if (e < ne) f *= y0[(ew >> (8 * e)) & 0xff];
                if (e < np) b *= y0[(pw >> (8 * e)) & 0xff];
            }
            const double net = f - b;
            const double* nur = tb.nu[rr];
#pragma unroll
            for (int r = 0; r < S; ++r) F[r] = fma(nur[r], net, F[r]);
        }
#pragma unroll
        for (int r = 0; r < S; ++r) out[r] = F[r];
        if (tb.use_migration) out[S] = -(y0[S] - gp - rho * k.hi) * grow_scale(ws.cs, k.hi);
        else out[S] = -y0[S];
        if (ST) out[NB - 1] = -(y0[NB - 1] - ym[NB - 1] - y0[S] * k.him);
    }
    __syncwarp();
    if (inv_gamma != 0.0) {
        // mass term of the BDF corrector, psi streamed from global memory; every lane finishes the
        // node records it produced above (no cross-lane dependency, hence no barrier in between)
        for (int i = ws.vlane; i < n - 1; i += ws.vstride) {
#pragma unroll
            for (int r = 0; r < S; ++r) {
                const size_t idx = (size_t)i * NB + r;
                ws.zb[idx] -= (y[idx] + ws.psi[idx]) * inv_gamma;
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// Solve sweeps with the stored factors.  The node records stream from global memory (L2/HBM)
// into a shared-memory ring by cp.async, RING-1 nodes ahead of use, so that the sequential
// chain over the nodes never waits for DRAM.  Lane r (< NB) owns row r.
//
// All sweeps take a node range and a direction so that the same code serves the one-warp
// elimination (top-down over all nodes) and the twisted elimination of a warp pair: the "top"
// warp eliminates nodes 0..m-1 downwards, the "bottom" warp nodes n-1..m+1 upwards, node m
// couples the two halves (see pnp_kernels.cuh, PAIR).
template <int NB, bool ST>
struct FactorRow {
    double v[NB];
    double4 co;      // l, a (A_L) and ud, ua (A_U) of this row
};

template <int NB, bool ST>
__device__ __forceinline__ void ring_row(const WarpState<NB, ST>& ws, int slot, int r, FactorRow<NB, ST>& f) {
    constexpr int NBP = padded<NB, ST>();
    const double* rec = ws.ring + (size_t)slot * fac_rec<NB, ST>();
    const double2* p = reinterpret_cast<const double2*>(rec + r * NBP);
#pragma unroll
    for (int c = 0; c < NBP / 2; ++c) {
        const double2 t = p[c];
        f.v[2 * c] = t.x;
        if (2 * c + 1 < NB) f.v[2 * c + 1] = t.y;
    }
    f.co = reinterpret_cast<const double4*>(rec + NB * NBP)[r];
}

template <int NB, bool ST>
__device__ __forceinline__ double row_dot(const FactorRow<NB, ST>& f, const double* tt) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int c = 0; c < NB; c += 2) {
        s0 = fma(f.v[c], tt[c], s0);
        if (c + 1 < NB) s1 = fma(f.v[c + 1], tt[c + 1], s1);
    }
    return s0 + s1;
}

// elimination of the right-hand side over `count` nodes starting at `first` in direction dir:
//   dir=+1:  z_i = inv_i*(rhs_i - A_L z_{i-1})      (the first node of the range has no predecessor
//   dir=-1:  z_i = inv_i*(rhs_i - A_U z_{i+1})       inside the range: its coupling term is skipped)
// zb <- z.
template <int NB, bool ST>
__device__ void forward_solve(WarpState<NB, ST>& ws, int first, int count, int dir) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    constexpr int REC = fac_rec<NB, ST>();
    const int lane = ws.lane;
    const int n = ws.cs.n;
    const bool act = lane < NB;
    const int r = act ? lane : 0;
    double* tbuf = ws.scratch;           // 2*NB doubles (the pivot buffer is free here)
    auto issue = [&](int k) {
        const int i = first + dir * k;
        if (k < count && i >= 0 && i < n) {
            const double* src = ws.fac + (size_t)i * REC;
            double* dst = ws.ring + (size_t)(k & (RING - 1)) * REC;
            for (int c = lane; c < REC / 2; c += 32) __pipeline_memcpy_async(dst + 2 * c, src + 2 * c, 16);
        }
        __pipeline_commit();
    };
#pragma unroll
    for (int p = 0; p < RING - 1; ++p) issue(p);
    double zprev = 0.0;
    for (int k = 0; k < count; ++k) {
        const int i = first + dir * k;
        __pipeline_wait_prior(RING - 2);
        __syncwarp();
        FactorRow<NB, ST> f;
        double z = 0.0;
        if (act) {
            ring_row<NB, ST>(ws, k & (RING - 1), r, f);
            double t = ws.zb[(size_t)i * NB + r];
            if (k > 0) {
                const double zs = ws.zb[(size_t)(i - dir) * NB + S];
                if (dir > 0) t = fma(f.co.x, zprev, fma(f.co.y, zs, t));      // rhs - A_L z_{i-1}
                else t = fma(f.co.z, zprev, fma(f.co.w, zs, t));              // rhs - A_U z_{i+1}
            }
            tbuf[(k & 1) * NB + r] = t;
        }
        __syncwarp();
        if (act) {
            z = row_dot<NB, ST>(f, tbuf + (k & 1) * NB);
            ws.zb[(size_t)i * NB + r] = z;
        }
        zprev = z;
        issue(k + RING - 1);
    }
    __pipeline_wait_prior(0);
    __syncwarp();
}

// Back substitution over `count` nodes starting at `first` in direction dir, the solution of the
// node before `first` (first-dir) being final already in zb:
//   dir=-1:  d_i = z_i - inv_i*(A_U d_{i+1})   (node 1: dense W_1; node 0: extra wall block V_0)
//   dir=+1:  d_i = z_i - inv_i*(A_L d_{i-1})
// y += scale*d, zb <- d.  Fused with the weighted max norms of the Newton update (|scale*d|*w)
// and of the accumulated correction (|y-zn0|*w) over the error-controlled unknowns
// (concentrations of nodes 0..n-2), accumulated into dmax/amax (per lane, reduce afterwards):
// the weights and zn0 of each node ride in a second cp.async ring.
// wmode 0: weights ws.ewt; wmode 1 (steady polish): w = 1/(prtol*|y|+patol).
template <int NB, bool ST>
__device__ void backward_solve(WarpState<NB, ST>& ws, double scale, int first, int count, int dir,
                               double& dmax, double& amax, int wmode, double prtol, double patol) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    constexpr int NBP = padded<NB, ST>();
    constexpr int REC = fac_rec<NB, ST>();
    constexpr int R2 = 2 * NBP;                  // doubles per node in the weight ring
    const int lane = ws.lane;
    const int n = ws.cs.n;
    const bool act = lane < NB;
    const int r = act ? lane : 0;
    double* tbuf = ws.scratch;
    double* ring2 = ws.ring + (size_t)RING * REC;
    auto issue = [&](int k) {
        const int i = first + dir * k;
        if (k < count && i >= 0 && i < n) {
            const int slot = k & (RING - 1);
            const double* src = ws.fac + (size_t)i * REC;
            double* dst = ws.ring + (size_t)slot * REC;
            for (int c = lane; c < REC / 2; c += 32) __pipeline_memcpy_async(dst + 2 * c, src + 2 * c, 16);
            if (wmode == 0 && lane < 2 * NB) {
                const double* s2 = (lane < NB ? ws.ewt : ws.zn - NB) + (size_t)i * NB + lane;
                __pipeline_memcpy_async(ring2 + (size_t)slot * R2 + (lane < NB ? lane : NBP + lane - NB), s2, 8);
            }
        }
        __pipeline_commit();
    };
#pragma unroll
    for (int p = 0; p < RING - 1; ++p) issue(p);
    for (int k = 0; k < count; ++k) {
        const int i = first + dir * k;
        __pipeline_wait_prior(RING - 2);
        __syncwarp();
        const int slot = k & (RING - 1);
        FactorRow<NB, ST> f;
        double d = 0.0;
        if (dir < 0 && i == 1) {
            if (act) {
                const double* dn = ws.zb + 2 * NB;
                const double* Wr = ws.W1 + (size_t)r * NBP;
                double s = ws.zb[NB + r];
#pragma unroll
                for (int c = 0; c < NB; ++c) s = fma(-Wr[c], dn[c], s);
                d = s;
            }
        } else {
            if (act) {
                ring_row<NB, ST>(ws, slot, r, f);
                const double* dn = ws.zb + (size_t)(i - dir) * NB;       // final solution of the neighbour
                double t;
                if (dir < 0) t = -(f.co.z * dn[r] + f.co.w * dn[S]);      // A_U d_{i+1} (ua = 0 off the c rows)
                else t = -(f.co.x * dn[r] + f.co.y * dn[S]);              // A_L d_{i-1}
                tbuf[(k & 1) * NB + r] = t;
            }
            __syncwarp();
            if (act) {
                d = ws.zb[(size_t)i * NB + r] - row_dot<NB, ST>(f, tbuf + (k & 1) * NB);
                if (dir < 0 && i == 0) {
                    const double* d2 = ws.zb + 2 * NB;
                    const double* Vr = ws.V0 + (size_t)r * NBP;
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < NB; ++c) s = fma(Vr[c], d2[c], s);
                    d -= s;
                }
            }
        }
        if (act) {
            const size_t idx = (size_t)i * NB + r;
            ws.zb[idx] = d;
            const double ds = d * scale;
            const double yn = ws.y[idx] + ds;
            ws.y[idx] = yn;
            if (r < S && i < n - 1) {
                double w, z0 = 0.0;
                if (wmode == 0) { w = ring2[(size_t)slot * R2 + r]; z0 = ring2[(size_t)slot * R2 + NBP + r]; }
                else w = 1.0 / (prtol * fabs(yn) + patol);
                double ad = fabs(ds) * w;
                if (!(ad <= 1e300)) ad = INFINITY;          // NaN/Inf must not be lost in fmax
                dmax = fmax(dmax, ad);
                if (wmode == 0) amax = fmax(amax, fabs(yn - z0) * w);
            }
        }
        issue(k + RING - 1);
    }
    __pipeline_wait_prior(0);
    __syncwarp();
}

// update of one node whose solution d already sits in zb (the bulk node of the one-warp sweep,
// the coupling node of the twisted sweep): y += scale*d and the norms.
template <int NB, bool ST>
__device__ void apply_node(WarpState<NB, ST>& ws, double scale, int i, double& dmax, double& amax,
                           int wmode, double prtol, double patol) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    const int r = ws.lane;
    if (r < NB) {
        const size_t idx = (size_t)i * NB + r;
        const double ds = ws.zb[idx] * scale;
        const double yn = ws.y[idx] + ds;
        ws.y[idx] = yn;
        if (r < S && i < ws.cs.n - 1) {
            const double w = (wmode == 0) ? ws.ewt[idx] : 1.0 / (prtol * fabs(yn) + patol);
            double ad = fabs(ds) * w;
            if (!(ad <= 1e300)) ad = INFINITY;
            dmax = fmax(dmax, ad);
            if (wmode == 0) amax = fmax(amax, fabs(yn - ws.zn[idx]) * w);
        }
    }
    __syncwarp();
}

// solution of the coupling node m of the twisted sweep (done by the top warp):
//   d_m = inv_m*(rhs_m - A_L z_{m-1} - A_U z_{m+1}),  zb[m] <- d_m
template <int NB, bool ST>
__device__ void solve_middle(WarpState<NB, ST>& ws, int m) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    constexpr int NBP = padded<NB, ST>();
    const int lane = ws.lane;
    const bool act = lane < NB;
    const int r = act ? lane : 0;
    double* tbuf = ws.scratch;
    const double* rec = ws.fac + (size_t)m * fac_rec<NB, ST>();
    double row[NB];
    double4 co = make_double4(0, 0, 0, 0);
    if (act) {
#pragma unroll
        for (int c = 0; c < NB; ++c) row[c] = rec[r * NBP + c];
        co = reinterpret_cast<const double4*>(rec + NB * NBP)[r];
        const double* zm = ws.zb + (size_t)(m - 1) * NB;
        const double* zp = ws.zb + (size_t)(m + 1) * NB;
        tbuf[r] = ws.zb[(size_t)m * NB + r] + co.x * zm[r] + co.y * zm[S] + co.z * zp[r] + co.w * zp[S];
    }
    __syncwarp();
    if (act) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < NB; ++c) s = fma(row[c], tbuf[c], s);
        ws.zb[(size_t)m * NB + r] = s;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------
// Weighted max norms of the Newton update (scale*zb) and of the accumulated correction
// (y - zn0) over the error-controlled unknowns (concentrations of nodes 0..n-2).
// wmode 0: weights ws.ewt; wmode 1 (steady polish): w = 1/(prtol*|y|+patol).
template <int NB, bool ST>
__device__ void newton_norms(const WarpState<NB, ST>& ws, double scale, bool want_acn, double& dnorm, double& anorm,
                             int wmode, double prtol, double patol) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    const int n = ws.cs.n;
    double dmax = 0.0, amax = 0.0;
    for (int idx = ws.lane; idx < ws.N; idx += 32) {
        const int i = idx / NB, r = idx - i * NB;
        if (r < S && i < n - 1) {
            const double yv = ws.y[idx];
            const double w = (wmode == 0) ? ws.ewt[idx] : 1.0 / (prtol * fabs(yv) + patol);
            double ad = fabs(ws.zb[idx] * scale) * w;
            if (!(ad <= 1e300)) ad = INFINITY;          // NaN/Inf must not be lost in fmax
            dmax = fmax(dmax, ad);
            if (want_acn) amax = fmax(amax, fabs(yv - ws.zn[idx]) * w);
        }
    }
    dnorm = warp_max(dmax);
    anorm = warp_max(amax);
}

}  // namespace catint
