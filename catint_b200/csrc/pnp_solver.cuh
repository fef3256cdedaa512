// pnp_solver.cuh -- fused per-cell implicit integrator (K2+K3 inside a BDF/Newton loop).
//
// One warp integrates one cell from t=0 to t_end: variable-order (1..5),
// variable-step BDF in Nordsieck form with LSODA/CVODE-style step and order
// control (weighted max norm over the concentrations, rtol/atol as passed to
// scipy odeint at /root/reference/catint/calculator_old.py:947), Newton corrector
// with the analytic block-tridiagonal Jacobian, block-Thomas solve with partial
// pivoting inside the NBxNB blocks.  No tensor cores: the blocks are 9..13 wide
// and the chain over nodes is sequential.
//
// Lane mapping of the forward (elimination) sweep: the augmented block row
// [A_D | A_U | rhs] has 2*NB+1 columns; lane j owns column j (NB registers).
// Gauss-Jordan with row pivoting turns it into [I | W_i | z_i]; W_i goes to
// global memory for the backward sweep, its columns are handed to the A_D lanes
// of the next node by warp shuffles for the Schur update A_D' = A_D - A_L*W.
// A_L and A_U are diagonal + one column (g) (+ one g-row entry), which keeps the
// Schur update at O(NB) per lane.
// Backward sweep: lane r owns row r:  d_i[r] = z_i[r] - sum_j W_i[r][j]*d_{i+1}[j].
#pragma once
#include "pnp_device.cuh"

namespace catint {

constexpr int QMAX = 5;
constexpr int LMAX = QMAX + 1;     // Nordsieck vectors zn[0..QMAX]
constexpr unsigned FULL = 0xffffffffu;

// integrator constants (VODE/CVODE family)
constexpr double ADDON = 1e-6, BIAS1 = 6.0, BIAS2 = 6.0, BIAS3 = 10.0;
constexpr double ETAMX1 = 1e4, ETAMX2 = 10.0, ETAMXF = 0.2, ETAMIN = 0.1, ETACF = 0.25, THRESH = 1.5;
constexpr int MXNCF = 10, MXNEF = 7, MXNEF1 = 3, SMALL_NEF = 2, LONG_WAIT = 10;
constexpr int MAXCOR = 3;
constexpr double CRDOWN = 0.3, RDIV = 2.0, NLSCOEF = 0.1;

struct SolveParams {
    DevTables tb;
    // per-cell inputs
    const double* par; const int* nx; const int* mesh_id; const double* mesh_xi;
    const double* y0;          // optional [B][nx_max][S]
    long long n_cells;
    // control
    int mode, max_steps, n_out, polish_max_iter;
    double rtol, atol, h0, polish_rtol;
    const double* t_out;       // device [n_out]
    // outputs
    double* c_out; double* phi_out; double* g_out; double* flux_out;
    int* status; int* n_steps; int* n_newton;
    // workspace (global): per cell  zn[LMAX][N], ewt[N], W[nx_max][NB][NB], V0[NB][NB] (+ y,psi,zb if not in smem)
    double* ws; long long ws_stride;   // doubles per cell
    int state_in_smem;
};

// ---------------------------------------------------------------------------
template <int NB>
struct WarpState {
    // pointers (shared or global)
    double* y;      // current Newton iterate              [n*NB]
    double* psi;    // rl1*zn1 - zn0 (mass rows)           [n*NB]
    double* zb;     // rhs -> z -> delta                   [n*NB]
    double* zn;     // Nordsieck array                     [LMAX][N]
    double* ewt;    // error weights                       [N]
    double* W;      // [n][NB][NB]
    double* V0;     // [NB][NB]
    double* scratch;            // shared, per warp: 2*(NB+1) + 6*NB doubles
    const CellSpecies* sp;      // shared, per warp
    const DevTables* tb;        // shared, per block
    CellScalars cs;
    int N;          // n*NB
    int lane;
};

template <int NB>
__device__ __forceinline__ int scratch_doubles() { return 2 * (NB + 1) + 6 * NB; }

// ---------------------------------------------------------------------------
// Gauss-Jordan elimination with partial (row) pivoting, lane j owns column j of
// the NB x (2*NB+1) augmented block.  On exit columns NB..2NB hold
// A_D^{-1}*[A_U | rhs].  Returns false on a zero/non-finite pivot.
template <int NB>
__device__ __forceinline__ bool gauss_jordan(double (&A)[NB], int lane, double* pivbuf) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < NB; ++k) {
        double* buf = pivbuf + (k & 1) * (NB + 1);
        if (lane == k) {
            double best = fabs(A[k]);
            int p = k;
#pragma unroll
            for (int r = k + 1; r < NB; ++r) {
                const double v = fabs(A[r]);
                if (v > best) { best = v; p = r; }
            }
#pragma unroll
            for (int r = 0; r < NB; ++r) buf[r] = A[r];
            buf[NB] = (double)p;
        }
        __syncwarp();
        const int p = (int)buf[NB];
        double col[NB];
#pragma unroll
        for (int r = 0; r < NB; ++r) col[r] = buf[r];
        // swap rows k and p (p >= k)
        double ck = col[k], ak = A[k];
#pragma unroll
        for (int r = k + 1; r < NB; ++r) {
            if (r == p) {
                double t = col[r]; col[r] = ck; ck = t;
                t = A[r]; A[r] = ak; ak = t;
            }
        }
        const double inv = 1.0 / ck;
        ok = ok && (ck != 0.0) && (fabs(inv) < 1e300);
        ak *= inv;
        A[k] = ak;
#pragma unroll
        for (int r = 0; r < NB; ++r)
            if (r != k) A[r] = fma(-col[r], ak, A[r]);
    }
    return ok;
}

// ---------------------------------------------------------------------------
// Row quantities of node i, computed by lane r (< NB); everything the column
// assembly needs is published through the per-warp scratch:
//   sl[r], sa[r]   A_L: diagonal and g-column entries of row r
//   sud[r], sua[r] A_U: diagonal and g-column entries of row r
//   sf[r]          residual F_r (transport rows: dc/dt; algebraic rows: -constraint)
// Returns F_r for lane r (0 for other lanes).
template <int NB>
__device__ __forceinline__ double node_rows(const WarpState<NB>& ws, const double* y, int i,
                                            double* sl, double* sa, double* sud, double* sua, double* sf) {
    constexpr int S = NB - 1;
    const int r = ws.lane;
    const int n = ws.cs.n;
    const DevTables& tb = *ws.tb;
    double F = 0.0, l = 0.0, a = 0.0, ud = 0.0, ua = 0.0;
    if (r < NB) {
        const double* y0 = y + (size_t)i * NB;
        if (i == 0) {
            const WallCoef w = wall_coef(ws.cs);
            const double* y1 = y + NB;
            const double* y2 = y + 2 * NB;
            if (r < S) {
                const double Dr = ws.sp->D[r], bq = tb.use_migration ? ws.sp->bq[r] : 0.0;
                const double g1 = y1[S];
                F = (Dr * ((y2[r] - y0[r]) * w.w0 + bq * y1[r] * g1) + ws.sp->J[r]) * w.ih0;
                ud = Dr * bq * g1 * w.ih0;         // dF/dc_{r,1}
                ua = Dr * bq * y1[r] * w.ih0;      // dF/dg_1
                l = Dr * w.w0 * w.ih0;             // dF/dc_{r,2}   (extra wall block, diagonal)
                a = -Dr * w.w0 * w.ih0;            // dF/dc_{r,0}   (A_D diagonal, without mass)
            } else {   // g row: g_0 - g_1 - (g_1-g_2)*ext = 0
                if (tb.use_migration) {
                    F = -(y0[S] - y1[S] - (y1[S] - y2[S]) * w.ext);
                    ud = 1.0 + w.ext;              // dF/dg_1
                    l = -w.ext;                    // dF/dg_2
                } else {
                    F = -y0[S];
                }
                a = -1.0;                          // dF/dg_0
            }
        } else if (i == n - 1) {
            if (r < S) F = ws.sp->cb[r] - y0[r];
            else F = (tb.use_migration ? ws.cs.g_bulk : 0.0) - y0[S];
        } else {
            const NodeCoef k = interior_coef(ws.cs, i);
            const double* ym = y0 - NB;
            const double* yp = y0 + NB;
            if (r < S) {
                const double Dr = ws.sp->D[r], bq = tb.use_migration ? ws.sp->bq[r] : 0.0;
                const double cm = ym[r], c0 = y0[r], cp = yp[r], gm = ym[S], gp = yp[S];
                const double R = tb.R ? reaction_source(tb, r, y0) : 0.0;
                F = Dr * (k.am * cm - (k.am + k.ap) * c0 + k.ap * cp + bq * k.ac * (cp * gp - cm * gm)) + R;
                l = Dr * (k.am - bq * k.ac * gm);      // dF/dc_{r,i-1}
                a = -Dr * bq * k.ac * cm;              // dF/dg_{i-1}
                ud = Dr * (k.ap + bq * k.ac * gp);     // dF/dc_{r,i+1}
                ua = Dr * bq * k.ac * cp;              // dF/dg_{i+1}
            } else {   // g row: g_i - g_{i+1} - rho_i*h_i = 0, rho = sum q c / eps
                if (tb.use_migration) {
                    double rho = 0.0;
                    for (int s = 0; s < S; ++s) rho += ws.sp->q[s] * y0[s];
                    rho /= ws.cs.eps;
                    F = -(y0[S] - yp[S] - rho * k.hi);
                    ud = 1.0;                          // dF/dg_{i+1}
                } else {
                    F = -y0[S];
                }
            }
        }
        sl[r] = l; sa[r] = a; sud[r] = ud; sua[r] = ua; sf[r] = F;
    }
    return F;
}

// ---------------------------------------------------------------------------
// One Newton linear solve  (Mass*inv_gamma - dF/dy) * delta = F - Mass*(y+psi)*inv_gamma,
// forward elimination part.  z_i -> zb, W_i -> global.  Returns false if a
// block was singular.
template <int NB>
__device__ bool forward_sweep(WarpState<NB>& ws, double inv_gamma) {
    constexpr int S = NB - 1;
    const int lane = ws.lane;
    const int n = ws.cs.n;
    const DevTables& tb = *ws.tb;
    double* pivbuf = ws.scratch;
    double* sl = ws.scratch + 2 * (NB + 1);
    double* sa = sl + NB; double* sud = sa + NB; double* sua = sud + NB; double* sf = sua + NB;
    const bool isD = lane < NB, isU = lane >= NB && lane < 2 * NB, isR = lane == 2 * NB;
    const int j = isD ? lane : lane - NB;    // column index inside its block
    double A[NB], Wp[NB];
#pragma unroll
    for (int r = 0; r < NB; ++r) Wp[r] = 0.0;
    bool ok = true;
    const double* y = ws.y;

    for (int i = 0; i < n; ++i) {
        __syncwarp();
        node_rows<NB>(ws, y, i, sl, sa, sud, sua, sf);
        __syncwarp();
        const double* yi = y + (size_t)i * NB;
        const bool mass_node = i < n - 1;
#pragma unroll
        for (int r = 0; r < NB; ++r) A[r] = 0.0;

        if (i == 0) {
            // ---- wall node: [A_D0 | A_U0 | r0] then [A_D0 | A_E0 | 0] ----
            for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
                for (int r = 0; r < NB; ++r) A[r] = 0.0;
                if (isD) {
#pragma unroll
                    for (int r = 0; r < NB; ++r)
                        if (r == j) A[r] = (j < S ? inv_gamma : 0.0) - sa[r];
                } else if (isU) {
                    if (pass == 0) {
#pragma unroll
                        for (int r = 0; r < NB; ++r) {
                            if (r == j) A[r] = -sud[r];
                            if (j == S && r < S) A[r] = -sua[r];
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < NB; ++r)
                            if (r == j) A[r] = -sl[r];
                    }
                } else if (isR && pass == 0) {
#pragma unroll
                    for (int r = 0; r < NB; ++r)
                        A[r] = sf[r] - (r < S ? (yi[r] + ws.psi[r]) * inv_gamma : 0.0);
                }
                ok = gauss_jordan<NB>(A, lane, pivbuf) && ok;
                if (pass == 0) {
                    if (isU) {
#pragma unroll
                        for (int r = 0; r < NB; ++r) ws.W[(size_t)r * NB + j] = A[r];
                    }
                    if (isR) {
#pragma unroll
                        for (int r = 0; r < NB; ++r) ws.zb[r] = A[r];
                    }
                    // hand W_0 columns to the A_D lanes, keep z_0 in the rhs lane
                    const int src = isD ? lane + NB : lane;
#pragma unroll
                    for (int r = 0; r < NB; ++r) Wp[r] = __shfl_sync(FULL, A[r], src);
                } else if (isU) {
                    // V_0 = A_D0^{-1} A_E0 stays in the A_U lanes for node 1
#pragma unroll
                    for (int r = 0; r < NB; ++r) { ws.V0[(size_t)r * NB + j] = A[r]; Wp[r] = A[r]; }
                }
                __syncwarp();
            }
            continue;
        }

        // ---- assemble column j of [A_D | A_U | rhs] ----
        if (i == n - 1) {
            if (isD) {
#pragma unroll
                for (int r = 0; r < NB; ++r) if (r == j) A[r] = 1.0;
            } else if (isR) {
#pragma unroll
                for (int r = 0; r < NB; ++r) A[r] = sf[r];
            }
        } else {
            if (isD) {
                if (j < S) {
                    const NodeCoef k = interior_coef(ws.cs, i);
                    const double Dj = ws.sp->D[j];
                    // reaction Jacobian column: -dR_k/dc_j
                    for (int t = tb.tbeg[j]; t < tb.tbeg[j + 1]; ++t) {
                        double v = tb.tcoef[t];
                        if (tb.ti1[t] >= 0) v *= yi[tb.ti1[t]];
                        if (tb.ti2[t] >= 0) v *= yi[tb.ti2[t]];
                        if (tb.ti3[t] >= 0) v *= yi[tb.ti3[t]];
                        const double* nur = tb.nu[tb.tr[t]];
#pragma unroll
                        for (int r = 0; r < S; ++r) A[r] = fma(-nur[r], v, A[r]);
                    }
#pragma unroll
                    for (int r = 0; r < NB; ++r) {
                        if (r == j) A[r] += inv_gamma + Dj * (k.am + k.ap);
                    }
                    if (tb.use_migration) A[S] = -(ws.sp->q[j] / ws.cs.eps) * k.hi;
                } else {
                    A[S] = 1.0;
                }
            } else if (isU) {
#pragma unroll
                for (int r = 0; r < NB; ++r) {
                    if (r == j) A[r] = -sud[r];
                    if (j == S && r < S) A[r] = -sua[r];
                }
            } else if (isR) {
#pragma unroll
                for (int r = 0; r < NB; ++r)
                    A[r] = sf[r] - (r < S ? (yi[r] + ws.psi[(size_t)i * NB + r]) * inv_gamma : 0.0);
            }
            // ---- Schur update with the previous node: A -= A_L * Wp,  A_L = -dF/dy_{i-1} ----
            if (isD || isR || (isU && i == 1)) {
                const double wg = Wp[S];
#pragma unroll
                for (int r = 0; r < S; ++r) A[r] += sl[r] * Wp[r] + sa[r] * wg;
            }
        }
        (void)mass_node;
        ok = gauss_jordan<NB>(A, lane, pivbuf) && ok;
        if (isU) {
            double* Wi = ws.W + (size_t)i * NB * NB;
#pragma unroll
            for (int r = 0; r < NB; ++r) Wi[(size_t)r * NB + j] = A[r];
        }
        if (isR) {
#pragma unroll
            for (int r = 0; r < NB; ++r) ws.zb[(size_t)i * NB + r] = A[r];
        }
        const int src = isD ? lane + NB : lane;
#pragma unroll
        for (int r = 0; r < NB; ++r) Wp[r] = __shfl_sync(FULL, A[r], src);
    }
    __syncwarp();
    return __all_sync(FULL, ok);
}

// ---------------------------------------------------------------------------
// Backward substitution; y += delta, zb <- delta.  Returns the weighted max
// norms  |delta|*w  and  |y-zn0|*w  over the error-controlled unknowns
// (concentrations of nodes 0..n-2).  wmode 0: weights from ws.ewt;
// wmode 1 (steady polish): w = 1/(prtol*|y|+patol).
template <int NB>
__device__ void backward_sweep(WarpState<NB>& ws, double& dnorm, double& anorm, int wmode,
                               double prtol, double patol) {
    constexpr int S = NB - 1;
    const int lane = ws.lane;
    const int n = ws.cs.n;
    const bool act = lane < NB;
    const int r = act ? lane : 0;
    double dmax = 0.0, amax = 0.0;
    double Wrow[NB], Wnext[NB];
    // prefetch W row of node n-2
    if (n >= 2) {
        const double* Wi = ws.W + (size_t)(n - 2) * NB * NB + (size_t)r * NB;
#pragma unroll
        for (int c = 0; c < NB; ++c) Wnext[c] = act ? Wi[c] : 0.0;
    }
    const double* zn0 = ws.zn;
    for (int i = n - 1; i >= 0; --i) {
        double d = 0.0;
        if (i == n - 1) {
            if (act) d = ws.zb[(size_t)i * NB + r];
        } else {
#pragma unroll
            for (int c = 0; c < NB; ++c) Wrow[c] = Wnext[c];
            if (i >= 1) {
                const double* Wi = ws.W + (size_t)(i - 1) * NB * NB + (size_t)r * NB;
#pragma unroll
                for (int c = 0; c < NB; ++c) Wnext[c] = act ? Wi[c] : 0.0;
            }
            if (act) {
                const double* dn = ws.zb + (size_t)(i + 1) * NB;
                double s0 = ws.zb[(size_t)i * NB + r], s1 = 0.0;
#pragma unroll
                for (int c = 0; c < NB; c += 2) {
                    s0 = fma(-Wrow[c], dn[c], s0);
                    if (c + 1 < NB) s1 = fma(-Wrow[c + 1], dn[c + 1], s1);
                }
                d = s0 + s1;
                if (i == 0) {
                    const double* d2 = ws.zb + 2 * NB;
                    const double* Vr = ws.V0 + (size_t)r * NB;
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
            const double yn = ws.y[idx] + d;
            ws.y[idx] = yn;
            if (r < S && i < n - 1) {
                double w;
                if (wmode == 0) w = ws.ewt[idx];
                else w = 1.0 / (prtol * fabs(yn) + patol);
                double ad = fabs(d) * w;
                if (!(ad <= 1e300)) ad = INFINITY;      // NaN/Inf must not be lost in fmax
                dmax = fmax(dmax, ad);
                if (wmode == 0) amax = fmax(amax, fabs(yn - zn0[idx]) * w);
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dmax = fmax(dmax, __shfl_xor_sync(FULL, dmax, o));
        amax = fmax(amax, __shfl_xor_sync(FULL, amax, o));
    }
    // propagate NaN as failure
    dnorm = dmax; anorm = amax;
}

}  // namespace catint
