// pnp_kernels.cuh -- CUDA kernels (sm_100a) behind the C ABI of include/catint_pnp.h
//
//   K1  pnp_rhs_kernel       dc/dt of the reference ODE (streaming stencil, HBM bound)
//   K2  pnp_jacobian_kernel  residual + block-tridiagonal Jacobian blocks (parity/debug)
//   K3  pnp_bdf_kernel       per-cell variable-order BDF / modified-Newton integrator on the twisted
//                            block factorisation of pnp_solver.cuh, one warp per cell
#pragma once
#include <type_traits>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "../../include/catint_pnp.h"
#include "pnp_solver.cuh"

namespace catint {

constexpr int QMAX = 5;
constexpr int LMAX = QMAX + 1;     // Nordsieck vectors zn[0..QMAX]

// The step/order control below (set_bdf, increase_coef, decrease_coef, the step-attempt loop of
// pnp_bdf_kernel and these constants) restates the published variable-coefficient Nordsieck BDF scheme
// of VODE / SUNDIALS CVODE -- cvSetBDF, cvIncreaseBDF, cvDecreaseBDF, cvNls/cvDoErrorTest/cvPrepareNextStep
// and CVODE's constant names -- from P. N. Brown, G. D. Byrne, A. C. Hindmarsh, "VODE: a variable-coefficient
// ODE solver", SIAM J. Sci. Stat. Comput. 10 (1989) 1038-1051, and A. C. Hindmarsh et al., "SUNDIALS: suite of
// nonlinear and differential/algebraic equation solvers", ACM TOMS 31 (2005) 363-396 (the ODEPACK family
// that scipy.integrate.odeint = LSODA, the reference's integrator at catint/calculator_old.py:947, belongs
// to).  It is a third-party algorithm restated from its publications, not code taken from SUNDIALS;
// oracle/bdf_local.py is the CPU prototype of the same control flow.
// integrator constants (VODE/CVODE family)
constexpr double ADDON = 1e-6, BIAS1 = 6.0, BIAS2 = 6.0, BIAS3 = 10.0;
constexpr double ETAMX1 = 1e4, ETAMX2 = 10.0, ETAMXF = 0.2, ETAMIN = 0.1, ETACF = 0.25, THRESH = 1.5;
constexpr int MXNCF = 10, MXNEF = 7, MXNEF1 = 3, SMALL_NEF = 2, LONG_WAIT = 10;
#ifndef CATINT_HIST_SPECIALIZE
#define CATINT_HIST_SPECIALIZE 1       // order-specialised instances of the history / correction passes
#endif
#ifndef CATINT_MSBP
#define CATINT_MSBP 6
#endif
#ifndef CATINT_NLSCOEF
#define CATINT_NLSCOEF 0.1
#endif
#ifndef CATINT_DGMAX
#define CATINT_DGMAX 0.3
#endif
#ifndef CATINT_MAXCOR
#define CATINT_MAXCOR 3
#endif
constexpr int MAXCOR = CATINT_MAXCOR, MSBP = CATINT_MSBP;       // VODE default 20; measured on C2: 8 -> 18 % fewer steps; 6 is best once the 7x7 factorisation is cheap
constexpr double CRDOWN = 0.3, RDIV = 2.0, NLSCOEF = CATINT_NLSCOEF, DGMAX = CATINT_DGMAX;

struct SolveParams {
    DevTables tb;
    // per-cell inputs
    const double* par; const int* nx; const int* mesh_id; const double* mesh_xi;
    const double* y0;          // optional [B][nx_max][S]
    const int* order;          // optional [B]: cell handled by launch slot k (expensive cells first), or nullptr
    long long n_cells;
    // control
    int mode, max_steps, n_out, polish_max_iter;
    double rtol, atol, h0, polish_rtol;
    const double* t_out;       // device [n_out]
    // outputs
    double* c_out; double* phi_out; double* g_out; double* flux_out;
    int* status; int* n_steps; int* n_newton; int* n_setups;
    // workspace (global), per cell: zn[LMAX][N], ewt[N], inv[nx][NB][NB], W[nx][NB][NB], V0[NB][NB],
    // la[nx][2NB] (+ y,psi,zb if they do not fit in shared memory)
    double* ws; long long ws_stride;   // doubles per cell
    int state_in_smem;
    long long* prof;           // optional [B][8] cycle counters per phase (debug hook), or nullptr
};

// g from the concentrations stored in y (default Poisson BCs): g_{n-1}=g_bulk,
// g_i = g_{i+1} + h_i*sum_k q_k c_{k,i}/eps, g_0 by linear extrapolation
// (calculator_old.py:753-759,793-796).  Sequential, one lane; y is [n][NB].
template <int NB, bool ST>
__device__ void consistent_field(const WarpState<NB, ST>& ws, double* y) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    const int n = ws.cs.n;
    if (ws.lane == 0) {
        if (!ws.tb->use_migration) {
            for (int i = 0; i < n; ++i) {
                y[(size_t)i * NB + S] = 0.0;
                if (ST) y[(size_t)i * NB + S + 1] = 0.0;
            }
        } else {
            // T_i = sum_{j=i}^{n-2} lapl_j*h_j ; g_i = g_{n-1} - T_i
            double T = 0.0;
            y[(size_t)(n - 1) * NB + S] = 0.0;
            for (int i = n - 2; i >= 1; --i) {
                const NodeCoef k = interior_coef(ws.cs, i);
                double lapl = 0.0;
                for (int s = 0; s < S; ++s) lapl -= ws.sp->q[s] * y[(size_t)i * NB + s] / ws.cs.eps;
                T += lapl * k.hi;
                y[(size_t)i * NB + S] = -T;
            }
            const WallCoef w = wall_coef(ws.cs);
            double gb = ws.cs.g_bulk;
            if (ST) {
                // Robin wall + phi(L)=0 close the affine dependence on g_{n-1}:
                //   g_0 = gb - A,  phi_0 = -gb*L + B,  (eps/Cs)*g_0 + (phiM-phiPZC) - phi_0 = 0
                const double T1 = -y[NB + S], T2 = -y[2 * NB + S];
                const double A = T1 + (T1 - T2) * w.ext;
                double B = 0.0, L = 0.0;
                for (int i = 1; i <= n - 1; ++i) {
                    const NodeCoef k = interior_coef(ws.cs, i < n - 1 ? i : n - 2);
                    const double him = i < n - 1 ? k.him : k.hi;
                    B += (-y[(size_t)i * NB + S]) * him;
                    L += him;
                }
                const double ec = ws.cs.eps / ws.cs.cstern;
                gb = (ec * A - ws.cs.phi_wall + B) / (ec + L);
            }
            for (int i = 1; i <= n - 1; ++i) y[(size_t)i * NB + S] += gb;
            const double g1 = y[NB + S], g2 = y[2 * NB + S];
            y[S] = g1 + (g1 - g2) * w.ext;
            if (ST) {
                y[(size_t)(n - 1) * NB + S + 1] = 0.0;
                for (int i = n - 1; i >= 1; --i) {
                    const NodeCoef k = interior_coef(ws.cs, i < n - 1 ? i : n - 2);
                    const double him = i < n - 1 ? k.him : k.hi;
                    y[(size_t)(i - 1) * NB + S + 1] = y[(size_t)i * NB + S + 1] - y[(size_t)i * NB + S] * him;
                }
            }
        }
    }
    __syncwarp();
}

// ===========================================================================
// K3: the integrator
// ===========================================================================
template <int NB, bool ST>
struct Bdf {
    int q, qwait, nst;
    double h, hscale, t, etamax, saved_tq5;
    double tau[QMAX + 2], l[QMAX + 2], tq[6];
};

template <int NB, bool ST>
__device__ void set_bdf(Bdf<NB, ST>& B) {
    // every loop has static bounds and is fully unrolled (order tests as predicates): the coefficient arrays stay in
    // registers instead of local memory (the dynamically indexed version showed up with 2.7 % of the stall samples)
    const int q = B.q;
    const double h = B.h;
    double l[QMAX + 2], tau[QMAX + 2];
#pragma unroll
    for (int i = 0; i < QMAX + 2; ++i) { l[i] = 0.0; tau[i] = B.tau[i]; }
    l[0] = l[1] = 1.0;
    double xi_inv = 1.0, xistar_inv = 1.0, alpha0 = -1.0, alpha0_hat = -1.0, hsum = h;
    if (q > 1) {
#pragma unroll
        for (int j = 2; j < QMAX; ++j) {
            if (j < q) {
                hsum += tau[j - 1];
                xi_inv = h / hsum;
                alpha0 -= 1.0 / j;
#pragma unroll
                for (int i = QMAX; i >= 1; --i)
                    if (i <= j) l[i] += l[i - 1] * xi_inv;
            }
        }
        alpha0 -= 1.0 / q;
        xistar_inv = -l[1] - alpha0;
        double tq1 = 0.0;
#pragma unroll
        for (int j = 1; j <= QMAX; ++j) if (j == q - 1) tq1 = tau[j];
        hsum += tq1;
        xi_inv = h / hsum;
        alpha0_hat = -l[1] - xi_inv;
#pragma unroll
        for (int i = QMAX; i >= 1; --i)
            if (i <= q) l[i] += l[i - 1] * xistar_inv;
    }
    double lq = 0.0, tauq = 0.0;
#pragma unroll
    for (int j = 0; j < QMAX + 2; ++j) { if (j == q) { lq = l[j]; tauq = tau[j]; } }
    const double A1 = 1.0 - alpha0_hat + alpha0;
    const double A2 = 1.0 + q * A1;
    double* tq = B.tq;
    tq[2] = fabs(A1 / (alpha0 * A2));
    tq[5] = fabs(A2 * xistar_inv / (lq * xi_inv));
    tq[1] = 1.0; tq[3] = 1.0;
    if (B.qwait == 1) {
        if (q > 1) {
            const double C = xistar_inv / lq;
            const double A3 = alpha0 + 1.0 / q;
            const double A4 = alpha0_hat + xi_inv;
            tq[1] = fabs(C * (1.0 - A4 + A3) / A3);
        }
        hsum += tauq;
        xi_inv = h / hsum;
        const double A5 = alpha0 - 1.0 / (q + 1);
        const double A6 = alpha0_hat - xi_inv;
        tq[3] = fabs(((1.0 - A6 + A5) / A2) / (xi_inv * (q + 2) * A5));
    }
    tq[4] = NLSCOEF / tq[2];
#pragma unroll
    for (int i = 0; i < QMAX + 2; ++i) B.l[i] = l[i];
}

// coefficients for an order increase / decrease (applied inside the begin-step pass)
template <int NB, bool ST>
__device__ void increase_coef(const Bdf<NB, ST>& B, double* l, double& A1) {
    const int q = B.q;
    for (int i = 0; i < QMAX + 2; ++i) l[i] = 0.0;
    double alpha1 = 1.0, prod = 1.0, xiold = 1.0, alpha0 = -1.0, hsum = B.hscale;
    l[2] = 1.0;
    if (q > 1) {
        for (int j = 1; j < q; ++j) {
            hsum += B.tau[j + 1];
            const double xi = hsum / B.hscale;
            prod *= xi;
            alpha0 -= 1.0 / (j + 1);
            alpha1 += 1.0 / xi;
            for (int i = j + 2; i >= 2; --i) l[i] = l[i] * xiold + l[i - 1];
            xiold = xi;
        }
    }
    A1 = (-alpha0 - alpha1) / prod;
}

template <int NB, bool ST>
__device__ void decrease_coef(const Bdf<NB, ST>& B, double* l) {
    const int q = B.q;
    for (int i = 0; i < QMAX + 2; ++i) l[i] = 0.0;
    l[2] = 1.0;
    double hsum = 0.0;
    for (int j = 1; j <= q - 2; ++j) {
        hsum += B.tau[j];
        const double xi = hsum / B.hscale;
        for (int i = j + 2; i >= 2; --i) l[i] = l[i] * xi + l[i - 1];
    }
}

// The history pass that opens every step attempt.  In one sweep over the N unknowns:
//   (undo)   inverse Pascal of a rejected attempt
//   (order)  raise (+1, needs acor in zb) or lower (-1) the order
//   (scale)  zn[j] *= eta^j
//   (predict) Pascal triangle
//   y = zn[0], psi = rl1*zn[1] - zn[0]
// q_old is the order the array currently has, q_new the order after the change.
// The Nordsieck array lives in global memory (L2).  The element-wise passes over it are latency
// bound when every lane waits for its own loads, so each lane prefetches its elements of the next
// chunks (32 consecutive unknowns per warp) into the shared-memory ring of the solve sweeps with
// cp.async, STREAM_DEPTH chunks ahead; a lane only ever reads back what it copied itself, hence no
// barrier.  ring_doubles<NB,ST>() is the ring's size; small blocks (ring too short) load directly.
// (A ring enlarged to the full 8-chunk look-ahead of the history passes for small blocks -- 1792 instead of 896
// doubles at NB = 7 -- was measured neutral: history 10.7 -> 10.1 k, correction 14.2 -> 17.1 k cycles per step.)
template <int NB, bool ST>
__host__ __device__ constexpr int ring_doubles() {
    return RING * SweepRing<NB, ST>::SLOT;
}
template <int NB, bool ST, int NARR>
__host__ __device__ constexpr int stream_depth() {
    return ring_doubles<NB, ST>() / (NARR * 32) > 8 ? 8 : ring_doubles<NB, ST>() / (NARR * 32);
}

// QF > 0: instance specialised for the common case "order QF kept, nothing to undo, predict": all
// order tests fold at compile time and the loop body shrinks to the loads, the rescaling, the
// Pascal additions and the stores of exactly QF+1 arrays.  QF = 0: general run-time version.
template <int NB, bool ST, int QF>
__device__ void history_pass(WarpState<NB, ST>& ws, int q_old_rt, int dq_rt, bool undo_rt, double eta,
                             const double* lc, double A1, double rl1, bool predict_rt) {
    const int N = ws.N;
    const int q_old = QF ? QF : q_old_rt;
    const int dq = QF ? 0 : dq_rt;
    const bool undo = QF ? false : undo_rt;
    const bool predict = QF ? true : predict_rt;
    const int q_new = q_old + dq;
    constexpr int SD = stream_depth<NB, ST, LMAX>();
    constexpr bool STREAM = SD >= 3;
    const int nch = (N + 32 - 1) / 32;
    const unsigned sb = (unsigned)__cvta_generic_to_shared(ws.ring) + 8u * ws.lane;
    const double* sring = ws.ring + ws.lane;
    int islot = 0, ichunk = 0;
    auto issue = [&]() {
        const int idx = ws.lane + ichunk * 32;
        if (ichunk < nch && idx < N) {
            const unsigned d = sb + (unsigned)(islot * LMAX * 256);
#pragma unroll
            for (int j = 0; j < LMAX; ++j)
                if (j <= q_old) {
                    if (j == 0) cp_async8(d, ws.zn + idx);
                    else cp_async8_hint(d + 256u * j, ws.zn + (size_t)j * N + idx, ws.stream);
                }
        }
        cp_commit();
        ++ichunk;
        islot = (islot + 1 == SD) ? 0 : islot + 1;
    };
    if constexpr (STREAM) {
#pragma unroll 1
        for (int p = 0; p < SD - 1; ++p) issue();
    }
    int rslot = 0;
    for (int c = 0; c < nch; ++c) {
        const int idx = ws.lane + c * 32;
        if constexpr (STREAM) cp_wait<(SD >= 3 ? SD - 2 : 0)>();
        if (idx >= N) { if constexpr (STREAM) { issue(); rslot = (rslot + 1 == SD) ? 0 : rslot + 1; } continue; }
        double z[LMAX];
        if constexpr (STREAM) {
            const double* s = sring + rslot * LMAX * 32;
#pragma unroll
            for (int j = 0; j < LMAX; ++j) z[j] = (j <= q_old) ? s[j * 32] : 0.0;
            rslot = (rslot + 1 == SD) ? 0 : rslot + 1;
        } else {
#pragma unroll
            for (int j = 0; j < LMAX; ++j) z[j] = (j <= q_old) ? ws.zn[(size_t)j * N + idx] : 0.0;
        }
        if (undo) {
#pragma unroll
            for (int k = 1; k <= QMAX; ++k)
#pragma unroll
                for (int j = QMAX; j >= k; --j)
                    if (k <= q_old && j <= q_old) z[j - 1] -= z[j];
        }
        if (dq > 0) {
            const double zl = A1 * ws.zb[idx];
#pragma unroll
            for (int j = 2; j <= QMAX; ++j) {
                if (j <= q_old) z[j] += lc[j] * zl;
                if (j == q_new) z[j] = zl;
            }
            if (q_new == 1) z[1] = zl;   // cannot happen (q_old >= 1), kept for clarity
        } else if (dq < 0) {
            double zq = 0.0;
#pragma unroll
            for (int j = 0; j < LMAX; ++j) if (j == q_old) zq = z[j];
#pragma unroll
            for (int j = 2; j < QMAX; ++j)
                if (j < q_old) z[j] -= lc[j] * zq;
        }
        if (eta != 1.0) {
            double f = eta;
#pragma unroll
            for (int j = 1; j <= QMAX; ++j) {
                if (j <= q_new) z[j] *= f;
                f *= eta;
            }
        }
        if (predict) {
#pragma unroll
            for (int k = 1; k <= QMAX; ++k)
#pragma unroll
                for (int j = QMAX; j >= k; --j)
                    if (k <= q_new && j <= q_new) z[j - 1] += z[j];
        }
#pragma unroll
        for (int j = 0; j < LMAX; ++j)
            if (j <= q_new) {
                if (j == 0) ws.zn[idx] = z[0];
                else st_hint(ws.zn + (size_t)j * N + idx, z[j], ws.stream);
            }
        if (predict) {
            ws.y[idx] = z[0];
            ws.psi[idx] = rl1 * z[1] - z[0];
        }
        if constexpr (STREAM) issue();
    }
    if constexpr (STREAM) cp_wait<0>();
    __syncwarp();
}

// SMEM = true: the Newton iterate, psi and the rhs/update vector live in shared memory and are
// addressed with LDS/STS (a generic pointer would queue these critical-path accesses behind the
// global prefetch loads in the L1TEX pipeline); SMEM = false: large grids, state in the workspace.
// One warp per cell, four cells per block.  (A variant with two warps per cell, one per half of the
// twisted elimination, was measured slower in round 1 -- register cap and issue-slot contention --
// and has been removed; both halves now run in the two half warps of the one warp.)

// One linear solve with the stored factors + update of the iterate:  zb holds the rhs on entry and the
// (unscaled) update on exit, y += scale*update; returns the weighted max norms of the scaled update
// and of the accumulated correction (identical on both warps of a pair).
template <int NB, bool ST, bool GS>
__device__ void newton_solve(WarpState<NB, ST>& ws, double scale, int mid,
                             double& del, double& acn, int wmode, double prtol, double patol,
                             long long* pc, bool prof_on) {
    const int n = ws.cs.n;
    double dmax = 0.0, amax = 0.0;
    // twisted factors: both chains advance in the two halves of the warp
    long long t0 = prof_on ? clock64() : 0;
    forward_solve<NB, ST, GS>(ws, mid);
    solve_middle<NB, ST>(ws, mid);
    if (prof_on) pc[2] += clock64() - t0;
    t0 = prof_on ? clock64() : 0;
    apply_node<NB, ST>(ws, scale, mid, dmax, amax, wmode, prtol, patol);
    backward_solve<NB, ST, GS>(ws, scale, mid, dmax, amax, wmode, prtol, patol);
    del = warp_max(dmax);
    acn = warp_max(amax);
    if (prof_on) pc[3] += clock64() - t0;
}

// Cells (warps) per block.  Measured on the 1024-cell C2 launch: 1 cell per block (cells spread 6..7 per SM instead
// of 8 on 108 SMs and 4 on 40, expensive cells first via CatintPnpCells.order) is not faster than 4 (159.8 vs
// 160.4 ms) -- the launch lasts as long as its slowest cell either way -- and costs a table copy per cell.
#ifndef CATINT_CELLS_PER_BLOCK
#define CATINT_CELLS_PER_BLOCK 4
#endif
constexpr int BDF_WARPS = CATINT_CELLS_PER_BLOCK;

template <int NB, bool ST, bool SMEM>
__global__ void __launch_bounds__(32 * BDF_WARPS, 8 / BDF_WARPS) pnp_bdf_kernel(SolveParams P) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    constexpr int WARPS = BDF_WARPS;                       // cells per block
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = warp;                                 // cell slot inside the block
    const long long launch_slot = (long long)blockIdx.x * WARPS + slot;
    // launch order: expensive cells first (optional permutation from the host)
    const long long cell = (P.order && launch_slot < P.n_cells) ? P.order[launch_slot] : launch_slot;

    // shared layout: tables | per warp { CellSpecies | scratch | [y psi zb] }
    DevTables* tb = reinterpret_cast<DevTables*>(smem_raw);
    const size_t tbytes = (tables_prefix_bytes(P.tb) + 15) & ~size_t(15);
    {
        const int words = (int)(tbytes / 4);
        const int* src = reinterpret_cast<const int*>(&P.tb);
        int* dst = reinterpret_cast<int*>(tb);
        for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
    }
    __syncthreads();
    size_t off = tbytes;
    const int nxm = P.tb.nx_max;
    const int SF = P.tb.S_full;                            // species of the caller's model (outputs, y0)
    // per cell: CellSpecies | scratch | ring | [y zb]
    constexpr size_t RINGD = (size_t)ring_doubles<NB, ST>();
    constexpr size_t WARPD = scratch_doubles<NB, ST>() + RINGD;          // doubles private to the warp
    const size_t cell_doubles = WARPD + (SMEM ? (size_t)2 * nxm * NB : 0);
    const size_t per_cell = ((sizeof(CellSpecies) + 15) & ~size_t(15)) + cell_doubles * sizeof(double);
    unsigned char* mine = smem_raw + off + (size_t)slot * per_cell;
    if (launch_slot >= P.n_cells) return;

    const long long t_kernel0 = clock64();
    WarpState<NB, ST> ws;
    ws.lane = lane;
    ws.tb = tb;
    ws.keep = l2_policy_keep();
    ws.stream = l2_policy_stream();
    CellSpecies* sp = reinterpret_cast<CellSpecies*>(mine);
    ws.sp = sp;
    double* cellbase = reinterpret_cast<double*>(mine + ((sizeof(CellSpecies) + 15) & ~size_t(15)));
    ws.scratch = cellbase;
    load_cell(*tb, P.par, P.nx, P.mesh_id, P.mesh_xi, cell, lane, ws.cs, sp);
    const int n = ws.cs.n;
    if (n < CATINT_PNP_MIN_NODES || n > nxm) {           // bad nx[cell]: report, touch nothing else
        if (lane == 0) {
            P.status[cell] = CATINT_PNP_CELL_BAD_INPUT;
            P.n_steps[cell] = 0; P.n_newton[cell] = 0;
            if (P.n_setups) P.n_setups[cell] = 0;
        }
        return;
    }
    const int N = n * NB;
    ws.N = N;
    double* g = P.ws + (size_t)cell * P.ws_stride;
    constexpr int NBP = padded<NB, ST>();
    ws.zn = g;                 g += align4((size_t)LMAX * nxm * NB);
    ws.ewt = g;                g += align4((size_t)nxm * NB);
    ws.fac = g;                g += align4((size_t)nxm * fac_rec<NB, ST>());
    ws.V0 = g;                 g += align4((size_t)NB * NBP);
    ws.W1 = g;                 g += align4((size_t)NB * NBP);
    ws.dJ = g;                 g += align4((size_t)NB * NBP);
    ws.psi = g;                g += align4((size_t)nxm * NB);
    ws.ring = ws.scratch + scratch_doubles<NB, ST>();
    double* sdyn = cellbase + WARPD;
    if constexpr (SMEM) {
        ws.y = sdyn; ws.zb = sdyn + (size_t)nxm * NB;
        sdyn += (size_t)2 * nxm * NB;
    } else {
        ws.y = g; ws.zb = g + align4((size_t)nxm * NB);
    }
    const int mid = n / 2;                                // coupling node of the twisted sweeps
    SmemOffsets so;                                       // shared-memory pointers as offsets (factor_nodes)
    so.scratch = (unsigned)(reinterpret_cast<unsigned char*>(ws.scratch) - smem_raw);
    so.sp = (unsigned)(reinterpret_cast<const unsigned char*>(sp) - smem_raw);
    so.y = SMEM ? (unsigned)(reinterpret_cast<unsigned char*>(ws.y) - smem_raw) : 0u;
    so.ring = (unsigned)(reinterpret_cast<unsigned char*>(ws.ring) - smem_raw);
    constexpr int vstride = 32;
    const int vlane = lane;

    // ---- initial state: y0 (or bulk) with the consistent field --------------
    for (int idx = vlane; idx < N; idx += vstride) {
        const int i = idx / NB, r = idx - i * NB;
        double v = 0.0;
        if (r < S) v = P.y0 ? P.y0[((size_t)cell * nxm + i) * SF + tb->cmap[r]] : sp->cb[r];
        ws.y[idx] = v;
        ws.psi[idx] = 0.0;
    }
    __syncwarp();
    consistent_field<NB, ST>(ws, ws.y);
    refresh_wall_flux<NB, ST>(ws, ws.y);

    const double rtol = P.rtol, atol = P.atol;

    // zn[0]=y, zn[1]=h*f(y) (mass rows), ewt; first step size from the initial rate of change
    double fnorm = 0.0;
    for (int idx = vlane; idx < N; idx += vstride) {
        const int i = idx / NB, r = idx - i * NB;
        const bool mass = r < S && i < n - 1;
        const double F = mass ? row_residual<NB, ST>(ws, ws.y, i, r) : 0.0;
        const double yv = ws.y[idx];
        const double w = 1.0 / (rtol * fabs(yv) + atol);
        ws.ewt[idx] = w;
        ws.zn[idx] = yv;
        ws.zn[(size_t)N + idx] = F;                       // scaled by h below
        if (mass) fnorm = fmax(fnorm, fabs(F) * w);
    }
    __syncwarp();
    fnorm = warp_max(fnorm);

    Bdf<NB, ST> B;
    B.q = 1; B.qwait = 2; B.nst = 0; B.t = 0.0; B.etamax = ETAMX1; B.saved_tq5 = 0.0;
    for (int i = 0; i < QMAX + 2; ++i) { B.tau[i] = 0.0; B.l[i] = 0.0; }
    const double t_end = P.t_out[P.n_out - 1];
    double h0 = P.h0;
    if (!(h0 > 0.0)) {
        h0 = fnorm > 0.0 ? 1.0 / fnorm : 1e-6;
        h0 = fmin(h0, 1e-3 * (t_end > 0.0 ? t_end : 1.0));
    }
    B.h = B.hscale = h0;
    for (int idx = vlane; idx < N; idx += vstride) ws.zn[(size_t)N + idx] *= h0;
    __syncwarp();

    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};     // factor, residual, forward, backward, norms, history, correction, other
    const bool prof_on = P.prof != nullptr;
#define CATINT_TIC long long tic_ = prof_on ? clock64() : 0
#define CATINT_TOC(k_) if (prof_on) pc[k_] += clock64() - tic_
    int status = CATINT_PNP_CELL_CONVERGED;
    int nni = 0, iout = 0, nsetups = 0, nstlp = 0;
    bool have_factors = false, force_setup = false;
    double inv_gamma_p = 1.0;
    // pending transformation of the history array for the next attempt
    int pend_dq = 0; bool pend_undo = false; double pend_eta = 1.0;
    double lc[QMAX + 2]; double A1c = 0.0;
    for (int i = 0; i < QMAX + 2; ++i) lc[i] = 0.0;

    while (iout < P.n_out && status == CATINT_PNP_CELL_CONVERGED) {
        if (B.nst >= P.max_steps) { status = CATINT_PNP_CELL_MAX_STEPS; break; }
        int ncf = 0, nef = 0;
        const double saved_t = B.t;
        double dsm = 0.0;
        bool accepted = false;
        // ------------------------------------------------ attempts of one step
        while (true) {
            // order change (only on the first attempt) + rescale + predict, fused
            const int q_old = B.q;
            if (pend_dq > 0) increase_coef<NB, ST>(B, lc, A1c);
            else if (pend_dq < 0) decrease_coef<NB, ST>(B, lc);
            if (pend_dq != 0) { B.q += pend_dq; B.qwait = B.q + 1; }
            if (pend_eta != 1.0) { B.h = B.hscale * pend_eta; B.hscale = B.h; }
            if (saved_t + B.h == saved_t) { status = CATINT_PNP_CELL_STEP_UNDERFLOW; break; }   // t+h == t
            B.t = saved_t + B.h;
            set_bdf<NB, ST>(B);
            const double rl1 = 1.0 / B.l[1];
            { CATINT_TIC;
              if (CATINT_HIST_SPECIALIZE && !pend_undo && pend_dq == 0) {
                  switch (q_old) {
                      case 1: history_pass<NB, ST, 1>(ws, 1, 0, false, pend_eta, lc, A1c, rl1, true); break;
                      case 2: history_pass<NB, ST, 2>(ws, 2, 0, false, pend_eta, lc, A1c, rl1, true); break;
                      case 3: history_pass<NB, ST, 3>(ws, 3, 0, false, pend_eta, lc, A1c, rl1, true); break;
                      case 4: history_pass<NB, ST, 4>(ws, 4, 0, false, pend_eta, lc, A1c, rl1, true); break;
                      default: history_pass<NB, ST, 5>(ws, 5, 0, false, pend_eta, lc, A1c, rl1, true); break;
                  }
              } else {
                  history_pass<NB, ST, 0>(ws, q_old, pend_dq, pend_undo, pend_eta, lc, A1c, rl1, true);
              } CATINT_TOC(5); }
            pend_dq = 0; pend_undo = false; pend_eta = 1.0;
            const double inv_gamma = B.l[1] / B.h;

            // ---- modified Newton corrector (VODE/CVODE factor re-use) ---------------
            // re-factor on the first step, after a failed attempt, every MSBP steps, or when
            // gamma drifted by more than DGMAX since the factors were computed
            bool conv = false;
            double acnrm = 0.0;
            bool call_setup = !have_factors || force_setup || B.nst >= nstlp + MSBP ||
                              fabs(inv_gamma_p / inv_gamma - 1.0) > DGMAX;
            force_setup = false;
            for (int pass = 0; pass < 2 && !conv; ++pass) {
                bool jcur = false;
                if (pass == 1) {
                    // stale factors did not converge: restart the corrector from the prediction
                    for (int idx = vlane; idx < N; idx += vstride) ws.y[idx] = ws.zn[idx];
                    __syncwarp();
                    call_setup = true;
                }
                if (call_setup) {
                    CATINT_TIC;
                    bool ok;
                    ok = factor_nodes<NB, ST, SMEM>(ws, so, inv_gamma, mid, prof_on ? P.prof + cell * 8 + 4 : nullptr);
                    CATINT_TOC(0);
                    ++nsetups;
                    have_factors = ok;
                    inv_gamma_p = inv_gamma;
                    nstlp = B.nst;
                    jcur = true;
                    if (!ok) break;
                }
                const double gamrat = inv_gamma_p / inv_gamma;          // gamma/gamma_p
                const double dscale = (gamrat != 1.0) ? 2.0 * gamrat / (1.0 + gamrat) : 1.0;
                double crate = 1.0, delp = 0.0;
                bool bad = false;
                for (int m = 0; m < MAXCOR; ++m) {
                    { CATINT_TIC; residual_pass<NB, ST>(ws, inv_gamma); CATINT_TOC(1); }
                    double del = 0.0, acn = 0.0;
                    newton_solve<NB, ST, !SMEM>(ws, dscale, mid, del, acn, 0, 0.0, 0.0, pc, prof_on);
                    ++nni;
                    if (!(del <= 1e300)) { bad = true; break; }
                    if (m > 0) crate = fmax(CRDOWN * crate, del / delp);
                    const double dcon = del * fmin(1.0, crate) / B.tq[4];
                    if (dcon <= 1.0) { conv = true; acnrm = (m == 0) ? del : acn; break; }
                    if (m + 1 == MAXCOR || (m >= 1 && del > RDIV * delp)) break;
                    delp = del;
                }
                (void)bad;
                if (jcur) break;            // fresh factors: a failure is a real failure
            }
            if (!conv) {
                ++ncf;
                force_setup = true;
                B.etamax = 1.0;
                B.t = saved_t;
                if (ncf == MXNCF || B.h * ETACF < 1e-300) { status = CATINT_PNP_CELL_CORRECTOR_FAILED; break; }
                pend_undo = true; pend_eta = ETACF;
                continue;
            }
            dsm = acnrm * B.tq[2];
            if (dsm <= 1.0) { accepted = true; break; }
            // ---- error test failed ---------------------------------------------
            ++nef;
            force_setup = true;
            B.etamax = 1.0;
            B.t = saved_t;
            if (nef == MXNEF) { status = CATINT_PNP_CELL_ERROR_TEST_FAILED; break; }
            pend_undo = true;
            if (nef <= MXNEF1) {
                double eta = 1.0 / (pow(BIAS2 * dsm, 1.0 / (B.q + 1)) + ADDON);
                eta = fmax(ETAMIN, eta);
                if (nef >= SMALL_NEF) eta = fmin(eta, ETAMXF);
                pend_eta = eta;
            } else if (B.q > 1) {
                pend_dq = -1;
                pend_eta = ETAMIN;
            } else {
                // order 1 and still failing: restart the history from the last accepted state
                history_pass<NB, ST, 0>(ws, B.q, 0, true, 1.0, lc, 0.0, 1.0, false);
                pend_undo = false;
                B.h *= ETAMIN; B.hscale = B.h;
                B.qwait = LONG_WAIT;
                for (int idx = vlane; idx < N; idx += vstride) ws.y[idx] = ws.zn[idx];
                __syncwarp();
                refresh_wall_flux<NB, ST>(ws, ws.y);
                for (int idx = vlane; idx < N; idx += vstride) {
                    const int i = idx / NB, r = idx - i * NB;
                    const bool mass = r < S && i < n - 1;
                    ws.zn[(size_t)N + idx] = mass ? B.h * row_residual<NB, ST>(ws, ws.y, i, r) : 0.0;
                }
                __syncwarp();
            }
        }
        if (!accepted) break;

        // ------------------------------------------------ complete the step
        ++B.nst;
        const int q = B.q;
        for (int i = q; i >= 2; --i) B.tau[i] = B.tau[i - 1];
        if (q == 1 && B.nst > 1) B.tau[2] = B.tau[1];
        B.tau[1] = B.h;
        --B.qwait;
        const bool save_acor = (B.qwait == 1 && q != QMAX);
        const bool want_eta = (B.etamax != 1.0) && (B.qwait == 0);
        const bool want_up = want_eta && q != QMAX && B.saved_tq5 != 0.0;
        double cquot = 0.0;
        if (want_up) cquot = (B.tq[5] / B.saved_tq5) * pow(B.h / B.tau[2], (double)(q + 1));
        double ddn = 0.0, dup = 0.0;
        const long long tic_corr = prof_on ? clock64() : 0;
        // correction pass: zn[j] += l[j]*acor, norms for the order selection, new weights.
        // Two elements per lane and iteration, loads grouped ahead of the stores (memory-level
        // parallelism; the history array streams from L2/HBM).
        auto correction = [&](auto qf_) {
            // QF > 0: instance specialised for order QF (all order tests fold); QF = 0: run-time order
            constexpr int QF = decltype(qf_)::value;
            const int q = QF ? QF : B.q;
            double lreg[LMAX];
#pragma unroll
            for (int j = 0; j < LMAX; ++j) lreg[j] = B.l[j];
            double* __restrict__ zn = ws.zn;
            double* __restrict__ ewt = ws.ewt;
            constexpr int NARR = LMAX + 1 + (SMEM ? 0 : 1);          // zn[0..QMAX], the weights (+ y when it lives in global memory)
            constexpr int SD = stream_depth<NB, ST, NARR>();
            constexpr bool STREAM = SD >= 3;
            const int nch = (N + vstride - 1) / vstride;
            const unsigned sb = (unsigned)__cvta_generic_to_shared(ws.ring) + 8u * lane;
            const double* sring = ws.ring + lane;
            int islot = 0, ichunk = 0;
            auto issue = [&]() {
                const int ii = vlane + ichunk * vstride;
                if (ichunk < nch && ii < N) {
                    const unsigned d = sb + (unsigned)(islot * NARR * 256);
#pragma unroll
                    for (int j = 0; j < LMAX; ++j)
                        if (j <= q || (j == QMAX && want_up)) {
                            if (j == 0) cp_async8(d, zn + ii);
                            else cp_async8_hint(d + 256u * j, zn + (size_t)j * N + ii, ws.stream);
                        }
                    cp_async8(d + 256u * LMAX, ewt + ii);
                    if (!SMEM) cp_async8(d + 256u * (LMAX + 1), ws.y + ii);
                }
                cp_commit();
                ++ichunk;
                islot = (islot + 1 == SD) ? 0 : islot + 1;
            };
            if constexpr (STREAM) {
#pragma unroll 1
                for (int p_ = 0; p_ < SD - 1; ++p_) issue();
            }
            int rslot = 0;
            for (int c = 0; c < nch; ++c) {
                const int ii = vlane + c * vstride;
                if constexpr (STREAM) cp_wait<(SD >= 3 ? SD - 2 : 0)>();
                if (ii < N) {
                    double zj[LMAX], w;
                    if constexpr (STREAM) {
                        const double* s = sring + rslot * NARR * 32;
#pragma unroll
                        for (int j = 0; j < LMAX; ++j) zj[j] = (j <= q || (j == QMAX && want_up)) ? s[j * 32] : 0.0;
                        w = s[LMAX * 32];
                    } else {
#pragma unroll
                        for (int j = 0; j < LMAX; ++j)
                            zj[j] = (j <= q || (j == QMAX && want_up)) ? zn[(size_t)j * N + ii] : 0.0;
                        w = ewt[ii];
                    }
                    double yv;
                    if constexpr (SMEM || !STREAM) yv = ws.y[ii];
                    else yv = sring[rslot * NARR * 32 + (LMAX + 1) * 32];
                    const int i = ii / NB, r = ii - i * NB;
                    const bool mass = r < S && i < n - 1;
                    const double ac = yv - zj[0];
                    if (want_up && mass) dup = fmax(dup, fabs(ac - cquot * zj[QMAX]) * w);
                    zn[ii] = yv;
#pragma unroll
                    for (int j = 1; j <= QMAX; ++j) {
                        if (j <= q) {
                            const double v = zj[j] + lreg[j] * ac;
                            st_hint(zn + (size_t)j * N + ii, v, ws.stream);
                            if (j == q && want_eta && mass) ddn = fmax(ddn, fabs(v) * w);
                        }
                    }
                    if (save_acor) st_hint(zn + (size_t)QMAX * N + ii, ac, ws.stream);
                    ws.zb[ii] = ac;
                    ewt[ii] = 1.0 / (rtol * fabs(yv) + atol);
                }
                if constexpr (STREAM) { issue(); rslot = (rslot + 1 == SD) ? 0 : rslot + 1; }
            }
            if constexpr (STREAM) cp_wait<0>();
        };
        switch (CATINT_HIST_SPECIALIZE ? q : 0) {
            case 1: correction(std::integral_constant<int, 1>()); break;
            case 2: correction(std::integral_constant<int, 2>()); break;
            case 3: correction(std::integral_constant<int, 3>()); break;
            case 4: correction(std::integral_constant<int, 4>()); break;
            case 5: correction(std::integral_constant<int, 5>()); break;
            default: correction(std::integral_constant<int, 0>()); break;
        }
        __syncwarp();
        ddn = warp_max(ddn);
        dup = warp_max(dup);
        if (prof_on) pc[6] += clock64() - tic_corr;
        if (save_acor) B.saved_tq5 = B.tq[5];

        // ------------------------------------------------ next order and step size
        int qprime = q;
        double eta = 1.0;
        if (B.etamax == 1.0) {
            B.qwait = max(B.qwait, 2);
        } else {
            const double etaq = 1.0 / (pow(BIAS2 * dsm, 1.0 / (q + 1)) + ADDON);
            if (B.qwait != 0) {
                eta = etaq;
            } else {
                B.qwait = 2;
                ddn = ddn * B.tq[1];
                dup = dup * B.tq[3];
                double etaqm1 = 0.0, etaqp1 = 0.0;
                if (q > 1) etaqm1 = 1.0 / (pow(BIAS1 * ddn, 1.0 / q) + ADDON);
                if (want_up) etaqp1 = 1.0 / (pow(BIAS3 * dup, 1.0 / (q + 2)) + ADDON);
                const double etam = fmax(etaqm1, fmax(etaq, etaqp1));
                if (etam < THRESH) eta = 1.0;
                else if (etam == etaq) eta = etaq;
                else if (etam == etaqm1) { eta = etaqm1; qprime = q - 1; }
                else { eta = etaqp1; qprime = q + 1; }
            }
            if (eta < THRESH) eta = 1.0;
            else eta = fmin(eta, B.etamax);
        }
        B.etamax = ETAMX2;

        // ------------------------------------------------ dense output
        while (iout < P.n_out && B.t >= P.t_out[iout] * (1.0 - 1e-14)) {
            const double s = (P.t_out[iout] - B.t) / B.h;
            const bool last = (iout == P.n_out - 1);
            double* co = P.c_out + ((size_t)iout * P.n_cells + cell) * nxm * SF;
            double* go = P.g_out ? P.g_out + ((size_t)iout * P.n_cells + cell) * nxm : nullptr;
            for (int idx = vlane; idx < N; idx += vstride) {
                const int i = idx / NB, r = idx - i * NB;
                double v = 0.0;
#pragma unroll
                for (int j = QMAX; j >= 0; --j)
                    if (j <= q) v = v * s + ws.zn[(size_t)j * N + idx];
                if (last) ws.y[idx] = v;           // kept for the steady polish / final outputs
                if (!(last && P.mode == CATINT_PNP_MODE_STEADY)) {
                    if (r < S) co[(size_t)i * SF + tb->cmap[r]] = v;
                    else if (r == S) { if (go) go[i] = v; }
                    else if (P.phi_out) P.phi_out[((size_t)iout * P.n_cells + cell) * nxm + i] = v;
                }
            }
            __syncwarp();
            if (!ST && !last && P.phi_out && go && lane == 0) {
                // potential of an intermediate output: forward cumulative sum of the g just written
                double* po2 = P.phi_out + ((size_t)iout * P.n_cells + cell) * nxm;
                double v = ws.cs.phi_wall, vm1 = v, vm2 = v;
                po2[0] = v;
                for (int i = 1; i <= n - 2; ++i) {
                    const NodeCoef kc = interior_coef(ws.cs, i);
                    v = v + go[i] * kc.him;
                    po2[i] = v; vm2 = vm1; vm1 = v;
                }
                if (n >= 3) {
                    const double ratio = ws.cs.uniform ? 1.0 :
                        (ws.cs.xi[n - 1] - ws.cs.xi[n - 2]) / (ws.cs.xi[n - 2] - ws.cs.xi[n - 3]);
                    po2[n - 1] = vm1 + (vm1 - vm2) * ratio;
                }
            }
            __syncwarp();
            ++iout;
        }
        if (iout >= P.n_out) break;
        pend_dq = qprime - q;
        pend_eta = eta;
    }

    // ---------------------------------------------------- steady-state polish
    if (status == CATINT_PNP_CELL_CONVERGED && P.mode == CATINT_PNP_MODE_STEADY) {
        double cscale = lane < S ? fabs(sp->cb[lane]) : 0.0;
        cscale = warp_max(cscale);
        const double patol = 1e-12 * fmax(cscale, 1e-300);
        bool done = false;
        double del_prev = 0.0;
        for (int it = 0; it < P.polish_max_iter && !done; ++it) {
            // true Newton on the steady residual: inv_gamma = 0 removes the mass term
            bool ok;
            ok = factor_nodes<NB, ST, SMEM>(ws, so, 0.0, mid, nullptr);
            ++nsetups;
            if (!ok) break;
            residual_pass<NB, ST>(ws, 0.0);
            double del = 0.0, acn = 0.0;
            newton_solve<NB, ST, !SMEM>(ws, 1.0, mid, del, acn, 1, P.polish_rtol, patol, pc, false);
            ++nni;
            if (!(del <= 1e300)) break;
            // converged, or stagnating at the rounding floor of the linear solve with an update that is
            // already below 1e-8 relative (two orders under the 1e-6 parity tolerance)
            if (del <= 1.0 || (it >= 1 && del > 0.5 * del_prev && del * P.polish_rtol <= 1e-8)) done = true;
            del_prev = del;
        }
        if (!done) status = CATINT_PNP_CELL_POLISH_FAILED;
    }

    // ---------------------------------------------------- final outputs
    {
        const int io = P.n_out - 1;
        double* co = P.c_out + ((size_t)io * P.n_cells + cell) * nxm * SF;
        double* go = P.g_out ? P.g_out + ((size_t)io * P.n_cells + cell) * nxm : nullptr;
        double* po = P.phi_out ? P.phi_out + ((size_t)io * P.n_cells + cell) * nxm : nullptr;
        if (P.mode == CATINT_PNP_MODE_STEADY || status != CATINT_PNP_CELL_CONVERGED) {
            for (int idx = vlane; idx < N; idx += vstride) {
                const int i = idx / NB, r = idx - i * NB;
                const double v = ws.y[idx];
                if (r < S) co[(size_t)i * SF + tb->cmap[r]] = v;
                else if (r == S) { if (go) go[i] = v; }
                else if (po) po[i] = v;
            }
        }
        // passive species (neutral, no homogeneous reaction, not read by any flux equation; eliminated from the
        // block system in steady mode, pnp_capi.cu: reduce_passive): their discrete steady equations
        // D*(second difference) = 0 with the flux wall condition and the Dirichlet bulk node have the exactly
        // linear solution c(x_i) = c_bulk + (J/D)*(x_{n-1} - x_i) on any mesh
        for (int pp = 0; pp < tb->npas; ++pp) {
            const int k = tb->pmap[pp];
            const double* par = P.par + (size_t)cell * (3 * SF + 8);
            double Jp = par[SF + k];
            for (int e = 0; e < tb->fq.n_eq; ++e)
                Jp += tb->pcoef[pp][e] * fluxeq_eval(&tb->fq, e, ws.cs.fpar, ws.y, ST ? ws.y[S + 1] : ws.cs.phi_wall, -1, S, nullptr);
            const double slope = Jp / par[2 * SF + k];
            const double xl = ws.cs.uniform ? ws.cs.dx * (n - 1) : ws.cs.dx * ws.cs.xi[n - 1];
            for (int i = lane; i < n; i += 32) {
                const double x = ws.cs.uniform ? ws.cs.dx * i : ws.cs.dx * ws.cs.xi[i];
                co[(size_t)i * SF + k] = i == n - 1 ? par[k] : par[k] + slope * (xl - x);
            }
            if (P.flux_out && lane == 0) P.flux_out[(size_t)cell * SF + k] = Jp;
        }
        __syncwarp();
        // potential by the forward cumulative sum of the reference (calculator_old.py:798-800);
        // in Stern mode phi is an unknown of the state and has been written above
        if (!ST && po && lane == 0) {
            double v = ws.cs.phi_wall;
            po[0] = v;
            double vm1 = v, vm2 = v;
            for (int i = 1; i <= n - 2; ++i) {
                const NodeCoef k = interior_coef(ws.cs, i);
                v = v + ws.y[(size_t)i * NB + S] * k.him;
                po[i] = v;
                vm2 = vm1; vm1 = v;
            }
            if (n >= 3) {
                const double ratio = ws.cs.uniform ? 1.0 :
                    (ws.cs.xi[n - 1] - ws.cs.xi[n - 2]) / (ws.cs.xi[n - 2] - ws.cs.xi[n - 3]);
                po[n - 1] = vm1 + (vm1 - vm2) * ratio;
            }
        }
        if (P.flux_out && lane < S) {
            const WallCoef w = wall_coef(ws.cs);
            const double bq = tb->use_migration ? sp->bq[lane] : 0.0;
            P.flux_out[(size_t)cell * SF + tb->cmap[lane]] =
                -sp->D[lane] * ((ws.y[2 * NB + lane] - ws.y[lane]) * w.w0 + bq * ws.y[NB + lane] * ws.y[NB + S]);
        }
        if (lane == 0) {
            P.status[cell] = status;
            P.n_steps[cell] = B.nst;
            P.n_newton[cell] = nni;
            if (P.n_setups) P.n_setups[cell] = nsetups;
            if (prof_on) {
                // slot 4 (norms, fused into the sweeps long ago) holds the assembly part of the factorisation,
                // accumulated by factor_nodes itself
                pc[7] = clock64() - t_kernel0;
                for (int k_ = 0; k_ < 8; ++k_) if (k_ != 4) P.prof[cell * 8 + k_] = pc[k_];
            }
        }
    }
}

// ===========================================================================
// K2: residual and Jacobian blocks (one warp per cell), for parity checks
// ===========================================================================
struct JacParams {
    DevTables tb;
    const double* par; const int* nx; const int* mesh_id; const double* mesh_xi;
    const double* y; long long n_cells;
    double* F; double* Lb; double* Db; double* Ub;
};

template <int NB, bool ST>
__global__ void __launch_bounds__(128) pnp_jacobian_kernel(JacParams P) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    constexpr int WARPS = 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long cell = (long long)blockIdx.x * WARPS + warp;
    DevTables* tb = reinterpret_cast<DevTables*>(smem_raw);
    {
        const int words = (int)(sizeof(DevTables) / 4);
        const int* src = reinterpret_cast<const int*>(&P.tb);
        int* dst = reinterpret_cast<int*>(tb);
        for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
    }
    __syncthreads();
    size_t off = (sizeof(DevTables) + 15) & ~size_t(15);
    const size_t per_warp = ((sizeof(CellSpecies) + 15) & ~size_t(15)) + (size_t)scratch_doubles<NB, ST>() * sizeof(double);
    unsigned char* mine = smem_raw + off + (size_t)warp * per_warp;
    if (cell >= P.n_cells) return;
    WarpState<NB, ST> ws;
    ws.lane = lane; ws.tb = tb;
    ws.keep = 0; ws.stream = 0;
    CellSpecies* sp = reinterpret_cast<CellSpecies*>(mine);
    ws.sp = sp;
    ws.scratch = reinterpret_cast<double*>(mine + ((sizeof(CellSpecies) + 15) & ~size_t(15)));
    load_cell(*tb, P.par, P.nx, P.mesh_id, P.mesh_xi, cell, lane, ws.cs, sp);
    const int n = ws.cs.n, nxm = P.tb.nx_max;
    if (n < 4 || n > nxm) return;                        // bad nx[cell]: skip, outputs untouched
    const double* y = P.y + (size_t)cell * nxm * NB;
    refresh_wall_flux<NB, ST>(ws, y);
    double* sl = ws.scratch + 2 * (NB + 2);
    double* sa = sl + NB; double* sud = sa + NB; double* sua = sud + NB;
    for (int i = 0; i < n; ++i) {
        __syncwarp();
        node_coeffs<NB, ST>(ws, y, i, sl, sa, sud, sua);
        __syncwarp();
        const size_t nb = ((size_t)cell * nxm + i);
        if (P.F && lane < NB) P.F[nb * NB + lane] = row_residual<NB, ST>(ws, y, i, lane);
        // blocks are d(row)/d(col) of F itself (not of the Newton matrix)
        if (lane < NB) {
            const int j = lane;     // column
            const double* yi = y + (size_t)i * NB;
            double Dcol[NB], Lcol[NB], Ucol[NB];
#pragma unroll
            for (int r = 0; r < NB; ++r) { Dcol[r] = 0.0; Lcol[r] = 0.0; Ucol[r] = 0.0; }
            if (i == 0) {
#pragma unroll
                for (int r = 0; r < NB; ++r) {
                    if (r == j) { Dcol[r] = sa[r]; Ucol[r] = sud[r]; Lcol[r] = sl[r]; }   // Lcol = extra block (0,2)
                    if (j == S && r < S) Ucol[r] = sua[r];
                }
                if (tb->fq.n_eq > 0 && j < S) {
                    // flux equations: dF_r/dc_j(0) += ih0*dJ_r/dc_j(0)
                    const double ih0 = wall_coef(ws.cs).ih0;
                    for (int e = 0; e < tb->fq.n_eq; ++e) {
                        double ge = 0.0;
                        fluxeq_eval(&tb->fq, e, ws.cs.fpar, y, ws.cs.phi_wall, j, S, &ge);
#pragma unroll
                        for (int r = 0; r < S; ++r) Dcol[r] += ih0 * tb->fq.coef[r][e] * ge;
                    }
                }
            } else if (i == n - 1) {
#pragma unroll
                for (int r = 0; r < NB; ++r) if (r == j) Dcol[r] = -1.0;
            } else {
                const NodeCoef k = interior_coef(ws.cs, i);
                if (j < S) {
                    for (int t = tb->tbeg[j]; t < tb->tbeg[j + 1]; ++t) {
                        double v = tb->tcoef[t];
                        if (tb->ti1[t] >= 0) v *= yi[tb->ti1[t]];
                        if (tb->ti2[t] >= 0) v *= yi[tb->ti2[t]];
                        if (tb->ti3[t] >= 0) v *= yi[tb->ti3[t]];
                        const double* nur = tb->nu[tb->tr[t]];
#pragma unroll
                        for (int r = 0; r < S; ++r) Dcol[r] = fma(nur[r], v, Dcol[r]);
                    }
#pragma unroll
                    for (int r = 0; r < NB; ++r) if (r == j) Dcol[r] -= sp->D[j] * (k.am + k.ap);
                    if (tb->use_migration) Dcol[S] = (sp->q[j] / ws.cs.eps) * k.hi;
                } else {
                    Dcol[S] = -1.0;
                }
#pragma unroll
                for (int r = 0; r < NB; ++r) {
                    if (r == j) { Lcol[r] = sl[r]; Ucol[r] = sud[r]; }
                    if (j == S && r < S) { Lcol[r] = sa[r]; Ucol[r] = sua[r]; }
                }
            }
#pragma unroll
            for (int r = 0; r < NB; ++r) {
                if (P.Db) P.Db[(nb * NB + r) * NB + j] = Dcol[r];
                if (P.Lb) P.Lb[(nb * NB + r) * NB + j] = Lcol[r];
                if (P.Ub) P.Ub[(nb * NB + r) * NB + j] = Ucol[r];
            }
        }
    }
}

}  // namespace catint


// ===========================================================================
// launchers (explicitly instantiated per block size in pnp_inst.cu)
// ===========================================================================
namespace catint {
template <int NB, bool ST>
int launch_bdf(SolveParams& P, cudaStream_t st) {
    const size_t base = (tables_prefix_bytes(P.tb) + 15) & ~size_t(15);
    const size_t species = ((sizeof(CellSpecies) + 15) & ~size_t(15));
    const size_t warpd = (size_t)scratch_doubles<NB, ST>() + (size_t)ring_doubles<NB, ST>();
    const size_t state = (size_t)2 * P.tb.nx_max * NB;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return CATINT_PNP_ECUDA;
    int max_optin = 0;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    // one warp per cell, four cells per block; the Newton iterate and the update vector live in shared
    // memory when they fit (SMEM = true), else in the workspace.  The shared-memory opt-in is a
    // per-device function attribute: it is set on every launch (cheap, no cached state to race on).
    constexpr int W = BDF_WARPS;
    const size_t smem_single = base + W * (species + (warpd + state) * sizeof(double));
    const unsigned grid = (unsigned)((P.n_cells + W - 1) / W);
    // the state stays in shared memory as long as at least 4 cells per SM fit that way
    if (smem_single * (4 / W > 0 ? 4 / W : 1) + 4096 <= (size_t)max_optin) {
        P.state_in_smem = 1;
        cudaFuncSetAttribute(pnp_bdf_kernel<NB, ST, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_single);
        pnp_bdf_kernel<NB, ST, true><<<grid, 32 * W, smem_single, st>>>(P);
    } else {
        P.state_in_smem = 0;
        const size_t smem = base + W * (species + warpd * sizeof(double));
        cudaFuncSetAttribute(pnp_bdf_kernel<NB, ST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        pnp_bdf_kernel<NB, ST, false><<<grid, 32 * W, smem, st>>>(P);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : CATINT_PNP_ECUDA;
}

template <int NB, bool ST>
int launch_jac(JacParams& P, cudaStream_t st) {
    const int WARPS = 4;
    const size_t base = ((sizeof(DevTables) + 15) & ~size_t(15));
    const size_t per_warp = ((sizeof(CellSpecies) + 15) & ~size_t(15)) + (size_t)scratch_doubles<NB, ST>() * sizeof(double);
    const size_t smem = base + WARPS * per_warp;
    const unsigned grid = (unsigned)((P.n_cells + WARPS - 1) / WARPS);
    pnp_jacobian_kernel<NB, ST><<<grid, WARPS * 32, smem, st>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : CATINT_PNP_ECUDA;
}

}  // namespace catint
