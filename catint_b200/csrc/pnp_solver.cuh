// pnp_solver.cuh -- per-cell linear algebra of the implicit integrator (K2+K3 building blocks).
//
// One warp owns one cell.  The Newton matrix  A = Mass/gamma - dF/dy  is block
// tridiagonal (+ one wall block).  It is FACTORED by a twisted block elimination (from both
// ends towards the coupling node n/2, one end per half warp) with threshold pivoting inside the
// NB x NB blocks, and the factors are kept in global memory (L2) so that the modified-Newton
// iterations of the following steps only run the cheap SOLVE sweeps (VODE/CVODE re-use policy):
//
//   factor_nodes   phase 1, all lanes: assembly of every node's record [A_D | l,a,ud,ua] (one
//                  (node, unknown) pair per lane at a time);
//                  phase 2, sequential: records stream back through the cp.async ring; in each half
//                  warp lane l < NB owns column l of A_D' and ends with column l of inv_i = A_D'^{-1}
//                  (in-place rolled Gauss-Jordan), lane NB carries the g column of the coupling block
//                  (-> W[:,g]); inv_i goes to global memory as [16-byte chunk][row], the W columns
//                  (inv_i*A_U top-down, inv_i*A_L bottom-up) are handed to the next node in registers /
//                  by shuffles for the Schur update A_D' = A_D - A_L*W (resp. A_U*W^b).
//   residual_pass  lane per node: rhs = F(y) - Mass*(y+psi)/gamma  (K1 arithmetic on the
//                  shared-memory state)
//   forward_solve  two chains, one per half warp, lane = row:  z_i = inv_i*(rhs_i - A_L z_{i-1})
//   backward_solve same mapping:  d_i = z_i - inv_i*(A_U d_{i+1});  y += d; weighted max norms
//
// No tensor cores: the blocks are 7..13 wide and the chain over nodes is sequential.
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
    double* dJ;     // flux equations: dJ_r/dc_j(0), dJ_r/dphi(0)   [S][S+1]  (global)
    double* ring;   // shared: node records staged by cp.async ahead of the sweeps (+ weights, history chunks)
    double* scratch;            // shared, per warp
    const CellSpecies* sp;      // shared, per warp
    const DevTables* tb;        // shared, per block
    CellScalars cs;
    unsigned long long keep, stream;   // L2 cache policies: node records (evict_last) / Nordsieck history (evict_first)
    int N;          // n*NB
    int lane;
};

template <int NB, bool ST>
__host__ __device__ constexpr int scratch_doubles() { return 2 * (NB + 2) + 4 * NB + 2; }   // >= 4*padded (sweep buffers)
// sub-arrays of the per-cell workspace start on 32-byte boundaries
__host__ __device__ constexpr size_t align4(size_t doubles) { return (doubles + 3) & ~size_t(3); }
// padded row length of the stored blocks (even -> 16-byte aligned rows, double2 loads)
template <int NB, bool ST>
__host__ __device__ constexpr int padded() { return NB + (NB & 1); }
// doubles per node of the stored factors: inverse block (padded rows) + 4 coefficients per row
template <int NB, bool ST>
__host__ __device__ constexpr int fac_rec() { return NB * padded<NB, ST>() + NB * 4; }
// position of inv_i[r][c] inside a node record: [chunk c/2][row r] pairs of columns
template <int NB, bool ST>
__host__ __device__ constexpr int inv_off(int r, int c) { return ((c >> 1) * NB + r) * 2 + (c & 1); }
// Node records in flight in the solve sweeps: four per chain (three iterations of look-ahead).  One warp
// per cell walks two chains (8 slots).
constexpr int RING_CHAIN = 4;
constexpr int RING = 2 * RING_CHAIN;   // slots of the one-warp-per-cell ring

// ---------------------------------------------------------------------------
// Gauss-Jordan elimination with threshold partial pivoting; lane j owns column j
// (columns >= NB are right-hand sides).  The pivot column travels through shared
// memory (one writer, broadcast reads).  A row swap is a warp-uniform branch, so
// it only costs when it happens.  Returns false on a zero/non-finite pivot.
// Reciprocal of a pivot: hardware seed (MUFU.RCP64H, ~2^-20) + two Newton steps (2^-40, 2^-80), no slow path.
// Accurate to ~1 ulp for normal operands, which is all a factorisation needs (the correctly rounded
// __drcp_rn costs ~45 dependent instructions on the critical path of every elimination step);
// zero / denormal / huge pivots give Inf or NaN and are rejected by the caller.
__device__ __forceinline__ double pivot_rcp(double a) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    double e = fma(-a, x, 1.0);
    x = fma(x, e, x);
    e = fma(-a, x, 1.0);
    x = fma(x, e, x);
    return x;
}

// IN-PLACE Gauss-Jordan inversion with threshold row pivoting, TWO independent systems per warp: one per
// half warp (lanes 0..15 / 16..31; hbase = lane & 16, l = lane & 15).  Lane l < NB owns column l of the
// NB x NB matrix and ends up with column l of its inverse; lanes l >= NB carry right-hand-side columns
// (-> inverse times that column).  Compared with eliminating [A | I | rhs] this needs NB+1 instead of
// 2*NB+1 lanes per matrix, which is what lets the two halves of the twisted factorisation run
// concurrently in one warp (see factor_nodes).
//
// ROLLED form: the pivot row always sits in register slot 0 and the rows rotate up by one slot per
// step (after NB steps the pivot row of step k sits in slot k), so one loop body with static register
// indices serves all NB steps; the fully unrolled form is ~8x larger and made the factorisation loop
// overflow the SM's instruction cache.  Slots 0..NB-1-k hold the rows not yet used as
// pivots.  The pivot column is broadcast from lane k of the half warp by shuffles.  In place: when
// the row that originally was row p becomes the pivot row of step k, column k of the left part turns
// into a unit vector and column p of the (implicit) identity on the right stops being one; the latter
// is stored in the former's place, i.e. lane k finishes with column p_k of the inverse.  The pivot slots
// are only RECORDED in the loop (4 bits per step); when some half warp did swap rows (rare: the algebraic
// rows are pre-scaled so that the diagonal is an acceptable pivot) the bookkeeping is replayed afterwards
// and a final shuffle hands every lane its own column.  Returns false on a zero/non-finite pivot.
// 64-bit shuffle from two explicit 32-bit ones (the compiler's own expansion of __shfl_sync(double) came with
// register-pair swaps: three XORs per value in the elimination loop)
__device__ __forceinline__ double shfl_f64(double v, int src) {
    const int lo = __shfl_sync(FULL, __double2loint(v), src);
    const int hi = __shfl_sync(FULL, __double2hiint(v), src);
    return __hiloint2double(hi, lo);
}

template <int NB, bool ST>
__device__ __forceinline__ bool gauss_jordan(double (&A)[NB], int l, int hbase) {
    static_assert(NB <= 15, "NB+1 columns per half warp");
    bool bad = false;
    unsigned long long plist = 0;                      // pivot slot of every step, 4 bits each (all zero: no swap)
#pragma unroll 1
    for (int k = 0; k < NB; ++k) {
        // reciprocal of the diagonal candidate, computed by every lane for its own column BEFORE the pivot
        // decision: in the common case (no swap) the owner's value is simply broadcast with the column and
        // the MUFU + Newton chain (~60 cycles) runs beside the pivot search instead of after the shuffles
        double myinv = pivot_rcp(A[0]);
        // magnitude keys: high word of |a| with the slot index in the low 4 bits; tree reduction
        int key[NB];
        key[0] = __double2hiint(A[0]) & 0x7ffffff0;
#pragma unroll
        for (int r = 1; r < NB; ++r) key[r] = (r < NB - k) ? ((__double2hiint(A[r]) & 0x7ffffff0) | r) : 0;
        const int diag = key[0];
#pragma unroll
        for (int w = 1; w < NB; w <<= 1)
#pragma unroll
            for (int r = 0; r + w < NB; r += 2 * w) key[r] = max(key[r], key[r + w]);
        const int best = key[0];
        // keep the diagonal unless another entry is more than 8x larger (3 exponent steps)
        int p = best & 0xf;
        if (diag + (3 << 20) >= best) p = 0;
        p = __shfl_sync(FULL, p, hbase + k);          // the decision of the lane that owns column k
        if (p != 0) {                                  // uniform per half warp: slots 0 and p change places
            double a0 = A[0];
#pragma unroll
            for (int r = 1; r < NB; ++r) {
                if (r == p) { const double t = A[r]; A[r] = a0; a0 = t; }
            }
            A[0] = a0;
            myinv = pivot_rcp(a0);
            plist |= (unsigned long long)p << (4 * k);
        }
        double col[NB];
#pragma unroll
        for (int r = 0; r < NB; ++r) col[r] = shfl_f64(A[r], hbase + k);
        const double inv = shfl_f64(myinv, hbase + k);
        bad = bad || !(fabs(inv) < 1e300);             // zero / denormal / non-finite pivot
        const bool own = l == k;                       // the pivot column itself: continue with e_k in its place
        const double a0 = (own ? 1.0 : A[0]) * inv;
#pragma unroll
        for (int r = 1; r < NB; ++r) A[r - 1] = fma(-col[r], a0, own ? 0.0 : A[r]);     // eliminate and rotate
        A[NB - 1] = a0;
    }
    if (__any_sync(FULL, plist != 0)) {
        // rows were swapped in some half warp (rare: the algebraic rows are pre-scaled): replay the bookkeeping --
        // `orig` tracks the original row of every slot -- to find which lane holds my column of the inverse
        unsigned long long orig = 0xFEDCBA9876543210ull;
        int mysrc = l;
        for (int k = 0; k < NB; ++k) {
            const int p = (int)((plist >> (4 * k)) & 0xfull);
            if (p != 0) {
                const unsigned long long x = ((orig >> (4 * p)) ^ orig) & 0xfull;
                orig ^= x | (x << (4 * p));
            }
            const int pk = (int)(orig & 0xfull);       // original row of this step's pivot row
            if (pk == l) mysrc = k;
            orig = ((orig >> 4) & ~(0xfull << (4 * (NB - 1)))) | ((unsigned long long)pk << (4 * (NB - 1)));
        }
        const int src = hbase + (l < NB ? mysrc : l);
#pragma unroll
        for (int r = 0; r < NB; ++r) A[r] = shfl_f64(A[r], src);
    }
    return !bad;
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
#pragma unroll 1
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
__device__ __forceinline__ double4 node_coeff_row(const WarpState<NB, ST>& ws, const double* y, int i, int r) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    const int n = ws.cs.n;
    const bool mig = ws.tb->use_migration;
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
    return make_double4(l, a, ud, ua);
}

template <int NB, bool ST>
__device__ __forceinline__ void node_coeffs(const WarpState<NB, ST>& ws, const double* y, int i,
                                            double* sl, double* sa, double* sud, double* sua) {
    const int r = ws.lane;
    if (r >= NB) return;
    const double4 c = node_coeff_row<NB, ST>(ws, y, i, r);
    sl[r] = c.x; sa[r] = c.y; sud[r] = c.z; sua[r] = c.w;
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

// ---- asynchronous global->shared copies (LDGSTS) with explicit shared addresses ------------------
__device__ __forceinline__ void cp_async16(unsigned saddr, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(saddr), "l"(__cvta_generic_to_global(g)) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned saddr, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(saddr), "l"(__cvta_generic_to_global(g)) : "memory");
}
// ---- L2 eviction priorities ------------------------------------------------------------------------
// The per-cell workspace of a full launch (C2: 162 KB x 1024 cells = 166 MB) exceeds the 126 MB L2.  The node
// records are read by every solve sweep (twice per Newton iteration, by a sequential chain whose look-ahead is
// three nodes: a DRAM miss stalls the chain), the higher Nordsieck vectors only once per step by passes with
// independent, deeply prefetched chunks.  So the records are loaded and stored with L2::evict_last and the
// Nordsieck vectors zn[1..5] with L2::evict_first: the L2 then holds the records (104 MB) and the history
// streams from DRAM, where its latency is hidden.  CATINT_L2_HINTS=0 turns the hints off (A/B measurements).
#ifndef CATINT_L2_HINTS
#define CATINT_L2_HINTS 1
#endif
__device__ __forceinline__ unsigned long long l2_policy_keep() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_stream() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void cp_async16_hint(unsigned saddr, const void* g, unsigned long long pol) {
#if CATINT_L2_HINTS
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;"
                 :: "r"(saddr), "l"(__cvta_generic_to_global(g)), "l"(pol) : "memory");
#else
    cp_async16(saddr, g);
#endif
}
__device__ __forceinline__ void cp_async8_hint(unsigned saddr, const void* g, unsigned long long pol) {
#if CATINT_L2_HINTS
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;"
                 :: "r"(saddr), "l"(__cvta_generic_to_global(g)), "l"(pol) : "memory");
#else
    cp_async8(saddr, g);
#endif
}
__device__ __forceinline__ void st_hint(double* p, double v, unsigned long long pol) {
#if CATINT_L2_HINTS
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" :: "l"(__cvta_generic_to_global(p)), "d"(v), "l"(pol) : "memory");
#else
    *p = v;
#endif
}
__device__ __forceinline__ void st_hint4(double* p, double4 v, unsigned long long pol) {
#if CATINT_L2_HINTS
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" :: "l"(__cvta_generic_to_global(p)), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" :: "l"(__cvta_generic_to_global(p + 2)), "d"(v.z), "d"(v.w), "l"(pol) : "memory");
#else
    *reinterpret_cast<double4*>(p) = v;
#endif
}
// predicated copy (no branch): pred != 0 -> copy
__device__ __forceinline__ void cp_async16_if(unsigned saddr, const void* g, unsigned pred, unsigned long long pol) {
#if CATINT_L2_HINTS
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %3;\n\t}"
                 :: "r"(saddr), "l"(__cvta_generic_to_global(g)), "r"(pred), "l"(pol) : "memory");
#else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                 :: "r"(saddr), "l"(__cvta_generic_to_global(g)), "r"(pred) : "memory");
#endif
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// A sweep walks up to TWO independent chains at once, one per half warp (lanes 0..15 and 16..31, row
// r = lane & 15): with the twisted factorisation the upper chain (wall side) and the lower chain
// (bulk side) have no data dependence, so one warp advances both in the same instruction stream and
// the sequential depth of a sweep halves.  A chain is (first node, number of nodes, direction).
struct Chain { int first, count, dir; };

// Feeds the node records of a sweep into the shared-memory ring: per lane the running source
// pointers of its 16-byte chunks, one commit group per sweep iteration (both chains' records).
template <int NB, bool ST>
struct RecordFeed {
    static constexpr int REC = fac_rec<NB, ST>();
    static constexpr int CH = REC / 2;                   // 16-byte chunks per record
    static constexpr int ROUNDS = (CH + 31) / 32;        // rounds of the warp per record
    const double* s0; const double* s1;                  // this lane's first chunk of the next record, per chain
    long long st0, st1;                                  // doubles between consecutive records (+-REC)
    unsigned dst;                                        // shared address of this lane's first chunk in slot 0
    int n0, n1, two;
    int lane;
    unsigned long long pol;
    __device__ __forceinline__ void init(const double* fac, const double* ring, int lane, Chain c0, Chain c1) {
        s0 = fac + (long long)c0.first * REC + 2 * lane; st0 = (long long)c0.dir * REC;
        s1 = fac + (long long)c1.first * REC + 2 * lane; st1 = (long long)c1.dir * REC;
        dst = (unsigned)__cvta_generic_to_shared(ring) + 16u * lane;
        n0 = c0.count; n1 = c1.count; two = c1.count > 0 ? 2 : 1;
        this->lane = lane;
    }
    // records of iteration k (calls must come with k = 0, 1, 2, ...)
    __device__ __forceinline__ void issue(int k) {
        if (k < n0) {
            const unsigned d = dst + (unsigned)(((two * k) & (RING_CHAIN * two - 1)) * REC * 8);
#pragma unroll
            for (int q = 0; q < ROUNDS; ++q)
                if (q + 1 < ROUNDS || lane + 32 * q < CH) cp_async16_hint(d + 512u * q, s0 + 64 * q, pol);
            s0 += st0;
        }
        if (k < n1) {
            const unsigned d = dst + (unsigned)(((2 * k + 1) & (RING - 1)) * REC * 8);
#pragma unroll
            for (int q = 0; q < ROUNDS; ++q)
                if (q + 1 < ROUNDS || lane + 32 * q < CH) cp_async16_hint(d + 512u * q, s1 + 64 * q, pol);
            s1 += st1;
        }
    }
};

// ---------------------------------------------------------------------------
// Flux equations (pnp_fluxeq.cuh): the inward wall flux J = Jfix + sum_e coef[.][e]*E_e(c(0), phi(0)) of the
// current iterate, refreshed in the per-warp species table before every residual evaluation; lane r owns J_r.
// Out of line, plain arguments: models without flux equations (the common case) only pay a uniform test.
static __device__ __noinline__ void fluxeq_refresh(const DevTables* tb, const double* fpar, const double* y0,
                                                   double phi0, int S, int lane, CellSpecies* sp) {
    double add = 0.0;
    for (int e = 0; e < tb->fq.n_eq; ++e) {
        const double E = fluxeq_eval(&tb->fq, e, fpar, y0, phi0, -1, S, nullptr);
        if (lane < S) add += tb->fq.coef[lane][e] * E;
    }
    if (lane < S) sp->J[lane] = sp->Jfix[lane] + add;
    __syncwarp();
}

// dJ_r/dy_0 for the wall block of the Newton matrix -> dJ[r*(S+1) + comp]; lane = component comp
// (0..S-1: c_comp(0), S: phi(0))
static __device__ __noinline__ void fluxeq_jacobian(const DevTables* tb, const double* fpar, const double* y0,
                                                    double phi0, int S, int lane, double* dJ) {
    for (int r = 0; r < S; ++r) {
        double col = 0.0;
        for (int e = 0; e < tb->fq.n_eq; ++e) {
            double g = 0.0;
            fluxeq_eval(&tb->fq, e, fpar, y0, phi0, lane, S, &g);
            col = fma(tb->fq.coef[r][e], g, col);
        }
        if (lane <= S) dJ[r * (S + 1) + lane] = col;
    }
}

template <int NB, bool ST>
__device__ __forceinline__ void refresh_wall_flux(const WarpState<NB, ST>& ws, const double* y) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    if (ws.tb->fq.n_eq <= 0) return;
    fluxeq_refresh(ws.tb, ws.cs.fpar, y, ST ? y[S + 1] : ws.cs.phi_wall, S, ws.lane, const_cast<CellSpecies*>(ws.sp));
}

template <int NB, bool ST>
__device__ __forceinline__ void wall_flux_jacobian(const WarpState<NB, ST>& ws, const double* y) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    if (ws.tb->fq.n_eq <= 0) return;
    fluxeq_jacobian(ws.tb, ws.cs.fpar, y, ST ? y[S + 1] : ws.cs.phi_wall, S, ws.lane, ws.dJ);
}

// ---------------------------------------------------------------------------
// Twisted block factorisation of the Newton matrix, BOTH halves at once: the lower half warp
// (lanes 0..15) eliminates the nodes 0..mid downwards, the upper half warp (lanes 16..31) the nodes
// n-1..mid+1 upwards, in the same instruction stream.  Inside a half warp (l = lane & 15):
// lanes l < NB own column l of A_D' (-> column l of inv_i, in-place Gauss-Jordan), lane NB carries the
// g column u_g of the coupling block (-> W[:,g] = inv_i*u_g).
//
//   top    (nodes 0..mid):    A_D'_i = A_D,i - A_L,i * W_{i-1},   W_i = inv_i*A_U,i,  A_U = -(diag ud + ua e_g^T)
//       node 0 carries the extra wall block A_E (V_0 = inv_0*A_E), which makes A_U of node 1 dense:
//       W_1 = inv_1*(A_U,1 - A_L,1*V_0) is formed as an explicit product (inv_1 read back from the
//       node record) and stored; the coupling node `mid` also subtracts A_U * W^b_{mid+1}, which it
//       receives from the other half warp by shuffles.
//   bottom (nodes n-1..mid+1): A_D'_i = A_D,i - A_U,i * W^b_{i+1},  W^b_i = inv_i*A_L,i, A_L = -(diag l + a e_g^T)
//
// Because A_L and A_U are diagonal plus the g column, column j of W is the scaled column j of inv_i
// (already in lane j) or the result of the g lane: the Schur update of the next node needs no data
// from other lanes except that one column.  Stores inv_i and the sparse coefficients of A_L/A_U per
// node, V_0 and W_1.  Every interior node of either half passes through the SAME compact elimination
// code; the wall/bulk pair, node 1 and the coupling node take the SPECIAL instance, which keeps their
// rarely executed code out of the hot loop's instruction footprint (instruction cache).
template <int NB, bool ST, bool SPECIAL>
__device__ __forceinline__ bool eliminate_pair(const WarpState<NB, ST>& ws, double (&Wp)[NB],
                                               int k, int iters, int i, bool live, int slot, double inv_gamma) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    constexpr int NBP = padded<NB, ST>();
    constexpr int REC = fac_rec<NB, ST>();
    const int lane = ws.lane;
    const int n = ws.cs.n;
    const int l = lane & 15, hbase = lane & 16;
    const bool bottom = hbase != 0;
    const bool isD = l < NB, isG = l == NB;
    const int j = isD ? l : 0;
    const bool wall = SPECIAL && !bottom && k == 0;
    const bool bulk = SPECIAL && bottom && k == 0;
    const bool node1 = SPECIAL && !bottom && k == 1;
    const bool couple = SPECIAL && k == iters - 1;                  // warp-uniform; concerns the top half
    double A[NB];
    // the node's record in the ring: raw A_D and the coefficient rows (ud of the g row is scaled)
    const double* rr_ = ws.ring + (size_t)slot * REC;
    const double* cof = rr_ + NB * NBP;
#define C_L(r_) cof[4 * (r_)]
#define C_A(r_) cof[4 * (r_) + 1]
#define C_UD(r_) cof[4 * (r_) + 2]
#define C_UA(r_) cof[4 * (r_) + 3]
    double Wb[NB];
    if (SPECIAL && couple) {
        // W^b_{mid+1} column j from the lane that owns it in the other half warp
#pragma unroll
        for (int r = 0; r < NB; ++r) Wb[r] = __shfl_sync(FULL, Wp[r], (lane + 16) & 31);
    }
#pragma unroll
    for (int r = 0; r < NB; ++r) A[r] = 0.0;

    // ---- column of [A_D' | u_g] ----
    if (wall) {
        // A_D0 is diag(mass*inv_gamma - a); A_U0 = diag(-ud) + g-column(-ua); A_E = diag(-l)
        if (isD) {
#pragma unroll
            for (int r = 0; r < NB; ++r)
                if (r == j) A[r] = (j < S ? inv_gamma : 0.0) - C_A(r);
            if (ST && j == S) A[NB - 1] = ws.cs.eps / ws.cs.cstern;             // -dF_phi/dg_0 (Robin row)
            if (ws.tb->fq.n_eq > 0) {
                // flux equations: -dF_r/dy_0 = -ih0*dJ_r/dy_0 makes the wall block dense
                const int comp = j < S ? j : ((ST && j == NB - 1) ? S : -1);
                if (comp >= 0) {
                    const double ih0 = wall_coef(ws.cs).ih0;
#pragma unroll
                    for (int r = 0; r < S; ++r) A[r] -= ih0 * __ldcg(ws.dJ + r * (S + 1) + comp);
                }
            }
        } else if (isG) {
#pragma unroll
            for (int r = 0; r < NB; ++r) A[r] = r < S ? -C_UA(r) : -C_UD(r);
        }
    } else if (bulk) {
        // bulk node: c rows identity; default Poisson BCs: g row identity;
        // Stern: row S: phi_{n-1} = 0; row P: phi recursion with h_{n-2}
        if (isD) {
#pragma unroll
            for (int r = 0; r < S; ++r)
                if (r == j) A[r] = 1.0;
            if (!ST) {
                if (j == S) A[S] = 1.0;
            } else {
                const NodeCoef kc = interior_coef(ws.cs, n - 2);
                if (j == S) A[NB - 1] = -kc.hi;
                if (j == NB - 1) { A[S] = 1.0; A[NB - 1] = 1.0; }
            }
        } else if (isG) {
#pragma unroll
            for (int r = 0; r < NB; ++r) A[r] = r < S ? -C_A(r) : 0.0;
        }
    } else if (isD) {
#pragma unroll
        for (int r = 0; r < NB; ++r) A[r] = rr_[r * NBP + j];
        const double wg = Wp[S];
        if (bottom) {
            // - A_U * W^b_{i+1}:  A_U = -(diag ud + ua e_g^T), g row: -ud*sg
#pragma unroll
            for (int r = 0; r < S; ++r) A[r] += C_UD(r) * Wp[r] + C_UA(r) * wg;
            A[S] += C_UD(S) * wg;
        } else {
            // - A_L * W_{i-1}:  A_L = -(diag l + a e_g^T)
#pragma unroll
            for (int r = 0; r < S; ++r) A[r] += C_L(r) * Wp[r] + C_A(r) * wg;
            if (ST) A[NB - 1] += C_L(NB - 1) * Wp[NB - 1];                      // phi row: A_L = -1 on the diagonal
            if (SPECIAL && couple) {
                const double xg = Wb[S];
#pragma unroll
                for (int r = 0; r < S; ++r) A[r] += C_UD(r) * Wb[r] + C_UA(r) * xg;
                A[S] += C_UD(S) * xg;
            }
        }
    } else if (isG) {
        if (bottom) {
            // g-column of A_L: -a_r on the transport rows
#pragma unroll
            for (int r = 0; r < NB; ++r) A[r] = r < S ? -C_A(r) : 0.0;
        } else {
#pragma unroll
            for (int r = 0; r < NB; ++r) A[r] = r < S ? -C_UA(r) : -C_UD(r);
        }
    }
    // scale of this lane's column of the coupling block, read before the ring slot may be refilled:
    // top A_U = diag(-ud) on the c columns (no phi column), bottom A_L = diag(-l) on the c and phi columns;
    // the g column comes from the G lane
    const double c_l = C_L(j), c_ud = C_UD(j);
    double cj = bottom ? -c_l : (j < S ? -c_ud : 0.0);
    if (j == S) cj = 1.0;
    // node 1: column j of A_U1' = A_U1 - A_L1*V_0 (dense in general)
    double U1[NB];
    if (SPECIAL && node1) {
        // column j of V_0, written by the wall iteration (a warp barrier ago)
        double V0p[NB];
#pragma unroll
        for (int r = 0; r < NB; ++r) V0p[r] = __ldcg(ws.V0 + r * NBP + j);
        const double vg = V0p[S];
#pragma unroll
        for (int r = 0; r < NB; ++r) {
            double v = 0.0;
            if (r == j) v = -C_UD(r);
            if (j == S && r < S) v = -C_UA(r);
            if (r < S) v += C_L(r) * V0p[r] + C_A(r) * vg;
            if (ST && r == NB - 1) v += C_L(r) * V0p[r];
            U1[r] = v;
        }
    }
#undef C_L
#undef C_A
#undef C_UD
#undef C_UA

    const bool ok = gauss_jordan<NB, ST>(A, l, hbase);

    // ---- store the inverse, form the W column for the next node ----
    double* rec = ws.fac + (size_t)i * REC;
    if (isD && live) {
#pragma unroll
        for (int r = 0; r < NB; ++r) st_hint(rec + inv_off<NB, ST>(r, 0) + ((j >> 1) * NB * 2 + (j & 1)), A[r], ws.keep);
    }
    if (SPECIAL && wall && isD) {
        // V_0 = inv_0*A_E, A_E = diag(-l): needed for the modified A_U of node 1 and by the back substitution
        double* v0col = ws.V0 + j;
#pragma unroll
        for (int r = 0; r < NB; ++r) v0col[r * NBP] = A[r] * (-c_l);
    }
    const int src = (isD && j == S) ? hbase + NB : lane;
#pragma unroll
    for (int r = 0; r < NB; ++r) {
        const double t = __shfl_sync(FULL, A[r], src);
        if (live) Wp[r] = t * cj;
    }
    if (SPECIAL && k == 1) {
        // W_1 = inv_1 * A_U1' as an explicit product; inv_1 comes back from the node record (L2)
        __syncwarp();
        if (node1 && isD) {
#pragma unroll
            for (int r = 0; r < NB; ++r) {
                double s_ = 0.0;
#pragma unroll
                for (int c = 0; c < NB; ++c) s_ = fma(__ldcg(rec + inv_off<NB, ST>(r, c)), U1[c], s_);
                Wp[r] = s_;
            }
            double* w1col = ws.W1 + j;
#pragma unroll
            for (int r = 0; r < NB; ++r) w1col[r * NBP] = Wp[r];
        }
    }
    return ok || !live;
}

// The warp state is passed BY VALUE: handing out a reference would force the caller's copy into
// local memory and turn every shared-memory access of the whole kernel into a generic one.  Pointers
// that cross the call lose their address space, so the shared-memory ones are rebuilt here from
// byte offsets into the block's dynamic shared memory (SmemOffsets) and come out as LDS/STS again.
//
// Two phases.  ASSEMBLY: the Jacobian blocks of all nodes are independent of each other, so every
// lane assembles one column of A_D and one row of coefficients of a different (node, unknown) pair
// at a time, with the whole warp busy, straight into the node records in global memory.
// ELIMINATION: the sequential sweep then only streams the records back through the cp.async ring
// (three nodes ahead, both chains), applies the Schur update and eliminates.  Doing the assembly
// inside the sequential sweep (one node at a time, every load latency exposed) cost more than the
// elimination itself.
struct SmemOffsets { unsigned scratch, sp, y, ring; };

template <int NB, bool ST, bool SMEM>
__device__ __noinline__ bool factor_nodes(const WarpState<NB, ST> ws_in, const SmemOffsets so, double inv_gamma,
                                          int mid, long long* prof_assembly) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    constexpr int NBP = padded<NB, ST>();
    constexpr int REC = fac_rec<NB, ST>();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpState<NB, ST> ws = ws_in;
    ws.scratch = reinterpret_cast<double*>(smem_raw + so.scratch);
    ws.ring = reinterpret_cast<double*>(smem_raw + so.ring);
    ws.sp = reinterpret_cast<const CellSpecies*>(smem_raw + so.sp);
    ws.tb = reinterpret_cast<const DevTables*>(smem_raw);
    if constexpr (SMEM) ws.y = reinterpret_cast<double*>(smem_raw + so.y);
    const int lane = ws.lane;
    const int n = ws.cs.n;
    const bool mig = ws.tb->use_migration;
    const double* y = ws.y;

    const long long t_asm = prof_assembly ? clock64() : 0;
    wall_flux_jacobian<NB, ST>(ws, y);            // flux equations: dJ/dy_0 for the wall block (no-op without them)
    // ---- assembly of all node records: [ A_D (raw, interior nodes) | l, a, ud, ua per row ] ----
    for (int item = lane; item < n * NB; item += 32) {
        const int i = item / NB, r = item - i * NB;
        double* rec = ws.fac + (size_t)i * REC;
        double4 co = node_coeff_row<NB, ST>(ws, y, i, r);
        if (i > 0 && i < n - 1) {
            const NodeCoef k = interior_coef(ws.cs, i);
            const double sg = mig ? grow_scale(ws.cs, k.hi) : 1.0;
            if (r == S) co.z *= sg;                                  // g row of A_U carries the row scale
            double C[NB];
            interior_diag_column<NB, ST>(ws, y + (size_t)i * NB, r, k, inv_gamma, sg, C);
#pragma unroll
            for (int rr = 0; rr < NB; ++rr) st_hint(rec + rr * NBP + r, C[rr], ws.keep);
        }
        st_hint4(rec + NB * NBP + 4 * r, co, ws.keep);
    }
    __syncwarp();
    if (prof_assembly && lane == 0) *prof_assembly += clock64() - t_asm;     // debug profile: cycles of the assembly

    // ---- elimination: top chain 0..mid in the lower half warp, bottom chain n-1..mid+1 in the upper ----
    const int grp = lane >> 4;
    const int iters = mid + 1;                   // nodes of the top chain (>= those of the bottom chain)
    const int nbot = n - 1 - mid;
    double Wp[NB];
#pragma unroll
    for (int r = 0; r < NB; ++r) Wp[r] = 0.0;
    bool ok = true;
    RecordFeed<NB, ST> feed;
    feed.init(ws.fac, ws.ring, lane, Chain{0, iters, +1}, Chain{n - 1, nbot, -1});
    feed.pol = ws.keep;
#pragma unroll 1
    for (int p_ = 0; p_ < RING_CHAIN - 1; ++p_) { feed.issue(p_); cp_commit(); }
#pragma unroll 1
    for (int k = 0; k < iters; ++k) {
        cp_wait<RING_CHAIN - 2>();
        __syncwarp();
        const int slot = (2 * k + grp) & (RING - 1);
        const int i = grp ? n - 1 - k : k;
        const bool live = grp ? k < nbot : true;
        const bool special = (k <= 1) || (k == iters - 1);
        if (special) ok = eliminate_pair<NB, ST, true>(ws, Wp, k, iters, i, live, slot, inv_gamma) && ok;
        else ok = eliminate_pair<NB, ST, false>(ws, Wp, k, iters, i, live, slot, inv_gamma) && ok;
        __syncwarp();                           // every lane is done with this iteration's ring slots
        feed.issue(k + RING_CHAIN - 1);
        cp_commit();
    }
    cp_wait<0>();
    __syncwarp();
    return __all_sync(FULL, ok);
}

// ---------------------------------------------------------------------------
// rhs = F(y) - Mass*(y+psi)*inv_gamma for every unknown; one lane per node (mass term fused).
// The g-row of interior nodes carries the same scale as in factor_sweep.
#ifndef CATINT_RES_UNROLL
#define CATINT_RES_UNROLL 1
#endif
constexpr int RES_UNROLL = CATINT_RES_UNROLL;     // unroll factor of the reaction loop in residual_pass (measured: 2 neutral,
                                                  // 5 slows EVERY phase by ~5 % -- instruction-cache footprint)

template <int NB, bool ST>
__device__ void residual_pass(WarpState<NB, ST>& ws, double inv_gamma) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    const int n = ws.cs.n;
    const DevTables& tb = *ws.tb;
    const double* y = ws.y;
    refresh_wall_flux<NB, ST>(ws, y);
    for (int i = ws.lane; i < n; i += 32) {
        const double* y0 = y + (size_t)i * NB;
        double* out = ws.zb + (size_t)i * NB;
        // mass term of the BDF corrector: psi streams from global memory, loaded first so that its
        // latency hides behind the stencil arithmetic
        double ps[S];
#pragma unroll
        for (int r = 0; r < S; ++r) ps[r] = (inv_gamma != 0.0 && i < n - 1) ? ws.psi[(size_t)i * NB + r] : 0.0;
        if (i == 0 || i == n - 1) {
#pragma unroll 1
            for (int r = 0; r < NB; ++r) {
                double v = row_residual<NB, ST>(ws, y, i, r);
                if (inv_gamma != 0.0 && r < S && i == 0) v -= (y0[r] + ws.psi[r]) * inv_gamma;
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
            F[r] = fma(-(c0 + ps[r]), inv_gamma, F[r]);
            rho = fma(ws.sp->qe[r], c0, rho);
        }
#pragma unroll RES_UNROLL
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
}

// ---------------------------------------------------------------------------
// Solve sweeps with the stored factors.  Both chains of the twisted factorisation advance in the same
// instruction stream: the lower half warp (lanes 0..15) walks the wall-side chain, the upper half warp
// (lanes 16..31) the bulk-side chain; lane l = lane & 15 (< NB) owns row l of its chain's node.
//
// The node records stream from global memory (L2/HBM) into a shared-memory ring by cp.async, three
// iterations ahead of use, so that the sequential chain over the nodes never waits for DRAM.  (Loading the
// rows straight into registers with a software pipeline does not work: ptxas puts all the loads of all
// stages on one scoreboard, so the first use waits for the youngest load -- measured 1.8x slower.)
// Each half warp copies the records of its own chain (lane l: 16-byte chunks l, l+16, ...), with
// predicated copies instead of branches (unrolling the loop by the ring depth to get immediate slot offsets
// was measured slower: instruction-cache footprint).  One commit group and one warp barrier per iteration; the barrier
// publishes the right-hand sides tt (double buffered), makes the NEXT record visible (every lane has
// just waited for its own copies of it) and tells that all lanes are done with the previous one.  The
// predecessor's solution travels in registers (own row) and by one shuffle (g row).
template <int NB, bool ST>
struct SweepRing {
    static constexpr int NBP = padded<NB, ST>();
    static constexpr int REC = fac_rec<NB, ST>();
    static constexpr int CH = REC / 2;                    // 16-byte chunks per record
    static constexpr int R16 = (CH + 15) / 16;            // copy rounds of a half warp per record
    static constexpr int SLOT = R16 * 32 + 4 * NBP;       // doubles per slot: record (padded) | weights | zn0 | zb | y
    const double* src;                                    // my first chunk of the next record to copy
    // per-lane side streams of my unknown: error weight and zn0 (backward sweep), and -- when the iterate and the
    // rhs/update vector live in global memory (large grids) -- zb and y, so that no global load sits in the chain
    const double* wsrc; const double* zsrc; const double* bsrc; const double* ysrc;
    unsigned dst, wdst;                                   // shared address of my first chunk / my weight in slot (0, grp)
    int rstride, zstep, left;                             // doubles between records / unknowns; records left to copy
    unsigned lastp;                                       // copy predicate of the last round
    unsigned long long pol;
    bool wl, bl, yl;
    __device__ __forceinline__ void init(const WarpState<NB, ST>& ws, int first, int dir, int count, bool weights, int r,
                                         bool rowlane, bool stage_zb = false, bool stage_y = false) {
        const int l = ws.lane & 15, grp = ws.lane >> 4;
        src = ws.fac + (size_t)first * REC + 2 * l;
        rstride = dir * REC;
        zstep = dir * NB;
        left = count;
        dst = (unsigned)__cvta_generic_to_shared(ws.ring) + (unsigned)(grp * SLOT * 8 + 16 * l);
        wdst = (unsigned)__cvta_generic_to_shared(ws.ring) + (unsigned)((grp * SLOT + R16 * 32 + l) * 8);
        lastp = (l + 16 * (R16 - 1) < CH) ? 1u : 0u;
        wl = weights && rowlane;
        bl = stage_zb && rowlane;
        yl = stage_y && rowlane;
        pol = ws.keep;
        wsrc = ws.ewt + (size_t)first * NB + r;
        zsrc = ws.zn + (size_t)first * NB + r;
        bsrc = ws.zb + (size_t)first * NB + r;
        ysrc = ws.y + (size_t)first * NB + r;
    }
    // copy the next record of my chain into ring slot `slot` (0..RING_CHAIN-1); past the end of the chain
    // the last record is copied again (never consumed by a live iteration)
    __device__ __forceinline__ void issue(int slot) {
        const unsigned d = dst + (unsigned)(slot * 2 * SLOT * 8);
#pragma unroll
        for (int q = 0; q < R16; ++q) {
            if (q + 1 < R16) cp_async16_hint(d + 256u * q, src + 32 * q, pol);
            else cp_async16_if(d + 256u * q, src + 32 * q, lastp, pol);
        }
        const unsigned dw = wdst + (unsigned)(slot * 2 * SLOT * 8);
        if (wl) {
            cp_async8(dw, wsrc);
            cp_async8(dw + 8u * NBP, zsrc);
        }
        if (bl) cp_async8(dw + 16u * NBP, bsrc);
        if (yl) cp_async8(dw + 24u * NBP, ysrc);
        cp_commit();
        --left;
        if (left > 0) { src += rstride; wsrc += zstep; zsrc += zstep; bsrc += zstep; ysrc += zstep; }
    }
};

template <int NB, bool ST>
__device__ __forceinline__ double row_dot(const double2 (&m)[padded<NB, ST>() / 2], const double* tt) {
    constexpr int H = padded<NB, ST>() / 2;
    const double2* t2 = reinterpret_cast<const double2*>(tt);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int c = 0; c < H; ++c) {
        const double2 t = t2[c];
        if (c % 3 == 0) s0 = fma(m[c].x, t.x, s0);
        else if (c % 3 == 1) s1 = fma(m[c].x, t.x, s1);
        else s2 = fma(m[c].x, t.x, s2);
        if (2 * c + 1 < NB) {
            if (c % 3 == 0) s1 = fma(m[c].y, t.y, s1);
            else if (c % 3 == 1) s2 = fma(m[c].y, t.y, s2);
            else s0 = fma(m[c].y, t.y, s0);
        }
    }
    return (s0 + s1) + s2;
}

// elimination of the right-hand side along both chains (zb <- z):
//   wall side,  nodes 0..mid-1 upwards:      z_i = inv_i*(rhs_i - A_L z_{i-1})
//   bulk side,  nodes n-1..mid+1 downwards:  z_i = inv_i*(rhs_i - A_U z_{i+1})
// (the first node of a chain has no predecessor).  Lanes without a row and the shorter chain in its
// missing last iteration run the same arithmetic on valid dummy operands and only skip the stores.
template <int NB, bool ST, bool GS>
__device__ void forward_solve(WarpState<NB, ST>& ws, int mid) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    constexpr int NBP = padded<NB, ST>();
    constexpr int H = NBP / 2;
    constexpr int SLOT = SweepRing<NB, ST>::SLOT;
    static_assert(NB <= 16, "a chain occupies half a warp");
    const int lane = ws.lane, n = ws.cs.n;
    const int grp = lane >> 4, r0 = lane & 15;
    const bool rowlane = r0 < NB;
    const int r = rowlane ? r0 : 0;
    const int cnt0 = mid, cnt1 = n - 1 - mid;
    const int count = grp ? cnt1 : cnt0;
    const int iters = max(cnt0, cnt1);
    const int dir = grp ? -1 : 1;
    const int first = grp ? n - 1 : 0;
    SweepRing<NB, ST> feed;
    feed.init(ws, first, dir, count, false, r, rowlane, GS);
#pragma unroll
    for (int p = 0; p < RING_CHAIN - 1; ++p) feed.issue(p);
    // my row of the record in slot (0, grp): chunk c at [c][r]; coupling coefficients (diagonal, g column)
    // A_L = -(diag l + a e_g^T) on the wall side, A_U = -(diag ud + ua e_g^T) on the bulk side
    const double2* rowp = reinterpret_cast<const double2*>(ws.ring + grp * SLOT) + r;
    const double2* cop = reinterpret_cast<const double2*>(ws.ring + grp * SLOT + NB * NBP + 4 * r + (grp ? 2 : 0));
    const double* bp = ws.ring + grp * SLOT + SweepRing<NB, ST>::R16 * 32 + 2 * NBP + r;     // staged zb (GS)
    double* tbuf = ws.scratch + grp * 2 * NBP;           // [parity][NBP] per chain
    const int zstep = dir * NB;
    const int gsrc = (lane & 16) + S;                    // lane that owns the g component of my chain
    int zo = first * NB + r;                             // this lane's unknown of the current node
    double zprev = 0.0, zs = 0.0;                        // my row / the g row of the previous node (none at k = 0)
    cp_wait<RING_CHAIN - 2>();
    __syncwarp();                                        // record 0 is visible to every lane
    double rhs = GS ? bp[0] : ws.zb[zo];
    int rs = 1, is = RING_CHAIN - 1, tp = 0;             // ring slot of the NEXT record / slot refilled in this iteration, tt parity offset
    // my row of inv and my coupling coefficients travel one iteration ahead in registers: they are loaded right
    // after the barrier that makes their record visible, so no shared-memory latency sits at the head of the
    // next iteration's dependency chain (the loads reuse the registers row_dot has just read)
    double2 m[H];
#pragma unroll
    for (int c = 0; c < H; ++c) m[c] = rowp[c * NB];
    double2 cf = cop[0];
#pragma unroll 1
    for (int k = 0; k < iters; ++k) {
        feed.issue(is);                                  // refills the slot of record k-1
        is = (is + 1) & (RING_CHAIN - 1);
        const double t = fma(cf.y, zs, fma(cf.x, zprev, rhs));
        double* tt = tbuf + tp;
        tp ^= NBP;
        if (rowlane) tt[r] = t;
        cp_wait<RING_CHAIN - 2>();                       // my copies of record k+1 have landed
        __syncwarp();
        const bool live = k < count;
        const int zo_now = zo;
        if (k + 1 < count) zo += zstep;
        rhs = GS ? bp[rs * 2 * SLOT] : ws.zb[zo];        // GS: my copy of zb of node k+1 landed with its record
        const double z = row_dot<NB, ST>(m, tt);
#pragma unroll
        for (int c = 0; c < H; ++c) m[c] = rowp[rs * SLOT + c * NB];
        cf = cop[rs * SLOT];
        rs = (rs + 1) & (RING_CHAIN - 1);
        if (rowlane && live) ws.zb[zo_now] = z;
        zprev = z;
        zs = __shfl_sync(FULL, z, gsrc);
    }
    cp_wait<0>();
    __syncwarp();
}

// Back substitution along both chains, the solution of the coupling node `mid` being final already in zb:
//   wall side,  nodes mid-1..0 downwards:   d_i = z_i - inv_i*(A_U d_{i+1})   (node 1: dense W_1; node 0: extra wall block V_0)
//   bulk side,  nodes mid+1..n-1 upwards:   d_i = z_i - inv_i*(A_L d_{i-1})
// y += scale*d, zb <- d.  Fused with the weighted max norms of the Newton update (|scale*d|*w)
// and of the accumulated correction (|y-zn0|*w) over the error-controlled unknowns
// (concentrations of nodes 0..n-2), accumulated into dmax/amax (per lane, reduce afterwards):
// the weight and zn0 of each unknown ride in the ring slot of their node (a lane reads back only what it
// copied itself).  wmode 0: weights ws.ewt; wmode 1 (steady polish): w = 1/(prtol*|y|+patol).
template <int NB, bool ST, bool GS>
__device__ void backward_solve(WarpState<NB, ST>& ws, double scale, int mid,
                               double& dmax, double& amax, int wmode, double prtol, double patol) {
    constexpr int S = NB - 1 - (ST ? 1 : 0);
    constexpr int NBP = padded<NB, ST>();
    constexpr int H = NBP / 2;
    constexpr int SLOT = SweepRing<NB, ST>::SLOT;
    constexpr int WOFF = SweepRing<NB, ST>::R16 * 32;    // weights / zn0 inside a slot
    const int lane = ws.lane, n = ws.cs.n;
    const int grp = lane >> 4, r0 = lane & 15;
    const bool rowlane = r0 < NB;
    const int r = rowlane ? r0 : 0;
    const int cnt0 = mid, cnt1 = n - 1 - mid;
    const int count = grp ? cnt1 : cnt0;
    const int iters = max(cnt0, cnt1);
    const int dir = grp ? 1 : -1;
    const int first = grp ? mid + 1 : mid - 1;
    const bool wl = wmode == 0;
    SweepRing<NB, ST> feed;
    feed.init(ws, first, dir, count, wl, r, rowlane, GS, GS);
#pragma unroll
    for (int p = 0; p < RING_CHAIN - 1; ++p) feed.issue(p);
    const double2* rowp = reinterpret_cast<const double2*>(ws.ring + grp * SLOT) + r;
    const double2* cop = reinterpret_cast<const double2*>(ws.ring + grp * SLOT + NB * NBP + 4 * r + (grp ? 0 : 2));
    const double* wp = ws.ring + grp * SLOT + WOFF + r;  // staged: +0 weight, +NBP zn0, +2 NBP zb, +3 NBP y
    double* tbuf = ws.scratch + grp * 2 * NBP;
    const int zstep = dir * NB;
    const int gsrc = (lane & 16) + S;
    int zo = first * NB + r;
    const bool crow = rowlane && r < S;                  // this lane owns a concentration unknown
    // final solution of the node before the chain (the coupling node): my row and its g row
    double dprev = ws.zb[zo - zstep];
    double dg = ws.zb[zo - zstep + (S - r)];
    cp_wait<RING_CHAIN - 2>();
    __syncwarp();                                        // record 0 is visible to every lane
    double zcur = GS ? wp[2 * NBP] : ws.zb[zo];          // forward-eliminated value of the current node
    double ynext = GS ? wp[3 * NBP] : 0.0;               // GS: iterate of the current node, staged with its record
    int rs = 1, is = RING_CHAIN - 1, tp = 0;
    double2 m[H];                                        // one iteration ahead in registers, see forward_solve
#pragma unroll
    for (int c = 0; c < H; ++c) m[c] = rowp[c * NB];
    double2 cf = cop[0];
    double wn = 0.0, z0n = 0.0;
    if (wl) { wn = wp[0]; z0n = wp[NBP]; }
#pragma unroll 1
    for (int k = 0; k < iters; ++k) {
        feed.issue(is);
        is = (is + 1) & (RING_CHAIN - 1);
        double w = wn;
        const double z0 = z0n;
        double* tt = tbuf + tp;
        tp ^= NBP;
        if (rowlane) tt[r] = -fma(cf.x, dprev, cf.y * dg);
        cp_wait<RING_CHAIN - 2>();                       // my copies of record k+1 (my weights of node k landed earlier)
        __syncwarp();
        const bool live = k < count;
        const int i = first + dir * k;
        const int zo_now = zo;
        if (k + 1 < count) zo += zstep;
        const double znext = GS ? wp[rs * 2 * SLOT + 2 * NBP] : ws.zb[zo];
        const double yold = GS ? ynext : ws.y[zo_now];
        if (GS) ynext = wp[rs * 2 * SLOT + 3 * NBP];
        double d = zcur - row_dot<NB, ST>(m, tt);
#pragma unroll
        for (int c = 0; c < H; ++c) m[c] = rowp[rs * SLOT + c * NB];
        cf = cop[rs * SLOT];
        if (wl) { wn = wp[rs * 2 * SLOT]; z0n = wp[rs * 2 * SLOT + NBP]; }
        rs = (rs + 1) & (RING_CHAIN - 1);
        // wall end of the wall-side chain (warp-uniform test): node 1 couples through the dense W_1,
        // node 0 has the extra block V_0 towards node 2
        const int i0 = mid - 1 - k;
        if (i0 <= 1 && i0 >= 0) {
            if (grp == 0 && rowlane) {
                const double* d2 = ws.zb + 2 * NB;
                const double* Mr = (i0 == 1 ? ws.W1 : ws.V0) + (size_t)r * NBP;
                double s_ = 0.0;
#pragma unroll
                for (int c = 0; c < NB; ++c) s_ = fma(Mr[c], d2[c], s_);
                d = (i0 == 1 ? zcur : d) - s_;
            }
        }
        const double ds = d * scale;
        const double yn = yold + ds;
        if (rowlane && live) { ws.zb[zo_now] = d; ws.y[zo_now] = yn; }
        dprev = d;
        dg = __shfl_sync(FULL, d, gsrc);
        zcur = znext;
        {
            if (!wl) w = 1.0 / (prtol * fabs(yn) + patol);
            double ad = fabs(ds) * w;
            if (!(ad <= 1e300)) ad = INFINITY;          // NaN/Inf must not be lost in fmax
            const bool counted = crow && live && i < n - 1;
            dmax = fmax(dmax, counted ? ad : 0.0);
            if (wl) amax = fmax(amax, counted ? fabs(yn - z0) * w : 0.0);
        }
    }
    cp_wait<0>();
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
        for (int c = 0; c < NB; ++c) row[c] = rec[inv_off<NB, ST>(0, c) + 2 * r];
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

}  // namespace catint
