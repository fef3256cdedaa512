// pnp_explicit.cuh -- K4: the reference's fixed-step steppers and its Poisson routine on the device.
//
//   potential_and_gradient   get_potential_and_gradient (/root/reference/catint/calculator_old.py:680-819) for EVERY
//                            combination of pb_bound the reference accepts (:776-803): potential at wall and bulk
//                            (tridiagonal Poisson solve, central gradient, extrapolated ends), or one gradient
//                            (wall: forward sum, bulk: backward sum) + one potential (wall: forward sum, bulk:
//                            backward sum).  Uniform mesh (the reference's scalar dx).
//   pnp_explicit_kernel      integrate_FTCS (:976-1029) and integrate_Crank_Nicolson (:457-564): per time step the
//                            field of the current state, then every species on its own -- explicit Euler (optionally
//                            with the Lax-Friedrichs 1/2 terms) or a tridiagonal Crank-Nicolson solve with the lagged
//                            field; Robin wall condition from the imposed flux, Dirichlet bulk.
//   pnp_potential_kernel     the Poisson routine alone (parity / a13 entry point).
//
// The restatement is literal, quirks included (they are part of "the same results as the reference"):
//   * the Robin wall value uses v[1] - vzeta with vzeta = the wall potential (the reference reads system['vzeta'],
//     a key its own Transport rejects -- the steppers are orphaned upstream);
//   * Crank-Nicolson indexes the field by the INTERIOR index (unknown j sits on node j+1 but uses grad_v[j],
//     lapl_v[j]), multiplies the old state from the left (np.dot(C, B1), i.e. with the transposed matrix), and
//     has no reaction term;
//   * FTCS evaluates the rates before the wall update and uses grad_v[i+1] / grad_v[i-1] in the east/west weights.
// One warp per cell, state species-major in shared memory (lanes run over the nodes: conflict-free).  HBM only sees
// the initial state and the requested outputs; the step loop is on chip.
#pragma once
#include "pnp_device.cuh"

namespace catint {

struct ExplicitParams {
    DevTables tb;
    const double* par; const int* nx;
    const double* c0;          // [B][nx_max][S]
    long long n_cells;
    int method;                // 0 FTCS, 1 Crank-Nicolson
    int lax_friedrich;
    int nt, n_out;
    double dt;
    const int* itout;          // device [n_out] increasing step indices
    double* c_out;             // [n_out][B][nx_max][S]
    double* phi_out; double* g_out;   // optional [n_out][B][nx_max]
    int cells_per_block, np;   // np: padded nodes per species row in shared memory
    int bc;                    // CATINT_PNP_BC_* (all pb_bound combinations but Stern)
};

struct PotentialParams {
    DevTables tb;
    const double* par; const int* nx;
    const double* c;           // [B][nx_max][S]
    long long n_cells;
    double* v; double* grad; double* lapl;   // [B][nx_max]
    int np, bc;
};

__device__ __forceinline__ double wscan(double v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// var[i] = var[0] + sum_{j=1..i} f[j]*dx, i = 1..n-2 (integrate_1d_func, n=1, forward, :753-759)
__device__ __forceinline__ void sum_forward(double* var, const double* f, double dx, int n, int lane) {
    double carry = var[0];
    for (int base = 1; base <= n - 2; base += 32) {
        const int i = base + lane;
        const double s = wscan(i <= n - 2 ? f[i] * dx : 0.0, lane);
        if (i <= n - 2) var[i] = carry + s;
        carry += __shfl_sync(FULL, s, 31);
    }
    __syncwarp();
}
// var[i] = var[n-1] - sum_{j=i..n-2} f[j]*dx, i = n-2..1 (inverse direction)
__device__ __forceinline__ void sum_backward(double* var, const double* f, double dx, int n, int lane) {
    double carry = var[n - 1];
    for (int top = n - 2; top >= 1; top -= 32) {
        const int i = top - lane;
        const double s = wscan(i >= 1 ? f[i] * dx : 0.0, lane);
        if (i >= 1) var[i] = carry - s;
        carry -= __shfl_sync(FULL, s, 31);
    }
    __syncwarp();
}

// C: [S][np] species-major; outputs v, g (= grad_v), lp (= lapl_v) per node.  bc: CATINT_PNP_BC_*.
__device__ void potential_and_gradient(int bc, int n, int S, int np, double dx, const double* qe, const double* C,
                                       double phi_wall, double phi_bulk, double g_wall, double g_bulk,
                                       double* v, double* g, double* lp, int lane) {
    for (int i = lane; i < n; i += 32) {
        double r = 0.0;
        for (int k = 0; k < S; ++k) r -= qe[k] * C[k * np + i];            // :762-766
        lp[i] = r;
        v[i] = 0.0; g[i] = 0.0;
    }
    __syncwarp();
    const bool PW = bc == CATINT_PNP_BC_DIRICHLET_WALL_NEUMANN_BULK || bc == CATINT_PNP_BC_DIRICHLET_BOTH ||
                    bc == CATINT_PNP_BC_DIRICHLET_WALL_NEUMANN_WALL;
    const bool PB = bc == CATINT_PNP_BC_DIRICHLET_BOTH || bc == CATINT_PNP_BC_DIRICHLET_BULK_NEUMANN_WALL ||
                    bc == CATINT_PNP_BC_DIRICHLET_BULK_NEUMANN_BULK;
    const bool GW = bc == CATINT_PNP_BC_DIRICHLET_WALL_NEUMANN_WALL || bc == CATINT_PNP_BC_DIRICHLET_BULK_NEUMANN_WALL;
    const bool GB = bc == CATINT_PNP_BC_DIRICHLET_WALL_NEUMANN_BULK || bc == CATINT_PNP_BC_DIRICHLET_BULK_NEUMANN_BULK;
    if (lane == 0) {
        if (PW) v[0] = phi_wall;                                             // :771-774
        if (PB) v[n - 1] = phi_bulk;
    }
    __syncwarp();
    if (PW && PB) {
        // v'' = rhs with both ends given (solve_poisson, :718-731): v_i = v_0 + i*d_1 + T_i,
        // T_i = sum_{m=1}^{i-1} sum_{j=1}^{m} rhs_j*dx^2 -- two prefix sums instead of the dense solve
        if (lane == 0) g[0] = 0.0;
        __syncwarp();
        sum_forward(g, lp, dx * dx, n, lane);                                // g[m] = P_m, m = 1..n-2 (scratch use of g)
        if (lane == 0) { g[0] = 0.0; v[0] = 0.0; }
        __syncwarp();
        // T_i = sum_{m=1}^{i-1} P_m for i = 1..n-1: prefix sum of P shifted by one
        {
            double carry = 0.0;
            for (int base = 1; base <= n - 1; base += 32) {
                const int i = base + lane;
                const double s = wscan((i <= n - 1 && i >= 2) ? g[i - 1] : 0.0, lane);
                if (i <= n - 1) v[i] = carry + s;                            // T_i
                carry += __shfl_sync(FULL, s, 31);
            }
        }
        __syncwarp();
        const double d1 = (phi_bulk - phi_wall - v[n - 1]) / (double)(n - 1);
        __syncwarp();
        for (int i = lane; i < n; i += 32) v[i] = phi_wall + i * d1 + (i >= 1 ? v[i] : 0.0);
        __syncwarp();
        if (lane == 0) v[n - 1] = phi_bulk;
        __syncwarp();
        for (int i = 1 + lane; i <= n - 2; i += 32) g[i] = 1.0 / (2.0 * dx) * (v[i + 1] - v[i - 1]);      // :777-778
        __syncwarp();
        if (lane == 0) {
            g[0] = g[1] + (g[1] - g[2]);                                     // :779-780
            g[n - 1] = g[n - 2] + (g[n - 2] - g[n - 3]);
        }
        __syncwarp();
        return;
    }
    if (GW) {                                                                // :783-786
        if (lane == 0) g[0] = g_wall;
        __syncwarp();
        sum_forward(g, lp, dx, n, lane);
        if (lane == 0) g[n - 1] = g[n - 2] + (g[n - 2] - g[n - 3]);
        __syncwarp();
    }
    if (GB) {                                                                // :787-790
        if (lane == 0) g[n - 1] = g_bulk;
        __syncwarp();
        sum_backward(g, lp, dx, n, lane);
        if (lane == 0) g[0] = g[1] + (g[1] - g[2]);
        __syncwarp();
    }
    if (PW) {                                                                // :792-794
        sum_forward(v, g, dx, n, lane);
        if (lane == 0) v[n - 1] = v[n - 2] + (v[n - 2] - v[n - 3]);
        __syncwarp();
    }
    if (PB) {                                                                // :795-797
        sum_backward(v, g, dx, n, lane);                                     // var[i] = var[i+1] - g[i]*dx
        if (lane == 0) v[0] = v[1] + (v[1] - v[2]);
        __syncwarp();
    }
}

struct ExplicitCell {
    double D[MAXS], mu[MAXS], qe[MAXS], J[MAXS], cb[MAXS];
    double beta, eps, phi_wall, g_bulk, phi_bulk, g_wall, dx;
    int n;
};

__device__ __forceinline__ void load_explicit_cell(const DevTables& tb, const double* par, const int* nx, long long cell,
                                                   int lane, ExplicitCell* ce) {
    const int S = tb.S;
    const double* p = par + (size_t)cell * (3 * S + 8);
    if (lane == 0) {
        ce->n = nx[cell];
        ce->beta = p[3 * S + 0]; ce->eps = p[3 * S + 1];
        ce->phi_wall = p[3 * S + 2]; ce->g_bulk = p[3 * S + 3];
        ce->dx = p[3 * S + 5];
        ce->phi_bulk = p[3 * S + 6]; ce->g_wall = p[3 * S + 7];
    }
    if (lane < S) {
        const double q = tb.z[lane] * UNIT_F;
        ce->cb[lane] = p[lane]; ce->J[lane] = p[S + lane]; ce->D[lane] = p[2 * S + lane];
        ce->mu[lane] = p[2 * S + lane] * q * p[3 * S + 0];                    // transport.py:436
        ce->qe[lane] = q / p[3 * S + 1];
    }
    __syncwarp();
}

__global__ void __launch_bounds__(128) pnp_potential_kernel(PotentialParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ DevTables tbs;
    {
        const int words = (int)(sizeof(DevTables) / 4);
        const int* src = reinterpret_cast<const int*>(&P.tb);
        int* dst = reinterpret_cast<int*>(&tbs);
        for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long cell = (long long)blockIdx.x * 4 + warp;
    if (cell >= P.n_cells) return;
    const int S = tbs.S, np = P.np, nxm = tbs.nx_max;
    const size_t per_cell = ((sizeof(ExplicitCell) + 15) & ~size_t(15)) + (size_t)(S + 3) * np * sizeof(double);
    unsigned char* mine = smem_raw + (size_t)warp * per_cell;
    ExplicitCell* ce = reinterpret_cast<ExplicitCell*>(mine);
    double* C = reinterpret_cast<double*>(mine + ((sizeof(ExplicitCell) + 15) & ~size_t(15)));
    double* v = C + (size_t)S * np; double* g = v + np; double* lp = g + np;
    load_explicit_cell(tbs, P.par, P.nx, cell, lane, ce);
    const int n = ce->n;
    if (n < 4 || n > nxm) return;
    const double* src = P.c + (size_t)cell * nxm * S;
    for (int idx = lane; idx < n * S; idx += 32) { const int i = idx / S, k = idx - i * S; C[k * np + i] = src[idx]; }
    __syncwarp();
    potential_and_gradient(P.bc, n, S, np, ce->dx, ce->qe, C, ce->phi_wall, ce->phi_bulk, ce->g_wall, ce->g_bulk,
                           v, g, lp, lane);
    for (int i = lane; i < n; i += 32) {
        if (P.v) P.v[(size_t)cell * nxm + i] = v[i];
        if (P.grad) P.grad[(size_t)cell * nxm + i] = g[i];
        if (P.lapl) P.lapl[(size_t)cell * nxm + i] = lp[i];
    }
}

__global__ void __launch_bounds__(128) pnp_explicit_kernel(ExplicitParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ DevTables tbs;
    {
        const int words = (int)(sizeof(DevTables) / 4);
        const int* src = reinterpret_cast<const int*>(&P.tb);
        int* dst = reinterpret_cast<int*>(&tbs);
        for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
    }
    __syncthreads();
    const DevTables& tb = tbs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long cell = (long long)blockIdx.x * P.cells_per_block + warp;
    if (warp >= P.cells_per_block || cell >= P.n_cells) return;
    const int S = tb.S, np = P.np, nxm = tb.nx_max;
    const size_t per_cell = ((sizeof(ExplicitCell) + 15) & ~size_t(15)) + (size_t)(3 * S + 3) * np * sizeof(double);
    unsigned char* mine = smem_raw + (size_t)warp * per_cell;
    ExplicitCell* ce = reinterpret_cast<ExplicitCell*>(mine);
    double* C = reinterpret_cast<double*>(mine + ((sizeof(ExplicitCell) + 15) & ~size_t(15)));   // [S][np]
    double* T = C + (size_t)S * np;            // FTCS: new state; CN: right-hand side B, then the solution
    double* Rt = T + (size_t)S * np;           // FTCS: rates; CN: modified diagonal of the Thomas sweep
    double* v = Rt + (size_t)S * np; double* g = v + np; double* lp = g + np;
    load_explicit_cell(tb, P.par, P.nx, cell, lane, ce);
    const int n = ce->n;
    if (n < 4 || n > nxm) return;
    const double dx = ce->dx, dt = P.dt;
    const bool mig = tb.use_migration;
    const double* src = P.c0 + (size_t)cell * nxm * S;
    for (int idx = lane; idx < n * S; idx += 32) { const int i = idx / S, k = idx - i * S; C[k * np + i] = src[idx]; }
    for (int i = lane; i < n; i += 32) { v[i] = 0.0; g[i] = 0.0; lp[i] = 0.0; }
    __syncwarp();
    // Crank-Nicolson keeps the previous boundary values (COLD, :524-526); lane k owns species k
    double c0_old = 0.0, c1_old = 0.0;
    int iout = 0;
    const int n_first = P.method == 0 ? 0 : 1;                                // FTCS range(0,nt), CN range(1,nt)
    for (int step = n_first; step < P.nt; ++step) {
        if (mig) potential_and_gradient(P.bc, n, S, np, dx, ce->qe, C, ce->phi_wall, ce->phi_bulk, ce->g_wall,
                                        ce->g_bulk, v, g, lp, lane);
        const double vz = v[1] - ce->phi_wall;                                // v[1] - vzeta
        if (P.method == 0) {
            // ---- FTCS (:993-1023) ----
            for (int idx = lane; idx < n * S; idx += 32) {                    // rates = get_rates(C), all nodes
                const int k = idx / n, i = idx - k * n;
                double cn[MAXS];
                for (int s = 0; s < S; ++s) cn[s] = C[s * np + i];
                double r = 0.0;
                for (int rr = 0; rr < tb.R; ++rr) {
                    const double nu = tb.nu[rr][k];
                    if (nu != 0.0) r += nu * net_rate(tb, rr, cn);
                }
                Rt[k * np + i] = r;
            }
            __syncwarp();
            if (lane < S) {
                const int k = lane;
                const double D = ce->D[k], mu = ce->mu[k];
                const double divisor = 2.0 * D - mu * vz;                     // :1003
                C[k * np] = ((2.0 * D + mu * vz) * C[k * np + 1] + ce->J[k] * 2.0 * dx) / divisor;
                C[k * np + n - 1] = ce->cb[k];                                // :1008 (c0 = bulk everywhere)
            }
            __syncwarp();
            for (int idx = lane; idx < n * S; idx += 32) {
                const int k = idx / n, i = idx - k * n;
                const double* Ck = C + k * np;
                double out;
                if (i == 0 || i == n - 1) {
                    out = Ck[i];
                } else {
                    const double D = ce->D[k], mu = ce->mu[k];
                    double W = D * dt / (dx * dx) - dt / (2.0 * dx) * mu * g[i + 1] + 0.5;      // :1013-1018
                    double M = -2.0 * D * dt / (dx * dx);
                    double E = D * dt / (dx * dx) + dt / (2.0 * dx) * mu * g[i - 1] + 0.5;
                    if (!P.lax_friedrich) { W -= 0.5; E -= 0.5; M += 1.0; }
                    out = E * Ck[i - 1] + M * Ck[i] + W * Ck[i + 1] + Rt[k * np + i] * dt;
                }
                T[k * np + i] = out;
            }
            __syncwarp();
            for (int idx = lane; idx < n * S; idx += 32) { const int k = idx / n, i = idx - k * n; C[k * np + i] = T[k * np + i]; }
            __syncwarp();
        } else {
            // ---- Crank-Nicolson (:513-558), lane = species ----
            if (lane < S) {
                const int k = lane;
                double* Ck = C + k * np; double* B = T + k * np; double* dp = Rt + k * np;
                const double D = ce->D[k], mu = ce->mu[k];
                if (step == 1) { c0_old = Ck[0]; c1_old = Ck[n - 1]; }
                Ck[0] = (-2.0 * D - mu * vz) / (-2.0 * D + mu * vz) * Ck[1] - 2.0 * ce->J[k] * dx / (-2.0 * D + mu * vz);
                Ck[n - 1] = ce->cb[k];
                double s = D * dt / (dx * dx);
                if (P.lax_friedrich) s += 0.5;
                const double ee = mig ? tb.z[k] * UNIT_F * ce->beta * dt * D : 0.0;
                const double f4 = ee / (4.0 * dx);
                const int m = n - 2;                                           // interior unknowns j = 0..m-1 on nodes j+1
                // B = np.dot(C[k,1:-1], B1): B_j = C_{j-1}*B1[j-1,j] + C_j*B1[j,j] + C_{j+1}*B1[j+1,j]
                for (int j = 0; j < m; ++j) {
                    double b = Ck[j + 1] * (1.0 - s + ee * lp[j]);
                    if (j >= 1) b += Ck[j] * (0.5 * s + f4 * g[j - 1]);
                    if (j <= m - 2) b += Ck[j + 2] * (0.5 * s - f4 * g[j + 1]);
                    B[j] = b;
                }
                B[0] += (0.5 * s + f4 * g[0]) * (Ck[0] + c0_old);             // add_boundary_values
                B[m - 1] += (0.5 * s - f4 * g[n - 1]) * (Ck[n - 1] + c1_old);
                // A: diag 1+s, A[j,j-1] = -s/2 + f4*g[j], A[j,j+1] = -s/2 - f4*g[j]; Thomas without pivoting
                dp[0] = 1.0 + s;
                for (int j = 1; j < m; ++j) {
                    const double w = (-0.5 * s + f4 * g[j]) / dp[j - 1];
                    dp[j] = (1.0 + s) - w * (-0.5 * s - f4 * g[j - 1]);
                    B[j] -= w * B[j - 1];
                }
                B[m - 1] /= dp[m - 1];
                for (int j = m - 2; j >= 0; --j) B[j] = (B[j] - (-0.5 * s - f4 * g[j]) * B[j + 1]) / dp[j];
                for (int j = 0; j < m; ++j) Ck[j + 1] = B[j];
                c0_old = Ck[0]; c1_old = Ck[n - 1];
            }
            __syncwarp();
        }
        if (iout < P.n_out && step == P.itout[iout]) {
            double* co = P.c_out + ((size_t)iout * P.n_cells + cell) * nxm * S;
            for (int idx = lane; idx < n * S; idx += 32) { const int i = idx / S, k = idx - i * S; co[idx] = C[k * np + i]; }
            if (P.phi_out) for (int i = lane; i < n; i += 32) P.phi_out[((size_t)iout * P.n_cells + cell) * nxm + i] = v[i];
            if (P.g_out) for (int i = lane; i < n; i += 32) P.g_out[((size_t)iout * P.n_cells + cell) * nxm + i] = g[i];
            ++iout;
        }
        __syncwarp();
    }
}

inline int launch_explicit(ExplicitParams& P, cudaStream_t st) {
    int dev = 0, max_optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return CATINT_PNP_ECUDA;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    P.np = P.tb.nx_max | 1;                                                    // odd pitch
    const size_t per_cell = ((sizeof(ExplicitCell) + 15) & ~size_t(15)) + (size_t)(3 * P.tb.S + 3) * P.np * sizeof(double);
    int cpb = 4;
    while (cpb > 1 && per_cell * cpb > (size_t)max_optin) cpb >>= 1;
    if (per_cell * cpb > (size_t)max_optin) return CATINT_PNP_EINVAL;
    P.cells_per_block = cpb;
    const size_t smem = per_cell * cpb;
    cudaFuncSetAttribute(pnp_explicit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const unsigned grid = (unsigned)((P.n_cells + cpb - 1) / cpb);
    pnp_explicit_kernel<<<grid, 128, smem, st>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : CATINT_PNP_ECUDA;
}

// Potential of the pair "potential and gradient in the bulk" (calculator_old.py:795-797) from a gradient the implicit
// integrator has written: the dynamics of that pair are those of the default pair (the field is the backward sum from
// the bulk gradient either way, :787-790), only the potential is integrated from the other end:
//   v[n-1] = phi_bulk,  v[i] = v[i+1] - g[i]*(x_{i+1}-x_i)  (i = n-2..1),  v[0] = v[1] + (v[1]-v[2])*h_0/h_1.
// One warp per (output time, cell) row of g_out / phi_out.
struct BulkPotentialParams {
    const double* par; const int* nx; const int* mesh_id; const double* mesh_xi;
    const double* g; double* v;          // [rows][nx_max], rows = n_out * n_cells, cell = row % n_cells
    long long n_cells, rows;
    int S, nx_max;
};

__global__ void __launch_bounds__(128) pnp_bulk_potential_kernel(BulkPotentialParams P) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 4 + warp;
    if (row >= P.rows) return;
    const long long cell = row % P.n_cells;
    const double* p = P.par + (size_t)cell * (3 * P.S + 8);
    const int n = P.nx[cell];
    if (n < 4 || n > P.nx_max) return;
    const double dx = p[3 * P.S + 5], phi_bulk = p[3 * P.S + 6];
    const int mid = P.mesh_id ? P.mesh_id[cell] : -1;
    const double* xi = mid < 0 ? nullptr : P.mesh_xi + (size_t)mid * P.nx_max;
    const double* g = P.g + (size_t)row * P.nx_max;
    double* v = P.v + (size_t)row * P.nx_max;
    double carry = phi_bulk;
    for (int top = n - 2; top >= 1; top -= 32) {
        const int i = top - lane;
        double t = 0.0;
        if (i >= 1) t = g[i] * (xi ? dx * (xi[i + 1] - xi[i]) : dx);
        const double s = wscan(t, lane);
        if (i >= 1) v[i] = carry - s;
        carry -= __shfl_sync(FULL, s, 31);
    }
    __syncwarp();
    if (lane == 0) {
        v[n - 1] = phi_bulk;
        const double ratio = xi ? (xi[1] - xi[0]) / (xi[2] - xi[1]) : 1.0;
        v[0] = v[1] + (v[1] - v[2]) * ratio;
    }
}

inline int launch_bulk_potential(BulkPotentialParams& P, cudaStream_t st) {
    const unsigned grid = (unsigned)((P.rows + 3) / 4);
    pnp_bulk_potential_kernel<<<grid, 128, 0, st>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : CATINT_PNP_ECUDA;
}

inline int launch_potential(PotentialParams& P, cudaStream_t st) {
    int dev = 0, max_optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return CATINT_PNP_ECUDA;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    P.np = P.tb.nx_max | 1;
    const size_t per_cell = ((sizeof(ExplicitCell) + 15) & ~size_t(15)) + (size_t)(P.tb.S + 3) * P.np * sizeof(double);
    const size_t smem = per_cell * 4;
    if (smem > (size_t)max_optin) return CATINT_PNP_EINVAL;
    cudaFuncSetAttribute(pnp_potential_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const unsigned grid = (unsigned)((P.n_cells + 3) / 4);
    pnp_potential_kernel<<<grid, 128, smem, st>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : CATINT_PNP_ECUDA;
}

}  // namespace catint
