// pnp_rhs.cuh -- K1: right-hand side of the reference ODE, dc/dt = f(c)
// (/root/reference/catint/calculator_old.py:827-935, Poisson :680-819, rates :159-208).
//
// Pure streaming stencil: 16*S*n algorithmic bytes per cell (read c, write dc/dt), ~1.4 flop/B,
// HBM bound.  Design for B200:
//   * a block of 256 threads works on up to RHS_NODES = 512 nodes at a time, one node per thread and
//     round: either a GROUP of whole small cells (n <= 512: five 101-node cells = 505 of 512 lanes busy,
//     where a warp-per-cell mapping left 21 % of the lanes idle), or one TILE of a large cell (n > 512,
//     e.g. the 1001/5001-node grids of C3/C5), tiles walked from the bulk towards the wall so that the
//     backward cumulative sum of the reference's Poisson solve (:753-759,793-796) is carried from tile to
//     tile and the state is read exactly once;
//   * every thread reads its node record (S contiguous doubles, 16-byte loads: a warp covers 32*8*S
//     contiguous bytes, every sector is used) straight into registers, keeps it there for the stencil and
//     drops a transposed copy [species][node] (odd pitch, conflict-free lane-per-node access) into shared
//     memory for its neighbours and for the mass-action products; the node's charge density comes from the
//     registers for free;
//   * g (the field gradient) by a warp-per-cell shuffle scan (groups) or a block-wide scan with carry (tiles);
//   * stencil + mass-action rates from registers/shared memory, results stored straight from registers with
//     16-byte stores (each thread writes its own contiguous record; the half-filled sectors of one store
//     instruction are completed by the next one and merge in L2: DRAM traffic = algorithmic bytes);
//   * the reaction table is pre-compiled per block into byte offsets into the transposed state, so a
//     product of concentrations costs one add, one LDS and one multiply per factor, no predication.
// Grid = resident blocks x SMs, persistent loop over groups / cells.
#pragma once
#include "pnp_kernels.cuh"

namespace catint {

struct RhsParams {
    DevTables tb;
    const double* par; const int* nx; const int* mesh_id; const double* mesh_xi;
    const double* c; long long n_cells;
    double* dcdt; double* g_out; double* phi_out;
};

constexpr int RHS_THREADS = 256;
constexpr int RHS_WARPS = RHS_THREADS / 32;
constexpr int RHS_NODES = 512;                 // node slots per block iteration (two rounds of 256)
constexpr int RHS_ROUNDS = RHS_NODES / RHS_THREADS;
constexpr int RHS_NP = RHS_NODES + 3;          // odd pitch of the transposed state (slot = node + 1)
constexpr int RHS_MAXG = 16;                   // cells per group (tiny grids use fewer slots)
#ifndef CATINT_RHS_CONST_TABLES
#define CATINT_RHS_CONST_TABLES 1              // reaction tables read from the kernel parameters instead of shared memory
#endif
#ifndef CATINT_RHS_MINB
#define CATINT_RHS_MINB 3                      // resident blocks per SM the register allocation aims at
#endif

// per-cell constants of the cells of a group (shared memory)
struct RhsCell {
    double D[MAXS], J[MAXS];
    double bF, Feps;            // beta*F (0 without migration), F/eps
    double dx, g_bulk, phi_wall;
    double u_am, u_ac;          // uniform mesh: 1/dx^2, 1/(2dx) (the divisions are done once per cell)
    const double* xi;           // normalised mesh row or nullptr (uniform)
    const double* fpar;         // flux-equation parameters of the cell or nullptr
    int n, first;               // nodes (0: cell skipped, bad nx), first node slot of this cell in the group
};

// reaction program: byte offsets of the factors into the transposed state
struct RhsProg {
    int ne[MAXR], np[MAXR];
    int off[MAXR][2 * MAXRT];   // educts then products
};

template <int S>
__host__ __device__ inline size_t rhs_smem_bytes() {
    size_t b = (sizeof(DevTables) + 15) & ~size_t(15);     // upper bound (the kernel copies tables_prefix_bytes)
    b += (sizeof(RhsProg) + 15) & ~size_t(15);
    b += sizeof(RhsCell) * RHS_MAXG;
    b += sizeof(double) * ((size_t)S * RHS_NP + 2 * (RHS_NODES + 4) + 2 * RHS_WARPS + 8);
    return b;
}

__device__ __forceinline__ double mesh_h(const RhsCell& ce, int i) {       // h_i = x_{i+1} - x_i
    return ce.xi ? ce.dx * (ce.xi[i + 1] - ce.xi[i]) : ce.dx;
}

// inclusive warp scan (sum) by shuffles
__device__ __forceinline__ double warp_scan(double t, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double v = __shfl_up_sync(FULL, t, o);
        if (lane >= o) t += v;
    }
    return t;
}

// potential by the reference's forward cumulative sum (:798-800): v_0 = phi_wall, v_i = v_{i-1} + g_i*h_{i-1}
// for i = 1..n-2, linear extrapolation to the bulk node.  One warp; g indexed by node.
__device__ __forceinline__ void potential_scan(const RhsCell& ce, const double* g, double* po, int lane) {
    const int n = ce.n;
    double run = ce.phi_wall, vm1 = run, vm2 = run;
    for (int bot = 1; bot <= n - 2; bot += 32) {
        const int i = bot + lane;
        const double t = warp_scan(i <= n - 2 ? g[i] * mesh_h(ce, i - 1) : 0.0, lane);
        const double v = run + t;
        if (i <= n - 2) po[i] = v;
        const int last = (n - 2 - bot) < 31 ? (n - 2 - bot) : 31;
        const double vl = __shfl_sync(FULL, v, last);
        const double vprev = __shfl_sync(FULL, v, last > 0 ? last - 1 : 0);
        vm2 = last > 0 ? vprev : run;
        vm1 = vl;
        run = vl;
    }
    if (lane == 0) {
        po[0] = ce.phi_wall;
        const double ratio = ce.xi ? mesh_h(ce, n - 2) / mesh_h(ce, n - 3) : 1.0;
        po[n - 1] = vm1 + (vm1 - vm2) * ratio;
    }
}

template <int S, bool LARGE>
__global__ void __launch_bounds__(RHS_THREADS, CATINT_RHS_MINB) pnp_rhs_kernel(RhsParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    DevTables* tb = reinterpret_cast<DevTables*>(smem_raw);
    {
        const int words = (int)(sizeof(DevTables) / 4);
        const int* src = reinterpret_cast<const int*>(&P.tb);
        int* dst = reinterpret_cast<int*>(tb);
        for (int w = tid; w < words; w += RHS_THREADS) dst[w] = src[w];
    }
    size_t off = (sizeof(DevTables) + 15) & ~size_t(15);
    RhsProg* prog = reinterpret_cast<RhsProg*>(smem_raw + off);   off += (sizeof(RhsProg) + 15) & ~size_t(15);
    RhsCell* cells_all = reinterpret_cast<RhsCell*>(smem_raw + off);  off += sizeof(RhsCell) * RHS_MAXG;
    double* cs_ = reinterpret_cast<double*>(smem_raw + off);      // [S][RHS_NP] transposed state, slot = node + 1
    double* term = cs_ + (size_t)S * RHS_NP;                      // [RHS_NODES + 4] (F/eps)*sum z c * h_i per slot
    double* gs = term + RHS_NODES + 4;                            // [RHS_NODES + 4] g per slot
    double* wsum = gs + RHS_NODES + 4;                            // [2*RHS_WARPS] block scan partials
    double* carry_s = wsum + 2 * RHS_WARPS;                       // [8] carry between tiles
    __syncthreads();
    if (tid < tb->R) {
        const int r = tid;
        prog->ne[r] = tb->ned[r]; prog->np[r] = tb->npr[r];
        for (int e = 0; e < MAXRT; ++e) {
            prog->off[r][e] = (e < tb->ned[r] ? (int)tb->ed[r][e] : 0) * RHS_NP * 8;
            prog->off[r][MAXRT + e] = (e < tb->npr[r] ? (int)tb->pr[r][e] : 0) * RHS_NP * 8;
        }
    }
    __syncthreads();
    const int nxm = P.tb.nx_max;
    const bool mig = tb->use_migration;
#ifdef CATINT_RHS_NOREACT
    const int R = 0;
#else
    const int R = tb->R;
#endif

    // one node: stencil + rates -> global.  c0: this node's record (registers); slot: its slot in the
    // transposed state (neighbours at slot-1 / slot+1); i: node index in its cell.
    auto node_rhs = [&](const RhsCell& ce, const double (&c0)[S], int slot, int i, double* __restrict__ o) {
        double res[S];
        const int n = ce.n;
        if (i == n - 1) {
#pragma unroll
            for (int s = 0; s < S; ++s) res[s] = 0.0;                                    // frozen bulk node (:886)
        } else if (i == 0) {
            double w0, ih0;
            if (ce.xi) { const double h0 = mesh_h(ce, 0), h1 = mesh_h(ce, 1); w0 = 1.0 / (h0 + h1); ih0 = 1.0 / h0; }
            else { w0 = ce.u_ac; ih0 = 2.0 * ce.u_ac; }
            const double g1 = gs[slot + 1];
#pragma unroll
            for (int s = 0; s < S; ++s)                                                  // :902-909, inward flux
                res[s] = (ce.D[s] * ((cs_[s * RHS_NP + slot + 2] - c0[s]) * w0
                                     + ce.bF * tb->z[s] * cs_[s * RHS_NP + slot + 1] * g1) + ce.J[s]) * ih0;
        } else {
            double am, ap, ac;
            if (ce.xi) {
                const double hm = mesh_h(ce, i - 1), hp = mesh_h(ce, i);
                am = 2.0 / (hm * (hm + hp)); ap = 2.0 / (hp * (hm + hp)); ac = 1.0 / (hm + hp);
            } else {
                am = ap = ce.u_am; ac = ce.u_ac;
            }
            const double gm = gs[slot - 1], gp = gs[slot + 1];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const double cm = cs_[s * RHS_NP + slot - 1], cp = cs_[s * RHS_NP + slot + 1];
                double d2;
                if (!ce.xi) d2 = (cp - 2.0 * c0[s] + cm) * am;                           // :890 (reciprocal hoisted)
                else d2 = am * cm - (am + ap) * c0[s] + ap * cp;
                const double dcg = (cp * gp - cm * gm) * ac;                             // :892
                res[s] = ce.D[s] * (d2 + ce.bF * (CATINT_RHS_CONST_TABLES ? P.tb.z[s] : tb->z[s]) * dcg);
            }
            // mass-action rates (:159-208): factors by pre-compiled byte offsets into the transposed state
            const unsigned char* mine = reinterpret_cast<const unsigned char*>(cs_ + slot);
#pragma unroll 1
            for (int rr = 0; rr < R; ++rr) {
#if CATINT_RHS_CONST_TABLES
                // rate constants, stoichiometry and factor offsets straight from the kernel parameters (constant
                // bank, warp-uniform index): no shared-memory wavefronts for data every lane reads alike
                double f = P.tb.kf[rr], b = P.tb.kr[rr];
                const int ne = P.tb.ned[rr], np = P.tb.npr[rr];
#pragma unroll 1
                for (int e = 0; e < ne; ++e) f *= *reinterpret_cast<const double*>(mine + (int)P.tb.ed[rr][e] * (RHS_NP * 8));
#pragma unroll 1
                for (int e = 0; e < np; ++e) b *= *reinterpret_cast<const double*>(mine + (int)P.tb.pr[rr][e] * (RHS_NP * 8));
                const double net = f - b;
#pragma unroll
                for (int s = 0; s < S; ++s) res[s] = fma(P.tb.nu[rr][s], net, res[s]);        // :920-927
#else
                double f = tb->kf[rr], b = tb->kr[rr];
                const int ne = prog->ne[rr], np = prog->np[rr];
                const int* po = prog->off[rr];
#pragma unroll 1
                for (int e = 0; e < ne; ++e) f *= *reinterpret_cast<const double*>(mine + po[e]);
#pragma unroll 1
                for (int e = 0; e < np; ++e) b *= *reinterpret_cast<const double*>(mine + po[MAXRT + e]);
                const double net = f - b;
                const double2* nur = reinterpret_cast<const double2*>(tb->nu[rr]);   // 16-byte aligned rows (MAXS even)
#pragma unroll
                for (int s = 0; s < S; s += 2) {                                         // :920-927
                    const double2 nn = nur[s / 2];
                    res[s] = fma(nn.x, net, res[s]);
                    if (s + 1 < S) res[s + 1] = fma(nn.y, net, res[s + 1]);
                }
#endif
            }
        }
        if ((S % 2 == 0) && ((((size_t)o) & 15) == 0)) {
#pragma unroll
            for (int s = 0; s + 1 < S; s += 2) *reinterpret_cast<double2*>(o + s) = make_double2(res[s], res[s + 1]);
        } else {
#pragma unroll
            for (int s = 0; s < S; ++s) o[s] = res[s];
        }
    };

    // node record -> registers
    auto fetch_node = [&](const double* __restrict__ src, double (&c0)[S]) {
        if ((S % 2 == 0) && ((((size_t)src) & 15) == 0)) {
#pragma unroll
            for (int s = 0; s + 1 < S; s += 2) {
                const double2 v = *reinterpret_cast<const double2*>(src + s);
                c0[s] = v.x; c0[s + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int s = 0; s < S; ++s) c0[s] = src[s];
        }
    };
    // registers -> transposed copy + charge term
    auto publish_node = [&](const RhsCell& ce, int slot, int i, const double (&c0)[S]) {
        double rho = 0.0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            cs_[s * RHS_NP + slot] = c0[s];
            rho = fma(tb->z[s], c0[s], rho);                                             // :767-771
        }
        // g_i = g_{i+1} + (F/eps)*sum z c * h_i for i = n-2..1; nodes 0 and n-1 add nothing
        term[slot] = (mig && i >= 1 && i <= ce.n - 2) ? rho * ce.Feps * mesh_h(ce, i) : 0.0;
    };

    auto load_node = [&](const RhsCell& ce, const double* __restrict__ src, int slot, int i, double (&c0)[S]) {
        fetch_node(src, c0);
        publish_node(ce, slot, i, c0);
    };

    auto fill_cell = [&](RhsCell& ce, long long cell, int first) {
        const int NPAR = 3 * S + 8;
        const double* p = P.par + (size_t)cell * NPAR;
        const int n = P.nx[cell];
        ce.n = (n < 4 || n > nxm) ? 0 : n;                   // bad nx[cell]: the cell is skipped, outputs untouched
        ce.first = first;
        const int mid = P.mesh_id ? P.mesh_id[cell] : -1;
        ce.xi = mid < 0 ? nullptr : P.mesh_xi + (size_t)mid * nxm;
        ce.fpar = (tb->fq.n_eq > 0 && tb->fpar) ? tb->fpar + (size_t)cell * tb->fq.n_par : nullptr;
        const double beta = p[3 * S + 0], eps = p[3 * S + 1];
        ce.bF = mig ? beta * UNIT_F : 0.0;
        ce.Feps = UNIT_F / eps;
        ce.phi_wall = p[3 * S + 2];
        ce.g_bulk = p[3 * S + 3];
        ce.dx = p[3 * S + 5];
        ce.u_am = 1.0 / (ce.dx * ce.dx);
        ce.u_ac = 1.0 / (2.0 * ce.dx);
        for (int s = 0; s < S; ++s) { ce.J[s] = p[S + s]; ce.D[s] = p[2 * S + s]; }
    };

    if constexpr (!LARGE) {
        // ------------------------------------------------------------------ groups of whole cells
        // (A software-pipelined variant -- node records of group g+1 in flight into registers while group g is
        // computed -- was measured slower: 64 more live registers per thread spill at 128 registers / thread.)
        RhsCell* cells = cells_all;
        int G = RHS_NODES / nxm;
        if (G > RHS_MAXG) G = RHS_MAXG;
        const long long n_groups = (P.n_cells + G - 1) / G;
        for (long long grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
            const long long cell0 = grp * G;
            const int gc = (int)((P.n_cells - cell0) < G ? (P.n_cells - cell0) : G);
            if (tid < gc) fill_cell(cells[tid], cell0 + tid, tid * nxm);
            __syncthreads();
            double c0[RHS_ROUNDS][S];
            int ci[RHS_ROUNDS], ii[RHS_ROUNDS];
#pragma unroll
            for (int rd = 0; rd < RHS_ROUNDS; ++rd) {
                const int j = tid + rd * RHS_THREADS;
                const int cg = j / nxm;
                const int i = j - cg * nxm;
                const bool valid = cg < gc && i < cells[cg < gc ? cg : 0].n;
                ci[rd] = valid ? cg : -1; ii[rd] = i;
                if (valid) load_node(cells[cg], P.c + ((size_t)(cell0 + cg) * nxm + i) * S, j + 1, i, c0[rd]);
            }
            __syncthreads();
            // g per cell: warp-per-cell suffix scan from the bulk (:753-759,793-796)
            for (int cg = warp; cg < gc; cg += RHS_WARPS) {
                const RhsCell& ce = cells[cg];
                const int n = ce.n;
                if (n == 0) continue;
                double* gc_ = gs + ce.first + 1;                    // g of this cell, indexed by node
                const double* tc_ = term + ce.first + 1;
                // four consecutive nodes per lane (serial partial sums) + one warp scan per 128 nodes: a 101-node
                // cell needs one round of five shuffle steps instead of four rounds
                double carry = 0.0;
                for (int top = n - 2; top >= 1; top -= 128) {
                    const int i0 = top - 4 * lane;
                    double t[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) t[q] = (i0 - q >= 1) ? tc_[i0 - q] : 0.0;
                    t[1] += t[0]; t[2] += t[1]; t[3] += t[2];
                    const double incl = warp_scan(t[3], lane);
                    const double base = carry + (incl - t[3]);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (i0 - q >= 1) gc_[i0 - q] = mig ? ce.g_bulk + (base + t[q]) : 0.0;
                    carry += __shfl_sync(FULL, incl, 31);
                }
                if (lane == 0) gc_[n - 1] = mig ? ce.g_bulk : 0.0;
                __syncwarp();
                if (lane == 0) {
                    const double ext = ce.xi ? mesh_h(ce, 0) / mesh_h(ce, 1) : 1.0;
                    gc_[0] = mig ? gc_[1] + (gc_[1] - gc_[2]) * ext : 0.0;                       // :796
                }
                __syncwarp();
                if (P.g_out) for (int i = lane; i < n; i += 32) P.g_out[(size_t)(cell0 + cg) * nxm + i] = gc_[i];
                if (P.phi_out) potential_scan(ce, gc_, P.phi_out + (size_t)(cell0 + cg) * nxm, lane);
            }
            __syncthreads();
#pragma unroll
            for (int rd = 0; rd < RHS_ROUNDS; ++rd) {
                if (ci[rd] >= 0) {
                    const int j = tid + rd * RHS_THREADS;
                    node_rhs(cells[ci[rd]], c0[rd], j + 1, ii[rd], P.dcdt + ((size_t)(cell0 + ci[rd]) * nxm + ii[rd]) * S);
                }
            }
            __syncthreads();
        }
    } else {
        // ------------------------------------------------------------------ tiles of one large cell
        constexpr int TILE = RHS_NODES - 2;                       // nodes computed per tile (+ one halo node each side)
        for (long long cell = blockIdx.x; cell < P.n_cells; cell += gridDim.x) {
            __syncthreads();
            if (tid == 0) fill_cell(cells_all[0], cell, 0);
            __syncthreads();
            const RhsCell& ce = cells_all[0];
            const int n = ce.n;
            if (n == 0) continue;                                 // block-uniform
            const double* csrc = P.c + (size_t)cell * nxm * S;
            double* cdst = P.dcdt + (size_t)cell * nxm * S;
            const int n_tiles = (n + TILE - 1) / TILE;
            if (tid == 0) carry_s[0] = mig ? ce.g_bulk : 0.0;
            // tiles from the bulk end to the wall; tile t covers nodes [a, b]; slot of node i: i - a + 1
            for (int t = n_tiles - 1; t >= 0; --t) {
                const int a = t * TILE;
                const int b = (a + TILE - 1 < n - 1) ? a + TILE - 1 : n - 1;
                const int lo = a > 0 ? a - 1 : 0, hi = b < n - 1 ? b + 1 : n - 1;     // with halo
                __syncthreads();
                double c0[RHS_ROUNDS][S];
                int ii[RHS_ROUNDS];
#pragma unroll
                for (int rd = 0; rd < RHS_ROUNDS; ++rd) {
                    const int i = lo + tid + rd * RHS_THREADS;
                    ii[rd] = i <= hi ? i : -1;
                    if (i <= hi) load_node(ce, csrc + (size_t)i * S, i - a + 1, i, c0[rd]);
                }
                __syncthreads();
                // block-wide suffix scan over the nodes hi..lo (reversed index r = hi - i): the carry is g at
                // node hi (computed by the previous tile, or g_bulk at the bulk end), so node hi adds nothing
                {
                    const int M = hi - lo + 1;
                    const int r0 = 2 * tid, r1 = 2 * tid + 1;
                    const double t0 = (r0 < M && r0 > 0) ? term[hi - r0 - a + 1] : 0.0;
                    const double t1 = r1 < M ? term[hi - r1 - a + 1] : 0.0;
                    const double s2 = warp_scan(t0 + t1, lane);
                    if (lane == 31) wsum[warp] = s2;
                    __syncthreads();
                    double pre = 0.0;
                    for (int w = 0; w < warp; ++w) pre += wsum[w];
                    const double base = carry_s[0];
                    const double incl1 = pre + s2, incl0 = incl1 - t1;
                    if (r0 < M) gs[hi - r0 - a + 1] = mig ? base + incl0 : 0.0;
                    if (r1 < M) gs[hi - r1 - a + 1] = mig ? base + incl1 : 0.0;
                    __syncthreads();
                    if (tid == 0) {
                        if (a == 0) {
                            const double ext = ce.xi ? mesh_h(ce, 0) / mesh_h(ce, 1) : 1.0;
                            gs[1] = mig ? gs[2] + (gs[2] - gs[3]) * ext : 0.0;                  // g_0 (:796)
                        }
                        carry_s[0] = gs[1];                        // g at node a = the next tile's node hi
                    }
                    __syncthreads();
                }
                if (P.g_out) for (int i = a + tid; i <= b; i += RHS_THREADS) P.g_out[(size_t)cell * nxm + i] = gs[i - a + 1];
#pragma unroll
                for (int rd = 0; rd < RHS_ROUNDS; ++rd) {
                    const int i = ii[rd];
                    if (i >= a && i <= b) node_rhs(ce, c0[rd], i - a + 1, i, cdst + (size_t)i * S);
                }
            }
            __syncthreads();
            if (P.phi_out && P.g_out && warp == 0) {
                // potential from the g this block has just written (needs g_out): one warp, scan from the wall
                potential_scan(ce, P.g_out + (size_t)cell * nxm, P.phi_out + (size_t)cell * nxm, lane);
            }
        }
    }
}

// Flux equations (pnp_fluxeq.cuh): the wall rows get ih0 * sum_e coef[.][e]*E_e(c(0), phi_wall) on top of the
// fixed flux.  A separate one-thread-per-cell kernel after the streaming kernel: inside it, the call into the
// expression interpreter cost every launch 15 % (register pressure around the call), with or without flux equations.
__global__ void __launch_bounds__(128) pnp_rhs_fluxeq_kernel(RhsParams P) {
    const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= P.n_cells) return;
    const DevTables& tb = P.tb;
    const int S = tb.S, nxm = tb.nx_max;
    const int n = P.nx[cell];
    if (n < 4 || n > nxm) return;
    const double* p = P.par + (size_t)cell * (3 * S + 8);
    const int mid = P.mesh_id ? P.mesh_id[cell] : -1;
    const double dx = p[3 * S + 5];
    const double h0 = mid < 0 ? dx : dx * (P.mesh_xi[(size_t)mid * nxm + 1] - P.mesh_xi[(size_t)mid * nxm]);
    const double ih0 = 1.0 / h0;
    const double* c0 = P.c + (size_t)cell * nxm * S;
    double* out = P.dcdt + (size_t)cell * nxm * S;
    const double* fpar = tb.fpar ? tb.fpar + (size_t)cell * tb.fq.n_par : nullptr;
    for (int e = 0; e < tb.fq.n_eq; ++e) {
        const double E = fluxeq_eval(&tb.fq, e, fpar, c0, p[3 * S + 2], -1, S, nullptr) * ih0;
        for (int s = 0; s < S; ++s) out[s] = fma(tb.fq.coef[s][e], E, out[s]);
    }
}

template <int S>
int launch_rhs(RhsParams& P, cudaStream_t st) {
    const size_t smem = rhs_smem_bytes<S>();
    // per-device launch geometry, queried on every call (a few microseconds; no static state to race on,
    // correct when one process drives several GPUs)
    int dev = 0, sms = 148, max_optin = 0, per_sm = 1;
    if (cudaGetDevice(&dev) != cudaSuccess) return CATINT_PNP_ECUDA;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (smem > (size_t)max_optin) return CATINT_PNP_EINVAL;
    const bool large = P.tb.nx_max > RHS_NODES;
    long long want;
    if (large) {
        cudaFuncSetAttribute(pnp_rhs_kernel<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pnp_rhs_kernel<S, true>, RHS_THREADS, smem);
        want = P.n_cells;
    } else {
        cudaFuncSetAttribute(pnp_rhs_kernel<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pnp_rhs_kernel<S, false>, RHS_THREADS, smem);
        int G = RHS_NODES / P.tb.nx_max;
        if (G > RHS_MAXG) G = RHS_MAXG;
        want = (P.n_cells + G - 1) / G;
    }
    if (per_sm < 1) per_sm = 1;
    const long long cap = (long long)sms * per_sm;            // one resident wave, persistent loop inside
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (large) pnp_rhs_kernel<S, true><<<grid, RHS_THREADS, smem, st>>>(P);
    else pnp_rhs_kernel<S, false><<<grid, RHS_THREADS, smem, st>>>(P);
    if (P.tb.fq.n_eq > 0) pnp_rhs_fluxeq_kernel<<<(unsigned)((P.n_cells + 127) / 128), 128, 0, st>>>(P);
    return cudaGetLastError() == cudaSuccess ? 0 : CATINT_PNP_ECUDA;
}

}  // namespace catint
