"""Golden for the 1001-node uniform grid (C1 chemistry, nx=1000, dx = 50 nm): the state does not fit in
shared memory, the GPU takes the global-state kernel variant.  Dense odeint is impractical at N = 8008, so the
checker is the CPU BDF of oracle/bdf_local.py (pinned against scipy odeint at 101 nodes: tests/test_oracle.py
and make_c3_golden.py 101) + the Newton root of the same discrete residual.

    OMP_NUM_THREADS=1 python tests/golden/make_n1001_golden.py

Transient outputs and their CONDITIONING.  On its way to the steady state the reference's discretisation
passes through a phase in which a grid-scale mode grows from rounding noise and saturates.  The
time at which that happens depends on the size of the seed, i.e. on rounding: two runs of the SAME integrator
whose initial states differ by 1e-13 relative already differ by percents inside that window, while they agree
before it and after it.  The script therefore integrates twice (bdf_c and bdf_c_perturbed, initial
concentrations multiplied by 1 + 1e-13*N(0,1)) and stores both, so that the test can use the measured
sensitivity as its tolerance at every output time instead of a guessed one.
"""
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
os.environ.setdefault('CATINT_QUIET', '1')

T_OUT = [1e-3, 1e-2, 1.0, 200.0]


def main():
    from catint_b200.transport import Transport
    from catint_b200.calculator import build_cell_batch
    from catint_b200 import workloads
    from oracle.fixtures import system_from_batch
    from oracle.pnp_local import LocalForm
    from oracle.bdf_local import BdfIntegrator
    tp = Transport(resultsdir=tempfile.mkdtemp(), **workloads.co2r_inputs(nx=1000))
    batch, _ = build_cell_batch(tp)
    s = system_from_batch(batch, 0)
    lf = LocalForm(s)
    runs = []
    rng = np.random.default_rng(0)
    for pert in (0.0, 1e-13):
        y0 = lf.y_bulk()
        if pert:
            C = y0[:, :s.S].T * (1.0 + pert * rng.standard_normal((s.S, s.n)))
            C[:, -1] = s.c_bulk                      # the bulk node stays on its Dirichlet value
            y0 = lf.y_from_c(C)
        t0 = time.time()
        integ = BdfIntegrator(lf, fresh_jacobian=False)
        outs = integ.integrate(T_OUT, y0=y0, max_steps=400000)
        runs.append(np.stack([o[:, :s.S] for o in outs]))
        print('perturbation %.0e: %d steps, %.0f s' % (pert, integ.stats.nst, time.time() - t0), flush=True)
        y_end = outs[-1]
    y, info = lf.solve_steady(y0=y_end, pure_newton=True)
    assert info['converged']
    C, v, g = lf.unpack(y)
    cs = np.max(np.abs(s.c_bulk))
    sens = [float(np.max(np.abs(runs[0][k] - runs[1][k]) / (np.abs(runs[0][k]) + 1e-9 * cs))) for k in range(len(T_OUT))]
    print('relative difference of the two runs at t =', T_OUT, ':', sens)
    np.savez_compressed(os.path.join(HERE, 'oracle_n1001.npz'), par=batch.par, nx=batch.nx, t_out=np.array(T_OUT),
                        bdf_c=runs[0], bdf_c_perturbed=runs[1], sensitivity=np.array(sens),
                        newton_c=C.T.copy(), phi=v, g=g)


if __name__ == '__main__':
    main()
