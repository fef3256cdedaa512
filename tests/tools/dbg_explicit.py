import os, sys
HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE))); sys.path.insert(0, os.path.dirname(HERE))
import numpy as np, warnings
warnings.simplefilter('ignore')
from conftest import load_golden
from test_explicit import model_from_setup, batch_with_pair, PAIRS
from catint_b200 import backend as be
from oracle import explicit_oracle as eo
bk = be.PnpBackend('cuda:0')
su = load_golden('ref_c1.npz'); S, n = len(su['z']), int(su['nx'])
for stepper, lf in (('ftcs', False), ('ftcs', True), ('cn', False), ('cn', True)):
    batch = batch_with_pair(su, 0, B=1); db = bk.upload(batch)
    nt = 6; itout = list(range(0 if stepper == 'ftcs' else 1, nt))
    out = bk.step(db, be.STEPPER_FTCS if stepper == 'ftcs' else be.STEPPER_CRANK_NICOLSON, 2e-11, nt, itout, lax_friedrich=lf)
    m = model_from_setup(su, PAIRS[0])
    ref = (eo.integrate_ftcs if stepper == 'ftcs' else eo.integrate_crank_nicolson)(m, 2e-11, nt, itout, lax_friedrich=lf)
    gc = out['c'].cpu().numpy(); gv = out['phi'].cpu().numpy(); gg = out['g'].cpu().numpy()
    for k, (C, v, g) in enumerate(ref):
        e = np.abs(gc[k, 0, :n].T - C)
        kk, ii = np.unravel_index(np.argmax(e), e.shape)
        print(stepper, lf, 'step', itout[k], 'max|dC| %.3e at species %d node %d (C=%.6g, max|C| %.4g)' % (e.max(), kk, ii, C[kk, ii], np.abs(C).max()),
              'dv1 %.3e dg1 %.3e g1 %.4g' % (abs(gv[k, 0, 1] - v[1]), abs(gg[k, 0, 1] - g[1]), g[1]))
