"""Result persistence for the transport model.

Same on-disk artefacts as the reference (one pickle per container inside the
results folder: alldata, species, system, descriptors, xmesh, tmesh,
electrode_reactions, electrolyte_reactions, comsol_outputs;
/root/reference/catint/catint_io.py:23-30,77-131) so that the reference's
post-processing tools keep working on results produced by this backend.
The reference's mpi4py helpers (catint_io.py:154-193) are disabled upstream;
multi-GPU gathering lives in ``catint_b200.distributed``.
"""
import os
import pickle

_CONTAINERS = ('alldata', 'species', 'system', 'descriptors', 'xmesh', 'tmesh',
               'electrode_reactions', 'electrolyte_reactions')


def save_obj(folder, obj, name):
    with open(os.path.join(folder, name + '.pkl'), 'wb') as f:
        pickle.dump(obj, f, pickle.HIGHEST_PROTOCOL)


def load_obj(name, fname):
    with open(os.path.join(fname, name + '.pkl'), 'rb') as f:
        return pickle.load(f)


def save_all(tp, only=None):
    """pickle the result containers of ``tp`` into tp.outputfoldername."""
    if only is not None:
        save_obj(tp.outputfoldername, tp.alldata, only)
        return
    for name in _CONTAINERS:
        save_obj(tp.outputfoldername, getattr(tp, name, None), name)
    save_obj(tp.outputfoldername, tp.comsol_args['outputs'], 'comsol_outputs')


def read_all(tp, fname, only=None):
    """load containers written by ``save_all`` back onto ``tp``."""
    if only is None:
        for name in _CONTAINERS:
            setattr(tp, name, load_obj(name, fname))
        tp.xmax = max(tp.xmesh)
        tp.nx = len(tp.xmesh)
        tp.dx = tp.xmesh[1] - tp.xmesh[0]
        if tp.tmesh is not None and len(tp.tmesh) > 1:
            tp.tmax = max(tp.tmesh)
            tp.nt = len(tp.tmesh)
            tp.dt = tp.tmesh[1] - tp.tmesh[0]
    elif isinstance(only, (list, tuple)):
        for name in only:
            setattr(tp, name, load_obj(name, fname))
    else:
        setattr(tp, only, load_obj(only, fname))
