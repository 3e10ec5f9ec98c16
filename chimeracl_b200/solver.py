"""Solver wrapper class (API of the reference's chimeraCL/solver.py)."""
import numpy as np

from .grid import Grid
from .transformer import Transformer
from .methods.solver_methods_cl import SolverMethodsCL


def psatd_coefficients(A):
    """cos(w dt), w sin(w dt), 1/w^2 per azimuthal mode (reference solver.py:41-52)."""
    dt = A['dt']
    for m in range(A['M'] + 1):
        w = A['w_m' + str(m)]
        ms = '_m' + str(m)
        A['MxSlv_cos(wdt)' + ms] = np.cos(w * dt)
        A['MxSlv_sin(wdt)*w' + ms] = np.sin(w * dt) * w
        A['MxSlv_1/w**2' + ms] = 1. / w ** 2
        A['dont_keep'] += ['MxSlv_cos(wdt)' + ms, 'MxSlv_sin(wdt)*w' + ms, 'MxSlv_1/w**2' + ms]
    return A


class Solver(Grid, Transformer, SolverMethodsCL):
    def __init__(self, configs_in, comm):
        self.import_comm(comm)
        self._process_configs(configs_in)
        self.Args['vec_comps'] = ['x', 'y', 'z']
        self.init_solver_methods()
        self.init_grid_methods()
        self.DataDev = {}
        self._init_grid_data_on_dev()
        self.init_transformer()
        self._make_ms_coefficients()
        self.send_args_to_dev()
        self.pad_operator_matrices()

    def push_fields(self):
        self.advance_fields(vecs=['E', 'G', 'J', 'dN0', 'dN1'])

    def damp_fields(self):
        if self.damp_fields_fused(['E', 'G']):
            return
        self.fb_transform(vects=['E', 'G'], dir=1, mode='half')
        self.profile_edges(['E', 'G'])
        self.fb_transform(vects=['E', 'G'], dir=0, mode='half')

    def restore_B_fb(self):
        self.field_rot('G', 'B')
        self.field_poiss_vec('B')

    def _make_ms_coefficients(self):
        psatd_coefficients(self.Args)
