"""Particles wrapper class (API of the reference's chimeraCL/particles.py)."""
import numpy as np
from scipy.constants import m_e, c, e, epsilon_0

from .methods.generic_methods_cl import ArgsDict
from .methods.particles_methods_cl import ParticleMethodsCL
from .methods.particles_methods_cl import sqrt  # noqa: F401


class Particles(ParticleMethodsCL):
    def __init__(self, configs_in, comm=None):
        if comm is None:
            raise ValueError("Particles needs a Communicator")
        self.import_comm(comm)
        self.DataDev = {}
        self.init_particle_methods()
        self._process_configs(configs_in)
        self.send_args_to_dev()
        self._init_data_on_dev()

    def sort_parts(self, grid):
        if self.Args['Np'] == 0:
            self.flag_sorted = True
        if self.flag_sorted == False:  # noqa: E712
            self.index_sort(grid)
            self.flag_sorted = True

    def add_particles(self, domain_in=None, beam_in=None, source=None):
        if source is not None:
            self.add_new_particles(source=source)
            return
        if domain_in is not None:
            self.make_new_domain(domain_in)
        elif beam_in is not None:
            self.make_new_beam(beam_in)
        self.add_new_particles()

    def align_parts(self):
        if self.Args['Np'] == 0:
            return
        if 'Immobile' in self.Args.keys():
            comps = ['x', 'y', 'z', 'w']
        else:
            comps = ['x', 'y', 'z', 'px', 'py', 'pz', 'g_inv', 'w']
        self.align_and_damp(comps_align=comps)

    def _process_configs(self, configs_in):
        """Defaults and derived constants of reference particles.py:53-100.  The reference
        keeps the caller's dict itself as Args; here Args is an ArgsDict (lazy Np_stay)
        built from it, and the derived keys are written back into the caller's dict once --
        later changes (Np, right_lim, InjectorSource) are visible through parts.Args only."""
        A = ArgsDict(configs_in)
        self._user_configs = configs_in
        A['Np'] = 0
        A['Np_stay'] = 0
        for key, default in (('dt', 1.), ('dx', 1.), ('dr', 1.), ('charge', -1.),
                             ('mass', 1.), ('dens', 1.)):
            if key not in A:
                A[key] = default
        A['dt_2'] = 0.5 * A['dt']
        if 'Nppc' in A:
            A['Nppc'] = np.array(A['Nppc'], dtype=np.uint32)
            A['w0'] = 2 * np.pi * A['dx'] * A['dr'] * A['dens'] / np.prod(A['Nppc'])
            A['ddx'] = A['dx'] / A['Nppc'][0]
        else:
            A['ddx'] = 1.
        A['FactorPush'] = 2 * np.pi * A['dt'] * A['charge'] / A['mass']
        A['right_lim'] = 0.0
        A['w2pC'] = 4 * np.pi ** 2 * m_e * c ** 2 * epsilon_0 * 1e6 / e
        A['dont_send'] = ['InjectorSource', 'charge', 'mass', 'dens', 'Immobile, w2pC']
        A['dont_keep'] = []
        self.flag_sorted = False
        self.Args = A
        try:
            configs_in.update({k: v for k, v in A.items() if k not in configs_in})
        except Exception:
            pass

    def _init_data_on_dev(self):
        for arg in self._attr_names():
            self.DataDev[arg] = self.dev_arr(shape=0, dtype=np.double)
