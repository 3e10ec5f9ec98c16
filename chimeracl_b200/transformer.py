"""Transformer wrapper class (API of the reference's chimeraCL/transformer.py).

Spectral axes, filters and the DHT / dDHT matrices are one-off host-side NumPy/SciPy
constructions, exactly the formulas of reference transformer.py:28-97 (they are data
of the algorithm, uploaded once)."""
import numpy as np
from scipy.special import jn_zeros, jn

from .methods.transformer_methods_cl import TransformerMethodsCL


def spectral_axes(A):
    """kx, kr_m, w_m, Poisson and smoothing filters (reference transformer.py:28-63)."""
    Nx, Nr, M = A['Nx'], A['Nr'], A['M']
    kx = 2 * np.pi * np.fft.fftfreq(Nx, A['dx'])
    R_period = A['Rgrid'][-1] + 0.5 * A['dr']
    A['kx'] = kx
    for m in range(M + 2):
        A['kr_m' + str(m)] = jn_zeros(m, Nr - 1) / R_period
    sx = 1 - np.sin(0.5 * np.pi * kx[None, :] / kx.max()) ** 2
    for m in range(M + 1):
        kr = A['kr_m' + str(m)]
        w = np.sqrt(kx[None, :] ** 2 + kr[:, None] ** 2)
        A['w_m' + str(m)] = w
        A['Poiss_m' + str(m)] = 1. / w ** 2
        A['SmoothingFilter_m' + str(m)] = sx * (1 - np.sin(0.5 * np.pi * kr[:, None] / kr.max()) ** 2)
    for m in range(M + 2):
        A['dont_send'] += ['kr_m' + str(m), 'w_m' + str(m)]
        A['dont_keep'] += ['Poiss_m' + str(m), 'SmoothingFilter_m' + str(m)]
    return A


def hankel_matrices(A):
    """DHT_inv_m = J_m(r_j k_i), DHT_m = pinv(DHT_inv_m) and the radial-derivative
    operators dDHT_plus/minus_m (reference transformer.py:65-97)."""
    r = A['Rgrid'][1:, None]
    R_period = r[-1] + 0.5 * A['dr']
    n = A['Nr'] - 1
    for m in range(A['M'] + 1):
        k0, kp, km = (jn_zeros(mm, n) / R_period for mm in (m, m + 1, m - 1))
        inv = jn(m, r * k0)
        fwd = np.linalg.pinv(inv)
        A['DHT_inv_m' + str(m)] = inv
        A['DHT_m' + str(m)] = fwd
        A['dDHT_plus_m' + str(m)] = fwd.dot(0.5 * kp * jn(m, r * kp))
        A['dDHT_minus_m' + str(m)] = fwd.dot(0.5 * km * jn(m, r * km))
        A['dont_keep'] += [s + str(m) for s in ('DHT_inv_m', 'DHT_m', 'dDHT_plus_m', 'dDHT_minus_m')]
    return A


class Transformer(TransformerMethodsCL):
    def init_transformer(self):
        self._init_transformer_data_on_dev()
        self.init_transformer_methods()
        self._make_spectral_axes()
        self._make_DHT()

    def fb_transform(self, scals=[], vects=[], dir=0, mode='full', smooth=False, partial=False):
        """Reference transformer.py:17-26.  On a kr-row sharded solver (Solver.
        enable_spectral_sharding) the forward transform fills the owned spectral rows and
        the full backward transform leaves this rank's partial sum in the grid arrays;
        unless partial=True (PIC_loop overlaps the exchange itself) the partials are summed
        over the ranks here, so a direct caller (Diagnostics, a user script) gets complete
        grid fields."""
        comps = list(scals)
        for vect in vects:
            comps += [vect + comp for comp in self.Args['vec_comps']]
        self.transform_fields(comps, dir=dir, mode=mode, smooth=smooth)
        if dir == 1 and mode == 'full' and not partial:
            st = self.__dict__.get('_sharding')
            if st is not None and not st['emulate'] and st['world'] > 1:
                from .parallel import allreduce_each_async
                work = allreduce_each_async(
                    [self.DataDev[c + '_m' + str(m)].t[1:] for c in comps
                     for m in range(self.Args['M'] + 1)], self.comm.process_group)
                if work is not None:
                    work.wait()

    def pad_operator_matrices(self):
        """Re-house the (Nr-1) x (Nr-1) DHT / dDHT matrices in storage with a leading
        dimension rounded up to a multiple of 8 (zeros in the pad) and expose the same
        (Nr-1, Nr-1) view under the same DataDev key: rows become 16-byte aligned, which
        lets the contraction kernel load them 128 bits at a time.  Called after
        send_args_to_dev()."""
        import torch
        from .devarray import DevArray
        K = self.Args['Nr'] - 1
        Kp = (K + 7) // 8 * 8
        for m in range(self.Args['M'] + 1):
            for name in ('DHT_m', 'DHT_inv_m', 'dDHT_plus_m', 'dDHT_minus_m'):
                key = name + str(m)
                if key not in self.DataDev:
                    continue
                store = torch.zeros((K, Kp), dtype=torch.float64, device=self.comm.device)
                store[:, :K] = self.DataDev[key].t
                self.DataDev[key] = DevArray(store[:, :K])

    def _make_spectral_axes(self):
        spectral_axes(self.Args)

    def _make_DHT(self):
        hankel_matrices(self.Args)

    def _init_transformer_data_on_dev(self):
        shape = (self.Args['Nr'] - 1, self.Args['Nx'])
        names = ['rho'] + [f + c for f in ('E', 'B', 'G', 'J', 'dN0', 'dN1')
                           for c in self.Args['vec_comps']]
        for name in names:
            for m in range(self.Args['M'] + 1):
                self.DataDev[name + '_fb_m' + str(m)] = self.dev_arr(
                    val=0, dtype=np.complex128, shape=shape)
        for comp in self.Args['vec_comps']:
            self.DataDev['buff_fb_m-1_' + comp] = self.dev_arr(val=0, dtype=np.complex128,
                                                               shape=shape)
        self.DataDev['phs_shft'] = self.dev_arr(dtype=np.complex128, val=0,
                                                shape=self.Args['Nx'])
        for i in range(2):
            self.DataDev['fld_buff%d_d' % i] = self.dev_arr(val=0, shape=shape, dtype=np.double)
            self.DataDev['fld_buff%d_c' % i] = self.dev_arr(val=0, shape=shape,
                                                            dtype=np.complex128)
