"""The HOST side of the GPU parity tests, on the CPU: the bodies of selected tests of
tests/test_gpu_parity.py run with the C ABI served by tests/cabi_emulator.py (oracle
kernels / NumPy on host memory).  What is under test here is everything ABOVE the C ABI --
the wrapper classes, the mixins' call sequences, argument marshalling, workspace and
sort-state bookkeeping, the one-pass particle side, the fused damping, the moving window,
the laser initialiser -- NOT the CUDA kernels (they are covered by `-m gpu`, through the
same bodies).  A regression in the Python layer shows up here without a GPU."""
import pytest

import cabi_emulator as emu
import test_gpu_parity as gpu_tests
import test_gpu_parity_shapes as shape_tests

CASES = [
    ("test_push_and_sort_bit_exact", dict(M=0, fused=False)),
    ("test_push_and_sort_bit_exact", dict(M=1, fused=True)),
    ("test_sort_edge_cases", {}),
    ("test_deposit_and_gather", dict(M=0)),
    ("test_deposit_and_gather", dict(M=1)),
    ("test_spectral_pipeline_against_oracle", dict(M=0)),
    ("test_spectral_pipeline_against_oracle", dict(M=1)),
    ("test_laser_group_velocity", {}),
    ("test_lwfa_moving_window_run", {}),
    ("test_particle_creation_matches_oracle", {}),
    ("test_fused_push_deposit_equals_push_sort_deposit", dict(M=1)),
    ("test_one_pass_particle_side_equals_reference_sequence", dict(M=0)),
    ("test_one_pass_particle_side_equals_reference_sequence", dict(M=1)),
    ("test_fused_damp_fields_equals_three_calls", dict(M=1, Nx=512)),
    ("test_reference_named_transform_helpers", {}),
    ("test_hermitian_contraction_matches_full", {}),
]


@pytest.mark.filterwarnings("ignore:invalid value encountered in cast")
@pytest.mark.parametrize("name,kwargs", CASES,
                         ids=["%s%s" % (n[5:], "".join("-%s%s" % kv for kv in k.items()))
                              for n, k in CASES])
def test_host_side_of_gpu_test(monkeypatch, name, kwargs):
    from chimeracl_b200 import _lib as real_lib
    emu.patch_cuda_host_calls(monkeypatch)
    comm = emu.EmulatedComm()
    monkeypatch.setattr(real_lib, "_lib", comm.lib)
    getattr(gpu_tests, name)(comm, **kwargs)


# bodies of tests/test_gpu_parity_shapes.py at reduced sizes: the oracle Frame restatement
# against the product's moving window + injector, the cfg3-shape full steps, the uneven-
# filling particle sequence -- host logic only, see the module docstring
SHAPE_CASES = [
    ("test_particle_kernels_long_rows_uneven_filling", dict(M=1, Nx=160, Nr=80, n=60000)),
    ("test_cfg3_shape_two_steps_against_reference_kernels", dict(Nx=128, Nr=40)),
    ("test_cfg1_lwfa_moving_window_against_oracle", dict(Nx=200, Nr=30, checkpoints=(1, 20, 21))),
]


@pytest.mark.filterwarnings("ignore:invalid value encountered in cast")
@pytest.mark.parametrize("name,kwargs", SHAPE_CASES, ids=[n[5:] for n, _ in SHAPE_CASES])
def test_host_side_of_shape_test(monkeypatch, name, kwargs):
    from chimeracl_b200 import _lib as real_lib
    emu.patch_cuda_host_calls(monkeypatch)
    comm = emu.EmulatedComm()
    monkeypatch.setattr(real_lib, "_lib", comm.lib)
    getattr(shape_tests, name)(comm, **kwargs)
