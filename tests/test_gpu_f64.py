"""GPU parity of the float64 entry points (csrc/abk_f64.cu, cuFFT D2Z) against outputs of the unmodified reference."""

import pytest

import f64_checks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('pd,gd,wd', [('f8', 'f8', 'f8'), ('f4', 'f8', None), ('f8', 'f4', 'f4'), ('f4', 'f8', 'f8')])
def test_tsc_parallel_float64(pd, gd, wd):
    f64_checks.check_tsc_parallel(pd, gd, wd)


@pytest.mark.parametrize('name', ['auto_w', 'cross'])
def test_calc_power_float64(name):
    f64_checks.check_calc_power(name)
