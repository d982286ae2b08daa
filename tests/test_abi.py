"""CPU-side checks of the C-ABI boundary: libabk.so builds for sm_100a, loads, and exports exactly
the entry points include/abk.h declares (no compute calls: there is no GPU here)."""

import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope='module')
def lib():
    from abacusutils_b200 import _build, _lib

    _build.build()
    return _lib.load_library()


def declared_functions():
    text = (ROOT / 'include' / 'abk.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(abk_[a-z0-9_]+)\s*\(', text)))


def test_header_and_binding_agree(lib):
    from abacusutils_b200 import _lib

    declared = declared_functions()
    assert len(declared) >= 25
    assert sorted(_lib.SIGNATURES) == declared


def test_every_declared_symbol_is_exported(lib):
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_version_and_error_string(lib):
    assert lib.abk_version() == 100
    assert isinstance(lib.abk_last_error(), bytes)


def test_fft_backend_is_opened_by_path(lib):
    """libabk does not link cuFFT (a process that imported torch would hand it torch's bundled copy): it opens the toolkit's
    library by path; no compute call is made here."""
    import ctypes as C
    import subprocess

    from abacusutils_b200 import _lib

    needed = subprocess.run(['ldd', str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert 'cufft' not in needed
    buf, ver = C.create_string_buffer(512), C.c_int()
    rc = lib.abk_fft_backend(buf, 512, C.byref(ver))
    if rc != 0:
        pytest.skip('no cuFFT on this machine: ' + lib.abk_last_error().decode())
    assert b'libcufft' in buf.value and ver.value >= 11000


def test_struct_layouts_match_header():
    """sizeof of the ctypes mirrors == what the C compiler lays out (checked with a tiny C program)."""
    import ctypes
    import subprocess
    import tempfile

    from abacusutils_b200 import _lib

    src = '#include <stdio.h>\n#include "abk.h"\nint main(){printf("%zu %zu\\n", sizeof(abk_kmesh), sizeof(abk_bin_request));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = Path(d) / 't.c'
        c.write_text(src)
        exe = Path(d) / 't'
        subprocess.run(['/usr/bin/gcc', '-I', str(ROOT / 'include'), str(c), '-o', str(exe)], check=True)
        a, b = map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split())
    assert ctypes.sizeof(_lib.KMesh) == a
    assert ctypes.sizeof(_lib.BinRequest) == b


def test_no_cpu_fallback_without_gpu():
    import numpy as np
    import torch

    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from abacusutils_b200._lib import AbkError
    from abacusutils_b200.analysis.tsc import tsc_parallel

    with pytest.raises(AbkError):
        tsc_parallel(np.zeros((4, 3), dtype=np.float32), 8, 1.0)


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/ (test infrastructure)."""
    for py in (ROOT / 'abacusutils_b200').rglob('*.py'):
        text = py.read_text()
        assert 'import oracle' not in text and 'from oracle' not in text and 'abk_oracle' not in text, py


def test_host_tables_match_reference(golden):
    import numpy as np

    from abacusutils_b200.analysis import power_spectrum as ps

    for n, L, inter in [(8, 100.0, True), (8, 100.0, False), (72, 1000.0, True), (128, 1000.0, False),
                        (250, 2000.0, True)]:
        np.testing.assert_array_equal(ps.get_W_compensated(L, n, 'TSC', inter), golden[f'W/{n}_{int(inter)}'])
        np.testing.assert_array_equal(ps.get_W_compensated(L, n, 'CIC', inter), golden[f'Wcic/{n}_{int(inter)}'])
    with pytest.raises(ValueError):
        ps.get_W_compensated(100.0, 8, 'NGP', True)
    k, mu = ps.get_k_mu_edges(1000.0, 0.5, 10, 4, False)
    np.testing.assert_array_equal(k, np.linspace(0, 0.5, 11))
    np.testing.assert_array_equal(mu, np.linspace(0, 1, 5))
    k, _ = ps.get_k_mu_edges(1000.0, 0.5, 10, 4, True)
    np.testing.assert_array_equal(k, np.geomspace((1 - 1e-4) * 2 * np.pi / 1000.0, 0.5, 11))
    arr = np.array([0.1, 0.2])
    assert ps.get_k_mu_edges(1000.0, 0.5, arr, arr, False)[0] is arr


def test_legendre_coefficients(golden):
    import numpy as np

    import cases
    from abacusutils_b200.analysis.power_spectrum import legendre_coefficients

    co = legendre_coefficients(list(range(11))).astype(np.float64)
    mu = np.sqrt(cases.PN_X.astype(np.float64))
    for ell in range(11):
        val = sum(co[ell, m] * mu**m for m in range(11)) / (2 * ell + 1)
        np.testing.assert_allclose(val, golden[f'P_n/{ell}'], rtol=1e-5, atol=3e-6 * 4**(ell // 2))


def _strip_comments(text):
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return re.sub(r'//[^\n]*', '', text)


def _param_types(sig):
    out = []
    for a in sig.split(','):
        a = a.strip()
        if a in ('', 'void'):
            continue
        a = re.sub(r'\b[A-Za-z_][A-Za-z0-9_]*$', '', a).strip()  # drop the parameter name
        out.append(re.sub(r'\s+', ' ', a).replace(' *', '*').replace('* ', '*'))
    return out


def test_header_prototypes_match_definitions_and_bindings():
    """Every prototype in include/abk.h has a definition with the same parameter types in csrc/*.cu, and the
    ctypes binding passes the same number of arguments."""
    from abacusutils_b200 import _lib

    hdr = _strip_comments((ROOT / 'include' / 'abk.h').read_text())
    protos = {m.group(1): m.group(2) for m in
              re.finditer(r'\b(abk_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;', hdr, flags=re.S)}
    defs = {}
    for f in (ROOT / 'abacusutils_b200' / 'csrc').glob('*.cu'):
        src = _strip_comments(f.read_text())
        for m in re.finditer(r'extern "C"\s+[A-Za-z0-9_ ]+\*?\s*(abk_[a-z0-9_]+)\s*\(([^{;]*?)\)\s*\{', src, flags=re.S):
            defs[m.group(1)] = m.group(2)
    assert set(protos) == set(defs) == set(_lib.SIGNATURES)
    for name, sig in protos.items():
        assert _param_types(sig) == _param_types(defs[name]), name
        assert len(_param_types(sig)) == len(_lib.SIGNATURES[name][1]), name


def test_tile_shape_matches_the_sharded_path_constants():
    """dist.py cuts slab boundaries on tile columns and slices tile-offset tables: its constants must be the library's."""
    import ctypes as C

    from abacusutils_b200 import _lib, dist

    lib = _lib.load_library()
    tx, ty, tz = C.c_int(), C.c_int(), C.c_int()
    assert lib.abk_tsc_tile_shape(C.byref(tx), C.byref(ty), C.byref(tz)) == 0
    assert (tx.value, ty.value, tz.value) == (dist.TILE_X, dist.TILE_Y, dist.TILE_Z)
    nt = C.c_int64()
    assert lib.abk_tsc_num_tiles(100, 50, 61, C.byref(nt)) == 0
    assert nt.value == -(-100 // tx.value) * -(-50 // ty.value) * -(-61 // tz.value)
