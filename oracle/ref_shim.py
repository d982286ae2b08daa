"""
Import the UNMODIFIED reference (abacusnbody.analysis.{tsc,power_spectrum}) from /root/reference.

/root/reference does not exist on the GPU box; there the modules staged by ``oracle/make_ref.py``
under ``oracle/_ref/`` (git-ignored build output, byte-for-byte copies) are used instead.  Callers:
``tests/golden/make_golden.py`` (generates the committed fixtures), the ``not gpu`` oracle-vs-reference
tests (skip when no reference is available) and ``bench.py --impl reference`` / the ``cpu_baseline`` leg,
which time the reference's own Numba path on the box's host cores.  Nothing under ``-m gpu`` tests or
``smoke()`` calls this.

Two shims, neither of which touches the reference sources:
  * ``abacusnbody/__init__.py`` imports a setuptools_scm-generated ``version`` module that does not
    exist in a source checkout -> register bare package modules with the right ``__path__``.
  * ``power_spectrum.py:11`` imports ``astropy.table.Table`` (not installed) -> a dict-with-meta stub.

Oracle hygiene: the reference's TSC has a lost-update race when it picks ``npartition == n//2``
(``2*threads >= n//2``), and the interlaced paints ignore ``nthread`` and use the import-time
``NUMBA_NUM_THREADS``; so the thread count is fixed through the environment BEFORE numba loads.
"""

import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')   # oracle/make_ref.py: travels to the GPU box


def _pick_root():
    env = os.environ.get('ABK_REFERENCE_ROOT')
    for root in ([env] if env else []) + ['/root/reference', _STAGED]:
        if os.path.isfile(os.path.join(root, 'abacusnbody', 'analysis', 'tsc.py')):
            return root
    return env or '/root/reference'


REF_ROOT = _pick_root()


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'abacusnbody', 'analysis', 'tsc.py'))


def load(num_threads=4):
    """Return (tsc_module, power_spectrum_module) of the reference."""
    if not available():
        raise RuntimeError(f'reference tree not found under {REF_ROOT}')
    if 'numba' not in sys.modules:
        os.environ.setdefault('NUMBA_NUM_THREADS', str(num_threads))
    if 'abacusnbody.analysis.power_spectrum' in sys.modules:
        return sys.modules['abacusnbody.analysis.tsc'], sys.modules['abacusnbody.analysis.power_spectrum']

    pkg = types.ModuleType('abacusnbody')
    pkg.__path__ = [os.path.join(REF_ROOT, 'abacusnbody')]
    sub = types.ModuleType('abacusnbody.analysis')
    sub.__path__ = [os.path.join(REF_ROOT, 'abacusnbody', 'analysis')]
    sys.modules.setdefault('abacusnbody', pkg)
    sys.modules.setdefault('abacusnbody.analysis', sub)

    if 'astropy' not in sys.modules:
        try:
            import astropy.table  # noqa: F401
        except Exception:
            class Table(dict):
                def __init__(self, d, meta=None):
                    super().__init__(d)
                    self.meta = meta or {}

            ast = types.ModuleType('astropy')
            ast.__path__ = []
            tab = types.ModuleType('astropy.table')
            tab.Table = Table
            ast.table = tab
            sys.modules['astropy'] = ast
            sys.modules['astropy.table'] = tab

    from abacusnbody.analysis import power_spectrum, tsc  # type: ignore

    return tsc, power_spectrum
