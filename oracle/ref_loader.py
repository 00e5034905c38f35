"""TEST INFRASTRUCTURE. Imports the reference's own Python hot-path modules UNMODIFIED from /root/reference
(present only in the build container; never on the GPU box) so that the oracle restatement can be validated against
them and golden vectors can be generated (tests/golden/make_golden.py).

The reference cannot be imported as is: `import imate` (mixed_correlation.py:17; absent, unpinned
third-party package) and matplotlib/seaborn (_utilities/plot_utilities.py:16-24) are missing. This module installs
  * an `imate` shim restating the two deterministic methods the reference reaches: 'eigenvalue' (reductions over the
    eigenvalues that mixed_correlation.py:76-79 itself computes with scipy.linalg.eigh) and 'cholesky'; every
    function returns the pre-release `(value, info)` tuple the reference unpacks (mixed_correlation.py:109,178,245);
  * empty matplotlib / mpl_toolkits / seaborn shims (plotting is out of scope);
and loads the package under the alias `gpref_py` straight from the reference tree, bypassing gaussian_proc/__init__.py
(whose import guard, :14-65, is a Cython-build artefact) and the compiled generators (those come from oracle/_ref).
"""

import importlib.util
import os
import sys
import types

import numpy
import scipy.linalg

REF = os.environ.get('GP_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REF, 'gaussian_proc', '_likelihood'))


def _imate_shim():
    m = types.ModuleType('imate')

    class AffineMatrixFunction(object):
        def __init__(self, K):
            self.K = K

    def _dense(A):
        return A.toarray() if hasattr(A, 'toarray') else numpy.asarray(A)

    def logdet(A, method='cholesky', eigenvalues=None, exponent=1, **kw):
        if method == 'eigenvalue':
            return exponent * numpy.sum(numpy.log(eigenvalues)), {}
        L = numpy.linalg.cholesky(_dense(A))
        return exponent * 2.0 * numpy.sum(numpy.log(numpy.diag(L))), {}

    def traceinv(A, method='cholesky', eigenvalues=None, exponent=1, **kw):
        if method == 'eigenvalue':
            return numpy.sum(1.0 / eigenvalues ** exponent), {}
        Ai = numpy.linalg.inv(_dense(A))
        return numpy.trace(numpy.linalg.matrix_power(Ai, exponent)), {}

    def trace(A, method='exact', eigenvalues=None, exponent=1, **kw):
        if method == 'eigenvalue':
            return numpy.sum(eigenvalues ** exponent), {}
        Ad = _dense(A)
        return numpy.trace(numpy.linalg.matrix_power(Ad, exponent)), {}

    m.AffineMatrixFunction = AffineMatrixFunction
    m.InterpolateTraceInv = None
    m.logdet, m.traceinv, m.trace = logdet, traceinv, trace
    return m


def _plot_shims():
    class _Any(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith('__'):
                raise AttributeError(name)
            return lambda *a, **k: None

    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.ticker', 'matplotlib.font_manager', 'mpl_toolkits',
                 'mpl_toolkits.mplot3d', 'mpl_toolkits.axes_grid1', 'mpl_toolkits.axes_grid1.inset_locator',
                 'seaborn'):
        if name not in sys.modules:
            mod = _Any(name)
            mod.__path__ = []
            sys.modules[name] = mod
    for name in list(sys.modules):
        if '.' in name and isinstance(sys.modules[name], _Any):
            parent, child = name.rsplit('.', 1)
            if isinstance(sys.modules.get(parent), _Any):
                setattr(sys.modules[parent], child, sys.modules[name])
    sys.modules['matplotlib'].get_backend = lambda: 'agg'
    if 'distutils.spawn' not in sys.modules:
        try:
            import distutils.spawn  # noqa: F401  (setuptools' shim on Python >= 3.12)
        except ImportError:
            ds = types.ModuleType('distutils.spawn')
            ds.find_executable = lambda name: None
            d0 = types.ModuleType('distutils')
            d0.spawn = ds
            sys.modules['distutils'], sys.modules['distutils.spawn'] = d0, ds
    il = sys.modules['mpl_toolkits.axes_grid1.inset_locator']
    il.mark_inset = il.InsetPosition = il.inset_axes = lambda *a, **k: None


_loaded = None


def load():
    """Returns a namespace with the reference's DirectLikelihood, ProfileLikelihood, MixedCorrelation, root finders."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError('reference tree not available at %s' % REF)
    sys.modules.setdefault('imate', _imate_shim())
    _plot_shims()
    root = os.path.join(REF, 'gaussian_proc')
    pkg = types.ModuleType('gpref_py')
    pkg.__path__ = [root]
    sys.modules['gpref_py'] = pkg

    def sub(name, is_pkg=False):
        path = os.path.join(root, *name.split('.'))
        path = os.path.join(path, '__init__.py') if is_pkg else path + '.py'
        full = 'gpref_py.' + name
        spec = importlib.util.spec_from_file_location(
            full, path, submodule_search_locations=[os.path.dirname(path)] if is_pkg else None)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
        return mod

    sub('_utilities', True)
    sub('_mixed_correlation', True)
    sub('_likelihood', True)
    ns = types.SimpleNamespace()
    ns.MixedCorrelation = sys.modules['gpref_py._mixed_correlation'].MixedCorrelation
    lk = sys.modules['gpref_py._likelihood']
    ns.DirectLikelihood = sys.modules['gpref_py._likelihood._direct_likelihood'].DirectLikelihood
    ns.ProfileLikelihood = sys.modules['gpref_py._likelihood._profile_likelihood'].ProfileLikelihood
    ns.Likelihood = lk.Likelihood
    rf = sys.modules['gpref_py._likelihood._root_finding']
    ns.find_interval_with_sign_change = rf.find_interval_with_sign_change
    ns.chandrupatla_method = rf.chandrupatla_method
    _loaded = ns
    return ns


def load_cython():
    """The compiled reference generators from oracle/_ref (built by oracle/build_ref.py); travels to the GPU box."""
    here = os.path.dirname(os.path.abspath(__file__))
    p = os.path.join(here, '_ref')
    if not os.path.isdir(os.path.join(p, 'gpref')):
        raise RuntimeError('oracle/_ref not built; run python oracle/build_ref.py where /root/reference exists')
    if p not in sys.path:
        sys.path.insert(0, p)
    from gpref.generate_correlation._generate_dense_correlation import generate_dense_correlation
    from gpref.generate_correlation._generate_sparse_correlation import generate_sparse_correlation
    return types.SimpleNamespace(generate_dense_correlation=generate_dense_correlation,
                                 generate_sparse_correlation=generate_sparse_correlation)
