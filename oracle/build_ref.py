"""
TEST INFRASTRUCTURE -- not part of the product path.

Recipe that compiles the reference's own Cython correlation generators (the only compiled code on the reference's
hot path: gaussian_proc/generate_correlation/{_kernels,_generate_dense_correlation,_generate_sparse_correlation}.pyx)
from the sources where they lie under /root/reference into ``oracle/_ref/`` (git-ignored; only the built ``.so``
files and generated empty ``__init__.py`` files land there, never reference sources).

* Build directives are the reference's own (setup.py:993-999: boundscheck/wraparound off, cdivision on, ...) plus
  ``legacy_implicit_noexcept=True`` -- the Cython-0.29 semantics the code was written for (pyproject.toml:3); without
  it Cython 3 re-acquires the GIL on every ``cdef ... nogil`` call and the OpenMP loop serialises.
* The sparse generator is broken as shipped (two call sites pass the wrong number of arguments,
  _generate_sparse_correlation.pyx:390 and :542-545). The build applies exactly those two call-site
  fixes to the TEMPORARY copy under /tmp that Cython reads; "reference + these two fixes" is the sparse oracle.
* The compiler is /usr/bin/g++ (the /opt/gcc wrapper cannot link -fopenmp in this image).

Usage: python oracle/build_ref.py            (no-op with a message when /root/reference is absent)
"""

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('GP_REFERENCE_ROOT', '/root/reference')
OUT = os.path.join(HERE, '_ref')
PKG = 'gpref'

SETUP = r'''
import numpy
from setuptools import setup, Extension
from Cython.Build import cythonize
exts = [Extension("%(pkg)s.generate_correlation." + m, ["%(pkg)s/generate_correlation/" + m + ".pyx"], language="c++",
                  include_dirs=[numpy.get_include()],
                  extra_compile_args=["-O3", "-fopenmp"], extra_link_args=["-fopenmp"])
        for m in ("_kernels", "_generate_dense_correlation", "_generate_sparse_correlation")]
setup(name="%(pkg)s", ext_modules=cythonize(exts, language_level="3", include_path=[numpy.get_include(), "."],
      compiler_directives=dict(boundscheck=False, cdivision=True, wraparound=False, nonecheck=False,
                               embedsignature=True, legacy_implicit_noexcept=True)))
'''


def build(verbose=True):
    src = os.path.join(REF, 'gaussian_proc', 'generate_correlation')
    if not os.path.isdir(src):
        if verbose:
            print('oracle/build_ref.py: %s not present; keeping whatever is in oracle/_ref/' % src)
        return False
    work = tempfile.mkdtemp(prefix='gpref_build_')
    try:
        pk = os.path.join(work, PKG, 'generate_correlation')
        os.makedirs(pk)
        open(os.path.join(work, PKG, '__init__.py'), 'w').close()
        open(os.path.join(pk, '__init__.py'), 'w').close()
        for f in os.listdir(src):
            if f.endswith(('.pyx', '.pxd')):
                shutil.copy(os.path.join(src, f), os.path.join(pk, f))
        sp = os.path.join(pk, '_generate_sparse_correlation.pyx')
        os.chmod(sp, 0o644)
        text = open(sp).read()
        a = '_ball_volume(geometric_mean_radius)'
        b = '_estimate_max_nnz(\n            matrix_size,\n            dimension,\n            density)'
        assert text.count(a) == 1 and text.count(b) == 1, 'reference sparse generator changed; review the fixes'
        text = text.replace(a, '_ball_volume(geometric_mean_radius, dimension)')
        text = text.replace(b, '_estimate_max_nnz(\n            matrix_size,\n            correlation_scale,\n'
                               '            dimension,\n            density)')
        open(sp, 'w').write(text)
        open(os.path.join(work, 'setup_ref.py'), 'w').write(SETUP % {'pkg': PKG})
        env = dict(os.environ, CC='/usr/bin/gcc', CXX='/usr/bin/g++', LDSHARED='/usr/bin/g++ -shared')
        r = subprocess.run([sys.executable, 'setup_ref.py', 'build_ext', '--inplace'], cwd=work, env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            print(r.stdout[-4000:])
            raise RuntimeError('reference Cython build failed')
        dst = os.path.join(OUT, PKG, 'generate_correlation')
        os.makedirs(dst, exist_ok=True)
        open(os.path.join(OUT, PKG, '__init__.py'), 'w').close()
        open(os.path.join(dst, '__init__.py'), 'w').close()
        n = 0
        for f in os.listdir(pk):
            if f.endswith('.so'):
                shutil.copy(os.path.join(pk, f), os.path.join(dst, f))
                n += 1
        if verbose:
            print('oracle/build_ref.py: built %d reference extension modules into %s' % (n, dst))
        return n == 3
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == '__main__':
    ok = build()
    sys.exit(0 if ok or not os.path.isdir(REF) else 1)
