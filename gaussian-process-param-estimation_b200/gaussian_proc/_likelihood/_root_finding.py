"""
Scalar bracketing + Chandrupatla root finder that drives d l/d eta = 0.

Same algorithm and argument meaning as the reference's gaussian_proc/_likelihood/_root_finding.py
(find_interval_with_sign_change :21-148, chandrupatla_method :155-309; the latter follows Chandrupatla 1997 /
Scherer 2010). Host-only control flow: every f() call is one full likelihood-derivative evaluation on the GPU.
Only scalar f is supported (the reference's array branch is never reached on this path).
"""

import numpy

__all__ = ['find_interval_with_sign_change', 'chandrupatla_method']


def _opposite(u, v):
    return numpy.sign(u) != numpy.sign(v)


def find_interval_with_sign_change(f, bracket, num_bracket_trials, args=(), ):
    """Returns (bracket_found, bracket, bracket_values). Each trial first probes the midpoint; if that neither
    brackets a root nor improves on both ends, it probes half an interval beyond the end with the smaller |f| and
    shifts the window there (_root_finding.py:47-139)."""
    lo, hi = bracket[0], bracket[1]
    f_lo, f_hi = f(lo, *args), f(hi, *args)
    for trial in range(1, num_bracket_trials + 1):
        if _opposite(f_lo, f_hi):
            return True, [lo, hi], [f_lo, f_hi]
        print('bracket was not found. Search for bracket. Iteration: %d' % trial)
        mid = lo * 0.5 + hi * 0.5
        f_mid = f(mid, *args)
        lo_is_smaller = numpy.abs(f_lo) < numpy.abs(f_hi)
        if _opposite(f_lo, f_mid):
            if lo_is_smaller:
                return True, [lo, mid], [f_lo, f_mid]
            return True, [mid, hi], [f_mid, f_hi]
        if numpy.abs(f_mid) < numpy.min([numpy.abs(f_lo), numpy.abs(f_hi)]):
            # midpoint is closer to zero than both ends: shrink towards it, keeping the smaller end
            if lo_is_smaller:
                hi, f_hi = mid, f_mid
            else:
                lo, f_lo = mid, f_mid
            continue
        # extrapolate beyond the end with the smaller |f|
        t = 1.5 if numpy.abs(f_lo) > numpy.abs(f_hi) else -0.5
        ext = lo * (1 - t) + hi * t
        f_ext = f(ext, *args)
        if _opposite(f_lo, f_ext):
            if t > 0:
                return True, [ext, lo], [f_ext, f_lo]
            return True, [hi, ext], [f_hi, f_ext]
        if t > 0:
            lo, f_lo, hi, f_hi = hi, f_hi, ext, f_ext
        else:
            hi, f_hi, lo, f_lo = lo, f_lo, ext, f_ext
    return False, [lo, hi], [f_lo, f_hi]


def chandrupatla_method(f, bracket, bracket_values, verbose=False, eps_m=None, eps_a=None, maxiter=50, args=(), ):
    """Chandrupatla's hybrid of bisection and inverse quadratic interpolation. Returns {'root', 'iterations'}.
    State: a = newest point, b = the bracketing counterpart, c = the point dropped last; x(t) = a + t (b - a)."""
    b, a = bracket[0], bracket[1]
    if bracket_values is None:
        fa, fb = f(a, *args), f(b, *args)
    else:
        fb, fa = bracket_values[0], bracket_values[1]
    if numpy.sign(fa) * numpy.sign(fb) > 0:
        raise AssertionError('the bracket does not enclose a sign change')
    c, fc = a, fa
    eps = numpy.finfo(float).eps
    eps_m = eps if eps_m is None else eps_m
    eps_a = 2 * eps if eps_a is None else eps_a
    t = 0.5
    iterations = 0
    root = a
    for _ in range(maxiter):
        xt = a + t * (b - a)
        ft = f(xt, *args)
        if verbose:
            print('t=%s xt=%s ft=%s a=%s b=%s c=%s' % (t, xt, ft, a, b, c))
        if numpy.sign(ft) == numpy.sign(fa):
            c, fc = a, fa                 # xt replaces a on the same side; old a is dropped into c
        else:
            c, fc, b, fb = b, fb, a, fa   # xt lands on b's side: old a becomes the counterpart
        a, fa = xt, ft
        if numpy.abs(fa) < numpy.abs(fb):
            root, f_root = a, fa
        else:
            root, f_root = b, fb
        tol = 2 * eps_m * numpy.abs(root) + eps_a
        tlim = tol / numpy.abs(b - c)
        if f_root == 0 or tlim > 0.5:
            break
        iterations += 1
        xi = (a - b) / (c - b)
        phi = (fa - fb) / (fc - fb)
        if phi ** 2 < xi and (1 - phi) ** 2 < 1 - xi:
            t = fa / (fb - fa) * fc / (fb - fc) + (c - a) / (b - a) * fa / (fc - fa) * fb / (fc - fb)
        else:
            t = 0.5
        t = numpy.minimum(1 - tlim, numpy.maximum(tlim, t))
    return {'root': root, 'iterations': iterations}
