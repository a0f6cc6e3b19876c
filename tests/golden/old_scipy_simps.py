"""``scipy.integrate.simps`` as shipped by SciPy 0.19 - 1.5 (the versions contemporary with the reference:
``/root/reference/.travis.yml:4-8`` Python 2.7 / 3.4-3.6, ``pip-requirements.txt:2`` un-pinned ``scipy``).

TEST INFRASTRUCTURE ONLY.  ``simps`` was removed from SciPy 1.14 and the ``simpson`` that replaced it
integrates an even number of samples differently (since 1.11), so the installed SciPy cannot say what the
reference computed for its default ``Ns = nz`` (even: 30, 128, 256).  This file is a transcription of the old
routine (``scipy/integrate/quadrature.py``: ``tupleset``, ``_basic_simps``, ``simps``; BSD-3-Clause,
Copyright (c) 2001-2002 Enthought, Inc. 2003-2019 SciPy Developers) so that the golden vectors are produced
by the arithmetic the reference actually ran, and so that the oracle's and the kernels' even-N rule is pinned
on NON-uniform abscissae (``simps_even.npz``), not only by the uniform-x answers of the old docstring.
"""
import numpy as np


def tupleset(t, i, value):
    l = list(t)
    l[i] = value
    return tuple(l)


def _basic_simps(y, start, stop, x, dx, axis):
    nd = len(y.shape)
    if start is None:
        start = 0
    step = 2
    slice_all = (slice(None),) * nd
    slice0 = tupleset(slice_all, axis, slice(start, stop, step))
    slice1 = tupleset(slice_all, axis, slice(start + 1, stop + 1, step))
    slice2 = tupleset(slice_all, axis, slice(start + 2, stop + 2, step))

    if x is None:  # Even spaced Simpson's rule.
        result = np.sum(dx / 3.0 * (y[slice0] + 4 * y[slice1] + y[slice2]), axis=axis)
    else:
        # Account for possibly different spacings.
        #    Simpson's rule changes a bit.
        h = np.diff(x, axis=axis)
        sl0 = tupleset(slice_all, axis, slice(start, stop, step))
        sl1 = tupleset(slice_all, axis, slice(start + 1, stop + 1, step))
        h0 = h[sl0]
        h1 = h[sl1]
        hsum = h0 + h1
        hprod = h0 * h1
        h0divh1 = h0 / h1
        tmp = hsum / 6.0 * (y[slice0] * (2 - 1.0 / h0divh1) +
                            y[slice1] * hsum * hsum / hprod +
                            y[slice2] * (2 - h0divh1))
        result = np.sum(tmp, axis=axis)
    return result


def simps(y, x=None, dx=1, axis=-1, even='avg'):
    y = np.asarray(y)
    nd = len(y.shape)
    N = y.shape[axis]
    last_dx = dx
    first_dx = dx
    returnshape = 0
    if x is not None:
        x = np.asarray(x)
        if len(x.shape) == 1:
            shapex = [1] * nd
            shapex[axis] = x.shape[0]
            saveshape = x.shape
            returnshape = 1
            x = x.reshape(tuple(shapex))
        elif len(x.shape) != len(y.shape):
            raise ValueError("If given, shape of x must be 1-d or the "
                             "same as y.")
        if x.shape[axis] != N:
            raise ValueError("If given, length of x along axis must be the "
                             "same as y.")
    if N % 2 == 0:
        val = 0.0
        result = 0.0
        slice1 = (slice(None),) * nd
        slice2 = (slice(None),) * nd
        if even not in ['avg', 'last', 'first']:
            raise ValueError("Parameter 'even' must be "
                             "'avg', 'last', or 'first'.")
        # Compute using Simpson's rule on first intervals
        if even in ['avg', 'first']:
            slice1 = tupleset(slice1, axis, -1)
            slice2 = tupleset(slice2, axis, -2)
            if x is not None:
                last_dx = x[slice1] - x[slice2]
            val += 0.5 * last_dx * (y[slice1] + y[slice2])
            result = _basic_simps(y, 0, N - 3, x, dx, axis)
        # Compute using Simpson's rule on last set of intervals
        if even in ['avg', 'last']:
            slice1 = tupleset(slice1, axis, 0)
            slice2 = tupleset(slice2, axis, 1)
            if x is not None:
                first_dx = x[tuple(slice2)] - x[tuple(slice1)]
            val += 0.5 * first_dx * (y[slice2] + y[slice1])
            result += _basic_simps(y, 1, N - 2, x, dx, axis)
        if even == 'avg':
            val /= 2.0
            result /= 2.0
        result = result + val
    else:
        result = _basic_simps(y, 0, N - 2, x, dx, axis)
    if returnshape:
        x = x.reshape(saveshape)
    return result
