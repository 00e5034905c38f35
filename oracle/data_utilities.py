"""TEST INFRASTRUCTURE. Synthetic inputs, restating examples/_utilities/data_utilities.py:22-185 of the reference
(generate_points :22-69, generate_data :76-101 with its numpy.random.seed(31), generate_basis_functions :136-185)."""

import itertools

import numpy


def generate_points(num_points, dimension=2, grid=True):
    """grid: num_points per axis on linspace(0,1) meshgrid ('xy' indexing, raveled) -- data_utilities.py:55-64;
    otherwise numpy.random.rand(num_points, dimension) -- :67."""
    if not grid:
        return numpy.random.rand(num_points, dimension)
    axis = numpy.linspace(0, 1, num_points)
    mesh = numpy.meshgrid(*([axis] * dimension))
    return numpy.stack([m.ravel() for m in mesh], axis=1).astype(float)


def generate_data(points, noise_magnitude):
    """z = sum_k sin(pi x_k) + noise * randn, generator re-seeded with 31 (data_utilities.py:93-101)."""
    z = numpy.zeros(points.shape[0])
    for k in range(points.shape[1]):
        z += numpy.sin(points[:, k] * numpy.pi)
    numpy.random.seed(31)
    z += noise_magnitude * numpy.random.randn(points.shape[0])
    return z


def generate_basis_functions(points, polynomial_degree=2, trigonometric=False):
    """Monomials of total degree <= p in the column order of data_utilities.py:150-173 (meshgrid of the powers,
    raveled, filtered by sum), optionally followed by the trigonometric block of :176-183 (including the reference's
    overlapping column indices i+0 / i+1, which leave the last column uninitialised for dimension > 1: we restate the
    defined part only and zero the rest)."""
    n, dim = points.shape
    grids = numpy.meshgrid(*([numpy.arange(polynomial_degree + 1)] * dim))
    powers = numpy.array([g.ravel() for g in grids])
    powers = powers[:, powers.sum(axis=0) <= polynomial_degree]
    X = numpy.ones((n, powers.shape[1]))
    for j in range(powers.shape[1]):
        for i in range(dim):
            X[:, j] *= points[:, i] ** powers[i, j]
    if trigonometric:
        T = numpy.zeros((n, 2 * dim))
        for i in range(dim):
            T[:, i] = numpy.sin(points[:, i] * numpy.pi)
            T[:, i + 1] = numpy.cos(points[:, i] * numpy.pi)
        X = numpy.c_[X, T]
    return X
