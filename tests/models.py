"""Model builders shared by parity tests: the same programme is built once with the product
package and once with the oracle (both expose the reference's names)."""
import numpy as np


def f3_model(m):
    """examples/process_decomposition/script.jl:10-14 with the kernels of
    test/gaussian_process_probabilistic_programme.jl:19-21."""
    return m.gppp(lambda GP: (lambda f1, f2: dict(f1=f1, f2=f2, f3=f1 + f2))(
        GP(m.SEKernel()), GP(m.Matern52Kernel())))


def toy_model(m):
    """test/gaussian_process_probabilistic_programme.jl:18-24: f3 = f1 + 3*f2 with means."""
    return m.gppp(lambda GP: (lambda f1, f2: dict(f1=f1, f2=f2, f3=f1 + 3 * f2))(
        GP(np.sin, m.SEKernel()), GP(np.cos, m.Matern52Kernel())))


def mixing_model(m, coeffs=((.2, .8), (.3, .7), (.9, .2), (.5, .5), (.7, .1))):
    """examples/naive-linear-mixing/script.jl:11-17 extended to 5 outputs (SURVEY 8d cfg 5)."""
    def build(GP):
        f1 = GP(m.SEKernel())
        f2 = GP(m.with_lengthscale(m.SEKernel(), 0.1))
        out = dict(f1=f1, f2=f2)
        for i, (a, b) in enumerate(coeffs):
            out[f"g{i + 1}"] = a * f1 + b * f2
        return out
    return m.gppp(build)


def rich_model(m):
    """Exercises every lowering rule: sum, const/function scaling, negation, stretch, shift,
    periodic, kernel sums/scales, white + constant kernels, constant and function means."""
    def build(GP):
        f1 = GP(m.SEKernel())
        f2 = GP(1.5, 0.7 * m.Matern32Kernel() + 0.1 * m.WhiteKernel())
        f3 = GP(np.cos, m.with_lengthscale(m.Matern12Kernel(), 2.0) + m.ConstantKernel(0.3))
        f4 = GP(m.Matern52Kernel())  # only ever seen through the 2-D periodic embedding
        g1 = m.stretch(f1, 0.5) + 2.0 * f2
        g2 = (lambda x: 1.0 + 0.1 * x * x) * m.shift(f1, 0.3) - f3
        g3 = m.periodic(f4, 0.25) + np.sin + f2
        g4 = g1 - g2
        return dict(f1=f1, f2=f2, f3=f3, f4=f4, g1=g1, g2=g2, g3=g3, g4=g4)
    return m.gppp(build)
