"""Evaluation points of the reconstruction path.

Mirrors ``EvalPoint`` (polynomial.py:385-423) and ``get_omega``
(polynomial.py:253-268) of the reference: same attributes (``n``, ``order``,
``omega``, ``omega2``, ``field``, ``use_omega_powers``), ``point(i)`` returns a
``GFElement``.  The omega for a given (field, n) is identical to the
reference's because it is derived from the same ``Random(0)`` draw.
The points are cached (the reference recomputes ``omega2.value ** (2*i)`` as an
unreduced big integer on every call, polynomial.py:418).
"""

from .field import GF, GFElement


def get_omega(field, n, seed=None):
    """An n-th primitive root of unity of ``field`` (n a power of two)."""
    assert n & (n - 1) == 0, "n must be a power of 2"
    while True:
        x = field.random(seed)
        y = x ** ((field.modulus - 1) // n)
        if y == 1 or y ** (n // 2) == 1:
            seed = None  # the reference retries unseeded (polynomial.py:264-265)
            continue
        return y


class EvalPoint:
    def __init__(self, field, n, use_omega_powers=False):
        if not isinstance(field, GF):
            field = GF(field)
        self.field = field
        self.n = n
        self.use_omega_powers = use_omega_powers
        if use_omega_powers:
            self.order = n if n & (n - 1) == 0 else 2 ** n.bit_length()
            self.omega2 = get_omega(field, 2 * self.order, seed=0)
            self.omega = self.omega2 ** 2
        else:
            self.order = n
            self.omega2 = None
            self.omega = None
        self._cache = {}

    def __call__(self, i):
        v = self._cache.get(i)
        if v is None:
            if self.use_omega_powers:
                v = pow(self.omega2.value, 2 * i, self.field.modulus)
            else:
                v = (i + 1) % self.field.modulus
            self._cache[i] = v
        return GFElement(v, self.field)

    def zero(self):
        return self.field(0)


class OpenedPolynomial(list):
    """What ``robust_reconstruct`` returns: the coefficients of the opened polynomial (a
    ``list`` of ints, lowest degree first) that can also be called like the reference's
    ``polynomials_over(field)(coeffs)`` object -- ``Mpc.open_share`` evaluates it at zero
    (mpc.py:157).  ``coeffs``: the coefficients as field elements, trailing zeros stripped
    (polynomial.py:36).  Not the reference's general ``Polynomial`` class (arithmetic,
    interpolation: outside this path)."""

    def __init__(self, coeffs, field):
        super().__init__(int(c) for c in coeffs)
        self.field = field

    @property
    def coeffs(self):
        vals = list(self)
        while vals and vals[-1] == 0:
            vals.pop()
        return [self.field(v) for v in vals]

    def degree(self):
        return max(len(self.coeffs) - 1, 0)

    def __call__(self, x):
        p = self.field.modulus
        xv = int(x) % p
        acc = 0
        for c in reversed(self):
            acc = (acc * xv + c) % p
        return self.field(acc)
