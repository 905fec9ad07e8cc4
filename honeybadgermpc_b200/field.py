"""Prime field objects at the Python boundary of the reconstruction path.

Mirrors the interface of the reference's ``honeybadgermpc/field.py`` (``GF``
multiton :41-66, ``GFElement`` :68-290) as far as the path needs it: shares
enter ``batch_reconstruct`` as ``GFElement`` and leave it as ``GFElement``
(batch_reconstruction.py:126, :227); ``EvalPoint`` hands out ``GFElement``
points.  Values cross into the CUDA library as ``.value`` ints.  No gmpy2: the
primality check is a deterministic-base Miller-Rabin.
"""

from random import Random


class FieldsNotIdentical(Exception):
    pass


_SMALL_PRIMES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47)


def is_probable_prime(n):
    if n < 2:
        return False
    for q in _SMALL_PRIMES:
        if n % q == 0:
            return n == q
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in _SMALL_PRIMES + (53, 59, 61, 67, 71):
        if a % n == 0:
            continue
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def sqrt_mod_prime(a, p):
    """A square root of ``a`` modulo the odd prime ``p`` (Tonelli-Shanks on Python ints; the
    smaller of the two roots).  Raises ``ValueError`` for a non-residue."""
    a %= p
    if a == 0:
        return 0
    if pow(a, (p - 1) // 2, p) != 1:
        raise ValueError("not a quadratic residue")
    q, s = p - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    g = 2
    while pow(g, (p - 1) // 2, p) != p - 1:
        g += 1
    m, c, t, r = s, pow(g, q, p), pow(a, q, p), pow(a, (q + 1) // 2, p)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
        b = pow(c, 1 << (m - i - 1), p)
        m, c = i, b * b % p
        t, r = t * c % p, r * b % p
    return min(r, p - r)


class GF:
    """One object per modulus (field.py:41-58)."""

    _cache = {}

    def __new__(cls, modulus):
        obj = cls._cache.get(modulus)
        if obj is None:
            if not is_probable_prime(int(modulus)):
                raise ValueError(f"{modulus} is not a prime")
            obj = super().__new__(cls)
            obj.modulus = int(modulus)
            cls._cache[modulus] = obj
        return obj

    def __call__(self, value):
        return GFElement(value, self)

    def __reduce__(self):
        return (GF, (self.modulus,))

    def random(self, seed=None):
        # same draw as the reference (field.py:64-65) so get_omega(seed=0) agrees
        return GFElement(Random(seed).randint(0, self.modulus - 1), self)

    def wrap_canonical(self, values):
        """``[GFElement(v, self) for v in values]`` for ints already in [0, p) --
        what ``batch_reconstruct`` returns.  Skips the per-element ``int() % p``
        of the constructor: wrapping B results is otherwise the single largest
        cost of a large open once the kernels are fast."""
        new = object.__new__
        cls = GFElement
        modulus = self.modulus
        out = []
        append = out.append
        for v in values:
            e = new(cls)
            e.value = v
            e.field = self
            e.modulus = modulus
            append(e)
        return out


class GFElement:
    __slots__ = ("value", "field", "modulus")

    def __init__(self, value, gf):
        self.field = gf
        self.modulus = gf.modulus
        self.value = int(value) % gf.modulus

    def _coerce(self, other):
        if isinstance(other, GFElement):
            if other.field is not self.field:
                raise FieldsNotIdentical
            return other.value
        if isinstance(other, int):
            return other
        return None

    def __int__(self):
        return self.value

    def __add__(self, other):
        v = self._coerce(other)
        return NotImplemented if v is None else GFElement(self.value + v, self.field)

    __radd__ = __add__

    def __sub__(self, other):
        v = self._coerce(other)
        return NotImplemented if v is None else GFElement(self.value - v, self.field)

    def __rsub__(self, other):
        return GFElement(other - self.value, self.field)

    def __mul__(self, other):
        v = self._coerce(other)
        return NotImplemented if v is None else GFElement(self.value * v, self.field)

    __rmul__ = __mul__

    def __neg__(self):
        return GFElement(-self.value, self.field)

    def __pow__(self, e):
        return GFElement(pow(self.value, e, self.modulus), self.field)

    def inverse(self):
        if self.value == 0:
            raise ZeroDivisionError("inverse of 0")
        return GFElement(pow(self.value, -1, self.modulus), self.field)

    __invert__ = inverse  # the reference spells the inverse ~x (field.py:126-150)

    def sqrt(self):
        """field.py:170-208: a square root (either one; the offline bit generation only needs
        every party to take the same root of the same public value).  A non-residue fails the
        reference's assertion."""
        assert self.modulus % 2 == 1, "Modulus must be odd"
        assert pow(self.value, (self.modulus - 1) // 2, self.modulus) in (0, 1)
        return GFElement(sqrt_mod_prime(self.value, self.modulus), self.field)

    def __truediv__(self, other):
        v = self._coerce(other)
        if v is None:
            return NotImplemented
        return self * GFElement(v, self.field).inverse()

    __floordiv__ = __truediv__  # field.py:151-162: every division is the field division

    def __rtruediv__(self, other):
        return GFElement(other, self.field) * self.inverse()

    __rfloordiv__ = __rtruediv__

    def bit(self, index):
        return (self.value >> index) & 1

    def signed(self):
        """value, or value - p when it is above (p - 1) / 2 (field.py:214-222)"""
        return self.value - self.modulus if 2 * self.value > self.modulus - 1 else self.value

    def unsigned(self):
        return self.value

    def __eq__(self, other):
        # field.py:239-246: elements of different fields do not compare (FieldsNotIdentical);
        # anything else is compared with the raw value
        if isinstance(other, GFElement):
            if other.field is not self.field:
                raise FieldsNotIdentical
            return self.value == other.value
        return self.value == other

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((self.value, self.modulus))

    def __bool__(self):
        return self.value != 0

    def __repr__(self):
        return "{%d}" % self.value

    __str__ = __repr__
