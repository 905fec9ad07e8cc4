/* Python int lists <-> 4 x u64 little-endian limb buffers, in C (CPython API).
 *
 * Host-side marshalling only (the reference does this element by element in
 * Cython: intToZZp / ZZpToInt, hbmpc_ntl_helpers.pyx:20-35).  Loaded with
 * ctypes.PyDLL, so the GIL is held and PyObject* arguments are passed as is.
 * honeybadgermpc_b200/ntl/__init__.py falls back to its pure-Python version
 * when this helper is not built -- it computes nothing on the batch.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <string.h>
#include <structmember.h>

static int ge_le32(const unsigned char* a, const unsigned char* b) {
  for (int i = 31; i >= 0; i--) {
    if (a[i] != b[i]) return a[i] > b[i];
  }
  return 1;
}

#if PY_VERSION_HEX >= 0x030C0000 && PYLONG_BITS_IN_DIGIT == 30
/* A non-negative int below 2^256 -> 32 little-endian bytes straight from its 30-bit digits
 * (CPython >= 3.12 layout, cpython/longintrepr.h); 0 when the value needs the general path
 * (negative: the caller raises; >= 2^256: reduced mod p there). */
static int fast_le32(PyObject* v, unsigned char* out) {
  const PyLongObject* lv = (const PyLongObject*)v;
  const uintptr_t tag = lv->long_value.lv_tag;
  if ((tag & _PyLong_SIGN_MASK) == 2) return 0;
  const Py_ssize_t nd = (Py_ssize_t)(tag >> _PyLong_NON_SIZE_BITS);
  if (nd > 9) return 0;
  unsigned long long l[4] = {0, 0, 0, 0};
  if ((tag & _PyLong_SIGN_MASK) != 1) { /* not zero */
    for (Py_ssize_t k = 0; k < nd; k++) {
      const unsigned long long d = lv->long_value.ob_digit[k];
      const int bit = 30 * (int)k, limb = bit >> 6, off = bit & 63;
      l[limb] |= d << off;
      if (off > 34) {
        const unsigned long long hi = d >> (64 - off);
        if (limb < 3)
          l[limb + 1] |= hi;
        else if (hi)
          return 0; /* >= 2^256 */
      }
    }
  }
  memcpy(out, l, 32); /* little-endian host */
  return 1;
}
#endif

/* one element -> 32 bytes, reduced mod p; returns 0, or -1 with an exception set */
static int pack_one(PyObject* v, PyObject* p_obj, const unsigned char* p_le, unsigned char* out) {
  if (!PyLong_Check(v)) {
    PyErr_Format(PyExc_TypeError, "expected int, got %s", Py_TYPE(v)->tp_name);
    return -1;
  }
  if (_PyLong_Sign(v) < 0) {
    PyErr_SetString(PyExc_OverflowError, "can't convert negative int to unsigned");
    return -1;
  }
#if PY_VERSION_HEX >= 0x030C0000 && PYLONG_BITS_IN_DIGIT == 30
  if (fast_le32(v, out)) {
    if (!ge_le32(out, p_le)) return 0;
  } else
#endif
  if (_PyLong_NumBits(v) <= 256) {
    if (_PyLong_AsByteArray((PyLongObject*)v, out, 32, 1, 0) < 0) return -1;
    if (!ge_le32(out, p_le)) return 0;
  }
  /* v >= p: reduce like to_ZZ_p */
  PyObject* r = PyNumber_Remainder(v, p_obj);
  if (!r) return -1;
  int rc = _PyLong_AsByteArray((PyLongObject*)r, out, 32, 1, 0);
  Py_DECREF(r);
  return rc < 0 ? -1 : 0;
}

/* rows: sequence of sequences of ints -> out[len(rows)][width][32]; short rows are
 * zero padded, long rows truncated.  Returns None. */
PyObject* hbg_py_pack_rows(PyObject* rows, Py_ssize_t width, PyObject* p_obj, unsigned char* out) {
  unsigned char p_le[32];
  if (!PyLong_Check(p_obj) || _PyLong_AsByteArray((PyLongObject*)p_obj, p_le, 32, 1, 0) < 0) {
    if (!PyErr_Occurred()) PyErr_SetString(PyExc_TypeError, "modulus must be an int below 2**256");
    return NULL;
  }
  PyObject* outer = PySequence_Fast(rows, "rows must be a sequence");
  if (!outer) return NULL;
  Py_ssize_t nrows = PySequence_Fast_GET_SIZE(outer);
  for (Py_ssize_t i = 0; i < nrows; i++) {
    PyObject* row = PySequence_Fast(PySequence_Fast_GET_ITEM(outer, i), "each row must be a sequence");
    if (!row) {
      Py_DECREF(outer);
      return NULL;
    }
    Py_ssize_t m = PySequence_Fast_GET_SIZE(row);
    unsigned char* dst = out + (size_t)i * width * 32;
    Py_ssize_t upto = m < width ? m : width;
    for (Py_ssize_t j = 0; j < upto; j++) {
      if (pack_one(PySequence_Fast_GET_ITEM(row, j), p_obj, p_le, dst + j * 32) < 0) {
        Py_DECREF(row);
        Py_DECREF(outer);
        return NULL;
      }
    }
    if (upto < width) memset(dst + upto * 32, 0, (size_t)(width - upto) * 32);
    Py_DECREF(row);
  }
  Py_DECREF(outer);
  Py_RETURN_NONE;
}

/* 32 little-endian bytes -> int.  _PyLong_FromByteArray walks the value byte by byte (~90 ns per
 * 256-bit element, the largest single cost of unpacking a big batch); from CPython 3.12 on
 * _PyLong_FromDigits takes the 30-bit digits directly, which four 64-bit limbs yield with a few
 * shifts.  Same value either way (tests/test_marshal.py compares both on edge cases). */
static PyObject* long_from_le32(const unsigned char* in) {
#if PY_VERSION_HEX >= 0x030C0000 && PYLONG_BITS_IN_DIGIT == 30
  unsigned long long l[4];
  memcpy(l, in, 32); /* little-endian host (x86-64 / aarch64) */
  digit d[9];
  for (int k = 0; k < 9; k++) {
    const int bit = 30 * k, limb = bit >> 6, off = bit & 63;
    unsigned long long v = l[limb] >> off;
    if (off > 34 && limb < 3) v |= l[limb + 1] << (64 - off);
    d[k] = (digit)(v & 0x3fffffffULL);
  }
  Py_ssize_t n = 9;
  while (n > 0 && d[n - 1] == 0) n--;
  if (n <= 1) return PyLong_FromUnsignedLong(n ? (unsigned long)d[0] : 0ul); /* small ints stay the cached ones */
  return (PyObject*)_PyLong_FromDigits(0, n, d);
#else
  return _PyLong_FromByteArray(in, 32, 1, 0);
#endif
}

/* for the tests: the two conversions side by side on one 32-byte value */
PyObject* hbg_py_long_from_le32(const unsigned char* in, int reference) {
  return reference ? _PyLong_FromByteArray(in, 32, 1, 0) : long_from_le32(in);
}

/* in[batch][width][32] -> list of `batch` lists of `width` ints */
static PyObject* unpack_rows_inner(const unsigned char* in, Py_ssize_t batch, Py_ssize_t width) {
  PyObject* outer = PyList_New(batch);
  if (!outer) return NULL;
  for (Py_ssize_t i = 0; i < batch; i++) {
    PyObject* row = PyList_New(width);
    if (!row) {
      Py_DECREF(outer);
      return NULL;
    }
    for (Py_ssize_t j = 0; j < width; j++) {
      PyObject* v = long_from_le32(in + ((size_t)i * width + j) * 32);
      if (!v) {
        Py_DECREF(row);
        Py_DECREF(outer);
        return NULL;
      }
      PyList_SET_ITEM(row, j, v);
    }
    PyList_SET_ITEM(outer, i, row);
  }
  return outer;
}

PyObject* hbg_py_unpack_rows(const unsigned char* in, Py_ssize_t batch, Py_ssize_t width) {
  /* tens of thousands of row lists allocated in one go trigger a young-generation collection
   * every 700 of them, none of which can free anything: pause the collector for the loop */
#if PY_VERSION_HEX >= 0x030A0000
  const int was_enabled = PyGC_Disable();
  PyObject* r = unpack_rows_inner(in, batch, width);
  if (was_enabled) PyGC_Enable();
  return r;
#else
  return unpack_rows_inner(in, batch, width);
#endif
}

/* Offset of a __slots__ member of a class, or -1 with an exception set. */
static Py_ssize_t slot_offset(PyObject* cls, const char* name) {
  PyObject* d = PyObject_GetAttrString(cls, name);
  if (!d) return -1;
  if (Py_TYPE(d) != &PyMemberDescr_Type || ((PyMemberDescrObject*)d)->d_member->type != T_OBJECT_EX) {
    Py_DECREF(d);
    PyErr_Format(PyExc_TypeError, "%s is not a __slots__ member", name);
    return -1;
  }
  Py_ssize_t off = ((PyMemberDescrObject*)d)->d_member->offset;
  Py_DECREF(d);
  return off;
}

/* elems: sequence of objects with an int attribute `value` (GFElement: read from its slot; any
 * other object: getattr) -> out[width][32], reduced mod p, zero padded to `width`.  The inverse of
 * hbg_py_wrap_elements: what batch_reconstruct does with its input shares
 * ([share.value for share in shares] + pack in one pass).  Returns None. */
PyObject* hbg_py_pack_elements(PyObject* elems, Py_ssize_t width, PyObject* cls, PyObject* p_obj,
                               unsigned char* out) {
  unsigned char p_le[32];
  if (!PyLong_Check(p_obj) || _PyLong_AsByteArray((PyLongObject*)p_obj, p_le, 32, 1, 0) < 0) {
    if (!PyErr_Occurred()) PyErr_SetString(PyExc_TypeError, "modulus must be an int below 2**256");
    return NULL;
  }
  if (!PyType_Check(cls)) {
    PyErr_SetString(PyExc_TypeError, "cls must be a class");
    return NULL;
  }
  const Py_ssize_t ov = slot_offset(cls, "value");
  if (ov < 0) return NULL;
  PyObject* seq = PySequence_Fast(elems, "shares must be a sequence");
  if (!seq) return NULL;
  const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
  const Py_ssize_t upto = n < width ? n : width;
  for (Py_ssize_t i = 0; i < upto; i++) {
    PyObject* e = PySequence_Fast_GET_ITEM(seq, i);
    int rc;
    if (Py_TYPE(e) == (PyTypeObject*)cls) {
      PyObject* v = *(PyObject**)((char*)e + ov);
      if (!v) {
        PyErr_SetString(PyExc_AttributeError, "value");
        rc = -1;
      } else {
        rc = pack_one(v, p_obj, p_le, out + (size_t)i * 32);
      }
    } else {
      PyObject* v = PyObject_GetAttrString(e, "value");
      rc = v ? pack_one(v, p_obj, p_le, out + (size_t)i * 32) : -1;
      Py_XDECREF(v);
    }
    if (rc < 0) {
      Py_DECREF(seq);
      return NULL;
    }
  }
  if (upto < width) memset(out + (size_t)upto * 32, 0, (size_t)(width - upto) * 32);
  Py_DECREF(seq);
  Py_RETURN_NONE;
}

/* in[count][32] (canonical residues) -> list of `count` instances of `cls`, a class with the
 * slots (value, field, modulus) -- GFElement -- built without running its __init__:
 * value = the int, field / modulus = the given objects.  What batch_reconstruct returns;
 * in Python this loop is the largest single cost of a big open once the kernels are fast. */
PyObject* hbg_py_wrap_elements(const unsigned char* in, Py_ssize_t count, PyObject* cls, PyObject* field,
                               PyObject* modulus) {
  if (!PyType_Check(cls)) {
    PyErr_SetString(PyExc_TypeError, "cls must be a class");
    return NULL;
  }
  PyTypeObject* tp = (PyTypeObject*)cls;
  const Py_ssize_t ov = slot_offset(cls, "value"), of = slot_offset(cls, "field"),
                   om = slot_offset(cls, "modulus");
  if (ov < 0 || of < 0 || om < 0) return NULL;
  PyObject* out = PyList_New(count);
  if (!out) return NULL;
  /* the allocations below count towards the collector's young-generation threshold although
   * every element is untracked at once: pause it for the loop (see hbg_py_unpack_rows) */
#if PY_VERSION_HEX >= 0x030A0000
  const int gc_was_enabled = PyGC_Disable();
#else
  const int gc_was_enabled = 0;
#endif
  for (Py_ssize_t i = 0; i < count; i++) {
    PyObject* v = long_from_le32(in + (size_t)i * 32);
    PyObject* e = v ? tp->tp_alloc(tp, 0) : NULL;
    if (!e) {
      Py_XDECREF(v);
      Py_DECREF(out);
#if PY_VERSION_HEX >= 0x030A0000
      if (gc_was_enabled) PyGC_Enable();
#endif
      return NULL;
    }
    /* An element refers to an int and to its (immortal, per-modulus) field object: it can never
     * be part of a reference cycle, so it does not need to be tracked by the cyclic GC -- and
     * hundreds of thousands of tracked elements make every full collection of the process
     * slower (measured: a 49 152-share open spent more time in collections triggered by its own
     * bulk allocations than in anything else). */
    PyObject_GC_UnTrack(e);
    *(PyObject**)((char*)e + ov) = v;  /* steals the reference */
    Py_INCREF(field);
    *(PyObject**)((char*)e + of) = field;
    Py_INCREF(modulus);
    *(PyObject**)((char*)e + om) = modulus;
    PyList_SET_ITEM(out, i, e);
  }
#if PY_VERSION_HEX >= 0x030A0000
  if (gc_was_enabled) PyGC_Enable();
#endif
  return out;
}
