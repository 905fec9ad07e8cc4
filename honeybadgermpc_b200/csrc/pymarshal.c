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

static int ge_le32(const unsigned char* a, const unsigned char* b) {
  for (int i = 31; i >= 0; i--) {
    if (a[i] != b[i]) return a[i] > b[i];
  }
  return 1;
}

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

/* in[batch][width][32] -> list of `batch` lists of `width` ints */
PyObject* hbg_py_unpack_rows(const unsigned char* in, Py_ssize_t batch, Py_ssize_t width) {
  PyObject* outer = PyList_New(batch);
  if (!outer) return NULL;
  for (Py_ssize_t i = 0; i < batch; i++) {
    PyObject* row = PyList_New(width);
    if (!row) {
      Py_DECREF(outer);
      return NULL;
    }
    for (Py_ssize_t j = 0; j < width; j++) {
      PyObject* v = _PyLong_FromByteArray(in + ((size_t)i * width + j) * 32, 32, 1, 0);
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
