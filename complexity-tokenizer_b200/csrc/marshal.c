/* _ctk_marshal -- CPython helpers for the list-based methods of the Tokenizer shim.
 *
 * The reference's PyO3 layer (src/bindings/tokenizer.rs:203-238) converts list[str] -> Vec<String> and
 * Vec<Vec<u32>> -> list[list[int]] in compiled code; with the GPU doing the work in a millisecond these conversions are
 * what a caller of encode_batch / decode_batch waits for.  Same conversions here, straight between Python objects and
 * the packed buffers of the C ABI (include/ctk.h): no intermediate NumPy arrays, one pass per direction.
 * Marshalling only -- nothing here encodes or decodes.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

/* pack_strs(seq of str) -> (bytes text, bytes offsets_u64[n+1]) */
static PyObject* pack_strs(PyObject* self, PyObject* arg) {
    (void)self;
    PyObject* seq = PySequence_Fast(arg, "argument 'texts': expected a sequence of str");
    if (!seq) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    PyObject** items = PySequence_Fast_ITEMS(seq);
    PyObject* offs = PyBytes_FromStringAndSize(NULL, (n + 1) * 8);
    if (!offs) { Py_DECREF(seq); return NULL; }
    uint64_t* off = (uint64_t*)PyBytes_AS_STRING(offs);
    uint64_t total = 0;
    off[0] = 0;
    for (Py_ssize_t i = 0; i < n; ++i) {
        if (!PyUnicode_Check(items[i])) {
            PyErr_Format(PyExc_TypeError, "argument 'texts': '%s' object cannot be converted to 'PyString'", Py_TYPE(items[i])->tp_name);
            Py_DECREF(offs); Py_DECREF(seq); return NULL;
        }
        Py_ssize_t len;
        if (!PyUnicode_AsUTF8AndSize(items[i], &len)) { Py_DECREF(offs); Py_DECREF(seq); return NULL; }   /* lone surrogate: UnicodeEncodeError */
        total += (uint64_t)len;
        off[i + 1] = total;
    }
    PyObject* text = PyBytes_FromStringAndSize(NULL, (Py_ssize_t)total);
    if (!text) { Py_DECREF(offs); Py_DECREF(seq); return NULL; }
    char* dst = PyBytes_AS_STRING(text);
    for (Py_ssize_t i = 0; i < n; ++i) {
        Py_ssize_t len;
        const char* s = PyUnicode_AsUTF8AndSize(items[i], &len);     /* cached by the first call */
        memcpy(dst + off[i], s, (size_t)len);
    }
    Py_DECREF(seq);
    PyObject* out = PyTuple_Pack(2, text, offs);
    Py_DECREF(text); Py_DECREF(offs);
    return out;
}

/* pack_id_lists(seq of seq of int) -> (bytes ids_u32, bytes offsets_u64[n+1]); ids outside u32 raise OverflowError */
static PyObject* pack_id_lists(PyObject* self, PyObject* arg) {
    (void)self;
    PyObject* seq = PySequence_Fast(arg, "expected a sequence of id sequences");
    if (!seq) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    PyObject** items = PySequence_Fast_ITEMS(seq);
    PyObject** inner = (PyObject**)PyMem_Malloc((size_t)(n ? n : 1) * sizeof(PyObject*));
    PyObject* offs = PyBytes_FromStringAndSize(NULL, (n + 1) * 8);
    if (!inner || !offs) { PyMem_Free(inner); Py_XDECREF(offs); Py_DECREF(seq); return PyErr_NoMemory(); }
    uint64_t* off = (uint64_t*)PyBytes_AS_STRING(offs);
    uint64_t total = 0;
    Py_ssize_t got = 0;
    PyObject* ids = NULL;
    off[0] = 0;
    for (; got < n; ++got) {
        inner[got] = PySequence_Fast(items[got], "expected a sequence of ids");
        if (!inner[got]) goto fail;
        total += (uint64_t)PySequence_Fast_GET_SIZE(inner[got]);
        off[got + 1] = total;
    }
    ids = PyBytes_FromStringAndSize(NULL, (Py_ssize_t)(total * 4));
    if (!ids) goto fail;
    {
        uint32_t* dst = (uint32_t*)PyBytes_AS_STRING(ids);
        for (Py_ssize_t i = 0; i < n; ++i) {
            const Py_ssize_t m = PySequence_Fast_GET_SIZE(inner[i]);
            PyObject** v = PySequence_Fast_ITEMS(inner[i]);
            for (Py_ssize_t k = 0; k < m; ++k) {
                unsigned long x;
                if (PyLong_CheckExact(v[k])) x = PyLong_AsUnsignedLong(v[k]);   /* negative: error set */
                else {                                                         /* numpy integers, anything with __index__ (PyO3's u32 extraction takes them too) */
                    PyObject* ix = PyNumber_Index(v[k]);
                    if (!ix) goto fail;
                    x = PyLong_AsUnsignedLong(ix);
                    Py_DECREF(ix);
                }
                if ((x == (unsigned long)-1 && PyErr_Occurred()) || x > 0xFFFFFFFFul) {
                    if (!PyErr_Occurred()) PyErr_SetString(PyExc_OverflowError, "out of range integral type conversion attempted");
                    goto fail;
                }
                *dst++ = (uint32_t)x;
            }
        }
    }
    for (Py_ssize_t i = 0; i < n; ++i) Py_DECREF(inner[i]);
    PyMem_Free(inner);
    Py_DECREF(seq);
    {
        PyObject* out = PyTuple_Pack(2, ids, offs);
        Py_DECREF(ids); Py_DECREF(offs);
        return out;
    }
fail:
    for (Py_ssize_t i = 0; i < got; ++i) Py_XDECREF(inner[i]);
    PyMem_Free(inner);
    Py_XDECREF(ids); Py_DECREF(offs); Py_DECREF(seq);
    return NULL;
}

/* unpack_ids(ids_addr, off_addr, n[, width = 4[, out, first]]) -> list[list[int]] from one part of a packed encode result
 * (include/ctk.h: ctk_result_part).  width = bytes per id (2 or 4: ctk_result_id_width); with `out` (a list of the whole
 * batch's length) the rows are stored at out[first ..] and out is returned. */
static PyObject* unpack_ids(PyObject* self, PyObject* args) {
    (void)self;
    unsigned long long a_ids, a_off;
    Py_ssize_t n, first = 0;
    int width = 4;
    PyObject* into = NULL;
    if (!PyArg_ParseTuple(args, "KKn|iOn", &a_ids, &a_off, &n, &width, &into, &first)) return NULL;
    if (width != 2 && width != 4) { PyErr_SetString(PyExc_ValueError, "width must be 2 or 4"); return NULL; }
    const uint32_t* ids32 = (const uint32_t*)(uintptr_t)a_ids;
    const uint16_t* ids16 = (const uint16_t*)(uintptr_t)a_ids;
    const uint64_t* off = (const uint64_t*)(uintptr_t)a_off;
    PyObject* out;
    if (into && into != Py_None) {
        if (!PyList_CheckExact(into) || first < 0 || PyList_GET_SIZE(into) < first + n) { PyErr_SetString(PyExc_ValueError, "bad output list"); return NULL; }
        out = into;
        Py_INCREF(out);
    } else {
        out = PyList_New(n);
        first = 0;
        if (!out) return NULL;
    }
    for (Py_ssize_t i = 0; i < n; ++i) {
        const uint64_t lo = off[i], hi = off[i + 1];
        PyObject* row = PyList_New((Py_ssize_t)(hi - lo));
        if (!row) { Py_DECREF(out); return NULL; }
        for (uint64_t k = lo; k < hi; ++k) {
            PyObject* v = PyLong_FromUnsignedLong(width == 2 ? (unsigned long)ids16[k] : (unsigned long)ids32[k]);
            if (!v) { Py_DECREF(row); Py_DECREF(out); return NULL; }
            PyList_SET_ITEM(row, (Py_ssize_t)(k - lo), v);
        }
        if (into && into != Py_None) { if (PyList_SetItem(out, first + i, row) < 0) { Py_DECREF(out); return NULL; } }
        else PyList_SET_ITEM(out, i, row);
    }
    return out;
}

/* unpack_strs(bytes_addr, off_addr, n) -> list[str] from the packed result of ctk_decode_batch (valid UTF-8) */
static PyObject* unpack_strs(PyObject* self, PyObject* args) {
    (void)self;
    unsigned long long a_b, a_off;
    Py_ssize_t n;
    if (!PyArg_ParseTuple(args, "KKn", &a_b, &a_off, &n)) return NULL;
    const char* b = (const char*)(uintptr_t)a_b;
    const uint64_t* off = (const uint64_t*)(uintptr_t)a_off;
    PyObject* out = PyList_New(n);
    if (!out) return NULL;
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject* s = PyUnicode_DecodeUTF8(b + off[i], (Py_ssize_t)(off[i + 1] - off[i]), "strict");
        if (!s) { Py_DECREF(out); return NULL; }
        PyList_SET_ITEM(out, i, s);
    }
    return out;
}

static PyMethodDef methods[] = {
    {"pack_strs", pack_strs, METH_O, "list[str] -> (utf-8 bytes, uint64 offsets)"},
    {"pack_id_lists", pack_id_lists, METH_O, "list[list[int]] -> (uint32 ids, uint64 offsets)"},
    {"unpack_ids", unpack_ids, METH_VARARGS, "(ids address, offsets address, n[, width, out, first]) -> list[list[int]]"},
    {"unpack_strs", unpack_strs, METH_VARARGS, "(bytes address, offsets address, n) -> list[str]"},
    {NULL, NULL, 0, NULL}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_ctk_marshal", "list <-> packed buffer conversions for complexity_tokenizer", -1, methods,
                                    NULL, NULL, NULL, NULL};
PyMODINIT_FUNC PyInit__ctk_marshal(void) { return PyModule_Create(&moddef); }
