/* Host side of the Python boundary, the part the reference does in PyO3 (src/python/bindings.rs:337-350: extract
 * Vec<String> from the list, collect Vec<Vec<u32>> back into lists): packing list[str] into one UTF-8 buffer with
 * offsets, and turning the id stream back into list[list[int]].  Plain CPython C API; no device code here.
 *
 *   pack_sizes(texts, offsets_addr) -> total bytes     offsets: uint64[len(texts) + 1], written here
 *   pack_copy(texts, dst_addr)      -> None            copies every text's UTF-8 bytes to dst (capacity = that total)
 *   ids_to_lists(ids_addr, offsets_addr, n_docs) -> list[list[int]]      ids uint32, offsets uint64[n_docs + 1]
 *
 * A non-str item raises TypeError with PyO3's wording; a lone surrogate raises UnicodeEncodeError (as str.encode and
 * PyO3's extraction do).  PyUnicode_AsUTF8AndSize returns the string's own buffer for ASCII text (no copy, no
 * allocation) and caches the UTF-8 form otherwise. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static PyObject* not_a_str(PyObject* item) {
    PyErr_Format(PyExc_TypeError, "argument 'texts': '%s' object cannot be converted to 'PyString'", Py_TYPE(item)->tp_name);
    return NULL;
}

static PyObject* pack_sizes(PyObject* self, PyObject* args) {
    PyObject* texts; unsigned long long off_addr;
    if (!PyArg_ParseTuple(args, "OK", &texts, &off_addr)) return NULL;
    if (!PyList_Check(texts)) { PyErr_SetString(PyExc_TypeError, "texts must be a list"); return NULL; }
    uint64_t* off = (uint64_t*)(uintptr_t)off_addr;
    const Py_ssize_t n = PyList_GET_SIZE(texts);
    uint64_t run = 0;
    off[0] = 0;
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject* it = PyList_GET_ITEM(texts, i);
        if (!PyUnicode_Check(it)) return not_a_str(it);
        Py_ssize_t len;
        if (!PyUnicode_AsUTF8AndSize(it, &len)) return NULL;
        run += (uint64_t)len;
        off[i + 1] = run;
    }
    return PyLong_FromUnsignedLongLong(run);
}

static PyObject* pack_copy(PyObject* self, PyObject* args) {
    PyObject* texts; unsigned long long dst_addr;
    if (!PyArg_ParseTuple(args, "OK", &texts, &dst_addr)) return NULL;
    if (!PyList_Check(texts)) { PyErr_SetString(PyExc_TypeError, "texts must be a list"); return NULL; }
    uint8_t* dst = (uint8_t*)(uintptr_t)dst_addr;
    const Py_ssize_t n = PyList_GET_SIZE(texts);
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject* it = PyList_GET_ITEM(texts, i);
        if (!PyUnicode_Check(it)) return not_a_str(it);
        Py_ssize_t len;
        const char* p = PyUnicode_AsUTF8AndSize(it, &len);
        if (!p) return NULL;
        memcpy(dst, p, (size_t)len);
        dst += len;
    }
    Py_RETURN_NONE;
}

/* One int object per token id, made on first use and shared by every list after that (ints are immutable, so nobody can
 * tell): a list entry then costs a reference count and a pointer instead of an allocation.  28 M PyLong_FromUnsignedLong
 * calls per 100 MB of English text were the whole cost of encode_batch's list[list[int]] result. */
#define ID_CACHE_MAX (1u << 21)                 /* token ids are below this (21-bit symbols); larger ones get fresh objects */
static PyObject** g_id_cache = NULL;

static PyObject* ids_to_lists(PyObject* self, PyObject* args) {
    unsigned long long ids_addr, off_addr; Py_ssize_t n_docs;
    if (!PyArg_ParseTuple(args, "KKn", &ids_addr, &off_addr, &n_docs)) return NULL;
    const uint32_t* ids = (const uint32_t*)(uintptr_t)ids_addr;
    const uint64_t* off = (const uint64_t*)(uintptr_t)off_addr;
    if (!g_id_cache) {
        g_id_cache = (PyObject**)calloc(ID_CACHE_MAX, sizeof(PyObject*));
        if (!g_id_cache) return PyErr_NoMemory();
    }
    PyObject* out = PyList_New(n_docs);
    if (!out) return NULL;
    /* lists of ints cannot form cycles: no collections while 100 000 of them are being made (each young collection
     * would walk the items of the lists made since the last one) */
    const int gc_was_on = PyGC_Disable();
    for (Py_ssize_t d = 0; d < n_docs; ++d) {
        const uint64_t lo = off[d], hi = off[d + 1];
        PyObject* row = PyList_New((Py_ssize_t)(hi - lo));
        if (!row) { Py_DECREF(out); if (gc_was_on) PyGC_Enable(); return NULL; }
        for (uint64_t k = lo; k < hi; ++k) {
            const uint32_t id = ids[k];
            PyObject* v;
            if (id < ID_CACHE_MAX) {
                v = g_id_cache[id];
                if (!v) {
                    v = PyLong_FromUnsignedLong(id);
                    if (!v) { Py_DECREF(row); Py_DECREF(out); if (gc_was_on) PyGC_Enable(); return NULL; }
                    g_id_cache[id] = v;                       /* the cache keeps its own reference for good */
                }
                Py_INCREF(v);
            } else {
                v = PyLong_FromUnsignedLong(id);
                if (!v) { Py_DECREF(row); Py_DECREF(out); if (gc_was_on) PyGC_Enable(); return NULL; }
            }
            PyList_SET_ITEM(row, (Py_ssize_t)(k - lo), v);
        }
        PyList_SET_ITEM(out, d, row);
    }
    if (gc_was_on) PyGC_Enable();
    return out;
}

static PyMethodDef methods[] = {
    {"pack_sizes", pack_sizes, METH_VARARGS, "byte offsets of every text's UTF-8 form; returns the total"},
    {"pack_copy", pack_copy, METH_VARARGS, "copy every text's UTF-8 bytes to one buffer"},
    {"ids_to_lists", ids_to_lists, METH_VARARGS, "id stream + offsets -> list[list[int]]"},
    {NULL, NULL, 0, NULL}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_pyhost", NULL, -1, methods};
PyMODINIT_FUNC PyInit__pyhost(void) { return PyModule_Create(&moddef); }
