/* fastbuild.c -- C-level construction of the reference's result containers (SURVEY 8f N3 / DESIGN section 10): the
 * list of Op NamedTuples of one stage from the flat op tables the C ABI returns.  What the reference's binding does in
 * C++ (bindings.cc:106-151: one Python call per op) and what da4ml_b200.types.pipeline_from_arrays does in a Python
 * comprehension (~2 us per op), here ~0.3 us per op: tuple subclasses are allocated and filled directly.
 * Optional acceleration of host-side object construction only -- no solver arithmetic here. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

static PyObject *alloc_tuple(PyTypeObject *tp, Py_ssize_t n) {
    return tp->tp_alloc(tp, n); /* a tuple subclass instance with n NULL items */
}

/* build_ops(ops_i: int64[n,4] buffer, ops_f: float32[n,5] buffer, Op, QInterval) -> list[Op] */
static PyObject *build_ops(PyObject *self, PyObject *args) {
    PyObject *oi_obj, *of_obj, *op_t, *q_t;
    if (!PyArg_ParseTuple(args, "OOOO", &oi_obj, &of_obj, &op_t, &q_t))
        return NULL;
    if (!PyType_Check(op_t) || !PyType_Check(q_t) || !PyType_IsSubtype((PyTypeObject *)op_t, &PyTuple_Type) ||
        !PyType_IsSubtype((PyTypeObject *)q_t, &PyTuple_Type)) {
        PyErr_SetString(PyExc_TypeError, "Op and QInterval must be tuple subclasses");
        return NULL;
    }
    Py_buffer bi, bf;
    if (PyObject_GetBuffer(oi_obj, &bi, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) < 0)
        return NULL;
    if (PyObject_GetBuffer(of_obj, &bf, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) < 0) {
        PyBuffer_Release(&bi);
        return NULL;
    }
    PyObject *list = NULL;
    if (bi.itemsize != 8 || bf.itemsize != 4 || bi.len % 32 != 0 || bf.len % 20 != 0 || bi.len / 32 != bf.len / 20) {
        PyErr_SetString(PyExc_ValueError, "expected int64 [n,4] and float32 [n,5]");
        goto done;
    }
    {
        const Py_ssize_t n = bi.len / 32;
        const int64_t *oi = (const int64_t *)bi.buf;
        const float *of = (const float *)bf.buf;
        list = PyList_New(n);
        if (!list)
            goto done;
        for (Py_ssize_t k = 0; k < n; ++k) {
            PyObject *q = alloc_tuple((PyTypeObject *)q_t, 3), *op = alloc_tuple((PyTypeObject *)op_t, 7);
            if (!q || !op) {
                Py_XDECREF(q);
                Py_XDECREF(op);
                Py_CLEAR(list);
                goto done;
            }
            PyTuple_SET_ITEM(q, 0, PyFloat_FromDouble((double)of[5 * k + 0]));
            PyTuple_SET_ITEM(q, 1, PyFloat_FromDouble((double)of[5 * k + 1]));
            PyTuple_SET_ITEM(q, 2, PyFloat_FromDouble((double)of[5 * k + 2]));
            PyTuple_SET_ITEM(op, 0, PyLong_FromLongLong(oi[4 * k + 0]));
            PyTuple_SET_ITEM(op, 1, PyLong_FromLongLong(oi[4 * k + 1]));
            PyTuple_SET_ITEM(op, 2, PyLong_FromLongLong(oi[4 * k + 2]));
            PyTuple_SET_ITEM(op, 3, PyLong_FromLongLong(oi[4 * k + 3]));
            PyTuple_SET_ITEM(op, 4, q);
            PyTuple_SET_ITEM(op, 5, PyFloat_FromDouble((double)of[5 * k + 3]));
            PyTuple_SET_ITEM(op, 6, PyFloat_FromDouble((double)of[5 * k + 4]));
            PyList_SET_ITEM(list, k, op);
        }
    }
done:
    PyBuffer_Release(&bi);
    PyBuffer_Release(&bf);
    return list;
}

static PyMethodDef methods[] = {{"build_ops", build_ops, METH_VARARGS, "list of Op tuples from flat op tables"}, {NULL, NULL, 0, NULL}};
static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_fastbuild", NULL, -1, methods};
PyMODINIT_FUNC PyInit__fastbuild(void) { return PyModule_Create(&module); }
