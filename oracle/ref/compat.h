/*
 * TEST INFRASTRUCTURE ONLY -- never linked into the product (reveal_b200/).
 *
 * Force-included (gcc -include) in front of the UNMODIFIED reference sources
 * /root/reference/reveallib/{interface.c,reveal.c} so that they compile against
 * CPython 3.12 (the reference is written for Python 2).  Nothing here changes
 * the arithmetic of the reference; it only renames Python-2 C-API spellings:
 *
 *   PyString_Check / PyInt_AS_LONG      interface.c:29, reveal.c:923
 *   PyObject_HEAD_INIT(NULL) 0,         interface.c:842-843 (py2 type header layout)
 *   Py_InitModule3                      interface.c:902,925
 *   PyMODINIT_FUNC void + bare return;  interface.c:894-937
 *   "s#" without PY_SSIZE_T_CLEAN       interface.c:58
 */
#ifndef REVEAL_REF_COMPAT_H
#define REVEAL_REF_COMPAT_H
#include <Python.h>
#include <ctype.h>
#include <assert.h>
#include <stdint.h>
#include <limits.h>

#define PyString_Check PyUnicode_Check
#define PyInt_AS_LONG PyLong_AsLong

/* py2 spelled the static type header as `PyObject_HEAD_INIT(NULL) 0,` (two
 * positional fields).  In py3 the header is one nested struct, so turn the
 * pair into a designated initialiser of ob_size; positional initialisers
 * that follow continue with tp_name. ob_type is filled by PyType_Ready. */
#undef PyObject_HEAD_INIT
#define PyObject_HEAD_INIT(x) .ob_base.ob_size =

#undef PyMODINIT_FUNC
#define PyMODINIT_FUNC void

PyObject *compat_InitModule3(const char *name, PyMethodDef *methods, const char *doc);
#define Py_InitModule3 compat_InitModule3

int compat_ParseTuple(PyObject *args, const char *fmt, ...);
#define PyArg_ParseTuple compat_ParseTuple

#endif
