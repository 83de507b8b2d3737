/*
 * TEST INFRASTRUCTURE ONLY -- never linked into the product (reveal_b200/).
 *
 * Python-3 module shell around the UNMODIFIED reference extension
 * (/root/reference/reveallib/interface.c + reveal.c + divsufsort/), compiled by
 * path by oracle/ref/Makefile into oracle/_ref/.  The reference's own
 * `initreveallib[64]()` (interface.c:894-937) still registers its own type
 * `index` and its own `error`; this file only supplies the three py2 C-API
 * entry points that no longer exist (see compat.h) and the py3 PyInit_ hook.
 * Compiled WITHOUT compat.h.
 */
#include <Python.h>
#include <stdarg.h>
#include <string.h>

#ifdef SA64
extern void initreveallib64(void);
#define REF_INIT initreveallib64
#define REF_PYINIT PyInit__reveallib64_ref
#define REF_NAME "_reveallib64_ref"
#else
extern void initreveallib(void);
#define REF_INIT initreveallib
#define REF_PYINIT PyInit__reveallib_ref
#define REF_NAME "_reveallib_ref"
#endif

static PyObject *g_mod = NULL;
static struct PyModuleDef g_def = {PyModuleDef_HEAD_INIT, REF_NAME, NULL, -1, NULL};

PyObject *compat_InitModule3(const char *name, PyMethodDef *methods, const char *doc)
{
    (void)name;
    g_def.m_doc = doc;
    g_def.m_methods = methods;
    g_mod = PyModule_Create(&g_def);
    return g_mod;
}

/* PyArg_ParseTuple stand-in: identical to the real one except for the py2
 * "s#" + int* spelling used by addsequence (interface.c:55-59). */
int compat_ParseTuple(PyObject *args, const char *fmt, ...)
{
    va_list va;
    int ok = 0;
    va_start(va, fmt);
    if (strcmp(fmt, "s#") == 0) {
        char **s = va_arg(va, char **);
        int *l = va_arg(va, int *);
        PyObject *o = (PyTuple_Check(args) && PyTuple_Size(args) == 1) ? PyTuple_GetItem(args, 0) : NULL;
        Py_ssize_t sz = 0;
        const char *p = NULL;
        if (o && PyUnicode_Check(o)) {
            p = PyUnicode_AsUTF8AndSize(o, &sz);
        } else if (o && PyBytes_Check(o)) {
            p = PyBytes_AsString(o);
            sz = PyBytes_Size(o);
        } else {
            PyErr_SetString(PyExc_TypeError, "addsequence expects one str/bytes argument");
        }
        if (p) {
            *s = (char *)p;
            *l = (int)sz;
            ok = 1;
        }
    } else {
        ok = PyArg_VaParse(args, fmt, va);
    }
    va_end(va);
    return ok;
}

PyMODINIT_FUNC REF_PYINIT(void)
{
    REF_INIT();
    if (g_mod) {
        /* the reference's type object is static with refcount 0 (see compat.h):
         * leak one reference so module teardown never deallocates it */
        PyObject *t = PyObject_GetAttrString(g_mod, "index");
        (void)t;
    }
    return g_mod;
}
