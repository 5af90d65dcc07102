/*
 * cvxopt_b200._misc_solvers — CPython extension mirroring CVXOPT's `cvxopt.misc_solvers`
 * (reference src/C/misc_solvers.c:1155-1173: the 12-entry method table) on top of the B200
 * library's C ABI (include/cvxopt_b200.h).  Same function names, same keyword lists, same
 * in-place semantics on `cvxopt.matrix` buffers, so an unmodified CVXOPT can be pointed at it:
 *
 *     import cvxopt_b200._misc_solvers as ms            # needs `import cvxopt` to succeed
 *     sys.modules['cvxopt.misc_solvers'] = ms           # before cvxopt.misc is first imported, or
 *     for f in ms.__all__: setattr(cvxopt.misc, f, getattr(ms, f))      # afterwards
 *
 * It talks to CVXOPT the way any third-party CVXOPT extension does: `import_cvxopt()` fetches the
 * "base_API" capsule of cvxopt.base (reference src/C/cvxopt.h:93-113, exported at base.c:2049-2067)
 * and matrix fields are read through the public struct (cvxopt.h:48-56, MAT_BUFD/nrows/ncols
 * :121-132).  The ABI is restated below (no reference header is needed to build this file).
 *
 * The arithmetic runs on the GPU (CVXB_HOST space: the library stages the buffers); there is no
 * CPU fallback — without a B200 every call raises RuntimeError.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/cvxopt_b200.h"

/* ---- CVXOPT's C API, restated (reference src/C/cvxopt.h:42-56, 76-113) ------------------- */
#define CVX_DOUBLE 1
typedef struct {
    PyObject_HEAD
    void *buffer;              /* column-major array of type `id` */
    int nrows, ncols;
    int id;                    /* 0 int, 1 double, 2 complex */
    Py_ssize_t shape[2];
    Py_ssize_t strides[2];
    Py_ssize_t ob_exports;
} cvx_matrix;

static void **cvx_api = NULL;
/* slot 3 of the capsule's table is Matrix_Check (cvxopt.h:79) */
#define CVX_MATRIX_CHECK(o) ((*(int (*)(void *))cvx_api[3])((void *)(o)))
#define MBUF(o) ((double *)((cvx_matrix *)(o))->buffer)
#define MLEN(o) ((Py_ssize_t)((cvx_matrix *)(o))->nrows * (Py_ssize_t)((cvx_matrix *)(o))->ncols)

static int import_cvxopt(void)
{
    PyObject *mod = PyImport_ImportModule("cvxopt.base");
    if (!mod) return -1;
    PyObject *cap = PyObject_GetAttrString(mod, "_C_API");
    Py_DECREF(mod);
    if (!cap) return -1;
    if (!PyCapsule_IsValid(cap, "base_API")) {
        Py_DECREF(cap);
        PyErr_SetString(PyExc_ImportError, "cvxopt.base._C_API is not a 'base_API' capsule");
        return -1;
    }
    cvx_api = (void **)PyCapsule_GetPointer(cap, "base_API");
    Py_DECREF(cap);
    return cvx_api ? 0 : -1;
}

/* ---- helpers ------------------------------------------------------------------------------ */
static int is_dmat(PyObject *o) { return o && CVX_MATRIX_CHECK(o) && ((cvx_matrix *)o)->id == CVX_DOUBLE; }

static PyObject *need_dmat(const char *name)
{
    PyErr_Format(PyExc_TypeError, "%s must be a matrix with typecode 'd'", name);
    return NULL;
}

/* map a C-ABI return code onto the exception the reference raises for the same condition */
static PyObject *raise_rc(int rc, const char *what)
{
    const char *msg = cvxb_last_error();
    if (rc > 0) { PyErr_SetObject(PyExc_ArithmeticError, PyLong_FromLong(rc)); return NULL; }
    PyObject *exc = PyExc_RuntimeError;
    if (rc == CVXB_E_ARG) exc = PyExc_ValueError;
    else if (rc == CVXB_E_NOMEM) exc = PyExc_MemoryError;
    else if (rc == CVXB_E_UNSUP) exc = PyExc_NotImplementedError;
    PyErr_Format(exc, "%s: %s", what, msg ? msg : "");
    return NULL;
}

typedef struct {
    cvxb_dims d;
    int *q, *s;
    Py_ssize_t cdim, cdim_pckd, nlam, sums;    /* unpacked / packed cone dimension, length of lambda */
} dims_t;

static void dims_free(dims_t *t) { free(t->q); free(t->s); t->q = t->s = NULL; }

/* dims dict {'l': int, 'q': [int], 's': [int]} (+ mnl) -> cvxb_dims */
static int dims_parse(PyObject *dims, int mnl, dims_t *t)
{
    memset(t, 0, sizeof(*t));
    if (!PyDict_Check(dims)) { PyErr_SetString(PyExc_TypeError, "dims must be a dictionary"); return -1; }
    PyObject *l = PyDict_GetItemString(dims, "l"), *q = PyDict_GetItemString(dims, "q"),
             *s = PyDict_GetItemString(dims, "s");
    if (!l || !q || !s) { PyErr_SetString(PyExc_KeyError, "dims must have keys 'l', 'q' and 's'"); return -1; }
    if (!PyList_Check(q) || !PyList_Check(s)) {
        PyErr_SetString(PyExc_TypeError, "dims['q'] and dims['s'] must be lists");
        return -1;
    }
    long ml = PyLong_AsLong(l);
    if (ml == -1 && PyErr_Occurred()) return -1;
    if (ml < 0 || mnl < 0) { PyErr_SetString(PyExc_ValueError, "dims['l'] and mnl must be nonnegative"); return -1; }
    Py_ssize_t nq = PyList_GET_SIZE(q), ns = PyList_GET_SIZE(s);
    t->q = (int *)calloc((size_t)(nq > 0 ? nq : 1), sizeof(int));
    t->s = (int *)calloc((size_t)(ns > 0 ? ns : 1), sizeof(int));
    if (!t->q || !t->s) { dims_free(t); PyErr_NoMemory(); return -1; }
    Py_ssize_t sq = 0, s2 = 0, sp = 0, ss = 0;
    for (Py_ssize_t k = 0; k < nq; ++k) {
        long v = PyLong_AsLong(PyList_GET_ITEM(q, k));
        if (v < 1) { if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "dims['q'] must be positive integers"); dims_free(t); return -1; }
        t->q[k] = (int)v; sq += v;
    }
    for (Py_ssize_t k = 0; k < ns; ++k) {
        long v = PyLong_AsLong(PyList_GET_ITEM(s, k));
        if (v < 0) { if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "dims['s'] must be nonnegative integers"); dims_free(t); return -1; }
        t->s[k] = (int)v; s2 += v * v; sp += v * (v + 1) / 2; ss += v;
    }
    t->d.mnl = mnl; t->d.ml = (int)ml; t->d.nq = (int)nq; t->d.q = t->q; t->d.ns = (int)ns; t->d.s = t->s;
    t->cdim = mnl + ml + sq + s2;
    t->cdim_pckd = mnl + ml + sq + sp;
    t->nlam = mnl + ml + nq + ss;
    t->sums = ss;
    return 0;
}

/* the scaling dict W (coneprog.py:327-334; 'dnl'/'dnli' in cvxprog, misc.py:45-56) -> flat arrays.
 * Sizes come from W itself, as in the reference's scale() (misc_solvers.c:117-214). */
typedef struct {
    cvxb_scaling w;
    dims_t dm;
    double *v, *beta, *r, *rti;      /* owned concatenations */
} scal_t;

static void scal_free(scal_t *s)
{
    free(s->v); free(s->beta); free(s->r); free(s->rti);
    dims_free(&s->dm);
    memset(s, 0, sizeof(*s));
}

static int scal_parse(PyObject *W, scal_t *out)
{
    memset(out, 0, sizeof(*out));
    if (!PyDict_Check(W)) { PyErr_SetString(PyExc_TypeError, "W must be a dictionary"); return -1; }
    PyObject *dnl = PyDict_GetItemString(W, "dnl"), *dnli = PyDict_GetItemString(W, "dnli");
    PyObject *d = PyDict_GetItemString(W, "d"), *di = PyDict_GetItemString(W, "di");
    PyObject *v = PyDict_GetItemString(W, "v"), *beta = PyDict_GetItemString(W, "beta");
    PyObject *r = PyDict_GetItemString(W, "r"), *rti = PyDict_GetItemString(W, "rti");
    if (!d || !di) { PyErr_SetString(PyExc_KeyError, "missing item W['d'] or W['di']"); return -1; }   /* :134 */
    if (!is_dmat(d) || !is_dmat(di)) { need_dmat("W['d'] and W['di']"); return -1; }
    int mnl = 0;
    if (dnl) {
        if (!is_dmat(dnl) || !dnli || !is_dmat(dnli)) { need_dmat("W['dnl'] and W['dnli']"); return -1; }
        mnl = (int)MLEN(dnl);
    }
    if (!v || !beta || !r || !rti || !PyList_Check(v) || !PyList_Check(beta) || !PyList_Check(r) ||
        !PyList_Check(rti)) {
        PyErr_SetString(PyExc_KeyError, "W must have list items 'v', 'beta', 'r' and 'rti'");
        return -1;
    }
    Py_ssize_t nq = PyList_GET_SIZE(v), ns = PyList_GET_SIZE(r);
    if (PyList_GET_SIZE(beta) != nq || PyList_GET_SIZE(rti) != ns) {
        PyErr_SetString(PyExc_ValueError, "W['beta'] / W['rti'] do not match W['v'] / W['r']");
        return -1;
    }
    dims_t *dm = &out->dm;
    dm->q = (int *)calloc((size_t)(nq > 0 ? nq : 1), sizeof(int));
    dm->s = (int *)calloc((size_t)(ns > 0 ? ns : 1), sizeof(int));
    if (!dm->q || !dm->s) { scal_free(out); PyErr_NoMemory(); return -1; }
    Py_ssize_t sq = 0, s2 = 0, sp = 0, ss = 0;
    for (Py_ssize_t k = 0; k < nq; ++k) {
        PyObject *vk = PyList_GET_ITEM(v, k);
        if (!is_dmat(vk)) { scal_free(out); need_dmat("W['v'][k]"); return -1; }
        dm->q[k] = (int)MLEN(vk); sq += dm->q[k];
    }
    for (Py_ssize_t k = 0; k < ns; ++k) {
        PyObject *rk = PyList_GET_ITEM(r, k), *tk = PyList_GET_ITEM(rti, k);
        if (!is_dmat(rk) || !is_dmat(tk)) { scal_free(out); need_dmat("W['r'][k] and W['rti'][k]"); return -1; }
        int m = ((cvx_matrix *)rk)->nrows;
        if (((cvx_matrix *)rk)->ncols != m || MLEN(tk) != (Py_ssize_t)m * m) {
            scal_free(out);
            PyErr_SetString(PyExc_ValueError, "W['r'][k] and W['rti'][k] must be square and of equal order");
            return -1;
        }
        dm->s[k] = m; s2 += (Py_ssize_t)m * m; sp += (Py_ssize_t)m * (m + 1) / 2; ss += m;
    }
    const Py_ssize_t ml = MLEN(d);
    if (MLEN(di) != ml || (dnl && MLEN(dnli) != mnl)) {
        scal_free(out);
        PyErr_SetString(PyExc_ValueError, "W['d'] / W['di'] (or W['dnl'] / W['dnli']) differ in length");
        return -1;
    }
    dm->d.mnl = mnl; dm->d.ml = (int)ml; dm->d.nq = (int)nq; dm->d.q = dm->q; dm->d.ns = (int)ns; dm->d.s = dm->s;
    dm->cdim = mnl + ml + sq + s2; dm->cdim_pckd = mnl + ml + sq + sp; dm->nlam = mnl + ml + nq + ss; dm->sums = ss;
    out->v = (double *)malloc(sizeof(double) * (size_t)(sq > 0 ? sq : 1));
    out->beta = (double *)malloc(sizeof(double) * (size_t)(nq > 0 ? nq : 1));
    out->r = (double *)malloc(sizeof(double) * (size_t)(s2 > 0 ? s2 : 1));
    out->rti = (double *)malloc(sizeof(double) * (size_t)(s2 > 0 ? s2 : 1));
    if (!out->v || !out->beta || !out->r || !out->rti) { scal_free(out); PyErr_NoMemory(); return -1; }
    Py_ssize_t o = 0;
    for (Py_ssize_t k = 0; k < nq; ++k) {
        PyObject *vk = PyList_GET_ITEM(v, k);
        memcpy(out->v + o, MBUF(vk), sizeof(double) * (size_t)dm->q[k]);
        o += dm->q[k];
        out->beta[k] = PyFloat_AsDouble(PyList_GET_ITEM(beta, k));
        if (out->beta[k] == -1.0 && PyErr_Occurred()) { scal_free(out); return -1; }
    }
    o = 0;
    for (Py_ssize_t k = 0; k < ns; ++k) {
        const size_t cnt = (size_t)dm->s[k] * (size_t)dm->s[k];
        memcpy(out->r + o, MBUF(PyList_GET_ITEM(r, k)), sizeof(double) * cnt);
        memcpy(out->rti + o, MBUF(PyList_GET_ITEM(rti, k)), sizeof(double) * cnt);
        o += (Py_ssize_t)cnt;
    }
    out->w.dnl = dnl ? MBUF(dnl) : NULL; out->w.dnli = dnl ? MBUF(dnli) : NULL;
    out->w.d = MBUF(d); out->w.di = MBUF(di);
    out->w.v = out->v; out->w.beta = out->beta; out->w.r = out->r; out->w.rti = out->rti;
    return 0;
}

static int flag_of(int c, const char *name, const char *allowed)
{
    if (!strchr(allowed, c) || c == 0) {
        PyErr_Format(PyExc_ValueError, "possible values of %s are: %s", name, allowed);
        return -1;
    }
    return c;
}

/* ---- the 12 functions ----------------------------------------------------------------------- */
static const char doc_scale[] =
    "scale(x, W, trans = 'N', inverse = 'N')\n\n"
    "In place x := W*x ('N','N'), W'*x ('T','N'), W^{-1}*x ('N','I'), W^{-T}*x ('T','I') for the\n"
    "Nesterov-Todd scaling W; x is a 'd' matrix whose columns are cone vectors.  Mirrors\n"
    "cvxopt.misc_solvers.scale (reference src/C/misc_solvers.c:85-244); computed on the B200.";
static PyObject *ms_scale(PyObject *self, PyObject *args, PyObject *kw)
{
    static char *kwlist[] = {"x", "W", "trans", "inverse", NULL};
    PyObject *x, *W;
    int trans = 'N', inverse = 'N';
    if (!PyArg_ParseTupleAndKeywords(args, kw, "OO|CC", kwlist, &x, &W, &trans, &inverse)) return NULL;
    if (!is_dmat(x)) return need_dmat("x");
    if (flag_of(trans, "trans", "NT") < 0 || flag_of(inverse, "inverse", "NI") < 0) return NULL;
    scal_t sc;
    if (scal_parse(W, &sc) < 0) return NULL;
    cvx_matrix *xm = (cvx_matrix *)x;
    if (xm->nrows < sc.dm.cdim) {
        scal_free(&sc);
        PyErr_SetString(PyExc_TypeError, "length of x is too small");
        return NULL;
    }
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = cvxb_scale(MBUF(x), xm->nrows, xm->ncols, &sc.dm.d, &sc.w, trans, inverse, CVXB_HOST);
    Py_END_ALLOW_THREADS
    scal_free(&sc);
    if (rc) return raise_rc(rc, "scale");
    Py_RETURN_NONE;
}

static const char doc_scale2[] =
    "scale2(lmbda, x, dims, mnl = 0, inverse = 'N')\n\n"
    "x := H(lambda^{1/2}) * x ('N') or H(lambda^{-1/2}) * x ('I'), in place.\n"
    "Mirrors cvxopt.misc_solvers.scale2 (reference src/C/misc_solvers.c:256-401).";
static PyObject *ms_scale2(PyObject *self, PyObject *args, PyObject *kw)
{
    static char *kwlist[] = {"lmbda", "x", "dims", "mnl", "inverse", NULL};
    PyObject *lm, *x, *dims;
    int mnl = 0, inverse = 'N';
    if (!PyArg_ParseTupleAndKeywords(args, kw, "OOO|iC", kwlist, &lm, &x, &dims, &mnl, &inverse)) return NULL;
    if (!is_dmat(lm)) return need_dmat("lmbda");
    if (!is_dmat(x)) return need_dmat("x");
    if (flag_of(inverse, "inverse", "NI") < 0) return NULL;
    dims_t t;
    if (dims_parse(dims, mnl, &t) < 0) return NULL;
    if (MLEN(lm) < t.nlam || MLEN(x) < t.cdim) {
        dims_free(&t);
        PyErr_SetString(PyExc_TypeError, "length of lmbda or x is too small");
        return NULL;
    }
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = cvxb_scale2(MBUF(lm), MBUF(x), &t.d, inverse, CVXB_HOST);     /* one cone vector, as the reference */
    Py_END_ALLOW_THREADS
    dims_free(&t);
    if (rc) return raise_rc(rc, "scale2");
    Py_RETURN_NONE;
}

static PyObject *pack_common(PyObject *args, PyObject *kw, int do_pack)
{
    static char *kwlist[] = {"x", "y", "dims", "mnl", "offsetx", "offsety", NULL};
    PyObject *x, *y, *dims;
    int mnl = 0, ox = 0, oy = 0;
    if (!PyArg_ParseTupleAndKeywords(args, kw, "OOO|iii", kwlist, &x, &y, &dims, &mnl, &ox, &oy)) return NULL;
    if (!is_dmat(x)) return need_dmat("x");
    if (!is_dmat(y)) return need_dmat("y");
    dims_t t;
    if (dims_parse(dims, mnl, &t) < 0) return NULL;
    const Py_ssize_t nx = do_pack ? t.cdim : t.cdim_pckd, ny = do_pack ? t.cdim_pckd : t.cdim;
    if (ox < 0 || oy < 0 || MLEN(x) < ox + nx || MLEN(y) < oy + ny) {
        dims_free(&t);
        PyErr_SetString(PyExc_TypeError, "length of x or y is too small");
        return NULL;
    }
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = do_pack ? cvxb_pack(MBUF(x) + ox, MBUF(y) + oy, &t.d, CVXB_HOST)
                 : cvxb_unpack(MBUF(x) + ox, MBUF(y) + oy, &t.d, CVXB_HOST);
    Py_END_ALLOW_THREADS
    dims_free(&t);
    if (rc) return raise_rc(rc, do_pack ? "pack" : "unpack");
    Py_RETURN_NONE;
}
static const char doc_pack[] =
    "pack(x, y, dims, mnl = 0, offsetx = 0, offsety = 0)\n\n"
    "Copies x to y with the 's' blocks in packed lower storage, off-diagonal entries scaled by\n"
    "sqrt(2).  Mirrors cvxopt.misc_solvers.pack (reference src/C/misc_solvers.c:412-465).";
static PyObject *ms_pack(PyObject *self, PyObject *args, PyObject *kw) { return pack_common(args, kw, 1); }
static const char doc_unpack[] =
    "unpack(x, y, dims, mnl = 0, offsetx = 0, offsety = 0)\n\n"
    "The inverse of pack: packed x to unpacked 'L' storage in y, off-diagonals scaled by 1/sqrt(2).\n"
    "Mirrors cvxopt.misc_solvers.unpack (reference src/C/misc_solvers.c:552-601).";
static PyObject *ms_unpack(PyObject *self, PyObject *args, PyObject *kw) { return pack_common(args, kw, 0); }

static const char doc_pack2[] =
    "pack2(x, dims, mnl = 0)\n\n"
    "In-place version of pack on the columns of x; the rows are compacted to the packed cone\n"
    "dimension.  Mirrors cvxopt.misc_solvers.pack2 (reference src/C/misc_solvers.c:476-541).";
static PyObject *ms_pack2(PyObject *self, PyObject *args, PyObject *kw)
{
    static char *kwlist[] = {"x", "dims", "mnl", NULL};
    PyObject *x, *dims;
    int mnl = 0;
    if (!PyArg_ParseTupleAndKeywords(args, kw, "OO|i", kwlist, &x, &dims, &mnl)) return NULL;
    if (!is_dmat(x)) return need_dmat("x");
    dims_t t;
    if (dims_parse(dims, mnl, &t) < 0) return NULL;
    cvx_matrix *xm = (cvx_matrix *)x;
    if (xm->nrows < t.cdim) {
        dims_free(&t);
        PyErr_SetString(PyExc_TypeError, "length of x is too small");
        return NULL;
    }
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = cvxb_pack2(MBUF(x), xm->nrows, xm->ncols, &t.d, CVXB_HOST);
    Py_END_ALLOW_THREADS
    dims_free(&t);
    if (rc) return raise_rc(rc, "pack2");
    Py_RETURN_NONE;
}

static const char doc_symm[] =
    "symm(x, n, offset = 0)\n\n"
    "Fills the upper triangle of the n x n matrix stored at x[offset:] from its lower triangle.\n"
    "Mirrors cvxopt.misc_solvers.symm (reference src/C/misc_solvers.c:610-625).";
static PyObject *ms_symm(PyObject *self, PyObject *args, PyObject *kw)
{
    static char *kwlist[] = {"x", "n", "offset", NULL};
    PyObject *x;
    int n, off = 0;
    if (!PyArg_ParseTupleAndKeywords(args, kw, "Oi|i", kwlist, &x, &n, &off)) return NULL;
    if (!is_dmat(x)) return need_dmat("x");
    if (n < 0 || off < 0 || MLEN(x) < off + (Py_ssize_t)n * n) {
        PyErr_SetString(PyExc_TypeError, "length of x is too small");
        return NULL;
    }
    if (n > 1) {
        int rc;
        Py_BEGIN_ALLOW_THREADS
        rc = cvxb_symm(MBUF(x) + off, n, CVXB_HOST);
        Py_END_ALLOW_THREADS
        if (rc) return raise_rc(rc, "symm");
    }
    Py_RETURN_NONE;
}

static const char doc_sprod[] =
    "sprod(x, y, dims, mnl = 0, diag = 'N')\n\n"
    "The cone product x := y o x; with diag = 'D' the 's' part of y is diagonal and stored as a\n"
    "vector.  Mirrors cvxopt.misc_solvers.sprod (reference src/C/misc_solvers.c:634-767).";
static PyObject *ms_sprod(PyObject *self, PyObject *args, PyObject *kw)
{
    static char *kwlist[] = {"x", "y", "dims", "mnl", "diag", NULL};
    PyObject *x, *y, *dims;
    int mnl = 0, diag = 'N';
    if (!PyArg_ParseTupleAndKeywords(args, kw, "OOO|iC", kwlist, &x, &y, &dims, &mnl, &diag)) return NULL;
    if (!is_dmat(x)) return need_dmat("x");
    if (!is_dmat(y)) return need_dmat("y");
    if (flag_of(diag, "diag", "ND") < 0) return NULL;
    dims_t t;
    if (dims_parse(dims, mnl, &t) < 0) return NULL;
    if (MLEN(x) < t.cdim || MLEN(y) < (diag == 'D' ? t.nlam : t.cdim)) {
        dims_free(&t);
        PyErr_SetString(PyExc_TypeError, "length of x or y is too small");
        return NULL;
    }
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = cvxb_sprod(MBUF(x), MBUF(y), &t.d, diag, CVXB_HOST);
    Py_END_ALLOW_THREADS
    dims_free(&t);
    if (rc) return raise_rc(rc, "sprod");
    Py_RETURN_NONE;
}

static const char doc_sinv[] =
    "sinv(x, y, dims, mnl = 0)\n\n"
    "The inverse of the cone product, x := y o\\ x, with the 's' components of y diagonal (stored as\n"
    "a vector).  Mirrors cvxopt.misc_solvers.sinv (reference src/C/misc_solvers.c:775-878).";
static PyObject *ms_sinv(PyObject *self, PyObject *args, PyObject *kw)
{
    static char *kwlist[] = {"x", "y", "dims", "mnl", NULL};
    PyObject *x, *y, *dims;
    int mnl = 0;
    if (!PyArg_ParseTupleAndKeywords(args, kw, "OOO|i", kwlist, &x, &y, &dims, &mnl)) return NULL;
    if (!is_dmat(x)) return need_dmat("x");
    if (!is_dmat(y)) return need_dmat("y");
    dims_t t;
    if (dims_parse(dims, mnl, &t) < 0) return NULL;
    if (MLEN(x) < t.cdim || MLEN(y) < t.nlam) {
        dims_free(&t);
        PyErr_SetString(PyExc_TypeError, "length of x or y is too small");
        return NULL;
    }
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = cvxb_sinv(MBUF(x), MBUF(y), &t.d, CVXB_HOST);
    Py_END_ALLOW_THREADS
    dims_free(&t);
    if (rc) return raise_rc(rc, "sinv");
    Py_RETURN_NONE;
}

static PyObject *trisc_common(PyObject *args, PyObject *kw, int undo)
{
    static char *kwlist[] = {"x", "dims", "offset", NULL};
    PyObject *x, *dims;
    int off = 0;
    if (!PyArg_ParseTupleAndKeywords(args, kw, "OO|i", kwlist, &x, &dims, &off)) return NULL;
    if (!is_dmat(x)) return need_dmat("x");
    dims_t t;
    if (dims_parse(dims, 0, &t) < 0) return NULL;
    if (off < 0 || MLEN(x) < off + t.cdim) {
        dims_free(&t);
        PyErr_SetString(PyExc_TypeError, "length of x is too small");
        return NULL;
    }
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = undo ? cvxb_triusc(MBUF(x) + off, &t.d, CVXB_HOST) : cvxb_trisc(MBUF(x) + off, &t.d, CVXB_HOST);
    Py_END_ALLOW_THREADS
    dims_free(&t);
    if (rc) return raise_rc(rc, undo ? "triusc" : "trisc");
    Py_RETURN_NONE;
}
static const char doc_trisc[] =
    "trisc(x, dims, offset = 0)\n\n"
    "Zeroes the strict upper triangles of the 's' blocks of x and doubles their strict lower\n"
    "triangles.  Mirrors cvxopt.misc_solvers.trisc (reference src/C/misc_solvers.c:887-935).";
static PyObject *ms_trisc(PyObject *self, PyObject *args, PyObject *kw) { return trisc_common(args, kw, 0); }
static const char doc_triusc[] =
    "triusc(x, dims, offset = 0)\n\n"
    "Undoes trisc: halves the strict lower triangles of the 's' blocks of x.\n"
    "Mirrors cvxopt.misc_solvers.triusc (reference src/C/misc_solvers.c:940-986).";
static PyObject *ms_triusc(PyObject *self, PyObject *args, PyObject *kw) { return trisc_common(args, kw, 1); }

static const char doc_sdot[] =
    "sdot(x, y, dims, mnl = 0)\n\n"
    "Inner product of two cone vectors ('s' blocks: trace inner product of the symmetric matrices\n"
    "given by their lower triangles).  Mirrors cvxopt.misc_solvers.sdot (src/C/misc_solvers.c:991-1039).";
static PyObject *ms_sdot(PyObject *self, PyObject *args, PyObject *kw)
{
    static char *kwlist[] = {"x", "y", "dims", "mnl", NULL};
    PyObject *x, *y, *dims;
    int mnl = 0;
    if (!PyArg_ParseTupleAndKeywords(args, kw, "OOO|i", kwlist, &x, &y, &dims, &mnl)) return NULL;
    if (!is_dmat(x)) return need_dmat("x");
    if (!is_dmat(y)) return need_dmat("y");
    dims_t t;
    if (dims_parse(dims, mnl, &t) < 0) return NULL;
    if (MLEN(x) < t.cdim || MLEN(y) < t.cdim) {
        dims_free(&t);
        PyErr_SetString(PyExc_TypeError, "length of x or y is too small");
        return NULL;
    }
    double out = 0.0;
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = cvxb_sdot(MBUF(x), MBUF(y), &t.d, &out, CVXB_HOST);
    Py_END_ALLOW_THREADS
    dims_free(&t);
    if (rc) return raise_rc(rc, "sdot");
    return PyFloat_FromDouble(out);
}

static const char doc_max_step[] =
    "max_step(x, dims, mnl = 0, sigma = None)\n\n"
    "Returns min {t | x + t*e >= 0} (e the identity of the cone).  With sigma (a 'd' matrix of length\n"
    "sum(dims['s'])) the eigenvalues of the 's' blocks go to sigma and their eigenvectors overwrite\n"
    "the blocks of x.  Mirrors cvxopt.misc_solvers.max_step (reference src/C/misc_solvers.c:1052-1153).";
static PyObject *ms_max_step(PyObject *self, PyObject *args, PyObject *kw)
{
    static char *kwlist[] = {"x", "dims", "mnl", "sigma", NULL};
    PyObject *x, *dims, *sigma = NULL;
    int mnl = 0;
    if (!PyArg_ParseTupleAndKeywords(args, kw, "OO|iO", kwlist, &x, &dims, &mnl, &sigma)) return NULL;
    if (!is_dmat(x)) return need_dmat("x");
    if (sigma == Py_None) sigma = NULL;
    if (sigma && !is_dmat(sigma)) return need_dmat("sigma");
    dims_t t;
    if (dims_parse(dims, mnl, &t) < 0) return NULL;
    if (MLEN(x) < t.cdim || (sigma && MLEN(sigma) < t.sums)) {
        dims_free(&t);
        PyErr_SetString(PyExc_TypeError, "length of x or sigma is too small");
        return NULL;
    }
    double out = 0.0;
    int rc;
    Py_BEGIN_ALLOW_THREADS
    rc = cvxb_max_step(MBUF(x), &t.d, sigma ? MBUF(sigma) : NULL, &out, CVXB_HOST);
    Py_END_ALLOW_THREADS
    dims_free(&t);
    if (rc) return raise_rc(rc, "max_step");
    return PyFloat_FromDouble(out);
}

#define ENTRY(name) {#name, (PyCFunction)(void (*)(void))ms_##name, METH_VARARGS | METH_KEYWORDS, doc_##name}
static PyMethodDef ms_methods[] = {
    ENTRY(scale), ENTRY(scale2), ENTRY(pack), ENTRY(pack2), ENTRY(unpack), ENTRY(symm), ENTRY(sprod),
    ENTRY(sinv), ENTRY(trisc), ENTRY(triusc), ENTRY(sdot), ENTRY(max_step), {NULL, NULL, 0, NULL}};

static struct PyModuleDef ms_module = {
    PyModuleDef_HEAD_INIT, "_misc_solvers",
    "B200-backed mirror of cvxopt.misc_solvers (reference src/C/misc_solvers.c): the cone algebra of\n"
    "CVXOPT's cone solvers executed by libcvxopt_b200.so on cvxopt.matrix buffers.",
    -1, ms_methods, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__misc_solvers(void)
{
    PyObject *m = PyModule_Create(&ms_module);
    if (!m) return NULL;
    if (import_cvxopt() < 0) { Py_DECREF(m); return NULL; }
    PyObject *all = Py_BuildValue("[ssssssssssss]", "scale", "scale2", "pack", "pack2", "unpack", "symm", "sprod",
                                  "sinv", "trisc", "triusc", "sdot", "max_step");
    if (!all || PyModule_AddObject(m, "__all__", all) < 0) { Py_XDECREF(all); Py_DECREF(m); return NULL; }
    return m;
}
