/*
 * _orph_refnative -- CPython extension that gives the UNMODIFIED reference python its two native dependencies at
 * native speed, for timing it on the bench box (bench.py --impl reference, kind "reference").
 *
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Nothing under orpheum_b200/ may import it.
 *
 * The reference calls sourmash.minhash.hash_murmur (a Rust/C extension function; orpheum/translate.py:187,
 * orpheum/index.py:118) and khmer.Nodegraph.add / .get (a C++ extension type; orpheum/index.py:101,119,
 * orpheum/translate.py:190) once per k-mer.  The pure-python stand-ins of oracle/shims are fine for golden vectors but
 * ten times too slow to be a fair clock.  This module wraps the C restatement in oracle/orpheum_oracle.c (pinned
 * against the reference's own fixtures by tests/test_oracle.py) as real extension objects, so the per-call cost is
 * what the pinned khmer / sourmash wheels would cost: one C call per k-mer.  oracle/shims_native/{khmer,sourmash}
 * re-export it under the names the reference imports.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

typedef struct orc_nodegraph orc_nodegraph;
uint64_t orc_hash_murmur_seed(const uint8_t *data, uint64_t len, uint32_t seed);
orc_nodegraph *orc_ng_new(uint32_t ksize, uint64_t tablesize, uint32_t n_tables);
orc_nodegraph *orc_ng_load(const char *path);
void orc_ng_free(orc_nodegraph *g);
uint32_t orc_ng_ksize(const orc_nodegraph *g);
uint32_t orc_ng_n_tables(const orc_nodegraph *g);
uint64_t orc_ng_table_size(const orc_nodegraph *g, uint32_t i);
uint64_t orc_ng_n_unique_kmers(const orc_nodegraph *g);
int orc_ng_add(orc_nodegraph *g, uint64_t h);
int orc_ng_get(const orc_nodegraph *g, uint64_t h);
uint64_t orc_ng_popcount(const orc_nodegraph *g, uint32_t i);
int orc_ng_save(const orc_nodegraph *g, const char *path);

/* hash_murmur(kmer, seed=42) -> int   (sourmash.minhash.hash_murmur) */
static PyObject *py_hash_murmur(PyObject *self, PyObject *args, PyObject *kwargs)
{
    static char *kw[] = {"kmer", "seed", NULL};
    PyObject *kmer;
    unsigned int seed = 42;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "O|I", kw, &kmer, &seed)) return NULL;
    const char *data;
    Py_ssize_t len;
    if (PyUnicode_Check(kmer)) {
        data = PyUnicode_AsUTF8AndSize(kmer, &len);
        if (!data) return NULL;
    } else if (PyBytes_Check(kmer)) {
        data = PyBytes_AS_STRING(kmer);
        len = PyBytes_GET_SIZE(kmer);
    } else {
        PyErr_SetString(PyExc_TypeError, "hash_murmur: kmer must be str or bytes");
        return NULL;
    }
    return PyLong_FromUnsignedLongLong(orc_hash_murmur_seed((const uint8_t *)data, (uint64_t)len, seed));
}

typedef struct {
    PyObject_HEAD
    orc_nodegraph *g;
} NodegraphObject;

static void Nodegraph_dealloc(NodegraphObject *self)
{
    if (self->g) orc_ng_free(self->g);
    Py_TYPE(self)->tp_free((PyObject *)self);
}

/* khmer.Nodegraph(k, starting_size, n_tables) */
static int Nodegraph_init(NodegraphObject *self, PyObject *args, PyObject *kwargs)
{
    static char *kw[] = {"k", "starting_size", "n_tables", NULL};
    unsigned int k, n_tables;
    double size;
    if (!PyArg_ParseTupleAndKeywords(args, kwargs, "IdI", kw, &k, &size, &n_tables)) return -1;
    if (self->g) orc_ng_free(self->g);
    self->g = orc_ng_new(k, (uint64_t)size, n_tables);
    if (!self->g) {
        PyErr_SetString(PyExc_MemoryError, "Nodegraph: cannot allocate the tables");
        return -1;
    }
    return 0;
}

static int parse_hash(PyObject *arg, uint64_t *h)
{
    *h = PyLong_AsUnsignedLongLongMask(arg);
    return PyErr_Occurred() ? -1 : 0;
}

static PyObject *Nodegraph_add(NodegraphObject *self, PyObject *arg)
{
    uint64_t h;
    if (parse_hash(arg, &h)) return NULL;
    return PyBool_FromLong(orc_ng_add(self->g, h));
}

static PyObject *Nodegraph_get(NodegraphObject *self, PyObject *arg)
{
    uint64_t h;
    if (parse_hash(arg, &h)) return NULL;
    return PyLong_FromLong(orc_ng_get(self->g, h));
}

static PyObject *Nodegraph_ksize(NodegraphObject *self, PyObject *ignored) { return PyLong_FromUnsignedLong(orc_ng_ksize(self->g)); }
static PyObject *Nodegraph_n_tables(NodegraphObject *self, PyObject *ignored) { return PyLong_FromUnsignedLong(orc_ng_n_tables(self->g)); }
static PyObject *Nodegraph_n_unique_kmers(NodegraphObject *self, PyObject *ignored)
{
    return PyLong_FromUnsignedLongLong(orc_ng_n_unique_kmers(self->g));
}
static PyObject *Nodegraph_n_occupied(NodegraphObject *self, PyObject *ignored) { return PyLong_FromUnsignedLongLong(orc_ng_popcount(self->g, 0)); }

static PyObject *Nodegraph_hashsizes(NodegraphObject *self, PyObject *ignored)
{
    const uint32_t n = orc_ng_n_tables(self->g);
    PyObject *out = PyList_New(n);
    if (!out) return NULL;
    for (uint32_t i = 0; i < n; i++) PyList_SET_ITEM(out, i, PyLong_FromUnsignedLongLong(orc_ng_table_size(self->g, i)));
    return out;
}

static PyObject *Nodegraph_save(NodegraphObject *self, PyObject *arg)
{
    PyObject *path = NULL;
    if (!PyUnicode_FSConverter(arg, &path)) return NULL;
    const int rc = orc_ng_save(self->g, PyBytes_AS_STRING(path));
    Py_DECREF(path);
    if (rc != 0) {
        PyErr_SetString(PyExc_OSError, "Nodegraph.save: cannot write the file");
        return NULL;
    }
    Py_RETURN_NONE;
}

static PyObject *Nodegraph_load(PyObject *cls, PyObject *arg)
{
    PyObject *path = NULL;
    if (!PyUnicode_FSConverter(arg, &path)) return NULL;
    orc_nodegraph *g = orc_ng_load(PyBytes_AS_STRING(path));
    Py_DECREF(path);
    if (!g) {
        PyErr_SetString(PyExc_OSError, "Nodegraph.load: not an OXLI nodegraph file");
        return NULL;
    }
    NodegraphObject *self = (NodegraphObject *)((PyTypeObject *)cls)->tp_alloc((PyTypeObject *)cls, 0);
    if (!self) {
        orc_ng_free(g);
        return NULL;
    }
    self->g = g;
    return (PyObject *)self;
}

static PyMethodDef Nodegraph_methods[] = {
    {"add", (PyCFunction)Nodegraph_add, METH_O, "add(hash) -> True if a new bit was set"},
    {"get", (PyCFunction)Nodegraph_get, METH_O, "get(hash) -> 1 if every table has the bit"},
    {"ksize", (PyCFunction)Nodegraph_ksize, METH_NOARGS, ""},
    {"n_tables", (PyCFunction)Nodegraph_n_tables, METH_NOARGS, ""},
    {"n_unique_kmers", (PyCFunction)Nodegraph_n_unique_kmers, METH_NOARGS, ""},
    {"n_occupied", (PyCFunction)Nodegraph_n_occupied, METH_NOARGS, ""},
    {"hashsizes", (PyCFunction)Nodegraph_hashsizes, METH_NOARGS, ""},
    {"save", (PyCFunction)Nodegraph_save, METH_O, ""},
    {"load", (PyCFunction)Nodegraph_load, METH_O | METH_CLASS, ""},
    {NULL, NULL, 0, NULL},
};

static PyTypeObject NodegraphType = {
    PyVarObject_HEAD_INIT(NULL, 0).tp_name = "_orph_refnative.Nodegraph",
    .tp_basicsize = sizeof(NodegraphObject),
    .tp_flags = Py_TPFLAGS_DEFAULT | Py_TPFLAGS_BASETYPE,
    .tp_new = PyType_GenericNew,
    .tp_init = (initproc)Nodegraph_init,
    .tp_dealloc = (destructor)Nodegraph_dealloc,
    .tp_methods = Nodegraph_methods,
};

static PyMethodDef module_methods[] = {
    {"hash_murmur", (PyCFunction)py_hash_murmur, METH_VARARGS | METH_KEYWORDS, "MurmurHash3_x64_128(kmer, seed=42), low 64 bits"},
    {NULL, NULL, 0, NULL},
};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_orph_refnative", "native khmer / sourmash stand-ins (measurement only)", -1, module_methods};

PyMODINIT_FUNC PyInit__orph_refnative(void)
{
    if (PyType_Ready(&NodegraphType) < 0) return NULL;
    PyObject *m = PyModule_Create(&moduledef);
    if (!m) return NULL;
    Py_INCREF(&NodegraphType);
    if (PyModule_AddObject(m, "Nodegraph", (PyObject *)&NodegraphType) < 0) {
        Py_DECREF(&NodegraphType);
        Py_DECREF(m);
        return NULL;
    }
    return m;
}
