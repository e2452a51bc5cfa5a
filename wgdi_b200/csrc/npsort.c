/*
 * npsort.c — segmented argsort that calls NUMPY'S OWN argsort routine, without the GIL.
 *
 * The reference orders block extraction with pandas' sort_values(by=['score1'], ascending=False)
 * (wgdi/collinearity.py:49,60), which ends in numpy.argsort(kind='quicksort') on an int64 array.  That
 * sort is unstable and its tie permutation depends on the numpy build and on the CPU (AVX-512 / AVX2 /
 * scalar kernels differ; SURVEY.md hazard H1), so it cannot be re-implemented — it has to be the same
 * routine.  numpy.argsort on a contiguous aligned 1-D int64 array does exactly this
 * (numpy/_core/src/multiarray/item_selection.c, _new_argsortlike): fill the index array with 0..n-1 and
 * call   PyDataType_GetArrFuncs(descr)->argsort[NPY_QUICKSORT](values, index, n, arr).
 * This module fetches that function pointer from the running numpy and calls it for every chromosome-pair
 * segment from a pthread pool.  The Python side verifies on start-up that the result is identical to
 * numpy.argsort and otherwise falls back to one numpy call per segment (wgdi_b200/engine.py).
 *
 * Python API:  seg_argsort(keys:int32[n], off:int64[S+1], seg_order:int32[S], out:int32[n], threads:int) -> 0
 *   out[off[s] + j] = argsort(keys[off[s]:off[s+1]].astype(int64), kind='quicksort')[j]   (narrowed to int32)
 * The keys cross PCIe as int32 and are widened per segment here: the reference's score column is int64 and
 * numpy's sort kernels (hence the tie permutation) are dtype specific.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/arrayobject.h>
#include <immintrin.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct {
    PyArray_ArgSortFunc *fn;
    const int32_t *keys;
    const int64_t *off;
    const int32_t *seg_order;
    int32_t *out;
    int64_t n_seg;
    volatile int64_t next;
    volatile int err;
    int flush;
} job_t;

static void *worker(void *vp)
{
    job_t *j = (job_t *)vp;
    npy_intp *idx = NULL;
    int64_t *wide = NULL;
    int64_t cap = 0;
    for (;;) {
        int64_t s = __sync_fetch_and_add(&j->next, 1);
        if (s >= j->n_seg) break;
        int64_t seg = j->seg_order ? j->seg_order[s] : s;
        int64_t a = j->off[seg], n = j->off[seg + 1] - a;
        if (n <= 0) continue;
        if (n > cap) {
            free(idx);
            free(wide);
            cap = n + n / 4 + 64;
            idx = (npy_intp *)malloc((size_t)cap * sizeof(npy_intp));
            wide = (int64_t *)malloc((size_t)cap * sizeof(int64_t));
            if (!idx || !wide) { j->err = 1; break; }
        }
        const int32_t *k32 = j->keys + a;
        if (j->flush == 2) {                   /* non-temporal prefetch: the keys pass through L1 only */
            for (int64_t i = 0; i < n; i += 16) _mm_prefetch((const char *)(k32 + i), _MM_HINT_NTA);
        }
        for (int64_t i = 0; i < n; i++) { idx[i] = i; wide[i] = k32[i]; }   /* index fill as _new_argsortlike does */
        /* The key buffer is the target of the next device-to-host DMA: drop its lines from the CPU caches now so
           that the DMA writes do not have to snoop them out (measured: 15 MB in 0.70 ms with cached lines vs
           0.28 ms for untouched pinned memory). */
        if (j->flush == 1)
            for (const char *q = (const char *)((uintptr_t)k32 & ~(uintptr_t)63); q < (const char *)(k32 + n); q += 64) _mm_clflush(q);
#ifdef __CLFLUSHOPT__
        if (j->flush == 3)
            for (const char *q = (const char *)((uintptr_t)k32 & ~(uintptr_t)63); q < (const char *)(k32 + n); q += 64) _mm_clflushopt((void *)q);
#endif
        if (j->fn((void *)wide, idx, (npy_intp)n, NULL) < 0) { j->err = 2; break; }
        int32_t *o = j->out + a;
        /* streaming stores: the argsort goes to the device next, keep it out of the caches */
        for (int64_t i = 0; i < n; i++) _mm_stream_si32((int *)&o[i], (int)idx[i]);
    }
    free(idx);
    free(wide);
    return NULL;
}

static PyObject *seg_argsort(PyObject *self, PyObject *args)
{
    PyArrayObject *keys, *off, *order, *out;
    int threads, flush = 0;
    if (!PyArg_ParseTuple(args, "O!O!O!O!i|i", &PyArray_Type, &keys, &PyArray_Type, &off, &PyArray_Type, &order,
                          &PyArray_Type, &out, &threads, &flush))
        return NULL;
    if (PyArray_TYPE(keys) != NPY_INT32 || PyArray_TYPE(off) != NPY_INT64 || PyArray_TYPE(order) != NPY_INT32 ||
        PyArray_TYPE(out) != NPY_INT32 || !PyArray_IS_C_CONTIGUOUS(keys) || !PyArray_IS_C_CONTIGUOUS(off) ||
        !PyArray_IS_C_CONTIGUOUS(order) || !PyArray_IS_C_CONTIGUOUS(out) || !PyArray_ISALIGNED(keys) ||
        PyArray_NDIM(keys) != 1 || PyArray_NDIM(out) != 1 || PyArray_SIZE(out) < PyArray_SIZE(keys) ||
        PyArray_SIZE(off) < 1 || PyArray_SIZE(order) != PyArray_SIZE(off) - 1) {
        PyErr_SetString(PyExc_TypeError, "seg_argsort(keys:int32[n], off:int64[S+1], seg_order:int32[S], out:int32[n], threads)");
        return NULL;
    }
    PyArray_Descr *d64 = PyArray_DescrFromType(NPY_INT64);
    PyArray_ArgSortFunc *fn = PyDataType_GetArrFuncs(d64)->argsort[NPY_QUICKSORT];
    Py_DECREF(d64);
    if (!fn) {
        PyErr_SetString(PyExc_RuntimeError, "numpy exposes no int64 quicksort argsort routine");
        return NULL;
    }
    job_t j;
    j.fn = fn;
    j.keys = (const int32_t *)PyArray_DATA(keys);
    j.off = (const int64_t *)PyArray_DATA(off);
    j.seg_order = (const int32_t *)PyArray_DATA(order);
    j.out = (int32_t *)PyArray_DATA(out);
    j.n_seg = PyArray_SIZE(off) - 1;
    j.next = 0;
    j.err = 0;
    j.flush = flush;
    if (threads < 1) threads = 1;
    if (threads > 128) threads = 128;
    if (threads > j.n_seg) threads = (int)(j.n_seg > 0 ? j.n_seg : 1);
    Py_BEGIN_ALLOW_THREADS
    pthread_t th[128];
    int started = 0;
    for (int t = 1; t < threads; t++)
        if (pthread_create(&th[started], NULL, worker, &j) == 0) started++;
    worker(&j);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    _mm_sfence();
    Py_END_ALLOW_THREADS
    if (j.err) {
        PyErr_SetString(PyExc_RuntimeError, j.err == 1 ? "seg_argsort: out of memory" : "seg_argsort: numpy argsort failed");
        return NULL;
    }
    return PyLong_FromLong(0);
}

/* CPUID.(EAX=7,ECX=0):EBX bit 23 */
static PyObject *has_clflushopt(PyObject *self, PyObject *args)
{
    unsigned a = 7, b = 0, c = 0, d = 0;
#if defined(__x86_64__)
    __asm__ volatile("cpuid" : "+a"(a), "=b"(b), "+c"(c), "=d"(d));
#endif
    return PyBool_FromLong((b >> 23) & 1u);
}

static PyMethodDef methods[] = {
    {"has_clflushopt", has_clflushopt, METH_NOARGS, "CPU supports CLFLUSHOPT"},
    {"seg_argsort", seg_argsort, METH_VARARGS, "segmented numpy argsort(kind='quicksort') on int64, multi-threaded"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_npsort", NULL, -1, methods, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__npsort(void)
{
    import_array();
    return PyModule_Create(&moddef);
}
