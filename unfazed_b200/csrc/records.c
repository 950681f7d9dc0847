/* Host-side record builder: device evidence lists -> the record dicts phase_snvs / phase_svs return.
 *
 * The reference builds one record per phased DNM in Python (snv_phaser.py:169-203: region, vartype, kid,
 * dad, mom, per-parent sorted lists of str(pos) and of read names).  After the kernels a 10 k-DNM batch is a
 * few hundred thousand small strings and ~7 k dicts; creating them from the interpreter was the largest
 * part of an end-to-end step.  This CPython extension does exactly what phaser.BatchPhaser.records did for
 * the read-backed entries, from the same flat arrays, with the C API (no per-string bytecode).
 *
 * Python stays the host language of the drop-in (the reference's own); nothing here does phasing
 * arithmetic: the lists arrive compacted per DNM and per parent from evidence_lists_kernel (chain.cu).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

static PyObject *s_region, *s_chrom, *s_start, *s_end, *s_vartype, *s_kid, *s_dad, *s_mom, *s_dad_sites, *s_mom_sites,
    *s_evidence_type, *s_dad_reads, *s_mom_reads, *s_cnv_dad_sites, *s_cnv_mom_sites, *s_cnv_evidence_type,
    *s_readbacked, *s_empty, *s_NA, *s_SEXCHROM, *s_us;

/* decimal text of v into buf (no terminator), returns the length */
static inline int fmt_i64(char *buf, int64_t v) {
    char tmp[24];
    int n = 0, neg = v < 0;
    uint64_t u = neg ? (uint64_t)(-(v + 1)) + 1u : (uint64_t)v;
    do { tmp[n++] = (char)('0' + u % 10u); u /= 10u; } while (u);
    int k = 0;
    if (neg) buf[k++] = '-';
    while (n) buf[k++] = tmp[--n];
    return k;
}

static inline PyObject *ascii_str(const char *p, int n) {
    PyObject *o = PyUnicode_New(n, 127);
    if (o) memcpy(PyUnicode_1BYTE_DATA(o), p, (size_t)n);
    return o;
}

typedef struct { char s[12]; int n; } PosTxt;

static int postxt_cmp(const PosTxt *a, const PosTxt *b) {       /* str.__lt__ on ASCII digits: code-point order, shorter prefix first */
    int m = a->n < b->n ? a->n : b->n;
    int c = memcmp(a->s, b->s, (size_t)m);
    return c ? c : a->n - b->n;
}

/* sorted(set(str(p) for p in pos[a:b])) -- or the list as it comes when it has fewer than two entries */
static PyObject *site_list(const int32_t *pos, int64_t a, int64_t b) {
    int64_t n = b - a;
    if (n <= 0) return PyList_New(0);
    PosTxt stack_buf[64];
    PosTxt *t = n <= 64 ? stack_buf : (PosTxt *)PyMem_Malloc((size_t)n * sizeof(PosTxt));
    if (!t) return PyErr_NoMemory();
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) {                           /* insertion sort with duplicate drop: the lists are short */
        PosTxt x;
        x.n = fmt_i64(x.s, pos[a + i]);
        int64_t j = m;
        int dup = 0;
        while (j > 0) {
            int c = postxt_cmp(&t[j - 1], &x);
            if (c == 0) { dup = 1; break; }
            if (c < 0) break;
            --j;
        }
        if (dup) continue;
        memmove(&t[j + 1], &t[j], (size_t)(m - j) * sizeof(PosTxt));
        t[j] = x;
        ++m;
    }
    PyObject *l = PyList_New(m);
    if (l)
        for (int64_t i = 0; i < m; ++i) {
            PyObject *o = ascii_str(t[i].s, t[i].n);
            if (!o) { Py_CLEAR(l); break; }
            PyList_SET_ITEM(l, i, o);
        }
    if (t != stack_buf) PyMem_Free(t);
    return l;
}

/* [names[i] for i in idx[a:b]]  or  ["q%d" % pair_ids[i] ...] */
static PyObject *name_list(PyObject *names, const int64_t *pair_ids, int64_t n_reads, const int32_t *idx, int64_t a, int64_t b) {
    int64_t n = b > a ? b - a : 0;
    PyObject *l = PyList_New(n);
    if (!l) return NULL;
    for (int64_t i = 0; i < n; ++i) {
        int64_t r = idx[a + i];
        PyObject *o;
        if (r < 0 || r >= n_reads) {
            PyErr_SetString(PyExc_IndexError, "evidence read index out of range");
            Py_DECREF(l);
            return NULL;
        }
        if (names) {
            o = PyList_GET_ITEM(names, r);
            Py_INCREF(o);
        } else {
            char buf[24];
            buf[0] = 'q';
            int k = 1 + fmt_i64(buf + 1, pair_ids[r]);
            o = ascii_str(buf, k);
            if (!o) { Py_DECREF(l); return NULL; }
        }
        PyList_SET_ITEM(l, i, o);
    }
    return l;
}

static int set_steal(PyObject *d, PyObject *k, PyObject *v) {   /* d[k] = v, consuming v */
    if (!v) return -1;
    int rc = PyDict_SetItem(d, k, v);
    Py_DECREF(v);
    return rc;
}

typedef struct { Py_buffer v; int ok; } Buf;
static int get_buf(PyObject *o, Buf *b, Py_ssize_t itemsize, const char *what) {
    b->ok = 0;
    if (PyObject_GetBuffer(o, &b->v, PyBUF_C_CONTIGUOUS) < 0) return -1;
    b->ok = 1;
    if (b->v.itemsize != itemsize) {
        PyErr_Format(PyExc_TypeError, "%s: expected %zd-byte items, got %zd", what, itemsize, b->v.itemsize);
        return -1;
    }
    return 0;
}
static void rel_buf(Buf *b) { if (b->ok) PyBuffer_Release(&b->v); b->ok = 0; }

/* read_records(out, entries, ped, live, auto, names, pair_ids, read_dad, read_mom, pos_dad, pos_mom, off) -> number added
 *   out      dict filled with key -> record
 *   entries  list of DNM dicts (plan.entries), indexed by the values of live
 *   ped      {kid: {"dad":..., "mom":...}}
 *   live     int64[k] entries that get a record, auto uint8[k] = 1 for autophased ones (sex chromosome outside the PARs)
 *   names    list of query names by read index, or None with pair_ids int64[n_reads] ("q%d" names of the synthetic tables)
 *   read_* / pos_*  int32 flat evidence lists;  off int64[4][n+1]: read_dad, read_mom, pos_dad, pos_mom offsets per entry */
static PyObject *read_records(PyObject *self, PyObject *args) {
    PyObject *out, *entries, *ped, *o_live, *o_auto, *names, *o_pid, *o_rd, *o_rm, *o_pd, *o_pm, *o_off;
    if (!PyArg_ParseTuple(args, "O!O!O!OOOOOOOOO", &PyDict_Type, &out, &PyList_Type, &entries, &PyDict_Type, &ped, &o_live, &o_auto,
                          &names, &o_pid, &o_rd, &o_rm, &o_pd, &o_pm, &o_off))
        return NULL;
    Buf live = {0}, au = {0}, pid = {0}, rd = {0}, rm = {0}, pd = {0}, pm = {0}, off = {0};
    PyObject *ret = NULL;
    int64_t n_reads = 0;
    if (names == Py_None) {
        names = NULL;
        if (get_buf(o_pid, &pid, 8, "pair_ids") < 0) goto done;
        n_reads = pid.v.len / 8;
    } else {
        if (!PyList_Check(names)) { PyErr_SetString(PyExc_TypeError, "names must be a list or None"); goto done; }
        n_reads = PyList_GET_SIZE(names);
    }
    if (get_buf(o_live, &live, 8, "live") < 0 || get_buf(o_auto, &au, 1, "auto") < 0 || get_buf(o_rd, &rd, 4, "read_dad") < 0 ||
        get_buf(o_rm, &rm, 4, "read_mom") < 0 || get_buf(o_pd, &pd, 4, "pos_dad") < 0 || get_buf(o_pm, &pm, 4, "pos_mom") < 0 ||
        get_buf(o_off, &off, 8, "off") < 0)
        goto done;
    {
        const int64_t k = live.v.len / 8, n1 = off.v.len / 32;
        const int64_t *lv = (const int64_t *)live.v.buf, *of = (const int64_t *)off.v.buf;
        const uint8_t *isa = (const uint8_t *)au.v.buf;
        const int64_t *o_rd_ = of, *o_rm_ = of + n1, *o_sd_ = of + 2 * n1, *o_sm_ = of + 3 * n1;
        const int64_t len_rd = rd.v.len / 4, len_rm = rm.v.len / 4, len_pd = pd.v.len / 4, len_pm = pm.v.len / 4;
        const Py_ssize_t n_ent = PyList_GET_SIZE(entries);
        if (au.v.len != k) { PyErr_SetString(PyExc_ValueError, "live / auto lengths differ"); goto done; }
        int64_t added = 0;
        for (int64_t i = 0; i < k; ++i) {
            const int64_t d = lv[i];
            if (d < 0 || d >= n_ent || (!isa[i] && d + 1 >= n1)) {
                PyErr_SetString(PyExc_IndexError, "entry index out of range");
                goto done;
            }
            PyObject *dn = PyList_GET_ITEM(entries, d);
            PyObject *chrom = PyObject_GetItem(dn, s_chrom), *start = PyObject_GetItem(dn, s_start), *end = PyObject_GetItem(dn, s_end),
                     *kid = PyObject_GetItem(dn, s_kid), *vt = PyObject_GetItem(dn, s_vartype);
            PyObject *key = NULL, *rec = NULL, *region = NULL, *p = NULL, *dad = NULL, *mom = NULL, *parts = NULL;
            int bad = !(chrom && start && end && kid && vt);
            if (!bad) {
                p = PyObject_GetItem(ped, kid);
                if (p) { dad = PyObject_GetItem(p, s_dad); mom = PyObject_GetItem(p, s_mom); }
                bad = !(p && dad && mom);
            }
            if (!bad) {                                          /* dnm_key: "_".join([str(chrom), str(start), str(end), kid, vartype]) */
                parts = PyList_New(5);
                PyObject *a0 = PyObject_Str(chrom), *a1 = PyObject_Str(start), *a2 = PyObject_Str(end);
                if (parts && a0 && a1 && a2) {
                    PyList_SET_ITEM(parts, 0, a0); PyList_SET_ITEM(parts, 1, a1); PyList_SET_ITEM(parts, 2, a2);
                    Py_INCREF(kid); PyList_SET_ITEM(parts, 3, kid);
                    Py_INCREF(vt); PyList_SET_ITEM(parts, 4, vt);
                    key = PyUnicode_Join(s_us, parts);
                } else { Py_XDECREF(a0); Py_XDECREF(a1); Py_XDECREF(a2); }
                bad = !key;
            }
            if (!bad) {
                region = PyDict_New();
                rec = PyDict_New();
                bad = !(region && rec) || PyDict_SetItem(region, s_chrom, chrom) < 0 || PyDict_SetItem(region, s_start, start) < 0 ||
                      PyDict_SetItem(region, s_end, end) < 0 || PyDict_SetItem(rec, s_region, region) < 0 ||
                      PyDict_SetItem(rec, s_vartype, vt) < 0 || PyDict_SetItem(rec, s_kid, kid) < 0 ||
                      PyDict_SetItem(rec, s_dad, dad) < 0 || PyDict_SetItem(rec, s_mom, mom) < 0;
            }
            if (!bad) {
                if (isa[i]) {                                    /* phaser._auto_record */
                    bad = PyDict_SetItem(rec, s_cnv_dad_sites, s_NA) < 0 || PyDict_SetItem(rec, s_cnv_mom_sites, s_NA) < 0 ||
                          PyDict_SetItem(rec, s_cnv_evidence_type, s_SEXCHROM) < 0 || PyDict_SetItem(rec, s_dad_sites, s_empty) < 0 ||
                          PyDict_SetItem(rec, s_mom_sites, s_empty) < 0 || PyDict_SetItem(rec, s_evidence_type, s_SEXCHROM) < 0 ||
                          set_steal(rec, s_dad_reads, PyList_New(0)) < 0 || set_steal(rec, s_mom_reads, PyList_New(0)) < 0;
                } else if (o_rd_[d + 1] > len_rd || o_rm_[d + 1] > len_rm || o_sd_[d + 1] > len_pd || o_sm_[d + 1] > len_pm ||
                           o_rd_[d] < 0 || o_rm_[d] < 0 || o_sd_[d] < 0 || o_sm_[d] < 0) {
                    PyErr_SetString(PyExc_IndexError, "evidence offsets exceed the lists");
                    bad = 1;
                } else {
                    bad = set_steal(rec, s_dad_sites, site_list((const int32_t *)pd.v.buf, o_sd_[d], o_sd_[d + 1])) < 0 ||
                          set_steal(rec, s_mom_sites, site_list((const int32_t *)pm.v.buf, o_sm_[d], o_sm_[d + 1])) < 0 ||
                          PyDict_SetItem(rec, s_evidence_type, s_readbacked) < 0 ||
                          set_steal(rec, s_dad_reads, name_list(names, (const int64_t *)pid.v.buf, n_reads, (const int32_t *)rd.v.buf, o_rd_[d], o_rd_[d + 1])) < 0 ||
                          set_steal(rec, s_mom_reads, name_list(names, (const int64_t *)pid.v.buf, n_reads, (const int32_t *)rm.v.buf, o_rm_[d], o_rm_[d + 1])) < 0 ||
                          PyDict_SetItem(rec, s_cnv_dad_sites, s_empty) < 0 || PyDict_SetItem(rec, s_cnv_mom_sites, s_empty) < 0 ||
                          PyDict_SetItem(rec, s_cnv_evidence_type, s_empty) < 0;
                }
            }
            if (!bad) bad = PyDict_SetItem(out, key, rec) < 0;
            Py_XDECREF(chrom); Py_XDECREF(start); Py_XDECREF(end); Py_XDECREF(kid); Py_XDECREF(vt);
            Py_XDECREF(p); Py_XDECREF(dad); Py_XDECREF(mom); Py_XDECREF(parts); Py_XDECREF(key); Py_XDECREF(region); Py_XDECREF(rec);
            if (bad) goto done;
            ++added;
        }
        ret = PyLong_FromLongLong(added);
    }
done:
    rel_buf(&live); rel_buf(&au); rel_buf(&pid); rel_buf(&rd); rel_buf(&rm); rel_buf(&pd); rel_buf(&pm); rel_buf(&off);
    return ret;
}

static PyMethodDef methods[] = {
    {"read_records", read_records, METH_VARARGS, "device evidence lists -> record dicts (see phaser.BatchPhaser.records)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_records", "native record builder of unfazed_b200", -1, methods};

PyMODINIT_FUNC PyInit__records(void) {
#define S(var, txt) if (!(var = PyUnicode_InternFromString(txt))) return NULL
    S(s_region, "region"); S(s_chrom, "chrom"); S(s_start, "start"); S(s_end, "end"); S(s_vartype, "vartype"); S(s_kid, "kid");
    S(s_dad, "dad"); S(s_mom, "mom"); S(s_dad_sites, "dad_sites"); S(s_mom_sites, "mom_sites"); S(s_evidence_type, "evidence_type");
    S(s_dad_reads, "dad_reads"); S(s_mom_reads, "mom_reads"); S(s_cnv_dad_sites, "cnv_dad_sites"); S(s_cnv_mom_sites, "cnv_mom_sites");
    S(s_cnv_evidence_type, "cnv_evidence_type"); S(s_readbacked, "readbacked"); S(s_empty, ""); S(s_NA, "NA");
    S(s_SEXCHROM, "SEX-CHROM"); S(s_us, "_");
#undef S
    return PyModule_Create(&moddef);
}
