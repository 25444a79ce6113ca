/* _svimfastobj: builds the Python objects the reference's downstream stages read (SVSignature.py:36-310 attribute surface)
 * from the record arrays the C ABI returns, in one C loop.
 *
 * Materialising 250 k Signature objects + 14 k SignatureCluster objects with Python-level code cost 380 ms on BASELINE
 * configs[1] — more than twice the whole GPU path ("e2e excludes Python object materialisation").  The classes
 * in svim_b200/SVSignature.py declare __slots__ (plus __dict__, so attribute assignment from downstream code still works);
 * here every object is tp_alloc + pointer stores into its slots.  Nothing is computed: values are copied out of svim_sig /
 * svim_cluster records (include/svimgpu.h).  svim_b200/SVIM_COLLECT.py::materialize_signatures_py is the same thing in
 * Python and the tests compare the two attribute by attribute.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <structmember.h>
#include <stdint.h>
#include <string.h>

#pragma pack(push, 1)
typedef struct {            /* svim_sig, 48 bytes */
    int32_t start, end, pos, contig1, contig2;
    uint32_t aln_idx, qname_id, ordinal;
    uint64_t seq_off;
    uint32_t seq_len;
    uint8_t type, flags;
    uint16_t copies;
} sig_t;
typedef struct {            /* svim_cluster, 72 bytes */
    int64_t start, end, dest_start, dest_end;
    double score, std_span, std_pos;
    uint32_t member_off, size;
    uint8_t type, dir1_rev, dir2_rev, pad0;
    uint32_t pad1;
} cluster_t;
#pragma pack(pop)

enum { F_SUPPL = 1, F_FULLY = 2, F_DIR1_REV = 4, F_DIR2_REV = 8, F_INVDIR_SHIFT = 4 };
enum { T_DEL = 0, T_INS = 1, T_INV = 2, T_DUP_TAN = 3, T_BND = 4, T_DUP_INT = 5 };

static Py_ssize_t slot_offset(PyObject* cls, const char* name) {
    PyObject* d = PyObject_GetAttrString(cls, name);
    if (!d) return -1;
    Py_ssize_t off = -1;
    if (Py_TYPE(d) == &PyMemberDescr_Type) off = ((PyMemberDescrObject*)d)->d_member->offset;
    else PyErr_Format(PyExc_TypeError, "%R.%s is not a slot", cls, name);
    Py_DECREF(d);
    return off;
}

#define SLOT(obj, off) (*(PyObject**)((char*)(obj) + (off)))
/* store a NEW reference */
#define PUT_NEW(obj, off, v) do { PyObject* _v = (v); if (!_v) goto fail; SLOT(obj, off) = _v; } while (0)
/* store a borrowed reference */
#define PUT(obj, off, v) do { PyObject* _v = (v); Py_INCREF(_v); SLOT(obj, off) = _v; } while (0)

/* signatures(sig_buffer, ins_buffer, contig_names: list[str], qnames: list[str] | None, classes: 6-tuple in type-code order,
 *            inv_directions: tuple[str]) -> list */
static PyObject* fo_signatures(PyObject* self, PyObject* args) {
    Py_buffer sb, ib;
    PyObject *names, *qnames, *classes, *invdirs;
    if (!PyArg_ParseTuple(args, "y*y*OOOO", &sb, &ib, &names, &qnames, &classes, &invdirs)) return NULL;
    PyObject* out = NULL;
    PyObject *s_cigar = NULL, *s_suppl = NULL, *s_fwd = NULL, *s_rev = NULL;
    PyObject** name_cache = NULL; size_t name_cap = 0;
    const Py_ssize_t n = sb.len / (Py_ssize_t)sizeof(sig_t);
    const sig_t* sg = (const sig_t*)sb.buf;
    const char* ins = (const char*)ib.buf;
    if (!PyList_Check(names) || !PyTuple_Check(classes) || PyTuple_GET_SIZE(classes) != 6 || !PyTuple_Check(invdirs)) {
        PyErr_SetString(PyExc_TypeError, "signatures(): bad arguments"); goto fail0;
    }
    PyTypeObject* cls[6]; Py_ssize_t o_contig[6], o_start[6], o_end[6], o_sig[6], o_read[6];
    Py_ssize_t o_seq, o_dir, o_copies, o_fully, o_c2, o_pos, o_b[8];
    for (int t = 0; t < 6; ++t) {
        cls[t] = (PyTypeObject*)PyTuple_GET_ITEM(classes, t);
        PyObject* c = (PyObject*)cls[t];
        o_sig[t] = slot_offset(c, "signature"); o_read[t] = slot_offset(c, "read");
        if (o_sig[t] < 0 || o_read[t] < 0) goto fail0;
        if (t == T_BND) continue;
        o_contig[t] = slot_offset(c, t == T_DUP_INT ? "contig1" : "contig"); o_start[t] = slot_offset(c, "start"); o_end[t] = slot_offset(c, "end");
        if (o_contig[t] < 0 || o_start[t] < 0 || o_end[t] < 0) goto fail0;
    }
    o_seq = slot_offset((PyObject*)cls[T_INS], "sequence"); o_dir = slot_offset((PyObject*)cls[T_INV], "direction");
    o_copies = slot_offset((PyObject*)cls[T_DUP_TAN], "copies"); o_fully = slot_offset((PyObject*)cls[T_DUP_TAN], "fully_covered");
    o_c2 = slot_offset((PyObject*)cls[T_DUP_INT], "contig2"); o_pos = slot_offset((PyObject*)cls[T_DUP_INT], "pos");
    {
        const char* bn[6] = {"contig1", "pos1", "direction1", "contig2", "pos2", "direction2"};
        for (int k = 0; k < 6; ++k) { o_b[k] = slot_offset((PyObject*)cls[T_BND], bn[k]); if (o_b[k] < 0) goto fail0; }
    }
    if (o_seq < 0 || o_dir < 0 || o_copies < 0 || o_fully < 0 || o_c2 < 0 || o_pos < 0) goto fail0;
    s_cigar = PyUnicode_InternFromString("cigar"); s_suppl = PyUnicode_InternFromString("suppl");
    s_fwd = PyUnicode_InternFromString("fwd"); s_rev = PyUnicode_InternFromString("rev");
    const Py_ssize_t n_names = PyList_GET_SIZE(names);
    const int have_q = qnames != Py_None;
    if (have_q && !PyList_Check(qnames)) { PyErr_SetString(PyExc_TypeError, "qnames must be a list or None"); goto fail0; }
    const Py_ssize_t n_q = have_q ? PyList_GET_SIZE(qnames) : 0;
    if (!have_q) {
        uint32_t mx = 0;
        for (Py_ssize_t k = 0; k < n; ++k) if (sg[k].qname_id > mx) mx = sg[k].qname_id;
        name_cap = (size_t)mx + 1;
        name_cache = (PyObject**)calloc(name_cap, sizeof(PyObject*));
        if (!name_cache) { PyErr_NoMemory(); goto fail0; }
    }
    out = PyList_New(n);
    if (!out) goto fail0;
    for (Py_ssize_t k = 0; k < n; ++k) {
        const sig_t* s = sg + k;
        const int t = s->type;
        if (t > 5 || s->contig1 < 0 || s->contig1 >= n_names) { PyErr_Format(PyExc_ValueError, "signature %zd: bad type/contig", k); goto fail; }
        PyObject* o = cls[t]->tp_alloc(cls[t], 0);
        if (!o) goto fail;
        PyList_SET_ITEM(out, k, o);              /* the list owns it from here on; slots are NULL (= unset) until stored */
        PUT(o, o_sig[t], (s->flags & F_SUPPL) ? s_suppl : s_cigar);
        if (have_q) {
            if ((Py_ssize_t)s->qname_id >= n_q) { PyErr_Format(PyExc_IndexError, "signature %zd: read id %u out of range", k, s->qname_id); goto fail; }
            PUT(o, o_read[t], PyList_GET_ITEM(qnames, s->qname_id));
        } else {                               /* synthetic input without a name table: "read<id>", one string per id */
            const uint32_t id = s->qname_id;
            if (id >= name_cap) { PyErr_SetString(PyExc_RuntimeError, "signatures(): read id above the scanned maximum"); goto fail; }
            PyObject* r = name_cache[id];
            if (!r) {
                char buf[24];
                const int len = snprintf(buf, sizeof buf, "read%u", id);
                r = PyUnicode_FromStringAndSize(buf, len);
                if (!r) goto fail;
                name_cache[id] = r;             /* the cache owns one reference, released at the end */
            }
            PUT(o, o_read[t], r);
        }
        PyObject* c1 = PyList_GET_ITEM(names, s->contig1);
        if (t == T_BND) {
            if (s->contig2 < 0 || s->contig2 >= n_names) { PyErr_Format(PyExc_ValueError, "signature %zd: bad contig2", k); goto fail; }
            PUT(o, o_b[0], c1); PUT_NEW(o, o_b[1], PyLong_FromLong(s->start)); PUT(o, o_b[2], (s->flags & F_DIR1_REV) ? s_rev : s_fwd);
            PUT(o, o_b[3], PyList_GET_ITEM(names, s->contig2)); PUT_NEW(o, o_b[4], PyLong_FromLong(s->pos)); PUT(o, o_b[5], (s->flags & F_DIR2_REV) ? s_rev : s_fwd);
            continue;
        }
        PUT(o, o_contig[t], c1); PUT_NEW(o, o_start[t], PyLong_FromLong(s->start)); PUT_NEW(o, o_end[t], PyLong_FromLong(s->end));
        if (t == T_INS) {
            if (s->seq_off + s->seq_len > (uint64_t)ib.len) { PyErr_Format(PyExc_ValueError, "signature %zd: sequence outside the blob", k); goto fail; }
            /* the blob holds the kernel's own 4-bit -> "=ACMGRSVTWYHKDBN" decode: 7-bit by construction, no validation pass */
            PyObject* u = PyUnicode_New((Py_ssize_t)s->seq_len, 127);
            if (!u) goto fail;
            memcpy(PyUnicode_1BYTE_DATA(u), ins + s->seq_off, s->seq_len);
            SLOT(o, o_seq) = u;
        } else if (t == T_INV) {
            const Py_ssize_t d = (s->flags >> F_INVDIR_SHIFT) & 7;
            if (d >= PyTuple_GET_SIZE(invdirs)) { PyErr_Format(PyExc_ValueError, "signature %zd: bad inversion direction", k); goto fail; }
            PUT(o, o_dir, PyTuple_GET_ITEM(invdirs, d));
        } else if (t == T_DUP_TAN) {
            PUT_NEW(o, o_copies, PyLong_FromLong(s->copies));
            PUT(o, o_fully, (s->flags & F_FULLY) ? Py_True : Py_False);
        } else if (t == T_DUP_INT) {
            if (s->contig2 < 0 || s->contig2 >= n_names) { PyErr_Format(PyExc_ValueError, "signature %zd: bad contig2", k); goto fail; }
            PUT(o, o_c2, PyList_GET_ITEM(names, s->contig2)); PUT_NEW(o, o_pos, PyLong_FromLong(s->pos));
        }
    }
    goto done;
fail:
    Py_CLEAR(out);
fail0:
done:
    Py_XDECREF(s_cigar); Py_XDECREF(s_suppl); Py_XDECREF(s_fwd); Py_XDECREF(s_rev);
    if (name_cache) { for (size_t k = 0; k < name_cap; ++k) Py_XDECREF(name_cache[k]); free(name_cache); }
    PyBuffer_Release(&sb); PyBuffer_Release(&ib);
    return out;
}

/* clusters(cluster_buffer, members_buffer (uint32), signatures: list, uni_cls, bi_cls, type_names: 6-tuple[str]) -> list of 6 lists
 * (type-code order).  Contig names come from the first member like the reference's consolidate_* (SVIM_clustering.py:214-303). */
static PyObject* fo_clusters(PyObject* self, PyObject* args) {
    Py_buffer cb, mb;
    PyObject *sigs, *uni, *bi, *tnames;
    if (!PyArg_ParseTuple(args, "y*y*OOOO", &cb, &mb, &sigs, &uni, &bi, &tnames)) return NULL;
    PyObject* out = NULL;
    PyObject *s_fwd = NULL, *s_rev = NULL, *m_src = NULL, *m_dst = NULL, *s_d1 = NULL, *s_d2 = NULL;
    const Py_ssize_t n = cb.len / (Py_ssize_t)sizeof(cluster_t), n_mem = mb.len / 4;
    const cluster_t* cl = (const cluster_t*)cb.buf;
    const uint32_t* mem = (const uint32_t*)mb.buf;
    if (!PyList_Check(sigs) || !PyTuple_Check(tnames) || PyTuple_GET_SIZE(tnames) != 6) { PyErr_SetString(PyExc_TypeError, "clusters(): bad arguments"); goto fail0; }
    const Py_ssize_t n_sigs = PyList_GET_SIZE(sigs);
    PyTypeObject* ucls = (PyTypeObject*)uni; PyTypeObject* bcls = (PyTypeObject*)bi;
    const char* un[9] = {"contig", "start", "end", "score", "size", "members", "type", "std_span", "std_pos"};
    const char* bn[12] = {"source_contig", "source_start", "source_end", "dest_contig", "dest_start", "dest_end", "score", "size", "members", "type", "std_span", "std_pos"};
    Py_ssize_t ou[9], ob[12];
    for (int k = 0; k < 9; ++k) { ou[k] = slot_offset(uni, un[k]); if (ou[k] < 0) goto fail0; }
    for (int k = 0; k < 12; ++k) { ob[k] = slot_offset(bi, bn[k]); if (ob[k] < 0) goto fail0; }
    s_fwd = PyUnicode_InternFromString("fwd"); s_rev = PyUnicode_InternFromString("rev");
    m_src = PyUnicode_InternFromString("get_source"); m_dst = PyUnicode_InternFromString("get_destination");
    out = PyList_New(6);
    if (!out) goto fail0;
    for (int t = 0; t < 6; ++t) { PyObject* l = PyList_New(0); if (!l) goto fail; PyList_SET_ITEM(out, t, l); }
    s_d1 = PyUnicode_InternFromString("direction1"); s_d2 = PyUnicode_InternFromString("direction2");
    for (Py_ssize_t k = 0; k < n; ++k) {
        const cluster_t* c = cl + k;
        const int t = c->type;
        if (t > 5 || c->size == 0 || (Py_ssize_t)c->member_off + c->size > n_mem) { PyErr_Format(PyExc_ValueError, "cluster %zd: bad record", k); goto fail; }
        PyObject* ms = PyList_New(c->size);
        if (!ms) goto fail;
        for (uint32_t i = 0; i < c->size; ++i) {
            const uint32_t g = mem[c->member_off + i];
            if ((Py_ssize_t)g >= n_sigs) { Py_DECREF(ms); PyErr_Format(PyExc_IndexError, "cluster %zd: member outside the signature list", k); goto fail; }
            PyObject* m = PyList_GET_ITEM(sigs, g); Py_INCREF(m); PyList_SET_ITEM(ms, i, m);
        }
        PyObject* first = PyList_GET_ITEM(ms, 0);
        PyObject* src = PyObject_CallMethodNoArgs(first, m_src);
        if (!src) { Py_DECREF(ms); goto fail; }
        PyObject* o = (t <= T_INV ? ucls : bcls)->tp_alloc(t <= T_INV ? ucls : bcls, 0);
        if (!o) { Py_DECREF(ms); Py_DECREF(src); goto fail; }
        if (PyList_Append(PyList_GET_ITEM(out, t), o) < 0) { Py_DECREF(o); Py_DECREF(ms); Py_DECREF(src); goto fail; }
        Py_DECREF(o);                            /* the per-type list owns it */
        PyObject* sc = PyTuple_GetItem(src, 0);
        if (!sc) { Py_DECREF(ms); Py_DECREF(src); goto fail; }
        PyObject* sd_span = isnan(c->std_span) ? Py_NewRef(Py_None) : PyFloat_FromDouble(c->std_span);
        PyObject* sd_pos = isnan(c->std_pos) ? Py_NewRef(Py_None) : PyFloat_FromDouble(c->std_pos);
        if (t <= T_INV) {
            SLOT(o, ou[5]) = ms;
            PUT(o, ou[0], sc); Py_DECREF(src);
            SLOT(o, ou[7]) = sd_span; SLOT(o, ou[8]) = sd_pos;
            PUT_NEW(o, ou[1], PyLong_FromLongLong(c->start)); PUT_NEW(o, ou[2], PyLong_FromLongLong(c->end));
            PUT_NEW(o, ou[3], PyFloat_FromDouble(c->score)); PUT_NEW(o, ou[4], PyLong_FromUnsignedLong(c->size));
            PUT(o, ou[6], PyTuple_GET_ITEM(tnames, t));
        } else {
            SLOT(o, ob[8]) = ms;
            PUT(o, ob[0], sc); Py_DECREF(src);
            SLOT(o, ob[10]) = sd_span; SLOT(o, ob[11]) = sd_pos;
            if (t == T_DUP_TAN) PUT(o, ob[3], SLOT(o, ob[0]));
            else {
                PyObject* dst = PyObject_CallMethodNoArgs(first, m_dst);
                if (!dst) goto fail;
                PyObject* dc = PyTuple_GetItem(dst, 0);
                if (!dc) { Py_DECREF(dst); goto fail; }
                PUT(o, ob[3], dc); Py_DECREF(dst);
            }
            PUT_NEW(o, ob[1], PyLong_FromLongLong(c->start)); PUT_NEW(o, ob[2], PyLong_FromLongLong(c->end));
            PUT_NEW(o, ob[4], PyLong_FromLongLong(c->dest_start)); PUT_NEW(o, ob[5], PyLong_FromLongLong(c->dest_end));
            PUT_NEW(o, ob[6], PyFloat_FromDouble(c->score)); PUT_NEW(o, ob[7], PyLong_FromUnsignedLong(c->size));
            PUT(o, ob[9], PyTuple_GET_ITEM(tnames, t));
            if (t == T_BND) {                    /* attributes added after construction in the reference (:300-301) */
                if (PyObject_SetAttr(o, s_d1, c->dir1_rev ? s_rev : s_fwd) < 0 || PyObject_SetAttr(o, s_d2, c->dir2_rev ? s_rev : s_fwd) < 0) goto fail;
            }
        }
    }
    goto done;
fail:
    Py_CLEAR(out);
fail0:
done:
    Py_XDECREF(s_fwd); Py_XDECREF(s_rev); Py_XDECREF(m_src); Py_XDECREF(m_dst); Py_XDECREF(s_d1); Py_XDECREF(s_d2);
    PyBuffer_Release(&cb); PyBuffer_Release(&mb);
    return out;
}

static PyMethodDef methods[] = {
    {"signatures", fo_signatures, METH_VARARGS, "svim_sig records -> list of Signature objects"},
    {"clusters", fo_clusters, METH_VARARGS, "svim_cluster records -> six lists of SignatureCluster objects"},
    {NULL, NULL, 0, NULL}};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_svimfastobj", "object materialisation for svim_b200 (host side)", -1, methods};
PyMODINIT_FUNC PyInit__svimfastobj(void) { return PyModule_Create(&moddef); }
