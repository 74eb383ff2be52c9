/*
 * _avbhost -- CPython extension: the per-frame host side of ImageProcessingPipeline.stereo_callback in C.
 *
 * The CUDA frame takes ~0.15 ms; building 300 FeatureMeasurement objects and integrating the gyro window in
 * interpreted Python took longer than that.  This module keeps the reference's Python surface (message objects in,
 * list of FeatureMeasurement out; reference image_processing/pipeline.py:46-150, feature_publisher.py:109-121,
 * imu_processor.py:28-67) and does the per-frame work natively:
 *
 *   process_frame(ctx, img0, img1, R|None (3x3 or 2x3x3), FeatureMeasurement) -> (features, header tuple)
 *   submit_images(ctx, img0, img1) + finish_frame(ctx, R|None, FeatureMeasurement): the same in two halves, so that the
 *       caller integrates the gyro window between them, while the images are on the bus
 *       copies the two host images into libavb's pinned staging block (GIL released), runs avb_process_frame
 *       (H2D + CUDA-graph frame + D2H), and materialises the FeatureMeasurement list straight from the pinned
 *       result block
 *   integrate_imu(buffer, t_prev, t_curr, R_cam0_imu, R_cam1_imu, out0, out1) -> end index | -1
 *       IMUProcessor.integrate_imu_data: mean gyro over the window rule of imu_processor.py:37-66, Rodrigues,
 *       transposed; the caller trims the buffer
 *
 * Built by uav-airvision_b200/build.py with gcc against libavb.so (C-ABI in include/avb.h).  No numpy C-API:
 * arrays travel through the buffer protocol.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <descrobject.h>
#include <structmember.h>

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/avb.h"

/* ---- FeatureMeasurement: the stereo measurement the MSCKF reads (id, u0, v0, u1, v1) -------------------------
 * Same attribute set as the reference's plain Python class (image_processing/feature_measurment.py:1-9, read at
 * msckf.py:430-438); here a C struct with member descriptors so that a frame's 300 measurements cost 300 allocations
 * instead of 1800 (object + int + four floats each). */
typedef struct {
    PyObject_HEAD
    long long id;
    double u0, v0, u1, v1;
} FMObject;

static PyMemberDef FM_members[] = {
    {"id", T_LONGLONG, offsetof(FMObject, id), 0, "feature id"},
    {"u0", T_DOUBLE, offsetof(FMObject, u0), 0, "cam0 normalized x"},
    {"v0", T_DOUBLE, offsetof(FMObject, v0), 0, "cam0 normalized y"},
    {"u1", T_DOUBLE, offsetof(FMObject, u1), 0, "cam1 normalized x"},
    {"v1", T_DOUBLE, offsetof(FMObject, v1), 0, "cam1 normalized y"},
    {NULL}};

static int FM_init(FMObject* self, PyObject* args, PyObject* kw) {
    static char* names[] = {"id", "u0", "v0", "u1", "v1", NULL};
    long long id = 0;
    double u0 = 0, v0 = 0, u1 = 0, v1 = 0;
    if (!PyArg_ParseTupleAndKeywords(args, kw, "|Ldddd", names, &id, &u0, &v0, &u1, &v1)) return -1;
    self->id = id;
    self->u0 = u0;
    self->v0 = v0;
    self->u1 = u1;
    self->v1 = v1;
    return 0;
}

static PyObject* FM_repr(FMObject* self) {
    char buf[200];
    snprintf(buf, sizeof buf, "FeatureMeasurement(id=%lld, u0=%.9g, v0=%.9g, u1=%.9g, v1=%.9g)", self->id, self->u0, self->v0,
             self->u1, self->v1);
    return PyUnicode_FromString(buf);
}

/* A frame publishes a few hundred of these and the consumer drops them a frame later: freed objects of the exact type
 * wait on a free list (as CPython does for its own small types) instead of going back to the allocator. */
#ifndef FM_FREELIST_MAX
#define FM_FREELIST_MAX 16384
#endif
static PyTypeObject FMType;
static FMObject* fm_free[FM_FREELIST_MAX + 1];
static int fm_nfree = 0;

static void FM_dealloc(FMObject* self) {
    if (Py_IS_TYPE((PyObject*)self, &FMType) && fm_nfree < FM_FREELIST_MAX) {       /* subclasses go the normal way */
        fm_free[fm_nfree++] = self;
        return;
    }
    Py_TYPE(self)->tp_free((PyObject*)self);
}

static PyTypeObject FMType = {
    PyVarObject_HEAD_INIT(NULL, 0).tp_name = "image_processing.FeatureMeasurement",
    .tp_dealloc = (destructor)FM_dealloc,
    .tp_basicsize = sizeof(FMObject),
    .tp_flags = Py_TPFLAGS_DEFAULT | Py_TPFLAGS_BASETYPE,
    .tp_doc = "Stereo measurement of one feature in normalized coordinates: id, u0, v0, u1, v1.",
    .tp_members = FM_members,
    .tp_init = (initproc)FM_init,
    .tp_new = PyType_GenericNew,
    .tp_repr = (reprfunc)FM_repr,
};

/* Any other class with id/u0/v0/u1/v1 attributes can be requested instead (generic setattr path). */
static PyObject* g_names[5];

static PyObject* make_feature(PyTypeObject* tp, long long id, const double* m) {
    if (tp == &FMType) {
        FMObject* o;
        if (fm_nfree) {
            o = fm_free[--fm_nfree];
            _Py_NewReference((PyObject*)o);
        } else {
            o = (FMObject*)FMType.tp_alloc(&FMType, 0);
            if (!o) return NULL;
        }
        o->id = id;
        o->u0 = m[0];
        o->v0 = m[1];
        o->u1 = m[2];
        o->v1 = m[3];
        return (PyObject*)o;
    }
    PyObject* vals[5];
    vals[0] = PyLong_FromLongLong(id);
    vals[1] = PyFloat_FromDouble(m[0]);
    vals[2] = PyFloat_FromDouble(m[1]);
    vals[3] = PyFloat_FromDouble(m[2]);
    vals[4] = PyFloat_FromDouble(m[3]);
    PyObject* o = NULL;
    if (vals[0] && vals[1] && vals[2] && vals[3] && vals[4]) {
        o = PyObject_CallNoArgs((PyObject*)tp);
        if (o) {
            for (int i = 0; i < 5; ++i) {
                if (PyObject_SetAttr(o, g_names[i], vals[i]) < 0) {
                    Py_CLEAR(o);
                    break;
                }
            }
        }
    }
    for (int i = 0; i < 5; ++i) Py_XDECREF(vals[i]);
    return o;
}

static int resolve_layout(PyTypeObject* tp) { (void)tp; return 0; }

static PyObject* build_list(PyTypeObject* tp, const avb_frame_header* h, const int64_t* ids, const double* meas) {
    const Py_ssize_t n = (Py_ssize_t)h->n_features;
    PyObject* list = PyList_New(n);
    if (!list) return NULL;
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject* o = make_feature(tp, (long long)ids[i], meas + 4 * i);
        if (!o) {
            Py_DECREF(list);
            return NULL;
        }
        PyList_SET_ITEM(list, i, o);
    }
    return list;
}

static PyObject* header_tuple(const avb_frame_header* h) {
    return Py_BuildValue("(LLiiiiiiii)", (long long)h->n_features, (long long)h->next_feature_id, h->before_tracking,
                         h->after_tracking, h->after_matching, h->after_ransac, h->has_new, h->n_fast, h->n_candidates,
                         h->frame_index);
}

static int copy_image(uint8_t* dst, Py_buffer* b, int W, int H) {
    if (b->ndim != 2 || b->itemsize != 1 || b->shape[0] != H || b->shape[1] != W || b->strides[1] != 1) return -1;
    const uint8_t* src = (const uint8_t*)b->buf;
    if (b->strides[0] == W) {
        memcpy(dst, src, (size_t)W * H);
    } else {
        for (int y = 0; y < H; ++y) memcpy(dst + (size_t)y * W, src + (size_t)y * b->strides[0], W);
    }
    return 0;
}

/* R|None -> (R[18], haveR, haveR1); 0 on success, -1 with a Python error set */
static int parse_R(PyObject* Robj, double* R, int* haveR, int* haveR1) {
    *haveR = *haveR1 = 0;
    if (Robj == Py_None) return 0;
    Py_buffer bR;
    if (PyObject_GetBuffer(Robj, &bR, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) < 0) return -1;
    if ((bR.len != 72 && bR.len != 144) || bR.itemsize != 8) {
        PyBuffer_Release(&bR);
        PyErr_SetString(PyExc_ValueError, "R must be a C-contiguous float64 array: 3x3 (cam0_R_p_c) or 2x3x3 (cam0, cam1)");
        return -1;
    }
    memcpy(R, bR.buf, (size_t)bR.len);
    *haveR1 = bR.len == 144;
    *haveR = 1;
    PyBuffer_Release(&bR);
    return 0;
}

/* (features, header tuple) of stream 0 from the result block of the frame just processed */
static PyObject* frame_result(avb_ctx* ctx, PyObject* tpobj) {
    const avb_frame_header* h;
    const int64_t* ids;
    const double* meas;
    const int rc = avb_get_result(ctx, 0, &h, &ids, &meas);
    if (rc != AVB_OK) {
        PyErr_Format(PyExc_RuntimeError, "libavb error %d: %s", rc, avb_last_error(ctx));
        return NULL;
    }
    resolve_layout((PyTypeObject*)tpobj);
    PyObject* list = build_list((PyTypeObject*)tpobj, h, ids, meas);
    if (!list) return NULL;
    PyObject* hd = header_tuple(h);
    if (!hd) {
        Py_DECREF(list);
        return NULL;
    }
    PyObject* out = PyTuple_Pack(2, list, hd);
    Py_DECREF(list);
    Py_DECREF(hd);
    return out;
}

/* The two images of a single-stream frame through `call(ctx, &p0, &p1, stride, ...)`; returns the libavb status or -100
 * with a Python error set. */
typedef int (*frame_call)(avb_ctx*, const uint8_t* const*, const uint8_t* const*, int, const double*, const double*);
static int submit_only(avb_ctx* ctx, const uint8_t* const* p0, const uint8_t* const* p1, int stride, const double* R0, const double* R1) {
    (void)R0;
    (void)R1;
    return avb_submit_images(ctx, p0, p1, stride);
}
static int with_images(avb_ctx* ctx, PyObject* img0, PyObject* img1, frame_call call, const double* R0, const double* R1) {
    int W = 0, H = 0, S = 0;
    avb_get_geometry(ctx, &W, &H, &S);
    if (S != 1) {
        PyErr_SetString(PyExc_RuntimeError, "process_frame drives single-stream contexts");
        return -100;
    }
    Py_buffer b0, b1;
    if (PyObject_GetBuffer(img0, &b0, PyBUF_STRIDES) < 0) return -100;
    if (PyObject_GetBuffer(img1, &b1, PyBUF_STRIDES) < 0) {
        PyBuffer_Release(&b0);
        return -100;
    }
    int rc = 0;
    const int bad = b0.ndim != 2 || b0.itemsize != 1 || b0.shape[0] != H || b0.shape[1] != W || b0.strides[1] != 1 ||
                    b1.ndim != 2 || b1.itemsize != 1 || b1.shape[0] != H || b1.shape[1] != W || b1.strides[1] != 1 ||
                    b0.strides[0] != b1.strides[0] || b0.strides[0] < W;
    if (!bad) {
        const uint8_t* p0 = (const uint8_t*)b0.buf;
        const uint8_t* p1 = (const uint8_t*)b1.buf;
        const int stride = (int)b0.strides[0];
        Py_BEGIN_ALLOW_THREADS
        rc = call(ctx, &p0, &p1, stride, R0, R1);      /* staging + pipelined H2D [+ frame + D2H] */
        Py_END_ALLOW_THREADS
    }
    PyBuffer_Release(&b0);
    PyBuffer_Release(&b1);
    if (bad) {
        PyErr_Format(PyExc_RuntimeError, "images must be (%d, %d) uint8 arrays with unit column stride and equal row stride", H, W);
        return -100;
    }
    if (rc != AVB_OK) {
        PyErr_Format(PyExc_RuntimeError, "libavb error %d: %s", rc, avb_last_error(ctx));
        return -100;
    }
    return AVB_OK;
}

/* process_frame(ctx:int, img0, img1, R|None, fm_type) */
static PyObject* py_process_frame(PyObject* self, PyObject* args) {
    unsigned long long handle;
    PyObject *img0, *img1, *Robj, *tpobj;
    if (!PyArg_ParseTuple(args, "KOOOO", &handle, &img0, &img1, &Robj, &tpobj)) return NULL;
    avb_ctx* ctx = (avb_ctx*)(uintptr_t)handle;
    if (!ctx || !PyType_Check(tpobj)) {
        PyErr_SetString(PyExc_TypeError, "process_frame(ctx, img0, img1, R|None, FeatureMeasurement)");
        return NULL;
    }
    double R[18];                       /* cam0_R_p_c [, cam1_R_p_c] */
    int haveR, haveR1;
    if (parse_R(Robj, R, &haveR, &haveR1) < 0) return NULL;
    if (with_images(ctx, img0, img1, avb_process_frame, haveR ? R : NULL, haveR1 ? R + 9 : NULL) != AVB_OK) return NULL;
    return frame_result(ctx, tpobj);
}

/* submit_images(ctx:int, img0, img1): the first half of process_frame (avb_submit_images): the copies are on their way
 * and, in the steady state of a single stream, FAST runs behind cam0's; the arrays must stay alive until finish_frame. */
static PyObject* py_submit_images(PyObject* self, PyObject* args) {
    unsigned long long handle;
    PyObject *img0, *img1;
    if (!PyArg_ParseTuple(args, "KOO", &handle, &img0, &img1)) return NULL;
    avb_ctx* ctx = (avb_ctx*)(uintptr_t)handle;
    if (!ctx) {
        PyErr_SetString(PyExc_TypeError, "submit_images(ctx, img0, img1)");
        return NULL;
    }
    if (with_images(ctx, img0, img1, submit_only, NULL, NULL) != AVB_OK) return NULL;
    Py_RETURN_NONE;
}

/* finish_frame(ctx:int, R|None, fm_type) -> (features, header tuple): the second half (avb_process_submitted) */
static PyObject* py_finish_frame(PyObject* self, PyObject* args) {
    unsigned long long handle;
    PyObject *Robj, *tpobj;
    if (!PyArg_ParseTuple(args, "KOO", &handle, &Robj, &tpobj)) return NULL;
    avb_ctx* ctx = (avb_ctx*)(uintptr_t)handle;
    if (!ctx || !PyType_Check(tpobj)) {
        PyErr_SetString(PyExc_TypeError, "finish_frame(ctx, R|None, FeatureMeasurement)");
        return NULL;
    }
    double R[18];
    int haveR, haveR1, rc;
    if (parse_R(Robj, R, &haveR, &haveR1) < 0) return NULL;
    Py_BEGIN_ALLOW_THREADS
    rc = avb_process_submitted(ctx, haveR ? R : NULL, haveR1 ? R + 9 : NULL);
    Py_END_ALLOW_THREADS
    if (rc != AVB_OK) {
        PyErr_Format(PyExc_RuntimeError, "libavb error %d: %s", rc, avb_last_error(ctx));
        return NULL;
    }
    return frame_result(ctx, tpobj);
}

/* process_frames(ctx:int, imgs0:sequence, imgs1:sequence, R|None (S x 9 doubles, C-contiguous), fm_type)
 *   -> [(features, header), ...] for the S lock-stepped streams of a context (one launch chain for all of them). */
static PyObject* py_process_frames(PyObject* self, PyObject* args) {
    unsigned long long handle;
    PyObject *l0, *l1, *Robj, *tpobj;
    if (!PyArg_ParseTuple(args, "KOOOO", &handle, &l0, &l1, &Robj, &tpobj)) return NULL;
    avb_ctx* ctx = (avb_ctx*)(uintptr_t)handle;
    if (!ctx || !PyType_Check(tpobj)) {
        PyErr_SetString(PyExc_TypeError, "process_frames(ctx, imgs0, imgs1, R|None, FeatureMeasurement)");
        return NULL;
    }
    int W = 0, H = 0, S = 0;
    avb_get_geometry(ctx, &W, &H, &S);
    PyObject* s0 = PySequence_Fast(l0, "imgs0 must be a sequence");
    if (!s0) return NULL;
    PyObject* s1 = PySequence_Fast(l1, "imgs1 must be a sequence");
    if (!s1) {
        Py_DECREF(s0);
        return NULL;
    }
    PyObject* out = NULL;
    Py_buffer bR;
    int haveR = 0;
    if (PySequence_Fast_GET_SIZE(s0) != S || PySequence_Fast_GET_SIZE(s1) != S) {
        PyErr_Format(PyExc_ValueError, "expected %d images per camera", S);
        goto done;
    }
    if (Robj != Py_None) {
        if (PyObject_GetBuffer(Robj, &bR, PyBUF_C_CONTIGUOUS) < 0) goto done;
        haveR = 1;
        if ((bR.len != (Py_ssize_t)S * 72 && bR.len != (Py_ssize_t)S * 144) || bR.itemsize != 8) {
            PyErr_Format(PyExc_ValueError, "R must be (%d,3,3) float64 (cam0_R_p_c) or (2,%d,3,3) (cam0, cam1), C-contiguous", S, S);
            goto done;
        }
    }
    {
        uint8_t* st = avb_input_staging(ctx);
        const size_t ib = (size_t)W * H;
        for (int s = 0; s < S; ++s) {
            for (int cam = 0; cam < 2; ++cam) {
                Py_buffer b;
                PyObject* img = PySequence_Fast_GET_ITEM(cam ? s1 : s0, s);
                if (PyObject_GetBuffer(img, &b, PyBUF_STRIDES) < 0) goto done;
                int bad;
                Py_BEGIN_ALLOW_THREADS
                bad = copy_image(st + ((size_t)s * 2 + cam) * ib, &b, W, H);
                Py_END_ALLOW_THREADS
                PyBuffer_Release(&b);
                if (bad) {
                    PyErr_Format(PyExc_RuntimeError, "images must be (%d, %d) uint8 arrays with unit column stride", H, W);
                    goto done;
                }
            }
        }
        int rc;
        Py_BEGIN_ALLOW_THREADS
        rc = avb_process_frame(ctx, NULL, NULL, W, haveR ? (const double*)bR.buf : NULL,
                               haveR && bR.len == (Py_ssize_t)S * 144 ? (const double*)bR.buf + (size_t)S * 9 : NULL);
        Py_END_ALLOW_THREADS
        if (rc != AVB_OK) {
            PyErr_Format(PyExc_RuntimeError, "libavb error %d: %s", rc, avb_last_error(ctx));
            goto done;
        }
        resolve_layout((PyTypeObject*)tpobj);
        PyObject* res = PyList_New(S);
        if (!res) goto done;
        for (int s = 0; s < S; ++s) {
            const avb_frame_header* h;
            const int64_t* ids;
            const double* meas;
            avb_get_result(ctx, s, &h, &ids, &meas);
            PyObject* list = build_list((PyTypeObject*)tpobj, h, ids, meas);
            PyObject* hd = list ? header_tuple(h) : NULL;
            PyObject* pair = (list && hd) ? PyTuple_Pack(2, list, hd) : NULL;
            Py_XDECREF(list);
            Py_XDECREF(hd);
            if (!pair) {
                Py_DECREF(res);
                goto done;
            }
            PyList_SET_ITEM(res, s, pair);
        }
        out = res;
    }
done:
    if (haveR) PyBuffer_Release(&bR);
    Py_DECREF(s0);
    Py_DECREF(s1);
    return out;
}

/* features_from_result(ctx:int, s:int, fm_type) -> (features, header tuple): list construction only, for contexts
 * driven through avb_process_frame* elsewhere (multi-stream drivers). */
static PyObject* py_features_from_result(PyObject* self, PyObject* args) {
    unsigned long long handle;
    int s;
    PyObject* tpobj;
    if (!PyArg_ParseTuple(args, "KiO", &handle, &s, &tpobj)) return NULL;
    avb_ctx* ctx = (avb_ctx*)(uintptr_t)handle;
    if (!ctx || !PyType_Check(tpobj)) {
        PyErr_SetString(PyExc_TypeError, "features_from_result(ctx, stream, FeatureMeasurement)");
        return NULL;
    }
    const avb_frame_header* h;
    const int64_t* ids;
    const double* meas;
    int rc = avb_get_result(ctx, s, &h, &ids, &meas);
    if (rc != AVB_OK) {
        PyErr_Format(PyExc_RuntimeError, "libavb error %d: %s", rc, avb_last_error(ctx));
        return NULL;
    }
    resolve_layout((PyTypeObject*)tpobj);
    PyObject* list = build_list((PyTypeObject*)tpobj, h, ids, meas);
    if (!list) return NULL;
    PyObject* hd = header_tuple(h);
    if (!hd) {
        Py_DECREF(list);
        return NULL;
    }
    PyObject* out = PyTuple_Pack(2, list, hd);
    Py_DECREF(list);
    Py_DECREF(hd);
    return out;
}

/* features_from_arrays(ids: int64[n], meas: float64[n, 4], fm_type) -> list: the list construction alone, from caller
 * arrays (what a sweep's estimator side holds, estimator_pool.py). */
static PyObject* py_features_from_arrays(PyObject* self, PyObject* args) {
    Py_buffer bi, bm;
    PyObject* tpobj;
    if (!PyArg_ParseTuple(args, "y*y*O", &bi, &bm, &tpobj)) return NULL;
    PyObject* out = NULL;
    const Py_ssize_t n = bi.len / 8;
    if (!PyType_Check(tpobj) || bi.len % 8 || bm.len != n * 32) {
        PyErr_SetString(PyExc_ValueError, "features_from_arrays(ids int64[n], meas float64[n, 4] (both C-contiguous), type)");
    } else {
        avb_frame_header h;
        memset(&h, 0, sizeof h);
        h.n_features = n;
        out = build_list((PyTypeObject*)tpobj, &h, (const int64_t*)bi.buf, (const double*)bm.buf);
    }
    PyBuffer_Release(&bi);
    PyBuffer_Release(&bm);
    return out;
}

/* ---- gyro integration ---------------------------------------------------------------------------------- */

static PyObject *s_timestamp, *s_angular_velocity;

static int msg_time(PyObject* msg, double* t) {
    PyObject* v = PyObject_GetAttr(msg, s_timestamp);
    if (!v) return -1;
    *t = PyFloat_AsDouble(v);
    Py_DECREF(v);
    return (*t == -1.0 && PyErr_Occurred()) ? -1 : 0;
}

static int msg_gyro(PyObject* msg, double* w) {
    PyObject* v = PyObject_GetAttr(msg, s_angular_velocity);
    if (!v) return -1;
    Py_buffer b;
    int rc = -1;
    if (PyObject_GetBuffer(v, &b, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) == 0) {
        if (b.len == 24 && b.itemsize == 8 && b.format && (b.format[0] == 'd' || (b.format[0] && b.format[1] == 'd'))) {
            memcpy(w, b.buf, 24);
            rc = 0;
        }
        PyBuffer_Release(&b);
    } else {
        PyErr_Clear();
    }
    if (rc != 0) {                                   /* any 3-sequence of numbers */
        PyObject* seq = PySequence_Fast(v, "angular_velocity must be a 3-vector");
        if (seq && PySequence_Fast_GET_SIZE(seq) == 3) {
            rc = 0;
            for (int i = 0; i < 3; ++i) {
                w[i] = PyFloat_AsDouble(PySequence_Fast_GET_ITEM(seq, i));
                if (w[i] == -1.0 && PyErr_Occurred()) rc = -1;
            }
        } else if (seq) {
            PyErr_SetString(PyExc_ValueError, "angular_velocity must be a 3-vector");
        }
        Py_XDECREF(seq);
    }
    Py_DECREF(v);
    return rc;
}

/* cv2.Rodrigues(v)[0] transposed, written row-major into out[9] */
static void rodrigues_T(const double* v, double* out) {
    const double theta = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double R[9];
    if (theta < 2.220446049250313e-16) {
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    } else {
        const double r[3] = {v[0] / theta, v[1] / theta, v[2] / theta};
        const double c = cos(theta), s = sin(theta), c1 = 1.0 - c;
        const double rx[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R[i * 3 + j] = (c * (i == j ? 1.0 : 0.0) + c1 * (r[i] * r[j])) + s * rx[i * 3 + j];
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) out[i * 3 + j] = R[j * 3 + i];
}

static int get_mat9(PyObject* o, double* m, int writable, Py_buffer* keep) {
    if (PyObject_GetBuffer(o, keep, (writable ? PyBUF_WRITABLE : 0) | PyBUF_C_CONTIGUOUS) < 0) return -1;
    if (keep->len != 72 || keep->itemsize != 8) {
        PyBuffer_Release(keep);
        PyErr_SetString(PyExc_ValueError, "expected a C-contiguous 3x3 float64 array");
        return -1;
    }
    if (m) memcpy(m, keep->buf, 72);
    return 0;
}

/* integrate_imu(buffer:list, t_prev, t_curr, R_cam0_imu, R_cam1_imu, out0, out1) -> end | -1 */
static PyObject* py_integrate_imu(PyObject* self, PyObject* args) {
    PyObject *buf, *R0o, *R1o, *o0, *o1;
    double t_prev, t_curr;
    if (!PyArg_ParseTuple(args, "O!ddOOOO", &PyList_Type, &buf, &t_prev, &t_curr, &R0o, &R1o, &o0, &o1)) return NULL;
    double R0[9], R1[9];
    Py_buffer b;
    if (get_mat9(R0o, R0, 0, &b) < 0) return NULL;
    PyBuffer_Release(&b);
    if (get_mat9(R1o, R1, 0, &b) < 0) return NULL;
    PyBuffer_Release(&b);
    const Py_ssize_t n = PyList_GET_SIZE(buf);
    Py_ssize_t begin = -1, end = -1;
    const double lo = t_prev - 0.01, hi = t_curr - 0.004;
    for (Py_ssize_t i = 0; i < n; ++i) {
        double t;
        if (msg_time(PyList_GET_ITEM(buf, i), &t) < 0) return NULL;
        if (begin < 0 && t >= lo) begin = i;
        if (end < 0 && t >= hi) end = i;
        if (begin >= 0 && end >= 0) break;
    }
    double out0[9], out1[9];
    long ret = -1;
    if (begin < 0 || end < 0) {
        for (int i = 0; i < 9; ++i) out0[i] = out1[i] = (i % 4 == 0) ? 1.0 : 0.0;
    } else {
        double w[3] = {0, 0, 0};
        for (Py_ssize_t i = begin; i < end; ++i) {
            double g[3];
            if (msg_gyro(PyList_GET_ITEM(buf, i), g) < 0) return NULL;
            w[0] += g[0];
            w[1] += g[1];
            w[2] += g[2];
        }
        if (end - begin > 0) {
            const double cnt = (double)(end - begin);
            w[0] /= cnt;
            w[1] /= cnt;
            w[2] /= cnt;
        }
        const double dt = t_curr - t_prev;
        double v0[3], v1[3];
        for (int j = 0; j < 3; ++j) {               /* (R^T w) * dt */
            v0[j] = ((R0[0 * 3 + j] * w[0] + R0[1 * 3 + j] * w[1]) + R0[2 * 3 + j] * w[2]) * dt;
            v1[j] = ((R1[0 * 3 + j] * w[0] + R1[1 * 3 + j] * w[1]) + R1[2 * 3 + j] * w[2]) * dt;
        }
        rodrigues_T(v0, out0);
        rodrigues_T(v1, out1);
        ret = (long)end;
    }
    if (get_mat9(o0, NULL, 1, &b) < 0) return NULL;
    memcpy(b.buf, out0, 72);
    PyBuffer_Release(&b);
    if (get_mat9(o1, NULL, 1, &b) < 0) return NULL;
    memcpy(b.buf, out1, 72);
    PyBuffer_Release(&b);
    return PyLong_FromLong(ret);
}

/* ---- PNG scanline reconstruction (EuRoC cam0/cam1 frames are 8-bit grayscale PNGs) ------------------------- */

/* png_unfilter(raw: bytes-like (h * (1 + w*bpp)), w, h, bpp, out: writable buffer of h*w*bpp bytes)
 * Undoes the per-row PNG filters 0-4 (None, Sub, Up, Average, Paeth; PNG spec section 6) into `out`, which may be
 * libavb's pinned staging block.  The inflate step stays in Python's zlib. */
static PyObject* py_png_unfilter(PyObject* self, PyObject* args) {
    Py_buffer raw, out;
    int w, h, bpp;
    if (!PyArg_ParseTuple(args, "y*iiiw*", &raw, &w, &h, &bpp, &out)) return NULL;
    const Py_ssize_t stride = (Py_ssize_t)w * bpp;
    int err = 0;
    if (w <= 0 || h <= 0 || bpp <= 0 || raw.len != (Py_ssize_t)h * (stride + 1) || out.len < (Py_ssize_t)h * stride) {
        err = 1;
    } else {
        const unsigned char* src = (const unsigned char*)raw.buf;
        unsigned char* dst = (unsigned char*)out.buf;
        Py_BEGIN_ALLOW_THREADS
        for (int y = 0; y < h && !err; ++y) {
            const unsigned char ft = src[(Py_ssize_t)y * (stride + 1)];
            const unsigned char* in = src + (Py_ssize_t)y * (stride + 1) + 1;
            unsigned char* cur = dst + (Py_ssize_t)y * stride;
            const unsigned char* up = y ? cur - stride : NULL;
            switch (ft) {
                case 0:
                    memcpy(cur, in, stride);
                    break;
                case 1:
                    for (Py_ssize_t x = 0; x < stride; ++x) cur[x] = (unsigned char)(in[x] + (x >= bpp ? cur[x - bpp] : 0));
                    break;
                case 2:
                    for (Py_ssize_t x = 0; x < stride; ++x) cur[x] = (unsigned char)(in[x] + (up ? up[x] : 0));
                    break;
                case 3:
                    for (Py_ssize_t x = 0; x < stride; ++x) {
                        const int a = x >= bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0;
                        cur[x] = (unsigned char)(in[x] + ((a + b) >> 1));
                    }
                    break;
                case 4:
                    for (Py_ssize_t x = 0; x < stride; ++x) {
                        const int a = x >= bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
                        const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
                        const int pr = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                        cur[x] = (unsigned char)(in[x] + pr);
                    }
                    break;
                default:
                    err = 2;
            }
        }
        Py_END_ALLOW_THREADS
    }
    PyBuffer_Release(&raw);
    PyBuffer_Release(&out);
    if (err) {
        PyErr_SetString(PyExc_ValueError, err == 1 ? "png_unfilter: size mismatch" : "png_unfilter: unknown filter type");
        return NULL;
    }
    Py_RETURN_NONE;
}

static PyMethodDef methods[] = {
    {"process_frame", py_process_frame, METH_VARARGS, "One stereo frame: host images in, (FeatureMeasurement list, header) out."},
    {"submit_images", py_submit_images, METH_VARARGS, "First half of process_frame: image intake (copies, cam0-only kernels)."},
    {"finish_frame", py_finish_frame, METH_VARARGS, "Second half of process_frame: rotations in, (FeatureMeasurement list, header) out."},
    {"process_frames", py_process_frames, METH_VARARGS, "One stereo frame for every stream of a multi-stream context."},
    {"features_from_result", py_features_from_result, METH_VARARGS, "FeatureMeasurement list of stream s from the last frame."},
    {"features_from_arrays", py_features_from_arrays, METH_VARARGS, "FeatureMeasurement list from ids / measurement arrays."},
    {"png_unfilter", py_png_unfilter, METH_VARARGS, "Undo PNG row filters into a caller buffer."},
    {"integrate_imu", py_integrate_imu, METH_VARARGS, "Gyro window integration (imu_processor.py:28-67)."},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_avbhost", "libavb host-side frame driver", -1, methods};

PyMODINIT_FUNC PyInit__avbhost(void) {
    static const char* names[5] = {"id", "u0", "v0", "u1", "v1"};
    for (int i = 0; i < 5; ++i) g_names[i] = PyUnicode_InternFromString(names[i]);
    s_timestamp = PyUnicode_InternFromString("timestamp");
    s_angular_velocity = PyUnicode_InternFromString("angular_velocity");
    if (PyType_Ready(&FMType) < 0) return NULL;
    PyObject* m = PyModule_Create(&moddef);
    if (!m) return NULL;
    Py_INCREF(&FMType);
    if (PyModule_AddObject(m, "FeatureMeasurement", (PyObject*)&FMType) < 0) {
        Py_DECREF(&FMType);
        Py_DECREF(m);
        return NULL;
    }
    return m;
}
