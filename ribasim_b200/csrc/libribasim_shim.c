/* libribasim_shim.c — `libribasim`-compatible shared library on top of the B200 path.
 *
 * Exports exactly the 19 C symbols of the reference's library
 * (/root/reference/build/libribasim.jl:74-212; declared in include/libribasim.h) with its
 * conventions: every function returns 0 on success and 1 on failure (`@try_c`, :40-52), the
 * message of the last failure is read with get_last_bmi_error (:184-189), strings are written
 * NUL-terminated into caller-owned buffers sized by the constants of
 * python/ribasim_api/ribasim_api/ribasim_api.py:8-19.
 *
 * The reference's library is a Julia system image with a global `model`; this one embeds
 * CPython and forwards to ribasim_b200/capi.py, whose global model is a `RibasimBmi` over the
 * device-resident `Model` (the CUDA library librbs.so does the stepping).  Loaded into a process
 * that already runs Python (ctypes / xmipy) it attaches to that interpreter; loaded by a C or Rust
 * host (build/cli/src/main.rs dlopens the library the same way) it starts one.
 */
#define PY_SSIZE_T_CLEAN
#define _GNU_SOURCE
#include <Python.h>
#include <dlfcn.h>
#include <libgen.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#ifndef RBS_REPO_ROOT
#define RBS_REPO_ROOT "."
#endif
#ifndef RBS_SITE_PACKAGES
#define RBS_SITE_PACKAGES ""
#endif

#define RBS_EXPORT __attribute__((visibility("default")))
#define BMI_LENERRMESSAGE 1025

static char g_last_error[BMI_LENERRMESSAGE] = "";
static PyObject* g_capi = NULL; /* ribasim_b200.capi */
static int g_own_interpreter = 0;

static void set_error(const char* msg) {
    strncpy(g_last_error, msg ? msg : "unknown error", BMI_LENERRMESSAGE - 1);
    g_last_error[BMI_LENERRMESSAGE - 1] = 0;
}

/* message of the pending Python exception -> last error (sprint(showerror, e)) */
static void take_py_error(void) {
    PyObject *type = NULL, *value = NULL, *tb = NULL;
    PyErr_Fetch(&type, &value, &tb);
    PyErr_NormalizeException(&type, &value, &tb);
    PyObject* s = value ? PyObject_Str(value) : NULL;
    const char* c = s ? PyUnicode_AsUTF8(s) : NULL;
    if (c && *c) set_error(c);
    else if (type) {
        PyObject* n = PyObject_GetAttrString(type, "__name__");
        const char* cn = n ? PyUnicode_AsUTF8(n) : NULL;
        set_error(cn ? cn : "Python error");
        Py_XDECREF(n);
    } else set_error("Python error");
    PyErr_Clear();
    Py_XDECREF(s);
    Py_XDECREF(type);
    Py_XDECREF(value);
    Py_XDECREF(tb);
}

/* directory that holds the `ribasim_b200` package: two levels above this library, wherever the
 * tree was copied to (falls back to the path of the build) */
static const char* repo_root(void) {
    static char root[4096] = "";
    if (root[0]) return root;
    Dl_info info;
    if (dladdr((void*)&repo_root, &info) && info.dli_fname) {
        char tmp[4096], *rp = realpath(info.dli_fname, tmp);
        if (rp) {
            char* pkg = dirname(rp);        /* .../ribasim_b200 */
            char* top = dirname(pkg);       /* repo root */
            strncpy(root, top, sizeof root - 1);
            return root;
        }
    }
    strncpy(root, RBS_REPO_ROOT, sizeof root - 1);
    return root;
}

static void add_sys_path(const char* dir) {
    if (!dir || !*dir) return;
    PyObject* sys_path = PySys_GetObject("path"); /* borrowed */
    PyObject* d = PyUnicode_FromString(dir);
    if (sys_path && d && !PySequence_Contains(sys_path, d)) PyList_Insert(sys_path, 0, d);
    Py_XDECREF(d);
}

/* interpreter + module; returns with the GIL held (state in *gil), or -1 */
static int enter(PyGILState_STATE* gil) {
    if (!Py_IsInitialized()) {
        Py_InitializeEx(0);
        g_own_interpreter = 1;
        /* the embedded interpreter does not know the virtual environment of the build */
        const char* sp = getenv("RIBASIM_B200_SITE_PACKAGES");
        add_sys_path(sp && *sp ? sp : RBS_SITE_PACKAGES);
        /* release the GIL taken by Py_Initialize so that PyGILState_Ensure works from any thread */
        PyEval_SaveThread();
    }
    *gil = PyGILState_Ensure();
    if (!g_capi) {
        add_sys_path(repo_root());
        g_capi = PyImport_ImportModule("ribasim_b200.capi");
        if (!g_capi) {
            take_py_error();
            PyGILState_Release(*gil);
            return -1;
        }
    }
    return 0;
}

/* call capi.<fn>(args...) -> new reference or NULL (error recorded) */
static PyObject* call(const char* fn, const char* fmt, ...) {
    PyObject* f = PyObject_GetAttrString(g_capi, fn);
    if (!f) { take_py_error(); return NULL; }
    PyObject* args;
    if (fmt && *fmt) {
        va_list va;
        va_start(va, fmt);
        args = Py_VaBuildValue(fmt, va);
        va_end(va);
        if (args && !PyTuple_Check(args)) {
            PyObject* t = PyTuple_Pack(1, args);
            Py_DECREF(args);
            args = t;
        }
    } else args = PyTuple_New(0);
    PyObject* r = args ? PyObject_CallObject(f, args) : NULL;
    Py_XDECREF(args);
    Py_DECREF(f);
    if (!r) take_py_error();
    return r;
}

static void write_cstring(char* dest, const char* src) { /* unsafe_write_to_cstring!, :227-233 */
    if (dest) memcpy(dest, src, strlen(src) + 1);
}

/* int-returning call with optional string / double argument */
static int call_int(const char* fn, const char* fmt, const void* arg, double darg) {
    PyGILState_STATE gil;
    if (enter(&gil)) return 1;
    PyObject* r = !fmt ? call(fn, NULL) : (fmt[0] == 's' ? call(fn, "(s)", (const char*)arg) : call(fn, "(d)", darg));
    int rc = 1;
    if (r) {
        long v = PyLong_AsLong(r);
        if (v == -1 && PyErr_Occurred()) take_py_error();
        else rc = (int)v;
        Py_DECREF(r);
    }
    PyGILState_Release(gil);
    return rc;
}

static int call_double_out(const char* fn, double* out) {
    PyGILState_STATE gil;
    if (enter(&gil)) return 1;
    PyObject* r = call(fn, NULL);
    int rc = 1;
    if (r) {
        double v = PyFloat_AsDouble(r);
        if (v == -1.0 && PyErr_Occurred()) take_py_error();
        else { if (out) *out = v; rc = 0; }
        Py_DECREF(r);
    }
    PyGILState_Release(gil);
    return rc;
}

static int call_string_out(const char* fn, const char* name, char* out) {
    PyGILState_STATE gil;
    if (enter(&gil)) return 1;
    PyObject* r = name ? call(fn, "(s)", name) : call(fn, NULL);
    int rc = 1;
    if (r) {
        const char* c = PyUnicode_AsUTF8(r);
        if (!c) take_py_error();
        else { write_cstring(out, c); rc = 0; }
        Py_DECREF(r);
    }
    PyGILState_Release(gil);
    return rc;
}

/* ---- the 19 exported symbols (libribasim.jl:74-212) ------------------------------------------- */
RBS_EXPORT int initialize(const char* path) { return call_int("initialize", "s", path, 0.0); }
RBS_EXPORT int finalize(void) { return call_int("finalize", NULL, NULL, 0.0); }
RBS_EXPORT int update(void) { return call_int("update", NULL, NULL, 0.0); }
RBS_EXPORT int update_until(double time) { return call_int("update_until", "d", NULL, time); }
RBS_EXPORT int update_subgrid_level(void) { return call_int("update_subgrid_level", NULL, NULL, 0.0); }
RBS_EXPORT int get_current_time(double* time) { return call_double_out("get_current_time", time); }
RBS_EXPORT int get_start_time(double* time) { return call_double_out("get_start_time", time); }
RBS_EXPORT int get_end_time(double* time) { return call_double_out("get_end_time", time); }
RBS_EXPORT int get_time_step(double* time_step) { return call_double_out("get_time_step", time_step); }
RBS_EXPORT int get_var_type(const char* name, char* var_type) { return call_string_out("get_var_type", name, var_type); }

RBS_EXPORT int get_var_rank(const char* name, int* rank) {
    PyGILState_STATE gil;
    if (enter(&gil)) return 1;
    PyObject* r = call("get_var_rank", "(s)", name);
    int rc = 1;
    if (r) {
        long v = PyLong_AsLong(r);
        if (v == -1 && PyErr_Occurred()) take_py_error();
        else { if (rank) *rank = (int)v; rc = 0; }
        Py_DECREF(r);
    }
    PyGILState_Release(gil);
    return rc;
}

RBS_EXPORT int get_value_ptr(const char* name, void** value_ptr) {
    PyGILState_STATE gil;
    if (enter(&gil)) return 1;
    PyObject* r = call("value_address", "(s)", name);
    int rc = 1;
    if (r) {
        void* p = PyLong_AsVoidPtr(r);
        if (!p && PyErr_Occurred()) take_py_error();
        else { if (value_ptr) *value_ptr = p; rc = 0; }
        Py_DECREF(r);
    }
    PyGILState_Release(gil);
    return rc;
}

RBS_EXPORT int get_value_ptr_double(const char* name, void** value_ptr) { return get_value_ptr(name, value_ptr); }

RBS_EXPORT int get_var_shape(const char* name, int* shape_ptr) {
    PyGILState_STATE gil;
    if (enter(&gil)) return 1;
    PyObject* r = call("get_var_shape", "(s)", name);
    int rc = 1;
    if (r) {
        Py_ssize_t nd = PyTuple_Check(r) ? PyTuple_GET_SIZE(r) : -1;
        if (nd < 0) set_error("get_var_shape: not a tuple");
        else {
            rc = 0;
            for (Py_ssize_t k = 0; k < nd; k++) {
                long v = PyLong_AsLong(PyTuple_GET_ITEM(r, k));
                if (v == -1 && PyErr_Occurred()) { take_py_error(); rc = 1; break; }
                if (shape_ptr) shape_ptr[k] = (int)v;
            }
        }
        Py_DECREF(r);
    }
    PyGILState_Release(gil);
    return rc;
}

RBS_EXPORT int get_component_name(char* name) { write_cstring(name, "Ribasim"); return 0; }

RBS_EXPORT int get_version(char* version) { return call_string_out("get_version", NULL, version); }

RBS_EXPORT int get_last_bmi_error(char* error_message) { write_cstring(error_message, g_last_error); return 0; }

RBS_EXPORT int execute(const char* toml_path) { return call_int("execute", "s", toml_path, 0.0); }

/* build/set_julia_threads.c:12-34: thread count of the host runtime, settable only before the
 * runtime starts (1 = already initialised, 2 = out of range).  The device path has no host
 * thread pool; the value is kept for the embedded interpreter's numerical libraries. */
RBS_EXPORT int ribasim_set_julia_threads(int32_t threads) {
    if (g_capi || g_own_interpreter) return 1;
    if (threads < 1 || threads > INT16_MAX) return 2;
    char buf[16];
    snprintf(buf, sizeof buf, "%d", (int)threads);
    setenv("OMP_NUM_THREADS", buf, 1);
    return 0;
}
