/* capi_host.c — a plain C host (no Python in the process) that binds libribasim.so the way the
 * reference's Rust CLI does (build/cli/src/main.rs:59-87: dlopen + symbol lookup) and walks
 * through the BMI calls of python/ribasim_api/tests/test_bmi.py.
 *   capi_host <libribasim.so>              : calls that need no model (CPU-only check)
 *   capi_host <libribasim.so> <model.toml> : initialize / update / get_value_ptr / finalize
 * Prints one "key value" line per check; exit code 0 when every call behaved. */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

typedef int (*fn_s)(const char*);
typedef int (*fn_v)(void);
typedef int (*fn_d)(double);
typedef int (*fn_pd)(double*);
typedef int (*fn_buf)(char*);
typedef int (*fn_name_buf)(const char*, char*);
typedef int (*fn_name_pi)(const char*, int*);
typedef int (*fn_name_pp)(const char*, void**);

#define SYM(type, name) type name = (type)dlsym(lib, #name); if (!name) { printf("missing %s\n", #name); return 2; }

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    void* lib = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { printf("dlopen failed: %s\n", dlerror()); return 2; }
    SYM(fn_s, initialize) SYM(fn_v, finalize) SYM(fn_v, update) SYM(fn_d, update_until)
    SYM(fn_v, update_subgrid_level) SYM(fn_pd, get_current_time) SYM(fn_pd, get_start_time)
    SYM(fn_pd, get_end_time) SYM(fn_pd, get_time_step) SYM(fn_name_buf, get_var_type)
    SYM(fn_name_pi, get_var_rank) SYM(fn_name_pp, get_value_ptr) SYM(fn_name_pp, get_value_ptr_double)
    SYM(fn_name_pi, get_var_shape) SYM(fn_buf, get_component_name) SYM(fn_buf, get_version)
    SYM(fn_buf, get_last_bmi_error) SYM(fn_s, execute)
    int (*ribasim_set_julia_threads)(int) = (int (*)(int))dlsym(lib, "ribasim_set_julia_threads");
    if (!ribasim_set_julia_threads) { printf("missing ribasim_set_julia_threads\n"); return 2; }
    char buf[1025];
    int bad = 0;
    printf("set_threads_range %d\n", ribasim_set_julia_threads(0));        /* 2: out of range */
    printf("set_threads %d\n", ribasim_set_julia_threads(4));              /* 0 */
    bad |= get_component_name(buf); printf("component %s\n", buf);
    bad |= get_version(buf); printf("version %s\n", buf);
    printf("set_threads_late %d\n", ribasim_set_julia_threads(4));         /* 1: runtime is up */
    int rc = update(); get_last_bmi_error(buf); printf("update_uninitialized %d %s\n", rc, buf);
    rc = finalize(); printf("finalize_uninitialized %d\n", rc);
    if (argc < 3) return bad;
    rc = initialize(argv[2]);
    if (rc) { get_last_bmi_error(buf); printf("initialize %d %s\n", rc, buf); return 1; }
    printf("initialize 0\n");
    double t0, t1, te, dt;
    bad |= get_start_time(&t0) | get_current_time(&t1) | get_end_time(&te) | get_time_step(&dt);
    printf("times %.17g %.17g %.17g %.17g\n", t0, t1, te, dt);
    bad |= get_var_type("basin.storage", buf); printf("var_type %s\n", buf);
    int rank = -1, shape[4] = {0};
    bad |= get_var_rank("basin.storage", &rank) | get_var_shape("basin.storage", shape);
    printf("rank_shape %d %d\n", rank, shape[0]);
    double *storage = NULL, *level = NULL, *storage2 = NULL, *infil = NULL;
    bad |= get_value_ptr("basin.storage", (void**)&storage) | get_value_ptr("basin.level", (void**)&level);
    printf("storage0"); for (int k = 0; k < shape[0]; k++) printf(" %.17g", storage[k]); printf("\n");
    printf("level0"); for (int k = 0; k < shape[0]; k++) printf(" %.17g", level[k]); printf("\n");
    bad |= update(); bad |= get_current_time(&t1); printf("after_update %.17g\n", t1);
    bad |= update_until(86400.0); bad |= get_current_time(&t1); printf("after_update_until %.17g\n", t1);
    bad |= get_value_ptr_double("basin.storage", (void**)&storage2);
    printf("pointer_stable %d\n", storage == storage2);
    printf("storage1"); for (int k = 0; k < shape[0]; k++) printf(" %.17g", storage[k]); printf("\n");
    bad |= get_value_ptr("basin.infiltration", (void**)&infil);
    infil[0] = 1e-6;                                   /* writable forcing (core/src/bmi.jl:64-88) */
    bad |= update();
    rc = get_var_type("unknown_node.unknown_variable", buf); get_last_bmi_error(buf);
    printf("unknown_var %d %s\n", rc, buf);
    rc = update_until(10.0); get_last_bmi_error(buf); printf("backwards %d %s\n", rc, buf);
    bad |= update_subgrid_level();
    bad |= finalize();
    rc = update(); get_last_bmi_error(buf); printf("update_finalized %d %s\n", rc, buf);
    printf("bad %d\n", bad);
    return bad;
}
