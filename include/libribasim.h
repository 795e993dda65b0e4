/* libribasim.h — C ABI of the `libribasim` stand-in (ribasim_b200/libribasim.so).
 *
 * These are the exports of the reference's shared library, /root/reference/build/libribasim.jl
 * :74-212 (18 `Base.@ccallable` functions) and build/set_julia_threads.c:12 — same names,
 * argument meaning and conventions, so that the existing consumers bind it unchanged: xmipy /
 * python/ribasim_api/ribasim_api/ribasim_api.py:5-26 (ctypes.CDLL) and the Rust CLI
 * (build/cli/src/main.rs:59-87, dlopen + `execute`).
 *
 * Conventions (libribasim.jl:40-72): every function returns 0 on success and 1 on failure; the
 * message of the last failure is written by get_last_bmi_error into a caller-owned buffer of
 * BMI_LENERRMESSAGE bytes.  Calls other than initialize / finalize / get_component_name /
 * get_version / get_last_bmi_error / execute fail with "Model not initialized" before initialize.
 * Strings are NUL-terminated ASCII in caller-owned buffers (sizes: ribasim_api.py:8-19).
 * get_value_ptr returns the address of the model's own Float64 array for the variable: stable
 * from initialize to finalize, readable at any time, writable for basin.infiltration,
 * basin.drainage, basin.surface_runoff and user_demand.demand (core/src/bmi.jl:59-89).
 */
#ifndef LIBRIBASIM_H
#define LIBRIBASIM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define BMI_LENVARTYPE 51
#define BMI_LENGRIDTYPE 17
#define BMI_LENVARADDRESS 68
#define BMI_LENCOMPONENTNAME 256
#define BMI_LENVERSION 256
#define BMI_LENERRMESSAGE 1025

int initialize(const char* config_path);                       /* libribasim.jl:74-80 */
int finalize(void);                                            /* :82-90 */
int update(void);                                              /* :92-97, one solver step */
int update_until(double time);                                 /* :99-104, seconds since start */
int update_subgrid_level(void);                                /* :106-111 */
int get_current_time(double* time);                            /* :113-119 */
int get_start_time(double* time);                              /* :121-127 */
int get_end_time(double* time);                                /* :129-135 */
int get_time_step(double* time_step);                          /* :137-143 */
int get_var_type(const char* name, char* var_type);            /* :145-152, "double" */
int get_var_rank(const char* name, int* rank);                 /* :154-160 */
int get_value_ptr(const char* name, void** value_ptr);         /* :162-169 */
int get_var_shape(const char* name, int* shape);               /* :171-179 */
int get_component_name(char* name);                            /* :181-186, "Ribasim" */
int get_version(char* version);                                /* :188-193 */
int get_last_bmi_error(char* error_message);                   /* :195-200 */
int execute(const char* toml_path);                            /* :202-204, Ribasim.main */
int get_value_ptr_double(const char* name, void** value_ptr);  /* :206-211 */
int ribasim_set_julia_threads(int32_t threads);                /* build/set_julia_threads.c:12 */

#ifdef __cplusplus
}
#endif
#endif
