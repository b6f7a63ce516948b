/* napi/stub/node_api.h -- DECLARATIONS ONLY, for a syntax / type check of openclaw_gov_napi.c where Node.js is absent
 * (tests/test_oracle_golden.py::test_napi_shim_compiles).  Signatures as in Node's node_api.h / js_native_api.h (N-API
 * version 8); never linked, never shipped -- a real build uses Node's own header. */
#ifndef CG_STUB_NODE_API_H
#define CG_STUB_NODE_API_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_callback_info__ *napi_callback_info;
typedef enum { napi_ok, napi_invalid_arg, napi_generic_failure } napi_status;
typedef enum { napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object, napi_function, napi_external, napi_bigint } napi_valuetype;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array, napi_int32_array, napi_uint32_array,
               napi_float32_array, napi_float64_array, napi_bigint64_array, napi_biguint64_array } napi_typedarray_type;
typedef enum { napi_default = 0 } napi_property_attributes;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void *finalize_data, void *finalize_hint);
typedef struct { const char *utf8name; napi_value name; napi_callback method; napi_callback getter; napi_callback setter; napi_value value;
                 napi_property_attributes attributes; void *data; } napi_property_descriptor;
napi_status napi_throw_error(napi_env env, const char *code, const char *msg);
napi_status napi_throw_type_error(napi_env env, const char *code, const char *msg);
napi_status napi_throw_range_error(napi_env env, const char *code, const char *msg);
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t *argc, napi_value *argv, napi_value *this_arg, void **data);
napi_status napi_typeof(napi_env env, napi_value value, napi_valuetype *result);
napi_status napi_get_value_int32(napi_env env, napi_value value, int32_t *result);
napi_status napi_get_value_uint32(napi_env env, napi_value value, uint32_t *result);
napi_status napi_get_value_double(napi_env env, napi_value value, double *result);
napi_status napi_get_value_bool(napi_env env, napi_value value, bool *result);
napi_status napi_get_value_bigint_uint64(napi_env env, napi_value value, uint64_t *result, bool *lossless);
napi_status napi_get_value_string_utf8(napi_env env, napi_value value, char *buf, size_t bufsize, size_t *result);
napi_status napi_get_value_external(napi_env env, napi_value value, void **result);
napi_status napi_get_array_length(napi_env env, napi_value value, uint32_t *result);
napi_status napi_get_element(napi_env env, napi_value object, uint32_t index, napi_value *result);
napi_status napi_get_named_property(napi_env env, napi_value object, const char *utf8name, napi_value *result);
napi_status napi_has_named_property(napi_env env, napi_value object, const char *utf8name, bool *result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char *utf8name, napi_value value);
napi_status napi_is_typedarray(napi_env env, napi_value value, bool *result);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type *type, size_t *length, void **data, napi_value *arraybuffer, size_t *byte_offset);
napi_status napi_create_arraybuffer(napi_env env, size_t byte_length, void **data, napi_value *result);
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value arraybuffer, size_t byte_offset, napi_value *result);
napi_status napi_create_object(napi_env env, napi_value *result);
napi_status napi_create_external(napi_env env, void *data, napi_finalize finalize_cb, void *finalize_hint, napi_value *result);
napi_status napi_create_double(napi_env env, double value, napi_value *result);
napi_status napi_create_int32(napi_env env, int32_t value, napi_value *result);
napi_status napi_create_bigint_uint64(napi_env env, uint64_t value, napi_value *result);
napi_status napi_get_boolean(napi_env env, bool value, napi_value *result);
napi_status napi_define_properties(napi_env env, napi_value object, size_t property_count, const napi_property_descriptor *properties);
#define NAPI_MODULE(modname, regfunc) napi_value napi_register_module_v1(napi_env env, napi_value exports) { return regfunc(env, exports); }
#endif
