/* Minimal stand-in for <node_api.h>, for SYNTAX-CHECKING napi/openclaw_gov_napi.c only (Node.js is absent in this image). */
#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>
typedef struct napi_env__ *napi_env; typedef struct napi_value__ *napi_value; typedef struct napi_callback_info__ *napi_callback_info;
typedef enum { napi_ok } napi_status; typedef enum { napi_default } napi_property_attributes;
typedef enum { napi_int8_array, napi_uint8_array, napi_int32_array, napi_uint32_array, napi_biguint64_array } napi_typedarray_type;
typedef enum { napi_undefined, napi_number, napi_string, napi_object } napi_valuetype;
typedef napi_value (*napi_callback)(napi_env, napi_callback_info);
typedef void (*napi_finalize)(napi_env, void *, void *);
typedef struct { const char *utf8name; napi_value name; napi_callback method, getter, setter; napi_value value; napi_property_attributes attributes; void *data; } napi_property_descriptor;
napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t *, napi_value *, napi_value *, void **);
napi_status napi_throw_error(napi_env, const char *, const char *); napi_status napi_throw_type_error(napi_env, const char *, const char *);
napi_status napi_get_value_external(napi_env, napi_value, void **); napi_status napi_create_external(napi_env, void *, napi_finalize, void *, napi_value *);
napi_status napi_create_arraybuffer(napi_env, size_t, void **, napi_value *); napi_status napi_create_typedarray(napi_env, napi_typedarray_type, size_t, napi_value, size_t, napi_value *);
napi_status napi_get_typedarray_info(napi_env, napi_value, napi_typedarray_type *, size_t *, void **, napi_value *, size_t *);
napi_status napi_create_object(napi_env, napi_value *); napi_status napi_set_named_property(napi_env, napi_value, const char *, napi_value);
napi_status napi_get_named_property(napi_env, napi_value, const char *, napi_value *); napi_status napi_has_named_property(napi_env, napi_value, const char *, bool *);
napi_status napi_get_array_length(napi_env, napi_value, uint32_t *); napi_status napi_get_element(napi_env, napi_value, uint32_t, napi_value *);
napi_status napi_get_value_string_utf8(napi_env, napi_value, char *, size_t, size_t *); napi_status napi_get_value_uint32(napi_env, napi_value, uint32_t *); napi_status napi_get_value_int32(napi_env, napi_value, int32_t *);
napi_status napi_typeof(napi_env, napi_value, napi_valuetype *); napi_status napi_get_undefined(napi_env, napi_value *);
napi_status napi_define_properties(napi_env, napi_value, size_t, const napi_property_descriptor *);
#define NAPI_MODULE(name, fn)
#define NODE_GYP_MODULE_NAME x
