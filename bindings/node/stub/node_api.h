/* Minimal stand-in for Node's <node_api.h>: only the declarations bindings/node/planet_b200_addon.cc uses,
 * with the signatures of Node-API version 8.  This image has no Node toolchain; the stub lets
 * `g++ -fsyntax-only` type-check the addon (tests/test_abi.py).  Build against the real header with node-gyp. */
#ifndef STUB_NODE_API_H
#define STUB_NODE_API_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef enum { napi_ok = 0, napi_invalid_arg, napi_object_expected, napi_string_expected, napi_name_expected, napi_function_expected,
               napi_number_expected, napi_boolean_expected, napi_array_expected, napi_generic_failure, napi_pending_exception } napi_status;
typedef enum { napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object, napi_function,
               napi_external, napi_bigint } napi_valuetype;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array,
               napi_int32_array, napi_uint32_array, napi_float32_array, napi_float64_array } napi_typedarray_type;
typedef enum { napi_default = 0 } napi_property_attributes;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef struct {
    const char* utf8name; napi_value name; napi_callback method; napi_callback getter; napi_callback setter; napi_value value;
    napi_property_attributes attributes; void* data;
} napi_property_descriptor;
napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_typeof(napi_env, napi_value, napi_valuetype* result);
napi_status napi_get_value_double(napi_env, napi_value, double* result);
napi_status napi_get_value_int32(napi_env, napi_value, int32_t* result);
napi_status napi_get_value_bool(napi_env, napi_value, bool* result);
napi_status napi_get_value_string_utf8(napi_env, napi_value, char* buf, size_t bufsize, size_t* result);
napi_status napi_get_named_property(napi_env, napi_value object, const char* utf8name, napi_value* result);
napi_status napi_has_named_property(napi_env, napi_value object, const char* utf8name, bool* result);
napi_status napi_set_named_property(napi_env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_get_typedarray_info(napi_env, napi_value, napi_typedarray_type* type, size_t* length, void** data,
                                     napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_create_arraybuffer(napi_env, size_t byte_length, void** data, napi_value* result);
napi_status napi_create_typedarray(napi_env, napi_typedarray_type type, size_t length, napi_value arraybuffer, size_t byte_offset,
                                   napi_value* result);
napi_status napi_create_object(napi_env, napi_value* result);
napi_status napi_create_double(napi_env, double value, napi_value* result);
napi_status napi_get_undefined(napi_env, napi_value* result);
napi_status napi_throw_error(napi_env, const char* code, const char* msg);
napi_status napi_throw_type_error(napi_env, const char* code, const char* msg);
napi_status napi_define_properties(napi_env, napi_value object, size_t property_count, const napi_property_descriptor* properties);
#define NAPI_MODULE_INIT() extern "C" napi_value napi_register_module_v1(napi_env env, napi_value exports)
#ifdef __cplusplus
}
#endif
#endif
