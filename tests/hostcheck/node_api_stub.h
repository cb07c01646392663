/* Declarations of the N-API (node_api.h, Node >= 10) entry points integration/nzcb_napi.c uses, as recalled --
 * ONLY so that the shim can be type-checked with `gcc -fsyntax-only` in an image without Node.  Test infrastructure. */
#ifndef NODE_API_STUB_H
#define NODE_API_STUB_H
#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>
typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef enum { napi_ok, napi_invalid_arg, napi_generic_failure } napi_status;
typedef enum { napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object, napi_function,
               napi_external, napi_bigint } napi_valuetype;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);
#define NAPI_AUTO_LENGTH SIZE_MAX
napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_get_buffer_info(napi_env env, napi_value value, void** data, size_t* length);
napi_status napi_create_buffer(napi_env env, size_t length, void** data, napi_value* result);
napi_status napi_create_buffer_copy(napi_env env, size_t length, const void* data, void** result_data, napi_value* result);
napi_status napi_create_external(napi_env env, void* data, napi_finalize finalize_cb, void* finalize_hint, napi_value* result);
napi_status napi_get_value_external(napi_env env, napi_value value, void** result);
napi_status napi_get_value_uint32(napi_env env, napi_value value, uint32_t* result);
napi_status napi_get_value_string_utf8(napi_env env, napi_value value, char* buf, size_t bufsize, size_t* result);
napi_status napi_create_string_utf8(napi_env env, const char* str, size_t length, napi_value* result);
napi_status napi_create_object(napi_env env, napi_value* result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_create_function(napi_env env, const char* utf8name, size_t length, napi_callback cb, void* data, napi_value* result);
napi_status napi_get_boolean(napi_env env, bool value, napi_value* result);
napi_status napi_typeof(napi_env env, napi_value value, napi_valuetype* result);
napi_status napi_throw_error(napi_env env, const char* code, const char* msg);
#define NAPI_MODULE(modname, regfunc) napi_value nzcb_napi_register_stub(napi_env env, napi_value exports) { return regfunc(env, exports); }
#endif
