// TEST INFRASTRUCTURE ONLY — a minimal in-process implementation of the Node-API calls that
// bindings/node/planet_b200_addon.cc makes (the declarations of bindings/node/stub/node_api.h), so that the addon can be
// EXECUTED in an image without Node: registration, argument unpacking, typed-array validation, result objects, exceptions.
// Values live in an arena owned by the one environment; typed arrays either own their bytes or view caller memory (numpy),
// which gives the in-place semantics the reference's stage functions rely on.  Driven from Python through the nh_* entries
// (tests/napi_host/host.py).  It is not a JavaScript engine: no GC, no prototypes, no property attributes.
#include <node_api.h>

#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <utility>
#include <vector>

struct napi_value__ {
    napi_valuetype type = napi_undefined;
    double number = 0;
    bool boolean = false;
    std::string string;
    std::vector<std::pair<std::string, napi_value>> props;     // insertion order, like JS own string keys
    napi_callback fn = nullptr;
    // array buffers and typed arrays are objects with these set
    bool isBuffer = false, isTyped = false;
    std::shared_ptr<std::vector<uint8_t>> bytes;               // owned storage (null for external views)
    uint8_t* data = nullptr;
    size_t byteLength = 0, length = 0, byteOffset = 0;
    napi_typedarray_type elem = napi_uint8_array;
    napi_value buffer = nullptr;
};
struct napi_callback_info__ {
    std::vector<napi_value> args;
    napi_value thisArg = nullptr;
};
struct napi_env__ {
    std::vector<std::unique_ptr<napi_value__>> arena;
    size_t permanent = 0;                                       // values below this index survive nh_release
    napi_value undefinedValue = nullptr, nullValue = nullptr, exports = nullptr;
    bool pending = false, pendingIsTypeError = false;
    std::string pendingMessage;
    napi_value make(napi_valuetype t) {
        arena.emplace_back(new napi_value__());
        arena.back()->type = t;
        return arena.back().get();
    }
};

namespace {
napi_env__ g_env;
size_t elem_size(napi_typedarray_type t) {
    switch (t) {
        case napi_int8_array: case napi_uint8_array: case napi_uint8_clamped_array: return 1;
        case napi_int16_array: case napi_uint16_array: return 2;
        case napi_int32_array: case napi_uint32_array: case napi_float32_array: return 4;
        default: return 8;
    }
}
napi_value find(napi_value o, const char* name) {
    for (auto& p : o->props) if (p.first == name) return p.second;
    return nullptr;
}
}  // namespace

extern "C" {

// ---- the Node-API subset -------------------------------------------------------------------------------------------------
napi_status napi_get_cb_info(napi_env, napi_callback_info info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data) {
    if (argc) {
        const size_t cap = *argc;
        if (argv) for (size_t i = 0; i < cap; i++) argv[i] = i < info->args.size() ? info->args[i] : g_env.undefinedValue;
        *argc = info->args.size();
    }
    if (this_arg) *this_arg = info->thisArg;
    if (data) *data = nullptr;
    return napi_ok;
}
napi_status napi_typeof(napi_env, napi_value v, napi_valuetype* result) { *result = v->type; return napi_ok; }
napi_status napi_get_value_double(napi_env, napi_value v, double* result) {
    if (v->type != napi_number) return napi_number_expected;
    *result = v->number; return napi_ok;
}
napi_status napi_get_value_int32(napi_env, napi_value v, int32_t* result) {       // ECMAScript ToInt32 of a number value
    if (v->type != napi_number) return napi_number_expected;
    const double d = v->number;
    if (d != d || std::isinf(d)) { *result = 0; return napi_ok; }
    double m = std::fmod(std::trunc(d), 4294967296.0);
    if (m < 0) m += 4294967296.0;
    *result = (int32_t)(uint32_t)m; return napi_ok;
}
napi_status napi_get_value_bool(napi_env, napi_value v, bool* result) {
    if (v->type != napi_boolean) return napi_boolean_expected;
    *result = v->boolean; return napi_ok;
}
napi_status napi_get_value_string_utf8(napi_env, napi_value v, char* buf, size_t bufsize, size_t* result) {
    if (v->type != napi_string) return napi_string_expected;
    if (!buf) { if (result) *result = v->string.size(); return napi_ok; }
    const size_t n = bufsize ? std::min(bufsize - 1, v->string.size()) : 0;
    if (bufsize) { memcpy(buf, v->string.data(), n); buf[n] = 0; }
    if (result) *result = n;
    return napi_ok;
}
napi_status napi_get_named_property(napi_env, napi_value o, const char* name, napi_value* result) {
    if (o->type != napi_object && o->type != napi_function) return napi_object_expected;
    napi_value v = find(o, name);
    *result = v ? v : g_env.undefinedValue; return napi_ok;
}
napi_status napi_has_named_property(napi_env, napi_value o, const char* name, bool* result) {
    if (o->type != napi_object && o->type != napi_function) return napi_object_expected;
    *result = find(o, name) != nullptr; return napi_ok;
}
napi_status napi_set_named_property(napi_env, napi_value o, const char* name, napi_value value) {
    if (o->type != napi_object && o->type != napi_function) return napi_object_expected;
    for (auto& p : o->props) if (p.first == name) { p.second = value; return napi_ok; }
    o->props.emplace_back(name, value); return napi_ok;
}
napi_status napi_get_typedarray_info(napi_env, napi_value v, napi_typedarray_type* type, size_t* length, void** data,
                                     napi_value* arraybuffer, size_t* byte_offset) {
    if (v->type != napi_object || !v->isTyped) return napi_invalid_arg;
    if (type) *type = v->elem;
    if (length) *length = v->length;
    if (data) *data = v->data;
    if (arraybuffer) *arraybuffer = v->buffer;
    if (byte_offset) *byte_offset = v->byteOffset;
    return napi_ok;
}
napi_status napi_create_arraybuffer(napi_env, size_t byte_length, void** data, napi_value* result) {
    napi_value b = g_env.make(napi_object);
    b->isBuffer = true;
    b->bytes = std::make_shared<std::vector<uint8_t>>(byte_length ? byte_length : 1, (uint8_t)0);
    b->data = b->bytes->data(); b->byteLength = byte_length;
    if (data) *data = b->data;
    *result = b; return napi_ok;
}
napi_status napi_create_typedarray(napi_env, napi_typedarray_type type, size_t length, napi_value arraybuffer, size_t byte_offset,
                                   napi_value* result) {
    if (arraybuffer->type != napi_object || !arraybuffer->isBuffer) return napi_invalid_arg;
    const size_t es = elem_size(type);
    if (byte_offset % es || byte_offset + length * es > arraybuffer->byteLength) return napi_invalid_arg;   // RangeError in Node
    napi_value t = g_env.make(napi_object);
    t->isTyped = true; t->elem = type; t->length = length; t->byteOffset = byte_offset; t->buffer = arraybuffer;
    t->bytes = arraybuffer->bytes; t->data = arraybuffer->data + byte_offset; t->byteLength = length * es;
    *result = t; return napi_ok;
}
napi_status napi_create_object(napi_env, napi_value* result) { *result = g_env.make(napi_object); return napi_ok; }
napi_status napi_create_double(napi_env, double value, napi_value* result) {
    napi_value v = g_env.make(napi_number); v->number = value; *result = v; return napi_ok;
}
napi_status napi_get_undefined(napi_env, napi_value* result) { *result = g_env.undefinedValue; return napi_ok; }
napi_status napi_throw_error(napi_env, const char*, const char* msg) {
    if (g_env.pending) return napi_pending_exception;
    g_env.pending = true; g_env.pendingIsTypeError = false; g_env.pendingMessage = msg ? msg : ""; return napi_ok;
}
napi_status napi_throw_type_error(napi_env, const char*, const char* msg) {
    if (g_env.pending) return napi_pending_exception;
    g_env.pending = true; g_env.pendingIsTypeError = true; g_env.pendingMessage = msg ? msg : ""; return napi_ok;
}
napi_status napi_define_properties(napi_env env, napi_value object, size_t n, const napi_property_descriptor* props) {
    for (size_t i = 0; i < n; i++) {
        napi_value v = props[i].value;
        if (props[i].method) { v = g_env.make(napi_function); v->fn = props[i].method; }
        if (!v || !props[i].utf8name) return napi_invalid_arg;
        napi_set_named_property(env, object, props[i].utf8name, v);
    }
    return napi_ok;
}

napi_value napi_register_module_v1(napi_env env, napi_value exports);      // the addon's NAPI_MODULE_INIT

// ---- driver (what `require()` and a JS caller would do) ----------------------------------------------------------------------
int nh_init(void) {
    if (g_env.exports) return 0;
    g_env.undefinedValue = g_env.make(napi_undefined);
    g_env.nullValue = g_env.make(napi_null);
    napi_value ex = g_env.make(napi_object);
    napi_value r = napi_register_module_v1(&g_env, ex);
    g_env.exports = r ? r : ex;
    g_env.permanent = g_env.arena.size();
    return g_env.pending ? -1 : 0;
}
// drops every value created since nh_init (handles held by the caller become invalid)
void nh_release(void) { g_env.arena.resize(g_env.permanent); }
napi_value nh_undefined(void) { return g_env.undefinedValue; }
napi_value nh_null(void) { return g_env.nullValue; }
napi_value nh_number(double d) { napi_value v; napi_create_double(&g_env, d, &v); return v; }
napi_value nh_bool(int b) { napi_value v = g_env.make(napi_boolean); v->boolean = b != 0; return v; }
napi_value nh_string(const char* s) { napi_value v = g_env.make(napi_string); v->string = s; return v; }
napi_value nh_object(void) { return g_env.make(napi_object); }
// typed array over caller memory (no copy): writes by the addon are visible to the caller, like a JS typed array
napi_value nh_typed_view(int type, size_t length, void* data) {
    napi_value b = g_env.make(napi_object);
    b->isBuffer = true; b->data = static_cast<uint8_t*>(data); b->byteLength = length * elem_size((napi_typedarray_type)type);
    napi_value t = g_env.make(napi_object);
    t->isTyped = true; t->elem = (napi_typedarray_type)type; t->length = length; t->buffer = b; t->data = b->data; t->byteLength = b->byteLength;
    return t;
}
int nh_set(napi_value o, const char* name, napi_value v) { return napi_set_named_property(&g_env, o, name, v); }
napi_value nh_get(napi_value o, const char* name) { return (o->type == napi_object || o->type == napi_function) ? find(o, name) : nullptr; }
int nh_typeof(napi_value v) { return v->type; }
double nh_number_value(napi_value v) { return v->number; }
int nh_bool_value(napi_value v) { return v->boolean; }
const char* nh_string_value(napi_value v) { return v->string.c_str(); }
int nh_is_typed(napi_value v) { return v->type == napi_object && v->isTyped; }
int nh_typed_info(napi_value v, int* type, size_t* length, void** data) {
    if (!nh_is_typed(v)) return -1;
    *type = v->elem; *length = v->length; *data = v->data; return 0;
}
size_t nh_num_keys(napi_value o) { return o->props.size(); }
const char* nh_key(napi_value o, size_t i) { return o->props[i].first.c_str(); }
size_t nh_num_exports(void) { return g_env.exports->props.size(); }
const char* nh_export_name(size_t i) { return g_env.exports->props[i].first.c_str(); }
// exports[name](...argv): the result, or NULL with an exception pending (also when the export does not exist)
napi_value nh_call(const char* name, int argc, napi_value* argv) {
    napi_value f = find(g_env.exports, name);
    if (!f || f->type != napi_function) { napi_throw_type_error(&g_env, nullptr, "not a function"); return nullptr; }
    napi_callback_info__ info;
    info.args.assign(argv, argv + argc);
    info.thisArg = g_env.exports;
    napi_value r = f->fn(&g_env, &info);
    if (g_env.pending) return nullptr;
    return r ? r : g_env.undefinedValue;
}
// 0 = none, 1 = Error, 2 = TypeError; the message is copied out and the exception cleared
int nh_take_exception(char* buf, size_t cap) {
    if (!g_env.pending) return 0;
    if (cap) { strncpy(buf, g_env.pendingMessage.c_str(), cap - 1); buf[cap - 1] = 0; }
    g_env.pending = false;
    return g_env.pendingIsTypeError ? 2 : 1;
}

}  // extern "C"
