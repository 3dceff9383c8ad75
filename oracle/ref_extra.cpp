// oracle/ref_extra.cpp -- TEST INFRASTRUCTURE ONLY.
// Extra C entry points on top of the reference's own exports.cpp (src/exports.cpp:42-311, compiled in place
// into oracle/_ref/liboracle_ref.so) for Model knobs that the reference's FFI does not expose but its apps
// set directly on the public members (src/onnxstream.h:944-968): attention slicing parts, range data,
// m_requires_upcast (src/llm.cpp:385-389), int64 tensor read-back.
#include "onnxstream.h"
#include <cstring>
#include <cstdlib>

using namespace onnxstream;

// src/exports.cpp:28-40 -- ModelContext's first member is the Model.
static Model& M(void* ctx) { return *reinterpret_cast<Model*>(ctx); }

static char* dup_err(const std::exception& e) { char* b = (char*)malloc(strlen(e.what()) + 1); strcpy(b, e.what()); return b; }

extern "C" {

void model_ext_set_attention_parts(void* ctx, unsigned parts) { M(ctx).m_attention_fused_ops_parts = parts; }

void model_ext_set_range(void* ctx, const char* op_name, float mn, float mx) { M(ctx).m_range_data[op_name] = { mn, mx }; }

char* model_ext_read_range_data(void* ctx, const char* fn)
{
    try { M(ctx).read_range_data(fn); return nullptr; } catch (const std::exception& e) { return dup_err(e); }
}

void oracle_set_bool(void* ctx, const char* name, int v)
{
    if (!strcmp(name, "range_data_calibrate")) M(ctx).m_range_data_calibrate = v != 0;
}

// m_requires_upcast(op_type, op_name) := op_name contains `pattern` (what src/llm.cpp:385-389 installs).
void model_ext_add_upcast_pattern(void* ctx, const char* pattern)
{
    std::string pat(pattern);
    auto prev = M(ctx).m_requires_upcast;
    M(ctx).m_requires_upcast = [pat, prev](const std::string& type, const std::string& name) {
        if (name.find(pat) != std::string::npos) return true;
        return prev ? prev(type, name) : false;
    };
}

void oracle_add_outputs_convert(void* ctx, const char* name) { M(ctx).m_outputs_convert_set.insert(name); }

void oracle_add_force_uint8_storage(void* ctx, const char* name) { M(ctx).m_force_uint8_storage_set.insert(name); }

// returns element count, -1 if absent / not int64; copies up to `cap` elements and the shape.
long long model_ext_get_tensor_i64(void* ctx, const char* name, long long* dst, long long cap, size_t* dims, size_t* ndims)
{
    for (auto& t : M(ctx).m_data)
        if (t.m_name == name) {
            if (t.m_type != TensorDataType::int64) return -1;
            auto& v = t.get_vector<int64_t>();
            for (size_t i = 0; i < v.size() && (long long)i < cap; i++) dst[i] = v[i];
            *ndims = t.m_shape.size();
            for (size_t i = 0; i < t.m_shape.size() && i < 8; i++) dims[i] = t.m_shape[i];
            return (long long)v.size();
        }
    return -1;
}

// Push a float32 / int64 input through Model::push_tensor without reading it back.  The reference's own model_add_tensor
// (src/exports.cpp:169-203) cannot be used once m_use_fp16_arithmetic is set: push_tensor converts the tensor to fp16
// (src/onnxstream.cpp:3029-3034) and the subsequent get_vector<float>() throws.  sd.cpp pushes exactly like this
// (src/sd.cpp:1488-1516).
void model_ext_push_tensor(void* ctx, const char* type, const char* name, unsigned dims_num, const unsigned* dims, const void* data)
{
    Tensor t;
    t.m_name = name;
    size_t n = 1;
    for (unsigned i = 0; i < dims_num; i++) { t.m_shape.push_back(dims[i]); n *= dims[i]; }
    if (!strcmp(type, "float32")) {
        tensor_vector<float> v(n);
        memcpy(v.data(), data, n * sizeof(float));
        t.set_vector(std::move(v));
    } else {
        tensor_vector<int64_t> v(n);
        memcpy(v.data(), data, n * sizeof(int64_t));
        t.set_vector(std::move(v));
    }
    M(ctx).push_tensor(std::move(t));
}

// type of a tensor left in m_data: 0 none, 1 u8, 2 f16, 3 f32, 4 i64 (TensorDataType order, src/onnxstream.h:147-154)
int model_ext_get_tensor_type(void* ctx, const char* name)
{
    for (auto& t : M(ctx).m_data) if (t.m_name == name) return (int)t.m_type;
    return -1;
}

}
