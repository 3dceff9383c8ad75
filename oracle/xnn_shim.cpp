// oracle/xnn_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// The reference (vitoplantamura/OnnxStream) computes every Conv/MatMul/Softmax/elementwise op through
// google/XNNPACK (pinned 5671db05..., src/CMakeLists.txt:43-50), which is an un-vendored download.  The
// only XNNPACK available offline is the one bundled in libtorch_cpu.so; it exports 36 of the 66 entry
// points the reference's onnxstream.cpp needs.  This translation unit supplies the 30 missing ones so that
// the reference's own Model::run() can be linked and executed as the parity oracle (oracle/_ref):
//
//   * xnn_{create,reshape,setup}_softmax_nc_{f32,f16,qu8}            (called at src/onnxstream.cpp:1958-2051)
//   * xnn_{create,reshape,setup}_dynamic_fully_connected_nc_{f32,f16} (src/onnxstream.cpp:929-1023)
//   * xnn_{create,reshape,setup}_scaled_dot_product_attention_nhtc_{f32,f16} (src/onnxstream.cpp:2053-2149)
//   * xnn_run_transpose_nd_x{8,16,32}                                (src/onnxstream.cpp:1748-1809)
//   * xnn_*_convolution2d_nchw_{f32,f16}: "unsupported" stubs (path unusable with file weights,
//     src/onnxstream.cpp:2686-2689 throws).
//
// These are restatements of XNNPACK's *documented* semantics in scalar fp32 C++ (f16 variants:
// convert -> fp32 compute -> convert), i.e. for these ops parity is pinned by this restatement, not by the
// XNNPACK binaries the reference author used ("parity unpinned by the reference" for these ops).
//
// It additionally interposes xnn_{create,reshape,setup}_{convolution2d_nhwc,fully_connected_nc}_f16 because
// the f16 Conv/FC micro-kernels of torch's bundled XNNPACK return wrong values on AVX512-FP16 hosts
// (SURVEY.md section 8c): they are recomputed as fp16 storage / fp32 XNNPACK arithmetic / fp16 rounding of
// the result.  Set OSB200_ORACLE_NATIVE_F16=1 to disable the interposition (used by the self-test that
// documents the breakage).
//
// Operators created here are tagged objects smuggled through xnn_operator_t; xnn_run_operator and
// xnn_delete_operator are interposed (link with -Wl,-Bsymbolic) and forward everything else to the real
// library.

#include <xnnpack.h>
#include <dlfcn.h>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <mutex>
#include <unordered_set>
#include <vector>

namespace {

typedef _Float16 f16_t;

inline float h2f(uint16_t h) { f16_t v; std::memcpy(&v, &h, 2); return (float)v; }
inline uint16_t f2h(float f) { f16_t v = (f16_t)f; uint16_t h; std::memcpy(&h, &v, 2); return h; }

void* real_sym(const char* name)
{
    static void* handle = nullptr;
    if (!handle) {
        handle = dlopen("libtorch_cpu.so", RTLD_LAZY | RTLD_NOLOAD);
        if (!handle) handle = dlopen("libtorch_cpu.so", RTLD_LAZY);
    }
    void* p = handle ? dlsym(handle, name) : nullptr;
    if (!p) { fprintf(stderr, "xnn_shim: cannot resolve %s in libtorch_cpu.so\n", name); abort(); }
    return p;
}

#define REAL(name) ((decltype(&name))real_sym(#name))

enum class Kind { softmax_f32, softmax_f16, softmax_qu8, dynfc_f32, dynfc_f16, sdpa_f32, sdpa_f16, conv_f16, fc_f16 };

struct ShimOp
{
    Kind kind;
    // softmax
    size_t channels = 0, batch = 0;
    float in_scale = 0, out_scale = 0; uint8_t out_zp = 0;
    // dynamic fc / fc
    size_t K = 0, N = 0; uint32_t flags = 0;
    // sdpa
    size_t B = 0, Hq = 0, Tq = 0, Hkv = 0, Tk = 0, Dqk = 0, Dv = 0;
    // pointers
    const void* in0 = nullptr; const void* in1 = nullptr; const void* in2 = nullptr;
    const void* in3 = nullptr; const void* in4 = nullptr; void* out = nullptr;
    // wrapped real f32 operator for conv_f16 / fc_f16
    xnn_operator_t inner = nullptr;
    size_t in_elems = 0, out_elems = 0, cin = 0, cout = 0;
    std::vector<float> tmp_in, tmp_out;
};

std::mutex g_mutex;
std::unordered_set<void*> g_ops;

xnn_operator_t register_op(ShimOp* op)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    g_ops.insert(op);
    return (xnn_operator_t)op;
}

ShimOp* as_shim(xnn_operator_t op)
{
    std::lock_guard<std::mutex> lk(g_mutex);
    return g_ops.count((void*)op) ? (ShimOp*)op : nullptr;
}

bool native_f16()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("OSB200_ORACLE_NATIVE_F16"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

template <typename LOAD, typename STORE>
void softmax_rows(size_t batch, size_t channels, LOAD load, STORE store)
{
    #pragma omp parallel for schedule(static)
    for (long long b = 0; b < (long long)batch; b++) {
        float mx = -std::numeric_limits<float>::infinity();
        for (size_t c = 0; c < channels; c++) mx = std::fmax(mx, load(b * channels + c));
        float sum = 0;
        std::vector<float> e(channels);
        for (size_t c = 0; c < channels; c++) { e[c] = std::exp(load(b * channels + c) - mx); sum += e[c]; }
        float inv = 1.0f / sum;
        for (size_t c = 0; c < channels; c++) store(b * channels + c, e[c] * inv);
    }
}

// out[m,n] = sum_k in[m,k] * w[k,n] (+ bias[n]); w is [K,N] when XNN_FLAG_TRANSPOSE_WEIGHTS is set, else [N,K].
template <typename LOADI, typename LOADW, typename LOADB, typename STORE>
void fc_rows(size_t M, size_t K, size_t N, bool w_is_kn, bool has_bias, LOADI li, LOADW lw, LOADB lb, STORE st)
{
    #pragma omp parallel for schedule(static)
    for (long long m = 0; m < (long long)M; m++) {
        std::vector<float> acc(N);
        for (size_t n = 0; n < N; n++) acc[n] = has_bias ? lb(n) : 0.0f;
        if (w_is_kn) {
            for (size_t k = 0; k < K; k++) {
                float a = li(m * K + k);
                for (size_t n = 0; n < N; n++) acc[n] += a * lw(k * N + n);
            }
        } else {
            for (size_t n = 0; n < N; n++) {
                float s = 0;
                for (size_t k = 0; k < K; k++) s += li(m * K + k) * lw(n * K + k);
                acc[n] += s;
            }
        }
        for (size_t n = 0; n < N; n++) st(m * N + n, acc[n]);
    }
}

// XNNPACK scaled_dot_product_attention_nhtc: Q [B,Hq,Tq,D], K [B,Hkv,Tk,D], V [B,Hkv,Tk,Dv],
// scale [D] applied to Q per channel, mask [Tq,Tk] added to the logits, softmax over Tk, times V.
template <typename LOAD, typename STORE>
void sdpa(const ShimOp& o, LOAD ld, STORE st)
{
    const size_t B = o.B, Hq = o.Hq, Tq = o.Tq, Hkv = o.Hkv, Tk = o.Tk, D = o.Dqk, Dv = o.Dv;
    const size_t group = Hq / Hkv;
    #pragma omp parallel for schedule(static)
    for (long long bh = 0; bh < (long long)(B * Hq); bh++) {
        size_t b = bh / Hq, h = bh % Hq, hk = (Hkv == 1) ? 0 : (Hkv == Hq ? h : h / group);
        std::vector<float> qs(D), logit(Tk), acc(Dv);
        for (size_t t = 0; t < Tq; t++) {
            size_t qoff = ((b * Hq + h) * Tq + t) * D;
            for (size_t c = 0; c < D; c++) qs[c] = ld(o.in0, qoff + c) * ld(o.in3, c);
            float mx = -std::numeric_limits<float>::infinity();
            for (size_t s = 0; s < Tk; s++) {
                size_t koff = ((b * Hkv + hk) * Tk + s) * D;
                float dot = 0;
                for (size_t c = 0; c < D; c++) dot += qs[c] * ld(o.in1, koff + c);
                dot += ld(o.in4, t * Tk + s);
                logit[s] = dot;
                mx = std::fmax(mx, dot);
            }
            float sum = 0;
            for (size_t s = 0; s < Tk; s++) { logit[s] = std::exp(logit[s] - mx); sum += logit[s]; }
            float inv = 1.0f / sum;
            for (size_t c = 0; c < Dv; c++) acc[c] = 0;
            for (size_t s = 0; s < Tk; s++) {
                size_t voff = ((b * Hkv + hk) * Tk + s) * Dv;
                float p = logit[s] * inv;
                for (size_t c = 0; c < Dv; c++) acc[c] += p * ld(o.in2, voff + c);
            }
            size_t ooff = ((b * Hq + h) * Tq + t) * Dv;
            for (size_t c = 0; c < Dv; c++) st(ooff + c, acc[c]);
        }
    }
}

xnn_status run_shim(ShimOp& o, pthreadpool_t tp)
{
    switch (o.kind) {
    case Kind::softmax_f32: {
        const float* in = (const float*)o.in0; float* out = (float*)o.out;
        softmax_rows(o.batch, o.channels, [&](size_t i) { return in[i]; }, [&](size_t i, float v) { out[i] = v; });
        return xnn_status_success;
    }
    case Kind::softmax_f16: {
        const uint16_t* in = (const uint16_t*)o.in0; uint16_t* out = (uint16_t*)o.out;
        softmax_rows(o.batch, o.channels, [&](size_t i) { return h2f(in[i]); }, [&](size_t i, float v) { out[i] = f2h(v); });
        return xnn_status_success;
    }
    case Kind::softmax_qu8: {
        // NOTE: plain restatement (dequantise -> softmax -> requantise); XNNPACK's LUT rounding is not reproduced.
        const uint8_t* in = (const uint8_t*)o.in0; uint8_t* out = (uint8_t*)o.out;
        softmax_rows(o.batch, o.channels, [&](size_t i) { return (float)in[i] * o.in_scale; },
            [&](size_t i, float v) {
                long q = std::lrintf(v / o.out_scale) + o.out_zp;
                out[i] = (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q));
            });
        return xnn_status_success;
    }
    case Kind::dynfc_f32: {
        const float* in = (const float*)o.in0; const float* w = (const float*)o.in1; const float* b = (const float*)o.in2; float* out = (float*)o.out;
        fc_rows(o.batch, o.K, o.N, (o.flags & XNN_FLAG_TRANSPOSE_WEIGHTS) != 0, b != nullptr,
            [&](size_t i) { return in[i]; }, [&](size_t i) { return w[i]; }, [&](size_t i) { return b[i]; },
            [&](size_t i, float v) { out[i] = v; });
        return xnn_status_success;
    }
    case Kind::dynfc_f16: {
        const uint16_t* in = (const uint16_t*)o.in0; const uint16_t* w = (const uint16_t*)o.in1; const uint16_t* b = (const uint16_t*)o.in2; uint16_t* out = (uint16_t*)o.out;
        // convert the weights once (they are re-read M times)
        std::vector<float> wf(o.K * o.N);
        for (size_t i = 0; i < wf.size(); i++) wf[i] = h2f(w[i]);
        fc_rows(o.batch, o.K, o.N, (o.flags & XNN_FLAG_TRANSPOSE_WEIGHTS) != 0, b != nullptr,
            [&](size_t i) { return h2f(in[i]); }, [&](size_t i) { return wf[i]; }, [&](size_t i) { return h2f(b[i]); },
            [&](size_t i, float v) { out[i] = f2h(v); });
        return xnn_status_success;
    }
    case Kind::sdpa_f32: {
        float* out = (float*)o.out;
        sdpa(o, [](const void* p, size_t i) { return ((const float*)p)[i]; }, [&](size_t i, float v) { out[i] = v; });
        return xnn_status_success;
    }
    case Kind::sdpa_f16: {
        uint16_t* out = (uint16_t*)o.out;
        sdpa(o, [](const void* p, size_t i) { return h2f(((const uint16_t*)p)[i]); }, [&](size_t i, float v) { out[i] = f2h(v); });
        return xnn_status_success;
    }
    case Kind::conv_f16:
    case Kind::fc_f16: {
        const uint16_t* in = (const uint16_t*)o.in0; uint16_t* out = (uint16_t*)o.out;
        o.tmp_in.resize(o.in_elems + 16); o.tmp_out.resize(o.out_elems + 16);
        #pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)o.in_elems; i++) o.tmp_in[i] = h2f(in[i]);
        xnn_status st;
        if (o.kind == Kind::conv_f16) st = REAL(xnn_setup_convolution2d_nhwc_f32)(o.inner, nullptr, o.tmp_in.data(), o.tmp_out.data());
        else st = REAL(xnn_setup_fully_connected_nc_f32)(o.inner, o.tmp_in.data(), o.tmp_out.data());
        if (st != xnn_status_success) return st;
        st = REAL(xnn_run_operator)(o.inner, tp);
        if (st != xnn_status_success) return st;
        #pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)o.out_elems; i++) out[i] = f2h(o.tmp_out[i]);
        return xnn_status_success;
    }
    }
    return xnn_status_invalid_state;
}

template <typename T>
xnn_status transpose_nd(const void* input, void* output, size_t num_dims, const size_t* shape, const size_t* perm)
{
    if (num_dims == 0 || num_dims > 8) return xnn_status_unsupported_parameter;
    size_t in_stride[8], out_shape[8], src_stride[8];
    size_t total = 1;
    for (size_t i = num_dims; i-- > 0;) { in_stride[i] = total; total *= shape[i]; }
    for (size_t i = 0; i < num_dims; i++) { out_shape[i] = shape[perm[i]]; src_stride[i] = in_stride[perm[i]]; }
    const T* in = (const T*)input; T* out = (T*)output;
    size_t outer = out_shape[0];
    size_t inner_total = total / (outer ? outer : 1);
    #pragma omp parallel for schedule(static)
    for (long long o = 0; o < (long long)outer; o++) {
        size_t idx[8] = { 0 };
        size_t src = (size_t)o * src_stride[0];
        T* dst = out + (size_t)o * inner_total;
        for (size_t n = 0; n < inner_total; n++) {
            dst[n] = in[src];
            // increment the multi-index over dims 1..num_dims-1
            for (size_t d = num_dims; d-- > 1;) {
                idx[d]++; src += src_stride[d];
                if (idx[d] < out_shape[d]) break;
                src -= src_stride[d] * out_shape[d]; idx[d] = 0;
            }
        }
    }
    return xnn_status_success;
}

} // namespace

extern "C" {

// ---- run / delete interposition ------------------------------------------------------------------------

enum xnn_status xnn_run_operator(xnn_operator_t op, pthreadpool_t threadpool)
{
    if (ShimOp* s = as_shim(op)) return run_shim(*s, threadpool);
    return REAL(xnn_run_operator)(op, threadpool);
}

enum xnn_status xnn_delete_operator(xnn_operator_t op)
{
    if (ShimOp* s = as_shim(op)) {
        { std::lock_guard<std::mutex> lk(g_mutex); g_ops.erase((void*)op); }
        xnn_status st = xnn_status_success;
        if (s->inner) st = REAL(xnn_delete_operator)(s->inner);
        delete s;
        return st;
    }
    return REAL(xnn_delete_operator)(op);
}

// ---- softmax -------------------------------------------------------------------------------------------

enum xnn_status xnn_create_softmax_nc_f32(uint32_t flags, xnn_operator_t* out)
{ auto* o = new ShimOp(); o->kind = Kind::softmax_f32; *out = register_op(o); return xnn_status_success; }
enum xnn_status xnn_create_softmax_nc_f16(uint32_t flags, xnn_operator_t* out)
{ auto* o = new ShimOp(); o->kind = Kind::softmax_f16; *out = register_op(o); return xnn_status_success; }
enum xnn_status xnn_create_softmax_nc_qu8(float input_scale, uint8_t output_zero_point, float output_scale, uint32_t flags, xnn_operator_t* out)
{
    auto* o = new ShimOp(); o->kind = Kind::softmax_qu8; o->in_scale = input_scale; o->out_zp = output_zero_point; o->out_scale = output_scale;
    *out = register_op(o); return xnn_status_success;
}

static xnn_status reshape_softmax(xnn_operator_t op, size_t channels, size_t input_stride, size_t output_stride, size_t batch_size)
{
    ShimOp* o = as_shim(op);
    if (!o || input_stride != channels || output_stride != channels) return xnn_status_invalid_parameter;
    o->channels = channels; o->batch = batch_size; return xnn_status_success;
}
enum xnn_status xnn_reshape_softmax_nc_f32(xnn_operator_t op, size_t c, size_t is, size_t os, size_t b, pthreadpool_t) { return reshape_softmax(op, c, is, os, b); }
enum xnn_status xnn_reshape_softmax_nc_f16(xnn_operator_t op, size_t c, size_t is, size_t os, size_t b, pthreadpool_t) { return reshape_softmax(op, c, is, os, b); }
enum xnn_status xnn_reshape_softmax_nc_qu8(xnn_operator_t op, size_t c, size_t is, size_t os, size_t b, pthreadpool_t) { return reshape_softmax(op, c, is, os, b); }

static xnn_status setup_io(xnn_operator_t op, const void* in, void* out)
{ ShimOp* o = as_shim(op); if (!o) return xnn_status_invalid_parameter; o->in0 = in; o->out = out; return xnn_status_success; }
enum xnn_status xnn_setup_softmax_nc_f32(xnn_operator_t op, const float* in, float* out) { return setup_io(op, in, out); }
enum xnn_status xnn_setup_softmax_nc_f16(xnn_operator_t op, const void* in, void* out) { return setup_io(op, in, out); }
enum xnn_status xnn_setup_softmax_nc_qu8(xnn_operator_t op, const uint8_t* in, uint8_t* out) { return setup_io(op, in, out); }

// ---- dynamic fully connected ---------------------------------------------------------------------------

enum xnn_status xnn_create_dynamic_fully_connected_nc_f32(float, float, uint32_t flags, xnn_operator_t* out)
{ auto* o = new ShimOp(); o->kind = Kind::dynfc_f32; o->flags = flags; *out = register_op(o); return xnn_status_success; }
enum xnn_status xnn_create_dynamic_fully_connected_nc_f16(float, float, uint32_t flags, xnn_operator_t* out)
{ auto* o = new ShimOp(); o->kind = Kind::dynfc_f16; o->flags = flags; *out = register_op(o); return xnn_status_success; }

static xnn_status reshape_dynfc(xnn_operator_t op, size_t batch, size_t ic, size_t oc, size_t is, size_t os, size_t* ws, size_t* wa)
{
    ShimOp* o = as_shim(op);
    if (!o || is != ic || os != oc) return xnn_status_invalid_parameter;
    o->batch = batch; o->K = ic; o->N = oc; if (ws) *ws = 0; if (wa) *wa = 1; return xnn_status_success;
}
enum xnn_status xnn_reshape_dynamic_fully_connected_nc_f32(xnn_operator_t op, size_t b, size_t ic, size_t oc, size_t is, size_t os, size_t* ws, size_t* wa, pthreadpool_t)
{ return reshape_dynfc(op, b, ic, oc, is, os, ws, wa); }
enum xnn_status xnn_reshape_dynamic_fully_connected_nc_f16(xnn_operator_t op, size_t b, size_t ic, size_t oc, size_t is, size_t os, size_t* ws, size_t* wa, pthreadpool_t)
{ return reshape_dynfc(op, b, ic, oc, is, os, ws, wa); }

static xnn_status setup_dynfc(xnn_operator_t op, const void* in, const void* k, const void* b, void* out)
{ ShimOp* o = as_shim(op); if (!o) return xnn_status_invalid_parameter; o->in0 = in; o->in1 = k; o->in2 = b; o->out = out; return xnn_status_success; }
enum xnn_status xnn_setup_dynamic_fully_connected_nc_f32(xnn_operator_t op, void*, const float* in, const float* k, const float* b, float* out) { return setup_dynfc(op, in, k, b, out); }
enum xnn_status xnn_setup_dynamic_fully_connected_nc_f16(xnn_operator_t op, void*, const void* in, const void* k, const void* b, void* out) { return setup_dynfc(op, in, k, b, out); }

// ---- scaled dot product attention ----------------------------------------------------------------------

enum xnn_status xnn_create_scaled_dot_product_attention_nhtc_f32(enum xnn_attention_logits_cap_type cap, const void*, uint32_t, xnn_operator_t* out)
{ if (cap != xnn_attention_logits_cap_type_none) return xnn_status_unsupported_parameter; auto* o = new ShimOp(); o->kind = Kind::sdpa_f32; *out = register_op(o); return xnn_status_success; }
enum xnn_status xnn_create_scaled_dot_product_attention_nhtc_f16(enum xnn_attention_logits_cap_type cap, const void*, uint32_t, xnn_operator_t* out)
{ if (cap != xnn_attention_logits_cap_type_none) return xnn_status_unsupported_parameter; auto* o = new ShimOp(); o->kind = Kind::sdpa_f16; *out = register_op(o); return xnn_status_success; }

static xnn_status reshape_sdpa(xnn_operator_t op, size_t B, size_t Hq, size_t Tq, size_t Hkv, size_t Tk, size_t D, size_t Dv, size_t* ws, size_t* wa)
{
    ShimOp* o = as_shim(op);
    if (!o || Hkv == 0 || (Hkv != 1 && Hq % Hkv != 0)) return xnn_status_invalid_parameter;
    o->B = B; o->Hq = Hq; o->Tq = Tq; o->Hkv = Hkv; o->Tk = Tk; o->Dqk = D; o->Dv = Dv; if (ws) *ws = 0; if (wa) *wa = 1; return xnn_status_success;
}
enum xnn_status xnn_reshape_scaled_dot_product_attention_nhtc_f32(xnn_operator_t op, size_t B, size_t Hq, size_t Tq, size_t Hkv, size_t Tk, size_t D, size_t Dv, size_t* ws, size_t* wa, pthreadpool_t)
{ return reshape_sdpa(op, B, Hq, Tq, Hkv, Tk, D, Dv, ws, wa); }
enum xnn_status xnn_reshape_scaled_dot_product_attention_nhtc_f16(xnn_operator_t op, size_t B, size_t Hq, size_t Tq, size_t Hkv, size_t Tk, size_t D, size_t Dv, size_t* ws, size_t* wa, pthreadpool_t)
{ return reshape_sdpa(op, B, Hq, Tq, Hkv, Tk, D, Dv, ws, wa); }

static xnn_status setup_sdpa(xnn_operator_t op, const void* q, const void* k, const void* v, const void* scale, const void* mask, void* out)
{ ShimOp* o = as_shim(op); if (!o) return xnn_status_invalid_parameter; o->in0 = q; o->in1 = k; o->in2 = v; o->in3 = scale; o->in4 = mask; o->out = out; return xnn_status_success; }
enum xnn_status xnn_setup_scaled_dot_product_attention_nhtc_f32(xnn_operator_t op, void*, const float* q, const float* k, const float* v, const float* s, const float* m, float* out)
{ return setup_sdpa(op, q, k, v, s, m, out); }
enum xnn_status xnn_setup_scaled_dot_product_attention_nhtc_f16(xnn_operator_t op, void*, const void* q, const void* k, const void* v, const void* s, const void* m, void* out)
{ return setup_sdpa(op, q, k, v, s, m, out); }

// ---- transpose -----------------------------------------------------------------------------------------

enum xnn_status xnn_run_transpose_nd_x8(const void* in, void* out, size_t nd, const size_t* shape, const size_t* perm, uint32_t, pthreadpool_t)
{ return transpose_nd<uint8_t>(in, out, nd, shape, perm); }
enum xnn_status xnn_run_transpose_nd_x16(const void* in, void* out, size_t nd, const size_t* shape, const size_t* perm, uint32_t, pthreadpool_t)
{ return transpose_nd<uint16_t>(in, out, nd, shape, perm); }
enum xnn_status xnn_run_transpose_nd_x32(const void* in, void* out, size_t nd, const size_t* shape, const size_t* perm, uint32_t, pthreadpool_t)
{ return transpose_nd<uint32_t>(in, out, nd, shape, perm); }

// ---- NCHW convolution: unsupported stubs ---------------------------------------------------------------

enum xnn_status xnn_create_convolution2d_nchw_f32(uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, size_t, size_t, size_t, size_t, const float*, const float*, float, float, uint32_t, xnn_code_cache_t, xnn_weights_cache_t, xnn_operator_t*)
{ return xnn_status_unsupported_hardware; }
enum xnn_status xnn_create_convolution2d_nchw_f16(uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, size_t, size_t, size_t, size_t, const void*, const void*, float, float, uint32_t, xnn_code_cache_t, xnn_weights_cache_t, xnn_operator_t*)
{ return xnn_status_unsupported_hardware; }
enum xnn_status xnn_reshape_convolution2d_nchw_f32(xnn_operator_t, size_t, size_t, size_t, size_t*, size_t*, pthreadpool_t) { return xnn_status_unsupported_hardware; }
enum xnn_status xnn_reshape_convolution2d_nchw_f16(xnn_operator_t, size_t, size_t, size_t, size_t*, size_t*, pthreadpool_t) { return xnn_status_unsupported_hardware; }
enum xnn_status xnn_setup_convolution2d_nchw_f32(xnn_operator_t, const float*, float*) { return xnn_status_unsupported_hardware; }
enum xnn_status xnn_setup_convolution2d_nchw_f16(xnn_operator_t, const void*, void*) { return xnn_status_unsupported_hardware; }

// ---- f16 Conv / FC interposition (fp16 storage, real XNNPACK fp32 arithmetic, fp16 rounding of the result) ----

enum xnn_status xnn_create_convolution2d_nhwc_f16(
    uint32_t pt, uint32_t pr, uint32_t pb, uint32_t pl, uint32_t kh, uint32_t kw, uint32_t sh, uint32_t sw, uint32_t dh, uint32_t dw,
    uint32_t groups, size_t gic, size_t goc, size_t ics, size_t ocs, const void* kernel, const void* bias,
    float omin, float omax, uint32_t flags, xnn_code_cache_t cc, xnn_weights_cache_t wc, xnn_operator_t* out)
{
    if (native_f16())
        return REAL(xnn_create_convolution2d_nhwc_f16)(pt, pr, pb, pl, kh, kw, sh, sw, dh, dw, groups, gic, goc, ics, ocs, kernel, bias, omin, omax, flags, cc, wc, out);
    size_t wn = (size_t)groups * goc * kh * kw * gic, bn = (size_t)groups * goc;
    std::vector<float> wf(wn + 16), bf(bn + 16);
    const uint16_t* w = (const uint16_t*)kernel; const uint16_t* b = (const uint16_t*)bias;
    for (size_t i = 0; i < wn; i++) wf[i] = h2f(w[i]);
    if (b) for (size_t i = 0; i < bn; i++) bf[i] = h2f(b[i]);
    auto* o = new ShimOp(); o->kind = Kind::conv_f16; o->cin = ics; o->cout = ocs;
    xnn_status st = REAL(xnn_create_convolution2d_nhwc_f32)(pt, pr, pb, pl, kh, kw, sh, sw, dh, dw, groups, gic, goc, ics, ocs, wf.data(), b ? bf.data() : nullptr, omin, omax, 0, nullptr, nullptr, &o->inner);
    if (st != xnn_status_success) { delete o; return st; }
    *out = register_op(o); return xnn_status_success;
}

enum xnn_status xnn_reshape_convolution2d_nhwc_f16(xnn_operator_t op, size_t batch, size_t ih, size_t iw, size_t* ws, size_t* wa, size_t* oh_out, size_t* ow_out, pthreadpool_t tp)
{
    ShimOp* o = as_shim(op);
    if (!o) return REAL(xnn_reshape_convolution2d_nhwc_f16)(op, batch, ih, iw, ws, wa, oh_out, ow_out, tp);
    size_t oh = 0, ow = 0, ws2 = 0, wa2 = 0;
    xnn_status st = REAL(xnn_reshape_convolution2d_nhwc_f32)(o->inner, batch, ih, iw, &ws2, &wa2, &oh, &ow, tp);
    if (st != xnn_status_success) return st;
    o->in_elems = batch * ih * iw * o->cin; o->out_elems = batch * oh * ow * o->cout;
    if (ws) *ws = 0; if (wa) *wa = 1; if (oh_out) *oh_out = oh; if (ow_out) *ow_out = ow;
    return xnn_status_success;
}

enum xnn_status xnn_setup_convolution2d_nhwc_f16(xnn_operator_t op, void* workspace, const void* in, void* out)
{
    ShimOp* o = as_shim(op);
    if (!o) return REAL(xnn_setup_convolution2d_nhwc_f16)(op, workspace, in, out);
    o->in0 = in; o->out = out; return xnn_status_success;
}

enum xnn_status xnn_create_fully_connected_nc_f16(size_t ic, size_t oc, size_t is, size_t os, const void* kernel, const void* bias,
    float omin, float omax, uint32_t flags, xnn_code_cache_t cc, xnn_weights_cache_t wc, xnn_operator_t* out)
{
    if (native_f16())
        return REAL(xnn_create_fully_connected_nc_f16)(ic, oc, is, os, kernel, bias, omin, omax, flags, cc, wc, out);
    size_t wn = ic * oc;
    std::vector<float> wf(wn + 16), bf(oc + 16);
    const uint16_t* w = (const uint16_t*)kernel; const uint16_t* b = (const uint16_t*)bias;
    for (size_t i = 0; i < wn; i++) wf[i] = h2f(w[i]);
    if (b) for (size_t i = 0; i < oc; i++) bf[i] = h2f(b[i]);
    auto* o = new ShimOp(); o->kind = Kind::fc_f16; o->cin = is; o->cout = os;
    xnn_status st = REAL(xnn_create_fully_connected_nc_f32)(ic, oc, is, os, wf.data(), b ? bf.data() : nullptr, omin, omax, flags, nullptr, nullptr, &o->inner);
    if (st != xnn_status_success) { delete o; return st; }
    *out = register_op(o); return xnn_status_success;
}

enum xnn_status xnn_reshape_fully_connected_nc_f16(xnn_operator_t op, size_t batch, pthreadpool_t tp)
{
    ShimOp* o = as_shim(op);
    if (!o) return REAL(xnn_reshape_fully_connected_nc_f16)(op, batch, tp);
    xnn_status st = REAL(xnn_reshape_fully_connected_nc_f32)(o->inner, batch, tp);
    if (st != xnn_status_success) return st;
    o->in_elems = batch * o->cin; o->out_elems = batch * o->cout; return xnn_status_success;
}

enum xnn_status xnn_setup_fully_connected_nc_f16(xnn_operator_t op, const void* in, void* out)
{
    ShimOp* o = as_shim(op);
    if (!o) return REAL(xnn_setup_fully_connected_nc_f16)(op, in, out);
    o->in0 = in; o->out = out; return xnn_status_success;
}

} // extern "C"
