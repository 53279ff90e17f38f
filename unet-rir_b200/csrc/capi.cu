// extern "C" surface of liburir (declared in include/urir.h): argument validation, path selection
// (tcgen05 implicit GEMM vs CUDA-core direct conv), error strings, launch accounting.
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>

#include "urir_common.cuh"

namespace urir {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches_all{0}, g_launches_tc{0};

void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
int fail(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}
// urir_conv2d_dgrad_sums: the statistics epilogue of the call in flight on this host thread needs channel sums only
static thread_local bool g_sums_only = false;
bool stats_sums_only() { return g_sums_only; }
static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
    int v = g_pdl.load();
    if (v < 0) { const char* e = getenv("URIR_NO_PDL"); v = (e && e[0] == '1') ? 0 : 1; g_pdl.store(v); }
    return v == 1;
}
int sm_count() {
    static std::atomic<int> n{0};
    int v = n.load();
    if (v <= 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
            v = 4 * 37;                     // no device visible (host-side queries in the CPU test suite): B200's count
        cudaGetLastError();
        n.store(v);
    }
    return v;
}

// deterministic mode: see urir_common.cuh
static std::atomic<int> g_det{-1};
bool deterministic() {
    int v = g_det.load();
    if (v < 0) { const char* e = getenv("URIR_DETERMINISTIC"); v = (e && e[0] == '1') ? 1 : 0; g_det.store(v); }
    return v == 1;
}
constexpr int GATE_SLOTS = 4096;
__device__ unsigned int g_gate_slots[GATE_SLOTS];
unsigned int* next_gate() {
    if (!deterministic()) return nullptr;
    static unsigned int* base = nullptr;
    static std::atomic<unsigned int> next{0};
    if (base == nullptr) {
        void* p = nullptr;
        if (cudaGetSymbolAddress(&p, g_gate_slots) != cudaSuccess) return nullptr;
        base = static_cast<unsigned int*>(p);
    }
    return base + (next.fetch_add(1) % GATE_SLOTS);
}
void count_launch(int kind) { g_launches_all++; if (kind == 1) g_launches_tc++; }

// per-kernel-family call counters of the convolution entry points (urir_family_calls): parity tests assert through them
// that the dispatch they exercised is the one the benchmark times (halo / stride-2 halo / up-2 / deep / ...)
static const char* const g_family_names[URIR_FAM_COUNT] = {
    "simt", "igemm", "halo", "halo_s2_fprop", "halo_up2", "thin_gemm", "head_fprop", "wgrad_tc", "wgrad_halo",
    "wgrad_halo_s2", "thin_wgrad", "deep"};
static std::atomic<long long> g_family_calls[URIR_FAM_COUNT];
static inline int fam(int family, int rc) { if (rc == URIR_OK) g_family_calls[family]++; return rc; }

// implemented in the other translation units
int conv_fprop_simt_dispatch(const urir_conv_desc*, const void*, const void*, const float*, void*, float*, cudaStream_t);
int conv_dgrad_simt_dispatch(const urir_conv_desc*, const void*, const void*, const float*, void*, float*, cudaStream_t, const void*);
int conv_wgrad_simt_dispatch(const urir_conv_desc*, const void*, const void*, float*, cudaStream_t);
int weight_prep(const float*, void*, void*, int, int, int, cudaStream_t);
int weight_prep_batched(const long long*, int, cudaStream_t);
int weight_fold_bn_batched(const long long*, int, cudaStream_t);
bool igemm_fprop_supported(const urir_conv_desc*);
bool igemm_dgrad_supported(const urir_conv_desc*);
int conv_fprop_igemm(const urir_conv_desc*, const void*, const void*, const float*, void*, float*, cudaStream_t);
int conv_dgrad_igemm(const urir_conv_desc*, const void*, const void*, const float*, void*, float*, cudaStream_t);
bool wgrad_tc_supported(const urir_conv_desc*);
bool thin_supported(const urir_conv_desc*, int op);
bool halo_supported(const urir_conv_desc*, int op, bool forced);
int conv_halo(const urir_conv_desc*, int op, const void*, const void*, const float*, void*, float*, cudaStream_t);
bool halo_up2_supported(const urir_conv_desc*);
bool halo_s2_fprop_supported(const urir_conv_desc*, bool forced);
int conv_halo_s2_fprop(const urir_conv_desc*, const void*, const void*, const float*, void*, float*, cudaStream_t);
int conv_halo_up2(const urir_conv_desc*, const void*, const void*, const float*, void*, cudaStream_t);
int weight_prep_up2(const float*, void*, int, int, cudaStream_t);
bool deep_supported(const urir_conv_desc*, int op, bool wide_ok);
int conv_deep(const urir_conv_desc*, int op, const void*, const void*, const float*, void*, float*, cudaStream_t);
bool head_fprop_supported(const urir_conv_desc*);
int head_fprop(const urir_conv_desc*, const void*, const void*, const float*, void*, cudaStream_t);
int thin_gemm(const urir_conv_desc*, const void*, const void*, const float*, void*, bool, cudaStream_t);
int thin_wgrad(const urir_conv_desc*, const void*, const void*, float*, cudaStream_t);
int conv_wgrad_tc(const urir_conv_desc*, const void*, const void*, float*, cudaStream_t);
bool wgrad_halo_supported(const urir_conv_desc*, bool forced);
bool wgrad_halo_s2_supported(const urir_conv_desc*, bool forced);
int conv_wgrad_halo_s2(const urir_conv_desc*, const void*, const void*, float*, cudaStream_t);
int conv_wgrad_halo(const urir_conv_desc*, const void*, const void*, float*, cudaStream_t);
int bn_finalize(const float*, double, const float*, const float*, float*, float*, float, float, int, float*, float*, int, cudaStream_t);
int bn_relu_fwd(const void*, int, int, const float*, void*, int, int, long long, int, int, cudaStream_t);
int bn_relu_bwd_reduce(const void*, int, int, const void*, int, int, const float*, const float*, float*, long long, int, int, cudaStream_t);
int bn_relu_fwd_train(const void*, int, int, const float*, double, const float*, const float*, float*, float*, float, float, int,
                      float*, float*, void*, int, int, long long, int, cudaStream_t);
int bn_relu_bwd_apply(const void*, int, int, const void*, int, int, const float*, const float*, const float*, const float*,
                      void*, int, int, float*, float*, float*, long long, int, int, cudaStream_t);
int channel_sum(const void*, int, long long, int, int, int, float*, cudaStream_t);
int ampphase_loss(const float*, const float*, long long, float, float, int, float*, float*, void*, int, cudaStream_t);
int mse2_loss(const float*, const float*, long long, float, int, float*, float*, cudaStream_t);
int adam(float*, const float*, float*, float*, long long, const float*, const int*, float, float, float, cudaStream_t);
int sgd(float*, const float*, long long, const float*, cudaStream_t);
int nadam(float*, const float*, float*, float*, long long, const float*, const int*, float*, float, float, float, cudaStream_t);
int lamb(float*, const float*, float*, float*, float*, const long long*, int, float*, const float*, const int*, float, float,
         float, float, cudaStream_t);
int step_increment(int*, cudaStream_t);
int axpy(float*, const float*, float, long long, cudaStream_t);
int sumsq(const float*, long long, float, float*, int, cudaStream_t);
int add_bf16(const void*, const void*, void*, long long, cudaStream_t);
int add_bf16_strided(const void*, int, int, const void*, int, int, void*, int, int, long long, int, cudaStream_t);
int l2_reg_batched(const long long*, int, float, float*, cudaStream_t);
int cast_f32_to_bf16(const float*, void*, long long, cudaStream_t);
int cast_pad_bf16(const float*, void*, long long, int, int, cudaStream_t);
int embedding_fwd(const int*, const float*, void*, int, int, int, int, cudaStream_t);
int embedding_bwd(const int*, const void*, int, float*, int, int, int, int, cudaStream_t);
int dense_fwd(const void*, const void*, const void*, const float*, const float*, void*, int, int, int, cudaStream_t);
int dense_bwd(const void*, const void*, const void*, const void*, const float*, void*, float*, float*, void*, int, int, int, cudaStream_t);
int dropout_mask(float*, long long, float, uint64_t, const int*, cudaStream_t);
int stft_ampphase(const float*, int, const urir_stft_desc*, float*, cudaStream_t);
int istft_from_ampphase(const float*, int, const urir_stft_desc*, float*, cudaStream_t);

static int check_conv(const urir_conv_desc* d, const char* who) {
    URIR_CHECK_ARG(d != nullptr, "%s: null descriptor", who);
    URIR_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0 && d->C > 0 && d->K > 0 && d->R > 0 && d->S > 0 && d->P > 0 && d->Q > 0,
                   "%s: non-positive dimension", who);
    URIR_CHECK_ARG(d->stride >= 1 && d->stride <= 2, "%s: stride %d unsupported (1 or 2)", who, d->stride);
    URIR_CHECK_ARG(d->pad_top >= 0 && d->pad_left >= 0 && d->pad_top < d->R && d->pad_left < d->S, "%s: bad padding", who);
    URIR_CHECK_ARG((d->P - 1) * d->stride - d->pad_top < d->H && (d->Q - 1) * d->stride - d->pad_left < d->W,
                   "%s: output grid (%d,%d) reaches past the input", who, d->P, d->Q);
    URIR_CHECK_ARG(d->x_ld >= d->x_coff + d->C && d->y_ld >= d->y_coff + d->K && d->x_coff >= 0 && d->y_coff >= 0,
                   "%s: channel slice exceeds the buffer pitch", who);
    URIR_CHECK_ARG((d->x_dtype == URIR_F32 || d->x_dtype == URIR_BF16) && (d->y_dtype == URIR_F32 || d->y_dtype == URIR_BF16),
                   "%s: bad dtype", who);
    return URIR_OK;
}

static int env_force_simt() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("URIR_FORCE_SIMT"); v = (e && e[0] == '1') ? 1 : 0; }
    return v;
}

}  // namespace urir

using namespace urir;

extern "C" {

int urir_version(void) { return URIR_VERSION; }
const char* urir_last_error(void) { return g_err; }
int urir_set_deterministic(int enabled) { const int prev = deterministic() ? 1 : 0; g_det.store(enabled ? 1 : 0); return prev; }
int urir_set_pdl(int enabled) { const int prev = pdl_enabled() ? 1 : 0; g_pdl.store(enabled ? 1 : 0); return prev; }
long long urir_launch_count(int kind) { return kind == 1 ? g_launches_tc.load() : g_launches_all.load(); }
long long urir_family_calls(int family) { return (family >= 0 && family < URIR_FAM_COUNT) ? g_family_calls[family].load() : -1; }
const char* urir_family_name(int family) { return (family >= 0 && family < URIR_FAM_COUNT) ? g_family_names[family] : ""; }

int urir_conv2d_fprop(const urir_conv_desc* d, const void* x, const void* w_ck, const void* w_kc, const float* bias,
                      void* y, float* stats, void* stream) {
    int rc = check_conv(d, "conv2d_fprop"); if (rc) return rc;
    URIR_CHECK_ARG(x && y, "conv2d_fprop: null tensor");
    cudaStream_t st = (cudaStream_t)stream;
    if (d->impl != URIR_IMPL_SIMT && !env_force_simt() && w_ck && !stats && thin_supported(d, 0))
        return fam(URIR_FAM_THIN_GEMM, thin_gemm(d, x, w_ck, bias, y, true, st));                       // the 2-channel stem
    if (d->impl != URIR_IMPL_SIMT && !env_force_simt() && w_ck && !stats && head_fprop_supported(d))
        return fam(URIR_FAM_HEAD_FPROP, head_fprop(d, x, w_ck, bias, y, st));                            // the 2-channel head
    if (w_kc && d->stride == 2 && d->impl != URIR_IMPL_SIMT && d->impl != URIR_IMPL_TC && !env_force_simt() &&
        halo_s2_fprop_supported(d, d->impl == URIR_IMPL_HALO))
        return fam(URIR_FAM_HALO_S2_FPROP, conv_halo_s2_fprop(d, x, w_kc, bias, y, stats, st));
    if (d->impl == URIR_IMPL_DEEP && !(w_kc && deep_supported(d, 0, true)))
        return fail(URIR_ERR_UNSUP, "conv2d_fprop: shape not supported by the deep-layer tcgen05 path");
    if (w_kc && (d->impl == URIR_IMPL_DEEP || (d->impl == URIR_IMPL_AUTO && !env_force_simt() && deep_supported(d, 0, !halo_supported(d, 0, false)))))
        return fam(URIR_FAM_DEEP, conv_deep(d, 0, x, w_kc, bias, y, stats, st));
    if (d->impl == URIR_IMPL_HALO && !(w_kc && halo_supported(d, 0, true)))
        return fail(URIR_ERR_UNSUP, "conv2d_fprop: shape not supported by the halo-tile tcgen05 path");
    if (w_kc && ((d->impl == URIR_IMPL_HALO) || (d->impl == URIR_IMPL_AUTO && !env_force_simt() && halo_supported(d, 0, false))))
        return fam(URIR_FAM_HALO, conv_halo(d, 0, x, w_kc, bias, y, stats, st));
    const bool tc_ok = igemm_fprop_supported(d) && w_kc != nullptr;
    if (d->impl == URIR_IMPL_TC && !tc_ok) return fail(URIR_ERR_UNSUP, "conv2d_fprop: shape not supported by the tcgen05 path");
    const bool use_tc = d->impl == URIR_IMPL_TC || (d->impl == URIR_IMPL_AUTO && tc_ok && !env_force_simt());
    return use_tc ? fam(URIR_FAM_IGEMM, conv_fprop_igemm(d, x, w_kc, bias, y, stats, st))
                  : fam(URIR_FAM_SIMT, conv_fprop_simt_dispatch(d, x, w_ck, bias, y, stats, st));
}

int urir_conv2d_dgrad(const urir_conv_desc* d, const void* dy, const void* w_ck, const void* w_kc, const float* bias,
                      void* dx, float* stats, void* stream) {
    int rc = check_conv(d, "conv2d_dgrad"); if (rc) return rc;
    URIR_CHECK_ARG(dy && dx, "conv2d_dgrad: null tensor");
    URIR_CHECK_ARG(d->act == URIR_ACT_NONE, "conv2d_dgrad: activation not supported");
    cudaStream_t st = (cudaStream_t)stream;
    if (d->impl != URIR_IMPL_SIMT && !env_force_simt() && w_ck && !stats && !bias && thin_supported(d, 1))
        return fam(URIR_FAM_THIN_GEMM, thin_gemm(d, dy, w_ck, nullptr, dx, false, st));                 // the 2-channel head
    const bool deep_dg_ok = !(stats && d->stride == 2);            // no statistics epilogue on the stride-2 classes
    if (d->impl == URIR_IMPL_DEEP && !(w_ck && deep_dg_ok && deep_supported(d, 1, true)))
        return fail(URIR_ERR_UNSUP, "conv2d_dgrad: shape not supported by the deep-layer tcgen05 path");
    if (w_ck && (d->impl == URIR_IMPL_DEEP || (d->impl == URIR_IMPL_AUTO && !env_force_simt() && deep_dg_ok && deep_supported(d, 1, !halo_supported(d, 1, false)))))
        return fam(URIR_FAM_DEEP, conv_deep(d, 1, dy, w_ck, bias, dx, stats, st));
    if (d->impl == URIR_IMPL_HALO && !(w_ck && halo_supported(d, 1, true)))
        return fail(URIR_ERR_UNSUP, "conv2d_dgrad: shape not supported by the halo-tile tcgen05 path");
    if (w_ck && ((d->impl == URIR_IMPL_HALO) || (d->impl == URIR_IMPL_AUTO && !env_force_simt() && halo_supported(d, 1, false))))
        return fam(URIR_FAM_HALO, conv_halo(d, 1, dy, w_ck, bias, dx, stats, st));
    const bool tc_ok = igemm_dgrad_supported(d) && w_ck != nullptr;
    if (d->impl == URIR_IMPL_TC && !tc_ok) return fail(URIR_ERR_UNSUP, "conv2d_dgrad: shape not supported by the tcgen05 path");
    const bool use_tc = d->impl == URIR_IMPL_TC || (d->impl == URIR_IMPL_AUTO && tc_ok && !env_force_simt());
    return use_tc ? fam(URIR_FAM_IGEMM, conv_dgrad_igemm(d, dy, w_ck, bias, dx, stats, st))
                  : fam(URIR_FAM_SIMT, conv_dgrad_simt_dispatch(d, dy, w_kc, bias, dx, stats, st, w_ck));
}

int urir_conv2d_dgrad_sums(const urir_conv_desc* d, const void* dy, const void* w_ck, const void* w_kc, const float* bias,
                           void* dx, float* stats, void* stream) {
    urir::g_sums_only = true;
    const int rc = urir_conv2d_dgrad(d, dy, w_ck, w_kc, bias, dx, stats, stream);
    urir::g_sums_only = false;
    return rc;
}

int urir_conv2d_wgrad(const urir_conv_desc* d, const void* x, const void* dy, float* dw, void* stream) {
    int rc = check_conv(d, "conv2d_wgrad"); if (rc) return rc;
    URIR_CHECK_ARG(x && dy && dw, "conv2d_wgrad: null tensor");
    cudaStream_t st = (cudaStream_t)stream;
    if (d->impl != URIR_IMPL_SIMT && !env_force_simt() && thin_supported(d, 2)) return fam(URIR_FAM_THIN_WGRAD, thin_wgrad(d, x, dy, dw, st));
    if (d->stride == 2 && (d->impl == URIR_IMPL_HALO || (d->impl == URIR_IMPL_AUTO && !env_force_simt())) &&
        wgrad_halo_s2_supported(d, d->impl == URIR_IMPL_HALO))
        return fam(URIR_FAM_WGRAD_HALO_S2, conv_wgrad_halo_s2(d, x, dy, dw, st));
    if (d->impl == URIR_IMPL_HALO && !wgrad_halo_supported(d, true))
        return fail(URIR_ERR_UNSUP, "conv2d_wgrad: shape not supported by the halo-tile tcgen05 path");
    if (d->impl == URIR_IMPL_HALO || (d->impl == URIR_IMPL_AUTO && !env_force_simt() && wgrad_halo_supported(d, false)))
        return fam(URIR_FAM_WGRAD_HALO, conv_wgrad_halo(d, x, dy, dw, st));
    const bool tc_ok = wgrad_tc_supported(d);
    if (d->impl == URIR_IMPL_TC && !tc_ok) return fail(URIR_ERR_UNSUP, "conv2d_wgrad: shape not supported by the tcgen05 path");
    const bool use_tc = d->impl == URIR_IMPL_TC || (d->impl == URIR_IMPL_AUTO && tc_ok && !env_force_simt());
    return use_tc ? fam(URIR_FAM_WGRAD_TC, conv_wgrad_tc(d, x, dy, dw, st)) : fam(URIR_FAM_SIMT, conv_wgrad_simt_dispatch(d, x, dy, dw, st));
}

int urir_conv2d_dgrad_up2(const urir_conv_desc* d, const void* dy, const void* w_up2, const float* bias, void* dx, void* stream) {
    int rc = check_conv(d, "conv2d_dgrad_up2"); if (rc) return rc;
    URIR_CHECK_ARG(dy && dx && w_up2, "conv2d_dgrad_up2: null tensor");
    if (!halo_up2_supported(d)) return fail(URIR_ERR_UNSUP, "conv2d_dgrad_up2: shape not supported (3x3 stride 2 on an even input, 4C <= 512, resident weights)");
    return fam(URIR_FAM_HALO_UP2, conv_halo_up2(d, dy, w_up2, bias, dx, (cudaStream_t)stream));
}
int urir_weight_prep_up2(const float* w_hwio, void* w_up2, int C, int K, void* stream) {
    URIR_CHECK_ARG(w_hwio && w_up2 && C > 0 && K > 0, "weight_prep_up2: bad args");
    return weight_prep_up2(w_hwio, w_up2, C, K, (cudaStream_t)stream);
}

int urir_conv_path(const urir_conv_desc* d, int op) {
    if (op == 3) return (d && d->impl != URIR_IMPL_SIMT && !(d->impl == URIR_IMPL_AUTO && env_force_simt()) && halo_up2_supported(d)) ? 1 : 0;
    if (!d || d->impl == URIR_IMPL_SIMT || (d->impl == URIR_IMPL_AUTO && env_force_simt())) return 0;
    if (thin_supported(d, op)) return 1;
    if (op < 2 && deep_supported(d, op, true)) return 1;
    if (op < 2 && halo_supported(d, op, d->impl == URIR_IMPL_HALO)) return 1;
    if (op == 0 && halo_s2_fprop_supported(d, d->impl == URIR_IMPL_HALO)) return 1;
    if (op == 2 && wgrad_halo_supported(d, d->impl == URIR_IMPL_HALO)) return 1;
    if (op == 2 && wgrad_halo_s2_supported(d, d->impl == URIR_IMPL_HALO)) return 1;
    if (op == 0 && head_fprop_supported(d)) return 1;
    if (op == 0) return igemm_fprop_supported(d) ? 1 : 0;
    if (op == 1) return igemm_dgrad_supported(d) ? 1 : 0;
    return wgrad_tc_supported(d) ? 1 : 0;
}

int urir_weight_prep(const float* w, void* w_ck, void* w_kc, int taps, int C, int K, void* stream) {
    URIR_CHECK_ARG(w && (w_ck || w_kc) && taps > 0 && C > 0 && K > 0, "weight_prep: bad args");
    return weight_prep(w, w_ck, w_kc, taps, C, K, (cudaStream_t)stream);
}

int urir_weight_prep_batched(const int64_t* table_dev, int n_entries, void* stream) {
    URIR_CHECK_ARG(table_dev && n_entries > 0, "weight_prep_batched: bad args");
    return weight_prep_batched(reinterpret_cast<const long long*>(table_dev), n_entries, (cudaStream_t)stream);
}

int urir_weight_fold_bn_batched(const int64_t* table_dev, int n_entries, void* stream) {
    URIR_CHECK_ARG(table_dev && n_entries > 0, "weight_fold_bn_batched: bad args");
    return weight_fold_bn_batched(reinterpret_cast<const long long*>(table_dev), n_entries, (cudaStream_t)stream);
}

int urir_channel_sum(const void* x, int dtype, long long npix, int C, int ld, int coff, float* out, void* stream) {
    URIR_CHECK_ARG(x && out, "channel_sum: null tensor");
    return channel_sum(x, dtype, npix, C, ld, coff, out, (cudaStream_t)stream);
}

int urir_bn_finalize(const float* stats, double count, const float* gamma, const float* beta, float* mm, float* mv,
                     float momentum, float eps, int unbiased, float* scale_shift, float* mean_rstd, int C, void* stream) {
    return bn_finalize(stats, count, gamma, beta, mm, mv, momentum, eps, unbiased, scale_shift, mean_rstd, C, (cudaStream_t)stream);
}
int urir_bn_relu_fwd(const void* x, int x_ld, int x_coff, const float* ss, void* y, int y_ld, int y_coff, long long npix,
                     int C, int relu, void* stream) {
    URIR_CHECK_ARG(x && y && ss && npix > 0, "bn_relu_fwd: bad args");
    return bn_relu_fwd(x, x_ld, x_coff, ss, y, y_ld, y_coff, npix, C, relu, (cudaStream_t)stream);
}
int urir_bn_relu_fwd_train(const void* x, int x_ld, int x_coff, const float* stats, double count, const float* gamma,
                           const float* beta, float* moving_mean, float* moving_var, float momentum, float eps,
                           int unbiased_moving_var, float* scale_shift, float* mean_rstd, void* y, int y_ld, int y_coff,
                           long long npix, int C, void* stream) {
    return bn_relu_fwd_train(x, x_ld, x_coff, stats, count, gamma, beta, moving_mean, moving_var, momentum, eps,
                             unbiased_moving_var, scale_shift, mean_rstd, y, y_ld, y_coff, npix, C, (cudaStream_t)stream);
}
int urir_bn_relu_bwd_reduce(const void* dy, int dy_ld, int dy_coff, const void* x, int x_ld, int x_coff, const float* ss,
                            const float* mr, float* sums, long long npix, int C, int prezeroed, void* stream) {
    URIR_CHECK_ARG(dy && x && ss && mr && sums && npix > 0, "bn_relu_bwd_reduce: bad args");
    return bn_relu_bwd_reduce(dy, dy_ld, dy_coff, x, x_ld, x_coff, ss, mr, sums, npix, C, prezeroed, (cudaStream_t)stream);
}
int urir_bn_relu_bwd_apply(const void* dy, int dy_ld, int dy_coff, const void* x, int x_ld, int x_coff, const float* ss,
                           const float* mr, const float* gamma, const float* sums, void* dx, int dx_ld, int dx_coff,
                           float* dgamma, float* dbeta, float* dbias, long long npix, int C, int prezeroed, void* stream) {
    URIR_CHECK_ARG(dy && x && ss && mr && sums && dx && npix > 0, "bn_relu_bwd_apply: bad args");
    return bn_relu_bwd_apply(dy, dy_ld, dy_coff, x, x_ld, x_coff, ss, mr, gamma, sums, dx, dx_ld, dx_coff, dgamma, dbeta,
                             dbias, npix, C, prezeroed, (cudaStream_t)stream);
}

int urir_embedding_fwd(const int32_t* idx, const float* table, void* out, int B, int T, int D, int vocab, void* stream) {
    URIR_CHECK_ARG(idx && table && out && B > 0 && T > 0, "embedding_fwd: bad args");
    return embedding_fwd(idx, table, out, B, T, D, vocab, (cudaStream_t)stream);
}
int urir_embedding_bwd(const int32_t* idx, const void* dx, int dx_dtype, float* dtable, int B, int T, int D, int vocab,
                       void* stream) {
    URIR_CHECK_ARG(idx && dx && dtable && B > 0 && T > 0, "embedding_bwd: bad args");
    URIR_CHECK_ARG(dx_dtype == URIR_F32 || dx_dtype == URIR_BF16, "embedding_bwd: bad dtype");
    return embedding_bwd(idx, dx, dx_dtype, dtable, B, T, D, vocab, (cudaStream_t)stream);
}
int urir_dense_fwd(const void* x, const void* w_kn, const void* w_nk, const float* bias, const float* mask, void* out,
                   int B, int Kd, int N, void* stream) {
    URIR_CHECK_ARG(x && w_kn && w_nk && out && B > 0 && Kd > 0 && N > 0, "dense_fwd: bad args");
    return dense_fwd(x, w_kn, w_nk, bias, mask, out, B, Kd, N, (cudaStream_t)stream);
}
int urir_dense_bwd(const void* x, const void* w_kn, const void* w_nk, const void* dy, const float* mask, void* dy_eff,
                   float* dw, float* db, void* dx, int B, int Kd, int N, void* stream) {
    URIR_CHECK_ARG(x && w_kn && w_nk && dy && B > 0 && Kd > 0 && N > 0, "dense_bwd: bad args");
    return dense_bwd(x, w_kn, w_nk, dy, mask, dy_eff, dw, db, dx, B, Kd, N, (cudaStream_t)stream);
}
int urir_dropout_mask(float* mask, long long n, float rate, uint64_t seed, const int32_t* step_dev, void* stream) {
    URIR_CHECK_ARG(mask && n > 0, "dropout_mask: bad args");
    return dropout_mask(mask, n, rate, seed, step_dev, (cudaStream_t)stream);
}

int urir_ampphase_loss(const float* y_true, const float* y_pred, long long npix, float w_amp, float w_ph,
                       int sigmoid_bwd, float* losses, float* grad, void* grad_bf16, int grad_bf16_ld, void* stream) {
    URIR_CHECK_ARG(y_true && y_pred && losses, "ampphase_loss: null tensor");
    return ampphase_loss(y_true, y_pred, npix, w_amp, w_ph, sigmoid_bwd, losses, grad, grad_bf16, grad_bf16_ld,
                         (cudaStream_t)stream);
}

int urir_mse2_loss(const float* y_true, const float* y_pred, long long npix, float w, int sigmoid_bwd, float* losses,
                   float* grad, void* stream) {
    URIR_CHECK_ARG(y_true && y_pred && losses, "mse2_loss: null tensor");
    return mse2_loss(y_true, y_pred, npix, w, sigmoid_bwd, losses, grad, (cudaStream_t)stream);
}

int urir_adam(float* p, const float* g, float* m, float* v, long long n, const float* lr_dev, const int32_t* step_dev,
              float beta1, float beta2, float eps, void* stream) {
    URIR_CHECK_ARG(p && g && m && v && lr_dev && step_dev, "adam: null tensor");
    return adam(p, g, m, v, n, lr_dev, step_dev, beta1, beta2, eps, (cudaStream_t)stream);
}
int urir_nadam(float* p, const float* g, float* m, float* v, long long n, const float* lr_dev, const int32_t* step_dev,
               float* coef_dev, float beta1, float beta2, float eps, void* stream) {
    URIR_CHECK_ARG(p && g && m && v && lr_dev && step_dev && coef_dev && n > 0, "nadam: bad args");
    return nadam(p, g, m, v, n, lr_dev, step_dev, coef_dev, beta1, beta2, eps, (cudaStream_t)stream);
}
int urir_lamb(float* p, const float* g, float* m, float* v, float* upd, const int64_t* table_dev, int n_vars, float* norms,
              const float* lr_dev, const int32_t* step_dev, float beta1, float beta2, float eps, float weight_decay,
              void* stream) {
    URIR_CHECK_ARG(p && g && m && v && upd && table_dev && norms && lr_dev && step_dev && n_vars > 0, "lamb: bad args");
    return lamb(p, g, m, v, upd, reinterpret_cast<const long long*>(table_dev), n_vars, norms, lr_dev, step_dev, beta1,
                beta2, eps, weight_decay, (cudaStream_t)stream);
}
int urir_sgd(float* p, const float* g, long long n, const float* lr_dev, void* stream) {
    URIR_CHECK_ARG(p && g && lr_dev && n > 0, "sgd: bad args");
    return sgd(p, g, n, lr_dev, (cudaStream_t)stream);
}
int urir_step_increment(int32_t* step_dev, void* stream) {
    URIR_CHECK_ARG(step_dev, "step_increment: null");
    return step_increment(step_dev, (cudaStream_t)stream);
}
int urir_axpy(float* y, const float* x, float a, long long n, void* stream) {
    URIR_CHECK_ARG(y && x && n > 0, "axpy: bad args");
    return axpy(y, x, a, n, (cudaStream_t)stream);
}
int urir_sumsq(const float* x, long long n, float scale, float* out, int accumulate, void* stream) {
    URIR_CHECK_ARG(x && out && n > 0, "sumsq: bad args");
    return sumsq(x, n, scale, out, accumulate, (cudaStream_t)stream);
}
int urir_l2_reg_batched(const int64_t* table_dev, int n_entries, float coef, float* out, void* stream) {
    URIR_CHECK_ARG(table_dev && out && n_entries > 0, "l2_reg_batched: bad args");
    return l2_reg_batched(reinterpret_cast<const long long*>(table_dev), n_entries, coef, out, (cudaStream_t)stream);
}
int urir_add_bf16_strided(const void* a, int a_ld, int a_coff, const void* b, int b_ld, int b_coff, void* out, int out_ld,
                          int out_coff, long long npix, int C, void* stream) {
    URIR_CHECK_ARG(a && b && out && npix > 0 && C > 0, "add_bf16_strided: bad args");
    return add_bf16_strided(a, a_ld, a_coff, b, b_ld, b_coff, out, out_ld, out_coff, npix, C, (cudaStream_t)stream);
}
int urir_add_bf16(const void* a, const void* b, void* out, long long n, void* stream) {
    URIR_CHECK_ARG(a && b && out && n > 0, "add_bf16: bad args");
    return add_bf16(a, b, out, n, (cudaStream_t)stream);
}
int urir_cast_f32_to_bf16(const float* x, void* y, long long n, void* stream) {
    URIR_CHECK_ARG(x && y && n > 0, "cast: bad args");
    return cast_f32_to_bf16(x, y, n, (cudaStream_t)stream);
}

int urir_cast_pad_bf16(const float* x, void* y, long long npix, int C, int ld, void* stream) {
    URIR_CHECK_ARG(x && y, "cast_pad_bf16: null tensor");
    return cast_pad_bf16(x, y, npix, C, ld, (cudaStream_t)stream);
}

int urir_stft_ampphase(const float* wav, int B, const urir_stft_desc* d, float* spec, void* stream) {
    URIR_CHECK_ARG(wav && spec, "stft: null tensor");
    return stft_ampphase(wav, B, d, spec, (cudaStream_t)stream);
}
int urir_istft_from_ampphase(const float* spec, int B, const urir_stft_desc* d, float* wav, void* stream) {
    URIR_CHECK_ARG(wav && spec, "istft: null tensor");
    return istft_from_ampphase(spec, B, d, wav, (cudaStream_t)stream);
}

}  // extern "C"
