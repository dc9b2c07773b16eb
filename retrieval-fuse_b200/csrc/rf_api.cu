// C-ABI plumbing: error string, device query, kNN dispatch.
#include <stdarg.h>
#include <string.h>

#include "rf_common.cuh"

static thread_local char g_err[512] = "";

void rf_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* rf_last_error(void) { return g_err; }
extern "C" int rf_version(void) { return 100; }

extern "C" int rf_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
    cudaDeviceProp p;
    RF_CUDA_OK(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    RF_CHECK_ARG(p.major == 10, "rf_b200 kernels are built for sm_100a only; device %d is sm_%d%d", device, p.major, p.minor);
    return 0;
}

int rf_tc_conv_init();
int rf_tc_conv_halo_init();
int rf_tc_linear_init();
int rf_tc_mlp_init();
int rf_knn_tc_init();

struct rf_handle_s {
    int device, sm_count;
};

/* The library keeps no mutable global state: kernels take everything through their arguments, the only per-device
 * setup is the opt-in to large dynamic shared memory, which rf_create performs for every kernel of `device` (the launch
 * sites repeat it lazily, so calls without a handle work too).  The handle records the device it was created for. */
extern "C" int rf_create(int device, rf_handle** out) {
    RF_CHECK_ARG(out, "rf_create: null output pointer");
    *out = nullptr;
    int sm = 0, major = 0, minor = 0;
    if (int rc = rf_device_info(device, &sm, &major, &minor)) return rc;
    int prev = 0;
    RF_CUDA_OK(cudaGetDevice(&prev));
    RF_CUDA_OK(cudaSetDevice(device));
    int rc = rf_tc_conv_init();
    if (!rc) rc = rf_tc_conv_halo_init();
    if (!rc) rc = rf_tc_linear_init();
    if (!rc) rc = rf_tc_mlp_init();
    if (!rc) rc = rf_knn_tc_init();
    cudaSetDevice(prev);
    if (rc) return rc;
    rf_handle* h = new rf_handle_s;
    h->device = device;
    h->sm_count = sm;
    *out = h;
    return 0;
}

extern "C" int rf_destroy(rf_handle* h) {
    delete h;
    return 0;
}

extern "C" int rf_handle_device(const rf_handle* h, int* device, int* sm_count) {
    RF_CHECK_ARG(h, "rf_handle_device: null handle");
    if (device) *device = h->device;
    if (sm_count) *sm_count = h->sm_count;
    return 0;
}

int rf_knn_exact_launch(const float* bank, long n_rows, long row_offset, const float* q, long Q, int k, const int* q_sel,
                        int* out_idx, double* out_d, void* workspace, size_t workspace_bytes, cudaStream_t s);
int rf_knn_exact_nsplit(long Q, long n_rows);
size_t rf_knn_tc_workspace_bytes(long Q, long n_rows, int k, int kblk, int with_image);
size_t rf_knn_tc_image_bytes(long n_rows, int kblk);
size_t rf_knn_tc_prepare_scratch_bytes(long n_rows, int kblk);
int rf_knn_tc_prepare(const float* bank, long n_rows, int kblk, const float* q_sample, long n_sample, void* image,
                      size_t image_bytes, void* scratch, size_t scratch_bytes, cudaStream_t s);
int rf_knn_tc_launch(const float* bank, long n_rows, long row_offset, const void* image, const float* q, long Q, int k, int kblk,
                     int* out_idx, double* out_d, void* workspace, size_t workspace_bytes, cudaStream_t s);

static size_t exact_ws(long Q, long n_rows, int k) {
    const int ns = rf_knn_exact_nsplit(Q, n_rows);
    return ns == 1 ? 256 : (size_t)ns * Q * k * (sizeof(int) + sizeof(double)) + 256;
}

// 0 = auto: the tensor-core candidate pass whenever it applies (k <= 32, enough work to amortise the operand
// images), as ONE fp16 GEMM (method 2).  Measured on a 1 M-row isotropic bank (BASELINE configs[4]) the fp16 pass
// proves every query and is ~3x faster than the bf16 hi/lo split (method 3: three times the MMAs, two instead of six
// pipeline stages); method 3 stays available for banks so dense in near-ties that the 1e-3 score bound of method 2
// sends many queries to the exact re-check (ops.last_knn_stats / rf_knn_tc_stats report that count).
static int resolve_method(int method, long Q, long n_rows, int k) {
    if (method != 0) return method;
    if (k > 32 || n_rows < 1024 || Q * n_rows < (1L << 24)) return 1;
    return 2;
}
// the method a PREPARED image of this bank is built for (no query count yet): tensor cores unless the bank is tiny
static int resolve_bank_method(int method, long n_rows) {
    if (method != 0) return method;
    return n_rows < 1024 ? 1 : 2;
}

extern "C" size_t rf_knn_workspace_bytes(long Q, long n_rows, int k, int method) {
    if (Q <= 0 || n_rows <= 0 || k <= 0) return 0;
    method = resolve_method(method, Q, n_rows, k);
    if (method == 1) return exact_ws(Q, n_rows, k);
    return rf_knn_tc_workspace_bytes(Q, n_rows, k, method == 2 ? 1 : 3, 1);
}

static int check_knn_args(const float* bank, long n_rows, long row_offset, const float* q, long Q, int D, int k, int method,
                          const int* out_idx, const double* out_d) {
    RF_CHECK_ARG(bank && q && out_idx && out_d, "rf_knn_l2_topk: null pointer");
    RF_CHECK_ARG(D == 64, "rf_knn_l2_topk: latent_dim must be 64 (got %d)", D);
    RF_CHECK_ARG(Q > 0 && n_rows > 0, "rf_knn_l2_topk: empty bank or query set");
    RF_CHECK_ARG(k >= 1 && k <= 32 && k <= n_rows, "rf_knn_l2_topk: k=%d out of range (1..min(32, n_rows=%ld))", k, n_rows);
    RF_CHECK_ARG(row_offset >= 0 && row_offset + n_rows < (1L << 31), "rf_knn_l2_topk: row ids exceed int32");
    RF_CHECK_ARG(((uintptr_t)bank & 15) == 0 && ((uintptr_t)q & 15) == 0, "rf_knn_l2_topk: bank / q must be 16-byte aligned");
    RF_CHECK_ARG(method >= 0 && method <= 3, "rf_knn_l2_topk: bad method %d", method);
    return 0;
}

extern "C" int rf_knn_l2_topk(const float* bank, long n_rows, long row_offset, const float* q, long Q, int D, int k,
                              int method, int* out_idx, double* out_d, void* workspace, size_t workspace_bytes,
                              void* stream) {
    if (int rc = check_knn_args(bank, n_rows, row_offset, q, Q, D, k, method, out_idx, out_d)) return rc;
    method = resolve_method(method, Q, n_rows, k);
    if (method >= 2)
        return rf_knn_tc_launch(bank, n_rows, row_offset, nullptr, q, Q, k, method == 2 ? 1 : 3, out_idx, out_d, workspace,
                                workspace_bytes, (cudaStream_t)stream);
    return rf_knn_exact_launch(bank, n_rows, row_offset, q, Q, k, nullptr, out_idx, out_d, workspace, workspace_bytes, (cudaStream_t)stream);
}

/* ---- prepared banks: the tensor-core operand image of a static bank is built once and reused by every lookup */
extern "C" int rf_knn_bank_method(long n_rows, int method) { return resolve_bank_method(method, n_rows); }

extern "C" size_t rf_knn_bank_image_bytes(long n_rows, int method) {
    method = resolve_bank_method(method, n_rows);
    return method >= 2 && n_rows > 0 ? rf_knn_tc_image_bytes(n_rows, method == 2 ? 1 : 3) : 0;
}

extern "C" size_t rf_knn_bank_scratch_bytes(long n_rows, int method) {
    method = resolve_bank_method(method, n_rows);
    return method >= 2 && n_rows > 0 ? rf_knn_tc_prepare_scratch_bytes(n_rows, method == 2 ? 1 : 3) : 0;
}

extern "C" int rf_knn_bank_prepare(const float* bank, long n_rows, int method, const float* q_sample, long n_sample, void* image,
                                   size_t image_bytes, void* scratch, size_t scratch_bytes, void* stream) {
    RF_CHECK_ARG(bank && image && scratch, "rf_knn_bank_prepare: null pointer");
    RF_CHECK_ARG(n_rows > 0 && n_rows < (1L << 31), "rf_knn_bank_prepare: bad row count %ld", n_rows);
    RF_CHECK_ARG(((uintptr_t)bank & 15) == 0 && (!q_sample || ((uintptr_t)q_sample & 15) == 0), "rf_knn_bank_prepare: pointers must be 16-byte aligned");
    method = resolve_bank_method(method, n_rows);
    RF_CHECK_ARG(method == 2 || method == 3, "rf_knn_bank_prepare: method %d keeps no image (tensor-core methods 2 / 3 only)", method);
    return rf_knn_tc_prepare(bank, n_rows, method == 2 ? 1 : 3, q_sample, q_sample ? n_sample : 0, image, image_bytes, scratch,
                             scratch_bytes, (cudaStream_t)stream);
}

extern "C" size_t rf_knn_prepared_workspace_bytes(long Q, long n_rows, int k, int method) {
    if (Q <= 0 || n_rows <= 0 || k <= 0) return 0;
    method = resolve_bank_method(method, n_rows);
    if (method == 1) return exact_ws(Q, n_rows, k);
    return rf_knn_tc_workspace_bytes(Q, n_rows, k, method == 2 ? 1 : 3, 0);
}

extern "C" int rf_knn_l2_topk_prepared(const float* bank, long n_rows, long row_offset, const void* image, int method,
                                       const float* q, long Q, int D, int k, int* out_idx, double* out_d, void* workspace,
                                       size_t workspace_bytes, void* stream) {
    if (int rc = check_knn_args(bank, n_rows, row_offset, q, Q, D, k, method, out_idx, out_d)) return rc;
    method = resolve_bank_method(method, n_rows);
    RF_CHECK_ARG(image && (method == 2 || method == 3), "rf_knn_l2_topk_prepared: needs an image built by rf_knn_bank_prepare (method 2 / 3)");
    return rf_knn_tc_launch(bank, n_rows, row_offset, image, q, Q, k, method == 2 ? 1 : 3, out_idx, out_d, workspace, workspace_bytes,
                            (cudaStream_t)stream);
}
