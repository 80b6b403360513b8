// C-ABI glue: error state, device query, dispatch of the streaming products between the tcgen05 and the
// SIMT implementations.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace xb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
    cached_dev = dev;
  }
  return cached;
}

// project_simt.cu
int project_S_simt(const float*, int64_t, int64_t, int64_t, const float*, const float*, const float*, const uint8_t*,
                   const float*, int64_t, int64_t, float*, int64_t, float*, cudaStream_t);
int project_T_simt(const float*, int64_t, int64_t, int64_t, const float*, const float*, const float*, const uint8_t*,
                   const float*, int64_t, int64_t, float*, int64_t, float*, cudaStream_t);
// project_tc.cu
bool tc_supported(int64_t T, int64_t S, int64_t ldx, const float* X, int64_t l);
bool tensor_maps_available();
static inline bool tensor_maps_ok() { return tensor_maps_available(); }
int64_t tc_workspace_bytes(int64_t T, int64_t S, int64_t l, int algo);
int project_S_tc(const float*, int64_t, int64_t, int64_t, const float*, const float*, const float*, const uint8_t*,
                 const float*, int64_t, int64_t, float*, int64_t, void*, int64_t, int, cudaStream_t);
int project_T_tc(const float*, int64_t, int64_t, int64_t, const float*, const float*, const float*, const uint8_t*,
                 const float*, int64_t, int64_t, float*, int64_t, void*, int64_t, int, bool, cudaStream_t, uint16_t*, int64_t,
                 const float*);
int h16_scales(const float*, const float*, int64_t, float*, float*, cudaStream_t);
int project_S16_tc(const uint16_t*, int64_t, int64_t, int64_t, const float*, const float*, const float*, int64_t, int64_t, float*,
                   int64_t, void*, cudaStream_t);
int project_T16_tc(const uint16_t*, int64_t, int64_t, int64_t, const float*, const float*, const float*, int64_t, int64_t, float*,
                   int64_t, void*, cudaStream_t);

int project_S_stats_tc(const float*, int64_t, int64_t, int64_t, const double*, int, const float*, int64_t, int64_t, float*,
                       float*, uint8_t*, float*, float*, float*, double*, int32_t*, float*, int64_t, void*, cudaStream_t,
                       uint16_t*, int64_t, float*, float*, float*);
}  // namespace xb

using namespace xb;

extern "C" int xeofs_b200_version(void) { return 100; }
extern "C" const char* xeofs_b200_last_error(void) { return g_err; }

extern "C" int xeofs_b200_has_tcgen05(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

static int resolve_algo(int algo, int64_t T, int64_t S, int64_t ldx, const float* X, int64_t l) {
  if ((algo == XEOFS_ALGO_TF32X2 || algo == XEOFS_ALGO_TF32X1R || algo == XEOFS_ALGO_TF32X1F) && !(xeofs_b200_has_tcgen05() && tc_supported(T, S, ldx, X, l)))
    return XEOFS_ALGO_SIMT;
  if (algo == XEOFS_ALGO_AUTO || algo == XEOFS_ALGO_AUTO_FAST) {
    const bool tc = xeofs_b200_has_tcgen05() && tc_supported(T, S, ldx, X, l);
    if (!tc) return XEOFS_ALGO_SIMT;
    return algo == XEOFS_ALGO_AUTO ? XEOFS_ALGO_TF32X3 : XEOFS_ALGO_TF32X1;
  }
  return algo;
}

namespace xb {
__global__ void round_tf32_kernel(float* __restrict__ M, int64_t rows, int64_t cols, int64_t ld) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
    float* q = M + r * ld + c;
    *q = __uint_as_float(__float_as_uint(*q) & 0xffffe000u);
  }
}
}  // namespace xb

extern "C" int xeofs_b200_round_tf32(float* M, int64_t rows, int64_t cols, int64_t ld, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(M && rows > 0 && cols > 0 && ld >= cols, "round_tf32: bad arguments");
  dim3 grid((unsigned)ceil_div(cols, 256), (unsigned)imin(rows, 16384));
  round_tf32_kernel<<<grid, 256, 0, stream>>>(M, rows, cols, ld);
  XB_LAUNCH_CHECK();
  return XEOFS_OK;
}

extern "C" int64_t xeofs_b200_project_workspace_bytes(int64_t T, int64_t S, int64_t l, int algo) {
  algo &= ~XEOFS_ALGO_FLAG_NO_NAN;
  int64_t simt = 2 * lpad(l) * (int64_t)sizeof(float) + 256;
  int64_t tc = tc_workspace_bytes(T, S, l, algo);
  return simt > tc ? simt : tc;
}

static int check_project_args(const char* who, const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot,
                              const float* dscale, const void* a, const void* b, int64_t lda, int64_t ldb, int64_t l,
                              void* ws, int64_t ws_bytes, int algo) {
  XB_CHECK_ARG(X && pivot && dscale && a && b && ws, "%s: null pointer", who);
  XB_CHECK_ARG(T > 0 && S > 0 && ldx >= S, "%s: bad field shape T=%lld S=%lld ldx=%lld", who, (long long)T, (long long)S, (long long)ldx);
  XB_CHECK_ARG(l > 0 && l <= 128, "%s: l=%lld must be in 1..128", who, (long long)l);
  XB_CHECK_ARG(lda % 4 == 0 && ldb % 4 == 0, "%s: leading dimensions of the small matrices must be multiples of 4", who);
  XB_CHECK_ARG(((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) && ((uintptr_t)ws % 256 == 0), "%s: misaligned buffer", who);
  XB_CHECK_ARG(algo >= XEOFS_ALGO_AUTO && algo <= XEOFS_ALGO_TF32X1F, "%s: unknown algo %d", who, algo);
  if (ws_bytes < xeofs_b200_project_workspace_bytes(T, S, l, algo)) {
    set_error("%s: workspace too small (%lld < %lld bytes)", who, (long long)ws_bytes,
              (long long)xeofs_b200_project_workspace_bytes(T, S, l, algo));
    return XEOFS_E_WORKSPACE;
  }
  return XEOFS_OK;
}

extern "C" int xeofs_b200_project_S(const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot,
                                    const float* dscale, const float* ccorr, const uint8_t* row_valid, const float* W,
                                    int64_t ldw, int64_t l, float* Yt, int64_t ldy, void* workspace, int64_t workspace_bytes, int algo,
                                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  algo &= ~XEOFS_ALGO_FLAG_NO_NAN;  // (project_S tests for NaN only in stages that hold an all-NaN sample)
  int rc = check_project_args("project_S", X, T, S, ldx, pivot, dscale, W, Yt, ldw, ldy, l, workspace, workspace_bytes, algo);
  if (rc) return rc;
  XB_CHECK_ARG(ldw >= lpad(l) && ldy >= S, "project_S: ldw=%lld must be >= lp and ldy=%lld >= S", (long long)ldw, (long long)ldy);
  algo = resolve_algo(algo, T, S, ldx, X, l);
  if (algo == XEOFS_ALGO_SIMT)
    return project_S_simt(X, T, S, ldx, pivot, dscale, ccorr, row_valid, W, ldw, l, Yt, ldy, (float*)workspace, stream);
  if (!xeofs_b200_has_tcgen05() || !tc_supported(T, S, ldx, X, l)) {
    set_error("project_S: tcgen05 path unavailable for this device/shape (need sm_100, ldx %% 4 == 0, 16-byte aligned X)");
    return XEOFS_E_UNSUPPORTED;
  }
  return project_S_tc(X, T, S, ldx, pivot, dscale, ccorr, row_valid, W, ldw, l, Yt, ldy, workspace, workspace_bytes, algo, stream);
}

extern "C" int xeofs_b200_project_T(const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot,
                                    const float* dscale, const float* ccorr, const uint8_t* row_valid, const float* Yt,
                                    int64_t ldy, int64_t l, float* Z, int64_t ldz, void* workspace, int64_t workspace_bytes, int algo,
                                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool no_nan = (algo & XEOFS_ALGO_FLAG_NO_NAN) != 0;
  algo &= ~XEOFS_ALGO_FLAG_NO_NAN;
  int rc = check_project_args("project_T", X, T, S, ldx, pivot, dscale, Yt, Z, ldy, ldz, l, workspace, workspace_bytes, algo);
  if (rc) return rc;
  XB_CHECK_ARG(ldz >= lpad(l) && ldy >= S, "project_T: ldz=%lld must be >= lp and ldy=%lld >= S", (long long)ldz, (long long)ldy);
  algo = resolve_algo(algo, T, S, ldx, X, l);
  if (algo == XEOFS_ALGO_SIMT)
    return project_T_simt(X, T, S, ldx, pivot, dscale, ccorr, row_valid, Yt, ldy, l, Z, ldz, (float*)workspace + lpad(l), stream);
  if (!xeofs_b200_has_tcgen05() || !tc_supported(T, S, ldx, X, l)) {
    set_error("project_T: tcgen05 path unavailable for this device/shape (need sm_100, ldx %% 4 == 0, 16-byte aligned X)");
    return XEOFS_E_UNSUPPORTED;
  }
  return project_T_tc(X, T, S, ldx, pivot, dscale, ccorr, row_valid, Yt, ldy, l, Z, ldz, workspace, workspace_bytes, algo,
                      no_nan, stream, nullptr, 0, nullptr);
}

// ---- the half-precision copy for the power iterations (include/xeofs_b200.h)
extern "C" int xeofs_b200_h16_scales(const float* dscale, const float* std, int64_t S, float* e16, float* ic16, void* stream_) {
  XB_CHECK_ARG(dscale && std && e16 && ic16 && S > 0, "h16_scales: bad arguments");
  return h16_scales(dscale, std, S, e16, ic16, (cudaStream_t)stream_);
}

extern "C" int xeofs_b200_project_T_h16copy(const float* X, int64_t T, int64_t S, int64_t ldx, const float* pivot,
                                            const float* dscale, const float* Yt, int64_t ldy, int64_t l, float* Z,
                                            int64_t ldz, void* workspace, int64_t workspace_bytes, int no_nan,
                                            const float* e16, void* copy16, int64_t ldc, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_project_args("project_T_h16copy", X, T, S, ldx, pivot, dscale, Yt, Z, ldy, ldz, l, workspace, workspace_bytes,
                              XEOFS_ALGO_TF32X1);
  if (rc) return rc;
  XB_CHECK_ARG(e16 && copy16 && ((uintptr_t)copy16 % 16 == 0) && ldz >= lpad(l) && ldy >= S, "project_T_h16copy: bad arguments");
  if (!xeofs_b200_has_tcgen05() || !tc_supported(T, S, ldx, X, l)) {
    set_error("project_T_h16copy: needs the tcgen05 path");
    return XEOFS_E_UNSUPPORTED;
  }
  return project_T_tc(X, T, S, ldx, pivot, dscale, nullptr, nullptr, Yt, ldy, l, Z, ldz, workspace, workspace_bytes,
                      XEOFS_ALGO_TF32X1, no_nan != 0, stream, (uint16_t*)copy16, ldc, e16);
}

static int check_h16_args(const char* who, const void* A16, int64_t T, int64_t S, int64_t ldc, const float* ic16, const void* a,
                          const void* b, int64_t l, void* ws, int64_t ws_bytes) {
  XB_CHECK_ARG(A16 && ic16 && a && b && ws && T > 0 && S > 0 && l > 0 && l <= 128, "%s: bad arguments", who);
  XB_CHECK_ARG(ldc >= S && ldc % 8 == 0 && ((uintptr_t)A16 % 16 == 0) && ((uintptr_t)ws % 256 == 0), "%s: misaligned copy / workspace", who);
  if (ws_bytes < xeofs_b200_project_workspace_bytes(T, S, l, XEOFS_ALGO_TF32X1)) {
    set_error("%s: workspace too small", who);
    return XEOFS_E_WORKSPACE;
  }
  if (!xeofs_b200_has_tcgen05() || !tensor_maps_ok()) {
    set_error("%s: needs the tcgen05 path", who);
    return XEOFS_E_UNSUPPORTED;
  }
  return XEOFS_OK;
}

extern "C" int xeofs_b200_project_S16(const void* A16, int64_t T, int64_t S, int64_t ldc, const float* ic16, const float* cc16,
                                      const float* W, int64_t ldw, int64_t l, float* Yt, int64_t ldy, void* workspace,
                                      int64_t workspace_bytes, void* stream_) {
  int rc = check_h16_args("project_S16", A16, T, S, ldc, ic16, W, Yt, l, workspace, workspace_bytes);
  if (rc) return rc;
  XB_CHECK_ARG(ldw >= lpad(l) && ldy >= S, "project_S16: ldw=%lld must be >= lp and ldy=%lld >= S", (long long)ldw, (long long)ldy);
  return project_S16_tc((const uint16_t*)A16, T, S, ldc, ic16, cc16, W, ldw, l, Yt, ldy, workspace, (cudaStream_t)stream_);
}

extern "C" int xeofs_b200_project_T16(const void* A16, int64_t T, int64_t S, int64_t ldc, const float* ic16, const float* cc16,
                                      const float* Yt, int64_t ldy, int64_t l, float* Z, int64_t ldz, void* workspace,
                                      int64_t workspace_bytes, void* stream_) {
  int rc = check_h16_args("project_T16", A16, T, S, ldc, ic16, Yt, Z, l, workspace, workspace_bytes);
  if (rc) return rc;
  XB_CHECK_ARG(ldz >= lpad(l) && ldy >= S, "project_T16: ldz=%lld must be >= lp and ldy=%lld >= S", (long long)ldz, (long long)ldy);
  return project_T16_tc((const uint16_t*)A16, T, S, ldc, ic16, cc16, Yt, ldy, l, Z, ldz, workspace, (cudaStream_t)stream_);
}

extern "C" int xeofs_b200_project_S_stats(const float* X, int64_t T, int64_t S, int64_t ldx, const double* featw, int flags,
                                          const float* W, int64_t ldw, int64_t l, float* mean, float* std, uint8_t* valid,
                                          float* pivot, float* dscale, float* ccorr, double* scalars_out, int32_t* row_nan,
                                          float* Yt, int64_t ldy, void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(X && W && mean && std && valid && pivot && dscale && ccorr && scalars_out && row_nan && Yt && workspace,
               "project_S_stats: null pointer");
  XB_CHECK_ARG(T > 0 && S > 0 && ldx >= S && l > 0 && l <= 128 && ldw >= lpad(l) && ldy >= S && ldw % 4 == 0,
               "project_S_stats: bad shape");
  if (workspace_bytes < xeofs_b200_project_workspace_bytes(T, S, l, XEOFS_ALGO_TF32X1)) {
    set_error("project_S_stats: workspace too small");
    return XEOFS_E_WORKSPACE;
  }
  if (!xeofs_b200_has_tcgen05() || !tc_supported(T, S, ldx, X, l)) {
    set_error("project_S_stats: needs the tcgen05 path (sm_100, 16-byte aligned X, ldx %% 4 == 0)");
    return XEOFS_E_UNSUPPORTED;
  }
  return project_S_stats_tc(X, T, S, ldx, featw, flags, W, ldw, l, mean, std, valid, pivot, dscale, ccorr, scalars_out,
                            row_nan, Yt, ldy, workspace, stream, nullptr, 0, nullptr, nullptr, nullptr);
}

extern "C" int xeofs_b200_project_S_stats_h16copy(const float* X, int64_t T, int64_t S, int64_t ldx, const double* featw,
                                                  int flags, const float* W, int64_t ldw, int64_t l, float* mean, float* std,
                                                  uint8_t* valid, float* pivot, float* dscale, float* ccorr,
                                                  double* scalars_out, int32_t* row_nan, float* Yt, int64_t ldy,
                                                  void* workspace, int64_t workspace_bytes, void* copy16, int64_t ldc,
                                                  float* c0, float* ic16, float* cc16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  XB_CHECK_ARG(X && W && mean && std && valid && pivot && dscale && ccorr && scalars_out && row_nan && Yt && workspace &&
                   copy16 && c0 && ic16 && cc16,
               "project_S_stats_h16copy: null pointer");
  XB_CHECK_ARG(T > 0 && S > 0 && ldx >= S && l > 0 && l <= 128 && ldw >= lpad(l) && ldy >= S && ldw % 4 == 0 && ldc >= S &&
                   ldc % 8 == 0 && ((uintptr_t)copy16 % 16 == 0),
               "project_S_stats_h16copy: bad shape");
  if (workspace_bytes < xeofs_b200_project_workspace_bytes(T, S, l, XEOFS_ALGO_TF32X1)) {
    set_error("project_S_stats_h16copy: workspace too small");
    return XEOFS_E_WORKSPACE;
  }
  if (!xeofs_b200_has_tcgen05() || !tc_supported(T, S, ldx, X, l)) {
    set_error("project_S_stats_h16copy: needs the tcgen05 path (sm_100, 16-byte aligned X, ldx %% 4 == 0)");
    return XEOFS_E_UNSUPPORTED;
  }
  return project_S_stats_tc(X, T, S, ldx, featw, flags, W, ldw, l, mean, std, valid, pivot, dscale, ccorr, scalars_out,
                            row_nan, Yt, ldy, workspace, stream, (uint16_t*)copy16, ldc, c0, ic16, cc16);
}
