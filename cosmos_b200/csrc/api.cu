// extern "C" surface of libcosmos_b200.so (declared in include/cosmos_b200.h).
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "infonce.h"
#include "internal.h"
#include "tma_host.h"

namespace {

thread_local int g_last_cuda = 0;   // diagnostic only: last CUDA error code seen by this host thread
inline bool cu_fail(cudaError_t e) {
  if (e == cudaSuccess) return false;
  g_last_cuda = static_cast<int>(e);
  cudaGetLastError();
  return true;
}

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    // Always (re)bind: a host thread that never touched the runtime - PyTorch's autograd worker runs our
    // backward - has no current context yet, and cuTensorMapEncodeTiled needs one (CUDA_ERROR_INVALID_CONTEXT).
    ok = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (ok && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

// Diagnostic switches are read from the environment ONCE per process (function-local statics: thread-safe since C++11), not on
// every call: the library keeps no other host state.
int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
int dbg_flags() {
  static const int v = env_int("COSMOS_B200_DBG", 0);
  return v;
}
int bwd_t_splits() {
  static const int v = env_int("COSMOS_B200_TSPLIT", 1);
  return v;
}
int gemm_cta_cap() {                     // diagnostics (tools/overlap_probe.py): persistent CTAs of cosmos_gemm
  static const int v = env_int("COSMOS_B200_GEMM_CTAS", 0);
  return v;
}

int sm_count_of(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
  return n;
}

}  // namespace

extern "C" {

int cosmos_abi_version(void) { return COSMOS_B200_ABI_VERSION; }

int cosmos_last_cuda_error(void) { return g_last_cuda; }
const char* cosmos_cuda_error_string(int code) { return cudaGetErrorString(static_cast<cudaError_t>(code)); }

const char* cosmos_status_string(int status) {
  switch (status) {
    case COSMOS_OK: return "ok";
    case COSMOS_ERR_INVALID_ARGUMENT: return "invalid argument (shape, null pointer or alignment)";
    case COSMOS_ERR_UNSUPPORTED: return "unsupported dtype or size";
    case COSMOS_ERR_CUDA: return "CUDA runtime or launch failure";
    case COSMOS_ERR_NO_DEVICE: return "device is not an sm_100 (B200) part";
    case COSMOS_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown status";
  }
}

int cosmos_device_check(int device) {
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) {
    cudaGetLastError();
    return COSMOS_ERR_CUDA;
  }
  return major == 10 ? COSMOS_OK : COSMOS_ERR_NO_DEVICE;
}

int64_t cosmos_ema_table_entries(int64_t n_tensors, const int64_t* numel) {
  if (n_tensors < 0 || (n_tensors > 0 && numel == nullptr)) return -1;
  int64_t n = 0;
  for (int64_t i = 0; i < n_tensors; ++i) {
    if (numel[i] < 0) return -1;
    n += (numel[i] + COSMOS_EMA_CHUNK - 1) / COSMOS_EMA_CHUNK;
  }
  return n;
}

int cosmos_ema_table_fill(int64_t n_tensors, const uint64_t* teacher_ptrs, const uint64_t* student_ptrs,
                          const int64_t* numel, int elem_size, cosmos_ema_chunk* table_host) {
  if (n_tensors < 0 || (elem_size != 2 && elem_size != 4)) return COSMOS_ERR_INVALID_ARGUMENT;
  if (n_tensors > 0 && (!teacher_ptrs || !student_ptrs || !numel || !table_host)) return COSMOS_ERR_INVALID_ARGUMENT;
  int64_t e = 0;
  for (int64_t i = 0; i < n_tensors; ++i) {
    if (numel[i] < 0) return COSMOS_ERR_INVALID_ARGUMENT;
    if (numel[i] > 0 && (teacher_ptrs[i] == 0 || student_ptrs[i] == 0)) return COSMOS_ERR_INVALID_ARGUMENT;
    if ((teacher_ptrs[i] % elem_size) || (student_ptrs[i] % elem_size)) return COSMOS_ERR_INVALID_ARGUMENT;
    for (int64_t off = 0; off < numel[i]; off += COSMOS_EMA_CHUNK) {
      cosmos_ema_chunk& c = table_host[e++];
      c.teacher = teacher_ptrs[i] + static_cast<uint64_t>(off) * elem_size;
      c.student = student_ptrs[i] + static_cast<uint64_t>(off) * elem_size;
      const int64_t left = numel[i] - off;
      c.count = static_cast<uint32_t>(left < COSMOS_EMA_CHUNK ? left : COSMOS_EMA_CHUNK);
      c.aligned = ((c.teacher | c.student) & 15) == 0 ? 1u : 0u;
    }
  }
  return COSMOS_OK;
}

int cosmos_ema_apply(const cosmos_ema_chunk* table_dev, int64_t n_entries, double momentum, int dtype, int device,
                     void* stream) {
  if (n_entries < 0 || n_entries > 0x7fffffff || (n_entries > 0 && table_dev == nullptr)) return COSMOS_ERR_INVALID_ARGUMENT;
  if (dtype != COSMOS_DTYPE_F32 && dtype != COSMOS_DTYPE_BF16 && dtype != COSMOS_DTYPE_F16) return COSMOS_ERR_UNSUPPORTED;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  cudaError_t e = cb::launch_ema(table_dev, static_cast<int>(n_entries), momentum, dtype, sm_count_of(device),
                                 static_cast<cudaStream_t>(stream));
  return cu_fail(e) ? COSMOS_ERR_CUDA : COSMOS_OK;
}

int cosmos_clamp_scalars(const uint64_t* ptrs, int32_t n, double lo, double hi, int dtype, int device, void* stream) {
  if (n < 0 || n > COSMOS_CLAMP_MAX || (n > 0 && ptrs == nullptr) || !(lo <= hi)) return COSMOS_ERR_INVALID_ARGUMENT;
  if (dtype != COSMOS_DTYPE_F32 && dtype != COSMOS_DTYPE_BF16 && dtype != COSMOS_DTYPE_F16) return COSMOS_ERR_UNSUPPORTED;
  cb::ClampTable t = {};
  t.n = n;
  const uintptr_t align = dtype == COSMOS_DTYPE_F32 ? 3 : 1;
  for (int i = 0; i < n; ++i) {
    if (ptrs[i] == 0 || (ptrs[i] & align) != 0) return COSMOS_ERR_INVALID_ARGUMENT;
    t.ptr[i] = ptrs[i];
  }
  if (n == 0) return COSMOS_OK;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  return cu_fail(cb::launch_clamp_scalars(t, lo, hi, dtype, static_cast<cudaStream_t>(stream))) ? COSMOS_ERR_CUDA : COSMOS_OK;
}

}  // extern "C"

namespace {

struct Dims {
  int pairs, n_row_tiles, n_col_tiles_fwd, n_col_tiles_bwd, n_slabs, ks, n_parts;
};

int check_problem(const cosmos_infonce_problem* p, Dims* d) {
  if (p == nullptr) return COSMOS_ERR_INVALID_ARGUMENT;
  if (p->dtype != COSMOS_DTYPE_BF16 && p->dtype != COSMOS_DTYPE_F16) return COSMOS_ERR_UNSUPPORTED;
  if (p->dim < 64 || p->dim > 512 || (p->dim % 64) != 0) return COSMOS_ERR_UNSUPPORTED;
  if (p->gx <= 0 || p->gy <= 0 || p->n_rows <= 0 || p->n_cols <= 0 || p->label_offset < 0) return COSMOS_ERR_INVALID_ARGUMENT;
  if (static_cast<int64_t>(p->label_offset) + p->n_rows > p->n_cols) return COSMOS_ERR_INVALID_ARGUMENT;
  if (p->x == 0 || p->y == 0 || p->scale == 0) return COSMOS_ERR_INVALID_ARGUMENT;
  if ((p->x & 15) || (p->y & 15) || (p->scale & 3)) return COSMOS_ERR_INVALID_ARGUMENT;
  if (static_cast<int64_t>(p->gx) * p->gy > (1 << 20)) return COSMOS_ERR_UNSUPPORTED;
  d->pairs = p->gx * p->gy;
  d->n_row_tiles = (p->n_rows + cb::kFwdBM - 1) / cb::kFwdBM;
  d->n_col_tiles_fwd = (p->n_cols + cb::kFwdBN - 1) / cb::kFwdBN;
  d->n_col_tiles_bwd = (p->n_cols + cb::kBwdBN - 1) / cb::kBwdBN;
  d->n_slabs = d->n_row_tiles;      // column partials: one per 128-row tile (merged over its four 32-row slabs in the CTA)
  d->ks = p->dim / 64;
  d->n_parts = (p->dim + cb::kBwdDP - 1) / cb::kBwdDP;
  return COSMOS_OK;
}

int64_t fwd_workspace(const cosmos_infonce_problem* p, const Dims& d) {
  return static_cast<int64_t>(d.pairs) * d.n_slabs * p->n_cols * static_cast<int64_t>(sizeof(float2));
}
constexpr int kMaxTSplits = 4;
int64_t bwd_partials_bytes(const cosmos_infonce_problem* p, const Dims& d) {   // dscale partials, 256-byte aligned
  return ((static_cast<int64_t>(p->gx) * d.n_row_tiles * kMaxTSplits * static_cast<int64_t>(sizeof(float))) + 255) & ~int64_t(255);
}
int64_t bwd_workspace(const cosmos_infonce_problem* p, const Dims& d) {       // + fp32 dX accumulation buffer (split column sweeps)
  return bwd_partials_bytes(p, d) + static_cast<int64_t>(p->gx) * p->n_rows * p->dim * static_cast<int64_t>(sizeof(float));
}
// Column-sweep split of the pair kernel: the fewest slices for which the clusters fill whole waves of SM pairs.
int choose_t_splits(int clusters, int slots, int T_all) {
  int best = 1;
  double best_eff = 0.0;
  for (int ts = 1; ts <= kMaxTSplits && 2 * ts <= T_all; ++ts) {
    const int c = clusters * ts;
    const double eff = static_cast<double>(c) / (static_cast<double>((c + slots - 1) / slots) * slots) - 0.01 * (ts - 1);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = ts; }
  }
  return best;
}

// Slices of the column sweep of the stored-exponential row-side backward: the fewest for which its CTA pairs (one per 256
// rows of one row tensor) fill whole waves.  COSMOS_B200_ROWS_SPLITS forces a value (diagnostics).
int bwd_e_splits(const cosmos_infonce_problem* p, const Dims& d, int device) {
  static const int forced = env_int("COSMOS_B200_ROWS_SPLITS", 0);
  const int steps = p->gy * d.n_col_tiles_bwd;
  if (forced > 0) return (forced <= kMaxTSplits && 2 * forced <= steps) ? forced : 1;
  // at most two slices: measured at the per-rank shapes of an 8-GPU job (b = 4096, N = 32768), two slices take the
  // distillation group from 3.46 to 6.92 waves (8.85 -> 8.24 ms); four slices of the CLIP group's 128 pairs cost more in
  // fp32 partials and per-item prologues than their fuller last wave returns (2.18 -> 2.30 ms)
  const int clusters = p->gx * ((d.n_row_tiles + 1) / 2), slots = sm_count_of(device) / 2;
  if (steps < 4) return 1;
  auto eff = [&](int c) { return static_cast<double>(c) / (static_cast<double>((c + slots - 1) / slots) * slots); };
  return eff(2 * clusters) - 0.01 > eff(clusters) + 1e-9 ? 2 : 1;
}

}  // namespace

extern "C" {

int64_t cosmos_infonce_bwd_e_workspace_bytes(const cosmos_infonce_problem* p, int device) {
  Dims d;
  if (check_problem(p, &d) != COSMOS_OK) return -1;
  const int splits = bwd_e_splits(p, d, device);
  return bwd_partials_bytes(p, d) + (splits > 1 ? static_cast<int64_t>(splits) * p->gx * p->n_rows * 512 * static_cast<int64_t>(sizeof(float)) : 0);
}

int64_t cosmos_infonce_workspace_bytes(const cosmos_infonce_problem* p) {
  Dims d;
  if (check_problem(p, &d) != COSMOS_OK) return -1;
  const int64_t a = fwd_workspace(p, d), b = bwd_workspace(p, d);
  return ((a > b ? a : b) + 255) & ~static_cast<int64_t>(255);
}

int cosmos_infonce_fwd(const cosmos_infonce_problem* p, float* row_lse2, float* diag_raw, float* col_lse2, void* workspace,
                       int64_t workspace_bytes, int device, void* stream) {
  return cosmos_infonce_fwd_e(p, row_lse2, diag_raw, col_lse2, nullptr, nullptr, workspace, workspace_bytes, device, stream);
}

int64_t cosmos_infonce_e_bytes(const cosmos_infonce_problem* p) {
  Dims d;
  if (check_problem(p, &d) != COSMOS_OK) return -1;
  return static_cast<int64_t>(d.pairs) * d.n_row_tiles * d.n_col_tiles_bwd * 32768;
}

int cosmos_infonce_fwd_e(const cosmos_infonce_problem* p, float* row_lse2, float* diag_raw, float* col_lse2, void* e_out,
                         float* off_out, void* workspace, int64_t workspace_bytes, int device, void* stream) {
  Dims d;
  int st = check_problem(p, &d);
  if (st != COSMOS_OK) return st;
  if (!row_lse2 || !diag_raw || !col_lse2 || !workspace) return COSMOS_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return COSMOS_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < fwd_workspace(p, d)) return COSMOS_ERR_WORKSPACE;
  if (e_out != nullptr) {
    if (off_out == nullptr || (reinterpret_cast<uintptr_t>(e_out) & 15) != 0) return COSMOS_ERR_INVALID_ARGUMENT;
  }
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  CUtensorMap tmX, tmY;
  const int bf = p->dtype == COSMOS_DTYPE_BF16;
  {
    const int m1 = cb::make_stack_map(&tmX, reinterpret_cast<const void*>(p->x), bf, p->dim, p->n_rows, p->gx, cb::kFwdBM);
    const bool pair0 = !(dbg_flags() & 4);   // COSMOS_B200_DBG=4: single-CTA forward (diagnostics)
    const int m2 = cb::make_stack_map(&tmY, reinterpret_cast<const void*>(p->y), bf, p->dim, p->n_cols, p->gy,
                                      pair0 ? cb::kFwdBN / 2 : cb::kFwdBN);
    if (m1 != 0 || m2 != 0) {
      g_last_cuda = 100000 + (m1 != 0 ? m1 : m2);   // 100000 + CUresult of cuTensorMapEncodeTiled
      return COSMOS_ERR_CUDA;
    }
  }
  cb::FwdParams fp;
  fp.gx = p->gx; fp.gy = p->gy; fp.n_rows = p->n_rows; fp.n_cols = p->n_cols; fp.ks = d.ks;
  fp.label_offset = p->label_offset;
  fp.n_row_tiles = d.n_row_tiles; fp.n_col_tiles = d.n_col_tiles_fwd; fp.n_slabs = d.n_slabs;
  const bool pair = !(dbg_flags() & 4);
  fp.idesc = cb::make_idesc(bf, 0, 0, pair ? 2 * cb::kFwdBM : cb::kFwdBM, (dbg_flags() & 8192) ? 16 : cb::kFwdBN);   // 8192: issue-rate diagnostics (wrong results)
  fp.dbg = dbg_flags();
  fp.scale = reinterpret_cast<const float*>(p->scale);
  fp.row_lse2 = row_lse2; fp.diag_raw = diag_raw;
  fp.col_part = reinterpret_cast<float2*>(workspace);
  fp.e_out = static_cast<uint16_t*>(e_out);
  fp.off_out = off_out;
  fp.n_steps = d.n_col_tiles_bwd;
  fp.n_chunks = (p->n_cols + 31) / 32;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // (the single-CTA kernel, COSMOS_B200_DBG=4, only has the first-generation epilogue)
  if (cu_fail(cb::launch_infonce_fwd(tmX, tmY, fp, pair, s)))
    return COSMOS_ERR_CUDA;
  if (cu_fail(cb::launch_col_combine(fp.col_part, col_lse2, d.pairs, d.n_slabs, p->n_cols, s))) return COSMOS_ERR_CUDA;
  return COSMOS_OK;
}

int cosmos_scale16(const void* src, void* dst, const float* num, float den, int32_t dtype, int64_t n, int device, void* stream) {
  if (!src || !dst || !num || n < 0 || (n & 7) != 0 || !(den != 0.f)) return COSMOS_ERR_INVALID_ARGUMENT;
  if (dtype != COSMOS_DTYPE_BF16 && dtype != COSMOS_DTYPE_F16) return COSMOS_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(src) & 15) != 0 || (reinterpret_cast<uintptr_t>(dst) & 15) != 0) return COSMOS_ERR_INVALID_ARGUMENT;
  if (n == 0) return COSMOS_OK;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  return cu_fail(cb::launch_scale16(src, dst, num, den, dtype, static_cast<size_t>(n), static_cast<cudaStream_t>(stream)))
             ? COSMOS_ERR_CUDA : COSMOS_OK;
}

int cosmos_lse2_merge(const float* parts, float* out, int32_t n_parts, int64_t n, int device, void* stream) {
  if (!parts || !out || n_parts <= 0 || n < 0) return COSMOS_ERR_INVALID_ARGUMENT;
  if (n == 0) return COSMOS_OK;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  return cu_fail(cb::launch_lse2_merge(parts, out, n_parts, static_cast<size_t>(n), static_cast<cudaStream_t>(stream)))
             ? COSMOS_ERR_CUDA : COSMOS_OK;
}

int cosmos_infonce_loss_sums(const cosmos_infonce_problem* p, const float* row_lse2, const float* diag_raw,
                             const float* col_lse2, int32_t use_rows, int32_t use_cols, float* out, void* workspace,
                             int device, void* stream) {
  (void)workspace;
  Dims d;
  int st = check_problem(p, &d);
  if (st != COSMOS_OK) return st;
  if (!row_lse2 || !diag_raw || !col_lse2 || !out) return COSMOS_ERR_INVALID_ARGUMENT;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  if (cu_fail(cb::launch_loss_sums(row_lse2, diag_raw, col_lse2, reinterpret_cast<const float*>(p->scale), d.pairs, p->n_rows,
                                   p->n_cols, p->label_offset, use_rows, use_cols, out, static_cast<cudaStream_t>(stream))))
    return COSMOS_ERR_CUDA;
  return COSMOS_OK;
}

int cosmos_infonce_bwd(const cosmos_infonce_problem* p, const float* row_lse2, const float* col_lse2, float a_row, float a_col,
                       float s_row, float s_col, float weight, const float* upstream, void* dx, float* dscale,
                       void* workspace, int64_t workspace_bytes, int device, void* stream) {
  return cosmos_infonce_bwd_g(p, row_lse2, col_lse2, a_row, a_col, s_row, s_col, weight, upstream, dx, dscale, nullptr, 0, workspace,
                              workspace_bytes, device, stream);
}

int cosmos_infonce_bwd_g(const cosmos_infonce_problem* p, const float* row_lse2, const float* col_lse2, float a_row, float a_col,
                         float s_row, float s_col, float weight, const float* upstream, void* dx, float* dscale, void* g_out,
                         int64_t g_ld, void* workspace, int64_t workspace_bytes, int device, void* stream) {
  Dims d;
  int st = check_problem(p, &d);
  if (st != COSMOS_OK) return st;
  if (!row_lse2 || !col_lse2 || !upstream) return COSMOS_ERR_INVALID_ARGUMENT;
  if (dx == nullptr && dscale == nullptr) return COSMOS_OK;
  if (dx != nullptr && (reinterpret_cast<uintptr_t>(dx) & 15) != 0) return COSMOS_ERR_INVALID_ARGUMENT;
  if (workspace == nullptr || workspace_bytes < bwd_workspace(p, d)) return COSMOS_ERR_WORKSPACE;
  if (g_out != nullptr) {
    // only the dim-512 cluster kernel keeps the tiles in a shape that can be stored cheaply; 16-byte pieces must not straddle
    if (d.ks != 8 || dx == nullptr || (dbg_flags() & (8 | 128))) return COSMOS_ERR_UNSUPPORTED;
    if ((p->n_cols & 7) != 0 || (g_ld & 7) != 0 || g_ld < static_cast<int64_t>(p->gy) * p->n_cols ||
        (reinterpret_cast<uintptr_t>(g_out) & 15) != 0)
      return COSMOS_ERR_INVALID_ARGUMENT;
  }
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  CUtensorMap tmX, tmY, tmY64;
  const int bf = p->dtype == COSMOS_DTYPE_BF16;
  // CTA-pair kernel for dims whose 256-wide parts split evenly over two CTAs; COSMOS_B200_DBG=8 forces the 1-CTA kernel
  const bool pair = (d.ks == 2 || d.ks == 4 || d.ks == 8) && !(dbg_flags() & 8);
  {
    const int m1 = cb::make_stack_map(&tmX, reinterpret_cast<const void*>(p->x), bf, p->dim, p->n_rows, p->gx, cb::kFwdBM);
    const int m2 = cb::make_stack_map(&tmY, reinterpret_cast<const void*>(p->y), bf, p->dim, p->n_cols, p->gy, cb::kBwdBN);
    const int m3 = cb::make_stack_map(&tmY64, reinterpret_cast<const void*>(p->y), bf, p->dim, p->n_cols, p->gy, cb::kBwdBN / 2);
    if (m1 != 0 || m2 != 0 || m3 != 0) {
      g_last_cuda = 100000 + (m1 != 0 ? m1 : (m2 != 0 ? m2 : m3));   // 100000 + CUresult of cuTensorMapEncodeTiled
      return COSMOS_ERR_CUDA;
    }
  }
  cb::BwdParams bp;
  bp.gx = p->gx; bp.gy = p->gy; bp.n_rows = p->n_rows; bp.n_cols = p->n_cols; bp.ks = d.ks;
  bp.label_offset = p->label_offset;
  bp.n_row_tiles = d.n_row_tiles; bp.n_col_tiles = d.n_col_tiles_bwd;
  bp.n_parts = dx != nullptr ? d.n_parts : 1;
  bp.dtype = p->dtype;
  bp.dbg = dbg_flags();
  if (pair) {
    const int nh = d.ks < 4 ? d.ks : 4;
    bp.idesc_s = cb::make_idesc(bf, 0, 0, 2 * cb::kFwdBM, (dbg_flags() & 8192) ? 16 : cb::kBwdBN);
    bp.idesc_g = cb::make_idesc(bf, 0, 1, 2 * cb::kFwdBM, (dbg_flags() & 8192) ? 16 : nh * 64);
  } else {
    bp.idesc_s = cb::make_idesc(bf, 0, 0, cb::kFwdBM, cb::kBwdBN);
    bp.idesc_g = cb::make_idesc(bf, 0, 1, cb::kFwdBM, 64);
  }
  bp.a_row = a_row; bp.a_col = a_col; bp.s_row = s_row; bp.s_col = s_col; bp.weight = weight;
  bp.scale = reinterpret_cast<const float*>(p->scale);
  bp.upstream = upstream;
  bp.row_lse2 = row_lse2; bp.col_lse2 = col_lse2;
  bp.dx = dx;
  bp.dscale_part = dscale != nullptr ? reinterpret_cast<float*>(workspace) : nullptr;
  bp.t_splits = 1;
  bp.dx32 = nullptr;
  bp.g_out = g_out;
  bp.g_ld = g_ld;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (pair && dx != nullptr && !(dbg_flags() & 64)) {
    const int clusters = p->gx * ((d.n_row_tiles + 1) / 2) * bp.n_parts;
    (void)clusters;   // measured: splitting the sweep does not pay (per-item prologue + atomics eat the wave-tail gain)
    bp.t_splits = bwd_t_splits();   // diagnostics (COSMOS_B200_TSPLIT)
    if (bp.t_splits > 1) {
      bp.dx32 = reinterpret_cast<float*>(static_cast<char*>(workspace) + bwd_partials_bytes(p, d));
      if (cu_fail(cudaMemsetAsync(bp.dx32, 0, static_cast<size_t>(p->gx) * p->n_rows * p->dim * sizeof(float), s))) return COSMOS_ERR_CUDA;
    }
  }
  // dim 512 with dX wanted: 4-CTA clusters that share the softmax-gradient tiles between the two embedding parts
  const bool quad = pair && d.ks == 8 && dx != nullptr && !(dbg_flags() & 128);   // COSMOS_B200_DBG=128: pair kernel (diagnostics)
  int ds_slots = bp.t_splits;
  if (quad) {
    bp.n_parts = 2;
    ds_slots = 2;
    if (cu_fail(cb::launch_infonce_bwd_quad(tmX, tmY64, bp, s))) return COSMOS_ERR_CUDA;
  } else if (cu_fail(pair ? cb::launch_infonce_bwd_pair(tmX, tmY64, tmY, bp, s) : cb::launch_infonce_bwd(tmX, tmY, bp, s))) {
    return COSMOS_ERR_CUDA;
  }
  if (bp.t_splits > 1 &&
      cu_fail(cb::launch_convert_dx(bp.dx32, dx, p->dtype, static_cast<size_t>(p->gx) * p->n_rows * p->dim, s)))
    return COSMOS_ERR_CUDA;
  if (dscale != nullptr &&
      cu_fail(cb::launch_dscale_reduce(bp.dscale_part, p->gx * d.n_row_tiles * ds_slots, weight, upstream, dscale, s)))
    return COSMOS_ERR_CUDA;
  return COSMOS_OK;
}

int cosmos_infonce_bwd_e(const cosmos_infonce_problem* p, const void* e, const float* off, const float* diag_raw,
                         const float* row_lse2, const float* col_lse2, float a_row, float a_col, float s_row, float s_col,
                         float weight, const float* upstream, void* dx, float* dscale, void* g_out, int64_t g_ld, void* workspace,
                         int64_t workspace_bytes, int device, void* stream) {
  Dims d;
  int st = check_problem(p, &d);
  if (st != COSMOS_OK) return st;
  if (!e || !off || !diag_raw || !row_lse2 || !col_lse2 || !upstream || !dx) return COSMOS_ERR_INVALID_ARGUMENT;
  if (p->dim != 512) return COSMOS_ERR_UNSUPPORTED;     // the dX accumulator of 128 rows x 512 columns is the whole tensor memory
  // d(scale) is read off the dX accumulators, which weigh the two softmax terms like G: the mixes must be proportional
  if (dscale != nullptr && (fabsf(a_row * s_col - a_col * s_row) > 1e-12f || a_row + a_col == 0.f)) return COSMOS_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(e) & 15) != 0 || (reinterpret_cast<uintptr_t>(dx) & 15) != 0) return COSMOS_ERR_INVALID_ARGUMENT;
  if (g_out != nullptr && ((p->n_cols & 7) != 0 || (g_ld & 7) != 0 || g_ld < static_cast<int64_t>(p->gy) * p->n_cols ||
                           (reinterpret_cast<uintptr_t>(g_out) & 15) != 0))
    return COSMOS_ERR_INVALID_ARGUMENT;
  const int splits = bwd_e_splits(p, d, device);
  const int64_t part_bytes = splits > 1 ? static_cast<int64_t>(splits) * p->gx * p->n_rows * 512 * static_cast<int64_t>(sizeof(float)) : 0;
  if (workspace == nullptr || workspace_bytes < bwd_partials_bytes(p, d) + part_bytes) return COSMOS_ERR_WORKSPACE;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  CUtensorMap tmY64;
  const int bf = p->dtype == COSMOS_DTYPE_BF16;
  {
    // (E is bf16 whatever the stack dtype is - fp16 has no range for 2^(s2 - offset); the kernel converts G in registers)
    const int m2 = cb::make_stack_map(&tmY64, reinterpret_cast<const void*>(p->y), bf, p->dim, p->n_cols, p->gy, 64);
    if (m2 != 0) {
      g_last_cuda = 100000 + m2;
      return COSMOS_ERR_CUDA;
    }
  }
  cb::BwdEParams bp;
  bp.gx = p->gx; bp.gy = p->gy; bp.n_rows = p->n_rows; bp.n_cols = p->n_cols; bp.label_offset = p->label_offset;
  bp.n_row_tiles = d.n_row_tiles; bp.n_col_tiles = d.n_col_tiles_bwd;
  bp.n_chunks = (p->n_cols + 31) / 32;
  bp.dtype = p->dtype;
  bp.dbg = dbg_flags();
  bp.idesc_g = cb::make_idesc(bf, 0, 1, 2 * cb::kFwdBM, 256);
  bp.a_row = a_row; bp.a_col = a_col; bp.s_row = s_row; bp.s_col = s_col; bp.weight = weight;
  bp.scale = reinterpret_cast<const float*>(p->scale);
  bp.upstream = upstream;
  bp.e = e; bp.off = off; bp.row_lse2 = row_lse2; bp.col_lse2 = col_lse2; bp.diag_raw = diag_raw;
  bp.x = reinterpret_cast<const void*>(p->x);
  bp.dx = dx;
  bp.dscale_part = dscale != nullptr ? reinterpret_cast<float*>(workspace) : nullptr;
  bp.g_out = g_out;
  bp.g_ld = g_ld;
  bp.t_splits = splits;
  bp.dx32 = splits > 1 ? reinterpret_cast<float*>(static_cast<char*>(workspace) + bwd_partials_bytes(p, d)) : nullptr;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (cu_fail(cb::launch_infonce_bwd_e2(tmY64, bp, s))) return COSMOS_ERR_CUDA;
  if (splits > 1 && cu_fail(cb::launch_reduce_dx(bp.dx32, splits, dx, p->dtype, static_cast<size_t>(p->gx) * p->n_rows * 512, s)))
    return COSMOS_ERR_CUDA;
  // partial sums hold <G, raw> with G's mix; (s_row + s_col) / (a_row + a_col) turns it into the requested one
  if (dscale != nullptr && cu_fail(cb::launch_dscale_reduce(bp.dscale_part, splits * p->gx * d.n_row_tiles,
                                                            weight * (s_row + s_col) / (a_row + a_col), upstream, dscale, s)))
    return COSMOS_ERR_CUDA;
  return COSMOS_OK;
}

int32_t cosmos_infonce_bwd_e_cols_splits(const cosmos_infonce_problem* p, int device) {
  Dims d;
  if (check_problem(p, &d) != COSMOS_OK) return -1;
  // the fewest slices of the row sweep for which the CTA pairs (one per 256 columns of one column tensor) fill whole waves
  const int pairs = p->gy * ((d.n_col_tiles_bwd + 1) / 2);
  const int steps = p->gx * d.n_row_tiles;
  static const int forced = env_int("COSMOS_B200_COLS_SPLITS", 0);      // diagnostics
  if (forced > 0) return forced <= kMaxTSplits && forced <= steps ? forced : 1;
  return choose_t_splits(pairs, sm_count_of(device) / 2, steps);
}

int cosmos_infonce_bwd_e_cols(const cosmos_infonce_problem* p, const void* e, const float* off, const float* diag_raw,
                              const float* row_lse2, const float* col_lse2, float a_row, float a_col, float* dy, int32_t splits,
                              int device, void* stream) {
  Dims d;
  int st = check_problem(p, &d);
  if (st != COSMOS_OK) return st;
  if (!e || !off || !diag_raw || !row_lse2 || !col_lse2 || !dy) return COSMOS_ERR_INVALID_ARGUMENT;
  if (p->dim != 512) return COSMOS_ERR_UNSUPPORTED;
  if (splits < 1 || splits > kMaxTSplits || splits > p->gx * d.n_row_tiles) return COSMOS_ERR_INVALID_ARGUMENT;
  if ((reinterpret_cast<uintptr_t>(e) & 15) != 0 || (reinterpret_cast<uintptr_t>(dy) & 15) != 0) return COSMOS_ERR_INVALID_ARGUMENT;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  CUtensorMap tmX64;
  const int bf = p->dtype == COSMOS_DTYPE_BF16;
  {
    const int m = cb::make_stack_map(&tmX64, reinterpret_cast<const void*>(p->x), bf, p->dim, p->n_rows, p->gx, 64);
    if (m != 0) {
      g_last_cuda = 100000 + m;
      return COSMOS_ERR_CUDA;
    }
  }
  cb::BwdEParams bp = {};
  bp.gx = p->gx; bp.gy = p->gy; bp.n_rows = p->n_rows; bp.n_cols = p->n_cols; bp.label_offset = p->label_offset;
  bp.n_row_tiles = d.n_row_tiles; bp.n_col_tiles = d.n_col_tiles_bwd;
  bp.n_chunks = (p->n_cols + 31) / 32;
  bp.dtype = p->dtype;
  bp.dbg = dbg_flags();
  bp.idesc_g = cb::make_idesc(bf, 1, 1, 2 * cb::kFwdBM, 256);     // A = G^T and B = X both read with M / N contiguous
  bp.a_row = a_row; bp.a_col = a_col;
  bp.scale = reinterpret_cast<const float*>(p->scale);
  bp.e = e; bp.off = off; bp.row_lse2 = row_lse2; bp.col_lse2 = col_lse2; bp.diag_raw = diag_raw;
  bp.x = reinterpret_cast<const void*>(p->x);
  bp.dx = dy;
  bp.t_splits = splits;
  return cu_fail(cb::launch_infonce_bwd_e2t(tmX64, bp, static_cast<cudaStream_t>(stream))) ? COSMOS_ERR_CUDA : COSMOS_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// pooler building blocks
// ------------------------------------------------------------------------------------------------
namespace {
bool dtype16(int d) { return d == COSMOS_DTYPE_BF16 || d == COSMOS_DTYPE_F16; }
bool dtype_any(int d) { return d == COSMOS_DTYPE_F32 || dtype16(d); }
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
}  // namespace

extern "C" {

int cosmos_gemm(const void* a, const void* b, void* d, const float* bias, int32_t M, int32_t N, int32_t K, int64_t lda,
                int64_t ldb, int64_t ldd, int32_t a_kmajor, int32_t b_kmajor, int32_t in_dtype, int32_t out_dtype,
                int32_t splits, float alpha, int device, void* stream) {
  if (!a || !b || !d || M <= 0 || N <= 0 || K <= 0 || splits <= 0) return COSMOS_ERR_INVALID_ARGUMENT;
  if (!dtype16(in_dtype) || !dtype_any(out_dtype)) return COSMOS_ERR_UNSUPPORTED;
  if ((lda & 7) || (ldb & 7) || !aligned16(a) || !aligned16(b) || !aligned16(d)) return COSMOS_ERR_INVALID_ARGUMENT;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  cb::GemmArgs ga;
  ga.a = a; ga.b = b; ga.d = d; ga.bias = bias;
  ga.M = M; ga.N = N; ga.K = K; ga.lda = lda; ga.ldb = ldb; ga.ldd = ldd;
  ga.a_kmajor = a_kmajor; ga.b_kmajor = b_kmajor; ga.in_dtype = in_dtype; ga.out_dtype = out_dtype;
  ga.splits = splits; ga.alpha = alpha;
  cudaError_t e = cudaSuccess;
  int ctas = sm_count_of(device);
  if (const int c = gemm_cta_cap(); c > 0 && c < ctas) ctas = c;
  const int r = cb::launch_gemm(ga, ctas, static_cast<cudaStream_t>(stream), &e);
  if (r == 0) return COSMOS_OK;
  if (r > 0) g_last_cuda = r; else cu_fail(e);
  return COSMOS_ERR_CUDA;
}

int cosmos_gemm_ex(const cosmos_gemm_desc* q, int device, void* stream) {
  if (!q) return COSMOS_ERR_INVALID_ARGUMENT;
  const int batch_in = q->batch_in > 0 ? q->batch_in : 1;
  if (!q->a || !q->b || !q->d || q->M <= 0 || q->N <= 0 || q->K <= 0 || q->splits <= 0 || q->batch <= 0 || q->K2 < 0)
    return COSMOS_ERR_INVALID_ARGUMENT;
  if (!dtype16(q->in_dtype) || !dtype_any(q->out_dtype)) return COSMOS_ERR_UNSUPPORTED;
  if ((q->lda & 7) || (q->ldb & 7) || (q->stride_a & 7) || (q->stride_b & 7) || (q->stride_a_in & 7) || (q->stride_b_in & 7) ||
      !aligned16(q->a) || !aligned16(q->b) || !aligned16(q->d))
    return COSMOS_ERR_INVALID_ARGUMENT;
  if (q->K2 > 0 && (!q->a2 || !q->b2 || (q->lda2 & 7) || (q->ldb2 & 7) || (q->stride_a2 & 7) || (q->stride_b2 & 7) ||
                    (q->stride_a2_in & 7) || (q->stride_b2_in & 7) || !aligned16(q->a2) || !aligned16(q->b2)))
    return COSMOS_ERR_INVALID_ARGUMENT;
  if (q->accumulate && q->splits != 1) return COSMOS_ERR_INVALID_ARGUMENT;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  cb::GemmArgs ga;
  ga.a = q->a; ga.b = q->b; ga.d = q->d; ga.bias = q->bias;
  ga.M = q->M; ga.N = q->N; ga.K = q->K; ga.lda = q->lda; ga.ldb = q->ldb; ga.ldd = q->ldd;
  ga.a_kmajor = q->a_kmajor; ga.b_kmajor = q->b_kmajor; ga.in_dtype = q->in_dtype; ga.out_dtype = q->out_dtype;
  ga.splits = q->splits; ga.alpha = q->alpha;
  ga.batch = q->batch; ga.sa = q->stride_a; ga.sb = q->stride_b; ga.sd = q->stride_d; ga.sbias = q->stride_bias;
  ga.accumulate = q->accumulate;
  ga.a2 = q->a2; ga.b2 = q->b2; ga.K2 = q->K2; ga.lda2 = q->lda2; ga.ldb2 = q->ldb2; ga.sa2 = q->stride_a2; ga.sb2 = q->stride_b2;
  ga.batch_in = batch_in; ga.sa_in = q->stride_a_in; ga.sb_in = q->stride_b_in; ga.sd_in = q->stride_d_in;
  ga.sbias_in = q->stride_bias_in; ga.sa2_in = q->stride_a2_in; ga.sb2_in = q->stride_b2_in;
  cudaError_t e = cudaSuccess;
  int ctas = sm_count_of(device);
  if (const int c = gemm_cta_cap(); c > 0 && c < ctas) ctas = c;
  const int r = cb::launch_gemm(ga, ctas, static_cast<cudaStream_t>(stream), &e);
  if (r == 0) return COSMOS_OK;
  if (r > 0) g_last_cuda = r; else cu_fail(e);
  return COSMOS_ERR_CUDA;
}

int cosmos_gemm_batched(const void* a, const void* b, void* d, const float* bias, int32_t M, int32_t N, int32_t K, int64_t lda,
                        int64_t ldb, int64_t ldd, int32_t batch, int64_t stride_a, int64_t stride_b, int64_t stride_d,
                        int64_t stride_bias, int32_t a_kmajor, int32_t b_kmajor, int32_t in_dtype, int32_t out_dtype, int32_t splits,
                        int32_t accumulate, float alpha, const void* a2, const void* b2, int32_t K2, int64_t lda2, int64_t ldb2,
                        int64_t stride_a2, int64_t stride_b2, int device, void* stream) {
  cosmos_gemm_desc q = {};
  q.a = a; q.b = b; q.d = d; q.bias = bias; q.a2 = a2; q.b2 = b2;
  q.M = M; q.N = N; q.K = K; q.K2 = K2;
  q.lda = lda; q.ldb = ldb; q.ldd = ldd; q.lda2 = lda2; q.ldb2 = ldb2;
  q.batch = batch; q.batch_in = 1;
  q.stride_a = stride_a; q.stride_b = stride_b; q.stride_d = stride_d; q.stride_bias = stride_bias;
  q.stride_a2 = stride_a2; q.stride_b2 = stride_b2;
  q.a_kmajor = a_kmajor; q.b_kmajor = b_kmajor; q.in_dtype = in_dtype; q.out_dtype = out_dtype; q.splits = splits;
  q.accumulate = accumulate; q.alpha = alpha;
  return cosmos_gemm_ex(&q, device, stream);
}

int cosmos_colsoftmax_fwd(const float* s, int64_t s_stride, int32_t lds, void* p, int64_t p_stride, int32_t ldp, int32_t p_dtype,
                          int32_t n_sets, int32_t L, int32_t n_cols, int32_t zero_key, int device, void* stream) {
  if (!s || !p || n_sets < 0 || L <= 0 || n_cols <= 0 || lds < n_cols || ldp < n_cols) return COSMOS_ERR_INVALID_ARGUMENT;
  if (!dtype16(p_dtype)) return COSMOS_ERR_UNSUPPORTED;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  return cu_fail(cb::launch_colsoftmax_fwd(s, s_stride, lds, p, p_stride, ldp, p_dtype, n_sets, L, n_cols, zero_key != 0,
                                           static_cast<cudaStream_t>(stream)))
             ? COSMOS_ERR_CUDA : COSMOS_OK;
}

int cosmos_colsoftmax_bwd(const void* p, int64_t p_stride, int32_t ldp, const float* dp, int64_t dp_stride, int32_t lddp, void* ds,
                          int64_t ds_stride, int32_t ldds, int32_t dtype, int32_t n_sets, int32_t L, int32_t n_cols, int device,
                          void* stream) {
  if (!p || !dp || !ds || n_sets < 0 || L <= 0 || n_cols <= 0 || ldp < n_cols || lddp < n_cols || ldds < n_cols)
    return COSMOS_ERR_INVALID_ARGUMENT;
  if (!dtype16(dtype)) return COSMOS_ERR_UNSUPPORTED;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  return cu_fail(cb::launch_colsoftmax_bwd(p, p_stride, ldp, dp, dp_stride, lddp, ds, ds_stride, ldds, dtype, n_sets, L, n_cols,
                                           static_cast<cudaStream_t>(stream)))
             ? COSMOS_ERR_CUDA : COSMOS_OK;
}

int cosmos_layernorm_fwd(const void* x, int32_t x_dtype, const float* w, const float* b, void* y, int32_t y_dtype, float* mean,
                         float* rstd, int64_t rows, int32_t dim, float eps, int device, void* stream) {
  if (!x || !w || !b || !y || !mean || !rstd || rows < 0 || dim <= 0 || !(eps >= 0.f)) return COSMOS_ERR_INVALID_ARGUMENT;
  if (dim > 1024 || !dtype_any(x_dtype) || !dtype_any(y_dtype)) return COSMOS_ERR_UNSUPPORTED;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  return cu_fail(cb::launch_layernorm_fwd(x, x_dtype, w, b, y, y_dtype, mean, rstd, rows, dim, eps, static_cast<cudaStream_t>(stream)))
             ? COSMOS_ERR_CUDA : COSMOS_OK;
}

int cosmos_layernorm_bwd(const void* dy, int32_t dy_dtype, const void* x, int32_t x_dtype, const float* w, const float* mean,
                         const float* rstd, void* dx, int32_t dx_dtype, int32_t accumulate, float* dw, float* db, int64_t rows,
                         int32_t dim, int device, void* stream) {
  if (!dy || !x || !w || !mean || !rstd || !dx || rows < 0 || dim <= 0) return COSMOS_ERR_INVALID_ARGUMENT;
  if ((dw == nullptr) != (db == nullptr)) return COSMOS_ERR_INVALID_ARGUMENT;
  if (dim > 1024 || !dtype_any(dy_dtype) || !dtype_any(x_dtype) || !dtype_any(dx_dtype)) return COSMOS_ERR_UNSUPPORTED;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  return cu_fail(cb::launch_layernorm_bwd(dy, dy_dtype, x, x_dtype, w, mean, rstd, dx, dx_dtype, accumulate, dw, db, rows, dim,
                                          static_cast<cudaStream_t>(stream)))
             ? COSMOS_ERR_CUDA : COSMOS_OK;
}

static int check_attn(int32_t dtype, int32_t n_sets, int32_t L, int32_t dim, int32_t heads, int32_t q_per_set) {
  if (n_sets <= 0 || L <= 0 || dim <= 0 || heads <= 0 || q_per_set <= 0) return COSMOS_ERR_INVALID_ARGUMENT;
  if (!dtype16(dtype) || dim % heads != 0) return COSMOS_ERR_UNSUPPORTED;
  const int hd = dim / heads;
  if (!(hd == 16 || hd == 32 || hd == 64 || hd == 128) || L > 1024 || q_per_set > 4096) return COSMOS_ERR_UNSUPPORTED;
  return COSMOS_OK;
}

int cosmos_attn_core_fwd(const void* q, const void* kv, void* o, float* lse, int32_t dtype, int32_t n_sets, int32_t L, int32_t dim,
                         int32_t heads, int32_t q_per_set, int64_t q_stride_set, int64_t q_stride_q, int device, void* stream) {
  if (!q || !kv || !o || !lse) return COSMOS_ERR_INVALID_ARGUMENT;
  int st = check_attn(dtype, n_sets, L, dim, heads, q_per_set);
  if (st != COSMOS_OK) return st;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  return cu_fail(cb::launch_attn_core_fwd(q, kv, o, lse, dtype, n_sets, L, dim, heads, q_per_set, q_stride_set, q_stride_q,
                                          static_cast<cudaStream_t>(stream)))
             ? COSMOS_ERR_CUDA : COSMOS_OK;
}

int cosmos_attn_core_bwd(const void* q, const void* kv, const void* d_o, const float* lse, void* dq, void* dkv, int32_t dtype,
                         int32_t n_sets, int32_t L, int32_t dim, int32_t heads, int32_t q_per_set, int64_t q_stride_set,
                         int64_t q_stride_q, int device, void* stream) {
  if (!q || !kv || !d_o || !lse || !dq || !dkv) return COSMOS_ERR_INVALID_ARGUMENT;
  int st = check_attn(dtype, n_sets, L, dim, heads, q_per_set);
  if (st != COSMOS_OK) return st;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  return cu_fail(cb::launch_attn_core_bwd(q, kv, d_o, lse, dq, dkv, dtype, n_sets, L, dim, heads, q_per_set, q_stride_set,
                                          q_stride_q, static_cast<cudaStream_t>(stream)))
             ? COSMOS_ERR_CUDA : COSMOS_OK;
}

int cosmos_addnorm_fwd(const void* f, int32_t f_dtype, const float* pooled, void* out, float* inv_norm, int64_t rows, int32_t dim,
                       int device, void* stream) {
  if (!f || !pooled || !out || !inv_norm || rows < 0 || dim <= 0) return COSMOS_ERR_INVALID_ARGUMENT;
  if (dim > 1024 || !dtype_any(f_dtype)) return COSMOS_ERR_UNSUPPORTED;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  return cu_fail(cb::launch_addnorm_fwd(f, f_dtype, pooled, out, inv_norm, rows, dim, static_cast<cudaStream_t>(stream)))
             ? COSMOS_ERR_CUDA : COSMOS_OK;
}

int cosmos_addnorm_bwd(const void* g_out, const void* out, int32_t f_dtype, const float* inv_norm, float* g_z32, void* g_z16,
                       int32_t g_dtype, int64_t rows, int32_t dim, int device, void* stream) {
  if (!g_out || !out || !inv_norm || !g_z32 || !g_z16 || rows < 0 || dim <= 0) return COSMOS_ERR_INVALID_ARGUMENT;
  if (dim > 1024 || !dtype_any(f_dtype) || !dtype16(g_dtype)) return COSMOS_ERR_UNSUPPORTED;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  return cu_fail(cb::launch_addnorm_bwd(g_out, out, f_dtype, inv_norm, g_z32, g_z16, g_dtype, rows, dim,
                                        static_cast<cudaStream_t>(stream)))
             ? COSMOS_ERR_CUDA : COSMOS_OK;
}

int cosmos_colsum(const void* src, int32_t dtype, float* dst, int64_t rows, int32_t n, int64_t ld, int device, void* stream) {
  if (!src || !dst || rows < 0 || n <= 0) return COSMOS_ERR_INVALID_ARGUMENT;
  if (!dtype_any(dtype)) return COSMOS_ERR_UNSUPPORTED;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  return cu_fail(cb::launch_colsum(src, dtype, dst, rows, n, ld, static_cast<cudaStream_t>(stream))) ? COSMOS_ERR_CUDA : COSMOS_OK;
}

int cosmos_retrieval_ranks(const void* q, const void* gal, int dtype, int32_t M, int32_t N, int32_t D, int64_t ldq, int64_t ldg,
                           const int32_t* gt_offsets, const int32_t* gt_index, float* best, int32_t* best_col, int32_t* ranks,
                           int device, void* stream) {
  if (!q || !gal || !best || !best_col || !ranks || M <= 0 || N <= 0 || D <= 0 || ldq < D || ldg < D) return COSMOS_ERR_INVALID_ARGUMENT;
  if (gt_index != nullptr && gt_offsets == nullptr) return COSMOS_ERR_INVALID_ARGUMENT;
  if (!dtype_any(dtype)) return COSMOS_ERR_UNSUPPORTED;
  if (M > 65535 * 128) return COSMOS_ERR_UNSUPPORTED;
  DeviceGuard g(device);
  if (!g.ok) return COSMOS_ERR_CUDA;
  return cu_fail(cb::launch_retrieval_ranks(q, gal, dtype, M, N, D, ldq, ldg, gt_offsets, gt_index, best, best_col, ranks,
                                            static_cast<cudaStream_t>(stream)))
             ? COSMOS_ERR_CUDA : COSMOS_OK;
}

}  // extern "C"
