// extern "C" entry points of liblist_b200.so (declared in include/list_b200.h).
// Argument validation, the row layout, chunk loops; the kernels live in the other .cu files.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "hoist.cuh"
#include "tgemm.cuh"

namespace list {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return LIST_ECUDA;
}

// kernels / launchers defined in the other translation units
int prep_maps(const float* const* maps, const int32_t* ch, const int32_t* size, int n_maps, int B, int S, void* out,
              int dtype, cudaStream_t st);
int prep_volume(const float* in, int B, int C, int R, void* out, int dtype, cudaStream_t st);
int prep_volume_bwd(const float* g, int B, int C, int R, float* out, cudaStream_t st);
int prep_maps_bwd(const float* g, const int32_t* ch, const int32_t* size, int n_maps, int B, int S, float* const* outs, cudaStream_t st);
int gather_fwd(const ListCtx* ctx, const float* q, int q_is_raw, void* X, int64_t ldx, int B, int64_t N, cudaStream_t st);
int gather_grid_fwd(const ListCtx* ctx, int image, int res, double bb_min, double bb_max, int64_t begin, int64_t count,
                    void* X, int64_t ldx, cudaStream_t st);
int gather_grid_walk(const ListCtx* ctx, int image, int res, double bb_min, double bb_max, int64_t begin, int64_t count,
                     void* X, int64_t ldx, cudaStream_t st);
int grid_points(float* q, int res, double lo, double hi, int64_t begin, int64_t count, cudaStream_t st);
size_t mlp_f32_workspace_bytes(const ListWeights* w, int64_t rows);
int mlp_f32_fwd(const ListWeights* w, const float* X, int64_t ldx, int64_t rows, float* sdf, float out_div, float* ws,
                int exact, cudaStream_t st);
size_t mlp_f32_bwd_workspace_bytes(const ListWeights* w, int64_t rows);
int mlp_f32_bwd(const ListWeights* w, const float* X, int64_t ldx, int64_t rows, const float* fwd_ws, const float* d_sdf,
                const ListGrads* g, float* ws, cudaStream_t st);
int gather_bwd(const ListCtx* ctx, const float* q, int q_is_raw, int B, int64_t N, const float* dX, int64_t ldd,
               const ListGrads* g, cudaStream_t st);
int mlp_tc_fwd(const ListWeights* w, const void* X, int64_t ldx, int64_t rows, float* sdf, float out_div, int variant,
               float* dbg1, float* dbg2, float* dbg3, cudaStream_t st);
int mlp_tc_fwd_hoisted(const ListWeights* w, int col0, int k, const void* Xh, int64_t ldx, int64_t rows, float* sdf,
                       float out_div, int variant, float* dbg1, float* dbg2, float* dbg3, long long* trace, cudaStream_t st);

size_t mlp_f32_workspace_bytes(const ListWeights* w, int64_t rows);
static size_t elem_size(int dtype) { return dtype == LIST_BF16 ? 2 : 4; }
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int check_ctx(const ListCtx* c) {
  LIST_CHECK_ARG(c != nullptr, "ctx is NULL");
  LIST_CHECK_ARG(c->dtype == LIST_F32 || c->dtype == LIST_BF16, "ctx.dtype %d is not LIST_F32/LIST_BF16", c->dtype);
  LIST_CHECK_ARG(c->B >= 1, "ctx.B %d < 1", c->B);
  LIST_CHECK_ARG(c->map_size >= 2, "ctx.map_size %d < 2", c->map_size);
  LIST_CHECK_ARG(c->map_channels > 0 && c->map_channels % 8 == 0, "ctx.map_channels %d must be a positive multiple of 8",
                 c->map_channels);
  LIST_CHECK_ARG(c->maps != nullptr && (reinterpret_cast<uintptr_t>(c->maps) & 15) == 0, "ctx.maps NULL or not 16B aligned");
  LIST_CHECK_ARG(c->n_levels >= 1 && c->n_levels <= LIST_MAX_LEVELS, "ctx.n_levels %d out of range", c->n_levels);
  for (int l = 0; l < c->n_levels; ++l) {
    LIST_CHECK_ARG(c->vol_res[l] >= 1, "ctx.vol_res[%d]=%d < 1", l, c->vol_res[l]);
    LIST_CHECK_ARG(c->vol_ch[l] >= 1 && c->vol_ch[l] <= 128, "ctx.vol_ch[%d]=%d not in [1,128]", l, c->vol_ch[l]);
    LIST_CHECK_ARG(c->vols[l] != nullptr && (reinterpret_cast<uintptr_t>(c->vols[l]) & 15) == 0,
                   "ctx.vols[%d] NULL or not 16B aligned", l);
  }
  LIST_CHECK_ARG(c->trans_mat != nullptr, "ctx.trans_mat is NULL");
  return LIST_OK;
}

static int check_weights(const ListWeights* w, int k_pad) {
  LIST_CHECK_ARG(w != nullptr, "weights is NULL");
  LIST_CHECK_ARG(w->dtype == LIST_F32 || w->dtype == LIST_BF16, "weights.dtype %d invalid", w->dtype);
  LIST_CHECK_ARG(k_pad < 0 || w->k_pad == k_pad, "weights.k_pad %d does not match the feature layout (%d)", w->k_pad, k_pad);
  LIST_CHECK_ARG(w->n0 > 0 && w->n1 > 0 && w->n2 > 0 && w->n0 % 4 == 0 && w->n1 % 4 == 0 && w->n2 % 4 == 0,
                 "weights layer widths %d/%d/%d must be positive multiples of 4", w->n0, w->n1, w->n2);
  LIST_CHECK_ARG(w->w0 && w->w1 && w->w2 && w->w3 && w->b0 && w->b1 && w->b2 && w->b3, "a weights pointer is NULL");
  LIST_CHECK_ARG(((reinterpret_cast<uintptr_t>(w->w0) | reinterpret_cast<uintptr_t>(w->w1) |
                   reinterpret_cast<uintptr_t>(w->w2)) & 15) == 0, "weight matrices must be 16B aligned");
  return LIST_OK;
}

// Dense-grid gather: the z-run walker kernel (gather_grid.cu); configurations it does not cover, or
// LIST_B200_GRID_GENERIC=1 (A/B aid), use the generic per-point kernel in grid mode.
static int grid_gather(const ListCtx* ctx, int image, int res, double bb_min, double bb_max, int64_t begin, int64_t count,
                       void* X, int64_t ldx, cudaStream_t st) {
  const char* e = getenv("LIST_B200_GRID_GENERIC");
  if (!(e && e[0] == '1')) {
    const int rc = gather_grid_walk(ctx, image, res, bb_min, bb_max, begin, count, X, ldx, st);
    if (rc != LIST_ENOSYS) return rc;
  }
  return gather_grid_fwd(ctx, image, res, bb_min, bb_max, begin, count, X, ldx, st);
}

static int mlp_variant() {
  // LIST_B200_MLP_VARIANT=1 selects the single-CTA tcgen05 kernel (bring-up aid), 3 the CTA pair with a
  // 3-stage operand ring; default CTA pair with 4 stages.
  const char* e = getenv("LIST_B200_MLP_VARIANT");
  if (e && e[0] == '1') return 1;
  if (e && e[0] == '3') return 3;
  return 2;
}

// ---- two-stream chunk pipeline ------------------------------------------------------------
// The gather is issue/HBM-write bound and the tensor-core MLP is tensor-pipe bound, so the gather of
// chunk i+1 runs CONCURRENTLY with the MLP of chunk i: the gather is enqueued on a low-priority
// auxiliary stream, the MLP on a high-priority one, X is double buffered and events order them.  Both
// auxiliary streams are forked from / joined to the caller's stream with events, so the caller still
// sees plain stream semantics (everything enqueued, nothing synchronised, graph-capturable).
// The streams and events are created lazily once per (host thread, device): thread_local keeps the
// entry points re-entrant (nn.DataParallel replicas call from parallel host threads).
struct Pipe {
  int device = -1;
  cudaStream_t lo = nullptr, hi = nullptr;
  cudaEvent_t fork = nullptr, ready[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
  // host-buffer entry point (list_sdf_grid_host): copy stream, "coarse tensors uploaded", "all tensors uploaded",
  // "these SDF values are final", "last download enqueued"
  cudaStream_t cp = nullptr;
  cudaEvent_t cfork = nullptr, small = nullptr, big = nullptr, item = nullptr, cdone = nullptr;
};
static constexpr int kMaxPipeDevices = 16;
static thread_local Pipe g_pipes[kMaxPipeDevices];

static int get_pipe(Pipe** out) {
  int dev = 0;
  LIST_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxPipeDevices) { set_error("device %d outside the pipeline table", dev); return LIST_ENOSYS; }
  Pipe& p = g_pipes[dev];
  if (p.device != dev) {
    int least = 0, greatest = 0;
    LIST_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    LIST_CUDA(cudaStreamCreateWithPriority(&p.lo, cudaStreamNonBlocking, least));
    LIST_CUDA(cudaStreamCreateWithPriority(&p.hi, cudaStreamNonBlocking, greatest));
    LIST_CUDA(cudaStreamCreateWithFlags(&p.cp, cudaStreamNonBlocking));
    LIST_CUDA(cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming));
    for (cudaEvent_t* e : {&p.cfork, &p.small, &p.big, &p.item, &p.cdone}) LIST_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
      LIST_CUDA(cudaEventCreateWithFlags(&p.ready[i], cudaEventDisableTiming));
      LIST_CUDA(cudaEventCreateWithFlags(&p.freed[i], cudaEventDisableTiming));
    }
    p.device = dev;
  }
  *out = &p;
  return LIST_OK;
}

// Hoisted fc_0 (hoist.cu) is the default dense-grid path in bf16 mode; LIST_B200_HOIST=0 selects the
// plain full-row gather + MLP (A/B aid).
static bool hoist_enabled() {
  const char* e = getenv("LIST_B200_HOIST");
  return !(e && e[0] == '0');
}

// Line-table path (lines.cu + grid_tc.cu): the hoisted terms are interpolated on the tensor cores inside the MLP kernel.
// Default for bf16 dense grids; LIST_B200_LINES=0 selects the round-1 addend-kernel path (A/B aid).
static bool lines_enabled() {
  const char* e = getenv("LIST_B200_LINES");
  return !(e && e[0] == '0');
}
constexpr int kLinesLevels = 3, kLinesMaxRes = 32;     // hoisted levels of the line-table path

static bool overlap_enabled() {
  const char* e = getenv("LIST_B200_OVERLAP");
  return !(e && e[0] == '0');
}

// Hooks of the host-buffer entry point into the dense-grid evaluation.  Upload, evaluation and download overlap:
//   * the coarse per-image tensors (maps, levels with R <= 32, T) are uploaded and prepared first; the projection and
//     the addend part of the first chunk's gather only need those, so they run while the big volumes are still on the
//     wire.  `late` (wait for the upload, prepare the remaining levels) runs on the gather stream right before the
//     first kernel that reads them;
//   * every chunk's SDF values go to the host on the copy stream as soon as their MLP launch is done.
struct GridHooks {
  cudaEvent_t uploaded = nullptr;                   // recorded on the copy stream after the last H2D copy
  int (*late)(void* user, cudaStream_t s) = nullptr;
  void* user = nullptr;
  float* sdf_host = nullptr;                        // [B, count] like the device-side sdf; NULL: no download
  cudaStream_t copy = nullptr;
  cudaEvent_t item = nullptr;
};
static int hook_late(const GridHooks* h, cudaStream_t s) {
  if (!h || !h->late) return LIST_OK;
  if (h->uploaded) LIST_CUDA(cudaStreamWaitEvent(s, h->uploaded, 0));
  return h->late(h->user, s);
}
// values [off, off + n) of the SDF buffer are final once the work enqueued on `s` so far is done
static int hook_download(const GridHooks* h, const float* sdf_dev, int64_t off, int64_t n, cudaStream_t s) {
  if (!h || !h->sdf_host || n <= 0) return LIST_OK;
  LIST_CUDA(cudaEventRecord(h->item, s));
  LIST_CUDA(cudaStreamWaitEvent(h->copy, h->item, 0));
  LIST_CUDA(cudaMemcpyAsync(h->sdf_host + off, sdf_dev + off, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost, h->copy));
  return LIST_OK;
}

// Runs `n_items` (gather_i -> mlp_i) pairs.  gather(i, X, stream) fills X; mlp(i, X, stream) consumes it.
// xbuf[0], xbuf[1]: two feature-row buffers; with overlap == false only xbuf[0] is used, serially on `st`.
// pre (optional part of gather(i) that may run early): with `pre_ahead` the pre-stage of item 1 is enqueued right after
// item 0's, before gather(0) -- its buffer is free at the start, and whatever gather(0) has to wait for (tensors that are
// still being uploaded) then has two items' worth of independent work in front of it.
template <typename P, typename G, typename M>
static int run_chunks(int64_t n_items, void* const xbuf[2], bool overlap, cudaStream_t st, bool pre_ahead, P pre, G gather, M mlp) {
  int rc;
  if (!overlap || n_items < 2) {
    for (int64_t i = 0; i < n_items; ++i) {
      if ((rc = pre(i, xbuf[0], st))) return rc;
      if ((rc = gather(i, xbuf[0], st))) return rc;
      if ((rc = mlp(i, xbuf[0], st))) return rc;
    }
    return LIST_OK;
  }
  Pipe* p = nullptr;
  if ((rc = get_pipe(&p))) return rc;
  LIST_CUDA(cudaEventRecord(p->fork, st));
  LIST_CUDA(cudaStreamWaitEvent(p->lo, p->fork, 0));
  LIST_CUDA(cudaStreamWaitEvent(p->hi, p->fork, 0));
  for (int64_t i = 0; i < n_items; ++i) {
    const int b = static_cast<int>(i & 1);
    if (i >= 2) LIST_CUDA(cudaStreamWaitEvent(p->lo, p->freed[b], 0));   // MLP of item i-2 has read xbuf[b]
    if (!(i == 1 && pre_ahead) && (rc = pre(i, xbuf[b], p->lo))) return rc;
    if (i == 0 && pre_ahead && (rc = pre(1, xbuf[1], p->lo))) return rc;
    if ((rc = gather(i, xbuf[b], p->lo))) return rc;
    LIST_CUDA(cudaEventRecord(p->ready[b], p->lo));
    LIST_CUDA(cudaStreamWaitEvent(p->hi, p->ready[b], 0));
    if ((rc = mlp(i, xbuf[b], p->hi))) return rc;
    LIST_CUDA(cudaEventRecord(p->freed[b], p->hi));
  }
  // join: every gather is ordered before an MLP on `hi`; its last event joins the caller's stream
  LIST_CUDA(cudaStreamWaitEvent(st, p->freed[(n_items - 1) & 1], 0));
  return LIST_OK;
}

}  // namespace list

using namespace list;

extern "C" {

int list_b200_abi_version(void) { return LIST_B200_ABI_VERSION; }
const char* list_b200_last_error(void) { return g_err; }

int list_b200_device_ok(void) {
  int dev = 0;
  LIST_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  LIST_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; list_b200 is built for sm_100a (B200) only", dev, prop.major, prop.minor);
    return LIST_ENOSYS;
  }
  return LIST_OK;
}

int list_feature_layout(int32_t map_channels, int32_t n_levels, const int32_t* vol_ch, ListLayout* layout, int32_t* perm) {
  LIST_CHECK_ARG(layout != nullptr && vol_ch != nullptr, "layout/vol_ch is NULL");
  LIST_CHECK_ARG(n_levels >= 1 && n_levels <= LIST_MAX_LEVELS, "n_levels %d out of range", n_levels);
  LIST_CHECK_ARG(map_channels > 0 && map_channels % 8 == 0, "map_channels %d must be a positive multiple of 8", map_channels);
  int sum_c = 0;
  for (int l = 0; l < n_levels; ++l) {
    LIST_CHECK_ARG(vol_ch[l] >= 1, "vol_ch[%d]=%d < 1", l, vol_ch[l]);
    sum_c += vol_ch[l];
  }
  memset(layout, 0, sizeof(*layout));
  int col = 0;
  layout->map_off = col;
  col += map_channels;
  // vector levels (C % 8 == 0) first, widest-index first, then the scalar levels, then q
  for (int l = n_levels - 1; l >= 0; --l)
    if (vol_ch[l] % 8 == 0) { layout->vol_off[l] = col; col += LIST_NUM_DISP * vol_ch[l]; }
  for (int l = n_levels - 1; l >= 0; --l)
    if (vol_ch[l] % 8 != 0) { layout->vol_off[l] = col; col += LIST_NUM_DISP * vol_ch[l]; }
  layout->xyz_off = col;
  col += 3;
  layout->k_out = col;
  layout->k_pad = (col + 63) / 64 * 64;
  if (perm) {
    // reference column of (level l, channel c, displacement d) = (cum_c[l] + c)*7 + d  (modules.py:270-273);
    // percep follows at 7*sum_c, q at 7*sum_c + map_channels (modules.py:275).
    int cum = 0;
    for (int l = 0; l < n_levels; ++l) {
      for (int d = 0; d < LIST_NUM_DISP; ++d)
        for (int c = 0; c < vol_ch[l]; ++c) perm[layout->vol_off[l] + d * vol_ch[l] + c] = (cum + c) * LIST_NUM_DISP + d;
      cum += vol_ch[l];
    }
    for (int c = 0; c < map_channels; ++c) perm[layout->map_off + c] = LIST_NUM_DISP * sum_c + c;
    for (int j = 0; j < 3; ++j) perm[layout->xyz_off + j] = LIST_NUM_DISP * sum_c + map_channels + j;
  }
  return LIST_OK;
}

int list_prep_maps(const float* const* maps_nchw, const int32_t* ch, const int32_t* size, int32_t n_maps, int32_t B,
                   int32_t map_size, void* out, int32_t dtype, void* stream) {
  LIST_CHECK_ARG(maps_nchw && ch && size && out, "list_prep_maps: NULL argument");
  LIST_CHECK_ARG(n_maps >= 1 && n_maps <= LIST_MAX_MAPS, "list_prep_maps: n_maps %d out of range", n_maps);
  LIST_CHECK_ARG(dtype == LIST_F32 || dtype == LIST_BF16, "list_prep_maps: bad dtype %d", dtype);
  LIST_CHECK_ARG(B >= 1 && map_size >= 2, "list_prep_maps: B=%d map_size=%d", B, map_size);
  for (int i = 0; i < n_maps; ++i) {
    LIST_CHECK_ARG(maps_nchw[i] != nullptr && ch[i] >= 1 && size[i] >= 1 && size[i] <= 1024,
                   "list_prep_maps: map %d invalid (C=%d, size=%d)", i, ch[i], size[i]);
  }
  return prep_maps(maps_nchw, ch, size, n_maps, B, map_size, out, dtype, static_cast<cudaStream_t>(stream));
}

int list_prep_volume(const float* vol, int32_t B, int32_t C, int32_t R, void* out, int32_t dtype, void* stream) {
  LIST_CHECK_ARG(vol && out, "list_prep_volume: NULL argument");
  LIST_CHECK_ARG(dtype == LIST_F32 || dtype == LIST_BF16, "list_prep_volume: bad dtype %d", dtype);
  LIST_CHECK_ARG(B >= 1 && C >= 1 && C <= 128 && R >= 1, "list_prep_volume: B=%d C=%d R=%d", B, C, R);
  return prep_volume(vol, B, C, R, out, dtype, static_cast<cudaStream_t>(stream));
}

int list_prep_maps_bwd(const float* g, const int32_t* ch, const int32_t* size, int32_t n_maps, int32_t B, int32_t map_size,
                       float* const* grads_nchw, void* stream) {
  LIST_CHECK_ARG(g && ch && size && grads_nchw, "list_prep_maps_bwd: NULL argument");
  LIST_CHECK_ARG(n_maps >= 1 && n_maps <= LIST_MAX_MAPS, "list_prep_maps_bwd: n_maps %d out of range", n_maps);
  LIST_CHECK_ARG(B >= 1 && map_size >= 2 && map_size <= 256, "list_prep_maps_bwd: B=%d map_size=%d (2..256)", B, map_size);
  for (int i = 0; i < n_maps; ++i)
    LIST_CHECK_ARG(grads_nchw[i] && ch[i] >= 1 && size[i] >= 1 && size[i] <= 1024 && static_cast<int64_t>(B) * ((ch[i] + 31) / 32) < 65536,
                   "list_prep_maps_bwd: map %d invalid (C=%d, size=%d)", i, ch[i], size[i]);
  return prep_maps_bwd(g, ch, size, n_maps, B, map_size, grads_nchw, static_cast<cudaStream_t>(stream));
}

int list_prep_volume_bwd(const float* g, int32_t B, int32_t C, int32_t R, float* grad_ncdhw, void* stream) {
  LIST_CHECK_ARG(g && grad_ncdhw, "list_prep_volume_bwd: NULL argument");
  LIST_CHECK_ARG(B >= 1 && B < 65536 && C >= 1 && C <= 128 && R >= 1, "list_prep_volume_bwd: B=%d C=%d R=%d", B, C, R);
  return prep_volume_bwd(g, B, C, R, grad_ncdhw, static_cast<cudaStream_t>(stream));
}

int list_grid_points(float* q, int32_t res, double bb_min, double bb_max, int64_t begin, int64_t count, void* stream) {
  LIST_CHECK_ARG(q != nullptr || count == 0, "list_grid_points: q is NULL");
  LIST_CHECK_ARG(res >= 1 && res <= 2048, "list_grid_points: res %d out of range", res);
  const int64_t total = static_cast<int64_t>(res) * res * res;
  LIST_CHECK_ARG(begin >= 0 && count >= 0 && begin + count <= total, "list_grid_points: [%lld,+%lld) outside res^3",
                 (long long)begin, (long long)count);
  return grid_points(q, res, bb_min, bb_max, begin, count, static_cast<cudaStream_t>(stream));
}

int list_gather_fwd(const ListCtx* ctx, const float* q, int32_t q_is_raw, void* X, int64_t ldx, int32_t B, int64_t N,
                    void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  LIST_CHECK_ARG(B == ctx->B, "list_gather_fwd: B %d != ctx.B %d", B, ctx->B);
  LIST_CHECK_ARG(N >= 0, "list_gather_fwd: N < 0");
  if (N == 0) return LIST_OK;
  ListLayout lay;
  if ((rc = list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr))) return rc;
  LIST_CHECK_ARG(q != nullptr && X != nullptr, "list_gather_fwd: q/X is NULL");
  LIST_CHECK_ARG(ldx >= lay.k_pad && ldx % 8 == 0, "list_gather_fwd: ldx %lld must be >= %d and a multiple of 8", (long long)ldx, lay.k_pad);
  LIST_CHECK_ARG((reinterpret_cast<uintptr_t>(X) & 15) == 0, "list_gather_fwd: X not 16B aligned");
  return gather_fwd(ctx, q, q_is_raw, X, ldx, B, N, static_cast<cudaStream_t>(stream));
}

int list_gather_grid_fwd(const ListCtx* ctx, int32_t image, int32_t res, double bb_min, double bb_max, int64_t begin,
                         int64_t count, void* X, int64_t ldx, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  LIST_CHECK_ARG(image >= 0 && image < ctx->B, "list_gather_grid_fwd: image %d outside [0,%d)", image, ctx->B);
  LIST_CHECK_ARG(res >= 1 && res <= 2048, "list_gather_grid_fwd: res %d out of range", res);
  const int64_t total = static_cast<int64_t>(res) * res * res;
  LIST_CHECK_ARG(begin >= 0 && count >= 0 && begin + count <= total, "list_gather_grid_fwd: [%lld,+%lld) outside res^3",
                 (long long)begin, (long long)count);
  if (count == 0) return LIST_OK;
  ListLayout lay;
  if ((rc = list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr))) return rc;
  LIST_CHECK_ARG(X != nullptr && ldx >= lay.k_pad && ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0,
                 "list_gather_grid_fwd: X NULL/unaligned or ldx %lld < %d", (long long)ldx, lay.k_pad);
  return grid_gather(ctx, image, res, bb_min, bb_max, begin, count, X, ldx, static_cast<cudaStream_t>(stream));
}

size_t list_mlp_workspace_bytes(const ListWeights* w, int64_t rows) {
  if (!w || rows <= 0) return 0;
  return w->dtype == LIST_F32 ? mlp_f32_workspace_bytes(w, rows) : 0;
}

static int mlp_fwd_impl(const ListWeights* w, const void* X, int64_t ldx, int64_t rows, float* sdf, float out_div, void* workspace,
                        size_t workspace_bytes, int exact, void* stream) {
  int rc = check_weights(w, -1);
  if (rc) return rc;
  LIST_CHECK_ARG(rows >= 0, "list_mlp_fwd: rows < 0");
  if (rows == 0) return LIST_OK;
  LIST_CHECK_ARG(X != nullptr && sdf != nullptr, "list_mlp_fwd: X/sdf is NULL");
  LIST_CHECK_ARG(out_div != 0.f, "list_mlp_fwd: out_div must be non-zero");
  LIST_CHECK_ARG(ldx >= w->k_pad && ldx % 8 == 0, "list_mlp_fwd: ldx %lld must be >= k_pad %d and a multiple of 8", (long long)ldx, w->k_pad);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (w->dtype == LIST_F32) {
    LIST_CHECK_ARG(rows < (1LL << 31), "list_mlp_fwd: rows too large for one call");
    const size_t need = mlp_f32_workspace_bytes(w, rows);
    if (workspace == nullptr || workspace_bytes < need) {
      set_error("list_mlp_fwd: workspace %zu B < required %zu B", workspace_bytes, need);
      return LIST_ENOMEM;
    }
    return mlp_f32_fwd(w, static_cast<const float*>(X), ldx, rows, sdf, out_div, static_cast<float*>(workspace), exact, st);
  }
  return mlp_tc_fwd(w, X, ldx, rows, sdf, out_div, mlp_variant(), nullptr, nullptr, nullptr, st);
}

int list_mlp_fwd(const ListWeights* w, const void* X, int64_t ldx, int64_t rows, float* sdf, float out_div, void* workspace,
                 size_t workspace_bytes, void* stream) {
  return mlp_fwd_impl(w, X, ldx, rows, sdf, out_div, workspace, workspace_bytes, 0, stream);
}

int list_mlp_fwd_train(const ListWeights* w, const void* X, int64_t ldx, int64_t rows, float* sdf, float out_div, void* workspace,
                       size_t workspace_bytes, void* stream) {
  LIST_CHECK_ARG(w != nullptr && w->dtype == LIST_F32, "list_mlp_fwd_train: the training path is fp32");
  return mlp_fwd_impl(w, X, ldx, rows, sdf, out_div, workspace, workspace_bytes, 1, stream);
}

int list_mlp_fwd_debug(const ListWeights* w, const void* X, int64_t ldx, int64_t rows, float* sdf, float out_div,
                       float* h1, float* h2, float* h3, void* stream) {
  int rc = check_weights(w, -1);
  if (rc) return rc;
  LIST_CHECK_ARG(w->dtype == LIST_BF16, "list_mlp_fwd_debug: bf16 tensor-core kernel only");
  LIST_CHECK_ARG(rows >= 1 && X && sdf && out_div != 0.f, "list_mlp_fwd_debug: bad arguments");
  LIST_CHECK_ARG(ldx >= w->k_pad && ldx % 8 == 0, "list_mlp_fwd_debug: bad ldx %lld", (long long)ldx);
  return mlp_tc_fwd(w, X, ldx, rows, sdf, out_div, mlp_variant(), h1, h2, h3, static_cast<cudaStream_t>(stream));
}

size_t list_sdf_workspace_bytes(const ListCtx* ctx, const ListWeights* w, int64_t chunk_rows) {
  if (!ctx || !w || chunk_rows <= 0) return 0;
  ListLayout lay;
  if (list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr)) return 0;
  const size_t xb = align_up(static_cast<size_t>(chunk_rows) * lay.k_pad * elem_size(ctx->dtype), 256);
  // bf16: two feature-row buffers (gather of chunk i+1 overlaps the MLP of chunk i) + the hoisted-fc_0 tensors of
  // list_sdf_grid (projected maps / coarse volumes, W0h); fp32: one buffer + MLP activations
  size_t extra = 0;
  hoist::Plan pl;
  if (ctx->dtype == LIST_BF16 && hoist::make_plan(ctx, w, &pl) == LIST_OK) extra = align_up(pl.total, 256);
  if (ctx->dtype == LIST_BF16 && hoist::make_plan(ctx, w, &pl, kLinesLevels, kLinesMaxRes) == LIST_OK && align_up(pl.total, 256) > extra)
    extra = align_up(pl.total, 256);
  return (ctx->dtype == LIST_BF16 ? 2 * xb : xb) + align_up(list_mlp_workspace_bytes(w, chunk_rows), 256) + extra;
}

// Workspace layout of the dense-grid call for a given resolution: which path runs (env switches are read here, at the
// time of the call) and where its pieces live.  [chunk buffer 0 | chunk buffer 1 (bf16) | MLP workspace | hoisted tensors]
namespace {
struct GridWs {
  int path;                          // 0: full feature rows (fp32, or no hoisted path), 1: addend-kernel path, 2: line-table path
  size_t xb, mlp_off, hoist_off, total;
  size_t xr_bytes, g_bytes;          // line-table path: pieces of a chunk buffer [Xr | G | plans]
  hoist::Plan pl;
};
int grid_ws_plan(const ListCtx* ctx, const ListWeights* w, int res, int64_t chunk_rows, GridWs* g) {
  ListLayout lay;
  const int rc = list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr);
  if (rc) return rc;
  const bool two = ctx->dtype == LIST_BF16;
  const size_t mlp_ws = align_up(list_mlp_workspace_bytes(w, chunk_rows), 256);
  g->path = 0;
  g->xr_bytes = g->g_bytes = 0;
  g->xb = align_up(static_cast<size_t>(chunk_rows) * lay.k_pad * elem_size(ctx->dtype), 256);
  size_t hoisted = 0;
  if (two && hoist_enabled()) {
    if (lines_enabled() && hoist::make_plan(ctx, w, &g->pl, kLinesLevels, kLinesMaxRes) == LIST_OK &&
        hoist::check_gather(ctx, g->pl, res) == LIST_OK) {
      const int k_f = g->pl.k_h - 512;
      g->xr_bytes = align_up(static_cast<size_t>(chunk_rows) * k_f * 2, 256);
      g->g_bytes = align_up(hoist::lines_bytes(g->pl, res, res - 1, chunk_rows), 256);   // worst case: the chunk starts on a line's last point
      const size_t need = align_up(g->xr_bytes + g->g_bytes + grid_plan_bytes(res, res - 1, chunk_rows), 256);
      if (need <= g->xb) {             // (tiny grids: a line's tables outweigh its few rows -- the addend path takes those)
        g->xb = need;
        g->path = 2;
        hoisted = align_up(g->pl.total, 256);
      }
    }
    if (g->path == 0 && hoist::make_plan(ctx, w, &g->pl) == LIST_OK && hoist::check_gather(ctx, g->pl, res) == LIST_OK) {
      g->xb = align_up(static_cast<size_t>(chunk_rows) * g->pl.k_h * 2, 256);
      g->path = 1;
      hoisted = align_up(g->pl.total, 256);
    }
  }
  g->mlp_off = two ? 2 * g->xb : g->xb;
  g->hoist_off = g->mlp_off + mlp_ws;
  g->total = g->hoist_off + hoisted;
  return LIST_OK;
}
}  // namespace

size_t list_sdf_grid_workspace_bytes(const ListCtx* ctx, const ListWeights* w, int32_t res, int64_t chunk_rows) {
  if (!ctx || !w || chunk_rows <= 0 || res < 1 || check_ctx(ctx) || check_weights(w, -1)) return 0;
  GridWs g;
  return grid_ws_plan(ctx, w, res, chunk_rows, &g) == LIST_OK ? g.total : 0;
}

size_t list_hoist_bytes(const ListCtx* ctx, const ListWeights* w) {
  if (!ctx || !w) return 0;
  hoist::Plan pl;
  if (check_ctx(ctx) || check_weights(w, -1) || hoist::make_plan(ctx, w, &pl) != LIST_OK) return 0;
  return pl.total;
}

int list_hoist_layout(const ListCtx* ctx, const ListWeights* w, int32_t* hoist_cols, int32_t* k_h) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if ((rc = check_weights(w, -1))) return rc;
  hoist::Plan pl;
  if (hoist::make_plan(ctx, w, &pl) != LIST_OK) {
    set_error("list_hoist_layout: this configuration has no hoisted path (bf16, fc_0 width 512, coarse levels with C %% 64 == 0)");
    return LIST_ENOSYS;
  }
  if (hoist_cols) *hoist_cols = pl.hoist_cols;
  if (k_h) *k_h = pl.k_h;
  return LIST_OK;
}

int list_hoist_prepare(const ListCtx* ctx, const ListWeights* w, void* hoist_buf, size_t hoist_bytes, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if ((rc = check_weights(w, -1))) return rc;
  hoist::Plan pl;
  if (hoist::make_plan(ctx, w, &pl) != LIST_OK) {
    set_error("list_hoist_prepare: this configuration has no hoisted path (bf16, fc_0 width 512, coarse levels with C %% 64 == 0)");
    return LIST_ENOSYS;
  }
  if (!hoist_buf || hoist_bytes < pl.total || (reinterpret_cast<uintptr_t>(hoist_buf) & 255)) {
    set_error("list_hoist_prepare: hoist_buf NULL, not 256B aligned or %zu B < required %zu B", hoist_bytes, pl.total);
    return LIST_ENOMEM;
  }
  return hoist::prepare(ctx, w, pl, hoist_buf, static_cast<cudaStream_t>(stream));
}

int list_mlp_hoisted_fwd(const ListWeights* w, int32_t hoist_cols, const void* Xh, int64_t ldx, int64_t rows, float* sdf,
                         float out_div, void* stream) {
  int rc = check_weights(w, -1);
  if (rc) return rc;
  LIST_CHECK_ARG(w->dtype == LIST_BF16, "list_mlp_hoisted_fwd: bf16 tensor-core kernel only");
  LIST_CHECK_ARG(rows >= 0, "list_mlp_hoisted_fwd: rows < 0");
  if (rows == 0) return LIST_OK;
  LIST_CHECK_ARG(Xh && sdf && out_div != 0.f, "list_mlp_hoisted_fwd: X/sdf NULL or out_div == 0");
  LIST_CHECK_ARG(hoist_cols > 0 && hoist_cols < w->k_pad && hoist_cols % 64 == 0, "list_mlp_hoisted_fwd: hoist_cols %d invalid for k_pad %d",
                 hoist_cols, w->k_pad);
  return mlp_tc_fwd_hoisted(w, hoist_cols, w->k_pad - hoist_cols, Xh, ldx, rows, sdf, out_div, mlp_variant(), nullptr, nullptr, nullptr,
                            nullptr, static_cast<cudaStream_t>(stream));
}

int list_mlp_hoisted_trace(const ListWeights* w, int32_t hoist_cols, const void* Xh, int64_t ldx, int64_t rows, float* sdf,
                           float out_div, int64_t* trace, void* stream) {
  int rc = check_weights(w, -1);
  if (rc) return rc;
  LIST_CHECK_ARG(w->dtype == LIST_BF16 && rows >= 1 && Xh && sdf && trace && out_div != 0.f, "list_mlp_hoisted_trace: bad arguments");
  LIST_CHECK_ARG(hoist_cols > 0 && hoist_cols < w->k_pad && hoist_cols % 64 == 0, "list_mlp_hoisted_trace: hoist_cols %d invalid", hoist_cols);
  return mlp_tc_fwd_hoisted(w, hoist_cols, w->k_pad - hoist_cols, Xh, ldx, rows, sdf, out_div, mlp_variant(), nullptr, nullptr, nullptr,
                            reinterpret_cast<long long*>(trace), static_cast<cudaStream_t>(stream));
}

// ---- line-table path, stage by stage (tests, per-kernel timing) ----
static int lines_plan(const ListCtx* ctx, const ListWeights* w, hoist::Plan* pl, const char* who) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if ((rc = check_weights(w, -1))) return rc;
  if (hoist::make_plan(ctx, w, pl, kLinesLevels, kLinesMaxRes) != LIST_OK) {
    set_error("%s: this configuration has no hoisted path (bf16, fc_0 width 512, coarse levels with C %% 64 == 0)", who);
    return LIST_ENOSYS;
  }
  return LIST_OK;
}
static int check_range(int32_t res, int64_t begin, int64_t count, const char* who) {
  LIST_CHECK_ARG(res >= 1 && res <= 2048, "%s: res %d out of range", who, res);
  const int64_t total = static_cast<int64_t>(res) * res * res;
  LIST_CHECK_ARG(begin >= 0 && count >= 0 && begin + count <= total, "%s: [%lld,+%lld) outside res^3", who, (long long)begin, (long long)count);
  return LIST_OK;
}

size_t list_lines_hoist_bytes(const ListCtx* ctx, const ListWeights* w) {
  hoist::Plan pl;
  if (!ctx || !w || lines_plan(ctx, w, &pl, "list_lines_hoist_bytes")) return 0;
  return pl.total;
}

int list_lines_layout(const ListCtx* ctx, const ListWeights* w, int32_t* hoist_cols, int32_t* k_f, int32_t* rows_per_line) {
  hoist::Plan pl;
  const int rc = lines_plan(ctx, w, &pl, "list_lines_layout");
  if (rc) return rc;
  if (hoist_cols) *hoist_cols = pl.hoist_cols;
  if (k_f) *k_f = pl.k_h - 512;
  if (rows_per_line) *rows_per_line = pl.rpl;
  return LIST_OK;
}

int list_lines_prepare(const ListCtx* ctx, const ListWeights* w, void* hoist_buf, size_t hoist_bytes, void* stream) {
  hoist::Plan pl;
  const int rc = lines_plan(ctx, w, &pl, "list_lines_prepare");
  if (rc) return rc;
  if (!hoist_buf || hoist_bytes < pl.total || (reinterpret_cast<uintptr_t>(hoist_buf) & 255)) {
    set_error("list_lines_prepare: hoist_buf NULL, not 256B aligned or %zu B < required %zu B", hoist_bytes, pl.total);
    return LIST_ENOMEM;
  }
  return hoist::prepare(ctx, w, pl, hoist_buf, static_cast<cudaStream_t>(stream));
}

size_t list_lines_table_bytes(const ListCtx* ctx, const ListWeights* w, int32_t res, int64_t begin, int64_t count) {
  hoist::Plan pl;
  if (!ctx || !w || res < 1 || lines_plan(ctx, w, &pl, "list_lines_table_bytes")) return 0;
  return hoist::lines_bytes(pl, res, begin, count);
}

int list_lines_table(const ListCtx* ctx, const ListWeights* w, const void* hoist_buf, int32_t image, int32_t res, double bb_min,
                     double bb_max, int64_t begin, int64_t count, void* G, size_t G_bytes, void* stream) {
  hoist::Plan pl;
  int rc = lines_plan(ctx, w, &pl, "list_lines_table");
  if (rc) return rc;
  if ((rc = check_range(res, begin, count, "list_lines_table"))) return rc;
  LIST_CHECK_ARG(image >= 0 && image < ctx->B, "list_lines_table: image %d outside [0,%d)", image, ctx->B);
  LIST_CHECK_ARG(hoist_buf && G && (reinterpret_cast<uintptr_t>(G) & 15) == 0, "list_lines_table: hoist_buf / G NULL or G unaligned");
  if (G_bytes < hoist::lines_bytes(pl, res, begin, count)) {
    set_error("list_lines_table: G %zu B < required %zu B", G_bytes, hoist::lines_bytes(pl, res, begin, count));
    return LIST_ENOMEM;
  }
  return hoist::lines(ctx, pl, hoist_buf, image, res, bb_min, bb_max, begin, count, G, static_cast<cudaStream_t>(stream));
}

int list_lines_rest(const ListCtx* ctx, const ListWeights* w, int32_t image, int32_t res, double bb_min, double bb_max, int64_t begin,
                    int64_t count, void* Xr, int64_t ldx, void* stream) {
  hoist::Plan pl;
  int rc = lines_plan(ctx, w, &pl, "list_lines_rest");
  if (rc) return rc;
  if ((rc = check_range(res, begin, count, "list_lines_rest"))) return rc;
  if (hoist::check_gather(ctx, pl, res) != LIST_OK) {
    set_error("list_lines_rest: res %d not covered by the rest kernel", res);
    return LIST_ENOSYS;
  }
  LIST_CHECK_ARG(image >= 0 && image < ctx->B, "list_lines_rest: image %d outside [0,%d)", image, ctx->B);
  LIST_CHECK_ARG(Xr && ldx >= pl.k_h - 512 && ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(Xr) & 15) == 0,
                 "list_lines_rest: Xr NULL / unaligned or ldx %lld < %d", (long long)ldx, pl.k_h - 512);
  // the rest kernel addresses columns relative to a row that starts with the 512 addend columns of the round-1 layout
  return hoist::gather(ctx, w, pl, nullptr, image, res, bb_min, bb_max, begin, count, static_cast<__nv_bfloat16*>(Xr) - 512, ldx,
                       hoist::kPartRest | hoist::kPartOnes, static_cast<cudaStream_t>(stream));
}

size_t list_grid_plan_bytes(int32_t res, int64_t begin, int64_t count) {
  if (res < 1 || begin < 0 || count <= 0) return 0;
  return grid_plan_bytes(res, begin, count);
}

int list_grid_plan(const ListCtx* ctx, const ListWeights* w, const void* hoist_buf, int32_t image, int32_t res, double bb_min,
                   double bb_max, int64_t begin, int64_t count, const void* G, void* plan, size_t plan_bytes, void* stream) {
  hoist::Plan pl;
  int rc = lines_plan(ctx, w, &pl, "list_grid_plan");
  if (rc) return rc;
  if ((rc = check_range(res, begin, count, "list_grid_plan"))) return rc;
  LIST_CHECK_ARG(image >= 0 && image < ctx->B, "list_grid_plan: image %d outside [0,%d)", image, ctx->B);
  LIST_CHECK_ARG(hoist_buf && G && plan, "list_grid_plan: NULL argument");
  if (plan_bytes < grid_plan_bytes(res, begin, count)) {
    set_error("list_grid_plan: plan buffer %zu B < required %zu B", plan_bytes, grid_plan_bytes(res, begin, count));
    return LIST_ENOMEM;
  }
  return grid_plan(ctx, pl, hoist_buf, image, res, bb_min, bb_max, begin, count, G, plan, static_cast<cudaStream_t>(stream));
}

int list_grid_tc_fwd(const ListCtx* ctx, const ListWeights* w, const void* hoist_buf, int32_t res, double bb_min, double bb_max,
                     int64_t begin, int64_t count, const void* Xr, int64_t ldx, const void* plan, float* sdf, float out_div,
                     float* dbg_h1, int64_t* trace, uint64_t* stats, void* stream) {
  hoist::Plan pl;
  int rc = lines_plan(ctx, w, &pl, "list_grid_tc_fwd");
  if (rc) return rc;
  if ((rc = check_range(res, begin, count, "list_grid_tc_fwd"))) return rc;
  LIST_CHECK_ARG(hoist_buf && Xr && plan && sdf && out_div != 0.f, "list_grid_tc_fwd: NULL argument or out_div == 0");
  return grid_tc_fwd(ctx, w, pl, hoist_buf, res, bb_min, bb_max, begin, count, Xr, ldx, plan, sdf, out_div, dbg_h1,
                     reinterpret_cast<long long*>(trace), reinterpret_cast<unsigned long long*>(stats), static_cast<cudaStream_t>(stream));
}

int list_hoist_gather_grid_fwd(const ListCtx* ctx, const ListWeights* w, const void* hoist_buf, int32_t image, int32_t res,
                               double bb_min, double bb_max, int64_t begin, int64_t count, void* X, int64_t ldx, int32_t parts,
                               void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  if ((rc = check_weights(w, -1))) return rc;
  hoist::Plan pl;
  if (hoist::make_plan(ctx, w, &pl) != LIST_OK || hoist::check_gather(ctx, pl, res) != LIST_OK) {
    set_error("list_hoist_gather_grid_fwd: configuration / res %d not covered by the hoisted gather", res);
    return LIST_ENOSYS;
  }
  LIST_CHECK_ARG(image >= 0 && image < ctx->B, "list_hoist_gather_grid_fwd: image %d outside [0,%d)", image, ctx->B);
  LIST_CHECK_ARG(res >= 1 && res <= 2048, "list_hoist_gather_grid_fwd: res %d out of range", res);
  const int64_t total = static_cast<int64_t>(res) * res * res;
  LIST_CHECK_ARG(begin >= 0 && count >= 0 && begin + count <= total, "list_hoist_gather_grid_fwd: [%lld,+%lld) outside res^3",
                 (long long)begin, (long long)count);
  LIST_CHECK_ARG(hoist_buf && X && ldx >= pl.k_h && ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0,
                 "list_hoist_gather_grid_fwd: hoist_buf/X NULL, X unaligned or ldx %lld < %d", (long long)ldx, pl.k_h);
  LIST_CHECK_ARG(parts >= 1 && parts <= 3, "list_hoist_gather_grid_fwd: parts %d must be 1 (addend), 2 (rest) or 3 (both)", parts);
  return hoist::gather(ctx, w, pl, hoist_buf, image, res, bb_min, bb_max, begin, count, X, ldx, parts, static_cast<cudaStream_t>(stream));
}

int list_sdf_fwd(const ListCtx* ctx, const ListWeights* w, const float* q, int32_t q_is_raw, int32_t B, int64_t N, float* sdf,
                 float out_div, int64_t chunk_rows, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  ListLayout lay;
  if ((rc = list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr))) return rc;
  if ((rc = check_weights(w, lay.k_pad))) return rc;
  LIST_CHECK_ARG(w->dtype == ctx->dtype, "list_sdf_fwd: weights.dtype %d != ctx.dtype %d", w->dtype, ctx->dtype);
  LIST_CHECK_ARG(B == ctx->B && N >= 0, "list_sdf_fwd: B %d != ctx.B %d or N < 0", B, ctx->B);
  if (N == 0) return LIST_OK;
  LIST_CHECK_ARG(q && sdf && out_div != 0.f, "list_sdf_fwd: q/sdf NULL or out_div == 0");
  LIST_CHECK_ARG(chunk_rows >= 1, "list_sdf_fwd: chunk_rows < 1");
  const size_t need = list_sdf_workspace_bytes(ctx, w, chunk_rows);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("list_sdf_fwd: workspace %zu B < required %zu B", workspace_bytes, need);
    return LIST_ENOMEM;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t xb = align_up(static_cast<size_t>(chunk_rows) * lay.k_pad * elem_size(ctx->dtype), 256);
  const bool two = ctx->dtype == LIST_BF16;
  void* const xbuf[2] = {workspace, two ? static_cast<char*>(workspace) + xb : workspace};
  const size_t mlp_off = two ? 2 * xb : xb;
  void* mlp_ws = static_cast<char*>(workspace) + mlp_off;
  // per image, chunks of the point range; one image at a time keeps the ctx indexing trivial
  const int64_t per_image = (N + chunk_rows - 1) / chunk_rows;
  auto one_image = [&](int b) {
    ListCtx one = *ctx;
    one.B = 1;
    one.maps = static_cast<const char*>(ctx->maps) + static_cast<size_t>(b) * ctx->map_size * ctx->map_size * ctx->map_channels * elem_size(ctx->dtype);
    for (int l = 0; l < ctx->n_levels; ++l) {
      const size_t vox = static_cast<size_t>(ctx->vol_res[l]) * ctx->vol_res[l] * ctx->vol_res[l] * ctx->vol_ch[l];
      one.vols[l] = static_cast<const char*>(ctx->vols[l]) + static_cast<size_t>(b) * vox * elem_size(ctx->dtype);
    }
    one.trans_mat = ctx->trans_mat + b * 12;
    return one;
  };
  auto span = [&](int64_t i, int& b, int64_t& n0, int64_t& n) {
    b = static_cast<int>(i / per_image);
    n0 = (i % per_image) * chunk_rows;
    n = (N - n0 < chunk_rows) ? (N - n0) : chunk_rows;
  };
  return run_chunks(
      per_image * B, xbuf, two && overlap_enabled(), st, false,
      [](int64_t, void*, cudaStream_t) { return LIST_OK; },
      [&](int64_t i, void* X, cudaStream_t s) {
        int b; int64_t n0, n;
        span(i, b, n0, n);
        const ListCtx one = one_image(b);
        return gather_fwd(&one, q + (static_cast<int64_t>(b) * N + n0) * 3, q_is_raw, X, lay.k_pad, 1, n, s);
      },
      [&](int64_t i, void* X, cudaStream_t s) {
        int b; int64_t n0, n;
        span(i, b, n0, n);
        return list_mlp_fwd(w, X, lay.k_pad, n, sdf + static_cast<int64_t>(b) * N + n0, out_div, mlp_ws,
                            workspace_bytes - mlp_off, s);
      });
}

static int grid_impl(const ListCtx* ctx, const ListWeights* w, int32_t res, double bb_min, double bb_max, int64_t begin,
                     int64_t count, float* sdf, float sdf_scale, int64_t chunk_rows, void* workspace, size_t workspace_bytes,
                     void* stream, const GridHooks* hooks) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  ListLayout lay;
  if ((rc = list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr))) return rc;
  if ((rc = check_weights(w, lay.k_pad))) return rc;
  LIST_CHECK_ARG(w->dtype == ctx->dtype, "list_sdf_grid: weights.dtype %d != ctx.dtype %d", w->dtype, ctx->dtype);
  LIST_CHECK_ARG(res >= 1 && res <= 2048, "list_sdf_grid: res %d out of range", res);
  const int64_t total = static_cast<int64_t>(res) * res * res;
  LIST_CHECK_ARG(begin >= 0 && count >= 0 && begin + count <= total, "list_sdf_grid: [%lld,+%lld) outside res^3",
                 (long long)begin, (long long)count);
  if (count == 0) return LIST_OK;
  LIST_CHECK_ARG(sdf != nullptr && sdf_scale != 0.f && chunk_rows >= 1, "list_sdf_grid: sdf NULL, sdf_scale 0 or chunk_rows < 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool late_done = false;
  GridWs gw;
  if ((rc = grid_ws_plan(ctx, w, res, chunk_rows, &gw))) return rc;
  if (workspace == nullptr || workspace_bytes < gw.total) {
    set_error("list_sdf_grid: workspace %zu B < required %zu B (list_sdf_grid_workspace_bytes)", workspace_bytes, gw.total);
    return LIST_ENOMEM;
  }
  const size_t xb = gw.xb;
  const bool two = ctx->dtype == LIST_BF16;
  void* const xbuf[2] = {workspace, two ? static_cast<char*>(workspace) + xb : workspace};
  const size_t mlp_off = gw.mlp_off;
  void* mlp_ws = static_cast<char*>(workspace) + mlp_off;
  void* const hbuf = static_cast<char*>(workspace) + gw.hoist_off;
  const int64_t per_image = (count + chunk_rows - 1) / chunk_rows;
  auto span = [&](int64_t i, int& b, int64_t& n0, int64_t& n) {
    b = static_cast<int>(i / per_image);
    n0 = (i % per_image) * chunk_rows;
    n = (count - n0 < chunk_rows) ? (count - n0) : chunk_rows;
  };
  // bf16, line-table path (default): project the maps and the coarse levels (R <= 32) through their W0 blocks once per
  // call (hoist::prepare); per chunk, lines.cu reduces the projected levels to one column table per z-line, the rest
  // kernel writes the non-hoisted feature columns, and grid_tc.cu interpolates the hoisted terms on the tensor cores inside
  // the MLP kernel.  A chunk's buffer holds [Xr: rows x k_f | G: line tables].
  if (gw.path == 2) {
    const hoist::Plan& pl3 = gw.pl;
    const int k_f = pl3.k_h - 512;
    const size_t xr_bytes = gw.xr_bytes, g_bytes = gw.g_bytes;
    {
      if ((rc = hoist::prepare(ctx, w, pl3, hbuf, st))) return rc;
      // the line tables and the plans only read the projected tensors; everything uploaded late is first read by the rest
      // kernel -- with a late upload pending, two chunks' tables and plans are enqueued in front of it
      const bool late_pending = !late_done && hooks && hooks->late;
      return run_chunks(
          per_image * ctx->B, xbuf, overlap_enabled(), st, late_pending,
          [&](int64_t i, void* X, cudaStream_t s) -> int {
            int b; int64_t n0, n;
            span(i, b, n0, n);
            char* const G = static_cast<char*>(X) + xr_bytes;
            const int r2 = hoist::lines(ctx, pl3, hbuf, b, res, bb_min, bb_max, begin + n0, n, G, s);
            return r2 ? r2 : grid_plan(ctx, pl3, hbuf, b, res, bb_min, bb_max, begin + n0, n, G, G + g_bytes, s);
          },
          [&](int64_t i, void* X, cudaStream_t s) -> int {
            int b; int64_t n0, n;
            span(i, b, n0, n);
            int r2;
            if (i == 0 && !late_done && hooks && hooks->late && (r2 = hook_late(hooks, s))) return r2;
            return hoist::gather(ctx, w, pl3, hbuf, b, res, bb_min, bb_max, begin + n0, n, static_cast<__nv_bfloat16*>(X) - 512, k_f,
                                 hoist::kPartRest | hoist::kPartOnes, s);
          },
          [&](int64_t i, void* X, cudaStream_t s) -> int {
            int b; int64_t n0, n;
            span(i, b, n0, n);
            const int r2 = grid_tc_fwd(ctx, w, pl3, hbuf, res, bb_min, bb_max, begin + n0, n, X, k_f, static_cast<char*>(X) + xr_bytes + g_bytes,
                                       sdf + static_cast<int64_t>(b) * count + n0, sdf_scale, nullptr, nullptr, nullptr, s);
            return r2 ? r2 : hook_download(hooks, sdf, static_cast<int64_t>(b) * count + n0, n, s);
          });
    }
  }
  // bf16, addend-kernel path (round 1; LIST_B200_LINES=0 or a configuration the line-table path does not cover): the
  // per-chunk gather writes the hoisted row [addend 512 | 832 columns]; fc_0 runs on the 832 columns and adds the addend
  // block in its epilogue.
  if (gw.path == 1) {
    const hoist::Plan& pl = gw.pl;
    if ((rc = hoist::prepare(ctx, w, pl, hbuf, st))) return rc;
    return run_chunks(
        per_image * ctx->B, xbuf, overlap_enabled(), st, false,
        [](int64_t, void*, cudaStream_t) { return LIST_OK; },
        [&](int64_t i, void* X, cudaStream_t s) -> int {
          int b; int64_t n0, n;
          span(i, b, n0, n);
          if (i == 0 && !late_done && hooks && hooks->late) {
            // the addend only reads the projected tensors; everything uploaded late is first read by the other part
            int r2 = hoist::gather(ctx, w, pl, hbuf, b, res, bb_min, bb_max, begin + n0, n, X, pl.k_h, hoist::kPartAddend, s);
            if (r2) return r2;
            if ((r2 = hook_late(hooks, s))) return r2;
            return hoist::gather(ctx, w, pl, hbuf, b, res, bb_min, bb_max, begin + n0, n, X, pl.k_h, hoist::kPartRest, s);
          }
          return hoist::gather(ctx, w, pl, hbuf, b, res, bb_min, bb_max, begin + n0, n, X, pl.k_h, 3, s);
        },
        [&](int64_t i, void* X, cudaStream_t s) -> int {
          int b; int64_t n0, n;
          span(i, b, n0, n);
          const int r2 = mlp_tc_fwd_hoisted(w, pl.hoist_cols, pl.k_h - 512, X, pl.k_h, n, sdf + static_cast<int64_t>(b) * count + n0,
                                            sdf_scale, mlp_variant(), nullptr, nullptr, nullptr, nullptr, s);
          return r2 ? r2 : hook_download(hooks, sdf, static_cast<int64_t>(b) * count + n0, n, s);
        });
  }
  if (!late_done && (rc = hook_late(hooks, st))) return rc;
  return run_chunks(
      per_image * ctx->B, xbuf, two && overlap_enabled(), st, false,
      [](int64_t, void*, cudaStream_t) { return LIST_OK; },
      [&](int64_t i, void* X, cudaStream_t s) {
        int b; int64_t n0, n;
        span(i, b, n0, n);
        return grid_gather(ctx, b, res, bb_min, bb_max, begin + n0, n, X, lay.k_pad, s);
      },
      [&](int64_t i, void* X, cudaStream_t s) -> int {
        int b; int64_t n0, n;
        span(i, b, n0, n);
        const int r2 = list_mlp_fwd(w, X, lay.k_pad, n, sdf + static_cast<int64_t>(b) * count + n0, sdf_scale, mlp_ws,
                                    workspace_bytes - mlp_off, s);
        return r2 ? r2 : hook_download(hooks, sdf, static_cast<int64_t>(b) * count + n0, n, s);
      });
}

int list_sdf_grid(const ListCtx* ctx, const ListWeights* w, int32_t res, double bb_min, double bb_max, int64_t begin,
                  int64_t count, float* sdf, float sdf_scale, int64_t chunk_rows, void* workspace, size_t workspace_bytes,
                  void* stream) {
  return grid_impl(ctx, w, res, bb_min, bb_max, begin, count, sdf, sdf_scale, chunk_rows, workspace, workspace_bytes, stream,
                   nullptr);
}

// Device-resident variant of the staged evaluation: the caller has the coarse tensors prepared in `ctx` already and is
// still producing (uploading, all-gathering ...) the reference-layout volumes of the other levels on another stream.
int list_sdf_grid_late(const ListCtx* ctx, const ListWeights* w, int32_t res, double bb_min, double bb_max, int64_t begin,
                       int64_t count, float* sdf, float sdf_scale, int64_t chunk_rows, void* workspace, size_t workspace_bytes,
                       void* stream, void* late_event, const float* const* late_vols_ncdhw, float* sdf_host) {
  LIST_CHECK_ARG(ctx != nullptr, "list_sdf_grid_late: ctx is NULL");
  Pipe* pp = nullptr;
  int rc;
  if ((rc = get_pipe(&pp))) return rc;
  struct Late { const ListCtx* ctx; const float* const* raw; } late{ctx, late_vols_ncdhw};
  GridHooks hooks;
  if (late_vols_ncdhw) {
    LIST_CHECK_ARG(w != nullptr && ctx->n_levels >= 1 && ctx->n_levels <= LIST_MAX_LEVELS, "list_sdf_grid_late: bad ctx / weights");
    hoist::Plan pl;                                     // the projection runs first and reads the hoisted levels
    if (ctx->dtype == LIST_BF16 && hoist_enabled() && hoist::make_plan(ctx, w, &pl, kLinesLevels, kLinesMaxRes) == LIST_OK)
      for (int h = 0; h < pl.nh; ++h)
        LIST_CHECK_ARG(late_vols_ncdhw[pl.lev[h]] == nullptr,
                       "list_sdf_grid_late: level %d (R <= 32, C %% 64 == 0) is read by the projection and cannot be late", pl.lev[h]);
    hooks.uploaded = static_cast<cudaEvent_t>(late_event);
    hooks.user = &late;
    hooks.late = [](void* u, cudaStream_t s) -> int {
      auto* l = static_cast<Late*>(u);
      for (int i = 0; i < l->ctx->n_levels; ++i) {
        if (!l->raw[i]) continue;
        const int r2 = list_prep_volume(l->raw[i], l->ctx->B, l->ctx->vol_ch[i], l->ctx->vol_res[i],
                                        const_cast<void*>(l->ctx->vols[i]), l->ctx->dtype, s);
        if (r2) return r2;
      }
      return LIST_OK;
    };
  }
  hooks.sdf_host = sdf_host;
  hooks.copy = pp->cp;
  hooks.item = pp->item;
  if ((rc = grid_impl(ctx, w, res, bb_min, bb_max, begin, count, sdf, sdf_scale, chunk_rows, workspace, workspace_bytes, stream,
                      &hooks)))
    return rc;
  if (sdf_host) {                                       // join: the caller's stream continues after the last download
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LIST_CUDA(cudaEventRecord(pp->cdone, pp->cp));
    LIST_CUDA(cudaStreamWaitEvent(st, pp->cdone, 0));
  }
  return LIST_OK;
}

// ---- host-buffer variant -----------------------------------------------------------------
struct HostPlan {
  size_t raw_maps[LIST_MAX_MAPS], raw_vols[LIST_MAX_LEVELS], raw_T, maps_cl, vols_cl[LIST_MAX_LEVELS], sdf, ws, total;
};
static void plan_host(const int32_t* map_ch, const int32_t* map_in, int n_maps, int S, int n_levels, const int32_t* vol_ch,
                      const int32_t* vol_res, int B, int dtype, int64_t count, int64_t chunk_rows, HostPlan* p) {
  size_t off = 0;
  int cm = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += align_up(bytes, 256); return o; };
  for (int i = 0; i < n_maps; ++i) { p->raw_maps[i] = take(static_cast<size_t>(B) * map_ch[i] * map_in[i] * map_in[i] * 4); cm += map_ch[i]; }
  for (int l = 0; l < n_levels; ++l)
    p->raw_vols[l] = take(static_cast<size_t>(B) * vol_ch[l] * vol_res[l] * vol_res[l] * vol_res[l] * 4);
  p->raw_T = take(static_cast<size_t>(B) * 12 * 4);
  p->maps_cl = take(static_cast<size_t>(B) * S * S * cm * elem_size(dtype));
  for (int l = 0; l < n_levels; ++l)
    p->vols_cl[l] = take(static_cast<size_t>(B) * vol_ch[l] * vol_res[l] * vol_res[l] * vol_res[l] * elem_size(dtype));
  p->sdf = take(static_cast<size_t>(B) * count * 4);
  ListLayout lay;
  list_feature_layout(cm, n_levels, vol_ch, &lay, nullptr);
  size_t ws = align_up(static_cast<size_t>(chunk_rows) * lay.k_pad * elem_size(dtype), 256);
  if (dtype == LIST_BF16) {
    ws *= 2;                                           // double-buffered feature rows (run_chunks)
    // hoisted-fc_0 tensors (hoist.cu): the larger of the two plans, as list_sdf_workspace_bytes counts them
    ListCtx cdim{};
    cdim.B = B; cdim.dtype = dtype; cdim.map_size = S; cdim.map_channels = cm; cdim.n_levels = n_levels;
    for (int l = 0; l < n_levels; ++l) { cdim.vol_res[l] = vol_res[l]; cdim.vol_ch[l] = vol_ch[l]; }
    ListWeights wdim{};
    wdim.dtype = LIST_BF16; wdim.k_pad = lay.k_pad; wdim.n0 = 512; wdim.n1 = 256; wdim.n2 = 256;
    size_t h = 0;
    hoist::Plan pl;
    if (hoist::make_plan(&cdim, &wdim, &pl) == LIST_OK) h = align_up(pl.total, 256);
    if (hoist::make_plan(&cdim, &wdim, &pl, kLinesLevels, kLinesMaxRes) == LIST_OK && align_up(pl.total, 256) > h) h = align_up(pl.total, 256);
    ws += h + 256;
  }
  if (dtype == LIST_F32) {                             // activations + lo copies of the fp32 MLP (mlp_f32_workspace_bytes)
    ListWeights wdim{};
    wdim.dtype = LIST_F32; wdim.k_pad = lay.k_pad; wdim.n0 = 512; wdim.n1 = 256; wdim.n2 = 256;
    ws += align_up(mlp_f32_workspace_bytes(&wdim, chunk_rows), 256);
  }
  p->ws = take(ws);
  p->total = off;
}

size_t list_sdf_grid_host_bytes(const int32_t* map_ch, const int32_t* map_size_in, int32_t n_maps, int32_t map_size,
                                int32_t n_levels, const int32_t* vol_ch, const int32_t* vol_res, int32_t B, int32_t dtype,
                                int64_t count, int64_t chunk_rows) {
  if (!map_ch || !map_size_in || !vol_ch || !vol_res || n_maps < 1 || n_maps > LIST_MAX_MAPS || n_levels < 1 ||
      n_levels > LIST_MAX_LEVELS || B < 1 || count < 0 || chunk_rows < 1)
    return 0;
  HostPlan p;
  plan_host(map_ch, map_size_in, n_maps, map_size, n_levels, vol_ch, vol_res, B, dtype, count, chunk_rows, &p);
  return p.total;
}

int list_sdf_grid_host(const float* const* maps_host, const int32_t* map_ch, const int32_t* map_size_in, int32_t n_maps,
                       int32_t map_size, const float* const* vols_host, int32_t n_levels, const int32_t* vol_ch,
                       const int32_t* vol_res, const float* trans_mat_host, int32_t B, int32_t dtype, const ListWeights* w_dev,
                       int32_t res, double bb_min, double bb_max, int64_t begin, int64_t count, float sdf_scale,
                       int64_t chunk_rows, float* sdf_host, void* dev_scratch, size_t dev_scratch_bytes, void* stream) {
  LIST_CHECK_ARG(maps_host && map_ch && map_size_in && vols_host && vol_ch && vol_res && trans_mat_host && w_dev && sdf_host,
                 "list_sdf_grid_host: NULL argument");
  LIST_CHECK_ARG(n_maps >= 1 && n_maps <= LIST_MAX_MAPS && n_levels >= 1 && n_levels <= LIST_MAX_LEVELS && B >= 1 &&
                 count >= 0 && chunk_rows >= 1, "list_sdf_grid_host: bad sizes");
  LIST_CHECK_ARG(dtype == LIST_F32 || dtype == LIST_BF16, "list_sdf_grid_host: bad dtype %d", dtype);
  LIST_CHECK_ARG(w_dev->n0 == 512 && w_dev->n1 == 256 && w_dev->n2 == 256, "list_sdf_grid_host: layer widths must be 512/256/256");
  HostPlan p;
  plan_host(map_ch, map_size_in, n_maps, map_size, n_levels, vol_ch, vol_res, B, dtype, count, chunk_rows, &p);
  if (!dev_scratch || dev_scratch_bytes < p.total) {
    set_error("list_sdf_grid_host: dev_scratch %zu B < required %zu B", dev_scratch_bytes, p.total);
    return LIST_ENOMEM;
  }
  LIST_CHECK_ARG((reinterpret_cast<uintptr_t>(dev_scratch) & 255) == 0, "list_sdf_grid_host: dev_scratch must be 256B aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(dev_scratch);
  Pipe* pp = nullptr;
  int rc;
  if ((rc = get_pipe(&pp))) return rc;
  // Uploads run on the copy stream, ordered after the caller's stream: T, maps and the coarse levels (R <= 32: all the
  // projection and the addend gather read) first, then the other levels from coarse to fine.
  LIST_CUDA(cudaEventRecord(pp->cfork, st));
  LIST_CUDA(cudaStreamWaitEvent(pp->cp, pp->cfork, 0));
  LIST_CUDA(cudaMemcpyAsync(base + p.raw_T, trans_mat_host, static_cast<size_t>(B) * 48, cudaMemcpyHostToDevice, pp->cp));
  const float* dmaps[LIST_MAX_MAPS];
  int cm = 0;
  for (int i = 0; i < n_maps; ++i) {
    const size_t bytes = static_cast<size_t>(B) * map_ch[i] * map_size_in[i] * map_size_in[i] * 4;
    LIST_CUDA(cudaMemcpyAsync(base + p.raw_maps[i], maps_host[i], bytes, cudaMemcpyHostToDevice, pp->cp));
    dmaps[i] = reinterpret_cast<const float*>(base + p.raw_maps[i]);
    cm += map_ch[i];
  }
  int order[LIST_MAX_LEVELS];
  for (int l = 0; l < n_levels; ++l) order[l] = l;
  auto vol_bytes = [&](int l) { return static_cast<size_t>(B) * vol_ch[l] * vol_res[l] * vol_res[l] * vol_res[l] * 4; };
  for (int i = 1; i < n_levels; ++i)                                      // insertion sort: coarse levels first, then by size
    for (int j = i; j > 0; --j) {
      const int x = order[j - 1], y = order[j];
      const bool sx = vol_res[x] <= kLinesMaxRes, sy = vol_res[y] <= kLinesMaxRes;
      if ((sy && !sx) || (sx == sy && vol_bytes(y) < vol_bytes(x))) { order[j - 1] = y; order[j] = x; } else break;
    }
  int n_small = 0;
  while (n_small < n_levels && vol_res[order[n_small]] <= kLinesMaxRes) ++n_small;
  for (int i = 0; i < n_small; ++i)
    LIST_CUDA(cudaMemcpyAsync(base + p.raw_vols[order[i]], vols_host[order[i]], vol_bytes(order[i]), cudaMemcpyHostToDevice, pp->cp));
  LIST_CUDA(cudaEventRecord(pp->small, pp->cp));
  for (int i = n_small; i < n_levels; ++i)
    LIST_CUDA(cudaMemcpyAsync(base + p.raw_vols[order[i]], vols_host[order[i]], vol_bytes(order[i]), cudaMemcpyHostToDevice, pp->cp));
  LIST_CUDA(cudaEventRecord(pp->big, pp->cp));

  LIST_CUDA(cudaStreamWaitEvent(st, pp->small, 0));
  if ((rc = list_prep_maps(dmaps, map_ch, map_size_in, n_maps, B, map_size, base + p.maps_cl, dtype, stream))) return rc;
  ListCtx ctx{};
  ctx.B = B;
  ctx.dtype = dtype;
  ctx.map_size = map_size;
  ctx.map_channels = cm;
  ctx.maps = base + p.maps_cl;
  ctx.n_levels = n_levels;
  for (int l = 0; l < n_levels; ++l) {
    ctx.vol_res[l] = vol_res[l];
    ctx.vol_ch[l] = vol_ch[l];
    ctx.vols[l] = base + p.vols_cl[l];
  }
  ctx.trans_mat = reinterpret_cast<const float*>(base + p.raw_T);
  auto prep_levels = [&](int i0, int i1, cudaStream_t s) -> int {
    for (int i = i0; i < i1; ++i) {
      const int l = order[i];
      const int r2 = list_prep_volume(reinterpret_cast<const float*>(base + p.raw_vols[l]), B, vol_ch[l], vol_res[l],
                                      base + p.vols_cl[l], dtype, s);
      if (r2) return r2;
    }
    return LIST_OK;
  };
  if ((rc = prep_levels(0, n_small, st))) return rc;
  struct Late { decltype(prep_levels)* f; int i0, i1; } late{&prep_levels, n_small, n_levels};
  GridHooks hooks;
  hooks.uploaded = pp->big;
  hooks.late = [](void* u, cudaStream_t s) -> int { auto* l = static_cast<Late*>(u); return (*l->f)(l->i0, l->i1, s); };
  hooks.user = &late;
  hooks.sdf_host = sdf_host;
  hooks.copy = pp->cp;
  hooks.item = pp->item;
  float* dsdf = reinterpret_cast<float*>(base + p.sdf);
  if ((rc = grid_impl(&ctx, w_dev, res, bb_min, bb_max, begin, count, dsdf, sdf_scale, chunk_rows, base + p.ws, p.total - p.ws,
                      stream, &hooks)))
    return rc;
  // join: the caller's stream continues after the last download
  LIST_CUDA(cudaEventRecord(pp->cdone, pp->cp));
  LIST_CUDA(cudaStreamWaitEvent(st, pp->cdone, 0));
  return LIST_OK;
}

// ---- fp32-accurate tensor-core GEMM, exposed for tests ------------------------------------------------
int list_gemm_f32_tc(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int32_t M, int32_t N, int32_t K,
                     int32_t accumulate, float* lo_workspace, size_t lo_bytes, void* stream) {
  LIST_CHECK_ARG(A && B && C && lo_workspace && M >= 1 && N >= 1 && K >= 1, "list_gemm_f32_tc: NULL argument or empty problem");
  const int64_t a_elems = static_cast<int64_t>(M) * lda, b_elems = static_cast<int64_t>(N) * ldb;
  LIST_CHECK_ARG(lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0, "list_gemm_f32_tc: leading dimensions must be multiples of 4");
  if (lo_bytes < static_cast<size_t>(a_elems + b_elems) * 4) {
    set_error("list_gemm_f32_tc: lo workspace %zu B < required %zu B", lo_bytes, static_cast<size_t>(a_elems + b_elems) * 4);
    return LIST_ENOMEM;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* Alo = lo_workspace;
  float* Blo = lo_workspace + a_elems;
  int rc;
  if ((rc = split_lo(A, Alo, a_elems, st))) return rc;
  if ((rc = split_lo(B, Blo, b_elems, st))) return rc;
  GemmEpilogue ep{};
  ep.accumulate = accumulate;
  return tgemm(A, Alo, lda, B, Blo, ldb, C, ldc, M, N, K, ep, st);
}

// ---- backward ------------------------------------------------------------------------------
size_t list_bwd_workspace_bytes(const ListWeights* w, int64_t rows) {
  if (!w || rows <= 0) return 0;
  return mlp_f32_bwd_workspace_bytes(w, rows);
}

int list_sdf_bwd(const ListCtx* ctx, const ListWeights* w, const float* q, int32_t q_is_raw, int32_t B, int64_t N, const void* X,
                 int64_t ldx, const void* fwd_workspace, const float* d_sdf, const ListGrads* grads, void* workspace,
                 size_t workspace_bytes, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  ListLayout lay;
  if ((rc = list_feature_layout(ctx->map_channels, ctx->n_levels, ctx->vol_ch, &lay, nullptr))) return rc;
  if ((rc = check_weights(w, lay.k_pad))) return rc;
  LIST_CHECK_ARG(ctx->dtype == LIST_F32 && w->dtype == LIST_F32, "list_sdf_bwd: the backward path is fp32 only");
  LIST_CHECK_ARG(B == ctx->B && N >= 0, "list_sdf_bwd: B %d != ctx.B %d or N < 0", B, ctx->B);
  const int64_t rows = static_cast<int64_t>(B) * N;
  if (rows == 0) return LIST_OK;
  LIST_CHECK_ARG(rows < (1LL << 31), "list_sdf_bwd: too many rows for one call");
  LIST_CHECK_ARG(q && X && fwd_workspace && d_sdf && grads, "list_sdf_bwd: NULL argument");
  LIST_CHECK_ARG(ldx >= lay.k_pad && ldx % 4 == 0, "list_sdf_bwd: bad ldx %lld", (long long)ldx);
  const size_t need = mlp_f32_bwd_workspace_bytes(w, rows);
  if (!workspace || workspace_bytes < need) {
    set_error("list_sdf_bwd: workspace %zu B < required %zu B", workspace_bytes, need);
    return LIST_ENOMEM;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((rc = mlp_f32_bwd(w, static_cast<const float*>(X), ldx, rows, static_cast<const float*>(fwd_workspace), d_sdf, grads,
                        static_cast<float*>(workspace), st)))
    return rc;
  return gather_bwd(ctx, q, q_is_raw, B, N, static_cast<const float*>(workspace), lay.k_pad, grads, st);
}

}  // extern "C"
