// Context, error reporting, staging and pyramid-batch memory management of libsvo_cuda.
#include "common.cuh"

int svoFail(svo_cuda_ctx* ctx, int code, const char* what, const char* file, int line) {
  if (ctx) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s (%s:%d)", what, file, line);
    ctx->last_error = buf;
  }
  return code;
}

Stager::Stager(svo_cuda_ctx* c, svo_mem m) : ctx_(c), mem_(m) {
  if (!c || m != SVO_MEM_HOST || c->stage_busy) return;
  if (!c->stage_host) {  // first host-memory call of the context
    cudaSetDevice(c->device);
    void* h = nullptr;
    void* d = nullptr;
    if (cudaHostAlloc(&h, 2 * kStageHalf + kZeroCopyRegion, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); return; }
    if (cudaMalloc(&d, 2 * kStageHalf) != cudaSuccess) { cudaGetLastError(); cudaFreeHost(h); return; }
    void* hd = nullptr;
    if (cudaHostGetDevicePointer(&hd, h, 0) != cudaSuccess) { cudaGetLastError(); hd = nullptr; }  // no mapping: results go through the device arena
    c->stage_host = (uint8_t*)h;
    c->stage_host_dev = (uint8_t*)hd;
    c->stage_dev = (uint8_t*)d;
  }
  c->stage_busy = true;
  arena_ = true;
}
void* Stager::arenaIn(const void* p, size_t bytes) {
  const size_t padded = (bytes + 15) & ~(size_t)15;
  if (!arena_ || sent_ || bytes > kStageSmall || in_used_ + padded > kStageHalf) return nullptr;
  memcpy(ctx_->stage_host + in_used_, p, bytes);
  void* d = ctx_->stage_dev + in_used_;
  in_used_ += padded;
  return d;
}
void* Stager::zeroCopyOut(void* p, size_t bytes) {
  const size_t padded = (bytes + 15) & ~(size_t)15;
  if (!arena_ || !ctx_->stage_host_dev || bytes > kZeroCopyOut || zc_used_ + padded > kZeroCopyRegion) return nullptr;
  arena_outs_.push_back({p, ctx_->stage_host + 2 * kStageHalf + zc_used_, bytes});
  void* d = ctx_->stage_host_dev + 2 * kStageHalf + zc_used_;
  zc_used_ += padded;
  return d;
}
void* Stager::arenaOut(void* p, size_t bytes) {
  const size_t padded = (bytes + 15) & ~(size_t)15;
  if (!arena_ || bytes > kStageSmall || out_used_ + padded > kStageHalf) return nullptr;
  arena_outs_.push_back({p, ctx_->stage_host + kStageHalf + out_used_, bytes});
  void* d = ctx_->stage_dev + kStageHalf + out_used_;
  out_used_ += padded;
  return d;
}
bool Stager::send() {
  if (arena_ && !sent_) {
    sent_ = true;
    if (in_used_ && cudaMemcpyAsync(ctx_->stage_dev, ctx_->stage_host, in_used_, cudaMemcpyHostToDevice, ctx_->stream) != cudaSuccess) failed_ = true;
  }
  return !failed_;
}
void* Stager::alloc(size_t bytes) {
  void* d = nullptr;
  if (cudaMallocAsync(&d, bytes ? bytes : 1, ctx_->stream) != cudaSuccess) {
    failed_ = true;
    return nullptr;
  }
  allocs_.push_back(d);
  return d;
}
void Stager::release() {
  for (void* d : allocs_) cudaFreeAsync(d, ctx_->stream);
  allocs_.clear();
  if (arena_) { ctx_->stage_busy = false; arena_ = false; }
}
int Stager::finish() {
  if (!send()) { release(); return SVO_FAIL(ctx_, SVO_ERR_OUT_OF_MEMORY, "staging allocation or copy failed"); }
  if (mem_ == SVO_MEM_HOST) {
    if (out_used_)
      SVO_CUDA_TRY(ctx_, cudaMemcpyAsync(ctx_->stage_host + kStageHalf, ctx_->stage_dev + kStageHalf, out_used_, cudaMemcpyDeviceToHost, ctx_->stream));
    for (const Out& o : outs_) SVO_CUDA_TRY(ctx_, cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, ctx_->stream));
    outs_.clear();
    for (void* d : allocs_) cudaFreeAsync(d, ctx_->stream);
    allocs_.clear();
    SVO_CUDA_TRY(ctx_, cudaStreamSynchronize(ctx_->stream));
    for (const Out& o : arena_outs_) memcpy(o.host, o.dev, o.bytes);
    arena_outs_.clear();
  }
  release();
  return SVO_OK;
}

int svoEnsureSideStreams(svo_cuda_ctx* ctx) {
  if (ctx->ev_fork) return SVO_OK;
  for (int i = 0; i < svo_cuda_ctx::kSideStreams; ++i) {
    SVO_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->side_stream[i], cudaStreamNonBlocking));
    SVO_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_join[i], cudaEventDisableTiming));
  }
  SVO_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  return SVO_OK;
}

extern "C" {

int svo_cuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int svo_cuda_sizeof(const char* n) {
  if (!n) return -1;
#define SVO_SZ(t) if (strcmp(n, #t) == 0) return (int)sizeof(t)
  SVO_SZ(svo_camera); SVO_SZ(svo_corner); SVO_SZ(svo_detector_options); SVO_SZ(svo_sparse_align_options);
  SVO_SZ(svo_align_prior); SVO_SZ(svo_align_result); SVO_SZ(svo_matcher_options); SVO_SZ(svo_feature);
  SVO_SZ(svo_match_out); SVO_SZ(svo_depth_filter_options);
  SVO_SZ(svo_reproj_map); SVO_SZ(svo_reprojector_options); SVO_SZ(svo_reproj_result); SVO_SZ(svo_reproj_stats);
  SVO_SZ(svo_pose_optimizer_options); SVO_SZ(svo_pose_opt_result); SVO_SZ(svo_stereo_result); SVO_SZ(svo_stereo_stats);
#undef SVO_SZ
  return -1;
}

int svo_cuda_ctx_create(int device, svo_cuda_ctx** out) {
  if (!out) return SVO_ERR_INVALID_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return SVO_ERR_NO_DEVICE;  // no CPU fallback exists
  if (device < 0 || device >= n) return SVO_ERR_INVALID_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return SVO_ERR_CUDA;
  svo_cuda_ctx* c = new svo_cuda_ctx();
  c->device = device;
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return SVO_ERR_CUDA;
  }
  c->stream = c->own_stream;
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  // keep freed staging buffers in the stream-ordered pool instead of returning them to the driver
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  *out = c;
  return SVO_OK;
}

int svo_cuda_ctx_destroy(svo_cuda_ctx* ctx) {
  if (!ctx) return SVO_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->angle_bins) cudaFree(ctx->angle_bins);
  if (ctx->stage_dev) cudaFree(ctx->stage_dev);
  if (ctx->stage_host) cudaFreeHost(ctx->stage_host);
  for (int i = 0; i < svo_cuda_ctx::kSideStreams; ++i) {
    if (ctx->side_stream[i]) { cudaStreamSynchronize(ctx->side_stream[i]); cudaStreamDestroy(ctx->side_stream[i]); }
    if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return SVO_OK;
}

int svo_cuda_ctx_set_stream(svo_cuda_ctx* ctx, void* s) {
  if (!ctx) return SVO_ERR_INVALID_ARG;
  ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
  return SVO_OK;
}

int svo_cuda_ctx_synchronize(svo_cuda_ctx* ctx) {
  if (!ctx) return SVO_ERR_INVALID_ARG;
  SVO_BIND(ctx);
  SVO_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return SVO_OK;
}

int svo_cuda_host_alloc(svo_cuda_ctx* ctx, size_t bytes, int write_combined, void** out) {
  if (!ctx || !out) return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_host_alloc: bad arguments");
  *out = nullptr;
  SVO_BIND(ctx);
  const unsigned flags = cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0u);
  if (cudaHostAlloc(out, bytes ? bytes : 1, flags) != cudaSuccess) {
    cudaGetLastError();
    return SVO_FAIL(ctx, SVO_ERR_OUT_OF_MEMORY, "svo_cuda_host_alloc: cudaHostAlloc failed");
  }
  return SVO_OK;
}

int svo_cuda_host_free(svo_cuda_ctx* ctx, void* ptr) {
  if (!ptr) return SVO_OK;
  if (ctx) cudaSetDevice(ctx->device);
  return cudaFreeHost(ptr) == cudaSuccess ? SVO_OK : SVO_ERR_CUDA;
}

const char* svo_cuda_last_error(const svo_cuda_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }
long long svo_cuda_launch_count(const svo_cuda_ctx* ctx) { return ctx ? ctx->launches : 0; }

int svo_cuda_grid_cells(int width, int height, int cell_size, int* n_cols, int* n_rows) {
  if (cell_size <= 0 || width <= 0 || height <= 0) return SVO_ERR_INVALID_ARG;
  const int c = (width + cell_size - 1) / cell_size, r = (height + cell_size - 1) / cell_size;
  if (n_cols) *n_cols = c;
  if (n_rows) *n_rows = r;
  return c * r;
}

int svo_cuda_pyr_create(svo_cuda_ctx* ctx, int n_frames, int width, int height, int n_levels, int halfsample_mode,
                        svo_cuda_pyr** out) {
  if (!ctx || !out || n_frames <= 0 || width <= 0 || height <= 0 || n_levels <= 0 || n_levels > SVO_MAX_LEVELS)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_pyr_create: bad arguments");
  cudaSetDevice(ctx->device);
  svo_cuda_pyr* p = new svo_cuda_pyr();
  p->n_frames = n_frames;
  p->n_levels = n_levels;
  p->halfsample_mode = halfsample_mode;
  int c = width, r = height;
  for (int l = 0; l < n_levels; ++l) {
    if (c <= 0 || r <= 0) {
      delete p;
      return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_pyr_create: image too small for n_levels");
    }
    p->cols[l] = c;
    p->rows[l] = r;
    p->pitch[l] = (size_t(c) + 15) & ~size_t(15);
    p->frame_stride[l] = (p->pitch[l] * r + 255) & ~size_t(255);
    c /= 2;
    r /= 2;
  }
  for (int l = 0; l < n_levels; ++l) {
    const size_t bytes = p->frame_stride[l] * n_frames + 256;  // slack: row loaders read whole aligned words
    if (cudaMalloc(&p->data[l], bytes) != cudaSuccess) {
      for (int k = 0; k < l; ++k) cudaFree(p->data[k]);
      delete p;
      return SVO_FAIL(ctx, SVO_ERR_OUT_OF_MEMORY, "svo_cuda_pyr_create: cudaMalloc failed");
    }
    cudaMemsetAsync(p->data[l], 0, bytes, ctx->stream);
  }
  *out = p;
  return SVO_OK;
}

int svo_cuda_pyr_destroy(svo_cuda_ctx* ctx, svo_cuda_pyr* pyr) {
  if (!pyr) return SVO_ERR_INVALID_ARG;
  if (ctx) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
  }
  for (int l = 0; l < pyr->n_levels; ++l) cudaFree(pyr->data[l]);
  delete pyr;
  return SVO_OK;
}

int svo_cuda_pyr_upload(svo_cuda_ctx* ctx, svo_cuda_pyr* pyr, int first, int count, const uint8_t* src, size_t src_pitch,
                        size_t src_frame_stride, svo_mem mem) {
  if (!ctx || !pyr || !src || first < 0 || count < 0 || first + count > pyr->n_frames || src_pitch < size_t(pyr->cols[0]))
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_pyr_upload: bad arguments");
  SVO_BIND(ctx);
  const cudaMemcpyKind kind = mem == SVO_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  const size_t w = pyr->cols[0], h = pyr->rows[0];
  if (src_pitch == w && pyr->pitch[0] == w && src_frame_stride == w * h && pyr->frame_stride[0] == w * h) {
    // both sides are tightly packed: one linear copy runs at full PCIe / HBM copy speed (a 2D copy of 752-byte rows does not)
    SVO_CUDA_TRY(ctx, cudaMemcpyAsync(pyr->data[0] + pyr->frame_stride[0] * first, src, w * h * count, kind, ctx->stream));
  } else if (src_frame_stride == src_pitch * h && pyr->frame_stride[0] == pyr->pitch[0] * h) {
    // frames are back to back on both sides: one 2D copy of count*h rows
    SVO_CUDA_TRY(ctx, cudaMemcpy2DAsync(pyr->data[0] + pyr->frame_stride[0] * first, pyr->pitch[0], src, src_pitch, w,
                                        h * count, kind, ctx->stream));
  } else {
    for (int i = 0; i < count; ++i)
      SVO_CUDA_TRY(ctx, cudaMemcpy2DAsync(pyr->data[0] + pyr->frame_stride[0] * (first + i), pyr->pitch[0],
                                          src + src_frame_stride * i, src_pitch, w, h, kind, ctx->stream));
  }
  return SVO_OK;
}

int svo_cuda_pyr_download(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int frame, int level, uint8_t* dst, size_t dst_pitch,
                          svo_mem mem) {
  if (!ctx || !pyr || !dst || frame < 0 || frame >= pyr->n_frames || level < 0 || level >= pyr->n_levels ||
      dst_pitch < size_t(pyr->cols[level]))
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_pyr_download: bad arguments");
  SVO_BIND(ctx);
  const cudaMemcpyKind kind = mem == SVO_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  SVO_CUDA_TRY(ctx, cudaMemcpy2DAsync(dst, dst_pitch, pyr->data[level] + pyr->frame_stride[level] * frame, pyr->pitch[level],
                                      pyr->cols[level], pyr->rows[level], kind, ctx->stream));
  if (mem == SVO_MEM_HOST) SVO_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return SVO_OK;
}

int svo_cuda_pyr_level_info(const svo_cuda_pyr* pyr, int level, int* cols, int* rows, size_t* pitch, size_t* frame_stride,
                            void** device_ptr) {
  if (!pyr || level < 0 || level >= pyr->n_levels) return SVO_ERR_INVALID_ARG;
  if (cols) *cols = pyr->cols[level];
  if (rows) *rows = pyr->rows[level];
  if (pitch) *pitch = pyr->pitch[level];
  if (frame_stride) *frame_stride = pyr->frame_stride[level];
  if (device_ptr) *device_ptr = pyr->data[level];
  return SVO_OK;
}

}  // extern "C"
