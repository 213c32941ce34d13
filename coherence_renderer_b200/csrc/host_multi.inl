// host_multi.inl — part of coherence_b200.cu: several GPUs of one box from ONE host process (coh_multi_*), and the
// CUDA-IPC export / import of framebuffers for hosts that run one process per GPU.
//
// A frame shards by horizontal scanline bands: every device holds the whole (small) scene and renders
// the rows of its band; the only exchange is the gather of the RGBA8 strips, and it is fused into the rendering
// kernels — the framebuffers are peer-mapped over NVLink and every finished pixel is stored to all of them as it is
// produced (coh_fb_set_peers), so after the frame every device holds the whole picture and no collective follows.
// One worker thread per device issues that device's launches, so the host-side launch chains of the devices overlap.
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

struct coh_multi {
  int n = 0;
  std::vector<int> dev;
  std::vector<coh_ctx*> ctx;
  std::vector<uint32_t*> fb;          // cudaMalloc'ed (peer-accessible) framebuffers, one per device
  std::vector<int> cuts;              // band k = rows [cuts[k], cuts[k + 1])
  int W = 0, H = 0;
  std::string err;
  // workers
  std::vector<std::thread> th;
  std::mutex mu; std::condition_variable cv_job, cv_done;
  std::vector<std::function<int()>> job; std::vector<int> rc; std::vector<char> pending;
  bool quit = false;
};
struct MultiScene { std::vector<coh_scene_t> per_dev; };

static void multi_worker(coh_multi* m, int i) {
  cudaSetDevice(m->dev[i]);
  for (;;) {
    std::function<int()> f;
    {
      std::unique_lock<std::mutex> lk(m->mu);
      m->cv_job.wait(lk, [&] { return m->quit || m->pending[i]; });
      if (m->quit) return;
      f = m->job[i];
    }
    const int r = f();
    {
      std::lock_guard<std::mutex> lk(m->mu);
      m->rc[i] = r; m->pending[i] = 0;
    }
    m->cv_done.notify_all();
  }
}
// run fn(i) on every device's worker and wait until all have returned; the first failure is reported
static int multi_run(coh_multi* m, const std::function<int(int)>& fn) {
  {
    std::lock_guard<std::mutex> lk(m->mu);
    for (int i = 0; i < m->n; i++) { m->job[i] = [fn, i] { return fn(i); }; m->pending[i] = 1; }
  }
  m->cv_job.notify_all();
  std::unique_lock<std::mutex> lk(m->mu);
  m->cv_done.wait(lk, [&] { for (int i = 0; i < m->n; i++) if (m->pending[i]) return false; return true; });
  for (int i = 0; i < m->n; i++)
    if (m->rc[i]) { m->err = "device " + std::to_string(m->dev[i]) + ": " + coh_last_error(m->ctx[i]); return 1; }
  return 0;
}
#define MFAIL(msg) do { m->err = (msg); return 1; } while (0)
#define MCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { m->err = std::string(#call) + " failed: " + cudaGetErrorString(e_); return 1; } } while (0)

extern "C" {
static std::string g_multi_err;
const char* coh_multi_last_error(coh_multi* m) { return m ? m->err.c_str() : g_multi_err.c_str(); }
int coh_multi_device_count(coh_multi* m) { return m->n; }
coh_ctx* coh_multi_ctx(coh_multi* m, int32_t i) { return (i >= 0 && i < m->n) ? m->ctx[i] : nullptr; }

int coh_multi_shutdown(coh_multi* m) {
  if (!m) return 0;
  { std::lock_guard<std::mutex> lk(m->mu); m->quit = true; }
  m->cv_job.notify_all();
  for (auto& t : m->th) if (t.joinable()) t.join();
  for (int i = 0; i < (int)m->ctx.size(); i++) {
    if (!m->ctx[i]) continue;
    cudaSetDevice(m->dev[i]);
    coh_fb_set_peers(m->ctx[i], 0, nullptr);
    if (m->ctx[i]->fb && !m->ctx[i]->own_fb) { cudaStreamSynchronize(m->ctx[i]->stream); m->ctx[i]->fb = nullptr; m->ctx[i]->own_fb = true; }
    coh_shutdown(m->ctx[i]);
    if (i < (int)m->fb.size() && m->fb[i]) cudaFree(m->fb[i]);
  }
  delete m;
  return 0;
}
int coh_multi_init(int32_t n_devices, const int32_t* device_ids, coh_multi** out) {
  *out = nullptr;
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) { g_multi_err = "coh_multi_init: no CUDA device available; there is no CPU fallback"; return 1; }
  if (n_devices <= 0 || n_devices > COH_MAX_PEERS + 1 || n_devices > have) { g_multi_err = "coh_multi_init: 1 .. 8 devices of one box (and no more than are visible)"; return 1; }
  coh_multi* m = new coh_multi();
  m->n = n_devices;
  for (int i = 0; i < n_devices; i++) m->dev.push_back(device_ids ? device_ids[i] : i);
  m->ctx.assign(n_devices, nullptr); m->fb.assign(n_devices, nullptr);
  for (int i = 0; i < n_devices; i++)
    if (coh_init(m->dev[i], &m->ctx[i])) { g_multi_err = std::string("coh_multi_init: ") + coh_last_error(nullptr); coh_multi_shutdown(m); return 1; }
  // every device reads / writes every other device's framebuffer over NVLink
  for (int i = 0; i < n_devices; i++)
    for (int j = 0; j < n_devices; j++) {
      if (i == j) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, m->dev[i], m->dev[j]);
      if (!can) { g_multi_err = "coh_multi_init: devices " + std::to_string(m->dev[i]) + " and " + std::to_string(m->dev[j]) + " have no peer access"; coh_multi_shutdown(m); return 1; }
      cudaSetDevice(m->dev[i]);
      cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[j], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { g_multi_err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); coh_multi_shutdown(m); return 1; }
      cudaGetLastError();
    }
  m->job.resize(n_devices); m->rc.assign(n_devices, 0); m->pending.assign(n_devices, 0);
  for (int i = 0; i < n_devices; i++) m->th.emplace_back(multi_worker, m, i);
  *out = m;
  return 0;
}
// Framebuffer of width x height on every device; band k = rows [cuts[k], cuts[k + 1]) (cuts = NULL: equal bands,
// band k of N owns rows [floor (k H / N), floor ((k + 1) H / N))).
int coh_multi_configure(coh_multi* m, int32_t width, int32_t height, const int32_t* cuts) {
  if (width <= 0 || height <= 0) MFAIL("coh_multi_configure: bad size");
  m->cuts.resize(m->n + 1);
  for (int k = 0; k <= m->n; k++) m->cuts[k] = cuts ? cuts[k] : (int)((long long)k * height / m->n);
  if (m->cuts[0] != 0 || m->cuts[m->n] != height) MFAIL("coh_multi_configure: the bands must cover rows 0 .. height");
  for (int k = 0; k < m->n; k++) if (m->cuts[k] > m->cuts[k + 1]) MFAIL("coh_multi_configure: band cuts must not decrease");
  const bool resize = width != m->W || height != m->H;
  for (int i = 0; i < m->n; i++) {
    cudaSetDevice(m->dev[i]);
    if (resize) {
      if (m->fb[i]) { coh_fb_set_peers(m->ctx[i], 0, nullptr); cudaStreamSynchronize(m->ctx[i]->stream); cudaFree(m->fb[i]); m->fb[i] = nullptr; }
      MCK(cudaMalloc(&m->fb[i], sizeof(uint32_t) * (size_t)width * height));
      MCK(cudaMemset(m->fb[i], 0, sizeof(uint32_t) * (size_t)width * height));
    }
  }
  m->W = width; m->H = height;
  for (int i = 0; i < m->n; i++) {
    if (coh_fb_configure(m->ctx[i], width, height, m->cuts[i], m->cuts[i + 1]) || coh_fb_attach(m->ctx[i], m->fb[i])) MFAIL(std::string("coh_multi_configure: ") + coh_last_error(m->ctx[i]));
    void* peers[COH_MAX_PEERS]; int np = 0;
    for (int j = 0; j < m->n; j++) if (j != i) peers[np++] = m->fb[j];
    if (coh_fb_set_peers(m->ctx[i], np, peers)) MFAIL(coh_last_error(m->ctx[i]));
  }
  return 0;
}
int coh_multi_scene_create(coh_multi* m, const coh_object* objs, int32_t n_objs, int32_t n_background, const int32_t* edges, int32_t n_edges,
                           const int32_t* points, int32_t n_points, coh_scene_t* out) {
  *out = 0;
  MultiScene* s = new MultiScene();
  s->per_dev.assign(m->n, 0);
  if (multi_run(m, [&](int i) { return coh_scene_create(m->ctx[i], objs, n_objs, n_background, edges, n_edges, points, n_points, &s->per_dev[i]); })) {
    for (int i = 0; i < m->n; i++) if (s->per_dev[i]) coh_scene_free(m->ctx[i], s->per_dev[i]);
    delete s;
    return 1;
  }
  *out = (coh_scene_t)s;
  return 0;
}
int coh_multi_scene_free(coh_multi* m, coh_scene_t scene) {
  MultiScene* s = (MultiScene*)scene;
  if (!s) return 0;
  int rc = multi_run(m, [&](int i) { return coh_scene_free(m->ctx[i], s->per_dev[i]); });
  delete s;
  return rc;
}
// Render.render_frame over update = Sprite.box ux uy uw uh, every device its band; returns when all launches are
// issued.  After coh_multi_sync every device's framebuffer holds the WHOLE frame (peer stores over NVLink).
int coh_multi_render_frame(coh_multi* m, coh_scene_t scene, int32_t ux, int32_t uy, int32_t uw, int32_t uh, int32_t flags) {
  MultiScene* s = (MultiScene*)scene;
  if (!s) MFAIL("coh_multi_render_frame: null scene");
  if (!m->W) MFAIL("coh_multi_render_frame: call coh_multi_configure first");
  return multi_run(m, [&](int i) { return coh_render_frame(m->ctx[i], s->per_dev[i], ux, uy, uw, uh, flags); });
}
int coh_multi_scene_translate_object(coh_multi* m, coh_scene_t scene, int32_t obj_index, int32_t dx, int32_t dy) {
  MultiScene* s = (MultiScene*)scene;
  if (!s) MFAIL("coh_multi_scene_translate_object: null scene");
  return multi_run(m, [&](int i) { return coh_scene_translate_object(m->ctx[i], s->per_dev[i], obj_index, dx, dy); });
}
int coh_multi_sync(coh_multi* m) { return multi_run(m, [&](int i) { return coh_sync(m->ctx[i]); }); }
// A rectangle of the finished frame from device 0 (which holds every band): RGBA8 / RGB888 as coh_fb_read_*.
int coh_multi_fb_read_rgba(coh_multi* m, int32_t x, int32_t y, int32_t w, int32_t h, uint8_t* out) {
  if (coh_multi_sync(m)) return 1;
  if (coh_fb_read_rgba(m->ctx[0], x, y, w, h, out)) MFAIL(coh_last_error(m->ctx[0]));
  return 0;
}
int coh_multi_fb_read_rgb888(coh_multi* m, int32_t x, int32_t y, int32_t w, int32_t h, uint8_t* out) {
  if (coh_multi_sync(m)) return 1;
  if (coh_fb_read_rgb888(m->ctx[0], x, y, w, h, out)) MFAIL(coh_last_error(m->ctx[0]));
  return 0;
}

// ---- one process per GPU: framebuffers cross process boundaries as CUDA IPC handles (64 bytes each) ----
// The framebuffer must be one this library allocated with cudaMalloc for the purpose (coh_fb_alloc_shared); the pool
// allocations behind coh_fb_configure cannot be exported.
// Behind its W x H pixels a shared framebuffer carries COH_SIGNAL_SLOTS 32-bit frame counters: the cross-process
// "this band has landed" / "this frame is consumed" signals of coh_frame_signal / coh_frame_wait.
static size_t shared_fb_bytes(const coh_ctx* ctx) { return sizeof(uint32_t) * ((size_t)ctx->fr.W * ctx->fr.H + COH_SIGNAL_SLOTS); }
int coh_fb_alloc_shared(coh_ctx* ctx, uint8_t handle_out[64]) {
  CK(cudaSetDevice(ctx->device));
  if (!ctx->fr.W) FAIL("coh_fb_alloc_shared: call coh_fb_configure first");
  uint32_t* p = nullptr;
  CK(cudaMalloc(&p, shared_fb_bytes(ctx)));
  CK(cudaMemset(p, 0, shared_fb_bytes(ctx)));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, p));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  memcpy(handle_out, &h, 64);
  if (coh_fb_attach(ctx, p)) return 1;
  ctx->shared_fb = p;
  return 0;
}
int coh_fb_open_peer(coh_ctx* ctx, const uint8_t handle[64], void** device_ptr_out) {
  CK(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CK(cudaIpcOpenMemHandle(device_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  ctx->opened_peers.push_back(*device_ptr_out);
  return 0;
}
}  // extern "C"

// ---- frame signals between processes (one per GPU): counters in the shared framebuffers, written over NVLink ----
struct SignalTargets { int* flag[COH_MAX_PEERS + 1]; int n; };
struct SignalSlots { int slot[COH_SIGNAL_SLOTS]; int n; };
// Runs behind the frame's kernels in the stream: their stores (local and peer) are complete; the fence orders them
// before the counter for observers on other devices, the counter is released at system scope.
__global__ void k_frame_signal(SignalTargets T, int epoch) {
  if (threadIdx.x < T.n) {
    __threadfence_system();
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(T.flag[threadIdx.x]), "r"(epoch) : "memory");
  }
}
// One thread per awaited counter; gives up after ~4 s (a peer that died must not hang the GPU) and raises the error flag.
__global__ void k_frame_wait(const int* flags, SignalSlots L, int epoch, int* error_flag) {
  if (threadIdx.x < L.n) {
    const int* f = flags + L.slot[threadIdx.x];
    unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      int v; asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (v - epoch >= 0) break;   // (counters may wrap)
      __nanosleep(200);
      unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > 4000000000ull) { *error_flag = 5; break; }
    }
  }
}
extern "C" {
// coh_frame_signal: after everything issued so far on the context's stream, counter `slot` of every target framebuffer
// (this context's own shared framebuffer, or pointers from coh_fb_open_peer) becomes `epoch`.
int coh_frame_signal(coh_ctx* ctx, int32_t n_targets, void* const* target_fbs, int32_t slot, int32_t epoch) {
  CK(cudaSetDevice(ctx->device));
  if (!ctx->fr.W) FAIL("coh_frame_signal: call coh_fb_configure first");
  if (slot < 0 || slot >= COH_SIGNAL_SLOTS) FAIL("coh_frame_signal: slot out of range");
  if (n_targets < 0 || n_targets > COH_MAX_PEERS + 1) FAIL("coh_frame_signal: too many targets");
  if (n_targets == 0) return 0;
  SignalTargets T; memset(&T, 0, sizeof T); T.n = n_targets;
  for (int k = 0; k < n_targets; k++) {
    void* p = target_fbs[k];
    const bool known = (p && p == (void*)ctx->shared_fb) || std::find(ctx->opened_peers.begin(), ctx->opened_peers.end(), p) != ctx->opened_peers.end();
    if (!known) FAIL("coh_frame_signal: a target is neither this context's shared framebuffer nor a pointer from coh_fb_open_peer");
    T.flag[k] = (int*)((uint32_t*)p + (size_t)ctx->fr.W * ctx->fr.H) + slot;
  }
  k_frame_signal<<<1, 32, 0, ctx->stream>>>(T, epoch); LAUNCHED();
  return 0;
}
// coh_frame_wait: the context's stream waits (on the device, no host round trip) until the listed counters of its own shared
// framebuffer have reached `epoch`.
int coh_frame_wait(coh_ctx* ctx, int32_t n_slots, const int32_t* slots, int32_t epoch) {
  CK(cudaSetDevice(ctx->device));
  if (!ctx->shared_fb || ctx->fb != ctx->shared_fb) FAIL("coh_frame_wait: the framebuffer is not a shared one (coh_fb_alloc_shared)");
  if (n_slots < 0 || n_slots > COH_SIGNAL_SLOTS) FAIL("coh_frame_wait: too many slots");
  if (n_slots == 0) return 0;
  SignalSlots L; memset(&L, 0, sizeof L); L.n = n_slots;
  for (int k = 0; k < n_slots; k++) { if (slots[k] < 0 || slots[k] >= COH_SIGNAL_SLOTS) FAIL("coh_frame_wait: slot out of range"); L.slot[k] = slots[k]; }
  k_frame_wait<<<1, 32, 0, ctx->stream>>>((const int*)(ctx->shared_fb + (size_t)ctx->fr.W * ctx->fr.H), L, epoch, ctx->d_error); LAUNCHED();
  return 0;
}
}  // extern "C"
