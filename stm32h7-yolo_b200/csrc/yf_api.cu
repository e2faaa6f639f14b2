// yf_api.cu -- the C-ABI of libyoloface_b200.so: X-CUBE-AI style entry points (include/network.h,
// include/network_data.h) plus the B200 extensions (include/yoloface_b200.h), on top of the plan
// (yf_plan.cc) and the sm_100a kernels (yf_kernels.cu).  No CPU execution path exists: without a
// usable CUDA device ai_network_create() fails with AI_ERROR_CREATE_FAILED.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <thread>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/network.h"
#include "../../include/network_data.h"
#include "../../include/yoloface_b200.h"
#include "yf_kernels.cuh"
#include "yf_plan.h"
#include "yf_requant.cuh"

// the .tflite the reference deploys (yoloface/tflite/yoloface_int8.tflite), embedded at build time
extern "C" const unsigned char yf_embedded_model[];
extern "C" const unsigned int yf_embedded_model_len;

namespace {

using namespace yf;

thread_local std::string g_text;
void set_text(const std::string& s) { g_text = s; }

// One compiled resolution: plan + device-side constants + arena.
struct PlanDev {
  Plan plan;
  uint32_t cap = 0;                     // images per chunk the arena holds
  bool observer = false;
  uint8_t* d_wblob = nullptr;
  uint8_t* d_luts = nullptr;
  uint8_t* d_arena = nullptr;           // activation arena of lane 0 (= d_arena_l[0])
  // Layer-by-layer path: independent chunks run side by side on up to kLayerLanes streams, each with its own arena and
  // TMA descriptors (allocated on first use), and every (input, output, images, lane) combination is captured once in
  // a CUDA graph so that a chunk costs one graph launch instead of 26 kernel launches.
  static constexpr int kLayerLanes = 4;
  uint8_t* d_arena_l[kLayerLanes] = {};
  std::vector<CUtensorMap> tmaps_l[kLayerLanes];
  struct GraphKey { const void* in; void* out; uint32_t nb; int lane;
                    bool operator<(const GraphKey& o) const { return std::tie(in, out, nb, lane) < std::tie(o.in, o.out, o.nb, o.lane); } };
  std::map<GraphKey, cudaGraphExec_t> graphs;
  EpiCh* d_epi = nullptr;               // this plan's requant tables (general and 16-byte form)
  EpiChF* d_epif = nullptr;
  int8_t* d_in = nullptr;               // staging of the synchronous paths for host inputs [cap,H,W,3] (never a ring slot)
  int8_t* d_head = nullptr;             // staging of the synchronous paths for heads [cap,GH,GW,C]
  static constexpr int kRing = 6;       // pipelined host path: slots of (input, head) staging + events
  int8_t* r_in[kRing] = {};
  int8_t* r_head[kRing] = {};
  cudaEvent_t ev_h2d[kRing] = {}, ev_comp[kRing] = {}, ev_d2h[kRing] = {};
  bool busy[kRing] = {};
  bool ring_ready = false;              // set only once every slot and event exists
  uint64_t seq = 0;
  std::vector<CUtensorMap> tmaps;       // per step (conv1x1 only)
  std::vector<float> step_ms;
  Plan fplan;                           // same steps with 16-aligned concat slots, for the fused kernel
  FusedProgram fprog;
  uint8_t* d_fparams = nullptr;
  FusedPhase* d_fphases = nullptr;      // this plan's phase descriptors (global memory: nothing is shared between plans)
  bool fused_spec = false;              // the plan is the one the specialised kernel was generated from
  // the same plan laid out for the latency shape (512-thread CTAs, one per SM): launches of at most one image per SM
  // that run alone.  Exists only where the specialised latency kernel does (deployed model at its own resolution).
  FusedProgram fprog_lat;
  uint8_t* d_fparams_lat = nullptr;
  FusedPhase* d_fphases_lat = nullptr;
  bool fused_lat = false;
  // and for the cluster shape (the front phases of one image shared by a cluster of CTAs over distributed shared memory):
  // launches of so few images that every image can have a cluster of its own
  FusedProgram fprog_cl;
  uint8_t* d_fparams_cl = nullptr;
  FusedPhase* d_fphases_cl = nullptr;
  int cluster_images = 0;               // images one launch of the cluster shape can take (0: shape unavailable)
  ~PlanDev() {
    for (auto& kv : graphs) cudaGraphExecDestroy(kv.second);
    for (int l = 1; l < kLayerLanes; ++l) cudaFree(d_arena_l[l]);
    cudaFree(d_epi); cudaFree(d_epif);
    cudaFree(d_wblob); cudaFree(d_luts); cudaFree(d_arena); cudaFree(d_in); cudaFree(d_head); cudaFree(d_fparams); cudaFree(d_fphases);
    cudaFree(d_fparams_lat); cudaFree(d_fphases_lat); cudaFree(d_fparams_cl); cudaFree(d_fphases_cl);
    for (int i = 0; i < kRing; ++i) { cudaFree(r_in[i]); cudaFree(r_head[i]); }
    for (int i = 0; i < kRing; ++i) { if (ev_h2d[i]) cudaEventDestroy(ev_h2d[i]); if (ev_comp[i]) cudaEventDestroy(ev_comp[i]); if (ev_d2h[i]) cudaEventDestroy(ev_d2h[i]); }
  }
};

// One persistent host thread per additional GPU of a multi-GPU context ("internally one worker
// thread + stream per GPU"): it owns the CUDA device selection of its thread and runs the jobs the caller's thread posts.
struct Worker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int32_t()> job;
  int32_t result = 0;
  std::string text;                     // thread-local error text of the job, handed back to the caller
  bool pending = false, quit = false;
  explicit Worker(int device);
  ~Worker() { { std::lock_guard<std::mutex> lk(mu); quit = true; } cv.notify_all(); if (th.joinable()) th.join(); }
  void post(std::function<int32_t()> j) { { std::lock_guard<std::mutex> lk(mu); job = std::move(j); pending = true; } cv.notify_all(); }
  int32_t wait(std::string* t) { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return !pending; }); if (t) *t = text; return result; }
};

size_t head_bytes(const PlanDev* pd) { const Plan& P = pd->plan; return static_cast<size_t>(P.GH) * P.GW * P.buffers[P.output_buf].C; }

struct Network {
  bool initialized = false;
  ai_error err{AI_ERROR_NONE, AI_ERROR_CODE_NONE};
  TflModel model;
  std::vector<uint8_t> blob;            // weights handed to ai_network_init (ST layout)
  int device = 0, sm_count = 148;
  uint32_t chunk = 1024;
  bool observer = false, step_profiling = false;
  int mode = 0;                         // 0 auto (fused when possible), 1 layer-by-layer, 2 fused only
  bool st_act = false;                  // ST-style LeakyReLU tables instead of TFLite's
  bool custom_model = false;            // the model came from a file (tflite_path / YF_B200_TFLITE), not from the library
  int H = 56, W = 56;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // Pipeline error word: mapped pinned host memory.  Kernels write it (atomicCAS across PCIe) only when a bounded wait
  // gives up, the host reads it after any synchronisation -- no copy on any path.
  int* h_err = nullptr;                 // host view
  int* d_err = nullptr;                 // device view of the same word
  long long* d_trace = nullptr; bool trace_on = false;
  float* d_dets = nullptr; int* d_counts = nullptr; size_t dets_cap = 0; uint32_t dets_max = 0;
  float anchors[6] = {9.f, 14.f, 12.f, 17.f, 22.f, 21.f};   // yoloface.c:20 (yf_b200_set_decode_params overrides)
  float stride = 0.f;                   // 0: input height / head rows (8 for this model)
  uint8_t* d_frames = nullptr; size_t frames_cap = 0;
  // small host batches: the kernels read the images from / write the heads to mapped pinned memory (no DMA copies)
  int8_t* h_small = nullptr; size_t small_cap = 0;      // [in | out | completion words], device-visible at the same address (UVA)
  uint32_t* done_words = nullptr; uint32_t done_seq = 0; int done_grid = 0;   // armed by the small-batch path around one launch
  std::map<std::pair<int, int>, std::unique_ptr<PlanDev>> plans;
  cudaStream_t own_stream = nullptr;    // created by the library; `stream` may be a caller's
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;   // copy streams of the pipelined host path
  cudaStream_t s_h2d_b = nullptr;       // second host-to-device stream: consecutive chunks alternate, so one copy's set-up hides behind the other's transfer
  int h2d_streams = 2;                  // YF_B200_H2D_STREAMS=1 restores the single stream (256-image steps: 4.94 -> 5.10 M img/s sustained with two)
  // Kernel lanes: one fused launch covers 256 of the GPU's 296 CTA slots for one image latency, so
  // independent chunks alternate over two streams and the head of one overlaps the tail of the other.
  static constexpr int kLanes = 8;      // streams created; `lanes` of them are used (YF_B200_LANES, default below)
  int lanes = 8;                        // 20 queued 256-image batches: 4 lanes 6.37 M img/s, 6 lanes 6.58 M, 8 lanes 6.62 M (profiles/r02_lanes_pairing.txt)
  cudaStream_t lane[kLanes] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[kLanes] = {};
  uint64_t lane_seq = 0;
  uint32_t last_run_n = 0;
  ai_buffer rep_in{}, rep_out{};        // I/O descriptors handed out by ai_network_get_report
  uint64_t launches = 0, images = 0;
  uint32_t lat_launches = 0;            // fused launches that took the latency shape
  uint32_t cl_launches = 0;             // ... of which with a cluster per image
  float last_ms = 0.f;
  // Multi-GPU context (yf_b200_config.device_mask / YF_B200_DEVICES): members[0] is this object, the others are
  // contexts on the other GPUs owned by it and driven by workers[i]; a member's `owner` points back here.
  std::vector<Network*> members;
  std::vector<std::unique_ptr<Worker>> workers;
  Network* owner = nullptr;
  void latch(int type, int code) { if (err.type == AI_ERROR_NONE) { err.type = type; err.code = code; } }
  // Everything the context owns on its device; shared by ai_network_destroy and the failure paths of ai_network_create
  // (the caller has selected n->device).  Safe on a partially constructed object: every handle starts out null.
  ~Network() {
    workers.clear();                                        // joins the threads before their contexts go away
    for (size_t i = 1; i < members.size(); ++i) { cudaSetDevice(members[i]->device); delete members[i]; }
    if (members.size() > 1) cudaSetDevice(device);
    members.clear();
    plans.clear();
    if (h_err) cudaFreeHost(h_err);
    if (h_small) cudaFreeHost(h_small);
    cudaFree(d_trace); cudaFree(d_dets); cudaFree(d_counts); cudaFree(d_frames);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (own_stream) cudaStreamDestroy(own_stream);
    if (s_h2d) cudaStreamDestroy(s_h2d);
    if (s_h2d_b) cudaStreamDestroy(s_h2d_b);
    if (s_d2h) cudaStreamDestroy(s_d2h);
    for (int l = 0; l < kLanes; ++l) { if (lane[l]) cudaStreamDestroy(lane[l]); if (ev_join[l]) cudaEventDestroy(ev_join[l]); }
    if (ev_fork) cudaEventDestroy(ev_fork);
  }
};

Worker::Worker(int device) {
  th = std::thread([this, device] {
    cudaSetDevice(device);
    for (;;) {
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return pending || quit; });
      if (quit) return;
      std::function<int32_t()> j = std::move(job);
      lk.unlock();
      g_text.clear();
      const int32_t r = j();
      lk.lock();
      result = r; text = g_text; pending = false;
      lk.unlock();
      cv.notify_all();
    }
  });
}

bool make_lanes(Network* n) {
  for (int l = 0; l < Network::kLanes; ++l)
    if (cudaStreamCreateWithFlags(&n->lane[l], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&n->ev_join[l], cudaEventDisableTiming) != cudaSuccess) return false;
  return cudaEventCreateWithFlags(&n->ev_fork, cudaEventDisableTiming) == cudaSuccess;
}

// Locking: g_mu guards the registry of live contexts only.  Host-side work on a device is serialised by that device's
// mutex (API calls are short: they queue work and return, or wait for their own streams); contexts on different GPUs
// driven from different host threads run in parallel.  No device-side state is shared between contexts -- phase
// descriptors, requant tables and weights are per plan in global memory -- so two models on one GPU interleave freely.
// One context must not be used from two threads at once (same as ST).
std::mutex g_mu;
std::mutex g_dev_mu[64];
// ST's runtime has one static context (g_network, network.c:36); a GPU process may want one per
// device or per configuration, so every create returns a fresh context and all stay valid.
std::vector<Network*> g_nets;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

bool cuda_ok(Network* n, cudaError_t e, const char* what, int type = AI_ERROR_INVALID_STATE, int code = AI_ERROR_CODE_NETWORK) {
  if (e == cudaSuccess) return true;
  set_text(std::string(what) + ": " + cudaGetErrorString(e));
  if (n) n->latch(type, code);
  return false;
}

Network* as_net(ai_handle h) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (Network* n : g_nets) if (n == h) return n;
  return nullptr;
}

bool is_device_ptr(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// device memory of another GPU than the context's (no peer mapping is set up by this library)
bool is_foreign_device_ptr(const Network* n, const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice && a.device != n->device;
}
bool is_pageable_ptr(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
}

// arena (cleared once: pad channels of every buffer must read as defined bytes) + TMA descriptors of one layer lane
bool prepare_layer_lane(Network* n, PlanDev* pd, int lane) {
  const Plan& P = pd->plan;
  const size_t per_img = pd->observer ? P.arena_bytes_per_image_observer : P.arena_bytes_per_image;
  if (!pd->d_arena_l[lane] && !cuda_ok(n, cudaMalloc(&pd->d_arena_l[lane], per_img * pd->cap + 1024), "cudaMalloc arena", AI_ERROR_ALLOCATION_FAILED, AI_ERROR_CODE_NETWORK_ACTIVATIONS)) return false;
  if (!pd->tmaps_l[lane].empty()) return true;
  if (!cuda_ok(n, cudaMemsetAsync(pd->d_arena_l[lane], 0, per_img * pd->cap + 1024, n->stream), "clear arena")) return false;
  // TMA descriptors of the 1x1-conv A operands: 2-D [rows = cap*H*W, CP bytes], box 128 rows x 16 B
  std::vector<CUtensorMap> tm(P.steps.size());
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_text("cuTensorMapEncodeTiled entry point unavailable"); n->latch(AI_ERROR_INIT_FAILED, AI_ERROR_CODE_NETWORK); return false; }
  for (size_t i = 0; i < P.steps.size(); ++i) {
    const Step& s = P.steps[i];
    if (s.kind != STEP_CONV1X1) continue;
    const PBuffer& b = P.buffers[s.in_buf];
    void* base = pd->d_arena_l[lane] + b.offset * pd->cap;
    cuuint64_t gdim[2] = {static_cast<cuuint64_t>(b.CP), static_cast<cuuint64_t>(pd->cap) * b.H * b.W};
    cuuint64_t gstr[1] = {static_cast<cuuint64_t>(b.CP)};
    cuuint32_t box[2] = {16, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_text("cuTensorMapEncodeTiled failed for step " + s.name + " (CUresult " + std::to_string(r) + ")");
      n->latch(AI_ERROR_INIT_FAILED, AI_ERROR_CODE_TENSOR); return false; }
  }
  if (!cuda_ok(n, cudaStreamSynchronize(n->stream), "clear arena")) return false;   // the lanes' streams use it next
  pd->tmaps_l[lane] = std::move(tm);
  return true;
}

// ---- plan instantiation ------------------------------------------------------------------
PlanDev* get_plan(Network* n, int H, int W) {
  auto key = std::make_pair(H, W);
  auto it = n->plans.find(key);
  if (it != n->plans.end() && it->second->observer == n->observer && it->second->cap == n->chunk) return it->second.get();
  if (it != n->plans.end()) {
    n->plans.erase(it);
  }
  std::unique_ptr<PlanDev> pd(new PlanDev);
  std::string perr;
  if (!build_plan(n->model, H, W, n->blob.empty() ? nullptr : n->blob.data(), n->blob.size(), &pd->plan, &perr, 4, n->st_act)) {
    set_text("plan: " + perr); n->latch(AI_ERROR_INIT_FAILED, AI_ERROR_CODE_NETWORK); return nullptr;
  }
  Plan& P = pd->plan;
  pd->cap = n->chunk; pd->observer = n->observer;
  const size_t per_img = n->observer ? P.arena_bytes_per_image_observer : P.arena_bytes_per_image;
  const int al = AI_ERROR_ALLOCATION_FAILED, ac = AI_ERROR_CODE_NETWORK_ACTIVATIONS;
  if (!cuda_ok(n, cudaMalloc(&pd->d_wblob, std::max<size_t>(P.wblob.size(), 16)), "cudaMalloc weights", al, AI_ERROR_CODE_NETWORK_WEIGHTS)) return nullptr;
  if (!cuda_ok(n, cudaMalloc(&pd->d_luts, std::max<size_t>(P.luts.size(), 256)), "cudaMalloc luts", al, ac)) return nullptr;
  if (!cuda_ok(n, cudaMalloc(&pd->d_arena, per_img * pd->cap + 1024), "cudaMalloc arena", al, ac)) return nullptr;
  if (!cuda_ok(n, cudaMalloc(&pd->d_in, static_cast<size_t>(H) * W * 3 * pd->cap), "cudaMalloc input staging", al, ac)) return nullptr;
  if (!cuda_ok(n, cudaMalloc(&pd->d_head, head_bytes(pd.get()) * pd->cap), "cudaMalloc head staging", al, ac)) return nullptr;
  if (!cuda_ok(n, cudaMemcpyAsync(pd->d_wblob, P.wblob.data(), P.wblob.size(), cudaMemcpyHostToDevice, n->stream), "upload weights")) return nullptr;
  if (!P.luts.empty() && !cuda_ok(n, cudaMemcpyAsync(pd->d_luts, P.luts.data(), P.luts.size(), cudaMemcpyHostToDevice, n->stream), "upload luts")) return nullptr;
  pd->d_arena_l[0] = pd->d_arena;
  if (!prepare_layer_lane(n, pd.get(), 0)) return nullptr;
  {
    const std::vector<EpiChF> lean = lean_epi_table(P.epi.data(), static_cast<int>(P.epi.size()));
    if (!cuda_ok(n, cudaMalloc(&pd->d_epi, std::max<size_t>(P.epi.size(), 1) * sizeof(EpiCh)), "cudaMalloc requant table", al, AI_ERROR_CODE_NETWORK_WEIGHTS)) return nullptr;
    if (!cuda_ok(n, cudaMalloc(&pd->d_epif, lean.size() * sizeof(EpiChF)), "cudaMalloc requant table", al, AI_ERROR_CODE_NETWORK_WEIGHTS)) return nullptr;
    if (!P.epi.empty() && !cuda_ok(n, cudaMemcpyAsync(pd->d_epi, P.epi.data(), P.epi.size() * sizeof(EpiCh), cudaMemcpyHostToDevice, n->stream), "upload requant table")) return nullptr;
    if (!cuda_ok(n, cudaMemcpyAsync(pd->d_epif, lean.data(), lean.size() * sizeof(EpiChF), cudaMemcpyHostToDevice, n->stream), "upload requant table")) return nullptr;
    if (!cuda_ok(n, cudaStreamSynchronize(n->stream), "upload requant table")) return nullptr;      // `lean` dies with this scope
  }
  pd->step_ms.assign(P.steps.size(), -1.f);
  // fused single-kernel program (falls back to the layered path when it cannot be built)
  std::string ferr;
  if (build_plan(n->model, H, W, n->blob.empty() ? nullptr : n->blob.data(), n->blob.size(), &pd->fplan, &ferr, 16, n->st_act) &&
      build_fused(pd->fplan, &pd->fprog) && pd->fprog.ok) {
    if (!cuda_ok(n, cudaMalloc(&pd->d_fparams, pd->fprog.params.size()), "cudaMalloc fused params", al, AI_ERROR_CODE_NETWORK_WEIGHTS)) return nullptr;
    if (!cuda_ok(n, cudaMemcpyAsync(pd->d_fparams, pd->fprog.params.data(), pd->fprog.params.size(), cudaMemcpyHostToDevice, n->stream), "upload fused params")) return nullptr;
    const size_t dbytes = pd->fprog.phases.size() * sizeof(FusedPhase);
    if (!cuda_ok(n, cudaMalloc(&pd->d_fphases, dbytes), "cudaMalloc fused descriptors", al, AI_ERROR_CODE_NETWORK_WEIGHTS)) return nullptr;
    if (!cuda_ok(n, cudaMemcpyAsync(pd->d_fphases, pd->fprog.phases.data(), dbytes, cudaMemcpyHostToDevice, n->stream), "upload fused descriptors")) return nullptr;
    static const bool spec_off = [] { const char* e = std::getenv("YF_B200_FUSED_SPEC"); return e && !std::atoi(e); }();
    pd->fused_spec = !spec_off && fused_spec_matches(pd->fprog);
    if (!cuda_ok(n, fused_init(pd->fprog, pd->fused_spec), "fused kernel attributes")) return nullptr;
    const char* lat_env = std::getenv("YF_B200_FUSED_LAT");      // diagnostics / tests: 0 keeps every launch on the throughput shape
    const bool lat_off = lat_env && !std::atoi(lat_env);
    if (pd->fused_spec && !lat_off && build_fused(pd->fplan, &pd->fprog_lat, kFusedLatThreads) && pd->fprog_lat.ok && fused_spec_matches(pd->fprog_lat)) {
      const FusedProgram& FL = pd->fprog_lat;
      if (!cuda_ok(n, cudaMalloc(&pd->d_fparams_lat, FL.params.size()), "cudaMalloc fused params", al, AI_ERROR_CODE_NETWORK_WEIGHTS)) return nullptr;
      if (!cuda_ok(n, cudaMemcpyAsync(pd->d_fparams_lat, FL.params.data(), FL.params.size(), cudaMemcpyHostToDevice, n->stream), "upload fused params")) return nullptr;
      if (!cuda_ok(n, cudaMalloc(&pd->d_fphases_lat, dbytes), "cudaMalloc fused descriptors", al, AI_ERROR_CODE_NETWORK_WEIGHTS)) return nullptr;
      if (!cuda_ok(n, cudaMemcpyAsync(pd->d_fphases_lat, FL.phases.data(), dbytes, cudaMemcpyHostToDevice, n->stream), "upload fused descriptors")) return nullptr;
      if (!cuda_ok(n, fused_init(FL, true), "fused kernel attributes")) return nullptr;
      pd->fused_lat = true;
      // The cluster shape is OPT-IN (YF_B200_FUSED_CLUSTER=1): measured on B200 it is slower than the plain latency shape
      // (one image 49 vs 42 us): a cluster barrier plus the remote stores cost ~1.2 k cycles per shared phase, more than
      // sharing the phase's work over four SMs saves.
      const char* cl_env = std::getenv("YF_B200_FUSED_CLUSTER");
      if (cl_env && std::atoi(cl_env) && build_fused(pd->fplan, &pd->fprog_cl, kFusedLatThreads, kFusedMaxCluster) && pd->fprog_cl.ok &&
          fused_spec_matches(pd->fprog_cl)) {
        const FusedProgram& FC = pd->fprog_cl;
        if (!cuda_ok(n, cudaMalloc(&pd->d_fparams_cl, FC.params.size()), "cudaMalloc fused params", al, AI_ERROR_CODE_NETWORK_WEIGHTS)) return nullptr;
        if (!cuda_ok(n, cudaMemcpyAsync(pd->d_fparams_cl, FC.params.data(), FC.params.size(), cudaMemcpyHostToDevice, n->stream), "upload fused params")) return nullptr;
        if (!cuda_ok(n, cudaMalloc(&pd->d_fphases_cl, dbytes), "cudaMalloc fused descriptors", al, AI_ERROR_CODE_NETWORK_WEIGHTS)) return nullptr;
        if (!cuda_ok(n, cudaMemcpyAsync(pd->d_fphases_cl, FC.phases.data(), dbytes, cudaMemcpyHostToDevice, n->stream), "upload fused descriptors")) return nullptr;
        if (!cuda_ok(n, fused_init(FC, true), "fused kernel attributes")) return nullptr;
        pd->cluster_images = fused_max_clusters(FC);
      }
    }
  } else {
    pd->fprog.ok = false;
    if (pd->fprog.why.empty()) pd->fprog.why = ferr;
  }
  if (n->mode == 2 && !pd->fprog.ok) { set_text("fused path unavailable: " + pd->fprog.why); n->latch(AI_ERROR_INIT_FAILED, AI_ERROR_CODE_NETWORK); return nullptr; }
  if (!cuda_ok(n, cudaStreamSynchronize(n->stream), "plan upload")) return nullptr;   // lanes read these buffers too
  PlanDev* raw = pd.get();
  n->plans[key] = std::move(pd);
  return raw;
}

int8_t* buf_ptr(const PlanDev* pd, int buf, const int8_t* in, int8_t* head, int lane = 0) {
  if (buf < 0) return nullptr;
  const PBuffer& b = pd->plan.buffers[buf];
  if (b.is_input) return const_cast<int8_t*>(in);
  if (b.is_output) return head;
  if (b.observer_only && !pd->observer) return nullptr;
  return reinterpret_cast<int8_t*>(pd->d_arena_l[lane] + b.offset * pd->cap);
}

EpiOut make_epi_out(const PlanDev* pd, const Step& s, const int8_t* in, int8_t* head, int lane = 0) {
  const Plan& P = pd->plan;
  EpiOut eo{};
  const PBuffer& ob = P.buffers[s.out_buf];
  eo.epi_tab = pd->d_epi; eo.epif_tab = pd->d_epif;
  eo.out = buf_ptr(pd, s.out_buf, in, head, lane); eo.out_pitch = ob.CP; eo.out_coff = s.out_coff; eo.cout = s.Cout;
  // the step owns the pad channels of a buffer it writes from channel 0 alone
  eo.fill_to = (s.out_coff == 0 && ob.C == s.Cout && !ob.is_output) ? ob.CP : s.Cout;
  eo.epi_base = s.epi_base;
  const bool obs = pd->observer;
  if (obs && s.raw_buf >= 0) { eo.raw = buf_ptr(pd, s.raw_buf, in, head, lane); eo.raw_pitch = P.buffers[s.raw_buf].CP; }
  if (obs && s.mid_buf >= 0) { eo.mid = buf_ptr(pd, s.mid_buf, in, head, lane); eo.mid_pitch = P.buffers[s.mid_buf].CP; }
  if (obs && s.pre_add_buf >= 0) { eo.pre_add = buf_ptr(pd, s.pre_add_buf, in, head, lane); eo.pre_add_pitch = P.buffers[s.pre_add_buf].CP; }
  if (s.add.enabled) {
    eo.add = s.add; eo.add_in = buf_ptr(pd, s.add_buf, in, head, lane); eo.add_pitch = P.buffers[s.add_buf].CP; eo.add_coff = s.add_coff;
  }
  if (obs) {
    eo.lut1 = s.lut1 >= 0 ? pd->d_luts + static_cast<size_t>(s.lut1) * 256 : nullptr;
    eo.lut2 = s.lut2 >= 0 ? pd->d_luts + static_cast<size_t>(s.lut2) * 256 : nullptr;
  } else {
    eo.lut1 = s.lut_fused >= 0 ? pd->d_luts + static_cast<size_t>(s.lut_fused) * 256 : nullptr;
    eo.lut2 = nullptr;
  }
  // lean conv epilogue: not observing, at most one of {table, ADD}, every channel in the lean requant form, and
  // 4-byte aligned skip operand / output slot (the ADD reads the skip tensor a word at a time)
  eo.fast = 0;
  if (!obs && (s.kind == STEP_CONV1X1 || s.kind == STEP_CONV_IM2COL || s.kind == STEP_DW) && !(eo.lut1 && s.add.enabled) &&
      (!s.add.enabled || (eo.add_coff % 4 == 0 && eo.add_pitch % 4 == 0))) {
    eo.fast = 1;
    for (int c = 0; c < s.Cout && eo.fast; ++c) { int32_t b; if (!epi_lean_form(P.epi[s.epi_base + c], &b)) eo.fast = 0; }
  }
  return eo;
}

// run the fused steps for nb images whose input is at d_in (device) writing heads to d_head (device)
// true when chunks of this plan go through the single fused kernel (no shared activation arena:
// independent chunks may then run concurrently)
bool uses_fused(const Network* n, const PlanDev* pd) { return !pd->observer && !n->step_profiling && n->mode != 1 && pd->fprog.ok; }

// the 26 (yoloface) layer kernels of one chunk, queued on `st`; with `timed`, every step between its own events
bool launch_layer_steps(Network* n, PlanDev* pd, const int8_t* d_in, int8_t* d_head, uint32_t nb, cudaStream_t st, int lane, bool timed) {
  const Plan& P = pd->plan;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (timed) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
  bool ok = true;
  for (size_t i = 0; i < P.steps.size() && ok; ++i) {
    const Step& s = P.steps[i];
    EpiOut eo = make_epi_out(pd, s, d_in, d_head, lane);
    eo.err_word = n->d_err;
    cudaError_t e = cudaSuccess;
    if (timed) cudaEventRecord(e0, st);
    switch (s.kind) {
      case STEP_CONV1X1: {
        Conv1x1Args a{};
        a.w_img = pd->d_wblob + s.w_off; a.w_bytes = static_cast<int>(s.w_bytes);
        a.nchunk = P.buffers[s.in_buf].CP / 16; a.nk = s.Kpad / 32;
        a.M = static_cast<long long>(nb) * s.Hout * s.Wout; a.num_tiles = static_cast<int>((a.M + 127) / 128);
        a.eo = eo; a.err = n->d_err;
        e = launch_conv1x1(pd->tmaps_l[lane][i], a, s.Npad, n->sm_count, st);
        break; }
      case STEP_CONV_IM2COL: {
        ConvIm2colArgs a{};
        a.in = d_in; a.w_img = pd->d_wblob + s.w_off; a.w_bytes = static_cast<int>(s.w_bytes);
        a.n_img = static_cast<int>(nb); a.Hin = s.Hin; a.Win = s.Win; a.Hout = s.Hout; a.Wout = s.Wout;
        a.band_rows = s.band_rows; a.bands = s.bands; a.in_zp = s.in_zp; a.eo = eo; a.err = n->d_err;
        e = launch_conv_im2col(a, s.Npad, n->sm_count, st);
        break; }
      case STEP_DW: {
        DwArgs a{};
        a.in = buf_ptr(pd, s.in_buf, d_in, d_head, lane); a.in_pitch = P.buffers[s.in_buf].CP;
        a.w1h = reinterpret_cast<const uint32_t*>(pd->d_wblob + s.w_off);
        a.n_img = static_cast<int>(nb); a.Hin = s.Hin; a.Win = s.Win; a.Hout = s.Hout; a.Wout = s.Wout;
        a.stride = s.stride; a.pad_t = s.pad_t; a.pad_l = s.pad_l; a.in_zp = s.in_zp; a.words = (s.Cout + 3) / 4; a.eo = eo;
        e = launch_dw(a, st);
        break; }
      case STEP_MAXPOOL: {
        PoolArgs a{};
        a.in = buf_ptr(pd, s.in_buf, d_in, d_head, lane); a.in_pitch = P.buffers[s.in_buf].CP;
        a.n_img = static_cast<int>(nb); a.Hin = s.Hin; a.Win = s.Win; a.Hout = s.Hout; a.Wout = s.Wout;
        a.k = s.kh; a.stride = s.stride; a.pad_t = s.pad_t; a.pad_l = s.pad_l; a.words = (s.Cout + 3) / 4; a.eo = eo;
        e = launch_pool(a, st);
        break; }
      case STEP_LUT: {
        LutArgs a{};
        a.in = buf_ptr(pd, s.in_buf, d_in, d_head, lane); a.in_pitch = P.buffers[s.in_buf].CP; a.in_coff = s.in_coff;
        a.rows = static_cast<long long>(nb) * s.Hout * s.Wout; a.words = (s.Cout + 3) / 4; a.eo = eo;
        e = launch_lut(a, st);
        break; }
    }
    if (!cuda_ok(n, e, s.name.c_str())) { ok = false; break; }
    if (timed) {
      cudaEventRecord(e1, st); cudaEventSynchronize(e1);
      float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); pd->step_ms[i] = ms;
    }
  }
  if (e0) { cudaEventDestroy(e0); cudaEventDestroy(e1); }
  return ok;
}

// overlapped: the caller queues further launches around this one (kernel lanes / the pipelined host ring)
// lane: which activation arena the layer kernels use (independent chunks on different lanes run side by side)
bool run_steps(Network* n, PlanDev* pd, const int8_t* d_in, int8_t* d_head, uint32_t nb, cudaStream_t st = nullptr, bool overlapped = false, int lane = 0) {
  const Plan& P = pd->plan;
  if (!st) st = n->stream;
  if (uses_fused(n, pd)) {
    FusedLaunch L{};
    L.d_in = d_in; L.d_out = d_head; L.d_params = pd->d_fparams; L.d_phases = pd->d_fphases; L.n_img = static_cast<int>(nb);
    L.sm_count = n->sm_count; L.d_err = n->d_err; L.stream = st; L.d_trace = n->trace_on ? n->d_trace : nullptr;
    L.use_spec = pd->fused_spec; L.overlapped = overlapped;
    L.d_done = n->done_words; L.done_seq = n->done_seq; L.grid_out = &n->done_grid;
    // a launch that runs alone with at most one image per SM: the latency shape (twice the warps on each image)
    const bool lat = pd->fused_lat && !overlapped && static_cast<int>(nb) <= n->sm_count;
    // ... and with so few images that each can have a cluster of CTAs: the cluster shape
    const bool cl = lat && static_cast<int>(nb) <= pd->cluster_images;
    if (cl) { L.d_params = pd->d_fparams_cl; L.d_phases = pd->d_fphases_cl; }
    else if (lat) { L.d_params = pd->d_fparams_lat; L.d_phases = pd->d_fphases_lat; }
    if (!cuda_ok(n, launch_fused(cl ? pd->fprog_cl : lat ? pd->fprog_lat : pd->fprog, L), "fused kernel")) return false;
    ++n->launches;
    if (lat) ++n->lat_launches;
    if (cl) ++n->cl_launches;
    return true;
  }
  if (!prepare_layer_lane(n, pd, lane)) return false;
  static const bool graphs_on = [] { const char* e = std::getenv("YF_B200_GRAPH"); return !(e && !std::atoi(e)); }();
  if (!graphs_on || n->step_profiling || pd->observer) {
    if (!launch_layer_steps(n, pd, d_in, d_head, nb, st, lane, n->step_profiling)) return false;
    n->launches += P.steps.size();
    return true;
  }
  // One CUDA graph per (input, output, images, lane): the first use captures the launches, every later use replays
  // them with one call.  A caller cycling through more buffers than the cache holds just re-captures.
  const PlanDev::GraphKey key{d_in, d_head, nb, lane};
  auto it = pd->graphs.find(key);
  if (it == pd->graphs.end()) {
    if (pd->graphs.size() >= 512) { for (auto& kv : pd->graphs) cudaGraphExecDestroy(kv.second); pd->graphs.clear(); }
    cudaGraph_t g = nullptr; cudaGraphExec_t ge = nullptr;
    if (!cuda_ok(n, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal), "graph capture")) return false;
    const bool ok = launch_layer_steps(n, pd, d_in, d_head, nb, st, lane, false);
    const cudaError_t ce = cudaStreamEndCapture(st, &g);
    if (!ok || !cuda_ok(n, ce, "graph capture")) { if (g) cudaGraphDestroy(g); return false; }
    const cudaError_t ie = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphDestroy(g);
    if (!cuda_ok(n, ie, "graph instantiate")) return false;
    it = pd->graphs.emplace(key, ge).first;
  }
  if (!cuda_ok(n, cudaGraphLaunch(it->second, st), "graph launch")) return false;
  n->launches += P.steps.size();
  return true;
}

// after a synchronisation that covers the kernels in question: read (and clear) the error word
bool check_mirrored_err(Network* n) {
  const int herr = *static_cast<volatile int*>(n->h_err);
  if (herr == 0) return true;
  set_text("device pipeline watchdog fired (code " + std::to_string(herr) + ")");
  *static_cast<volatile int*>(n->h_err) = 0;
  n->latch(AI_ERROR_INVALID_STATE, AI_ERROR_CODE_LAYER);
  return false;
}
bool check_device_err(Network* n) {
  if (!cuda_ok(n, cudaStreamSynchronize(n->stream), "synchronize")) return false;
  return check_mirrored_err(n);
}

// ---- independent device-resident chunks: fork from n->stream over the kernel lanes, join back ----
struct DevChunk { const int8_t* in; int8_t* out; uint32_t nb; };
bool run_chunks(Network* n, PlanDev* pd, const std::vector<DevChunk>& ch) {
  const bool fused = uses_fused(n, pd);
  const bool layer_lanes = !fused && !pd->observer && !n->step_profiling;      // every lane has its own activation arena
  if (ch.size() < 2 || (!fused && !layer_lanes)) {
    for (const DevChunk& c : ch) { if (!run_steps(n, pd, c.in, c.out, c.nb)) return false; n->last_run_n = c.nb; }
    return true;
  }
  const int nl = fused ? n->lanes : std::min<int>(n->lanes, std::min<int>(PlanDev::kLayerLanes, static_cast<int>(ch.size())));
  if (!fused) for (int l = 0; l < nl; ++l) if (!prepare_layer_lane(n, pd, l)) return false;
  if (!cuda_ok(n, cudaEventRecord(n->ev_fork, n->stream), "fork")) return false;
  for (int l = 0; l < nl; ++l) cudaStreamWaitEvent(n->lane[l], n->ev_fork, 0);
  for (size_t i = 0; i < ch.size(); ++i) {
    const int l = static_cast<int>(i % nl);
    if (!run_steps(n, pd, ch[i].in, ch[i].out, ch[i].nb, n->lane[l], true, fused ? 0 : l)) return false;
    n->last_run_n = ch[i].nb;
  }
  for (int l = 0; l < nl; ++l) {
    cudaEventRecord(n->ev_join[l], n->lane[l]);
    if (!cuda_ok(n, cudaStreamWaitEvent(n->stream, n->ev_join[l], 0), "join")) return false;
  }
  return true;
}

// ---- pipelined host path: H2D (s_h2d) -> kernels (lanes) -> D2H (s_d2h) over a ring of staging slots ----
bool ring_prepare(Network* n, PlanDev* pd) {
  if (pd->ring_ready) return true;
  const size_t in_sz = static_cast<size_t>(pd->plan.H) * pd->plan.W * 3 * pd->cap, out_sz = head_bytes(pd) * pd->cap;
  // The ring owns its buffers: the synchronous paths (d_in / d_head) may run while submissions are still in flight.
  // A partial failure leaves ring_ready false; what was allocated is reused by the next attempt / freed with the plan.
  for (int i = 0; i < PlanDev::kRing; ++i) {
    if (!pd->r_in[i] && !cuda_ok(n, cudaMalloc(&pd->r_in[i], in_sz), "cudaMalloc ring", AI_ERROR_ALLOCATION_FAILED)) return false;
    if (!pd->r_head[i] && !cuda_ok(n, cudaMalloc(&pd->r_head[i], out_sz), "cudaMalloc ring", AI_ERROR_ALLOCATION_FAILED)) return false;
    if (!pd->ev_h2d[i] && !cuda_ok(n, cudaEventCreateWithFlags(&pd->ev_h2d[i], cudaEventDisableTiming), "event")) return false;
    if (!pd->ev_comp[i] && !cuda_ok(n, cudaEventCreateWithFlags(&pd->ev_comp[i], cudaEventDisableTiming), "event")) return false;
    if (!pd->ev_d2h[i] && !cuda_ok(n, cudaEventCreateWithFlags(&pd->ev_d2h[i], cudaEventDisableTiming), "event")) return false;
  }
  pd->ring_ready = true;
  return true;
}
// queue one chunk (nb <= cap) from host memory; returns without waiting.  out may be NULL (heads stay in the slot).
// alone: nothing else is queued around this chunk (a blocking call that fits one piece) -> the launch may take the latency shape
bool ring_submit(Network* n, PlanDev* pd, const int8_t* in_host, int8_t* out_host, uint32_t nb, int8_t** slot_heads, bool alone = false) {
  if (!ring_prepare(n, pd)) return false;
  const int s = static_cast<int>(pd->seq++ % PlanDev::kRing);
  const size_t in_sz = static_cast<size_t>(pd->plan.H) * pd->plan.W * 3, out_sz = head_bytes(pd);
  if (pd->busy[s] && !cuda_ok(n, cudaEventSynchronize(pd->ev_d2h[s]), "ring slot wait")) return false;   // slot's previous user has drained
  cudaStream_t hs = (n->h2d_streams > 1 && (pd->seq & 1)) ? n->s_h2d_b : n->s_h2d;
  if (!cuda_ok(n, cudaMemcpyAsync(pd->r_in[s], in_host, nb * in_sz, cudaMemcpyHostToDevice, hs), "H2D input", AI_ERROR_INVALID_INPUT, AI_ERROR_CODE_INVALID_PTR)) return false;
  cudaEventRecord(pd->ev_h2d[s], hs);
  cudaStream_t ks = uses_fused(n, pd) ? n->lane[n->lane_seq++ % n->lanes] : n->stream;
  cudaStreamWaitEvent(ks, pd->ev_h2d[s], 0);
  if (!run_steps(n, pd, pd->r_in[s], pd->r_head[s], nb, ks, !alone)) return false;
  cudaEventRecord(pd->ev_comp[s], ks);
  cudaStreamWaitEvent(n->s_d2h, pd->ev_comp[s], 0);
  if (out_host && !cuda_ok(n, cudaMemcpyAsync(out_host, pd->r_head[s], nb * out_sz, cudaMemcpyDeviceToHost, n->s_d2h), "D2H output", AI_ERROR_INVALID_OUTPUT, AI_ERROR_CODE_INVALID_PTR)) return false;
  cudaEventRecord(pd->ev_d2h[s], n->s_d2h);
  pd->busy[s] = true; n->last_run_n = nb;
  if (slot_heads) *slot_heads = pd->r_head[s];
  return true;
}
bool ring_wait(Network* n, PlanDev* pd) {
  for (int i = 0; i < PlanDev::kRing; ++i)
    if (pd->busy[i]) { if (!cuda_ok(n, cudaEventSynchronize(pd->ev_d2h[i]), "ring wait")) return false; pd->busy[i] = false; }
  return true;
}

// failure path: whatever a partly queued submission left on the copy / kernel streams must finish before the caller
// gets its buffers back
void ring_drain(Network* n, PlanDev* pd) {
  ring_wait(n, pd);
  cudaStreamSynchronize(n->s_h2d);
  cudaStreamSynchronize(n->s_h2d_b);
  for (int l = 0; l < Network::kLanes; ++l) cudaStreamSynchronize(n->lane[l]);
  cudaStreamSynchronize(n->stream);
  cudaStreamSynchronize(n->s_d2h);
  for (int i = 0; i < PlanDev::kRing; ++i) pd->busy[i] = false;
}

// inference of n images; in/out host or device
int32_t run_images(Network* n, const void* in, void* out, uint32_t count, bool keep_heads_on_device, int8_t** dev_heads) {
  PlanDev* pd = get_plan(n, n->H, n->W);
  if (!pd) return -1;
  const Plan& P = pd->plan;
  const size_t in_sz = static_cast<size_t>(P.H) * P.W * 3, out_sz = head_bytes(pd);
  const bool in_dev = is_device_ptr(in);
  const bool out_dev = out ? is_device_ptr(out) : true;
  if (in_dev && (reinterpret_cast<uintptr_t>(in) & 15)) { set_text("device input must be 16-byte aligned"); n->latch(AI_ERROR_INVALID_INPUT, AI_ERROR_CODE_INVALID_PTR); return -1; }
  if (in_dev && is_foreign_device_ptr(n, in)) { set_text("input lives on another GPU than this context"); n->latch(AI_ERROR_INVALID_INPUT, AI_ERROR_CODE_INVALID_PTR); return -1; }
  if (out && out_dev && is_foreign_device_ptr(n, out)) { set_text("output lives on another GPU than this context"); n->latch(AI_ERROR_INVALID_OUTPUT, AI_ERROR_CODE_INVALID_PTR); return -1; }
  // zero-copy pays up to 8 images when the caller's buffers are page-locked and up to 32 when they are pageable
  // (measured: 1 image 90 / 98 us -> 80 us; 32 pageable images 117 -> 103 us); YF_B200_SMALL overrides both
  static const int small_env = [] { const char* e = getenv("YF_B200_SMALL"); return e ? atoi(e) : -1; }();
  const uint32_t small_max = small_env >= 0 ? static_cast<uint32_t>(small_env) : (is_pageable_ptr(in) || is_pageable_ptr(out) ? 32u : 8u);
  if (!in_dev && out && !out_dev && !n->step_profiling && !pd->observer && count <= small_max && count <= pd->cap) {
    // The reference's own call pattern (one frame per ai_network_run): the image is staged in mapped pinned memory that
    // the kernel's bulk copy reads across PCIe, the head is written straight back to it, and one stream synchronise
    // ends the call -- no cudaMemcpy of the payload, no events.
    constexpr size_t kDoneWords = 64;                       // >= small_max CTAs (one image each)
    const size_t in_bytes = (count * in_sz + 255) & ~size_t(255), out_bytes = (count * out_sz + 255) & ~size_t(255);
    const size_t need = in_bytes + out_bytes + kDoneWords * sizeof(uint32_t);
    if (need > n->small_cap) {
      if (n->h_small) cudaFreeHost(n->h_small);
      n->h_small = nullptr; n->small_cap = 0;
      if (!cuda_ok(n, cudaHostAlloc(&n->h_small, need, cudaHostAllocMapped), "cudaHostAlloc small-batch staging", AI_ERROR_ALLOCATION_FAILED)) return -1;
      n->small_cap = need;
      std::memset(n->h_small, 0, need);                     // completion words start at 0: no sequence number is ever 0
    }
    std::memcpy(n->h_small, in, count * in_sz);
    // Fused path: every CTA stores a completion word after its heads (cta_teardown); the host polls those words in the
    // mapped buffer instead of paying a stream synchronisation's wake-up, and only falls back to it after 2 ms.
    const bool poll = uses_fused(n, pd) && count <= kDoneWords;
    volatile uint32_t* words = reinterpret_cast<volatile uint32_t*>(n->h_small + in_bytes + out_bytes);
    if (poll) { n->done_words = const_cast<uint32_t*>(words); n->done_seq = n->done_seq + 1 ? n->done_seq + 1 : 1; n->done_grid = 0; }
    const bool launched = run_steps(n, pd, n->h_small, n->h_small + in_bytes, count);
    const uint32_t seq = n->done_seq; const int grid = n->done_grid;
    n->done_words = nullptr;
    if (!launched) return -1;
    bool done = false;
    if (poll && grid > 0 && grid <= static_cast<int>(kDoneWords)) {
      const auto t0 = std::chrono::steady_clock::now();
      for (uint32_t spins = 0; !done; ++spins) {
        done = true;
        for (int b = 0; b < grid; ++b) if (words[b] != seq) { done = false; break; }
        if (!done && (spins & 1023u) == 1023u && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2)) break;
      }
    }
    if (!done && !cuda_ok(n, cudaStreamSynchronize(n->stream), "synchronize")) return -1;
    if (!check_mirrored_err(n)) return -1;
    std::memcpy(out, n->h_small + in_bytes, count * out_sz);
    n->last_run_n = count; n->last_ms = 0.f; n->images += count;
    return static_cast<int32_t>(count);
  }
  if (!in_dev && out && !out_dev && !n->step_profiling && !pd->observer) {
    // host -> host: pipeline the chunks (copy of chunk i+1 overlaps the kernels of chunk i).  On the fused path a
    // call that fits one chunk is still cut into up to four pieces of >= 256 images: their kernels run side by side
    // on the lanes, so the copy-in of the later pieces and the copy-out of the earlier ones hide behind them
    // (1,024 images: 498 -> 334 us; below 256 images per piece the ~14 us of host-side submission per piece lose).
    uint32_t piece = pd->cap;
    if (uses_fused(n, pd) && count >= 512) piece = std::min<uint32_t>(pd->cap, std::max<uint32_t>(256, ((count + 3) / 4 + 15) & ~15u));
    for (uint32_t done = 0; done < count; done += piece) {
      const uint32_t nb = std::min<uint32_t>(piece, count - done);
      if (!ring_submit(n, pd, static_cast<const int8_t*>(in) + done * in_sz, static_cast<int8_t*>(out) + done * out_sz, nb, nullptr, count <= piece)) {
        ring_drain(n, pd);                                  // copies into the caller's buffers must not stay pending
        return -1;
      }
    }
    if (!ring_wait(n, pd)) return -1;
    if (!check_mirrored_err(n)) return -1;
    n->last_ms = 0.f;                     // device time of a pipelined call is not a single interval
    n->images += count;
    return static_cast<int32_t>(count);
  }
  cudaEventRecord(n->ev0, n->stream);
  for (uint32_t done = 0; done < count; done += pd->cap) {
    const uint32_t nb = std::min<uint32_t>(pd->cap, count - done);
    const int8_t* din = in_dev ? static_cast<const int8_t*>(in) + done * in_sz : pd->d_in;
    if (!in_dev && !cuda_ok(n, cudaMemcpyAsync(pd->d_in, static_cast<const int8_t*>(in) + done * in_sz, nb * in_sz, cudaMemcpyHostToDevice, n->stream), "H2D input", AI_ERROR_INVALID_INPUT, AI_ERROR_CODE_INVALID_PTR)) return -1;
    int8_t* dhead = (out && out_dev) ? static_cast<int8_t*>(out) + done * out_sz : pd->d_head;
    if (!run_steps(n, pd, din, dhead, nb)) return -1;
    if (out && !out_dev && !cuda_ok(n, cudaMemcpyAsync(static_cast<int8_t*>(out) + done * out_sz, pd->d_head, nb * out_sz, cudaMemcpyDeviceToHost, n->stream), "D2H output", AI_ERROR_INVALID_OUTPUT, AI_ERROR_CODE_INVALID_PTR)) return -1;
    if (keep_heads_on_device && dev_heads) *dev_heads = dhead;
    n->last_run_n = nb;
  }
  cudaEventRecord(n->ev1, n->stream);
  if (!check_device_err(n)) return -1;
  cudaEventElapsedTime(&n->last_ms, n->ev0, n->ev1);
  n->images += count;
  return static_cast<int32_t>(count);
}

bool ensure_dets(Network* n, uint32_t count, uint32_t max_det) {
  const size_t need = static_cast<size_t>(count) * max_det;
  if (need > n->dets_cap || max_det != n->dets_max) {
    cudaFree(n->d_dets); cudaFree(n->d_counts); n->d_dets = nullptr; n->d_counts = nullptr;
    if (!cuda_ok(n, cudaMalloc(&n->d_dets, need * 5 * sizeof(float)), "cudaMalloc detections", AI_ERROR_ALLOCATION_FAILED)) return false;
    if (!cuda_ok(n, cudaMalloc(&n->d_counts, static_cast<size_t>(count) * sizeof(int)), "cudaMalloc counts", AI_ERROR_ALLOCATION_FAILED)) return false;
    n->dets_cap = need; n->dets_max = max_det;
  }
  return true;
}

// decode + NMS launch arguments from the plan (head geometry, output quantisation) and the context (anchors, stride)
bool fill_decode_args(Network* n, const PlanDev* pd, DecodeArgs* a, const int8_t* d_heads, uint32_t count, float conf_thr, float iou_thr,
                      uint32_t flags, float* d_dets, int* d_counts, uint32_t max_det) {
  const Plan& P = pd->plan;
  if (P.buffers[P.output_buf].C != 18) {   // 3 anchors x {x, y, w, h, conf, class} (yoloface.c:116)
    set_text("decode: the model's head does not have 3 x 6 channels"); n->latch(AI_ERROR_INVALID_STATE, AI_ERROR_CODE_INVALID_FORMAT); return false;
  }
  a->head = d_heads; a->n_img = static_cast<int>(count); a->gh = P.GH; a->gw = P.GW;
  a->scale = P.out_scale; a->zp = P.out_zp;
  for (int i = 0; i < 6; ++i) a->anchors[i] = n->anchors[i];
  a->stride = n->stride > 0.f ? n->stride : static_cast<float>(P.H) / static_cast<float>(P.GH);
  a->conf_thr = conf_thr; a->iou_thr = iou_thr; a->plus_one = (flags & YF_B200_NMS_PLUS_ONE) ? 1 : 0;
  a->dets = d_dets; a->counts = d_counts; a->max_det = static_cast<int>(max_det);
  return true;
}

int32_t decode_on_device(Network* n, const int8_t* d_heads, uint32_t count, int gh, int gw, float conf_thr, float iou_thr,
                         uint32_t flags, yf_b200_det* dets, int32_t* counts, uint32_t max_det) {
  if (!ensure_dets(n, count, max_det)) return -1;
  PlanDev* pd = get_plan(n, n->H, n->W); if (!pd) return -1;
  DecodeArgs a{};
  if (!fill_decode_args(n, pd, &a, d_heads, count, conf_thr, iou_thr, flags, n->d_dets, n->d_counts, max_det)) return -1;
  (void)gh; (void)gw;
  if (!cuda_ok(n, launch_decode_nms(a, n->stream), "decode_nms (head too large for the on-device decode?)")) return -1;
  ++n->launches;
  if (!cuda_ok(n, cudaMemcpyAsync(counts, n->d_counts, sizeof(int) * count, cudaMemcpyDeviceToHost, n->stream), "D2H counts")) return -1;
  if (!cuda_ok(n, cudaMemcpyAsync(dets, n->d_dets, sizeof(float) * 5 * count * max_det, cudaMemcpyDeviceToHost, n->stream), "D2H detections")) return -1;
  if (!cuda_ok(n, cudaStreamSynchronize(n->stream), "synchronize")) return -1;
  long total = 0; for (uint32_t i = 0; i < count; ++i) total += counts[i];
  return static_cast<int32_t>(total);
}

// weights handle -> blob pointer: {MARKER, blob, MARKER} table (network_data.c:395-401) or raw blob
const uint8_t* resolve_weights(const void* data) {
  if (!data) return nullptr;
  const uintptr_t* t = static_cast<const uintptr_t*>(data);
  if (t[0] == static_cast<uintptr_t>(AI_MAGIC_MARKER) && t[2] == static_cast<uintptr_t>(AI_MAGIC_MARKER))
    return reinterpret_cast<const uint8_t*>(t[1]);
  return static_cast<const uint8_t*>(data);
}

std::vector<uint8_t>& own_blob() {       // ST-layout blob regenerated from the embedded model
  static std::vector<uint8_t> blob;
  static std::once_flag once;
  std::call_once(once, [] {
    TflModel m; std::string e;
    if (m.parse(yf_embedded_model, yf_embedded_model_len, &e)) blob = st_blob_from_model(m);
    blob.resize((blob.size() + 7) & ~size_t(7));
  });
  return blob;
}

void fill_report(Network* n, ai_network_report* r) {
  ai_buffer& in_desc = n->rep_in; ai_buffer& out_desc = n->rep_out;   // the report points at descriptors owned by the context
  std::memset(r, 0, sizeof *r);
  r->model_name = AI_NETWORK_MODEL_NAME;
  r->model_signature = "yoloface_int8.tflite";
  r->model_datetime = ""; r->compile_datetime = __DATE__ " " __TIME__;
  r->runtime_revision = "yoloface-b200 " YF_B200_BACKEND;
  r->runtime_version = ai_platform_version{7, 0, 0, 0};
  r->tool_revision = "yf_plan";
  r->tool_version = ai_platform_version{AI_TOOLS_VERSION_MAJOR, AI_TOOLS_VERSION_MINOR, AI_TOOLS_VERSION_MICRO, 0};
  r->tool_api_version = ai_platform_version{AI_TOOLS_API_VERSION_MAJOR, AI_TOOLS_API_VERSION_MINOR, AI_TOOLS_API_VERSION_MICRO, 0};
  r->api_version = ai_platform_version{1, 1, 0, 0};
  r->interface_api_version = ai_platform_version{1, 3, 0, 0};
  // The numbers describe the model and input size this context actually runs (not the macros of the deployed model).
  // MACC in ST's convention: conv / depthwise MACs + the window elements of every pooled output + one op per
  // LEAKY_RELU / ADD output element (network_generate_report.txt:484-519).  For yoloface at 56x56 this gives 1,343,776
  // against ST's 1,344,320 (network_generate_report.txt:20): the generator's exact rule for the last 0.04 % is not
  // documented, so the figure is computed, not quoted.
  int in_c = 3, out_c = 18, gh = n->H / 8, gw = n->W / 8, nodes = AI_NETWORK_N_NODES; long long macc = 0; size_t wbytes = AI_NETWORK_DATA_WEIGHTS_SIZE;
  auto it = n->plans.find(std::make_pair(n->H, n->W));
  if (it != n->plans.end()) {
    const Plan& P = it->second->plan;
    in_c = P.buffers[P.input_buf].C; out_c = P.buffers[P.output_buf].C; gh = P.GH; gw = P.GW; nodes = static_cast<int>(P.steps.size());
    macc = P.macs_per_image;
    for (const Step& s : P.steps) {
      const long long opix = static_cast<long long>(s.Hout) * s.Wout;
      if (s.kind == STEP_MAXPOOL) macc += opix * s.Cout * s.kh * s.kw;
      for (int op : s.ops) {
        const int code = n->model.ops[static_cast<size_t>(op)].opcode;
        if (code == OP_LEAKY_RELU || code == OP_ADD) macc += opix * s.Cout;
      }
    }
    st_blob_layout(n->model, &wbytes);
  }
  r->n_macc = static_cast<ai_u32>(macc);
  in_desc = ai_buffer{AI_BUFFER_FORMAT_S8, 1, static_cast<ai_u16>(n->H), static_cast<ai_u16>(n->W), static_cast<ai_u32>(in_c), nullptr, nullptr};
  out_desc = ai_buffer{AI_BUFFER_FORMAT_S8, 1, static_cast<ai_u16>(gh), static_cast<ai_u16>(gw), static_cast<ai_u32>(out_c), nullptr, nullptr};
  r->n_inputs = 1; r->n_outputs = 1; r->inputs = &in_desc; r->outputs = &out_desc;
  r->params = ai_buffer{AI_BUFFER_FORMAT_U8, 1, 1, 1, static_cast<ai_u32>(wbytes), nullptr, nullptr};
  r->activations = ai_buffer{AI_BUFFER_FORMAT_U8, 1, 1, 1, AI_NETWORK_DATA_ACTIVATIONS_SIZE, nullptr, nullptr};
  r->n_nodes = static_cast<ai_u32>(nodes);
  r->signature = 0;
}

// ---- multi-GPU contexts --------------------------------------------------------------------------------------
bool is_group(const Network* n) { return n->members.size() > 1; }

// Split `count` images into contiguous ranges, one per member; fn(member, first, n) runs on the calling thread for
// the primary (whose device lock the caller holds) and on the members' worker threads, each under its own device's
// lock.  Returns the sum of the results, or -1 with the first failing member's error latched on the primary.
template <class F>
int32_t group_dispatch(Network* n, uint32_t count, F fn) {
  const size_t m = n->members.size();
  std::vector<uint32_t> first(m + 1, 0);
  for (size_t i = 0; i < m; ++i) first[i + 1] = first[i] + count / static_cast<uint32_t>(m) + (i < count % m ? 1u : 0u);
  for (size_t i = 1; i < m; ++i) {
    Network* c = n->members[i];
    const uint32_t f0 = first[i], cnt = first[i + 1] - first[i];
    n->workers[i]->post([c, f0, cnt, fn]() -> int32_t {
      if (!cnt) return 0;
      std::lock_guard<std::mutex> lk(g_dev_mu[c->device & 63]);
      cudaSetDevice(c->device);
      return fn(c, f0, cnt);
    });
  }
  int32_t total = first[1] ? fn(n, 0u, first[1]) : 0;
  bool ok = total >= 0;
  for (size_t i = 1; i < m; ++i) {
    std::string text;
    const int32_t r = n->workers[i]->wait(&text);
    Network* c = n->members[i];
    if (r < 0) {
      if (ok) { n->latch(c->err.type != AI_ERROR_NONE ? c->err.type : AI_ERROR_INVALID_STATE, c->err.code); set_text("device " + std::to_string(c->device) + ": " + text); }
      ok = false;
    } else if (ok) total += r;
    c->err = ai_error{AI_ERROR_NONE, AI_ERROR_CODE_NONE};
  }
  cudaSetDevice(n->device);
  return ok ? total : -1;
}
// Host -> host inference over several GPUs with DYNAMIC balancing: the GPUs of a box do not get equal shares of the
// host's copy bandwidth (measured: 20 to 28 GB/s per GPU with eight copying at once), so a static equal split waits for
// the slowest one.  Every member pulls the next chunk from a shared counter and queues it on its pipelined ring
// (ring_submit returns at once unless all of that GPU's staging slots are busy -- which is the back-pressure that does the
// balancing); heads land at the chunk's offset in the caller's buffer.
int32_t group_run_dynamic(Network* n, const int8_t* in, int8_t* out, uint32_t count, size_t in_sz, size_t out_sz) {
  std::atomic<uint32_t> next{0};
  std::atomic<uint32_t>* pnext = &next;
  auto body = [pnext, in, out, count, in_sz, out_sz](Network* c) -> int32_t {
    PlanDev* pd = get_plan(c, c->H, c->W);
    if (!pd) return -1;
    uint32_t mine = 0;
    for (;;) {
      const uint32_t f0 = pnext->fetch_add(pd->cap);
      if (f0 >= count) break;
      const uint32_t nb = std::min<uint32_t>(pd->cap, count - f0);
      if (!ring_submit(c, pd, in + static_cast<size_t>(f0) * in_sz, out + static_cast<size_t>(f0) * out_sz, nb, nullptr)) { ring_drain(c, pd); return -1; }
      mine += nb;
    }
    if (!ring_wait(c, pd) || !check_mirrored_err(c)) return -1;
    c->images += mine; c->last_ms = 0.f;
    return static_cast<int32_t>(mine);
  };
  const size_t m = n->members.size();
  for (size_t i = 1; i < m; ++i) {
    Network* c = n->members[i];
    n->workers[i]->post([c, body]() -> int32_t {
      std::lock_guard<std::mutex> lk(g_dev_mu[c->device & 63]);
      cudaSetDevice(c->device);
      return body(c);
    });
  }
  int32_t total = body(n);
  bool ok = total >= 0;
  for (size_t i = 1; i < m; ++i) {
    std::string text;
    const int32_t r = n->workers[i]->wait(&text);
    Network* c = n->members[i];
    if (r < 0) {
      if (ok) { n->latch(c->err.type != AI_ERROR_NONE ? c->err.type : AI_ERROR_INVALID_STATE, c->err.code); set_text("device " + std::to_string(c->device) + ": " + text); }
      ok = false;
    } else if (ok) total += r;
    c->err = ai_error{AI_ERROR_NONE, AI_ERROR_CODE_NONE};
  }
  cudaSetDevice(n->device);
  return ok ? total : -1;
}
bool dynamic_ok(const Network* n) { return !n->step_profiling && !n->observer; }

// run fn(member) on every member (settings that all devices must share); false if any fails
template <class F>
bool group_each(Network* n, F fn) {
  bool ok = fn(n);
  for (size_t i = 1; i < n->members.size(); ++i) {
    Network* c = n->members[i];
    n->workers[i]->post([c, fn]() -> int32_t {
      std::lock_guard<std::mutex> lk(g_dev_mu[c->device & 63]);
      cudaSetDevice(c->device);
      return fn(c) ? 0 : -1;
    });
  }
  for (size_t i = 1; i < n->members.size(); ++i) {
    std::string text;
    if (n->workers[i]->wait(&text) < 0) {
      Network* c = n->members[i];
      if (ok) { n->latch(c->err.type != AI_ERROR_NONE ? c->err.type : AI_ERROR_INVALID_STATE, c->err.code); set_text("device " + std::to_string(c->device) + ": " + text); }
      c->err = ai_error{AI_ERROR_NONE, AI_ERROR_CODE_NONE};
      ok = false;
    }
  }
  cudaSetDevice(n->device);
  return ok;
}
// host pointers only: a device pointer belongs to one GPU and the whole call then runs there
bool splittable(const Network* n, const void* in, const void* out, uint32_t count) {
  return is_group(n) && count >= 2 * n->members.size() && !is_device_ptr(in) && !(out && is_device_ptr(out));
}
Network* member_for_ptr(Network* n, const void* p) {
  if (!is_group(n) || !p) return n;
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return n; }
  if (a.type != cudaMemoryTypeDevice) return n;
  for (Network* c : n->members) if (c->device == a.device) return c;
  return n;
}

}  // namespace

// ============================================================================================
// X-CUBE-AI style API
// ============================================================================================
extern "C" {

// one context on one device (dev_forced >= 0 overrides the configured ordinal: members of a multi-GPU context)
static Network* create_on_device(const yf_b200_config* cfg, int dev_forced, ai_error* perr) {
  ai_error& err = *perr;
  std::unique_ptr<Network> n(new Network);
  if (cfg && cfg->chunk_images) n->chunk = cfg->chunk_images;
  else if (const char* e = std::getenv("YF_B200_CHUNK")) n->chunk = static_cast<uint32_t>(std::max(1, std::atoi(e)));
  n->observer = cfg && (cfg->flags & YF_B200_FLAG_OBSERVER);
  n->st_act = (cfg && (cfg->flags & YF_B200_FLAG_ST_ACTIVATIONS)) || (std::getenv("YF_B200_ST_ACTIVATIONS") && std::atoi(std::getenv("YF_B200_ST_ACTIVATIONS")));
  if (cfg && (cfg->flags & YF_B200_FLAG_LAYERED)) n->mode = 1;
  else if (cfg && (cfg->flags & YF_B200_FLAG_FUSED_ONLY)) n->mode = 2;
  else if (const char* e = std::getenv("YF_B200_MODE")) n->mode = !std::strcmp(e, "layered") ? 1 : (!std::strcmp(e, "fused") ? 2 : 0);
  const char* path = cfg && cfg->tflite_path ? cfg->tflite_path : std::getenv("YF_B200_TFLITE");
  std::string perr_text; bool ok;
  if (path && *path) {
    FILE* f = std::fopen(path, "rb");
    if (!f) { set_text(std::string("cannot open ") + path); err.type = AI_ERROR_CREATE_FAILED; err.code = AI_ERROR_CODE_INVALID_PTR; return nullptr; }
    std::vector<uint8_t> buf; uint8_t tmp[65536]; size_t k;
    while ((k = std::fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + k);
    std::fclose(f);
    ok = n->model.parse(buf.data(), buf.size(), &perr_text);
    n->custom_model = true;
  } else {
    ok = n->model.parse(yf_embedded_model, yf_embedded_model_len, &perr_text);
  }
  if (!ok) { set_text("model: " + perr_text); err.type = AI_ERROR_CREATE_FAILED; err.code = AI_ERROR_CODE_INVALID_FORMAT; return nullptr; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    set_text("no CUDA device: libyoloface_b200 has no CPU path");
    err.type = AI_ERROR_CREATE_FAILED; err.code = AI_ERROR_CODE_NETWORK; return nullptr;
  }
  int dev = dev_forced;
  if (dev < 0 && cfg && cfg->device >= 0) dev = cfg->device;
  else if (dev < 0) { if (const char* e = std::getenv("YF_B200_DEVICE")) dev = std::atoi(e); }
  if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
  if (dev >= ndev) { set_text("CUDA device ordinal out of range"); err.type = AI_ERROR_CREATE_FAILED; err.code = AI_ERROR_CODE_OUT_OF_RANGE; return nullptr; }
  std::lock_guard<std::mutex> lk(g_dev_mu[dev & 63]);
  cudaDeviceProp prop{};
  int prev_dev = -1;
  if (cudaGetDevice(&prev_dev) != cudaSuccess) prev_dev = -1;
  if (cudaSetDevice(dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    set_text("cannot select CUDA device"); err.type = AI_ERROR_CREATE_FAILED; err.code = AI_ERROR_CODE_NETWORK; return nullptr;
  }
  if (prop.major != 10) {
    set_text("libyoloface_b200 carries sm_100a code only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor));
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    err.type = AI_ERROR_CREATE_FAILED; err.code = AI_ERROR_CODE_NETWORK; return nullptr;
  }
  n->device = dev; n->sm_count = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&n->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&n->s_h2d, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&n->s_d2h, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&n->s_h2d_b, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&n->ev0) != cudaSuccess ||
      cudaEventCreate(&n->ev1) != cudaSuccess || cudaHostAlloc(&n->h_err, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
      (*n->h_err = 0, cudaHostGetDevicePointer(reinterpret_cast<void**>(&n->d_err), n->h_err, 0)) != cudaSuccess ||
      (n->h2d_streams = [] { const char* e = std::getenv("YF_B200_H2D_STREAMS"); return e ? std::max(1, std::min(2, std::atoi(e))) : 2; }(), false) ||
      !make_lanes(n.get()) ||
      kernels_init() != cudaSuccess) {
    set_text(std::string("CUDA setup: ") + cudaGetErrorString(cudaGetLastError()));
    n.reset();                                              // ~Network releases whatever was created, on this device
    if (prev_dev >= 0) cudaSetDevice(prev_dev);             // leave the caller's current device as it was
    err.type = AI_ERROR_CREATE_FAILED; err.code = AI_ERROR_CODE_NETWORK; return nullptr;
  }
  n->stream = n->own_stream;
  if (const char* e = std::getenv("YF_B200_LANES")) n->lanes = std::max(1, std::min(static_cast<int>(Network::kLanes), std::atoi(e)));
  if (dev_forced >= 0 && prev_dev >= 0) cudaSetDevice(prev_dev);   // members are driven by their own threads
  return n.release();
}

// devices of a multi-GPU context: yf_b200_config.device_mask (when the caller's struct carries it), else
// YF_B200_DEVICES = "all" | "0,1,3"; empty = single device
static std::vector<int> group_devices(const ai_buffer* network_config, const yf_b200_config* cfg) {
  std::vector<int> devs;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); return devs; }
  uint64_t mask = 0;
  const size_t cfg_bytes = network_config ? static_cast<size_t>(network_config->height) * network_config->width * network_config->channels : 0;
  if (cfg && cfg_bytes >= offsetof(yf_b200_config, device_mask) + sizeof(uint32_t)) mask = cfg->device_mask;
  if (!mask && !(cfg && cfg->device >= 0)) {
    if (const char* e = std::getenv("YF_B200_DEVICES")) {
      if (!std::strcmp(e, "all")) mask = ndev >= 64 ? ~0ull : ((1ull << ndev) - 1);
      else for (const char* p = e; *p;) { char* q; long d = std::strtol(p, &q, 10); if (q == p) break; if (d >= 0 && d < 64) mask |= 1ull << d; p = *q ? q + 1 : q; }
    }
  }
  for (int d = 0; d < ndev && d < 64; ++d) if (mask & (1ull << d)) devs.push_back(d);
  if (devs.size() < 2) devs.clear();
  return devs;
}

AI_API_ENTRY ai_error ai_network_create(ai_handle* network, const ai_buffer* network_config) {
  ai_error err{AI_ERROR_NONE, AI_ERROR_CODE_NONE};
  if (!network) { err.type = AI_ERROR_CREATE_FAILED; err.code = AI_ERROR_CODE_INVALID_PTR; return err; }
  *network = AI_HANDLE_NULL;
  const yf_b200_config* cfg = nullptr;
  if (network_config && network_config->data) {
    cfg = static_cast<const yf_b200_config*>(network_config->data);
    if (cfg->magic != YF_B200_CONFIG_MAGIC) { err.type = AI_ERROR_CREATE_FAILED; err.code = AI_ERROR_CODE_INVALID_FORMAT; set_text("network_config is not a yf_b200_config"); return err; }
  }
  const std::vector<int> devs = group_devices(network_config, cfg);
  Network* n = create_on_device(cfg, devs.empty() ? -1 : devs[0], &err);
  if (!n) return err;
  if (!devs.empty()) {
    // ONE handle, several GPUs: the caller keeps the reference's single-handle call sequence (yoloface.c:188-240)
    // and the batch of every run is split by image over the devices (no collective: SURVEY.md 8e).
    cudaSetDevice(devs[0]);
    n->members.push_back(n);
    n->workers.emplace_back(nullptr);
    for (size_t i = 1; i < devs.size(); ++i) {
      Network* c = create_on_device(cfg, devs[i], &err);
      if (!c) { cudaSetDevice(devs[0]); delete n; return err; }
      c->owner = n;
      n->members.push_back(c);
      n->workers.emplace_back(new Worker(devs[i]));
    }
    cudaSetDevice(devs[0]);
  }
  { std::lock_guard<std::mutex> rk(g_mu); g_nets.push_back(n); }
  *network = n;
  return err;
}

AI_API_ENTRY ai_handle ai_network_destroy(ai_handle network) {
  Network* n = as_net(network);
  if (!n) return network;
  std::lock_guard<std::mutex> lk(g_dev_mu[n->device & 63]);
  cudaSetDevice(n->device);
  cudaStreamSynchronize(n->stream);
  for (int l = 0; l < Network::kLanes; ++l) cudaStreamSynchronize(n->lane[l]);
  cudaStreamSynchronize(n->s_d2h);
  for (Network* c : n->members.empty() ? std::vector<Network*>{n} : n->members) {
    if (c != n) { cudaSetDevice(c->device); cudaDeviceSynchronize(); }
  }
  cudaSetDevice(n->device);
  { std::lock_guard<std::mutex> rk(g_mu); g_nets.erase(std::remove(g_nets.begin(), g_nets.end(), n), g_nets.end()); }
  delete n;                                                 // ~Network (members and their worker threads included)
  return AI_HANDLE_NULL;
}

AI_API_ENTRY ai_error ai_network_get_error(ai_handle network) {
  Network* n = as_net(network);
  if (!n) return ai_error{AI_ERROR_INVALID_HANDLE, AI_ERROR_CODE_NETWORK};
  std::lock_guard<std::mutex> lk(g_dev_mu[n->device & 63]);
  ai_error e = n->err;
  n->err = ai_error{AI_ERROR_NONE, AI_ERROR_CODE_NONE};
  return e;
}

AI_API_ENTRY ai_bool ai_network_init(ai_handle network, const ai_network_params* params) {
  Network* n = as_net(network);
  if (!n) return false;
  std::lock_guard<std::mutex> lk(g_dev_mu[n->device & 63]);
  if (!params) { n->latch(AI_ERROR_INVALID_PARAM, AI_ERROR_CODE_NETWORK_PARAMS); return false; }
  const ai_buffer* wbuf = &params->params; const ai_buffer* abuf = &params->activations;
  if (params->map_signature == AI_MAGIC_SIGNATURE) {            // ai_network_data_params_get() form
    if (!params->map_weights.buffer || params->map_weights.size < 1) { n->latch(AI_ERROR_INIT_FAILED, AI_ERROR_CODE_NETWORK_WEIGHTS); return false; }
    wbuf = &params->map_weights.buffer[0];
    abuf = (params->map_activations.buffer && params->map_activations.size) ? &params->map_activations.buffer[0] : nullptr;
  }
  const uint8_t* blob = resolve_weights(wbuf->data);
  size_t need = 0; st_blob_layout(n->model, &need);
  if (!blob && !n->custom_model) { n->latch(AI_ERROR_INIT_FAILED, AI_ERROR_CODE_NETWORK_WEIGHTS); set_text("weights handle is NULL"); return false; }
  // A model given by path (yf_b200_config.tflite_path) carries its own weights: a NULL weights handle keeps them.  A
  // non-NULL handle must be that model's blob in ST's layout (conv operators in graph order, weights then bias).
  if (blob) {
    const size_t wsize = static_cast<size_t>(wbuf->height) * wbuf->width * wbuf->channels;
    if (wsize < need) { n->latch(AI_ERROR_INIT_FAILED, AI_ERROR_CODE_NETWORK_WEIGHTS); set_text("weights buffer smaller than the model's blob"); return false; }
  }
  if (abuf) {
    // caller-owned arena: validated like ST's runtime does, but unused (activations live in HBM)
    const size_t asize = static_cast<size_t>(abuf->height) * abuf->width * abuf->channels;
    if (abuf->data && asize < AI_NETWORK_DATA_ACTIVATIONS_SIZE) { n->latch(AI_ERROR_INIT_FAILED, AI_ERROR_CODE_NETWORK_ACTIVATIONS); set_text("activations buffer too small"); return false; }
  }
  cudaSetDevice(n->device);
  const std::vector<uint8_t> weights = blob ? std::vector<uint8_t>(blob, blob + need) : std::vector<uint8_t>();
  auto init_one = [weights](Network* c) {
    c->blob = weights;
    c->plans.clear();
    if (!get_plan(c, c->H, c->W)) return false;
    if (!cuda_ok(c, cudaStreamSynchronize(c->stream), "init synchronize", AI_ERROR_INIT_FAILED)) return false;
    c->initialized = true;
    return true;
  };
  return is_group(n) ? group_each(n, init_one) : init_one(n);
}

static ai_i32 process(ai_handle network, const ai_buffer* input, ai_buffer* output) {
  Network* n = as_net(network);
  if (!n) return 0;
  std::lock_guard<std::mutex> lk(g_dev_mu[n->device & 63]);
  if (!n->initialized) { n->latch(AI_ERROR_INVALID_STATE, AI_ERROR_CODE_MISSED_INIT); return 0; }
  if (!input || !input->data) { n->latch(AI_ERROR_INVALID_INPUT, AI_ERROR_CODE_INVALID_PTR); return 0; }
  if (AI_BUFFER_FMT_GET(input->format) != AI_BUFFER_FORMAT_S8) { n->latch(AI_ERROR_INVALID_INPUT, AI_ERROR_CODE_INVALID_FORMAT); return 0; }
  if (input->height != n->H || input->width != n->W || input->channels != 3) { n->latch(AI_ERROR_INVALID_INPUT, AI_ERROR_CODE_INVALID_SIZE); return 0; }
  if (input->n_batches == 0) { n->latch(AI_ERROR_INVALID_INPUT, AI_ERROR_CODE_INVALID_BATCH); return 0; }
  if (output) {
    if (!output->data) { n->latch(AI_ERROR_INVALID_OUTPUT, AI_ERROR_CODE_INVALID_PTR); return 0; }
    if (AI_BUFFER_FMT_GET(output->format) != AI_BUFFER_FORMAT_S8) { n->latch(AI_ERROR_INVALID_OUTPUT, AI_ERROR_CODE_INVALID_FORMAT); return 0; }
    if (output->height != n->H / 8 || output->width != n->W / 8 || output->channels != 18) { n->latch(AI_ERROR_INVALID_OUTPUT, AI_ERROR_CODE_INVALID_SIZE); return 0; }
    if (output->n_batches < input->n_batches) { n->latch(AI_ERROR_INVALID_OUTPUT, AI_ERROR_CODE_INVALID_BATCH); return 0; }
  }
  cudaSetDevice(n->device);
  int32_t r;
  if (splittable(n, input->data, output ? output->data : nullptr, input->n_batches)) {
    PlanDev* pd = get_plan(n, n->H, n->W); if (!pd) return 0;
    const size_t in_sz = static_cast<size_t>(pd->plan.H) * pd->plan.W * 3, out_sz = head_bytes(pd);
    const int8_t* in = static_cast<const int8_t*>(input->data); int8_t* out = output ? static_cast<int8_t*>(output->data) : nullptr;
    if (out && dynamic_ok(n) && input->n_batches >= 2 * pd->cap) r = group_run_dynamic(n, in, out, input->n_batches, in_sz, out_sz);
    else r = group_dispatch(n, input->n_batches, [in, out, in_sz, out_sz](Network* c, uint32_t f0, uint32_t cnt) {
      return run_images(c, in + f0 * in_sz, out ? out + f0 * out_sz : nullptr, cnt, false, nullptr);
    });
  } else {
    Network* c = member_for_ptr(n, input->data);
    if (c != n) { std::lock_guard<std::mutex> ck(g_dev_mu[c->device & 63]); cudaSetDevice(c->device); r = run_images(c, input->data, output ? output->data : nullptr, input->n_batches, false, nullptr); cudaSetDevice(n->device); }
    else r = run_images(n, input->data, output ? output->data : nullptr, input->n_batches, false, nullptr);
  }
  return r < 0 ? 0 : r;
}

AI_API_ENTRY ai_i32 ai_network_run(ai_handle network, const ai_buffer* input, ai_buffer* output) {
  if (!output) { Network* n = as_net(network); if (n) n->latch(AI_ERROR_INVALID_OUTPUT, AI_ERROR_CODE_INVALID_PTR); return 0; }
  return process(network, input, output);
}
AI_API_ENTRY ai_i32 ai_network_forward(ai_handle network, const ai_buffer* input) { return process(network, input, nullptr); }

AI_API_ENTRY ai_bool ai_network_get_report(ai_handle network, ai_network_report* report) {
  Network* n = as_net(network);
  if (!n || !report) return false;
  std::lock_guard<std::mutex> lk(g_dev_mu[n->device & 63]);
  fill_report(n, report);
  return true;
}
AI_API_ENTRY ai_bool ai_network_get_info(ai_handle network, ai_network_report* report) { return ai_network_get_report(network, report); }

AI_API_ENTRY ai_handle ai_network_data_weights_get(void) {
  static const void* table[3];
  table[0] = reinterpret_cast<const void*>(static_cast<uintptr_t>(AI_MAGIC_MARKER));
  table[1] = own_blob().data();
  table[2] = reinterpret_cast<const void*>(static_cast<uintptr_t>(AI_MAGIC_MARKER));
  return AI_HANDLE_PTR(table);
}

AI_API_ENTRY ai_bool ai_network_data_params_get(ai_handle network, ai_network_params* params) {
  if (!network || !params) return false;
  static ai_buffer w[1], a[1];
  w[0] = ai_buffer{AI_BUFFER_FORMAT_U8, 1, 1, 1, AI_NETWORK_DATA_WEIGHTS_SIZE, own_blob().data(), nullptr};
  a[0] = ai_buffer{AI_BUFFER_FORMAT_U8, 1, 1, 1, AI_NETWORK_DATA_ACTIVATIONS_SIZE, nullptr, nullptr};
  std::memset(params, 0, sizeof *params);
  params->map_signature = AI_MAGIC_SIGNATURE;
  params->map_weights = ai_buffer_array{AI_FLAG_NONE, 1, w};
  params->map_activations = ai_buffer_array{AI_FLAG_NONE, 1, a};
  return true;
}

// ============================================================================================
// B200 extensions
// ============================================================================================
#define YF_NET_OR_FAIL(n, network)                                                             \
  Network* n = as_net(network);                                                                \
  if (!n) return -1;                                                                           \
  std::lock_guard<std::mutex> lk(g_dev_mu[n->device & 63]);                                    \
  cudaSetDevice(n->device);

AI_API_ENTRY int32_t yf_b200_set_input_size(ai_handle network, int32_t height, int32_t width) {
  YF_NET_OR_FAIL(n, network)
  if (height < 8 || width < 8 || height % 8 || width % 8 || height > 4096 || width > 4096) { n->latch(AI_ERROR_INVALID_PARAM, AI_ERROR_CODE_INVALID_SIZE); return -1; }
  auto set_one = [height, width](Network* c) { c->H = height; c->W = width; return !(c->initialized && !get_plan(c, c->H, c->W)); };
  return (is_group(n) ? group_each(n, set_one) : set_one(n)) ? 0 : -1;
}

AI_API_ENTRY int32_t yf_b200_run(ai_handle network, const void* in, void* out, uint32_t count) {
  YF_NET_OR_FAIL(n, network)
  if (!n->initialized) { n->latch(AI_ERROR_INVALID_STATE, AI_ERROR_CODE_MISSED_INIT); return -1; }
  if (!in) { n->latch(AI_ERROR_INVALID_INPUT, AI_ERROR_CODE_INVALID_PTR); return -1; }
  if (!out) { n->latch(AI_ERROR_INVALID_OUTPUT, AI_ERROR_CODE_INVALID_PTR); return -1; }
  if (count == 0) return 0;
  if (splittable(n, in, out, count)) {
    PlanDev* pd = get_plan(n, n->H, n->W); if (!pd) return -1;
    const size_t in_sz = static_cast<size_t>(pd->plan.H) * pd->plan.W * 3, out_sz = head_bytes(pd);
    const int8_t* pi = static_cast<const int8_t*>(in); int8_t* po = static_cast<int8_t*>(out);
    if (dynamic_ok(n) && count >= 2 * pd->cap) return group_run_dynamic(n, pi, po, count, in_sz, out_sz);
    return group_dispatch(n, count, [pi, po, in_sz, out_sz](Network* c, uint32_t f0, uint32_t cnt) {
      return run_images(c, pi + f0 * in_sz, po + f0 * out_sz, cnt, false, nullptr);
    });
  }
  Network* c = member_for_ptr(n, in);
  if (c != n) { std::lock_guard<std::mutex> ck(g_dev_mu[c->device & 63]); cudaSetDevice(c->device); const int32_t r = run_images(c, in, out, count, false, nullptr); cudaSetDevice(n->device); return r; }
  return run_images(n, in, out, count, false, nullptr);
}

AI_API_ENTRY int32_t yf_b200_set_stream(ai_handle network, void* cuda_stream) {
  YF_NET_OR_FAIL(n, network)
  cudaStreamSynchronize(n->stream);
  n->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : n->own_stream;
  return 0;
}

AI_API_ENTRY int32_t yf_b200_enqueue(ai_handle network, const void* d_in, void* d_out, uint32_t count) {
  YF_NET_OR_FAIL(n, network)
  if (!n->initialized) { n->latch(AI_ERROR_INVALID_STATE, AI_ERROR_CODE_MISSED_INIT); return -1; }
  // the kernels dereference these pointers: device, managed or page-locked host memory -- never pageable host memory
  if (!d_in || (reinterpret_cast<uintptr_t>(d_in) & 15) || is_pageable_ptr(d_in) || is_foreign_device_ptr(n, d_in)) { set_text("enqueue: input must be 16-byte aligned device-accessible memory"); n->latch(AI_ERROR_INVALID_INPUT, AI_ERROR_CODE_INVALID_PTR); return -1; }
  if (!d_out || is_pageable_ptr(d_out) || is_foreign_device_ptr(n, d_out)) { set_text("enqueue: output must be device-accessible memory"); n->latch(AI_ERROR_INVALID_OUTPUT, AI_ERROR_CODE_INVALID_PTR); return -1; }
  PlanDev* pd = get_plan(n, n->H, n->W); if (!pd) return -1;
  const size_t in_sz = static_cast<size_t>(pd->plan.H) * pd->plan.W * 3, out_sz = head_bytes(pd);
  std::vector<DevChunk> ch;
  for (uint32_t done = 0; done < count; done += pd->cap)
    ch.push_back({static_cast<const int8_t*>(d_in) + done * in_sz, static_cast<int8_t*>(d_out) + done * out_sz, std::min<uint32_t>(pd->cap, count - done)});
  if (!run_chunks(n, pd, ch)) return -1;
  n->images += count;
  return static_cast<int32_t>(count);
}

AI_API_ENTRY int32_t yf_b200_enqueue_batches(ai_handle network, const void* const* d_in, void* const* d_out, const uint32_t* counts, uint32_t n_batches) {
  YF_NET_OR_FAIL(n, network)
  if (!n->initialized) { n->latch(AI_ERROR_INVALID_STATE, AI_ERROR_CODE_MISSED_INIT); return -1; }
  if (!d_in || !d_out || !counts) { n->latch(AI_ERROR_INVALID_PARAM, AI_ERROR_CODE_INVALID_PTR); return -1; }
  PlanDev* pd = get_plan(n, n->H, n->W); if (!pd) return -1;
  const size_t in_sz = static_cast<size_t>(pd->plan.H) * pd->plan.W * 3, out_sz = head_bytes(pd);
  std::vector<DevChunk> ch;
  uint64_t total = 0;
  for (uint32_t b = 0; b < n_batches; ++b) {
    if (!d_in[b] || (reinterpret_cast<uintptr_t>(d_in[b]) & 15) || is_pageable_ptr(d_in[b]) || is_foreign_device_ptr(n, d_in[b])) { set_text("enqueue_batches: input must be 16-byte aligned device-accessible memory"); n->latch(AI_ERROR_INVALID_INPUT, AI_ERROR_CODE_INVALID_PTR); return -1; }
    if (!d_out[b] || is_pageable_ptr(d_out[b]) || is_foreign_device_ptr(n, d_out[b])) { set_text("enqueue_batches: output must be device-accessible memory"); n->latch(AI_ERROR_INVALID_OUTPUT, AI_ERROR_CODE_INVALID_PTR); return -1; }
    for (uint32_t done = 0; done < counts[b]; done += pd->cap)
      ch.push_back({static_cast<const int8_t*>(d_in[b]) + done * in_sz, static_cast<int8_t*>(d_out[b]) + done * out_sz, std::min<uint32_t>(pd->cap, counts[b] - done)});
    total += counts[b];
  }
  if (total > 0x7fffffffull) { n->latch(AI_ERROR_INVALID_PARAM, AI_ERROR_CODE_OUT_OF_RANGE); return -1; }
  if (!run_chunks(n, pd, ch)) return -1;
  n->images += total;
  return static_cast<int32_t>(total);
}

AI_API_ENTRY int32_t yf_b200_submit(ai_handle network, const void* in_host, void* out_host, uint32_t count) {
  YF_NET_OR_FAIL(n, network)
  if (!n->initialized) { n->latch(AI_ERROR_INVALID_STATE, AI_ERROR_CODE_MISSED_INIT); return -1; }
  if (!in_host || is_device_ptr(in_host)) { n->latch(AI_ERROR_INVALID_INPUT, AI_ERROR_CODE_INVALID_PTR); return -1; }
  if (!out_host || is_device_ptr(out_host)) { n->latch(AI_ERROR_INVALID_OUTPUT, AI_ERROR_CODE_INVALID_PTR); return -1; }
  PlanDev* pd = get_plan(n, n->H, n->W); if (!pd) return -1;
  const size_t in_sz = static_cast<size_t>(pd->plan.H) * pd->plan.W * 3, out_sz = head_bytes(pd);
  for (uint32_t done = 0; done < count; done += pd->cap) {
    const uint32_t nb = std::min<uint32_t>(pd->cap, count - done);
    if (!ring_submit(n, pd, static_cast<const int8_t*>(in_host) + done * in_sz, static_cast<int8_t*>(out_host) + done * out_sz, nb, nullptr)) {
      ring_drain(n, pd);                                    // drain what was queued: nothing may land in the caller's buffers later
      return -1;
    }
  }
  n->images += count;
  return static_cast<int32_t>(count);
}

AI_API_ENTRY int32_t yf_b200_wait(ai_handle network) {
  YF_NET_OR_FAIL(n, network)
  for (auto& kv : n->plans) if (!ring_wait(n, kv.second.get())) return -1;
  return check_mirrored_err(n) ? 0 : -1;
}

AI_API_ENTRY int32_t yf_b200_sync(ai_handle network) {
  YF_NET_OR_FAIL(n, network)
  return check_device_err(n) ? 0 : -1;
}

AI_API_ENTRY int32_t yf_b200_decode(ai_handle network, const void* heads, uint32_t count, float conf_thr, float iou_thr,
                                    uint32_t flags, yf_b200_det* dets, int32_t* counts, uint32_t max_det) {
  YF_NET_OR_FAIL(n, network)
  if (!n->initialized) { n->latch(AI_ERROR_INVALID_STATE, AI_ERROR_CODE_MISSED_INIT); return -1; }
  if (!heads || !dets || !counts || !max_det) { n->latch(AI_ERROR_INVALID_PARAM, AI_ERROR_CODE_INVALID_PTR); return -1; }
  if (count == 0) return 0;
  PlanDev* pd = get_plan(n, n->H, n->W); if (!pd) return -1;
  const size_t hsz = head_bytes(pd);
  int32_t total = 0;
  for (uint32_t done = 0; done < count; done += pd->cap) {
    const uint32_t nb = std::min<uint32_t>(pd->cap, count - done);
    const int8_t* dh;
    if (is_device_ptr(heads)) dh = static_cast<const int8_t*>(heads) + done * hsz;
    else {
      if (!cuda_ok(n, cudaMemcpyAsync(pd->d_head, static_cast<const int8_t*>(heads) + done * hsz, nb * hsz, cudaMemcpyHostToDevice, n->stream), "H2D heads")) return -1;
      dh = pd->d_head;
    }
    int32_t r = decode_on_device(n, dh, nb, pd->plan.GH, pd->plan.GW, conf_thr, iou_thr, flags, dets + static_cast<size_t>(done) * max_det, counts + done, max_det);
    if (r < 0) return -1;
    total += r;
  }
  return total;
}

}  // extern "C"
// inference + decode + NMS of `count` images on ONE context (the caller holds its device's lock and has selected it)
static int32_t detect_impl(Network* n, const void* in, uint32_t count, float conf_thr, float iou_thr,
                           uint32_t flags, yf_b200_det* dets, int32_t* counts, uint32_t max_det, void* heads_out) {
  PlanDev* pd = get_plan(n, n->H, n->W); if (!pd) return -1;
  const size_t in_sz = static_cast<size_t>(pd->plan.H) * pd->plan.W * 3, hsz = head_bytes(pd);
  const bool in_dev = is_device_ptr(in);
  int32_t total = 0;
  if (in_dev && !heads_out && count > pd->cap && uses_fused(n, pd) && !(reinterpret_cast<uintptr_t>(in) & 15)) {
    // Device-resident images, several chunks, detections only: chunk i runs inference then decode + NMS on lane i % 4
    // (its heads stay in that lane's staging buffer), so the inference of one chunk overlaps the decode of its
    // neighbours; ONE copy brings all detections back.
    if (!ring_prepare(n, pd) || !ring_wait(n, pd) || !ensure_dets(n, count, max_det)) return -1;   // the slots' head buffers must be idle
    if (!cuda_ok(n, cudaEventRecord(n->ev_fork, n->stream), "fork")) return -1;
    const int nl = std::min(n->lanes, static_cast<int>(PlanDev::kRing));
    for (int l = 0; l < nl; ++l) cudaStreamWaitEvent(n->lane[l], n->ev_fork, 0);
    uint32_t ci = 0;
    for (uint32_t done = 0; done < count; done += pd->cap, ++ci) {
      const uint32_t nb = std::min<uint32_t>(pd->cap, count - done);
      const int l = static_cast<int>(ci % nl);
      int8_t* heads = pd->r_head[l];                           // ring slots 0..3 double as the lanes' head buffers
      if (!run_steps(n, pd, static_cast<const int8_t*>(in) + done * in_sz, heads, nb, n->lane[l], true)) return -1;
      DecodeArgs a{};
      if (!fill_decode_args(n, pd, &a, heads, nb, conf_thr, iou_thr, flags, n->d_dets + static_cast<size_t>(done) * max_det * 5, n->d_counts + done, max_det)) return -1;
      if (!cuda_ok(n, launch_decode_nms(a, n->lane[l]), "decode_nms")) return -1;
      ++n->launches;
    }
    for (int l = 0; l < nl; ++l) {
      cudaEventRecord(n->ev_join[l], n->lane[l]);
      if (!cuda_ok(n, cudaStreamWaitEvent(n->stream, n->ev_join[l], 0), "join")) return -1;
    }
    if (!cuda_ok(n, cudaMemcpyAsync(counts, n->d_counts, sizeof(int) * count, cudaMemcpyDeviceToHost, n->stream), "D2H counts")) return -1;
    if (!cuda_ok(n, cudaMemcpyAsync(dets, n->d_dets, sizeof(float) * 5 * count * max_det, cudaMemcpyDeviceToHost, n->stream), "D2H detections")) return -1;
    if (!check_device_err(n)) return -1;
    n->images += count; n->last_run_n = std::min<uint32_t>(pd->cap, count);
    for (uint32_t i = 0; i < count; ++i) total += counts[i];
    return total;
  }
  for (uint32_t done = 0; done < count; done += pd->cap) {
    const uint32_t nb = std::min<uint32_t>(pd->cap, count - done);
    const int8_t* src = static_cast<const int8_t*>(in) + done * in_sz;
    int8_t* dh = nullptr;
    void* hout = heads_out ? static_cast<int8_t*>(heads_out) + done * hsz : nullptr;
    // heads stay in the staging buffer unless the caller gave a device destination
    if (hout && is_device_ptr(hout)) { if (run_images(n, src, hout, nb, false, nullptr) < 0) return -1; dh = static_cast<int8_t*>(hout); }
    else {
      if (run_images(n, src, in_dev ? static_cast<void*>(pd->d_head) : nullptr, nb, true, &dh) < 0) return -1;
      dh = pd->d_head;
      if (hout && !cuda_ok(n, cudaMemcpyAsync(hout, pd->d_head, nb * hsz, cudaMemcpyDeviceToHost, n->stream), "D2H heads")) return -1;
    }
    int32_t r = decode_on_device(n, dh, nb, pd->plan.GH, pd->plan.GW, conf_thr, iou_thr, flags, dets + static_cast<size_t>(done) * max_det, counts + done, max_det);
    if (r < 0) return -1;
    total += r;
  }
  return total;
}
extern "C" {

AI_API_ENTRY int32_t yf_b200_detect(ai_handle network, const void* in, uint32_t count, float conf_thr, float iou_thr,
                                    uint32_t flags, yf_b200_det* dets, int32_t* counts, uint32_t max_det, void* heads_out) {
  YF_NET_OR_FAIL(n, network)
  if (!n->initialized) { n->latch(AI_ERROR_INVALID_STATE, AI_ERROR_CODE_MISSED_INIT); return -1; }
  if (!in || !dets || !counts || !max_det) { n->latch(AI_ERROR_INVALID_PARAM, AI_ERROR_CODE_INVALID_PTR); return -1; }
  if (count == 0) return 0;
  if (splittable(n, in, heads_out, count)) {
    // every device runs inference + decode + NMS on its range of images; only detections (and, if asked for, heads)
    // come back, each into its range of the caller's arrays -- "only detections are gathered to the host"
    PlanDev* pd = get_plan(n, n->H, n->W); if (!pd) return -1;
    const size_t in_sz = static_cast<size_t>(pd->plan.H) * pd->plan.W * 3, hsz = head_bytes(pd);
    const int8_t* pi = static_cast<const int8_t*>(in); int8_t* ph = static_cast<int8_t*>(heads_out);
    return group_dispatch(n, count, [=](Network* c, uint32_t f0, uint32_t cnt) {
      return detect_impl(c, pi + f0 * in_sz, cnt, conf_thr, iou_thr, flags, dets + static_cast<size_t>(f0) * max_det, counts + f0, max_det,
                         ph ? ph + f0 * hsz : nullptr);
    });
  }
  Network* c = member_for_ptr(n, in);
  if (c != n) {
    std::lock_guard<std::mutex> ck(g_dev_mu[c->device & 63]); cudaSetDevice(c->device);
    const int32_t r = detect_impl(c, in, count, conf_thr, iou_thr, flags, dets, counts, max_det, heads_out);
    cudaSetDevice(n->device);
    return r;
  }
  return detect_impl(n, in, count, conf_thr, iou_thr, flags, dets, counts, max_det, heads_out);
}

AI_API_ENTRY int32_t yf_b200_set_decode_params(ai_handle network, const float* anchors6, float stride) {
  YF_NET_OR_FAIL(n, network)
  if (anchors6) for (int i = 0; i < 6; ++i) {
    if (!(anchors6[i] > 0.f)) { n->latch(AI_ERROR_INVALID_PARAM, AI_ERROR_CODE_OUT_OF_RANGE); return -1; }
    n->anchors[i] = anchors6[i];
  }
  if (stride < 0.f) { n->latch(AI_ERROR_INVALID_PARAM, AI_ERROR_CODE_OUT_OF_RANGE); return -1; }
  n->stride = stride;
  for (size_t i = 1; i < n->members.size(); ++i) { for (int k = 0; k < 6; ++k) n->members[i]->anchors[k] = n->anchors[k]; n->members[i]->stride = stride; }
  return 0;
}

AI_API_ENTRY int32_t yf_b200_preprocess_rgb565(ai_handle network, const void* frames, void* out, uint32_t count) {
  YF_NET_OR_FAIL(n, network)
  if (!frames || !out) { n->latch(AI_ERROR_INVALID_PARAM, AI_ERROR_CODE_INVALID_PTR); return -1; }
  if (count == 0) return 0;
  const size_t fsz = 112 * 112 * 2, osz = 56 * 56 * 3;
  const bool fdev = is_device_ptr(frames), odev = is_device_ptr(out);
  const size_t need = (fdev ? 0 : fsz * count) + (odev ? 0 : osz * count);
  if (need > n->frames_cap) {
    cudaFree(n->d_frames); n->d_frames = nullptr; n->frames_cap = 0;
    if (!cuda_ok(n, cudaMalloc(&n->d_frames, need), "cudaMalloc frames", AI_ERROR_ALLOCATION_FAILED)) return -1;
    n->frames_cap = need;
  }
  const uint8_t* dfr = static_cast<const uint8_t*>(frames);
  uint8_t* scratch = n->d_frames;
  if (!fdev) {
    if (!cuda_ok(n, cudaMemcpyAsync(scratch, frames, fsz * count, cudaMemcpyHostToDevice, n->stream), "H2D frames")) return -1;
    dfr = scratch; scratch += fsz * count;
  }
  PrepArgs a{dfr, odev ? static_cast<int8_t*>(out) : reinterpret_cast<int8_t*>(scratch), static_cast<int>(count)};
  if (!cuda_ok(n, launch_prep_rgb565(a, n->stream), "prep_rgb565")) return -1;
  ++n->launches;
  if (!odev && !cuda_ok(n, cudaMemcpyAsync(out, a.out, osz * count, cudaMemcpyDeviceToHost, n->stream), "D2H inputs")) return -1;
  if (!cuda_ok(n, cudaStreamSynchronize(n->stream), "synchronize")) return -1;
  return static_cast<int32_t>(count);
}

AI_API_ENTRY int32_t yf_b200_set_observer(ai_handle network, int32_t enable) {
  YF_NET_OR_FAIL(n, network)
  n->observer = enable != 0;
  if (n->initialized && !get_plan(n, n->H, n->W)) return -1;
  return 0;
}

AI_API_ENTRY int32_t yf_b200_tensor_shape(ai_handle network, int32_t t, int32_t dims[4]) {
  YF_NET_OR_FAIL(n, network)
  if (!n->initialized) return -1;
  PlanDev* pd = get_plan(n, n->H, n->W); if (!pd) return -1;
  const Plan& P = pd->plan;
  if (t < 0 || t >= static_cast<int>(P.loc.size()) || P.loc[t].buf < 0 || P.loc[t].view) return -1;
  const PBuffer& b = P.buffers[P.loc[t].buf];
  dims[0] = 1; dims[1] = b.H; dims[2] = b.W; dims[3] = P.loc[t].C;
  return 0;
}

AI_API_ENTRY int64_t yf_b200_get_tensor(ai_handle network, int32_t t, uint32_t count, void* dst, uint64_t dst_bytes) {
  YF_NET_OR_FAIL(n, network)
  if (!n->initialized || !n->observer) { n->latch(AI_ERROR_INVALID_STATE, AI_ERROR_CODE_MISSED_INIT); set_text("observer mode is off"); return -1; }
  PlanDev* pd = get_plan(n, n->H, n->W); if (!pd) return -1;
  const Plan& P = pd->plan;
  if (t < 0 || t >= static_cast<int>(P.loc.size()) || P.loc[t].buf < 0 || P.loc[t].view) { n->latch(AI_ERROR_INVALID_PARAM, AI_ERROR_CODE_TENSOR); return -1; }
  const TensorLoc& L = P.loc[t]; const PBuffer& b = P.buffers[L.buf];
  if (b.is_input) { n->latch(AI_ERROR_INVALID_PARAM, AI_ERROR_CODE_TENSOR); return -1; }
  if (count > n->last_run_n) { n->latch(AI_ERROR_INVALID_PARAM, AI_ERROR_CODE_INVALID_BATCH); return -1; }
  const size_t rows = static_cast<size_t>(count) * b.H * b.W, bytes = rows * L.C;
  if (dst_bytes < bytes) { n->latch(AI_ERROR_INVALID_PARAM, AI_ERROR_CODE_INVALID_SIZE); return -1; }
  const int8_t* src = b.is_output ? pd->d_head : reinterpret_cast<const int8_t*>(pd->d_arena + b.offset * pd->cap);
  if (!cuda_ok(n, cudaMemcpy2DAsync(dst, L.C, src + L.coff, b.CP, L.C, rows, cudaMemcpyDeviceToHost, n->stream), "D2H tensor")) return -1;
  if (!cuda_ok(n, cudaStreamSynchronize(n->stream), "synchronize")) return -1;
  return static_cast<int64_t>(bytes);
}

AI_API_ENTRY int32_t yf_b200_get_stats(ai_handle network, yf_b200_stats* st) {
  YF_NET_OR_FAIL(n, network)
  if (!st) return -1;
  st->kernel_launches = n->launches; st->images = n->images; st->last_run_device_ms = n->last_ms;
  for (size_t i = 1; i < n->members.size(); ++i) { st->kernel_launches += n->members[i]->launches; st->images += n->members[i]->images; }
  st->device = n->device; st->sm_count = n->sm_count; st->chunk_images = n->chunk;
  auto it = n->plans.find(std::make_pair(n->H, n->W));
  st->steps = it == n->plans.end() ? 0 : static_cast<int32_t>(it->second->plan.steps.size());
  st->fused = (it != n->plans.end() && !n->observer && !n->step_profiling && n->mode != 1 && it->second->fprog.ok) ? 1 : 0;
  st->fused_smem_bytes = it == n->plans.end() ? 0 : it->second->fprog.smem_bytes;
  st->fused_latency = (st->fused && it->second->fused_lat) ? 1 : 0;
  st->latency_launches = n->lat_launches; st->cluster_launches = n->cl_launches;
  st->cluster_images = (st->fused_latency && it != n->plans.end()) ? it->second->cluster_images : 0;
  for (size_t i = 1; i < n->members.size(); ++i) { st->latency_launches += n->members[i]->lat_launches; st->cluster_launches += n->members[i]->cl_launches; }
  return 0;
}

AI_API_ENTRY int32_t yf_b200_step_count(ai_handle network) {
  YF_NET_OR_FAIL(n, network)
  if (!n->initialized) return -1;
  PlanDev* pd = get_plan(n, n->H, n->W); if (!pd) return -1;
  return static_cast<int32_t>(pd->plan.steps.size());
}

AI_API_ENTRY int32_t yf_b200_step_info_get(ai_handle network, int32_t step, yf_b200_step_info* info) {
  YF_NET_OR_FAIL(n, network)
  if (!n->initialized || !info) return -1;
  PlanDev* pd = get_plan(n, n->H, n->W); if (!pd) return -1;
  const Plan& P = pd->plan;
  if (step < 0 || step >= static_cast<int>(P.steps.size())) return -1;
  const Step& s = P.steps[step];
  std::memset(info, 0, sizeof *info);
  std::snprintf(info->name, sizeof info->name, "%s", s.name.c_str());
  info->kind = s.kind; info->first_op = s.op_first; info->n_ops = static_cast<int32_t>(s.ops.size());
  const int64_t opix = static_cast<int64_t>(s.Hout) * s.Wout, ipix = static_cast<int64_t>(s.Hin) * s.Win;
  switch (s.kind) {
    case STEP_CONV1X1: info->macs = opix * s.Cout * s.Cin; break;
    case STEP_CONV_IM2COL: info->macs = opix * s.Cout * s.kh * s.kw * s.Cin; break;
    case STEP_DW: info->macs = opix * s.Cout * 9; break;
    default: info->macs = 0;
  }
  info->bytes_read = ipix * s.Cin + (s.add.enabled ? opix * s.Cout : 0);
  info->bytes_written = opix * s.Cout;
  info->last_ms = pd->step_ms[step];
  return 0;
}

// Per-phase cycle trace of the fused kernel (CTA 0, first image): enable, run, then read nphases+1 clock64 stamps.
AI_API_ENTRY int32_t yf_b200_fused_trace(ai_handle network, int32_t enable, int64_t* stamps, int32_t cap) {
  YF_NET_OR_FAIL(n, network)
  if (!n->d_trace && !cuda_ok(n, cudaMalloc(&n->d_trace, sizeof(long long) * 128), "cudaMalloc trace", AI_ERROR_ALLOCATION_FAILED)) return -1;
  n->trace_on = enable != 0;
  if (stamps && cap > 0) {
    if (!cuda_ok(n, cudaStreamSynchronize(n->stream), "synchronize")) return -1;
    const int k = std::min<int>(cap, 128);
    if (!cuda_ok(n, cudaMemcpy(stamps, n->d_trace, sizeof(long long) * k, cudaMemcpyDeviceToHost), "D2H trace")) return -1;
    return k;
  }
  return 0;
}

AI_API_ENTRY int32_t yf_b200_set_step_profiling(ai_handle network, int32_t enable) {
  YF_NET_OR_FAIL(n, network)
  n->step_profiling = enable != 0;
  return 0;
}

// ---- plan introspection (host only) ------------------------------------------------------
// the model the host-only introspection calls describe: YF_B200_TFLITE (a path, re-read on every call) or the embedded one
static bool host_model(TflModel* m) {
  std::string e;
  const char* path = std::getenv("YF_B200_TFLITE");
  if (path && *path) {
    FILE* f = std::fopen(path, "rb");
    if (!f) { set_text(std::string("cannot open ") + path); return false; }
    std::vector<uint8_t> buf; uint8_t tmp[65536]; size_t k;
    while ((k = std::fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + k);
    std::fclose(f);
    if (!m->parse(buf.data(), buf.size(), &e)) { set_text("model: " + e); return false; }
    return true;
  }
  if (!m->parse(yf_embedded_model, yf_embedded_model_len, &e)) { set_text("model: " + e); return false; }
  return true;
}

static bool host_plan(int32_t H, int32_t W, const void* blob, Plan* P) {
  TflModel model;
  if (!host_model(&model)) return false;
  size_t need = 0; st_blob_layout(model, &need);
  std::string perr;
  const bool st = std::getenv("YF_B200_ST_ACTIVATIONS") && std::atoi(std::getenv("YF_B200_ST_ACTIVATIONS"));
  if (!build_plan(model, H, W, static_cast<const uint8_t*>(blob), blob ? need : 0, P, &perr, 4, st)) { set_text("plan: " + perr); return false; }
  return true;
}

AI_API_ENTRY int64_t yf_b200_plan_json(int32_t H, int32_t W, const void* blob, char* dst, uint64_t cap) {
  Plan P;
  if (!host_plan(H, W, blob, &P)) return -1;
  std::string j = "{";
  auto kv = [&](const char* k, long long v, bool comma = true) { j += std::string("\"") + k + "\":" + std::to_string(v) + (comma ? "," : ""); };
  kv("H", P.H); kv("W", P.W); kv("GH", P.GH); kv("GW", P.GW); kv("input_buf", P.input_buf); kv("output_buf", P.output_buf);
  kv("arena_bytes_per_image", static_cast<long long>(P.arena_bytes_per_image));
  kv("arena_bytes_per_image_observer", static_cast<long long>(P.arena_bytes_per_image_observer));
  kv("macs_per_image", P.macs_per_image); kv("n_epi", static_cast<long long>(P.epi.size())); kv("n_luts", static_cast<long long>(P.luts.size() / 256));
  kv("out_zp", P.out_zp);
  j += "\"out_scale\":" + std::to_string(static_cast<double>(P.out_scale)) + ",";
  j += "\"buffers\":[";
  for (size_t i = 0; i < P.buffers.size(); ++i) {
    const PBuffer& b = P.buffers[i];
    j += "{"; kv("H", b.H); kv("W", b.W); kv("C", b.C); kv("CP", b.CP); kv("is_input", b.is_input); kv("is_output", b.is_output);
    kv("observer_only", b.observer_only); kv("offset", static_cast<long long>(b.offset), false); j += i + 1 < P.buffers.size() ? "}," : "}";
  }
  j += "],\"loc\":[";
  for (size_t i = 0; i < P.loc.size(); ++i) {
    j += "[" + std::to_string(P.loc[i].buf) + "," + std::to_string(P.loc[i].coff) + "," + std::to_string(P.loc[i].C) + "]";
    if (i + 1 < P.loc.size()) j += ",";
  }
  j += "],\"steps\":[";
  for (size_t i = 0; i < P.steps.size(); ++i) {
    const Step& s = P.steps[i];
    j += "{\"name\":\"" + s.name + "\",";
    kv("kind", s.kind); kv("op_first", s.op_first);
    j += "\"ops\":["; for (size_t k = 0; k < s.ops.size(); ++k) j += std::to_string(s.ops[k]) + (k + 1 < s.ops.size() ? "," : ""); j += "],";
    kv("in_buf", s.in_buf); kv("in_coff", s.in_coff); kv("add_buf", s.add_buf); kv("add_coff", s.add_coff);
    kv("out_buf", s.out_buf); kv("out_coff", s.out_coff); kv("raw_buf", s.raw_buf); kv("mid_buf", s.mid_buf); kv("pre_add_buf", s.pre_add_buf);
    kv("Hin", s.Hin); kv("Win", s.Win); kv("Cin", s.Cin); kv("Hout", s.Hout); kv("Wout", s.Wout); kv("Cout", s.Cout);
    kv("kh", s.kh); kv("kw", s.kw); kv("stride", s.stride); kv("pad_t", s.pad_t); kv("pad_l", s.pad_l); kv("in_zp", s.in_zp);
    kv("Kpad", s.Kpad); kv("Npad", s.Npad); kv("epi_base", s.epi_base); kv("lut1", s.lut1); kv("lut2", s.lut2); kv("lut_fused", s.lut_fused);
    kv("w_off", static_cast<long long>(s.w_off)); kv("w_bytes", static_cast<long long>(s.w_bytes));
    kv("band_rows", s.band_rows); kv("bands", s.bands);
    j += "\"add\":[" + std::to_string(s.add.enabled) + "," + std::to_string(s.add.zp1) + "," + std::to_string(s.add.zp2) + "," + std::to_string(s.add.zp_out) + "," +
         std::to_string(s.add.m1) + "," + std::to_string(s.add.m2) + "," + std::to_string(s.add.mo) + "," + std::to_string(s.add.s1) + "," +
         std::to_string(s.add.s2) + "," + std::to_string(s.add.so) + "]";
    j += i + 1 < P.steps.size() ? "}," : "}";
  }
  j += "]}";
  if (dst && cap) { size_t k = std::min<size_t>(cap - 1, j.size()); std::memcpy(dst, j.data(), k); dst[k] = 0; }
  return static_cast<int64_t>(j.size() + 1);
}

static bool host_fused(int32_t H, int32_t W, const void* blob, Plan* P, FusedProgram* F, int threads = kFusedWorkerThreads, int cluster = 1) {
  TflModel model;
  if (!host_model(&model)) return false;
  size_t need = 0; st_blob_layout(model, &need);
  std::string perr;
  const bool st = std::getenv("YF_B200_ST_ACTIVATIONS") && std::atoi(std::getenv("YF_B200_ST_ACTIVATIONS"));
  if (!build_plan(model, H, W, static_cast<const uint8_t*>(blob), blob ? need : 0, P, &perr, 16, st)) { set_text("plan: " + perr); return false; }
  build_fused(*P, F, threads, cluster);
  if (!F->ok) { set_text("fused: " + F->why); return false; }
  return true;
}

AI_API_ENTRY int64_t yf_b200_fused_json(int32_t H, int32_t W, const void* blob, char* dst, uint64_t cap) {
  return yf_b200_fused_json_ex(H, W, blob, kFusedWorkerThreads, dst, cap);
}
AI_API_ENTRY int64_t yf_b200_fused_json_ex(int32_t H, int32_t W, const void* blob, int32_t threads, char* dst, uint64_t cap) {
  // threads: 256 | 512, plus 1000 x the cluster size for the cluster shape (4512 = 512-thread CTAs in clusters of 4)
  const int cluster = threads >= 1000 ? threads / 1000 : 1;
  threads %= 1000;
  Plan P; FusedProgram F;
  if (!host_fused(H, W, blob, &P, &F, threads, cluster)) return -1;
  std::string j = "{";
  auto kv = [&](const char* k, long long v, bool comma = true) { j += std::string("\"") + k + "\":" + std::to_string(v) + (comma ? "," : ""); };
  kv("in_off", F.in_off); kv("in_bytes", F.in_bytes); kv("arena_off", F.arena_off); kv("arena_bytes", F.arena_bytes);
  kv("slot_off", F.slot_off); kv("slot_bytes", F.slot_bytes); kv("smem_bytes", F.smem_bytes); kv("head_bytes", F.head_bytes);
  kv("threads", F.threads); kv("cluster", F.cluster); kv("warpgroups", F.threads / 128); kv("tmem_cols", F.tmem_cols); kv("param_slots", kFusedParamSlots); kv("desc_off", F.desc_off);
  kv("in_pf_phase", F.in_pf_phase); kv("split", F.split); kv("smem_bytes_spec", F.smem_bytes_spec); kv("spec", fused_spec_matches(F) ? 1 : 0);
  j += "\"phases\":[";
  for (size_t i = 0; i < F.phases.size(); ++i) {
    const FusedPhase& p = F.phases[i];
    j += "{";
    kv("kind", p.kind); kv("Hin", p.Hin); kv("Win", p.Win); kv("Hout", p.Hout); kv("Wout", p.Wout); kv("rows_in", p.rows_in); kv("rows_out", p.rows_out);
    kv("stride", p.stride); kv("pad_t", p.pad_t); kv("pad_l", p.pad_l); kv("ksize", p.ksize); kv("in_off", p.in_off); kv("in_cs", p.in_cs);
    kv("out_off", p.out_off); kv("out_cs", p.out_cs); kv("add_off", p.add_off); kv("add_cs", p.add_cs); kv("nk", p.nk); kv("npad", p.npad);
    kv("cout", p.cout); kv("chunks_out", p.chunks_out); kv("epi_base", p.epi_base); kv("has_lut", p.has_lut); kv("in_zp", p.in_zp);
    kv("to_global", p.to_global); kv("param_off", p.param_off); kv("param_bytes", p.param_bytes); kv("w_off", p.w_off); kv("lut_off", p.lut_off);
    kv("dw_off", p.dw_off); kv("dwepi_off", p.dwepi_off); kv("epi_off", p.epi_off); kv("scratch_off", p.scratch_off); kv("nw", p.nw); kv("per", p.per); kv("in_wp", p.in_wp); kv("out_wp", p.out_wp); kv("out_zp", p.out_zp);
    kv("in_ws", p.in_ws); kv("out_ws", p.out_ws); kv("scratch_ws", p.scratch_ws); kv("tpg", p.tpg); kv("ntiles", p.ntiles);
    kv("pair", p.pair); kv("sep_y", p.sep_y); kv("rows_a", p.rows_a); kv("row_b0", p.row_b0); kv("rows_single", p.rows_single);
    kv("out_pair_shift", p.out_pair_shift); kv("grp_warps", p.grp_warps); kv("grp_warps_single", p.grp_warps_single);
    kv("own0", p.own[0]); kv("own1", p.own[1]); kv("own_single0", p.own_single[0]); kv("own_single1", p.own_single[1]);
    j += "\"add\":[" + std::to_string(p.add.enabled) + "," + std::to_string(p.add.zp1) + "," + std::to_string(p.add.zp2) + "," + std::to_string(p.add.zp_out) + "," +
         std::to_string(p.add.m1) + "," + std::to_string(p.add.m2) + "," + std::to_string(p.add.mo) + "," + std::to_string(p.add.s1) + "," +
         std::to_string(p.add.s2) + "," + std::to_string(p.add.so) + "]";
    j += i + 1 < F.phases.size() ? "}," : "}";
  }
  j += "]}";
  if (dst && cap) { size_t k = std::min<size_t>(cap - 1, j.size()); std::memcpy(dst, j.data(), k); dst[k] = 0; }
  return static_cast<int64_t>(j.size() + 1);
}

AI_API_ENTRY int64_t yf_b200_plan_blob(int32_t H, int32_t W, const void* blob, int32_t what, void* dst, uint64_t cap) {
  Plan P; FusedProgram F;
  if (what >= 3 ? !host_fused(H, W, blob, &P, &F, what >= 5 ? kFusedLatThreads : kFusedWorkerThreads, what == 6 ? kFusedMaxCluster : 1) : !host_plan(H, W, blob, &P)) return -1;
  const void* src; size_t n;
  switch (what) {
    case 0: src = P.epi.data(); n = P.epi.size() * sizeof(EpiCh); break;
    case 1: src = P.luts.data(); n = P.luts.size(); break;
    case 2: src = P.wblob.data(); n = P.wblob.size(); break;
    case 3: src = F.params.data(); n = F.params.size(); break;
    case 4: src = P.epi.data(); n = P.epi.size() * sizeof(EpiCh); break;
    case 5: case 6: src = F.params.data(); n = F.params.size(); break;
    default: return -1;
  }
  if (dst && cap) std::memcpy(dst, src, std::min<size_t>(cap, n));
  return static_cast<int64_t>(n);
}

AI_API_ENTRY void* yf_b200_host_alloc(uint64_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
AI_API_ENTRY void yf_b200_host_free(void* p) { if (p) cudaFreeHost(p); }

AI_API_ENTRY const char* yf_b200_last_error_text(void) { return g_text.c_str(); }

AI_API_ENTRY int32_t yf_b200_debug_raise(ai_handle network, int32_t code) {
  YF_NET_OR_FAIL(n, network)
  return cuda_ok(n, launch_raise_error(n->d_err, code, n->stream), "raise") ? 0 : -1;
}

}  // extern "C"
