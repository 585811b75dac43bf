// rs_host.cu -- the C ABI of libroadsurf_b200.so (include/roadsurf_b200.h): model upload, the
// device-resident launch, the batched host entry (pack -> H2D -> kernel -> D2H -> unpack, sharded
// over GPUs) and the reference's single-point `runsimulation`.
//
// There is no CPU implementation of the model in this library: every entry point either runs the
// CUDA kernel or returns an error.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "rs_kernel.cuh"
#include "rs_libm.h"

namespace
{
thread_local std::string g_err;
thread_local RsBatchStats g_stats;
thread_local RsLaunchInfo g_launch;
// The model of a device lives in one __constant__ symbol (rs_kernel.cu).  Ordering rule: the symbol is
// only ever overwritten (a) with DIFFERENT bytes and (b) after every kernel that was launched through
// roadsurf_run_device since the previous upload has finished (the asynchronous entry records an event
// per stream; the host entries synchronise their own streams before they return and hold
// g_device_mu while they run).  An upload of identical bytes is skipped.
struct DeviceModel
{
  RsModel m;
  bool valid = false;
  std::map<cudaStream_t, cudaEvent_t> last_launch;  // per caller stream: after its latest step kernel
};
std::mutex g_model_mu;
std::map<int, DeviceModel> g_models;  // device -> model currently in its constant memory
std::atomic<int> g_launches_total{0};

// Run-time options (roadsurf_set_option).  forcing_staging: 1 = full-resolution forcing through the
// per-warp TMA ring in shared memory, 0 = direct read-only loads (default: measured faster).
std::atomic<int> g_opt_staging{-1};
std::atomic<int> g_opt_max_slots{0};  // test hook: cap on points per device batch in roadsurf_run_batch (0 = memory bound)
// coupling_compaction_passes: number of compacted passes over the coupling window before the points
// still iterating are left to finish inside the last launch (0 = one launch, warps repeat in place).
std::atomic<int> g_opt_compaction{6};
// write_back_inputs: roadsurf_run_batch / runsimulation also reproduce the reference's in-place mutation of the
// caller's input arrays (VZ[0], SW_dir, and SW / SW_dir / LW of sky-view points).  Off by default.
std::atomic<int> g_opt_write_back{0};
int opt_compaction_passes() { return g_opt_compaction.load(); }
int opt_staging()
{
  int v = g_opt_staging.load();
  if (v < 0)
  {
    const char* e = std::getenv("ROADSURF_B200_FORCING_STAGING");
    v = (e && e[0] == '1') ? 1 : 0;
    g_opt_staging.store(v);
  }
  return v;
}

int fail(int code, const std::string& msg)
{
  g_err = msg;
  return code;
}

#define CU(call)                                                                                    \
  do                                                                                                \
  {                                                                                                 \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess)                                                                          \
      return fail(RS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                 \
  } while (0)

// One model run on `stream`: a single launch of the step kernel, or -- coupling with lane compaction --
// several.  A warp repeats the coupling window until its slowest lane has converged; when every
// coupled point of the batch has the same window end `wend` (> 0) and the state planes are there,
// the run is split at the window end instead: [step_begin, wend] for everyone, then passes over a
// compacted list of the points that want another iteration (state through the SoA planes), then
// [wend + 1, step_end] with the few points still iterating gathered in warps of their own.  Same
// arithmetic per point, bit-identical results; all launches are queued without a host
// synchronisation.  *launches receives the number of kernels queued.
int launch_model(const RsArgs& a, const RsArgsCold& ac, const RsModel& m, int wend, void* stream, RsLaunchInfo* li,
                 int* launches)
{
  const int passes = opt_compaction_passes();
  *launches = 0;
  if (!(m.use_coupling && wend > 0 && passes > 0 && ac.state && a.scratch && !(opt_staging() && a.forcing_mode == 0) &&
        a.forcing_step0 == 1 && a.step_begin <= 1 && wend + 1 < a.step_end))
  {
    CU(static_cast<cudaError_t>(rs_launch_run(&a, &ac, m.nlayers, opt_staging(), m.use_coupling, m.use_relaxation, m.depth_mode != 0, stream, &li->grid, &li->block,
                                              &li->regs_per_thread, &li->smem_bytes)));
    *launches = 1;
    return RS_OK;
  }
  int* index = reinterpret_cast<int*>(a.scratch + static_cast<size_t>(2 * m.nlayers + 16) * a.ld);
  int* n_index = index + a.ld;
  const double* flags = ac.state + static_cast<size_t>(m.nlayers + 2 + 18) * a.ld;
  RsArgs a1 = a;
  RsArgsCold c1 = ac;
  a1.step_end = wend;
  c1.mode = RS_MODE_SPLIT;
  c1.window_end = wend;
  CU(static_cast<cudaError_t>(rs_launch_run(&a1, &c1, m.nlayers, 0, m.use_coupling, m.use_relaxation, m.depth_mode != 0, stream, &li->grid, &li->block,
                                            &li->regs_per_thread, &li->smem_bytes)));
  RsArgs a2 = a;
  RsArgsCold c2 = c1;
  a2.step_begin = wend + 1;
  a2.step_end = wend;
  c2.mode = RS_MODE_SPLIT | RS_MODE_ONE_PASS;
  c2.index = index;
  c2.n_index = n_index;
  for (int k = 0; k < passes; ++k)
  {
    CU(static_cast<cudaError_t>(rs_launch_partition(flags, a.ld, a.npoints, 0, index, n_index, stream)));
    CU(static_cast<cudaError_t>(rs_launch_run(&a2, &c2, m.nlayers, 0, m.use_coupling, m.use_relaxation, m.depth_mode != 0, stream, &li->grid, &li->block,
                                              &li->regs_per_thread, &li->smem_bytes)));
  }
  CU(static_cast<cudaError_t>(rs_launch_partition(flags, a.ld, a.npoints, 1, index, n_index, stream)));
  RsArgs a3 = a;
  RsArgsCold c3 = c1;
  a3.step_begin = wend + 1;
  c3.index = index;
  c3.n_index = n_index;
  CU(static_cast<cudaError_t>(rs_launch_run(&a3, &c3, m.nlayers, 0, m.use_coupling, m.use_relaxation, m.depth_mode != 0, stream, &li->grid, &li->block,
                                            &li->regs_per_thread, &li->smem_bytes)));
  *launches = 2 * passes + 3;
  return RS_OK;
}

double now_ms()
{
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

int set_model_on_current_device(const InputSettings* settings, const InputParameters* params, RsModel* out)
{
  RsModel m;
  char err[256];
  const int rc = rs_build_model(settings, params, &m, err, sizeof err);
  if (rc != RS_OK) return fail(rc, err);
  int dev = 0;
  CU(cudaGetDevice(&dev));
  {
    std::lock_guard<std::mutex> lk(g_model_mu);
    DeviceModel& dm = g_models[dev];
    if (!dm.valid || std::memcmp(&dm.m, &m, sizeof m) != 0)
    {
      // kernels of the asynchronous entry may still be reading the old model: wait for them
      for (auto& kv : dm.last_launch) CU(cudaEventSynchronize(kv.second));
      CU(static_cast<cudaError_t>(rs_upload_model(&m)));
      dm.m = m;
      dm.valid = true;
    }
  }
  if (out) *out = m;
  return RS_OK;
}

// roadsurf_run_device: remember that `stream` has kernels in flight that read the device's model.
int note_async_launch(int dev, cudaStream_t stream)
{
  std::lock_guard<std::mutex> lk(g_model_mu);
  DeviceModel& dm = g_models[dev];
  cudaEvent_t& ev = dm.last_launch[stream];
  if (!ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CU(cudaEventRecord(ev, stream));
  return RS_OK;
}

// Run fn(k) for k in [0, n) on up to `nthreads` host threads (the caller's thread included).
void parallel_for(int n, int nthreads, const std::function<void(int)>& fn)
{
  if (n <= 0) return;
  nthreads = std::max(1, std::min(nthreads, n));
  if (nthreads == 1)
  {
    for (int k = 0; k < n; ++k) fn(k);
    return;
  }
  std::atomic<int> next(0);
  auto work = [&]() {
    for (;;)
    {
      const int k = next.fetch_add(1);
      if (k >= n) break;
      fn(k);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
}

int host_threads()
{
  const unsigned hc = std::thread::hardware_concurrency();
  return static_cast<int>(std::max(1u, std::min(hc ? hc : 4u, 32u)));
}

// What the kernel will decide about a point's coupling window (src/InputOutput.f90:34-36,
// src/Coupling.f90:510-519): used only to order slots so that every warp has one window.
long long window_key(const RsModel& m, const LocalParameters& lp)
{
  if (!m.use_coupling || lp.couplingTsurf < -100 || lp.couplingIndexI < 1) return -1;
  return lp.couplingIndexI;
}

struct Shard
{
  int first = 0, count = 0, device = 0;
  int rc = RS_OK;
  std::string err;
  RsBatchStats stats{};
  RsLaunchInfo launch{};
};

// Grow-only device and pinned-host work space, one buffer per (device, slot), reused across calls so
// that a steady-state call does no cudaMalloc / cudaMallocHost (both cost tens to hundreds of ms at
// these sizes).
struct PoolEntry
{
  void* p = nullptr;
  size_t cap = 0;
};
std::mutex g_pool_mu;
std::map<std::pair<int, int>, PoolEntry> g_pool;

cudaError_t pool_get(int device, int slot, size_t bytes, void** out)
{
  std::lock_guard<std::mutex> lk(g_pool_mu);
  PoolEntry& e = g_pool[{device, slot}];
  if (e.cap < bytes)
  {
    if (e.p) cudaFree(e.p);
    e.p = nullptr;
    e.cap = 0;
    cudaError_t rc = cudaMalloc(&e.p, bytes);
    if (rc != cudaSuccess) return rc;
    e.cap = bytes;
  }
  *out = e.p;
  return cudaSuccess;
}

std::map<std::pair<int, int>, PoolEntry> g_pinned_pool;

// The pooled buffers of a device are used by one call at a time: concurrent callers (the reference
// calls runsimulation from N host threads) are serialised per device, as the GPU would do anyway.
std::mutex g_device_mu[64];

cudaError_t pinned_get(int device, int slot, size_t bytes, void** out)
{
  std::lock_guard<std::mutex> lk(g_pool_mu);
  PoolEntry& e = g_pinned_pool[{device, slot}];
  if (e.cap < bytes)
  {
    if (e.p) cudaFreeHost(e.p);
    e.p = nullptr;
    e.cap = 0;
    cudaError_t rc = cudaMallocHost(&e.p, bytes);
    if (rc != cudaSuccess) return rc;
    e.cap = bytes;
  }
  *out = e.p;
  return cudaSuccess;
}

// Non-owning views with the DeviceMem / PinnedMem interface, backed by the pools above.
struct PooledDevice
{
  void* p = nullptr;
  cudaError_t alloc(int device, int slot, size_t bytes) { return pool_get(device, slot, bytes ? bytes : 8, &p); }
  template <class T>
  T* as() const
  {
    return static_cast<T*>(p);
  }
};
struct PooledPinned
{
  void* p = nullptr;
  cudaError_t alloc(int device, int slot, size_t bytes) { return pinned_get(device, slot, bytes ? bytes : 8, &p); }
  template <class T>
  T* as() const
  {
    return static_cast<T*>(p);
  }
};

// Run points [first, first+count) of the batch on `device`.
int run_shard(Shard& sh, OutputPointers* const* out, const InputPointers* const* in,
              const InputSettings* settings, const InputParameters* params,
              const LocalParameters* const* local, int* status)
{
  std::lock_guard<std::mutex> device_lock(g_device_mu[sh.device & 63]);
  const double setup0 = now_ms();
  CU(cudaSetDevice(sh.device));
  RsModel model;
  {
    const int rc = set_model_on_current_device(settings, params, &model);
    if (rc != RS_OK) return rc;
  }
  const int sim_len = settings->SimLen;
  const int nl = model.nlayers;
  cudaStream_t stream;
  CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  struct StreamGuard
  {
    cudaStream_t s;
    ~StreamGuard() { cudaStreamDestroy(s); }
  } sguard{stream};

  // ---- group points by time axis; inside a group order slots by coupling window ----------------
  struct Group
  {
    const InputPointers* rep;
    std::vector<int> points;
  };
  std::vector<Group> groups;
  for (int q = 0; q < sh.count; ++q)
  {
    const int p = sh.first + q;
    const InputPointers* ip = in[p];
    if (ip->inputLen < sim_len || out[p]->outputLen < sim_len)
      return fail(RS_ERR_BAD_ARGUMENT, "inputLen/outputLen shorter than SimLen");
    bool placed = false;
    for (auto& g : groups)
    {
      const InputPointers* r = g.rep;
      const bool same_ptr = r->c_year == ip->c_year && r->c_month == ip->c_month && r->c_day == ip->c_day &&
                            r->c_hour == ip->c_hour && r->c_minute == ip->c_minute && r->c_second == ip->c_second;
      const size_t nb = sizeof(int) * sim_len;
      if (same_ptr || (!std::memcmp(r->c_year, ip->c_year, nb) && !std::memcmp(r->c_month, ip->c_month, nb) &&
                       !std::memcmp(r->c_day, ip->c_day, nb) && !std::memcmp(r->c_hour, ip->c_hour, nb) &&
                       !std::memcmp(r->c_minute, ip->c_minute, nb) && !std::memcmp(r->c_second, ip->c_second, nb)))
      {
        g.points.push_back(p);
        placed = true;
        break;
      }
    }
    if (!placed) groups.push_back(Group{ip, {p}});
  }

  size_t free_b = 0, total_b = 0;
  CU(cudaMemGetInfo(&free_b, &total_b));

  for (auto& g : groups)
  {
    // slots: windows sorted, each window padded to a warp boundary; uncoupled points fill the gaps
    std::map<long long, std::vector<int>> by_window;
    for (int p : g.points) by_window[window_key(model, *local[p])].push_back(p);
    // inside a window, points with sky-view radiation go last: the solar geometry and the horizon
    // lookup are skipped by warps without such a point (7 % on a grid with 30 % of them scattered)
    for (auto& kv : by_window)
      std::stable_partition(kv.second.begin(), kv.second.end(), [&](int p) {
        const double sv = local[p]->sky_view;
        return !(sv < 1.0 && sv > -0.01f);
      });
    std::vector<int> slots;  // point index or -1 (padding)
    std::vector<int> loose;
    if (by_window.count(-1)) loose = by_window[-1];
    size_t loose_pos = 0;
    for (auto& kv : by_window)
    {
      if (kv.first == -1) continue;
      for (int p : kv.second) slots.push_back(p);
      while (slots.size() % 32 != 0)
      {
        if (loose_pos < loose.size())
          slots.push_back(loose[loose_pos++]);
        else
          slots.push_back(-1);
      }
    }
    while (loose_pos < loose.size()) slots.push_back(loose[loose_pos++]);
    while (slots.size() % 32 != 0) slots.push_back(-1);
    ++sh.stats.groups;
    // one coupling window in the whole group (the usual case: one analysis time): lane compaction
    // between coupling iterations (launch_model); needs the per-point state planes
    int wend = 0;
    {
      int nwin = 0;
      for (auto& kv : by_window)
        if (kv.first != -1)
        {
          ++nwin;
          wend = static_cast<int>(kv.first);
        }
      if (nwin != 1 || opt_compaction_passes() < 1) wend = 0;
    }

    std::atomic<bool> any_depth_a(false), any_sky_a(false);
    parallel_for(static_cast<int>(g.points.size()), host_threads(), [&](int k) {
      const int p = g.points[k];
      const LocalParameters& lp = *local[p];
      if (lp.sky_view < 1.0 && lp.sky_view > -0.01f) any_sky_a = true;
      if (!any_depth_a.load(std::memory_order_relaxed))  // depth(1), depth(SimLen) are read even with tsurfOutputDepth
      {
        const double* d = in[p]->c_Depth;
        for (int t = 0; t < sim_len; ++t)
          if (d[t] >= 0.0)
          {
            any_depth_a = true;
            break;
          }
      }
    });
    const bool any_depth = any_depth_a, any_sky = any_sky_a;
    const int nvar = any_depth ? RS_F_NVAR_DEPTH : RS_F_NVAR;

    // ---- device batches that fit the memory budget -------------------------------------------
    const size_t per_slot = sizeof(double) * (static_cast<size_t>(sim_len) * (nvar + RS_O_NVAR) + RS_L_NLOCAL +
                                              (any_sky ? 360 : 0) + RS_SCRATCH_NPLANES(nl) +
                                              (wend > 0 ? RS_STATE_NPLANES(nl) : 0)) + sizeof(int);
    // Two buffer sets are in flight (see the pipeline below), and a large group is cut into at least
    // four batches so that copying batch b back overlaps packing and uploading batch b+1.
    const size_t budget = static_cast<size_t>(free_b * 0.80);
    size_t max_slots = budget / 2 / per_slot / 32 * 32;
    if (max_slots < 32) return fail(RS_ERR_CUDA, "not enough device memory for one warp of points");
    {
      const size_t n = slots.size();
      const size_t parts = n >= 4 * 8192 ? 4 : (n >= 2 * 4096 ? 2 : 1);
      max_slots = std::min(max_slots, (n / parts + 31) / 32 * 32);
    }
    if (const int cap = g_opt_max_slots.load()) max_slots = std::min<size_t>(max_slots, (cap + 31) / 32 * 32);
    // staging chunk: <= 192 MiB of pinned memory per direction
    const size_t stage_budget = 192ull << 20;
    int chunk = static_cast<int>(std::max<size_t>(32, stage_budget / (sizeof(double) * sim_len * nvar) / 32 * 32));

    PooledDevice d_tf;
    CU(d_tf.alloc(sh.device, 209, sizeof(int) * 6 * sim_len));
    {
      const int* src[6] = {g.rep->c_year, g.rep->c_month, g.rep->c_day, g.rep->c_hour, g.rep->c_minute,
                           g.rep->c_second};
      for (int k = 0; k < 6; ++k)
        CU(cudaMemcpyAsync(d_tf.as<int>() + static_cast<size_t>(k) * sim_len, src[k], sizeof(int) * sim_len,
                           cudaMemcpyHostToDevice, stream));
      sh.stats.h2d_bytes += sizeof(int) * 6 * sim_len;
    }
    PooledDevice d_solar;
    CU(d_solar.alloc(sh.device, 210, sizeof(double) * 4 * sim_len));
    CU(static_cast<cudaError_t>(rs_launch_solar(d_tf.as<int>(), sim_len, d_solar.as<double>(), stream)));
    ++sh.stats.kernel_launches;

    // ---- software pipeline over the device batches: while a helper thread copies batch b back and
    // scatters it into the caller's arrays (its kernels may still be running), the calling thread packs
    // and uploads batch b+1 into the other buffer set.  Each set has its own stream.
    struct Ctx
    {
      PooledDevice d_forcing, d_out, d_local, d_hor, d_status, d_scratch, d_stage, d_stage2, d_counters, d_state;
      PooledPinned h_stage, h_stage2, h_local, h_hor, h_status;
      cudaStream_t st = nullptr;
      cudaEvent_t ev_free[2] = {nullptr, nullptr};                       // staging buffer reuse
      cudaEvent_t ev_in0 = nullptr, ev_in1 = nullptr, ev_k1 = nullptr, ev_o1 = nullptr;
      size_t s0 = 0;
      int ld = 0, chunk = 0;
      RsBatchStats out_stats{};   // what the helper thread measured
      int rc = RS_OK;
      std::string err;
      std::thread drain;
    };
    Ctx ctx[2];
    struct CtxGuard  // no thread may outlive its buffers; streams and events are per group
    {
      Ctx* c;
      ~CtxGuard()
      {
        for (int k = 0; k < 2; ++k)
        {
          if (c[k].drain.joinable()) c[k].drain.join();
          for (cudaEvent_t e : {c[k].ev_free[0], c[k].ev_free[1], c[k].ev_in0, c[k].ev_in1, c[k].ev_k1, c[k].ev_o1})
            if (e) cudaEventDestroy(e);
          if (c[k].st) cudaStreamDestroy(c[k].st);
        }
      }
    } ctx_guard{ctx};
    const int nthreads = host_threads();
    const size_t max_ld = std::min(max_slots, slots.size());
    const size_t nbatches = (slots.size() + max_slots - 1) / max_slots;
    const int dv = sh.device;
    for (int k = 0; k < (nbatches > 1 ? 2 : 1); ++k)
    {
      Ctx& c = ctx[k];
      const int o = 50 * k;  // pool slots of the second buffer set
      CU(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
      for (cudaEvent_t* e : {&c.ev_free[0], &c.ev_free[1]}) CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
      for (cudaEvent_t* e : {&c.ev_in0, &c.ev_in1, &c.ev_k1, &c.ev_o1}) CU(cudaEventCreate(e));
      c.chunk = static_cast<int>(std::min<size_t>(chunk, max_ld));
      CU(c.d_forcing.alloc(dv, 200 + o, sizeof(double) * sim_len * nvar * max_ld));
      CU(c.d_out.alloc(dv, 201 + o, sizeof(double) * RS_O_NVAR * sim_len * max_ld));
      CU(c.d_local.alloc(dv, 202 + o, sizeof(double) * RS_L_NLOCAL * max_ld));
      if (any_sky) CU(c.d_hor.alloc(dv, 203 + o, sizeof(double) * 360 * max_ld));
      CU(c.d_status.alloc(dv, 204 + o, sizeof(int) * max_ld));
      if (model.use_coupling) CU(c.d_scratch.alloc(dv, 205 + o, sizeof(double) * RS_SCRATCH_NPLANES(nl) * max_ld));
      if (wend > 0) CU(c.d_state.alloc(dv, 211 + o, sizeof(double) * RS_STATE_NPLANES(nl) * max_ld));
      const size_t stage_bytes =
          sizeof(double) * static_cast<size_t>(c.chunk) * sim_len * std::max(nvar, static_cast<int>(RS_O_NVAR));
      CU(c.d_stage.alloc(dv, 206 + o, stage_bytes));
      CU(c.d_stage2.alloc(dv, 207 + o, stage_bytes));
      CU(c.d_counters.alloc(dv, 208 + o, sizeof(unsigned long long) * RS_CNT_N));
      CU(c.h_stage.alloc(dv, 0 + o, stage_bytes));
      CU(c.h_stage2.alloc(dv, 1 + o, stage_bytes));
      CU(c.h_local.alloc(dv, 2 + o, sizeof(double) * RS_L_NLOCAL * max_ld));
      if (any_sky) CU(c.h_hor.alloc(dv, 3 + o, sizeof(double) * 360 * max_ld));
      CU(c.h_status.alloc(dv, 4 + o, sizeof(int) * max_ld));
    }
    // the time axis and the solar table were queued on `stream`: every batch stream reads them
    CU(cudaStreamSynchronize(stream));
    if (sh.stats.setup_ms == 0.0) sh.stats.setup_ms = now_ms() - setup0 - sh.stats.pack_ms;  // everything so far but packing

    // ---- input stage (calling thread): statics, then the forcing: caller's per-point arrays -> pinned
    // [var][point][time] -> device -> SoA.  Rows are packed by all host threads into one of two pinned
    // buffers while the previous buffer is in flight (H2D copy + device-side transpose).
    auto stage_in = [&](Ctx& c) -> int {
      const int ld = c.ld;
      const size_t s0 = c.s0;
      cudaStream_t st = c.st;
      double* h_st[2] = {c.h_stage.as<double>(), c.h_stage2.as<double>()};
      double* d_st[2] = {c.d_stage.as<double>(), c.d_stage2.as<double>()};
      CU(cudaMemsetAsync(c.d_counters.p, 0, sizeof(unsigned long long) * RS_CNT_N, st));
      double t0 = now_ms();
      {
        double* L = c.h_local.as<double>();
        for (int q = 0; q < ld; ++q)
        {
          const int p = slots[s0 + q];
          LocalParameters lp;
          std::memset(&lp, 0, sizeof lp);
          if (p >= 0) lp = *local[p];
          L[RS_L_TAIR_RELAX * (size_t)ld + q] = lp.tair_relax;
          L[RS_L_VZ_RELAX * (size_t)ld + q] = lp.VZ_relax;
          L[RS_L_RH_RELAX * (size_t)ld + q] = lp.RH_relax;
          L[RS_L_COUPLING_TSURF * (size_t)ld + q] = lp.couplingTsurf;
          L[RS_L_LAT * (size_t)ld + q] = lp.lat;
          L[RS_L_LON * (size_t)ld + q] = lp.lon;
          L[RS_L_SKY_VIEW * (size_t)ld + q] = (p >= 0) ? lp.sky_view : 1.0;
          L[RS_L_COUPLING_INDEX * (size_t)ld + q] = lp.couplingIndexI;
          L[RS_L_INIT_LEN * (size_t)ld + q] = lp.InitLenI;
          L[RS_L_ACTIVE * (size_t)ld + q] = (p >= 0) ? 1.0 : 0.0;
        }
        if (any_sky)
        {
          double* H = c.h_hor.as<double>();
          for (int q = 0; q < ld; ++q)
          {
            const int p = slots[s0 + q];
            const double* src = (p >= 0) ? in[p]->c_local_horizons : nullptr;
            for (int k = 0; k < 360; ++k) H[static_cast<size_t>(k) * ld + q] = src ? src[k] : 0.0;
          }
        }
      }
      sh.stats.pack_ms += now_ms() - t0;
      CU(cudaEventRecord(c.ev_in0, st));
      CU(cudaMemcpyAsync(c.d_local.p, c.h_local.p, sizeof(double) * RS_L_NLOCAL * ld, cudaMemcpyHostToDevice, st));
      sh.stats.h2d_bytes += sizeof(double) * RS_L_NLOCAL * ld;
      if (any_sky)
      {
        CU(cudaMemcpyAsync(c.d_hor.p, c.h_hor.p, sizeof(double) * 360 * ld, cudaMemcpyHostToDevice, st));
        sh.stats.h2d_bytes += sizeof(double) * 360 * ld;
      }
      int cidx = 0;
      for (int q0 = 0; q0 < ld; q0 += c.chunk, ++cidx)
      {
        const int npc = std::min(c.chunk, ld - q0);
        const int bsel = cidx & 1;
        if (cidx >= 2) CU(cudaEventSynchronize(c.ev_free[bsel]));
        t0 = now_ms();
        double* S = h_st[bsel];
        const size_t plane = static_cast<size_t>(npc) * sim_len;
        parallel_for(npc, nthreads, [&](int q) {
          const int p = slots[s0 + q0 + q];
          const size_t row = static_cast<size_t>(q) * sim_len;
          if (p < 0)
          {
            for (int v = 0; v < nvar; ++v) std::memset(S + v * plane + row, 0, sizeof(double) * sim_len);
            return;
          }
          const InputPointers* ip = in[p];
          const double* src[RS_F_NVAR_DEPTH] = {ip->c_tair, ip->c_tdew, ip->c_VZ,     ip->c_Rhz,
                                                ip->c_prec, ip->c_SW,   ip->c_LW,     ip->c_SW_dir,
                                                ip->c_LW_net, ip->c_TSurfObs, nullptr, ip->c_Depth};
          for (int v = 0; v < nvar; ++v)
          {
            double* dst = S + v * plane + row;
            if (v == RS_F_PHASE)
              for (int t = 0; t < sim_len; ++t) dst[t] = static_cast<double>(ip->c_PrecPhase[t]);
            else
              std::memcpy(dst, src[v], sizeof(double) * sim_len);
          }
        });
        sh.stats.pack_ms += now_ms() - t0;
        const size_t bytes = sizeof(double) * plane * nvar;
        CU(cudaMemcpyAsync(d_st[bsel], S, bytes, cudaMemcpyHostToDevice, st));
        CU(static_cast<cudaError_t>(
            rs_launch_pack_forcing(d_st[bsel], npc, q0, sim_len, nvar, c.d_forcing.as<double>(), ld, st)));
        CU(cudaEventRecord(c.ev_free[bsel], st));
        ++sh.stats.kernel_launches;
        sh.stats.h2d_bytes += bytes;
      }
      CU(cudaEventRecord(c.ev_in1, st));

      // ---- the step kernel(s), queued behind the input stage
      RsArgs a;
      std::memset(&a, 0, sizeof a);
      a.npoints = ld;
      a.ld = ld;
      a.sim_len = sim_len;
      a.n_records = sim_len;
      a.nvar = nvar;
      a.out_stride = 1;
      a.n_out = sim_len;
      a.forcing_mode = 0;
      a.step_begin = 1;
      a.step_end = sim_len;
      a.forcing_step0 = 1;
      a.out_slot0 = 0;
      a.forcing = c.d_forcing.as<double>();
      a.tf = d_tf.as<int>();
      a.local = c.d_local.as<double>();
      a.horizons = any_sky ? c.d_hor.as<double>() : nullptr;
      a.solar = d_solar.as<double>();
      a.out = c.d_out.as<double>();
      a.status = c.d_status.as<int>();
      a.scratch = model.use_coupling ? c.d_scratch.as<double>() : nullptr;
      RsArgsCold ac;
      std::memset(&ac, 0, sizeof ac);
      ac.counters = c.d_counters.as<unsigned long long>();
      ac.out_start = 0;
      ac.out_nvar = RS_O_NVAR;
      ac.state = wend > 0 ? c.d_state.as<double>() : nullptr;
      {
        int n = 0;
        const int rc = launch_model(a, ac, model, wend, st, &sh.launch, &n);
        if (rc != RS_OK) return rc;
        sh.stats.kernel_launches += n;
      }
      CU(cudaEventRecord(c.ev_k1, st));
      sh.launch.nlayers = nl;
      sh.launch.forcing_mode = 0;
      return RS_OK;
    };

    // ---- output stage (helper thread): SoA -> [var][point][time] -> pinned -> caller's arrays, double
    // buffered: chunk c-1 is scattered into the caller's arrays while chunk c is copied back
    auto stage_out = [&](Ctx& c) -> int {
      CU(cudaSetDevice(dv));
      const int ld = c.ld;
      const size_t s0 = c.s0;
      cudaStream_t st = c.st;
      double* h_st[2] = {c.h_stage.as<double>(), c.h_stage2.as<double>()};
      double* d_st[2] = {c.d_stage.as<double>(), c.d_stage2.as<double>()};
      RsBatchStats& os = c.out_stats;
      auto scatter = [&](int q0, int npc, const double* S) {
        const size_t plane = static_cast<size_t>(npc) * sim_len;
        parallel_for(npc, nthreads, [&](int q) {
          const int p = slots[s0 + q0 + q];
          if (p < 0) return;
          OutputPointers* op = out[p];
          double* dst[RS_O_NVAR] = {op->c_TsurfOut, op->c_SnowOut, op->c_WaterOut,
                                    op->c_IceOut,   op->c_DepositOut, op->c_Ice2Out};
          for (int v = 0; v < RS_O_NVAR; ++v)
            std::memcpy(dst[v], S + v * plane + static_cast<size_t>(q) * sim_len, sizeof(double) * sim_len);
        });
      };
      int cidx = 0, prev_q0 = -1, prev_npc = 0;
      for (int q0 = 0; q0 < ld; q0 += c.chunk, ++cidx)
      {
        const int npc = std::min(c.chunk, ld - q0);
        const int bsel = cidx & 1;
        const size_t bytes = sizeof(double) * static_cast<size_t>(npc) * sim_len * RS_O_NVAR;
        CU(static_cast<cudaError_t>(rs_launch_unpack_out(c.d_out.as<double>(), ld, sim_len, q0, npc, d_st[bsel], st)));
        ++os.kernel_launches;
        CU(cudaMemcpyAsync(h_st[bsel], d_st[bsel], bytes, cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(c.ev_free[bsel], st));
        os.d2h_bytes += bytes;
        if (prev_q0 >= 0)
        {
          // buffer of the previous chunk: its copy was enqueued before this one
          CU(cudaEventSynchronize(c.ev_free[bsel ^ 1]));
          const double t0 = now_ms();
          scatter(prev_q0, prev_npc, h_st[bsel ^ 1]);
          os.unpack_ms += now_ms() - t0;
        }
        prev_q0 = q0;
        prev_npc = npc;
      }
      CU(cudaEventRecord(c.ev_o1, st));
      if (prev_q0 >= 0)
      {
        CU(cudaEventSynchronize(c.ev_free[(cidx - 1) & 1]));
        const double t0 = now_ms();
        scatter(prev_q0, prev_npc, h_st[(cidx - 1) & 1]);
        os.unpack_ms += now_ms() - t0;
      }
      if (g_opt_write_back.load())
      {
        // the reference's caller-visible input mutations (see rs_mutation_kernel), chunk by chunk through the
        // first staging buffer; VZ(1) is clamped on the host (src/Initialization.f90:121-123)
        PooledDevice d_nvis;
        CU(d_nvis.alloc(dv, 212 + (&c == &ctx[1] ? 50 : 0), sizeof(int) * ld));
        for (int q0 = 0; q0 < ld; q0 += c.chunk)
        {
          const int npc = std::min(c.chunk, ld - q0);
          const size_t plane = static_cast<size_t>(npc) * sim_len;
          CU(static_cast<cudaError_t>(rs_launch_mutation(c.d_forcing.as<double>(), nvar, ld, sim_len, c.d_local.as<double>(),
                                                         any_sky ? c.d_hor.as<double>() : nullptr, d_solar.as<double>(),
                                                         c.d_out.as<double>(), d_nvis.as<int>(), q0, npc, d_st[0], st)));
          os.kernel_launches += q0 == 0 ? 2 : 1;
          CU(cudaMemcpyAsync(h_st[0], d_st[0], sizeof(double) * 3 * plane, cudaMemcpyDeviceToHost, st));
          CU(cudaStreamSynchronize(st));
          os.d2h_bytes += sizeof(double) * 3 * plane;
          const double* S = h_st[0];
          parallel_for(npc, nthreads, [&](int q) {
            const int p = slots[s0 + q0 + q];
            if (p < 0) return;
            const InputPointers* ip = in[p];
            double* dst[3] = {ip->c_SW, ip->c_SW_dir, ip->c_LW};
            for (int v = 0; v < 3; ++v)
              std::memcpy(dst[v], S + v * plane + static_cast<size_t>(q) * sim_len, sizeof(double) * sim_len);
            if (ip->c_VZ[0] < 0.4f) ip->c_VZ[0] = 0.4f;
          });
        }
      }
      CU(cudaMemcpyAsync(c.h_status.p, c.d_status.p, sizeof(int) * ld, cudaMemcpyDeviceToHost, st));
      unsigned long long cnt[RS_CNT_N];
      CU(cudaMemcpyAsync(cnt, c.d_counters.p, sizeof cnt, cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      float ms = 0.f;
      cudaEventElapsedTime(&ms, c.ev_in0, c.ev_in1);
      os.h2d_ms += ms;  // device-side time of the input stage (copies + transposes, overlapped with packing)
      cudaEventElapsedTime(&ms, c.ev_in1, c.ev_k1);
      os.kernel_ms += ms;
      cudaEventElapsedTime(&ms, c.ev_k1, c.ev_o1);
      os.d2h_ms += ms;
      os.d2h_bytes += sizeof(int) * ld;
      os.executed_steps += static_cast<int64_t>(cnt[RS_CNT_EXECUTED_STEPS]);
      if (status)
        for (int q = 0; q < ld; ++q)
        {
          const int p = slots[s0 + q];
          if (p >= 0) status[p] = c.h_status.as<int>()[q];
        }
      return RS_OK;
    };
    // joins the helper thread of a buffer set and folds its measurements into the shard's
    auto finish = [&](Ctx& c) -> int {
      if (!c.drain.joinable()) return RS_OK;
      c.drain.join();
      sh.stats.h2d_ms += c.out_stats.h2d_ms;
      sh.stats.kernel_ms += c.out_stats.kernel_ms;
      sh.stats.d2h_ms += c.out_stats.d2h_ms;
      sh.stats.unpack_ms += c.out_stats.unpack_ms;
      sh.stats.d2h_bytes += c.out_stats.d2h_bytes;
      sh.stats.executed_steps += c.out_stats.executed_steps;
      sh.stats.kernel_launches += c.out_stats.kernel_launches;
      c.out_stats = RsBatchStats{};
      if (c.rc != RS_OK) return fail(c.rc, c.err);
      return RS_OK;
    };

    // The helper threads reference this frame's lambdas and locals: whatever happens in the loop, they are
    // joined (and, after an error, the batch streams drained: pooled buffers may still be in use by queued
    // kernels and copies) before anything here goes out of scope.
    const int loop_rc = [&]() -> int {
      size_t b = 0;
      for (size_t s0 = 0; s0 < slots.size(); s0 += max_slots, ++b)
      {
        Ctx& c = ctx[b & 1];
        {
          const int rc = finish(c);  // the batch that used this buffer set two rounds ago
          if (rc != RS_OK) return rc;
        }
        c.s0 = s0;
        c.ld = static_cast<int>(std::min(max_slots, slots.size() - s0));
        {
          const int rc = stage_in(c);
          if (rc != RS_OK) return rc;
        }
        c.rc = RS_OK;
        c.drain = std::thread([&stage_out, &c]() {
          c.rc = stage_out(c);
          if (c.rc != RS_OK) c.err = g_err;
        });
      }
      for (int k = 0; k < 2; ++k)
      {
        const int rc = finish(ctx[k]);
        if (rc != RS_OK) return rc;
      }
      return RS_OK;
    }();
    if (loop_rc != RS_OK)
    {
      const std::string err = g_err;
      for (int k = 0; k < 2; ++k)
      {
        if (ctx[k].drain.joinable()) ctx[k].drain.join();
        if (ctx[k].st) cudaStreamSynchronize(ctx[k].st);
      }
      return fail(loop_rc, err);
    }
  }
  return RS_OK;
}
// Column chunks of a host-SoA shard: a multiple of 32 points, about 1/8 of the shard but at least 64 Ki points.
int host_soa_chunk(int count)
{
  int chunk = ((count + 7) / 8 + 31) / 32 * 32;
  if (chunk < 65536) chunk = std::min((count + 31) / 32 * 32, 65536);
  return chunk;
}

// roadsurf_prepare_statics: per device shard and per chunk of roadsurf_run_host_soa's pipeline, the local
// horizon table [360][ld] and the sky-view point order [ld], resident on the device.
struct StaticsChunk
{
  double* hor = nullptr;  // [360][ld] or null (no horizons given)
  int* order = nullptr;   // [ld]
  int ld = 0;
};
struct StaticsShard
{
  int device = 0, first = 0, count = 0, chunk = 0;
  std::vector<StaticsChunk> chunks;
};
struct RsStatics
{
  unsigned magic = 0x52535354u;  // "RSST"
  int npoints = 0, ngpus = 0;
  bool has_horizons = false;
  std::vector<StaticsShard> shards;
};

// Streams and events of the host-SoA pipeline, created once per device and reused by every call (the
// calls of one device are serialised by g_device_mu).
struct SoaStreams
{
  cudaStream_t st[2] = {nullptr, nullptr};
  cudaEvent_t ev[2][4] = {};
};
std::mutex g_soa_mu;
std::map<int, SoaStreams> g_soa_streams;
int soa_streams(int device, SoaStreams** out)
{
  std::lock_guard<std::mutex> lk(g_soa_mu);
  SoaStreams& s = g_soa_streams[device];
  if (!s.st[0])
    for (int k = 0; k < 2; ++k)
    {
      CU(cudaStreamCreateWithFlags(&s.st[k], cudaStreamNonBlocking));
      for (int e = 0; e < 4; ++e) CU(cudaEventCreate(&s.ev[k][e]));
    }
  *out = &s;
  return RS_OK;
}

// Points [first, first+count) of a host SoA batch on `device`, pipelined in column chunks over two
// streams: H2D(chunk c+1) overlaps kernel(chunk c) overlaps D2H(chunk c-1).
int run_host_soa_shard(Shard& sh, const RsHostBatch* b, const InputSettings* settings,
                       const InputParameters* params)
{
  std::lock_guard<std::mutex> device_lock(g_device_mu[sh.device & 63]);
  CU(cudaSetDevice(sh.device));
  RsModel model;
  {
    const int rc = set_model_on_current_device(settings, params, &model);
    if (rc != RS_OK) return rc;
  }
  const int nl = model.nlayers;
  const int n_out = (b->sim_len + b->out_stride - 1) / b->out_stride;
  const size_t np_all = static_cast<size_t>(b->npoints);
  constexpr int NSTREAM = 2;
  const int chunk = host_soa_chunk(sh.count);
  const int nchunks = (sh.count + chunk - 1) / chunk;
  const size_t frows = static_cast<size_t>(b->n_records) * b->nvar;
  const size_t orows = static_cast<size_t>(RS_O_NVAR) * n_out;
  // statics resident on the device (roadsurf_prepare_statics)?
  const StaticsShard* resident = nullptr;
  if (b->statics)
  {
    const RsStatics* h = static_cast<const RsStatics*>(b->statics);
    if (h->magic != 0x52535354u || h->npoints != b->npoints)
      return fail(RS_ERR_BAD_ARGUMENT, "statics handle does not belong to a grid of this many points");
    for (const StaticsShard& s : h->shards)
      if (s.device == sh.device && s.first == sh.first && s.count == sh.count && s.chunk == chunk) resident = &s;
    if (!resident || static_cast<int>(resident->chunks.size()) != nchunks)
      return fail(RS_ERR_BAD_ARGUMENT, "statics handle was prepared for another ngpus / device set");
  }
  const bool hor = resident ? resident->chunks[0].hor != nullptr : b->horizons != nullptr;

  SoaStreams* ss = nullptr;
  {
    const int rc = soa_streams(sh.device, &ss);
    if (rc != RS_OK) return rc;
  }
  cudaStream_t* streams = ss->st;
  cudaEvent_t(*ev)[4] = ss->ev;

  // per-stream device buffers
  double *d_forcing[NSTREAM], *d_out[NSTREAM], *d_local[NSTREAM], *d_hor[NSTREAM], *d_scratch[NSTREAM],
      *d_state[NSTREAM];
  const int wend = (model.use_coupling && opt_compaction_passes() > 0) ? b->coupling_window_end : 0;
  int* d_status[NSTREAM];
  int* d_order[NSTREAM];
  unsigned long long* d_counters = nullptr;
  int *d_tf = nullptr, *d_rs = nullptr;
  void* tmp = nullptr;
  for (int s = 0; s < NSTREAM; ++s)
  {
    CU(pool_get(sh.device, 10 * s + 0, sizeof(double) * frows * chunk, &tmp));
    d_forcing[s] = static_cast<double*>(tmp);
    CU(pool_get(sh.device, 10 * s + 1, sizeof(double) * orows * chunk, &tmp));
    d_out[s] = static_cast<double*>(tmp);
    CU(pool_get(sh.device, 10 * s + 2, sizeof(double) * RS_L_NLOCAL * chunk, &tmp));
    d_local[s] = static_cast<double*>(tmp);
    d_hor[s] = nullptr;
    if (hor && !resident)
    {
      CU(pool_get(sh.device, 10 * s + 3, sizeof(double) * 360 * chunk, &tmp));
      d_hor[s] = static_cast<double*>(tmp);
    }
    d_scratch[s] = nullptr;
    if (model.use_coupling)
    {
      CU(pool_get(sh.device, 10 * s + 4, sizeof(double) * RS_SCRATCH_NPLANES(nl) * chunk, &tmp));
      d_scratch[s] = static_cast<double*>(tmp);
    }
    CU(pool_get(sh.device, 10 * s + 5, sizeof(int) * chunk, &tmp));
    d_status[s] = static_cast<int*>(tmp);
    d_state[s] = nullptr;
    if (wend > 0)
    {
      CU(pool_get(sh.device, 10 * s + 6, sizeof(double) * RS_STATE_NPLANES(nl) * chunk, &tmp));
      d_state[s] = static_cast<double*>(tmp);
    }
    d_order[s] = nullptr;
    if (hor && !resident && b->forcing_mode == 1)  // sky-view points gathered at one end of the chunk (see roadsurf_order_points)
    {
      CU(pool_get(sh.device, 10 * s + 7, sizeof(int) * chunk, &tmp));
      d_order[s] = static_cast<int*>(tmp);
    }
  }
  CU(pool_get(sh.device, 100, sizeof(int) * 6 * b->sim_len, &tmp));
  d_tf = static_cast<int*>(tmp);
  CU(pool_get(sh.device, 101, sizeof(int) * std::max(1, b->n_records), &tmp));
  d_rs = static_cast<int*>(tmp);
  CU(pool_get(sh.device, 102, sizeof(unsigned long long) * RS_CNT_N, &tmp));
  d_counters = static_cast<unsigned long long*>(tmp);
  double* d_solar = nullptr;
  CU(pool_get(sh.device, 103, sizeof(double) * 4 * b->sim_len, &tmp));
  d_solar = static_cast<double*>(tmp);
  CU(cudaMemcpyAsync(d_tf, b->time_fields, sizeof(int) * 6 * b->sim_len, cudaMemcpyHostToDevice, streams[0]));
  if (b->forcing_mode == 1)
    CU(cudaMemcpyAsync(d_rs, b->record_step, sizeof(int) * b->n_records, cudaMemcpyHostToDevice, streams[0]));
  CU(cudaMemsetAsync(d_counters, 0, sizeof(unsigned long long) * RS_CNT_N, streams[0]));
  CU(static_cast<cudaError_t>(rs_launch_solar(d_tf, b->sim_len, d_solar, streams[0])));
  ++sh.stats.kernel_launches;
  CU(cudaStreamSynchronize(streams[0]));
  sh.stats.h2d_bytes += sizeof(int) * (6 * b->sim_len + b->n_records);

  for (int c = 0; c < nchunks; ++c)
  {
    const int s = c % NSTREAM;
    cudaStream_t st = streams[s];
    const int q0 = c * chunk;
    const int npc = std::min(chunk, sh.count - q0);
    const int ld = (npc + 31) / 32 * 32;
    const size_t col0 = static_cast<size_t>(sh.first) + q0;
    if (c >= NSTREAM)
    {
      // buffers of this stream are free once its previous chunk's D2H has finished
      CU(cudaStreamSynchronize(st));
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[s][0], ev[s][1]);
      sh.stats.h2d_ms += ms;
      cudaEventElapsedTime(&ms, ev[s][1], ev[s][2]);
      sh.stats.kernel_ms += ms;
      cudaEventElapsedTime(&ms, ev[s][2], ev[s][3]);
      sh.stats.d2h_ms += ms;
    }
    const size_t wbytes = sizeof(double) * npc;
    CU(cudaEventRecord(ev[s][0], st));
    CU(cudaMemcpy2DAsync(d_forcing[s], sizeof(double) * ld, b->forcing + col0, sizeof(double) * np_all, wbytes,
                         frows, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpy2DAsync(d_local[s], sizeof(double) * ld, b->local + col0, sizeof(double) * np_all, wbytes,
                         RS_L_NLOCAL, cudaMemcpyHostToDevice, st));
    if (hor && !resident)
      CU(cudaMemcpy2DAsync(d_hor[s], sizeof(double) * ld, b->horizons + col0, sizeof(double) * np_all, wbytes, 360,
                           cudaMemcpyHostToDevice, st));
    sh.stats.h2d_bytes += wbytes * (frows + RS_L_NLOCAL + ((hor && !resident) ? 360 : 0));
    CU(cudaEventRecord(ev[s][1], st));
    RsArgs a;
    std::memset(&a, 0, sizeof a);
    a.npoints = npc;
    a.ld = ld;
    a.sim_len = b->sim_len;
    a.n_records = b->n_records;
    a.nvar = b->nvar;
    a.out_stride = b->out_stride;
    a.n_out = n_out;
    a.forcing_mode = b->forcing_mode;
    a.step_begin = 1;
    a.step_end = b->sim_len;
    a.forcing_step0 = 1;
    a.out_slot0 = 0;
    a.forcing = d_forcing[s];
    a.record_step = d_rs;
    a.tf = d_tf;
    a.local = d_local[s];
    a.horizons = resident ? resident->chunks[c].hor : d_hor[s];
    a.solar = d_solar;
    a.out = d_out[s];
    a.status = d_status[s];
    a.scratch = d_scratch[s];
    RsArgsCold ac;
    std::memset(&ac, 0, sizeof ac);
    ac.counters = d_counters;
    ac.out_start = 0;
    ac.out_nvar = RS_O_NVAR;
    ac.state = d_state[s];
    if (resident && b->forcing_mode == 1 && resident->chunks[c].order)
    {
      ac.index = resident->chunks[c].order;  // built once by roadsurf_prepare_statics
      ac.n_fixed = ld;
    }
    else if (d_order[s])
    {
      CU(static_cast<cudaError_t>(rs_launch_partition(d_local[s] + static_cast<size_t>(RS_L_SKY_VIEW) * ld, ld, npc, 2,
                                                      d_order[s], nullptr, st)));
      ++sh.stats.kernel_launches;
      ac.index = d_order[s];
      ac.n_fixed = ld;
    }
    {
      int n = 0;
      const int rc = launch_model(a, ac, model, wend, st, &sh.launch, &n);
      if (rc != RS_OK) return rc;
      sh.stats.kernel_launches += n;
    }
    sh.launch.nlayers = nl;
    sh.launch.forcing_mode = b->forcing_mode;
    CU(cudaEventRecord(ev[s][2], st));
    CU(cudaMemcpy2DAsync(b->out + col0, sizeof(double) * np_all, d_out[s], sizeof(double) * ld, wbytes, orows,
                         cudaMemcpyDeviceToHost, st));
    if (b->status)
      CU(cudaMemcpyAsync(b->status + col0, d_status[s], sizeof(int) * npc, cudaMemcpyDeviceToHost, st));
    sh.stats.d2h_bytes += wbytes * orows + (b->status ? sizeof(int) * npc : 0);
    CU(cudaEventRecord(ev[s][3], st));
  }
  for (int s = 0; s < NSTREAM && s < nchunks; ++s)
  {
    CU(cudaStreamSynchronize(streams[s]));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev[s][0], ev[s][1]);
    sh.stats.h2d_ms += ms;
    cudaEventElapsedTime(&ms, ev[s][1], ev[s][2]);
    sh.stats.kernel_ms += ms;
    cudaEventElapsedTime(&ms, ev[s][2], ev[s][3]);
    sh.stats.d2h_ms += ms;
  }
  unsigned long long cnt[RS_CNT_N];
  CU(cudaMemcpy(cnt, d_counters, sizeof cnt, cudaMemcpyDeviceToHost));
  sh.stats.executed_steps += static_cast<int64_t>(cnt[RS_CNT_EXECUTED_STEPS]);
  sh.stats.groups = nchunks;
  return RS_OK;
}

int run_sharded(int npoints, int ngpus, const std::function<int(Shard&)>& fn)
{
  const double wall0 = now_ms();
  const int ndev = roadsurf_device_count();
  if (ndev < 1) return fail(RS_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU path)");
  if (ngpus <= 0 || ngpus > ndev) ngpus = ndev;
  if (ngpus > npoints) ngpus = npoints;
  int current = 0;
  cudaGetDevice(&current);
  std::vector<Shard> shards(ngpus);
  for (int g = 0; g < ngpus; ++g)
  {
    const long long a = static_cast<long long>(npoints) * g / ngpus;
    const long long b = static_cast<long long>(npoints) * (g + 1) / ngpus;
    shards[g].first = static_cast<int>(a);
    shards[g].count = static_cast<int>(b - a);
    shards[g].device = (ngpus == 1) ? current : g;
  }
  auto work = [&](int g) {
    Shard& sh = shards[g];
    sh.rc = fn(sh);
    if (sh.rc != RS_OK) sh.err = g_err;
  };
  if (ngpus == 1)
    work(0);
  else
  {
    std::vector<std::thread> th;
    for (int g = 0; g < ngpus; ++g) th.emplace_back(work, g);
    for (auto& t : th) t.join();
    cudaSetDevice(current);
  }
  for (auto& sh : shards)
  {
    g_stats.pack_ms = std::max(g_stats.pack_ms, sh.stats.pack_ms);
    g_stats.h2d_ms = std::max(g_stats.h2d_ms, sh.stats.h2d_ms);
    g_stats.kernel_ms = std::max(g_stats.kernel_ms, sh.stats.kernel_ms);
    g_stats.d2h_ms = std::max(g_stats.d2h_ms, sh.stats.d2h_ms);
    g_stats.unpack_ms = std::max(g_stats.unpack_ms, sh.stats.unpack_ms);
    g_stats.h2d_bytes += sh.stats.h2d_bytes;
    g_stats.d2h_bytes += sh.stats.d2h_bytes;
    g_stats.executed_steps += sh.stats.executed_steps;
    g_stats.kernel_launches += sh.stats.kernel_launches;
    g_stats.groups += sh.stats.groups;
    g_stats.setup_ms = std::max(g_stats.setup_ms, sh.stats.setup_ms);
    g_launches_total += sh.stats.kernel_launches;
    if (sh.launch.grid) g_launch = sh.launch;
  }
  g_stats.wall_ms = now_ms() - wall0;
  for (auto& sh : shards)
    if (sh.rc != RS_OK) return fail(sh.rc, sh.err);
  return RS_OK;
}
}  // namespace

extern "C" {

int roadsurf_run_host_soa(const RsHostBatch* b, const InputSettings* settings, const InputParameters* params,
                          int ngpus)
{
  g_err.clear();
  std::memset(&g_stats, 0, sizeof g_stats);
  if (!b || !settings || !params) return fail(RS_ERR_BAD_ARGUMENT, "null argument");
  if (b->npoints < 0 || b->sim_len < 1 || b->out_stride < 1) return fail(RS_ERR_BAD_ARGUMENT, "bad sizes");
  if (b->sim_len != settings->SimLen) return fail(RS_ERR_BAD_ARGUMENT, "batch sim_len != settings SimLen");
  if (b->nvar != RS_F_NVAR && b->nvar != RS_F_NVAR_DEPTH) return fail(RS_ERR_BAD_ARGUMENT, "nvar must be 11 or 12");
  if (b->forcing_mode != 0 && b->forcing_mode != 1)
    return fail(RS_ERR_UNSUPPORTED, "roadsurf_run_host_soa takes forcing_mode 0 or 1 (example2's rule, mode 2, is offered by "
                                    "the device entry: roadsurf_run_device / roadsurf_expand_records)");
  if (b->forcing_mode == 0 ? b->n_records != b->sim_len : (b->n_records < 2 || !b->record_step))
    return fail(RS_ERR_BAD_ARGUMENT, "bad n_records / record_step for the forcing mode");
  if (b->npoints > 0 && (!b->forcing || !b->time_fields || !b->local || !b->out))
    return fail(RS_ERR_BAD_ARGUMENT, "null host pointer in batch");
  if (b->npoints == 0) return roadsurf_device_count() < 1 ? fail(RS_ERR_NO_DEVICE, "no CUDA device visible") : RS_OK;
  return run_sharded(b->npoints, ngpus, [&](Shard& sh) { return run_host_soa_shard(sh, b, settings, params); });
}

int roadsurf_prepare_statics(const RsHostBatch* b, int ngpus, void** handle)
{
  g_err.clear();
  if (!b || !handle || b->npoints < 1 || !b->local) return fail(RS_ERR_BAD_ARGUMENT, "bad arguments");
  const int ndev = roadsurf_device_count();
  if (ndev < 1) return fail(RS_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU path)");
  if (ngpus <= 0 || ngpus > ndev) ngpus = ndev;
  if (ngpus > b->npoints) ngpus = b->npoints;
  int current = 0;
  cudaGetDevice(&current);
  RsStatics* h = new RsStatics;
  h->npoints = b->npoints;
  h->ngpus = ngpus;
  h->has_horizons = b->horizons != nullptr;
  const size_t np_all = static_cast<size_t>(b->npoints);
  int rc = RS_OK;
  for (int g = 0; g < ngpus && rc == RS_OK; ++g)
  {
    StaticsShard s;
    s.first = static_cast<int>(static_cast<long long>(b->npoints) * g / ngpus);
    s.count = static_cast<int>(static_cast<long long>(b->npoints) * (g + 1) / ngpus) - s.first;
    s.device = (ngpus == 1) ? current : g;
    s.chunk = host_soa_chunk(s.count);
    auto body = [&]() -> int {
      CU(cudaSetDevice(s.device));
      for (int q0 = 0; q0 < s.count; q0 += s.chunk)
      {
        StaticsChunk ch;
        const int npc = std::min(s.chunk, s.count - q0);
        ch.ld = (npc + 31) / 32 * 32;
        const size_t col0 = static_cast<size_t>(s.first) + q0;
        double* d_sky = nullptr;
        CU(cudaMalloc(&d_sky, sizeof(double) * ch.ld));
        CU(cudaMemset(d_sky, 0, sizeof(double) * ch.ld));
        CU(cudaMemcpy(d_sky, b->local + static_cast<size_t>(RS_L_SKY_VIEW) * np_all + col0, sizeof(double) * npc,
                      cudaMemcpyHostToDevice));
        CU(cudaMalloc(&ch.order, sizeof(int) * ch.ld));
        CU(static_cast<cudaError_t>(rs_launch_partition(d_sky, ch.ld, npc, 2, ch.order, nullptr, nullptr)));
        ++g_launches_total;
        if (b->horizons)
        {
          CU(cudaMalloc(&ch.hor, sizeof(double) * 360 * ch.ld));
          CU(cudaMemset(ch.hor, 0, sizeof(double) * 360 * ch.ld));
          CU(cudaMemcpy2D(ch.hor, sizeof(double) * ch.ld, b->horizons + col0, sizeof(double) * np_all,
                          sizeof(double) * npc, 360, cudaMemcpyHostToDevice));
        }
        CU(cudaDeviceSynchronize());
        cudaFree(d_sky);
        s.chunks.push_back(ch);
      }
      return RS_OK;
    };
    rc = body();
    h->shards.push_back(s);
  }
  cudaSetDevice(current);
  if (rc != RS_OK)
  {
    const std::string err = g_err;
    roadsurf_release_statics(h);
    return fail(rc, err);
  }
  *handle = h;
  return RS_OK;
}

void roadsurf_release_statics(void* handle)
{
  RsStatics* h = static_cast<RsStatics*>(handle);
  if (!h || h->magic != 0x52535354u) return;
  int current = 0;
  cudaGetDevice(&current);
  for (StaticsShard& s : h->shards)
  {
    cudaSetDevice(s.device);
    for (StaticsChunk& ch : s.chunks)
    {
      if (ch.hor) cudaFree(ch.hor);
      if (ch.order) cudaFree(ch.order);
    }
  }
  cudaSetDevice(current);
  h->magic = 0;
  delete h;
}

// ---- step-granular sessions -------------------------------------------------------------------------
namespace
{
struct RsSession
{
  unsigned magic = 0x52535345u;  // "RSSE"
  int device = 0, npoints = 0, ld = 0, sim_len = 0, nl = 0, nvar = 0;
  RsModel model;
  double *d_forcing = nullptr, *d_local = nullptr, *d_hor = nullptr, *d_out = nullptr, *d_state = nullptr,
         *d_scratch = nullptr, *d_solar = nullptr;
  int *d_tf = nullptr, *d_status = nullptr;
  unsigned long long* d_counters = nullptr;
  cudaStream_t stream = nullptr;
  std::vector<OutputPointers> outs;
  int done = 0, fetched = 0, chunk = 1;
  int wstart = 0, wend = 0;  // common coupling window of the coupled points (0 = none)
};
RsSession* session_of(void* p)
{
  RsSession* s = static_cast<RsSession*>(p);
  return (s && s->magic == 0x52535345u) ? s : nullptr;
}
}  // namespace

int roadsurf_session_open(int npoints, OutputPointers* const* out, const InputPointers* const* in,
                          const InputSettings* settings, const InputParameters* params,
                          const LocalParameters* const* local, void** session)
{
  g_err.clear();
  if (npoints < 1 || !out || !in || !settings || !params || !local || !session)
    return fail(RS_ERR_BAD_ARGUMENT, "null argument");
  if (roadsurf_device_count() < 1) return fail(RS_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU path)");
  std::unique_ptr<RsSession, void (*)(void*)> guard(new RsSession, roadsurf_session_close);
  RsSession* s = guard.get();
  CU(cudaGetDevice(&s->device));
  {
    std::lock_guard<std::mutex> device_lock(g_device_mu[s->device & 63]);
    const int rc = set_model_on_current_device(settings, params, &s->model);
    if (rc != RS_OK) return rc;
  }
  const int sim_len = settings->SimLen, nl = s->model.nlayers;
  s->npoints = npoints;
  s->ld = (npoints + 31) / 32 * 32;
  s->sim_len = sim_len;
  s->nl = nl;
  const size_t ld = s->ld;
  bool any_depth = false, any_sky = false;
  for (int p = 0; p < npoints; ++p)
  {
    const InputPointers* ip = in[p];
    if (ip->inputLen < sim_len || out[p]->outputLen < sim_len)
      return fail(RS_ERR_BAD_ARGUMENT, "inputLen/outputLen shorter than SimLen");
    const size_t nb = sizeof(int) * sim_len;
    const InputPointers* r = in[0];
    if (p > 0 && (std::memcmp(r->c_year, ip->c_year, nb) || std::memcmp(r->c_month, ip->c_month, nb) ||
                  std::memcmp(r->c_day, ip->c_day, nb) || std::memcmp(r->c_hour, ip->c_hour, nb) ||
                  std::memcmp(r->c_minute, ip->c_minute, nb) || std::memcmp(r->c_second, ip->c_second, nb)))
      return fail(RS_ERR_UNSUPPORTED, "the points of a session must share one time axis");
    const LocalParameters& lp = *local[p];
    if (lp.sky_view < 1.0 && lp.sky_view > -0.01f) any_sky = true;
    for (int t = 0; t < sim_len && !any_depth; ++t) any_depth = ip->c_Depth[t] >= 0.0;
    const long long w = window_key(s->model, lp);
    if (w > 0)
    {
      const int cend = static_cast<int>(w);
      const int cstart = (static_cast<double>(cend) <= s->model.coupling_span_real) ? 1 : cend - s->model.coupling_span;
      if (s->wend && (s->wend != cend || s->wstart != cstart))
        return fail(RS_ERR_UNSUPPORTED, "the coupled points of a session must share one coupling window");
      s->wend = cend;
      s->wstart = cstart;
    }
  }
  const int nvar = s->nvar = any_depth ? RS_F_NVAR_DEPTH : RS_F_NVAR;
  // ---- pack on the host: forcing [t][var][ld], statics, horizons; pre-fill the caller's outputs
  std::vector<double> hf(static_cast<size_t>(sim_len) * nvar * ld, 0.0), hl(RS_L_NLOCAL * ld, 0.0);
  std::vector<double> hh(any_sky ? 360 * ld : 0, 0.0);
  s->outs.resize(npoints);
  for (size_t q = 0; q < ld; ++q) hl[RS_L_SKY_VIEW * ld + q] = 1.0;
  for (int p = 0; p < npoints; ++p)
  {
    const InputPointers* ip = in[p];
    const double* src[RS_F_NVAR_DEPTH] = {ip->c_tair, ip->c_tdew, ip->c_VZ,     ip->c_Rhz,      ip->c_prec, ip->c_SW,
                                          ip->c_LW,   ip->c_SW_dir, ip->c_LW_net, ip->c_TSurfObs, nullptr,    ip->c_Depth};
    for (int t = 0; t < sim_len; ++t)
      for (int v = 0; v < nvar; ++v)
        hf[(static_cast<size_t>(t) * nvar + v) * ld + p] =
            (v == RS_F_PHASE) ? static_cast<double>(ip->c_PrecPhase[t]) : src[v][t];
    const LocalParameters& lp = *local[p];
    const double row[RS_L_NLOCAL] = {lp.tair_relax, lp.VZ_relax, lp.RH_relax, lp.couplingTsurf, lp.lat, lp.lon,
                                     lp.sky_view, static_cast<double>(lp.couplingIndexI),
                                     static_cast<double>(lp.InitLenI), 1.0};
    for (int k = 0; k < RS_L_NLOCAL; ++k) hl[k * ld + p] = row[k];
    if (any_sky)
      for (int k = 0; k < 360; ++k) hh[static_cast<size_t>(k) * ld + p] = ip->c_local_horizons ? ip->c_local_horizons[k] : 0.0;
    s->outs[p] = *out[p];
    double* o[RS_O_NVAR] = {out[p]->c_TsurfOut, out[p]->c_SnowOut,    out[p]->c_WaterOut,
                            out[p]->c_IceOut,   out[p]->c_DepositOut, out[p]->c_Ice2Out};
    for (int v = 0; v < RS_O_NVAR; ++v)
      for (int t = 0; t < sim_len; ++t) o[v][t] = -9999.0;
  }
  CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  CU(cudaMalloc(&s->d_forcing, sizeof(double) * hf.size()));
  CU(cudaMalloc(&s->d_local, sizeof(double) * hl.size()));
  if (any_sky) CU(cudaMalloc(&s->d_hor, sizeof(double) * hh.size()));
  CU(cudaMalloc(&s->d_out, sizeof(double) * RS_O_NVAR * sim_len * ld));
  CU(cudaMalloc(&s->d_state, sizeof(double) * RS_STATE_NPLANES(nl) * ld));
  CU(cudaMalloc(&s->d_scratch, sizeof(double) * RS_SCRATCH_NPLANES(nl) * ld));
  CU(cudaMalloc(&s->d_solar, sizeof(double) * 4 * sim_len));
  CU(cudaMalloc(&s->d_tf, sizeof(int) * 6 * sim_len));
  CU(cudaMalloc(&s->d_status, sizeof(int) * ld));
  CU(cudaMalloc(&s->d_counters, sizeof(unsigned long long) * RS_CNT_N));
  CU(cudaMemcpyAsync(s->d_forcing, hf.data(), sizeof(double) * hf.size(), cudaMemcpyHostToDevice, s->stream));
  CU(cudaMemcpyAsync(s->d_local, hl.data(), sizeof(double) * hl.size(), cudaMemcpyHostToDevice, s->stream));
  if (any_sky) CU(cudaMemcpyAsync(s->d_hor, hh.data(), sizeof(double) * hh.size(), cudaMemcpyHostToDevice, s->stream));
  const int* tf[6] = {in[0]->c_year, in[0]->c_month, in[0]->c_day, in[0]->c_hour, in[0]->c_minute, in[0]->c_second};
  for (int k = 0; k < 6; ++k)
    CU(cudaMemcpyAsync(s->d_tf + static_cast<size_t>(k) * sim_len, tf[k], sizeof(int) * sim_len, cudaMemcpyHostToDevice,
                       s->stream));
  CU(cudaMemsetAsync(s->d_status, 0, sizeof(int) * ld, s->stream));
  CU(cudaMemsetAsync(s->d_counters, 0, sizeof(unsigned long long) * RS_CNT_N, s->stream));
  CU(static_cast<cudaError_t>(rs_launch_fill(s->d_out, static_cast<long long>(RS_O_NVAR) * sim_len * ld, -9999.0, s->stream)));
  CU(static_cast<cudaError_t>(rs_launch_solar(s->d_tf, sim_len, s->d_solar, s->stream)));
  g_launches_total += 2;
  CU(cudaStreamSynchronize(s->stream));  // the host staging vectors die with this frame
  *session = guard.release();
  return RS_OK;
}

int roadsurf_session_set_chunk(void* session, int steps)
{
  RsSession* s = session_of(session);
  if (!s) return fail(RS_ERR_BAD_ARGUMENT, "not a session");
  s->chunk = steps > 1 ? steps : 1;
  return RS_OK;
}

int roadsurf_session_done(void* session)
{
  RsSession* s = session_of(session);
  return s ? s->done : RS_ERR_BAD_ARGUMENT;
}

int roadsurf_step(void* session, int i)
{
  RsSession* s = session_of(session);
  if (!s) return fail(RS_ERR_BAD_ARGUMENT, "not a session");
  if (i < 1 || i > s->sim_len) return fail(RS_ERR_BAD_ARGUMENT, "step outside [1, SimLen]");
  if (i <= s->done) return RS_OK;
  int current = 0;
  CU(cudaGetDevice(&current));
  if (current != s->device) CU(cudaSetDevice(s->device));
  const int begin = s->done + 1;
  int target = std::min(s->sim_len, std::max(i, s->done + s->chunk));
  // a coupling window [start, end + 1] is indivisible: its rewinds happen inside the kernel
  if (s->wend > 0 && begin <= s->wend + 1 && target >= s->wstart) target = std::max(target, std::min(s->wend + 1, s->sim_len));
  {
    std::lock_guard<std::mutex> device_lock(g_device_mu[s->device & 63]);
    RsModel m;
    // (another caller may have replaced the device's model since the session was opened)
    {
      std::lock_guard<std::mutex> lk(g_model_mu);
      DeviceModel& dm = g_models[s->device];
      if (!dm.valid || std::memcmp(&dm.m, &s->model, sizeof m) != 0)
      {
        for (auto& kv : dm.last_launch) CU(cudaEventSynchronize(kv.second));
        CU(static_cast<cudaError_t>(rs_upload_model(&s->model)));
        dm.m = s->model;
        dm.valid = true;
      }
    }
    RsArgs a;
    std::memset(&a, 0, sizeof a);
    a.npoints = s->npoints;
    a.ld = s->ld;
    a.sim_len = s->sim_len;
    a.n_records = s->sim_len;
    a.nvar = s->nvar;
    a.out_stride = 1;
    a.n_out = s->sim_len;
    a.forcing_mode = 0;
    a.step_begin = begin;
    a.step_end = target;
    a.forcing_step0 = 1;
    a.forcing = s->d_forcing;
    a.tf = s->d_tf;
    a.local = s->d_local;
    a.horizons = s->d_hor;
    a.solar = s->d_solar;
    a.out = s->d_out;
    a.status = s->d_status;
    a.scratch = s->d_scratch;
    RsArgsCold ac;
    std::memset(&ac, 0, sizeof ac);
    ac.state = s->d_state;
    ac.counters = s->d_counters;
    ac.out_nvar = RS_O_NVAR;
    RsLaunchInfo li;
    std::memset(&li, 0, sizeof li);
    int n = 0;
    const int rc = launch_model(a, ac, s->model, 0, s->stream, &li, &n);
    if (rc != RS_OK) return rc;
    g_launches_total += n;
    li.nlayers = s->nl;
    li.launches_total = g_launches_total;
    g_launch = li;
    CU(cudaStreamSynchronize(s->stream));
  }
  s->done = target;
  if (current != s->device) cudaSetDevice(current);
  return RS_OK;
}

int roadsurf_session_fetch(void* session, int i, int* status)
{
  RsSession* s = session_of(session);
  if (!s) return fail(RS_ERR_BAD_ARGUMENT, "not a session");
  if (i > s->done) return fail(RS_ERR_BAD_ARGUMENT, "step not executed yet: call roadsurf_step first");
  int current = 0;
  CU(cudaGetDevice(&current));
  if (current != s->device) CU(cudaSetDevice(s->device));
  // deliver every executed step not yet delivered; the steps of a coupling window were overwritten by the
  // kernel's re-runs before they became visible here
  const int t0 = s->fetched, n = s->done - s->fetched;
  if (n > 0)
  {
    std::vector<double> h(static_cast<size_t>(RS_O_NVAR) * n * s->ld);
    for (int v = 0; v < RS_O_NVAR; ++v)
      CU(cudaMemcpy(h.data() + static_cast<size_t>(v) * n * s->ld,
                    s->d_out + (static_cast<size_t>(v) * s->sim_len + t0) * s->ld, sizeof(double) * n * s->ld,
                    cudaMemcpyDeviceToHost));
    for (int p = 0; p < s->npoints; ++p)
    {
      OutputPointers& op = s->outs[p];
      double* dst[RS_O_NVAR] = {op.c_TsurfOut, op.c_SnowOut, op.c_WaterOut, op.c_IceOut, op.c_DepositOut, op.c_Ice2Out};
      for (int v = 0; v < RS_O_NVAR; ++v)
        for (int t = 0; t < n; ++t) dst[v][t0 + t] = h[(static_cast<size_t>(v) * n + t) * s->ld + p];
    }
    s->fetched = s->done;
  }
  if (status) CU(cudaMemcpy(status, s->d_status, sizeof(int) * s->npoints, cudaMemcpyDeviceToHost));
  if (current != s->device) cudaSetDevice(current);
  return RS_OK;
}

void roadsurf_session_close(void* session)
{
  RsSession* s = session_of(session);
  if (!s) return;
  int current = 0;
  cudaGetDevice(&current);
  cudaSetDevice(s->device);
  for (void* p : {static_cast<void*>(s->d_forcing), static_cast<void*>(s->d_local), static_cast<void*>(s->d_hor),
                  static_cast<void*>(s->d_out), static_cast<void*>(s->d_state), static_cast<void*>(s->d_scratch),
                  static_cast<void*>(s->d_solar), static_cast<void*>(s->d_tf), static_cast<void*>(s->d_status),
                  static_cast<void*>(s->d_counters)})
    if (p) cudaFree(p);
  if (s->stream) cudaStreamDestroy(s->stream);
  cudaSetDevice(current);
  s->magic = 0;
  delete s;
}

const char* roadsurf_last_error(void) { return g_err.c_str(); }

const char* roadsurf_version(void) { return "roadsurf_b200 0.1 (RoadSurf 1.6.1 per-point loop, sm_100a)"; }

int roadsurf_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
  {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int roadsurf_set_model(const InputSettings* settings, const InputParameters* params)
{
  if (!settings || !params) return fail(RS_ERR_BAD_ARGUMENT, "null settings/params");
  if (roadsurf_device_count() < 1) return fail(RS_ERR_NO_DEVICE, "no CUDA device visible");
  return set_model_on_current_device(settings, params, nullptr);
}

int roadsurf_run_device(const RsDeviceBatch* b, void* stream)
{
  if (!b) return fail(RS_ERR_BAD_ARGUMENT, "null batch");
  int dev = 0;
  CU(cudaGetDevice(&dev));
  RsModel m;
  {
    std::lock_guard<std::mutex> lk(g_model_mu);
    auto it = g_models.find(dev);
    if (it == g_models.end() || !it->second.valid)
      return fail(RS_ERR_BAD_ARGUMENT, "roadsurf_set_model was not called on this device");
    m = it->second.m;
  }
  if (b->ld < 32 || b->ld % 32 != 0 || b->npoints < 0 || b->npoints > b->ld)
    return fail(RS_ERR_BAD_ARGUMENT, "ld must be a positive multiple of 32 and >= npoints");
  const int step_begin = b->step_begin > 0 ? b->step_begin : 1;
  const int step_end = b->step_end > 0 ? b->step_end : b->sim_len;
  const int forcing_step0 = b->forcing_step0 > 0 ? b->forcing_step0 : 1;
  if (b->sim_len < 1 || b->out_stride < 1 || b->n_out < 1)
    return fail(RS_ERR_BAD_ARGUMENT, "bad sim_len / out_stride / n_out");
  if (step_begin > step_end || step_end > b->sim_len || forcing_step0 > step_begin || b->out_slot0 < 0)
    return fail(RS_ERR_BAD_ARGUMENT, "bad step_begin / step_end / forcing_step0 / out_slot0");
  if (b->out_start < 0 || b->out_start >= b->sim_len) return fail(RS_ERR_BAD_ARGUMENT, "out_start outside [0, sim_len)");
  if (b->out_nvar != 0 && b->out_nvar != RS_O_NVAR && b->out_nvar != RS_O_NVAR_EXT)
    return fail(RS_ERR_BAD_ARGUMENT, "out_nvar must be 0, RS_O_NVAR or RS_O_NVAR_EXT");
  if (step_end - 1 >= b->out_start)
  {
    const int k0 = step_begin - 1 - b->out_start;
    const int first_slot = k0 > 0 ? (k0 + b->out_stride - 1) / b->out_stride : 0;
    const int last_slot = (step_end - 1 - b->out_start) / b->out_stride;
    if (last_slot >= first_slot && (first_slot < b->out_slot0 || last_slot - b->out_slot0 >= b->n_out))
      return fail(RS_ERR_BAD_ARGUMENT, "out tensor does not cover the output slots of [step_begin, step_end]");
  }
  if (step_begin > 1 && !b->state) return fail(RS_ERR_BAD_ARGUMENT, "step_begin > 1 needs batch->state from the previous chunk");
  if (b->nvar != RS_F_NVAR && b->nvar != RS_F_NVAR_DEPTH) return fail(RS_ERR_BAD_ARGUMENT, "nvar must be 11 or 12");
  if (!b->forcing || !b->time_fields || !b->local || !b->out || !b->status || !b->solar)
    return fail(RS_ERR_BAD_ARGUMENT, "null device pointer in batch");
  if (b->forcing_mode == 0)
  {
    if (b->n_records < step_end - forcing_step0 + 1)
      return fail(RS_ERR_BAD_ARGUMENT, "forcing_mode 0 needs n_records >= step_end - forcing_step0 + 1");
  }
  else if (b->forcing_mode == 1 || b->forcing_mode == 2)
  {
    if (b->n_records < 2 || !b->record_step) return fail(RS_ERR_BAD_ARGUMENT, "coarse forcing needs >= 2 records");
    if (b->forcing_mode == 2)
    {
      if (!b->expand_workspace || b->expand_steps < 1)
        return fail(RS_ERR_BAD_ARGUMENT, "forcing_mode 2 needs expand_workspace and expand_steps >= 1");
      if (b->expand_steps < step_end - step_begin + 1 && !b->state)
        return fail(RS_ERR_BAD_ARGUMENT, "forcing_mode 2 in several chunks needs batch->state");
    }
  }
  else
    return fail(RS_ERR_BAD_ARGUMENT, "forcing_mode must be 0, 1 or 2");
  if (m.use_coupling && !b->scratch) return fail(RS_ERR_BAD_ARGUMENT, "coupling needs batch->scratch");
  RsArgs a;
  std::memset(&a, 0, sizeof a);
  a.npoints = b->npoints;
  a.ld = b->ld;
  a.sim_len = b->sim_len;
  a.n_records = b->n_records;
  a.nvar = b->nvar;
  a.out_stride = b->out_stride;
  a.n_out = b->n_out;
  a.forcing_mode = b->forcing_mode;
  a.step_begin = step_begin;
  a.step_end = step_end;
  a.forcing_step0 = forcing_step0;
  a.out_slot0 = b->out_slot0;
  a.forcing = b->forcing;
  a.record_step = b->record_step;
  a.tf = b->time_fields;
  a.local = b->local;
  a.horizons = b->horizons;
  a.solar = b->solar;
  a.out = b->out;
  a.status = b->status;
  a.scratch = b->scratch;
  RsArgsCold ac;
  std::memset(&ac, 0, sizeof ac);
  ac.state = b->state;
  ac.counters = b->counters;
  ac.out_start = b->out_start;
  ac.out_nvar = b->out_nvar == RS_O_NVAR_EXT ? RS_O_NVAR_EXT : RS_O_NVAR;
  if (b->order && !(opt_staging() && b->forcing_mode == 0))  // (the staged ring needs consecutive points per warp)
  {
    ac.index = b->order;  // thread t runs point order[t]
    ac.n_fixed = b->ld;
  }
  RsLaunchInfo li;
  std::memset(&li, 0, sizeof li);
  CU(static_cast<cudaError_t>(rs_launch_solar(b->time_fields, b->sim_len, b->solar, stream)));
  ++g_launches_total;

  if (b->forcing_mode == 2)
  {
    // expansion pass + full-resolution step kernel, chunk by chunk through the resumable path
    for (int cb = step_begin; cb <= step_end; cb += b->expand_steps)
    {
      const int ce = std::min(step_end, cb + b->expand_steps - 1);
      CU(static_cast<cudaError_t>(rs_launch_expand(b->forcing, b->record_step, b->n_records, b->nvar, b->ld, b->npoints, 2,
                                                    m.DT, cb, ce, b->expand_workspace, stream)));
      RsArgs ak = a;
      ak.forcing_mode = 0;
      ak.forcing = b->expand_workspace;
      ak.n_records = ce - cb + 1;
      ak.forcing_step0 = cb;
      ak.step_begin = cb;
      ak.step_end = ce;
      int n = 0;
      const int rc = launch_model(ak, ac, m, 0, stream, &li, &n);
      if (rc != RS_OK) return rc;
      g_launches_total += n + 1;
    }
    --g_launches_total;  // (the last one is counted below)
  }
  else
  {
    int n = 0;
    const int rc = launch_model(a, ac, m, b->coupling_window_end, stream, &li, &n);
    if (rc != RS_OK) return rc;
    g_launches_total += n - 1;  // (the last one is counted below)
  }
  {
    const int rc = note_async_launch(dev, static_cast<cudaStream_t>(stream));
    if (rc != RS_OK) return rc;
  }
  li.nlayers = m.nlayers;
  li.forcing_mode = b->forcing_mode;
  li.launches_total = ++g_launches_total;
  g_launch = li;
  return RS_OK;
}

int roadsurf_expand_records(const RsDeviceBatch* b, int rule, int step_begin, int step_end, double* dst, void* stream)
{
  if (!b || !dst || !b->forcing || !b->record_step || b->n_records < 2 || (rule != 1 && rule != 2) || step_begin < 1 ||
      step_end < step_begin || b->ld < 32 || b->ld % 32 != 0)
    return fail(RS_ERR_BAD_ARGUMENT, "bad arguments");
  int dev = 0;
  CU(cudaGetDevice(&dev));
  double DT;
  {
    std::lock_guard<std::mutex> lk(g_model_mu);
    auto it = g_models.find(dev);
    if (it == g_models.end() || !it->second.valid)
      return fail(RS_ERR_BAD_ARGUMENT, "roadsurf_set_model was not called on this device");
    DT = it->second.m.DT;
  }
  CU(static_cast<cudaError_t>(rs_launch_expand(b->forcing, b->record_step, b->n_records, b->nvar, b->ld, b->npoints, rule, DT,
                                                step_begin, step_end, dst, stream)));
  ++g_launches_total;
  return RS_OK;
}

int roadsurf_sun_position(const int* time_fields, int n_steps, const double* lat, const double* lon, int npoints,
                          double* elevation, double* azimuth, void* stream)
{
  if (!time_fields || !lat || !lon || !elevation || !azimuth || n_steps < 1 || npoints < 1 || n_steps > 65535)
    return fail(RS_ERR_BAD_ARGUMENT, "bad arguments (n_steps <= 65535 per call)");
  CU(static_cast<cudaError_t>(rs_launch_sun_position(time_fields, n_steps, lat, lon, npoints, elevation, azimuth, stream)));
  ++g_launches_total;
  return RS_OK;
}

int roadsurf_order_points(const double* local, int ld, int npoints, int* order, void* stream)
{
  if (!local || !order || ld < 32 || ld % 32 != 0 || npoints < 0 || npoints > ld)
    return fail(RS_ERR_BAD_ARGUMENT, "bad arguments");
  CU(static_cast<cudaError_t>(
      rs_launch_partition(local + static_cast<size_t>(RS_L_SKY_VIEW) * ld, ld, npoints, 2, order, nullptr, stream)));
  ++g_launches_total;
  return RS_OK;
}

int roadsurf_transpose_to_soa(const double* src, int64_t src_ld, int npoints, int n, double* dst, int ld,
                              void* stream)
{
  CU(static_cast<cudaError_t>(rs_launch_transpose_to_soa(src, src_ld, npoints, n, dst, ld, stream)));
  ++g_launches_total;
  return RS_OK;
}

int roadsurf_transpose_from_soa(const double* src, int ld, int npoints, int n, double* dst, int64_t dst_ld,
                                void* stream)
{
  CU(static_cast<cudaError_t>(rs_launch_transpose_from_soa(src, ld, npoints, n, dst, dst_ld, stream)));
  ++g_launches_total;
  return RS_OK;
}

int roadsurf_fill(double* dst, int64_t n, double value, void* stream)
{
  CU(static_cast<cudaError_t>(rs_launch_fill(dst, n, value, stream)));
  ++g_launches_total;
  return RS_OK;
}

double roadsurf_measure_fp64_tflops(int iterations)
{
  if (roadsurf_device_count() < 1)
  {
    fail(RS_ERR_NO_DEVICE, "no CUDA device visible");
    return -1.0;
  }
  return rs_measure_fp64(iterations > 0 ? iterations : 20000);
}

void roadsurf_release_workspace(void)
{
  // frees the pooled device and pinned buffers of every device (they are re-created on demand)
  std::lock_guard<std::mutex> lk(g_pool_mu);
  int current = 0;
  cudaGetDevice(&current);
  for (auto& kv : g_pool)
    if (kv.second.p)
    {
      cudaSetDevice(kv.first.first);
      cudaFree(kv.second.p);
    }
  g_pool.clear();
  for (auto& kv : g_pinned_pool)
    if (kv.second.p) cudaFreeHost(kv.second.p);
  g_pinned_pool.clear();
  cudaSetDevice(current);
}

int roadsurf_set_option(const char* name, int value)
{
  if (name && std::strcmp(name, "forcing_staging") == 0)
  {
    g_opt_staging = value ? 1 : 0;
    return RS_OK;
  }
  if (name && std::strcmp(name, "coupling_compaction_passes") == 0)
  {
    g_opt_compaction = value > 0 ? (value > 30 ? 30 : value) : 0;
    return RS_OK;
  }
  if (name && std::strcmp(name, "spread_small") == 0)
  {
    rs_set_spread_small(value);
    return RS_OK;
  }
  if (name && std::strcmp(name, "latency_body") == 0)
  {
    rs_set_latency_body(value);
    return RS_OK;
  }
  if (name && std::strcmp(name, "write_back_inputs") == 0)
  {
    g_opt_write_back = value ? 1 : 0;
    return RS_OK;
  }
  if (name && std::strcmp(name, "max_points_per_device_batch") == 0)
  {
    g_opt_max_slots = value > 0 ? value : 0;
    return RS_OK;
  }
  return fail(RS_ERR_BAD_ARGUMENT, "unknown option");
}

long long roadsurf_selftest_libm(long long n, unsigned long long seed, long long* mismatches)
{
  unsigned long long s = seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
  auto next = [&]() {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    return s;
  };
  auto same = [](double a, double b) { return !std::memcmp(&a, &b, sizeof a); };
  long long bad_e = 0, bad_l = 0;
  for (long long k = 0; k < n; ++k)
  {
    const double u = static_cast<double>(next() >> 11) * 0x1p-53;
    double x, y;
    switch (k & 3)
    {
      case 0: x = (u - 0.5) * 20; y = 1.0 + u * 0.2; break;               // Magnus exponents; weakly unstable
      case 1: x = (u - 0.5) * 1000; y = std::ldexp(1.0 + u, static_cast<int>(next() % 600) - 300); break;
      case 2: x = (u - 0.5) * 1e-3; y = 0.9 + u * 0.2; break;             // around exp(0), log(1)
      default: x = -u * 6; y = 1.0 + u * 40; break;                       // decays; strongly unstable
    }
    bool ok;
    const double e = rslibm::exp_fast(x, ok);
    if (ok && !same(e, std::exp(x))) ++bad_e;
    const double l = rslibm::log_fast(y, ok);
    if (ok && !same(l, std::log(y))) ++bad_l;
  }
  if (mismatches)
  {
    mismatches[0] = bad_e;
    mismatches[1] = bad_l;
  }
  return n;
}

long long roadsurf_selftest_arith(long long n, unsigned long long seed, long long* mismatches)
{
  if (roadsurf_device_count() < 1)
  {
    fail(RS_ERR_NO_DEVICE, "no CUDA device visible");
    return -1;
  }
  return rs_selftest_arith(n, seed, mismatches);
}

void roadsurf_last_launch(RsLaunchInfo* info)
{
  if (info)
  {
    *info = g_launch;
    info->launches_total = g_launches_total;
  }
}

int roadsurf_read_input_derive(int npoints, const InputPointers* const* in, const InputSettings* settings,
                               int forecast_step, const int* latest_obs_index, LocalParameters* const* local,
                               int* ok)
{
  if (npoints < 0 || !settings || (npoints > 0 && (!in || !local))) return fail(RS_ERR_BAD_ARGUMENT, "null argument");
  const int sim_len = settings->SimLen;
  const int span = static_cast<int>(settings->coupling_minutes * 60 / settings->DTSecs);
  auto missing = [](double x) { return std::isnan(x) || x < -9000; };  // roadrunner.cpp:40-43
  parallel_for(npoints, host_threads(), [&](int p) {
    const InputPointers* ip = in[p];
    LocalParameters* lp = local[p];
    const int n = std::min(sim_len, ip->inputLen);
    bool good = true;
    for (int t = 0; t < n && good; ++t)
      good = !(missing(ip->c_tair[t]) || missing(ip->c_Rhz[t]) || missing(ip->c_prec[t]) || missing(ip->c_SW[t]) ||
               missing(ip->c_LW[t]) || missing(ip->c_VZ[t]));
    if (ok) ok[p] = good ? 1 : 0;
    if (!good) return;
    lp->InitLenI = 1 + forecast_step;
    if (settings->use_relaxation == 1)
    {
      lp->tair_relax = lp->VZ_relax = lp->RH_relax = -9999.9;
      const int last = latest_obs_index ? latest_obs_index[p] : -9999;
      if (last > -1 && last < n)
      {
        lp->InitLenI = last;
        lp->tair_relax = ip->c_tair[last];
        lp->VZ_relax = ip->c_VZ[last];
        lp->RH_relax = ip->c_Rhz[last];
      }
    }
    if (settings->use_coupling == 1)
    {
      lp->couplingTsurf = -9999.9;
      lp->couplingIndexI = -9999;
      double* obs = ip->c_TSurfObs;
      int i = n - 1;
      while (i >= 0 && (missing(obs[i]) || obs[i] < -100)) --i;
      if (i >= span)
      {
        lp->couplingTsurf = obs[i];
        lp->couplingIndexI = i;
        for (int j = i; j > i - span; --j) obs[j] = -9999.9;
      }
    }
  });
  return RS_OK;
}

int roadsurf_read_input_derive_records(const RsHostBatch* b, const InputSettings* settings, int forecast_step,
                                       const int* latest_obs_index, double* local, int* window_end)
{
  if (!b || !settings || !local || !b->forcing || !b->record_step || b->forcing_mode != 1 || b->n_records < 2 ||
      b->npoints < 0 || (b->nvar != RS_F_NVAR && b->nvar != RS_F_NVAR_DEPTH))
    return fail(RS_ERR_BAD_ARGUMENT, "needs a coarse-record host batch");
  const int sim_len = settings->SimLen, nrec = b->n_records;
  const size_t np = static_cast<size_t>(b->npoints), nvar = static_cast<size_t>(b->nvar);
  const double DT = settings->DTSecs;
  const int span = static_cast<int>(settings->coupling_minutes * 60 / DT);
  const int* rs = b->record_step;
  auto rec = [&](int k, int v, size_t p) { return b->forcing[(static_cast<size_t>(k) * nvar + v) * np + p]; };
  // the value JsonSource's interpolation gives at vector index t (the kernel's fetch_coarse rule)
  auto value_at = [&](int v, size_t p, int t) {
    int k = 0;
    while (k + 2 < nrec && rs[k + 1] <= t) ++k;
    if (t >= rs[nrec - 1]) return -9999.9;  // at / after the last record: never filled (JsonSource.cpp:85)
    const double va = rec(k, v, p), vb = rec(k + 1, v, p);
    if (t == rs[k]) return (va > -100.0) ? va : -9999.9;
    if (!(va > -100.0 && vb > -100.0)) return -9999.9;
    const double spn = static_cast<double>(rs[k + 1] - rs[k]) * DT, dt_a = static_cast<double>(t - rs[k]) * DT;
    return va + (dt_a * (vb - va)) / spn;
  };
  // vector indices [lo, hi] served by bracket k (the first one reaches back to index 0); false if empty.
  // Indices at or after the last record are served by nobody: the reference's interpolation stops at its
  // last raw record (JsonSource.cpp:85), those steps stay missing and read_input rejects the point.
  auto bracket_range = [&](int k, int& lo, int& hi) {
    lo = (k == 0) ? 0 : std::max(rs[k], 0);
    hi = std::min(rs[k + 1] - 1, sim_len - 1);
    return lo <= hi;
  };
  const bool records_cover_run = rs[nrec - 1] > sim_len - 1;
  parallel_for(b->npoints, host_threads(), [&](int pi) {
    const size_t p = static_cast<size_t>(pi);
    double* L = local + p;
    // ---- screening.  Bracket k serves the vector indices [lo, hi]; an index above rs[k] needs both
    // records valid, the index rs[k] itself only record k
    bool good = records_cover_run;
    const int req[6] = {RS_F_TAIR, RS_F_RHZ, RS_F_PREC, RS_F_SW, RS_F_LW, RS_F_VZ};
    for (int k = 0; k + 1 < nrec && good; ++k)
    {
      int lo, hi;
      if (!bracket_range(k, lo, hi)) continue;
      const bool at_a = lo <= rs[k] && rs[k] <= hi, above = hi >= std::max(rs[k] + 1, lo);
      for (int r = 0; r < 6 && good; ++r)
      {
        const bool va = rec(k, req[r], p) > -100.0, vb = rec(k + 1, req[r], p) > -100.0;
        if ((at_a && !va) || (above && !(va && vb))) good = false;
      }
    }
    L[RS_L_ACTIVE * np] = good ? 1.0 : 0.0;
    if (!good) return;
    L[RS_L_INIT_LEN * np] = 1 + forecast_step;
    if (settings->use_relaxation == 1)
    {
      L[RS_L_TAIR_RELAX * np] = L[RS_L_VZ_RELAX * np] = L[RS_L_RH_RELAX * np] = -9999.9;
      const int last = latest_obs_index ? latest_obs_index[pi] : -9999;
      if (last > -1 && last < sim_len)
      {
        L[RS_L_INIT_LEN * np] = last;
        L[RS_L_TAIR_RELAX * np] = value_at(RS_F_TAIR, p, last);
        L[RS_L_VZ_RELAX * np] = value_at(RS_F_VZ, p, last);
        L[RS_L_RH_RELAX * np] = value_at(RS_F_RHZ, p, last);
      }
    }
    if (settings->use_coupling == 1)
    {
      L[RS_L_COUPLING_TSURF * np] = -9999.9;
      L[RS_L_COUPLING_INDEX * np] = -9999;
      // latest vector index with a valid interpolated observation
      int i = -1;
      for (int k = nrec - 2; k >= 0 && i < 0; --k)
      {
        int lo, hi;
        if (!bracket_range(k, lo, hi)) continue;
        const bool va = rec(k, RS_F_TSURFOBS, p) > -100.0, vb = rec(k + 1, RS_F_TSURFOBS, p) > -100.0;  // NaN fails
        if (va && vb && hi >= std::max(rs[k] + 1, lo))
          i = hi;
        else if (va && lo <= rs[k] && rs[k] <= hi)
          i = rs[k];
      }
      const double obs = (i >= 0) ? value_at(RS_F_TSURFOBS, p, i) : -9999.9;
      if (i >= span)
      {
        L[RS_L_COUPLING_TSURF * np] = obs;
        L[RS_L_COUPLING_INDEX * np] = i;
      }
    }
  });
  if (window_end)
  {
    int w = 0;
    bool differs = false;
    if (settings->use_coupling == 1)
      for (size_t p = 0; p < np; ++p)
      {
        const double ts = local[RS_L_COUPLING_TSURF * np + p], ix = local[RS_L_COUPLING_INDEX * np + p];
        if (local[RS_L_ACTIVE * np + p] == 0.0 || ts < -100 || ix < 1) continue;
        if (w == 0)
          w = static_cast<int>(ix);
        else if (w != static_cast<int>(ix))
          differs = true;
      }
    *window_end = differs ? 0 : w;
  }
  return RS_OK;
}

void roadsurf_last_batch_stats(RsBatchStats* stats)
{
  if (stats) *stats = g_stats;
}

int roadsurf_run_batch(int npoints, OutputPointers* const* out, const InputPointers* const* in,
                       const InputSettings* settings, const InputParameters* params,
                       const LocalParameters* const* local, int ngpus, int* status)
{
  g_err.clear();
  std::memset(&g_stats, 0, sizeof g_stats);
  if (npoints < 0 || !settings || !params || (npoints > 0 && (!out || !in || !local)))
    return fail(RS_ERR_BAD_ARGUMENT, "null argument");
  if (roadsurf_device_count() < 1)
    return fail(RS_ERR_NO_DEVICE, "no CUDA device visible (this library has no CPU path)");
  if (npoints == 0) return RS_OK;
  return run_sharded(npoints, ngpus,
                     [&](Shard& sh) { return run_shard(sh, out, in, settings, params, local, status); });
}

}  // extern "C" (entry points above)

// The reference's mains call runsimulation from a pool of host threads, one point per call
// (examples/example1/src/roadrunner.cpp:454-496).  One point per launch would leave the GPU at its
// single-warp latency (~8 us per model step), so concurrent callers are combined: the first caller
// becomes the leader and runs everything that is queued as ONE batch, the others sleep until their
// point is done; while a batch is on the device the next one accumulates.  A lone caller pays only a
// mutex.  Calls with different settings / parameters go into different batches.
namespace
{
struct PendingCall
{
  OutputPointers* out;
  const InputPointers* in;
  const InputSettings* settings;
  const InputParameters* params;
  const LocalParameters* local;
  int rc = RS_OK;
  bool done = false;
  std::string err;
};
std::mutex g_call_mu;
std::condition_variable g_call_cv;
std::vector<PendingCall*> g_call_queue;
bool g_call_leader = false;
size_t g_call_last_n = 0;     // calls in the last batch (0 = no batch yet); guarded by g_call_mu
double g_call_last_ms = 0.0;  // wall time of the last batch
std::atomic<long long> g_calls_total{0}, g_call_batches_total{0};

void mark_not_computed(OutputPointers* outPointers, const InputSettings* inSettings)
{
  // The reference signature has no error channel: leave the outputs at -9999.0, which is how the
  // reference marks "not computed" (src/Initialization.f90:397-412).
  if (!outPointers || !inSettings) return;
  double* o[RS_O_NVAR] = {outPointers->c_TsurfOut, outPointers->c_SnowOut,    outPointers->c_WaterOut,
                          outPointers->c_IceOut,   outPointers->c_DepositOut, outPointers->c_Ice2Out};
  for (int v = 0; v < RS_O_NVAR; ++v)
    if (o[v])
      for (int t = 0; t < inSettings->SimLen && t < outPointers->outputLen; ++t) o[v][t] = -9999.0;
}
}  // namespace

extern "C" void roadsurf_runsimulation_counters(long long* calls, long long* batches)
{
  if (calls) *calls = g_calls_total.load();
  if (batches) *batches = g_call_batches_total.load();
}

extern "C" void runsimulation(OutputPointers* outPointers, const InputPointers* inPointers,
                              const InputSettings* inSettings, const InputParameters* inputParam,
                              const LocalParameters* localParam)
{
  if (!outPointers || !inPointers || !inSettings || !inputParam || !localParam)
  {
    std::fprintf(stderr, "roadsurf_b200 runsimulation failed: null argument\n");
    mark_not_computed(outPointers, inSettings);
    return;
  }
  // arguments that would fail a whole batch are rejected here, for this caller only
  if (inPointers->inputLen < inSettings->SimLen || outPointers->outputLen < inSettings->SimLen)
  {
    std::fprintf(stderr, "roadsurf_b200 runsimulation failed: inputLen/outputLen shorter than SimLen\n");
    mark_not_computed(outPointers, inSettings);
    return;
  }
  PendingCall me;
  me.out = outPointers;
  me.in = inPointers;
  me.settings = inSettings;
  me.params = inputParam;
  me.local = localParam;
  std::unique_lock<std::mutex> lk(g_call_mu);
  g_call_queue.push_back(&me);
  while (!me.done)
  {
    if (g_call_leader)
    {
      // a leader is at work: it will run this call, or hand the leadership over when its own is done
      g_call_cv.wait(lk, [&] { return me.done || !g_call_leader; });
      continue;
    }
    // become the leader for ONE batch: everything queued with the same settings and parameters as this call
    g_call_leader = true;
    // Callers of a thread pool come back together after a batch, staggered by whatever each does between two
    // calls (reading the next point, writing the last one out).  A batch costs about the same whatever its
    // size, so the leader waits for the pool to queue up again: until as many calls wait as the last batch
    // held, bounded by a fifth of the last batch's run time (1..20 ms).  With no history (first batch) or a
    // lone caller (last batch of 1) the bound is: at least 1 ms resp. no wait at all, and stop once the queue
    // has not grown for 1 ms.
    {
      using clock = std::chrono::steady_clock;
      const size_t expected = g_call_last_n;
      if (expected != 1)
      {
        const double budget_ms = std::min(20.0, std::max(1.0, 0.2 * g_call_last_ms));
        const auto t0 = clock::now();
        auto last_growth = t0;
        size_t seen = g_call_queue.size();
        while (g_call_queue.size() < expected || expected == 0)
        {
          g_call_cv.wait_for(lk, std::chrono::microseconds(250), [] { return false; });
          const auto now = clock::now();
          if (g_call_queue.size() != seen)
          {
            seen = g_call_queue.size();
            last_growth = now;
          }
          const double waited = std::chrono::duration<double, std::milli>(now - t0).count();
          const double idle = std::chrono::duration<double, std::milli>(now - last_growth).count();
          if (waited >= budget_ms) break;
          if (expected == 0 && waited >= 1.0 && idle >= 1.0) break;
        }
      }
    }
    std::vector<PendingCall*> batch, rest;
    for (PendingCall* c : g_call_queue)
    {
      const bool same = !std::memcmp(c->settings, me.settings, sizeof(InputSettings)) &&
                        !std::memcmp(c->params, me.params, sizeof(InputParameters));
      (same ? batch : rest).push_back(c);
    }
    g_call_queue.swap(rest);
    lk.unlock();
    const int n = static_cast<int>(batch.size());
    g_calls_total += n;
    ++g_call_batches_total;
    std::vector<OutputPointers*> outs(n);
    std::vector<const InputPointers*> ins(n);
    std::vector<const LocalParameters*> locs(n);
    for (int k = 0; k < n; ++k)
    {
      outs[k] = batch[k]->out;
      ins[k] = batch[k]->in;
      locs[k] = batch[k]->local;
    }
    const auto run_t0 = std::chrono::steady_clock::now();
    const int rc = roadsurf_run_batch(n, outs.data(), ins.data(), me.settings, me.params, locs.data(), 1, nullptr);
    const std::string err = (rc != RS_OK) ? std::string(roadsurf_last_error()) : std::string();
    lk.lock();
    g_call_last_n = static_cast<size_t>(n);
    g_call_last_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - run_t0).count();
    for (PendingCall* c : batch)
    {
      c->rc = rc;
      c->err = err;
      c->done = true;
    }
    // this caller's own call is in `batch`, hence done: the leadership goes to whoever still waits
    g_call_leader = false;
    g_call_cv.notify_all();
  }
  lk.unlock();
  if (me.rc != RS_OK)
  {
    std::fprintf(stderr, "roadsurf_b200 runsimulation failed: %s\n", me.err.c_str());
    g_err = me.err;
    mark_not_computed(outPointers, inSettings);
  }
}

