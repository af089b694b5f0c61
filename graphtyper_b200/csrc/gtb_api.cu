// gtb_api.cu -- C-ABI entry points (include/gtb200.h) and host orchestration (product code).
//
// Host responsibilities: graph flattening checks, packing one region into a single device arena (the k-mer index itself is
// built on the device, gtb_index_dev.cu; the host builder of gtb_index_host.hpp is the alternative), gathering the record
// columns of a submit into pinned staging (plain copies -- alignment units, aligned orientations and link checks are derived
// on the device), chunking and stream orchestration, accumulator download + saturation, PHRED conversion, the phase-support
// map.  There is NO CPU fallback for the compute path: without a usable CUDA device gtb_create(device >= 0) fails and every
// compute entry point returns GTB_ERR_CUDA.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <functional>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>
#include <sched.h>
#include <time.h>
#include <dlfcn.h>

#include "gtb_device.cuh"

using namespace gtb;

namespace
{
thread_local std::string g_err;

int fail(int code, const std::string & msg)
{
  g_err = msg;
  return code;
}

#define CUDA_TRY(expr)                                                                              \
  do                                                                                                \
  {                                                                                                 \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return fail(GTB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));               \
  } while (0)

struct DeviceBuffer
{
  void * p = nullptr;
  size_t cap = 0;
  int reserve(size_t n)
  {
    if (n <= cap)
      return 0;
    if (p)
      cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t const want = n + n / 4 + 4096;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess)
      return fail(GTB_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    cap = want;
    return 0;
  }
  void release()
  {
    if (p)
      cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct PinnedBuffer
{
  void * p = nullptr;
  size_t cap = 0;
  int reserve(size_t n)
  {
    if (n <= cap)
      return 0;
    if (p)
      cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t const want = n + n / 4 + 4096;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess)
      return fail(GTB_ERR_CUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
    cap = want;
    return 0;
  }
  void release()
  {
    if (p)
      cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct Region
{
  int id = -1;
  int slot = -1; // index into the device region table
  // host copies
  std::vector<uint32_t> bubble_order, n_alleles, score_off, cov_off;
  HostIndex index;
  uint32_t n_bubbles = 0;
  uint32_t depth_size = 0, reference_offset = 0; // SV graphs only
  int n_samples = 0;
  bool pool_open = false;
  // pool states that refuse further work until gtb_pool_reset / gtb_pool_begin:
  bool poisoned = false; // a submit failed after (possibly) adding part of its batch to the accumulators
  bool reduced = false;  // gtb_allreduce_accumulators summed the ranks' accumulators in place: another submit or another
                         // reduce would count the other ranks' reads again
  bool borrowed = false; // gtb_region_attach: graph + index arenas belong to another context
  // device
  DeviceBuffer arena;   // graph + index (host-built index) or graph only (device-built index)
  DeviceBuffer index_arena; // device-built index: labels + distinct k-mers + table + bitmap
  bool dev_index = false;
  uint64_t dev_n_keys = 0, dev_n_labels = 0;
  IdxRegion idx{};      // graph pointers of the device index build
  DeviceBuffer accum;   // accumulators (allocated at pool_begin)
  DevRegion dev{};      // pointers into arena/accum
  size_t accum_bytes = 0;
  // phasing connections (gtb_set_connections): open-addressing table keys[cap] | vals[cap] | state[4]
  DeviceBuffer conn;
  uint32_t conn_cap = 0;   // slots (power of two), 0 = off for this pool
  uint64_t conn_used = 0;  // occupied slots after the last submit (read back with the submit's counters)
};

// Everything one in-flight (chunk of a) submit owns.  Slot 0 doubles as the "last batch" of the debug taps.
struct BatchState
{
  DeviceBuffer d_batch, d_summaries, d_pool, d_counters, d_seedrecs, d_slow, d_task_times;
  PinnedBuffer h_batch, h_counters;
  LaunchParams P{};
  uint32_t n_tasks = 0;
  std::vector<int> regions;            // region ids of this chunk, in batch order
  std::vector<uint32_t> unit_begin;    // per region: first unit index (size n+1)
  std::vector<uint32_t> rec_begin;
  bool with_conn = false;              // some region of this chunk collects phasing connections
  int gather_launches = 0;             // gather_columns_kernel launches of this chunk (zero-copy staging)
  PrepParams prep{};                   // device-side batch preparation of this chunk
  BamParams bam{};                     // record parsing (gtb_submit_bam_records); bam.n == 0: columns came from the host
  DeviceBuffer d_bam, d_bam_sort;
  size_t bam_sort_bytes = 0;
  // gtb_submit_bgzf: compressed bytes + tables | inflated bytes | record lists and sort buffers | cub temp
  DeviceBuffer d_bgzf_in, d_bgzf_out, d_bgzf_tmp, d_bgzf_cub;
  PinnedBuffer h_bgzf;
  BgzfParams bgzf{};
  uint32_t bgzf_n = 0; // records the last gtb_submit_bgzf built (gtb_debug_bgzf_records)
  uint32_t bgzf_stitched = 0; // files of that call whose per-block record walks stitched (no serial walk)
  DeviceBuffer d_scan_temp;
  size_t scan_temp_bytes = 0;
  cudaStream_t stream = nullptr;       // probe + chain of this chunk (chunks run concurrently, each on its own stream)
  cudaStream_t stream_gen = nullptr, stream_slow = nullptr; // chain_general_kernel / first slow_kernel launch beside the first score pass
  cudaEvent_t ev[13] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // ev: 0 h2d start (copy stream), 1 h2d done (copy stream), 2 kernels start, 3 after prep + probe, 4 after both chain tiers,
  //     5 after the first score pass, 6 after the batch preparation (chunk stream), 7 after counters D2H (main stream),
  //     8 after chain_kernel; 9 / 10 around chain_general_kernel (stream_gen); 11 / 12 around the first slow_kernel launch
  //     (stream_slow)
  void release()
  {
    d_batch.release();
    d_summaries.release();
    d_pool.release();
    d_counters.release();
    d_seedrecs.release();
    d_task_times.release();
    d_scan_temp.release();
    d_bam.release();
    d_bam_sort.release();
    d_bgzf_in.release();
    d_bgzf_out.release();
    d_bgzf_tmp.release();
    d_bgzf_cub.release();
    h_bgzf.release();
    d_slow.release();
    h_batch.release();
    h_counters.release();
    for (auto & e : ev)
      if (e)
        cudaEventDestroy(e);
    if (stream)
      cudaStreamDestroy(stream);
    if (stream_gen)
      cudaStreamDestroy(stream_gen);
    if (stream_slow)
      cudaStreamDestroy(stream_slow);
  }
};

struct Ctx
{
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t copy_stream2 = nullptr; // second H2D queue: per-region copies of < 1 MB do not reach PCIe peak one at a time
  cudaEvent_t ev_copy2 = nullptr;
  cudaEvent_t ev_block = nullptr; // wait_stream() in blocking mode
  BatchState bs[MAX_CHUNKS];
  int n_chunks_last = 0;
  std::map<int, std::unique_ptr<Region>> regions;
  std::vector<int> slot_region; // slot -> region id (-1 free)
  DeviceBuffer d_regions;       // DevRegion table
  bool regions_dirty = true;
  // batch
  DeviceBuffer d_tap_counts, d_tap_pool, d_spill, d_huge;
  std::vector<DeviceBuffer> buffer_cache; // arenas / accumulators of ended regions, reused by the next regions
  PinnedBuffer h_stage, h_accum, h_varstats;
  DeviceBuffer d_varstats;
  DeviceBuffer d_gather;
  bool have_last = false;
  bool debug = false;
  int forced_chunks = 0; // gtb_set_chunks / GTB_CHUNKS: 0 = automatic
  cudaEvent_t ev_slow[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}; // [0],[1] around slow_kernel + huge_kernel (main stream);
                                                        // [2] main-stream position when a submit / replay starts;
                                                        // [4], [3] around the second score pass
  int connections = 0;   // gtb_set_connections: 0 off, otherwise table slots reserved per submitted record
  PinnedBuffer h_conn_state;
  float t_h2d = 0, t_align = 0, t_score = 0, t_d2h = 0, t_probe = 0, t_chain = 0, t_slow = 0, t_total = 0, t_prep = 0, t_score0 = 0, t_score1 = 0;
  unsigned long long last_n_slow = 0, last_n_gen = 0, last_n_active = 0, t0_reasons[16] = {0};
  float t_chain_fast = 0;
  int gen_lanes = 8; // chain_general_kernel: tasks per warp (GTB_GEN_LANES)
  // nccl (loaded lazily with dlopen, see gtb_nccl.cpp part below)
  void * nccl_lib = nullptr;
  void * nccl_comm = nullptr;
  int nccl_rank = 0, nccl_size = 1;
  // device-side index build (gtb_set_index_build / GTB_INDEX_BUILD=host|device)
  int index_build_device = 1;
  DeviceBuffer d_idx_small, d_idx_jobs, d_idx_keys, d_idx_keys2, d_idx_labels, d_idx_idx, d_idx_idx2, d_idx_head, d_idx_temp;
  PinnedBuffer h_idx_small;
  // discovery re-alignment (gtb_sw_align_batch)
  DeviceBuffer d_sw_in, d_sw_out, d_sw_bt;
  PinnedBuffer h_sw;
  SwParams sw_last{};
  int sw_warps = 0;
  cudaEvent_t sw_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float t_sw_kernel = 0, t_sw_h2d = 0, t_sw_d2h = 0;
};

// Region arenas are recycled: cudaMalloc / cudaFree of ~10 MB blocks per region would otherwise dominate (and add
// multi-millisecond jitter to) the per-region setup of a long run.
int take_buffer(Ctx * c, DeviceBuffer & dst, size_t bytes)
{
  if (dst.cap >= bytes)
    return 0;
  dst.release();
  int best = -1;
  for (size_t i = 0; i < c->buffer_cache.size(); ++i)
    if (c->buffer_cache[i].cap >= bytes && (best < 0 || c->buffer_cache[i].cap < c->buffer_cache[best].cap))
      best = (int)i;
  if (best >= 0)
  {
    dst = c->buffer_cache[best];
    c->buffer_cache.erase(c->buffer_cache.begin() + best);
    return 0;
  }
  return dst.reserve(bytes);
}

void give_buffer(Ctx * c, DeviceBuffer & b)
{
  if (b.p)
  {
    if (c->buffer_cache.size() < 256)
      c->buffer_cache.push_back(b);
    else
      cudaFree(b.p);
  }
  b.p = nullptr;
  b.cap = 0;
}

// Waits for everything queued on the main stream.  cudaStreamSynchronize spins: lowest latency, but with several ranks and
// pool threads per node every waiting thread burns a core the staging threads need (8 ranks x 3 pool threads on a 32-core
// box).  A wait that sleeps on a cudaEventBlockingSync event frees the core but wakes up late (measured: 0.94 -> 1.27 ms per
// step with 4 pool threads).  Default = polite polling: cudaStreamQuery, giving the core away between polls (sched_yield),
// short naps once the wait has lasted 200 us.  GTB_SYNC=spin|block selects the other two.
cudaError_t wait_stream(Ctx * c)
{
  static int const mode = []() {
    const char * e = getenv("GTB_SYNC");
    return !e ? 2 : strcmp(e, "block") == 0 ? 1 : strcmp(e, "spin") == 0 ? 0 : 2;
  }();
  if (mode == 0)
    return cudaStreamSynchronize(c->stream);
  if (mode == 1)
  {
    if (!c->ev_block)
    {
      cudaError_t const e = cudaEventCreateWithFlags(&c->ev_block, cudaEventBlockingSync | cudaEventDisableTiming);
      if (e != cudaSuccess)
        return e;
    }
    cudaError_t e = cudaEventRecord(c->ev_block, c->stream);
    if (e == cudaSuccess)
      e = cudaEventSynchronize(c->ev_block);
    return e;
  }
  auto const t0 = std::chrono::steady_clock::now();
  for (unsigned n = 0;; ++n)
  {
    cudaError_t const e = cudaStreamQuery(c->stream);
    if (e != cudaErrorNotReady)
      return e;
    if ((n & 15u) == 15u && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(200))
    {
      timespec ts{0, 20000}; // 20 us
      nanosleep(&ts, nullptr);
    }
    else
      sched_yield();
  }
}

// Pools of different contexts (one per pool thread) share the device.  Left alone, their H2D transfers interleave copy by copy,
// so pools that were issued one after the other get their bases together, late.  A per-device token keeps the issue order:
// the transfers of a submit start after those of the submit issued before it, whichever context that was.  The same token
// for the wide kernels (probe_kernel owns every SM's register file, chain_kernel most of it) was measured and is off by
// default: end to end 0.999 -> 0.954 ms per step with both tokens, 0.968 with the copy token alone, but the device-resident
// replay loses (0.725 -> 0.775 ms): the fronts of two pools do overlap usefully at their edges.  A token is ONE event per
// device: wait for its latest record, queue the work, record it again -- under a mutex, so the chain of records is the issue
// order.  GTB_FIFO: bit 0 = kernels, bit 1 = copies (default 2).
struct DeviceFifo
{
  std::mutex m_front, m_copy;
  cudaEvent_t front = nullptr, copy = nullptr; // created on first use, never destroyed (another context may still wait on them)
};
DeviceFifo & fifo_of(int device)
{
  static DeviceFifo f[64];
  return f[device & 63];
}
int fifo_mode()
{
  static int const mode = []() { const char * e = getenv("GTB_FIFO"); return e ? atoi(e) : 2; }();
  return mode;
}
// waits (on stream s) for the token's latest record; creates the event on first use
cudaError_t fifo_wait(cudaEvent_t & ev, cudaStream_t s)
{
  if (!ev)
    return cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  return cudaStreamWaitEvent(s, ev, 0);
}

int upload_region_table(Ctx * c)
{
  if (!c->regions_dirty)
    return 0;
  size_t const n = c->slot_region.size();
  std::vector<DevRegion> tab(std::max<size_t>(n, 1));
  for (size_t s = 0; s < n; ++s)
    if (c->slot_region[s] >= 0)
      tab[s] = c->regions[c->slot_region[s]]->dev;
  if (int rc = c->d_regions.reserve(tab.size() * sizeof(DevRegion)))
    return rc;
  CUDA_TRY(cudaMemcpyAsync(c->d_regions.p, tab.data(), tab.size() * sizeof(DevRegion), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(wait_stream(c));
  c->regions_dirty = false;
  return 0;
}

int validate_graph(const gtb_graph_view * g)
{
  if (!g || g->n_ref == 0)
    return fail(GTB_ERR_ARG, "empty graph view");
  if (!g->ref_order || !g->ref_seq_off || !g->ref_var_off || !g->seq)
    return fail(GTB_ERR_ARG, "graph view has null arrays");
  if (g->n_var && (!g->var_order || !g->var_seq_off || !g->var_out_ref))
    return fail(GTB_ERR_ARG, "graph view has null var arrays");
  if (g->seq_len >= 0xFFFFFFFFull)
    return fail(GTB_ERR_ARG, "graph sequence larger than 4 GiB");
  if (g->ref_var_off[0] != 0 || g->ref_var_off[g->n_ref] != g->n_var)
    return fail(GTB_ERR_ARG, "ref_var_off is not a CSR over the var nodes");
  if (g->ref_var_off[g->n_ref - 1] != g->n_var)
    return fail(GTB_ERR_ARG, "the last ref node must have no out variants");
  for (uint32_t r = 0; r + 1 < g->n_ref; ++r)
  {
    uint32_t const deg = g->ref_var_off[r + 1] - g->ref_var_off[r];
    if (deg < 2)
      return fail(GTB_ERR_ARG, "every bubble needs a reference allele and at least one alternative allele "
                               "(the reference drops records without alts, graph.cpp:60-71)");
    if (deg > (uint32_t)MAX_ALLELES)
      return fail(GTB_ERR_CAPACITY, "bubble with more than 32 alleles: not supported by the device path yet");
    if (g->ref_order[r] > g->ref_order[r + 1])
      return fail(GTB_ERR_ARG, "ref node orders must be non-decreasing");
    for (uint32_t v = g->ref_var_off[r]; v < g->ref_var_off[r + 1]; ++v)
      if (g->var_out_ref[v] != r + 1)
        return fail(GTB_ERR_ARG, "var_out_ref does not match the bubble structure");
  }
  return 0;
}

// Persistent fork-join pool for the host-side staging work (regions are independent jobs).
class WorkPool
{
public:
  WorkPool()
  {
    // cores this process may use, shared with the other ranks of the node (torchrun exports LOCAL_WORLD_SIZE): 8 ranks
    // with 8 staging threads each on a 32-core box only fight each other
    int hw = (int)std::thread::hardware_concurrency();
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0)
      hw = CPU_COUNT(&set);
    int local_world = 1;
    if (const char * e = getenv("LOCAL_WORLD_SIZE"))
      local_world = std::max(1, atoi(e));
    if (const char * e = getenv("GTB_HOST_THREADS"))
      hw = std::max(1, atoi(e)) * local_world;
    n_workers_ = std::max(1, std::min(8, (hw > 0 ? hw : 1) / local_world)) - 1; // the caller thread works too
    for (int t = 0; t < n_workers_; ++t)
      threads_.emplace_back([this]() { worker(); });
  }
  ~WorkPool()
  {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      ++epoch_;
    }
    cv_.notify_all();
    for (auto & t : threads_)
      t.join();
  }
  void run(int n, const std::function<void(int)> & fn)
  {
    if (n <= 0)
      return;
    if (n == 1 || n_workers_ == 0)
    {
      for (int i = 0; i < n; ++i)
        fn(i);
      return;
    }
    // one fork-join at a time; a caller that finds the pool busy (another context's submit is staging) does its own
    // blocks inline instead of queueing behind it -- several pool threads then stage in parallel
    std::unique_lock<std::mutex> serial(run_m_, std::try_to_lock);
    if (!serial.owns_lock())
    {
      for (int i = 0; i < n; ++i)
        fn(i);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(m_);
      fn_ = &fn;
      n_ = n;
      next_.store(0);
      pending_ = n_workers_;
      ++epoch_;
    }
    cv_.notify_all();
    drain();
    std::unique_lock<std::mutex> lk(m_);
    done_cv_.wait(lk, [this]() { return pending_ == 0; });
    fn_ = nullptr;
  }

private:
  void drain()
  {
    for (;;)
    {
      int const i = next_.fetch_add(1);
      if (i >= n_)
        break;
      (*fn_)(i);
    }
  }
  void worker()
  {
    unsigned long seen = 0;
    for (;;)
    {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&]() { return epoch_ != seen; });
        seen = epoch_;
        if (stop_)
          return;
      }
      drain();
      {
        std::lock_guard<std::mutex> lk(m_);
        --pending_;
      }
      done_cv_.notify_one();
    }
  }
  std::vector<std::thread> threads_;
  std::mutex m_, run_m_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(int)> * fn_ = nullptr;
  std::atomic<int> next_{0};
  int n_ = 0, n_workers_ = 0, pending_ = 0;
  unsigned long epoch_ = 0;
  bool stop_ = false;
};

void parallel_for(int n, const std::function<void(int)> & fn)
{
  static WorkPool pool;
  pool.run(n, fn);
}

// Segment table of a gather / zero launch.  It is handed to the kernels by value (parameter space): nothing is staged, so
// nothing can be overwritten while an asynchronous copy of it is still in flight.
template <typename F>
std::vector<Segment> make_segments(int n, F make, unsigned long long & max_bytes)
{
  std::vector<Segment> seg((size_t)n);
  max_bytes = 0;
  for (int i = 0; i < n; ++i)
  {
    seg[i] = make(i);
    max_bytes = std::max(max_bytes, seg[i].bytes);
  }
  return seg;
}

// ---- phasing-connection table of one pool
size_t conn_table_bytes(uint32_t cap) { return (size_t)cap * 12 + 16; }

int conn_alloc(Ctx * c, Region & R, uint32_t cap)
{
  if (int rc = take_buffer(c, R.conn, conn_table_bytes(cap)))
    return rc;
  CUDA_TRY(cudaMemsetAsync(R.conn.p, 0, conn_table_bytes(cap), c->stream));
  uint8_t * d = static_cast<uint8_t *>(R.conn.p);
  R.dev.conn_keys = reinterpret_cast<unsigned long long *>(d);
  R.dev.conn_vals = reinterpret_cast<uint32_t *>(d + (size_t)cap * 8);
  R.dev.conn_state = reinterpret_cast<uint32_t *>(d + (size_t)cap * 12);
  R.dev.conn_mask = cap - 1;
  R.conn_cap = cap;
  c->regions_dirty = true;
  return 0;
}

// Keeps the table at most half full after `n_new` more records (c->connections slots budgeted per record; an insert that
// still finds the table full raises GTB_ERR_CAPACITY after the submit).  Growth = new zeroed table + re-insert on the device.
int conn_reserve(Ctx * c, Region & R, uint64_t n_new)
{
  if (!R.conn_cap)
    return 0;
  uint64_t const need = (R.conn_used + (uint64_t)c->connections * n_new) * 2;
  if (need <= R.conn_cap)
    return 0;
  uint64_t ncap = R.conn_cap;
  while (ncap < need)
    ncap <<= 1;
  if (ncap > (1ull << 30))
    return fail(GTB_ERR_CAPACITY, "phasing-connection table would exceed 2^30 slots");
  DeviceBuffer old = R.conn;
  DevRegion const old_dev = R.dev;
  uint32_t const old_cap = R.conn_cap;
  R.conn = DeviceBuffer();
  if (int rc = conn_alloc(c, R, (uint32_t)ncap))
    return rc;
  launch_conn_rehash(old_dev.conn_keys, old_dev.conn_vals, old_cap, R.dev, c->stream);
  CUDA_TRY(wait_stream(c));
  give_buffer(c, old);
  return 0;
}

template <typename T>
size_t place(size_t & off, size_t count)
{
  size_t const o = align_up(off, 16);
  off = o + count * sizeof(T);
  return o;
}

} // namespace

extern "C"
{
const char * gtb_last_error(void) { return g_err.c_str(); }
const char * gtb_version(void) { return "graphtyper_b200 0.1 (sm_100a)"; }

int gtb_create(int device_id, gtb_ctx ** out)
{
  if (!out)
    return fail(GTB_ERR_ARG, "null out");
  auto * c = new Ctx();
  c->device = device_id;
  if (device_id >= 0)
  {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= device_id)
    {
      delete c;
      return fail(GTB_ERR_CUDA, std::string("no usable CUDA device ") + std::to_string(device_id) + ": " +
                                  (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range") +
                                  " (the compute path has no CPU fallback)");
    }
    if (const char * gl = getenv("GTB_GEN_LANES"))
      c->gen_lanes = std::max(1, std::min(32, atoi(gl)));
    e = cudaSetDevice(device_id);
    if (e == cudaSuccess)
      e = (cudaError_t)upload_hash_tables_kernels();
    if (e == cudaSuccess)
      e = (cudaError_t)upload_hash_tables_index();
    if (e == cudaSuccess)
      e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess)
      e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess)
      e = cudaStreamCreateWithFlags(&c->copy_stream2, cudaStreamNonBlocking);
    if (e == cudaSuccess)
      e = cudaEventCreateWithFlags(&c->ev_copy2, cudaEventDisableTiming);

    for (int k = 0; k < MAX_CHUNKS; ++k)
    {
      for (int i = 0; i < 13 && e == cudaSuccess; ++i)
        e = cudaEventCreate(&c->bs[k].ev[i]);
      if (e == cudaSuccess)
        e = cudaStreamCreateWithFlags(&c->bs[k].stream, cudaStreamNonBlocking);
      if (e == cudaSuccess)
        e = cudaStreamCreateWithFlags(&c->bs[k].stream_gen, cudaStreamNonBlocking);
      if (e == cudaSuccess)
        e = cudaStreamCreateWithFlags(&c->bs[k].stream_slow, cudaStreamNonBlocking);
    }
    for (int i = 0; i < 5 && e == cudaSuccess; ++i)
      e = cudaEventCreate(&c->ev_slow[i]);
    if (e != cudaSuccess)
    {
      delete c;
      return fail(GTB_ERR_CUDA, std::string("CUDA init failed: ") + cudaGetErrorString(e));
    }
  }
  *out = reinterpret_cast<gtb_ctx *>(c);
  return 0;
}

void gtb_destroy(gtb_ctx * ctx)
{
  if (!ctx)
    return;
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (c->device >= 0)
  {
    cudaSetDevice(c->device);
    if (c->stream)
      cudaStreamSynchronize(c->stream);
    for (auto & kv : c->regions)
    {
      kv.second->arena.release();
      kv.second->index_arena.release();
      kv.second->accum.release();
      kv.second->conn.release();
    }
    c->d_regions.release();
    for (auto & b : c->buffer_cache)
      b.release();
    c->buffer_cache.clear();
    for (auto & B : c->bs)
      B.release();
    c->d_tap_counts.release();
    c->d_tap_pool.release();
    c->d_spill.release();
    c->d_huge.release();
    c->h_stage.release();
    c->h_accum.release();
    c->h_conn_state.release();
    c->h_varstats.release();
    c->d_varstats.release();
    c->d_gather.release();
    for (DeviceBuffer * b : {&c->d_idx_small, &c->d_idx_jobs, &c->d_idx_keys, &c->d_idx_keys2, &c->d_idx_labels, &c->d_idx_idx,
                             &c->d_idx_idx2, &c->d_idx_head, &c->d_idx_temp})
      b->release();
    c->h_idx_small.release();
    c->d_sw_in.release();
    c->d_sw_out.release();
    c->d_sw_bt.release();
    c->h_sw.release();
    for (auto & e : c->sw_ev)
      if (e)
        cudaEventDestroy(e);
    for (auto & e : c->ev_slow)
      if (e)
        cudaEventDestroy(e);
    if (c->copy_stream)
      cudaStreamDestroy(c->copy_stream);
    if (c->copy_stream2)
      cudaStreamDestroy(c->copy_stream2);
    if (c->ev_copy2)
      cudaEventDestroy(c->ev_copy2);
    if (c->ev_block)
      cudaEventDestroy(c->ev_block);

    if (c->stream)
      cudaStreamDestroy(c->stream);
  }
  delete c;
}

// Uploads one prepared region: graph + labels + distinct k-mer list in ONE H2D copy, then the k-mer table is
// built on the device (the 16-byte-slot table is 4-8x larger than the list it is built from).
static int upload_region(Ctx * c, Region & R, const gtb_graph_view * g, bool dev_index)
{
  size_t off = 0;
  size_t const o_ref_order = place<uint32_t>(off, g->n_ref);
  size_t const o_ref_seq_off = place<uint32_t>(off, g->n_ref + 1);
  size_t const o_ref_var_off = place<uint32_t>(off, g->n_ref + 1);
  size_t const o_var_order = place<uint32_t>(off, g->n_var);
  size_t const o_var_seq_off = place<uint32_t>(off, g->n_var + 1);
  size_t const o_var_out_ref = place<uint32_t>(off, g->n_var);
  size_t const o_seq = place<uint8_t>(off, g->seq_len + 16);
  size_t const o_actual = place<uint32_t>(off, g->n_special);
  size_t const o_rreach = place<uint32_t>(off, g->n_special);
  size_t const o_sp_keys = place<uint32_t>(off, g->n_sp_keys);
  size_t const o_sp_off = place<uint32_t>(off, g->n_sp_keys + 1);
  size_t const n_sp_list = g->n_sp_keys ? g->sp_off[g->n_sp_keys] : 0;
  size_t const o_sp_list = place<uint32_t>(off, n_sp_list);
  size_t const o_bubble_order = place<uint32_t>(off, R.n_bubbles);
  size_t const o_score_off = place<uint32_t>(off, R.n_bubbles + 1);
  size_t const o_cov_off = place<uint32_t>(off, R.n_bubbles + 1);
  // id2hap (vcf_writer.cpp:84) as a direct table: bubble index by (var order - first bubble order); the score kernels would
  // otherwise binary-search bubble_order for every bubble of every path (9 dependent loads)
  uint32_t hap_base = 0, hap_span = 0;
  if (R.n_bubbles > 0 && R.n_bubbles < 0xFFFFu)
  {
    uint32_t const lo = *std::min_element(R.bubble_order.begin(), R.bubble_order.end());
    uint32_t const hi = *std::max_element(R.bubble_order.begin(), R.bubble_order.end());
    if ((uint64_t)hi - lo < (1ull << 24))
    {
      hap_base = lo;
      hap_span = hi - lo + 1;
    }
  }
  size_t const o_hap = place<uint16_t>(off, hap_span);
  // allele number of every var node, and "last ref node at or before a position" per 16 positions (chain kernels)
  size_t const o_var_num = place<uint8_t>(off, g->n_var);
  uint32_t const pos_base = g->ref_order[0];
  uint32_t const n_pos_bucket = ((g->ref_order[g->n_ref - 1] - pos_base) >> 4) + 1;
  size_t const o_pos_bucket = place<uint32_t>(off, n_pos_bucket);
  // device index build: events + the sweep-order node table instead of a host-built index
  bool const have_ev = dev_index && g->var_ev_off && g->var_aev_off && g->var_ev && g->var_aev;
  size_t const n_ev = have_ev ? g->var_ev_off[g->n_var] : 0, n_aev = have_ev ? g->var_aev_off[g->n_var] : 0;
  uint32_t const n_sweep = dev_index ? g->n_ref + g->n_var : 0;
  size_t const o_ev_off = place<uint32_t>(off, have_ev ? g->n_var + 1 : 0);
  size_t const o_aev_off = place<uint32_t>(off, have_ev ? g->n_var + 1 : 0);
  size_t const o_ev = place<int64_t>(off, n_ev);
  size_t const o_aev = place<int64_t>(off, n_aev);
  size_t const o_sweep_node = place<uint32_t>(off, n_sweep);
  size_t const o_sweep_off = place<uint32_t>(off, dev_index ? n_sweep + 1 : 0);
  size_t const o_labels = place<DevLabel>(off, dev_index ? 0 : R.index.labels.size());
  size_t const o_uniq = place<IndexSlot>(off, dev_index ? 0 : R.index.uniq.size());
  size_t const upload_bytes = align_up(off, 256);
  size_t const o_table = place<IndexSlot>(off, dev_index ? 0 : R.index.table_cap); // device only
  size_t const o_bitmap = place<uint32_t>(off, dev_index ? 0 : (size_t)R.index.table_cap * 4 / 32 + 1); // follows the table
  size_t const total = align_up(off, 256);

  if (int rc = c->h_stage.reserve(upload_bytes))
    return rc;
  uint8_t * h = static_cast<uint8_t *>(c->h_stage.p);
  auto put32 = [&](size_t o, const uint32_t * src, size_t n)
  {
    if (n)
      memcpy(h + o, src, n * 4);
  };
  auto put64as32 = [&](size_t o, const uint64_t * src, size_t n)
  {
    uint32_t * d = reinterpret_cast<uint32_t *>(h + o);
    for (size_t i = 0; i < n; ++i)
      d[i] = (uint32_t)src[i];
  };
  put32(o_ref_order, g->ref_order, g->n_ref);
  put64as32(o_ref_seq_off, g->ref_seq_off, g->n_ref + 1);
  put32(o_ref_var_off, g->ref_var_off, g->n_ref + 1);
  put32(o_var_order, g->var_order, g->n_var);
  put64as32(o_var_seq_off, g->var_seq_off, g->n_var + 1);
  put32(o_var_out_ref, g->var_out_ref, g->n_var);
  memcpy(h + o_seq, g->seq, g->seq_len);
  memset(h + o_seq + g->seq_len, 0, 16);
  put32(o_actual, g->actual_poses, g->n_special);
  put32(o_rreach, g->ref_reach_poses, g->n_special);
  put32(o_sp_keys, g->sp_keys, g->n_sp_keys);
  if (g->n_sp_keys)
    put32(o_sp_off, g->sp_off, g->n_sp_keys + 1);
  else
    memset(h + o_sp_off, 0, 4);
  put32(o_sp_list, g->sp_list, n_sp_list);
  put32(o_bubble_order, R.bubble_order.data(), R.n_bubbles);
  put32(o_score_off, R.score_off.data(), R.n_bubbles + 1);
  put32(o_cov_off, R.cov_off.data(), R.n_bubbles + 1);
  if (hap_span)
  {
    uint16_t * tab = reinterpret_cast<uint16_t *>(h + o_hap);
    memset(tab, 0xFF, (size_t)hap_span * 2);
    for (uint32_t b = 0; b < R.n_bubbles; ++b)
      tab[R.bubble_order[b] - hap_base] = (uint16_t)b; // ascending b: on duplicate orders the last bubble wins, like id2hap
  }
  {
    uint8_t * vn = h + o_var_num;
    for (uint32_t r = 0; r + 1 < g->n_ref; ++r)
      for (uint32_t v = g->ref_var_off[r]; v < g->ref_var_off[r + 1]; ++v)
        vn[v] = (uint8_t)(v - g->ref_var_off[r]);
    uint32_t * pb = reinterpret_cast<uint32_t *>(h + o_pos_bucket);
    uint32_t r = 0;
    for (uint32_t b = 0; b < n_pos_bucket; ++b)
    {
      uint32_t const pos = pos_base + 16u * b;
      while (r + 1 < g->n_ref && g->ref_order[r + 1] <= pos)
        ++r;
      pb[b] = r;
    }
  }
  uint32_t n_jobs = 0;
  if (dev_index)
  {
    if (have_ev)
    {
      put32(o_ev_off, g->var_ev_off, g->n_var + 1);
      put32(o_aev_off, g->var_aev_off, g->n_var + 1);
      memcpy(h + o_ev, g->var_ev, n_ev * 8);
      memcpy(h + o_aev, g->var_aev, n_aev * 8);
    }
    // sweep order of indexer.cpp:246-291: ref node r, then the alleles of bubble r in order
    uint32_t * sn = reinterpret_cast<uint32_t *>(h + o_sweep_node);
    uint32_t * so = reinterpret_cast<uint32_t *>(h + o_sweep_off);
    uint32_t k = 0;
    for (uint32_t r = 0; r < g->n_ref; ++r)
    {
      sn[k] = r;
      so[k++] = n_jobs;
      n_jobs += (uint32_t)(g->ref_seq_off[r + 1] - g->ref_seq_off[r]);
      for (uint32_t v = g->ref_var_off[r]; v < g->ref_var_off[r + 1]; ++v)
      {
        sn[k] = v | 0x80000000u;
        so[k++] = n_jobs;
        n_jobs += (uint32_t)(g->var_seq_off[v + 1] - g->var_seq_off[v]);
      }
    }
    so[k] = n_jobs;
  }
  else
  {
    memcpy(h + o_labels, R.index.labels.data(), R.index.labels.size() * sizeof(DevLabel));
    memcpy(h + o_uniq, R.index.uniq.data(), R.index.uniq.size() * sizeof(IndexSlot));
  }

  if (int rc = take_buffer(c, R.arena, total))
    return rc;
  uint8_t * d = static_cast<uint8_t *>(R.arena.p);
  CUDA_TRY(cudaMemcpyAsync(d, h, upload_bytes, cudaMemcpyHostToDevice, c->stream));
  if (!dev_index)
  {
    CUDA_TRY(cudaMemsetAsync(d + o_table, 0, total - o_table, c->stream)); // table + bitmap
    launch_build_table(reinterpret_cast<const IndexSlot *>(d + o_uniq), (uint32_t)R.index.uniq.size(),
                       reinterpret_cast<IndexSlot *>(d + o_table), R.index.table_mask, R.index.table_shift,
                       reinterpret_cast<uint32_t *>(d + o_bitmap), c->stream);
  }
  CUDA_TRY(wait_stream(c)); // h_stage is reused by the next region
  CUDA_TRY(cudaGetLastError());

  DevRegion & D = R.dev;
  memset(&D, 0, sizeof(D));
  D.n_ref = g->n_ref;
  D.n_var = g->n_var;
  D.n_special = g->n_special;
  D.n_sp_keys = g->n_sp_keys;
  D.is_sv = g->is_sv_graph ? 1u : 0u;
  D.n_bubbles = R.n_bubbles;
  D.n_samples = 0;
  D.table_mask = R.index.table_mask;
  D.table_shift = R.index.table_shift;
  D.ref_order = reinterpret_cast<const uint32_t *>(d + o_ref_order);
  D.ref_seq_off = reinterpret_cast<const uint32_t *>(d + o_ref_seq_off);
  D.ref_var_off = reinterpret_cast<const uint32_t *>(d + o_ref_var_off);
  D.var_order = reinterpret_cast<const uint32_t *>(d + o_var_order);
  D.var_seq_off = reinterpret_cast<const uint32_t *>(d + o_var_seq_off);
  D.var_out_ref = reinterpret_cast<const uint32_t *>(d + o_var_out_ref);
  D.seq = d + o_seq;
  D.actual_poses = reinterpret_cast<const uint32_t *>(d + o_actual);
  D.ref_reach_poses = reinterpret_cast<const uint32_t *>(d + o_rreach);
  D.sp_keys = reinterpret_cast<const uint32_t *>(d + o_sp_keys);
  D.sp_off = reinterpret_cast<const uint32_t *>(d + o_sp_off);
  D.sp_list = reinterpret_cast<const uint32_t *>(d + o_sp_list);
  D.bubble_order = reinterpret_cast<const uint32_t *>(d + o_bubble_order);
  D.hap_of_order = hap_span ? reinterpret_cast<const uint16_t *>(d + o_hap) : nullptr;
  D.hap_base = hap_base;
  D.hap_span = hap_span;
  D.var_num = d + o_var_num;
  D.pos_bucket = reinterpret_cast<const uint32_t *>(d + o_pos_bucket);
  D.pos_base = pos_base;
  D.n_pos_bucket = n_pos_bucket;
  D.score_off = reinterpret_cast<const uint32_t *>(d + o_score_off);
  D.cov_off = reinterpret_cast<const uint32_t *>(d + o_cov_off);
  D.table = reinterpret_cast<const IndexSlot *>(d + o_table);
  D.bitmap = reinterpret_cast<const uint32_t *>(d + o_bitmap);
  D.labels = reinterpret_cast<const DevLabel *>(d + o_labels);
  D.depth_size = R.depth_size;
  D.reference_offset = R.reference_offset;
  R.dev_index = dev_index;
  if (dev_index)
  {
    IdxRegion & X = R.idx;
    memset(&X, 0, sizeof(X));
    X.n_ref = g->n_ref;
    X.n_var = g->n_var;
    X.n_sp_keys = g->n_sp_keys;
    X.n_sweep = n_sweep;
    X.ref_order = D.ref_order;
    X.ref_seq_off = D.ref_seq_off;
    X.ref_var_off = D.ref_var_off;
    X.var_order = D.var_order;
    X.var_seq_off = D.var_seq_off;
    X.var_out_ref = D.var_out_ref;
    X.seq = D.seq;
    X.sp_keys = D.sp_keys;
    X.sp_off = D.sp_off;
    X.sp_list = D.sp_list;
    if (have_ev)
    {
      X.var_ev_off = reinterpret_cast<const uint32_t *>(d + o_ev_off);
      X.var_aev_off = reinterpret_cast<const uint32_t *>(d + o_aev_off);
      X.var_ev = reinterpret_cast<const int64_t *>(d + o_ev);
      X.var_aev = reinterpret_cast<const int64_t *>(d + o_aev);
    }
    X.sweep_node = reinterpret_cast<const uint32_t *>(d + o_sweep_node);
    X.sweep_job_off = reinterpret_cast<const uint32_t *>(d + o_sweep_off);
    R.dev_n_labels = n_jobs; // number of END positions until the count pass has run
  }

  int slot = -1;
  for (size_t s = 0; s < c->slot_region.size(); ++s)
    if (c->slot_region[s] < 0)
    {
      slot = (int)s;
      break;
    }
  if (slot < 0)
  {
    slot = (int)c->slot_region.size();
    c->slot_region.push_back(-1);
  }
  if (slot > 0xFFFF)
    return fail(GTB_ERR_CAPACITY, "more than 65536 resident regions");
  c->slot_region[slot] = R.id;
  R.slot = slot;
  c->regions_dirty = true;
  return 0;
}

// Device-side index build of all regions of one gtb_region_begin_multi call (kernels: gtb_index_dev.cu).  Two stream
// synchronisations: after the count pass (the label totals size the per-region index arenas) and at the end.
static int build_indexes_on_device(Ctx * c, std::vector<std::unique_ptr<Region>> & regs)
{
  uint32_t const n = (uint32_t)regs.size();
  std::vector<uint32_t> job_off(n + 1, 0);
  for (uint32_t i = 0; i < n; ++i)
  {
    uint64_t const next = (uint64_t)job_off[i] + regs[i]->dev_n_labels; // = END positions (set by upload_region)
    if (next >= 0x7FFFFFFFull)
      return fail(GTB_ERR_CAPACITY, "more than 2^31 graph bases in one gtb_region_begin_multi call");
    job_off[i + 1] = (uint32_t)next;
  }
  uint32_t const total_jobs = job_off[n];
  // small arrays: [IdxRegion x n][region_job_off n+1][region_tuple_off n+1][region_n_uniq n][err 1]
  size_t so = 0;
  size_t const o_desc = place<IdxRegion>(so, n);
  size_t const o_rjo = place<uint32_t>(so, n + 1);
  size_t const o_rto = place<uint32_t>(so, n + 1);
  size_t const o_nu = place<uint32_t>(so, n);
  size_t const o_err = place<uint32_t>(so, 1);
  size_t const small_bytes = align_up(so, 256);
  if (int rc = c->d_idx_small.reserve(small_bytes))
    return rc;
  if (int rc = c->h_idx_small.reserve(small_bytes))
    return rc;
  uint8_t * hs = static_cast<uint8_t *>(c->h_idx_small.p);
  uint8_t * ds = static_cast<uint8_t *>(c->d_idx_small.p);
  memset(hs, 0, small_bytes);
  IdxRegion * hdesc = reinterpret_cast<IdxRegion *>(hs + o_desc);
  for (uint32_t i = 0; i < n; ++i)
    hdesc[i] = regs[i]->idx;
  memcpy(hs + o_rjo, job_off.data(), (n + 1) * 4);
  CUDA_TRY(cudaMemcpyAsync(ds, hs, small_bytes, cudaMemcpyHostToDevice, c->stream));
  const IdxRegion * d_desc = reinterpret_cast<const IdxRegion *>(ds + o_desc);
  const uint32_t * d_rjo = reinterpret_cast<const uint32_t *>(ds + o_rjo);
  uint32_t * d_rto = reinterpret_cast<uint32_t *>(ds + o_rto);
  uint32_t * d_nu = reinterpret_cast<uint32_t *>(ds + o_nu);
  uint32_t * d_err = reinterpret_cast<uint32_t *>(ds + o_err);

  // 1-2. count + scan
  size_t const jobs_bytes = align_up((size_t)(total_jobs + 1) * 4);
  if (int rc = c->d_idx_jobs.reserve(jobs_bytes * 2))
    return rc;
  uint32_t * d_cnt = static_cast<uint32_t *>(c->d_idx_jobs.p);
  uint32_t * d_joff = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(c->d_idx_jobs.p) + jobs_bytes);
  size_t scan_bytes = idx_scan_temp_bytes(total_jobs + 1);
  if (int rc = c->d_idx_temp.reserve(scan_bytes))
    return rc;
  CUDA_TRY(cudaMemsetAsync(d_cnt + total_jobs, 0, 4, c->stream));
  idx_launch_count(d_desc, n, d_rjo, total_jobs, d_cnt, d_err, c->stream);
  if (idx_exclusive_scan(c->d_idx_temp.p, c->d_idx_temp.cap, d_cnt, d_joff, total_jobs + 1, c->stream))
    return fail(GTB_ERR_CUDA, "index build: scan failed");
  idx_launch_region_totals(d_joff, d_rjo, n, d_rto, c->stream);
  CUDA_TRY(cudaMemcpyAsync(hs + o_rto, d_rto, (n + 1) * 4 + 0, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(hs + o_err, d_err, 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(wait_stream(c));
  CUDA_TRY(cudaGetLastError());
  const uint32_t * rto = reinterpret_cast<const uint32_t *>(hs + o_rto);
  if (*reinterpret_cast<const uint32_t *>(hs + o_err))
    return fail(GTB_ERR_ARG, "index build: inconsistent graph view (special position table / more than 64 nested empty nodes)");
  uint32_t const total = rto[n];

  // per-region index arenas (labels, distinct k-mers, table, bitmap), geometry as in HostIndex
  for (uint32_t i = 0; i < n; ++i)
  {
    Region & R = *regs[i];
    size_t const nl = rto[i + 1] - rto[i];
    size_t cap = 16;
    while (cap < nl * 4 + 2)
      cap <<= 1;
    int shift = 64;
    for (size_t x = cap; x > 1; x >>= 1)
      --shift;
    size_t ao = 0;
    size_t const o_labels = place<DevLabel>(ao, nl);
    size_t const o_uniq = place<IndexSlot>(ao, nl);
    size_t const o_table = place<IndexSlot>(ao, cap);
    size_t const o_bitmap = place<uint32_t>(ao, cap * 4 / 32 + 1);
    size_t const atotal = align_up(ao, 256);
    if (int rc = take_buffer(c, R.index_arena, atotal))
      return rc;
    uint8_t * a = static_cast<uint8_t *>(R.index_arena.p);
    CUDA_TRY(cudaMemsetAsync(a + o_table, 0, atotal - o_table, c->stream));
    IdxRegion & X = hdesc[i];
    X.labels = reinterpret_cast<DevLabel *>(a + o_labels);
    X.uniq = reinterpret_cast<IndexSlot *>(a + o_uniq);
    X.table = reinterpret_cast<IndexSlot *>(a + o_table);
    X.bitmap = reinterpret_cast<uint32_t *>(a + o_bitmap);
    X.table_mask = (uint32_t)(cap - 1);
    X.table_shift = shift;
    R.idx = X;
    R.dev.table = X.table;
    R.dev.bitmap = X.bitmap;
    R.dev.labels = X.labels;
    R.dev.table_mask = X.table_mask;
    R.dev.table_shift = X.table_shift;
    R.dev_n_labels = nl;
    R.index.table_cap = (uint32_t)cap;
    R.index.table_mask = X.table_mask;
    R.index.table_shift = shift;
  }
  CUDA_TRY(cudaMemcpyAsync(ds + o_desc, hs + o_desc, n * sizeof(IdxRegion), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemsetAsync(d_nu, 0, n * 4, c->stream));
  if (total)
  {
    // 3. emit  4. segmented stable sort  5-6. group + table
    if (int rc = c->d_idx_keys.reserve((size_t)total * 8))
      return rc;
    if (int rc = c->d_idx_keys2.reserve((size_t)total * 8))
      return rc;
    if (int rc = c->d_idx_labels.reserve((size_t)total * sizeof(DevLabel)))
      return rc;
    if (int rc = c->d_idx_idx.reserve((size_t)total * 4))
      return rc;
    if (int rc = c->d_idx_idx2.reserve((size_t)total * 4))
      return rc;
    if (int rc = c->d_idx_head.reserve((size_t)total * 16)) // head flags + inclusive scan; before that: the sort's aux arrays
      return rc;
    size_t const sort_bytes = idx_sort_temp_bytes(total, n);
    scan_bytes = idx_scan_temp_bytes(total);
    if (int rc = c->d_idx_temp.reserve(std::max(sort_bytes, scan_bytes)))
      return rc;
    uint64_t * k1 = static_cast<uint64_t *>(c->d_idx_keys.p);
    uint64_t * k2 = static_cast<uint64_t *>(c->d_idx_keys2.p);
    uint32_t * i1 = static_cast<uint32_t *>(c->d_idx_idx.p);
    uint32_t * i2 = static_cast<uint32_t *>(c->d_idx_idx2.p);
    DevLabel * le = static_cast<DevLabel *>(c->d_idx_labels.p);
    uint32_t * head = static_cast<uint32_t *>(c->d_idx_head.p);
    idx_launch_emit(d_desc, n, d_rjo, total_jobs, d_joff, k1, le, i1, d_err, c->stream);
    if (idx_sort(c->d_idx_temp.p, c->d_idx_temp.cap, k1, k2, i1, i2, head, total, n, d_rto, c->stream))
      return fail(GTB_ERR_CUDA, "index build: sort failed");
    idx_launch_group(d_desc, k2, i2, le, head, head + total, c->d_idx_temp.p, c->d_idx_temp.cap, d_rto, n, total, d_nu, c->stream);
  }
  CUDA_TRY(cudaMemcpyAsync(hs + o_nu, d_nu, n * 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(wait_stream(c));
  CUDA_TRY(cudaGetLastError());
  const uint32_t * nu = reinterpret_cast<const uint32_t *>(hs + o_nu);
  for (uint32_t i = 0; i < n; ++i)
    regs[i]->dev_n_keys = nu[i];
  c->regions_dirty = true;
  return 0;
}

// Several regions at once.  With a CUDA device the index of all regions is built on the device in one launch sequence
// (gtb_set_index_build(ctx, 0) / GTB_INDEX_BUILD=host selects the host builder, whose builds run in parallel).
int gtb_region_begin_multi(gtb_ctx * ctx, int n, const int * region_ids, const gtb_graph_view * graphs)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || n <= 0 || !region_ids || !graphs)
    return fail(GTB_ERR_ARG, "bad arguments");
  for (int i = 0; i < n; ++i)
  {
    if (c->regions.count(region_ids[i]))
      return fail(GTB_ERR_STATE, "region id already in use");
    for (int j = 0; j < i; ++j)
      if (region_ids[j] == region_ids[i])
        return fail(GTB_ERR_ARG, "duplicate region id");
    if (int rc = validate_graph(&graphs[i]))
      return rc;
  }
  static int const env_mode = []() {
    const char * e = getenv("GTB_INDEX_BUILD");
    return !e ? -1 : (strcmp(e, "host") == 0 ? 0 : 1);
  }();
  bool const dev_index = c->device >= 0 && (env_mode >= 0 ? env_mode == 1 : c->index_build_device != 0);
  std::vector<std::unique_ptr<Region>> regs(n);
  std::vector<const char *> errs(n, nullptr);
  parallel_for(n, [&](int i)
               {
                 const gtb_graph_view * g = &graphs[i];
                 auto R = std::make_unique<Region>();
                 R->id = region_ids[i];
                 R->n_bubbles = g->n_ref - 1;
                 if (g->is_sv_graph)
                 {
                   // ReferenceDepth: offset = first ref node order, size = graph.reference.size() (reference_depth.cpp:17-27)
                   uint32_t const last = g->n_ref - 1;
                   uint32_t const last_reach = g->ref_order[last] + (uint32_t)(g->ref_seq_off[last + 1] - g->ref_seq_off[last]) - 1;
                   R->reference_offset = g->ref_order[0];
                   R->depth_size = last_reach >= g->ref_order[0] ? last_reach - g->ref_order[0] + 1 : 0;
                 }
                 if (!dev_index)
                 {
                   IndexBuilder ib(*g);
                   const char * err = nullptr;
                   if (!ib.build(R->index, &err))
                   {
                     errs[i] = err ? err : "index build failed";
                     return;
                   }
                 }
                 R->score_off.assign(1, 0);
                 R->cov_off.assign(1, 0);
                 for (uint32_t b = 0; b < R->n_bubbles; ++b)
                 {
                   uint32_t const v0 = g->ref_var_off[b];
                   uint32_t const cnum = g->ref_var_off[b + 1] - v0;
                   R->bubble_order.push_back(g->var_order[v0]);
                   R->n_alleles.push_back(cnum);
                   R->score_off.push_back(R->score_off.back() + cnum * (cnum + 1) / 2);
                   R->cov_off.push_back(R->cov_off.back() + cnum);
                 }
                 regs[i] = std::move(R);
               });
  for (int i = 0; i < n; ++i)
    if (errs[i])
      return fail(GTB_ERR_ARG, errs[i]);
  if (c->device >= 0)
  {
    cudaSetDevice(c->device);
    int rc = 0;
    for (int i = 0; i < n && !rc; ++i)
      rc = upload_region(c, *regs[i], &graphs[i], dev_index);
    if (!rc && dev_index)
      rc = build_indexes_on_device(c, regs);
    if (rc)
    {
      // undo: free slots and recycle the arenas of the regions that were already uploaded
      for (int i = 0; i < n; ++i)
      {
        if (regs[i]->slot >= 0)
          c->slot_region[regs[i]->slot] = -1;
        give_buffer(c, regs[i]->arena);
        give_buffer(c, regs[i]->index_arena);
      }
      c->regions_dirty = true;
      return rc;
    }
  }
  for (int i = 0; i < n; ++i)
    c->regions[region_ids[i]] = std::move(regs[i]);
  return 0;
}

int gtb_region_begin(gtb_ctx * ctx, int region_id, const gtb_graph_view * g)
{
  if (!ctx)
    return fail(GTB_ERR_ARG, "null ctx");
  return gtb_region_begin_multi(ctx, 1, &region_id, g);
}

// A region of another context on the same device, shared: graph, index and lookup tables stay the owner's (read-only from
// here on), this context adds its own pool state (accumulators, connection table).  One pool thread = one context is the
// reference's own threading model (paw::Station workers, src/typer/caller.cpp:272-391); with attached regions N pool threads
// cost N sets of accumulators, not N copies of every region's graph + 8 MiB k-mer table.
int gtb_region_attach(gtb_ctx * ctx, int region_id, gtb_ctx * owner, int owner_region_id)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  auto * o = reinterpret_cast<Ctx *>(owner);
  if (!c || !o || c == o)
    return fail(GTB_ERR_ARG, "gtb_region_attach needs two different contexts");
  if (c->device < 0 || c->device != o->device)
    return fail(GTB_ERR_ARG, "gtb_region_attach: both contexts must use the same CUDA device");
  if (c->regions.count(region_id))
    return fail(GTB_ERR_STATE, "region id already in use");
  auto it = o->regions.find(owner_region_id);
  if (it == o->regions.end())
    return fail(GTB_ERR_STATE, "unknown region in the owner context");
  Region const & S = *it->second;
  auto R = std::make_unique<Region>();
  R->id = region_id;
  R->borrowed = true;
  R->bubble_order = S.bubble_order;
  R->n_alleles = S.n_alleles;
  R->score_off = S.score_off;
  R->cov_off = S.cov_off;
  R->index = S.index; // host-built index (empty for a device-built one): gtb_index_size / gtb_index_export keep working
  R->n_bubbles = S.n_bubbles;
  R->depth_size = S.depth_size;
  R->reference_offset = S.reference_offset;
  R->dev_index = S.dev_index;
  R->dev_n_keys = S.dev_n_keys;
  R->dev_n_labels = S.dev_n_labels;
  R->idx = S.idx;
  R->dev = S.dev;
  // the pool state is this context's own
  R->dev.n_samples = 0;
  R->dev.log_score = R->dev.gt_cov = R->dev.max_log_score = R->dev.amb = R->dev.amb_alt = R->dev.alt_pp = nullptr;
  R->dev.vs_clipped_reads = R->dev.vs_mapq_squared = R->dev.pa_clipped_bp = R->dev.pa_mapq_squared = nullptr;
  R->dev.pa_score_diff = R->dev.pa_mismatches = nullptr;
  R->dev.read_strand = nullptr;
  R->dev.ref_depth_delta = nullptr;
  R->dev.conn_keys = nullptr;
  R->dev.conn_vals = R->dev.conn_state = nullptr;
  R->dev.conn_mask = 0;
  cudaSetDevice(c->device);
  CUDA_TRY(cudaStreamSynchronize(o->stream)); // the owner's upload + index build are complete
  int slot = -1;
  for (size_t k = 0; k < c->slot_region.size(); ++k)
    if (c->slot_region[k] < 0)
    {
      slot = (int)k;
      break;
    }
  if (slot < 0)
  {
    slot = (int)c->slot_region.size();
    c->slot_region.push_back(-1);
  }
  if (slot > 0xFFFF)
    return fail(GTB_ERR_CAPACITY, "more than 65536 resident regions");
  c->slot_region[slot] = region_id;
  R->slot = slot;
  c->regions_dirty = true;
  c->regions[region_id] = std::move(R);
  return 0;
}

int gtb_region_end(gtb_ctx * ctx, int region_id)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  auto it = c->regions.find(region_id);
  if (it == c->regions.end())
    return fail(GTB_ERR_STATE, "unknown region");
  if (c->device >= 0)
  {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    give_buffer(c, it->second->arena);
    give_buffer(c, it->second->index_arena);
    give_buffer(c, it->second->accum);
    give_buffer(c, it->second->conn);
    if (it->second->slot >= 0)
      c->slot_region[it->second->slot] = -1;
    c->regions_dirty = true;
  }
  c->regions.erase(it);
  c->have_last = false;
  return 0;
}

int gtb_index_size(gtb_ctx * ctx, int region_id, uint64_t * n_keys, uint64_t * n_labels)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  auto it = c->regions.find(region_id);
  if (it == c->regions.end())
    return fail(GTB_ERR_STATE, "unknown region");
  if (it->second->dev_index)
  {
    *n_keys = it->second->dev_n_keys;
    *n_labels = it->second->dev_n_labels;
    return 0;
  }
  *n_keys = it->second->index.n_keys;
  *n_labels = it->second->index.labels.size();
  return 0;
}

int gtb_index_export(gtb_ctx * ctx, int region_id, uint64_t * keys, uint32_t * label_off, gtb_label * labels)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  auto it = c->regions.find(region_id);
  if (it == c->regions.end())
    return fail(GTB_ERR_STATE, "unknown region");
  if (it->second->dev_index)
  {
    // the device-built index is already grouped by ascending k-mer
    Region & R = *it->second;
    cudaSetDevice(c->device);
    std::vector<IndexSlot> u(R.dev_n_keys);
    CUDA_TRY(cudaMemcpy(u.data(), R.idx.uniq, u.size() * sizeof(IndexSlot), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(labels, R.idx.labels, R.dev_n_labels * sizeof(gtb_label), cudaMemcpyDeviceToHost));
    label_off[0] = 0;
    for (size_t i = 0; i < u.size(); ++i)
    {
      keys[i] = u[i].key;
      if (u[i].off != label_off[i])
        return fail(GTB_ERR_STATE, "device index: distinct k-mer slots are not contiguous");
      label_off[i + 1] = u[i].off + u[i].cnt;
    }
    return 0;
  }
  std::vector<uint64_t> k;
  std::vector<uint32_t> lo;
  std::vector<gtb_label> ll;
  it->second->index.export_sorted(k, lo, ll);
  memcpy(keys, k.data(), k.size() * 8);
  memcpy(label_off, lo.data(), lo.size() * 4);
  memcpy(labels, ll.data(), ll.size() * sizeof(gtb_label));
  return 0;
}

int gtb_pool_begin(gtb_ctx * ctx, int region_id, int n_samples)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  auto it = c->regions.find(region_id);
  if (it == c->regions.end())
    return fail(GTB_ERR_STATE, "unknown region");
  if (c->device < 0)
    return fail(GTB_ERR_CUDA, "host-only context: the genotyping path needs a CUDA device (no CPU fallback)");
  if (n_samples <= 0)
    return fail(GTB_ERR_ARG, "n_samples must be positive");
  Region & R = *it->second;
  cudaSetDevice(c->device);
  size_t const NB = R.n_bubbles, NS = (size_t)n_samples;
  size_t const n_scores = R.score_off.back(), n_cov = R.cov_off.back();
  size_t off = 0;
  size_t const o_ls = place<uint32_t>(off, n_scores * NS);
  size_t const o_gc = place<uint32_t>(off, n_cov * NS);
  size_t const o_ml = place<uint32_t>(off, NB * NS);
  size_t const o_amb = place<uint32_t>(off, NB * NS);
  size_t const o_amba = place<uint32_t>(off, NB * NS);
  size_t const o_altpp = place<uint32_t>(off, NB * NS);
  size_t const o_rdd = place<int>(off, R.depth_size ? ((size_t)R.depth_size + 1) * NS : 0);
  size_t const o_vcr = place<unsigned long long>(off, NB);
  size_t const o_vmq = place<unsigned long long>(off, NB);
  size_t const o_pcb = place<unsigned long long>(off, n_cov);
  size_t const o_pmq = place<unsigned long long>(off, n_cov);
  size_t const o_psd = place<unsigned long long>(off, n_cov);
  size_t const o_pmm = place<unsigned long long>(off, n_cov);
  size_t const o_rs = place<uint32_t>(off, n_cov * 4);
  size_t const total = align_up(off, 256);
  if (int rc = take_buffer(c, R.accum, total))
    return rc;
  CUDA_TRY(cudaMemsetAsync(R.accum.p, 0, total, c->stream));
  R.accum_bytes = total;
  uint8_t * d = static_cast<uint8_t *>(R.accum.p);
  DevRegion & D = R.dev;
  D.n_samples = (uint32_t)n_samples;
  D.log_score = reinterpret_cast<uint32_t *>(d + o_ls);
  D.gt_cov = reinterpret_cast<uint32_t *>(d + o_gc);
  D.max_log_score = reinterpret_cast<uint32_t *>(d + o_ml);
  D.amb = reinterpret_cast<uint32_t *>(d + o_amb);
  D.amb_alt = reinterpret_cast<uint32_t *>(d + o_amba);
  D.alt_pp = reinterpret_cast<uint32_t *>(d + o_altpp);
  D.ref_depth_delta = reinterpret_cast<int *>(d + o_rdd);
  D.vs_clipped_reads = reinterpret_cast<unsigned long long *>(d + o_vcr);
  D.vs_mapq_squared = reinterpret_cast<unsigned long long *>(d + o_vmq);
  D.pa_clipped_bp = reinterpret_cast<unsigned long long *>(d + o_pcb);
  D.pa_mapq_squared = reinterpret_cast<unsigned long long *>(d + o_pmq);
  D.pa_score_diff = reinterpret_cast<unsigned long long *>(d + o_psd);
  D.pa_mismatches = reinterpret_cast<unsigned long long *>(d + o_pmm);
  D.read_strand = reinterpret_cast<uint32_t *>(d + o_rs);
  R.n_samples = n_samples;
  R.pool_open = true;
  R.poisoned = R.reduced = false;
  c->regions_dirty = true;
  D.conn_keys = nullptr;
  D.conn_vals = nullptr;
  D.conn_state = nullptr;
  D.conn_mask = 0;
  R.conn_cap = 0;
  R.conn_used = 0;
  if (c->connections)
  {
    if (NB > CONN_MAX_BUBBLES || NS > CONN_MAX_SAMPLES)
      return fail(GTB_ERR_CAPACITY, "phasing connections need <= 65536 bubbles per region (the reference's uint16 haplotype "
                                    "index, vcf_writer.cpp:502) and <= 2^21 samples per pool");
    if (int rc = conn_alloc(c, R, 1u << 16))
      return rc;
  }
  return 0;
}

// Front half of one chunk on the chunk's own stream, as soon as its H2D copy has landed: probe + chain.  The chunks of a
// submit run concurrently -- chain_kernel is bound by the latency of its longest task (~0.25 ms whatever the chunk size),
// so a chunk's chain overlaps the next chunk's copy and probe instead of queueing behind them.
static int launch_front(Ctx * c, BatchState & B, cudaEvent_t after)
{
  LaunchParams & P = B.P;
  cudaStream_t const s = B.stream;
  CUDA_TRY(cudaStreamWaitEvent(s, after, 0)); // everything queued on the main stream before this submit
  CUDA_TRY(cudaStreamWaitEvent(s, B.ev[1], 0));
  CUDA_TRY(cudaMemsetAsync(B.d_counters.p, 0, sizeof(DevCounters), s));
  // orientations that are not aligned keep an all-zero summary (no paths)
  CUDA_TRY(cudaMemsetAsync(P.summaries, 0, (size_t)P.batch.n_units * 2 * (sizeof(TaskSummary) + 1), s)); // + pending[]
  if (P.tap.list_count)
    CUDA_TRY(cudaMemsetAsync(P.tap.list_count, 0, (size_t)P.batch.n_units * 2 * (NLISTS * 2 + 1) * 4, s));
  CUDA_TRY(cudaEventRecord(B.ev[2], s));
  if (B.bam.n) // record parsing: the chunk's columns from raw htslib records
    if (launch_bam_parse(B.bam, B.d_bam_sort.p, B.bam_sort_bytes, s) != 0)
      return fail(GTB_ERR_CUDA, "radix sort of the read-name hashes failed");
  // batch preparation: units, aligned orientations, link checks (exclusive scan of (is_unit << 32 | orientations))
  launch_prep_flags(B.prep, s);
  if (B.prep.n_records)
    if (exclusive_scan64(B.d_scan_temp.p, B.scan_temp_bytes, B.prep.scan, B.prep.scan, B.prep.n_records, s) != 0)
      return fail(GTB_ERR_CUDA, "prefix scan of the batch preparation failed");
  launch_prep_fill(B.prep, s);
  CUDA_TRY(cudaEventRecord(B.ev[6], s));
  std::unique_lock<std::mutex> front_lock;
  DeviceFifo * fifo = nullptr;
  if (fifo_mode() & 1)
  {
    fifo = &fifo_of(c->device);
    front_lock = std::unique_lock<std::mutex>(fifo->m_front);
    CUDA_TRY(fifo_wait(fifo->front, s)); // the front issued before this one, whichever context issued it
  }
  launch_probe(P, s);
  CUDA_TRY(cudaEventRecord(B.ev[3], s));
  // chain_kernel finishes ~95 % of the read orientations itself.  The rest runs beside the first score pass, on side
  // streams: chain_general_kernel over the queue of irregular reads, and slow_kernel over the tasks probe_kernel marked
  // (IUPAC / N seeds); both are long single-thread jobs that leave most issue slots free.  Records with a task in one of the
  // queues (pending[], written by chain_kernel only) wait for the second score pass (launch_back).
  P.defer = 1u;
  launch_chain(P, s);
  CUDA_TRY(cudaEventRecord(B.ev[8], s));
  if (fifo)
  {
    CUDA_TRY(cudaEventRecord(fifo->front, s));
    front_lock.unlock();
  }
  CUDA_TRY(cudaStreamWaitEvent(B.stream_gen, B.ev[8], 0));
  CUDA_TRY(cudaEventRecord(B.ev[9], B.stream_gen));
  launch_chain_general(P, B.stream_gen);
  CUDA_TRY(cudaEventRecord(B.ev[10], B.stream_gen));
  CUDA_TRY(cudaStreamWaitEvent(B.stream_slow, B.ev[8], 0));
  CUDA_TRY(cudaEventRecord(B.ev[11], B.stream_slow));
  {
    MultiLaunch M;
    M.n = 1;
    M.p[0] = P;
    launch_slow(M, 0, B.stream_slow);
  }
  CUDA_TRY(cudaEventRecord(B.ev[12], B.stream_slow));
  launch_score(P, B.with_conn, s);
  CUDA_TRY(cudaEventRecord(B.ev[5], s));
  CUDA_TRY(cudaEventRecord(B.ev[4], s));
  return 0;
}

// Back half of all chunks on the main stream: ONE slow_kernel + huge_kernel launch over every chunk's queue (a launch costs
// the latency of its slowest single-lane task, ~0.1 ms, however many chunks feed it), then score + counters per chunk.
static int launch_back(Ctx * c, int n_chunks)
{
  cudaStream_t const ts = c->stream;
  MultiLaunch M;
  M.n = n_chunks;
  bool with_conn[MAX_CHUNKS];
  for (int k = 0; k < n_chunks; ++k)
  {
    // the chunk's first score pass, its chain_general_kernel and its first slow_kernel launch
    CUDA_TRY(cudaStreamWaitEvent(ts, c->bs[k].ev[5], 0));
    CUDA_TRY(cudaStreamWaitEvent(ts, c->bs[k].ev[10], 0));
    CUDA_TRY(cudaStreamWaitEvent(ts, c->bs[k].ev[12], 0));
    M.p[k] = c->bs[k].P;
    with_conn[k] = c->bs[k].with_conn;
  }
  CUDA_TRY(cudaEventRecord(c->ev_slow[0], ts));
  launch_slow(M, 1, ts); // what chain_general_kernel re-queued (normally nothing), then huge_kernel
  CUDA_TRY(cudaEventRecord(c->ev_slow[1], ts));
  CUDA_TRY(cudaEventRecord(c->ev_slow[4], ts));
  launch_score_deferred(M, with_conn, ts); // second score pass: the records that waited for the slower tiers
  CUDA_TRY(cudaEventRecord(c->ev_slow[3], ts));
  for (int k = 0; k < n_chunks; ++k)
  {
    BatchState & B = c->bs[k];
    CUDA_TRY(cudaMemcpyAsync(B.h_counters.p, B.d_counters.p, sizeof(DevCounters), cudaMemcpyDeviceToHost, ts));
    CUDA_TRY(cudaEventRecord(B.ev[7], ts));
  }
  return 0;
}

// After the stream has been synchronised: timings, counters, errors of all chunks of the last submit/replay.
static int collect_chunks(Ctx * c, gtb_submit_stats * stats, bool record_h2d)
{
  CUDA_TRY(wait_stream(c));
  CUDA_TRY(cudaGetLastError());
  if (c->debug && c->n_chunks_last == 1 && c->bs[0].unit_begin.empty())
  {
    // debug taps address units per region: first unit of region i = unit of its first record
    BatchState & B = c->bs[0];
    size_t const nr = B.rec_begin.size() - 1;
    std::vector<int32_t> unit(B.P.batch.n_records);
    if (!unit.empty())
      CUDA_TRY(cudaMemcpy(unit.data(), B.P.batch.unit, unit.size() * 4, cudaMemcpyDeviceToHost));
    DevCounters kc0{};
    CUDA_TRY(cudaMemcpy(&kc0, B.d_counters.p, sizeof(kc0), cudaMemcpyDeviceToHost));
    B.unit_begin.assign(nr + 1, kc0.n_units);
    for (size_t i = nr; i-- > 0;)
      B.unit_begin[i] = B.rec_begin[i] < B.rec_begin[i + 1] ? (uint32_t)unit[B.rec_begin[i]] : B.unit_begin[i + 1];
  }
  if (record_h2d)
    c->t_h2d = 0;
  c->t_align = c->t_prep = c->t_probe = c->t_chain = c->t_chain_fast = c->t_slow = c->t_score = c->t_d2h = 0;
  c->last_n_slow = c->last_n_gen = c->last_n_active = 0;
  for (auto & q : c->t0_reasons)
    q = 0;
  auto kc_of = [](BatchState & B) { return static_cast<DevCounters const *>(B.h_counters.p); };
  gtb_submit_stats st{};
  unsigned long long n_overflow = 0, n_input_error = 0, reasons[12] = {0};
  uint32_t input_bits = 0;
  for (int k = 0; k < c->n_chunks_last; ++k)
  {
    BatchState & B = c->bs[k];
    float t = 0;
    if (record_h2d)
    {
      cudaEventElapsedTime(&t, B.ev[0], B.ev[1]);
      c->t_h2d += t;
    }
    cudaEventElapsedTime(&t, B.ev[2], B.ev[6]);
    c->t_prep += t;
    cudaEventElapsedTime(&t, B.ev[6], B.ev[3]);
    c->t_probe += t;
    cudaEventElapsedTime(&t, B.ev[3], B.ev[8]);
    c->t_chain_fast += t;
    c->t_chain += t;
    cudaEventElapsedTime(&t, B.ev[9], B.ev[10]);
    c->t_chain += t; // (chain_general_kernel runs beside the first score pass)
    cudaEventElapsedTime(&t, B.ev[11], B.ev[12]);
    c->t_slow += t;
    cudaEventElapsedTime(&t, B.ev[8], B.ev[5]);
    c->t_score += t;
    c->last_n_gen += kc_of(B)->n_gen;
    for (int q = 0; q < 16; ++q)
      c->t0_reasons[q] += kc_of(B)->t0_reasons[q];
    DevCounters const * kc = static_cast<DevCounters *>(B.h_counters.p);
    if (B.P.task_times && k == 0)
      if (const char * fn = getenv("GTB_TASK_TIMES"))
      {
        std::vector<unsigned long long> tt((size_t)kc->n_gen * 2); // tasks of chain_general_kernel, in queue order
        std::vector<uint32_t> at(kc->n_gen);
        cudaMemcpy(tt.data(), B.P.task_times, tt.size() * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(at.data(), B.P.gen_tasks, at.size() * 4, cudaMemcpyDeviceToHost);
        if (FILE * f = fopen(fn, "wb"))
        {
          uint64_t const n = kc->n_gen;
          fwrite(&n, 8, 1, f);
          fwrite(tt.data(), 8, tt.size(), f);
          fwrite(at.data(), 4, at.size(), f);
          fclose(f);
        }
      }
    c->last_n_slow += kc->n_slow + kc->n_slow2;
    c->last_n_active += kc->n_active;
    st.n_records += B.P.batch.n_records;
    st.n_alignments += kc->n_units;
    st.n_oriented += kc->n_active;
    input_bits |= kc->input_bits;
    st.n_pairs_scored += kc->n_pairs_scored;
    st.n_singles_scored += kc->n_singles_scored;
    st.n_capacity_overflow += kc->n_overflow;
    // prep_flags, scan, prep_fill, probe, chain, chain_general, slow (first launch), first score pass
    st.kernel_launches += B.P.batch.n_records ? 8 : 0;
    st.kernel_launches += B.bam.n ? 7 : 0; // parse, seq, dup, radix sort (3 kernels at these sizes), mate
    st.kernel_launches += record_h2d ? B.gather_launches : 0;
    n_overflow += kc->n_overflow;
    n_input_error += kc->n_input_error;
    for (int q = 0; q < 12; ++q)
      reasons[q] += kc->final_reasons[q];
  }
  {
    float t = 0;
    cudaEventElapsedTime(&t, c->ev_slow[0], c->ev_slow[1]);
    c->t_slow += t;
    c->t_score0 = c->t_score;
    cudaEventElapsedTime(&t, c->ev_slow[4], c->ev_slow[3]);
    c->t_score += t; // second pass
    c->t_score1 = t;
    cudaEventElapsedTime(&t, c->ev_slow[3], c->bs[c->n_chunks_last - 1].ev[7]);
    c->t_d2h = t;
    cudaEventElapsedTime(&t, c->bs[0].ev[2], c->ev_slow[3]);
    c->t_total = t; // device span of the whole launch sequence (slow_kernel overlaps the first score pass)
    st.kernel_launches += st.n_records ? 3 : 0; // slow_kernel, huge_kernel, second score pass: once per submit
  }
  c->t_align = c->t_prep + c->t_probe + c->t_chain_fast; // everything later overlaps: reported as one span (gtb_last_timing)
  if (stats)
    *stats = st;
  // input the batch-preparation kernels rejected (the offending links were cut / values replaced before the other kernels ran)
  if (input_bits & PREP_ERR_SAMPLE)
    return fail(GTB_ERR_ARG, "sample index out of range");
  if (input_bits & PREP_ERR_DUP)
    return fail(GTB_ERR_ARG, "dup_of must reference an earlier record of the same batch");
  if (input_bits & PREP_ERR_MATE)
    return fail(GTB_ERR_ARG, "mate must reference an earlier record of the same batch");
  if (input_bits & PREP_ERR_LEN)
    return fail(GTB_ERR_CAPACITY, "read longer than 152 bases (reference MAX_READ_LENGTH is 151)");
  if (input_bits & PREP_ERR_RECORD)
    return fail(GTB_ERR_ARG, "malformed record: qname + cigar + seq + qual do not fit its data block");
  if (n_input_error)
    return fail(GTB_ERR_INPUT, "two mates with the same IS_FIRST_IN_PAIR flag (the reference aborts here, "
                               "hts_parallel_reader.cpp:306-315)");
  if (n_overflow)
  {
    static const char * names[12] = {"refs", "vars", "paths", "locs", "labels", "cand_vars", "cands", "keys", "tap",
                                     "pool", "read_len", "-"};
    std::string why;
    for (int q = 0; q < 12; ++q)
      if (reasons[q])
        why += std::string(" ") + names[q] + "=" + std::to_string(reasons[q]);
    return fail(GTB_ERR_CAPACITY, std::to_string(n_overflow) +
                                    " read(s) exceeded a device working-set capacity (gtb_device.cuh):" + why);
  }
  return 0;
}

// Device layout of one chunk's record columns: [seq4] [columns that arrive from the host or from the record parser ...]
// [columns produced by the batch-preparation kernels ...]
struct ChunkLayout
{
  size_t o_seq4, o_lseq, o_flag, o_region, o_mapq, o_same, o_sd, o_clip, o_left, o_isize, o_sample, o_mate, o_dup, copy_end;
  size_t o_unit, o_urec, o_active, o_scan, bytes;
  explicit ChunkLayout(size_t total)
  {
    size_t off = 0;
    o_seq4 = place<uint8_t>(off, total * GTB_SEQ_STRIDE);
    o_lseq = place<uint16_t>(off, total);
    o_flag = place<uint16_t>(off, total);
    o_region = place<uint16_t>(off, total);
    o_mapq = place<uint8_t>(off, total);
    o_same = place<uint8_t>(off, total);
    o_sd = place<uint8_t>(off, total);
    o_clip = place<uint8_t>(off, total);
    o_left = place<uint8_t>(off, total);
    o_isize = place<int32_t>(off, total);
    o_sample = place<int32_t>(off, total);
    o_mate = place<int32_t>(off, total);
    o_dup = place<int32_t>(off, total);
    copy_end = off;
    o_unit = place<int32_t>(off, total);
    o_urec = place<int32_t>(off, total);
    o_active = place<uint32_t>(off, total * 2);
    o_scan = place<unsigned long long>(off, total);
    bytes = align_up(off, 256);
  }
};
static int bind_chunk(Ctx * c, BatchState & B, ChunkLayout const & Lo, size_t total, bool with_tap);

// Stages one chunk (regions [0, n) of the given arrays): the bases are DMA-ed straight from the caller's buffers when those
// are page-locked, the small columns are gathered into B's pinned buffer by the host pool (plain copies in blocks of 8192
// records; mate / duplicate links are rebased to chunk-global indices on the way) and follow in ONE H2D copy.  Everything
// that needs a look at each record -- alignment units, the list of aligned orientations, link validation -- happens on the
// device (prep_flags_kernel -> scan -> prep_fill_kernel, launch_front).
static int stage_chunk(Ctx * c, BatchState & B, int n, const int * region_ids, const gtb_read_batch * batches,
                       Region * const * regs, bool with_tap)
{
  size_t total = 0;
  for (int i = 0; i < n; ++i)
    total += batches[i].n_reads;
  ChunkLayout const Lo(total);
  size_t const o_seq4 = Lo.o_seq4, o_lseq = Lo.o_lseq, o_flag = Lo.o_flag, o_region = Lo.o_region, o_mapq = Lo.o_mapq,
               o_same = Lo.o_same, o_sd = Lo.o_sd, o_clip = Lo.o_clip, o_left = Lo.o_left, o_isize = Lo.o_isize,
               o_sample = Lo.o_sample, o_mate = Lo.o_mate, o_dup = Lo.o_dup, copy_end = Lo.copy_end, bytes = Lo.bytes;
  if (int rc = B.h_batch.reserve(align_up(copy_end, 256)))
    return rc;
  uint8_t * h = static_cast<uint8_t *>(B.h_batch.p);
  B.regions.assign(region_ids, region_ids + n);
  std::vector<size_t> rec_base(n + 1, 0);
  for (int i = 0; i < n; ++i)
    rec_base[i + 1] = rec_base[i] + batches[i].n_reads;
  B.rec_begin.assign(rec_base.begin(), rec_base.end());
  B.unit_begin.clear(); // filled after the kernels when the debug taps are on
  if (int rc = B.d_batch.reserve(bytes))
    return rc;
  // Bases are 3/4 of the bytes.  When the caller keeps them in pinned (page-locked) host memory -- gtb_host_alloc or
  // cudaHostRegister -- they are DMA-ed straight from the caller's buffers while the small columns are staged.
  bool direct_seq = n > 0;
  for (int i = 0; i < n && direct_seq; ++i)
    if (batches[i].n_reads)
    {
      cudaPointerAttributes at{};
      if (cudaPointerGetAttributes(&at, batches[i].seq4) != cudaSuccess || at.type != cudaMemoryTypeHost)
      {
        cudaGetLastError();
        direct_seq = false;
      }
    }
  // The small columns (27 bytes per record) as well: when every one of them is page-locked and mapped, the device can read
  // them straight from the caller's arrays (gather_columns_kernel) and the host touches no record at all.  Measured on one
  // GPU with the host to itself (4 pool threads, 2e5-record steps): staged 0.92-0.96 ms per step, zero copy 0.95-1.0 ms --
  // the staging threads have idle cores to run on and the gather kernel occupies a few SMs.  With several ranks on one host
  // the cores are what is short (8 ranks x 3 pool threads on 32 cores), so zero copy is the default exactly then
  // (LOCAL_WORLD_SIZE > 1); GTB_ZERO_COPY=0|1 overrides.
  bool const zero_copy_on = []() {
    if (const char * e = getenv("GTB_ZERO_COPY"))
      return atoi(e) != 0;
    const char * lw = getenv("LOCAL_WORLD_SIZE");
    return lw && atoi(lw) > 1;
  }();
  bool direct_cols = direct_seq && zero_copy_on;
  std::vector<ColumnJob> jobs;
  if (direct_cols)
  {
    jobs.reserve(n);
    auto dev_alias = [&](const void * p, bool optional) -> const void * {
      if (!p)
      {
        direct_cols = direct_cols && optional;
        return nullptr;
      }
      cudaPointerAttributes at{};
      if (cudaPointerGetAttributes(&at, p) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer)
      {
        cudaGetLastError();
        direct_cols = false;
        return nullptr;
      }
      return at.devicePointer;
    };
    for (int i = 0; i < n && direct_cols; ++i)
    {
      gtb_read_batch const & b = batches[i];
      if (!b.n_reads)
        continue;
      ColumnJob J{};
      J.lseq = static_cast<const uint16_t *>(dev_alias(b.lseq, false));
      J.flag = static_cast<const uint16_t *>(dev_alias(b.flag, false));
      J.mapq = static_cast<const uint8_t *>(dev_alias(b.mapq, false));
      J.same_tid = static_cast<const uint8_t *>(dev_alias(b.same_tid, false));
      J.score_diff = static_cast<const uint8_t *>(dev_alias(b.score_diff, false));
      J.clipped = static_cast<const uint8_t *>(dev_alias(b.clipped, true));
      J.leftover = static_cast<const uint8_t *>(dev_alias(b.leftover, true));
      J.isize = static_cast<const int32_t *>(dev_alias(b.isize, false));
      J.sample = static_cast<const int32_t *>(dev_alias(b.sample, false));
      J.mate = static_cast<const int32_t *>(dev_alias(b.mate, true));
      J.dup_of = static_cast<const int32_t *>(dev_alias(b.dup_of, true));
      J.n = b.n_reads;
      J.rec_base = (uint32_t)rec_base[i];
      J.slot = (uint32_t)regs[i]->slot;
      jobs.push_back(J);
    }
  }
  std::unique_lock<std::mutex> copy_lock;
  DeviceFifo * fifo = nullptr;
  if (direct_cols && (fifo_mode() & 2))
  {
    // nothing but queueing between here and the record below: the transfers of this submit go behind those of the submit
    // issued before it instead of sharing the link with them copy by copy
    fifo = &fifo_of(c->device);
    copy_lock = std::unique_lock<std::mutex>(fifo->m_copy);
    CUDA_TRY(fifo_wait(fifo->copy, c->copy_stream));
  }
  CUDA_TRY(cudaEventRecord(B.ev[0], c->copy_stream));
  B.gather_launches = 0;
  if (direct_cols)
  {
    // before the bases: the device can start on the columns while the copy engine is still being programmed
    for (size_t j0 = 0; j0 < jobs.size(); j0 += GATHER_JOBS_PER_LAUNCH)
    {
      ColumnGather G{};
      G.dst = static_cast<uint8_t *>(B.d_batch.p);
      G.o_lseq = o_lseq, G.o_flag = o_flag, G.o_region = o_region, G.o_mapq = o_mapq, G.o_same = o_same, G.o_sd = o_sd;
      G.o_clip = o_clip, G.o_left = o_left, G.o_isize = o_isize, G.o_sample = o_sample, G.o_mate = o_mate, G.o_dup = o_dup;
      G.n_jobs = (uint32_t)std::min<size_t>(GATHER_JOBS_PER_LAUNCH, jobs.size() - j0);
      uint32_t tiles = 0;
      for (uint32_t j = 0; j < G.n_jobs; ++j)
      {
        G.job[j] = jobs[j0 + j];
        G.job[j].tile_begin = tiles;
        tiles += (G.job[j].n + GATHER_TILE - 1) / GATHER_TILE;
      }
      G.n_tiles = tiles;
      launch_gather_columns(G, c->copy_stream);
      ++B.gather_launches;
    }
    CUDA_TRY(cudaGetLastError());
  }
  static int const n_copy_streams = []() { const char * e = getenv("GTB_COPY_STREAMS"); return e ? atoi(e) : 2; }();
  if (direct_seq)
  {
    bool used2 = false;
    if (n_copy_streams > 1)
    {
      CUDA_TRY(cudaEventRecord(c->ev_copy2, c->copy_stream)); // keep the two queues in step chunk by chunk
      CUDA_TRY(cudaStreamWaitEvent(c->copy_stream2, c->ev_copy2, 0));
    }
    for (int i = 0; i < n; ++i)
      if (batches[i].n_reads)
      {
        bool const second = n_copy_streams > 1 && (i & 1);
        used2 = used2 || second;
        CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(B.d_batch.p) + o_seq4 + rec_base[i] * GTB_SEQ_STRIDE, batches[i].seq4,
                                 (size_t)batches[i].n_reads * GTB_SEQ_STRIDE, cudaMemcpyHostToDevice,
                                 second ? c->copy_stream2 : c->copy_stream));
      }
    if (used2)
    {
      CUDA_TRY(cudaEventRecord(c->ev_copy2, c->copy_stream2));
      CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_copy2, 0));
    }
  }
  if (fifo)
  {
    CUDA_TRY(cudaEventRecord(fifo->copy, c->copy_stream));
    copy_lock.unlock();
  }
  // ---- gather the small columns: independent blocks of records, so the pool is busy whatever the number of regions
  struct Block
  {
    int region;
    uint32_t k0, k1;
  };
  std::vector<Block> blocks;
  constexpr uint32_t BLOCK = 8192;
  for (int i = 0; i < n && !direct_cols; ++i)
    for (uint32_t k0 = 0; k0 < batches[i].n_reads; k0 += BLOCK)
      blocks.push_back({i, k0, std::min<uint32_t>(k0 + BLOCK, batches[i].n_reads)});
  parallel_for((int)blocks.size(), [&](int bi)
               {
                 Block const & blk = blocks[bi];
                 gtb_read_batch const & b = batches[blk.region];
                 size_t const k0 = blk.k0, m = blk.k1 - blk.k0, base = rec_base[blk.region], at = base + k0;
                 if (!direct_seq)
                   memcpy(h + o_seq4 + at * GTB_SEQ_STRIDE, b.seq4 + k0 * GTB_SEQ_STRIDE, m * GTB_SEQ_STRIDE);
                 memcpy(h + o_lseq + at * 2, b.lseq + k0, m * 2);
                 memcpy(h + o_flag + at * 2, b.flag + k0, m * 2);
                 memcpy(h + o_mapq + at, b.mapq + k0, m);
                 memcpy(h + o_same + at, b.same_tid + k0, m);
                 memcpy(h + o_sd + at, b.score_diff + k0, m);
                 if (b.clipped)
                   memcpy(h + o_clip + at, b.clipped + k0, m);
                 else
                   memset(h + o_clip + at, 0, m);
                 if (b.leftover)
                   memcpy(h + o_left + at, b.leftover + k0, m);
                 else
                   memset(h + o_left + at, 0, m);
                 memcpy(h + o_isize + at * 4, b.isize + k0, m * 4);
                 memcpy(h + o_sample + at * 4, b.sample + k0, m * 4);
                 uint16_t * h_region = reinterpret_cast<uint16_t *>(h + o_region) + at;
                 int32_t * h_mate = reinterpret_cast<int32_t *>(h + o_mate) + at;
                 int32_t * h_dup = reinterpret_cast<int32_t *>(h + o_dup) + at;
                 uint16_t const slot = (uint16_t)regs[blk.region]->slot;
                 int32_t const rb = (int32_t)base;
                 for (size_t k = 0; k < m; ++k)
                   h_region[k] = slot;
                 // links are batch-local; a link at or beyond its own record stays >= its chunk-global index and is
                 // reported by prep_flags_kernel
                 if (b.mate)
                   for (size_t k = 0; k < m; ++k)
                     h_mate[k] = b.mate[k0 + k] < 0 ? -1 : rb + b.mate[k0 + k];
                 else
                   for (size_t k = 0; k < m; ++k)
                     h_mate[k] = -1;
                 if (b.dup_of)
                   for (size_t k = 0; k < m; ++k)
                     h_dup[k] = b.dup_of[k0 + k] < 0 ? -1 : rb + b.dup_of[k0 + k];
                 else
                   for (size_t k = 0; k < m; ++k)
                     h_dup[k] = -1;
               });

  size_t const copy_begin = direct_seq ? o_lseq : 0;
  if (copy_end > copy_begin && !direct_cols)
    CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(B.d_batch.p) + copy_begin, h + copy_begin, copy_end - copy_begin,
                             cudaMemcpyHostToDevice, c->copy_stream));
  CUDA_TRY(cudaEventRecord(B.ev[1], c->copy_stream));

  B.bam.n = 0;
  return bind_chunk(c, B, Lo, total, with_tap);
}

// Reserves the per-chunk device buffers and points B.P / B.prep at the chunk's columns.
static int bind_chunk(Ctx * c, BatchState & B, ChunkLayout const & Lo, size_t total, bool with_tap)
{
  size_t const o_seq4 = Lo.o_seq4, o_lseq = Lo.o_lseq, o_flag = Lo.o_flag, o_region = Lo.o_region, o_mapq = Lo.o_mapq,
               o_same = Lo.o_same, o_sd = Lo.o_sd, o_clip = Lo.o_clip, o_left = Lo.o_left, o_isize = Lo.o_isize,
               o_sample = Lo.o_sample, o_mate = Lo.o_mate, o_dup = Lo.o_dup, o_unit = Lo.o_unit, o_urec = Lo.o_urec,
               o_active = Lo.o_active, o_scan = Lo.o_scan;
  // upper bounds: every record its own unit, both orientations aligned; the exact counts are produced on the device
  uint32_t const n_units = (uint32_t)total;
  uint32_t const n_active = (uint32_t)total * 2;
  uint32_t const n_tasks = n_units * 2;
  if (int rc = B.d_seedrecs.reserve(align_up((size_t)n_active * SEED_REC_BYTES + 64) + (size_t)n_active * LAB_REC_BYTES))
    return rc;
  // queues: slow_tasks | huge_tasks | deferred records | gen_tasks | slow2_tasks
  if (int rc = B.d_slow.reserve((size_t)n_active * 16 + 512 + total * 4))
    return rc;
  if (int rc = B.d_summaries.reserve((size_t)n_tasks * (sizeof(TaskSummary) + 1) + 16))
    return rc;
  size_t const pool_words = (size_t)n_tasks * INLINE_WORDS + (size_t)n_units * 64 + 65536;
  if (int rc = B.d_pool.reserve(pool_words * 4))
    return rc;
  if (int rc = B.d_counters.reserve(sizeof(DevCounters)))
    return rc;
  if (int rc = c->d_huge.reserve(huge_state_bytes()))
    return rc;
  if (int rc = c->d_spill.reserve(align_spill_bytes()))
    return rc;
  if (int rc = B.h_counters.reserve(sizeof(DevCounters)))
    return rc;
  B.scan_temp_bytes = scan64_temp_bytes((uint32_t)std::max<size_t>(total, 1));
  if (int rc = B.d_scan_temp.reserve(B.scan_temp_bytes + 16))
    return rc;

  LaunchParams & P = B.P;
  memset(&P, 0, sizeof(P));
  uint8_t * d = static_cast<uint8_t *>(B.d_batch.p);
  P.regions = static_cast<const DevRegion *>(c->d_regions.p);
  P.batch.n_records = (uint32_t)total;
  P.batch.n_units = n_units;
  P.batch.seq4 = d + o_seq4;
  P.batch.lseq = reinterpret_cast<const uint16_t *>(d + o_lseq);
  P.batch.flag = reinterpret_cast<const uint16_t *>(d + o_flag);
  P.batch.region = reinterpret_cast<const uint16_t *>(d + o_region);
  P.batch.mapq = d + o_mapq;
  P.batch.same_tid = d + o_same;
  P.batch.score_diff = d + o_sd;
  P.batch.clipped = d + o_clip;
  P.batch.leftover = d + o_left;
  P.batch.isize = reinterpret_cast<const int32_t *>(d + o_isize);
  P.batch.sample = reinterpret_cast<const int32_t *>(d + o_sample);
  P.batch.mate = reinterpret_cast<const int32_t *>(d + o_mate);
  P.batch.unit = reinterpret_cast<const int32_t *>(d + o_unit);
  P.batch.unit_record = reinterpret_cast<const int32_t *>(d + o_urec);
  P.summaries = static_cast<TaskSummary *>(B.d_summaries.p);
  P.path_pool = static_cast<uint32_t *>(B.d_pool.p);
  P.path_pool_cap = pool_words;
  P.counters = static_cast<DevCounters *>(B.d_counters.p);
  P.cand_spill = c->d_spill.p;
  P.n_active = n_active;
  P.active_tasks = reinterpret_cast<const uint32_t *>(d + o_active);
  P.seed_recs = B.d_seedrecs.p;
  P.lab_recs = static_cast<uint8_t *>(B.d_seedrecs.p) + align_up((size_t)n_active * SEED_REC_BYTES + 64);
  {
    PrepParams & Q = B.prep;
    memset(&Q, 0, sizeof(Q));
    Q.n_records = (uint32_t)total;
    Q.dup_of = reinterpret_cast<const int32_t *>(d + o_dup);
    Q.mate = reinterpret_cast<int32_t *>(d + o_mate);
    Q.sample = reinterpret_cast<int32_t *>(d + o_sample);
    Q.lseq = P.batch.lseq;
    Q.flag = P.batch.flag;
    Q.same_tid = P.batch.same_tid;
    Q.isize = P.batch.isize;
    Q.region = P.batch.region;
    Q.regions = P.regions;
    Q.scan = reinterpret_cast<unsigned long long *>(d + o_scan);
    Q.unit = reinterpret_cast<int32_t *>(d + o_unit);
    Q.unit_record = reinterpret_cast<int32_t *>(d + o_urec);
    Q.active = reinterpret_cast<uint32_t *>(d + o_active);
    Q.counters = P.counters;
  }
  P.task_times = nullptr;
  if (getenv("GTB_TASK_TIMES")) // profiling aid: per-task start/end times of chain_kernel, dumped by collect_chunks
  {
    if (int rc = B.d_task_times.reserve((size_t)n_active * 16 + 16))
      return rc;
    P.task_times = static_cast<unsigned long long *>(B.d_task_times.p);
  }
  P.slow_tasks = static_cast<uint32_t *>(B.d_slow.p);
  P.huge_tasks = P.slow_tasks + n_active + 16;
  P.deferred = P.huge_tasks + n_active + 16;
  P.gen_tasks = P.deferred + total + 16;
  P.slow2_tasks = P.gen_tasks + n_active + 16;
  P.gen_lanes = (uint32_t)c->gen_lanes;
  P.pending = reinterpret_cast<uint8_t *>(P.summaries + n_tasks);
  P.huge_states = c->d_huge.p;
  if (with_tap)
  {
    size_t const tap_counts = (size_t)n_tasks * (NLISTS * 2 + 1) * 4;
    size_t const tap_labels = (size_t)n_tasks * 512 + 65536;
    if (int rc = c->d_tap_counts.reserve(tap_counts + 64))
      return rc;
    if (int rc = c->d_tap_pool.reserve(tap_labels * sizeof(DevLabel)))
      return rc;
    uint32_t * tc = static_cast<uint32_t *>(c->d_tap_counts.p);
    P.tap.list_count = tc;
    P.tap.list_off = tc + (size_t)n_tasks * NLISTS;
    P.tap.nslots = tc + (size_t)n_tasks * NLISTS * 2;
    P.tap.pool = static_cast<DevLabel *>(c->d_tap_pool.p);
    P.tap.pool_cap = tap_labels;
  }
  B.n_tasks = n_tasks;
  return 0;
}

// After a submit: occupancy + overflow flag of every connection table it touched (one small D2H each, one synchronisation).
static int conn_check(Ctx * c, int n, Region * const * regs)
{
  int m = 0;
  for (int i = 0; i < n; ++i)
    m += regs[i]->conn_cap != 0;
  if (m == 0)
    return 0;
  if (int rc = c->h_conn_state.reserve((size_t)m * 8))
    return rc;
  uint32_t * h = static_cast<uint32_t *>(c->h_conn_state.p);
  int k = 0;
  for (int i = 0; i < n; ++i)
    if (regs[i]->conn_cap)
      CUDA_TRY(cudaMemcpyAsync(h + 2 * k++, regs[i]->dev.conn_state, 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(wait_stream(c));
  k = 0;
  bool full = false;
  for (int i = 0; i < n; ++i)
    if (regs[i]->conn_cap)
    {
      regs[i]->conn_used = h[2 * k];
      full = full || h[2 * k + 1] != 0;
      ++k;
    }
  if (full)
    return fail(GTB_ERR_CAPACITY, "phasing-connection table full: raise the per-record budget with gtb_set_connections(ctx, n)");
  return 0;
}

int gtb_submit_reads_multi(gtb_ctx * ctx, int n, const int * region_ids, const gtb_read_batch * batches,
                           gtb_submit_stats * stats)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || n <= 0 || !region_ids || !batches)
    return fail(GTB_ERR_ARG, "bad arguments");
  if (c->device < 0)
    return fail(GTB_ERR_CUDA, "host-only context: the genotyping path needs a CUDA device (no CPU fallback)");
  cudaSetDevice(c->device);
  size_t total = 0;
  std::vector<Region *> regs(n);
  for (int i = 0; i < n; ++i)
  {
    auto it = c->regions.find(region_ids[i]);
    if (it == c->regions.end())
      return fail(GTB_ERR_STATE, "unknown region in submit");
    if (!it->second->pool_open)
      return fail(GTB_ERR_STATE, "gtb_pool_begin must precede gtb_submit_reads");
    if (it->second->poisoned)
      return fail(GTB_ERR_STATE, "an earlier submit on this pool failed: gtb_pool_reset before using it again");
    if (it->second->reduced)
      return fail(GTB_ERR_STATE, "the pool's accumulators were all-reduced in place: gtb_pool_reset before submitting again");
    regs[i] = it->second.get();
    if (batches[i].n_reads && batches[i].seq_stride != GTB_SEQ_STRIDE)
      return fail(GTB_ERR_ARG, "seq_stride must be GTB_SEQ_STRIDE (76)");
    total += batches[i].n_reads;
  }
  if (total >= 0x7FFFFFFFull)
    return fail(GTB_ERR_ARG, "batch too large");
  for (int i = 0; i < n; ++i)
    if (int rc = conn_reserve(c, *regs[i], batches[i].n_reads))
      return rc;
  if (int rc = upload_region_table(c))
    return rc;
  c->have_last = false;

  // ---- chunking: host staging of chunk k+1 overlaps the H2D copy and the kernels of chunk k.
  //      Region boundaries are natural cut points (mate / duplicate links never cross regions).
  int n_chunks = 1;
  if (!c->debug && n >= 2)
  {
    // Two chunks on two streams hide half of the H2D copy behind the first chunk's kernels.  Measured on B200/PCIe5 with
    // the 2e5-record step (tools/chunk_sweep.sh): 1 chunk 2.36-2.43 ms, 2 chunks 2.12-2.17 ms, 3 chunks 2.15-2.44 ms,
    // 4 chunks 2.35 ms end to end -- every chunk pays slow_kernel's fixed ~0.1 ms latency again, so more chunks lose what
    // the overlap wins.  gtb_set_chunks / GTB_CHUNKS override (1..4).
    static int const env_forced = []() { const char * e = getenv("GTB_CHUNKS"); return e ? atoi(e) : 0; }();
    int const forced = c->forced_chunks > 0 ? c->forced_chunks : env_forced;
    if (forced > 0)
      n_chunks = std::min({n, MAX_CHUNKS, forced});
    else if (total >= 65536)
      n_chunks = std::min(n, 4);
  }
  std::vector<int> cut(n_chunks + 1, n);
  cut[0] = 0;
  {
    // cumulative share of the records at the end of chunk k.  The first chunk is smaller (10 % of the records; GTB_CHUNK_HEAD
    // overrides, 0 = equal chunks) so that the device starts after a shorter copy; the rest is split evenly.  Measured
    // end to end on the bench workload, same run: equal chunks 1.92 / 1.86 ms, 10 % head 1.83 ms.
    static int const head_pct = []() { const char * e = getenv("GTB_CHUNK_HEAD"); return e ? std::max(0, std::min(90, atoi(e))) : 10; }();
    auto share = [&](int k) -> double {
      if (head_pct <= 0 || n_chunks < 2)
        return (double)k / n_chunks;
      double const h = head_pct / 100.0;
      return h + (1.0 - h) * (double)(k - 1) / (n_chunks - 1);
    };
    size_t acc = 0;
    int k = 1;
    for (int i = 0; i < n && k < n_chunks; ++i)
    {
      acc += batches[i].n_reads;
      while (k < n_chunks && (double)acc >= (double)total * share(k))
        cut[k++] = i + 1;
    }
  }
  for (int k = 1; k <= n_chunks; ++k)
    cut[k] = std::max(cut[k], cut[k - 1]);
  cut[n_chunks] = n;
  c->n_chunks_last = n_chunks;
  static bool const trace = getenv("GTB_TRACE") != nullptr; // host-side timeline of a submit (profiling aid)
  auto const t_begin = std::chrono::steady_clock::now();
  auto since = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count(); };
  std::string tr;
  CUDA_TRY(cudaEventRecord(c->ev_slow[2], c->stream)); // the chunk streams start after everything queued so far (resets)
  for (int k = 0; k < n_chunks; ++k)
  {
    int const b0 = cut[k], m = cut[k + 1] - cut[k];
    if (int rc = stage_chunk(c, c->bs[k], m, region_ids + b0, batches + b0, regs.data() + b0, c->debug))
      return rc;
    c->bs[k].with_conn = false;
    for (int i = b0; i < b0 + m; ++i)
      c->bs[k].with_conn = c->bs[k].with_conn || regs[i]->conn_cap != 0;
    if (trace)
      tr += " staged" + std::to_string(k) + "=" + std::to_string((int)since());
    if (int rc = launch_front(c, c->bs[k], c->ev_slow[2]))
      return rc;
    if (trace)
      tr += " front" + std::to_string(k) + "=" + std::to_string((int)since());
  }
  if (int rc = launch_back(c, n_chunks))
    return rc;
  if (trace)
    tr += " back=" + std::to_string((int)since());
  c->have_last = true;
  int const rc_collect = collect_chunks(c, stats, true);
  if (trace)
    fprintf(stderr, "[gtb trace us]%s synced=%d | dev: h2d %.0f prep %.0f probe %.0f chain %.0f score0 %.0f slow %.0f score1 %.0f total %.0f\n",
            tr.c_str(), (int)since(), c->t_h2d * 1e3, c->t_prep * 1e3, c->t_probe * 1e3, c->t_chain * 1e3, c->t_score0 * 1e3,
            c->t_slow * 1e3, c->t_score1 * 1e3, c->t_total * 1e3);
  if (rc_collect)
  {
    for (Region * r : regs)
      r->poisoned = true; // part of the batch may already sit in the accumulators (include/gtb200.h, gtb_pool_begin)
    return rc_collect;
  }
  return conn_check(c, n, regs.data());
}

// Device block of a record-parsing submit: the raw records (core fields, data blocks, offsets, read groups), the name-hash
// sort buffers and the small per-region tables.
struct BamLayout
{
  size_t o_core, o_doff, o_rg, o_hash, o_hash2, o_idx, o_idx2, o_tab, o_data, bytes;
  BamLayout(size_t total, size_t n_regions, size_t n_data)
  {
    size_t off = 0;
    o_core = place<gtb_bam_core>(off, total);
    o_doff = place<unsigned long long>(off, total + n_regions);
    o_rg = place<int32_t>(off, total);
    o_hash = place<unsigned long long>(off, total);
    o_hash2 = place<unsigned long long>(off, total);
    o_idx = place<uint32_t>(off, total);
    o_idx2 = place<uint32_t>(off, total);
    o_tab = place<unsigned long long>(off, n_regions * 3 + 4); // data_base[n] | rec_begin[n + 1] | slots[n]
    o_data = place<uint8_t>(off, n_data + 16);
    bytes = align_up(off, 256);
  }
};

// small tables of a record-parsing submit through the chunk's pinned staging buffer (copy stream)
static int upload_bam_tables(Ctx * c, BatchState & B, BamLayout const & L, int n, const unsigned long long * data_base,
                             const uint32_t * rec_begin, Region * const * regs)
{
  size_t const tab_bytes = (size_t)n * 8 + (size_t)(n + 1) * 4 + (size_t)n * 2;
  if (int rc = B.h_batch.reserve(align_up(tab_bytes, 256)))
    return rc;
  uint8_t * ht = static_cast<uint8_t *>(B.h_batch.p);
  memcpy(ht, data_base, (size_t)n * 8);
  memcpy(ht + (size_t)n * 8, rec_begin, (size_t)(n + 1) * 4);
  for (int i = 0; i < n; ++i)
    reinterpret_cast<uint16_t *>(ht + (size_t)n * 8 + (size_t)(n + 1) * 4)[i] = (uint16_t)regs[i]->slot;
  CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(B.d_bam.p) + L.o_tab, ht, tab_bytes, cudaMemcpyHostToDevice, c->copy_stream));
  return 0;
}

// The raw records of chunk B are (or will be, once B.ev[1] has fired) on the device in layout L: parse them and run the
// launch sequence of a submit.
static int run_bam_chunk(Ctx * c, BatchState & B, ChunkLayout const & Lo, BamLayout const & L, size_t total, int n,
                         Region * const * regs, gtb_submit_stats * stats)
{
  uint8_t * r = static_cast<uint8_t *>(B.d_bam.p);
  B.bam_sort_bytes = bam_sort_temp_bytes((uint32_t)std::max<size_t>(total, 1));
  if (int rc = B.d_bam_sort.reserve(B.bam_sort_bytes + 16))
    return rc;
  if (int rc = bind_chunk(c, B, Lo, total, c->debug))
    return rc;
  uint8_t * d = static_cast<uint8_t *>(B.d_batch.p);
  BamParams & Q = B.bam;
  memset(&Q, 0, sizeof(Q));
  Q.n = (uint32_t)total;
  Q.n_regions = (uint32_t)n;
  Q.data_base = reinterpret_cast<const unsigned long long *>(r + L.o_tab);
  Q.rec_begin = reinterpret_cast<const uint32_t *>(r + L.o_tab + (size_t)n * 8);
  Q.slots = reinterpret_cast<const uint16_t *>(r + L.o_tab + (size_t)n * 8 + (size_t)(n + 1) * 4);
  Q.regions = B.P.regions;
  Q.core = reinterpret_cast<const gtb_bam_core *>(r + L.o_core);
  Q.data = r + L.o_data;
  Q.data_off = reinterpret_cast<const unsigned long long *>(r + L.o_doff);
  Q.rg = reinterpret_cast<const int32_t *>(r + L.o_rg);
  Q.seq4 = d + Lo.o_seq4;
  Q.lseq = reinterpret_cast<uint16_t *>(d + Lo.o_lseq);
  Q.flag = reinterpret_cast<uint16_t *>(d + Lo.o_flag);
  Q.region = reinterpret_cast<uint16_t *>(d + Lo.o_region);
  Q.mapq = d + Lo.o_mapq;
  Q.same_tid = d + Lo.o_same;
  Q.score_diff = d + Lo.o_sd;
  Q.clipped = d + Lo.o_clip;
  Q.leftover = d + Lo.o_left;
  Q.isize = reinterpret_cast<int32_t *>(d + Lo.o_isize);
  Q.mate = reinterpret_cast<int32_t *>(d + Lo.o_mate);
  Q.dup_of = reinterpret_cast<int32_t *>(d + Lo.o_dup);
  {
    const char * hb = getenv("GTB_BAM_HASH_BITS"); // tests: narrow the name hash so that different names share a run
    int const bits = hb ? std::max(1, std::min(64, atoi(hb))) : 64;
    Q.hash_mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1ull);
  }
  Q.name_hash = reinterpret_cast<unsigned long long *>(r + L.o_hash);
  Q.name_hash_sorted = reinterpret_cast<unsigned long long *>(r + L.o_hash2);
  Q.idx = reinterpret_cast<uint32_t *>(r + L.o_idx);
  Q.idx_sorted = reinterpret_cast<uint32_t *>(r + L.o_idx2);
  Q.counters = B.P.counters;
  B.with_conn = false;
  for (int i = 0; i < n; ++i)
    B.with_conn = B.with_conn || regs[i]->conn_cap != 0;
  if (int rc = launch_front(c, B, c->ev_slow[2]))
    return rc;
  if (int rc = launch_back(c, 1))
    return rc;
  c->have_last = true;
  if (int rc = collect_chunks(c, stats, true))
  {
    for (int i = 0; i < n; ++i)
      regs[i]->poisoned = true; // part of the batch may already sit in the accumulators
    return rc;
  }
  return conn_check(c, n, regs);
}

// Records of one pool per region as htslib holds them -> parsed, paired and de-duplicated on the device, then the same
// kernels as gtb_submit_reads_multi (one chunk: the mates of a pool pair within one call).
int gtb_submit_bam_records_multi(gtb_ctx * ctx, int n, const int * region_ids, const gtb_bam_batch * batches, gtb_submit_stats * stats)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || n <= 0 || !region_ids || !batches)
    return fail(GTB_ERR_ARG, "bad arguments");
  if (c->device < 0)
    return fail(GTB_ERR_CUDA, "host-only context: the genotyping path needs a CUDA device (no CPU fallback)");
  std::vector<Region *> regs(n);
  std::vector<uint32_t> rec_begin(n + 1, 0);
  std::vector<unsigned long long> data_base(n + 1, 0);
  for (int i = 0; i < n; ++i)
  {
    auto it = c->regions.find(region_ids[i]);
    if (it == c->regions.end())
      return fail(GTB_ERR_STATE, "unknown region in submit");
    if (!it->second->pool_open)
      return fail(GTB_ERR_STATE, "gtb_pool_begin must precede gtb_submit_bam_records");
    if (it->second->poisoned)
      return fail(GTB_ERR_STATE, "an earlier submit on this pool failed: gtb_pool_reset before using it again");
    if (it->second->reduced)
      return fail(GTB_ERR_STATE, "the pool's accumulators were all-reduced in place: gtb_pool_reset before submitting again");
    regs[i] = it->second.get();
    gtb_bam_batch const & b = batches[i];
    if (b.n_reads && (!b.core || !b.data || !b.data_off || !b.sample || !b.rg))
      return fail(GTB_ERR_ARG, "record batch has null arrays");
    if (b.n_reads && b.data_off[0] != 0)
      return fail(GTB_ERR_ARG, "data_off[0] must be 0");
    if ((unsigned long long)rec_begin[i] + b.n_reads >= 0x7FFFFFFFull)
      return fail(GTB_ERR_ARG, "batch too large");
    rec_begin[i + 1] = rec_begin[i] + b.n_reads;
    data_base[i + 1] = data_base[i] + (b.n_reads ? b.data_off[b.n_reads] : 0);
  }
  size_t const total = rec_begin[n], n_data = (size_t)data_base[n];
  cudaSetDevice(c->device);
  for (int i = 0; i < n; ++i)
    if (int rc = conn_reserve(c, *regs[i], batches[i].n_reads))
      return rc;
  if (int rc = upload_region_table(c))
    return rc;
  c->have_last = false;
  c->n_chunks_last = 1;
  BatchState & B = c->bs[0];
  ChunkLayout const Lo(total);
  if (int rc = B.d_batch.reserve(Lo.bytes))
    return rc;
  B.regions.assign(region_ids, region_ids + n);
  B.rec_begin = rec_begin;
  B.unit_begin.clear();
  BamLayout const L(total, (size_t)n, n_data);
  if (int rc = B.d_bam.reserve(L.bytes))
    return rc;
  uint8_t * r = static_cast<uint8_t *>(B.d_bam.p);
  uint8_t * d = static_cast<uint8_t *>(B.d_batch.p);
  CUDA_TRY(cudaEventRecord(c->ev_slow[2], c->stream));
  CUDA_TRY(cudaEventRecord(B.ev[0], c->copy_stream));
  if (int rc = upload_bam_tables(c, B, L, n, data_base.data(), rec_begin.data(), regs.data()))
    return rc;
  for (int i = 0; i < n; ++i)
  {
    gtb_bam_batch const & b = batches[i];
    size_t const m = b.n_reads, at = rec_begin[i];
    if (m == 0)
      continue;
    CUDA_TRY(cudaMemcpyAsync(r + L.o_core + at * sizeof(gtb_bam_core), b.core, m * sizeof(gtb_bam_core), cudaMemcpyHostToDevice, c->copy_stream));
    CUDA_TRY(cudaMemcpyAsync(r + L.o_doff + (at + i) * 8, b.data_off, (m + 1) * 8, cudaMemcpyHostToDevice, c->copy_stream));
    CUDA_TRY(cudaMemcpyAsync(r + L.o_rg + at * 4, b.rg, m * 4, cudaMemcpyHostToDevice, c->copy_stream));
    CUDA_TRY(cudaMemcpyAsync(r + L.o_data + data_base[i], b.data, (size_t)b.data_off[m], cudaMemcpyHostToDevice, c->copy_stream));
    CUDA_TRY(cudaMemcpyAsync(d + Lo.o_sample + at * 4, b.sample, m * 4, cudaMemcpyHostToDevice, c->copy_stream));
  }
  CUDA_TRY(cudaEventRecord(B.ev[1], c->copy_stream));
  return run_bam_chunk(c, B, Lo, L, total, n, regs.data(), stats);
}

int gtb_submit_bam_records(gtb_ctx * ctx, int region_id, const gtb_bam_batch * batch, gtb_submit_stats * stats)
{
  return gtb_submit_bam_records_multi(ctx, 1, &region_id, batch, stats);
}

// ---------------------------------------------------------------------------------------------------------------------
// BGZF entry: compressed blocks of the pool's files -> records in merge order on the device -> gtb_submit_bam_records' path.
namespace
{
struct BgzfPlan
{
  std::vector<BgzfBlock> blocks;
  std::vector<BgzfSegment> segs;
  std::vector<BgzfFile> files;
  std::vector<const uint8_t *> seg_src; // host bytes of every segment
  std::vector<unsigned long long> seg_comp_off, seg_comp_bytes;
  unsigned long long comp_bytes = 0, out_bytes = 0;
  uint32_t n_slots = 0, n_block_slots = 0;
};

// Walks the block headers of every segment (the only thing the host reads of the compressed bytes).
int plan_bgzf(int n_files, const gtb_bgzf_file * files, BgzfPlan & P)
{
  for (int fi = 0; fi < n_files; ++fi)
  {
    gtb_bgzf_file const & f = files[fi];
    if (f.n_segments && !f.segments)
      return fail(GTB_ERR_ARG, "gtb_submit_bgzf: file without segments array");
    BgzfFile F{};
    F.seg_begin = (uint32_t)P.segs.size();
    F.sample = f.sample;
    F.rg = f.rg;
    unsigned long long file_out = 0;
    for (uint32_t si = 0; si < f.n_segments; ++si)
    {
      gtb_bgzf_segment const & g = f.segments[si];
      if (g.comp_bytes && !g.comp)
        return fail(GTB_ERR_ARG, "gtb_submit_bgzf: segment without bytes");
      BgzfSegment S{};
      S.block_begin = (uint32_t)P.blocks.size();
      S.out_begin = P.out_bytes;
      S.v_end = g.v_end;
      S.first_offset = g.first_offset;
      S.to_eof = g.to_eof;
      unsigned long long const comp_base = P.comp_bytes;
      unsigned long long at = 0;
      while (at < g.comp_bytes)
      {
        BgzfBlockInfo info{};
        int const rc = bgzf_block_info(g.comp + at, g.comp_bytes - at, &info);
        if (rc == INF_ERR_INPUT && at > 0)
          break; // a partial block at the end of the bytes handed over: ignored (the walk reports it if it needs it)
        if (rc != INF_OK)
          return fail(GTB_ERR_INPUT, "gtb_submit_bgzf: not a BGZF block header");
        if (info.isize > 65536u)
          return fail(GTB_ERR_INPUT, "gtb_submit_bgzf: BGZF block announces more than 64 KiB");
        if (info.isize == 0)
        {
          // the end-of-file marker: bgzf_read stops at an empty block (bgzf.c:1242-1244); nothing behind it is looked at.
          // The offset behind the last record is then this block's address (bgzf.c:1266-1269), which `at` still is.
          S.to_eof = 1;
          break;
        }
        BgzfBlock b{};
        b.comp_off = comp_base + at;
        b.file_off = g.file_offset + at;
        b.out_off = P.out_bytes;
        b.comp_bytes = info.block_bytes;
        b.isize = info.isize;
        b.segment = (uint32_t)P.segs.size();
        b.slot_base = P.n_block_slots;
        P.n_block_slots += info.isize / 36 + 1;
        P.blocks.push_back(b);
        P.out_bytes += info.isize;
        at += info.block_bytes;
      }
      S.block_end = (uint32_t)P.blocks.size();
      S.out_end = P.out_bytes;
      S.end_file_off = g.file_offset + at;
      if (S.first_offset > S.out_end - S.out_begin)
        return fail(GTB_ERR_ARG, "gtb_submit_bgzf: first_offset beyond the first block");
      P.seg_src.push_back(g.comp);
      P.seg_comp_off.push_back(comp_base);
      P.seg_comp_bytes.push_back(at);
      P.comp_bytes = align_up(comp_base + at, 16);
      file_out += S.out_end - S.out_begin;
      P.out_bytes = align_up(P.out_bytes, 16) + 64; // slack between segments: no record read crosses into the next one
      P.segs.push_back(S);
    }
    F.seg_end = (uint32_t)P.segs.size();
    F.rec_base = P.n_slots;
    F.rec_cap = (uint32_t)std::min<unsigned long long>(file_out / 36 + 2, 0x7FFFFFFFull);
    if ((unsigned long long)P.n_slots + F.rec_cap >= 0x7FFFFFFFull)
      return fail(GTB_ERR_ARG, "gtb_submit_bgzf: too many bytes in one call");
    P.n_slots += F.rec_cap;
    P.files.push_back(F);
  }
  return 0;
}

BamQuery make_query(const gtb_bgzf_query * q)
{
  BamQuery Q{};
  Q.tid = q->tid;
  Q.beg = q->beg;
  Q.end = q->end;
  Q.flag_filter = q->flag_filter;
  Q.sv_filter = q->sv_read_filter;
  Q.max_lseq = (uint32_t)MAX_SEQ;
  Q.whole_file = q->whole_file;
  return Q;
}

uint32_t g_bgzf_host_stitched = 0; // files of the last gtb_debug_bgzf_host whose per-block walks stitched

const char * bgzf_error_text(int st)
{
  switch (st)
  {
  case INF_ERR_INPUT: return "a DEFLATE stream ends before its block does";
  case INF_ERR_TYPE: return "DEFLATE block type 3";
  case INF_ERR_STORED: return "stored DEFLATE block with LEN != ~NLEN";
  case INF_ERR_CODE: return "invalid Huffman code lengths";
  case INF_ERR_SYMBOL: return "invalid symbol in a DEFLATE stream";
  case INF_ERR_DIST: return "DEFLATE match distance reaches before the block";
  case INF_ERR_OUTPUT: return "a BGZF block inflates to more than its ISIZE";
  case INF_ERR_SIZE: return "a BGZF block inflates to less than its ISIZE";
  case INF_ERR_CRC: return "CRC-32 mismatch in a BGZF block";
  case INF_ERR_HEADER: return "not a BGZF block header";
  case SCAN_ERR_TRUNCATED: return "the bytes handed over end before the chunk does";
  case SCAN_ERR_RECORD: return "malformed BAM record";
  case SCAN_ERR_CAPACITY: return "more records than the byte count allows";
  case SCAN_ERR_KEY: return "contig index or read length beyond the sort key";
  case SCAN_ERR_UNSORTED: return "a file is not in coordinate order";
  default: return "unknown decode error";
  }
}
} // namespace

int gtb_submit_bgzf(gtb_ctx * ctx, int region_id, int n_files, const gtb_bgzf_file * files, const gtb_bgzf_query * query,
                    gtb_submit_stats * stats)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || n_files < 0 || (n_files > 0 && !files) || !query)
    return fail(GTB_ERR_ARG, "bad arguments");
  if (c->device < 0)
    return fail(GTB_ERR_CUDA, "host-only context: the genotyping path needs a CUDA device (no CPU fallback)");
  auto it = c->regions.find(region_id);
  if (it == c->regions.end())
    return fail(GTB_ERR_STATE, "unknown region in submit");
  Region * reg = it->second.get();
  if (!reg->pool_open)
    return fail(GTB_ERR_STATE, "gtb_pool_begin must precede gtb_submit_bgzf");
  if (reg->poisoned)
    return fail(GTB_ERR_STATE, "an earlier submit on this pool failed: gtb_pool_reset before using it again");
  if (reg->reduced)
    return fail(GTB_ERR_STATE, "the pool's accumulators were all-reduced in place: gtb_pool_reset before submitting again");
  for (int i = 0; i < n_files; ++i)
    if (files[i].sample < 0 || files[i].sample >= reg->n_samples)
      return fail(GTB_ERR_ARG, "gtb_submit_bgzf: sample index outside the pool");
  BgzfPlan P;
  if (int rc = plan_bgzf(n_files, files, P))
    return rc;
  cudaSetDevice(c->device);
  if (int rc = upload_region_table(c))
    return rc;
  c->have_last = false;
  c->n_chunks_last = 1;
  BatchState & B = c->bs[0];
  B.bgzf_n = 0;
  // ---- device blocks
  size_t off = 0;
  size_t const o_blocks = place<BgzfBlock>(off, P.blocks.size());
  size_t const o_segs = place<BgzfSegment>(off, P.segs.size());
  size_t const o_files = place<BgzfFile>(off, P.files.size());
  size_t const tab_bytes = off;
  size_t const o_comp = place<uint8_t>(off, P.comp_bytes + 16);
  if (int rc = B.d_bgzf_in.reserve(align_up(off, 256)))
    return rc;
  if (int rc = B.d_bgzf_out.reserve(align_up(P.out_bytes + 64, 256)))
    return rc;
  size_t const ns = P.n_slots;
  size_t t = 0;
  size_t const o_status = place<uint32_t>(t, 8);
  size_t const o_nrec = place<uint32_t>(t, P.files.size());
  size_t const o_start = place<unsigned long long>(t, ns);
  size_t const o_keep = place<uint32_t>(t, ns);
  size_t const o_kpos = place<uint32_t>(t, ns);
  size_t const o_filt = place<uint8_t>(t, ns);
  size_t const o_sel = place<unsigned long long>(t, ns);
  size_t const o_sfile = place<uint32_t>(t, ns);
  size_t const o_key = place<unsigned long long>(t, ns);
  size_t const o_key2 = place<unsigned long long>(t, ns);
  size_t const o_idx = place<uint32_t>(t, ns);
  size_t const o_idx2 = place<uint32_t>(t, ns);
  size_t const o_newg = place<uint32_t>(t, ns);
  size_t const o_rank = place<uint32_t>(t, ns);
  size_t const o_fsort = place<uint32_t>(t, ns);
  size_t const o_final = place<uint32_t>(t, ns);
  size_t const o_keep2 = place<uint32_t>(t, ns);
  size_t const o_k2pos = place<uint32_t>(t, ns);
  size_t const o_outidx = place<uint32_t>(t, ns);
  size_t const o_perm = place<uint32_t>(t, ns);
  size_t const o_bslot = place<unsigned long long>(t, P.n_block_slots);
  size_t const o_walks = place<BlockWalk>(t, P.blocks.size());
  size_t const o_take = place<uint32_t>(t, P.blocks.size());
  size_t const o_bdst = place<uint32_t>(t, P.blocks.size());
  if (int rc = B.d_bgzf_tmp.reserve(align_up(t, 256)))
    return rc;
  size_t const cub_bytes = bgzf_temp_bytes((uint32_t)std::max<size_t>(ns, 1));
  if (int rc = B.d_bgzf_cub.reserve(cub_bytes))
    return rc;
  if (int rc = B.h_bgzf.reserve(align_up(tab_bytes + 128, 256)))
    return rc;
  uint8_t * din = static_cast<uint8_t *>(B.d_bgzf_in.p);
  uint8_t * dt = static_cast<uint8_t *>(B.d_bgzf_tmp.p);
  uint8_t * hb = static_cast<uint8_t *>(B.h_bgzf.p);
  uint32_t * h_status = reinterpret_cast<uint32_t *>(hb + align_up(tab_bytes, 16));
  if (!P.blocks.empty())
    memcpy(hb + o_blocks, P.blocks.data(), P.blocks.size() * sizeof(BgzfBlock));
  if (!P.segs.empty())
    memcpy(hb + o_segs, P.segs.data(), P.segs.size() * sizeof(BgzfSegment));
  if (!P.files.empty())
    memcpy(hb + o_files, P.files.data(), P.files.size() * sizeof(BgzfFile));
  cudaStream_t const s = B.stream;
  CUDA_TRY(cudaEventRecord(c->ev_slow[2], c->stream));
  CUDA_TRY(cudaStreamWaitEvent(s, c->ev_slow[2], 0));
  CUDA_TRY(cudaEventRecord(B.ev[0], s)); // "H2D" of this entry = compressed bytes + decode, up to the resident record batch
  if (tab_bytes)
    CUDA_TRY(cudaMemcpyAsync(din, hb, tab_bytes, cudaMemcpyHostToDevice, s));
  for (size_t k = 0; k < P.seg_src.size(); ++k)
    if (P.seg_comp_bytes[k])
      CUDA_TRY(cudaMemcpyAsync(din + o_comp + P.seg_comp_off[k], P.seg_src[k], P.seg_comp_bytes[k], cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemsetAsync(dt + o_status, 0, 32, s));
  BgzfParams & Z = B.bgzf;
  memset(&Z, 0, sizeof(Z));
  Z.comp = din + o_comp;
  Z.out = static_cast<uint8_t *>(B.d_bgzf_out.p);
  Z.blocks = reinterpret_cast<const BgzfBlock *>(din + o_blocks);
  Z.segs = reinterpret_cast<const BgzfSegment *>(din + o_segs);
  Z.files = reinterpret_cast<const BgzfFile *>(din + o_files);
  Z.n_blocks = (uint32_t)P.blocks.size();
  Z.n_files = (uint32_t)P.files.size();
  Z.n_slots = (uint32_t)ns;
  Z.check_crc = query->check_crc;
  Z.q = make_query(query);
  Z.status = reinterpret_cast<int *>(dt + o_status);
  Z.n_too_long = reinterpret_cast<uint32_t *>(dt + o_status) + 1;
  Z.n_kept = reinterpret_cast<uint32_t *>(dt + o_status) + 2;
  Z.n_final = reinterpret_cast<uint32_t *>(dt + o_status) + 3;
  Z.need_host = reinterpret_cast<uint32_t *>(dt + o_status) + 4;
  Z.n_serial_files = reinterpret_cast<uint32_t *>(dt + o_status) + 5;
  Z.block_slot = reinterpret_cast<unsigned long long *>(dt + o_bslot);
  Z.walks = reinterpret_cast<BlockWalk *>(dt + o_walks);
  Z.block_take = reinterpret_cast<uint32_t *>(dt + o_take);
  Z.block_dst = reinterpret_cast<uint32_t *>(dt + o_bdst);
  Z.n_block_slots = P.n_block_slots;
  {
    static bool const serial = getenv("GTB_BGZF_SERIAL_WALK") != nullptr;
    Z.serial_walk = serial ? 1u : 0u;
  }
  Z.file_nrec = reinterpret_cast<uint32_t *>(dt + o_nrec);
  Z.rec_start = reinterpret_cast<unsigned long long *>(dt + o_start);
  Z.keep = reinterpret_cast<uint32_t *>(dt + o_keep);
  Z.keep_pos = reinterpret_cast<uint32_t *>(dt + o_kpos);
  Z.filtered = dt + o_filt;
  Z.sel_start = reinterpret_cast<unsigned long long *>(dt + o_sel);
  Z.sel_file = reinterpret_cast<uint32_t *>(dt + o_sfile);
  Z.key = reinterpret_cast<unsigned long long *>(dt + o_key);
  Z.key_sorted = reinterpret_cast<unsigned long long *>(dt + o_key2);
  Z.idx = reinterpret_cast<uint32_t *>(dt + o_idx);
  Z.idx_sorted = reinterpret_cast<uint32_t *>(dt + o_idx2);
  Z.new_group = reinterpret_cast<uint32_t *>(dt + o_newg);
  Z.rank = reinterpret_cast<uint32_t *>(dt + o_rank);
  Z.file_sorted = reinterpret_cast<uint32_t *>(dt + o_fsort);
  Z.idx_final = reinterpret_cast<uint32_t *>(dt + o_final);
  Z.keep2 = reinterpret_cast<uint32_t *>(dt + o_keep2);
  Z.keep2_pos = reinterpret_cast<uint32_t *>(dt + o_k2pos);
  Z.out_idx = reinterpret_cast<uint32_t *>(dt + o_outidx);
  static bool const trace = getenv("GTB_TRACE") != nullptr; // timeline of the decode (profiling aid)
  static thread_local cudaEvent_t tev[4] = {nullptr, nullptr, nullptr, nullptr};
  auto const t_begin = std::chrono::steady_clock::now();
  auto since = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count(); };
  if (trace && !tev[0])
    for (auto & e : tev)
      CUDA_TRY(cudaEventCreate(&e));
  if (launch_bgzf_front(Z, B.d_bgzf_cub.p, cub_bytes, s, trace ? reinterpret_cast<void * const *>(tev) : nullptr) != 0)
    return fail(GTB_ERR_CUDA, "gtb_submit_bgzf: decode launch failed");
  CUDA_TRY(cudaMemcpyAsync(h_status, dt + o_status, 32, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  CUDA_TRY(cudaGetLastError());
  if ((int)h_status[0] != 0)
    return fail(GTB_ERR_INPUT, std::string("gtb_submit_bgzf: ") + bgzf_error_text((int)h_status[0]));
  if (h_status[1] != 0)
    return fail(GTB_ERR_CAPACITY, "read longer than 152 bases (reference MAX_READ_LENGTH is 151)");
  double const us_front = since();
  uint32_t const m_order = ns ? h_status[2] : 0; // records the iterators return
  B.bgzf_stitched = (uint32_t)P.files.size() - h_status[5];
  size_t const total = ns ? h_status[3] : 0;     // records the pool loop keeps
  // ---- order of the pool's records
  const uint32_t * d_perm = nullptr;
  if (m_order)
  {
    if (launch_bgzf_order(Z, m_order, B.d_bgzf_cub.p, cub_bytes, s) != 0)
      return fail(GTB_ERR_CUDA, "gtb_submit_bgzf: sort launch failed");
    CUDA_TRY(cudaMemcpyAsync(h_status, dt + o_status, 32, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if ((int)h_status[0] != 0)
      return fail(GTB_ERR_INPUT, std::string("gtb_submit_bgzf: ") + bgzf_error_text((int)h_status[0]));
    static bool const force_merge = getenv("GTB_BGZF_FORCE_MERGE") != nullptr; // tests: always replay the reference's merge
    if (h_status[4] != 0 || force_merge)
    {
      // exact duplicates in different files / a position with more than 16 records: the reference's order depends on its
      // standard library's sort and heap, replayed here on (key, rank, file order, file) integers -- 20 bytes per record
      std::vector<unsigned long long> keys(m_order);
      std::vector<uint32_t> rank(m_order), order(m_order), file(m_order), perm;
      CUDA_TRY(cudaMemcpyAsync(keys.data(), Z.key_sorted, (size_t)m_order * 8, cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaMemcpyAsync(rank.data(), Z.rank, (size_t)m_order * 4, cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaMemcpyAsync(order.data(), Z.idx_sorted, (size_t)m_order * 4, cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaMemcpyAsync(file.data(), Z.file_sorted, (size_t)m_order * 4, cudaMemcpyDeviceToHost, s));
      CUDA_TRY(cudaStreamSynchronize(s));
      reference_merge_order(m_order, (uint32_t)P.files.size(), keys.data(), rank.data(), order.data(), file.data(), perm);
      if (perm.size() != m_order)
        return fail(GTB_ERR_INPUT, "gtb_submit_bgzf: merge replay lost records");
      CUDA_TRY(cudaMemcpyAsync(dt + o_perm, perm.data(), (size_t)m_order * 4, cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaStreamSynchronize(s)); // perm is a local
      d_perm = reinterpret_cast<const uint32_t *>(dt + o_perm);
    }
  }
  // ---- from here on: gtb_submit_bam_records with the batch already resident
  if (int rc = conn_reserve(c, *reg, total))
    return rc;
  ChunkLayout const Lo(total);
  if (int rc = B.d_batch.reserve(Lo.bytes))
    return rc;
  B.regions.assign(1, region_id);
  B.rec_begin.assign({0u, (uint32_t)total});
  B.unit_begin.clear();
  BamLayout const L(total, 1, (size_t)P.out_bytes);
  if (int rc = B.d_bam.reserve(L.bytes))
    return rc;
  uint8_t * r = static_cast<uint8_t *>(B.d_bam.p);
  Z.core = reinterpret_cast<gtb_bam_core *>(r + L.o_core);
  Z.data = r + L.o_data;
  Z.data_off = reinterpret_cast<unsigned long long *>(r + L.o_doff);
  Z.rg = reinterpret_cast<int32_t *>(r + L.o_rg);
  Z.sample = reinterpret_cast<int32_t *>(static_cast<uint8_t *>(B.d_batch.p) + Lo.o_sample);
  double const us_order = since();
  if (launch_bgzf_back(Z, m_order, (uint32_t)total, d_perm, B.d_bgzf_cub.p, cub_bytes, s) != 0)
    return fail(GTB_ERR_CUDA, "gtb_submit_bgzf: gather launch failed");
  if (trace)
  {
    float inflate = 0, walk = 0, compact = 0;
    cudaEventElapsedTime(&inflate, tev[0], tev[1]);
    cudaEventElapsedTime(&walk, tev[1], tev[2]);
    cudaEventElapsedTime(&compact, tev[2], tev[3]);
    fprintf(stderr, "[gtb trace us] bgzf: %zu blocks, %u + %u records | host: front synced %.0f, order + merge %.0f (replayed on the host: %d) | "
                    "dev: inflate %.0f walk %.0f classify + compact %.0f\n",
            P.blocks.size(), m_order, (unsigned)total, us_front, us_order - us_front, d_perm ? 1 : 0, inflate * 1e3, walk * 1e3, compact * 1e3);
  }
  B.bgzf_n = (uint32_t)total;
  unsigned long long const data_base[2] = {0, 0};
  uint32_t const rec_begin[2] = {0u, (uint32_t)total};
  CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, B.ev[0], 0));
  if (int rc = upload_bam_tables(c, B, L, 1, data_base, rec_begin, &reg))
    return rc;
  CUDA_TRY(cudaEventRecord(B.ev[4], s)); // the record batch is complete on the chunk stream ...
  CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, B.ev[4], 0));
  CUDA_TRY(cudaEventRecord(B.ev[1], c->copy_stream)); // ... and the tables on the copy stream: launch_front waits for ev[1]
  return run_bam_chunk(c, B, Lo, L, total, 1, &reg, stats);
}

int gtb_debug_bgzf_records(gtb_ctx * ctx, uint32_t * n_reads, uint64_t * n_data, gtb_bam_core * core, uint8_t * data,
                           uint64_t * data_off, int32_t * sample, int32_t * rg)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || c->device < 0 || !n_reads || !n_data)
    return fail(GTB_ERR_ARG, "bad arguments");
  cudaSetDevice(c->device);
  BatchState & B = c->bs[0];
  uint32_t const n = B.bgzf_n;
  *n_reads = n;
  *n_data = 0;
  if (n == 0)
    return 0;
  BgzfParams const & Z = B.bgzf;
  unsigned long long total = 0;
  CUDA_TRY(cudaMemcpy(&total, Z.data_off + n, 8, cudaMemcpyDeviceToHost));
  *n_data = total;
  if (core)
    CUDA_TRY(cudaMemcpy(core, Z.core, (size_t)n * sizeof(gtb_bam_core), cudaMemcpyDeviceToHost));
  if (data)
    CUDA_TRY(cudaMemcpy(data, Z.data, (size_t)total, cudaMemcpyDeviceToHost));
  if (data_off)
    CUDA_TRY(cudaMemcpy(data_off, Z.data_off, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost));
  if (sample)
    CUDA_TRY(cudaMemcpy(sample, Z.sample, (size_t)n * 4, cudaMemcpyDeviceToHost));
  if (rg)
    CUDA_TRY(cudaMemcpy(rg, Z.rg, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return 0;
}

// files whose per-block walks stitched: last gtb_debug_bgzf_host (ctx == NULL) or last gtb_submit_bgzf of ctx
int gtb_debug_bgzf_stitched(gtb_ctx * ctx, uint32_t * n_files_stitched)
{
  if (!n_files_stitched)
    return fail(GTB_ERR_ARG, "bad arguments");
  auto * c = reinterpret_cast<Ctx *>(ctx);
  *n_files_stitched = c ? c->bs[0].bgzf_stitched : g_bgzf_host_stitched;
  return 0;
}

int gtb_debug_bgzf_host(int n_files, const gtb_bgzf_file * files, const gtb_bgzf_query * query, uint32_t * n_reads,
                        uint64_t * n_data, gtb_bam_core * core, uint8_t * data, uint64_t * data_off, int32_t * sample, int32_t * rg,
                        uint64_t * n_inflated, uint8_t * inflated)
{
  if (n_files < 0 || (n_files > 0 && !files) || !query || !n_reads || !n_data)
    return fail(GTB_ERR_ARG, "bad arguments");
  BgzfPlan P;
  if (int rc = plan_bgzf(n_files, files, P))
    return rc;
  std::vector<uint8_t> comp(P.comp_bytes + 16, 0);
  for (size_t k = 0; k < P.seg_src.size(); ++k)
    if (P.seg_comp_bytes[k])
      memcpy(comp.data() + P.seg_comp_off[k], P.seg_src[k], P.seg_comp_bytes[k]);
  std::vector<uint8_t> out, d;
  std::vector<gtb_bam_core> cr;
  std::vector<unsigned long long> doff;
  std::vector<int32_t> smp, rgv;
  uint32_t too_long = 0, n_stitched = 0;
  int const st = bgzf_host_pipeline(comp.data(), P.blocks, P.segs, P.files, make_query(query), query->check_crc != 0, out, cr, d,
                                    doff, smp, rgv, &too_long, getenv("GTB_BGZF_FORCE_MERGE") != nullptr, &n_stitched);
  g_bgzf_host_stitched = n_stitched;
  if (st == -30)
    return fail(GTB_ERR_STATE, "gtb_debug_bgzf_host: the per-block walks disagree with the serial walk");
  if (st != 0)
    return fail(GTB_ERR_INPUT, std::string("gtb_debug_bgzf_host: ") + bgzf_error_text(st));
  if (too_long)
    return fail(GTB_ERR_CAPACITY, "read longer than 152 bases (reference MAX_READ_LENGTH is 151)");
  bool const fits = *n_reads >= cr.size() && *n_data >= d.size();
  *n_reads = (uint32_t)cr.size();
  *n_data = d.size();
  if (n_inflated)
  {
    bool const room = *n_inflated >= P.out_bytes;
    *n_inflated = P.out_bytes;
    if (inflated && room)
      memcpy(inflated, out.data(), (size_t)P.out_bytes);
  }
  if (!fits || !core)
    return 0; // sizes only
  memcpy(core, cr.data(), cr.size() * sizeof(gtb_bam_core));
  if (data && !d.empty())
    memcpy(data, d.data(), d.size());
  if (data_off)
    for (size_t i = 0; i < doff.size(); ++i)
      data_off[i] = doff[i];
  if (sample && !smp.empty())
    memcpy(sample, smp.data(), smp.size() * 4);
  if (rg && !rgv.empty())
    memcpy(rg, rgv.data(), rgv.size() * 4);
  return 0;
}

// The per-record columns the device derived in the last gtb_submit_bam_records (duplicate links resolved to the record whose
// alignment is re-used, as the reference's `prev` pointer would have it).
int gtb_debug_bam_columns(gtb_ctx * ctx, uint32_t n_reads, uint8_t * seq4, uint16_t * lseq, uint16_t * flag, uint8_t * mapq,
                          int32_t * isize, uint8_t * same_tid, uint8_t * score_diff, int32_t * mate, int32_t * dup_of,
                          uint8_t * leftover)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !c->have_last || c->bs[0].bam.n == 0 || c->bs[0].bam.n != n_reads)
    return fail(GTB_ERR_STATE, "the last submit was not a gtb_submit_bam_records call of that size");
  cudaSetDevice(c->device);
  BamParams const & Q = c->bs[0].bam;
  size_t const n = n_reads;
  if (seq4)
    CUDA_TRY(cudaMemcpy(seq4, Q.seq4, n * GTB_SEQ_STRIDE, cudaMemcpyDeviceToHost));
  if (lseq)
    CUDA_TRY(cudaMemcpy(lseq, Q.lseq, n * 2, cudaMemcpyDeviceToHost));
  if (flag)
    CUDA_TRY(cudaMemcpy(flag, Q.flag, n * 2, cudaMemcpyDeviceToHost));
  if (mapq)
    CUDA_TRY(cudaMemcpy(mapq, Q.mapq, n, cudaMemcpyDeviceToHost));
  if (isize)
    CUDA_TRY(cudaMemcpy(isize, Q.isize, n * 4, cudaMemcpyDeviceToHost));
  if (same_tid)
    CUDA_TRY(cudaMemcpy(same_tid, Q.same_tid, n, cudaMemcpyDeviceToHost));
  if (score_diff)
    CUDA_TRY(cudaMemcpy(score_diff, Q.score_diff, n, cudaMemcpyDeviceToHost));
  if (mate)
    CUDA_TRY(cudaMemcpy(mate, Q.mate, n * 4, cudaMemcpyDeviceToHost));
  if (leftover)
    CUDA_TRY(cudaMemcpy(leftover, Q.leftover, n, cudaMemcpyDeviceToHost));
  if (dup_of)
  {
    CUDA_TRY(cudaMemcpy(dup_of, Q.dup_of, n * 4, cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < n; ++k) // the device links each duplicate to its predecessor; roots are earlier, already final
      if (dup_of[k] >= 0 && dup_of[dup_of[k]] >= 0)
        dup_of[k] = dup_of[dup_of[k]];
  }
  return 0;
}

int gtb_submit_reads(gtb_ctx * ctx, int region_id, const gtb_read_batch * batch, gtb_submit_stats * stats)
{
  return gtb_submit_reads_multi(ctx, 1, &region_id, batch, stats);
}

// Re-runs the kernels on the batch that is already resident in HBM (bench: device-resident throughput).
int gtb_replay_last(gtb_ctx * ctx, gtb_submit_stats * stats)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !c->have_last)
    return fail(GTB_ERR_STATE, "no resident batch to replay");
  cudaSetDevice(c->device);
  CUDA_TRY(cudaEventRecord(c->ev_slow[2], c->stream)); // orders the chunk streams after earlier main-stream work
  for (int k = 0; k < c->n_chunks_last; ++k)
    if (int rc = launch_front(c, c->bs[k], c->ev_slow[2]))
      return rc;
  if (int rc = launch_back(c, c->n_chunks_last))
    return rc;
  return collect_chunks(c, stats, false);
}

// Device times (ms) of the last submit/replay measured with CUDA events on the library's streams.
int gtb_last_timing(gtb_ctx * ctx, float * h2d_ms, float * align_ms, float * score_ms, float * d2h_ms)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (h2d_ms)
    *h2d_ms = c->t_h2d;
  if (align_ms)
    *align_ms = c->t_align;
  if (score_ms)
    *score_ms = c->t_total - c->t_align; // everything after chain_kernel as a span: slow/huge beside the first score pass,
                                         // then the second pass -- so align_ms + score_ms = device time of the launch sequence
  if (d2h_ms)
    *d2h_ms = c->t_d2h;
  return 0;
}

// Per-kernel device times (ms) of the last submit/replay and the number of tasks the slow kernel handled.
int gtb_last_kernel_timing(gtb_ctx * ctx, float * probe_ms, float * chain_ms, float * slow_ms, float * score_ms,
                           uint64_t * n_slow)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (probe_ms)
    *probe_ms = c->t_probe;
  if (chain_ms)
    *chain_ms = c->t_chain;
  if (slow_ms)
    *slow_ms = c->t_slow;
  if (score_ms)
    *score_ms = c->t_score;
  if (n_slow)
    *n_slow = c->last_n_slow;
  return 0;
}

// Device time (ms) of the batch-preparation kernels (units, aligned orientations, link checks) of the last submit/replay.
int gtb_last_prep_timing(gtb_ctx * ctx, float * prep_ms)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !prep_ms)
    return fail(GTB_ERR_ARG, "bad arguments");
  *prep_ms = c->t_prep;
  return 0;
}

// Diagnostic counters of the last submit/replay (chunk 0): out[0..11] = why chain_kernel handed tasks to slow_kernel
// (refs vars paths locs labels cand_vars cands keys tap pool read_len probe-flag), out[12..23] = slow_kernel overflows.
int gtb_last_chain_timing(gtb_ctx * ctx, float * fast_ms, float * general_ms, uint64_t * n_tasks, uint64_t * n_general,
                          uint64_t * reasons16)
{
  Ctx * c = reinterpret_cast<Ctx *>(ctx);
  if (!c)
    return fail(GTB_ERR_ARG, "null context");
  if (fast_ms)
    *fast_ms = c->t_chain_fast;
  if (general_ms)
    *general_ms = c->t_chain - c->t_chain_fast;
  if (n_tasks)
    *n_tasks = c->last_n_active;
  if (n_general)
    *n_general = c->last_n_gen;
  if (reasons16)
    for (int q = 0; q < 16; ++q)
      reasons16[q] = c->t0_reasons[q];
  return 0;
}

int gtb_debug_counters(gtb_ctx * ctx, uint64_t * out24)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !c->have_last || !out24)
    return fail(GTB_ERR_STATE, "no submit yet");
  DevCounters const * k = static_cast<DevCounters *>(c->bs[0].h_counters.p);
  for (int q = 0; q < 12; ++q)
  {
    out24[q] = k->fast_reasons[q];
    out24[12 + q] = k->reasons[q];
  }
  return 0;
}

// Forces the number of pipeline chunks of gtb_submit_reads_multi (1..4; 0 = automatic).
int gtb_set_chunks(gtb_ctx * ctx, int n_chunks)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || n_chunks < 0 || n_chunks > MAX_CHUNKS)
    return fail(GTB_ERR_ARG, "n_chunks must be 0..8");
  c->forced_chunks = n_chunks;
  return 0;
}

// 1 (default): region indexes are built on the device; 0: by the host builder (gtb_index_host.hpp).
int gtb_set_index_build(gtb_ctx * ctx, int on_device)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c)
    return fail(GTB_ERR_ARG, "null ctx");
  c->index_build_device = on_device ? 1 : 0;
  return 0;
}

// Page-locked host memory for callers that want their batch columns DMA-ed without an intermediate staging copy.
int gtb_host_alloc(size_t bytes, void ** out)
{
  if (!out)
    return fail(GTB_ERR_ARG, "null out");
  cudaError_t const e = cudaMallocHost(out, bytes ? bytes : 1);
  if (e != cudaSuccess)
    return fail(GTB_ERR_CUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
  return 0;
}

int gtb_host_free(void * p)
{
  if (p)
    cudaFreeHost(p);
  return 0;
}

// Zeroes a region's accumulators (bench: repeated passes over the same batch).
int gtb_pool_reset(gtb_ctx * ctx, int region_id)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  auto it = c->regions.find(region_id);
  if (it == c->regions.end() || !it->second->pool_open)
    return fail(GTB_ERR_STATE, "unknown region / pool not open");
  cudaSetDevice(c->device);
  it->second->poisoned = it->second->reduced = false;
  CUDA_TRY(cudaMemsetAsync(it->second->accum.p, 0, it->second->accum_bytes, c->stream));
  if (it->second->conn_cap)
    CUDA_TRY(cudaMemsetAsync(it->second->conn.p, 0, conn_table_bytes(it->second->conn_cap), c->stream));
  it->second->conn_used = 0;
  return 0;
}

int gtb_accumulator_sizes(gtb_ctx * ctx, int region_id, uint32_t * n_bubbles, uint64_t * n_scores, uint64_t * n_cov)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !n_bubbles || !n_scores || !n_cov)
    return fail(GTB_ERR_ARG, "bad arguments");
  auto it = c->regions.find(region_id);
  if (it == c->regions.end())
    return fail(GTB_ERR_STATE, "unknown region");
  *n_bubbles = it->second->n_bubbles;
  *n_scores = it->second->score_off.back();
  *n_cov = it->second->cov_off.back();
  return 0;
}

static void convert_accumulators(Region const & R, const uint8_t * h, gtb_accumulators * out)
{
  uint8_t const * d0 = static_cast<uint8_t *>(R.accum.p);
  auto host_of = [&](const void * devp) { return h + (static_cast<const uint8_t *>(devp) - d0); };
  uint32_t const NB = R.n_bubbles, NS = (uint32_t)R.n_samples;
  uint64_t const n_scores = R.score_off.back(), n_cov = R.cov_off.back();
  out->n_bubbles = NB;
  out->n_samples = NS;
  for (uint32_t b = 0; b < NB; ++b)
  {
    out->bubble_id[b] = R.bubble_order[b];
    out->n_alleles[b] = R.n_alleles[b];
  }
  for (uint32_t b = 0; b <= NB; ++b)
  {
    out->score_off[b] = R.score_off[b];
    out->cov_off[b] = R.cov_off[b];
  }
  const uint32_t * ls = reinterpret_cast<const uint32_t *>(host_of(R.dev.log_score));
  const uint32_t * gc = reinterpret_cast<const uint32_t *>(host_of(R.dev.gt_cov));
  const uint32_t * ml = reinterpret_cast<const uint32_t *>(host_of(R.dev.max_log_score));
  const uint32_t * amb = reinterpret_cast<const uint32_t *>(host_of(R.dev.amb));
  const uint32_t * amba = reinterpret_cast<const uint32_t *>(host_of(R.dev.amb_alt));
  const uint32_t * altpp = reinterpret_cast<const uint32_t *>(host_of(R.dev.alt_pp));
  for (uint64_t i = 0; i < n_scores * NS; ++i)
    out->log_score[i] = (uint16_t)std::min<uint32_t>(ls[i], 0xFFFFu);
  for (uint64_t i = 0; i < n_cov * NS; ++i)
    out->gt_coverage[i] = (uint16_t)std::min<uint32_t>(gc[i], 0xFFFFu); // HapSample::increment_allele_depth saturates
  for (uint64_t i = 0; i < (uint64_t)NB * NS; ++i)
  {
    // explain_to_score stops adding reads once max_log_score >= 0xFFFF - eps (haplotype.cpp:561): with eps <= 8
    // nothing was dropped while the total stays below 0xFFFF - 8; beyond that the result is order dependent.
    out->saturated[i] = ml[i] >= (0xFFFFu - 8u) ? 1u : 0u;
    out->max_log_score[i] = (uint16_t)std::min<uint32_t>(ml[i], 0xFFFFu);
    out->ambiguous_depth[i] = (uint8_t)std::min<uint32_t>(amb[i], 0xFFu);
    out->ambiguous_depth_alt[i] = (uint8_t)std::min<uint32_t>(amba[i], 0xFFu);
    out->alt_proper_pair_depth[i] = (uint8_t)std::min<uint32_t>(altpp[i], 0xFFu);
  }
  memcpy(out->vs_clipped_reads, host_of(R.dev.vs_clipped_reads), (size_t)NB * 8);
  memcpy(out->vs_mapq_squared, host_of(R.dev.vs_mapq_squared), (size_t)NB * 8);
  memcpy(out->pa_clipped_bp, host_of(R.dev.pa_clipped_bp), n_cov * 8);
  memcpy(out->pa_mapq_squared, host_of(R.dev.pa_mapq_squared), n_cov * 8);
  memcpy(out->pa_score_diff, host_of(R.dev.pa_score_diff), n_cov * 8);
  memcpy(out->pa_mismatches, host_of(R.dev.pa_mismatches), n_cov * 8);
  memcpy(out->read_strand, host_of(R.dev.read_strand), n_cov * 16);
  out->depth_size = R.depth_size;
  out->reference_offset = R.reference_offset;
  if (out->ref_depth && R.depth_size)
  {
    const int * delta = reinterpret_cast<const int *>(host_of(R.dev.ref_depth_delta));
    for (uint32_t s = 0; s < NS; ++s)
    {
      long long run = 0;
      const int * dl = delta + (size_t)s * ((size_t)R.depth_size + 1);
      uint16_t * o = out->ref_depth + (size_t)s * R.depth_size;
      for (uint32_t k = 0; k < R.depth_size; ++k)
      {
        run += dl[k];
        o[k] = (uint16_t)std::min<long long>(std::max<long long>(run, 0), 0xFFFF);
      }
    }
  }
}

// Downloads the accumulators of several regions with ONE stream synchronisation; conversion runs in parallel.
int gtb_pool_finish_multi(gtb_ctx * ctx, int n, const int * region_ids, gtb_accumulators * outs)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || n <= 0 || !region_ids || !outs)
    return fail(GTB_ERR_ARG, "bad arguments");
  if (c->device < 0)
    return fail(GTB_ERR_CUDA, "host-only context");
  cudaSetDevice(c->device);
  std::vector<Region *> regs(n);
  std::vector<size_t> off(n + 1, 0);
  for (int i = 0; i < n; ++i)
  {
    auto it = c->regions.find(region_ids[i]);
    if (it == c->regions.end() || !it->second->pool_open)
      return fail(GTB_ERR_STATE, "unknown region / pool not open");
    regs[i] = it->second.get();
    off[i + 1] = off[i] + align_up(regs[i]->accum_bytes, 256);
  }
  if (int rc = c->h_accum.reserve(off[n]))
    return rc;
  uint8_t * h = static_cast<uint8_t *>(c->h_accum.p);
  if (n == 1)
    CUDA_TRY(cudaMemcpyAsync(h, regs[0]->accum.p, regs[0]->accum_bytes, cudaMemcpyDeviceToHost, c->stream));
  else
  {
    // one gather kernel + ONE D2H copy instead of a small (latency-bound) copy per region
    if (int rc = c->d_gather.reserve(off[n]))
      return rc;
    unsigned long long max_bytes = 0;
    std::vector<Segment> const seg = make_segments(n, [&](int i) { return Segment{regs[i]->accum.p, off[i], regs[i]->accum_bytes}; }, max_bytes);
    launch_gather_segments(seg.data(), n, max_bytes, c->d_gather.p, c->stream);
    CUDA_TRY(cudaMemcpyAsync(h, c->d_gather.p, off[n], cudaMemcpyDeviceToHost, c->stream));
  }
  CUDA_TRY(wait_stream(c));
  CUDA_TRY(cudaGetLastError());
  parallel_for(n, [&](int i) { convert_accumulators(*regs[i], h + off[i], &outs[i]); });
  return 0;
}

int gtb_ref_depth_size(gtb_ctx * ctx, int region_id, uint32_t * depth_size, uint32_t * reference_offset)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  auto it = c->regions.find(region_id);
  if (it == c->regions.end())
    return fail(GTB_ERR_STATE, "unknown region");
  if (depth_size)
    *depth_size = it->second->depth_size;
  if (reference_offset)
    *reference_offset = it->second->reference_offset;
  return 0;
}

int gtb_pool_finish(gtb_ctx * ctx, int region_id, gtb_accumulators * out)
{
  return gtb_pool_finish_multi(ctx, 1, &region_id, out);
}

// Zeroes the accumulators of several regions (one call, no synchronisation).
int gtb_pool_reset_multi(gtb_ctx * ctx, int n, const int * region_ids)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || n <= 0 || !region_ids)
    return fail(GTB_ERR_ARG, "bad arguments");
  if (c->device < 0)
    return fail(GTB_ERR_CUDA, "host-only context");
  cudaSetDevice(c->device);
  std::vector<Region *> regs(n);
  for (int i = 0; i < n; ++i)
  {
    auto it = c->regions.find(region_ids[i]);
    if (it == c->regions.end() || !it->second->pool_open)
      return fail(GTB_ERR_STATE, "unknown region / pool not open");
    regs[i] = it->second.get();
    regs[i]->poisoned = regs[i]->reduced = false;
    if (regs[i]->conn_cap)
      CUDA_TRY(cudaMemsetAsync(regs[i]->conn.p, 0, conn_table_bytes(regs[i]->conn_cap), c->stream));
    regs[i]->conn_used = 0;
  }
  if (n <= 2)
  {
    for (int i = 0; i < n; ++i)
      CUDA_TRY(cudaMemsetAsync(regs[i]->accum.p, 0, regs[i]->accum_bytes, c->stream));
    return 0;
  }
  unsigned long long max_bytes = 0; // one launch for all regions instead of one memset each
  std::vector<Segment> const seg = make_segments(n, [&](int i) { return Segment{regs[i]->accum.p, 0, regs[i]->accum_bytes}; }, max_bytes);
  launch_zero_segments(seg.data(), n, max_bytes, c->stream);
  return 0;
}

// ------------------------------------------------------------------------------------------------ phasing connections
int gtb_set_connections(gtb_ctx * ctx, int on)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || on < 0)
    return fail(GTB_ERR_ARG, "bad arguments");
  c->connections = on == 1 ? 16 : on; // table slots budgeted per submitted record
  return 0;
}

// Compacts the pool's table on the device, downloads the entries and sorts them by key (= by sample, hap1, allele1, hap2,
// allele2).  Counts are reported modulo 2^16 like the reference's uint16 counters; entries that are 0 modulo 2^16 are dropped.
static int conn_fetch(Ctx * c, Region & R, std::vector<gtb_connection> & out)
{
  out.clear();
  if (!R.conn_cap)
    return 0;
  cudaSetDevice(c->device);
  uint32_t st[2] = {0, 0};
  CUDA_TRY(cudaMemcpyAsync(st, R.dev.conn_state, 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(wait_stream(c));
  uint32_t const n = st[0];
  if (n == 0)
    return 0;
  DeviceBuffer d;
  if (int rc = take_buffer(c, d, (size_t)n * 12 + 16))
    return rc;
  uint8_t * dp = static_cast<uint8_t *>(d.p);
  auto * d_keys = reinterpret_cast<unsigned long long *>(dp);
  auto * d_vals = reinterpret_cast<uint32_t *>(dp + (size_t)n * 8);
  auto * d_n = reinterpret_cast<uint32_t *>(dp + (size_t)n * 12);
  CUDA_TRY(cudaMemsetAsync(d_n, 0, 4, c->stream));
  launch_conn_compact(R.dev, d_keys, d_vals, d_n, c->stream);
  std::vector<unsigned long long> keys(n);
  std::vector<uint32_t> vals(n);
  CUDA_TRY(cudaMemcpyAsync(keys.data(), d_keys, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(vals.data(), d_vals, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(wait_stream(c));
  CUDA_TRY(cudaGetLastError());
  give_buffer(c, d);
  std::vector<uint32_t> order(n);
  for (uint32_t i = 0; i < n; ++i)
    order[i] = i;
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
  out.reserve(n);
  for (uint32_t i : order)
  {
    uint32_t const cnt = vals[i] & 0xFFFFu;
    if (cnt == 0)
      continue;
    unsigned long long const k = keys[i];
    gtb_connection e;
    e.sample = (uint32_t)((k >> 42) & 0x1FFFFFu);
    e.hap1 = (uint16_t)((k >> 26) & 0xFFFFu);
    e.allele1 = (uint16_t)((k >> 21) & 0x1Fu);
    e.hap2 = (uint16_t)((k >> 5) & 0xFFFFu);
    e.allele2 = (uint16_t)(k & 0x1Fu);
    e.count = cnt;
    out.push_back(e);
  }
  return 0;
}

int gtb_connections_size(gtb_ctx * ctx, int region_id, uint64_t * n)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !n)
    return fail(GTB_ERR_ARG, "bad arguments");
  auto it = c->regions.find(region_id);
  if (it == c->regions.end() || !it->second->pool_open)
    return fail(GTB_ERR_STATE, "unknown region / pool not open");
  if (c->device < 0)
    return fail(GTB_ERR_CUDA, "host-only context");
  std::vector<gtb_connection> v;
  if (int rc = conn_fetch(c, *it->second, v))
    return rc;
  *n = v.size();
  return 0;
}

int gtb_connections(gtb_ctx * ctx, int region_id, gtb_connection * out)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !out)
    return fail(GTB_ERR_ARG, "bad arguments");
  auto it = c->regions.find(region_id);
  if (it == c->regions.end() || !it->second->pool_open)
    return fail(GTB_ERR_STATE, "unknown region / pool not open");
  if (c->device < 0)
    return fail(GTB_ERR_CUDA, "host-only context");
  std::vector<gtb_connection> v;
  if (int rc = conn_fetch(c, *it->second, v))
    return rc;
  if (!v.empty())
    memcpy(out, v.data(), v.size() * sizeof(gtb_connection));
  return 0;
}

// `ph` of parallel_reader_genotype_only (hts_parallel_reader.cpp:782-893).  The connection entries arrive sorted by
// (sample, hap1, allele1, hap2, allele2), so the support vector of (sample, hap1, allele1) -> hap2 is one contiguous run;
// runs are visited once and looked at only when the two bubbles are < 100 bp apart.
int gtb_phase_support(const gtb_accumulators * acc, uint64_t n_conn, const gtb_connection * conn, uint64_t * n_out,
                      gtb_phase_support_entry * out)
{
  if (!acc || !n_out || (n_conn && !conn))
    return fail(GTB_ERR_ARG, "bad arguments");
  uint64_t const NS = acc->n_samples;
  uint32_t const NB = acc->n_bubbles;
  std::map<uint64_t, int8_t> ph; // hap1 << 48 | allele1 << 32 | hap2 << 16 | allele2
  auto seen = [&](uint32_t b, uint64_t s, uint32_t a, bool & clearly, bool & not_seen) {
    uint32_t const num = acc->n_alleles[b];
    const uint16_t * cv = acc->gt_coverage + acc->cov_off[b] * NS + s * num;
    double total = 0.0;
    for (uint32_t k = 0; k < num; ++k)
      total += cv[k];
    double const frac = (double)cv[a] / total;
    clearly = cv[a] >= 4 || frac >= 0.28;
    not_seen = cv[a] <= 2 || frac < 0.22;
  };
  for (uint64_t i = 0; i < n_conn;)
  {
    gtb_connection const & f = conn[i];
    uint64_t j = i;
    long total_support = 0;
    while (j < n_conn && conn[j].sample == f.sample && conn[j].hap1 == f.hap1 && conn[j].allele1 == f.allele1 &&
           conn[j].hap2 == f.hap2)
      total_support += (long)conn[j++].count;
    if (f.hap1 >= NB || f.hap2 >= NB || f.sample >= NS || f.allele1 >= acc->n_alleles[f.hap1])
      return fail(GTB_ERR_ARG, "connection entry out of range");
    bool const near = f.hap2 > f.hap1 && (long long)acc->bubble_id[f.hap2] < (long long)acc->bubble_id[f.hap1] + 100;
    // the reference's ps2 loop stops at the FIRST bubble >= 100 bp away; bubble ids are non-decreasing, so "near" is the same set
    if (near && f.allele1 >= 1)
    {
      bool clearly1, not1;
      seen(f.hap1, f.sample, f.allele1, clearly1, not1);
      uint32_t const num2 = acc->n_alleles[f.hap2];
      uint64_t r = i;
      for (uint32_t a2 = 1; a2 < num2; ++a2)
      {
        while (r < j && conn[r].allele2 < a2)
          ++r;
        double const support = (r < j && conn[r].allele2 == a2) ? (double)conn[r].count : 0.0;
        bool clearly2, not2;
        seen(f.hap2, f.sample, a2, clearly2, not2);
        int8_t flag = 0;
        if (not1 && not2)
          continue;
        if ((not1 && clearly2) || (not2 && clearly1))
          flag = 2; // IS_ANY_ANTI_HAP_SUPPORT
        else if (total_support <= 2)
          continue;
        else if (clearly1 && clearly2 && support / (double)total_support > 0.78)
          flag = 1; // IS_ANY_HAP_SUPPORT
        else if (support / (double)total_support < 0.22)
          flag = 2;
        else
          continue;
        ph[((uint64_t)f.hap1 << 48) | ((uint64_t)f.allele1 << 32) | ((uint64_t)f.hap2 << 16) | a2] |= flag;
      }
    }
    i = j;
  }
  if (out)
  {
    if (*n_out < ph.size())
      return fail(GTB_ERR_ARG, "phase support output too small");
    for (auto const & kv : ph)
    {
      memset(out, 0, sizeof(*out));
      out->hap1 = (uint16_t)(kv.first >> 48);
      out->allele1 = (uint16_t)(kv.first >> 32);
      out->hap2 = (uint16_t)(kv.first >> 16);
      out->allele2 = (uint16_t)kv.first;
      out->flags = kv.second;
      ++out;
    }
  }
  *n_out = ph.size();
  return 0;
}

// Two sorted connection lists of the same pool (e.g. from two ranks that each saw a share of the reads) -> one.
int gtb_merge_connections(uint64_t n_a, const gtb_connection * a, uint64_t n_b, const gtb_connection * b, uint64_t * n_out,
                          gtb_connection * out)
{
  if (!n_out || (n_a && !a) || (n_b && !b))
    return fail(GTB_ERR_ARG, "bad arguments");
  auto key = [](const gtb_connection & c) {
    return std::make_tuple(c.sample, c.hap1, c.allele1, c.hap2, c.allele2);
  };
  uint64_t i = 0, j = 0, n = 0;
  uint64_t const cap = *n_out;
  auto emit = [&](gtb_connection c) -> bool {
    c.count &= 0xFFFFu;
    if (c.count == 0)
      return true;
    if (out)
    {
      if (n >= cap)
        return false;
      out[n] = c;
    }
    ++n;
    return true;
  };
  bool ok = true;
  while (ok && (i < n_a || j < n_b))
  {
    if (j >= n_b || (i < n_a && key(a[i]) < key(b[j])))
      ok = emit(a[i++]);
    else if (i >= n_a || key(b[j]) < key(a[i]))
      ok = emit(b[j++]);
    else
    {
      gtb_connection c = a[i++];
      c.count += b[j++].count;
      ok = emit(c);
    }
  }
  if (!ok)
    return fail(GTB_ERR_ARG, "merged connection output too small");
  *n_out = n;
  return 0;
}

// ------------------------------------------------------------------------------------------------ debug taps
int gtb_debug_enable(gtb_ctx * ctx, int on)
{
  reinterpret_cast<Ctx *>(ctx)->debug = on != 0;
  return 0;
}

// debug taps look at chunk 0 (a submit made with gtb_debug_enable is never chunked)
static int find_last_region(Ctx * c, int region_id)
{
  if (c->n_chunks_last != 1)
    return -1;
  for (size_t i = 0; i < c->bs[0].regions.size(); ++i)
    if (c->bs[0].regions[i] == region_id)
      return (int)i;
  return -1;
}

int gtb_debug_seed_sizes(gtb_ctx * ctx, int region_id, uint64_t * n_units, uint64_t * n_slots, uint64_t * n_labels)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  int const k = c->have_last ? find_last_region(c, region_id) : -1;
  if (k < 0 || !c->bs[0].P.tap.list_count)
    return fail(GTB_ERR_STATE, "no debug tap for this region (gtb_debug_enable before submit)");
  cudaSetDevice(c->device);
  uint32_t const u0 = c->bs[0].unit_begin[k], u1 = c->bs[0].unit_begin[k + 1];
  size_t const nt = (size_t)(u1 - u0) * 2;
  std::vector<uint32_t> counts(nt * NLISTS), nsl(nt);
  CUDA_TRY(cudaMemcpy(counts.data(), c->bs[0].P.tap.list_count + (size_t)u0 * 2 * NLISTS, counts.size() * 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(nsl.data(), c->bs[0].P.tap.nslots + (size_t)u0 * 2, nsl.size() * 4, cudaMemcpyDeviceToHost));
  uint64_t ns = 0, nl = 0;
  for (size_t t = 0; t < nt; ++t)
  {
    ns += (uint64_t)nsl[t] * 2;
    for (uint32_t l = 0; l < nsl[t] * 2; ++l)
      nl += counts[t * NLISTS + l];
  }
  *n_units = u1 - u0;
  *n_slots = ns;
  *n_labels = nl;
  return 0;
}

int gtb_debug_seeds(gtb_ctx * ctx, int region_id, uint32_t * unit_record, uint32_t * nslots, uint32_t * nlabels,
                    gtb_label * labels)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  int const k = c->have_last ? find_last_region(c, region_id) : -1;
  if (k < 0 || !c->bs[0].P.tap.list_count)
    return fail(GTB_ERR_STATE, "no debug tap for this region");
  cudaSetDevice(c->device);
  uint32_t const u0 = c->bs[0].unit_begin[k], u1 = c->bs[0].unit_begin[k + 1];
  size_t const nu = u1 - u0, nt = nu * 2;
  std::vector<uint32_t> counts(nt * NLISTS), offs(nt * NLISTS), nsl(nt);
  std::vector<int32_t> urec(nu);
  CUDA_TRY(cudaMemcpy(counts.data(), c->bs[0].P.tap.list_count + (size_t)u0 * 2 * NLISTS, counts.size() * 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(offs.data(), c->bs[0].P.tap.list_off + (size_t)u0 * 2 * NLISTS, offs.size() * 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(nsl.data(), c->bs[0].P.tap.nslots + (size_t)u0 * 2, nsl.size() * 4, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(urec.data(), c->bs[0].P.batch.unit_record + u0, nu * 4, cudaMemcpyDeviceToHost));
  DevCounters k_host;
  CUDA_TRY(cudaMemcpy(&k_host, c->bs[0].d_counters.p, sizeof(k_host), cudaMemcpyDeviceToHost));
  std::vector<DevLabel> pool(k_host.dbg_label_words);
  if (!pool.empty())
    CUDA_TRY(cudaMemcpy(pool.data(), c->bs[0].P.tap.pool, pool.size() * sizeof(DevLabel), cudaMemcpyDeviceToHost));
  size_t si = 0, li = 0;
  for (size_t u = 0; u < nu; ++u)
  {
    unit_record[u] = (uint32_t)(urec[u] - (int32_t)c->bs[0].rec_begin[k]);
    for (int o = 0; o < 2; ++o)
    {
      size_t const t = u * 2 + o;
      for (int hm = 0; hm < 2; ++hm)
      {
        nslots[(u * 2 + o) * 2 + hm] = nsl[t];
        for (uint32_t s = 0; s < nsl[t]; ++s)
        {
          uint32_t const l = s * 2 + hm;
          uint32_t const cnt = counts[t * NLISTS + l];
          nlabels[si++] = cnt;
          for (uint32_t q = 0; q < cnt; ++q)
          {
            DevLabel const & d = pool[offs[t * NLISTS + l] + q];
            labels[li].start = d.start;
            labels[li].end = d.end;
            labels[li].var_id = d.var;
            ++li;
          }
        }
      }
    }
  }
  return 0;
}

static int fetch_paths(Ctx * c, int k, std::vector<TaskSummary> & sums, std::vector<uint32_t> & pool)
{
  uint32_t const u0 = c->bs[0].unit_begin[k], u1 = c->bs[0].unit_begin[k + 1];
  sums.resize((size_t)(u1 - u0) * 2);
  if (!sums.empty())
    CUDA_TRY(cudaMemcpy(sums.data(), c->bs[0].P.summaries + (size_t)u0 * 2, sums.size() * sizeof(TaskSummary), cudaMemcpyDeviceToHost));
  DevCounters k_host;
  CUDA_TRY(cudaMemcpy(&k_host, c->bs[0].d_counters.p, sizeof(k_host), cudaMemcpyDeviceToHost));
  size_t const words = (size_t)c->bs[0].n_tasks * INLINE_WORDS + k_host.path_words;
  pool.resize(std::min<size_t>(words, c->bs[0].P.path_pool_cap));
  if (!pool.empty())
    CUDA_TRY(cudaMemcpy(pool.data(), c->bs[0].P.path_pool, pool.size() * 4, cudaMemcpyDeviceToHost));
  return 0;
}

int gtb_debug_path_sizes(gtb_ctx * ctx, int region_id, uint64_t * n_units, uint64_t * n_paths, uint64_t * n_vars,
                         uint64_t * n_nums)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  int const k = c->have_last ? find_last_region(c, region_id) : -1;
  if (k < 0)
    return fail(GTB_ERR_STATE, "region was not part of the last submit");
  cudaSetDevice(c->device);
  std::vector<TaskSummary> sums;
  std::vector<uint32_t> pool;
  if (int rc = fetch_paths(c, k, sums, pool))
    return rc;
  uint64_t np = 0, nv = 0, nn = 0;
  for (auto const & s : sums)
  {
    const uint32_t * w = pool.data() + s.path_off;
    for (uint32_t p = 0; p < s.npaths; ++p)
    {
      uint32_t const nvar = w[3] >> 16;
      ++np;
      nv += nvar;
      for (uint32_t q = 0; q < nvar; ++q)
        nn += (uint64_t)__builtin_popcount(w[4 + 2 * q + 1]);
      w += 4 + 2 * nvar;
    }
  }
  *n_units = sums.size() / 2;
  *n_paths = np;
  *n_vars = nv;
  *n_nums = nn;
  return 0;
}

int gtb_debug_paths(gtb_ctx * ctx, int region_id, uint32_t * gp_npaths, uint32_t * gp_longest, uint32_t * p_fields,
                    uint32_t * v_order, uint32_t * v_nnum, uint16_t * v_nums)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  int const k = c->have_last ? find_last_region(c, region_id) : -1;
  if (k < 0)
    return fail(GTB_ERR_STATE, "region was not part of the last submit");
  cudaSetDevice(c->device);
  std::vector<TaskSummary> sums;
  std::vector<uint32_t> pool;
  if (int rc = fetch_paths(c, k, sums, pool))
    return rc;
  size_t pi = 0, vi = 0, ni = 0;
  for (size_t t = 0; t < sums.size(); ++t)
  {
    TaskSummary const & s = sums[t];
    gp_npaths[t] = s.npaths;
    gp_longest[t] = s.longest;
    const uint32_t * w = pool.data() + s.path_off;
    for (uint32_t p = 0; p < s.npaths; ++p)
    {
      uint32_t const nvar = w[3] >> 16;
      uint32_t * f = p_fields + pi * 6;
      f[0] = w[0];
      f[1] = w[1];
      f[2] = w[2] & 0xFFFFu;
      f[3] = w[2] >> 16;
      f[4] = w[3] & 0xFFFFu;
      f[5] = nvar;
      ++pi;
      for (uint32_t q = 0; q < nvar; ++q)
      {
        v_order[vi] = w[4 + 2 * q];
        uint32_t const mask = w[4 + 2 * q + 1];
        v_nnum[vi] = (uint32_t)__builtin_popcount(mask);
        ++vi;
        for (int a = 0; a < 32; ++a)
          if ((mask >> a) & 1u)
            v_nums[ni++] = (uint16_t)a;
      }
      w += 4 + 2 * nvar;
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ host finalisation
// get_haplotype_phred (src/typer/vcf.cpp:47-81) + SampleCall::get_gt_call / get_gq (src/typer/sample_call.cpp:78-131)
int gtb_calls_from_accumulators(const gtb_accumulators * acc, uint8_t * phred, uint16_t * gt, uint8_t * gq)
{
  if (!acc || !phred || !gt || !gq)
    return fail(GTB_ERR_ARG, "null argument");
  uint32_t const NB = acc->n_bubbles, NS = acc->n_samples;
  for (uint32_t b = 0; b < NB; ++b)
  {
    uint64_t const tri = acc->score_off[b + 1] - acc->score_off[b];
    uint32_t const cnum = acc->n_alleles[b];
    for (uint32_t s = 0; s < NS; ++s)
    {
      const uint16_t * ls = acc->log_score + acc->score_off[b] * NS + (uint64_t)s * tri;
      uint8_t * ph = phred + acc->score_off[b] * NS + (uint64_t)s * tri;
      uint16_t mx = 0;
      bool all_same = true;
      for (uint64_t i = 0; i < tri; ++i)
        mx = std::max(mx, ls[i]);
      for (uint64_t i = 0; i < tri; ++i)
        all_same = all_same && ls[i] == mx;
      int zeros = 0;
      uint8_t second = 255;
      bool have_gt = false;
      uint64_t i = 0;
      for (uint32_t y = 0; y < cnum; ++y)
        for (uint32_t x = 0; x <= y; ++x, ++i)
        {
          uint8_t p = 0;
          if (!all_same)
          {
            long long const sc = std::llround((double)(mx - ls[i]) * 3.01029995663981195213738894724493026768189881462108541);
            p = sc < 255 ? (uint8_t)sc : (uint8_t)255;
          }
          ph[i] = p;
          if (p == 0)
          {
            ++zeros;
            if (!have_gt)
            {
              gt[((uint64_t)b * NS + s) * 2 + 0] = (uint16_t)x;
              gt[((uint64_t)b * NS + s) * 2 + 1] = (uint16_t)y;
              have_gt = true;
            }
          }
          else if (p < second)
            second = p;
        }
      gq[(uint64_t)b * NS + s] = zeros >= 2 ? 0 : second;
    }
  }
  return 0;
}

// Variant::scan_calls (src/typer/variant.cpp:230-428) for every bubble of one pool (non-SV graphs)
int gtb_sample_depths(const gtb_accumulators * acc, uint16_t * ref_total_depth, uint16_t * alt_total_depth)
{
  if (!acc || !ref_total_depth || !alt_total_depth)
    return fail(GTB_ERR_ARG, "bad arguments");
  uint64_t const NS = acc->n_samples;
  for (uint32_t b = 0; b < acc->n_bubbles; ++b)
  {
    uint32_t const cnum = acc->n_alleles[b];
    for (uint64_t s = 0; s < NS; ++s)
    {
      const uint16_t * cov = acc->gt_coverage + acc->cov_off[b] * NS + s * cnum;
      uint64_t const i = (uint64_t)b * NS + s;
      uint32_t const amb = acc->ambiguous_depth[i], amb_alt = acc->ambiguous_depth_alt[i];
      uint32_t alt = amb;
      for (uint32_t a = 1; a < cnum; ++a)
        alt += cov[a];
      ref_total_depth[i] = (uint16_t)std::min<uint32_t>(0xFFFFu, (uint32_t)cov[0] + amb - amb_alt);
      alt_total_depth[i] = (uint16_t)std::min<uint32_t>(0xFFFFu, alt);
    }
  }
  return 0;
}

int gtb_scan_calls(const gtb_accumulators * acc, const uint8_t * phred, uint64_t * var, uint64_t * allele, double * ratio)
{
  if (!acc || !phred || !var || !allele || !ratio)
    return fail(GTB_ERR_ARG, "null argument");
  uint32_t const NB = acc->n_bubbles, NS = acc->n_samples;
  memset(var, 0, (size_t)NB * 9 * 8);
  memset(allele, 0, (size_t)acc->cov_off[NB] * 13 * 8);
  for (uint64_t i = 0; i < acc->cov_off[NB]; ++i)
    ratio[i] = 0.0;
  for (uint32_t b = 0; b < NB; ++b)
  {
    uint64_t * V = var + (size_t)b * 9;
    uint64_t const c0 = acc->cov_off[b];
    uint32_t const cnum = acc->n_alleles[b];
    uint64_t const tri = acc->score_off[b + 1] - acc->score_off[b];
    auto A = [&](uint32_t a) { return allele + (c0 + a) * 13; };
    V[1] += NS; // n_calls
    for (uint32_t s = 0; s < NS; ++s)
    {
      const uint8_t * ph = phred + acc->score_off[b] * NS + (uint64_t)s * tri;
      const uint16_t * cov = acc->gt_coverage + c0 * NS + (uint64_t)s * cnum;
      uint8_t const amb = acc->ambiguous_depth[(uint64_t)b * NS + s];
      uint8_t const altpp = acc->alt_proper_pair_depth[(uint64_t)b * NS + s];
      // get_gt_call / get_gq (sample_call.cpp:78-131)
      uint32_t g1 = 0, g2 = 0;
      {
        uint64_t i = 0;
        bool done = false;
        for (uint32_t y = 0; y < cnum && !done; ++y)
          for (uint32_t x = 0; x <= y; ++x, ++i)
            if (ph[i] == 0)
            {
              g1 = x;
              g2 = y;
              done = true;
              break;
            }
      }
      long gq;
      {
        bool seen = false, two = false;
        uint8_t nl = 255;
        for (uint64_t i = 0; i < tri; ++i)
        {
          if (ph[i] == 0)
          {
            if (seen)
            {
              two = true;
              break;
            }
            seen = true;
          }
          else if (ph[i] < nl)
            nl = ph[i];
        }
        gq = two ? 0 : nl;
      }
      auto lowest_phred_not_with = [&](uint32_t al) // sample_call.cpp:131-154
      {
        long i = 0;
        uint8_t mn = 255;
        for (long y = 0; y < (long)cnum; ++y)
        {
          if (y == (long)al)
          {
            i += y + 1;
            continue;
          }
          for (long x = 0; x <= y; ++x, ++i)
          {
            if (x == (long)al)
              continue;
            if (ph[i] < mn)
              mn = ph[i];
          }
        }
        return mn;
      };
      if (tri > 0 && ph[0] > 0) // not homozygous reference
      {
        auto qd = [&](uint32_t al)
        {
          long const depth = std::min(10l, (long)cov[al] + (long)amb);
          if (depth > 0)
          {
            A(al)[0] += (uint64_t)std::min(25l * depth, (long)lowest_phred_not_with(al));
            A(al)[1] += (uint64_t)depth;
          }
        };
        if (g1 > 0)
          qd(g1);
        if (g1 != g2)
          qd(g2);
      }
      V[3] = std::max<uint64_t>(V[3], altpp); // MaxAltPP within the pool
      uint64_t total_depth = 0;
      for (uint32_t a = 0; a < cnum; ++a)
        total_depth += cov[a];
      for (uint32_t a = 1; a < cnum; ++a)
      {
        uint64_t * P = A(a);
        P[8] = std::max<uint64_t>(P[8], cov[a]);
        if (total_depth > 0)
          ratio[c0 + a] = std::max(ratio[c0 + a], (double)cov[a] / (double)total_depth);
        if (g1 == a || g2 == a)
        {
          if (g1 == g2)
            ++P[7];
          else
            ++P[6];
        }
        else
          ++P[5];
      }
      long const filter = gq >= 30 ? 0 : gq >= 20 ? 1 : gq >= 10 ? 2 : 3; // check_filter (sample_call.cpp:156-170)
      bool nonzero = false;
      for (uint64_t i = 0; i < tri; ++i)
        nonzero = nonzero || ph[i] != 0;
      if (nonzero)
        ++V[0];
      if (filter == 0)
        ++V[2];
      if (g1 != g2)
      {
        V[5] += cov[g1];
        V[6] += cov[g2];
        A(g1)[9] += cov[g1];
        A(g1)[10] += total_depth - cov[g1];
        A(g2)[9] += cov[g2];
        A(g2)[10] += total_depth - cov[g2];
      }
      else
      {
        V[7] += cov[g1];
        V[8] += total_depth - cov[g1];
        A(g1)[11] += cov[g1];
        A(g1)[12] += total_depth - cov[g1];
      }
      V[4] += total_depth + amb; // seqdepth += get_depth()
      for (uint32_t a = 1; a < cnum; ++a)
        A(a)[2] += cov[a];
      ++A(g1)[3];
      ++A(g2)[3];
      if (filter == 0)
      {
        ++A(g1)[4];
        ++A(g2)[4];
      }
    }
  }
  return 0;
}

// VarStats::add_stats (src/typer/var_stats.cpp:141-189) on the arrays of gtb_scan_calls: dst += src
// Per-pool summaries of several regions in one call (the regions are independent jobs for the host pool): PHRED calls, then
// Variant::scan_calls; rows of region i start at var + 9 * (bubbles before it) / allele + 13 * (alleles before it).
int gtb_scan_calls_multi(int n, const gtb_accumulators * accs, uint64_t * var, uint64_t * allele, double * ratio)
{
  if (n <= 0 || !accs || !var || !allele || !ratio)
    return fail(GTB_ERR_ARG, "bad arguments");
  std::vector<size_t> b0(n + 1, 0), a0(n + 1, 0);
  for (int i = 0; i < n; ++i)
  {
    b0[i + 1] = b0[i] + accs[i].n_bubbles;
    a0[i + 1] = a0[i] + accs[i].cov_off[accs[i].n_bubbles];
  }
  std::vector<int> rcs(n, 0);
  parallel_for(n, [&](int i)
               {
                 gtb_accumulators const & A = accs[i];
                 size_t const NS = A.n_samples, NB = A.n_bubbles;
                 std::vector<uint8_t> phred((size_t)A.score_off[NB] * NS), gq(NB * NS);
                 std::vector<uint16_t> gt(NB * NS * 2);
                 rcs[i] = gtb_calls_from_accumulators(&A, phred.data(), gt.data(), gq.data());
                 if (!rcs[i])
                   rcs[i] = gtb_scan_calls(&A, phred.data(), var + 9 * b0[i], allele + 13 * a0[i], ratio + a0[i]);
               });
  for (int i = 0; i < n; ++i)
    if (rcs[i])
      return fail(rcs[i], "gtb_scan_calls_multi: a region failed");
  return 0;
}

int gtb_merge_varstats(uint32_t n_bubbles, uint64_t n_alleles_total, uint64_t * var, uint64_t * allele, double * ratio,
                       const uint64_t * var_src, const uint64_t * allele_src, const double * ratio_src)
{
  for (uint32_t b = 0; b < n_bubbles; ++b)
    for (int k = 0; k < 9; ++k)
    {
      uint64_t & d = var[(size_t)b * 9 + k];
      d += var_src[(size_t)b * 9 + k];
      if (k == 3)
        d &= 0xFFu; // uint8_t n_max_alt_proper_pairs is summed (not max-ed) in the reference and wraps
    }
  for (uint64_t a = 0; a < n_alleles_total; ++a)
  {
    for (int k = 0; k < 13; ++k)
    {
      uint64_t & d = allele[a * 13 + k];
      uint64_t const sv = allele_src[a * 13 + k];
      if (k == 8)
        d = std::max(d, sv);
      else
        d += sv;
    }
    ratio[a] = std::max(ratio[a], ratio_src[a]);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ multi-GPU reduce
// NCCL is bound lazily (dlopen) so the library loads on hosts without it and never clashes with the NCCL copy
// PyTorch bundles.  The communicator is created here from a unique id the caller distributes (e.g. with
// torch.distributed.broadcast): gtb_nccl_unique_id on rank 0 -> gtb_nccl_init on every rank.
typedef struct
{
  char internal[128];
} gtb_nccl_id;

namespace
{
struct NcclApi
{
  int (*GetUniqueId)(gtb_nccl_id *) = nullptr;
  int (*CommInitRank)(void **, int, gtb_nccl_id, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char * (*GetErrorString)(int) = nullptr;
  void * lib = nullptr;
} g_nccl;

int load_nccl()
{
  if (g_nccl.lib)
    return 0;
  const char * names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char * n : names)
  {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (g_nccl.lib)
      break;
  }
  if (!g_nccl.lib)
    return fail(GTB_ERR_NCCL, std::string("cannot dlopen libnccl: ") + dlerror());
  g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(g_nccl.lib, "ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(g_nccl.lib, "ncclCommInitRank"));
  g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(g_nccl.lib, "ncclAllReduce"));
  g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(g_nccl.lib, "ncclCommDestroy"));
  g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(g_nccl.lib, "ncclGetErrorString"));
  g_nccl.GroupStart = reinterpret_cast<decltype(g_nccl.GroupStart)>(dlsym(g_nccl.lib, "ncclGroupStart"));
  g_nccl.GroupEnd = reinterpret_cast<decltype(g_nccl.GroupEnd)>(dlsym(g_nccl.lib, "ncclGroupEnd"));
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
    return fail(GTB_ERR_NCCL, "libnccl lacks required symbols");
  return 0;
}
} // namespace

int gtb_nccl_unique_id(uint8_t * id128)
{
  if (int rc = load_nccl())
    return rc;
  gtb_nccl_id id;
  int const r = g_nccl.GetUniqueId(&id);
  if (r != 0)
    return fail(GTB_ERR_NCCL, std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
  memcpy(id128, id.internal, 128);
  return 0;
}

int gtb_nccl_init(gtb_ctx * ctx, int n_ranks, int rank, const uint8_t * id128)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (c->device < 0)
    return fail(GTB_ERR_CUDA, "host-only context");
  if (int rc = load_nccl())
    return rc;
  cudaSetDevice(c->device);
  gtb_nccl_id id;
  memcpy(id.internal, id128, 128);
  int const r = g_nccl.CommInitRank(&c->nccl_comm, n_ranks, id, rank);
  if (r != 0)
    return fail(GTB_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
  c->nccl_rank = rank;
  c->nccl_size = n_ranks;
  return 0;
}

// Sum-reduce of the widened accumulators of one region over all ranks (additive: SURVEY.md section 8e).
// The accumulator arena is laid out as uint32 [..] then uint64 [..] then uint32 read_strand; it is reduced as
// three typed spans.  nccl_comm == NULL uses the communicator of gtb_nccl_init.
int gtb_allreduce_accumulators_multi(gtb_ctx * ctx, int n, const int * region_ids, void * nccl_comm)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || n <= 0 || !region_ids)
    return fail(GTB_ERR_ARG, "bad arguments");
  void * comm = nccl_comm ? nccl_comm : c->nccl_comm;
  if (!comm)
    return fail(GTB_ERR_STATE, "no NCCL communicator (gtb_nccl_init)");
  if (int rc = load_nccl())
    return rc;
  cudaSetDevice(c->device);
  std::vector<Region *> regs(n);
  for (int i = 0; i < n; ++i)
  {
    auto it = c->regions.find(region_ids[i]);
    if (it == c->regions.end() || !it->second->pool_open)
      return fail(GTB_ERR_STATE, "unknown region / pool not open");
    if (it->second->reduced)
      return fail(GTB_ERR_STATE, "the pool's accumulators are already all-reduced (a second reduce would count every rank again)");
    regs[i] = it->second.get();
  }
  // one NCCL group for all regions: three typed spans per accumulator arena (u32 | u64 | u32 read_strand)
  int r = g_nccl.GroupStart ? g_nccl.GroupStart() : 0;
  for (int i = 0; i < n && r == 0; ++i)
  {
    Region & R = *regs[i];
    uint8_t * base = static_cast<uint8_t *>(R.accum.p);
    uint8_t * u64_begin = reinterpret_cast<uint8_t *>(R.dev.vs_clipped_reads);
    uint8_t * rs_begin = reinterpret_cast<uint8_t *>(R.dev.read_strand);
    uint8_t * end = base + R.accum_bytes;
    // ncclUint32 = 3, ncclUint64 = 5, ncclSum = 0
    r = g_nccl.AllReduce(base, base, (size_t)(u64_begin - base) / 4, 3, 0, comm, c->stream);
    if (r == 0)
      r = g_nccl.AllReduce(u64_begin, u64_begin, (size_t)(rs_begin - u64_begin) / 8, 5, 0, comm, c->stream);
    if (r == 0)
      r = g_nccl.AllReduce(rs_begin, rs_begin, (size_t)(end - rs_begin) / 4, 3, 0, comm, c->stream);
  }
  if (g_nccl.GroupEnd)
  {
    int const r2 = g_nccl.GroupEnd();
    if (r == 0)
      r = r2;
  }
  if (r != 0)
    return fail(GTB_ERR_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
  CUDA_TRY(wait_stream(c));
  for (Region * R : regs)
    R->reduced = true;
  return 0;
}

int gtb_allreduce_accumulators(gtb_ctx * ctx, int region_id, void * nccl_comm)
{
  return gtb_allreduce_accumulators_multi(ctx, 1, &region_id, nccl_comm);
}

// Cross-pool merge of the per-bubble summaries over the ranks of a sample-sharded run = VarStats::add_stats
// (src/typer/var_stats.cpp:141-189) as three collectives in ONE NCCL group: sum over every counter, max over
// maximum_alt_support (column 8 of the allele rows) and over maximum_alt_support_ratio; n_max_alt_proper_pairs (column 3 of
// the var rows) is summed as the reference's uint8 does.  The arrays may cover any number of regions back to back.
int gtb_allreduce_varstats_multi(gtb_ctx * ctx, int n_src, uint64_t n_var_rows, uint64_t n_allele_rows, uint64_t * const * var,
                                 uint64_t * const * allele, double * const * ratio, void * nccl_comm)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || n_src <= 0 || !var || !allele || !ratio)
    return fail(GTB_ERR_ARG, "bad arguments");
  if (c->device < 0)
    return fail(GTB_ERR_CUDA, "host-only context");
  void * comm = nccl_comm ? nccl_comm : c->nccl_comm;
  if (!comm)
    return fail(GTB_ERR_STATE, "no NCCL communicator (gtb_nccl_init)");
  if (int rc = load_nccl())
    return rc;
  cudaSetDevice(c->device);
  size_t const nv = (size_t)n_var_rows * 9, na = (size_t)n_allele_rows * 13;
  size_t const n_sum = nv + na, n_max = (size_t)n_allele_rows;
  size_t const bytes = (n_sum + 2 * n_max) * 8;
  if (int rc = c->h_varstats.reserve(bytes))
    return rc;
  if (int rc = c->d_varstats.reserve(bytes))
    return rc;
  uint64_t * h = static_cast<uint64_t *>(c->h_varstats.p);
  uint64_t * hmax = h + n_sum;
  double * hratio = reinterpret_cast<double *>(hmax + n_max);
  // pack = the cross-pool merge of this rank's pools (VarStats::add_stats over the sources), in parallel blocks
  {
    constexpr size_t BLK = 1 << 15;
    size_t const n_blk_v = (nv + BLK - 1) / BLK, n_blk_a = (n_max + (BLK / 16) - 1) / (BLK / 16);
    parallel_for((int)(n_blk_v + n_blk_a), [&](int bi)
                 {
                   if ((size_t)bi < n_blk_v)
                   {
                     size_t const lo = (size_t)bi * BLK, hi = std::min(nv, lo + BLK);
                     for (size_t i = lo; i < hi; ++i)
                     {
                       uint64_t v = 0;
                       for (int k = 0; k < n_src; ++k)
                         v += var[k][i];
                       h[i] = v;
                     }
                     return;
                   }
                   size_t const lo = ((size_t)bi - n_blk_v) * (BLK / 16), hi = std::min(n_max, lo + BLK / 16);
                   for (size_t a = lo; a < hi; ++a)
                   {
                     uint64_t mx = 0;
                     double mr = 0.0;
                     for (int col = 0; col < 13; ++col)
                     {
                       uint64_t v = 0;
                       for (int k = 0; k < n_src; ++k)
                         v += allele[k][a * 13 + col];
                       h[nv + a * 13 + col] = v;
                     }
                     for (int k = 0; k < n_src; ++k)
                     {
                       mx = std::max(mx, allele[k][a * 13 + 8]);
                       mr = std::max(mr, ratio[k][a]);
                     }
                     h[nv + a * 13 + 8] = 0; // maximum_alt_support travels in the max block
                     hmax[a] = mx;
                     hratio[a] = mr;
                   }
                 });
  }
  uint64_t * d = static_cast<uint64_t *>(c->d_varstats.p);
  CUDA_TRY(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream));
  // ncclUint64 = 5, ncclFloat64 = 8, ncclSum = 0, ncclMax = 2
  int r = g_nccl.GroupStart ? g_nccl.GroupStart() : 0;
  if (r == 0 && n_sum)
    r = g_nccl.AllReduce(d, d, n_sum, 5, 0, comm, c->stream);
  if (r == 0 && n_max)
    r = g_nccl.AllReduce(d + n_sum, d + n_sum, n_max, 5, 2, comm, c->stream);
  if (r == 0 && n_max)
    r = g_nccl.AllReduce(d + n_sum + n_max, d + n_sum + n_max, n_max, 8, 2, comm, c->stream);
  if (g_nccl.GroupEnd)
  {
    int const r2 = g_nccl.GroupEnd();
    if (r == 0)
      r = r2;
  }
  if (r != 0)
    return fail(GTB_ERR_NCCL, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
  CUDA_TRY(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(wait_stream(c));
  memcpy(var[0], h, nv * 8);
  memcpy(allele[0], h + nv, na * 8);
  for (size_t b2 = 0; b2 < n_var_rows; ++b2)
    var[0][b2 * 9 + 3] &= 0xFFu; // n_max_alt_proper_pairs is a uint8 the reference sums
  for (size_t a = 0; a < n_max; ++a)
  {
    allele[0][a * 13 + 8] = hmax[a];
    ratio[0][a] = hratio[a];
  }
  return 0;
}

int gtb_allreduce_varstats(gtb_ctx * ctx, uint64_t n_var_rows, uint64_t n_allele_rows, uint64_t * var, uint64_t * allele,
                           double * ratio, void * nccl_comm)
{
  return gtb_allreduce_varstats_multi(ctx, 1, n_var_rows, n_allele_rows, &var, &allele, &ratio, nccl_comm);
}

// ---------------------------------------------------------------------------------------------------------------------
// Discovery re-alignment: batch of (read, haplotype window) pairs through sw_kernel (gtb_sw.cu).
int gtb_sw_align_batch(gtb_ctx * ctx, int n_pairs, const uint8_t * query, const int32_t * q_off,
                       const uint8_t * database, const int32_t * d_off, gtb_sw_result * out)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || n_pairs < 0 || (n_pairs > 0 && (!query || !q_off || !database || !d_off || !out)))
    return fail(GTB_ERR_ARG, "gtb_sw_align_batch: null argument");
  if (n_pairs == 0)
    return 0;
  int max_db = 0;
  for (int k = 0; k < n_pairs; ++k)
  {
    int const m = q_off[k + 1] - q_off[k], n = d_off[k + 1] - d_off[k];
    if (m < 1 || m > GTB_SW_MAX_QUERY)
      return fail(GTB_ERR_ARG, "gtb_sw_align_batch: query length must be 1.." + std::to_string(GTB_SW_MAX_QUERY));
    if (n < 1 || n > GTB_SW_MAX_DATABASE)
      return fail(GTB_ERR_ARG, "gtb_sw_align_batch: database length must be 1.." + std::to_string(GTB_SW_MAX_DATABASE));
    max_db = std::max(max_db, n);
  }
  if (q_off[0] != 0 || d_off[0] != 0)
    return fail(GTB_ERR_ARG, "gtb_sw_align_batch: offsets must start at 0");
  if (c->device < 0)
    return fail(GTB_ERR_CUDA, "host-only context: the re-alignment kernel needs a CUDA device (no CPU fallback)");
  cudaSetDevice(c->device);
  for (auto & e : c->sw_ev)
    if (!e)
      CUDA_TRY(cudaEventCreate(&e));
  size_t const qb = (size_t)q_off[n_pairs], db = (size_t)d_off[n_pairs], ob = (size_t)(n_pairs + 1) * 4;
  size_t const o_q = 0, o_d = align_up(o_q + qb), o_qo = align_up(o_d + db), o_do = align_up(o_qo + ob);
  size_t const in_bytes = align_up(o_do + ob);
  max_db = (max_db + 63) / 64 * 64;
  c->sw_warps = sw_resident_warps(max_db);
  if (int rc = c->d_sw_in.reserve(in_bytes))
    return rc;
  if (int rc = c->d_sw_out.reserve((size_t)n_pairs * sizeof(gtb_sw_result)))
    return rc;
  if (int rc = c->d_sw_bt.reserve((size_t)c->sw_warps * (size_t)(max_db + 5) * 128))
    return rc;
  size_t const out_bytes = (size_t)n_pairs * sizeof(gtb_sw_result);
  if (int rc = c->h_sw.reserve(std::max(in_bytes, out_bytes)))
    return rc;
  uint8_t * h = static_cast<uint8_t *>(c->h_sw.p);
  memcpy(h + o_q, query, qb);
  memcpy(h + o_d, database, db);
  memcpy(h + o_qo, q_off, ob);
  memcpy(h + o_do, d_off, ob);
  uint8_t * dv = static_cast<uint8_t *>(c->d_sw_in.p);
  CUDA_TRY(cudaEventRecord(c->sw_ev[0], c->stream));
  CUDA_TRY(cudaMemcpyAsync(dv, h, in_bytes, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaEventRecord(c->sw_ev[1], c->stream));
  SwParams P{};
  P.n_pairs = n_pairs;
  P.max_db = max_db;
  P.q = dv + o_q;
  P.d = dv + o_d;
  P.q_off = reinterpret_cast<const int32_t *>(dv + o_qo);
  P.d_off = reinterpret_cast<const int32_t *>(dv + o_do);
  P.out = static_cast<gtb_sw_result *>(c->d_sw_out.p);
  P.bt = static_cast<uint32_t *>(c->d_sw_bt.p);
  launch_sw(P, c->sw_warps, c->stream);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(c->sw_ev[2], c->stream));
  // the input staging area is reused for the results once the H2D copy has been consumed (same stream => ordered)
  CUDA_TRY(cudaMemcpyAsync(h, c->d_sw_out.p, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaEventRecord(c->sw_ev[3], c->stream));
  CUDA_TRY(wait_stream(c));
  memcpy(out, h, out_bytes);
  cudaEventElapsedTime(&c->t_sw_h2d, c->sw_ev[0], c->sw_ev[1]);
  cudaEventElapsedTime(&c->t_sw_kernel, c->sw_ev[1], c->sw_ev[2]);
  cudaEventElapsedTime(&c->t_sw_d2h, c->sw_ev[2], c->sw_ev[3]);
  c->sw_last = P;
  return 0;
}

int gtb_sw_replay_last(gtb_ctx * ctx)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c || c->sw_last.n_pairs <= 0)
    return fail(GTB_ERR_STATE, "gtb_sw_replay_last: no previous gtb_sw_align_batch");
  cudaSetDevice(c->device);
  CUDA_TRY(cudaEventRecord(c->sw_ev[1], c->stream));
  launch_sw(c->sw_last, c->sw_warps, c->stream);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(c->sw_ev[2], c->stream));
  CUDA_TRY(wait_stream(c));
  cudaEventElapsedTime(&c->t_sw_kernel, c->sw_ev[1], c->sw_ev[2]);
  return 0;
}

int gtb_sw_last_timing(gtb_ctx * ctx, float * kernel_ms, float * h2d_ms, float * d2h_ms)
{
  auto * c = reinterpret_cast<Ctx *>(ctx);
  if (!c)
    return fail(GTB_ERR_ARG, "null ctx");
  if (kernel_ms)
    *kernel_ms = c->t_sw_kernel;
  if (h2d_ms)
    *h2d_ms = c->t_sw_h2d;
  if (d2h_ms)
    *d2h_ms = c->t_sw_d2h;
  return 0;
}

} // extern "C"
