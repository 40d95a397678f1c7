// ptb200 — multi-GPU render entry point of the C ABI (SURVEY.md §8e / §8b "ptb_render_multi").
// The path shards by samples: context r of n renders its share of the sample range of every pixel (scene and BVH are
// replicated: each context committed its own copy; the build is deterministic), then the per-GPU accumulators — sums, not
// means — are combined by ONE ncclReduce(sum) to context 0 over NVLink. That reduce is the only collective of the path.
// A request with fewer samples than GPUs is split along the image instead (bands of pixel rows, every sample): pixels keep
// their full-image coordinates, a band only ever adds to its own rows, so the same reduce combines the tiles.
// One host thread per GPU drives its context; NCCL is bound at run time (dlopen) so libptb200.so carries no link-time
// dependency on it and single-GPU users never load it.
#include <dlfcn.h>

#include <mutex>
#include <thread>
#include <vector>

#include "ptb_internal.h"

namespace ptb {
namespace {

// the slice of nccl.h this file needs (NCCL 2.x ABI)
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
constexpr int kNcclFloat32 = 7, kNcclSum = 0;
struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;  // optional (NCCL >= 2.4)
  std::vector<int> devices;        // communicators are cached for the last device list
  std::vector<ncclComm_t> comms;
  std::mutex mu;
};
Nccl g_nccl;

bool load_nccl(std::string& why) {
  if (g_nccl.lib) return true;
  const char* names[] = {getenv("PTB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n) continue;
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) { why = "libnccl.so.2 not found (set PTB_NCCL_LIB)"; return false; }
#define PTB_SYM(field, name)                                                                 \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.lib, name));           \
  if (!g_nccl.field) { why = std::string("libnccl lacks ") + name; g_nccl.lib = nullptr; return false; }
  PTB_SYM(CommInitAll, "ncclCommInitAll")
  PTB_SYM(CommDestroy, "ncclCommDestroy")
  PTB_SYM(Reduce, "ncclReduce")
  PTB_SYM(GroupStart, "ncclGroupStart")
  PTB_SYM(GroupEnd, "ncclGroupEnd")
  PTB_SYM(GetErrorString, "ncclGetErrorString")
#undef PTB_SYM
  g_nccl.CommGetAsyncError = reinterpret_cast<decltype(g_nccl.CommGetAsyncError)>(dlsym(g_nccl.lib, "ncclCommGetAsyncError"));
  return true;
}

}  // namespace
}  // namespace ptb

using namespace ptb;

extern "C" {

void ptb_shard_samples(uint32_t samples_per_pixel, uint32_t sample_offset, int32_t rank, int32_t world, uint32_t* first,
                       uint32_t* count) {
  // rank r renders [off + r*spp/G, off + (r+1)*spp/G): the ranges tile the request, sizes differ by <= 1
  const uint64_t lo = (uint64_t)rank * samples_per_pixel / (uint64_t)world;
  const uint64_t hi = ((uint64_t)rank + 1u) * samples_per_pixel / (uint64_t)world;
  if (first) *first = sample_offset + (uint32_t)lo;
  if (count) *count = (uint32_t)(hi - lo);
}

void ptb_shard_rows(uint32_t rows, uint32_t row_begin, int32_t rank, int32_t world, uint32_t* first, uint32_t* count) {
  const uint64_t lo = (uint64_t)rank * rows / (uint64_t)world;
  const uint64_t hi = ((uint64_t)rank + 1u) * rows / (uint64_t)world;
  if (first) *first = row_begin + (uint32_t)lo;
  if (count) *count = (uint32_t)(hi - lo);
}

int32_t ptb_render_multi(ptb_ctx* const* ctxs, int32_t n, const ptb_render_opts* opts) {
  if (!ctxs || n < 1 || !opts) return PTB_ERR_INVALID;
  for (int32_t i = 0; i < n; ++i)
    if (!ctxs[i]) return PTB_ERR_INVALID;
  Ctx* root = &ctxs[0]->c;
  for (int32_t i = 0; i < n; ++i)
    for (int32_t j = 0; j < i; ++j)
      if (ctxs[i]->c.device == ctxs[j]->c.device) return set_error(root, PTB_ERR_INVALID, "contexts %d and %d share GPU %d", j, i, ctxs[i]->c.device);

  // ---- render: one host thread per GPU, each on its own sample range
  std::vector<int32_t> rcs((size_t)n, PTB_OK);
  std::vector<std::thread> threads;
  for (int32_t r = 0; r < n; ++r)
    threads.emplace_back([&, r]() {
      ptb_render_opts o = *opts;
      bool work = true;
      if (opts->samples_per_pixel >= (uint32_t)n) {  // sample axis
        ptb_shard_samples(opts->samples_per_pixel, opts->sample_offset, r, n, &o.sample_offset, &o.samples_per_pixel);
      } else {  // image-tile axis: every sample of a band of rows
        const uint32_t rows = opts->row_count ? opts->row_count : (opts->row_begin < opts->height ? opts->height - opts->row_begin : 0u);
        ptb_shard_rows(rows, opts->row_begin, r, n, &o.row_begin, &o.row_count);
        work = o.row_count != 0u;  // more GPUs than rows
        if (!work) o.row_begin = 0u;
      }
      int32_t rc;
      {  // zero passes: only sizes this context's accumulator
        ptb_render_opts none = *opts;
        none.samples_per_pixel = 0;
        rc = ptb_render(ctxs[r], &none, nullptr, nullptr);
      }
      if (rc == PTB_OK) rc = ptb_accum_clear(ctxs[r]);
      if (rc == PTB_OK && work && o.samples_per_pixel) rc = ptb_render(ctxs[r], &o, nullptr, nullptr);
      if (rc == PTB_OK) rc = ptb_synchronize(ctxs[r]);
      rcs[(size_t)r] = rc;
    });
  for (auto& t : threads) t.join();
  for (int32_t r = 0; r < n; ++r)
    if (rcs[(size_t)r] != PTB_OK) {
      if (r != 0) set_error(root, rcs[(size_t)r], "GPU %d: %s", ctxs[r]->c.device, ctxs[r]->c.last_error.c_str());
      return rcs[(size_t)r];
    }
  if (n == 1) return PTB_OK;

  // ---- combine: ncclReduce(sum) of W*H*3 f32 to context 0
  std::lock_guard<std::mutex> lock(g_nccl.mu);
  std::string why;
  if (!load_nccl(why)) return set_error(root, PTB_ERR_UNSUPPORTED, "ptb_render_multi needs NCCL: %s", why.c_str());
  std::vector<int> devices;
  for (int32_t r = 0; r < n; ++r) devices.push_back(ctxs[r]->c.device);
  if (devices != g_nccl.devices) {
    for (ncclComm_t cm : g_nccl.comms) g_nccl.CommDestroy(cm);
    g_nccl.comms.assign((size_t)n, nullptr);
    g_nccl.devices.clear();
    const ncclResult_t e = g_nccl.CommInitAll(g_nccl.comms.data(), n, devices.data());
    if (e != 0) { g_nccl.comms.clear(); return set_error(root, PTB_ERR_CUDA, "ncclCommInitAll: %s", g_nccl.GetErrorString(e)); }
    g_nccl.devices = devices;
  }
  const size_t count = (size_t)opts->width * opts->height * 3;
  ncclResult_t e = g_nccl.GroupStart();
  for (int32_t r = 0; r < n && e == 0; ++r) {
    Ctx* c = &ctxs[r]->c;
    cudaSetDevice(c->device);
    e = g_nccl.Reduce(c->d_accum.p, c->d_accum.p, count, kNcclFloat32, kNcclSum, 0, g_nccl.comms[(size_t)r], c->stream);
  }
  const ncclResult_t e2 = g_nccl.GroupEnd();
  if (e != 0 || e2 != 0) return set_error(root, PTB_ERR_CUDA, "ncclReduce: %s", g_nccl.GetErrorString(e != 0 ? e : e2));
  for (int32_t r = 0; r < n; ++r) {
    const int32_t rc = ptb_synchronize(ctxs[r]);
    if (rc != PTB_OK) return rc;
  }
  // a failed collective (peer fault, NVLink error) surfaces asynchronously on the communicator, not on the stream
  if (g_nccl.CommGetAsyncError)
    for (int32_t r = 0; r < n; ++r) {
      ncclResult_t async = 0;
      const ncclResult_t q = g_nccl.CommGetAsyncError(g_nccl.comms[(size_t)r], &async);
      if (q != 0 || async != 0) {
        const ncclResult_t bad = q != 0 ? q : async;
        for (ncclComm_t cm : g_nccl.comms) g_nccl.CommDestroy(cm);  // the communicators are unusable after an async error
        g_nccl.comms.clear();
        g_nccl.devices.clear();
        return set_error(root, PTB_ERR_CUDA, "ncclReduce failed asynchronously on GPU %d: %s", ctxs[r]->c.device, g_nccl.GetErrorString(bad));
      }
    }
  root->accum_samples = opts->samples_per_pixel;
  root->stats.kernel_launches += 1;
  return PTB_OK;
}

}  // extern "C"
