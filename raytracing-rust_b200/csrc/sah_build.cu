// ptb200 — device SAH builder: a top-down surface-area-heuristic binary hierarchy in the LBVH's node format, the
// device-side replacement for the QUALITY of the reference's builder (implementations/src/acceleration/mod.rs:97-160
// `build_bvh` + split.rs:78-187 `Split::Sah`: buckets along an axis of a node's box, cost n_l * area_l + n_r * area_r,
// split.rs:161-163,176). CPU definition, operation by operation: oracle/sah_ref.hpp — the device tree is compared with it
// bit for bit (tests/test_gpu_sah.py). Everything here is deterministic: bins are min / max / integer counts (order
// independent atomics), partitions are stable (prefix sums), nodes are named after the gap they split at.
//
//   S1  k_sah_root        root task: all positions of the Morton order, box = union of the primitive boxes (k_prim_bounds)
//   per level of LARGE tasks (more than 32 positions), tasks in position order:
//   S2  k_sah_clear / k_sah_bin    8 bins on each axis of the task's box: union of boxes + count (block- / warp-aggregated atomics)
//   S3  k_sah_eval        one thread per task: first minimum of the cost over axes x planes; names and links the node; the
//                         children become leaves, SMALL tasks (appended to a list) or next-level large tasks
//   S4  k_sah_task_scan / k_sah_emit   next level's task list, in position order
//   S5  k_sah_flag_scan + scan_block_sums + k_sah_scatter   stable partition of every task's positions, all tasks at once
//   S6  k_sah_small       one WARP per small task: exact sweep over every (member, axis) candidate, the whole subtree in
//                         registers / shuffles, a stable partition per split through shared memory
//   S7  k_sah_finish      sphere bits of the leaf references
// The host reads one word per level (the number of next-level tasks). 1 M triangles: 21 levels, <= 16 193 tasks per level,
// 44 595 small tasks.
#include <cstdlib>
#include <cstring>

#include "ptb_internal.h"

namespace ptb {

namespace {

constexpr int kSahBins = 8;
constexpr uint32_t kSahSmall = 32;
// The traversal kernels' stacks hold 64 entries (ptb_traverse.cuh kStackDepth, ptb_packet.cuh): a split whose children could
// not both be finished by halving within kSahMaxDepth levels is replaced by the halving split (oracle/sah_ref.hpp).
constexpr uint32_t kSahMaxDepth = 60;
constexpr int kBinWords = 7;                          // min xyz, max xyz (order-preserving uint), count
constexpr int kTaskBinWords = 3 * kSahBins * kBinWords;  // 168

struct SahTask {  // 48 bytes
  uint32_t lo, hi;  // positions [lo, hi)
  uint32_t parent, side;
  float mn[3], mx[3];
  uint32_t depth;  // of the task's node (root: 0)
  uint32_t _pad;
};
struct SahSplit {  // per task of the current level, written by k_sah_eval / k_sah_emit
  int32_t axis;    // -1: halved in its current order
  int32_t bin;
  uint32_t cl;     // positions going left
  float scale, mn;
  uint32_t lo;
  uint32_t child[2];  // next-level task index of the left / right child, kNone when it is a leaf or a small task
};
// counters: [0] small tasks appended, [1] the root's gap, [2] tasks of the next level

__device__ __forceinline__ uint32_t sah_flip(float f) {  // order-preserving float -> uint
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float sah_unflip(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u); }
constexpr uint32_t kFlipPosInf = 0xFF800000u, kFlipNegInf = 0x007FFFFFu;

__device__ __forceinline__ float half_area(const float* mn, const float* mx) {
  const float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
  return (dx * dy + dy * dz) + dz * dx;
}
__device__ __forceinline__ bool axis_scale(float mn, float mx, float& scale) {
  const float ext = mx - mn;
  scale = 0.0f;
  if (!(ext > 0.0f)) return false;
  scale = (float)kSahBins / ext;
  return scale < 1.0e30f;
}
__device__ __forceinline__ int bin_of(float c, float mn, float scale) { return (int)fminf((c - mn) * scale, (float)(kSahBins - 1)); }
__host__ __device__ __forceinline__ uint32_t sah_ceil_log2(uint32_t n) {
  uint32_t k = 0u;
  while (k < 32u && (1ull << k) < (unsigned long long)n) ++k;
  return k;
}
__device__ __forceinline__ bool sah_fits(uint32_t depth, uint32_t n, uint32_t max_depth) {
  return depth + (n <= 1u ? 0u : 32u - (uint32_t)__clz(n - 1u)) <= max_depth;
}
__device__ __forceinline__ uint32_t node_id(uint32_t gap, uint32_t root_gap) {
  return gap == root_gap ? 0u : (gap == 0u ? root_gap : gap);
}
__device__ __forceinline__ void link_child(BvhNode* nodes, uint32_t parent, uint32_t side, uint32_t ref) {
  if (parent == kNone) return;
  uint32_t* w = reinterpret_cast<uint32_t*>(&nodes[parent].n3);
  w[side] = ref;
}

// ------------------------------------------------------------------------------------------ S1
__global__ void k_sah_root(const uint32_t* __restrict__ box6, uint32_t n, SahTask* __restrict__ task, uint32_t* __restrict__ counters) {
  SahTask t;
  t.lo = 0u; t.hi = n; t.parent = kNone; t.side = 0u;
  for (int k = 0; k < 3; ++k) { t.mn[k] = sah_unflip(box6[k]); t.mx[k] = sah_unflip(box6[3 + k]); }
  t.depth = 0u;
  t._pad = 0u;
  *task = t;
  counters[0] = 0u;
  counters[1] = kNone;
  counters[2] = 0u;
}
__global__ void __launch_bounds__(256) k_sah_fill(uint32_t* __restrict__ p, uint32_t n, uint32_t v) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------ S2
__global__ void __launch_bounds__(256) k_sah_clear(uint32_t* __restrict__ bins, uint32_t n_tasks) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tasks * (uint32_t)kTaskBinWords) return;
  const uint32_t w = i % (uint32_t)kBinWords;
  bins[i] = w < 3u ? kFlipPosInf : (w < 6u ? kFlipNegInf : 0u);
}
// One thread per position. A block whose positions all belong to ONE task (the upper levels: tasks of thousands of
// positions) folds its bins in shared memory and sends 168 atomics at most; other warps aggregate per (axis, bin) group when
// the warp's positions share a task, and fall back to one set of atomics per lane otherwise.
template <class Atom>
__device__ __forceinline__ void bin_warp(bool active, const int* bi, const uint32_t* v, uint32_t lane, Atom atom) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    for (int b = 0; b < kSahBins; ++b) {
      const bool in = active && bi[a] == b;
      const uint32_t m = __ballot_sync(0xffffffffu, in);
      if (!m) continue;
      uint32_t r[6];
#pragma unroll
      for (int k = 0; k < 3; ++k) r[k] = __reduce_min_sync(0xffffffffu, in ? v[k] : kFlipPosInf);
#pragma unroll
      for (int k = 3; k < 6; ++k) r[k] = __reduce_max_sync(0xffffffffu, in ? v[k] : kFlipNegInf);
      if (lane == (uint32_t)(__ffs(m) - 1)) atom((a * kSahBins + b) * kBinWords, r, (uint32_t)__popc(m));
    }
  }
}
__global__ void __launch_bounds__(256)
k_sah_bin(uint32_t n, const uint32_t* __restrict__ order, const uint32_t* __restrict__ slot_task, const SahTask* __restrict__ tasks,
          const float4* __restrict__ bmin, const float4* __restrict__ bmax, uint32_t* __restrict__ bins) {
  __shared__ uint32_t sb[kTaskBinWords];
  __shared__ uint32_t s_t0;
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t t = slot < n ? slot_task[slot] : kNone;
  const bool active = t != kNone;
  if (threadIdx.x == 0u) s_t0 = t;
  if (threadIdx.x < (uint32_t)kTaskBinWords) {
    const uint32_t w = threadIdx.x % (uint32_t)kBinWords;
    sb[threadIdx.x] = w < 3u ? kFlipPosInf : (w < 6u ? kFlipNegInf : 0u);
  }
  uint32_t v[6] = {kFlipPosInf, kFlipPosInf, kFlipPosInf, kFlipNegInf, kFlipNegInf, kFlipNegInf};
  int bi[3] = {0, 0, 0};
  if (active) {
    const uint32_t p = order[slot];
    const float4 a = bmin[p], b = bmax[p];
    const SahTask& tk = tasks[t];
    const float lo3[3] = {a.x, a.y, a.z}, hi3[3] = {b.x, b.y, b.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      v[k] = sah_flip(lo3[k]);
      v[3 + k] = sah_flip(hi3[k]);
      float scale;
      const float mn = tk.mn[k];
      if (axis_scale(mn, tk.mx[k], scale)) bi[k] = bin_of(0.5f * (lo3[k] + hi3[k]), mn, scale);
    }
  }
  __syncthreads();
  const uint32_t bt = s_t0;
  const bool block_uniform = __syncthreads_and(!active || t == bt) != 0;
  if (block_uniform) {
    if (bt == kNone) return;  // nothing active in this block
    bin_warp(active, bi, v, lane, [&](int w0, const uint32_t* r, uint32_t c) {
      for (int k = 0; k < 3; ++k) atomicMin(sb + w0 + k, r[k]);
      for (int k = 3; k < 6; ++k) atomicMax(sb + w0 + k, r[k]);
      atomicAdd(sb + w0 + 6, c);
    });
    __syncthreads();
    if (threadIdx.x < (uint32_t)kTaskBinWords) {
      const uint32_t w = threadIdx.x % (uint32_t)kBinWords;
      if (sb[threadIdx.x - w + 6u] != 0u) {
        uint32_t* g = bins + (size_t)bt * kTaskBinWords + threadIdx.x;
        if (w < 3u) atomicMin(g, sb[threadIdx.x]);
        else if (w < 6u) atomicMax(g, sb[threadIdx.x]);
        else atomicAdd(g, sb[threadIdx.x]);
      }
    }
    return;
  }
  const uint32_t any = __ballot_sync(0xffffffffu, active);
  if (!any) return;
  const uint32_t t0 = __shfl_sync(0xffffffffu, t, __ffs(any) - 1);
  if (__all_sync(0xffffffffu, !active || t == t0)) {
    uint32_t* tb = bins + (size_t)t0 * kTaskBinWords;
    bin_warp(active, bi, v, lane, [&](int w0, const uint32_t* r, uint32_t c) {
      for (int k = 0; k < 3; ++k) atomicMin(tb + w0 + k, r[k]);
      for (int k = 3; k < 6; ++k) atomicMax(tb + w0 + k, r[k]);
      atomicAdd(tb + w0 + 6, c);
    });
  } else if (active) {
    uint32_t* tb = bins + (size_t)t * kTaskBinWords;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      uint32_t* w = tb + (a * kSahBins + bi[a]) * kBinWords;
#pragma unroll
      for (int k = 0; k < 3; ++k) atomicMin(w + k, v[k]);
#pragma unroll
      for (int k = 3; k < 6; ++k) atomicMax(w + k, v[k]);
      atomicAdd(w + 6, 1u);
    }
  }
}

// ------------------------------------------------------------------------------------------ S3
__global__ void __launch_bounds__(128)
k_sah_eval(uint32_t n_tasks, const SahTask* __restrict__ tasks, const uint32_t* __restrict__ bins, SahSplit* __restrict__ splits,
           SahTask* __restrict__ child_tasks, uint32_t* __restrict__ n_large, BvhNode* nodes, uint32_t* __restrict__ leaf_parent,
           SahTask* __restrict__ small, uint32_t* counters, uint32_t max_depth) {
  const uint32_t ti = blockIdx.x * blockDim.x + threadIdx.x;
  if (ti >= n_tasks) return;
  const SahTask t = tasks[ti];
  const uint32_t len = t.hi - t.lo;
  const uint32_t* tb = bins + (size_t)ti * kTaskBinWords;
  float best = __int_as_float(0x7f800000);
  int best_axis = -1, best_bin = 0;
  uint32_t best_cl = 0u;
  float best_scale = 0.0f;
  for (int a = 0; a < 3; ++a) {
    float scale;
    if (!axis_scale(t.mn[a], t.mx[a], scale)) continue;
    const uint32_t* ab = tb + a * kSahBins * kBinWords;
    float ra[kSahBins];
    uint32_t rc[kSahBins];
    float mn[3] = {__int_as_float(0x7f800000), __int_as_float(0x7f800000), __int_as_float(0x7f800000)};
    float mx[3] = {__int_as_float(0xff800000), __int_as_float(0xff800000), __int_as_float(0xff800000)};
    uint32_t c = 0u;
#pragma unroll
    for (int i = kSahBins - 1; i > 0; --i) {
      const uint32_t* w = ab + i * kBinWords;
#pragma unroll
      for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], sah_unflip(w[k])); mx[k] = fmaxf(mx[k], sah_unflip(w[3 + k])); }
      c += w[6];
      ra[i] = c ? half_area(mn, mx) : 0.0f;
      rc[i] = c;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { mn[k] = __int_as_float(0x7f800000); mx[k] = __int_as_float(0xff800000); }
    c = 0u;
#pragma unroll
    for (int i = 0; i < kSahBins - 1; ++i) {
      const uint32_t* w = ab + i * kBinWords;
#pragma unroll
      for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], sah_unflip(w[k])); mx[k] = fmaxf(mx[k], sah_unflip(w[3 + k])); }
      c += w[6];
      if (c == 0u || rc[i + 1] == 0u) continue;
      const float cost = half_area(mn, mx) * (float)c + ra[i + 1] * (float)rc[i + 1];
      if (cost < best) { best = cost; best_axis = a; best_bin = i; best_cl = c; best_scale = scale; }
    }
  }
  if (best_axis >= 0 && !(sah_fits(t.depth + 1u, best_cl, max_depth) && sah_fits(t.depth + 1u, len - best_cl, max_depth))) best_axis = -1;
  SahTask L, R;
  uint32_t cl;
  if (best_axis < 0) {
    cl = len >> 1;
    for (int k = 0; k < 3; ++k) { L.mn[k] = R.mn[k] = t.mn[k]; L.mx[k] = R.mx[k] = t.mx[k]; }
  } else {
    cl = best_cl;
    for (int k = 0; k < 3; ++k) {
      L.mn[k] = R.mn[k] = __int_as_float(0x7f800000);
      L.mx[k] = R.mx[k] = __int_as_float(0xff800000);
    }
    const uint32_t* ab = tb + best_axis * kSahBins * kBinWords;
    for (int i = 0; i < kSahBins; ++i) {
      const uint32_t* w = ab + i * kBinWords;
      SahTask& d = i <= best_bin ? L : R;
      for (int k = 0; k < 3; ++k) { d.mn[k] = fminf(d.mn[k], sah_unflip(w[k])); d.mx[k] = fmaxf(d.mx[k], sah_unflip(w[3 + k])); }
    }
  }
  const uint32_t gap = t.lo + cl - 1u;
  uint32_t id;
  if (t.parent == kNone) {
    counters[1] = gap;
    id = 0u;
  } else {
    id = node_id(gap, counters[1]);
  }
  nodes[id].n3.z = t.parent;
  nodes[id].n3.w = 0u;
  link_child(nodes, t.parent, t.side, id);
  L.lo = t.lo; L.hi = t.lo + cl; L.parent = id; L.side = 0u;
  R.lo = t.lo + cl; R.hi = t.hi; R.parent = id; R.side = 1u;
  L.depth = R.depth = t.depth + 1u;
  L._pad = R._pad = 0u;
  SahSplit sp;
  sp.axis = best_axis;
  sp.bin = best_bin;
  sp.cl = cl;
  sp.scale = best_scale;
  sp.mn = best_axis < 0 ? 0.0f : t.mn[best_axis];
  sp.lo = t.lo;
  sp.child[0] = sp.child[1] = kNone;
  uint32_t large = 0u;
  for (int s = 0; s < 2; ++s) {
    const SahTask& c = s ? R : L;
    const uint32_t cn = c.hi - c.lo;
    if (cn == 1u) {
      link_child(nodes, id, (uint32_t)s, PTB_LEAF_BIT | c.lo);
      leaf_parent[c.lo] = id;
    } else if (cn <= kSahSmall) {
      small[atomicAdd(counters + 0, 1u)] = c;
    } else {
      child_tasks[2u * ti + large] = c;
      sp.child[s] = large;  // index among this task's large children; k_sah_emit adds the base
      ++large;
    }
  }
  n_large[ti] = large;
  splits[ti] = sp;
}

// ------------------------------------------------------------------------------------------ S4
// exclusive scan of n_large[0 .. n_tasks) by ONE block (a level has at most n / 33 tasks), total -> counters[2]
__global__ void __launch_bounds__(1024) k_sah_task_scan(uint32_t* __restrict__ n_large, uint32_t n_tasks, uint32_t* counters) {
  __shared__ uint32_t warp_sum[33];
  const uint32_t per = (n_tasks + 1023u) / 1024u;
  const uint32_t b = threadIdx.x * per, e = min(b + per, n_tasks);
  uint32_t mine = 0u;
  for (uint32_t i = b; i < e; ++i) mine += n_large[i];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += v;
  }
  if (lane == 31u) warp_sum[warp] = inc;
  __syncthreads();
  if (warp == 0u) {
    const uint32_t v = warp_sum[lane];
    uint32_t winc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= (uint32_t)o) winc += u;
    }
    warp_sum[lane] = winc - v;
    if (lane == 31u) warp_sum[32] = winc;
  }
  __syncthreads();
  uint32_t run = warp_sum[warp] + inc - mine;
  for (uint32_t i = b; i < e; ++i) {
    const uint32_t v = n_large[i];
    n_large[i] = run;
    run += v;
  }
  if (threadIdx.x == 0u) counters[2] = warp_sum[32];
}
__global__ void __launch_bounds__(128)
k_sah_emit(uint32_t n_tasks, const uint32_t* __restrict__ base, const SahTask* __restrict__ child_tasks, SahSplit* __restrict__ splits,
           SahTask* __restrict__ next) {
  const uint32_t ti = blockIdx.x * blockDim.x + threadIdx.x;
  if (ti >= n_tasks) return;
  const uint32_t b = base[ti];
  for (int s = 0; s < 2; ++s) {
    const uint32_t k = splits[ti].child[s];
    if (k == kNone) continue;
    next[b + k] = child_tasks[2u * ti + k];
    splits[ti].child[s] = b + k;
  }
}

// ------------------------------------------------------------------------------------------ S5
__device__ __forceinline__ bool goes_left(const SahSplit& sp, uint32_t slot, uint32_t p, const float4* __restrict__ bmin,
                                          const float4* __restrict__ bmax) {
  if (sp.axis < 0) return slot - sp.lo < sp.cl;
  const float4 a = bmin[p], b = bmax[p];
  const float lo = sp.axis == 0 ? a.x : (sp.axis == 1 ? a.y : a.z), hi = sp.axis == 0 ? b.x : (sp.axis == 1 ? b.y : b.z);
  return bin_of(0.5f * (lo + hi), sp.mn, sp.scale) <= sp.bin;
}
// left flags and their exclusive prefix sums in one pass: 4096 positions per block, prefix[] = the sum inside the block,
// block_sum[] = the block's total (scanned by scan_block_sums; k_sah_scatter adds it back)
constexpr int kFlagBlock = 1024, kFlagItems = 4, kFlagTile = kFlagBlock * kFlagItems;
__global__ void __launch_bounds__(kFlagBlock)
k_sah_flag_scan(uint32_t n, const uint32_t* __restrict__ order, const uint32_t* __restrict__ slot_task, const SahSplit* __restrict__ splits,
                const float4* __restrict__ bmin, const float4* __restrict__ bmax, uint32_t* __restrict__ prefix, uint32_t* __restrict__ block_sum) {
  __shared__ uint32_t warp_sum[33];
  const uint32_t base = blockIdx.x * kFlagTile + threadIdx.x * kFlagItems;
  uint32_t f[kFlagItems], mine = 0u;
#pragma unroll
  for (int k = 0; k < kFlagItems; ++k) {
    const uint32_t slot = base + k;
    f[k] = 0u;
    if (slot < n) {
      const uint32_t t = slot_task[slot];
      if (t != kNone && goes_left(splits[t], slot, order[slot], bmin, bmax)) f[k] = 1u;
    }
    mine += f[k];
  }
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += u;
  }
  if (lane == 31u) warp_sum[warp] = inc;
  __syncthreads();
  if (warp == 0u) {
    const uint32_t v = warp_sum[lane];
    uint32_t winc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= (uint32_t)o) winc += u;
    }
    warp_sum[lane] = winc - v;
    if (lane == 31u) warp_sum[32] = winc;
  }
  __syncthreads();
  uint32_t run = warp_sum[warp] + inc - mine;
#pragma unroll
  for (int k = 0; k < kFlagItems; ++k) {
    if (base + k < n) prefix[base + k] = run;
    run += f[k];
  }
  if (threadIdx.x == 0u) block_sum[blockIdx.x] = warp_sum[32];
}
// global prefix of a position = positions going left before it (whole array): the rank inside a task is the difference to
// the prefix of the task's first position
__global__ void __launch_bounds__(256)
k_sah_scatter(uint32_t n, const uint32_t* __restrict__ order, const uint32_t* __restrict__ slot_task, const SahSplit* __restrict__ splits,
              const float4* __restrict__ bmin, const float4* __restrict__ bmax, const uint32_t* __restrict__ prefix,
              const uint32_t* __restrict__ block_sum, uint32_t* __restrict__ order_out, uint32_t* __restrict__ slot_task_out) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n) return;
  const uint32_t t = slot_task[slot];
  const uint32_t p = order[slot];
  if (t == kNone) {
    order_out[slot] = p;
    slot_task_out[slot] = kNone;
    return;
  }
  const SahSplit sp = splits[t];
  const bool left = goes_left(sp, slot, p, bmin, bmax);
  const uint32_t rank_left = (prefix[slot] + block_sum[slot / kFlagTile]) - (prefix[sp.lo] + block_sum[sp.lo / kFlagTile]);
  const uint32_t dest = left ? sp.lo + rank_left : sp.lo + sp.cl + ((slot - sp.lo) - rank_left);
  order_out[dest] = p;
  slot_task_out[dest] = sp.child[left ? 0 : 1];
}

// ------------------------------------------------------------------------------------------ S6
constexpr int kSmallWarps = 4;
constexpr int kSmallWords = 15;  // primitive, box (6), centroid (3), segment lo / hi, parent, side, depth
__global__ void __launch_bounds__(32 * kSmallWarps)
k_sah_small(uint32_t n_small, const SahTask* __restrict__ small, uint32_t* __restrict__ order, const float4* __restrict__ bmin,
            const float4* __restrict__ bmax, BvhNode* nodes, uint32_t* __restrict__ leaf_parent, uint32_t* counters, uint32_t max_depth) {
  __shared__ uint32_t sm[kSmallWarps][kSmallWords][32];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t ti = blockIdx.x * kSmallWarps + warp;
  if (ti >= n_small) return;
  const SahTask t = small[ti];
  const uint32_t m = t.hi - t.lo;
  const bool active = lane < m;
  uint32_t p = 0u;
  float bx[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, c[3] = {0.f, 0.f, 0.f};
  if (active) {
    p = order[t.lo + lane];
    const float4 a = bmin[p], b = bmax[p];
    bx[0] = a.x; bx[1] = a.y; bx[2] = a.z; bx[3] = b.x; bx[4] = b.y; bx[5] = b.z;
#pragma unroll
    for (int k = 0; k < 3; ++k) c[k] = 0.5f * (bx[k] + bx[3 + k]);
  }
  uint32_t seg_lo = active ? 0u : lane, seg_hi = active ? m : lane, par = t.parent, side = t.side, depth = t.depth;
  uint32_t root_gap = t.parent == kNone ? kNone : counters[1];
  const float inf = __int_as_float(0x7f800000);
  for (;;) {
    const uint32_t len = seg_hi - seg_lo;
    const bool need = active && len >= 2u;
    if (!__any_sync(0xffffffffu, need)) break;
    // candidates (this lane, axis a): members j of the segment with (c_a[j], j) <= (c_a[lane], lane) go left
    float Lmn[3][3], Lmx[3][3], Rmn[3][3], Rmx[3][3];
    uint32_t cl[3] = {0u, 0u, 0u};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int k = 0; k < 3; ++k) { Lmn[a][k] = Rmn[a][k] = inf; Lmx[a][k] = Rmx[a][k] = -inf; }
    // every lane walks the members of ITS OWN segment (source lane seg_lo + i): the trip count is the longest segment of the
    // warp, which halves from level to level, not the size of the task
    const uint32_t max_len = __reduce_max_sync(0xffffffffu, need ? len : 0u);
    for (uint32_t i = 0; i < max_len; ++i) {
      const uint32_t j = min(seg_lo + i, 31u);
      float jb[6], jc[3];
#pragma unroll
      for (int k = 0; k < 6; ++k) jb[k] = __shfl_sync(0xffffffffu, bx[k], j);
#pragma unroll
      for (int k = 0; k < 3; ++k) jc[k] = __shfl_sync(0xffffffffu, c[k], j);
      const bool same = need && i < len;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const bool left = same && (jc[a] < c[a] || (jc[a] == c[a] && j <= lane));
        const bool right = same && !left;
        if (left) {
          ++cl[a];
#pragma unroll
          for (int k = 0; k < 3; ++k) { Lmn[a][k] = fminf(Lmn[a][k], jb[k]); Lmx[a][k] = fmaxf(Lmx[a][k], jb[3 + k]); }
        }
        if (right) {
#pragma unroll
          for (int k = 0; k < 3; ++k) { Rmn[a][k] = fminf(Rmn[a][k], jb[k]); Rmx[a][k] = fmaxf(Rmx[a][k], jb[3 + k]); }
        }
      }
    }
    float best = inf;
    int best_a = -1;
    uint32_t best_cl = 0u;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (!need || cl[a] == len) continue;
      const float cost = half_area(Lmn[a], Lmx[a]) * (float)cl[a] + half_area(Rmn[a], Rmx[a]) * (float)(len - cl[a]);
      if (cost < best) { best = cost; best_a = a; best_cl = cl[a]; }
    }
    const float my_pivot_c = best_a == 0 ? c[0] : (best_a == 1 ? c[1] : c[2]);
    // the segment's winner: smallest (cost, lane)
    float seg_best = inf;
    uint32_t pl = kNone;
    for (uint32_t i = 0; i < max_len; ++i) {
      const uint32_t j = min(seg_lo + i, 31u);
      const float k = __shfl_sync(0xffffffffu, best, j);
      if (need && i < len && k < seg_best) { seg_best = k; pl = j; }
    }
    const uint32_t src = pl == kNone ? lane : pl;
    const int pa = __shfl_sync(0xffffffffu, best_a, src);
    const float pc = __shfl_sync(0xffffffffu, my_pivot_c, src);
    uint32_t scl = __shfl_sync(0xffffffffu, best_cl, src);
    if (pl != kNone && !(sah_fits(depth + 1u, scl, max_depth) && sah_fits(depth + 1u, len - scl, max_depth))) pl = kNone;  // depth bound: halve
    bool mine;
    if (pl == kNone) {
      scl = len >> 1;
      mine = lane - seg_lo < scl;
    } else {
      const float mc = pa == 0 ? c[0] : (pa == 1 ? c[1] : c[2]);
      mine = mc < pc || (mc == pc && lane <= pl);
    }
    const uint32_t leftmask = __ballot_sync(0xffffffffu, need && mine);
    uint32_t newpos = lane, nlo = seg_lo, nhi = seg_hi, npar = par, nside = side, ndepth = depth;
    if (need) {
      const uint32_t below = leftmask & ((1u << lane) - 1u) & ~((1u << seg_lo) - 1u);
      const uint32_t rank_left = (uint32_t)__popc(below);
      newpos = mine ? seg_lo + rank_left : seg_lo + scl + ((lane - seg_lo) - rank_left);
      const uint32_t gap = t.lo + seg_lo + scl - 1u;
      uint32_t id;
      if (root_gap == kNone) {  // the root itself is a small task: its first split names node 0
        root_gap = gap;
        id = 0u;
      } else {
        id = node_id(gap, root_gap);
      }
      if (lane == seg_lo) {
        nodes[id].n3.z = par;
        nodes[id].n3.w = 0u;
        link_child(nodes, par, side, id);
      }
      npar = id;
      ndepth = depth + 1u;
      if (mine) { nhi = seg_lo + scl; nside = 0u; } else { nlo = seg_lo + scl; nside = 1u; }
    }
    root_gap = __shfl_sync(0xffffffffu, root_gap, 0);  // (root task: one segment at its first split, lane 0 is a member)
    // stable partition: every item moves to its new lane through shared memory
    uint32_t* s = &sm[warp][0][0];
    s[0 * 32 + newpos] = p;
#pragma unroll
    for (int k = 0; k < 6; ++k) s[(1 + k) * 32 + newpos] = __float_as_uint(bx[k]);
#pragma unroll
    for (int k = 0; k < 3; ++k) s[(7 + k) * 32 + newpos] = __float_as_uint(c[k]);
    s[10 * 32 + newpos] = nlo;
    s[11 * 32 + newpos] = nhi;
    s[12 * 32 + newpos] = npar;
    s[13 * 32 + newpos] = nside;
    s[14 * 32 + newpos] = ndepth;
    const uint32_t was_need = __ballot_sync(0xffffffffu, need);
    __syncwarp();
    p = s[0 * 32 + lane];
#pragma unroll
    for (int k = 0; k < 6; ++k) bx[k] = __uint_as_float(s[(1 + k) * 32 + lane]);
#pragma unroll
    for (int k = 0; k < 3; ++k) c[k] = __uint_as_float(s[(7 + k) * 32 + lane]);
    seg_lo = s[10 * 32 + lane];
    seg_hi = s[11 * 32 + lane];
    par = s[12 * 32 + lane];
    side = s[13 * 32 + lane];
    depth = s[14 * 32 + lane];
    __syncwarp();
    // an item whose segment has just become a single position is a leaf (a segment is split as a whole, so the lanes of
    // a segment that needed a split are exactly the lanes that hold its items afterwards)
    if (((was_need >> lane) & 1u) && seg_hi - seg_lo == 1u) {
      link_child(nodes, par, side, PTB_LEAF_BIT | (t.lo + lane));
      leaf_parent[t.lo + lane] = par;
    }
  }
  if (active) order[t.lo + lane] = p;
  if (t.parent == kNone && lane == 0u) counters[1] = root_gap;
}

// ------------------------------------------------------------------------------------------ S7
__global__ void __launch_bounds__(256)
k_sah_finish(uint32_t n, uint32_t n_spheres, const uint32_t* __restrict__ order, const uint32_t* __restrict__ leaf_parent, BvhNode* nodes) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n || order[slot] >= n_spheres) return;
  uint32_t* w = reinterpret_cast<uint32_t*>(&nodes[leaf_parent[slot]].n3);
  const uint32_t ref = PTB_LEAF_BIT | slot;
  if (w[0] == ref) w[0] = ref | kSphereBit;
  else w[1] = ref | kSphereBit;
}

}  // namespace

// Builds the SAH hierarchy over the primitives whose Morton order is order_a (n >= 2): links into nodes[].n3, leaf_parent,
// and the final primitive order, returned in *order_out (one of order_a / order_b). k_refit then fills the boxes.
// Scratch of an n-primitive SAH build (grow-only; called outside the timed part of the commit).
int32_t reserve_sah(Ctx* c, uint32_t n) {
  const uint32_t max_tasks = n / (kSahSmall + 1u) + 2u;
  DevBuf* b = c->sah_scratch;
  PTB_CUDA_TRY(c, b[0].reserve((size_t)n * 4));                                 // position -> task, ping
  PTB_CUDA_TRY(c, b[1].reserve((size_t)n * 4));                                 // ... pong
  PTB_CUDA_TRY(c, b[2].reserve((size_t)n * 4));                                 // left flags / their prefix sums
  PTB_CUDA_TRY(c, b[3].reserve((size_t)max_tasks * sizeof(SahTask)));           // tasks of the level, ping
  PTB_CUDA_TRY(c, b[4].reserve((size_t)max_tasks * sizeof(SahTask)));           // ... pong
  PTB_CUDA_TRY(c, b[5].reserve((size_t)max_tasks * kTaskBinWords * 4));         // bins
  PTB_CUDA_TRY(c, b[6].reserve((size_t)max_tasks * sizeof(SahSplit)));
  PTB_CUDA_TRY(c, b[7].reserve((size_t)max_tasks * 2 * sizeof(SahTask)));       // large children before compaction
  PTB_CUDA_TRY(c, b[8].reserve((size_t)max_tasks * 4));                         // large children per task / their bases
  PTB_CUDA_TRY(c, b[9].reserve(((size_t)n / 2 + 2) * sizeof(SahTask)));         // small tasks
  PTB_CUDA_TRY(c, b[10].reserve(16 * 4));                                       // counters + the root box
  PTB_CUDA_TRY(c, b[11].reserve(4100 * 4));                                     // block sums of the scan
  if (!c->h_sah) PTB_CUDA_TRY(c, cudaMallocHost(&c->h_sah, 4 * sizeof(uint32_t)));
  return PTB_OK;
}

int32_t build_sah(Ctx* c, const SahBuildInputs& in, const uint32_t** order_out) {
  const uint32_t n = in.n_prims;
  cudaStream_t st = c->stream;
  if (n > (1u << 24)) return set_error(c, PTB_ERR_UNSUPPORTED, "SAH build: more than 2^24 primitives");
  const uint32_t max_tasks = n / (kSahSmall + 1u) + 2u;
  {
    const int32_t rc = reserve_sah(c, n);
    if (rc != PTB_OK) return rc;
  }
  DevBuf &slot_task_a = c->sah_scratch[0], &slot_task_b = c->sah_scratch[1], &prefix = c->sah_scratch[2], &tasks_a = c->sah_scratch[3],
         &tasks_b = c->sah_scratch[4], &bins = c->sah_scratch[5], &splits = c->sah_scratch[6], &child_tasks = c->sah_scratch[7],
         &n_large = c->sah_scratch[8], &small = c->sah_scratch[9], &counters = c->sah_scratch[10], &block_sum = c->sah_scratch[11];

  // depth bound (test hook: PTB_SAH_MAX_DEPTH lowers it, never below what halving alone needs)
  uint32_t max_depth = kSahMaxDepth;
  if (const char* e = getenv("PTB_SAH_MAX_DEPTH")) {
    const int v = atoi(e);
    if (v > 0 && (uint32_t)v < kSahMaxDepth) max_depth = (uint32_t)v;
  }
  if (max_depth < sah_ceil_log2(n)) max_depth = sah_ceil_log2(n);
  const int T = 256;
  const uint32_t gn = (n + T - 1) / T;
  uint32_t* cnt = counters.as<uint32_t>();
  const uint32_t* box6 = in.box6;  // union of the primitive boxes, left by k_prim_bounds
  SahTask *cur = tasks_a.as<SahTask>(), *nxt = tasks_b.as<SahTask>();
  const bool root_small = n <= kSahSmall;
  k_sah_root<<<1, 1, 0, st>>>(box6, n, root_small ? small.as<SahTask>() : cur, cnt);
  c->stats.kernel_launches += 1;
  uint32_t *order = in.order_a, *order2 = in.order_b;
  uint32_t *stask = slot_task_a.as<uint32_t>(), *stask2 = slot_task_b.as<uint32_t>();
  uint32_t n_tasks = root_small ? 0u : 1u, n_small = root_small ? 1u : 0u;
  if (!root_small) {
    k_sah_fill<<<gn, T, 0, st>>>(stask, n, 0u);
    c->stats.kernel_launches += 1;
  }
  uint32_t levels = 0;
  while (n_tasks) {
    if (n_tasks > max_tasks) return set_error(c, PTB_ERR_INVALID, "SAH build: %u tasks in a level (capacity %u)", n_tasks, max_tasks);
    k_sah_clear<<<(n_tasks * kTaskBinWords + T - 1) / T, T, 0, st>>>(bins.as<uint32_t>(), n_tasks);
    k_sah_bin<<<gn, T, 0, st>>>(n, order, stask, cur, in.bmin, in.bmax, bins.as<uint32_t>());
    k_sah_eval<<<(n_tasks + 127) / 128, 128, 0, st>>>(n_tasks, cur, bins.as<uint32_t>(), splits.as<SahSplit>(), child_tasks.as<SahTask>(),
                                                      n_large.as<uint32_t>(), in.nodes, in.leaf_parent, small.as<SahTask>(), cnt, max_depth);
    k_sah_task_scan<<<1, 1024, 0, st>>>(n_large.as<uint32_t>(), n_tasks, cnt);
    k_sah_emit<<<(n_tasks + 127) / 128, 128, 0, st>>>(n_tasks, n_large.as<uint32_t>(), child_tasks.as<SahTask>(), splits.as<SahSplit>(), nxt);
    const uint32_t n_tiles = (n + kFlagTile - 1) / kFlagTile;
    k_sah_flag_scan<<<n_tiles, kFlagBlock, 0, st>>>(n, order, stask, splits.as<SahSplit>(), in.bmin, in.bmax, prefix.as<uint32_t>(),
                                                    block_sum.as<uint32_t>());
    scan_block_sums(c, block_sum.as<uint32_t>(), n_tiles, cnt + 3);
    k_sah_scatter<<<gn, T, 0, st>>>(n, order, stask, splits.as<SahSplit>(), in.bmin, in.bmax, prefix.as<uint32_t>(),
                                    block_sum.as<uint32_t>(), order2, stask2);
    c->stats.kernel_launches += 7;
    PTB_CUDA_TRY(c, cudaMemcpyAsync(c->h_sah, cnt, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PTB_CUDA_TRY(c, cudaStreamSynchronize(st));
    n_small = c->h_sah[0];
    n_tasks = c->h_sah[2];
    { uint32_t* t = order; order = order2; order2 = t; }
    { uint32_t* t = stask; stask = stask2; stask2 = t; }
    { SahTask* t = cur; cur = nxt; nxt = t; }
    if (++levels > 4096u) return set_error(c, PTB_ERR_INVALID, "SAH build: more than 4096 levels");
  }
  if (n_small) {
    k_sah_small<<<(n_small + kSmallWarps - 1) / kSmallWarps, 32 * kSmallWarps, 0, st>>>(n_small, small.as<SahTask>(), order, in.bmin, in.bmax,
                                                                                        in.nodes, in.leaf_parent, cnt, max_depth);
    c->stats.kernel_launches += 1;
  }
  k_sah_finish<<<gn, T, 0, st>>>(n, in.n_spheres, order, in.leaf_parent, in.nodes);
  c->stats.kernel_launches += 1;
  c->sah_levels = levels;
  *order_out = order;
  return PTB_OK;
}

}  // namespace ptb
