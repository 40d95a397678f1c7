// ptb200 — device build of the compressed 8-wide BVH (sm_100a): the "SAH collapse to a wide BVH" of the north star.
// Input: the LBVH of lbvh_build.cu (bit-exact against oracle/lbvh_ref.hpp). Output: 96-byte CwNodes + the final primitive
// order, bit-exact against oracle/cwbvh_ref.hpp (tests/test_gpu_lbvh.py) — every step below is a pure function of the LBVH:
//   W1  k_cw_level     top-down, one launch per wide level: a wide root starts from the two children of its LBVH node and
//                      repeatedly opens the inner child with the LARGEST SURFACE AREA (the SAH's "most likely to be
//                      entered"; the reference's cost model is traversal 0.125 / intersection 1, split.rs:161-163,176)
//                      until it has 8 children or only leaf groups (LBVH subtrees of <= max_leaf primitives);
//   W2  k_cw_place     children -> slots by the side of the node centre they lie on (greedy maximum of
//                      (centroid - centre) . (+-1, +-1, +-1)): octant-ordered traversal needs no distance sort;
//                      inner-child and primitive counts per wide root;
//   W3  k_scan_*       exclusive scans of both counts over the LBVH index: node index = 1 + inner children of all earlier
//                      wide roots (+ rank among its siblings), primitive base likewise — deterministic, no atomics;
//   W4  k_cw_assign    a node's index is handed down by its parent;
//   W5  k_cw_write     quantisation (8 bits per plane on a power-of-two grid, rounded outwards and verified against the
//                      f32 decode), meta bytes, the node itself, and the final primitive order of its leaf groups.
// The launch order of W1's levels is the only host-driven loop (one 4-byte read-back per level, ~20 levels at 1 M
// triangles); the order in which a level's threads append to the next frontier does not matter, every output is indexed
// by LBVH node.
#include "ptb_internal.h"

namespace ptb {

constexpr uint32_t kCwEmpty = 0xFFFFFFFEu;

struct CwIn {
  const BvhNode* nodes;
  const uint2* range;
  const float4 *nbmin, *nbmax, *bmin, *bmax;
  const uint32_t* prim_sorted;
  uint32_t max_leaf;
};
__device__ __forceinline__ uint32_t cw_count(const CwIn& in, uint32_t ref) {
  if (ref & PTB_LEAF_BIT) return 1u;
  const uint2 r = in.range[ref];
  return r.y - r.x + 1u;
}
__device__ __forceinline__ uint32_t cw_first(const CwIn& in, uint32_t ref) {
  return (ref & PTB_LEAF_BIT) ? (ref & kSlotMask) : in.range[ref].x;
}
__device__ __forceinline__ bool cw_group(const CwIn& in, uint32_t ref) { return cw_count(in, ref) <= in.max_leaf; }
__device__ __forceinline__ void cw_box(const CwIn& in, uint32_t ref, v3& mn, v3& mx) {
  if (ref & PTB_LEAF_BIT) {
    const uint32_t p = in.prim_sorted[ref & kSlotMask];
    mn = from4(in.bmin[p]);
    mx = from4(in.bmax[p]);
  } else {
    mn = from4(in.nbmin[ref]);
    mx = from4(in.nbmax[ref]);
  }
}
__device__ __forceinline__ float cw_area(v3 mn, v3 mx) {  // 2 (dx dy + dy dz + dz dx), the shape of aabb.rs:69-73
  const float dx = mx.x - mn.x, dy = mx.y - mn.y, dz = mx.z - mn.z;
  return 2.0f * (dx * dy + dy * dz + dz * dx);
}

// ------------------------------------------------------------------------------------------ W1
__global__ void __launch_bounds__(128) k_cw_level(CwIn in, const uint32_t* __restrict__ frontier, uint32_t n_frontier,
                                                   uint32_t* __restrict__ next, uint32_t* n_next, uint32_t* __restrict__ wchild,
                                                   uint32_t* __restrict__ wcount) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_frontier) return;
  const uint32_t r = frontier[i];
  uint32_t child[8];
  int n = 2;
  child[0] = in.nodes[r].n3.x & ~kSphereBit;
  child[1] = in.nodes[r].n3.y & ~kSphereBit;
  while (n < 8) {
    int best = -1;
    float best_sa = -1.0f;
    for (int k = 0; k < n; ++k) {
      if (cw_group(in, child[k])) continue;
      v3 mn, mx;
      cw_box(in, child[k], mn, mx);
      const float sa = cw_area(mn, mx);
      if (best < 0 || sa > best_sa || (sa == best_sa && child[k] < child[best])) { best = k; best_sa = sa; }
    }
    if (best < 0) break;
    const uint32_t open = child[best];
    child[best] = in.nodes[open].n3.x & ~kSphereBit;
    child[n++] = in.nodes[open].n3.y & ~kSphereBit;
  }
  uint32_t inner = 0;
  for (int k = 0; k < 8; ++k) wchild[8u * (size_t)r + k] = k < n ? child[k] : kCwEmpty;
  for (int k = 0; k < n; ++k) inner += cw_group(in, child[k]) ? 0u : 1u;
  wcount[r] = (uint32_t)n;
  if (inner) {
    uint32_t at = atomicAdd(n_next, inner);
    for (int k = 0; k < n; ++k)
      if (!cw_group(in, child[k])) next[at++] = child[k];
  }
}

// ------------------------------------------------------------------------------------------ W2
// One thread per LBVH node; only wide roots (wcount != 0) work. Rewrites wchild in SLOT order.
__global__ void __launch_bounds__(128) k_cw_place(CwIn in, uint32_t n_bin, uint32_t* __restrict__ wchild,
                                                   const uint32_t* __restrict__ wcount, uint32_t* __restrict__ inner_cnt,
                                                   uint32_t* __restrict__ prim_cnt) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_bin) return;
  const int n = (int)wcount[r];
  if (n == 0) { inner_cnt[r] = 0u; prim_cnt[r] = 0u; return; }
  uint32_t child[8];
  v3 cen[8];
  const float inf = __int_as_float(0x7f800000);
  v3 nmn = mk(inf, inf, inf), nmx = mk(-inf, -inf, -inf);
  for (int k = 0; k < n; ++k) {
    child[k] = wchild[8u * (size_t)r + k];
    v3 mn, mx;
    cw_box(in, child[k], mn, mx);
    nmn = vmin(nmn, mn);
    nmx = vmax(nmx, mx);
    cen[k] = 0.5f * (mn + mx);
  }
  const v3 centre = 0.5f * (nmn + nmx);
  uint32_t placed[8];
  for (int s = 0; s < 8; ++s) placed[s] = kCwEmpty;
  uint32_t done = 0u;  // bit k: child k has its slot
  for (int round = 0; round < n; ++round) {
    int bk = -1, bs = -1;
    float bc = 0.0f;
    for (int k = 0; k < n; ++k) {
      if ((done >> k) & 1u) continue;
      const v3 d = cen[k] - centre;
      for (int s = 0; s < 8; ++s) {
        if (placed[s] != kCwEmpty) continue;
        const float cost = ((s & 1) ? d.x : -d.x) + ((s & 2) ? d.y : -d.y) + ((s & 4) ? d.z : -d.z);
        if (bk < 0 || cost > bc) { bk = k; bs = s; bc = cost; }
      }
    }
    placed[bs] = child[bk];
    done |= 1u << bk;
  }
  uint32_t inner = 0, prims = 0;
  for (int s = 0; s < 8; ++s) {
    wchild[8u * (size_t)r + s] = placed[s];
    if (placed[s] == kCwEmpty) continue;
    if (cw_group(in, placed[s])) prims += cw_count(in, placed[s]);
    else ++inner;
  }
  inner_cnt[r] = inner;
  prim_cnt[r] = prims;
}

// ------------------------------------------------------------------------------------------ W3: exclusive scan (uint32)
// 4096 elements per block: local exclusive scan in place + block total; one block then scans the totals (<= 4096 blocks,
// i.e. 16 Mi elements — more LBVH nodes than the 2^30-slot limit ever needs would take a third level); then the add.
constexpr int kScanBlock = 1024, kScanItems = 4, kScanTile = kScanBlock * kScanItems;
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t mine, uint32_t* warp_sum, uint32_t& total) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += t;
  }
  if (lane == 31u) warp_sum[warp] = inc;
  __syncthreads();
  if (warp == 0u) {
    const uint32_t v = warp_sum[lane];
    uint32_t winc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= (uint32_t)o) winc += t;
    }
    warp_sum[lane] = winc - v;
    if (lane == 31u) warp_sum[32] = winc;
  }
  __syncthreads();
  total = warp_sum[32];
  return warp_sum[warp] + inc - mine;
}
__global__ void __launch_bounds__(kScanBlock) k_scan_local(uint32_t* __restrict__ data, uint32_t n, uint32_t* __restrict__ block_sum) {
  __shared__ uint32_t warp_sum[33];
  const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  uint32_t v[kScanItems], mine = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { v[k] = base + k < n ? data[base + k] : 0u; mine += v[k]; }
  uint32_t total;
  uint32_t run = block_exclusive_scan(mine, warp_sum, total);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) data[base + k] = run;
    run += v[k];
  }
  if (threadIdx.x == 0) block_sum[blockIdx.x] = total;
}
__global__ void __launch_bounds__(kScanBlock) k_scan_sums(uint32_t* __restrict__ block_sum, uint32_t n_blocks, uint32_t* total_out) {
  __shared__ uint32_t warp_sum[33];
  const uint32_t base = threadIdx.x * kScanItems;
  uint32_t v[kScanItems], mine = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { v[k] = base + k < n_blocks ? block_sum[base + k] : 0u; mine += v[k]; }
  uint32_t total;
  uint32_t run = block_exclusive_scan(mine, warp_sum, total);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n_blocks) block_sum[base + k] = run;
    run += v[k];
  }
  if (threadIdx.x == 0) *total_out = total;
}
__global__ void __launch_bounds__(kScanBlock) k_scan_add(uint32_t* __restrict__ data, uint32_t n, const uint32_t* __restrict__ block_sum) {
  const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  const uint32_t add = block_sum[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < n) data[base + k] += add;
}
// the middle step alone: exclusive scan of n_blocks <= 4096 block totals in place (the caller scanned its blocks itself)
void scan_block_sums(Ctx* c, uint32_t* block_sum, uint32_t n_blocks, uint32_t* total_out) {
  k_scan_sums<<<1, kScanBlock, 0, c->stream>>>(block_sum, n_blocks, total_out);
  c->stats.kernel_launches += 1;
}
int32_t exclusive_scan(Ctx* c, uint32_t* data, uint32_t n, uint32_t* block_sum, uint32_t* total_out) {
  const uint32_t n_blocks = (n + kScanTile - 1) / kScanTile;
  if (n_blocks > (uint32_t)kScanTile) return set_error(c, PTB_ERR_INVALID, "scan of %u elements needs a third level", n);
  k_scan_local<<<n_blocks, kScanBlock, 0, c->stream>>>(data, n, block_sum);
  k_scan_sums<<<1, kScanBlock, 0, c->stream>>>(block_sum, n_blocks, total_out);
  k_scan_add<<<n_blocks, kScanBlock, 0, c->stream>>>(data, n, block_sum);
  c->stats.kernel_launches += 3;
  return PTB_OK;
}

// ------------------------------------------------------------------------------------------ W4
__global__ void __launch_bounds__(128) k_cw_assign(CwIn in, uint32_t n_bin, const uint32_t* __restrict__ wchild,
                                                    const uint32_t* __restrict__ wcount, const uint32_t* __restrict__ inner_off,
                                                    uint32_t* __restrict__ widx) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_bin || wcount[r] == 0u) return;
  if (r == 0u) widx[0] = 0u;
  uint32_t rank = 0;
  for (int s = 0; s < 8; ++s) {
    const uint32_t ref = wchild[8u * (size_t)r + s];
    if (ref == kCwEmpty || cw_group(in, ref)) continue;
    widx[ref] = 1u + inner_off[r] + rank++;
  }
}

// ------------------------------------------------------------------------------------------ W5
// quantisation helpers — oracle/cwbvh_ref.hpp cell_exponent / quantise_lo / quantise_hi, operation for operation
__device__ __forceinline__ uint32_t cw_cell_exponent(float extent) {
  const float x = extent / 255.0f;
  const uint32_t b = __float_as_uint(x);
  const uint32_t E = (b >> 23) & 255u, m = b & 0x7FFFFFu;
  uint32_t eb = m ? E + 1u : E;
  if (eb < 1u) eb = 1u;
  if (eb > 254u) eb = 254u;
  return eb;
}
__device__ __forceinline__ uint32_t cw_quantise_lo(float lo, float origin, float cell) {
  float f = floorf((lo - origin) / cell);
  if (!(f > 0.0f)) f = 0.0f;
  if (f > 255.0f) f = 255.0f;
  uint32_t q = (uint32_t)f;
  while (q > 0u && origin + (float)q * cell > lo) --q;
  return q;
}
__device__ __forceinline__ uint32_t cw_quantise_hi(float hi, float origin, float cell) {
  float f = ceilf((hi - origin) / cell);
  if (!(f > 0.0f)) f = 0.0f;
  if (f > 256.0f) f = 256.0f;
  uint32_t q = (uint32_t)f;
  while (q < 256u && origin + (float)q * cell < hi) ++q;
  return q;
}
// Writes the node of wide root `r` (children already in slot order in `placed`) at `at`.
__device__ void cw_write_node(const CwIn& in, const uint32_t placed[8], uint32_t at, uint32_t child_base, uint32_t prim_base,
                              CwNode* __restrict__ out, uint32_t* __restrict__ slot_morton) {
  v3 cmn[8], cmx[8];
  const float inf = __int_as_float(0x7f800000);
  v3 nmn = mk(inf, inf, inf), nmx = mk(-inf, -inf, -inf);
  for (int s = 0; s < 8; ++s) {
    if (placed[s] == kCwEmpty) continue;
    cw_box(in, placed[s], cmn[s], cmx[s]);
    nmn = vmin(nmn, cmn[s]);
    nmx = vmax(nmx, cmx[s]);
  }
  CwNode nd;
  nd.p[0] = nmn.x; nd.p[1] = nmn.y; nd.p[2] = nmn.z;
  const float origin[3] = {nmn.x, nmn.y, nmn.z};
  const float ext[3] = {nmx.x - nmn.x, nmx.y - nmn.y, nmx.z - nmn.z};
  uint32_t e_imask = 0u;
  for (int w = 0; w < 12; ++w) nd.q[w] = 0u;
  for (int a = 0; a < 3; ++a) {
    uint32_t eb = cw_cell_exponent(ext[a]);
    for (;;) {  // the f32 division may land one binade low: widen until every plane fits 8 bits
      bool fits = true;
      const float cell = __uint_as_float(eb << 23);
      for (int s = 0; s < 8 && fits; ++s) {
        if (placed[s] == kCwEmpty) continue;
        const float hi = a == 0 ? cmx[s].x : a == 1 ? cmx[s].y : cmx[s].z;
        if (cw_quantise_hi(hi, origin[a], cell) > 255u) fits = false;
      }
      if (fits) break;
      ++eb;
    }
    e_imask |= eb << (8 * a);
    const float cell = __uint_as_float(eb << 23);
    for (int s = 0; s < 8; ++s) {
      uint32_t qlo = 255u, qhi = 0u;  // empty slot: inverted
      if (placed[s] != kCwEmpty) {
        const float lo = a == 0 ? cmn[s].x : a == 1 ? cmn[s].y : cmn[s].z;
        const float hi = a == 0 ? cmx[s].x : a == 1 ? cmx[s].y : cmx[s].z;
        qlo = cw_quantise_lo(lo, origin[a], cell);
        qhi = cw_quantise_hi(hi, origin[a], cell);
      }
      nd.q[2 * a + (s >> 2)] |= qlo << (8 * (s & 3));
      nd.q[6 + 2 * a + (s >> 2)] |= qhi << (8 * (s & 3));
    }
  }
  nd.child_base = child_base;
  nd.prim_base = prim_base;
  nd.meta[0] = nd.meta[1] = 0u;
  uint32_t off = 0;
  for (int s = 0; s < 8; ++s) {
    const uint32_t ref = placed[s];
    if (ref == kCwEmpty) continue;
    uint32_t meta;
    if (!cw_group(in, ref)) {
      e_imask |= 1u << (24 + s);
      meta = 0x20u | (24u + (uint32_t)s);
    } else {
      const uint32_t cnt = cw_count(in, ref), first = cw_first(in, ref);
      meta = (((1u << cnt) - 1u) << 5) | off;
      for (uint32_t j = 0; j < cnt; ++j) slot_morton[prim_base + off + j] = first + j;
      off += cnt;
    }
    nd.meta[s >> 2] |= meta << (8 * (s & 3));
  }
  nd.e_imask = e_imask;
  nd.pad[0] = nd.pad[1] = nd.pad[2] = nd.pad[3] = 0u;
  out[at] = nd;
}
__global__ void __launch_bounds__(128) k_cw_write(CwIn in, uint32_t n_bin, const uint32_t* __restrict__ wchild,
                                                   const uint32_t* __restrict__ wcount, const uint32_t* __restrict__ inner_off,
                                                   const uint32_t* __restrict__ prim_off, const uint32_t* __restrict__ widx,
                                                   CwNode* __restrict__ out, uint32_t* __restrict__ slot_morton) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_bin || wcount[r] == 0u) return;
  uint32_t placed[8];
  for (int s = 0; s < 8; ++s) placed[s] = wchild[8u * (size_t)r + s];
  cw_write_node(in, placed, widx[r], 1u + inner_off[r], prim_off[r], out, slot_morton);
}
// The whole scene fits one leaf group (n <= max_leaf, including n == 1): a single node with one leaf child.
__global__ void k_cw_whole_tree(CwIn in, uint32_t n_prims, CwNode* __restrict__ out, uint32_t* __restrict__ slot_morton) {
  uint32_t placed[8];
  for (int s = 0; s < 8; ++s) placed[s] = kCwEmpty;
  // slot of the only child: its centroid IS the node centre, every cost is 0, the greedy takes slot 0
  placed[0] = n_prims == 1u ? (PTB_LEAF_BIT | 0u) : 0u;
  cw_write_node(in, placed, 0u, 1u, 0u, out, slot_morton);
}
__global__ void __launch_bounds__(256) k_cw_final_prims(const uint32_t* __restrict__ slot_morton, const uint32_t* __restrict__ prim_sorted,
                                                         uint32_t n, uint32_t* __restrict__ final_prim) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) final_prim[i] = prim_sorted[slot_morton[i]];
}

// ------------------------------------------------------------------------------------------ host
int32_t build_wide(Ctx* c, const CwBuildInputs& bi, uint32_t* final_prim) {
  const uint32_t n = bi.n_prims;
  cudaStream_t st = c->stream;
  CwIn in{bi.nodes, bi.range, bi.nbmin, bi.nbmax, bi.bmin, bi.bmax, bi.prim_sorted, c->cw_max_leaf};
  const uint32_t n_bin = n > 1u ? n - 1u : 0u;
  DevBuf &wchild = c->cw_scratch[0], &wcount = c->cw_scratch[1], &inner_cnt = c->cw_scratch[2], &prim_cnt = c->cw_scratch[3],
         &widx = c->cw_scratch[4], &frontier = c->cw_scratch[5], &slot_morton = c->cw_scratch[6], &misc = c->cw_scratch[7];
  PTB_CUDA_TRY(c, slot_morton.reserve((size_t)n * 4));
  PTB_CUDA_TRY(c, misc.reserve(((size_t)kScanTile + 16) * 4));  // scan block sums + counters
  uint32_t* d_misc = misc.as<uint32_t>();
  uint32_t* d_counters = d_misc + kScanTile;  // [0] next-frontier size, [1] total inner children, [2] total primitives
  if (n <= c->cw_max_leaf) {
    PTB_CUDA_TRY(c, c->d_cw_nodes.reserve(sizeof(CwNode)));
    k_cw_whole_tree<<<1, 1, 0, st>>>(in, n, c->d_cw_nodes.as<CwNode>(), slot_morton.as<uint32_t>());
    c->stats.kernel_launches += 1;
    c->n_cw_nodes = 1;
  } else {
    PTB_CUDA_TRY(c, wchild.reserve((size_t)n_bin * 32));
    PTB_CUDA_TRY(c, wcount.reserve((size_t)n_bin * 4));
    PTB_CUDA_TRY(c, inner_cnt.reserve((size_t)n_bin * 4));
    PTB_CUDA_TRY(c, prim_cnt.reserve((size_t)n_bin * 4));
    PTB_CUDA_TRY(c, widx.reserve((size_t)n_bin * 4));
    PTB_CUDA_TRY(c, frontier.reserve((size_t)n_bin * 8));
    PTB_CUDA_TRY(c, cudaMemsetAsync(wcount.p, 0, (size_t)n_bin * 4, st));
    // ---- W1: level by level from the root
    uint32_t* fa = frontier.as<uint32_t>();
    uint32_t* fb = fa + n_bin;
    const uint32_t root = 0u;
    PTB_CUDA_TRY(c, cudaMemcpyAsync(fa, &root, 4, cudaMemcpyHostToDevice, st));
    uint32_t n_front = 1u;
    for (int level = 0; n_front != 0u; ++level) {
      if (level > 128) return set_error(c, PTB_ERR_INVALID, "wide collapse did not terminate");
      PTB_CUDA_TRY(c, cudaMemsetAsync(d_counters, 0, 4, st));
      k_cw_level<<<(n_front + 127u) / 128u, 128, 0, st>>>(in, fa, n_front, fb, d_counters, wchild.as<uint32_t>(), wcount.as<uint32_t>());
      c->stats.kernel_launches += 1;
      PTB_CUDA_TRY(c, cudaMemcpyAsync(&n_front, d_counters, 4, cudaMemcpyDeviceToHost, st));
      PTB_CUDA_TRY(c, cudaStreamSynchronize(st));
      uint32_t* t = fa; fa = fb; fb = t;
    }
    // ---- W2 .. W4
    const uint32_t gb = (n_bin + 127u) / 128u;
    k_cw_place<<<gb, 128, 0, st>>>(in, n_bin, wchild.as<uint32_t>(), wcount.as<uint32_t>(), inner_cnt.as<uint32_t>(), prim_cnt.as<uint32_t>());
    c->stats.kernel_launches += 1;
    int32_t rc = exclusive_scan(c, inner_cnt.as<uint32_t>(), n_bin, d_misc, d_counters + 1);
    if (rc == PTB_OK) rc = exclusive_scan(c, prim_cnt.as<uint32_t>(), n_bin, d_misc, d_counters + 2);
    if (rc != PTB_OK) return rc;
    uint32_t totals[2] = {0u, 0u};
    PTB_CUDA_TRY(c, cudaMemcpyAsync(totals, d_counters + 1, 8, cudaMemcpyDeviceToHost, st));
    PTB_CUDA_TRY(c, cudaStreamSynchronize(st));
    if (totals[1] != n) return set_error(c, PTB_ERR_INVALID, "wide collapse lost primitives (%u of %u)", totals[1], n);
    c->n_cw_nodes = 1u + (uint64_t)totals[0];
    PTB_CUDA_TRY(c, c->d_cw_nodes.reserve(c->n_cw_nodes * sizeof(CwNode)));
    k_cw_assign<<<gb, 128, 0, st>>>(in, n_bin, wchild.as<uint32_t>(), wcount.as<uint32_t>(), inner_cnt.as<uint32_t>(), widx.as<uint32_t>());
    // ---- W5
    k_cw_write<<<gb, 128, 0, st>>>(in, n_bin, wchild.as<uint32_t>(), wcount.as<uint32_t>(), inner_cnt.as<uint32_t>(), prim_cnt.as<uint32_t>(),
                                   widx.as<uint32_t>(), c->d_cw_nodes.as<CwNode>(), slot_morton.as<uint32_t>());
    c->stats.kernel_launches += 2;
  }
  k_cw_final_prims<<<(n + 255u) / 256u, 256, 0, st>>>(slot_morton.as<uint32_t>(), bi.prim_sorted, n, final_prim);
  c->stats.kernel_launches += 1;
  PTB_CUDA_TRY(c, cudaGetLastError());
  return PTB_OK;
}

}  // namespace ptb
