// Family 2: HBM-bound data movement.
//
//   t10_relayout_blocks  block-sparse strided permute between two sector layouts of the
//                        same symmetric tensor (reference: the interpreted element loop of
//                        UniTensor.Contiguous, UniTensor.py:2108-2129, ~4e4 elements/s).
//   t10_permute_dense    dense N-d strided permute into a contiguous tensor
//                        (reference: Tensor.contiguous(), UniTensor.py:2057/:2144, and the
//                        implicit copy inside torch.tensordot, :3769).
//   t10_diag_scale       row/column scaling by a diagonal operand (replaces torch.diag +
//                        dense GEMM, UniTensor.py:3756-3765).
//
// Relayout formulation (gather form): the source address of destination
// element (sector s, row r, col c) is additively separable into a row part and a column
// part BEFORE the block-map lookup:
//     src_row = RR[r] + CR[c]     src_col = RC[r] + CC[c]
//     addr    = rowbase[src_row] + coloff[src_col]
// RR/RC (per destination row) and CR/CC (per destination column) are small int32 tables
// built once per plan (t10_relayout_tables); rowbase/coloff belong to the source layout
// (t10_blockmap_build).  Every thread owns one destination column (coalesced stores, CR/CC
// in registers) and walks the rows of its tile with the loads of several rows in flight.
// Source sectors (32 B) touched by neighbouring rows of a tile are served from L1.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "t10_common.cuh"

namespace t10 {

constexpr int RL_THREADS = 256;                 // 8 warps: 4 column groups x 2 row halves
constexpr int RL_TR = T10_RELAYOUT_TR;          // 32 rows per tile
constexpr int RL_TC = T10_RELAYOUT_TC;          // 128 columns per tile
constexpr int RL_ROWS_PER_THREAD = RL_TR / (RL_THREADS / RL_TC);   // 16

// Dependent-load chain per tile (what bounds a 50 MB relayout after an L2 flush): tile record ->
// {RR/RC, CR/CC} -> {rowbase, coloff} -> source element.  The tile record carries everything the tile
// needs (no sector-table indirection) and a thread keeps all its 16 row gathers in flight at once.
template <typename T, bool CHECK, bool EMIT>
__global__ void __launch_bounds__(RL_THREADS)
k_relayout_blocks(const T* __restrict__ src, T* __restrict__ dst, int32_t* __restrict__ emit, int64_t ntiles,
                  const t10_relayout_tile* __restrict__ tiles,
                  const int32_t* __restrict__ RR, const int32_t* __restrict__ RC,
                  const int32_t* __restrict__ CR, const int32_t* __restrict__ CC,
                  const int64_t* __restrict__ rowbase, const int32_t* __restrict__ coloff,
                  const int64_t* __restrict__ ket_map, const int64_t* __restrict__ bra_map) {
  __shared__ int32_t s_rr[RL_TR], s_rc[RL_TR];
  const int tc = threadIdx.x & (RL_TC - 1);
  const int rbeg = (threadIdx.x / RL_TC) * RL_ROWS_PER_THREAD;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const t10_relayout_tile tile = tiles[t];
    const bool col_ok = tc < tile.nc;
    int cr = 0, cc = 0;
    if (col_ok) {
      cr = CR[tile.cbase + tc];
      cc = CC[tile.cbase + tc];
    }
    __syncthreads();  // previous tile done with s_rr/s_rc
    if (threadIdx.x < RL_TR) {
      const int r = threadIdx.x;
      s_rr[r] = r < tile.nr ? RR[tile.rbase + r] : 0;
      s_rc[r] = r < tile.nr ? RC[tile.rbase + r] : 0;
    }
    __syncthreads();
    if (!col_ok) continue;
    const int rend = min(tile.nr, rbeg + RL_ROWS_PER_THREAD);
    int64_t addr[RL_ROWS_PER_THREAD];
#pragma unroll
    for (int u = 0; u < RL_ROWS_PER_THREAD; ++u) {
      const int r = rbeg + u;
      addr[u] = -1;
      if (r < rend) {
        const int srow = s_rr[r] + cr, scol = s_rc[r] + cc;
        const int64_t base = rowbase[srow];
        const int32_t co = coloff[scol];
        bool ok = base >= 0 && co >= 0;
        if (CHECK && ok)  // Zn layouts (quirk Q3): row and column may sit in different sectors
          ok = ket_map[2 * (int64_t)srow] == bra_map[2 * (int64_t)scol];
        if (ok) addr[u] = base + co;
      }
    }
    if constexpr (EMIT) {  // plan time: record the source element index of every destination element
      int32_t* out = emit + tile.dst + tc;
#pragma unroll
      for (int u = 0; u < RL_ROWS_PER_THREAD; ++u)
        if (rbeg + u < rend) out[(int64_t)(rbeg + u) * tile.dld] = (int32_t)addr[u];
    } else {
      T v[RL_ROWS_PER_THREAD];
#pragma unroll
      for (int u = 0; u < RL_ROWS_PER_THREAD; ++u) v[u] = addr[u] >= 0 ? src[addr[u]] : T(0);
      T* out = dst + tile.dst + tc;
#pragma unroll
      for (int u = 0; u < RL_ROWS_PER_THREAD; ++u)
        if (rbeg + u < rend) out[(int64_t)(rbeg + u) * tile.dld] = v[u];
    }
  }
}

// Replay of a relayout through its precomputed element index (built once per plan by the kernel above
// in EMIT mode): dst[i] = src[index[i]] (index -1: write 0, -2: leave untouched).  Dependent-load chain
// of length two (index -> element), no tiles, no shared memory: this is what runs on the hot path when
// the index (4 B per destination element) fits the plan-memory budget.
constexpr int RI_THREADS = 256;
constexpr int RI_UNROLL = 4;
template <typename T>
__global__ void __launch_bounds__(RI_THREADS)
k_relayout_index(const T* __restrict__ src, T* __restrict__ dst, const int2* __restrict__ index, int64_t npairs) {
  struct alignas(2 * sizeof(T)) Pair { T x, y; };
  const int64_t stride = (int64_t)gridDim.x * RI_THREADS;
  for (int64_t i0 = (int64_t)blockIdx.x * RI_THREADS + threadIdx.x; i0 < npairs; i0 += stride * RI_UNROLL) {
    int2 ix[RI_UNROLL];
#pragma unroll
    for (int u = 0; u < RI_UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      ix[u] = i < npairs ? index[i] : make_int2(-2, -2);
    }
    Pair v[RI_UNROLL];
#pragma unroll
    for (int u = 0; u < RI_UNROLL; ++u) {
      v[u].x = ix[u].x >= 0 ? src[ix[u].x] : T(0);
      v[u].y = ix[u].y >= 0 ? src[ix[u].y] : T(0);
    }
#pragma unroll
    for (int u = 0; u < RI_UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= npairs) continue;
      if (ix[u].x != -2 && ix[u].y != -2) {
        *reinterpret_cast<Pair*>(dst + 2 * i) = v[u];
      } else {
        if (ix[u].x != -2) dst[2 * i] = v[u].x;
        if (ix[u].y != -2) dst[2 * i + 1] = v[u].y;
      }
    }
  }
}

// Replay through RUNS: a relayout that only regroups bonds keeps the innermost bond's charge sectors
// contiguous, so the element index is piecewise linear -- C3a: 16 376 runs of 190 elements on average for
// 3.1 M elements.  The plan stores (dst_start, src_start) per run (sorted by dst_start, sentinel at the end)
// and, per 1024-element destination chunk, the run that contains its first element.  A thread finds the run
// of its first element by bisection inside the chunk's few runs (L1-resident), walks forward for its four
// elements, and moves 32 bytes: coalesced on both sides, no per-element index traffic (12.5 MB of the
// 62 MB the index replay moves).  src_start -1: zero fill, -2: leave the destination untouched.
constexpr int RR_THREADS = 256, RR_PER_THREAD = 4, RR_CHUNK = RR_THREADS * RR_PER_THREAD, RR_MAXR = 254;
template <typename T>
__global__ void __launch_bounds__(RR_THREADS)
k_relayout_runs(const T* __restrict__ src, T* __restrict__ dst, int64_t n, const int2* __restrict__ runs,
                const int* __restrict__ chunk_first, int64_t nchunks) {
  __shared__ int2 s_runs[RR_MAXR + 2];
  for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const int64_t i0 = chunk * RR_CHUNK + (int64_t)threadIdx.x * RR_PER_THREAD;
    const int first = chunk_first[chunk], last = chunk_first[chunk + 1];
    const int cnt = last - first + 2;                 // the chunk's runs plus the one that follows
    const bool cached = cnt <= RR_MAXR + 2;
    const int2* tab = runs + first;                   // table this chunk searches: shared copy when it fits
    if (cached) {
      __syncthreads();                                // previous chunk's readers are done
      for (int k = threadIdx.x; k < cnt; k += RR_THREADS) s_runs[k] = runs[first + k];
      __syncthreads();
      tab = s_runs;
    }
    if (i0 >= n) continue;
    int lo = 0, hi = last - first;
    while (lo < hi) {                                 // last run with dst_start <= i0
      const int mid = (lo + hi + 1) >> 1;
      if (tab[mid].x <= i0) lo = mid; else hi = mid - 1;
    }
    int r = lo;
    int2 run = tab[r];
    int next = tab[r + 1].x;
    T v[RR_PER_THREAD];
    int kind[RR_PER_THREAD];                          // 1 store, 0 skip
#pragma unroll
    for (int j = 0; j < RR_PER_THREAD; ++j) {
      const int64_t i = i0 + j;
      kind[j] = 0;
      v[j] = T(0);
      if (i < n) {
        while (i >= next) {
          ++r;
          run = tab[r];
          next = tab[r + 1].x;
        }
        if (run.y >= 0) {
          v[j] = src[(int64_t)run.y + (i - run.x)];
          kind[j] = 1;
        } else if (run.y == -1) {
          kind[j] = 1;
        }
      }
    }
    bool all = true;
#pragma unroll
    for (int j = 0; j < RR_PER_THREAD; ++j) all = all && kind[j] == 1;
    if (all) {
      if (sizeof(T) == 8) {
        struct alignas(16) P2 { T a, b; };
        *reinterpret_cast<P2*>(dst + i0) = P2{v[0], v[1]};
        *reinterpret_cast<P2*>(dst + i0 + 2) = P2{v[2], v[3]};
      } else {
        struct alignas(16) P4 { T a, b, c, d; };
        *reinterpret_cast<P4*>(dst + i0) = P4{v[0], v[1], v[2], v[3]};
      }
    } else {
#pragma unroll
      for (int j = 0; j < RR_PER_THREAD; ++j)
        if (kind[j] == 1) dst[i0 + j] = v[j];
    }
  }
}

// Replay through SEGMENTS (default for relayouts that keep long runs): the plan lists every maximal piece of the
// element index that is contiguous on both sides, cut at 256 elements, as {dst, src, len, source}.  One WARP per
// segment: the descriptor is one broadcast load, then every lane has up to eight independent loads in flight
// (2 KB per warp, 128 KB per SM at full occupancy) and the stores follow -- no per-element index traffic
// (12.5 MB of the 62 MB the index replay moves on C3a), no bisection, and a dependent chain of two loads.
// `source` picks the base pointer of the arena the piece is read from: 0 on one GPU; in a sector-resident
// multi-GPU run the pieces owned by another rank are read through that rank's NVLink peer mapping, so the
// exchange a relayout of a sharded tensor needs is part of the copy itself (no collective, no staging).
// src < 0: zero fill.  The kernel opens the programmatic launch window of the GEMM that follows it.
constexpr int RS_THREADS = 512, RS_MAXLEN = 256, RS_MAXSRC = 8;
struct SegSources { const void* base[RS_MAXSRC]; };
template <typename T>
__global__ void __launch_bounds__(RS_THREADS)
k_relayout_segments(const __grid_constant__ SegSources srcs, T* __restrict__ dst, const int4* __restrict__ seg,
                    int nseg) {
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  const int lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * (RS_THREADS / 32);
  int s = (blockIdx.x * RS_THREADS + threadIdx.x) >> 5;
  if (s >= nseg) return;
  int4 d = __ldg(seg + s);                              // {dst, src, len, source}
  for (; s < nseg; s += nwarps) {
    // the next descriptor is in flight while this segment's data moves (one round trip less per segment)
    const int sn = s + nwarps;
    int4 dn = d;
    if (sn < nseg) dn = __ldg(seg + sn);
    T* o = dst + d.x;
    if (d.y < 0) {
#pragma unroll
      for (int u = 0; u < RS_MAXLEN / 32; ++u)
        if (lane + 32 * u < d.z) o[lane + 32 * u] = T(0);
    } else {
      const T* in = reinterpret_cast<const T*>(srcs.base[d.w]) + d.y;
      T v[RS_MAXLEN / 32];
#pragma unroll
      for (int u = 0; u < RS_MAXLEN / 32; ++u)
        if (lane + 32 * u < d.z) v[u] = in[lane + 32 * u];
#pragma unroll
      for (int u = 0; u < RS_MAXLEN / 32; ++u)
        if (lane + 32 * u < d.z) o[lane + 32 * u] = v[u];
    }
    d = dn;
  }
}

// ---------------------------------------------------------------- dense permute ----
struct PermDesc {
  int rank;                       // collapsed rank, dst order (last = dst innermost)
  int64_t shape[T10_MAX_RANK];
  int64_t sstride[T10_MAX_RANK];  // source stride of each dst dim (elements)
  int64_t total;
};

// dst innermost dim is also source-contiguous: vectorised row copy
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
k_permute_rows(const T* __restrict__ src, T* __restrict__ dst, const __grid_constant__ PermDesc pd) {
  const int64_t inner = pd.shape[pd.rank - 1] / VEC;  // vectors per row
  const int64_t nvec = pd.total / VEC;
  struct alignas(sizeof(T) * VEC) Vec { T v[VEC]; };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / inner, iv = i - row * inner;
    int64_t soff = 0;
#pragma unroll 1
    for (int d = pd.rank - 2; d >= 0; --d) {
      int64_t q = row / pd.shape[d];
      soff += (row - q * pd.shape[d]) * pd.sstride[d];
      row = q;
    }
    *reinterpret_cast<Vec*>(dst + i * VEC) = *reinterpret_cast<const Vec*>(src + soff + iv * VEC);
  }
}

// general case: 32x32 tile transposed through shared memory.  Tile axes: X = dst innermost
// dim (coalesced stores), Y = the dst dim whose source stride is 1 (coalesced loads).
template <typename T>
__global__ void __launch_bounds__(256)
k_permute_tiled(const T* __restrict__ src, T* __restrict__ dst, const __grid_constant__ PermDesc pd,
                int ydim, int64_t tiles_x, int64_t tiles_y, int64_t ntiles) {
  __shared__ T tile[32][33];
  const int X = pd.rank - 1;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    int64_t rem = t;
    const int64_t bx = rem % tiles_x;
    rem /= tiles_x;
    const int64_t by = rem % tiles_y;
    rem /= tiles_y;
    // remaining dims (all but X and ydim), row-major over dst order
    int64_t soff = 0, doff = 0, dstride = pd.shape[X];
#pragma unroll 1
    for (int d = pd.rank - 2; d >= 0; --d) {
      if (d != ydim) {
        int64_t q = rem / pd.shape[d];
        int64_t i = rem - q * pd.shape[d];
        soff += i * pd.sstride[d];
        doff += i * dstride;
        rem = q;
      }
      dstride *= pd.shape[d];
    }
    // dst stride of ydim
    int64_t ystride_d = pd.shape[X];
    for (int d = pd.rank - 2; d > ydim; --d) ystride_d *= pd.shape[d];
    const int64_t x0 = bx * 32, y0 = by * 32;
    __syncthreads();
    // load: threads fastest along Y (source contiguous)
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      int64_t x = x0 + ty + j, y = y0 + tx;
      if (x < pd.shape[X] && y < pd.shape[ydim])
        tile[ty + j][tx] = src[soff + x * pd.sstride[X] + y * pd.sstride[ydim]];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      int64_t y = y0 + ty + j, x = x0 + tx;
      if (x < pd.shape[X] && y < pd.shape[ydim]) dst[doff + y * ystride_d + x] = tile[tx][ty + j];
    }
  }
}

// fallback for shapes without a unit-stride source dim: plain gather, coalesced stores
template <typename T>
__global__ void __launch_bounds__(256)
k_permute_gather(const T* __restrict__ src, T* __restrict__ dst, const __grid_constant__ PermDesc pd) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pd.total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t rem = i, soff = 0;
#pragma unroll 1
    for (int d = pd.rank - 1; d >= 0; --d) {
      int64_t q = rem / pd.shape[d];
      soff += (rem - q * pd.shape[d]) * pd.sstride[d];
      rem = q;
    }
    dst[i] = src[soff];
  }
}

// a viewed as [outer, d, inner] (row-major): out[o, i, j] = a[o, i, j] * diag[i]
template <typename T>
__global__ void __launch_bounds__(256)
k_diag_scale(const T* __restrict__ a, const T* __restrict__ diag, T* __restrict__ out,
             int64_t n, int64_t d, int64_t inner) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    out[i] = a[i] * diag[(i / inner) % d];
  }
}

// permute_tma.cu: TMA load + TMA store ring; T10_ERR_UNSUPPORTED when the shape is not eligible
int permute_dense_tma(int dtype, const void* src, void* dst, int rank, const int64_t* shape,
                      const int64_t* sstride, cudaStream_t st);

template <typename T>
static int launch_permute(const T* src, T* dst, const PermDesc& pd, cudaStream_t st) {
  const int last = pd.rank - 1;
  const int grid_cap = sm_count() * 8;
  if (pd.sstride[last] == 1) {
    int rc = permute_dense_tma(sizeof(T) == 8 ? T10_F64 : T10_F32, src, dst, pd.rank, pd.shape, pd.sstride, st);
    if (rc != T10_ERR_UNSUPPORTED) return rc;
    // row copy; widest vector that divides the row, all source strides and both pointers
    int vec = (int)(16 / sizeof(T));
    auto ok = [&](int v) {
      if (pd.shape[last] % v) return false;
      for (int d = 0; d < last; ++d)
        if (pd.sstride[d] % v) return false;
      return ((uintptr_t)src % (v * sizeof(T)) == 0) && ((uintptr_t)dst % (v * sizeof(T)) == 0);
    };
    while (vec > 1 && !ok(vec)) vec /= 2;
    int64_t nvec = pd.total / vec;
    unsigned grid = (unsigned)std::min<int64_t>(ceil_div(nvec, 256), grid_cap);
    if (vec == 4) k_permute_rows<T, 4><<<grid, 256, 0, st>>>(src, dst, pd);
    else if (vec == 2) k_permute_rows<T, 2><<<grid, 256, 0, st>>>(src, dst, pd);
    else k_permute_rows<T, 1><<<grid, 256, 0, st>>>(src, dst, pd);
    T10_LAUNCH_CHECK();
    return T10_OK;
  }
  int ydim = -1;
  for (int d = 0; d < last; ++d)
    if (pd.sstride[d] == 1 && pd.shape[d] > 1) ydim = d;
  if (ydim >= 0) {
    int64_t tiles_x = ceil_div(pd.shape[last], 32), tiles_y = ceil_div(pd.shape[ydim], 32);
    int64_t outer = pd.total / (pd.shape[last] * pd.shape[ydim]);
    int64_t ntiles = tiles_x * tiles_y * outer;
    unsigned grid = (unsigned)std::min<int64_t>(ntiles, grid_cap);
    k_permute_tiled<T><<<grid, 256, 0, st>>>(src, dst, pd, ydim, tiles_x, tiles_y, ntiles);
    T10_LAUNCH_CHECK();
    return T10_OK;
  }
  unsigned grid = (unsigned)std::min<int64_t>(ceil_div(pd.total, 256), grid_cap);
  k_permute_gather<T><<<grid, 256, 0, st>>>(src, dst, pd);
  T10_LAUNCH_CHECK();
  return T10_OK;
}

}  // namespace t10

using namespace t10;

extern "C" int t10_relayout_blocks(int dtype, const void* d_src, void* d_dst, int64_t ntiles,
                                   const t10_relayout_tile* d_tiles, const int32_t* d_RR,
                                   const int32_t* d_RC, const int32_t* d_CR, const int32_t* d_CC,
                                   const int64_t* d_rowbase, const int32_t* d_coloff,
                                   const int64_t* d_src_ket_map, const int64_t* d_src_bra_map,
                                   int32_t* d_emit_index, t10_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (ntiles == 0) return T10_OK;
  T10_REQUIRE(ntiles > 0, T10_ERR_INVALID, "ntiles < 0");
  const bool chk = d_src_ket_map != nullptr && d_src_bra_map != nullptr;
  if (d_emit_index != nullptr) {  // plan-time pass: element index instead of data
    unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)sm_count() * 4);
    if (chk)
      k_relayout_blocks<double, true, true><<<grid, RL_THREADS, 0, st>>>(
          nullptr, nullptr, d_emit_index, ntiles, d_tiles, d_RR, d_RC, d_CR, d_CC, d_rowbase, d_coloff,
          d_src_ket_map, d_src_bra_map);
    else
      k_relayout_blocks<double, false, true><<<grid, RL_THREADS, 0, st>>>(
          nullptr, nullptr, d_emit_index, ntiles, d_tiles, d_RR, d_RC, d_CR, d_CC, d_rowbase, d_coloff,
          d_src_ket_map, d_src_bra_map);
    T10_LAUNCH_CHECK();
    return T10_OK;
  }
  static int occ_by_dev[T10_MAX_DEVICES][4] = {{0}};
  int* occ = occ_by_dev[cur_device()];
#define T10_RL_LAUNCH(T, CHK, SLOT)                                                                   \
  do {                                                                                                \
    auto kern = k_relayout_blocks<T, CHK, false>;                                                     \
    if (occ[SLOT] == 0) {                                                                             \
      T10_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[SLOT], kern, RL_THREADS, 0));       \
      if (occ[SLOT] < 1) occ[SLOT] = 1;                                                               \
    }                                                                                                 \
    unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)sm_count() * occ[SLOT]);             \
    kern<<<grid, RL_THREADS, 0, st>>>((const T*)d_src, (T*)d_dst, nullptr, ntiles, d_tiles, d_RR,     \
                                      d_RC, d_CR, d_CC, d_rowbase, d_coloff, d_src_ket_map,           \
                                      d_src_bra_map);                                                 \
  } while (0)
  if (dtype == T10_F64) {
    if (chk) T10_RL_LAUNCH(double, true, 0); else T10_RL_LAUNCH(double, false, 1);
  } else if (dtype == T10_F32) {
    if (chk) T10_RL_LAUNCH(float, true, 2); else T10_RL_LAUNCH(float, false, 3);
  } else {
    T10_REQUIRE(false, T10_ERR_UNSUPPORTED, "relayout: dtype %d not supported", dtype);
  }
#undef T10_RL_LAUNCH
  T10_LAUNCH_CHECK();
  return T10_OK;
}

extern "C" int t10_relayout_index(int dtype, const void* d_src, void* d_dst, int64_t n,
                                  const int32_t* d_index, t10_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) return T10_OK;
  T10_REQUIRE(n > 0 && (n & 1) == 0, T10_ERR_INVALID, "index length must be even");
  T10_REQUIRE(((uintptr_t)d_dst & 15) == 0 && ((uintptr_t)d_index & 7) == 0, T10_ERR_INVALID,
              "relayout_index: destination must be 16-byte aligned");
  const int64_t npairs = n / 2;
  static int occ_by_dev1[T10_MAX_DEVICES] = {0};
  int& occ = occ_by_dev1[cur_device()];
  if (occ == 0) {
    T10_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_relayout_index<double>, RI_THREADS, 0));
    if (occ < 1) occ = 1;
  }
  unsigned grid = (unsigned)std::min<int64_t>(ceil_div(npairs, (int64_t)RI_THREADS * RI_UNROLL),
                                              (int64_t)sm_count() * occ);
  if (dtype == T10_F64)
    k_relayout_index<double><<<grid, RI_THREADS, 0, st>>>((const double*)d_src, (double*)d_dst,
                                                          (const int2*)d_index, npairs);
  else if (dtype == T10_F32)
    k_relayout_index<float><<<grid, RI_THREADS, 0, st>>>((const float*)d_src, (float*)d_dst,
                                                         (const int2*)d_index, npairs);
  else
    T10_REQUIRE(false, T10_ERR_UNSUPPORTED, "relayout_index: dtype %d not supported", dtype);
  T10_LAUNCH_CHECK();
  return T10_OK;
}

// out[i] = sum over slices s (in order) of ws[s * n + i]: the deterministic second half of a split-K product
template <typename T>
__global__ void __launch_bounds__(256)
k_reduce_slices(const T* __restrict__ ws, T* __restrict__ out, int64_t n, int nslices) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    T acc = ws[i];
    for (int s2 = 1; s2 < nslices; ++s2) acc += ws[(int64_t)s2 * n + i];
    out[i] = acc;
  }
}

extern "C" int t10_reduce_slices(int dtype, const void* d_ws, void* d_out, int64_t n, int nslices,
                                 t10_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  T10_REQUIRE(n >= 0 && nslices >= 1, T10_ERR_INVALID, "reduce_slices: bad sizes");
  if (n == 0) return T10_OK;
  unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count() * 8);
  if (dtype == T10_F64) k_reduce_slices<double><<<grid, 256, 0, st>>>((const double*)d_ws, (double*)d_out, n, nslices);
  else if (dtype == T10_F32) k_reduce_slices<float><<<grid, 256, 0, st>>>((const float*)d_ws, (float*)d_out, n, nslices);
  else T10_REQUIRE(false, T10_ERR_UNSUPPORTED, "reduce_slices: dtype %d not supported", dtype);
  T10_LAUNCH_CHECK();
  return T10_OK;
}

extern "C" int t10_relayout_runs(int dtype, const void* d_src, void* d_dst, int64_t n, const int32_t* d_runs,
                                 int64_t nruns, const int32_t* d_chunk_first, t10_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) return T10_OK;
  T10_REQUIRE(n > 0 && n < (1ll << 31) && nruns >= 1, T10_ERR_INVALID, "relayout_runs: bad sizes");
  T10_REQUIRE(((uintptr_t)d_dst & 15) == 0, T10_ERR_INVALID, "relayout_runs: d_dst must be 16-byte aligned");
  const int64_t nchunks = ceil_div(n, (int64_t)RR_CHUNK);
  unsigned grid = (unsigned)std::min<int64_t>(nchunks, (int64_t)sm_count() * 64);
  if (dtype == T10_F64)
    k_relayout_runs<double><<<grid, RR_THREADS, 0, st>>>((const double*)d_src, (double*)d_dst, n, (const int2*)d_runs,
                                                         d_chunk_first, nchunks);
  else if (dtype == T10_F32)
    k_relayout_runs<float><<<grid, RR_THREADS, 0, st>>>((const float*)d_src, (float*)d_dst, n, (const int2*)d_runs,
                                                        d_chunk_first, nchunks);
  else
    T10_REQUIRE(false, T10_ERR_UNSUPPORTED, "relayout_runs: dtype %d not supported", dtype);
  T10_LAUNCH_CHECK();
  return T10_OK;
}

extern "C" int t10_relayout_segments(int dtype, int nsrc, const void* const* h_src, void* d_dst, int64_t nseg,
                                     const int32_t* d_seg, t10_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (nseg == 0) return T10_OK;
  T10_REQUIRE(nseg > 0 && nseg < (1ll << 31), T10_ERR_INVALID, "relayout_segments: bad segment count");
  T10_REQUIRE(nsrc >= 1 && nsrc <= RS_MAXSRC && h_src, T10_ERR_INVALID, "relayout_segments: 1..%d source arenas", RS_MAXSRC);
  T10_REQUIRE(((uintptr_t)d_seg & 15) == 0, T10_ERR_INVALID, "relayout_segments: table must be 16-byte aligned");
  SegSources ss;
  memset(&ss, 0, sizeof(ss));
  for (int i = 0; i < nsrc; ++i) ss.base[i] = h_src[i];
  const int64_t warps = nseg;
  static int rs_ctas = 0;        // CTAs of 512 threads per SM (TOR10_RELAYOUT_SEG_CTAS, default 3: 48 warps per SM)
  if (rs_ctas == 0) {
    const char* e = getenv("TOR10_RELAYOUT_SEG_CTAS");
    rs_ctas = e ? std::max(1, std::min(8, atoi(e))) : 3;
  }
  unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(warps, RS_THREADS / 32), (int64_t)sm_count() * rs_ctas));
  if (dtype == T10_F64)
    k_relayout_segments<double><<<grid, RS_THREADS, 0, st>>>(ss, (double*)d_dst, (const int4*)d_seg, (int)nseg);
  else if (dtype == T10_F32)
    k_relayout_segments<float><<<grid, RS_THREADS, 0, st>>>(ss, (float*)d_dst, (const int4*)d_seg, (int)nseg);
  else
    T10_REQUIRE(false, T10_ERR_UNSUPPORTED, "relayout_segments: dtype %d not supported", dtype);
  T10_LAUNCH_CHECK();
  return T10_OK;
}

// ---------------------------------------------------------------- peer gather -------
// Sector gather over NVLink peer memory (multi-GPU Contract): every rank's finished sectors sit in its
// own symmetric staging buffer; this kernel PULLS all the slices this rank does not own (P2P loads,
// 16-byte vectors, all peers in flight at once) into the local output arena.  One launch, no NCCL.
constexpr int PEER_MAXSEG = 64;
struct PeerSegs {
  int n;
  const void* src[PEER_MAXSEG];   // peer (or local) staging base pointers
  int64_t lo[PEER_MAXSEG];        // element range [lo, hi) of the slice, same offsets in staging and arena
  int64_t hi[PEER_MAXSEG];
  int64_t vec_start[PEER_MAXSEG + 1]; // prefix sum of slice lengths in 16-byte vectors
};

template <typename T>
__global__ void __launch_bounds__(256)
k_gather_peers(T* __restrict__ dst, const __grid_constant__ PeerSegs ps) {
  constexpr int VE = 16 / sizeof(T);
  const int64_t total = ps.vec_start[ps.n];
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
    int sgm = 0, hi_ = ps.n - 1;          // (bisection: up to 64 slices)
#pragma unroll 1
    while (sgm < hi_) {
      const int mid = (sgm + hi_) >> 1;
      if (v >= ps.vec_start[mid + 1]) sgm = mid + 1; else hi_ = mid;
    }
    const int64_t e = ps.lo[sgm] + (v - ps.vec_start[sgm]) * VE;
    const int4 val = *reinterpret_cast<const int4*>(reinterpret_cast<const T*>(ps.src[sgm]) + e);
    *reinterpret_cast<int4*>(dst + e) = val;
  }
}

// Consumer side of the fused GEMM -> all-gather: wait until every rank's flag reached `step` (the ranks'
// GEMM kernels raise them after their last tile store, t10_ggemm_allgather), then copy the completed staging
// arena into the result arena.  The spin is bounded (about two seconds): on expiry *err is set and the copy
// proceeds, so a dead peer cannot hang the device.
template <typename T>
__global__ void __launch_bounds__(256)
k_wait_copy(T* __restrict__ dst, const T* __restrict__ src, int64_t nvec, const unsigned long long* flags, int nflags,
            unsigned long long step, int* __restrict__ err) {
  if (threadIdx.x == 0) {
    for (int r = 0; r < nflags; ++r) {
      unsigned long long v = 0;
      for (long long spin = 0; spin < 20000000ll; ++spin) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + r) : "memory");
        if (v >= step) break;
        __nanosleep(100);
      }
      if (v < step) atomicExch(err, 1);
    }
  }
  __syncthreads();
  const int4* s4 = reinterpret_cast<const int4*>(src);
  int4* d4 = reinterpret_cast<int4*>(dst);
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x)
    d4[v] = s4[v];
}

extern "C" int t10_wait_copy(int dtype, void* d_dst, const void* d_src, int64_t n, const void* d_flags, int nflags,
                             uint64_t step, int32_t* d_err, t10_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  T10_REQUIRE(dtype == T10_F64 || dtype == T10_F32, T10_ERR_UNSUPPORTED, "wait_copy: dtype %d", dtype);
  const int ve = dtype == T10_F64 ? 2 : 4;
  T10_REQUIRE(n >= 0 && n % ve == 0 && nflags >= 0 && nflags <= 64, T10_ERR_INVALID, "wait_copy: bad sizes");
  T10_REQUIRE(((uintptr_t)d_dst & 15) == 0 && ((uintptr_t)d_src & 15) == 0, T10_ERR_INVALID, "wait_copy: alignment");
  const int64_t nvec = n / ve;
  unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(nvec, 256 * 4), (int64_t)sm_count() * 4));
  if (dtype == T10_F64)
    k_wait_copy<double><<<grid, 256, 0, st>>>((double*)d_dst, (const double*)d_src, nvec,
                                               (const unsigned long long*)d_flags, nflags, step, d_err);
  else
    k_wait_copy<float><<<grid, 256, 0, st>>>((float*)d_dst, (const float*)d_src, nvec,
                                              (const unsigned long long*)d_flags, nflags, step, d_err);
  T10_LAUNCH_CHECK();
  return T10_OK;
}

extern "C" int t10_gather_peers(int dtype, void* d_dst, int nseg, const void* const* h_src,
                                const int64_t* h_lo, const int64_t* h_hi, t10_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  T10_REQUIRE(nseg >= 0 && nseg <= PEER_MAXSEG, T10_ERR_INVALID, "gather_peers: at most %d segments", PEER_MAXSEG);
  T10_REQUIRE(dtype == T10_F64 || dtype == T10_F32, T10_ERR_UNSUPPORTED, "gather_peers: dtype %d", dtype);
  const int ve = dtype == T10_F64 ? 2 : 4;
  PeerSegs ps;
  memset(&ps, 0, sizeof(ps));
  ps.n = nseg;
  for (int i = 0; i < nseg; ++i) {
    T10_REQUIRE(h_lo[i] % ve == 0 && h_hi[i] % ve == 0 && h_hi[i] >= h_lo[i], T10_ERR_INVALID,
                "gather_peers: slice bounds must be multiples of 16 bytes");
    T10_REQUIRE(((uintptr_t)h_src[i] & 15) == 0 && ((uintptr_t)d_dst & 15) == 0, T10_ERR_INVALID,
                "gather_peers: buffers must be 16-byte aligned");
    ps.src[i] = h_src[i];
    ps.lo[i] = h_lo[i];
    ps.hi[i] = h_hi[i];
    ps.vec_start[i + 1] = ps.vec_start[i] + (h_hi[i] - h_lo[i]) / ve;
  }
  const int64_t total = ps.vec_start[nseg];
  if (total == 0) return T10_OK;
  unsigned grid = (unsigned)std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 8);
  if (dtype == T10_F64) k_gather_peers<double><<<grid, 256, 0, st>>>((double*)d_dst, ps);
  else k_gather_peers<float><<<grid, 256, 0, st>>>((float*)d_dst, ps);
  T10_LAUNCH_CHECK();
  return T10_OK;
}

extern "C" int t10_permute_dense(int dtype, const void* d_src, void* d_dst, int rank,
                                 const int64_t* h_shape, const int64_t* h_src_strides,
                                 t10_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  T10_REQUIRE(rank >= 0 && rank <= T10_MAX_RANK, T10_ERR_INVALID, "rank %d outside [0,%d]", rank,
              T10_MAX_RANK);
  T10_REQUIRE(dtype == T10_F64 || dtype == T10_F32, T10_ERR_UNSUPPORTED, "permute: dtype %d", dtype);
  // drop size-1 dims, merge dims that stay adjacent in the source
  PermDesc pd;
  pd.rank = 0;
  pd.total = 1;
  for (int d = 0; d < rank; ++d) {
    T10_REQUIRE(h_shape[d] >= 0, T10_ERR_INVALID, "negative extent");
    pd.total *= h_shape[d];
    if (h_shape[d] == 1) continue;
    if (pd.rank > 0 && pd.sstride[pd.rank - 1] == h_shape[d] * h_src_strides[d]) {
      pd.shape[pd.rank - 1] *= h_shape[d];
      pd.sstride[pd.rank - 1] = h_src_strides[d];
    } else {
      pd.shape[pd.rank] = h_shape[d];
      pd.sstride[pd.rank] = h_src_strides[d];
      ++pd.rank;
    }
  }
  if (pd.total == 0) return T10_OK;
  if (pd.rank == 0) {  // scalar / all-ones shape
    pd.rank = 1;
    pd.shape[0] = 1;
    pd.sstride[0] = 1;
  }
  if (dtype == T10_F64) return launch_permute<double>((const double*)d_src, (double*)d_dst, pd, st);
  return launch_permute<float>((const float*)d_src, (float*)d_dst, pd, st);
}

extern "C" int t10_diag_scale(int dtype, const void* d_a, const void* d_diag, void* d_out,
                              int64_t outer, int64_t d, int64_t inner, t10_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  T10_REQUIRE(outer >= 0 && d >= 0 && inner >= 0, T10_ERR_INVALID, "diag_scale: negative extent");
  int64_t n = outer * d * inner;
  if (n == 0) return T10_OK;
  unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count() * 8);
  if (dtype == T10_F64)
    k_diag_scale<double><<<grid, 256, 0, st>>>((const double*)d_a, (const double*)d_diag,
                                                (double*)d_out, n, d, inner);
  else if (dtype == T10_F32)
    k_diag_scale<float><<<grid, 256, 0, st>>>((const float*)d_a, (const float*)d_diag,
                                               (float*)d_out, n, d, inner);
  else
    T10_REQUIRE(false, T10_ERR_UNSUPPORTED, "diag_scale: dtype %d", dtype);
  T10_LAUNCH_CHECK();
  return T10_OK;
}
