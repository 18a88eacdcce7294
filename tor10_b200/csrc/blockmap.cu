// Family 1: quantum-number fusion, sector matching and block offsets (int64, bit-exact).
//
// Restates on the device what the reference does with numpy on the host:
//   Bond.combine                       Bond.py:367-412   (row-major outer fusion)
//   Bond.__mul__                       Bond.py:469-477   (sign flip, no modular reduction)
//   Bond.GetUniqueQnums / _fx_GetCommRows  Bond.py:423-447, :13-28 (ascending-lex sector list)
//   UniTensor.__init__ symmetric       UniTensor.py:393-441 (maps, inverse maps, block shapes)
//
// Algorithm (all on the device, two small host round trips for sizes):
//   1. fuse:   tq[x, s] for every flat row / column index x            (one thread per x)
//   2. ranges: per-symmetry min/max of the totals -> order-preserving packed key
//   3. presence bitmaps over the key space, AND of row/col presence, exclusive scan
//      -> ascending-lex sector ids and the sector charge table
//   4. stable counting rank of rows (cols) inside their sector: per-warp segment
//      histograms, column scan over segments, then a second pass that writes
//      (sector, offset), the grouped index lists and the arena address tables.
#include <stdarg.h>
#include <string.h>

#include <climits>

#include "t10_common.cuh"

namespace t10 {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int n_by_dev[T10_MAX_DEVICES] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int& n = n_by_dev[(dev >= 0 && dev < T10_MAX_DEVICES) ? dev : 0];
  if (n == 0 && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  return n;
}

struct SpaceDesc {
  int nb;
  int nsym;
  int64_t total;
  int64_t dims[T10_MAX_RANK];
  int64_t qoff[T10_MAX_RANK];
  int32_t sign[T10_MAX_RANK];
  int32_t kind[T10_MAX_NSYM];
  int64_t symn[T10_MAX_NSYM];
};

struct KeyDesc {
  int nsym;
  int64_t mn[T10_MAX_NSYM];
  int64_t mult[T10_MAX_NSYM];
  int64_t range[T10_MAX_NSYM];
};

__device__ __forceinline__ int64_t np_mod(int64_t a, int64_t n) {
  int64_t r = a % n;  // numpy remainder takes the sign of the divisor (n > 0)
  return r < 0 ? r + n : r;
}

// total charge of flat index x of a bond space, all symmetries      (Bond.py:367-378)
__device__ __forceinline__ void fuse_row(const SpaceDesc& sp, const int64_t* __restrict__ q,
                                         int64_t x, int64_t* out) {
  int64_t digit[T10_MAX_RANK];
#pragma unroll 1
  for (int b = sp.nb - 1; b >= 0; --b) {
    digit[b] = x % sp.dims[b];
    x /= sp.dims[b];
  }
  for (int s = 0; s < sp.nsym; ++s) {
    int64_t v = (int64_t)sp.sign[0] * q[(sp.qoff[0] + digit[0]) * sp.nsym + s];
#pragma unroll 1
    for (int b = 1; b < sp.nb; ++b) {
      int64_t w = (int64_t)sp.sign[b] * q[(sp.qoff[b] + digit[b]) * sp.nsym + s];
      v = (sp.kind[s] == T10_SYM_U1) ? (v + w) : np_mod(v + w, sp.symn[s]);
    }
    out[s] = v;
  }
}

__global__ void k_fuse(const __grid_constant__ SpaceDesc sp, const int64_t* __restrict__ q,
                       int64_t* __restrict__ out) {
  int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= sp.total) return;
  int64_t v[T10_MAX_NSYM];
  fuse_row(sp, q, x, v);
  for (int s = 0; s < sp.nsym; ++s) out[x * sp.nsym + s] = v[s];
}

__global__ void k_init_minmax(int64_t* mn, int64_t* mx, int nsym) {
  int s = threadIdx.x;
  if (s < nsym) {
    mn[s] = LLONG_MAX;
    mx[s] = LLONG_MIN;
  }
}

__global__ void k_minmax(const int64_t* __restrict__ tq, int64_t n, int nsym, int64_t* mn,
                         int64_t* mx) {
  int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int s = 0; s < nsym; ++s) {
    long long lo = LLONG_MAX, hi = LLONG_MIN;
    if (x < n) lo = hi = tq[x * nsym + s];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      long long lo2 = __shfl_xor_sync(0xffffffffu, lo, o);
      long long hi2 = __shfl_xor_sync(0xffffffffu, hi, o);
      lo = lo2 < lo ? lo2 : lo;
      hi = hi2 > hi ? hi2 : hi;
    }
    if ((threadIdx.x & 31) == 0 && lo != LLONG_MAX) {
      atomicMin((long long*)&mn[s], lo);
      atomicMax((long long*)&mx[s], hi);
    }
  }
}

__global__ void k_keys(const int64_t* __restrict__ tq, int64_t n, const __grid_constant__ KeyDesc kd,
                       int32_t* __restrict__ key, uint8_t* __restrict__ present) {
  int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  int64_t k = 0;
  for (int s = 0; s < kd.nsym; ++s) k += (tq[x * kd.nsym + s] - kd.mn[s]) * kd.mult[s];
  key[x] = (int32_t)k;
  present[k] = 1;
}

// ---- exclusive scan of (present_in & present_out) over the key space --------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[SCAN_THREADS / 32];
  __shared__ int block_total;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0) {
    int ws = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
    int winc = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = winc - ws;
    if (lane == 31) block_total = winc;
  }
  __syncthreads();
  int res = inc - v + warp_sums[w];
  *total = block_total;
  __syncthreads();
  return res;
}

__global__ void k_sector_count(const uint8_t* __restrict__ pin, const uint8_t* __restrict__ pout,
                               int64_t P, int32_t* __restrict__ blkcnt) {
  int64_t base = ((int64_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_ITEMS;
  int c = 0;
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t k = base + i;
    if (k < P) c += (pin[k] & pout[k]);
  }
  int total;
  block_exclusive_scan(c, &total);
  if (threadIdx.x == 0) blkcnt[blockIdx.x] = total;
}

// single block: in-place exclusive scan of blkcnt[0..nblk), total -> counts[0]
__global__ void k_scan_blocks(int32_t* blkcnt, int64_t nblk, int64_t* counts) {
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < nblk; base += SCAN_THREADS) {
    int64_t i = base + threadIdx.x;
    int v = i < nblk ? blkcnt[i] : 0;
    int total;
    int ex = block_exclusive_scan(v, &total);
    int carry = carry_s;
    if (i < nblk) blkcnt[i] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[0] = carry_s;
}

__global__ void k_sector_write(const uint8_t* __restrict__ pin, const uint8_t* __restrict__ pout,
                               int64_t P, const int32_t* __restrict__ blkoff,
                               const __grid_constant__ KeyDesc kd, int32_t* __restrict__ sec_of_key,
                               int64_t* __restrict__ block_qnums, int64_t cap_sec) {
  int64_t base = ((int64_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_ITEMS;
  int c = 0;
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t k = base + i;
    if (k < P) c += (pin[k] & pout[k]);
  }
  int total;
  int id = block_exclusive_scan(c, &total) + blkoff[blockIdx.x];
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t k = base + i;
    if (k >= P) break;
    if (pin[k] & pout[k]) {
      sec_of_key[k] = id;
      if (id < cap_sec) {
        int64_t rem = k;
        for (int s = 0; s < kd.nsym; ++s) {  // mult is descending: s = 0 most significant
          int64_t d = rem / kd.mult[s];
          rem -= d * kd.mult[s];
          block_qnums[(int64_t)id * kd.nsym + s] = d + kd.mn[s];
        }
      }
      ++id;
    } else {
      sec_of_key[k] = -1;
    }
  }
}

// ---- stable rank inside sectors -----------------------------------------------------
// one warp per segment of S consecutive flat indices (S multiple of 32)
__global__ void k_rank_hist(const int32_t* __restrict__ key, const int32_t* __restrict__ sec_of_key,
                            int64_t n, int64_t S, int64_t nseg, int64_t nsec,
                            int32_t* __restrict__ hist) {
  const int lane = threadIdx.x & 31;
  int64_t seg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (seg >= nseg) return;  // whole warp exits together
  int64_t beg = seg * S, end = beg + S < n ? beg + S : n;
  int32_t* h = hist + seg * nsec;
  for (int64_t x0 = beg; x0 < end; x0 += 32) {
    int64_t x = x0 + lane;
    int sec = x < end ? sec_of_key[key[x]] : -1;
    unsigned m = __match_any_sync(0xffffffffu, sec);
    if (sec >= 0 && lane == (__ffs(m) - 1)) h[sec] += __popc(m);
    __syncwarp();
  }
}

// block per sector: exclusive scan down the segments (in place), totals out
__global__ void k_rank_scan(int32_t* __restrict__ hist, int64_t nseg, int64_t nsec,
                            int64_t* __restrict__ totals) {
  __shared__ int carry_s;
  const int64_t s = blockIdx.x;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < nseg; base += SCAN_THREADS) {
    const int64_t g = base + threadIdx.x;
    const int v = g < nseg ? hist[g * nsec + s] : 0;
    int total;
    const int ex = block_exclusive_scan(v, &total);
    const int carry = carry_s;
    if (g < nseg) hist[g * nsec + s] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[s] = carry_s;
}

// sec_info [nsec,8] = {rows, cols, ket_start, bra_start, arena_off, ld, 0, 0}
// ld = cols rounded up to even: every stored row starts 16-byte aligned (fp64), so a sector can feed the
// grouped GEMM's 16-byte cp.async loaders in place, whatever its width
__global__ void k_sector_offsets(const int64_t* __restrict__ rows, const int64_t* __restrict__ cols,
                                 int64_t nsec, int64_t* __restrict__ sec_info,
                                 int64_t* __restrict__ counts) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int64_t ks = 0, bs = 0, off = 0;
  for (int64_t s = 0; s < nsec; ++s) {
    int64_t* o = sec_info + s * 8;
    o[0] = rows[s];
    o[1] = cols[s];
    o[2] = ks;
    o[3] = bs;
    const int64_t ld = cols[s] + (cols[s] & 1);
    o[4] = off;
    o[5] = ld;
    o[6] = o[7] = 0;
    ks += rows[s];
    bs += cols[s];
    int64_t sz = rows[s] * ld;
    off += (sz + T10_ARENA_ALIGN - 1) / T10_ARENA_ALIGN * T10_ARENA_ALIGN;
  }
  counts[1] = off;
  counts[2] = ks;
  counts[3] = bs;
}

// is_row = 1: writes ket_map/ket_idx/rowbase; 0: bra_map/bra_idx/coloff
__global__ void k_rank_write(const int32_t* __restrict__ key, const int32_t* __restrict__ sec_of_key,
                             int64_t n, int64_t S, int64_t nseg, int64_t nsec,
                             int32_t* __restrict__ hist, const int64_t* __restrict__ sec_info,
                             int is_row, int64_t* __restrict__ map, int64_t* __restrict__ idx,
                             int64_t* __restrict__ rowbase, int32_t* __restrict__ coloff) {
  const int lane = threadIdx.x & 31;
  int64_t seg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (seg >= nseg) return;
  int64_t beg = seg * S, end = beg + S < n ? beg + S : n;
  int32_t* h = hist + seg * nsec;
  for (int64_t x0 = beg; x0 < end; x0 += 32) {
    int64_t x = x0 + lane;
    int sec = x < end ? sec_of_key[key[x]] : -1;
    unsigned m = __match_any_sync(0xffffffffu, sec);
    int leader = __ffs(m) - 1;
    int rank = __popc(m & ((1u << lane) - 1u));
    int cnt = 0;
    if (sec >= 0 && lane == leader) {
      cnt = h[sec];
      h[sec] = cnt + __popc(m);
    }
    cnt = __shfl_sync(0xffffffffu, cnt, leader);
    if (x < end) {
      if (sec >= 0) {
        int64_t off = (int64_t)cnt + rank;
        const int64_t* si = sec_info + (int64_t)sec * 8;
        map[2 * x] = sec;
        map[2 * x + 1] = off;
        if (is_row) {
          idx[si[2] + off] = x;
          rowbase[x] = si[4] + off * si[5];
        } else {
          idx[si[3] + off] = x;
          coloff[x] = (int32_t)off;
        }
      } else {
        map[2 * x] = -1;
        map[2 * x + 1] = -1;
        if (is_row) rowbase[x] = -1;
        else coloff[x] = -1;
      }
    }
    __syncwarp();
  }
}

__global__ void k_unravel(int64_t n, int ns, const __grid_constant__ SpaceDesc sd /*dims = strides*/,
                          const int64_t* __restrict__ idx, int64_t* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int64_t x = idx[t];
  for (int p = 0; p < ns; ++p) {
    out[t * ns + p] = x / sd.dims[p];
    x = x % sd.dims[p];
  }
}

struct TableDesc {
  int nb;
  int64_t dims[T10_MAX_RANK];
  int64_t srow[T10_MAX_RANK];
  int64_t scol[T10_MAX_RANK];
};

__global__ void k_relayout_tables(int64_t n, const __grid_constant__ TableDesc td,
                                  const int64_t* __restrict__ idx, int32_t* __restrict__ to_row,
                                  int32_t* __restrict__ to_col) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int64_t x = idx[t];
  int64_t r = 0, c = 0;
#pragma unroll 1
  for (int b = td.nb - 1; b >= 0; --b) {
    int64_t d = x % td.dims[b];
    x /= td.dims[b];
    r += d * td.srow[b];
    c += d * td.scol[b];
  }
  to_row[t] = (int32_t)r;
  to_col[t] = (int32_t)c;
}

static int fill_space(SpaceDesc* sp, int nb, int nsym, const int64_t* dims, const int32_t* signs,
                      const int32_t* kind, const int64_t* symn, const int64_t* qoff) {
  T10_REQUIRE(nb >= 1 && nb <= T10_MAX_RANK, T10_ERR_INVALID, "bond count %d outside [1,%d]", nb,
              T10_MAX_RANK);
  T10_REQUIRE(nsym >= 1 && nsym <= T10_MAX_NSYM, T10_ERR_INVALID, "nsym %d outside [1,%d]", nsym,
              T10_MAX_NSYM);
  memset(sp, 0, sizeof(*sp));
  sp->nb = nb;
  sp->nsym = nsym;
  sp->total = 1;
  for (int b = 0; b < nb; ++b) {
    T10_REQUIRE(dims[b] >= 1, T10_ERR_INVALID, "bond %d has dim %lld", b, (long long)dims[b]);
    T10_REQUIRE(signs[b] == 1 || signs[b] == -1, T10_ERR_INVALID, "sign must be +1/-1");
    sp->dims[b] = dims[b];
    sp->qoff[b] = qoff[b];
    sp->sign[b] = signs[b];
    T10_REQUIRE(sp->total <= (int64_t)INT32_MAX / dims[b], T10_ERR_UNSUPPORTED,
                "fused bond space exceeds 2^31 states");
    sp->total *= dims[b];
  }
  for (int s = 0; s < nsym; ++s) {
    T10_REQUIRE(kind[s] == T10_SYM_U1 || kind[s] == T10_SYM_ZN, T10_ERR_INVALID, "bad symmetry kind");
    T10_REQUIRE(kind[s] == T10_SYM_U1 || symn[s] >= 2, T10_ERR_INVALID, "Zn needs n>=2");
    sp->kind[s] = kind[s];
    sp->symn[s] = symn[s];
  }
  return T10_OK;
}

}  // namespace t10

using namespace t10;

extern "C" int t10_version(void) { return T10_ABI_VERSION; }
extern "C" const char* t10_last_error(void) { return g_err; }

extern "C" int t10_device_props(int device, int* smc, int* major, int* minor, int* smem_optin) {
  T10_CUDA(cudaDeviceGetAttribute(smc, cudaDevAttrMultiProcessorCount, device));
  T10_CUDA(cudaDeviceGetAttribute(major, cudaDevAttrComputeCapabilityMajor, device));
  T10_CUDA(cudaDeviceGetAttribute(minor, cudaDevAttrComputeCapabilityMinor, device));
  T10_CUDA(cudaDeviceGetAttribute(smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  return T10_OK;
}

extern "C" int t10_fuse_qnums(int nbonds, int nsym, const int64_t* h_dims, const int32_t* h_signs,
                              const int32_t* h_sym_kind, const int64_t* h_sym_n,
                              const int64_t* h_qoff, const int64_t* d_qnums, int64_t* d_out,
                              t10_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SpaceDesc sp;
  int rc = fill_space(&sp, nbonds, nsym, h_dims, h_signs, h_sym_kind, h_sym_n, h_qoff);
  if (rc) return rc;
  k_fuse<<<(unsigned)ceil_div(sp.total, 256), 256, 0, st>>>(sp, d_qnums, d_out);
  T10_LAUNCH_CHECK();
  return T10_OK;
}

extern "C" int t10_blockmap_build(int rank, int rowrank, int nsym, const int64_t* h_dims,
                                  const int32_t* h_braket, const int32_t* h_sym_kind,
                                  const int64_t* h_sym_n, const int64_t* h_qoff,
                                  const int64_t* d_qnums, int64_t cap_sec, int64_t* d_block_qnums,
                                  int64_t* d_ket_map, int64_t* d_bra_map, int64_t* d_ket_idx,
                                  int64_t* d_bra_idx, int64_t* d_sec_info, int64_t* d_rowbase,
                                  int32_t* d_coloff, int64_t* h_counts, t10_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  T10_REQUIRE(rank >= 2 && rank <= T10_MAX_RANK, T10_ERR_INVALID, "rank %d outside [2,%d]", rank,
              T10_MAX_RANK);
  T10_REQUIRE(rowrank >= 1 && rowrank < rank, T10_ERR_INVALID,
              "symmetric tensor needs a row and a column space (rowrank=%d, rank=%d)", rowrank, rank);
  // UniTensor.py:2861-2868: row space charges * braket * KET(-1), column space * braket * BRA(+1)
  int32_t signs[T10_MAX_RANK];
  for (int b = 0; b < rank; ++b) {
    T10_REQUIRE(h_braket[b] == 1 || h_braket[b] == -1, T10_ERR_INVALID, "braket must be +1/-1");
    signs[b] = b < rowrank ? -h_braket[b] : h_braket[b];
  }
  SpaceDesc spr, spc;
  int rc = fill_space(&spr, rowrank, nsym, h_dims, signs, h_sym_kind, h_sym_n, h_qoff);
  if (rc) return rc;
  rc = fill_space(&spc, rank - rowrank, nsym, h_dims + rowrank, signs + rowrank, h_sym_kind, h_sym_n,
                  h_qoff + rowrank);
  if (rc) return rc;
  const int64_t R = spr.total, C = spc.total;
  T10_REQUIRE(cap_sec >= (R < C ? R : C), T10_ERR_INVALID, "cap_sec too small");

  Scratch scr(st);
  int64_t *tq_in, *tq_out, *mnmx, *d_counts, *tot_in, *tot_out;
  int32_t *key_in, *key_out;
  T10_CUDA(scr.alloc(&tq_in, R * nsym));
  T10_CUDA(scr.alloc(&tq_out, C * nsym));
  T10_CUDA(scr.alloc(&mnmx, 2 * T10_MAX_NSYM));
  T10_CUDA(scr.alloc(&d_counts, 4));
  T10_CUDA(scr.alloc(&key_in, R));
  T10_CUDA(scr.alloc(&key_out, C));
  T10_CUDA(cudaMemsetAsync(d_counts, 0, 4 * sizeof(int64_t), st));

  // 1. fuse
  k_fuse<<<(unsigned)ceil_div(R, 256), 256, 0, st>>>(spr, d_qnums, tq_in);
  T10_LAUNCH_CHECK();
  k_fuse<<<(unsigned)ceil_div(C, 256), 256, 0, st>>>(spc, d_qnums, tq_out);
  T10_LAUNCH_CHECK();
  // 2. ranges
  k_init_minmax<<<1, 32, 0, st>>>(mnmx, mnmx + T10_MAX_NSYM, nsym);
  k_minmax<<<(unsigned)ceil_div(R, 256), 256, 0, st>>>(tq_in, R, nsym, mnmx, mnmx + T10_MAX_NSYM);
  k_minmax<<<(unsigned)ceil_div(C, 256), 256, 0, st>>>(tq_out, C, nsym, mnmx, mnmx + T10_MAX_NSYM);
  T10_LAUNCH_CHECK();
  int64_t h_mnmx[2 * T10_MAX_NSYM];
  T10_CUDA(cudaMemcpyAsync(h_mnmx, mnmx, sizeof(h_mnmx), cudaMemcpyDeviceToHost, st));
  T10_CUDA(cudaStreamSynchronize(st));
  KeyDesc kd;
  memset(&kd, 0, sizeof(kd));
  kd.nsym = nsym;
  int64_t P = 1;
  for (int s = nsym - 1; s >= 0; --s) {
    kd.mn[s] = h_mnmx[s];
    // range computed in unsigned arithmetic: charges near the int64 limits cannot be keyed
    uint64_t range = (uint64_t)h_mnmx[T10_MAX_NSYM + s] - (uint64_t)h_mnmx[s] + 1ull;
    T10_REQUIRE(range >= 1 && range <= (uint64_t)T10_KEYSPACE_MAX &&
                    (uint64_t)P * range <= (uint64_t)T10_KEYSPACE_MAX,
                T10_ERR_KEYSPACE, "total-charge key space too large (> %lld)",
                (long long)T10_KEYSPACE_MAX);
    kd.range[s] = (int64_t)range;
    kd.mult[s] = P;
    P *= (int64_t)range;
  }
  // 3. presence + sector ids
  uint8_t *pin, *pout;
  int32_t *sec_of_key, *blkcnt;
  const int64_t nblk = ceil_div(P, (int64_t)SCAN_THREADS * SCAN_ITEMS);
  T10_CUDA(scr.alloc(&pin, P));
  T10_CUDA(scr.alloc(&pout, P));
  T10_CUDA(scr.alloc(&sec_of_key, P));
  T10_CUDA(scr.alloc(&blkcnt, nblk));
  T10_CUDA(cudaMemsetAsync(pin, 0, P, st));
  T10_CUDA(cudaMemsetAsync(pout, 0, P, st));
  k_keys<<<(unsigned)ceil_div(R, 256), 256, 0, st>>>(tq_in, R, kd, key_in, pin);
  k_keys<<<(unsigned)ceil_div(C, 256), 256, 0, st>>>(tq_out, C, kd, key_out, pout);
  T10_LAUNCH_CHECK();
  k_sector_count<<<(unsigned)nblk, SCAN_THREADS, 0, st>>>(pin, pout, P, blkcnt);
  k_scan_blocks<<<1, SCAN_THREADS, 0, st>>>(blkcnt, nblk, d_counts);
  k_sector_write<<<(unsigned)nblk, SCAN_THREADS, 0, st>>>(pin, pout, P, blkcnt, kd, sec_of_key,
                                                          d_block_qnums, cap_sec);
  T10_LAUNCH_CHECK();
  int64_t h_c[4];
  T10_CUDA(cudaMemcpyAsync(h_c, d_counts, sizeof(h_c), cudaMemcpyDeviceToHost, st));
  T10_CUDA(cudaStreamSynchronize(st));
  const int64_t nsec = h_c[0];
  T10_REQUIRE(nsec > 0, T10_ERR_NOBLOCK,
              "no valid block: row and column total qnums share no sector");
  T10_REQUIRE(nsec <= cap_sec, T10_ERR_INVALID, "internal: nsec %lld > cap %lld", (long long)nsec,
              (long long)cap_sec);
  // 4. ranks
  auto seg_plan = [&](int64_t n, int64_t* S, int64_t* nseg) {
    int64_t ns = ceil_div(n, 32);
    if (ns > 1024) ns = 1024;
    while (ns > 1 && ns * nsec > (1ll << 27)) ns /= 2;
    int64_t s = ceil_div(ceil_div(n, ns), 32) * 32;
    *S = s;
    *nseg = ceil_div(n, s);
  };
  int64_t S_in, nseg_in, S_out, nseg_out;
  seg_plan(R, &S_in, &nseg_in);
  seg_plan(C, &S_out, &nseg_out);
  int32_t *hist_in, *hist_out;
  T10_CUDA(scr.alloc(&hist_in, nseg_in * nsec));
  T10_CUDA(scr.alloc(&hist_out, nseg_out * nsec));
  T10_CUDA(scr.alloc(&tot_in, nsec));
  T10_CUDA(scr.alloc(&tot_out, nsec));
  T10_CUDA(cudaMemsetAsync(hist_in, 0, nseg_in * nsec * sizeof(int32_t), st));
  T10_CUDA(cudaMemsetAsync(hist_out, 0, nseg_out * nsec * sizeof(int32_t), st));
  k_rank_hist<<<(unsigned)ceil_div(nseg_in * 32, 128), 128, 0, st>>>(key_in, sec_of_key, R, S_in,
                                                                     nseg_in, nsec, hist_in);
  k_rank_hist<<<(unsigned)ceil_div(nseg_out * 32, 128), 128, 0, st>>>(key_out, sec_of_key, C, S_out,
                                                                      nseg_out, nsec, hist_out);
  k_rank_scan<<<(unsigned)nsec, SCAN_THREADS, 0, st>>>(hist_in, nseg_in, nsec, tot_in);
  k_rank_scan<<<(unsigned)nsec, SCAN_THREADS, 0, st>>>(hist_out, nseg_out, nsec, tot_out);
  k_sector_offsets<<<1, 32, 0, st>>>(tot_in, tot_out, nsec, d_sec_info, d_counts);
  T10_LAUNCH_CHECK();
  k_rank_write<<<(unsigned)ceil_div(nseg_in * 32, 128), 128, 0, st>>>(
      key_in, sec_of_key, R, S_in, nseg_in, nsec, hist_in, d_sec_info, 1, d_ket_map, d_ket_idx,
      d_rowbase, nullptr);
  k_rank_write<<<(unsigned)ceil_div(nseg_out * 32, 128), 128, 0, st>>>(
      key_out, sec_of_key, C, S_out, nseg_out, nsec, hist_out, d_sec_info, 0, d_bra_map, d_bra_idx,
      nullptr, d_coloff);
  T10_LAUNCH_CHECK();
  T10_CUDA(cudaMemcpyAsync(h_c, d_counts, sizeof(h_c), cudaMemcpyDeviceToHost, st));
  T10_CUDA(cudaStreamSynchronize(st));
  T10_REQUIRE(h_c[1] < (1ll << 40), T10_ERR_UNSUPPORTED, "arena too large");
  for (int i = 0; i < 4; ++i) h_counts[i] = h_c[i];
  return T10_OK;
}

extern "C" int t10_unravel(int64_t n, int nstrides, const int64_t* h_strides, const int64_t* d_idx,
                           int64_t* d_out, t10_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  T10_REQUIRE(nstrides >= 0 && nstrides <= T10_MAX_RANK, T10_ERR_INVALID, "bad stride count");
  if (n == 0 || nstrides == 0) return T10_OK;
  SpaceDesc sd;
  memset(&sd, 0, sizeof(sd));
  for (int p = 0; p < nstrides; ++p) {
    T10_REQUIRE(h_strides[p] >= 1, T10_ERR_INVALID, "stride must be >=1");
    sd.dims[p] = h_strides[p];
  }
  k_unravel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(n, nstrides, sd, d_idx, d_out);
  T10_LAUNCH_CHECK();
  return T10_OK;
}

extern "C" int t10_relayout_tables(int64_t n, int nb, const int64_t* h_dims, const int64_t* h_srow,
                                   const int64_t* h_scol, const int64_t* d_idx, int32_t* d_to_row,
                                   int32_t* d_to_col, t10_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  T10_REQUIRE(nb >= 1 && nb <= T10_MAX_RANK, T10_ERR_INVALID, "bad bond count");
  if (n == 0) return T10_OK;
  TableDesc td;
  memset(&td, 0, sizeof(td));
  td.nb = nb;
  for (int b = 0; b < nb; ++b) {
    td.dims[b] = h_dims[b];
    td.srow[b] = h_srow[b];
    td.scol[b] = h_scol[b];
  }
  k_relayout_tables<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(n, td, d_idx, d_to_row, d_to_col);
  T10_LAUNCH_CHECK();
  return T10_OK;
}
