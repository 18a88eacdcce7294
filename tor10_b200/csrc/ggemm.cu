// Family 3: grouped GEMM over the matched symmetry sectors (one persistent launch).
//
// Reference call site: the Python loop `out.Storage[ob] = torch.matmul(A_q, B_q)` over
// output sectors, UniTensor.py:3727-3741 (one library GEMM launch per sector), and the
// single GEMM inside torch.tensordot for dense tensors, UniTensor.py:3769 / linalg.py:733.
//
// FP64: tcgen05 has no f64 kind; the FP64 tensor path on sm_100a is the warp-level DMMA
// (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4; the m16n8k{4,8,16} PTX shapes lower to the
// same instruction).  Kernel: persistent CTAs walk a tile table (heaviest tiles first),
// cp.async multi-stage pipeline global -> shared (zero-filled at ragged edges through the
// src-size operand), conflict-free padded shared layouts, DMMA from shared fragments,
// predicated vector stores.
// FP32: tile_cfg 50 (default) = tcgen05 / TMEM kind::tf32 three-pass split, ggemm_tf32.cu;
// tile_cfg 0 / 100 = SIMT register-tiled fp32 FMA kernel (comparison path).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "t10_common.cuh"

namespace t10 {

// ------------------------------------------------------------------ PTX helpers ----
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// ------------------------------------------------------------------ FP64 DMMA ------
template <int BM_, int BN_, int WM_, int WN_, int STAGES_, int BK_ = 16>
struct F64Cfg {
  static constexpr int BM = BM_, BN = BN_, BK = BK_, WM = WM_, WN = WN_, STAGES = STAGES_;
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int THREADS = WARPS_M * WARPS_N * 32;
  static constexpr int LDA = BK + 4;  // doubles; row stride = 8 banks mod 32 -> conflict-free LDS.64
  static constexpr int LDB = BN + 4;  // row stride = 4 doubles mod 16
  static constexpr int A_STAGE = BM * LDA, B_STAGE = BK * LDB;
  static constexpr int SMEM_BYTES = STAGES * (A_STAGE + B_STAGE) * 8;
  static constexpr int MI = WM / 8, NI = WN / 8;
};

template <class Cfg>
__device__ __forceinline__ void load_stage_f64(double* As, double* Bs, const double* __restrict__ A,
                                               const double* __restrict__ B, int lda, int ldb,
                                               int m_rem, int n_rem, int k_rem, bool a16, bool b16) {
  // A tile: BM rows x BK cols (K contiguous), m_rem/k_rem = valid rows/cols from the tile origin
  constexpr int ACH = Cfg::BM * (Cfg::BK / 2);
#pragma unroll
  for (int i = 0; i < (ACH + Cfg::THREADS - 1) / Cfg::THREADS; ++i) {
    int ch = threadIdx.x + i * Cfg::THREADS;
    if (ACH % Cfg::THREADS != 0 && ch >= ACH) break;
    int r = ch / (Cfg::BK / 2), kc = (ch % (Cfg::BK / 2)) * 2;
    double* s = As + r * Cfg::LDA + kc;
    int valid = (r < m_rem) ? max(0, min(2, k_rem - kc)) : 0;
    const double* g = valid ? (A + (int64_t)r * lda + kc) : A;
    if (a16) {
      cp_async16(s, g, valid * 8);
    } else {
      cp_async8(s, g, valid >= 1 ? 8 : 0);
      cp_async8(s + 1, valid >= 2 ? g + 1 : A, valid >= 2 ? 8 : 0);
    }
  }
  constexpr int BCH = Cfg::BK * (Cfg::BN / 2);
#pragma unroll
  for (int i = 0; i < (BCH + Cfg::THREADS - 1) / Cfg::THREADS; ++i) {
    int ch = threadIdx.x + i * Cfg::THREADS;
    if (BCH % Cfg::THREADS != 0 && ch >= BCH) break;
    int r = ch / (Cfg::BN / 2), nc = (ch % (Cfg::BN / 2)) * 2;
    double* s = Bs + r * Cfg::LDB + nc;
    int valid = (r < k_rem) ? max(0, min(2, n_rem - nc)) : 0;
    const double* g = valid ? (B + (int64_t)r * ldb + nc) : B;
    if (b16) {
      cp_async16(s, g, valid * 8);
    } else {
      cp_async8(s, g, valid >= 1 ? 8 : 0);
      cp_async8(s + 1, valid >= 2 ? g + 1 : B, valid >= 2 ? 8 : 0);
    }
  }
}

// Lean loader of the aligned fast path: pointers, clamped row/column offsets and shared-memory
// destinations are set up once per tile; a k-step costs one LDGSTS.128 + one 64-bit add per chunk.
// Rows beyond m / columns beyond n are CLAMPED to the last valid row / chunk instead of predicated:
// they only feed accumulators that are never stored.  Only the K tail needs zero fill (it would
// pollute valid outputs); that one step per tile goes through the predicated loader above.
template <class Cfg>
struct LeanLoader {
  static constexpr int A_PASSES = (Cfg::BM * (Cfg::BK / 2)) / Cfg::THREADS;
  static constexpr int B_PASSES = (Cfg::BK * (Cfg::BN / 2)) / Cfg::THREADS;
  static constexpr int A_RPP = Cfg::THREADS / (Cfg::BK / 2);   // rows per pass
  static constexpr int B_RPP = Cfg::THREADS / (Cfg::BN / 2);
  static_assert((Cfg::BM * (Cfg::BK / 2)) % Cfg::THREADS == 0 && (Cfg::BK * (Cfg::BN / 2)) % Cfg::THREADS == 0, "");
  const double* pa;
  const double* pb;
  int a_off[A_PASSES];   // element offsets of the pass rows (clamped) relative to pa
  int b_off[B_PASSES];
  int sa, sb;            // shared-memory element offsets of pass 0 inside a stage
  int64_t b_kstep;       // BK * ldb

  __device__ __forceinline__ void setup(const double* A, const double* B, int lda, int ldb, int m_rem, int n_rem) {
    const int ar0 = threadIdx.x / (Cfg::BK / 2), akc = (threadIdx.x % (Cfg::BK / 2)) * 2;
    const int br0 = threadIdx.x / (Cfg::BN / 2), bnc = (threadIdx.x % (Cfg::BN / 2)) * 2;
    const int base_row = min(ar0, m_rem - 1);
    pa = A + (int64_t)base_row * lda + akc;
#pragma unroll
    for (int i = 0; i < A_PASSES; ++i) a_off[i] = (min(ar0 + i * A_RPP, m_rem - 1) - base_row) * lda;
    const int last_chunk = ((n_rem - 1) >> 1) << 1;
    pb = B + (int64_t)br0 * ldb + min(bnc, last_chunk);
#pragma unroll
    for (int i = 0; i < B_PASSES; ++i) b_off[i] = i * B_RPP * ldb;
    sa = ar0 * Cfg::LDA + akc;
    sb = br0 * Cfg::LDB + bnc;
    b_kstep = (int64_t)Cfg::BK * ldb;
  }
  // loads the k-step the pointers currently stand on, then advances them
  __device__ __forceinline__ void issue(double* As, double* Bs) {
#pragma unroll
    for (int i = 0; i < A_PASSES; ++i) cp_async16(As + sa + i * A_RPP * Cfg::LDA, pa + a_off[i], 16);
#pragma unroll
    for (int i = 0; i < B_PASSES; ++i) cp_async16(Bs + sb + i * B_RPP * Cfg::LDB, pb + b_off[i], 16);
    pa += Cfg::BK;
    pb += b_kstep;
  }
};

// Fused GEMM -> all-gather: when n > 0 every output tile is written to ALL of these arenas (this rank's and
// its peers', mapped through NVLink peer memory) instead of Cbase, tile by tile, so the transfer rides under
// the remaining tiles' math.  The tile is staged through shared memory so that each store instruction of a
// warp covers one full 512-byte row segment (large NVLink write packets).
// When the last CTA of the launch has stored its tiles it raises this rank's flag in every peer's flag
// array to `step` (system-scope release); the consumer (t10_wait_copy on each rank) acquires all flags.
constexpr int T10_MAX_PEERS = 8;
struct PeerOut {
  int n;
  int self;                                  // index of the destination in THIS GPU's memory, or -1 (v5 kernels:
                                             // written by plain vector stores, the others by bulk copies)
  double* ptr[T10_MAX_PEERS];
  unsigned long long* flag[T10_MAX_PEERS];   // this rank's slot in rank q's flag array (NULL: no signalling)
  unsigned long long step;
  int* done_ctr;                             // zero-initialised arrival counter of this launch (reset on exit)
};

template <class Cfg, int MIN_CTAS, bool RAGGED>
__global__ void __launch_bounds__(Cfg::THREADS, MIN_CTAS)
k_ggemm_f64(const double* __restrict__ Abase, const double* __restrict__ Bbase,
            double* __restrict__ Cbase, const t10_gemm_problem* __restrict__ prob,
            int64_t ntiles, const int4* __restrict__ tiles, const int32_t* __restrict__ slot_start,
            int32_t* __restrict__ sched, int nbins, unsigned long long* __restrict__ trace,
            const __grid_constant__ PeerOut peers) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_next[2];
  double* As = smem;
  double* Bs = smem + Cfg::STAGES * Cfg::A_STAGE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int wm0 = (warp / Cfg::WARPS_N) * Cfg::WM, wn0 = (warp % Cfg::WARPS_N) * Cfg::WN;

  // Tile schedule, one of
  //   sm-bins  (sched != NULL and slot_start != NULL): the slot table holds one LPT bin per SM; the CTAs
  //            resident on an SM pull that SM's tiles (heaviest first) through sched[2 + bin], so the SM's
  //            tensor pipe stays fed until the whole bin is done, then steal from other bins.  Balance is at
  //            SM granularity (6 tiles per bin on the north-star workload instead of 2 per CTA slot).
  //            Correctness never depends on which SM a CTA lands on: every tile is claimed exactly once.
  //   dynamic  (sched != NULL): CTAs pull the next tile of the heaviest-first list through an atomic
  //            counter (list scheduling on measured, not modelled, tile times)
  //   slots    (slot_start != NULL): this CTA's LPT slot, tiles[slot_start[b] .. slot_start[b+1])
  //   strided  grid-strided walk over the heaviest-first list
  // In both counter modes the last CTA to leave resets the counters for the next launch on the stream.
  const bool smbins = sched != nullptr && slot_start != nullptr;
  int64_t t_beg = blockIdx.x, t_end = ntiles, t_step = gridDim.x;
  if (slot_start != nullptr && !smbins) {
    t_beg = slot_start[blockIdx.x];
    t_end = slot_start[blockIdx.x + 1];
    t_step = 1;
  }
  int my_bin = 0, pending = 0;       // warp 0: this SM's bin; lane 0: the claim in flight on it
  bool own_open = true;              // warp 0, uniform: the own bin may still hold tiles
  if (smbins) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    my_bin = (int)(smid % (unsigned)nbins);
  }
  // warp 0 only: resolve the claim in flight on the own bin, or steal; returns a tile index or -1
  auto next_smbin = [&]() -> int {
    int t = -1;
    if (own_open) {
      const int idx = __shfl_sync(0xffffffffu, pending, 0);
      if (slot_start[my_bin] + idx < slot_start[my_bin + 1]) t = slot_start[my_bin] + idx;
      else own_open = false;
    }
    for (int base = 1; t < 0 && base < nbins; base += 32) {
      const int off = base + lane;
      int b2 = 0, left = 0;
      if (off < nbins) {
        b2 = (my_bin + off) % nbins;
        left = (slot_start[b2 + 1] - slot_start[b2]) - *((volatile int32_t*)&sched[2 + b2]);
      }
      unsigned cand = __ballot_sync(0xffffffffu, left > 0);
      while (t < 0 && cand) {
        const int src = __ffs(cand) - 1;
        int got = -1;
        if (lane == src) {
          const int idx = atomicAdd(&sched[2 + b2], 1);
          if (slot_start[b2] + idx < slot_start[b2 + 1]) got = slot_start[b2] + idx;
        }
        t = __shfl_sync(0xffffffffu, got, src);
        cand &= cand - 1;
      }
    }
    return t;
  };
  if (smbins) {
    if (warp == 0) {
      if (lane == 0) pending = atomicAdd(&sched[2 + my_bin], 1);
      const int t0 = next_smbin();
      if (lane == 0) s_next[0] = t0;
    }
    __syncthreads();
  } else if (sched != nullptr) {
    if (threadIdx.x == 0) s_next[0] = atomicAdd(&sched[0], 1);
    __syncthreads();
  }
  for (int64_t t = t_beg, it = 0;; t += t_step, ++it) {
    if (smbins) {
      t = s_next[it & 1];
      if (t < 0) break;
    } else if (sched != nullptr) {
      t = s_next[it & 1];
      if (t >= ntiles) break;
    } else if (t >= t_end) {
      break;
    }
    // Counter modes claim the NEXT tile a few k-steps before this one ends: late enough that the claim
    // reflects who really finishes first (claiming at the start of a tile would hand out two tiles per CTA
    // at t = 0, a static round-robin), early enough that the atomic's latency hides behind the last k-steps.
    auto claim_next = [&]() {
      if (smbins) {
        if (warp == 0 && own_open && lane == 0) pending = atomicAdd(&sched[2 + my_bin], 1);
      } else if (sched != nullptr) {
        if (threadIdx.x == 0) pending = atomicAdd(&sched[0], 1);
      }
    };
    const int4 tile = tiles[t];
    const t10_gemm_problem p = prob[tile.x];
    const int m0 = tile.y, n0 = tile.z;
    if (trace != nullptr && threadIdx.x == 0) {   // profiling aid (t10_ggemm_set_trace): {smid, cta, t0, t1} per tile
      unsigned smid;
      unsigned long long now;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      trace[4 * t + 0] = smid;
      trace[4 * t + 1] = blockIdx.x;
      trace[4 * t + 2] = now;
    }
    const double* A = Abase + p.a_off + (int64_t)m0 * p.lda;
    const double* B = Bbase + p.b_off + n0;
    const int m_rem = p.m - m0, n_rem = p.n - n0;
    const bool a16 = ((p.lda & 1) == 0) && ((p.a_off & 1) == 0) && (((uintptr_t)Abase & 15) == 0);
    const bool b16 = ((p.ldb & 1) == 0) && (((p.b_off + n0) & 1) == 0) && (((uintptr_t)Bbase & 15) == 0);
    const int KT = (p.k + Cfg::BK - 1) / Cfg::BK;
    // k-steps served by the lean loader: all full ones when both operands are 16-byte aligned
    const int KT_fast = (a16 && b16) ? p.k / Cfg::BK : 0;
    const int mi_valid = max(0, min(Cfg::MI, (m_rem - wm0 + 7) >> 3));
    const int ni_valid = max(0, min(Cfg::NI, (n_rem - wn0 + 7) >> 3));
    const bool full = !RAGGED || ((mi_valid == Cfg::MI) && (ni_valid == Cfg::NI));

    double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
    for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
      for (int j = 0; j < Cfg::NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    LeanLoader<Cfg> ld;
    ld.setup(A, B, p.lda, p.ldb, m_rem, n_rem);
    auto issue_stage = [&](int kt) {  // kt is issued in increasing order, so the lean pointers stay in step
      double* as_ = As + (kt % Cfg::STAGES) * Cfg::A_STAGE;
      double* bs_ = Bs + (kt % Cfg::STAGES) * Cfg::B_STAGE;
      if (kt < KT_fast)
        ld.issue(as_, bs_);
      else
        load_stage_f64<Cfg>(as_, bs_, A + kt * Cfg::BK, B + (int64_t)kt * Cfg::BK * p.ldb, p.lda, p.ldb, m_rem,
                            n_rem, p.k - kt * Cfg::BK, a16, b16);
    };

#pragma unroll
    for (int s = 0; s < Cfg::STAGES - 1; ++s) {
      if (s < KT) issue_stage(s);
      cp_async_commit();
    }
    const int claim_kt = max(KT - 4, 0);
    if (KT == 0) claim_next();
    for (int kt = 0; kt < KT; ++kt) {
      cp_async_wait<Cfg::STAGES - 2>();
      __syncthreads();
      if (kt + Cfg::STAGES - 1 < KT) issue_stage(kt + Cfg::STAGES - 1);
      cp_async_commit();
      if (kt == claim_kt) claim_next();
      const double* as = As + (kt % Cfg::STAGES) * Cfg::A_STAGE + (wm0 + g) * Cfg::LDA + tig;
      const double* bs = Bs + (kt % Cfg::STAGES) * Cfg::B_STAGE + tig * Cfg::LDB + wn0 + g;
      if (full) {
#pragma unroll
        for (int kk = 0; kk < Cfg::BK; kk += 4) {
          double af[Cfg::MI], bf[Cfg::NI];
#pragma unroll
          for (int i = 0; i < Cfg::MI; ++i) af[i] = as[i * 8 * Cfg::LDA + kk];
#pragma unroll
          for (int j = 0; j < Cfg::NI; ++j) bf[j] = bs[kk * Cfg::LDB + j * 8];
#pragma unroll
          for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
            for (int j = 0; j < Cfg::NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
      } else if (mi_valid > 0 && ni_valid > 0) {
        // ragged tile: 8x8 fragments that lie entirely outside the sector are skipped (warp-uniform)
#pragma unroll
        for (int kk = 0; kk < Cfg::BK; kk += 4) {
          double af[Cfg::MI], bf[Cfg::NI];
#pragma unroll
          for (int i = 0; i < Cfg::MI; ++i) af[i] = as[i * 8 * Cfg::LDA + kk];
#pragma unroll
          for (int j = 0; j < Cfg::NI; ++j) bf[j] = bs[kk * Cfg::LDB + j * 8];
#pragma unroll
          for (int i = 0; i < Cfg::MI; ++i) {
            if (i < mi_valid) {
#pragma unroll
              for (int j = 0; j < Cfg::NI; ++j)
                if (j < ni_valid) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
          }
        }
      }
    }
    cp_async_wait<0>();

    // epilogue: lane holds C[row g][cols 2*tig, 2*tig+1] of every 8x8 fragment
    const bool c16 = ((p.ldc & 1) == 0) && ((p.c_off & 1) == 0) && (((uintptr_t)Cbase & 15) == 0);
    if (peers.n > 0) {
      // Fused all-gather: registers -> a dedicated shared tile -> one TMA bulk copy (cp.async.bulk, shared ->
      // global) per tile row and rank, issued by warp 0 and left in flight while the CTA computes its next
      // tile.  The copy engine streams whole rows over NVLink; no LSU stores, no registers.
      constexpr int CLD = Cfg::BN + 2;
      double* Cs = smem + Cfg::STAGES * (Cfg::A_STAGE + Cfg::B_STAGE);
      if (warp == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous tile has left Cs
      __syncthreads();
#pragma unroll
      for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
        for (int j = 0; j < Cfg::NI; ++j)
          *reinterpret_cast<double2*>(Cs + (wm0 + i * 8 + g) * CLD + wn0 + j * 8 + 2 * tig) =
              make_double2(acc[i][j][0], acc[i][j][1]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> async-proxy reads
      __syncthreads();
      bool p16 = ((p.ldc & 1) == 0) && ((p.c_off & 1) == 0);
      for (int q = 0; q < peers.n; ++q) p16 = p16 && (((uintptr_t)peers.ptr[q] & 15) == 0);
      const int rows = min(Cfg::BM, m_rem), cols = min(Cfg::BN, n_rem);
      if (p16) {
        if (warp == 0) {
          const unsigned bytes = (unsigned)(cols & ~1) * 8u;
          for (int r = lane; r < rows; r += 32) {
            const int64_t off = p.c_off + (int64_t)(m0 + r) * p.ldc + n0;
            const unsigned src = (unsigned)__cvta_generic_to_shared(Cs + r * CLD);
#pragma unroll 1
            for (int q = 0; q < peers.n; ++q) {
              double* o = peers.ptr[q] + off;
              if (bytes)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(o), "r"(src), "r"(bytes)
                             : "memory");
              if (cols & 1) o[cols - 1] = Cs[r * CLD + cols - 1];
            }
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else {
        for (int idx = threadIdx.x; idx < Cfg::BM * Cfg::BN; idx += Cfg::THREADS) {
          const int r = idx / Cfg::BN, c = idx % Cfg::BN;
          if (r >= rows || c >= cols) continue;
          const double v = Cs[r * CLD + c];
          const int64_t off = p.c_off + (int64_t)(m0 + r) * p.ldc + n0 + c;
#pragma unroll 1
          for (int q = 0; q < peers.n; ++q) peers.ptr[q][off] = v;
        }
      }
    } else {
      double* C = Cbase + p.c_off;
#pragma unroll
      for (int i = 0; i < Cfg::MI; ++i) {
        const int r = m0 + wm0 + i * 8 + g;
        if (r >= p.m) continue;
#pragma unroll
        for (int j = 0; j < Cfg::NI; ++j) {
          const int c = n0 + wn0 + j * 8 + 2 * tig;
          double* o = C + (int64_t)r * p.ldc + c;
          if (c + 1 < p.n) {
            if (c16) *reinterpret_cast<double2*>(o) = make_double2(acc[i][j][0], acc[i][j][1]);
            else { o[0] = acc[i][j][0]; o[1] = acc[i][j][1]; }
          } else if (c < p.n) {
            o[0] = acc[i][j][0];
          }
        }
      }
    }
    if (trace != nullptr && threadIdx.x == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      trace[4 * t + 3] = now;
    }
    if (smbins && warp == 0) {
      const int tn = next_smbin();
      if (lane == 0) s_next[(it + 1) & 1] = tn;
    } else if (!smbins && sched != nullptr && threadIdx.x == 0) {
      s_next[(it + 1) & 1] = pending;
    }
    __syncthreads();  // stages free for the next tile; next tile index visible
  }
  if (peers.n > 0) {   // the bulk copies still in flight must have landed before this rank signals
    if (warp == 0) {
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();
  }
  if (peers.n > 0 && peers.done_ctr != nullptr && threadIdx.x == 0) {
    // every thread's tile stores are ordered before this point by the CTA barrier above
    __threadfence_system();
    if (atomicAdd(peers.done_ctr, 1) == (int)gridDim.x - 1) {
      *peers.done_ctr = 0;
      __threadfence_system();
      for (int q = 0; q < peers.n; ++q)
        if (peers.flag[q] != nullptr)
          asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers.flag[q]), "l"(peers.step) : "memory");
    }
  }
  if (sched != nullptr && threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&sched[1], 1) == (int)gridDim.x - 1) {
      sched[0] = 0;
      sched[1] = 0;
      for (int b = 0; b < nbins; ++b) sched[2 + b] = 0;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------ FP64 DMMA, aligned-only lean kernel
// Main loop organised like CUTLASS's multistage mainloop (the kernel cuBLAS runs for FP64 on this chip):
// register double-buffered fragments, the next k-group's LDS issued BEFORE the current group's DMMAs, the
// cp.asyncs of the stage being prefetched spread over the first three k-groups, the single wait+barrier
// of a k-step placed before its last group.  ONE loader: every operand row start is 16-byte aligned
// (host-checked), M/N edges are clamped, the K tail is zero-filled through the cp.async src-size operand
// in the same instruction stream (no second code path, no spills).  Used whenever every problem of the group has even lda/ldb and
// even operand offsets; other groups run k_ggemm_f64.
// One tile (m0, n0) of problem p through the software-pipelined main loop; `smem` = this CTA's dynamic shared
// memory, `cs_off` = offset (doubles) of the staging tile of the fused all-gather behind the pipeline stages.
template <class Cfg, bool SK = false>
__device__ __forceinline__ void v5_tile(const double* __restrict__ Abase, const double* __restrict__ Bbase,
                                        double* __restrict__ Cbase, const t10_gemm_problem& p, const int m0,
                                        const int n0, const int m_lim, const int n_lim, const int64_t t,
                                        double* smem, const int cs_off, unsigned long long* __restrict__ trace,
                                        const PeerOut& peers, const int4 piece = make_int4(0, 0, -1, 0),
                                        double* __restrict__ sk_ws = nullptr, int* __restrict__ sk_flags = nullptr,
                                        const int sk_epoch = 0) {
  static_assert(Cfg::BK == 16, "four k-groups per stage");
  double* As = smem;
  double* Bs = smem + Cfg::STAGES * Cfg::A_STAGE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int wm0 = (warp / Cfg::WARPS_N) * Cfg::WM, wn0 = (warp % Cfg::WARPS_N) * Cfg::WN;
  constexpr int A_PASSES = (Cfg::BM * (Cfg::BK / 2)) / Cfg::THREADS;
  constexpr int B_PASSES = (Cfg::BK * (Cfg::BN / 2)) / Cfg::THREADS;
  constexpr int A_RPP = Cfg::THREADS / (Cfg::BK / 2), B_RPP = Cfg::THREADS / (Cfg::BN / 2);
  constexpr int NPASS = A_PASSES + B_PASSES;
  const int ar0 = threadIdx.x / (Cfg::BK / 2), akc = (threadIdx.x % (Cfg::BK / 2)) * 2;
  const int br0 = threadIdx.x / (Cfg::BN / 2), bnc = (threadIdx.x % (Cfg::BN / 2)) * 2;
  const unsigned sa0 = (unsigned)__cvta_generic_to_shared(As + ar0 * Cfg::LDA + akc);
  const unsigned sb0 = (unsigned)__cvta_generic_to_shared(Bs + br0 * Cfg::LDB + bnc);

  // valid extent of this tile: up to the sector edge, or up to the region the planner gave it (mixed shapes)
  const int m_rem = m_lim > 0 ? m_lim : p.m - m0, n_rem = n_lim > 0 ? n_lim : p.n - n0;
  const int m_end = m0 + m_rem, n_end = n0 + n_rem;
  const int KT = (p.k + Cfg::BK - 1) / Cfg::BK;
  const bool active = (m_rem > wm0) && (n_rem > wn0);
  // Stream-K (SK): this call covers k-steps [kb, ke) of the tile.  A piece without the tile's last k-step parks
  // its accumulators in workspace slot `pslot` (= this CTA) and raises that slot's flag; the piece with the last
  // k-step adds the `nfix` parked slots ffix, ffix+1, ... in that order (fixed summation order) and stores.
  const int kb = SK ? piece.x : 0, ke = SK ? piece.y : KT;
  const int pslot = SK ? piece.z : -1, nfix = SK ? (piece.w & 0xff) : 0, ffix = SK ? (piece.w >> 8) : 0;
  if (trace != nullptr && threadIdx.x == 0) {   // profiling aid (t10_ggemm_set_trace)
    unsigned smid;
    unsigned long long now;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    trace[4 * t + 0] = smid;
    trace[4 * t + 1] = blockIdx.x;
    trace[4 * t + 2] = now;
  }
  // per-thread source pointers; rows beyond m and column chunks beyond n are clamped (never stored)
  const double* pa[A_PASSES];
#pragma unroll
  for (int i = 0; i < A_PASSES; ++i)
    pa[i] = Abase + p.a_off + (int64_t)(m0 + min(ar0 + i * A_RPP, m_rem - 1)) * p.lda + akc;
  const double* pb = Bbase + p.b_off + n0 + min(bnc, ((n_rem - 1) >> 1) << 1) + (int64_t)br0 * p.ldb;
  const int64_t b_pass = (int64_t)B_RPP * p.ldb, b_kstep = (int64_t)Cfg::BK * p.ldb;

  double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
  for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
    for (int j = 0; j < Cfg::NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // passes [lo, hi) of stage kt; k validity through the src-size operand (zero fill)
  auto issue = [&](int kt, int lo, int hi) {
    const int k0 = kt * Cfg::BK;
    const unsigned so_a = (unsigned)((kt % Cfg::STAGES) * Cfg::A_STAGE * 8);
    const unsigned so_b = (unsigned)((kt % Cfg::STAGES) * Cfg::B_STAGE * 8);
    const int a_bytes = max(0, min(2, p.k - k0 - akc)) * 8;
#pragma unroll
    for (int q = 0; q < NPASS; ++q) {
      if (q < lo || q >= hi) continue;
      if (q < A_PASSES) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa0 + so_a + q * A_RPP * Cfg::LDA * 8),
                     "l"(a_bytes ? pa[q] + k0 : Abase), "r"(a_bytes));
      } else {
        const int i = q - A_PASSES;
        const int b_bytes = (k0 + br0 + i * B_RPP < p.k) ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sb0 + so_b + i * B_RPP * Cfg::LDB * 8),
                     "l"(b_bytes ? pb + (int64_t)kt * b_kstep + i * b_pass : Bbase), "r"(b_bytes));
      }
    }
  };

  // cs_off < 0: the staging tile of the fused all-gather aliases the pipeline stages (keeps two CTAs per SM for
  // the mixed-shape kernels); the previous tile's bulk copies must then have read it before new loads land
  if (cs_off < 0 && peers.n > 0 && warp == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncthreads();  // previous tile's stages free
#pragma unroll
  for (int s2 = 0; s2 < Cfg::STAGES - 1; ++s2) {
    if (kb + s2 < ke) issue(kb + s2, 0, NPASS);
    cp_async_commit();
  }
  cp_async_wait<Cfg::STAGES - 2>();
  __syncthreads();

  double af[2][Cfg::MI], bf[2][Cfg::NI];
  auto load_frags = [&](int kt, int grp, int buf) {
    const double* as = As + (kt % Cfg::STAGES) * Cfg::A_STAGE + (wm0 + g) * Cfg::LDA + tig + grp * 4;
    const double* bs = Bs + (kt % Cfg::STAGES) * Cfg::B_STAGE + (tig + grp * 4) * Cfg::LDB + wn0 + g;
#pragma unroll
    for (int i = 0; i < Cfg::MI; ++i) af[buf][i] = as[i * 8 * Cfg::LDA];
#pragma unroll
    for (int j = 0; j < Cfg::NI; ++j) bf[buf][j] = bs[j * 8];
  };
  if (ke > kb) load_frags(kb, 0, 0);

  for (int kt = kb; kt < ke; ++kt) {
    const int nk = kt + Cfg::STAGES - 1;
#pragma unroll
    for (int grp = 0; grp < 4; ++grp) {
      if (grp < 3) load_frags(kt, grp + 1, (grp + 1) & 1);
      else if (kt + 1 < ke) load_frags(kt + 1, 0, 0);
      if (grp < 3 && nk < ke) issue(nk, grp * NPASS / 3, (grp + 1) * NPASS / 3);
      if (grp == 2) {
        cp_async_commit();
        cp_async_wait<Cfg::STAGES - 2>();
        __syncthreads();
      }
      if (active) {
#pragma unroll
        for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
          for (int j = 0; j < Cfg::NI; ++j)
            dmma884(acc[i][j][0], acc[i][j][1], af[grp & 1][i], bf[grp & 1][j]);
      }
    }
  }
  cp_async_wait<0>();

  if (trace != nullptr && threadIdx.x == 0) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    trace[4 * t + 3] = now;
  }
  if (SK && pslot >= 0) {   // park the partial accumulators, publish, done with this piece
    if (active) {
      double2* w = reinterpret_cast<double2*>(sk_ws + (int64_t)pslot * (128 * 34)) + threadIdx.x;
#pragma unroll
      for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
        for (int j = 0; j < Cfg::NI; ++j) w[(i * Cfg::NI + j) * Cfg::THREADS] = make_double2(acc[i][j][0], acc[i][j][1]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(sk_flags + pslot), "r"(sk_epoch) : "memory");
    }
    return;
  }
  if (SK && nfix > 0) {
    if (threadIdx.x == 0) {
      for (int f = 0; f < nfix; ++f) {
        int v = sk_epoch - 1;
        for (int spin = 0; spin < (1 << 22); ++spin) {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(sk_flags + ffix + f) : "memory");
          if (v == sk_epoch) break;
          __nanosleep(64);
        }
        if (v != sk_epoch) sk_flags[gridDim.x] = 1;   // a parked partial never arrived (reported by the host)
      }
    }
    __syncthreads();
    if (active) {
      for (int f = 0; f < nfix; ++f) {
        const double2* w = reinterpret_cast<const double2*>(sk_ws + (int64_t)(ffix + f) * (128 * 34)) + threadIdx.x;
#pragma unroll
        for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
          for (int j = 0; j < Cfg::NI; ++j) {
            const double2 v = __ldcg(w + (i * Cfg::NI + j) * Cfg::THREADS);
            acc[i][j][0] += v.x;
            acc[i][j][1] += v.y;
          }
      }
    }
  }
  if (peers.n > 0) {
    // Fused all-gather: registers -> a dedicated shared tile -> one TMA bulk copy (cp.async.bulk, shared ->
    // global) per tile row and rank, issued by warp 0 and left in flight while the CTA computes its next
    // tile.  The copy engine streams whole 512-byte rows over NVLink; no LSU stores, no registers.
    constexpr int CLD = Cfg::BN + 2;
    double* Cs = smem + (cs_off < 0 ? 0 : cs_off);
    if (cs_off >= 0 && warp == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous tile has left Cs
    __syncthreads();   // (alias mode: every warp is done reading the last stage)
    if (active) {
#pragma unroll
      for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
        for (int j = 0; j < Cfg::NI; ++j)
          *reinterpret_cast<double2*>(Cs + (wm0 + i * 8 + g) * CLD + wn0 + j * 8 + 2 * tig) =
              make_double2(acc[i][j][0], acc[i][j][1]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> async-proxy reads
    }
    __syncthreads();
    bool p16 = ((p.ldc & 1) == 0) && ((p.c_off & 1) == 0);
    for (int q = 0; q < peers.n; ++q) p16 = p16 && (((uintptr_t)peers.ptr[q] & 15) == 0);
    const int rows = min(Cfg::BM, m_rem), cols = min(Cfg::BN, n_rem);
    if (p16) {
      if (warp == 0) {
        const unsigned bytes = (unsigned)(cols & ~1) * 8u;
        for (int r = lane; r < rows; r += 32) {
          const int64_t off = p.c_off + (int64_t)(m0 + r) * p.ldc + n0;
          const unsigned src = (unsigned)__cvta_generic_to_shared(Cs + r * CLD);
#pragma unroll 1
          for (int q = 0; q < peers.n; ++q) {
            if (q == peers.self) continue;
            double* o = peers.ptr[q] + off;
            if (bytes)
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(o), "r"(src), "r"(bytes)
                           : "memory");
            if (cols & 1) o[cols - 1] = Cs[r * CLD + cols - 1];
          }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (peers.self >= 0) {
        // the local destination: 16-byte stores by all threads (a 512-byte bulk copy costs the SM's copy engine
        // ~125 ns whatever its destination; 64 of them per tile are slower than the LSU path for local memory)
        double* o = peers.ptr[peers.self] + p.c_off + (int64_t)m0 * p.ldc + n0;
        for (int idx = threadIdx.x; idx < rows * (Cfg::BN / 2); idx += Cfg::THREADS) {
          const int r = idx / (Cfg::BN / 2), c2 = (idx % (Cfg::BN / 2)) * 2;
          if (c2 + 1 < cols)
            *reinterpret_cast<double2*>(o + (int64_t)r * p.ldc + c2) = *reinterpret_cast<const double2*>(Cs + r * CLD + c2);
          else if (c2 < cols)
            o[(int64_t)r * p.ldc + c2] = Cs[r * CLD + c2];
        }
      }
    } else {
      for (int idx = threadIdx.x; idx < Cfg::BM * Cfg::BN; idx += Cfg::THREADS) {
        const int r = idx / Cfg::BN, c = idx % Cfg::BN;
        if (r >= rows || c >= cols) continue;
        const double v = Cs[r * CLD + c];
        const int64_t off = p.c_off + (int64_t)(m0 + r) * p.ldc + n0 + c;
#pragma unroll 1
        for (int q = 0; q < peers.n; ++q) peers.ptr[q][off] = v;
      }
    }
    return;
  }
  if (!active) return;
  double* C = Cbase + p.c_off;
  const bool c16 = ((p.ldc & 1) == 0) && ((p.c_off & 1) == 0) && (((uintptr_t)Cbase & 15) == 0);
#pragma unroll
  for (int i = 0; i < Cfg::MI; ++i) {
    const int r = m0 + wm0 + i * 8 + g;
    if (r >= m_end) continue;
#pragma unroll
    for (int j = 0; j < Cfg::NI; ++j) {
      const int c = n0 + wn0 + j * 8 + 2 * tig;
      double* o = C + (int64_t)r * p.ldc + c;
      if (c + 1 < n_end) {
        if (c16) *reinterpret_cast<double2*>(o) = make_double2(acc[i][j][0], acc[i][j][1]);
        else { o[0] = acc[i][j][0]; o[1] = acc[i][j][1]; }
      } else if (c < n_end) {
        o[0] = acc[i][j][0];
      }
    }
  }
}

template <class Cfg, int MIN_CTAS>
__global__ void __launch_bounds__(Cfg::THREADS, MIN_CTAS)
k_ggemm_f64_v5(const double* __restrict__ Abase, const double* __restrict__ Bbase,
               double* __restrict__ Cbase, const t10_gemm_problem* __restrict__ prob,
               int64_t ntiles, const int4* __restrict__ tiles, const int32_t* __restrict__ slot_start,
               unsigned long long* __restrict__ trace, const __grid_constant__ PeerOut peers) {
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5;
  int64_t t_beg = blockIdx.x, t_end = ntiles, t_step = gridDim.x;
  if (slot_start != nullptr) {
    t_beg = slot_start[blockIdx.x];
    t_end = slot_start[blockIdx.x + 1];
    t_step = 1;
  }
  auto run_tile = [&](int64_t t) {
    const int4 tile = tiles[t];
    const t10_gemm_problem p = prob[tile.x];
    v5_tile<Cfg>(Abase, Bbase, Cbase, p, tile.y, tile.z, 0, 0, t, smem, Cfg::STAGES * (Cfg::A_STAGE + Cfg::B_STAGE), trace,
                 peers);
  };
  for (int64_t t = t_beg; t < t_end; t += t_step) run_tile(t);
  if (peers.n > 0) {   // the bulk copies still in flight must have landed before this rank signals
    if (warp == 0) {
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async;" ::: "memory");
    }
  }
  if (peers.n > 0 && peers.done_ctr != nullptr) {
    __syncthreads();   // every thread's tile stores precede thread 0's fence
    if (threadIdx.x == 0) {
      __threadfence_system();
      if (atomicAdd(peers.done_ctr, 1) == (int)gridDim.x - 1) {
        *peers.done_ctr = 0;
        __threadfence_system();
        for (int q = 0; q < peers.n; ++q)
          if (peers.flag[q] != nullptr)
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers.flag[q]), "l"(peers.step) : "memory");
      }
    }
  }
}

// Mixed tile shapes in ONE launch: tiles[t].w selects 64x64 (0), 32x128 (1) or 128x32 (2), all on four warps of
// 32x32 with the same main loop.  A ragged sector's last rows (columns), when at most 32 of them are left, are
// covered by 32x128 (128x32) strips with all four warps busy instead of 64x64 tiles with half of them idle:
// 6.8 % fewer tile k-steps on the north-star workload (padding 13.3 % -> 5.6 %).
using MixSq = F64Cfg<64, 64, 32, 32, 4>;
using MixRow = F64Cfg<32, 128, 32, 32, 4>;     // bottom strip: 1 x 4 warps
using MixCol = F64Cfg<128, 32, 32, 32, 4>;     // right strip: 4 x 1 warps
constexpr int MIX_STAGE_DOUBLES = 4 * (128 * 20 + 16 * 36);                  // largest pipeline (128x32)
constexpr int MIX_CS_DOUBLES = 128 * 34;                                      // largest staging tile
constexpr int MIX_SMEM_BYTES = MIX_STAGE_DOUBLES * 8;
static_assert(MixSq::STAGES * (MixSq::A_STAGE + MixSq::B_STAGE) <= MIX_STAGE_DOUBLES, "");
static_assert(MixRow::STAGES * (MixRow::A_STAGE + MixRow::B_STAGE) <= MIX_STAGE_DOUBLES, "");
static_assert(MixCol::STAGES * (MixCol::A_STAGE + MixCol::B_STAGE) <= MIX_STAGE_DOUBLES, "");
static_assert(64 * 66 <= MIX_CS_DOUBLES && 32 * 130 <= MIX_CS_DOUBLES && MIX_CS_DOUBLES <= MIX_STAGE_DOUBLES, "");

template <int MIN_CTAS>
__global__ void __launch_bounds__(128, MIN_CTAS)
k_ggemm_f64_mixed(const double* __restrict__ Abase, const double* __restrict__ Bbase,
                  double* __restrict__ Cbase, const t10_gemm_problem* __restrict__ prob,
                  int64_t ntiles, const int4* __restrict__ tiles, const int32_t* __restrict__ slot_start,
                  unsigned long long* __restrict__ trace, const __grid_constant__ PeerOut peers) {
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5;
  int64_t t_beg = blockIdx.x, t_end = ntiles, t_step = gridDim.x;
  if (slot_start != nullptr) {
    t_beg = slot_start[blockIdx.x];
    t_end = slot_start[blockIdx.x + 1];
    t_step = 1;
  }
  auto run_tile = [&](int64_t t) {
    const int4 tile = tiles[t];
    const t10_gemm_problem p = prob[tile.x];
    // tile.w = shape | valid rows << 8 | valid columns << 20 (0 = up to the sector edge)
    const int shape = tile.w & 3, m_lim = (tile.w >> 8) & 0xfff, n_lim = (tile.w >> 20) & 0xfff;
    if (shape == 1)
      v5_tile<MixRow>(Abase, Bbase, Cbase, p, tile.y, tile.z, m_lim, n_lim, t, smem, -1, trace, peers);
    else if (shape == 2)
      v5_tile<MixCol>(Abase, Bbase, Cbase, p, tile.y, tile.z, m_lim, n_lim, t, smem, -1, trace, peers);
    else
      v5_tile<MixSq>(Abase, Bbase, Cbase, p, tile.y, tile.z, m_lim, n_lim, t, smem, -1, trace, peers);
  };
  for (int64_t t = t_beg; t < t_end; t += t_step) run_tile(t);
  if (peers.n > 0) {   // the bulk copies still in flight must have landed before this rank signals
    if (warp == 0) {
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async;" ::: "memory");
    }
  }
  if (peers.n > 0 && peers.done_ctr != nullptr) {
    __syncthreads();   // every thread's tile stores precede thread 0's fence
    if (threadIdx.x == 0) {
      __threadfence_system();
      if (atomicAdd(peers.done_ctr, 1) == (int)gridDim.x - 1) {
        *peers.done_ctr = 0;
        __threadfence_system();
        for (int q = 0; q < peers.n; ++q)
          if (peers.flag[q] != nullptr)
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers.flag[q]), "l"(peers.step) : "memory");
      }
    }
  }
}

// Stream-K form of the mixed-shape kernel: the table lists PIECES of tiles (k-step ranges); every CTA gets the same
// number of k-steps, so all of them finish together whatever the tile weights are.  See t10_ggemm_streamk.
template <int MIN_CTAS>
__global__ void __launch_bounds__(128, MIN_CTAS)
k_ggemm_f64_mixed_sk(const double* __restrict__ Abase, const double* __restrict__ Bbase,
                     double* __restrict__ Cbase, const t10_gemm_problem* __restrict__ prob,
                     const int4* __restrict__ tiles, const int4* __restrict__ pieces,
                     const int32_t* __restrict__ slot_start, double* __restrict__ sk_ws, int* __restrict__ sk_flags,
                     int sk_epoch, unsigned long long* __restrict__ trace, const __grid_constant__ PeerOut peers) {
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5;
  const int64_t t_beg = slot_start[blockIdx.x], t_end = slot_start[blockIdx.x + 1];
  for (int64_t t = t_beg; t < t_end; ++t) {
    const int4 tile = tiles[t];
    const int4 pc = pieces[t];
    const t10_gemm_problem p = prob[tile.x];
    const int shape = tile.w & 3, m_lim = (tile.w >> 8) & 0xfff, n_lim = (tile.w >> 20) & 0xfff;
    if (shape == 1)
      v5_tile<MixRow, true>(Abase, Bbase, Cbase, p, tile.y, tile.z, m_lim, n_lim, t, smem, -1, trace, peers, pc, sk_ws,
                            sk_flags, sk_epoch);
    else if (shape == 2)
      v5_tile<MixCol, true>(Abase, Bbase, Cbase, p, tile.y, tile.z, m_lim, n_lim, t, smem, -1, trace, peers, pc, sk_ws,
                            sk_flags, sk_epoch);
    else
      v5_tile<MixSq, true>(Abase, Bbase, Cbase, p, tile.y, tile.z, m_lim, n_lim, t, smem, -1, trace, peers, pc, sk_ws,
                           sk_flags, sk_epoch);
  }

  if (peers.n > 0) {   // the bulk copies still in flight must have landed before this rank signals
    if (warp == 0) {
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async;" ::: "memory");
    }
  }
  if (peers.n > 0 && peers.done_ctr != nullptr) {
    __syncthreads();   // every thread's tile stores precede thread 0's fence
    if (threadIdx.x == 0) {
      __threadfence_system();
      if (atomicAdd(peers.done_ctr, 1) == (int)gridDim.x - 1) {
        *peers.done_ctr = 0;
        __threadfence_system();
        for (int q = 0; q < peers.n; ++q)
          if (peers.flag[q] != nullptr)
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers.flag[q]), "l"(peers.step) : "memory");
      }
    }
  }
}

// ------------------------------------------------------------------ SIMT (fp32, debug f64)
// 64x64 tile, BK = 16, 256 threads, 4x4 micro-tile per thread, double-buffered via registers.
template <typename T>
__global__ void __launch_bounds__(256)
k_ggemm_simt(const T* __restrict__ Abase, const T* __restrict__ Bbase, T* __restrict__ Cbase,
             const t10_gemm_problem* __restrict__ prob, int64_t ntiles,
             const int4* __restrict__ tiles) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ T As[BK][BM + 4];  // transposed: As[k][m]
  __shared__ T Bs[BK][BN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int4 tile = tiles[t];
    const t10_gemm_problem p = prob[tile.x];
    const int m0 = tile.y, n0 = tile.z;
    const T* A = Abase + p.a_off;
    const T* B = Bbase + p.b_off;
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
    for (int k0 = 0; k0 < p.k; k0 += BK) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < (BM * BK) / 256; ++i) {
        int e = threadIdx.x + i * 256;
        int r = e / BK, kk = e % BK;  // consecutive threads along K (A row-major)
        int gr = m0 + r, gk = k0 + kk;
        As[kk][r] = (gr < p.m && gk < p.k) ? A[(int64_t)gr * p.lda + gk] : T(0);
      }
#pragma unroll
      for (int i = 0; i < (BK * BN) / 256; ++i) {
        int e = threadIdx.x + i * 256;
        int kk = e / BN, c = e % BN;
        int gk = k0 + kk, gc = n0 + c;
        Bs[kk][c] = (gk < p.k && gc < p.n) ? B[(int64_t)gk * p.ldb + gc] : T(0);
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        T a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
    }
    T* C = Cbase + p.c_off;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int r = m0 + ty * 4 + i;
      if (r >= p.m) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int c = n0 + tx * 4 + j;
        if (c < p.n) C[(int64_t)r * p.ldc + c] = acc[i][j];
      }
    }
  }
}

// ------------------------------------------------------------------ peak probes ----
__global__ void k_probe_dmma(double* out, int iters) {
  double acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma884(acc[i][0], acc[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) out[0] = s;
}
__global__ void k_probe_dfma(double* out, int iters) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = i;
  double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456) out[0] = s;
}

// even warps: DMMA, odd warps: DFMA (8x the iterations = same FMA count per warp).  Tells whether the
// tensor DMMA path and the FP64 FMA pipe are separate execution resources on sm_100a.
__global__ void k_probe_mixed(double* out, int iters) {
  const int warp = threadIdx.x >> 5;
  double s = 0;
  if (warp & 1) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = i;
    double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
    for (int it = 0; it < iters * 8; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
  } else {
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
  }
  if (s == 123.456) out[0] = s;
}

using CfgA = F64Cfg<64, 64, 32, 32, 4>;       // tile_cfg 0 / 40: 128 thr, 75.8 KB smem, 2 CTAs/SM
using CfgB = F64Cfg<128, 128, 64, 32, 4>;     // tile_cfg 1: 256 thr, 149.5 KB smem
using CfgC = F64Cfg<128, 64, 64, 32, 4>;      // tile_cfg 2: 128 thr, 116.7 KB smem
using CfgD = F64Cfg<64, 64, 32, 32, 3>;       // tile_cfg 3: 128 thr, 56.8 KB smem, 3 CTAs/SM (default)
using CfgS = F64Cfg<32, 32, 16, 16, 3>;       // tile_cfg 8: 32x32 tiles, 4 warps: fine-grained units for sharded runs
using CfgL = F64Cfg<64, 128, 32, 64, 3>;      // tile_cfg 42: cuBLAS's large-GEMM shape, 128 thr, 82.9 KB

static unsigned long long* g_trace = nullptr;   // t10_ggemm_set_trace
unsigned long long* ggemm_trace_ptr() { return g_trace; }   // ggemm_tma.cu

// launch == false: only report the number of resident CTA slots (sm_count x CTAs/SM)
template <class Cfg, int MIN_CTAS, bool RAGGED = true>
static int launch_f64(bool launch, int* nslots_out, const double* A, const double* B, double* C,
                      const t10_gemm_problem* prob, int64_t ntiles, const int32_t* tiles, int64_t nslots,
                      const int32_t* slot_start, int32_t* sched, cudaStream_t st, const PeerOut& peers) {
  auto kern = k_ggemm_f64<Cfg, MIN_CTAS, RAGGED>;
  constexpr int PEER_TILE_BYTES = Cfg::BM * (Cfg::BN + 2) * 8;   // staging tile of the fused all-gather
  constexpr bool PEER_FITS = Cfg::SMEM_BYTES + PEER_TILE_BYTES <= 227 * 1024;
  T10_REQUIRE(peers.n == 0 || PEER_FITS, T10_ERR_UNSUPPORTED, "fused all-gather: tile configuration too large");
  static int ctas_by_dev[T10_MAX_DEVICES] = {0};
  int& ctas_per_sm = ctas_by_dev[cur_device()];
  if (ctas_per_sm == 0) {
    T10_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM_BYTES + (PEER_FITS ? PEER_TILE_BYTES : 0)));
    int occ = 0;
    T10_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::THREADS, Cfg::SMEM_BYTES));
    T10_REQUIRE(occ >= 1, T10_ERR_CUDA, "ggemm kernel does not fit on an SM");
    ctas_per_sm = occ;
  }
  if (!launch) {
    *nslots_out = sm_count() * ctas_per_sm;
    return T10_OK;
  }
  unsigned grid;
  int nbins = 0;
  if (slot_start != nullptr && sched != nullptr) {          // SM bins: every resident CTA slot, bins = nslots
    T10_REQUIRE(nslots >= 1, T10_ERR_INVALID, "bin table without bins");
    nbins = (int)nslots;
    grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)sm_count() * ctas_per_sm);
  } else if (slot_start != nullptr) {
    T10_REQUIRE(nslots >= 1, T10_ERR_INVALID, "slot table without slots");
    grid = (unsigned)nslots;
  } else {
    grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)sm_count() * ctas_per_sm);
  }
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES + (peers.n > 0 ? PEER_TILE_BYTES : 0), st>>>(
      A, B, C, prob, ntiles, (const int4*)tiles, slot_start, sched, nbins, g_trace, peers);
  T10_LAUNCH_CHECK();
  return T10_OK;
}

template <class Cfg, int MIN_CTAS>
static int launch_f64_v5(bool launch, int* nslots_out, const double* A, const double* B, double* C,
                         const t10_gemm_problem* prob, int64_t ntiles, const int32_t* tiles, int64_t nslots,
                         const int32_t* slot_start, int32_t* /*sched*/, cudaStream_t st, const PeerOut& peers) {
  auto kern = k_ggemm_f64_v5<Cfg, MIN_CTAS>;
  constexpr int PEER_TILE_BYTES = Cfg::BM * (Cfg::BN + 2) * 8;   // staging tile of the fused all-gather
  constexpr bool PEER_FITS = Cfg::SMEM_BYTES + PEER_TILE_BYTES <= 227 * 1024;
  T10_REQUIRE(peers.n == 0 || PEER_FITS, T10_ERR_UNSUPPORTED, "fused all-gather: tile configuration too large");
  static int ctas_by_dev[T10_MAX_DEVICES] = {0};
  int& ctas_per_sm = ctas_by_dev[cur_device()];
  if (ctas_per_sm == 0) {
    T10_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::SMEM_BYTES + (PEER_FITS ? PEER_TILE_BYTES : 0)));
    int occ = 0;
    T10_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::THREADS, Cfg::SMEM_BYTES));
    T10_REQUIRE(occ >= 1, T10_ERR_CUDA, "ggemm kernel does not fit on an SM");
    ctas_per_sm = occ;
  }
  if (!launch) {
    *nslots_out = sm_count() * ctas_per_sm;
    return T10_OK;
  }
  T10_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, T10_ERR_INVALID,
              "aligned ggemm kernel: operand base pointers must be 16-byte aligned");
  unsigned grid;
  if (slot_start != nullptr) {
    T10_REQUIRE(nslots >= 1, T10_ERR_INVALID, "slot table without slots");
    grid = (unsigned)nslots;
  } else {
    grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)sm_count() * ctas_per_sm);
  }
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES + (peers.n > 0 ? PEER_TILE_BYTES : 0), st>>>(
      A, B, C, prob, ntiles, (const int4*)tiles, slot_start, g_trace, peers);
  T10_LAUNCH_CHECK();
  return T10_OK;
}

static int launch_f64_mixed(bool launch, int* nslots_out, const double* A, const double* B, double* C,
                            const t10_gemm_problem* prob, int64_t ntiles, const int32_t* tiles, int64_t nslots,
                            const int32_t* slot_start, int32_t* /*sched*/, cudaStream_t st, const PeerOut& peers) {
  auto kern = k_ggemm_f64_mixed<2>;
  static int ctas_by_dev[T10_MAX_DEVICES] = {0};
  int& ctas_per_sm = ctas_by_dev[cur_device()];
  if (ctas_per_sm == 0) {
    T10_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MIX_SMEM_BYTES));
    int occ = 0;
    T10_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, MIX_SMEM_BYTES));
    T10_REQUIRE(occ >= 1, T10_ERR_CUDA, "ggemm kernel does not fit on an SM");
    ctas_per_sm = occ;
  }
  if (!launch) {
    *nslots_out = sm_count() * ctas_per_sm;
    return T10_OK;
  }
  T10_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, T10_ERR_INVALID,
              "aligned ggemm kernel: operand base pointers must be 16-byte aligned");
  unsigned grid;
  if (slot_start != nullptr) {
    T10_REQUIRE(nslots >= 1, T10_ERR_INVALID, "slot table without slots");
    grid = (unsigned)nslots;
  } else {
    grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)sm_count() * ctas_per_sm);
  }
  // the all-gather's staging tile aliases the pipeline stages (v5_tile, cs_off < 0): same footprint either way
  kern<<<grid, 128, MIX_SMEM_BYTES, st>>>(A, B, C, prob, ntiles, (const int4*)tiles, slot_start, g_trace, peers);
  T10_LAUNCH_CHECK();
  return T10_OK;
}

static int dispatch_f64(int tile_cfg, bool launch, int* nslots_out, const double* A, const double* B, double* C,
                        const t10_gemm_problem* prob, int64_t ntiles, const int32_t* tiles, int64_t nslots,
                        const int32_t* slot_start, int32_t* sched, cudaStream_t st,
                        const PeerOut& peers = PeerOut{}) {
  if (peers.n > 0) T10_REQUIRE((tile_cfg >= 0 && tile_cfg <= 3) || tile_cfg == 8 || tile_cfg == 40 || tile_cfg == 42 || tile_cfg == 46, T10_ERR_INVALID,
                             "fused all-gather needs an fp64 tensor-core kernel (0..3, 8, 40, 42, 46)");
  switch (tile_cfg) {
    case 0: return launch_f64<CfgA, 2>(launch, nslots_out, A, B, C, prob, ntiles, tiles, nslots, slot_start, sched, st, peers);
    case 1: return launch_f64<CfgB, 1>(launch, nslots_out, A, B, C, prob, ntiles, tiles, nslots, slot_start, sched, st, peers);
    case 2: return launch_f64<CfgC, 1>(launch, nslots_out, A, B, C, prob, ntiles, tiles, nslots, slot_start, sched, st, peers);
    case 3: return launch_f64<CfgD, 3>(launch, nslots_out, A, B, C, prob, ntiles, tiles, nslots, slot_start, sched, st, peers);
    case 8: return launch_f64<CfgS, 4>(launch, nslots_out, A, B, C, prob, ntiles, tiles, nslots, slot_start, sched, st, peers);
    case 40: return launch_f64_v5<CfgA, 2>(launch, nslots_out, A, B, C, prob, ntiles, tiles, nslots, slot_start, sched, st, peers);
    case 42: return launch_f64_v5<CfgL, 2>(launch, nslots_out, A, B, C, prob, ntiles, tiles, nslots, slot_start, sched, st, peers);
    case 46: return launch_f64_mixed(launch, nslots_out, A, B, C, prob, ntiles, tiles, nslots, slot_start, sched, st, peers);
    default: T10_REQUIRE(false, T10_ERR_INVALID, "unknown tile_cfg %d", tile_cfg);
  }
}

// ggemm_tf32.cu
int launch_ggemm_tf32x3(const float* A, const float* B, float* C, const t10_gemm_problem* prob, int64_t ntiles,
                        const int32_t* tiles, cudaStream_t st);

static bool tile_dims(int cfg, int* bm, int* bn, int* bk) {
  *bk = 16;
  switch (cfg) {
    case 50: *bm = 128; *bn = 128; *bk = 32; return true;
    case 0: case 3: case 40: case 46: case 100: *bm = 64; *bn = 64; return true;
    case 8: *bm = 32; *bn = 32; return true;
    case 1: *bm = 128; *bn = 128; return true;
    case 2: *bm = 128; *bn = 64; return true;
    case 42: *bm = 64; *bn = 128; return true;
    default: return false;
  }
}

}  // namespace t10

using namespace t10;

extern "C" int t10_ggemm_tile_shape(int tile_cfg, int* bm, int* bn) {
  int bk;
  T10_REQUIRE(tile_dims(tile_cfg, bm, bn, &bk), T10_ERR_INVALID, "unknown tile_cfg %d", tile_cfg);
  return T10_OK;
}

extern "C" int t10_ggemm_slots(int dtype, int tile_cfg, int* nslots) {
  if (dtype == T10_F64 && tile_cfg != 100)
    return dispatch_f64(tile_cfg, false, nslots, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0, nullptr, nullptr, 0);
  *nslots = sm_count() * 4;
  return T10_OK;
}

extern "C" int t10_ggemm_plan_tiles(int tile_cfg, int64_t nprob, const t10_gemm_problem* h_prob,
                                    int64_t nslots, int64_t cap, int32_t* h_tiles,
                                    int32_t* h_slot_start, int64_t* ntiles) {
  int bm, bn, bk;
  T10_REQUIRE(tile_dims(tile_cfg, &bm, &bn, &bk), T10_ERR_INVALID, "unknown tile_cfg %d", tile_cfg);
  struct Tl { int32_t p, m0, n0; int64_t cost; int32_t shape; };
  const int64_t fixed_steps = 2;
  const char* al = getenv("TOR10_GGEMM_ALPHA");
  const int64_t alpha = al ? atoll(al) : 75;     // measured on C3a: 0 -> 118.1 us, 50 -> 120.3, 75 -> 112.3, 100 -> 114.4
  const bool warp_granular = tile_cfg == 40 || tile_cfg == 42 || tile_cfg == 46;
  int wmd = bm, wnd = bn;          // warp tile of the kernel
  switch (tile_cfg) {
    case 0: case 3: case 40: case 46: wmd = 32; wnd = 32; break;
    case 1: case 2: wmd = 64; wnd = 32; break;
    case 8: wmd = 16; wnd = 16; break;
    case 42: wmd = 32; wnd = 64; break;
    default: break;
  }
  std::vector<Tl> v;
  for (int64_t i = 0; i < nprob; ++i) {
    const t10_gemm_problem& p = h_prob[i];
    T10_REQUIRE(p.m >= 0 && p.n >= 0 && p.k >= 0, T10_ERR_INVALID, "negative GEMM extent");
    if (p.m == 0 || p.n == 0) continue;
    const int64_t kt = (p.k + bk - 1) / bk;
    // cost of one tile of shape tbm x tbn holding mm x nn valid elements.  Each warp of a CTA sits on its own
    // scheduler, so a k-step lasts as long as its BUSIEST warp's DMMAs (x16 cycles), not their sum: a ragged
    // tile with one full warp tile costs what a full tile costs.  cost = k-steps x (alpha * busiest warp's
    // fragments * warps + (1 - alpha) * all valid fragments) + a fixed prologue/epilogue term.  The general
    // kernels skip invalid 8x8 fragments; the aligned kernels (40, 42, 46) skip whole warp tiles only.
    auto tile_cost = [&](int tbm, int tbn, int twm, int twn, int64_t mm, int64_t nn) {
      int64_t wmax = 0, frags = 0;
      for (int wm0 = 0; wm0 < tbm; wm0 += twm)
        for (int wn0 = 0; wn0 < tbn; wn0 += twn) {
          int64_t mi = std::max<int64_t>(0, std::min<int64_t>(twm / 8, (mm - wm0 + 7) / 8));
          int64_t ni = std::max<int64_t>(0, std::min<int64_t>(twn / 8, (nn - wn0 + 7) / 8));
          if (warp_granular && mi > 0 && ni > 0) { mi = twm / 8; ni = twn / 8; }
          wmax = std::max(wmax, mi * ni);
          frags += mi * ni;
        }
      const int64_t nwarps = (tbm / twm) * (tbn / twn);
      const int64_t per_step = (alpha * wmax * nwarps + (100 - alpha) * frags) / 100;
      return kt * (bk / 4) * per_step + fixed_steps * (bk / 4) * (tbm / 8) * (tbn / 8);
    };
    if (tile_cfg == 46) {
      // mixed shapes: 64x64 tiles; when at most 32 rows (columns) are left over they become 32x128 (128x32)
      // strips; the corner belongs to the bottom strip
      const int rm = p.m % 64, rn = p.n % 64;
      const bool sm = rm > 0 && rm <= 32, sn = rn > 0 && rn <= 32;
      const int m_sq = sm ? p.m - rm : p.m, n_sq = sn ? p.n - rn : p.n;    // region of the square tiles
      auto enc = [](int shape, int64_t mm, int64_t nn) { return (int32_t)(shape | ((int)mm << 8) | ((int)nn << 20)); };
      for (int m0 = 0; m0 < m_sq; m0 += 64)
        for (int n0 = 0; n0 < n_sq; n0 += 64) {
          const int64_t mm = std::min(64, m_sq - m0), nn = std::min(64, n_sq - n0);
          v.push_back({(int32_t)i, m0, n0, tile_cost(64, 64, 32, 32, mm, nn), enc(0, mm, nn)});
        }
      if (sm)      // bottom strip, corner included
        for (int n0 = 0; n0 < p.n; n0 += 128) {
          const int64_t nn = std::min(128, p.n - n0);
          v.push_back({(int32_t)i, m_sq, n0, tile_cost(32, 128, 32, 32, rm, nn), enc(1, rm, nn)});
        }
      if (sn)      // right strip, rows of the square region only
        for (int m0 = 0; m0 < m_sq; m0 += 128) {
          const int64_t mm = std::min(128, m_sq - m0);
          v.push_back({(int32_t)i, m0, n_sq, tile_cost(128, 32, 32, 32, mm, rn), enc(2, mm, rn)});
        }
      continue;
    }
    for (int m0 = 0; m0 < p.m; m0 += bm)
      for (int n0 = 0; n0 < p.n; n0 += bn) {
        int64_t mm = std::min(bm, p.m - m0), nn = std::min(bn, p.n - n0);
        v.push_back({(int32_t)i, m0, n0, tile_cost(bm, bn, wmd, wnd, mm, nn), 0});
      }
  }
  std::stable_sort(v.begin(), v.end(), [](const Tl& a, const Tl& b) { return a.cost > b.cost; });
  *ntiles = (int64_t)v.size();
  if (!h_tiles) return T10_OK;
  T10_REQUIRE(cap >= (int64_t)v.size(), T10_ERR_INVALID, "tile buffer too small");
  auto put = [&](size_t dst, const Tl& t) {
    h_tiles[4 * dst + 0] = t.p;
    h_tiles[4 * dst + 1] = t.m0;
    h_tiles[4 * dst + 2] = t.n0;
    h_tiles[4 * dst + 3] = t.shape;
  };
  if (nslots <= 0 || h_slot_start == nullptr) {
    for (size_t i = 0; i < v.size(); ++i) put(i, v[i]);
    return T10_OK;
  }
  // LPT: heaviest tile to the least-loaded slot; tiles are then stored slot-major
  std::vector<int64_t> load(nslots, 0);
  std::vector<std::vector<int>> members(nslots);
  // min-heap over (load, slot)
  std::vector<std::pair<int64_t, int>> heap;
  for (int s2 = 0; s2 < nslots; ++s2) heap.push_back({0, s2});
  auto cmp = [](const std::pair<int64_t, int>& a, const std::pair<int64_t, int>& b) { return a > b; };
  std::make_heap(heap.begin(), heap.end(), cmp);
  for (size_t i = 0; i < v.size(); ++i) {
    std::pop_heap(heap.begin(), heap.end(), cmp);
    auto& top = heap.back();
    members[top.second].push_back((int)i);
    top.first += v[i].cost;
    std::push_heap(heap.begin(), heap.end(), cmp);
  }
  size_t pos = 0;
  for (int s2 = 0; s2 < nslots; ++s2) {
    h_slot_start[s2] = (int32_t)pos;
    for (int idx : members[s2]) put(pos++, v[idx]);
  }
  h_slot_start[nslots] = (int32_t)pos;
  return T10_OK;
}

extern "C" int t10_ggemm(int dtype, int tile_cfg, const void* d_A, const void* d_B, void* d_C,
                         int64_t nprob, const t10_gemm_problem* d_prob, int64_t ntiles,
                         const int32_t* d_tiles, int64_t nslots, const int32_t* d_slot_start,
                         int32_t* d_sched, t10_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (ntiles == 0 || nprob == 0) return T10_OK;
  T10_REQUIRE(d_sched == nullptr || d_slot_start == nullptr || (dtype == T10_F64 && ((tile_cfg >= 0 && tile_cfg <= 3) || tile_cfg == 8)),
              T10_ERR_INVALID, "ggemm: SM-bin schedule (counters + bin table) needs the fp64 kernels 0..3");
  T10_REQUIRE(ntiles > 0 && nprob > 0, T10_ERR_INVALID, "negative counts");
  if (dtype == T10_F64 && tile_cfg != 100)
    return dispatch_f64(tile_cfg, true, nullptr, (const double*)d_A, (const double*)d_B, (double*)d_C, d_prob,
                        ntiles, d_tiles, nslots, d_slot_start, d_sched, st);
  // SIMT kernels walk the flat tile list (slot tables are ignored: tiles are stored slot-major, any order works)
  unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)sm_count() * 4);
  if (dtype == T10_F64) {  // cfg 100: cross-check path used by the tests
    k_ggemm_simt<double><<<grid, 256, 0, st>>>((const double*)d_A, (const double*)d_B, (double*)d_C, d_prob,
                                               ntiles, (const int4*)d_tiles);
  } else if (dtype == T10_F32 && tile_cfg == 50) {
    return launch_ggemm_tf32x3((const float*)d_A, (const float*)d_B, (float*)d_C, d_prob, ntiles, d_tiles, st);
  } else if (dtype == T10_F32) {
    T10_REQUIRE(tile_cfg == 0 || tile_cfg == 100, T10_ERR_INVALID, "fp32 ggemm: cfg 0 (SIMT) or 50 (tcgen05)");
    k_ggemm_simt<float><<<grid, 256, 0, st>>>((const float*)d_A, (const float*)d_B, (float*)d_C, d_prob,
                                              ntiles, (const int4*)d_tiles);
  } else {
    T10_REQUIRE(false, T10_ERR_UNSUPPORTED, "ggemm: dtype %d not supported", dtype);
  }
  T10_LAUNCH_CHECK();
  return T10_OK;
}

__global__ void k_raise_flags(PeerOut peers) {   // a rank without tiles still has to signal
  if (threadIdx.x < peers.n && peers.flag[threadIdx.x] != nullptr)
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers.flag[threadIdx.x]), "l"(peers.step) : "memory");
}

extern "C" int t10_ggemm_allgather(int tile_cfg, const void* d_A, const void* d_B, int npeers, int self_index,
                                   void* const* h_peer_C, void* const* h_peer_flag, uint64_t step,
                                   int32_t* d_done_ctr, int64_t nprob, const t10_gemm_problem* d_prob,
                                   int64_t ntiles, const int32_t* d_tiles, int64_t nslots,
                                   const int32_t* d_slot_start, int32_t* d_sched, t10_stream_t stream) {
  T10_REQUIRE(npeers >= 1 && npeers <= T10_MAX_PEERS, T10_ERR_INVALID, "ggemm_allgather: 1..%d arenas", T10_MAX_PEERS);
  T10_REQUIRE(ntiles >= 0 && nprob >= 0, T10_ERR_INVALID, "negative counts");
  T10_REQUIRE(h_peer_flag == nullptr || d_done_ctr != nullptr, T10_ERR_INVALID, "flags need an arrival counter");
  PeerOut po;
  memset(&po, 0, sizeof(po));
  po.n = npeers;
  po.self = (self_index >= 0 && self_index < npeers) ? self_index : -1;
  po.step = step;
  po.done_ctr = d_done_ctr;
  for (int i = 0; i < npeers; ++i) {
    po.ptr[i] = (double*)h_peer_C[i];
    po.flag[i] = h_peer_flag ? (unsigned long long*)h_peer_flag[i] : nullptr;
  }
  if (ntiles == 0 || nprob == 0) {
    if (h_peer_flag) {
      k_raise_flags<<<1, 32, 0, (cudaStream_t)stream>>>(po);
      T10_LAUNCH_CHECK();
    }
    return T10_OK;
  }
  return dispatch_f64(tile_cfg, true, nullptr, (const double*)d_A, (const double*)d_B, (double*)h_peer_C[0], d_prob,
                      ntiles, d_tiles, nslots, d_slot_start, d_sched, (cudaStream_t)stream, po);
}

static int launch_streamk(const void* d_A, const void* d_B, void* d_C, int64_t nprob, const t10_gemm_problem* d_prob,
                          int64_t npieces, const int32_t* d_tiles, const int32_t* d_pieces, int64_t nslots,
                          const int32_t* d_slot_start, void* d_ws, int32_t* d_flags, int32_t epoch,
                          const PeerOut& po, cudaStream_t stream) {
  T10_REQUIRE(npieces > 0 && nprob > 0 && nslots >= 1, T10_ERR_INVALID, "ggemm_streamk: bad counts");
  T10_REQUIRE(d_pieces && d_tiles && d_slot_start && d_ws && d_flags, T10_ERR_INVALID, "ggemm_streamk: missing table");
  T10_REQUIRE(((uintptr_t)d_A & 15) == 0 && ((uintptr_t)d_B & 15) == 0, T10_ERR_INVALID,
              "ggemm_streamk: operand base pointers must be 16-byte aligned");
  auto kern = k_ggemm_f64_mixed_sk<2>;
  static int resident_by_dev[T10_MAX_DEVICES] = {0};
  int& resident = resident_by_dev[cur_device()];
  if (resident == 0) {
    T10_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MIX_SMEM_BYTES));
    int occ = 0;
    T10_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, MIX_SMEM_BYTES));
    T10_REQUIRE(occ >= 1, T10_ERR_CUDA, "ggemm kernel does not fit on an SM");
    resident = occ * sm_count();
  }
  // a waiting piece spins on flags raised by other CTAs of the same launch: all of them must be co-resident
  T10_REQUIRE(nslots <= resident, T10_ERR_INVALID, "ggemm_streamk: %lld slots but only %d resident CTAs",
              (long long)nslots, resident);
  kern<<<(unsigned)nslots, 128, MIX_SMEM_BYTES, stream>>>(
      (const double*)d_A, (const double*)d_B, (double*)d_C, d_prob, (const int4*)d_tiles, (const int4*)d_pieces,
      d_slot_start, (double*)d_ws, d_flags, epoch, g_trace, po);
  T10_LAUNCH_CHECK();
  return T10_OK;
}

extern "C" int t10_ggemm_streamk(const void* d_A, const void* d_B, void* d_C, int64_t nprob,
                                 const t10_gemm_problem* d_prob, int64_t npieces, const int32_t* d_tiles,
                                 const int32_t* d_pieces, int64_t nslots, const int32_t* d_slot_start,
                                 void* d_ws, int32_t* d_flags, int32_t epoch, t10_stream_t stream) {
  if (npieces == 0 || nprob == 0) return T10_OK;
  PeerOut po;
  memset(&po, 0, sizeof(po));
  po.self = -1;
  return launch_streamk(d_A, d_B, d_C, nprob, d_prob, npieces, d_tiles, d_pieces, nslots, d_slot_start, d_ws, d_flags,
                        epoch, po, (cudaStream_t)stream);
}

extern "C" int t10_ggemm_streamk_allgather(const void* d_A, const void* d_B, int npeers, int self_index,
                                           void* const* h_peer_C,
                                           void* const* h_peer_flag, uint64_t step, int32_t* d_done_ctr,
                                           int64_t nprob, const t10_gemm_problem* d_prob, int64_t npieces,
                                           const int32_t* d_tiles, const int32_t* d_pieces, int64_t nslots,
                                           const int32_t* d_slot_start, void* d_ws, int32_t* d_flags, int32_t epoch,
                                           t10_stream_t stream) {
  T10_REQUIRE(npeers >= 1 && npeers <= T10_MAX_PEERS, T10_ERR_INVALID, "ggemm_streamk_allgather: 1..%d arenas",
              T10_MAX_PEERS);
  T10_REQUIRE(npieces >= 0 && nprob >= 0, T10_ERR_INVALID, "negative counts");
  T10_REQUIRE(h_peer_flag == nullptr || d_done_ctr != nullptr, T10_ERR_INVALID, "flags need an arrival counter");
  PeerOut po;
  memset(&po, 0, sizeof(po));
  po.n = npeers;
  po.self = (self_index >= 0 && self_index < npeers) ? self_index : -1;
  po.step = step;
  po.done_ctr = d_done_ctr;
  for (int i = 0; i < npeers; ++i) {
    po.ptr[i] = (double*)h_peer_C[i];
    po.flag[i] = h_peer_flag ? (unsigned long long*)h_peer_flag[i] : nullptr;
  }
  if (npieces == 0 || nprob == 0) {
    if (h_peer_flag) {
      k_raise_flags<<<1, 32, 0, (cudaStream_t)stream>>>(po);
      T10_LAUNCH_CHECK();
    }
    return T10_OK;
  }
  return launch_streamk(d_A, d_B, h_peer_C[0], nprob, d_prob, npieces, d_tiles, d_pieces, nslots, d_slot_start, d_ws,
                        d_flags, epoch, po, (cudaStream_t)stream);
}

extern "C" int t10_ggemm_set_trace(void* d_trace) {
  t10::g_trace = (unsigned long long*)d_trace;
  return T10_OK;
}

extern "C" int t10_probe_f64_occupancy(int warps_per_scheduler, int iters, double* tflops, t10_stream_t stream) {
  // DMMA-only probe with exactly `warps_per_scheduler` resident warps on every SM sub-partition
  cudaStream_t st = (cudaStream_t)stream;
  T10_REQUIRE(warps_per_scheduler >= 1 && warps_per_scheduler <= 16, T10_ERR_INVALID, "1..16 warps");
  double* d_out;
  T10_CUDA(cudaMalloc(&d_out, 8));
  cudaEvent_t e0, e1;
  T10_CUDA(cudaEventCreate(&e0));
  T10_CUDA(cudaEventCreate(&e1));
  const int blocks = sm_count(), threads = 128 * warps_per_scheduler;  // one CTA per SM
  for (int rep = 0; rep < 2; ++rep) {
    T10_CUDA(cudaEventRecord(e0, st));
    k_probe_dmma<<<blocks, threads, 0, st>>>(d_out, iters);
    T10_CUDA(cudaEventRecord(e1, st));
    T10_CUDA(cudaEventSynchronize(e1));
  }
  float ms = 0;
  T10_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  double warps = (double)blocks * threads / 32.0;
  *tflops = 2.0 * warps * (double)iters * 16.0 * 256.0 / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  return T10_OK;
}

extern "C" int t10_probe_f64_peak(int use_dmma, int iters, double* tflops, t10_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  double* d_out;
  T10_CUDA(cudaMalloc(&d_out, 8));
  cudaEvent_t e0, e1;
  T10_CUDA(cudaEventCreate(&e0));
  T10_CUDA(cudaEventCreate(&e1));
  const int blocks = sm_count() * 4, threads = 256;
  for (int rep = 0; rep < 2; ++rep) {  // first rep warms up
    T10_CUDA(cudaEventRecord(e0, st));
    if (use_dmma == 1) k_probe_dmma<<<blocks, threads, 0, st>>>(d_out, iters);
    else if (use_dmma == 2) k_probe_mixed<<<blocks, threads, 0, st>>>(d_out, iters);
    else k_probe_dfma<<<blocks, threads, 0, st>>>(d_out, iters);
    T10_CUDA(cudaEventRecord(e1, st));
    T10_CUDA(cudaEventSynchronize(e1));
  }
  float ms = 0;
  T10_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  // DMMA m8n8k4 = 256 FMA per warp instruction; DFMA = 32 FMA per warp instruction
  double warps = (double)blocks * threads / 32.0;
  double fma = warps * (double)iters * 16.0 * (use_dmma == 1 ? 256.0 : use_dmma == 2 ? 256.0 : 32.0);
  *tflops = 2.0 * fma / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  return T10_OK;
}
