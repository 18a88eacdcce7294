// Family 3, FP32 path: grouped GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), 3xTF32.
//
// C[m,n] (fp32) = A[m,k] * B[k,n] with fp32 accuracy (north-star bound 1e-5) from TF32 tensor-core MMAs:
// every operand element x is split as x = hi + lo, hi = x with the 13 low mantissa bits cleared (what the
// tensor core reads of an fp32 word anyway), lo = x - hi (exact in fp32), and the tile product is
// accumulated as  Ahi*Bhi + Ahi*Blo + Alo*Bhi  in the fp32 TMEM accumulator (the dropped lo*lo term is
// 2^-22 relative).  The tensor core adds into its fp32 accumulator with truncation (measured: about one
// ulp lost per MMA, always towards zero), so a long k loop would drift by K/8 * 3 ulps.  Two measures keep
// the result at fp32 quality for any K: the large term (hi*hi) and the 2^-11 smaller correction terms go to
// SEPARATE TMEM accumulators, and every TF_KC = 256 of k the pair is drained into round-to-nearest fp32
// registers (the drain of chunk j overlaps the MMAs of chunk j+1: the accumulators are double-buffered,
// 2 x (main + corr) x 128 columns = all 512 TMEM columns).
//
// One CTA (256 threads) per 128x128 output tile, persistent over the tile table:
//   * all threads: LDG.128 of the NEXT 32-deep k-slab into registers (A rows are K-contiguous; B is read
//     as 4k x 4n register blocks and transposed in registers), then split + STS.128 of the current slab's
//     hi and lo tiles into the canonical no-swizzle K-major UMMA layout ((8,n),2):((1,SBO),LBO) (16-byte
//     units) for BOTH operands -- kind::tf32 with an MN-major no-swizzle B returned zeros on B200
//     (tools/tc_probe.cu), so B is stored transposed.  LBO / SBO carry 16 bytes of padding so that the
//     eight lanes of every STS.128 phase hit eight different 16-byte bank groups;
//   * one thread: 12 x tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=128, K=8) per slab, then
//     tcgen05.commit onto the stage's mbarrier; two stages, so the tensor core works on slab kt while
//     slab kt+1 is split and stored and slab kt+2 is in flight from HBM/L2;
//   * drain / epilogue: warp w reads TMEM lanes 32*(w%4).. and columns 64*(w/4).. with
//     tcgen05.ld.32x32b.x16, adds into its 64 register accumulators, and at the end of the tile stores
//     32-byte row segments.
// SASS: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (commit).
#include <algorithm>
#include <cstdlib>

#include "t10_common.cuh"

namespace t10 {

constexpr int TF_BM = 128, TF_BN = 128, TF_BK = 32, TF_THREADS = 256, TF_STAGES = 2;
// byte strides of the K-major layouts: 16-byte chunk (row r, k-chunk c) at (r/8)*SBO + c*LBO + (r%8)*16
constexpr unsigned A_LBO = 144, A_SBO = (TF_BK / 4) * A_LBO;          // 1152
constexpr unsigned B_LBO = 128, B_SBO = (TF_BK / 4) * B_LBO + 16;     // 1040
constexpr int TF_A_TILE = (TF_BM / 8) * A_SBO;                        // bytes of one A (hi or lo) tile
constexpr int TF_B_TILE = (TF_BN / 8) * B_SBO;
constexpr int TF_STAGE_BYTES = 2 * TF_A_TILE + 2 * TF_B_TILE;         // Ahi, Alo, Bhi, Blo
constexpr int TF_SMEM = TF_STAGES * TF_STAGE_BYTES + 128;
constexpr int TF_A_CHUNKS = (TF_BM * TF_BK / 4) / TF_THREADS;         // 16-byte A chunks per thread per slab (4)
constexpr int TF_KC_SLABS = 8;                                        // slabs per accumulation chunk (TF_KC = 256)
constexpr unsigned TF_TMEM_COLS = 512;                                // 2 buffers x (main, corr) x 128 columns

__device__ __forceinline__ uint64_t umma_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;                 // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

__device__ __forceinline__ void umma_tf32(unsigned tmem_d, uint64_t adesc, uint64_t bdesc, unsigned idesc, unsigned accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(0u)
      : "memory");
}

// returns false if the barrier did not flip within the spin budget (never on a healthy run)
__device__ __forceinline__ bool tf_mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  for (int spin = 0; spin < (1 << 24); ++spin) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}

__device__ __forceinline__ void split_store(unsigned hi_addr, unsigned lo_addr, float4 v) {
  float4 h, l;
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
  l.x = v.x - h.x;
  l.y = v.y - h.y;
  l.z = v.z - h.z;
  l.w = v.w - h.w;
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(hi_addr), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(lo_addr), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
}

// one thread's share of a 32-deep k-slab: four 16-byte A chunks and one 4k x 4n block of B
struct TfSlab {
  float4 a[TF_A_CHUNKS];
  float4 b[4];
};

// 4 consecutive floats of a row with zero fill past `limit` columns (start < limit is NOT assumed)
__device__ __forceinline__ float4 tf_load4(const float* __restrict__ row, int start, int limit, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (start < limit) {
    const float* src = row + start;
    if (vec && start + 3 < limit) {
      v = __ldg(reinterpret_cast<const float4*>(src));
    } else {
      v.x = src[0];
      if (start + 1 < limit) v.y = src[1];
      if (start + 2 < limit) v.z = src[2];
      if (start + 3 < limit) v.w = src[3];
    }
  }
  return v;
}

__global__ void __launch_bounds__(TF_THREADS, 1)
k_ggemm_tf32x3(const float* __restrict__ Abase, const float* __restrict__ Bbase, float* __restrict__ Cbase,
               const t10_gemm_problem* __restrict__ prob, int64_t ntiles, const int4* __restrict__ tiles,
               int* __restrict__ err_flag) {
  extern __shared__ __align__(128) unsigned char tf_smem[];
  __shared__ __align__(8) uint64_t mma_bar[TF_STAGES];
  __shared__ unsigned tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned smem0 = ((unsigned)__cvta_generic_to_shared(tf_smem) + 127u) & ~127u;

  if (warp == 0) {
    unsigned dst = (unsigned)__cvta_generic_to_shared(&tmem_base_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(dst), "r"(TF_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    for (int s = 0; s < TF_STAGES; ++s) {
      unsigned a = (unsigned)__cvta_generic_to_shared(&mma_bar[s]);
      asm volatile("mbarrier.init.shared.b64 [%0], 1;\n" ::"r"(a) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const unsigned tmem_d = tmem_base_s;
  // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
  const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((TF_BN >> 3) << 17) | ((TF_BM >> 4) << 24);
  unsigned commits[TF_STAGES] = {0, 0};   // commits issued so far per stage (same value in every thread)
  bool healthy = true;

  // shared-memory offsets of this thread's chunks inside a tile (fixed for the whole kernel)
  unsigned a_off[TF_A_CHUNKS];
#pragma unroll
  for (int i = 0; i < TF_A_CHUNKS; ++i) {
    const int c = tid + i * TF_THREADS, r = c >> 3, kc = c & 7;
    a_off[i] = (r >> 3) * A_SBO + kc * A_LBO + (r & 7) * 16;
  }
  const int b_nchunk = tid & 31, b_kc = tid >> 5;          // B block: k rows 4*b_kc.., columns 4*b_nchunk..
  unsigned b_off[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = b_nchunk * 4 + j;
    b_off[j] = (n >> 3) * B_SBO + b_kc * B_LBO + (n & 7) * 16;
  }

  for (int64_t t = blockIdx.x; t < ntiles && healthy; t += gridDim.x) {
    const int4 tile = tiles[t];
    const t10_gemm_problem p = prob[tile.x];
    const int m0 = tile.y, n0 = tile.z;
    const float* A = Abase + p.a_off;
    const float* B = Bbase + p.b_off;
    const bool a_vec = ((p.lda & 3) == 0) && ((p.a_off & 3) == 0) && (((uintptr_t)Abase & 15) == 0);
    const bool b_vec = ((p.ldb & 3) == 0) && ((p.b_off & 3) == 0) && (((uintptr_t)Bbase & 15) == 0) && ((n0 & 3) == 0);
    const int KT = (p.k + TF_BK - 1) / TF_BK;

    float acc[TF_BN / 2];
#pragma unroll
    for (int i = 0; i < TF_BN / 2; ++i) acc[i] = 0.f;
    const int lane0 = (warp & 3) * 32, col0 = (warp >> 2) * (TF_BN / 2);
    // acc += main + corr of accumulator buffer `buf` (this warp's 32 lanes x 64 columns)
    auto drain = [&](int buf) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
      for (int cb = 0; cb < TF_BN / 2; cb += 16) {
        unsigned r[16], q[16];
        const unsigned taddr = tmem_d + ((unsigned)lane0 << 16) + (unsigned)(buf * 2 * TF_BN + col0 + cb);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr)
            : "memory");
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
            : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
              "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15])
            : "r"(taddr + (unsigned)TF_BN)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[cb + j] += __uint_as_float(r[j]) + __uint_as_float(q[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");   // ordered before the next CTA barrier
    };

    auto gload = [&](TfSlab& sl, int kt) {
      const int k0 = kt * TF_BK;
#pragma unroll
      for (int i = 0; i < TF_A_CHUNKS; ++i) {
        const int c = tid + i * TF_THREADS;
        const int gr = m0 + (c >> 3), gk = k0 + (c & 7) * 4;
        sl.a[i] = gr < p.m ? tf_load4(A + (int64_t)gr * p.lda, gk, p.k, a_vec) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gk = k0 + b_kc * 4 + j;
        sl.b[j] = gk < p.k ? tf_load4(B + (int64_t)gk * p.ldb, n0 + b_nchunk * 4, p.n, b_vec)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto sstore = [&](const TfSlab& sl, unsigned st_base) {
#pragma unroll
      for (int i = 0; i < TF_A_CHUNKS; ++i) split_store(st_base + a_off[i], st_base + TF_A_TILE + a_off[i], sl.a[i]);
      const unsigned bh = st_base + 2 * TF_A_TILE, bl = bh + TF_B_TILE;
      // register transpose of the 4k x 4n block: column j of the block is one K-major chunk
      split_store(bh + b_off[0], bl + b_off[0], make_float4(sl.b[0].x, sl.b[1].x, sl.b[2].x, sl.b[3].x));
      split_store(bh + b_off[1], bl + b_off[1], make_float4(sl.b[0].y, sl.b[1].y, sl.b[2].y, sl.b[3].y));
      split_store(bh + b_off[2], bl + b_off[2], make_float4(sl.b[0].z, sl.b[1].z, sl.b[2].z, sl.b[3].z));
      split_store(bh + b_off[3], bl + b_off[3], make_float4(sl.b[0].w, sl.b[1].w, sl.b[2].w, sl.b[3].w));
    };
    // one k-slab: prefetch slab kt+1 into `nxt`, store `cur` into stage kt&1, issue its MMAs
    auto step = [&](TfSlab& cur, TfSlab& nxt, int kt) -> bool {
      const int s = kt & 1;
      if (kt + 1 < KT) gload(nxt, kt + 1);
      if (commits[s] > 0) {   // the MMAs that read this stage last time must have finished
        if (!tf_mbar_wait(&mma_bar[s], (commits[s] - 1) & 1)) return false;
      }
      const unsigned st_base = smem0 + s * TF_STAGE_BYTES;
      sstore(cur, st_base);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy stores -> tensor-core reads
      __syncthreads();
      const int chunk = kt / TF_KC_SLABS, in_chunk = kt % TF_KC_SLABS;
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const unsigned ahi = st_base, alo = st_base + TF_A_TILE;
        const unsigned bhi = st_base + 2 * TF_A_TILE, blo = bhi + TF_B_TILE;
        const unsigned d_main = tmem_d + (unsigned)((chunk & 1) * 2 * TF_BN), d_corr = d_main + (unsigned)TF_BN;
#pragma unroll
        for (int kg = 0; kg < TF_BK / 8; ++kg) {      // one MMA consumes K = 8 = two k-chunks of each operand
          const uint64_t dah = umma_desc(ahi + kg * 2 * A_LBO, A_LBO, A_SBO);
          const uint64_t dal = umma_desc(alo + kg * 2 * A_LBO, A_LBO, A_SBO);
          const uint64_t dbh = umma_desc(bhi + kg * 2 * B_LBO, B_LBO, B_SBO);
          const uint64_t dbl = umma_desc(blo + kg * 2 * B_LBO, B_LBO, B_SBO);
          const unsigned first = (in_chunk > 0 || kg > 0) ? 1u : 0u;
          umma_tf32(d_main, dah, dbh, idesc, first);
          umma_tf32(d_corr, dah, dbl, idesc, first);
          umma_tf32(d_corr, dal, dbh, idesc, 1u);
        }
        unsigned bar = (unsigned)__cvta_generic_to_shared(&mma_bar[s]);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
      }
      commits[s] += 1;
      if (in_chunk == 0 && chunk > 0) {
        // the previous chunk's MMAs retire before this slab's start (in-order pipe): drain it while they run
        const int ps = (kt - 1) & 1;
        if (!tf_mbar_wait(&mma_bar[ps], (commits[ps] - 1) & 1)) return false;
        drain((chunk - 1) & 1);
      }
      return true;
    };

    TfSlab s0, s1;
    if (KT > 0) gload(s0, 0);
    for (int kt = 0; kt < KT && healthy; kt += 2) {
      healthy = step(s0, s1, kt);
      if (healthy && kt + 1 < KT) healthy = step(s1, s0, kt + 1);
    }
    if (!healthy) break;
    // all MMAs of the tile done: the last commit covers every earlier MMA of the issuing thread
    if (KT > 0) {
      const int s = (KT - 1) & 1;
      if (!tf_mbar_wait(&mma_bar[s], (commits[s] - 1) & 1)) { healthy = false; break; }
    }
    if (KT > 0) drain(((KT - 1) / TF_KC_SLABS) & 1);
    // ---- epilogue: this thread holds row m0 + lane0 + lane, columns n0 + col0 .. +63
    const int row = m0 + lane0 + lane;
    float* crow = Cbase + p.c_off + (int64_t)row * p.ldc;
    const bool c_vec = ((p.ldc & 3) == 0) && ((p.c_off & 3) == 0) && (((uintptr_t)Cbase & 15) == 0) && ((n0 & 3) == 0);
    if (row < p.m) {
#pragma unroll
      for (int cb = 0; cb < TF_BN / 2; cb += 4) {
        const int gn = n0 + col0 + cb;
        if (gn < p.n) {
          if (c_vec && gn + 3 < p.n) {
            *reinterpret_cast<float4*>(crow + gn) = make_float4(acc[cb], acc[cb + 1], acc[cb + 2], acc[cb + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (gn + j < p.n) crow[gn + j] = acc[cb + j];
          }
        }
      }
    }
    // TMEM is reused by the next tile: every warp's loads must have retired first (drain fenced them)
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  }
  if (!healthy && tid == 0) atomicExch(err_flag, 1);
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"(TF_TMEM_COLS) : "memory");
  }
}

static int* d_err_by_dev[T10_MAX_DEVICES] = {nullptr};    // error words, one per device

int launch_ggemm_tf32x3(const float* A, const float* B, float* C, const t10_gemm_problem* prob, int64_t ntiles,
                        const int32_t* tiles, cudaStream_t st) {
  int*& d_err = d_err_by_dev[cur_device()];
  if (d_err == nullptr) {
    T10_CUDA(cudaFuncSetAttribute(k_ggemm_tf32x3, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM));
    T10_CUDA(cudaMalloc(&d_err, 4 * sizeof(int)));
    T10_CUDA(cudaMemset(d_err, 0, 4 * sizeof(int)));
  }
  unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)sm_count());
  k_ggemm_tf32x3<<<grid, TF_THREADS, TF_SMEM, st>>>(A, B, C, prob, ntiles, (const int4*)tiles, d_err);
  T10_LAUNCH_CHECK();
  return T10_OK;
}

}  // namespace t10

// 1 if any CTA of any tcgen05 launch so far gave up waiting on an MMA barrier (never on a healthy run)
extern "C" int t10_ggemm_tf32_status(int* h_flag) {
  *h_flag = 0;
  int* d_err = t10::d_err_by_dev[t10::cur_device()];
  if (d_err == nullptr) return T10_OK;
  T10_CUDA(cudaMemcpy(h_flag, d_err, sizeof(int), cudaMemcpyDeviceToHost));
  return T10_OK;
}
