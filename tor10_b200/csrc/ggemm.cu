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
// FP32: SIMT register-tiled kernel (fp32 FMA keeps the 1e-5 bound of the north star; a
// tcgen05 kind::tf32 3-pass split is the planned replacement).
#include <algorithm>
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
template <int BM_, int BN_, int WM_, int WN_, int STAGES_>
struct F64Cfg {
  static constexpr int BM = BM_, BN = BN_, BK = 16, WM = WM_, WN = WN_, STAGES = STAGES_;
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

template <class Cfg, int MIN_CTAS>
__global__ void __launch_bounds__(Cfg::THREADS, MIN_CTAS)
k_ggemm_f64(const double* __restrict__ Abase, const double* __restrict__ Bbase,
            double* __restrict__ Cbase, const t10_gemm_problem* __restrict__ prob,
            int64_t ntiles, const int4* __restrict__ tiles) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + Cfg::STAGES * Cfg::A_STAGE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int wm0 = (warp / Cfg::WARPS_N) * Cfg::WM, wn0 = (warp % Cfg::WARPS_N) * Cfg::WN;

  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int4 tile = tiles[t];
    const t10_gemm_problem p = prob[tile.x];
    const int m0 = tile.y, n0 = tile.z;
    const double* A = Abase + p.a_off + (int64_t)m0 * p.lda;
    const double* B = Bbase + p.b_off + n0;
    const int m_rem = p.m - m0, n_rem = p.n - n0;
    const bool a16 = ((p.lda & 1) == 0) && ((p.a_off & 1) == 0) && (((uintptr_t)Abase & 15) == 0);
    const bool b16 = ((p.ldb & 1) == 0) && (((p.b_off + n0) & 1) == 0) && (((uintptr_t)Bbase & 15) == 0);
    const int KT = (p.k + Cfg::BK - 1) / Cfg::BK;

    double acc[Cfg::MI][Cfg::NI][2];
#pragma unroll
    for (int i = 0; i < Cfg::MI; ++i)
#pragma unroll
      for (int j = 0; j < Cfg::NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    __syncthreads();  // all warps done reading the previous tile's stages
#pragma unroll
    for (int s = 0; s < Cfg::STAGES - 1; ++s) {
      if (s < KT)
        load_stage_f64<Cfg>(As + s * Cfg::A_STAGE, Bs + s * Cfg::B_STAGE, A + s * Cfg::BK,
                            B + (int64_t)s * Cfg::BK * p.ldb, p.lda, p.ldb, m_rem, n_rem,
                            p.k - s * Cfg::BK, a16, b16);
      cp_async_commit();
    }
    for (int kt = 0; kt < KT; ++kt) {
      cp_async_wait<Cfg::STAGES - 2>();
      __syncthreads();
      {
        const int nk = kt + Cfg::STAGES - 1;
        if (nk < KT) {
          const int slot = nk % Cfg::STAGES;
          load_stage_f64<Cfg>(As + slot * Cfg::A_STAGE, Bs + slot * Cfg::B_STAGE, A + nk * Cfg::BK,
                              B + (int64_t)nk * Cfg::BK * p.ldb, p.lda, p.ldb, m_rem, n_rem,
                              p.k - nk * Cfg::BK, a16, b16);
        }
        cp_async_commit();
      }
      const double* as = As + (kt % Cfg::STAGES) * Cfg::A_STAGE + (wm0 + g) * Cfg::LDA + tig;
      const double* bs = Bs + (kt % Cfg::STAGES) * Cfg::B_STAGE + tig * Cfg::LDB + wn0 + g;
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
    }
    cp_async_wait<0>();

    // epilogue: lane holds C[row g][cols 2*tig, 2*tig+1] of every 8x8 fragment
    double* C = Cbase + p.c_off;
    const bool c16 = ((p.ldc & 1) == 0) && ((p.c_off & 1) == 0) && (((uintptr_t)Cbase & 15) == 0);
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

using CfgA = F64Cfg<64, 64, 32, 32, 4>;     // tile_cfg 0: 128 thr, 75.8 KB smem
using CfgB = F64Cfg<128, 128, 64, 32, 4>;   // tile_cfg 1: 256 thr, 149.5 KB smem
using CfgC = F64Cfg<128, 64, 64, 32, 4>;    // tile_cfg 2: 128 thr, 116.7 KB smem

template <class Cfg, int MIN_CTAS>
static int launch_f64(const double* A, const double* B, double* C, const t10_gemm_problem* prob,
                      int64_t ntiles, const int32_t* tiles, cudaStream_t st) {
  auto kern = k_ggemm_f64<Cfg, MIN_CTAS>;
  static int ctas_per_sm = 0;
  if (ctas_per_sm == 0) {
    T10_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    int occ = 0;
    T10_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::THREADS, Cfg::SMEM_BYTES));
    T10_REQUIRE(occ >= 1, T10_ERR_CUDA, "ggemm kernel does not fit on an SM");
    ctas_per_sm = occ;
  }
  unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)sm_count() * ctas_per_sm);
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(A, B, C, prob, ntiles, (const int4*)tiles);
  T10_LAUNCH_CHECK();
  return T10_OK;
}

static bool tile_dims(int cfg, int* bm, int* bn) {
  switch (cfg) {
    case 0: *bm = 64; *bn = 64; return true;
    case 1: *bm = 128; *bn = 128; return true;
    case 2: *bm = 128; *bn = 64; return true;
    default: return false;
  }
}

}  // namespace t10

using namespace t10;

extern "C" int t10_ggemm_tile_shape(int tile_cfg, int* bm, int* bn) {
  T10_REQUIRE(tile_dims(tile_cfg, bm, bn), T10_ERR_INVALID, "unknown tile_cfg %d", tile_cfg);
  return T10_OK;
}

extern "C" int t10_ggemm_plan_tiles(int tile_cfg, int64_t nprob, const t10_gemm_problem* h_prob,
                                    int64_t cap, int32_t* h_tiles, int64_t* ntiles) {
  int bm, bn;
  T10_REQUIRE(tile_dims(tile_cfg, &bm, &bn), T10_ERR_INVALID, "unknown tile_cfg %d", tile_cfg);
  struct Tl { int32_t p, m0, n0; int64_t cost; };
  std::vector<Tl> v;
  for (int64_t i = 0; i < nprob; ++i) {
    const t10_gemm_problem& p = h_prob[i];
    T10_REQUIRE(p.m >= 0 && p.n >= 0 && p.k >= 0, T10_ERR_INVALID, "negative GEMM extent");
    if (p.m == 0 || p.n == 0) continue;
    for (int m0 = 0; m0 < p.m; m0 += bm)
      for (int n0 = 0; n0 < p.n; n0 += bn) {
        int64_t mm = std::min(bm, p.m - m0), nn = std::min(bn, p.n - n0);
        // cost model: k-loop length dominates; partial tiles still pay the full MMA tile
        v.push_back({(int32_t)i, m0, n0, (int64_t)p.k * 1024 + mm * nn / 64});
      }
  }
  std::stable_sort(v.begin(), v.end(), [](const Tl& a, const Tl& b) { return a.cost > b.cost; });
  *ntiles = (int64_t)v.size();
  if (h_tiles) {
    T10_REQUIRE(cap >= (int64_t)v.size(), T10_ERR_INVALID, "tile buffer too small");
    for (size_t i = 0; i < v.size(); ++i) {
      h_tiles[4 * i + 0] = v[i].p;
      h_tiles[4 * i + 1] = v[i].m0;
      h_tiles[4 * i + 2] = v[i].n0;
      h_tiles[4 * i + 3] = 0;
    }
  }
  return T10_OK;
}

extern "C" int t10_ggemm(int dtype, int tile_cfg, const void* d_A, const void* d_B, void* d_C,
                         int64_t nprob, const t10_gemm_problem* d_prob, int64_t ntiles,
                         const int32_t* d_tiles, t10_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (ntiles == 0 || nprob == 0) return T10_OK;
  T10_REQUIRE(ntiles > 0 && nprob > 0, T10_ERR_INVALID, "negative counts");
  if (dtype == T10_F64) {
    const double* A = (const double*)d_A;
    const double* B = (const double*)d_B;
    double* C = (double*)d_C;
    switch (tile_cfg) {
      case 0: return launch_f64<CfgA, 2>(A, B, C, d_prob, ntiles, d_tiles, st);
      case 1: return launch_f64<CfgB, 1>(A, B, C, d_prob, ntiles, d_tiles, st);
      case 2: return launch_f64<CfgC, 1>(A, B, C, d_prob, ntiles, d_tiles, st);
      case 100: {  // SIMT cross-check path (64x64 tiles), used by tests only
        unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)sm_count() * 4);
        k_ggemm_simt<double><<<grid, 256, 0, st>>>(A, B, C, d_prob, ntiles, (const int4*)d_tiles);
        T10_LAUNCH_CHECK();
        return T10_OK;
      }
      default: T10_REQUIRE(false, T10_ERR_INVALID, "unknown tile_cfg %d", tile_cfg);
    }
  } else if (dtype == T10_F32) {
    T10_REQUIRE(tile_cfg == 0 || tile_cfg == 100, T10_ERR_INVALID, "fp32 ggemm uses 64x64 tiles (cfg 0)");
    unsigned grid = (unsigned)std::min<int64_t>(ntiles, (int64_t)sm_count() * 4);
    k_ggemm_simt<float><<<grid, 256, 0, st>>>((const float*)d_A, (const float*)d_B, (float*)d_C,
                                              d_prob, ntiles, (const int4*)d_tiles);
    T10_LAUNCH_CHECK();
    return T10_OK;
  }
  T10_REQUIRE(false, T10_ERR_UNSUPPORTED, "ggemm: dtype %d not supported", dtype);
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
    if (use_dmma) k_probe_dmma<<<blocks, threads, 0, st>>>(d_out, iters);
    else k_probe_dfma<<<blocks, threads, 0, st>>>(d_out, iters);
    T10_CUDA(cudaEventRecord(e1, st));
    T10_CUDA(cudaEventSynchronize(e1));
  }
  float ms = 0;
  T10_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  // DMMA m8n8k4 = 256 FMA per warp instruction; DFMA = 32 FMA per warp instruction
  double warps = (double)blocks * threads / 32.0;
  double fma = warps * (double)iters * 16.0 * (use_dmma ? 256.0 : 32.0);
  *tflops = 2.0 * fma / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  return T10_OK;
}
