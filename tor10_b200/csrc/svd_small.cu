// Small-matrix SVD on the device: one CTA per matrix, one-sided Jacobi (Hestenes) in shared memory.
//
// Reference call site: `torch.svd(a.Storage, some=True)` inside Svd / Svd_truncate, linalg.py:544-678 -- in iTEBD.py
// (BASELINE config 1) ONE 64x64 matrix per step.  cuSOLVER needs 1.8 ms for it on a B200 (gesvdj: 79 sweep kernels of
// 2-30 us each and a device-to-host poll per sweep; gesvd 5.0 ms; gesvda 0.96 ms), 80 % of the step.  A matrix of
// this size fits one SM: A (m x n) and V (n x n) live column-major in shared memory, one WARP per column pair
// rotates both (32 lanes over the rows, three dot products by shuffle reduction), the n/2 disjoint pairs of a
// round-robin round run concurrently, one __syncthreads per round, no global traffic and no host round trip until
// the factors are written out sorted.  Same accuracy class as LAPACK's dgesvj (columns are orthogonalised until no
// pair has |a_p . a_q| > m eps |a_p| |a_q|).  Larger matrices stay on torch / cuSOLVER.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "t10_common.cuh"

namespace t10 {

constexpr int SV_MAX_N = 64;          // columns (after the host has oriented the matrix so that m >= n)
constexpr int SV_MAX_SWEEPS = 40;

struct SvdProb {          // one matrix: element (r, c) of A at d_A[a_off + r * rs + c * cs]
  long long a_off, rs, cs;
  long long u_off, s_off, v_off;      // U [m, n] row-major, S [n], Vh [n, n] row-major (offsets into d_U / d_S / d_Vh)
  long long w_off;                    // warm-start state of this matrix in d_W (n * n + 1 elements), or -1
  int m, n;
};
constexpr int SV_WARM_RESET = 64;     // a warm-started chain restarts from the identity every so many calls
constexpr int SV_WARM_RING = 2;       // slots of a warm-start state, used in turn

constexpr int SV_BATCH = 64;          // problems per launch (descriptors travel as kernel parameters: no allocation,
struct SvdBatch {                     // no copy -- the launch can be captured into a CUDA graph)
  SvdProb p[SV_BATCH];
};

constexpr int SV_LPP = 16;            // lanes per column pair: two pairs share a warp (measured at n = 64: 5490 cycles per
                                      // round with 4 lanes, 3400 with 8, 2690 with 16, 3610 with 32; the rotation angle is a long
                                      // scalar dependency chain; one warp per pair is issue-bound on it at n = 64)

template <typename T>
__global__ void __launch_bounds__(SV_LPP * SV_MAX_N / 2, 1)
k_svd_small(const T* __restrict__ A, T* __restrict__ U, T* __restrict__ S, T* __restrict__ Vh, T* __restrict__ W,
            const __grid_constant__ SvdBatch batch, int* __restrict__ info, double abs_kappa) {
  extern __shared__ __align__(16) unsigned char sv_raw[];
  __shared__ int s_rot;
  __shared__ unsigned s_maxr;          // largest squared |cos| of a rotated pair in this sweep (float bits)
  __shared__ T s_f2;
  __shared__ T s_sig[SV_MAX_N];
  __shared__ T s_n2[SV_MAX_N];         // squared column norms, carried through the rotations of a sweep
  __shared__ int s_rank[SV_MAX_N];
  const SvdProb p = batch.p[blockIdx.x];
  const int m = p.m, n = p.n;
  const int ne = (n + 1) & ~1;                       // even number of columns (a zero column pads odd n)
  const int ms = m | 1, vs = ne | 1;                 // odd column pitches: the four pairs of a warp hit different banks
  T* As = reinterpret_cast<T*>(sv_raw);              // column-major [ne][ms]
  T* Vs = As + (size_t)ne * ms;                      // column-major [ne][vs]
  unsigned char* pairs = reinterpret_cast<unsigned char*>(Vs + (size_t)ne * vs);   // [ne - 1][ne / 2][2] column pairs
  const int tid = threadIdx.x, sub = tid & (SV_LPP - 1), grp = tid / SV_LPP, ngrp = blockDim.x / SV_LPP;
  const unsigned gmask = ((1u << SV_LPP) - 1u) << ((tid & 31) & ~(SV_LPP - 1));    // the lanes of this pair's group
  // round-robin tournament, tabulated once (an integer modulo per round sat on the critical path): column ne-1
  // stays, the others rotate
  for (int i = tid; i < (ne - 1) * (ne / 2); i += blockDim.x) {
    const int round = i / (ne / 2), k = i - round * (ne / 2);
    int pa = k == 0 ? ne - 1 : (round + k) % (ne - 1);
    int pb = (round + (ne - 1) - k) % (ne - 1);
    if (pa > pb) { const int t = pa; pa = pb; pb = t; }
    pairs[2 * i] = (unsigned char)pa;
    pairs[2 * i + 1] = (unsigned char)pb;
  }

  // WARM START (w_off >= 0): the rotations start from the right singular vectors V0 the previous call on this state
  // left behind -- the iteration converges from ANY orthogonal V0, and when consecutive matrices change slowly (an
  // imaginary-time evolution near its fixed point) A V0 is already nearly orthogonal: 2-3 sweeps instead of ~10.
  // The state restarts from the identity every SV_WARM_RESET calls so that rounding in V0 cannot accumulate.
  // The state is a ring of SV_WARM_RING slots used in turn (call counter in the last element): algorithms that
  // alternate between two families of matrices -- iTEBD updates the even and the odd bond in turn -- find in slot
  // (call mod 2) the vectors of the call before last, which is the matrix this one resembles.
  T* Wbase = (W != nullptr && p.w_off >= 0) ? W + p.w_off : nullptr;
  T* Wst = nullptr;
  T calls = T(0);
  if (Wbase != nullptr) {
    calls = Wbase[(size_t)SV_WARM_RING * ((size_t)n * n + 1)];
    Wst = Wbase + (size_t)(((long long)calls) % SV_WARM_RING) * ((size_t)n * n + 1);
  }
  bool warm = false;
  if (Wst != nullptr) {
    const T cnt = Wst[(size_t)n * n];
    warm = cnt >= T(1) && cnt < T(SV_WARM_RESET);
  }
  if (warm) {
    T* Ts = reinterpret_cast<T*>(pairs + (((size_t)(ne - 1) * ne + 15) & ~(size_t)15));      // A, column-major [n][ms]
    for (int i = tid; i < n * m; i += blockDim.x) {
      const int c = i / m, r = i - c * m;
      Ts[(size_t)c * ms + r] = A[p.a_off + (long long)r * p.rs + (long long)c * p.cs];
    }
    for (int i = tid; i < ne * ne; i += blockDim.x) {
      const int c = i / ne, r = i - c * ne;
      Vs[(size_t)c * vs + r] = (c < n && r < n) ? Wst[(size_t)r * n + c] : ((c == r) ? T(1) : T(0));
    }
    __syncthreads();
    for (int i = tid; i < ne * m; i += blockDim.x) {       // As = A V0
      const int c = i / m, r = i - c * m;
      T acc = 0;
      if (c < n)
        for (int k = 0; k < n; ++k) acc += Ts[(size_t)k * ms + r] * Vs[(size_t)c * vs + k];
      As[(size_t)c * ms + r] = acc;
    }
  } else {
    for (int i = tid; i < ne * m; i += blockDim.x) {
      const int c = i / m, r = i - c * m;
      As[(size_t)c * ms + r] = c < n ? A[p.a_off + (long long)r * p.rs + (long long)c * p.cs] : T(0);
    }
    for (int i = tid; i < ne * ne; i += blockDim.x) Vs[(size_t)(i / ne) * vs + i % ne] = (i / ne == i % ne) ? T(1) : T(0);
  }
  __syncthreads();

  const T eps = sizeof(T) == 8 ? T(2.220446049250313e-16) : T(1.1920929e-7f);
  const T tol = eps * (T)m;                 // |a_p . a_q| <= m eps |a_p| |a_q| counts as orthogonal (rounding floor of
  const T tol2 = tol * tol;                 // the dot product itself is ~sqrt(m) eps: a tighter bound never settles)
  // Columns whose norm is below m eps ||A||_F are rounding noise of the larger ones (the numerical null space of a
  // rank-deficient or fast-decaying matrix -- every iTEBD step produces one: 32 of its 64 columns).  Their
  // directions carry no information and are re-randomised by every call's rounding; orthogonalising them costs most
  // of the sweeps and never settles.  A pair that involves such a column is left alone: the column contributes less
  // than m eps ||A||_F to A = U S Vh whatever its direction, V stays orthogonal (a product of rotations), and the
  // triplets of the columns that matter are unaffected to that level.  (U columns of such singular values are
  // therefore not orthonormal -- LAPACK returns an arbitrary orthonormal completion there.)
  if (tid == 0) s_f2 = T(0);
  __syncthreads();
  {
    T part = 0;
    for (int i = tid; i < ne * m; i += blockDim.x) {
      const T v = As[(size_t)(i / m) * ms + i % m];
      part += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) atomicAdd(&s_f2, part);
  }
  __syncthreads();
  const T floor2 = tol2 * s_f2;             // squared norm below which a column counts as noise
  // ... and a pair whose inner product is below eps ||A||_F^2 in ABSOLUTE terms is left alone as well: what it
  // leaves perturbs singular values by less than eps ||A|| -- the normwise accuracy of LAPACK's dgesdd / cuSOLVER.
  // (The purely relative test would keep polishing the tiny triplets of a graded matrix, whose directions a 1e-6
  // change of A scrambles anyway: 8-10 sweeps per iTEBD step even from a warm start.)
  const T abs2 = (eps * (T)abs_kappa * s_f2) * (eps * (T)abs_kappa * s_f2);
  const int npairs = ne / 2, nrounds = ne - 1;
#ifdef T10_SVD_DEBUG
  const long long dbg_c0 = clock64();
  unsigned long long dbg_t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
#endif
  int sweep = 0;
  for (; sweep < SV_MAX_SWEEPS; ++sweep) {
    if (tid == 0) { s_rot = 0; s_maxr = 0u; }
    // squared column norms: computed exactly once per sweep, then carried through the rotations (a rotation turns
    // (alpha, beta, gamma) into c^2 alpha - 2cs gamma + s^2 beta and s^2 alpha + 2cs gamma + c^2 beta), so a round
    // needs ONE dot product per pair instead of three
    for (int c = grp; c < ne; c += ngrp) {
      T a = 0;
      for (int r = sub; r < m; r += SV_LPP) a += As[(size_t)c * ms + r] * As[(size_t)c * ms + r];
#pragma unroll
      for (int o = SV_LPP / 2; o > 0; o >>= 1) a += __shfl_xor_sync(gmask, a, o);
      if (sub == 0) s_n2[c] = a;
    }
    __syncthreads();
    for (int round = 0; round < nrounds; ++round) {
      for (int k = grp; k < npairs; k += ngrp) {
        const int pa = pairs[2 * (round * npairs + k)], pb = pairs[2 * (round * npairs + k) + 1];
        T* ap = As + (size_t)pa * ms;
        T* aq = As + (size_t)pb * ms;
        const T alpha = s_n2[pa], beta = s_n2[pb];
        T gamma = 0;
#pragma unroll 4
        for (int r = sub; r < m; r += SV_LPP) gamma += ap[r] * aq[r];
#pragma unroll
        for (int o = SV_LPP / 2; o > 0; o >>= 1) gamma += __shfl_xor_sync(gmask, gamma, o);
        if (fmin(alpha, beta) >= floor2 && gamma * gamma > fmax(tol2 * alpha * beta, abs2)) {
          // Rotation: t = tan(theta) only has to be CLOSE to the annihilating tangent (what it leaves is removed by
          // the next sweep), so it is computed in fast fp32 arithmetic on operands scaled into range -- an fp64
          // divide or square root is a ~30-instruction dependent chain and the exact formula needs three of each.
          // (c, s) must be a rotation to full precision: c = 1 / sqrt(1 + t^2) is refined from its fp32 seed by two
          // Newton steps.  Far from the diagonal regime (|zeta| large, e.g. a tiny column against a big one) the
          // tangent is gamma / (beta - alpha) and is taken in full precision.
          // (operands scaled by a power of two taken from the exponent of max(alpha, beta): no fp64 divide)
          const T big = fmax(alpha, beta);
          T scale;
          if (sizeof(T) == 8) {
            const int e = (__double2hiint((double)big) >> 20) & 0x7ff;
            scale = (T)__hiloint2double((2046 - max(1, min(e, 2045))) << 20, 0);       // 2^(1023 - e)
          } else {
            scale = T(1) / big;
          }
          const float df = (float)((beta - alpha) * scale), gf = (float)(gamma * scale) * 2.0f;
          T t;
          if (fabsf(gf) < 1e-6f * fabsf(df) || fabsf(gf) < 1e-30f) {
            t = gamma / (beta - alpha);
          } else {
            const float zf = __fdividef(df, gf);
            const float hyp = sqrtf(1.0f + zf * zf);
            t = (T)(__fdividef(1.0f, fabsf(zf) + hyp) * (zf >= 0.0f ? 1.0f : -1.0f));
          }
          const T w = T(1) + t * t;
          T c = (T)rsqrtf((float)w);
          c = c * (T(1.5) - T(0.5) * w * c * c);
          c = c * (T(1.5) - T(0.5) * w * c * c);
          const T s = c * t;
#pragma unroll 4
          for (int r = sub; r < m; r += SV_LPP) {
            const T x = ap[r], y = aq[r];
            ap[r] = c * x - s * y;
            aq[r] = s * x + c * y;
          }
          T* vp = Vs + (size_t)pa * vs;
          T* vq = Vs + (size_t)pb * vs;
#pragma unroll 4
          for (int r = sub; r < ne; r += SV_LPP) {
            const T x = vp[r], y = vq[r];
            vp[r] = c * x - s * y;
            vq[r] = s * x + c * y;
          }
          if (sub == 0) {
            const T cc = c * c, ss = s * s, cs2 = T(2) * c * s * gamma;
            s_n2[pa] = fmax(cc * alpha - cs2 + ss * beta, T(0));
            s_n2[pb] = fmax(ss * alpha + cs2 + cc * beta, T(0));
            s_rot = 1;
            const float abf = (float)(alpha * scale) * (float)(beta * scale);
            const float r2 = abf > 0.0f ? fminf(__fdividef(0.25f * gf * gf, abf), 1.0f) : 1.0f;
            atomicMax(&s_maxr, __float_as_uint(r2));
          }
        }
      }
      __syncthreads();
    }
    // converged: nothing rotated -- or everything that did was already orthogonal to sqrt-of-tolerance level: what a
    // rotation with an fp32 tangent leaves is ~1e-7 of what it found (plus second-order terms), so after a sweep whose
    // largest |cos| was below 1e-8 (1e-4 in fp32) every pair is below the tolerance and the checking sweep is skipped
    const float done_below = sizeof(T) == 8 ? 1e-16f : 1e-8f;
    if (s_rot == 0 || __uint_as_float(s_maxr) < done_below) break;
    __syncthreads();
  }
#ifdef T10_SVD_DEBUG
  if (tid == 0 && info != nullptr) {
    unsigned long long dbg_t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t1));
    info[1] = (int)(clock64() - dbg_c0);
    info[2] = (int)(dbg_t1 - dbg_t0);
  }
#endif
  // info: bit 0 = still rotating after SV_MAX_SWEEPS sweeps, bits 1.. = sweeps run
  if (tid == 0 && info != nullptr) info[blockIdx.x] = (sweep >= SV_MAX_SWEEPS ? 1 : 0) | ((sweep + 1) << 1);

  // singular values = column norms; sorted descending (stable in the column index)
  for (int c = grp; c < n; c += ngrp) {
    T a = 0;
    for (int r = sub; r < m; r += SV_LPP) a += As[(size_t)c * ms + r] * As[(size_t)c * ms + r];
#pragma unroll
    for (int o = SV_LPP / 2; o > 0; o >>= 1) a += __shfl_xor_sync(gmask, a, o);
    if (sub == 0) s_sig[c] = sqrt(a);
  }
  __syncthreads();
  for (int c = tid; c < n; c += blockDim.x) {
    const T sc = s_sig[c];
    int rk = 0;
    for (int j = 0; j < n; ++j) rk += (s_sig[j] > sc || (s_sig[j] == sc && j < c)) ? 1 : 0;
    s_rank[c] = rk;
    S[p.s_off + rk] = sc;
  }
  __syncthreads();
  for (int i = tid; i < n * m; i += blockDim.x) {
    const int c = i / m, r = i - c * m;
    const T sc = s_sig[c];
    // a zero singular value has no direction: its U column is written as zeros (cuSOLVER completes the basis)
    U[p.u_off + (long long)r * n + s_rank[c]] = sc > T(0) ? As[(size_t)c * ms + r] / sc : T(0);
  }
  for (int i = tid; i < n * n; i += blockDim.x) {
    const int c = i / n, r = i - c * n;          // Vh[rank(c), r] = V[r, c]
    Vh[p.v_off + (long long)s_rank[c] * n + r] = Vs[(size_t)c * vs + r];
    if (Wst != nullptr) Wst[(size_t)r * n + c] = Vs[(size_t)c * vs + r];
  }
  if (Wst != nullptr && tid == 0) {
    Wst[(size_t)n * n] = warm ? Wst[(size_t)n * n] + T(1) : T(1);
    Wbase[(size_t)SV_WARM_RING * ((size_t)n * n + 1)] = calls + T(1) < T(1 << 20) ? calls + T(1) : T(0);
  }
}

}  // namespace t10

using namespace t10;

extern "C" int t10_svd_small_limits(int* max_n, int* max_elems) {
  *max_n = SV_MAX_N;
  *max_elems = (200 * 1024) / 8;         // m * ne + ne * ne doubles of shared memory
  return T10_OK;
}

extern "C" int t10_svd_small(int dtype, int64_t nprob, const int64_t* h_prob, const void* d_A, void* d_U, void* d_S,
                             void* d_Vh, void* d_W, int32_t* d_info, t10_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  T10_REQUIRE(dtype == T10_F64 || dtype == T10_F32, T10_ERR_UNSUPPORTED, "svd_small: dtype %d", dtype);
  T10_REQUIRE(nprob >= 0 && (nprob == 0 || h_prob), T10_ERR_INVALID, "svd_small: bad arguments");
  if (nprob == 0) return T10_OK;
  const size_t esz = dtype == T10_F64 ? 8 : 4;
  static bool attr_set[64] = {false};
  int dev = 0;
  T10_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    T10_CUDA(cudaFuncSetAttribute(k_svd_small<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    T10_CUDA(cudaFuncSetAttribute(k_svd_small<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set[dev] = true;
  }
  // absolute floor of the rotation test, in units of eps ||A||_F^2 (TOR10_SVD_ABS; 0 = purely relative test)
  static double kappa = -1.0;
  if (kappa < 0.0) {
    const char* e = getenv("TOR10_SVD_ABS");
    kappa = e ? atof(e) : 1e-4;
  }
  for (int64_t base = 0; base < nprob; base += SV_BATCH) {
    const int cnt = (int)std::min<int64_t>(SV_BATCH, nprob - base);
    SvdBatch b;
    memset(&b, 0, sizeof(b));
    size_t smem = 0;
    int maxn = 0;
    for (int i = 0; i < cnt; ++i) {
      const int64_t* q = h_prob + 9 * (base + i);       // {m, n, a_off, rs, cs, u_off, s_off, v_off, w_off}
      SvdProb& p = b.p[i];
      p.m = (int)q[0];
      p.n = (int)q[1];
      p.a_off = q[2];
      p.rs = q[3];
      p.cs = q[4];
      p.u_off = q[5];
      p.s_off = q[6];
      p.v_off = q[7];
      p.w_off = d_W ? q[8] : -1;
      T10_REQUIRE(p.m >= p.n && p.n >= 1 && p.n <= SV_MAX_N, T10_ERR_UNSUPPORTED,
                  "svd_small: matrix %lld is %d x %d (needs m >= n, 1 <= n <= %d)", (long long)(base + i), p.m, p.n,
                  SV_MAX_N);
      const size_t ne = (size_t)((p.n + 1) & ~1);
      smem = std::max(smem, (ne * (size_t)(p.m | 1) + ne * (ne | 1)) * esz + (((ne - 1) * ne + 15) & ~(size_t)15) +
                                (p.w_off >= 0 ? (size_t)p.n * (p.m | 1) * esz : 0));
      maxn = std::max(maxn, (int)ne);
    }
    T10_REQUIRE(smem <= 200 * 1024, T10_ERR_UNSUPPORTED, "svd_small: %zu bytes of shared memory needed", smem);
    const int threads = std::max(32, (SV_LPP * (maxn / 2) + 31) / 32 * 32);
    int* info = d_info ? d_info + base : nullptr;
    if (dtype == T10_F64)
      k_svd_small<double><<<(unsigned)cnt, threads, smem, st>>>((const double*)d_A, (double*)d_U, (double*)d_S,
                                                               (double*)d_Vh, (double*)d_W, b, info, kappa);
    else
      k_svd_small<float><<<(unsigned)cnt, threads, smem, st>>>((const float*)d_A, (float*)d_U, (float*)d_S,
                                                              (float*)d_Vh, (float*)d_W, b, info, kappa);
    T10_LAUNCH_CHECK();
  }
  return T10_OK;
}
