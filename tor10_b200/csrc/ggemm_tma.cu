// Family 3, FP64 grouped GEMM fed by the Tensor Memory Accelerator (tile_cfg 47).
//
// Reference call site: the per-sector `torch.matmul` loop of Contract, UniTensor.py:3727-3741.
//
// Same math as k_ggemm_f64_mixed_sk (ggemm.cu): four DMMA warps of 32x32 per CTA, mixed tile shapes
// (64x64, 32x128, 128x32), stream-K pieces with parked partials added in a fixed order.  What changes is how
// the operands reach shared memory and how the warps synchronise:
//   * a FIFTH warp is the producer: one thread issues `cp.async.bulk.tensor.2d` (SASS UTMALDG) boxes through
//     one tensor map per sector and operand (built by the host, t10_ggemm_tma_encode).  The DMMA warps issue
//     no global loads, no address arithmetic and no cp.async bookkeeping: LDS + DMMA + one mbarrier wait and
//     one arrive per k-step;
//   * the boxes are 32 rows x 20 doubles (A) and 16 k-rows x 36 doubles (B): 4 doubles wider than the warp
//     tile they feed, so the shared-memory row pitch is the padded, bank-conflict-free pitch of the LSU kernel
//     (20 / 36 doubles) without any swizzle -- the over-fetched columns are never read;
//   * sector edges and the K tail need no clamping or predication: elements outside the sector's tensor map
//     arrive as zeros (TMA out-of-bounds fill);
//   * full/empty mbarriers per stage instead of __syncthreads: the four DMMA warps (one per SM sub-partition)
//     drift against each other by up to the ring depth instead of meeting at a CTA barrier every k-step;
//   * the ring is carried ACROSS pieces: while the DMMA warps finish a tile and store / park it, the producer
//     is already loading the first k-steps of the CTA's next piece (the LSU kernel drains and refills its
//     pipeline at every piece: ~3 k-steps for each of ~1000 pieces on the north-star workload).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "t10_common.cuh"

namespace t10 {

constexpr int TG_ABOX_ROWS = 32, TG_ABOX_COLS = 20;      // A box: 32 rows x (16 + 4) k
constexpr int TG_BBOX_ROWS = 16, TG_BBOX_COLS = 36;      // B box: 16 k x (32 + 4) columns
constexpr int TG_ABOX = TG_ABOX_ROWS * TG_ABOX_COLS * 8;   // 5120 B
constexpr int TG_BBOX = TG_BBOX_ROWS * TG_BBOX_COLS * 8;   // 4608 B
constexpr int TG_STAGE_BYTES = 4 * TG_ABOX + TG_BBOX;      // 25088 B: the largest shape (128x32)
constexpr int tg_smem_bytes(int stages) { return stages * TG_STAGE_BYTES + 128; }   // + slack to align the ring
constexpr int TG_THREADS = 160;                          // 4 DMMA warps + 1 producer warp
static_assert(TG_ABOX % 128 == 0 && TG_BBOX % 128 == 0 && TG_STAGE_BYTES % 128 == 0, "TMA destinations: 128 B");
static_assert(2 * TG_ABOX + 2 * TG_BBOX <= TG_STAGE_BYTES && TG_ABOX + 4 * TG_BBOX <= TG_STAGE_BYTES, "");

__device__ __forceinline__ unsigned tg_smem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tg_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void tg_mbar_expect(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tg_mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tg_mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tg_tma_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tg_dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void tg_consumer_sync() { asm volatile("bar.sync 1, 128;\n" ::: "memory"); }

// tile.w = shape | valid rows << 8 | valid columns << 20 (t10_ggemm_plan_tiles, cfg 46)
// shape 0: 64x64 = 2 A boxes x 2 B boxes; 1: 32x128 = 1 x 4; 2: 128x32 = 4 x 1
__device__ __forceinline__ int tg_na(int shape) { return shape == 0 ? 2 : shape == 1 ? 1 : 4; }
__device__ __forceinline__ int tg_nb(int shape) { return shape == 0 ? 2 : shape == 1 ? 4 : 1; }

// <4 stages, 2 CTAs per SM> (default: two DMMA warps per SM sub-partition) or <3 stages, 3 CTAs per SM, 128
// registers> (three warps per sub-partition: more cover for each warp's LDS / epilogue phases, more pieces)
// Completion signal of a sector-resident multi-GPU run: when the last CTA of the launch has stored its tiles it
// raises this rank's progress word on every rank (its own included; peers through their NVLink mappings) to
// `value` with system-scope release semantics.  Consumers on any rank acquire it before they read this rank's
// sectors, and this rank's later launches acquire the peers' words before they overwrite memory the peers read.
struct TgSignal {
  int n;
  unsigned long long* flag[8];
  unsigned long long value;
  int* done_ctr;                   // zero-initialised arrival counter of this launch (reset on exit)
  // Fused GEMM -> all-gather ("push" transport): every finished tile is ALSO stored, straight from the accumulator
  // registers, into the result arenas of the other ranks (npush NVLink peer mappings, same element offsets), so the
  // transfer rides under the remaining pieces' math; the completion signal then covers the remote stores too
  // (system-scope fences), and a rank that has acquired every peer's progress word holds the whole result.
  int npush;
  double* push[7];
};

template <int TG_STAGES, int MIN_CTAS>
__global__ void __launch_bounds__(TG_THREADS, MIN_CTAS)
k_ggemm_f64_tma(const CUtensorMap* __restrict__ maps, double* __restrict__ Cbase,
                const t10_gemm_problem* __restrict__ prob, const int4* __restrict__ tiles,
                const int4* __restrict__ pieces, const int32_t* __restrict__ slot_start,
                double* __restrict__ sk_ws, int* __restrict__ sk_flags, int sk_epoch, int pdl,
                unsigned long long* __restrict__ trace, const __grid_constant__ TgSignal sig) {
  extern __shared__ unsigned char tg_raw[];
  __shared__ __align__(8) unsigned long long bars[2 * TG_STAGES];
  unsigned char* ring = tg_raw + ((128u - (tg_smem(tg_raw) & 127u)) & 127u);   // (keeps the shared address space)
  const unsigned ring_u = tg_smem(ring);
  const unsigned full0 = tg_smem(&bars[0]), empty0 = tg_smem(&bars[TG_STAGES]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

#ifdef T10_TRACE_PHASES
  unsigned long long ph_entry;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ph_entry));
#endif
  const int t_beg = slot_start[blockIdx.x], t_end = slot_start[blockIdx.x + 1];   // (in flight across the set-up)
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < TG_STAGES; ++s) {
      tg_mbar_init(full0 + 8 * s, 1);
      tg_mbar_init(empty0 + 8 * s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (warp == 4) {
    // ------------------------------------------------------------------ producer ----
    if (lane != 0) return;
    // programmatic dependent launch: everything above overlapped the tail of the kernel that writes A
    // (the block relayout); its results are visible after this point
    if (pdl) asm volatile("griddepcontrol.wait;\n" ::: "memory");
    unsigned it = 0;
    for (int t = t_beg; t < t_end; ++t) {
      const int4 tile = tiles[t];
      const int4 pc = pieces[t];
      const int shape = tile.w & 3;
      const int na = tg_na(shape), nb = tg_nb(shape);
      const CUtensorMap* ma = maps + 2 * tile.x;
      const CUtensorMap* mb = ma + 1;
      const unsigned bytes = (unsigned)(na * TG_ABOX + nb * TG_BBOX);
      for (int kt = pc.x; kt < pc.y; ++kt, ++it) {
        const unsigned s = it % TG_STAGES, ph = (it / TG_STAGES) & 1;
        tg_mbar_wait(empty0 + 8 * s, ph ^ 1);            // the DMMA warps have released this slot
        const unsigned bar = full0 + 8 * s;
        tg_mbar_expect(bar, bytes);
        unsigned dst = ring_u + s * TG_STAGE_BYTES;
        for (int a = 0; a < na; ++a, dst += TG_ABOX) tg_tma_2d(dst, ma, kt * 16, tile.y + 32 * a, bar);
        for (int b = 0; b < nb; ++b, dst += TG_BBOX) tg_tma_2d(dst, mb, tile.z + 32 * b, kt * 16, bar);
      }
    }
    return;
  }

  // -------------------------------------------------------------------- DMMA warps ----
  const int g = lane >> 2, tig = lane & 3;
  unsigned it = 0;
  for (int t = t_beg; t < t_end; ++t) {
    const int4 tile = tiles[t];
    const int4 pc = pieces[t];
    const int shape = tile.w & 3, m_lim = (tile.w >> 8) & 0xfff, n_lim = (tile.w >> 20) & 0xfff;
    const int na = tg_na(shape);
    const int ab = shape == 0 ? (warp >> 1) : shape == 1 ? 0 : warp;      // this warp's A box / B box
    const int bb = shape == 0 ? (warp & 1) : shape == 1 ? warp : 0;
    const int wm0 = 32 * ab, wn0 = 32 * bb;
    const int a_off = ab * TG_ABOX + (g * TG_ABOX_COLS + tig) * 8;
    const int b_off = na * TG_ABOX + bb * TG_BBOX + (tig * TG_BBOX_COLS + g) * 8;
    const int m0 = tile.y, n0 = tile.z;
    const int pm = prob[tile.x].m, pn = prob[tile.x].n;
    const int m_rem = m_lim > 0 ? m_lim : pm - m0, n_rem = n_lim > 0 ? n_lim : pn - n0;
    const bool active = (m_rem > wm0) && (n_rem > wn0);
    const int kb = pc.x, ke = pc.y;
    const int pslot = pc.z, nfix = pc.w & 0xff, ffix = pc.w >> 8;
    if (trace != nullptr && threadIdx.x == 0) {
      unsigned smid;
      unsigned long long now;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      trace[4 * t + 0] = smid;
      trace[4 * t + 1] = blockIdx.x;
      trace[4 * t + 2] = now;
    }

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    double af[2][4], bf[2][4];
    auto load_frags = [&](unsigned s, int grp, int buf) {
      const unsigned char* st = ring + s * TG_STAGE_BYTES;
      const double* as = reinterpret_cast<const double*>(st + a_off) + grp * 4;
      const double* bs = reinterpret_cast<const double*>(st + b_off) + grp * 4 * TG_BBOX_COLS;
#pragma unroll
      for (int i = 0; i < 4; ++i) af[buf][i] = as[i * 8 * TG_ABOX_COLS];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[buf][j] = bs[j * 8];
    };
    if (ke > kb) {
      tg_mbar_wait(full0 + 8 * (it % TG_STAGES), (it / TG_STAGES) & 1);
      if (active) load_frags(it % TG_STAGES, 0, 0);
    }
    for (int kt = kb; kt < ke; ++kt, ++it) {
      const unsigned s = it % TG_STAGES;
#pragma unroll
      for (int grp = 0; grp < 4; ++grp) {
        if (grp < 3) {
          if (active) load_frags(s, grp + 1, (grp + 1) & 1);
        } else if (kt + 1 < ke) {
          const unsigned n = it + 1;
          tg_mbar_wait(full0 + 8 * (n % TG_STAGES), (n / TG_STAGES) & 1);
          if (active) load_frags(n % TG_STAGES, 0, 0);
        }
        if (active) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) tg_dmma(acc[i][j][0], acc[i][j][1], af[grp & 1][i], bf[grp & 1][j]);
        }
      }
      // every fragment of this stage has been consumed by a DMMA (so its loads have completed): release it
      __syncwarp();
      if (lane == 0) tg_mbar_arrive(empty0 + 8 * s);
    }
    if (trace != nullptr && threadIdx.x == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      trace[4 * t + 3] = now;
    }

    if (pslot >= 0) {   // park the partial accumulators, publish, done with this piece
      if (active) {
        double2* w = reinterpret_cast<double2*>(sk_ws + (int64_t)pslot * (128 * 34)) + threadIdx.x;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) w[(i * 4 + j) * 128] = make_double2(acc[i][j][0], acc[i][j][1]);
      }
      tg_consumer_sync();
      if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(sk_flags + pslot), "r"(sk_epoch) : "memory");
#ifdef T10_TRACE_PHASES
        if (trace != nullptr) {
          unsigned long long now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          trace[4 * t + 0] = now;
          trace[4 * t + 1] = ph_entry;
        }
#endif
      }
      continue;
    }
    if (nfix > 0) {
      if (warp == 0) {
        // one lane per parked slot: the flags are polled side by side (one L2 round trip instead of nfix)
        for (int f = lane; f < nfix; f += 32) {
          int v = sk_epoch - 1;
          for (int spin = 0; spin < (1 << 22); ++spin) {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(sk_flags + ffix + f) : "memory");
            if (v == sk_epoch) break;
            __nanosleep(32);
          }
          if (v != sk_epoch) sk_flags[gridDim.x] = 1;   // a parked partial never arrived (reported by the host)
        }
      }
      tg_consumer_sync();
      if (active) {
#pragma unroll 2
        for (int f = 0; f < nfix; ++f) {
          const double2* w = reinterpret_cast<const double2*>(sk_ws + (int64_t)(ffix + f) * (128 * 34)) + threadIdx.x;
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const double2 v = __ldcg(w + (i * 4 + j) * 128);
              acc[i][j][0] += v.x;
              acc[i][j][1] += v.y;
            }
        }
      }
    }
#ifdef T10_TRACE_PHASES
    if (trace != nullptr && threadIdx.x == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      trace[4 * t + 0] = now;            // fix-up done (stores follow)
      trace[4 * t + 1] = ph_entry;
    }
#endif
    if (!active) continue;
    const t10_gemm_problem p = prob[tile.x];
    const int m_end = m0 + m_rem, n_end = n0 + n_rem;
    double* C = Cbase + p.c_off;
    const bool c16 = ((p.ldc & 1) == 0) && ((p.c_off & 1) == 0) && (((uintptr_t)Cbase & 15) == 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = m0 + wm0 + i * 8 + g;
      if (r >= m_end) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
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
    if (sig.npush > 0) {
      const int64_t base = p.c_off;
#pragma unroll 1
      for (int q = 0; q < sig.npush; ++q) {
        double* Cq = sig.push[q] + base;
        const bool q16 = ((p.ldc & 1) == 0) && ((p.c_off & 1) == 0) && (((uintptr_t)sig.push[q] & 15) == 0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = m0 + wm0 + i * 8 + g;
          if (r >= m_end) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = n0 + wn0 + j * 8 + 2 * tig;
            double* o = Cq + (int64_t)r * p.ldc + c;
            if (c + 1 < n_end) {
              if (q16) *reinterpret_cast<double2*>(o) = make_double2(acc[i][j][0], acc[i][j][1]);
              else { o[0] = acc[i][j][0]; o[1] = acc[i][j][1]; }
            } else if (c < n_end) {
              o[0] = acc[i][j][0];
            }
          }
        }
      }
    }
  }
  if (sig.n > 0) {
    // remote tile stores: EVERY storing thread orders its own at system scope (a fence by thread 0 alone behind the
    // CTA barrier let a consumer on the peer see the progress word before the tiles: measured, 2-GPU chain test)
    if (sig.npush > 0) __threadfence_system();
    tg_consumer_sync();              // every DMMA warp's tile stores precede thread 0's fence
    if (threadIdx.x == 0) {
      if (sig.npush > 0) __threadfence_system();
      __threadfence();               // release this CTA's tile stores before the arrival ...
      if (atomicAdd(sig.done_ctr, 1) == (int)gridDim.x - 1) {
        *sig.done_ctr = 0;
        __threadfence();             // ... acquire everybody's; the system-scope release below is cumulative over them
        for (int q = 0; q < sig.n; ++q)
          asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(sig.flag[q]), "l"(sig.value) : "memory");
      }
    }
  }
}

__global__ void k_tg_signal(const __grid_constant__ TgSignal sig) {
  if ((int)threadIdx.x < sig.n) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(sig.flag[threadIdx.x]), "l"(sig.value) : "memory");
  }
}

// Acquire: *flag[i] >= need[i] for i < n (flag[i]: a rank's progress word -- this GPU's own, or a peer's through its
// NVLink mapping: the reader polls remotely, so a producer that nobody waits for never touches the link).
// Bounded: after `timeout_ns` (0 = wait for ever, like NCCL) the kernel reports through *err and TRAPS -- the
// context dies loudly instead of letting later kernels read sectors that never arrived.
struct TgWait {
  int n;
  const unsigned long long* flag[8];
  unsigned long long need[8];
  unsigned long long timeout_ns;
  int* err;
};
__global__ void k_tg_wait(const __grid_constant__ TgWait w) {
  const int i = threadIdx.x;
  if (i >= w.n) return;
  unsigned long long t0, now, v = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(w.flag[i]) : "memory");
    if (v >= w.need[i]) break;
    __nanosleep(100);
    if (w.timeout_ns) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > w.timeout_ns) {
        if (w.err) { *w.err = 1 + i; __threadfence_system(); }
        asm volatile("trap;");
      }
    }
  }
}

typedef CUresult (*TgEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TgEncodeFn tg_encode_fn() {
  static TgEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (TgEncodeFn)p;
  }
  return fn;
}

extern unsigned long long* ggemm_trace_ptr();   // ggemm.cu (t10_ggemm_set_trace)

}  // namespace t10

using namespace t10;

static int tg_variant() {   // CTAs per SM: TOR10_TMA_CTAS=2 (default) | 3
  static int v = 0;
  if (v == 0) {
    const char* e = getenv("TOR10_TMA_CTAS");
    v = (e && atoi(e) == 3) ? 3 : 2;
  }
  return v;
}

extern "C" int t10_ggemm_tma_slots(int* nslots) {
  int dev = 0;
  T10_CUDA(cudaGetDevice(&dev));
  static int resident[64] = {0};
  T10_REQUIRE(dev >= 0 && dev < 64, T10_ERR_UNSUPPORTED, "device ordinal %d", dev);
  if (resident[dev] == 0) {
    int occ = 0;
    if (tg_variant() == 3) {
      T10_CUDA(cudaFuncSetAttribute(k_ggemm_f64_tma<3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg_smem_bytes(3)));
      T10_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ggemm_f64_tma<3, 3>, TG_THREADS, tg_smem_bytes(3)));
    } else {
      T10_CUDA(cudaFuncSetAttribute(k_ggemm_f64_tma<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg_smem_bytes(4)));
      T10_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ggemm_f64_tma<4, 2>, TG_THREADS, tg_smem_bytes(4)));
    }
    T10_REQUIRE(occ >= 1, T10_ERR_CUDA, "TMA ggemm kernel does not fit on an SM");
    resident[dev] = occ * sm_count();
  }
  *nslots = resident[dev];
  return T10_OK;
}

extern "C" int t10_ggemm_tma_encode(int64_t nprob, const t10_gemm_problem* h_prob, const void* d_A, const void* d_B,
                                    void* h_maps) {
  T10_REQUIRE(nprob >= 0 && h_prob && h_maps, T10_ERR_INVALID, "ggemm_tma_encode: bad arguments");
  T10_REQUIRE(((uintptr_t)d_A & 15) == 0 && ((uintptr_t)d_B & 15) == 0, T10_ERR_INVALID,
              "ggemm_tma_encode: operand base pointers must be 16-byte aligned");
  TgEncodeFn enc = tg_encode_fn();
  T10_REQUIRE(enc != nullptr, T10_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available");
  CUtensorMap* out = (CUtensorMap*)h_maps;
  const cuuint32_t estr[2] = {1, 1};
  for (int64_t i = 0; i < nprob; ++i) {
    const t10_gemm_problem& p = h_prob[i];
    T10_REQUIRE(p.m >= 0 && p.n >= 0 && p.k >= 0, T10_ERR_INVALID, "negative GEMM extent");
    // empty sectors keep a valid (never used) map: extents of at least one element
    const cuuint64_t m = (cuuint64_t)std::max(p.m, 1), n = (cuuint64_t)std::max(p.n, 1), k = (cuuint64_t)std::max(p.k, 1);
    T10_REQUIRE((p.lda & 1) == 0 && (p.ldb & 1) == 0 && (p.a_off & 1) == 0 && (p.b_off & 1) == 0, T10_ERR_INVALID,
                "ggemm_tma: sector %lld: operand rows must be 16-byte aligned (even lda / ldb / offsets)", (long long)i);
    const cuuint64_t adim[2] = {k, m}, bdim[2] = {n, k};
    const cuuint64_t astr[1] = {(cuuint64_t)std::max(p.lda, 2) * 8}, bstr[1] = {(cuuint64_t)std::max(p.ldb, 2) * 8};
    const cuuint32_t abox[2] = {TG_ABOX_COLS, TG_ABOX_ROWS}, bbox[2] = {TG_BBOX_COLS, TG_BBOX_ROWS};
    CUresult r1 = enc(&out[2 * i], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)((const double*)d_A + p.a_off), adim, astr,
                      abox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&out[2 * i + 1], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)((const double*)d_B + p.b_off), bdim,
                      bstr, bbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    T10_REQUIRE(r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS, T10_ERR_UNSUPPORTED,
                "cuTensorMapEncodeTiled failed for sector %lld (A: %d, B: %d; m %d n %d k %d lda %d ldb %d)", (long long)i,
                (int)r1, (int)r2, p.m, p.n, p.k, p.lda, p.ldb);
  }
  return T10_OK;
}

static int tg_fill_signal(TgSignal* sig, int nsig, void* const* h_sig_flags, uint64_t sig_value, int32_t* d_done_ctr,
                          int npush = 0, void* const* h_push = nullptr) {
  memset(sig, 0, sizeof(*sig));
  T10_REQUIRE(npush >= 0 && npush <= 7 && (npush == 0 || (h_push && nsig > 0)), T10_ERR_INVALID,
              "push: at most 7 peer arenas, and a completion signal");
  sig->npush = npush;
  for (int i = 0; i < npush; ++i) sig->push[i] = (double*)h_push[i];
  T10_REQUIRE(nsig >= 0 && nsig <= 8, T10_ERR_INVALID, "signal: at most 8 ranks");
  T10_REQUIRE(nsig == 0 || (h_sig_flags && d_done_ctr), T10_ERR_INVALID, "signal: flags need an arrival counter");
  sig->n = nsig;
  sig->value = sig_value;
  sig->done_ctr = d_done_ctr;
  for (int i = 0; i < nsig; ++i) sig->flag[i] = (unsigned long long*)h_sig_flags[i];
  return T10_OK;
}

extern "C" int t10_signal_flags(int nsig, void* const* h_sig_flags, uint64_t value, t10_stream_t stream) {
  TgSignal sig;
  int dummy = 0;
  int rc = tg_fill_signal(&sig, nsig, h_sig_flags, value, &dummy);
  if (rc != T10_OK) return rc;
  if (nsig == 0) return T10_OK;
  k_tg_signal<<<1, 32, 0, (cudaStream_t)stream>>>(sig);
  T10_LAUNCH_CHECK();
  return T10_OK;
}

extern "C" int t10_wait_flags(int n, const void* const* h_flags, const uint64_t* h_need, int64_t timeout_ms,
                              int32_t* d_err, t10_stream_t stream) {
  T10_REQUIRE(n >= 0 && n <= 8 && (n == 0 || (h_flags && h_need)), T10_ERR_INVALID, "wait_flags: 0..8 flags");
  if (n == 0) return T10_OK;
  TgWait w;
  memset(&w, 0, sizeof(w));
  w.n = n;
  for (int i = 0; i < n; ++i) {
    w.flag[i] = (const unsigned long long*)h_flags[i];
    w.need[i] = h_need[i];
  }
  w.timeout_ns = timeout_ms > 0 ? (unsigned long long)timeout_ms * 1000000ull : 0ull;
  w.err = d_err;
  k_tg_wait<<<1, 32, 0, (cudaStream_t)stream>>>(w);
  T10_LAUNCH_CHECK();
  return T10_OK;
}

extern "C" int t10_ggemm_tma(const void* d_maps, void* d_C, int64_t nprob, const t10_gemm_problem* d_prob,
                             int64_t npieces, const int32_t* d_tiles, const int32_t* d_pieces, int64_t nslots,
                             const int32_t* d_slot_start, void* d_ws, int32_t* d_flags, int32_t epoch, int pdl,
                             int nsig, void* const* h_sig_flags, uint64_t sig_value, int32_t* d_done_ctr,
                             int npush, void* const* h_push, t10_stream_t stream) {
  TgSignal sig;
  {
    int rc = tg_fill_signal(&sig, nsig, h_sig_flags, sig_value, d_done_ctr, npush, h_push);
    if (rc != T10_OK) return rc;
  }
  if (npieces == 0 || nprob == 0) {      // a rank without work still has to signal
    if (nsig > 0) {
      k_tg_signal<<<1, 32, 0, (cudaStream_t)stream>>>(sig);
      T10_LAUNCH_CHECK();
    }
    return T10_OK;
  }
  T10_REQUIRE(npieces > 0 && nprob > 0 && nslots >= 1, T10_ERR_INVALID, "ggemm_tma: bad counts");
  T10_REQUIRE(d_maps && d_pieces && d_tiles && d_slot_start && d_ws && d_flags, T10_ERR_INVALID, "ggemm_tma: missing table");
  T10_REQUIRE(((uintptr_t)d_maps & 63) == 0, T10_ERR_INVALID, "ggemm_tma: tensor maps must be 64-byte aligned");
  int resident = 0;
  int rc = t10_ggemm_tma_slots(&resident);
  if (rc != T10_OK) return rc;
  // a waiting piece spins on flags raised by other CTAs of the same launch: all of them must be co-resident
  T10_REQUIRE(nslots <= resident, T10_ERR_INVALID, "ggemm_tma: %lld slots but only %d resident CTAs", (long long)nslots,
              resident);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)nslots);
  cfg.blockDim = dim3(TG_THREADS);
  cfg.dynamicSmemBytes = tg_variant() == 3 ? tg_smem_bytes(3) : tg_smem_bytes(4);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  if (tg_variant() == 3)
    T10_CUDA(cudaLaunchKernelEx(&cfg, k_ggemm_f64_tma<3, 3>, (const CUtensorMap*)d_maps, (double*)d_C, d_prob,
                                (const int4*)d_tiles, (const int4*)d_pieces, d_slot_start, (double*)d_ws, (int*)d_flags,
                                (int)epoch, pdl, ggemm_trace_ptr(), sig));
  else
    T10_CUDA(cudaLaunchKernelEx(&cfg, k_ggemm_f64_tma<4, 2>, (const CUtensorMap*)d_maps, (double*)d_C, d_prob,
                                (const int4*)d_tiles, (const int4*)d_pieces, d_slot_start, (double*)d_ws, (int*)d_flags,
                                (int)epoch, pdl, ggemm_trace_ptr(), sig));
  return T10_OK;
}
